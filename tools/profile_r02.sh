#!/bin/bash
# the ncu evidence of round 2 (run under gpurun; every command first runs without ncu)
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-named-shapes > gpurun_out/prof_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-named-shapes > gpurun_out/prof_bench_ncu.log 2>&1
python tools/k2one.py 4 448 20000 3 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:self_moments_tile -c 1 -o gpurun_out/r02_k2_tile_cfg1 -f python tools/k2one.py 4 448 20000 2 > gpurun_out/prof_k2a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:self_moments_tile -c 1 -o gpurun_out/r02_k2_tile_50k -f python tools/k2one.py 4 2000 50000 2 > gpurun_out/prof_k2b.log 2>&1
python tools/k1probe.py > gpurun_out/k1probe.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:unwrap_cumsum_bulk -c 1 -o gpurun_out/r02_k1_unwrap_bulk -f python tools/k1probe.py 448 1088 20000 2 > gpurun_out/prof_k1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:frame_sums_bulk -c 1 -o gpurun_out/r02_k1_frame_sums_bulk -f python tools/k1probe.py 448 1088 20000 2 > gpurun_out/prof_k1f.log 2>&1
ls -la gpurun_out | tail -12
