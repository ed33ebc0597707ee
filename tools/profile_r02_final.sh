#!/bin/bash
# final evidence of round 2 (run under gpurun; every command first runs without ncu)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; tail -n 3 gpurun_out/final_pytest.log
python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -n 1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err || exit 1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-named-shapes > gpurun_out/prof_bench_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-named-shapes > gpurun_out/prof_bench_ncu.log 2>&1
python tools/k2one.py 4 448 20000 3 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:self_moments_tile -c 1 -o gpurun_out/r02_final_k2_tile_cfg1 -f python tools/k2one.py 4 448 20000 2 > gpurun_out/prof_k2a.log 2>&1
ls -la gpurun_out | tail -6
