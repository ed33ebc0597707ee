#!/bin/bash
# bench.py at N ranks with several numbers of analyses in flight (and spare SMs of the tiled moment kernel):
#   tools/inflight_sweep.sh N "4 6 8" -> gpurun_out/inflight_nN.txt
N=${1:-2}
OUT=gpurun_out/inflight_n$N.txt
: > $OUT
SPECS=${2:-4:2 6:2 8:2 8:4}
for spec in $SPECS; do
  f=${spec%%:*}; s=${spec##*:}
  echo "== in-flight $f spare SMs $s" >> $OUT
  KB2_TILE_SPARE_SMS=$s python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 80 --warmup 3 --no-cpu-baseline --no-named-shapes --in-flight $f 2>> gpurun_out/inflight_n$N.err | \
    python -c "import sys, json; d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(json.dumps({k: d[k] for k in ('value', 'ms_per_step', 'stage_ms', 'latency')}), json.dumps(d['e2e']))" >> $OUT
done
cat $OUT
