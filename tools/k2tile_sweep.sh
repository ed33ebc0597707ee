for cfg in "def"; do
  envs=""; for kv in $cfg; do [ "$kv" != def ] && envs="$envs KB2_TILE_$kv"; done
  echo "== $cfg"; env $envs KB2_PROBE_TAG=x timeout 600 python tools/k2win_probe.py > gpurun_out/probe_tmp.log 2>&1; grep FAIL gpurun_out/probe_tmp.log | tail -4; grep "^tile_" gpurun_out/probe_tmp.log | cut -c1-110
done
timeout 900 python -m pytest tests -m gpu -x -q -k "tiled or production or baseline_length or raw_sums or chunked or many_atoms" 2>&1 | tail -3
