for cfg in "RA=64" "RA=80" "RA=100" "RA=128"; do
  envs=""; for kv in $cfg; do envs="$envs KB2_TILE_$kv"; done
  echo "== $cfg"; env $envs KB2_PROBE_TAG=x timeout 600 python tools/k2win_probe.py > gpurun_out/probe_tmp.log 2>&1; grep FAIL gpurun_out/probe_tmp.log | tail -4; grep "^tile_lin" gpurun_out/probe_tmp.log | cut -c1-110
done
