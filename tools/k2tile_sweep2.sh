KB2_TILE_MULT=1 KB2_PROBE_TAG=x timeout 600 python tools/k2win_probe.py quick > gpurun_out/probe_m1.log 2>&1; grep "FAIL" gpurun_out/probe_m1.log | head -20
