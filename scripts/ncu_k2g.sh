#!/bin/bash
# one --set full capture: the validity-aware site scan on two gapped shards (second launch of each)
python scripts/ncu_targets.py k2g > gpurun_out/plain_k2g.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain_k2g.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:pfa_site_scan_tma -c 4 -f -o /tmp/prof_k2g python scripts/ncu_targets.py k2g > gpurun_out/ncu_k2g.log 2>&1
ncu -i /tmp/prof_k2g.ncu-rep --page raw --csv > gpurun_out/${1:-r2}_k2g_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_k2g.ncu-rep --page source --csv > gpurun_out/${1:-r2}_k2g_ncu_source.csv 2>/dev/null
ls -la gpurun_out/${1:-r2}_k2g_ncu_*
