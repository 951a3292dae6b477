#!/bin/bash
# one --set full capture: the pure site-scan kernel and the validity-aware one on the same clean shard
python scripts/ncu_targets.py k2v > gpurun_out/plain_k2v.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain_k2v.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:pfa_site_scan_tma -s 1 -c 2 -f -o /tmp/prof_k2v python scripts/ncu_targets.py k2v > gpurun_out/ncu_k2v.log 2>&1
ncu -i /tmp/prof_k2v.ncu-rep --page details --csv > gpurun_out/${1:-r2}_k2v_ncu_details.csv 2>/dev/null
ncu -i /tmp/prof_k2v.ncu-rep --page raw --csv > gpurun_out/${1:-r2}_k2v_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_k2v.ncu-rep --page source --csv > gpurun_out/${1:-r2}_k2v_ncu_source.csv 2>/dev/null
ls -la gpurun_out/${1:-r2}_k2v_ncu_*
