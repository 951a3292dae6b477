#!/bin/bash
# one --set full capture of the codon scan on the C3 shape (2 populations), details + per-line source counters as CSV
python scripts/ncu_targets.py k4 > gpurun_out/plain_k4.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain_k4.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:pfa_cds_scan -s 2 -c 1 -f -o /tmp/prof_k4 python scripts/ncu_targets.py k4 > gpurun_out/ncu_k4.log 2>&1
ncu -i /tmp/prof_k4.ncu-rep --page details --csv > gpurun_out/${1:-r2}_cds_ncu_details.csv 2>/dev/null
ncu -i /tmp/prof_k4.ncu-rep --page source --csv > gpurun_out/${1:-r2}_cds_ncu_source.csv 2>/dev/null
ls -la gpurun_out/${1:-r2}_cds_ncu_*
