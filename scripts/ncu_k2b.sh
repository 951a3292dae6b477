#!/bin/bash
# one --set full capture of the batched site scan on the C5 shape (2,500 resident loci of 100 x 5 kb)
python scripts/ncu_targets.py k2b > gpurun_out/plain_k2b.log 2>&1 || { echo "plain run failed"; tail gpurun_out/plain_k2b.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:pfa_batch_site -s 2 -c 1 -f -o /tmp/prof_k2b python scripts/ncu_targets.py k2b > gpurun_out/ncu_k2b.log 2>&1
ncu -i /tmp/prof_k2b.ncu-rep --page raw --csv > gpurun_out/${1:-r2}_k2b_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_k2b.ncu-rep --page source --csv > gpurun_out/${1:-r2}_k2b_ncu_source.csv 2>/dev/null
ls -la gpurun_out/${1:-r2}_k2b_ncu_*
