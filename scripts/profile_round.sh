#!/bin/bash
# ncu evidence of one round (run under gpurun, ONE GPU): launch list of the bench command, then --set full captures of the
# dominant kernels.  Outputs under gpurun_out/; summaries are extracted on the CPU box into profiles/.
set -x
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-check --e2e-steps 1 --e2e-sites 200000"
$B > gpurun_out/plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_launches.log 2>&1
B2="python bench.py --steps 2 --warmup 3 --no-cpu --no-check --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:pfa_site_scan_reg -s 3 -c 1 -f -o gpurun_out/prof_site $B2 > gpurun_out/ncu_site.log 2>&1
python scripts/ncu_targets.py k4 > gpurun_out/plain_k4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pfa_cds_scan -s 2 -c 1 -f -o gpurun_out/prof_cds python scripts/ncu_targets.py k4 > gpurun_out/ncu_cds.log 2>&1
python scripts/ncu_targets.py k1 > gpurun_out/plain_k1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pfa_encode -s 40 -c 4 -f -o gpurun_out/prof_enc python scripts/ncu_targets.py k1 > gpurun_out/ncu_enc.log 2>&1
ls -la gpurun_out/*.ncu-rep
