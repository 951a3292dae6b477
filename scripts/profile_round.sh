#!/bin/bash
# ncu evidence of one round (run under gpurun, ONE GPU): launch list of the bench command, then --set full captures of the
# dominant kernels, exported to CSV on the box (the .ncu-rep files are too large to travel back together).
set -x
cap() {  # cap <name> <kernel regex> <skip> <count> <command...>
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  "$@" > gpurun_out/plain_$name.log 2>&1 || return 1
  ncu --set full --clock-control none -k regex:$rx -s $skip -c $cnt -f -o /tmp/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page details --csv > gpurun_out/${name}_ncu_details.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/${name}_ncu_raw.csv 2>/dev/null
}
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-check --e2e-steps 1 --e2e-sites 200000"
$B > gpurun_out/plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv $B > gpurun_out/ncu_launches.log 2>&1
cap site pfa_site_scan_tma 3 1 python bench.py --steps 2 --warmup 3 --no-cpu --no-check --no-e2e
cap cds pfa_cds_scan 2 1 python scripts/ncu_targets.py k4
cap enc pfa_encode 40 4 python scripts/ncu_targets.py k1
cap pw pfa_pairwise_kernel 1 1 python scripts/ncu_targets.py k3
ls -la gpurun_out/ | head -30
