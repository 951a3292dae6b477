#!/bin/bash
# ncu evidence of one round (run under gpurun, ONE GPU): launch list of the bench command, then --set full captures of the
# dominant kernels, exported to CSV on the box (the .ncu-rep files are too large to travel back together).
# usage: scripts/profile_round.sh <tag>   (files gpurun_out/<tag>_*)
tag=${1:-r2}
set -x
cap() {  # cap <name> <kernel regex> <skip> <count> <command...>
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  "$@" > gpurun_out/plain_$name.log 2>&1 || return 1
  ncu --set full --import-source on --clock-control none -k regex:$rx -s $skip -c $cnt -f -o /tmp/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page details --csv > gpurun_out/${tag}_${name}_ncu_details.csv 2>/dev/null
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv | python scripts/ncu_pick.py > gpurun_out/${tag}_${name}_ncu_raw_picked.csv 2>/dev/null
}
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-check --e2e-steps 1 --e2e-sites 200000"
$B > gpurun_out/plain_bench.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench_c4.csv $B > gpurun_out/ncu_launches.log 2>&1
cap site pfa_site_scan_tma 3 1 python bench.py --steps 2 --warmup 3 --no-cpu --no-check --no-e2e
cap cds pfa_cds_scan 2 1 python scripts/ncu_targets.py k4
cap pw pfa_pairwise_kernel 1 1 python scripts/ncu_targets.py k3
cap batch pfa_batch_site 2 1 python scripts/ncu_targets.py k2b
ls -la gpurun_out/ | grep ${tag}_
