#!/bin/bash
# C3-shape K2 / K4 timings under the slot variants of the TMA kernels (run under gpurun)
for v in "1 0" "2 0" "2 1" "2 2" "1 2" "1 4" "3 1"; do
  set -- $v
  if [ "$2" = "0" ]; then
    echo "== stages=$1 m=default"; PFA_SITE_TMA=$1 PFA_CDS_TMA=$1 python scripts/probe_c3_c5.py --c3-only
  else
    echo "== stages=$1 m=$2"; PFA_SITE_TMA=$1 PFA_CDS_TMA=$1 PFA_SITE_TMA_M=$2 PFA_CDS_TMA_M=$2 python scripts/probe_c3_c5.py --c3-only
  fi
done
