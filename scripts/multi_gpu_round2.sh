#!/bin/bash
# multi-GPU evidence of round 2 (run under gpurun --gpus N): the 2-process tests on real peers, then the bench lines
N=${1:-2}
tag=${2:-r2m}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_xchg.py -m gpu -q -rs > gpurun_out/${tag}_xchg_tests_n$N.log 2>&1; echo "xchg tests rc=$?" >> gpurun_out/${tag}_xchg_tests_n$N.log; tail -6 gpurun_out/${tag}_xchg_tests_n$N.log
for wl in c4 c3 c5; do
  extra=""; [ $wl = c5 ] && extra="--steps 5"
  timeout 900 $TR bench.py --gpus $N --workload $wl --warmup 3 $extra > gpurun_out/${tag}_bench_${wl}_n$N.json 2> gpurun_out/${tag}_bench_${wl}_n$N.err; echo "$wl rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_${wl}_n$N.json"))
    e = d.get("e2e") or {}
    print("$wl N=$N: value %.3e ms/step %.4f kernel_ms %s frac %.3f e2e %.3e" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], e.get("value", 0)))
except Exception as ex:
    print("$wl: no line", ex)
PY
  tail -3 gpurun_out/${tag}_bench_${wl}_n$N.err
done
