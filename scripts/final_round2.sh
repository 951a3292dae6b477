#!/bin/bash
# round-2 closing run on ONE GPU: all GPU tests, smoke(), the bench lines of the three workloads and of the gap cases, then the
# ncu launch list of the default bench command.  usage: scripts/final_round2.sh <tag>
tag=${1:-r2final}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gpu_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/${tag}_gpu_tests.log; tail -3 gpurun_out/${tag}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/${tag}_bench_c4.json 2> gpurun_out/${tag}_bench_c4.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_c4_reference.json 2>/dev/null
python bench.py --gaps-ppm 1 --no-cpu > gpurun_out/${tag}_bench_c4_gaps1.json 2>/dev/null
python bench.py --gaps-ppm 10 --no-cpu > gpurun_out/${tag}_bench_c4_gaps10.json 2>/dev/null
python bench.py --gaps-ppm 100 --no-cpu > gpurun_out/${tag}_bench_c4_gaps100.json 2>/dev/null
python bench.py --force-validity --no-cpu > gpurun_out/${tag}_bench_c4_forcev.json 2>/dev/null
python bench.py --workload c3 > gpurun_out/${tag}_bench_c3.json 2>/dev/null
python bench.py --workload c5 > gpurun_out/${tag}_bench_c5.json 2>/dev/null
python scripts/probe_gaps5.py > gpurun_out/${tag}_probe_gaps.log 2>&1
for f in c4 c4_reference c4_gaps1 c4_gaps10 c4_gaps100 c4_forcev c3 c5; do python scripts/bench_line.py gpurun_out/${tag}_bench_$f.json; done
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-check --e2e-steps 1 --e2e-sites 200000"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench_c4.csv $B > gpurun_out/ncu_launches.log 2>&1
tail -4 gpurun_out/${tag}_launches_bench_c4.csv | cut -c1-200
