#!/bin/bash
# the driver's round-end sequence: both arms at N = 1, 2, 4, 8 on one box (run under gpurun --gpus 8)
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n"; fi
  $L bench.py --impl reference --gpus $n --steps 5 --warmup 1 > gpurun_out/scale_ref_n$n.json 2> gpurun_out/scale_ref_n$n.err
  $L bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  echo "N=$n rc=$?"
done
nproc; free -g | head -2
