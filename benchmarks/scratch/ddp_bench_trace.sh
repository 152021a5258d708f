#!/bin/bash
# 2-GPU bench with the per-step trace of the train leg: gpurun --gpus 2 -- bash benchmarks/scratch/ddp_bench_trace.sh
for i in 1 2; do
  PENEO_BENCH_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$i \
      bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/ddp_trace_$i.json 2> gpurun_out/ddp_trace_$i.err
done
