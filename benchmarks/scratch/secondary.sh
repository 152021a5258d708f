#!/bin/bash
timeout 300 python benchmarks/secondary_sweep.py > gpurun_out/r02_v4_secondary_sweep.jsonl 2>/dev/null; cut -c1-220 gpurun_out/r02_v4_secondary_sweep.jsonl
for args in "--seq-len 512 --batch 4 --no-shrink" "--seq-len 512 --batch 4 --layers 3" "--seq-len 512 --batch 4 --layers 1" "--seq-len 512 --batch 4 --precision fp32"; do
  timeout 300 python benchmarks/train_step.py $args --steps 10 --warmup 3 2>/dev/null | tail -1
done > gpurun_out/r02_v4_train_step_other_configs.jsonl
cut -c90-330 gpurun_out/r02_v4_train_step_other_configs.jsonl
timeout 300 python benchmarks/mixed_sweep.py --docs 2000 > gpurun_out/r02_v4_mixed_sweep_1gpu.json 2>/dev/null; cut -c1-400 gpurun_out/r02_v4_mixed_sweep_1gpu.json
