#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "training or fused_loss or dropout" 2>&1 | tail -3
for g in 0 24; do
  echo "== PENEO_SAVE_ACT_GB=$g"
  PENEO_SAVE_ACT_GB=$g python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 20 2>/dev/null | tail -1 | cut -c190-250,370-420
  PENEO_SAVE_ACT_GB=$g python benchmarks/train_step.py --seq-len 1024 --batch 4 --hin 960 --steps 20 2>/dev/null | tail -1 | cut -c190-250
done
PENEO_SAVE_ACT_GB=24 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_w9_train_launches.csv python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 1 --warmup 1 > /dev/null 2>&1
python benchmarks/scratch/launch_sum.py gpurun_out/r02_w9_train_launches.csv 2>/dev/null | head -8
