#!/bin/bash
# A/B of T1 build variants on the box: bash benchmarks/scratch/t1_ab.sh "<flags A>" "<flags B>" ...
i=0
for fl in "$@"; do
  i=$((i+1))
  PENEO_NVCC_EXTRA="$fl" python -m peneo_b200.build --force > /dev/null 2>&1 || { echo "build failed: $fl"; continue; }
  echo "== variant $i: $fl"
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "training_step_default or fused_loss" 2>&1 | tail -1
  PENEO_NVCC_EXTRA="$fl" python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 20 2>/dev/null | tail -1 | cut -c190-250
  PENEO_NVCC_EXTRA="$fl" ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/ab_$i.csv python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 1 --warmup 1 > /dev/null 2>&1
  python benchmarks/scratch/launch_sum.py gpurun_out/ab_$i.csv | sed -n 2,6p
done
