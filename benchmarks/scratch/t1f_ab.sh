#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "training or fused_loss or dropout or saved or fine_tuning" 2>&1 | tail -3
for f in 1 0; do
  echo "== PENEO_T1F=$f"
  for i in 1 2; do PENEO_T1F=$f timeout 200 python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 20 2>/dev/null | tail -1 | cut -c190-250; done
  PENEO_T1F=$f timeout 200 python benchmarks/train_step.py --seq-len 1024 --batch 4 --hin 960 --steps 20 2>/dev/null | tail -1 | cut -c190-250
done
PENEO_T1F=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/t1f_launches.csv python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 1 --warmup 1 > /dev/null 2>&1
python benchmarks/scratch/launch_sum.py gpurun_out/t1f_launches.csv 2>/dev/null | head -7
