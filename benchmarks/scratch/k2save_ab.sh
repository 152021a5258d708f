#!/bin/bash
# K2 time with the SAVE instantiation under build variants (timing only): bash benchmarks/scratch/k2save_ab.sh "<flags>" ...
i=0
for fl in "$@"; do
  i=$((i+1))
  PENEO_NVCC_EXTRA="$fl" python -m peneo_b200.build --force > /dev/null 2>&1 || { echo "build failed: $fl"; continue; }
  echo "== variant $i: $fl"
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pair_heads_tc -c 3 --csv --log-file gpurun_out/k2save_$i.csv python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 1 --warmup 1 > /dev/null 2>&1
  python benchmarks/scratch/launch_sum.py gpurun_out/k2save_$i.csv 2>/dev/null | sed -n 2,3p
done
