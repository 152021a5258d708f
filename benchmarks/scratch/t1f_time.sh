#!/bin/bash
# kernel time of T1f under build variants (timing only): bash benchmarks/scratch/t1f_time.sh "<flags>" ...
i=0
for fl in "$@"; do
  i=$((i+1))
  PENEO_NVCC_EXTRA="$fl" python -m peneo_b200.build --force > /dev/null 2>&1 || { echo "build failed: $fl"; continue; }
  echo "== variant $i: $fl"
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_ds_fused -c 8 --csv --log-file gpurun_out/t1ft_$i.csv python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 1 --warmup 0 > /dev/null 2>&1
  python benchmarks/scratch/launch_sum.py gpurun_out/t1ft_$i.csv 2>/dev/null | sed -n 2,3p
done
