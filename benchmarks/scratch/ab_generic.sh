#!/bin/bash
# bench A/B of build variants: bash benchmarks/scratch/ab_generic.sh "<flags A>" "<flags B>" ...   (100-step bench + train step)
i=0
for fl in "$@"; do
  i=$((i+1))
  PENEO_NVCC_EXTRA="$fl" python -m peneo_b200.build --force > /dev/null 2>&1 || { echo "build failed: $fl"; continue; }
  echo "== variant $i: $fl"
  python bench.py --no-sweep --no-train > gpurun_out/ab_bench_$i.json 2>/dev/null
  python benchmarks/scratch/brief.py gpurun_out/ab_bench_$i.json
  python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 20 2>/dev/null | tail -1 | cut -c190-250
done
