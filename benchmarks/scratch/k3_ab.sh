#!/bin/bash
# A/B of K3 build variants (kernel time alone under ncu): bash benchmarks/scratch/k3_ab.sh "<flags A>" ...
i=0
for fl in "$@"; do
  i=$((i+1))
  PENEO_NVCC_EXTRA="$fl" python -m peneo_b200.build --force > /dev/null 2>&1 || { echo "build failed: $fl"; continue; }
  echo "== variant: $fl"
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "decode or spot or pipeline" 2>&1 | tail -1
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:decode_spots -c 12 --csv --log-file gpurun_out/k3ab_$i.csv python bench.py --steps 3 --warmup 3 --no-sweep --no-train > /dev/null 2>&1
  python benchmarks/scratch/launch_sum.py gpurun_out/k3ab_$i.csv 2>/dev/null | sed -n 2,3p
done
