#!/bin/bash
# timing A/B of K2 build variants on the box: bash benchmarks/scratch/k2_ab.sh "<flags A>" "<flags B>" ...
for fl in "$@"; do
  PENEO_NVCC_EXTRA="$fl" python -m peneo_b200.build --force > /dev/null 2>&1 || { echo "build failed: $fl"; continue; }
  echo "== variant: $fl"
  python bench.py --steps 20 --warmup 5 --no-sweep --no-train 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value'],1), 'k2_ms', round(d['roofline']['kernel_ms'],4), 'frac_burst', round(d['roofline']['frac_burst'],4), d['clocks']['sm_mhz'])"
done
