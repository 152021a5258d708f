#!/bin/bash
for i in 1 2; do for f in 0 1; do
  PENEO_FUSED_SPOTS=$f python bench.py --no-sweep --no-train > gpurun_out/fused_${f}_$i.json 2>/dev/null
done; done
python benchmarks/scratch/brief.py gpurun_out/fused_*.json
