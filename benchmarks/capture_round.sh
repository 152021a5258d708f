#!/bin/bash
# Reproduces the files under profiles/ for one round: run on a B200 box from the repo root, e.g.
#   gpurun --timeout 1200 -- 'bash benchmarks/capture_round.sh r01_v5'
# Writes into gpurun_out/<tag>_*; copy what should be judged into profiles/.
tag=${1:-capture}
out=gpurun_out
mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.txt 2>&1; tail -2 $out/${tag}_pytest_gpu.txt
timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2>> $out/${tag}_bench.err
# launch list of the bench (per-launch times are cold-cache and serialised: only the kernels' shares are meaningful)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 > $out/${tag}_ncu_bench.log 2>&1
# one full capture of the dominant kernel (K2 on CTA pairs)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_heads_tc -s 3 -c 1 -o $out/${tag}_k2 \
    python bench.py --steps 2 --warmup 3 > $out/${tag}_ncu_k2.log 2>&1
# training step: three shapes (activations saved by the forward pass; the last line again with PENEO_SAVE_ACT_GB=0 = recompute route)
for args in "--seq-len 1024 --batch 4" "--seq-len 512 --batch 4" "--seq-len 512 --batch 32"; do
  timeout 200 python benchmarks/train_step.py $args 2>/dev/null | tail -1
done > $out/${tag}_train_step.jsonl
PENEO_SAVE_ACT_GB=0 timeout 200 python benchmarks/train_step.py --seq-len 512 --batch 32 2>/dev/null | tail -1 >> $out/${tag}_train_step.jsonl
cat $out/${tag}_train_step.jsonl | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/${tag}_train_launches_b32.csv \
    python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 2 --warmup 1 > $out/${tag}_ncu_train_b32.log 2>&1
PENEO_SAVE_ACT_GB=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/${tag}_train_launches_b32_recompute.csv \
    python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 2 --warmup 1 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_bwd_elem -s 2 -c 1 -o $out/${tag}_t1e \
    python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 1 --warmup 0 > $out/${tag}_ncu_t1.log 2>&1
PENEO_SAVE_ACT_GB=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_bwd_prep -s 2 -c 1 -o $out/${tag}_t1 \
    python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 1 --warmup 0 >> $out/${tag}_ncu_t1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pair_heads_tc -s 1 -c 1 -o $out/${tag}_k2save \
    python benchmarks/train_step.py --seq-len 512 --batch 32 --steps 1 --warmup 1 >> $out/${tag}_ncu_t1.log 2>&1
timeout 400 python benchmarks/decoder_microbench.py > $out/${tag}_decoder_microbench.jsonl 2>/dev/null; tail -3 $out/${tag}_decoder_microbench.jsonl | cut -c1-200
