"""Key metrics of an .ncu-rep (captured with `ncu --set full`) as text for profiles/:

    python benchmarks/ncu_metrics.py gpurun_out/r02_v2_k2.ncu-rep > profiles/r02_v2_k2_ncu_metrics.txt
"""
import csv
import subprocess
import sys

WANT = """Kernel Name
dram__bytes_read.sum
dram__bytes_write.sum
gpu__time_duration.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
launch__block_size
launch__cluster_size
launch__grid_size
launch__registers_per_thread
launch__shared_mem_per_block_dynamic
lts__t_sector_hit_rate.pct
lts__t_sectors_srcunit_tex_op_read.sum
sm__cycles_elapsed.max
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed
dram__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active""".split("\n")

for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        d = {a: (b, c) for a, b, c in zip(h, u, r)}
        for w in WANT:
            if w in d:
                print(f"{w:90s} {d[w][0]:12s} {d[w][1]}")
        print()
