"""Compact summary of an .ncu-rep: python ncu_key.py file.ncu-rep [kernel-index]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct", "sm__inst_executed_pipe_fp16.avg.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__pcsamp_warps_issue_stalled"]
for r in rows[2:]:
    print("=" * 100)
    for a, b, c in zip(h, u, r):
        if any(a.startswith(w) or a == w for w in want) and "not_issued" not in a:
            if a.startswith("smsp__pcsamp") and (c in ("0", "") ):
                continue
            print(f"{a:100s} {b:12s} {c}")
