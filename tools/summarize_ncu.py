"""Turn an .ncu-rep (ncu --set full) into the small tracked summaries under profiles/:
   <name>.csv  -- one row per captured launch with the metrics the roofline discussion uses
   (the .ncu-rep files themselves are large and stay in gpurun_out/, which is scratch).

   python tools/summarize_ncu.py gpurun_out/prof_r1c.ncu-rep profiles/r01_kernels_full
"""
import csv
import subprocess
import sys

METRICS = [
    ("Kernel Name", "kernel"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem_kb"),
    ("launch__shared_mem_per_block_static", "static_smem_kb"),
    ("gpu__time_duration.sum", "duration"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__inst_executed.sum", "warp_inst_executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "avg_active_lanes"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe_alu_pct"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "pipe_fmaheavy_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe_xu_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("lts__t_sectors_op_read.sum", "l2_sectors_read"),
    ("lts__t_sectors_op_write.sum", "l2_sectors_write"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_pipe"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(hdr.index(m), name, units[hdr.index(m)]) for m, name in METRICS if m in hdr]
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([f"{name} [{unit}]" if unit else name for _, name, unit in cols])
        for r in data:
            w.writerow([r[i][:90] for i, _, _ in cols])
    print("wrote", out + ".csv", len(data), "launches")


if __name__ == "__main__":
    main()
