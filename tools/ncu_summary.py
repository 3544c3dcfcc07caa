#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
KEYS = [
 ("gpu__time_duration.sum", "time"),
 ("launch__registers_per_thread", "regs"),
 ("launch__grid_size", "grid"), ("launch__block_size", "block"),
 ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy%"),
 ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
 ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
 ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_inst%"),
 ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fmaheavy_cyc%"),
 ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_cyc%"),
 ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu_cyc%"),
 ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "fmaheavy_inst%"),
 ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
 ("smsp__inst_executed.sum", "warp_insts"),
 ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
 ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
 ("lts__t_bytes.sum", "l2_bytes"),
 ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
 ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
 ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
 ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long_sb"),
 ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short_sb"),
 ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_barrier"),
 ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
 ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
 ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
 ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
 ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "st_notsel"),
 ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "st_dispatch"),
]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("##", rep)
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print("-", d["Kernel Name"].split("(")[0], "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
        print("   ", ", ".join("%s=%s%s" % (n, d[k], (" " + u[k]) if u.get(k) and n in ("time", "dram_rd", "dram_wr", "l2_bytes") else "") for k, n in KEYS if k in d and n not in ("grid", "block")))
