"""Condenses an `ncu --page raw --csv` export into the few counters DESIGN.md / bench.py quote, one block
per captured launch.  usage: ncu_summary.py raw.csv > summary.txt"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "fmaheavy pipe busy % (IMAD)"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "alu pipe busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sectors_srcunit_tex_op_write.sum", "L2 write sectors"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("zkodst::<unnamed>::", "").replace("void ", "")
    print("== %s  grid %s block %s" % (name, r[col["Grid Size"]], r[col["Block Size"]]))
    for key, label in KEYS:
        if key in col and r[col[key]] != "":
            print("  %-34s %s %s" % (label, r[col[key]], units[col[key]]))
