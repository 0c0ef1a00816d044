#!/usr/bin/env python
"""Write the per-kernel summary CSV kept under profiles/ from an `ncu --set full` report.
usage: ncu_summary.py REPORT.ncu-rep OUT.csv [pictures_per_launch]"""
import csv, json, subprocess, sys
rep, out_csv = sys.argv[1], sys.argv[2]
pics = int(sys.argv[3]) if len(sys.argv) > 3 else 256
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h = r[0]
ki = h.index("Kernel Name")
cols = [h.index(w) for w in want if w in h]
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["ID", "Kernel Name"] + [h[c] for c in cols] + ["dram_bytes_per_picture"])
    w.writerow(["", ""] + [r[1][c] for c in cols] + ["byte"])
    traffic = {}
    for row in r[2:]:
        rd = float(row[h.index("dram__bytes_read.sum")]); wr = float(row[h.index("dram__bytes_write.sum")])
        unit = r[1][h.index("dram__bytes_read.sum")]
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(unit, 1)
        per_pic = (rd + wr) * scale / pics
        w.writerow([row[0], row[ki]] + [row[c] for c in cols] + ["%.0f" % per_pic])
        traffic[row[ki].split("(")[0].replace("void ", "")] = per_pic
print(json.dumps(traffic, indent=1))
