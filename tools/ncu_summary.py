#!/usr/bin/env python3
"""One line per kernel launch of an ncu report: python tools/ncu_summary.py <report.ncu-rep> [out.csv]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs"),
        ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("smsp__inst_executed.sum", "warp_inst"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"), ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%")]
idx = [(n, hdr.index(k)) for k, n in want if k in hdr]
out = [[n for n, _ in idx], [units[i] for _, i in idx]]
for r in rows[2:]:
    out.append([r[i].split("(")[0].replace("void ", "").replace("og::", "") if n == "kernel" else r[i] for n, i in idx])
if len(sys.argv) > 2:
    csv.writer(open(sys.argv[2], "w")).writerows(out)
for r in out:
    print(" ".join(f"{(v[:20] if j == 0 else v[:9]):>{20 if j == 0 else 9}}" for j, v in enumerate(r)))
