#!/usr/bin/env python3
"""DRAM bytes per frame and stage from an `ncu --set full` capture of ONE extraction launch sequence at the bench batch
(tools/ncu_summary.py csv):  python tools/ncu_traffic.py <summary.csv> <batch> > profiles/traffic.json"""
import csv, json, sys
path, batch = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(path)))
hdr = rows[0]
ik, ir, iw, it = hdr.index("kernel"), hdr.index("dram_rd"), hdr.index("dram_wr"), hdr.index("time")
units = rows[1]
def to_bytes(v, u):
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
stage_of = {"k_level0": "pyramid", "k_resize4_mlp": "pyramid", "k_resize4_pp": "pyramid", "k_resize_tma": "pyramid", "k_resize4": "pyramid", "k_resize": "pyramid", "k_border_sides": "pyramid",
            "k_border_caps": "pyramid", "k_border": "pyramid", "k_fast_seg": "fast_cells", "k_blur_tma": "blur", "k_octree": "octree", "k_orient_desc": "orient_desc"}
out, ms = {}, {}
for r in rows[2:]:
    name = r[ik].split("<")[0].strip()
    st = stage_of.get(name)
    if st is None:
        continue
    out[st] = out.get(st, 0.0) + (to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])) / batch
    ms[st] = ms.get(st, 0.0) + float(r[it])
print(json.dumps({"source": f"{path} (ncu --set full --clock-control none, batch {batch}, one launch sequence): dram__bytes_read.sum + dram__bytes_write.sum per stage / {batch} frames",
                  "batch": batch, "dram_bytes_per_frame": out, "ncu_ms_per_stage": ms}, indent=1))
