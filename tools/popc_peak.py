#!/usr/bin/env python3
"""Measures the POPC / LOP3 / IADD3 issue peaks of this GPU (tools/popc_peak.cu) while sampling the SM clock with nvidia-smi, and
writes profiles/popc_peak.json: operations per second (the matching roofline's denominator) and per clock per SM at the clock
the run actually had.  Run on the GPU box:  python tools/popc_peak.py"""
import json, os, subprocess, threading, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
exe = os.path.join(ROOT, "tools", "_build", "popc_peak")
os.makedirs(os.path.dirname(exe), exist_ok=True)
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe, os.path.join(ROOT, "tools", "popc_peak.cu")])
rows = []
smi = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-lms", "50"],
                       stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.split(",")) for l in smi.stdout], daemon=True).start()
out = json.loads(subprocess.check_output([exe], text=True))
smi.terminate()
clk = [float(r[0]) for r in rows if r and r[0].strip().replace(".", "").isdigit()]
# the first and last samples may be idle clocks: the run is ~2 s of steady load, take the median of the upper half
clk.sort()
mhz = statistics.median(clk[len(clk) // 2:]) if clk else out["sm_clock_attr_mhz"]
res = dict(out)
res.update({"sm_mhz_during": mhz, "sm_mhz_samples": len(clk), "sm_mhz_min_max": [clk[0], clk[-1]] if clk else None,
            "popc_per_clk_per_sm": out["popc_per_s"] / (mhz * 1e6) / out["sms"], "lop3_per_clk_per_sm": out["lop3_per_s"] / (mhz * 1e6) / out["sms"],
            "iadd3_per_clk_per_sm": out["iadd3_per_s"] / (mhz * 1e6) / out["sms"],
            "note": "thread-level operations; SM clock = median of the upper half of nvidia-smi clocks.sm samples (50 ms period) taken while the kernels ran"})
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "popc_peak.json")
json.dump(res, open(path, "w"))
print(json.dumps(res))
