#!/usr/bin/env python3
"""Per-source-line summary of an ncu report: joins `ncu --page source --csv` (SASS rows with executed-instruction and
stall-sample counts) with `nvdisasm --print-line-info` of the same cubin (instruction offset -> file:line).

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N] [launch id]
"""
import csv, io, os, re, subprocess, sys, tempfile, glob

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "orb_slam2_with_comment_b200", "liborbgpu.so")

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
# the csv holds one block per launch: a 'Kernel Name' line, a header line, rows
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None and row and row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None and row and row[0].startswith("0x"):
        cur["rows"].append(row)
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
blk = blocks[which]
hdr = blk["hdr"]
ci, cs, cn = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
mangled = None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
fn = re.sub(r"\(.*", "", blk["name"]).split("::")[-1].split("<")[0]
lines = None
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
    m = list(re.finditer(r"^\.text\.(\S*%s\S*):$" % re.escape(fn), dis, re.M))
    for mm in m:
        seg = dis[mm.end():]
        end = seg.find("//--------------------- ")
        seg = seg[:end if end >= 0 else None]
        cur_line, lst = None, []
        for l in seg.splitlines():
            k = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if k:
                cur_line = (os.path.basename(k.group(1)), int(k.group(2)))
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
                lst.append(cur_line)
        if len(lst) == len(blk["rows"]):
            lines = lst
            break
    if lines:
        break
if not lines:
    sys.exit(f"could not match SASS of {fn}: ncu has {len(blk['rows'])} instructions")
agg = {}
tot_i = tot_s = 0
for row, ln in zip(blk["rows"], lines):
    i, s = int(row[ci]), int(row[cs])
    a = agg.setdefault(ln, [0, 0, 0])
    a[0] += i; a[1] += s; a[2] += 1
    tot_i += i; tot_s += s
print(f"{blk['name'][:80]}  launch {which}/{len(blocks)}: {tot_i} warp instructions, {tot_s} samples, {len(blk['rows'])} SASS instructions")
src_cache = {}
def src(ln):
    if not ln: return ""
    for d in ("orb_slam2_with_comment_b200/csrc", "include"):
        p = os.path.join(ROOT, d, ln[0])
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[ln[1] - 1].strip()[:90] if ln[1] - 1 < len(L) else ""
    return ""
key = (lambda kv: -kv[1][1]) if os.environ.get("SORT") == "stall" else (lambda kv: -kv[1][0])
for ln, (i, s, n) in sorted(agg.items(), key=key)[:top]:
    print(f"{100.0*i/tot_i:5.1f}% inst {100.0*s/max(tot_s,1):5.1f}% stall  {n:4d} sass  {ln[0] if ln else '?'}:{ln[1] if ln else 0:<4d} {src(ln)}")
