#!/usr/bin/env python3
"""The reference's own nondeterminism envelope (ORBextractor.cc:684): `sort(vSizeAndPointerToNode)` orders equal-size nodes by the
value of an ExtractorNode POINTER, i.e. by heap addresses, so which of several equally populated nodes is expanded last — and with
it which key points survive — depends on the allocator.  The oracle (liborbref.so) replaces that compare by a stable sort on the
size alone (creation order); liborbref_verbatim.so keeps the pointer compare.  This script runs both on the same frames and
reports how far the verbatim build is from the patched one: that distance is the size of the set of results the REFERENCE ITSELF
may legitimately produce, and the patched result is one member of it.

    python tools/envelope_report.py [frames per shape] > profiles/r2_envelope.json      (build container; CPU only)"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from orb_slam2_with_comment_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
patched, verbatim = ol.load_ref(""), ol.load_ref("_verbatim")
rows, tot = [], {"frames": 0, "frames_differing": 0, "keypoints": 0, "keypoints_only_in_one": 0, "descriptor_rows_differing_on_common_keypoints": 0}
for name, (w, h, nf) in (("KITTI 1241x376/2000", (1241, 376, 2000)), ("TUM 640x480/1000", (640, 480, 1000)), ("EuRoC 752x480/1200", (752, 480, 1200))):
    a = ol.Extractor(patched, "orbref", nf, 1.2, 8, 20, 7)
    b = ol.Extractor(verbatim, "orbref", nf, 1.2, 8, 20, 7)
    r = {"shape": name, "frames": n, "frames_differing": 0, "keypoints": 0, "keypoints_only_in_one": 0, "descriptor_rows_differing_on_common_keypoints": 0}
    for s in range(n):
        img = synth.g_rects(w, h, s)
        ka, da = a.extract(img)
        kb, db = b.extract(img)
        A = {(float(k["x"]), float(k["y"]), int(k["octave"])): i for i, k in enumerate(ka)}
        B = {(float(k["x"]), float(k["y"]), int(k["octave"])): i for i, k in enumerate(kb)}
        only = len(set(A) ^ set(B))
        common = set(A) & set(B)
        dd = sum(1 for k in common if not np.array_equal(da[A[k]], db[B[k]]))
        same_order = len(ka) == len(kb) and ka.tobytes() == kb.tobytes()
        r["frames_differing"] += 0 if same_order and only == 0 else 1
        r["keypoints"] += len(ka)
        r["keypoints_only_in_one"] += only
        r["descriptor_rows_differing_on_common_keypoints"] += dd
    rows.append(r)
    for k in tot:
        tot[k] += r[k]
tot["fraction_of_keypoints_only_in_one"] = tot["keypoints_only_in_one"] / max(tot["keypoints"], 1)
print(json.dumps({"what": "liborbref.so (stable tie-break, the oracle) vs liborbref_verbatim.so (pointer compare of ORBextractor.cc:684) on G_rects frames",
                  "per_shape": rows, "total": tot,
                  "reading": "key points that one build selects and the other does not come in pairs (an equally populated node expanded instead of another); "
                             "descriptors of common key points never differ"}, indent=1))
