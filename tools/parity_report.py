#!/usr/bin/env python3
"""Parity report on a B200: CUDA extraction vs the CPU checker over many seeded frames of the three BASELINE shapes —
key-point fields bit-exact, angle deviation, fraction of key points whose descriptor differs (north_star: <= 0.1 %).
The checker is the reference's own ORBextractor.cc (oracle/_ref/liborbref.so, stable tie-break at :684) when that build
travelled to the box, else the port; the line says which."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import oracle_lib as ol
from orb_slam2_with_comment_b200 import ORBextractor, synth

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 48
lib, prefix, checker = ol.load_ref(), "orbref", "reference ORBextractor.cc (oracle/_ref/liborbref.so)"
if lib is None:
    lib, prefix, checker = ol.load_port(), "orbo", "port (oracle/orb_oracle.cc)"
for name, (w, h, nf) in {"kitti 1241x376/2000": (1241, 376, 2000), "tum 640x480/1000": (640, 480, 1000), "euroc 752x480/1200": (752, 480, 1200)}.items():
    g = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=n_seeds)
    o = ol.Extractor(lib, prefix, nf, 1.2, 8, 20, 7)
    gens = [synth.g_rects, synth.g_blurnoise, synth.g_uniform]
    imgs = np.stack([gens[s % 3](w, h, 5000 + s) for s in range(n_seeds)])
    kp, desc, cnt = g.extract_batch(imgs)
    tot = dict(frames=n_seeds, keypoints=0, field_mismatches=0, angle_bit_mismatches=0, max_angle_dev_deg=0.0, descriptor_rows_differing=0, descriptor_bits_differing=0)
    for f in range(n_seeds):
        ekp, edesc = o.extract(imgs[f])
        k, d = kp[f, :cnt[f]], desc[f, :cnt[f]]
        assert len(k) == len(ekp)
        tot["keypoints"] += len(k)
        for fld in ("x", "y", "size", "response", "octave", "class_id"):
            tot["field_mismatches"] += int(np.count_nonzero(k[fld] != ekp[fld]))
        tot["angle_bit_mismatches"] += int(np.count_nonzero(k["angle"] != ekp["angle"]))
        tot["max_angle_dev_deg"] = max(tot["max_angle_dev_deg"], float(np.abs(k["angle"].astype(np.float64) - ekp["angle"]).max()))
        tot["descriptor_rows_differing"] += int(np.count_nonzero((d != edesc).any(1)))
        tot["descriptor_bits_differing"] += int(np.unpackbits(d ^ edesc).sum())
    tot["descriptor_rows_differing_pct"] = 100.0 * tot["descriptor_rows_differing"] / tot["keypoints"]
    print(json.dumps({"shape": name, "checker": checker, **tot}))
    g.close()
