"""Generates tests/golden/vocabulary_golden.npz from the REFERENCE's own DBoW2 (oracle/_ref/libdbowref.so, built from
/root/reference/Thirdparty/DBoW2 by oracle/Makefile) on the seeded cases of tests/vocab_cases.py.
Run in the build container:  python tests/golden/gen_vocabulary_golden.py"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib  # noqa: E402
import vocab_cases  # noqa: E402
from orb_slam2_with_comment_b200 import vocabulary  # noqa: E402

ref = oracle_lib.load_dbow_ref()
assert ref is not None, "needs /root/reference"
out = {}
with tempfile.TemporaryDirectory() as tmp:
    for name in sorted(vocab_cases.CASES):
        voc, scoring, weighting, levelsup, kp_off, desc = vocab_cases.make(name)
        path = os.path.join(tmp, name + ".txt")
        vocabulary.write_text_file(path, voc, scoring, weighting)
        r = oracle_lib.VocabularyRef(ref, path)
        out[f"{name}/words"] = np.int32(r.words())
        for f in range(len(kp_off) - 1):
            o = r.transform(desc[kp_off[f]:kp_off[f + 1]], levelsup)
            for k, v in o.items():
                out[f"{name}/{f}/{k}"] = v
np.savez_compressed(os.path.join(HERE, "vocabulary_golden.npz"), **out)
print("wrote", len(out), "arrays")
