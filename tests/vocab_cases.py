"""Seeded vocabulary workloads shared by the oracle tests, the golden generator and the GPU parity tests."""
from __future__ import annotations

import numpy as np

from orb_slam2_with_comment_b200 import synth

# name -> (tree kwargs, scoring, weighting, levelsup, descriptors per frame list)
CASES = {
    "k10_L3_tfidf_l1": (dict(k=10, L=3, seed=1), 0, 0, 2, [500, 0, 1, 777]),
    "k10_L4_levelsup4": (dict(k=10, L=4, seed=2), 0, 0, 4, [300, 1200]),        # L - levelsup = 0: every feature in node 0
    "k4_L5_ragged_stop": (dict(k=4, L=5, seed=3, ragged=True, stop_frac=0.15), 0, 0, 3, [900, 33]),
    "k3_L6_l2_tf": (dict(k=3, L=6, seed=4), 1, 1, 4, [600]),
    "k8_L3_idf_dot": (dict(k=8, L=3, seed=5, stop_frac=0.05), 5, 2, 1, [450, 450]),
    "k8_L3_tfidf_dot": (dict(k=8, L=3, seed=6), 5, 0, 1, [450]),
    "k20_L2_binary_chi": (dict(k=20, L=2, seed=7), 2, 3, 1, [800]),
    "k10_L3_levelsup0": (dict(k=10, L=3, seed=8), 0, 0, 0, [400]),              # FeatureVector keyed by the leaves themselves
    "k10_L3_levelsup9": (dict(k=10, L=3, seed=9), 0, 0, 9, [400]),              # nid_level < 0: node 0
}


def make(name):
    kw, scoring, weighting, levelsup, per = CASES[name]
    voc = synth.vocabulary_tree(**kw)
    descs = [synth.vocabulary_descriptors(voc, n, seed=1000 + i) if n else np.zeros((0, 32), np.uint8) for i, n in enumerate(per)]
    # exact duplicates and exact ties: a few descriptors equal to node descriptors
    if len(descs[0]) > 20:
        descs[0][:8] = voc["desc"][np.nonzero(voc["is_leaf"])[0][:8]]
        descs[0][8:12] = descs[0][0]
    kp_off = np.concatenate([[0], np.cumsum([len(d) for d in descs])]).astype(np.int32)
    return voc, scoring, weighting, levelsup, kp_off, np.concatenate(descs)
