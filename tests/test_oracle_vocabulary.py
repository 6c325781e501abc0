"""CPU: the vocabulary oracle (oracle/bow_oracle.cc) against the reference's own DBoW2 compiled from its sources
(oracle/_ref/libdbowref.so, where present) and against the committed fixtures generated from it."""
import os

import numpy as np
import pytest

import oracle_lib
import vocab_cases
from orb_slam2_with_comment_b200 import vocabulary

GOLD = os.path.join(os.path.dirname(__file__), "golden", "vocabulary_golden.npz")
KEYS = ("bv_word", "bv_value", "fv_node_id", "fv_feat_off", "fv_feat")


def _port_frames(name):
    voc, scoring, weighting, levelsup, kp_off, desc = vocab_cases.make(name)
    o = oracle_lib.VocabularyOracle(oracle_lib.load_port(), voc, scoring, weighting)
    return [o.transform(desc[kp_off[f]:kp_off[f + 1]], levelsup) for f in range(len(kp_off) - 1)], o


@pytest.mark.parametrize("name", sorted(vocab_cases.CASES))
def test_port_matches_golden(name):
    g = np.load(GOLD)
    frames, o = _port_frames(name)
    assert o.words() == int(g[f"{name}/words"])
    for f, out in enumerate(frames):
        for k in KEYS:
            exp = g[f"{name}/{f}/{k}"]
            assert out[k].dtype.kind == exp.dtype.kind and np.array_equal(out[k], exp), (name, f, k)   # doubles compared bit for bit


@pytest.mark.parametrize("name", sorted(vocab_cases.CASES))
def test_port_matches_reference_dbow2(name, tmp_path):
    ref = oracle_lib.load_dbow_ref()
    if ref is None:
        pytest.skip("oracle/_ref/libdbowref.so not available (needs /root/reference to build)")
    voc, scoring, weighting, levelsup, kp_off, desc = vocab_cases.make(name)
    path = str(tmp_path / "voc.txt")
    vocabulary.write_text_file(path, voc, scoring, weighting)
    r = oracle_lib.VocabularyRef(ref, path)
    frames, o = _port_frames(name)
    assert r.words() == o.words()
    for f, out in enumerate(frames):
        exp = r.transform(desc[kp_off[f]:kp_off[f + 1]], levelsup)
        for k in KEYS:
            assert np.array_equal(out[k], exp[k]), (name, f, k)


def test_text_file_round_trip(tmp_path):
    voc, scoring, weighting, *_ = vocab_cases.make("k4_L5_ragged_stop")
    path = str(tmp_path / "v.txt")
    vocabulary.write_text_file(path, voc, scoring, weighting)
    back = vocabulary.read_text_file(path)
    for k in ("parent", "is_leaf", "desc", "weight"):
        assert np.array_equal(back[k], voc[k]), k
    assert (back["k"], back["L"], back["scoring"], back["weighting"]) == (voc["k"], voc["L"], scoring, weighting)


def test_bow_vector_properties():
    frames, _ = _port_frames("k10_L3_tfidf_l1")
    out = frames[0]
    assert np.all(np.diff(out["bv_word"].astype(np.int64)) > 0) and np.all(np.diff(out["fv_node_id"]) > 0)
    assert abs(out["bv_value"].sum() - 1.0) < 1e-12          # L1-normalised
    assert sorted(out["fv_feat"].tolist()) == list(range(500))   # no stopped words in this vocabulary: every feature appears once
    fo = out["fv_feat_off"]
    for j in range(len(out["fv_node_id"])):
        assert np.all(np.diff(out["fv_feat"][fo[j]:fo[j + 1]]) > 0)
    assert len(frames[1]["bv_word"]) == 0 and len(frames[1]["fv_node_id"]) == 0   # empty frame
