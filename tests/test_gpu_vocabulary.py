"""GPU parity: ORBVocabulary::transform on the device (csrc/og_vocab.cu through the C ABI) against the CPU oracle
(oracle/bow_oracle.cc, itself pinned to the reference's DBoW2) — node / word ids and feature lists bit-exact, BowVector
values compared as raw doubles."""
import numpy as np
import pytest

import oracle_lib
import vocab_cases
from orb_slam2_with_comment_b200 import synth
from orb_slam2_with_comment_b200.vocabulary import ORBVocabulary

pytestmark = pytest.mark.gpu


def _check_batch(out, oracle, kp_off, desc, levelsup):
    for f in range(len(kp_off) - 1):
        exp = oracle.transform(desc[kp_off[f]:kp_off[f + 1]], levelsup)
        b0, b1 = out["bv_off"][f], out["bv_off"][f + 1]
        assert np.array_equal(out["bv_word"][b0:b1], exp["bv_word"]), f
        assert np.array_equal(out["bv_value"][b0:b1].view(np.uint64), exp["bv_value"].view(np.uint64)), f
        n0, n1 = out["fv_node_off"][f], out["fv_node_off"][f + 1]
        assert np.array_equal(out["fv_node_id"][n0:n1], exp["fv_node_id"]), f
        fo = out["fv_feat_off"][n0:n1 + 1]
        assert np.array_equal(fo - fo[0], exp["fv_feat_off"]), f
        assert np.array_equal(out["fv_feat"][fo[0]:fo[-1]], exp["fv_feat"]), f
        assert np.array_equal(out["word_of_feature"][kp_off[f]:kp_off[f + 1]], exp["word_of_feature"]), f
        assert np.array_equal(out["node_of_feature"][kp_off[f]:kp_off[f + 1]], exp["node_of_feature"]), f


@pytest.mark.parametrize("name", sorted(vocab_cases.CASES))
def test_transform_matches_oracle(name):
    voc, scoring, weighting, levelsup, kp_off, desc = vocab_cases.make(name)
    o = oracle_lib.VocabularyOracle(oracle_lib.load_port(), voc, scoring, weighting)
    v = ORBVocabulary().from_records(voc, scoring, weighting)
    assert v.size() == o.words()
    out = v.transform_batch(kp_off, desc, levelsup)
    assert v.last_launches == 4
    _check_batch(out, o, kp_off, desc, levelsup)


def test_frames_beyond_the_shared_memory_sort():
    """A frame with more than 8192 descriptors (the reference's transform has no size limit, TemplatedVocabulary.h:1127-1197): the
    batch runs the frame kernel on its HBM workspace; same results as the oracle for the large frame and for the small frames beside it."""
    voc, scoring, weighting, levelsup, _, _ = vocab_cases.make("k10_L3_tfidf_l1")
    rs = np.random.RandomState(77)
    kp_off = np.array([0, 700, 700 + 9001, 700 + 9001 + 1300], np.int32)
    desc = rs.randint(0, 256, (int(kp_off[-1]), 32)).astype(np.uint8)
    o = oracle_lib.VocabularyOracle(oracle_lib.load_port(), voc, scoring, weighting)
    v = ORBVocabulary().from_records(voc, scoring, weighting)
    out = v.transform_batch(kp_off, desc, levelsup)
    _check_batch(out, o, kp_off, desc, levelsup)
    out = v.transform_batch(kp_off[:2], desc[:700], levelsup)      # and back to the shared-memory path on the same handle
    _check_batch(out, o, kp_off[:2], desc[:700], levelsup)


def test_text_file_and_single_frame_maps(tmp_path):
    from orb_slam2_with_comment_b200 import vocabulary
    voc, scoring, weighting, levelsup, kp_off, desc = vocab_cases.make("k4_L5_ragged_stop")
    path = str(tmp_path / "voc.txt")
    vocabulary.write_text_file(path, voc, scoring, weighting)
    v = ORBVocabulary()
    assert v.loadFromTextFile(path) and not v.empty()
    bv, fv = v.transform(desc[:kp_off[1]], levelsup)
    exp = oracle_lib.VocabularyOracle(oracle_lib.load_port(), voc, scoring, weighting).transform(desc[:kp_off[1]], levelsup)
    assert list(bv.keys()) == exp["bv_word"].tolist() and list(bv.values()) == exp["bv_value"].tolist()
    assert list(fv.keys()) == exp["fv_node_id"].tolist()
    assert sum(len(x) for x in fv.values()) == len(exp["fv_feat"]) < kp_off[1]     # stopped words dropped some features


def test_orbvoc_sized_vocabulary():
    """k=10, L=6 (1.1 M nodes, the shape of ORBvoc.txt), levelsup 4 as Frame::ComputeBoW uses: 64 frames x ~2000 descriptors."""
    voc = synth.vocabulary_tree_full(10, 6, seed=11)
    rs = np.random.RandomState(5)
    per = rs.randint(1800, 2013, 64)
    kp_off = np.concatenate([[0], np.cumsum(per)]).astype(np.int32)
    desc = synth.vocabulary_descriptors_fast(voc, int(kp_off[-1]), seed=77)
    v = ORBVocabulary().from_records(voc)
    assert v.size() == 10 ** 6
    out = v.transform_batch(kp_off, desc, 4)
    o = oracle_lib.VocabularyOracle(oracle_lib.load_port(), voc)
    sub = np.array([0, 1, 31, 63])
    sub_off = np.concatenate([[0], np.cumsum(per[sub])]).astype(np.int32)
    for i, f in enumerate(sub):       # the oracle on four of the frames
        exp = o.transform(desc[kp_off[f]:kp_off[f + 1]], 4)
        b0, b1 = out["bv_off"][f], out["bv_off"][f + 1]
        assert np.array_equal(out["bv_word"][b0:b1], exp["bv_word"])
        assert np.array_equal(out["bv_value"][b0:b1].view(np.uint64), exp["bv_value"].view(np.uint64))
        n0, n1 = out["fv_node_off"][f], out["fv_node_off"][f + 1]
        assert np.array_equal(out["fv_node_id"][n0:n1], exp["fv_node_id"])
    # size-independent properties on all frames
    assert out["fv_feat_off"][-1] == kp_off[-1] and len(out["fv_feat"]) == kp_off[-1]      # no stopped words: every feature once
    for f in range(64):
        n0, n1 = out["fv_node_off"][f], out["fv_node_off"][f + 1]
        ids = out["fv_node_id"][n0:n1]
        assert np.all(np.diff(ids) > 0) and ids.min() >= 11 and ids.max() <= 110           # level-2 nodes of a breadth-first tree
        fo = out["fv_feat_off"][n0:n1 + 1]
        feats = out["fv_feat"][fo[0]:fo[-1]]
        assert np.array_equal(np.sort(feats), np.arange(per[f]))
        b0, b1 = out["bv_off"][f], out["bv_off"][f + 1]
        assert np.all(np.diff(out["bv_word"][b0:b1].astype(np.int64)) > 0)
        assert abs(out["bv_value"][b0:b1].sum() - 1.0) < 1e-9
    # node_of_feature is the level-2 ancestor of word_of_feature's leaf: parent chain in the records
    parent = np.concatenate([[0], voc["parent"]])
    leaf_node = np.nonzero(np.concatenate([[0], voc["is_leaf"]]))[0]
    anc = leaf_node[out["word_of_feature"][:5000]]
    for _ in range(4):
        anc = parent[anc]
    assert np.array_equal(anc, out["node_of_feature"][:5000])


def test_device_variant_feeds_frame_set():
    """_dev entry: descriptors and outputs stay on the device (torch tensors as plain device memory)."""
    import ctypes as C

    import torch
    from orb_slam2_with_comment_b200 import capi, vocabulary
    voc, scoring, weighting, levelsup, kp_off, desc = vocab_cases.make("k10_L3_tfidf_l1")
    v = ORBVocabulary().from_records(voc, scoring, weighting)
    dev = torch.device("cuda:0")
    n, nf = int(kp_off[-1]), len(kp_off) - 1
    d_off = torch.from_numpy(kp_off).to(dev)
    d_desc = torch.from_numpy(desc).to(dev)
    o = {"bv_off": torch.zeros(nf + 1, dtype=torch.int32, device=dev), "bv_word": torch.zeros(n, dtype=torch.int32, device=dev),
         "bv_value": torch.zeros(n, dtype=torch.float64, device=dev), "fv_node_off": torch.zeros(nf + 1, dtype=torch.int32, device=dev),
         "fv_node_id": torch.zeros(n, dtype=torch.int32, device=dev), "fv_feat_off": torch.zeros(n + 1, dtype=torch.int32, device=dev),
         "fv_feat": torch.zeros(n, dtype=torch.int32, device=dev)}
    torch.cuda.synchronize()
    L = vocabulary._lib()
    capi.check(L.orbgpu_bow_transform_dev(v._h, nf, d_off.data_ptr(), n, int(np.diff(kp_off).max()), d_desc.data_ptr(), levelsup,
                                          *[C.c_void_p(o[k].data_ptr()) for k in ("bv_off", "bv_word", "bv_value", "fv_node_off", "fv_node_id",
                                                                                   "fv_feat_off", "fv_feat")], None, None))
    capi.check(L.orbgpu_vocabulary_sync(v._h))
    host = v.transform_batch(kp_off, desc, levelsup)
    nw, nn = int(host["bv_off"][-1]), int(host["fv_node_off"][-1])
    assert np.array_equal(o["bv_off"].cpu().numpy(), host["bv_off"]) and np.array_equal(o["fv_node_off"].cpu().numpy(), host["fv_node_off"])
    assert np.array_equal(o["bv_word"].cpu().numpy()[:nw].view(np.uint32), host["bv_word"])
    assert np.array_equal(o["bv_value"].cpu().numpy()[:nw], host["bv_value"])
    assert np.array_equal(o["fv_node_id"].cpu().numpy()[:nn], host["fv_node_id"])
    assert np.array_equal(o["fv_feat_off"].cpu().numpy()[:nn + 1], host["fv_feat_off"])
    assert np.array_equal(o["fv_feat"].cpu().numpy()[:len(host["fv_feat"])], host["fv_feat"])


def test_argument_errors():
    from orb_slam2_with_comment_b200 import capi
    voc, *_ = vocab_cases.make("k10_L3_tfidf_l1")
    bad = dict(voc)
    bad["parent"] = voc["parent"].copy()
    bad["parent"][5] = 900          # a parent that comes after its child
    with pytest.raises(capi.OrbGpuError):
        ORBVocabulary().from_records(bad)
    v = ORBVocabulary().from_records(voc)
    with pytest.raises(capi.OrbGpuError):
        ORBVocabulary().transform_batch(np.array([0, 1], np.int32), np.zeros((1, 32), np.uint8))
