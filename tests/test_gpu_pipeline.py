"""GPU: extraction -> vocabulary -> matching chained on the device (orbgpu_frame_set_from_extraction), SURVEY §8d config #3
in miniature: consecutive frames are extracted in one batch, their FeatureVectors come from the device vocabulary transform, and
SearchForTriangulation / SearchByBoW run on the device-built frame set.  Checked against the same steps done through the host
(extraction downloaded, FeatureVector and searches by the CPU oracle)."""
import numpy as np
import pytest

import oracle_lib
from orb_slam2_with_comment_b200 import ORBextractor, synth
from orb_slam2_with_comment_b200.matcher import FrameSet, ORBmatcher, match_offsets
from orb_slam2_with_comment_b200.vocabulary import ORBVocabulary

pytestmark = pytest.mark.gpu
W, H, NF = 640, 480, 1000
K = np.array([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], np.float64)


def _frames(n):
    """Consecutive views: the same scene shifted by a few pixels per frame."""
    base = synth.g_rects(W + 64, H, 3)
    return np.stack([np.ascontiguousarray(base[:, 6 * f:6 * f + W]) for f in range(n)])


@pytest.fixture(scope="module")
def chain():
    import torch
    n = 5
    imgs = _frames(n)
    ex = ORBextractor(NF, 1.2, 8, 20, 7, device=0, max_width=W, max_height=H, max_batch=n)
    kp, desc, counts = ex.extract_batch(imgs)
    voc = synth.vocabulary_tree(k=10, L=2, seed=12345)          # the 10 x 10 two-level tree of config #3
    v = ORBVocabulary().from_records(voc)
    m = ORBmatcher(0.6, False)
    h = m.frame_set_from_extraction(ex, v, levelsup=0, kp_flag=0)
    yield dict(torch=torch, n=n, ex=ex, kp=kp, desc=desc, counts=counts, voc=voc, v=v, m=m, h=h)
    m.release(h)
    m.close()
    v.close()
    ex.close()


def _host_frame_set(c, kp_flag):
    """The same frame set built the host way: downloaded extraction + oracle FeatureVectors."""
    counts = c["counts"]
    kp_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    keys = np.concatenate([c["kp"][f][:counts[f]] for f in range(c["n"])])
    desc = np.concatenate([c["desc"][f][:counts[f]] for f in range(c["n"])])
    o = oracle_lib.VocabularyOracle(oracle_lib.load_port(), c["voc"])
    node_off, node_id, feat_off, feat = [0], [], [0], []
    for f in range(c["n"]):
        r = o.transform(desc[kp_off[f]:kp_off[f + 1]], 0)
        node_id.append(r["fv_node_id"])
        feat_off.append(feat_off[-1][-1] + r["fv_feat_off"][1:] if isinstance(feat_off[-1], np.ndarray) else feat_off[-1] + r["fv_feat_off"][1:])
        feat.append(r["fv_feat"])
        node_off.append(node_off[-1] + len(r["fv_node_id"]))
    feat_off = np.concatenate([[0]] + [np.atleast_1d(x) for x in feat_off[1:]]).astype(np.int32)
    fs = FrameSet(kp_off, keys, desc, kp_flags=np.full(len(keys), kp_flag, np.uint8), fv_node_off=np.array(node_off, np.int32),
                  fv_node_id=np.concatenate(node_id).astype(np.int32), fv_feat_off=feat_off, fv_feat=np.concatenate(feat).astype(np.int32))
    return fs


def test_device_frame_set_equals_host_built(chain):
    c = chain
    fs = _host_frame_set(c, 0)
    got = c["m"].frame_set_download(c["h"])
    assert np.array_equal(got["kp_off"], fs.kp_off)
    assert got["keys_un"].tobytes() == fs.keys_un.tobytes() and np.array_equal(got["desc"], fs.desc)
    for k in ("fv_node_off", "fv_node_id", "fv_feat_off", "fv_feat"):
        assert np.array_equal(got[k], getattr(fs, k)), k
    assert int(c["counts"].min()) > 500


def test_triangulation_on_the_chained_frame_set(chain):
    c, torch = chain, chain["torch"]
    fs = _host_frame_set(c, 0)
    n = c["n"]
    idx1, idx2 = np.arange(1, n, dtype=np.int32), np.arange(0, n - 1, dtype=np.int32)
    F, ep = synth.fundamental_and_epipole(K, np.eye(3), np.array([0.3, 0.01, 0.05]))
    F12, epi = np.tile(F, (n - 1, 1)).astype(np.float32), np.tile(ep, (n - 1, 1)).astype(np.float32)
    sf, s2 = synth.scale_tables()
    exp = oracle_lib.MatcherOracle(oracle_lib.load_port(), 0.6, False).SearchForTriangulation(fs, fs, idx1, idx2, F12, epi, sf, s2)
    off, total = match_offsets(fs, idx1)
    dev = torch.device("cuda:0")
    d12 = torch.full((total,), -7, dtype=torch.int32, device=dev)
    dd = torch.zeros(total, dtype=torch.int32, device=dev)
    dn = torch.zeros(n - 1, dtype=torch.int32, device=dev)
    c["m"].search_for_triangulation_dev(c["h"], c["h"], idx1, idx2, F12, epi, sf, s2, off, d12.data_ptr(), dd.data_ptr(), dn.data_ptr())
    c["m"].sync()
    assert np.array_equal(dn.cpu().numpy(), exp["nmatches"]) and int(exp["nmatches"].sum()) > 50
    assert np.array_equal(d12.cpu().numpy(), exp["match12"])
    assert np.array_equal(dd.cpu().numpy()[exp["match12"] >= 0], exp["match_dist"][exp["match12"] >= 0])


def test_bow_search_on_the_chained_frame_set(chain):
    c, torch = chain, chain["torch"]
    m = ORBmatcher(0.75, True)
    h = m.frame_set_from_extraction(c["ex"], c["v"], levelsup=0, kp_flag=1)     # every key point "has a MapPoint"
    fs = _host_frame_set(c, 1)
    n = c["n"]
    idx1, idx2 = np.arange(1, n, dtype=np.int32), np.arange(0, n - 1, dtype=np.int32)
    exp = oracle_lib.MatcherOracle(oracle_lib.load_port(), 0.75, True).SearchByBoW(fs, fs, idx1, idx2)
    off, total = match_offsets(fs, idx1)
    dev = torch.device("cuda:0")
    d12 = torch.full((total,), -7, dtype=torch.int32, device=dev)
    dd = torch.zeros(total, dtype=torch.int32, device=dev)
    dn = torch.zeros(n - 1, dtype=torch.int32, device=dev)
    m.search_by_bow_dev(h, h, idx1, idx2, off, d12.data_ptr(), dd.data_ptr(), dn.data_ptr())
    m.sync()
    assert np.array_equal(dn.cpu().numpy(), exp["nmatches"]) and int(exp["nmatches"].sum()) > 100
    assert np.array_equal(d12.cpu().numpy(), exp["match12"])
    m.release(h)
    m.close()


def test_errors(chain):
    from orb_slam2_with_comment_b200 import capi
    ex = ORBextractor(NF, 1.2, 8, 20, 7, device=0, max_width=W, max_height=H, max_batch=1)
    with pytest.raises(capi.OrbGpuError):
        chain["m"].frame_set_from_extraction(ex)          # nothing extracted yet
    ex.close()
