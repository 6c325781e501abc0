"""Single-process multi-GPU extraction (orbgpu_multi_extract_batch: one host thread + extractor per device, contiguous frame
ranges, host gather): an N-device run must equal the one-device run byte for byte and in input order."""
import numpy as np
import pytest

from orb_slam2_with_comment_b200 import ORBextractor, synth
from orb_slam2_with_comment_b200.extractor import MultiGpuExtractor

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _frames(n, w, h):
    return np.ascontiguousarray(np.stack([synth.g_rects(w, h, 100 + s) if s % 5 else synth.g_blurnoise(w, h, s) for s in range(n)]))


def test_one_device_dispatcher_equals_plain_call():
    w, h, nf, B = 640, 480, 1000, 13
    imgs = _frames(B, w, h)
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    kp, desc, cnt = ex.extract_batch(imgs)
    me = MultiGpuExtractor([0], nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch_per_device=5)   # ranges longer than a pass
    mkp, mdesc, mcnt = me.extract_batch(imgs)
    assert np.array_equal(cnt, mcnt) and cnt.min() > 500
    for f in range(B):
        assert kp[f, :cnt[f]].tobytes() == mkp[f, :cnt[f]].tobytes() and np.array_equal(desc[f, :cnt[f]], mdesc[f, :cnt[f]])
    assert me.last_launches() > 0
    me.close(); ex.close()


@pytest.mark.parametrize("shape", [(1241, 376, 2000), (752, 480, 1200)])
def test_n_device_run_equals_one_device_run(shape):
    n = _ndev()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    w, h, nf = shape
    B = 8 * n + 3          # ragged ranges
    imgs = _frames(B, w, h)
    one = MultiGpuExtractor([0], nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch_per_device=16)
    kp1, d1, c1 = one.extract_batch(imgs)
    one.close()
    for devs in (list(range(n)), list(range(n))[::-1], [n - 1, 0]):
        me = MultiGpuExtractor(devs, nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch_per_device=16)
        ranges = [me.frame_range(B, g) for g in range(len(devs))]
        assert ranges[0][0] == 0 and ranges[-1][1] == B and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        kp, d, c = me.extract_batch(imgs)
        assert np.array_equal(c, c1), devs
        for f in range(B):
            assert kp[f, :c[f]].tobytes() == kp1[f, :c[f]].tobytes() and np.array_equal(d[f, :c[f]], d1[f, :c[f]]), (devs, f)
        me.close()


def test_dispatcher_argument_errors():
    from orb_slam2_with_comment_b200.capi import OrbGpuError
    with pytest.raises(OrbGpuError):
        MultiGpuExtractor([0, 0], 1000, 1.2, 8, 20, 7)
    with pytest.raises(OrbGpuError):
        MultiGpuExtractor([_ndev() + 3], 1000, 1.2, 8, 20, 7)
