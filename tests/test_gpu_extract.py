"""Parity of the CUDA extraction path (through the C ABI) against the CPU oracle — stage by stage and end to end.

Bar (BASELINE.json north_star): keypoints (x, y, octave, response, size, order), pyramid, blur and descriptors
bit-exact; angle bit-exact expected (the fastAtan2 polynomial is reproduced), 1e-3 deg allowed; descriptors
BIT-EXACT: the device evaluates cos / sin of the key-point angle with glibc's sincosf algorithm (og_math.cuh
sincosf_glibc, equal to glibc 2.39 on every float in [0, 2 pi]), so the rotated-BRIEF offsets are the reference's.
"""
import os

import numpy as np
import pytest

import oracle_lib as ol
from orb_slam2_with_comment_b200 import synth

pytestmark = pytest.mark.gpu

SHAPES = {"kitti": (1241, 376, 2000), "tum": (640, 480, 1000), "euroc": (752, 480, 1200)}


@pytest.fixture(scope="module")
def gpu():
    from orb_slam2_with_comment_b200 import ORBextractor
    made = {}

    def get(nf, w, h, batch=1):
        key = (nf, w, h, batch)
        if key not in made:
            made[key] = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=batch)
        return made[key]
    yield get
    for e in made.values():
        e.close()


def compare_final(kp, desc, ekp, edesc, what=""):
    assert len(kp) == len(ekp), f"{what}: {len(kp)} keypoints vs oracle {len(ekp)}"
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(kp[f], ekp[f]), f"{what}: field {f} differs at {np.nonzero(kp[f] != ekp[f])[0][:5]}"
    dang = np.abs(kp["angle"].astype(np.float64) - ekp["angle"].astype(np.float64))
    assert (dang <= 1e-3).all(), f"{what}: angle differs by up to {dang.max()}"
    n_ang = int(np.count_nonzero(kp["angle"] != ekp["angle"]))
    bad_rows = int(np.count_nonzero((desc != edesc).any(1)))
    bad_bits = int(np.unpackbits(desc ^ edesc).sum())
    assert bad_rows == 0, f"{what}: {bad_rows} of {len(kp)} descriptors differ"
    return n_ang, bad_rows, bad_bits


@pytest.mark.parametrize("name", list(SHAPES))
def test_tables(gpu, oracle, name):
    w, h, nf = SHAPES[name]
    g = gpu(nf, w, h)
    s, q, u = ol.Extractor(oracle, "orbo", nf, 1.2, 8, 20, 7).tables()
    assert np.array_equal(g.GetScaleFactors(), s[0]) and np.array_equal(g.GetInverseScaleFactors(), s[1])
    assert np.array_equal(g.GetScaleSigmaSquares(), s[2]) and np.array_equal(g.GetInverseScaleSigmaSquares(), s[3])
    assert np.array_equal(g.mnFeaturesPerLevel, q) and np.array_equal(g.umax, u)


@pytest.mark.parametrize("name", list(SHAPES))
def test_stages_match_oracle(gpu, oracle, name):
    w, h, nf = SHAPES[name]
    g = gpu(nf, w, h)
    o = ol.Extractor(oracle, "orbo", nf, 1.2, 8, 20, 7)
    for seed, gen in ((0, synth.g_rects), (1, synth.g_blurnoise)):
        img = gen(w, h, seed)
        kp, desc = g(img)
        ekp, edesc = o.extract(img)
        for l in range(8):
            assert np.array_equal(g.level(l, bordered=True), o.level(l, True)), f"pyramid level {l} (bordered)"
            cand, ecand = g.level_points(l, 0), o.level_points(l, 0)
            assert len(cand) == len(ecand), f"level {l}: {len(cand)} candidates vs {len(ecand)}"
            for f in ("x", "y", "response"):
                assert np.array_equal(cand[f], ecand[f]), f"level {l} candidate {f}"
            sel, esel = g.level_points(l, 1), o.level_points(l, 1)
            assert len(sel) == len(esel), f"level {l}: {len(sel)} selected vs {len(esel)}"
            for f in ("x", "y", "response", "size", "octave", "angle"):
                assert np.array_equal(sel[f], esel[f]), f"level {l} selected {f}"
            eb = o.blurred(l)
            if eb is not None:
                assert np.array_equal(g.blurred(l), eb), f"blurred level {l}"
        compare_final(kp, desc, ekp, edesc, f"{name} seed {seed}")


@pytest.mark.parametrize("name", list(SHAPES))
def test_end_to_end_seeds(gpu, oracle, name):
    w, h, nf = SHAPES[name]
    g = gpu(nf, w, h)
    o = ol.Extractor(oracle, "orbo", nf, 1.2, 8, 20, 7)
    tot = [0, 0, 0, 0]
    for seed in range(100, 108):
        img = synth.g_rects(w, h, seed)
        kp, desc = g(img)
        ekp, edesc = o.extract(img)
        a, r, b = compare_final(kp, desc, ekp, edesc, f"{name} seed {seed}")
        tot[0] += a; tot[1] += r; tot[2] += b; tot[3] += len(kp)
    print(f"\n[{name}] keypoints {tot[3]}: angle bit-mismatches {tot[0]}, descriptor rows differing {tot[1]} "
          f"({100.0 * tot[1] / tot[3]:.4f} %), differing bits {tot[2]}")
    assert tot[1] == 0 and tot[0] == 0


def test_golden_fixtures(gpu):
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "extractor_golden.npz"))
    for name in [str(c) for c in G["cases"]]:
        w, h, seed, nf = [int(v) for v in G[name + "_meta"]]
        img = getattr(synth, str(G[name + "_gen"]))(w, h, seed)
        kp, desc = gpu(nf, w, h)(img)
        compare_final(kp, desc, G[name + "_kp"], G[name + "_desc"], name)


def test_edge_cases(gpu, oracle):
    g = gpu(1000, 640, 480)
    kp, desc = g(synth.g_flat(640, 480))
    assert len(kp) == 0 and desc.shape == (0, 32)           # _descriptors.release() path (:1064-1065)
    kp, desc = g(np.zeros((0, 0), np.uint8))
    assert len(kp) == 0                                      # empty image: silent return (:1046)
    o = ol.Extractor(oracle, "orbo", 1000, 1.2, 8, 20, 7)
    for img in (synth.g_half_flat(640, 480, 7), synth.g_uniform(640, 480, 8)):
        kp, desc = g(img)
        ekp, edesc = o.extract(img)
        compare_final(kp, desc, ekp, edesc, "edge")
    # a smaller frame through an extractor sized for a larger one, and a strided (non-contiguous rows) input
    big = synth.g_rects(640, 480, 9)
    sub = big[:300, :400]
    kp, desc = g(sub)
    ekp, edesc = o.extract(np.ascontiguousarray(sub))
    compare_final(kp, desc, ekp, edesc, "strided sub-image")


def test_rejects_bad_geometry(gpu):
    from orb_slam2_with_comment_b200.capi import OrbGpuError
    g = gpu(1000, 640, 480)
    with pytest.raises(OrbGpuError):
        g(np.zeros((100, 100), np.uint8) + 7)     # level 7 would be 28 x 28: no room for a cell
    with pytest.raises(OrbGpuError):
        g(np.zeros((481, 700), np.uint8))         # larger than max_width x max_height


def test_batch_equals_single_and_is_order_stable(gpu, oracle):
    from orb_slam2_with_comment_b200 import ORBextractor
    w, h, nf = SHAPES["euroc"]
    B = 6
    imgs = np.stack([synth.g_rects(w, h, 200 + i) for i in range(B)])
    gb = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    kp, desc, cnt = gb.extract_batch(imgs)
    o = ol.Extractor(oracle, "orbo", nf, 1.2, 8, 20, 7)
    for f in range(B):
        ekp, edesc = o.extract(imgs[f])
        compare_final(kp[f, :cnt[f]], desc[f, :cnt[f]], ekp, edesc, f"batch frame {f}")
    # idempotence: the same batch again gives the same bytes
    kp2, desc2, cnt2 = gb.extract_batch(imgs)
    assert np.array_equal(cnt, cnt2)
    for f in range(B):
        assert kp[f, :cnt[f]].tobytes() == kp2[f, :cnt[f]].tobytes() and np.array_equal(desc[f, :cnt[f]], desc2[f, :cnt[f]])
    gb.close()


def test_octree_stage_against_oracle(gpu, oracle):
    g = gpu(1000, 640, 480)
    rs = np.random.RandomState(3)
    for it in range(40):
        width, height = int(rs.randint(40, 500)), int(rs.randint(40, 200))
        if round(width / height) < 1:
            continue
        M, N = int(rs.randint(0, 1500)), int(rs.randint(1, 300))
        pts = np.unique(np.stack([rs.randint(3, height - 3, M * 2), rs.randint(3, width - 3, M * 2)], 1), axis=0)
        pts = pts[rs.permutation(len(pts))][:M]
        cand = np.zeros(len(pts), ol.KP_DTYPE)
        cand["x"], cand["y"] = pts[:, 1], pts[:, 0]
        cand["response"] = rs.randint(7, 80, len(pts))
        exp = ol.octree(oracle, "orbo", cand, 16, 16 + width, 16, 16 + height, N)
        got = g.octree(cand, 16, 16 + width, 16, 16 + height, N)
        assert len(got) == len(exp)
        for f in ("x", "y", "response"):
            assert np.array_equal(got[f], exp[f]), f"iteration {it} field {f}"


def test_frame_on_demand_and_eager_agree(gpu, oracle):
    """The 19-px reflect-101 frame of the levels (ORBextractor.cc:1122-1128) is written on the first bordered read-back by default
    and inside the call with set_eager_frame(True): same key points, same descriptors, same bordered levels as the oracle either
    way, also for the frames of a batch and after interior-only reads."""
    w, h, nf, B = 752, 480, 1200, 5
    imgs = np.stack([synth.g_rects(w, h, 40 + i) for i in range(B)])
    g = gpu(nf, w, h, B)
    o = ol.Extractor(oracle, "orbo", nf, 1.2, 8, 20, 7)
    res = {}
    for eager in (False, True):
        g.set_eager_frame(eager)
        kp, desc, cnt = g.extract_batch(imgs)
        inner = g.level(3, frame=2)                      # an interior read must not disturb the pending frame
        res[eager] = (kp.copy(), desc.copy(), cnt.copy(), [g.level(l, frame=B - 1, bordered=True) for l in range(8)], inner)
    g.set_eager_frame(False)
    assert np.array_equal(res[False][2], res[True][2])
    assert res[False][0].tobytes() == res[True][0].tobytes() and np.array_equal(res[False][1], res[True][1])
    o.extract(imgs[B - 1])
    for l in range(8):
        assert np.array_equal(res[False][3][l], o.level(l, True)), f"level {l}, frame on demand"
        assert np.array_equal(res[True][3][l], o.level(l, True)), f"level {l}, eager frame"
    assert np.array_equal(res[False][4], res[True][4])


def test_octree_both_device_formulations(gpu, oracle):
    """The pass-free construction (og_octree2.cuh) answers ordinary candidate sets; a tight cluster divides deeper than its cell
    histogram and goes to the division-pass state machine (og_octree.cuh).  Both equal the oracle; careful-phase-heavy cases
    (N close to the number of occupied cells) included."""
    g = gpu(1000, 640, 480)
    rs = np.random.RandomState(17)
    paths = {0: 0, 1: 0}
    for it in range(60):
        width, height = int(rs.randint(200, 1250)), int(rs.randint(100, 400))
        if round(width / height) < 1:
            continue
        M, N = int(rs.randint(50, 6000)), int(rs.randint(5, 900))
        if it % 4 == 0:      # everything inside a few pixels: nodes divide down to single pixels
            cx, cy, s = int(rs.randint(20, width - 20)), int(rs.randint(20, height - 20)), int(rs.randint(3, 12))
            ys, xs = rs.randint(cy - s, cy + s, M), rs.randint(cx - s, cx + s, M)
        elif it % 4 == 1:    # a few clusters
            k = int(rs.randint(2, 6))
            c = np.stack([rs.randint(10, height - 10, k), rs.randint(10, width - 10, k)], 1)[rs.randint(0, k, M)]
            ys = np.clip(c[:, 0] + rs.normal(0, 6, M), 3, height - 4).astype(int)
            xs = np.clip(c[:, 1] + rs.normal(0, 6, M), 3, width - 4).astype(int)
        else:
            ys, xs = rs.randint(3, height - 3, M), rs.randint(3, width - 3, M)
        pts = np.unique(np.stack([ys, xs], 1), axis=0)
        pts = pts[rs.permutation(len(pts))]
        cand = np.zeros(len(pts), ol.KP_DTYPE)
        cand["x"], cand["y"] = pts[:, 1], pts[:, 0]
        cand["response"] = rs.randint(7, 255 if it % 2 else 40, len(pts))
        exp = ol.octree(oracle, "orbo", cand, 16, 16 + width, 16, 16 + height, N)
        got = g.octree(cand, 16, 16 + width, 16, 16 + height, N)
        paths[g.octree_last_path()] += 1
        assert len(got) == len(exp), it
        for f in ("x", "y", "response"):
            assert np.array_equal(got[f], exp[f]), f"iteration {it} field {f}"
    assert paths[0] >= 3 and paths[1] >= 20, paths


def test_full_size_batch_properties(oracle):
    """BASELINE size (KITTI 1241x376, 2000 features) at a full 512-frame batch: size-independent properties instead of
    512 oracle runs — every copy of a frame inside the batch gives byte-identical results wherever it sits (chunk and
    stream boundaries of the pipelined host path included), the device-resident and host-pointer paths agree, and a
    sample of frames equals the oracle."""
    import torch
    from orb_slam2_with_comment_b200 import ORBextractor
    w, h, nf, B, D = 1241, 376, 2000, 512, 8
    base = np.stack([synth.g_rects(w, h, 300 + i) for i in range(D)])
    imgs = np.ascontiguousarray(base[np.arange(B) % D])
    ex = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    kp, desc, cnt = ex.extract_batch(imgs)
    for f in range(D, B):
        r = f % D
        assert cnt[f] == cnt[r] and kp[f, :cnt[f]].tobytes() == kp[r, :cnt[r]].tobytes() and np.array_equal(desc[f, :cnt[f]], desc[r, :cnt[r]]), f
    dev = torch.device("cuda", 0)
    d_img = torch.from_numpy(imgs).to(dev)
    d_kp = torch.zeros(B * ex.kp_cap * 28, dtype=torch.uint8, device=dev)
    d_desc = torch.zeros(B * ex.kp_cap * 32, dtype=torch.uint8, device=dev)
    d_cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    ex.extract_batch_dev(d_img.data_ptr(), B, w, h, d_kp.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr())
    ex.sync()
    assert np.array_equal(d_cnt.cpu().numpy(), cnt)
    dk = d_kp.cpu().numpy().reshape(B, ex.kp_cap, 28)
    for f in (0, 77, 511):
        assert dk[f, :cnt[f]].tobytes() == kp[f, :cnt[f]].tobytes()
    o = ol.Extractor(oracle, "orbo", nf, 1.2, 8, 20, 7)
    for f in (3, 6):
        ekp, edesc = o.extract(base[f])
        compare_final(kp[f, :cnt[f]], desc[f, :cnt[f]], ekp, edesc, f"full batch frame {f}")
    ex.close()


@pytest.mark.parametrize("w,h,nf,nlevels,scale,ini,mn", [
    (1920, 1080, 5000, 8, 1.2, 20, 7),     # full-HD, larger quota: wider cell rows, more segments per row, more blur tiles
    (801, 601, 1500, 5, 1.5, 30, 10),      # odd sizes, fewer levels, other thresholds
    (1000, 700, 1200, 3, 2.5, 20, 7),      # scale > 2: the generic resize kernel (the vectorised one needs scale <= 2)
    (500, 700, 600, 4, 1.3, 12, 5),        # portrait: nIni = round(w/h) = 1 (taller than 2:1 is rejected, the reference divides by zero there)
])
def test_other_geometries_vs_oracle(oracle, w, h, nf, nlevels, scale, ini, mn):
    from orb_slam2_with_comment_b200 import ORBextractor
    g = ORBextractor(nf, scale, nlevels, ini, mn, max_width=w, max_height=h)
    o = ol.Extractor(oracle, "orbo", nf, scale, nlevels, ini, mn)
    for seed, gen in ((11, synth.g_rects), (12, synth.g_blurnoise)):
        img = gen(w, h, seed)
        kp, desc = g(img)
        ekp, edesc = o.extract(img)
        for l in range(nlevels):
            assert np.array_equal(g.level(l, bordered=True), o.level(l, True)), f"pyramid level {l}"
        compare_final(kp, desc, ekp, edesc, f"{w}x{h} seed {seed}")
        assert len(kp) > nf // 2
    g.close()


def test_colour_input_matches_cv2_golden_and_gray_path():
    """orbgpu_extract_batch_color: the gray image equals cv2's cvtColor (golden vectors), and the extraction equals the
    extraction of that gray image (Tracking::GrabImageMonocular, Tracking.cc:214-228)."""
    import os
    from orb_slam2_with_comment_b200 import ORBextractor
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cvtcolor_golden.npz"))
    ex = ORBextractor(300, 1.2, 4, 20, 7, max_width=203, max_height=120, max_batch=2)
    for name in ("c3", "c4"):
        for rgb in (True, False):
            gray, *_ = ex.extract_batch_color(np.ascontiguousarray(g[name]), rgb=rgb)
            assert np.array_equal(gray, g[name + ("_rgb" if rgb else "_bgr")]), (name, rgb)
    ex.close()
    # a real-sized colour frame: tint the synthetic gray scene, convert + extract in one call, compare with the gray path
    W, H = 640, 480
    base = synth.g_rects(W, H, 5).astype(np.int32)
    rs = np.random.RandomState(1)
    col = np.clip(np.stack([base + rs.randint(-20, 20, base.shape), base, base + rs.randint(-30, 30, base.shape)], -1), 0, 255).astype(np.uint8)
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=1)
    gray, kp, desc, cnt = ex.extract_batch_color(col[None].copy(), rgb=True)
    kp2, desc2 = ex(gray[0])
    assert cnt[0] == len(kp2) > 500 and kp[0, :cnt[0]].tobytes() == kp2.tobytes() and np.array_equal(desc[0, :cnt[0]], desc2)
    ex.close()


@pytest.mark.gpu
def test_register_staged_resize_path_in_a_fresh_process():
    """The pyramid resize has two vectorised kernels: k_resize_tma (source box by TMA; the default wherever a 192 x 32 output tile
    reads at most a 256 x 42 source box) and k_resize4_pp (source words in registers; the fall-back).  ORBGPU_RESIZE_TMA is read once
    per process, so the fall-back is exercised in a child process: pyramid stages and golden fixtures against the oracle."""
    import os, subprocess, sys
    env = dict(os.environ, ORBGPU_RESIZE_TMA="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-x", "-q", "-k",
                        "test_stages_match_oracle or test_golden_fixtures"], cwd=root, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
