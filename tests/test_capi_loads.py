"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/orbgpu.h declares; without a GPU every compute entry fails loudly (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    out = []
    for hdr in ("orbgpu.h",):
        txt = open(os.path.join(ROOT, "include", hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        out += re.findall(r"\b(orbgpu_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(out))


def test_library_builds_and_exports_every_declared_symbol():
    from orb_slam2_with_comment_b200 import build, capi
    build.build()
    L = capi.lib()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/orbgpu.h but not exported by liborbgpu.so"


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from orb_slam2_with_comment_b200 import ORBextractor
    from orb_slam2_with_comment_b200.capi import OrbGpuError
    with pytest.raises(OrbGpuError, match="no CUDA device"):
        ORBextractor(2000, 1.2, 8, 20, 7)
    from orb_slam2_with_comment_b200.matcher import ORBmatcher
    with pytest.raises(OrbGpuError, match="no CUDA device"):
        ORBmatcher(0.6, True)


def test_keypoint_layout_is_cv_keypoint():
    from orb_slam2_with_comment_b200 import KP_DTYPE
    assert KP_DTYPE.itemsize == 28 and KP_DTYPE.names == ("x", "y", "size", "angle", "response", "octave", "class_id")


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: no product source (package, C ABI, shells) may import, include or link it."""
    pkg = os.path.join(ROOT, "orb_slam2_with_comment_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp", "Makefile")):
                continue
            text = open(os.path.join(base, f), errors="ignore").read()
            for needle in ("oracle_lib", "liborboracle", "liborbref", "libdbowref", "/oracle/", "import oracle", "from oracle"):
                if needle in text:
                    offenders.append((os.path.relpath(os.path.join(base, f), ROOT), needle))
    assert not offenders, offenders


def test_new_entries_fail_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    from orb_slam2_with_comment_b200 import capi, synth
    from orb_slam2_with_comment_b200.vocabulary import ORBVocabulary
    with pytest.raises(capi.OrbGpuError, match="no CUDA device"):
        ORBVocabulary().from_records(synth.vocabulary_tree(k=3, L=2, seed=1))
