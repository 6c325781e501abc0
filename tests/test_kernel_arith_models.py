"""Host models of two arithmetic shortcuts of the CUDA kernels (no GPU needed): the properties the kernels rely on are
checked here exhaustively / on random words, so a later edit of the formulas has something to fail against.

1. k_fast_seg's SWAR rejection compare (csrc/og_extract.cu, `far2`): for t < 127 the byte-wise "|I - Ic| > t" runs on the
   UNMASKED differences, bit 7 of ((d + K) | d) with K = (127 - t) in every byte.  A carry from the byte below may set the bit
   for d == t, nothing else may differ from the exact compare: the filter stays conservative (it never loses a pixel that
   FAST's own test, ORBextractor.cc:809 -> cv::FAST, could accept; the exact score decides afterwards).
2. k_orient_desc's sample address (`brief_row_bits` / `brief_col_bits`): cvRound (ORBextractor.cc:115-120) is done by adding
   1.5 * 2^23 and reading the low mantissa bits; the address is formed on the raw bits,
   ir * 64 + (iq + k) with k = centre - 65 * 0x4B400000 (mod 2^32) == round(r) * 64 + round(q) + centre.
"""
import numpy as np
import pytest


def swar_far(d_words: np.ndarray, t: int) -> np.ndarray:
    """bit 7 of every byte of the kernel's expression, as 0/1 per byte [n, 4]."""
    k = np.uint64((127 - t) * 0x01010101)
    s = (d_words.astype(np.uint64) + k) & np.uint64(0xFFFFFFFF)
    g = (s | d_words.astype(np.uint64)).astype(np.uint32)
    return np.stack([(g >> np.uint32(8 * i + 7)) & np.uint32(1) for i in range(4)], axis=1).astype(np.uint8)


def bytes_of(words: np.ndarray) -> np.ndarray:
    return np.stack([(words >> np.uint32(8 * i)) & np.uint32(0xFF) for i in range(4)], axis=1).astype(np.int64)


@pytest.mark.parametrize("t", [0, 1, 6, 7, 19, 20, 21, 63, 64, 100, 125, 126])
def test_fast_rejection_compare_is_conservative(t):
    rng = np.random.RandomState(1000 + t)
    # random words plus words built from the values around the threshold and around the carry boundary 129 + t
    special = np.array([0, 1, t - 1, t, t + 1, t + 2, 127, 128, 129, 128 + t, 129 + t, 130 + t, 254, 255]).clip(0, 255)
    sp = special[rng.randint(0, len(special), (200000, 4))]
    words = np.concatenate([rng.randint(0, 2 ** 32, 200000, dtype=np.uint64).astype(np.uint32),
                            (sp[:, 0] | (sp[:, 1] << 8) | (sp[:, 2] << 16) | (sp[:, 3] << 24)).astype(np.uint32)])
    d = bytes_of(words)
    got = swar_far(words, t)
    exact = (d > t).astype(np.uint8)
    assert not np.any(exact & ~got & 1), "a far pixel was lost"
    extra = (got == 1) & (exact == 0)
    # the only admissible extra: d == t in a byte whose lower neighbour carried (d_below >= 129 + t, or a chain of carries)
    assert np.all(d[extra] == t)
    lanes = np.nonzero(extra)
    assert np.all(lanes[1] > 0), "byte 0 has no carry in"
    below = d[lanes[0], lanes[1] - 1]
    assert np.all(below + (127 - t) + 1 >= 256), "an extra candidate needs a carry out of the byte below"


def test_fast_rejection_compare_exhaustive_two_bytes():
    # every (lower byte, upper byte, t): the upper byte's result only depends on the two (carry chains need d == 255 - kadd below)
    lo, hi = np.meshgrid(np.arange(256, dtype=np.uint32), np.arange(256, dtype=np.uint32), indexing="ij")
    words = (lo | (hi << np.uint32(8))).ravel().astype(np.uint32)
    for t in range(0, 127):
        got = swar_far(words, t)
        d = bytes_of(words)
        exact = (d > t)
        assert not np.any(exact[:, :2] & (got[:, :2] == 0)), t
        extra = (got[:, 1] == 1) & ~exact[:, 1]
        assert np.all((d[extra, 1] == t) & (d[extra, 0] >= 129 + t)), t
        assert np.array_equal(got[:, 0] == 1, exact[:, 0]), t


def test_descriptor_sample_address_on_rounding_bits():
    rng = np.random.RandomState(7)
    n = 400000
    # rotated pattern coordinates: |v| <= 15 * sqrt(2) + rounding, including exact halves (ties to even, as cvRound / lrintf)
    v = np.concatenate([rng.uniform(-22, 22, n).astype(np.float32), (rng.randint(-44, 45, n) * 0.5).astype(np.float32)])
    w = np.concatenate([rng.uniform(-22, 22, n).astype(np.float32), (rng.randint(-44, 45, n) * 0.5).astype(np.float32)])[::-1]
    magic = np.float32(12582912.0)
    ir = (v + magic).astype(np.float32).view(np.uint32).astype(np.uint64)
    iq = (w + magic).astype(np.float32).view(np.uint32).astype(np.uint64)
    r, q = np.rint(v).astype(np.int64), np.rint(w).astype(np.int64)   # numpy rint = round half to even = cvRound on SSE2
    assert np.array_equal(ir.astype(np.int64) - 0x4B400000, r) and np.array_equal(iq.astype(np.int64) - 0x4B400000, q)
    for centre in (0x1234 + 18 * 64 + 18, 0xFFFF - 40 * 64, 18 * 64 + 18 + 15 + 2560):
        k = np.uint64((centre - 65 * 0x4B400000) % 2 ** 32)
        got = (ir * np.uint64(64) + ((iq + k) & np.uint64(0xFFFFFFFF))) & np.uint64(0xFFFFFFFF)
        exp = (r * 64 + q + centre) % 2 ** 32
        assert np.array_equal(got.astype(np.int64), exp)
