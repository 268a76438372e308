"""Randomised differential test: small random pattern sets (tiny alphabets -> many overlaps,
duplicates, patterns that are prefixes / suffixes / infixes of each other, lengths 1..40) over
random texts of ragged lengths, every kernel and both sampled strides, against the oracle."""
import numpy as np
import pytest

import gpu_pattern_matching_b200 as g
from helpers import assert_same, build_oracle, build_product, gpu_scan, modes_for, sample_stride

pytestmark = pytest.mark.gpu


def _case(rng, alphabet, npat, lo, hi, nbytes):
    pats = []
    for i in range(npat):
        L = int(rng.integers(lo, hi + 1))
        pats.append((bytes(rng.choice(alphabet, size=L).tolist()), int(rng.integers(-5, 1000))))
    # duplicates and nested patterns on purpose
    if npat >= 4:
        pats[1] = (pats[0][0], 7)
        pats[2] = (pats[0][0][: max(lo, len(pats[0][0]) // 2)], 8)
        pats[3] = (pats[0][0][-max(lo, len(pats[0][0]) // 2):], 9)
    text = rng.choice(alphabet, size=nbytes).astype(np.uint8)
    # sprinkle real occurrences, some overlapping, some at the very ends
    for k in range(min(40, nbytes // 8)):
        p = np.frombuffer(pats[int(rng.integers(0, npat))][0], dtype=np.uint8)
        if p.size <= nbytes:
            pos = int(rng.integers(0, nbytes - p.size + 1))
            text[pos:pos + p.size] = p
    if nbytes >= len(pats[0][0]):
        p = np.frombuffer(pats[0][0], dtype=np.uint8)
        text[:p.size] = p
        text[nbytes - p.size:] = p
    return pats, text


@pytest.mark.parametrize("seed", range(12))
def test_random_sets_all_kernels(device, seed):
    rng = np.random.default_rng(1000 + seed)
    alphabet = [np.array([97, 98], dtype=np.uint8), np.array([0, 1, 255], dtype=np.uint8),
                np.arange(256, dtype=np.uint8), np.array([0], dtype=np.uint8)][seed % 4]
    lo, hi = [(1, 6), (7, 12), (10, 40), (2, 17), (10, 11), (7, 9)][seed % 6]
    for rep in range(6):
        npat = int(rng.integers(1, 60))
        nbytes = int(rng.choice([1, 2, 7, 15, 16, 17, 63, 64, 65, 200, 1000, 4097, 20000]))
        pats, text = _case(rng, alphabet, npat, lo, hi, nbytes)
        o, a = build_oracle(pats), build_product(pats)
        eo, ep, _, _ = o.search(text)
        variants = [(a, m) for m in modes_for(a)]
        if a.get_min_pattern_size() >= 10:
            a4 = build_product(pats, stride=4)
            assert (sample_stride(a), sample_stride(a4)) == (8, 4)
            variants.append((a4, g.MODE_SAMPLED4))
        for aut, mode in variants:
            # tiny buckets on purpose: exercises overflow -> exact two-pass path as well
            for kw in ({}, {"bucket_shift": 8, "bucket_cap": 32}):
                off, pat, res = gpu_scan(device, aut, text, mode, **kw)
                assert_same(off, pat, eo, ep, f"seed {seed} rep {rep} mode {mode} {kw} n={nbytes} npat={npat}")


def test_long_patterns_and_all_byte_values(device):
    rng = np.random.default_rng(5)
    pats = [(bytes(rng.integers(0, 256, size=L).tolist()), i) for i, L in enumerate([4000, 4096, 300, 187, 10, 11])]
    pats.append((bytes(range(256)), 100))
    pats.append((bytes([0] * 64), 101))
    o, a = build_oracle(pats), build_product(pats)
    text = rng.integers(0, 256, size=1 << 16).astype(np.uint8)
    for pos, k in ((0, 1), (5000, 0), (12000, 6), (20001, 2), (30000, 7), ((1 << 16) - 187, 3)):
        p = np.frombuffer(pats[k][0], dtype=np.uint8)
        text[pos:pos + p.size] = p
    eo, ep, _, _ = o.search(text)
    assert eo.size >= 6
    for mode in modes_for(a):
        off, pat, _ = gpu_scan(device, a, text, mode)
        assert_same(off, pat, eo, ep, f"long patterns mode {mode}")
