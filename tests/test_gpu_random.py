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


def _ushort_case(rng, npat, lo, hi, ntok, tokens):
    pats = [(rng.choice(tokens, size=int(rng.integers(lo, hi + 1))).astype(np.uint16), int(rng.integers(0, 5000)))
            for _ in range(npat)]
    if npat >= 3:
        pats[1] = (pats[0][0].copy(), 7)                        # duplicate
        pats[2] = (pats[0][0][-max(1, pats[0][0].size // 2):].copy(), 8)   # proper suffix
    text = rng.choice(np.concatenate([tokens, [2048, 0xFFFF]]).astype(np.uint16), size=ntok).astype(np.uint16)
    for _ in range(min(60, ntok // 4)):
        p = pats[int(rng.integers(0, npat))][0]
        if p.size <= ntok:
            pos = int(rng.integers(0, ntok - p.size + 1))
            text[pos:pos + p.size] = p
    p = pats[0][0]
    if ntok >= p.size:
        text[:p.size] = p
        text[ntok - p.size:] = p
    return pats, text


@pytest.mark.parametrize("seed", range(6))
def test_random_mixed_sets_forced_hybrid(device, seed, monkeypatch):
    """ACM_HYBRID=1: every set that has both a pattern under the split and one at or above it is
    scanned as sampled filter (long) + start filter (short), whatever the proportions."""
    monkeypatch.setenv("ACM_HYBRID", "1")
    rng = np.random.default_rng(3000 + seed)
    alphabet = [np.array([97, 98], dtype=np.uint8), np.array([0, 1, 255], dtype=np.uint8),
                np.arange(256, dtype=np.uint8)][seed % 3]
    hybrids = 0
    for rep in range(6):
        npat = int(rng.integers(2, 60))
        nbytes = int(rng.choice([1, 9, 16, 17, 64, 200, 1000, 4097, 20000]))
        pats, text = _case(rng, alphabet, npat, 1, 24, nbytes)
        o, a = build_oracle(pats), build_product(pats)
        eo, ep, _, _ = o.search(text)
        split = g.lib().acm_automaton_split_len(a.automaton)
        hybrids += split > 0
        variants = [(a, m) for m in [0] + modes_for(a)]
        if split == 10:
            a4 = build_product(pats, stride=4)
            if g.lib().acm_automaton_split_len(a4.automaton) == 7:
                variants.append((a4, g.MODE_SAMPLED4))
        for aut, mode in variants:
            for kw in ({}, {"bucket_shift": 8, "bucket_cap": 32}):
                off, pat, res = gpu_scan(device, aut, text, mode, **kw)
                assert_same(off, pat, eo, ep, f"hybrid seed {seed} rep {rep} mode {mode} {kw} n={nbytes} npat={npat}")
    assert hybrids >= 3


@pytest.mark.parametrize("seed", range(6))
def test_random_ushort_sets(device, seed):
    """AC_ushorts path (iacsm_*, alphabet 2048): random symbol-sequence signatures over random
    packet-size trains with out-of-alphabet separators, ragged lengths, every chunk size of the
    walk kernel, chunk cuts on and off the 16-byte grid -- against the oracle."""
    from oracle_lib import Oracle
    rng = np.random.default_rng(2000 + seed)
    tokens = [np.array([0, 1], dtype=np.uint16), np.array([0, 40, 52, 1448, 2047], dtype=np.uint16),
              np.arange(2048, dtype=np.uint16)][seed % 3]
    lo, hi = [(1, 5), (3, 30), (8, 64)][seed % 3]
    for rep in range(4):
        npat = int(rng.integers(1, 80))
        ntok = int(rng.choice([1, 2, 7, 8, 9, 127, 128, 129, 1000, 4097, 30000]))
        pats, text = _ushort_case(rng, npat, lo, hi, ntok, tokens)
        m, o = g.Iacsm(), Oracle(2048)
        for p, iid in pats:
            m.add_pattern(p, iid)
            o.add(p, iid)
        m.compile()
        o.compile()
        m.gen_state_table(0, device.handle, None)
        eo, ep, _, _ = o.search(text)
        d = device.alloc(text.nbytes + 128)
        try:
            device.h2d(d, text)
            for emit_lo in (0, 1, 5):                            # chunk cuts off the 16-byte grid
                if emit_lo >= ntok and emit_lo:
                    continue
                keep = eo >= emit_lo
                for chunk in (0, 16, 48, 1000):
                    sc = g.Scanner(device, m.automaton, max(ntok, 1), dfa_chunk=chunk)
                    res = sc.scan_device(d, ntok, emit_lo)
                    off, pat = sc.fetch()
                    sc.close()
                    assert res.mode == g.MODE_DFA
                    assert_same(off, pat, eo[keep], ep[keep], f"ushort seed {seed} rep {rep} n={ntok} "
                                                              f"npat={npat} emit_lo={emit_lo} chunk={chunk}")
        finally:
            device.free(d)
