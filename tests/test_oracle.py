"""Pins the CPU oracle (oracle/acsm_oracle.c): against the golden vectors generated from
the reference's own compiled builder (tests/golden/kat.json, made by make_golden.py), and --
where oracle/_ref/libacref.so exists -- live against that reference library, table for table
and match for match.  No GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

from gpu_pattern_matching_b200 import synth
from oracle_lib import (GOLDEN, Oracle, RefAcsm, RefIacsm, clamav_signatures, materialize,
                        parse_pattern_file, read_fixture, ref_available)

KAT = json.load(open(os.path.join(GOLDEN, "kat.json")))
FIXTURE_CASES = {
    "kat_pat_a.txt x kat_text_a.txt.gz": ("kat_pat_a.txt", "kat_text_a.txt.gz"),
    "kat_pat_b.txt x kat_text_b.txt.gz": ("kat_pat_b.txt", "kat_text_b.txt.gz"),
    "kat_pat_c.txt x kat_text_a.txt.gz": ("kat_pat_c.txt", "kat_text_a.txt.gz"),
    "kat_pat_two_words.txt x kat_pat_categorical_small.txt": ("kat_pat_two_words.txt",
                                                              "kat_pat_categorical_small.txt"),
    "sentiment_categorical.pat.gz x kat_text_a.txt.gz": ("sentiment_categorical.pat.gz",
                                                         "kat_text_a.txt.gz"),
    "sentiment_categorical.pat.gz x kat_text_readme.txt.gz": ("sentiment_categorical.pat.gz",
                                                              "kat_text_readme.txt.gz"),
}
HAND = [(b"abc", 0), (b"bc", 1), (b"c", 2), (b"abc", 3), (b"xbc", 4), (b"ab", 5)]


def _digest(off, pat):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(off, dtype=np.uint64).tobytes())
    h.update(np.ascontiguousarray(pat, dtype=np.uint32).tobytes())
    return h.hexdigest()


def _table_digest(t):
    a = t[:, :256]
    b = np.where(a < 0, t[:, 256:], 0)
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(a).tobytes())
    h.update(np.ascontiguousarray(b).tobytes())
    return h.hexdigest()


def _build(cls, pats):
    o = cls(256) if cls is Oracle else cls()
    for p, iid in pats:
        o.add(p, iid)
    o.compile()
    return o


def _inputs(name):
    if name in FIXTURE_CASES:
        pf, tf = FIXTURE_CASES[name]
        return parse_pattern_file(read_fixture(pf)), read_fixture(tf)
    if name.startswith("hand"):
        return HAND, b"zabcxbcab"
    n = int(name.split()[0][len("clamav"):])
    sigs = clamav_signatures(n)
    nbytes = 4 << 20
    buf = synth.stream(nbytes, 7)
    synth.Plants(sigs, nbytes, 512, 7, forced=[(0, 3), (nbytes - len(sigs[5]), 5)]).apply_host(buf)
    return [(s, i) for i, s in enumerate(sigs)], buf


@pytest.mark.parametrize("case", KAT["cases"], ids=[c["name"] for c in KAT["cases"]])
def test_oracle_matches_golden(case):
    pats, text = _inputs(case["name"])
    o = _build(Oracle, pats)
    off, pat, hits, fin = o.search(text)
    assert o.num_states == case["states"]
    assert o.max_pattern_len == case["max_pattern_len"]
    assert o.table_bytes == case["table_bytes"]
    assert hits == case["hits"] and off.size == case["matches"] and fin == case["final_state"]
    assert [[int(a), int(b)] for a, b in zip(off[:16], pat[:16])] == case["first"]
    assert _digest(off, pat) == case["digest"]
    b, _ = o.ml_csr()
    sizes = np.diff(b)
    assert int((sizes > 0).sum()) == case["final_states"]
    assert int((sizes > 1).sum()) == case["multi_pattern_finals"]
    assert int(sizes.max()) == case["max_list"]
    assert _table_digest(o.ref_table()) == case["table_digest"]
    # the bounded walk used as the timed CPU baseline counts the same matches
    assert o.walk_count(np.frombuffer(bytes(text), dtype=np.uint8)) == case["matches"]
    o.close()


def test_survey_known_answers():
    """SURVEY.md section 4 table (computed there with the reference's own builder)."""
    by = {c["name"]: c for c in KAT["cases"]}
    t = by["kat_pat_a.txt x kat_text_a.txt.gz"]
    assert (t["patterns"], t["states"], t["text_bytes"], t["hits"], t["matches"], t["first"][0]) == \
        (22, 198, 9479, 24, 24, [85, 0])
    t = by["kat_pat_b.txt x kat_text_b.txt.gz"]
    assert (t["patterns"], t["states"], t["hits"], t["matches"], t["first"][0]) == (25, 245, 25, 25, [236, 0])
    t = by["sentiment_categorical.pat.gz x kat_text_readme.txt.gz"]
    assert (t["patterns"], t["states"], t["hits"], t["matches"]) == (4376, 15704, 39, 40)
    t = by["hand x zabcxbcab"]
    assert t["first"] == [[2, 5], [3, 0], [3, 1], [3, 2], [3, 3], [6, 1], [6, 2], [6, 4], [8, 5]]
    for n, states, lmax in ((2000, 88783, 159), (10000, 378763, 187), (15000, 661298, 187)):
        c = by[f"clamav{n} x planted 4MiB seed 7"]
        assert (c["states"], c["max_pattern_len"]) == (states, lmax)
    assert KAT["ushort"]["matches"] == [[4, 1], [12, 2], [15, 0]] and KAT["ushort"]["states"] == 9


def test_ushort_oracle_matches_golden():
    o = Oracle(2048)
    for k, line in enumerate(read_fixture("ushort_signatures.txt").decode().splitlines()):
        o.add_csv(line.split(";")[0], k)
    o.compile()
    toks = np.array([9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 1, 2, 3, 4, 666, 676], dtype=np.uint16)
    off, pat, hits, fin = o.search(toks)
    got = [[int(a), o.pattern_iid(int(b))] for a, b in zip(off, pat)]
    assert got == KAT["ushort"]["matches"]
    assert o.num_states == KAT["ushort"]["states"] and fin == KAT["ushort"]["final_state"]


def test_halo_sharded_walk_is_exact():
    """cold start Lmax-1 bytes early + emit-from reproduces the serial walk."""
    sigs = clamav_signatures(2000)
    o = _build(Oracle, [(s, i) for i, s in enumerate(sigs)])
    n = 1 << 20
    buf = synth.stream(n, 21)
    synth.Plants(sigs, n, 200, 21, forced=[(n // 3 - 11, 7), (2 * n // 3 - 1, 9)]).apply_host(buf)
    eo, ep, _, _ = o.search(buf)
    halo = o.max_pattern_len - 1
    offs, pats = [], []
    for lo, hi in ((0, n // 3), (n // 3, 2 * n // 3), (2 * n // 3, n)):
        s = max(0, lo - halo)
        a, b, _, _ = o.search(buf[s:hi], emit_from=lo - s, base=s)
        offs.append(a)
        pats.append(b)
    assert np.array_equal(np.concatenate(offs), eo) and np.array_equal(np.concatenate(pats), ep)
    assert o.walk_count_mt(buf, 4) == eo.size
    o.close()


def test_c_pattern_parser_agrees_with_python_restatement(tmp_path):
    for name, hexp in (("kat_pat_a.txt", False), ("kat_pat_categorical_small.txt", False),
                       ("sentiment_categorical.pat.gz", False), ("clamav_sigs_15000.hex.gz", True)):
        path = materialize(name, tmp_path)
        o = Oracle(256)
        n = o.load_file(path, hexp)
        want = parse_pattern_file(read_fixture(name), hexp)
        assert n == len(want)
        for i in (0, 1, n // 2, n - 1):
            assert o.pattern(i) == want[i][0] and o.pattern_iid(i) == want[i][1]
        o.close()


needs_ref = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (no reference tree)")


@needs_ref
@pytest.mark.parametrize("name", list(FIXTURE_CASES) + ["hand x zabcxbcab",
                                                        "clamav2000 x planted 4MiB seed 7"])
def test_oracle_equals_compiled_reference(name):
    pats, text = _inputs(name)
    o, r = _build(Oracle, pats), _build(RefAcsm, pats)
    assert (o.num_states, o.max_pattern_len, o.table_bytes) == (r.num_states, r.max_pattern_len,
                                                               r.table_bytes)
    ht, ot = r.h_trans(), o.ref_table()
    neg = ht[:, :256] < 0
    assert np.array_equal(ht[:, :256], ot[:, :256])
    assert np.array_equal(ht[:, 256:][neg], ot[:, 256:][neg])     # head index of every final target
    (b1, i1), (b2, i2) = o.ml_csr(), r.ml_csr()
    assert np.array_equal(b1, b2) and np.array_equal(i1, i2)       # full lists, list order
    a, b = o.search(text), r.search(text)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2:] == b[2:]
    o.close()
    r.close()


@needs_ref
def test_ushort_oracle_equals_compiled_reference():
    o, r = Oracle(2048), RefIacsm()
    sigs = [[666, 676], [7, 6, 5], [1, 2, 3], [6, 5], [2047, 0, 2047]]
    for k, s in enumerate(sigs):
        o.add(s, 100 + k)
        r.add(s, 100 + k)
    o.compile()
    r.compile()
    assert o.num_states == r.num_states
    ht, ot = r.h_trans(), o.ref_table()
    neg = ht[:, :2048] < 0
    assert np.array_equal(ht[:, :2048], ot[:, :2048])
    assert np.array_equal(ht[:, 2048:][neg], ot[:, 2048:][neg])    # iid of the list head
    rng = np.random.default_rng(1)
    toks = rng.choice(np.array([1, 2, 3, 5, 6, 7, 666, 676, 2047, 0], dtype=np.uint16), size=4000)
    eo, ep, _, fin = o.search(toks)
    ro, ri, rfin = r.search(toks)
    assert np.array_equal(eo, ro) and fin == rfin
    assert [o.pattern_iid(int(p)) for p in ep] == ri.tolist()


@pytest.mark.parametrize("seed", range(6))
def test_oracle_equals_naive_search(seed):
    """The oracle's match list against the definition itself -- every (end offset, pattern index)
    such that the pattern occurs ending there, found by brute force -- on small random cases with
    overlapping, nested and duplicate patterns (bytes and ushort symbols)."""
    rng = np.random.default_rng(6000 + seed)
    ushort = seed % 2 == 1
    alpha = 2048 if ushort else 256
    symbols = [np.array([1, 2]), np.array([0, 7, 200]), np.arange(alpha)][seed % 3]
    dtype = np.uint16 if ushort else np.uint8
    for rep in range(8):
        npat = int(rng.integers(1, 30))
        pats = [rng.choice(symbols, size=int(rng.integers(1, 9))).astype(dtype) for _ in range(npat)]
        if npat >= 3:
            pats[1] = pats[0].copy()
            pats[2] = pats[0][-max(1, pats[0].size // 2):].copy()
        text = rng.choice(symbols, size=int(rng.choice([1, 5, 64, 700]))).astype(dtype)
        for _ in range(10):
            p = pats[int(rng.integers(0, npat))]
            if p.size <= text.size:
                pos = int(rng.integers(0, text.size - p.size + 1))
                text[pos:pos + p.size] = p
        o = Oracle(alpha)
        for i, p in enumerate(pats):
            o.add(p.tobytes() if not ushort else p, i)
        o.compile()
        eo, ep, _, _ = o.search(text)
        want = sorted((s + p.size - 1, i) for i, p in enumerate(pats)
                      for s in range(text.size - p.size + 1) if np.array_equal(text[s:s + p.size], p))
        assert list(zip(eo.tolist(), ep.tolist())) == want, f"seed {seed} rep {rep}"
        o.close()
