"""Shared helpers for the parity tests: build the same automaton in the oracle and in
the product, run a scan through the C ABI, compare."""
import numpy as np

import gpu_pattern_matching_b200 as g
from gpu_pattern_matching_b200 import synth
from oracle_lib import Oracle, clamav_signatures, parse_pattern_file, read_fixture

KAT_CASES = [
    # (pattern fixture, hex, text fixture)
    ("kat_pat_a.txt", False, "kat_text_a.txt.gz"),
    ("kat_pat_b.txt", False, "kat_text_b.txt.gz"),
    ("kat_pat_c.txt", False, "kat_text_a.txt.gz"),
    ("kat_pat_two_words.txt", False, "kat_pat_categorical_small.txt"),
    ("sentiment_categorical.pat.gz", False, "kat_text_a.txt.gz"),
    ("sentiment_categorical.pat.gz", False, "kat_text_readme.txt.gz"),
]

HAND_PATTERNS = [b"abc", b"bc", b"c", b"abc", b"xbc", b"ab"]
HAND_TEXT = b"zabcxbcab"


def load_patterns(name, hex_pat=False):
    return parse_pattern_file(read_fixture(name), hex_pat)


def build_oracle(pats, alphabet=256):
    o = Oracle(alphabet)
    for p, iid in pats:
        o.add(p, iid)
    o.compile()
    return o


def build_product(pats, upload=True, stride=None):
    """stride=4 forces the 4-byte / stride-4 sampled filter even when every pattern is >= 10
    bytes (the builder then would pick 3-byte windows at stride 8)."""
    import os
    a = g.Acsm()
    for p, iid in pats:
        a.add_pattern(p, iid)
    old = os.environ.get("ACM_SAMPLE_STRIDE")
    if stride is not None:
        os.environ["ACM_SAMPLE_STRIDE"] = str(stride)
    try:
        a.compile()
    finally:
        if stride is not None:
            if old is None:
                del os.environ["ACM_SAMPLE_STRIDE"]
            else:
                os.environ["ACM_SAMPLE_STRIDE"] = old
    if upload:
        a.gen_state_table()
    return a


def sample_stride(acsm):
    return g.lib().acm_automaton_sample_stride(acsm.automaton)


def clamav_pats(n):
    return [(s, i) for i, s in enumerate(clamav_signatures(n))]


def planted_stream(pats, nbytes, seed, plants, forced=()):
    buf = synth.stream(nbytes, seed)
    pl = synth.Plants([p for p, _ in pats], nbytes, plants, seed, forced)
    pl.apply_host(buf)
    return buf, pl


def modes_for(acsm):
    m = [g.MODE_START2, g.MODE_DFA]
    if acsm.get_min_pattern_size() >= 7 or (acsm.automaton and g.lib().acm_automaton_split_len(acsm.automaton) > 0):
        m.insert(0, g.MODE_SAMPLED4)             # mixed sets: sampled + start-filter pass
    if acsm.automaton and g.lib().acm_automaton_cdfa_classes(acsm.automaton) > 0:
        m.append(g.MODE_CDFA)
    return m


def gpu_scan(device, acsm, data, mode=0, emit_lo=0, emit_hi=None, valid_lo=0, **kw):
    """Upload `data` (uint8 array), scan, fetch.  Returns (off, pat, result)."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    n = data.size
    d = device.alloc(n + 64)
    try:
        device.h2d(d, data)
        sc = g.Scanner(device, acsm.automaton, max(n, 1), mode=mode, **kw)
        res = sc.scan_device(d, n, emit_lo, emit_hi, valid_lo)
        off, pat = sc.fetch()
        sc.close()
    finally:
        device.free(d)
    return off, pat, res


def assert_same(got_off, got_pat, exp_off, exp_pat, what=""):
    assert got_off.size == exp_off.size, f"{what}: {got_off.size} matches, oracle has {exp_off.size}"
    assert np.array_equal(got_off, exp_off), f"{what}: offsets differ"
    assert np.array_equal(got_pat, exp_pat), f"{what}: pattern indices differ"
