"""Host side of the product (no GPU): the C builder behind acsmx.h / iacsmx.h reproduces the
reference's automaton bit for bit (via acsm_export_ref_table), the pattern-file loader follows
the reference grammar, error paths return codes instead of exiting."""
import ctypes as C

import numpy as np
import pytest

import gpu_pattern_matching_b200 as g
from gpu_pattern_matching_b200 import _lib
from helpers import HAND_PATTERNS, build_oracle, build_product, clamav_pats, load_patterns
from oracle_lib import (Oracle, RefAcsm, materialize, parse_pattern_file, read_fixture,
                        ref_available)


def _same_table(at, ot, alpha=256):
    neg = ot[:, :alpha] < 0
    return (at.shape == ot.shape and np.array_equal(ot[:, :alpha], at[:, :alpha]) and
            np.array_equal(ot[:, alpha:][neg], at[:, alpha:][neg]))


@pytest.mark.parametrize("name", ["hand", "kat_pat_a.txt", "kat_pat_b.txt", "kat_pat_c.txt",
                                  "kat_pat_two_words.txt", "kat_pat_categorical_small.txt",
                                  "sentiment_categorical.pat.gz", "clamav2000"])
def test_exported_table_equals_oracle(name):
    if name == "hand":
        pats = [(p, i) for i, p in enumerate(HAND_PATTERNS)]
    elif name.startswith("clamav"):
        pats = clamav_pats(int(name[6:]))
    else:
        pats = load_patterns(name)
    o, a = build_oracle(pats), build_product(pats, upload=False)
    assert a.get_states() == o.num_states - 1       # highest id before gen_state_table (acsmx.c:615)
    assert a.get_max_pattern_size() == o.max_pattern_len
    assert _same_table(a.export_ref_table(), o.ref_table())
    a.free()
    o.close()


@pytest.mark.slow
def test_exported_table_equals_oracle_15000():
    pats = clamav_pats(15000)
    o, a = build_oracle(pats), build_product(pats, upload=False)
    assert o.num_states == 661298
    assert _same_table(a.export_ref_table(), o.ref_table())


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built")
def test_exported_table_equals_compiled_reference():
    pats = clamav_pats(2000)
    r = RefAcsm()
    for p, iid in pats:
        r.add(p, iid)
    r.compile()
    a = build_product(pats, upload=False)
    assert _same_table(a.export_ref_table(), r.h_trans())
    r.close()


def test_ushort_builder_table():
    sigs = [[666, 676], [7, 6, 5], [1, 2, 3], [6, 5], [2047, 0, 2047]]
    o = Oracle(2048)
    m = g.Iacsm()
    for k, s in enumerate(sigs):
        o.add(s, 100 + k)
        m.add_pattern(s, 100 + k)
    m.add_fullpattern("40,32,287", 7)
    o.add_csv("40,32,287", 7)
    o.compile()
    m.compile()
    assert m.get_states() == o.num_states - 1 and m.get_max_pattern_size() == 3
    assert _same_table(m.export_ref_table(), o.ref_table(), 2048)
    with pytest.raises(g.AcmError):
        g.Iacsm().add_pattern([1, 2048], 0)          # symbol outside the alphabet


def test_patterns_table_and_indices():
    a = g.Acsm()
    a.add_pattern(b"one", 11)
    a.add_pattern(b"tw\x00o", -5)                     # embedded NUL survives (reference truncates)
    a.add_pattern(b"one", 12)
    a.compile()
    tab = a.get_patterns_table()
    assert tab == [(b"one", 11, 0), (b"tw\x00o", -5, 1), (b"one", 12, 2)]
    assert a.num_patterns == 3 and a.get_min_pattern_size() == 3 and a.get_max_pattern_size() == 4


def test_pattern_file_loader_follows_reference_grammar(tmp_path):
    for name, hexp, limit in (("kat_pat_a.txt", False, -1), ("kat_pat_categorical_small.txt", False, -1),
                              ("sentiment_categorical.pat.gz", False, -1),
                              ("sentiment_categorical.pat.gz", False, 3),
                              ("clamav_sigs_15000.hex.gz", True, -1), ("clamav_sigs_15000.hex.gz", True, 12)):
        path = materialize(name, tmp_path)
        a = g.Acsm()
        n = a.load_pattern_file(path, hexp, limit)
        want = parse_pattern_file(read_fixture(name), hexp, limit)
        assert n == len(want)
        a.compile()
        got = a.get_patterns_table()
        assert [(p, i) for p, i, _ in got] == want
        a.free()


def test_pattern_file_errors(tmp_path):
    a = g.Acsm()
    with pytest.raises(g.AcmError):
        a.load_pattern_file(tmp_path / "missing.txt")
    bad = tmp_path / "odd.hex"
    bad.write_text("4d5a9\n")
    with pytest.raises(g.AcmError):
        a.load_pattern_file(bad, hex_pat=True)        # the reference exit()s here (utils.c:39-42)
    quoted = tmp_path / "q.txt"
    quoted.write_text('7 "a b"\n-3 " spaced "\n+12 plain\n')
    b = g.Acsm()
    assert b.load_pattern_file(quoted) == 3
    b.compile()
    assert [(p, i) for p, i, _ in b.get_patterns_table()] == [(b"a b", 7), (b" spaced ", -3), (b"plain", 12)]


def test_error_codes_instead_of_exit():
    L = g.lib()
    a = g.Acsm()
    a.add_pattern(b"", 1)                             # zero length: kept, flagged, never matches
    assert a.status() == -15                          # ACM_ERR_EMPTY_PATTERN
    a.add_pattern(b"ok", 2)
    a.compile()
    assert a.num_patterns == 2 and a.get_states() == 2
    if L.acm_device_count() == 0:
        with pytest.raises(g.AcmError):
            a.gen_state_table()                       # no GPU here: must fail loudly, not fall back
    h = L.printable_hex_to_bytes(b"4D5a90")
    assert h and bytes((C.c_ubyte * 3).from_address(h)) == b"\x4d\x5a\x90"
    assert not L.printable_hex_to_bytes(b"4d5")
    assert not L.printable_hex_to_bytes(b"zz")


@pytest.mark.parametrize("nsig,stride", [(2000, 8), (2000, 4), (10000, 8), (15000, 8)])
def test_filter_tables_consistent_with_patterns(nsig, stride):
    """For every pattern and every alignment exactly one window is indexed: its key is in both
    bitmaps and in the exact table, whose candidate list (count field, LAST flag) holds the
    pattern with its bytes at the window, length, tail and blob offset (acm_core_check_filters)."""
    a = build_product(clamav_pats(nsig), upload=False, stride=stride)
    assert a.check_filters() == 0


def test_filter_check_reports_no_filter_for_short_patterns():
    a = build_product(load_patterns("sentiment_categorical.pat.gz"), upload=False)
    assert a.check_filters() == -1          # patterns shorter than 7 bytes: no sampled filter


def test_filter_tables_of_mixed_sets_leave_out_the_short_patterns(monkeypatch):
    """A few patterns under 7 bytes among many long ones: the sampled filter is still built, over
    the long patterns only (stride 8 when at most 1/8 of the set is under 10 bytes, else stride
    4 with the split at 7); the short ones must be in the short-start bitmap."""
    short = [(b"a", 9001), (b"MZ", 9002), (b"\x00\x01\x02", 9003), (b"virus!", 9004)]
    a = build_product(clamav_pats(2000) + short, upload=False)
    assert a.get_min_pattern_size() == 1 and a.check_filters() == 0
    mid = [(bytes([200 + i, 7, 7, 7, 7, 7, 7, i]), 9100 + i) for i in range(40)]   # 8 bytes: under 10
    a = build_product(clamav_pats(2000)[:200] + short + mid, upload=False)          # 44 of 244 under 10
    assert a.check_filters() == 0
    monkeypatch.setenv("ACM_HYBRID", "0")
    a = build_product(clamav_pats(2000) + short, upload=False)
    assert a.check_filters() == -1          # switched off: no sampled filter for this set


def test_filter_tables_cover_every_pattern():
    """Every pattern's four leading 4-grams must be present in both bitmaps: checked through the
    exported reference table indirectly by test_gpu_parity; here: builder statistics."""
    pats = clamav_pats(2000)
    a = build_product(pats, upload=False)
    assert a.get_min_pattern_size() == 10 and a.get_max_pattern_size() == 159


def _random_pats(rng, alphabet, npat, lo, hi):
    pats = [(bytes(rng.choice(alphabet, size=int(rng.integers(lo, hi + 1))).tolist()), int(rng.integers(-9, 999)))
            for _ in range(npat)]
    if npat >= 4:                                    # duplicates, prefixes, suffixes, infixes on purpose
        p0 = pats[0][0]
        pats[1] = (p0, 7)
        pats[2] = (p0[:max(1, len(p0) // 2)], 8)
        pats[3] = (p0[-max(1, len(p0) // 2):], 9)
    return pats


@pytest.mark.parametrize("seed", range(8))
def test_random_sets_builder_equals_oracle_and_reference(seed):
    """Randomised differential test of the builder (no GPU): tiny alphabets make patterns overlap,
    nest and repeat; the exported table -- reference layout, reference numbering, head index of every
    final state -- must equal the oracle's and, where oracle/_ref is built, the table the
    reference's own acsm_compile / acsm_gen_state_table produce."""
    rng = np.random.default_rng(4000 + seed)
    alphabet = [np.array([97, 98], dtype=np.uint8), np.array([0, 1, 255], dtype=np.uint8),
                np.arange(256, dtype=np.uint8), np.array([0], dtype=np.uint8)][seed % 4]
    for rep in range(5):
        pats = _random_pats(rng, alphabet, int(rng.integers(1, 80)), 1, [6, 12, 40, 3][rep % 4])
        o, a = build_oracle(pats), build_product(pats, upload=False)
        assert a.get_states() == o.num_states - 1 and a.get_max_pattern_size() == o.max_pattern_len
        assert _same_table(a.export_ref_table(), o.ref_table()), f"seed {seed} rep {rep}: builder != oracle"
        if ref_available():
            r = RefAcsm()
            for p, iid in pats:
                r.add(p, iid)
            r.compile()
            assert _same_table(a.export_ref_table(), r.h_trans()), f"seed {seed} rep {rep}: builder != reference"
            r.close()
        a.free()
        o.close()


@pytest.mark.parametrize("seed", range(4))
def test_random_ushort_sets_builder_equals_oracle(seed):
    rng = np.random.default_rng(5000 + seed)
    tokens = [np.array([0, 1], dtype=np.uint16), np.array([0, 40, 52, 1448, 2047], dtype=np.uint16),
              np.arange(2048, dtype=np.uint16)][seed % 3]
    for rep in range(3):
        o, m = Oracle(2048), g.Iacsm()
        sigs = [rng.choice(tokens, size=int(rng.integers(1, 20))).astype(np.uint16) for _ in range(int(rng.integers(1, 40)))]
        if len(sigs) >= 2:
            sigs[1] = sigs[0].copy()
        for k, s in enumerate(sigs):
            o.add(s, 100 + k)
            m.add_pattern(s, 100 + k)
        o.compile()
        m.compile()
        assert m.get_states() == o.num_states - 1
        assert _same_table(m.export_ref_table(), o.ref_table(), 2048), f"seed {seed} rep {rep}"


def test_row_displaced_tables_are_the_automaton():
    """The two compact forms the DFA kernels walk -- xd (every automaton: k_scan_xd) and rd (word
    lists: k_scan_rd) -- against the dense table, over every (state, symbol), on the host."""
    import ctypes as C
    from helpers import build_product, clamav_pats, load_patterns
    import gpu_pattern_matching_b200 as g
    for name, hexp in (("kat_pat_a.txt", False), ("kat_pat_b.txt", False), ("kat_pat_c.txt", False),
                       ("kat_pat_two_words.txt", False), ("sentiment_categorical.pat.gz", False)):
        a = build_product(load_patterns(name, hexp), upload=False)
        slots = C.c_uint(0)
        assert a.L.acsm_check_xd(a._p, C.byref(slots)) == 0 and slots.value >= 256, name
        assert a.L.acsm_check_cdfa(a._p, None, None) in (0, -1), name        # -1: not a one-range word list
    a = build_product(load_patterns("sentiment_categorical.pat.gz"), upload=False)
    slots, dense = C.c_uint(0), C.c_uint(0)
    assert a.L.acsm_check_cdfa(a._p, C.byref(slots), C.byref(dense)) == 0
    assert slots.value * 4 < 176 * 1024 and 0 < dense.value <= 256          # fits shared memory
    a = build_product(clamav_pats(2000), upload=False)
    slots = C.c_uint(0)
    assert a.L.acsm_check_xd(a._p, C.byref(slots)) == 0
    assert slots.value * 4 < 4 << 20                                        # ~0.8 MB where the dense table is 87 MiB
    # ushort symbols (AC_ushorts): random sequences over a skewed alphabet
    rng = np.random.default_rng(6)
    w = 1.0 / np.arange(1, 2049) ** 1.1
    w /= w.sum()
    m = g.Iacsm()
    for i in range(300):
        m.add_pattern(rng.choice(2048, size=int(rng.integers(2, 17)), p=w).astype(np.uint16), i)
    m.compile()
    assert m.L.iacsm_check_xd(m._p, C.byref(slots)) == 0 and slots.value >= 2048
