"""GPU parity: every scan kernel, through the C ABI, against the CPU oracle on the same
bytes.  Bar: identical sorted (end offset, pattern index) lists -- bit exact."""
import numpy as np
import pytest

import gpu_pattern_matching_b200 as g
from gpu_pattern_matching_b200 import synth
from helpers import (HAND_PATTERNS, HAND_TEXT, KAT_CASES, assert_same, build_oracle, build_product,
                     clamav_pats, gpu_scan, load_patterns, modes_for, planted_stream, sample_stride)
from oracle_lib import read_fixture

pytestmark = pytest.mark.gpu


def test_hand_case(device):
    pats = [(p, i) for i, p in enumerate(HAND_PATTERNS)]
    o, a = build_oracle(pats), build_product(pats)
    eo, ep, _, _ = o.search(HAND_TEXT)
    assert list(zip(eo.tolist(), ep.tolist())) == [(2, 5), (3, 0), (3, 1), (3, 2), (3, 3), (6, 1),
                                                  (6, 2), (6, 4), (8, 5)]
    for mode in modes_for(a):
        off, pat, res = gpu_scan(device, a, np.frombuffer(HAND_TEXT, dtype=np.uint8), mode)
        assert_same(off, pat, eo, ep, f"hand mode {mode}")


@pytest.mark.parametrize("pf,hexp,tf", KAT_CASES)
def test_reference_fixtures(device, pf, hexp, tf):
    pats = load_patterns(pf, hexp)
    text = np.frombuffer(read_fixture(tf), dtype=np.uint8)
    o, a = build_oracle(pats), build_product(pats)
    eo, ep, _, _ = o.search(text)
    for mode in modes_for(a):
        off, pat, res = gpu_scan(device, a, text, mode)
        assert res.mode == mode
        assert_same(off, pat, eo, ep, f"{pf} x {tf} mode {mode}")


@pytest.mark.parametrize("nsig,nbytes,plants", [(2000, 4 << 20, 512), (10000, 2 << 20, 256)])
def test_clamav_planted(device, nsig, nbytes, plants):
    pats = clamav_pats(nsig)
    o, a = build_oracle(pats), build_product(pats)
    assert a.get_min_pattern_size() >= 7
    lmax = a.get_max_pattern_size()
    forced = [(0, 3), (nbytes - len(pats[5][0]), 5),            # first and last byte
              ((1 << 15) - 7, 11), ((1 << 16) - 1, 12),          # across result buckets
              ((1 << 20) - 3, 13)]
    buf, pl = planted_stream(pats, nbytes, seed=7, plants=plants, forced=forced)
    eo, ep, _, _ = o.search(buf)
    assert eo.size >= plants        # every plant is found (plus chance overlaps)
    for mode in modes_for(a):
        off, pat, res = gpu_scan(device, a, buf, mode)
        assert_same(off, pat, eo, ep, f"clamav{nsig} mode {mode}")
        assert not res.fallback
    # both sampled variants: 3-byte windows at stride 8 (default, min length 10) and 4-byte at stride 4
    a4 = build_product(pats, stride=4)
    assert (sample_stride(a), sample_stride(a4)) == (8, 4)
    off, pat, res = gpu_scan(device, a4, buf, g.MODE_SAMPLED4)
    assert_same(off, pat, eo, ep, f"clamav{nsig} sampled stride 4")


def test_unaligned_lengths_and_tails(device):
    pats = clamav_pats(2000)
    o, a = build_oracle(pats), build_product(pats)
    a4 = build_product(pats, stride=4)
    for n in (1, 3, 15, 16, 17, 31, 33, 64, 4095, 65537):
        buf = synth.stream(max(n, 256), 11)[:n].copy()
        # a signature that ends exactly at the last byte whenever it fits
        sig = pats[17][0]
        if n >= len(sig):
            buf[n - len(sig):] = np.frombuffer(sig, dtype=np.uint8)
        eo, ep, _, _ = o.search(buf)
        for mode in modes_for(a):
            off, pat, _ = gpu_scan(device, a, buf, mode)
            assert_same(off, pat, eo, ep, f"n={n} mode {mode}")
        off, pat, _ = gpu_scan(device, a4, buf, g.MODE_SAMPLED4)
        assert_same(off, pat, eo, ep, f"n={n} sampled stride 4")


def test_emit_window_sharding(device):
    """Contiguous shards with Lmax-1 bytes of leading halo reproduce the whole-stream list."""
    pats = clamav_pats(2000)
    o, a = build_oracle(pats), build_product(pats)
    n = 1 << 20
    cuts = [0, 300001, 300002 + 4096, 777777, n]
    forced = [(c - 20, 40 + k) for k, c in enumerate(cuts[1:-1])]     # straddle every cut
    buf, _ = planted_stream(pats, n, seed=3, plants=128, forced=forced)
    eo, ep, _, _ = o.search(buf)
    halo = a.get_max_pattern_size() - 1
    a4 = build_product(pats, stride=4)
    for aut, mode in [(a, m) for m in modes_for(a)] + [(a4, g.MODE_SAMPLED4)]:
        offs, pts = [], []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            start = max(0, lo - halo) & ~15       # device buffers are 16-byte aligned
            off, pat, _ = gpu_scan(device, aut, buf[start:hi], mode, emit_lo=lo - start, emit_hi=hi - start)
            offs.append(off + np.uint64(start))
            pts.append(pat)
        assert_same(np.concatenate(offs), np.concatenate(pts), eo, ep, f"sharded mode {mode}")


def test_valid_lo_hides_stale_prefix(device):
    pats = [(b"needle", 0), (b"dle", 1)]
    o, a = build_oracle(pats), build_product(pats)
    body = b"xxneedlexx needle"
    buf = np.frombuffer(b"nee" + b"dle....." + body, dtype=np.uint8)   # "needle" straddles valid_lo
    valid = 3
    eo, ep, _, _ = o.search(buf[valid:])
    for mode in modes_for(a):
        off, pat, _ = gpu_scan(device, a, buf, mode, emit_lo=valid, valid_lo=valid)
        assert_same(off - np.uint64(valid), pat, eo, ep, f"valid_lo mode {mode}")


def test_dense_matches_overflow_fallback(device):
    """Every byte matches several patterns: buckets overflow, the exact two-pass path runs."""
    pats = [(b"a", 0), (b"aa", 1), (b"aaa", 2), (b"ab", 3), (b"b", 4)]
    o, a = build_oracle(pats), build_product(pats)
    rng = np.random.default_rng(5)
    buf = rng.choice(np.frombuffer(b"aab", dtype=np.uint8), size=200000)
    eo, ep, _, _ = o.search(buf)
    for mode in modes_for(a):
        off, pat, res = gpu_scan(device, a, buf, mode, bucket_cap=32, bucket_shift=12)
        assert res.fallback == 1
        assert_same(off, pat, eo, ep, f"dense mode {mode}")
        off, pat, res = gpu_scan(device, a, buf, mode, bucket_cap=8192, bucket_shift=10)
        assert res.fallback == 0
        assert_same(off, pat, eo, ep, f"dense mode {mode} (buckets)")


def test_sentiment_text(device):
    pats = load_patterns("sentiment_categorical.pat.gz")
    o, a = build_oracle(pats), build_product(pats)
    words = [l.split(b"\t")[0] for l in read_fixture("english_top5000.txt.gz").split(b"\n") if l]
    text = synth.english_like(words, 1 << 20, seed=4)
    eo, ep, _, _ = o.search(text)
    assert eo.size > 5000
    for mode in modes_for(a):
        off, pat, res = gpu_scan(device, a, text, mode)
        assert_same(off, pat, eo, ep, f"sentiment mode {mode}")


def test_scan_host_pipeline(device):
    pats = clamav_pats(2000)
    o, a = build_oracle(pats), build_product(pats)
    n = (3 << 20) + 12345
    seg = 1 << 20
    forced = [(seg - 9, 21), (2 * seg - 30, 22), (3 * seg - 1, 23)]
    buf, _ = planted_stream(pats, n, seed=9, plants=300, forced=forced)
    eo, ep, _, _ = o.search(buf)
    sc = g.Scanner(device, a.automaton, seg)        # forces 4 segments
    off, pat, res = sc.scan_host(buf, base=1000)
    assert_same(off - np.uint64(1000), pat, eo, ep, "scan_host")
    sc.close()


def test_histogram(device):
    pats = clamav_pats(2000)
    o, a = build_oracle(pats), build_product(pats)
    buf, _ = planted_stream(pats, 1 << 20, seed=13, plants=400)
    eo, ep, _, _ = o.search(buf)
    d = device.alloc(buf.size + 64)
    device.h2d(d, buf)
    sc = g.Scanner(device, a.automaton, buf.size)
    sc.scan_device(d, buf.size)
    dc = device.alloc(8 * len(pats))
    device.h2d(dc, np.zeros(len(pats), dtype=np.uint64))
    sc.histogram_into(dc)
    counts = device.d2h(dc, 8 * len(pats), np.uint64)
    assert np.array_equal(counts, np.bincount(ep, minlength=len(pats)).astype(np.uint64))
    device.free(d)
    device.free(dc)
    sc.close()


def test_device_synth_matches_host(device):
    n = 1 << 16
    d = device.alloc(n)
    device.synth_fill(d, n, seed=99, offset=4096)
    device.sync()
    got = device.d2h(d, n)
    assert np.array_equal(got, synth.stream(n, 99, 4096))
    host = np.empty(1000, dtype=np.uint8)
    import ctypes as C
    g.lib().acm_synth_fill_host(host.ctypes.data_as(C.c_void_p), 1000, 99, 4099)
    assert np.array_equal(host, synth.stream(1000, 99, 4099))
    device.free(d)


def test_repetitive_input_dense_hit_fallback(device):
    """Zero pages and short-period fills make almost every aligned window a pattern gram; the
    sampled kernel hands such chunks to the automaton walk.  Results must not change."""
    pats = clamav_pats(10000)
    o, a = build_oracle(pats), build_product(pats)
    n = 1 << 19
    buf = synth.stream(n, 31)
    sig = lambda i: np.frombuffer(pats[i][0], dtype=np.uint8)
    buf[4096:4096 + 40000] = 0                                     # zero run
    buf[100000:100000 + 30000] = np.tile(np.frombuffer(bytes.fromhex("e800005d"), np.uint8), 7500)
    buf[200000:200000 + 16384] = np.tile(np.frombuffer(bytes.fromhex("b440cd21e8c2045a"), np.uint8), 2048)
    # real signatures inside, next to and straddling the dense regions
    for pos, i in ((5000, 504), (44090, 4385), (4096 - 30, 17), (99990, 23), (129990, 4390), (207000, 99),
                   (300000, 504), (300200, 4385)):
        s = sig(i)
        buf[pos:pos + s.size] = s
    eo, ep, _, _ = o.search(buf)
    assert eo.size >= 8
    for mode in modes_for(a):
        off, pat, res = gpu_scan(device, a, buf, mode)
        assert_same(off, pat, eo, ep, f"repetitive mode {mode}")
    off, pat, res = gpu_scan(device, build_product(pats, stride=4), buf, g.MODE_SAMPLED4)
    assert_same(off, pat, eo, ep, "repetitive sampled stride 4")


@pytest.mark.parametrize("dq_cap", [None, 0, 1, "resolve"])
def test_dense_chunks_next_to_filter_chunks_own_by_indexed_window(device, dq_cap, monkeypatch):
    """A dense 2 KiB chunk is walked by the automaton, its neighbours go through the gram filter,
    and an occurrence belongs to the chunk its INDEXED window lies in -- which for signatures that
    start with popular bytes (zero runs, common prologues) is not their first aligned window.
    Zero runs that cover part of a chunk, with signatures planted over every kind of border: the
    oracle's list, neither short nor with duplicates (tools/density_sweep.py found both in round 1's
    ownership rule).  ACM_DQ_CAP = 0 / 1: the scanning warp walks the chunk itself when its list
    of dense chunks is full; "resolve": k_resolve_queue walks them (the path of automata without a
    row-displaced table) instead of the scanning CTA at its end (s4_dense_phase)."""
    if dq_cap == "resolve":
        monkeypatch.setenv("ACM_DENSE_KERNEL", "0")
    elif dq_cap is not None:
        monkeypatch.setenv("ACM_DQ_CAP", str(dq_cap))
    pats = clamav_pats(10000)
    o, a = build_oracle(pats), build_product(pats)
    n = 8 << 20
    buf = synth.stream(n, 2)
    # what the sweep does: the first 2 % / 5 % of every 64 KiB block zeroed -> one dense chunk, the
    # next one partly dense, then random-looking ones
    v = buf.reshape(-1, 64 << 10)
    v[0::2, :1296] = 0
    v[1::2, :3264] = 0
    # signatures that start with zeros or with the popular prologues, over the borders of the zero runs
    zsig = [i for i, (p, _) in enumerate(pats) if p[:4] == b"\0\0\0\0" or p[:4] == bytes.fromhex("e800005d")][:64]
    assert len(zsig) >= 8
    rng = np.random.default_rng(4)
    planted = 0
    for b in range(0, n >> 16):
        base = b << 16
        edge = 1296 if b % 2 == 0 else 3264
        for k, i in enumerate(zsig[:6] + list(rng.integers(0, len(pats), 6))):
            s = np.frombuffer(pats[int(i)][0], dtype=np.uint8)
            pos = base + edge - s.size // 2 + 7 * k if k % 3 == 0 else base + (2048 * (1 + k)) - s.size // 3 + k
            if pos + s.size < base + (64 << 10):
                buf[pos:pos + s.size] = s
                planted += 1
    eo, ep, _, _ = o.search(buf)
    assert eo.size >= planted // 2
    off, pat, res = gpu_scan(device, a, buf, g.MODE_SAMPLED4)
    assert_same(off, pat, eo, ep, f"mixed dense / filter chunks, dq_cap {dq_cap}")
    off, pat, res = gpu_scan(device, build_product(pats, stride=4), buf, g.MODE_SAMPLED4)
    assert_same(off, pat, eo, ep, f"mixed dense / filter chunks, stride 4, dq_cap {dq_cap}")


def test_constant_runs_in_dense_chunks(device):
    """The dense walk skips the rest of a 16-byte vector of equal bytes once a step has left it in the
    same silent state.  Runs of 0x00, 0xFF and 0x90 with signatures that ARE such runs (every position
    of a long enough run reports), that end in one, and that begin with one: the oracle's list."""
    base = clamav_pats(10000)
    extra = [b"\0" * 12, b"\xff" * 16, b"\0" * 11 + b"\x01", b"\x02" + b"\0" * 13, b"\x90" * 10 + b"\xcc\xcc",
             b"\xff" * 9 + b"\0" * 9]
    pats = base + [(p, len(base) + i) for i, p in enumerate(extra)]
    o, a = build_oracle(pats), build_product(pats)
    n = 4 << 20
    buf = synth.stream(n, 5)
    v = buf.reshape(-1, 64 << 10)
    v[0::4, :5000] = 0
    v[1::4, 100:4200] = 0xFF
    v[2::4, 7:2300] = 0x90
    v[3::4, 2048:6144] = 0
    v[3::4, 4000] = 1                       # 0...0 01 inside a zero run
    v[2::4, 2300:2302] = 0xCC
    v[1::4, 4200:4212] = 0                  # ff..ff 00..00
    v[0::4, 6000] = 2
    v[0::4, 6001:6030] = 0                  # 02 00..00 behind a chunk border
    eo, ep, _, _ = o.search(buf)
    assert eo.size > 100000
    off, pat, res = gpu_scan(device, a, buf, g.MODE_SAMPLED4)
    assert_same(off, pat, eo, ep, "constant runs")
    off, pat, res = gpu_scan(device, a, buf, g.MODE_DFA)
    assert_same(off, pat, eo, ep, "constant runs, dfa")


@pytest.mark.parametrize("cap", [1, 3, 16])
def test_resolve_queue_overflow_takes_inline_path(device, cap, monkeypatch):
    """The sampled kernel queues its filter survivors for k_resolve_queue; when a warp's region is
    full it resolves them inline.  A tiny queue (test hook ACM_VQ_CAP) forces that path; planted,
    overlapping and repetitive input must give the oracle's list either way."""
    monkeypatch.setenv("ACM_VQ_CAP", str(cap))
    pats = clamav_pats(10000)
    o, a = build_oracle(pats), build_product(pats)
    n = 3 << 19
    buf, _ = planted_stream(pats, n, seed=11, plants=3000, forced=[(0, 3), (n - len(pats[5][0]), 5)])
    buf[300000:300000 + 20000] = np.tile(np.frombuffer(bytes.fromhex("e800005d81ed0000"), np.uint8), 2500)
    s = np.frombuffer(pats[504][0], dtype=np.uint8)
    for k in range(40):                                   # one signature over and over, 3 bytes apart from the next
        buf[400000 + k * (s.size + 3):400000 + k * (s.size + 3) + s.size] = s
    eo, ep, _, _ = o.search(buf)
    assert eo.size >= 3000
    for stride in (8, 4):
        off, pat, res = gpu_scan(device, build_product(pats, stride=stride) if stride == 4 else a, buf, g.MODE_SAMPLED4)
        assert_same(off, pat, eo, ep, f"queue cap {cap} stride {stride}")


def test_async_scan_two_scanners_and_push(device):
    """acm_scan_device_async / acm_scan_finish with two scanners in flight and the in-step push:
    same list as the synchronous call, including when the step has to be repaired on the host
    (bucket overflow -> two-pass path, output buffer growth) and the push is redone there."""
    pats = clamav_pats(2000)
    o, a = build_oracle(pats), build_product(pats)
    bufs = [planted_stream(pats, (1 << 20) + 123 * k, seed=20 + k, plants=2000 + 300 * k)[0] for k in range(3)]
    exp = [o.search(b)[:2] for b in bufs]
    cap = 1 << 14
    d_bufs = []
    for b in bufs:
        d = device.alloc(b.size + 64)
        device.h2d(d, b)
        d_bufs.append(d)
    region = device.alloc(2 * cap * 8)
    for kw in ({}, {"bucket_shift": 16, "bucket_cap": 32}):       # second shape overflows: host repair
        scs = [g.Scanner(device, a.automaton, 2 << 20, **kw) for _ in range(2)]
        order = [0, 1, 2, 0, 1]
        for i, k in enumerate(order):
            scs[i & 1].scan_async(d_bufs[k], bufs[k].size, push=(region + (i & 1) * cap * 8, cap, 5 << 24))
            if i > 0:
                j = i - 1
                res = scs[j & 1].finish()
                keys = device.d2h(region + (j & 1) * cap * 8, int(res.n_matches) * 8, dtype=np.uint64)
                eo, ep = exp[order[j]]
                assert_same((keys >> np.uint64(24)) - np.uint64(5), (keys & np.uint64(0xFFFFFF)).astype(np.uint32),
                            eo, ep, f"async step {j} {kw}")
                off, pat = scs[j & 1].fetch()
                assert_same(off, pat, eo, ep, f"async fetch {j} {kw}")
                if kw:
                    assert res.fallback == 1
        res = scs[0].finish()
        assert int(res.n_matches) == exp[1][0].size
        with pytest.raises(g.AcmError):
            scs[0].finish()                                   # nothing pending
        # a region that is too small is reported, not overrun
        scs[0].scan_async(d_bufs[0], bufs[0].size, push=(region, 16, 0))
        with pytest.raises(g.AcmError):
            scs[0].finish()
        for sc in scs:
            sc.close()
    for d in d_bufs + [region]:
        device.free(d)


def _word_text(rng, pats, nbytes, seps=b" \n.,"):
    parts, size = [], 0
    while size < nbytes:
        w = pats[int(rng.integers(0, len(pats)))][0]
        if rng.random() < 0.3:                            # a non-word now and then
            w = bytes(rng.integers(97, 123, size=int(rng.integers(1, 9))).tolist())
        sep = seps[int(rng.integers(0, len(seps)))]
        parts.append(w + bytes([sep]))
        size += len(w) + 1
    return np.frombuffer(b"".join(parts)[:nbytes], dtype=np.uint8).copy()


def test_cdfa_windows_shapes_and_class_maps(device):
    """The dense-output kernels (mode 4) on their own: the row-displaced table in shared memory
    (k_scan_rd + k_rd_expand, the default for word lists) and the class-compressed table with
    both byte->column forms (arithmetic range / replicated lookup table) and hot-row budgets from
    "nothing in shared memory" upwards; ragged lengths around chunk (256 B) and region (8 KiB)
    cuts, unaligned emit windows with a hidden stale prefix, log / bucket overflow (exact two-pass
    path), long patterns (halo of two staging pieces, 512-byte chunks) -- always the oracle's list."""
    import os
    rng = np.random.default_rng(77)
    lex = load_patterns("sentiment_categorical.pat.gz")[:1500]
    # long lower-case patterns: Lmax 120 -> halo 119 (two 64-byte staging pieces), 512-byte chunks
    longw = [(bytes(rng.integers(97, 123, size=int(rng.integers(2, 121))).tolist()), i) for i in range(120)]
    longw += [(p[:k], 500 + 10 * i + j) for i, (p, _) in enumerate(longw[:8]) for j, k in enumerate((2, 3, 5, 64, 65))
              if k < len(p)]
    # wide byte span (> 63 values) but few distinct bytes -> lookup-table form
    sym = np.array([0, 7, 65, 66, 200, 201, 255], dtype=np.uint8)
    wide = [(bytes(rng.choice(sym, size=int(rng.integers(4, 10))).tolist()), i) for i in range(300)]
    # four patterns ending in one state (the most this form takes), one of them twice over
    wide += [(bytes([7, 0, 255, 200, 65]), 900), (bytes([0, 255, 200, 65]), 901), (bytes([255, 200, 65]), 902),
             (bytes([200, 65]), 903), (bytes([66, 66]), 904), (bytes([66, 66]), 905)]
    hooks = ("ACM_CD_PLAIN", "ACM_CD_HOT_KB", "ACM_RD_WARPS", "ACM_RD_LOG_CAP")
    for pats, form in ((lex, "range"), (longw, "long"), (wide, "lut")):
        o, a = build_oracle(pats), build_product(pats)
        assert g.lib().acm_automaton_cdfa_classes(a.automaton) > 0
        assert g.lib().acm_automaton_default_mode(a.automaton) == g.MODE_CDFA
        # the row-displaced table exists exactly for the one-range sets, and is the automaton
        assert (a.L.acsm_check_cdfa(a._p, None, None) == 0) == (form != "lut")
        if form != "lut":
            text = _word_text(rng, pats, 300000)
        else:
            text = rng.choice(np.array([0, 7, 65, 66, 200, 201, 255, 33], dtype=np.uint8), size=300000)
        for n in (1, 15, 16, 255, 256, 257, 511, 513, 4096 + 17, 8191, 8192, 8193, 3 * 8192, 5 * 16384 + 1, 300000):
            t = text[:n]
            eo, ep, _, _ = o.search(t)
            # the row-displaced table wholly in shared memory (default; also with 3 warps per CTA),
            # then the plain table with 1 KiB / 64 KiB / the default amount of hot rows in shared memory
            for env in ({}, {"ACM_RD_WARPS": "3"}, {"ACM_CD_PLAIN": "1", "ACM_CD_HOT_KB": "1"},
                        {"ACM_CD_PLAIN": "1", "ACM_CD_HOT_KB": "64"}, {"ACM_CD_PLAIN": "1"}):
                for k in hooks:
                    os.environ.pop(k, None)
                os.environ.update(env)
                off, pat, res = gpu_scan(device, a, t, g.MODE_CDFA)
                assert_same(off, pat, eo, ep, f"cdfa {form} n={n} {env}")
        for k in hooks:
            os.environ.pop(k, None)
        # emit windows: any cut, with the bytes before valid_lo replaced by junk that would match
        eo, ep, _, _ = o.search(text)
        lmax = a.get_max_pattern_size()
        for lo, hi in ((0, 100), (1, 257), (255, 256), (256, 512), (300, 300000), (70001, 140003), (299990, 300000)):
            keep = (eo >= lo) & (eo < hi)
            off, pat, _ = gpu_scan(device, a, text, g.MODE_CDFA, emit_lo=lo, emit_hi=hi)
            assert_same(off, pat, eo[keep], ep[keep], f"cdfa {form} window [{lo},{hi})")
            vlo = max(0, lo - (lmax - 1))
            junk = text.copy()
            if vlo:
                # a pattern planted across valid_lo: matches only if the stale bytes are used
                p0 = np.frombuffer(max((p for p, _ in pats), key=len), dtype=np.uint8)
                cut = p0.size // 2
                if cut and vlo >= cut and vlo + p0.size - cut <= junk.size:
                    junk[vlo - cut:vlo] = p0[:cut]
                    junk[vlo:vlo + p0.size - cut] = p0[cut:]
            eo2, ep2, _, _ = o.search(junk[vlo:])
            keep2 = (eo2 + vlo >= lo) & (eo2 + vlo < hi)
            off, pat, _ = gpu_scan(device, a, junk, g.MODE_CDFA, emit_lo=lo, emit_hi=hi, valid_lo=vlo)
            assert_same(off, pat, eo2[keep2] + np.uint64(vlo), ep2[keep2], f"cdfa {form} valid_lo {vlo} window [{lo},{hi})")
        # overflow -> exact two-pass path (direct writes, no sort needed): a 320-entry region log for
        # k_scan_rd (a region of this text has ~800 hits), 32-hit rows for the plain kernel
        os.environ["ACM_RD_LOG_CAP"] = "320"
        off, pat, res = gpu_scan(device, a, text, g.MODE_CDFA, bucket_shift=10, bucket_cap=32)
        os.environ.pop("ACM_RD_LOG_CAP")
        assert res.fallback == 1
        assert_same(off, pat, eo, ep, f"cdfa {form} overflow")
        # a match list that outgrows the scanner's first output buffer (65 536 keys): grown, post-pass relaunched
        if form == "range":
            big = np.tile(text, 8)
            eb, pb, _, _ = o.search(big)
            assert eb.size > (1 << 16)
            off, pat, res = gpu_scan(device, a, big, g.MODE_CDFA)
            assert res.fallback == 0
            assert_same(off, pat, eb, pb, "cdfa range, output buffer grown")


def _split_len(a):
    return g.lib().acm_automaton_split_len(a.automaton)


def test_mixed_set_sampled_plus_short_pass(device):
    """ClamAV signatures plus a few patterns of 1..9 bytes: mode auto stays on the sampled filter
    (long patterns) and adds the start-filter pass for the short ones; same list as the oracle,
    and as the other kernels, also with tiny buckets (overflow -> exact two-pass path)."""
    base = clamav_pats(2000)
    short = [(b"\x00", 9000), (b"MZ", 9001), (b"\x90\x90\x90", 9002), (b"PE\x00\x00", 9003),
             (b"virus", 9004), (b"abcdef", 9005), (b"\xe8\x00\x00\x00\x00\x5d\x81", 9006),
             (base[5][0][:9], 9007), (base[5][0][-6:], 9008), (base[7][0][3:8], 9009)]
    pats = base + short
    o, a = build_oracle(pats), build_product(pats)
    assert _split_len(a) == 10 and sample_stride(a) == 8
    assert g.lib().acm_automaton_default_mode(a.automaton) == g.MODE_SAMPLED4
    buf, _ = planted_stream(pats, 3 << 20, seed=21, plants=600)
    buf[:2] = np.frombuffer(b"MZ", dtype=np.uint8)
    buf[-5:] = np.frombuffer(b"virus", dtype=np.uint8)
    eo, ep, _, _ = o.search(buf)
    assert (ep >= 2000).sum() > 1000 and (ep < 2000).sum() > 300
    for mode in [0] + modes_for(a):
        off, pat, res = gpu_scan(device, a, buf, mode)
        assert_same(off, pat, eo, ep, f"mixed set mode {mode}")
    off, pat, res = gpu_scan(device, a, buf, 0, bucket_shift=10, bucket_cap=4)
    assert res.mode == g.MODE_SAMPLED4 and res.fallback
    assert_same(off, pat, eo, ep, "mixed set, overflow fallback")
    # shard windows: the short pass has its own (shorter) lead-in
    for lo, hi in ((1, 2), (4096, 70001), ((1 << 20) + 3, 3 << 20)):
        keep = (eo >= lo) & (eo < hi)
        off, pat, _ = gpu_scan(device, a, buf, 0, emit_lo=lo, emit_hi=hi)
        assert_same(off, pat, eo[keep], ep[keep], f"mixed set window [{lo}, {hi})")
    # split at 7 / stride 4 when many patterns are under 10 bytes
    mid = [(bytes([200 + i, 7, 7, 7, 7, 7, 7, i]), 9100 + i) for i in range(40)]
    pats2 = base[:200] + short[:6] + mid
    o2, a2 = build_oracle(pats2), build_product(pats2)
    assert _split_len(a2) == 7 and sample_stride(a2) == 4
    buf2, _ = planted_stream(pats2, 1 << 20, seed=22, plants=500)
    eo2, ep2, _, _ = o2.search(buf2)
    for mode in [0] + modes_for(a2):
        off, pat, _ = gpu_scan(device, a2, buf2, mode)
        assert_same(off, pat, eo2, ep2, f"mixed set (split 7) mode {mode}")


def test_mixed_set_on_repetitive_input(device):
    """Dense chunks of a mixed set go through the in-kernel DFA fallback, which must leave the
    short patterns to the start-filter pass (no duplicates)."""
    base = clamav_pats(2000)
    pats = base + [(b"\x00", 9000), (b"\x00\x00\x00", 9001), (b"ab", 9002), (bytes(12), 9003)]
    o, a = build_oracle(pats), build_product(pats)
    assert _split_len(a) == 10
    buf = np.zeros(1 << 18, dtype=np.uint8)
    buf[100000:100000 + 4000] = np.frombuffer(b"ab" * 2000, dtype=np.uint8)
    sig = np.frombuffer(base[3][0], dtype=np.uint8)
    buf[50000:50000 + sig.size] = sig
    eo, ep, _, _ = o.search(buf)
    off, pat, res = gpu_scan(device, a, buf, 0)
    assert res.mode == g.MODE_SAMPLED4
    assert_same(off, pat, eo, ep, "mixed set on zero pages")
