"""GPU parity through the reference-shaped entry points: the five-call sequence of
cpu_worker() (reference ocl_aho_grep.c:116-137) over databuf / ocl_worker / ocl_aho_match,
the ushort-symbol twin, and the optional post-passes ocl_prefix_sum / ocl_compact_array /
ocl_bitonic_sort (checks modelled on the reference's own self-test, databuf.c:935-1021)."""
import ctypes as C
import os

import numpy as np
import pytest

import gpu_pattern_matching_b200 as g
from gpu_pattern_matching_b200 import _lib, synth
from helpers import build_oracle, clamav_pats, load_patterns, planted_stream
from oracle_lib import Oracle, materialize, read_fixture

pytestmark = pytest.mark.gpu


def _worker(L, pat_path, hex_pat, chunks, chunk_size, max_results=16, text_mode=0):
    w = L.ocl_worker_ctx_create(0)
    assert w, _lib.last_error()
    names = (C.c_char_p * 1)(b"stream")
    fds = (C.c_int * 1)(-1)
    rc = L.ocl_worker_ctx_init(w, 0, 1024, chunks, 0, str(pat_path).encode(), int(hex_pat), -1,
                               chunk_size, max_results, 0, text_mode, 0, 0, 1, 1, fds, names)
    assert rc == 0, _lib.last_error()
    return w, (names, fds)


def _run_rounds(L, w, data, collect):
    """databuf_add_fd from a pipe-like fd -> H2D -> match -> D2H -> process_results -> reset."""
    ctx = w.contents
    db = ctx.db
    r, wfd = os.pipe()
    total = 0

    @_lib.MATCH_CB
    def cb(file_idx, pat_idx, chunk_idx, offset, uarg):
        collect.append((file_idx, pat_idx, chunk_idx, offset))
        return 0
    # feed through a temp file (a pipe would block on large writes)
    os.close(r)
    os.close(wfd)
    import tempfile
    with tempfile.TemporaryFile() as f:
        f.write(data.tobytes())
        f.flush()
        f.seek(0)
        fd = f.fileno()
        os.lseek(fd, 0, os.SEEK_SET)
        done = False
        base = 0
        while not done:
            rd = C.c_size_t(0)
            e = L.databuf_add_fd(db, fd, 0, C.byref(rd))
            if rd.value == 0:
                done = True
            elif e not in (-1, -2):
                continue
            if db.contents.chunks > 0:
                nbytes = db.contents.bytes
                L.databuf_copy_host_to_device(db, ctx.cl.queue)
                L.ocl_aho_match(C.byref(ctx.cl), db, ctx.acsm, 1024, 1)
                assert L.databuf_status(db) == 0, _lib.last_error()
                L.databuf_copy_device_to_host(db, ctx.cl.queue)
                before = len(collect)
                n = L.databuf_process_results(db, cb, None)
                assert n == len(collect) - before == L.databuf_match_count(db)
                for k in range(before, len(collect)):
                    fi, pi, ci, off = collect[k]
                    collect[k] = (fi, pi, ci, off + base)
                total += n
                base += nbytes
                L.databuf_reset(db)
    return total


def test_worker_five_call_sequence_matches_oracle(lib, tmp_path):
    L = lib
    pats = clamav_pats(2000)
    o = build_oracle(pats)
    chunk, chunks = 4096, 64                       # 256 KiB buffers -> several rounds
    n = 5 * chunk * chunks                         # whole buffers: no zero padding inside the stream
    forced = [(chunk * chunks - 20, 31), (2 * chunk * chunks - 1, 32), (chunk - 3, 33)]
    buf, _ = planted_stream(pats, n, seed=17, plants=300, forced=forced)
    # the worker reads a pattern FILE: write the first 2000 signatures
    pf = tmp_path / "sigs2000.txt"
    pf.write_bytes(b"\n".join(read_fixture("clamav_sigs_15000.hex.gz").split(b"\n")[:2000]) + b"\n")
    w, keep = _worker(L, pf, True, chunks, chunk)
    assert w.contents.patterns_size == 2000 and w.contents.patterns[5].n == len(pats[5][0])
    got = []
    total = _run_rounds(L, w, buf, got)
    eo, ep, _, _ = o.search(buf)
    assert total == eo.size
    # callback offset = end offset + 1 (reference databuf.c:771), chunk = offset // chunk size
    assert [x[3] - 1 for x in got] == eo.tolist()
    assert [x[1] for x in got] == ep.tolist()
    assert all(x[2] == ((x[3] - 1) % (chunk * chunks)) // chunk for x in got)
    L.ocl_worker_ctx_free(w)


def test_databuf_bucket_and_compact_views(lib, tmp_path):
    """h_results/h_results2 (column-major buckets, ahomatch.cl:67-73) and the compact arrays
    (compactarray.cl:49-55) describe the same matches."""
    L = lib
    pf = tmp_path / "p.txt"
    pf.write_bytes(b"abc\nbc\nzz\n")
    # 9 x 64: databuf_add_chunk refuses a chunk that would exactly fill the buffer
    # (bytes + len >= size, reference databuf.c:503)
    w, keep = _worker(L, pf, False, 9, 64, max_results=4)
    db = w.contents.db
    text = (b"..abc...bc....zzzz..abc" + b"." * 41) * 8           # 64 bytes x 8 chunks
    text = text[:512]
    for i in range(8):
        assert L.databuf_add_chunk(db, text[64 * i:64 * i + 64], 64, 7, 0) != -3
    assert db.contents.chunks == 8 and db.contents.bytes == 512
    L.databuf_copy_host_to_device(db, None)
    L.ocl_aho_match(C.byref(w.contents.cl), db, w.contents.acsm, 256, 0)
    L.databuf_copy_device_to_host(db, None)
    o = Oracle(256)
    for i, p in enumerate([b"abc", b"bc", b"zz"]):
        o.add(p, i)
    o.compile()
    eo, ep, _, _ = o.search(text)
    d = db.contents
    total = d.h_results_comp[0]
    assert total == eo.size == d.h_results2_comp[0]
    assert [d.h_results_comp[1 + i] for i in range(total)] == ep.tolist()
    assert [d.h_results2_comp[1 + i] for i in range(total)] == eo.tolist()
    R, Cn = d.max_results, d.chunks
    for c in range(Cn):
        idx = np.nonzero(eo // 64 == c)[0]
        assert d.h_results[c] == idx.size == d.h_results2[c]
        for k, m in enumerate(idx[:R - 1]):
            assert d.h_results[(k + 1) * Cn + c] == ep[m] and d.h_results2[(k + 1) * Cn + c] == eo[m]
    assert [d.h_prefixsum[c] for c in range(Cn)] == np.concatenate(([0], np.cumsum(
        [d.h_results[c] for c in range(Cn)])[:-1])).tolist()
    L.ocl_worker_ctx_free(w)


def test_text_mode_lines(lib, tmp_path):
    L = lib
    pats = load_patterns("sentiment_categorical.pat.gz")
    pf = materialize("sentiment_categorical.pat.gz", tmp_path)
    w, keep = _worker(L, pf, False, 4096, 256, text_mode=1)
    text = read_fixture("kat_text_a.txt.gz")
    tf = tmp_path / "in.txt"
    tf.write_bytes(text)
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    fp = libc.fopen(str(tf).encode(), b"r")
    db = w.contents.db
    rb, rl = C.c_size_t(0), C.c_size_t(0)
    e = L.databuf_add_fp(db, fp, 0, 1, C.byref(rb), C.byref(rl))
    libc.fclose(fp)
    assert e > 0 and rb.value == len(text) and rl.value == text.count(b"\n")
    d = db.contents
    assert d.chunks == text.count(b"\n") and all(d.h_indices[i] % 16 == 0 for i in range(d.chunks))
    L.databuf_copy_host_to_device(db, None)
    L.ocl_aho_match(C.byref(w.contents.cl), db, w.contents.acsm, 256, 1)
    L.databuf_copy_device_to_host(db, None)
    # oracle over exactly the bytes the buffer holds (lines padded with zeros to 16)
    hbuf = np.ctypeslib.as_array(d.h_data, shape=(d.bytes,)).copy()
    o = build_oracle(pats)
    eo, ep, _, _ = o.search(hbuf)
    total = d.h_results_comp[0]
    assert total == eo.size and total >= 65
    assert [d.h_results2_comp[1 + i] for i in range(total)] == eo.tolist()
    assert [d.h_results_comp[1 + i] for i in range(total)] == ep.tolist()
    L.ocl_worker_ctx_free(w)


def test_ushort_automaton_on_device(lib, device):
    """AC_ushorts path: packet-size trains, alphabet 2048, matches report the iid."""
    L = lib
    m = g.Iacsm()
    o = Oracle(2048)
    for k, line in enumerate(read_fixture("ushort_signatures.txt").decode().splitlines()):
        m.add_fullpattern(line.split(";")[0], k)
        o.add_csv(line.split(";")[0], k)
    extra = [[6, 5], [2047, 0, 2047], [5, 4, 3, 2, 1, 0, 1, 2, 3, 4, 666]]
    for k, s in enumerate(extra):
        m.add_pattern(s, 50 + k)
        o.add(s, 50 + k)
    m.compile()
    o.compile()
    m.gen_state_table(0, device.handle, None)
    rng = np.random.default_rng(3)
    toks = rng.choice(np.array([0, 1, 2, 3, 4, 5, 6, 7, 666, 676, 2047, 3000], dtype=np.uint16), size=20000)
    toks[:16] = [9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 1, 2, 3, 4, 666, 676]
    eo, ep, _, fin = o.search(toks)
    d = device.alloc(toks.nbytes + 64)
    device.h2d(d, toks)
    for chunk in (0, 64, 1000):
        sc = g.Scanner(device, m.automaton, toks.size, dfa_chunk=chunk)
        res = sc.scan_device(d, toks.size)
        off, pat = sc.fetch()
        assert res.mode == g.MODE_DFA
        assert np.array_equal(off, eo) and np.array_equal(pat, ep)
        sc.close()
    # flow files shipped with the reference: one CSV line of tokens per packet train
    for flow in ("ushort_flow_333.txt", "ushort_flow_666.txt"):
        t = np.array([int(x) for line in read_fixture(flow).decode().split() for x in line.split(",") if x],
                     dtype=np.uint16)
        eo2, ep2, _, _ = o.search(t)
        d2 = device.alloc(t.nbytes + 64)
        device.h2d(d2, t)
        sc = g.Scanner(device, m.automaton, t.size)
        sc.scan_device(d2, t.size)
        off, pat = sc.fetch()
        assert np.array_equal(off, eo2) and np.array_equal(pat, ep2)
        sc.close()
        device.free(d2)
    device.free(d)


def test_exclusive_scan_matches_cumsum(lib, device):
    L = lib
    rng = np.random.default_rng(0)
    for n in (1, 2, 255, 2048, 2049, 100000, 3_000_001):
        for kind in ("rand", "zero", "max", "hot"):
            a = {"rand": rng.integers(0, 16, n), "zero": np.zeros(n), "max": np.full(n, 15),
                 "hot": np.eye(1, n, n // 2).ravel() * 1000}[kind].astype(np.uint32)
            di, do, dt = device.alloc(4 * n), device.alloc(4 * n), device.alloc(16)
            device.h2d(di, a)
            assert L.acm_exclusive_scan_u32(device.handle, di, do, n, dt) == 0
            got = device.d2h(do, 4 * n, np.uint32)
            tot = device.d2h(dt, 4, np.uint32)[0]
            exp = np.concatenate(([0], np.cumsum(a, dtype=np.uint64)[:-1])).astype(np.uint32)
            assert np.array_equal(got, exp) and tot == a.sum()
            for p in (di, do, dt):
                device.free(p)


def test_prefix_sum_and_compact_like_reference_selftest(lib, device):
    """reference databuf.c:935-1021: random per-chunk counts -> ocl_prefix_sum vs serial sum ->
    ocl_compact_array: comp[0] == total and the values arrive in chunk order."""
    L = lib
    conf = _lib.Clconf()
    L.clinitctx(C.byref(conf), 0, -1)
    assert conf.ctx
    chunks, R = 5000, 16
    db = L.databuf_new(chunks, 64, R, 0, C.byref(conf))
    assert db and L.databuf_alloc_postpass(db) == 0
    d = db.contents
    rng = np.random.default_rng(9)
    counts = rng.integers(0, R, chunks).astype(np.int32)          # < R: every match fits its bucket
    res = np.zeros(R * chunks + 1, dtype=np.int32)
    res2 = np.zeros(R * chunks + 1, dtype=np.int32)
    res[:chunks] = counts
    res2[:chunks] = counts
    run = 0
    for c in range(chunks):
        for k in range(counts[c]):
            res[(k + 1) * chunks + c] = run            # like the reference test: comp[i+1] == i
            res2[(k + 1) * chunks + c] = 1000000 + run
            run += 1
    res[R * chunks] = 4242                             # the "last state" slot
    res2[R * chunks] = 4242
    dev = g.Device.__new__(g.Device)
    dev.L, dev._h, dev.ordinal = L, C.c_void_p(conf.ctx), 0
    dev.h2d(d.d_results, res)
    dev.h2d(d.d_results2, res2)
    d.chunks = chunks
    L.ocl_prefix_sum(C.byref(conf), db, chunks)
    pre = dev.d2h(d.d_prefixsum, 4 * chunks, np.int32)
    assert np.array_equal(pre, np.concatenate(([0], np.cumsum(counts)[:-1])).astype(np.int32))
    L.ocl_compact_array(C.byref(conf), db, 1024)
    comp = dev.d2h(d.d_results_comp, 4 * (run + 2), np.int32)
    comp2 = dev.d2h(d.d_results2_comp, 4 * (run + 2), np.int32)
    assert comp[0] == run and comp2[0] == run
    assert np.array_equal(comp[1:run + 1], np.arange(run, dtype=np.int32))
    assert np.array_equal(comp2[1:run + 1], 1000000 + np.arange(run, dtype=np.int32))
    assert comp[run + 1] == 4242 and comp2[run + 1] == 4242
    L.databuf_free(db, 0, None)
    L.clfreectx(C.byref(conf))


def test_radix_sort_and_pair_sort(lib, device):
    L = lib
    rng = np.random.default_rng(2)
    for n in (1, 2, 1000, 4096, 4097, 1 << 20):
        keys = rng.integers(0, 1 << 63, n, dtype=np.uint64)
        keys[: n // 3] &= np.uint64(0xFFFF)                       # many equal high digits
        dk, dt = device.alloc(8 * n), device.alloc(8 * n)
        device.h2d(dk, keys)
        assert L.acm_radix_sort_u64(device.handle, dk, dt, n, 0, 64, 0) == 0
        assert np.array_equal(device.d2h(dk, 8 * n, np.uint64), np.sort(keys))
        device.h2d(dk, keys)
        assert L.acm_radix_sort_u64(device.handle, dk, dt, n, 0, 64, 1) == 0
        assert np.array_equal(device.d2h(dk, 8 * n, np.uint64), np.sort(keys)[::-1])
        device.free(dk)
        device.free(dt)
    # key/value pairs through the reference-shaped entry point (any length, both directions)
    conf = _lib.Clconf()
    L.clinitctx(C.byref(conf), 0, -1)
    for n in (2, 1000, 1 << 16, (1 << 16) + 3):
        k = rng.integers(0, 5000, n).astype(np.uint32)
        v = np.arange(n, dtype=np.uint32)
        dk, dv, ok, ov = (device.alloc(4 * n) for _ in range(4))
        device.h2d(dk, k)
        device.h2d(dv, v)
        for direction in (1, 0):
            assert L.ocl_bitonic_sort(C.byref(conf), ok, ov, dk, dv, 1, n, direction) == 0
            gk, gv = device.d2h(ok, 4 * n, np.uint32), device.d2h(ov, 4 * n, np.uint32)
            order = np.lexsort((v, k))
            if direction == 0:
                order = order[::-1]
            assert np.array_equal(gk, k[order]) and np.array_equal(gv, v[order])
        for p in (dk, dv, ok, ov):
            device.free(p)
    assert L.ocl_bitonic_sort(C.byref(conf), None, None, None, None, 1, 1, 1) == -2
    L.clfreectx(C.byref(conf))


def _five_calls(L, w, chunks_bytes=None):
    """copy_h2d -> match(stream=1) -> copy_d2h -> process_results (collect) -> reset; returns the callbacks"""
    ctx = w.contents
    db = ctx.db
    got = []

    @_lib.MATCH_CB
    def cb(file_idx, pat_idx, chunk_idx, offset, uarg):
        got.append((file_idx, pat_idx, chunk_idx, offset))
        return 0
    L.databuf_copy_host_to_device(db, ctx.cl.queue)
    L.ocl_aho_match(C.byref(ctx.cl), db, ctx.acsm, 1024, 1)
    assert L.databuf_status(db) == 0, _lib.last_error()
    L.databuf_copy_device_to_host(db, ctx.cl.queue)
    n = L.databuf_process_results(db, cb, None)
    assert n == len(got) == L.databuf_match_count(db)
    L.databuf_reset(db)
    return got


def test_per_file_semantics_and_the_reference_quirk(lib, tmp_path):
    """the carry is reset and the stream cut at every file change.  A signature whose
    halves are the end of file A and the start of file B must NOT be reported; the same split
    across two buffers of ONE file must; a signature completed by a chunk's zero padding must not.
    With the quirk switched back on (databuf_set_file_semantics(db, 0)) the reference's one-stream
    behaviour returns (ahomatch.cl:38-45,86-93: one state carried across everything)."""
    L = lib
    pf = tmp_path / "p.txt"
    pf.write_bytes(b"HELLOWORLD\nabcdefgh\nzz\n")
    # "tail0" ends with a zero byte: matches only if padding counts as stream
    pfx = tmp_path / "px.txt"
    pfx.write_bytes(b"48454c4c4f574f524c44\n6162636465666768\n7a7a\n7461696c00\n")      # the same three + "tail\0" (hex)
    chunk, chunks = 64, 4                                  # 256-byte buffers
    for quirk in (0, 1):
        w, keep = _worker(L, pfx, True, chunks, chunk)
        db = w.contents.db
        L.databuf_set_file_semantics(db, 0 if quirk else 1)
        fa = tmp_path / "a.bin"
        fb = tmp_path / "b.bin"
        # file A: exactly two chunks, ends with "HELLO"; file B starts with "WORLD"
        fa.write_bytes(b"." * (2 * chunk - 5) + b"HELLO")
        fb.write_bytes(b"WORLD" + b"." * 20 + b"zz" + b"." * 10 + b"tail")           # 41 bytes: padded chunk, ends "tail" + zeros
        rd = C.c_size_t(0)
        for fid, f in ((0, fa), (1, fb)):
            fd = os.open(f, os.O_RDONLY)
            L.databuf_add_fd(db, fd, fid, C.byref(rd))
            os.close(fd)
        assert db.contents.chunks == 3
        got = _five_calls(L, w)
        names = sorted((g_[0], g_[1]) for g_ in got)
        if quirk:
            # one stream: HELLO|WORLD across the files (pattern 0, reported in file 1), zz, and "tail\\0" made of padding
            assert names == [(1, 0), (1, 2), (1, 3)]
        else:
            assert names == [(1, 2)]                                   # only zz, inside file B
        # --- one file over two buffers: the split signature IS found, once, in the second buffer
        fc = tmp_path / "c.bin"
        fc.write_bytes(b"." * (chunks * chunk - 4) + b"abcdefgh" + b"." * 60)         # "abcd" | "efgh" across the buffer cut
        fd = os.open(fc, os.O_RDONLY)
        e = L.databuf_add_fd(db, fd, 5, C.byref(rd))
        assert e in (-1, -2) and rd.value == chunks * chunk
        first = _five_calls(L, w)
        L.databuf_add_fd(db, fd, 5, C.byref(rd))
        second = _five_calls(L, w)
        os.close(fd)
        assert [g_[1] for g_ in first] == [] and [(g_[0], g_[1], g_[3]) for g_ in second] == [(5, 1, 4)]
        # --- a different file right behind it: its first bytes complete nothing that file 5 left in the carry
        fd2 = tmp_path / "d.bin"
        fd2.write_bytes(b"." * (chunks * chunk - 5) + b"HELLO")
        fd3 = tmp_path / "e.bin"
        fd3.write_bytes(b"WORLD" + b"." * 59)
        fd = os.open(fd2, os.O_RDONLY)
        L.databuf_add_fd(db, fd, 7, C.byref(rd))
        os.close(fd)
        assert _five_calls(L, w) == []
        fd = os.open(fd3, os.O_RDONLY)
        L.databuf_add_fd(db, fd, 8, C.byref(rd))
        os.close(fd)
        cross = _five_calls(L, w)
        assert [(g_[0], g_[1]) for g_ in cross] == ([(8, 0)] if quirk else [])
        L.ocl_worker_ctx_free(w)
