"""Scratch timing of the AC_ushorts path (iacsm_* automaton, k_scan_dfa<uint16_t>) on synthetic
packet-size trains: signatures are sequences of 6..30 packet sizes drawn from a skewed size
distribution, the stream is flows of 50..2000 packets separated by an out-of-alphabet token
(what cli/b200_flow_grep lays out), with signatures planted.  Not bench.py.

    python tools/flow_bench.py [MiB of tokens] [signatures] [chunk,chunk,...]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_pattern_matching_b200 as g  # noqa: E402
from oracle_lib import Oracle  # noqa: E402


def workload(nsig, ntok, seed=6):
    rng = np.random.default_rng(seed)
    sizes = rng.permutation(2048)
    w = 1.0 / np.arange(1, 2049) ** 1.1                    # a few packet sizes dominate
    w /= w.sum()
    sigs = [sizes[rng.choice(2048, size=int(rng.integers(6, 31)), p=w)].astype(np.uint16) for _ in range(nsig)]
    text = sizes[rng.choice(2048, size=ntok, p=w)].astype(np.uint16)
    pos = 0
    while pos < ntok:                                       # flow boundaries
        pos += int(rng.integers(50, 2000))
        if pos < ntok:
            text[pos] = 0xFFFF
    for k in range(ntok // 6000):                           # planted signatures
        s = sigs[int(rng.integers(0, nsig))]
        p = int(rng.integers(0, ntok - s.size))
        text[p:p + s.size] = s
    return sigs, text


def main():
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    nsig = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    chunks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4096, 0]
    ntok = (mib << 20) // 2
    t0 = time.time()
    sigs, text = workload(nsig, ntok)
    m, o = g.Iacsm(), Oracle(2048)
    for i, s in enumerate(sigs):
        m.add_pattern(s, i)
        o.add(s, i)
    m.compile()
    o.compile()
    dev = g.Device(0)
    m.gen_state_table(0, dev.handle, None)
    print(f"[flows] {nsig} signatures, {m.get_states()} states, table {m.get_size() / 2**20:.0f} MiB "
          f"(reference layout), {ntok >> 20} Mi tokens, setup {time.time() - t0:.1f}s", flush=True)
    d = dev.alloc(text.nbytes + 64)
    dev.h2d(d, text)
    # parity on a prefix the oracle walks in a second
    npre = min(ntok, 8 << 20)
    eo, ep, _, _ = o.search(text[:npre])
    for chunk in chunks:
        sc = g.Scanner(dev, m.automaton, ntok, timing=True, dfa_chunk=chunk)
        r = sc.scan_device(d, npre)
        off, pat = sc.fetch()
        ok = np.array_equal(off, eo) and np.array_equal(pat, ep)
        best = None
        for _ in range(5):
            r = sc.scan_device(d, ntok)
            if best is None or r.ms_scan < best.ms_scan:
                best = g._lib.ScanResult.from_buffer_copy(r)
        gbs = text.nbytes / best.ms_scan / 1e6
        print(f"[flows] dfa_chunk={chunk:5d} prefix parity={'ok' if ok else 'MISMATCH'} ({eo.size} matches) "
              f"matches={best.n_matches} scan {best.ms_scan:.3f} ms = {gbs:.1f} GB/s = "
              f"{ntok / best.ms_scan / 1e6:.1f} G tokens/s, total {best.ms_total:.3f} ms", flush=True)
        sc.close()
    dev.free(d)


if __name__ == "__main__":
    main()
