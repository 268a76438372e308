"""Scratch: sampled4 throughput on adversarial (all-zero / periodic) input."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gpu_pattern_matching_b200 as g
from helpers import build_product, clamav_pats
dev = g.Device(0)
a = build_product(clamav_pats(10000))
n = 256 << 20
d = dev.alloc(n + 64)
for name, fill in (("zeros", b"\0"), ("e800005d", bytes.fromhex("e800005d")), ("text", b"The quick brown fox jumps over the lazy dog. ")):
    pat = np.frombuffer(fill, dtype=np.uint8)
    dev.h2d(d, np.tile(pat, n // pat.size + 1)[:n])
    for mode in (1, 2):
        sc = g.Scanner(dev, a.automaton, n, mode=mode, timing=True)
        best = min(sc.scan_device(d, n).ms_scan for _ in range(3))
        r = sc.last
        print(f"{name:10s} mode {g.MODE_NAMES[mode]:9s} {n/best/1e6:9.1f} GB/s matches {r.n_matches} fallback {r.fallback}")
        sc.close()
