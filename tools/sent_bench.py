"""Scratch: sentiment lexicon over English-like text (BASELINE config 4)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gpu_pattern_matching_b200 as g
from gpu_pattern_matching_b200 import synth
from helpers import build_product, load_patterns
from oracle_lib import read_fixture
dev = g.Device(0)
a = build_product(load_patterns("sentiment_categorical.pat.gz"))
words = [l.split(b"\t")[0] for l in read_fixture("english_top5000.txt.gz").split(b"\n") if l]
text = synth.english_like(words, 32 << 20, seed=4)
n = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 256 << 20
d = dev.alloc(n + 64)
dev.h2d(d, np.tile(text, n // text.size + 1)[:n])
for mode in (4, 2, 3):
    for shift, cap in (((0, 0),) if mode == 4 else ((0, 0), (12, 1024))):
        try:
            sc = g.Scanner(dev, a.automaton, n, mode=mode, timing=True, bucket_shift=shift, bucket_cap=cap)
        except Exception as e:
            print("skip", shift, cap, e); continue
        rs = [sc.scan_device(d, n) for _ in range(3)]
        r = min(rs, key=lambda r: r.ms_total)
        r = g._lib.ScanResult.from_buffer_copy(r)
        print(f"mode {g.MODE_NAMES[mode]:7s} shift {shift} cap {cap}: matches {r.n_matches} ({n / max(1, r.n_matches):.1f} B/match) fallback {r.fallback} "
              f"K1 {r.ms_scan:.3f} ms ({n / r.ms_scan / 1e6:.1f} GB/s) K2 {r.ms_prefix:.3f} K3/sort {r.ms_compact:.3f} total {r.ms_total:.3f} ms ({n / r.ms_total / 1e6:.1f} GB/s)", flush=True)
        sc.close()
