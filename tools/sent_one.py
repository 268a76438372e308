"""Scratch: one scan of the sentiment-over-text config in one mode (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gpu_pattern_matching_b200 as g
from gpu_pattern_matching_b200 import synth
from helpers import build_product, load_patterns
from oracle_lib import read_fixture
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = (int(sys.argv[2]) if len(sys.argv) > 2 else 256) << 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = g.Device(0)
a = build_product(load_patterns("sentiment_categorical.pat.gz"))
words = [l.split(b"\t")[0] for l in read_fixture("english_top5000.txt.gz").split(b"\n") if l]
text = synth.english_like(words, 32 << 20, seed=4)
d = dev.alloc(n + 64)
dev.h2d(d, np.tile(text, n // text.size + 1)[:n])
sc = g.Scanner(dev, a.automaton, n, mode=mode, timing=True)
for _ in range(reps):
    r = sc.scan_device(d, n)
    print(f"mode {g.MODE_NAMES[mode]} matches {r.n_matches} K1 {r.ms_scan:.3f} ms ({n / r.ms_scan / 1e6:.1f} GB/s) K2 {r.ms_prefix:.3f} K3 {r.ms_compact:.3f} total {r.ms_total:.3f} ms ({n / r.ms_total / 1e6:.1f} GB/s)", flush=True)
