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
n = 64 << 20
d = dev.alloc(n + 64)
dev.h2d(d, np.tile(text, n // text.size + 1)[:n])
sc = g.Scanner(dev, a.automaton, n, mode=2, timing=True, bucket_shift=12, bucket_cap=1024)
for _ in range(3):
    r = sc.scan_device(d, n)
print(r.n_matches, r.ms_scan, n / r.ms_scan / 1e6, "GB/s")
