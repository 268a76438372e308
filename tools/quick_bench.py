"""Scratch timing of the scan kernels on device-resident synthetic data (not bench.py)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_pattern_matching_b200 as g  # noqa: E402
from gpu_pattern_matching_b200 import synth  # noqa: E402
from helpers import build_product, clamav_pats, load_patterns  # noqa: E402
from oracle_lib import read_fixture  # noqa: E402


def run(dev, name, pats, n, modes, text=None, plants=0, iters=5):
    t0 = time.time()
    a = build_product(pats)
    print(f"[{name}] {len(pats)} patterns, {a.get_states()} states, table "
          f"{a.get_size() / 2**20:.1f} MiB, build+upload {time.time() - t0:.2f}s", flush=True)
    d = dev.alloc(n + 64)
    if text is None:
        dev.synth_fill(d, n, seed=2)
        if plants:
            pl = synth.Plants([p for p, _ in pats], n, plants, 2)
            dev.plant(d, n, 0, pl)
    else:
        reps = (n + text.size - 1) // text.size
        dev.h2d(d, np.tile(text, reps)[:n])
    dev.sync()
    for mode in modes:
        sc = g.Scanner(dev, a.automaton, n, mode=mode, timing=True)
        best = None
        for it in range(iters):
            r = sc.scan_device(d, n)
            if best is None or r.ms_scan < best.ms_scan:
                best = g._lib.ScanResult.from_buffer_copy(r)
        gbs = n / best.ms_scan / 1e6
        print(f"[{name}] mode {g.MODE_NAMES[mode]:9s} n={n >> 20} MiB matches={best.n_matches} "
              f"fallback={best.fallback} scan {best.ms_scan:.3f} ms = {gbs:.1f} GB/s "
              f"({gbs / 6531.6 * 100:.1f}% of measured HBM) prefix {best.ms_prefix:.3f} ms "
              f"compact {best.ms_compact:.3f} ms total {best.ms_total:.3f} ms", flush=True)
        if os.environ.get("ACM_TRACE") and mode == 1:
            import ctypes as C
            tr = np.zeros(148 * 4, dtype=np.uint64)
            g.lib().acm_scan_trace(sc._h, tr.ctypes.data_as(g._lib.u64p), 148)
            tr = tr.reshape(148, 4).astype(np.int64)
            t0 = tr[:, 0].min()
            print("  trace(us): entry min/max %.1f/%.1f ready min/max %.1f/%.1f exit min/max %.1f/%.1f chunks/CTA min/max %d/%d" % (
                (tr[:, 0].min() - t0) / 1e3, (tr[:, 0].max() - t0) / 1e3, (tr[:, 1].min() - t0) / 1e3,
                (tr[:, 1].max() - t0) / 1e3, (tr[:, 2].min() - t0) / 1e3, (tr[:, 2].max() - t0) / 1e3,
                tr[:, 3].min(), tr[:, 3].max()), flush=True)
            ex = np.sort((tr[:, 2] - t0) / 1e3)
            print("  exit times (us), sorted: " + " ".join(f"{x:.0f}" for x in ex), flush=True)
            order = np.argsort(tr[:, 2])
            print("  CTA ids by exit time:     " + " ".join(str(i) for i in order), flush=True)
            print("  chunks by exit time:      " + " ".join(str(int(tr[i, 3])) for i in order), flush=True)
        sc.close()
    dev.free(d)
    a.free()


if __name__ == "__main__":
    dev = g.Device(0)
    n = int(sys.argv[1]) << 20 if len(sys.argv) > 1 else 1 << 30
    which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["2k", "10k", "15k", "dfa", "sent"]
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    per_gib = n / float(1 << 30)
    if "2k" in which:
        run(dev, "clamav2k", clamav_pats(2000), n, [1, 2], plants=int(4096 * per_gib * 32), iters=iters)
    if "2k-s4" in which:
        run(dev, "clamav2k", clamav_pats(2000), n, [1], plants=int(4096 * per_gib * 32), iters=iters)
    if "10k" in which:
        run(dev, "clamav10k", clamav_pats(10000), n, [1, 2], plants=int(100000 * per_gib), iters=iters)
    if "10k-s4" in which:
        run(dev, "clamav10k", clamav_pats(10000), n, [1], plants=int(100000 * per_gib), iters=iters)
    if "10k-noplant" in which:
        run(dev, "clamav10k-noplant", clamav_pats(10000), n, [1, 2], plants=0, iters=iters)
    if "10k-mixed" in which:
        # a mixed set: ClamAV 10k plus a few short patterns -> sampled filter + start-filter pass (mode 1)
        short = [(b"MZ", 20000), (b"\x7fEL", 20001), (b"PE\x00\x00", 20002), (b"%PDF-", 20003), (b"virus!", 20004),
                 (b"\xe8\x00\x00\x00\x00\x5d\x81", 20005), (b"\x90\x90\x90\x90\x90\x90\x90\x90\x90", 20006)]
        run(dev, "clamav10k+7short", clamav_pats(10000) + short, n, [1, 2], plants=int(100000 * per_gib), iters=iters)
    if "15k" in which:
        run(dev, "clamav15k", clamav_pats(15000), n, [1, 2], plants=int(100000 * per_gib), iters=iters)
    if "dfa" in which:
        run(dev, "clamav10k-dfa", clamav_pats(10000), n >> 3, [3], plants=int(12500 * per_gib), iters=iters)
    if "sent" in which:
        words = [l.split(b"\t")[0] for l in read_fixture("english_top5000.txt.gz").split(b"\n") if l]
        text = synth.english_like(words, 16 << 20, seed=4)
        run(dev, "sentiment", load_patterns("sentiment_categorical.pat.gz"), n >> 2, [4, 2, 3], text=text,
            iters=iters)
