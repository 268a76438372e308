"""Scratch: time K1 on the shard a given (rank, world) would hold, on one GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gpu_pattern_matching_b200 as g
from gpu_pattern_matching_b200 import sharded, synth
from oracle_lib import clamav_signatures
rank, world = int(sys.argv[1]), int(sys.argv[2])
nsig = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
GIB = 1 << 30
sigs = clamav_signatures(nsig)
a = g.Acsm()
for i, s in enumerate(sigs): a.add_pattern(s, i)
a.compile()
dev = g.Device(0)
a.gen_state_table(0, dev.handle, None)
total = GIB * world
read_lo, lo, hi = sharded.shard_window(total, world, rank, a.get_max_pattern_size())
n = hi - read_lo
d = dev.alloc(n + 64)
dev.synth_fill(d, (n + 7) // 8 * 8, 2, read_lo)
pl = synth.Plants(sigs, total, 100000 * world, 2)
dev.plant(d, n, read_lo, pl)
dev.sync()
sc = g.Scanner(dev, a.automaton, hi - lo + 4096, timing=True)
ts = []
for it in range(6):
    r = sc.scan_device(d, n, lo - read_lo, n)
    ts.append(r.ms_scan)
print(f"rank {rank}/{world}: matches {r.n_matches} K1 ms min {min(ts):.3f} med {sorted(ts)[3]:.3f} total {r.ms_total:.3f}")
# which planted patterns fall in this shard, and how often the heavy ones
inshard = (pl.pos >= lo) & (pl.pos < hi)
ids, cnt = np.unique(pl.pid[inshard], return_counts=True)
print("plants in shard", int(inshard.sum()), "distinct", ids.size, "max repeats", cnt.max())
# experiments
for (el, label) in ((192, "same buffer, emit_lo=192"), (0, "same buffer, emit_lo=0"), (4096, "emit_lo=4096")):
    ts = []
    for it in range(4):
        r = sc.scan_device(d, n, el, n)
        ts.append(r.ms_scan)
    print(f"   {label}: matches {r.n_matches} K1 min {min(ts):.3f}")
