"""The N > 1 host path on CPU: world_size-2 and -3 gloo groups run the shard / halo / count
exchange / gather protocol of gpu_pattern_matching_b200.sharded with the CPU oracle standing
in for the per-rank device scan; rank 0's gathered list must equal the whole-stream walk."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpu_pattern_matching_b200 import sharded, synth
    from helpers import build_oracle, clamav_pats

    pats = clamav_pats(2000)
    o = build_oracle(pats)
    lmax = o.max_pattern_len
    # every rank can regenerate any part of the stream: it is a pure function of the offset
    cuts = [sharded.shard_bounds(total, world, r)[0] for r in range(1, world)]
    forced = [(c - 17, 30 + k) for k, c in enumerate(cuts)] + [(c - 1, 60 + k) for k, c in enumerate(cuts)]
    plants = synth.Plants([p for p, _ in pats], total, 64, 5, forced)
    read_lo, lo, hi = sharded.shard_window(total, world, rank, lmax)
    assert lo % 16 == 0 and read_lo % 16 == 0 and read_lo <= lo
    buf = synth.stream(hi - read_lo, 5, read_lo)
    plants.apply_host(buf, read_lo)
    off, pat, _, _ = o.search(buf, emit_from=lo - read_lo, base=read_lo)
    keys = torch.from_numpy(sharded.pack_keys(off, pat).astype(np.int64))
    cpu = torch.device("cpu")
    counts = sharded.exchange_counts(off.size, cpu)
    assert counts[rank] == off.size and len(counts) == world
    out = sharded.gather_keys(keys, counts, 0)
    hist = torch.from_numpy(np.bincount(pat, minlength=len(pats)).astype(np.int64))
    sharded.allreduce_histogram(hist)
    if rank == 0:
        whole = synth.stream(total, 5)
        plants.apply_host(whole)
        eo, ep, _, _ = o.search(whole)
        goff, gpat = sharded.unpack_keys(out)
        ok = np.array_equal(goff, eo) and np.array_equal(gpat, ep) and \
            np.array_equal(hist.numpy(), np.bincount(ep, minlength=len(pats)))
        open(os.path.join(outdir, "result"), "w").write(f"{int(ok)} {eo.size} {sum(counts)}")
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gather_equals_serial_walk(world, tmp_path):
    total = (1 << 20) + 4096 + 7           # not a multiple of anything convenient
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    ok, n_expected, n_gathered = open(tmp_path / "result").read().split()
    assert ok == "1" and n_expected == n_gathered and int(n_expected) >= 64


def test_shard_bounds_cover_the_stream():
    from gpu_pattern_matching_b200 import sharded
    for total in (1, 15, 16, 1000, (1 << 30) + 5):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = sharded.shard_bounds(total, world, r)
                assert lo == prev and lo <= hi and (lo % 16 == 0)
                prev = hi
            assert prev == total
