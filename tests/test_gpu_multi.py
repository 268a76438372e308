"""Two ranks (processes), sharded scan + peer gather through CUDA IPC, checked against the
whole-stream oracle walk.  Runs with both ranks on one GPU when only one is visible (the ranks'
kernels never wait on each other; only the hosts spin on shared-memory flags)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
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
    ngpu = torch.cuda.device_count()
    ordinal = rank % ngpu
    torch.cuda.set_device(ordinal)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gpu_pattern_matching_b200 as g
    from gpu_pattern_matching_b200 import sharded, synth
    from helpers import build_oracle, clamav_pats

    pats = clamav_pats(2000)
    dev = g.Device(ordinal)
    a = g.Acsm()
    for p, i in pats:
        a.add_pattern(p, i)
    a.compile()
    a.gen_state_table(0, dev.handle, None)
    lmax = a.get_max_pattern_size()
    cuts = [sharded.shard_bounds(total, world, r)[0] for r in range(1, world)]
    forced = [(c - 400, 30 + k) for k, c in enumerate(cuts)] + [(c - 5, 60 + k) for k, c in enumerate(cuts)]
    plants = synth.Plants([p for p, _ in pats], total, 256, 5, forced)
    read_lo, lo, hi = sharded.shard_window(total, world, rank, lmax)
    n = hi - read_lo
    d = dev.alloc(n + 64)
    dev.synth_fill(d, (n + 7) // 8 * 8, 5, read_lo)
    dev.plant(d, n, read_lo, plants)
    sc = g.Scanner(dev, a.automaton, hi - lo)
    peer = sharded.PeerGather(dev, 1 << 16, timeout_s=30)
    ok = True
    if rank == 0:
        o = build_oracle(pats)
        whole = synth.stream(total, 5)
        plants.apply_host(whole)
        eo, ep, _, _ = o.search(whole)
    for it in range(5):                       # several steps: generation counters, buffer reuse
        res = sc.scan_device(d, n, lo - read_lo, n)
        keys, tot = peer.gather(sc, int(res.n_matches), read_lo << sharded.KEY_PAT_BITS)
        if it == 4:
            keys = peer.flush()               # the list of the last step
        elif it == 0:
            assert keys is None               # gather() hands back the previous step's list
            continue
        if rank == 0:
            goff, gpat = sharded.unpack_keys(np.array(keys, copy=True))
            good = tot == eo.size and np.array_equal(goff, eo) and np.array_equal(gpat, ep)
            if not good:
                bad = np.nonzero(goff[:min(tot, eo.size)] != eo[:min(tot, eo.size)])[0]
                print(f"step {it}: gathered {tot} expected {eo.size}; first offset mismatch at {bad[:3]}: "
                      f"{goff[bad[:3]]} vs {eo[bad[:3]]}", flush=True)
            ok = ok and good
    peer.close()
    # the pipelined form: steps queued two deep, keys pushed by the step itself into per-rank regions
    pipe = sharded.StepPipeline(dev, a.automaton, hi - lo, 1 << 14, rank, world, timeout_s=30)
    outs = []
    for it in range(7):                       # > DEPTH steps: region reuse and the consumed counter
        pipe.submit(d, n, lo - read_lo, n, read_lo << sharded.KEY_PAT_BITS)
        if it > 0:
            outs.append(pipe.complete())
    outs.append(pipe.complete())
    for it, (res, tot, keys) in enumerate(outs):
        if rank == 0:
            goff, gpat = sharded.unpack_keys(np.array(keys, copy=True))
            good = tot == eo.size and np.array_equal(goff, eo) and np.array_equal(gpat, ep)
            if not good:
                print(f"pipeline step {it}: gathered {tot} expected {eo.size}", flush=True)
            ok = ok and good
        else:
            assert tot is None and keys is None
    pipe.close()
    if rank == 0:
        open(os.path.join(outdir, "result"), "w").write(f"{int(ok)} {eo.size}")
    dist.destroy_process_group()


def test_step_pipeline_single_rank(tmp_path):
    """world == 1: same pipeline, no torch.distributed, list checked against the oracle."""
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    import gpu_pattern_matching_b200 as g
    from gpu_pattern_matching_b200 import sharded
    from helpers import build_oracle, build_product, clamav_pats, planted_stream
    pats = clamav_pats(2000)
    o, a = build_oracle(pats), build_product(pats)
    buf, _ = planted_stream(pats, (2 << 20) + 77, seed=9, plants=500)
    eo, ep, _, _ = o.search(buf)
    dev = g.Device(0)
    d = dev.alloc(buf.size + 64)
    dev.h2d(d, buf)
    pipe = sharded.StepPipeline(dev, a.automaton, buf.size, 1 << 14)
    got = []
    for it in range(6):
        pipe.submit(d, buf.size, 0, buf.size, 0)
        if it > 0:
            got.append(pipe.complete())
    got.append(pipe.complete())
    for res, tot, keys in got:
        goff, gpat = sharded.unpack_keys(np.array(keys, copy=True))
        assert tot == eo.size and np.array_equal(goff, eo) and np.array_equal(gpat, ep)
    pipe.close()
    dev.free(d)
    dev.close()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_gather_two_ranks(world, tmp_path):
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    total = (3 << 20) + 4096 + 7
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    ok, n = open(tmp_path / "result").read().split()
    assert ok == "1" and int(n) >= 256
