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


@pytest.mark.parametrize("own_stream", [False, True])
def test_step_pipeline_single_rank(tmp_path, own_stream):
    """world == 1: same pipeline, no torch.distributed, list checked against the oracle.  own_stream:
    every scanner on its own stream -- the scan stage of a step waits for the scan stage of the step
    before it (the other scanner's), whose post-pass then runs under it on the SMs the persistent
    scan kernel leaves free; different input every step, so a stale list would show."""
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
    # a second, different buffer: steps alternate between the two
    buf2, _ = planted_stream(pats, buf.size, seed=10, plants=300)
    eo2, ep2, _, _ = o.search(buf2)
    d2 = dev.alloc(buf2.size + 64)
    dev.h2d(d2, buf2)
    dev.sync()
    # own streams also run with the host copy of a step's list left in flight (lazy_keys): a list is
    # read only after the next complete() / sync_keys()
    pipe = sharded.StepPipeline(dev, a.automaton, buf.size, 1 << 14, scanner_kwargs={"own_stream": own_stream},
                                lazy_keys=own_stream)
    got, pending = [], None

    def take(out):
        nonlocal pending
        if pending is not None:                 # the previous step's copy is complete now
            res, tot, keys = pending
            got.append((res, tot, np.array(keys, copy=True)))
        pending = out

    for it in range(8):
        pipe.submit(d2 if it % 2 else d, buf.size, 0, buf.size, 0)
        if it > 0:
            take(pipe.complete())
    take(pipe.complete())
    pipe.sync_keys()
    take(None)
    assert len(got) == 8
    for it, (res, tot, keys) in enumerate(got):
        goff, gpat = sharded.unpack_keys(keys)
        wo, wp = (eo2, ep2) if it % 2 else (eo, ep)
        assert tot == wo.size and np.array_equal(goff, wo) and np.array_equal(gpat, wp), f"step {it}"
    pipe.close()
    dev.free(d)
    dev.free(d2)
    dev.close()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_gather_two_ranks(world, tmp_path):
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    total = (3 << 20) + 4096 + 7
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    ok, n = open(tmp_path / "result").read().split()
    assert ok == "1" and int(n) >= 256


def test_multi_device_c_api_in_one_process():
    """acm_multi_*: one host thread per device, no torch.distributed.  Devices are distinct GPUs when
    the box has them, else the same GPU three times (separate streams and buffers).  Device-resident
    shards and the host-stream form, both against the whole-stream oracle walk; plants across every cut."""
    if not torch.cuda.is_available():
        pytest.fail("gpu test selected but no CUDA device is visible")
    import ctypes as C
    import gpu_pattern_matching_b200 as g
    from gpu_pattern_matching_b200 import _lib, sharded, synth
    from helpers import build_oracle, build_product, clamav_pats
    L = g.lib()
    pats = clamav_pats(2000)
    o, a = build_oracle(pats), build_product(pats, upload=False)
    ngpu = torch.cuda.device_count()
    for ndev in (1, 2, 3):
        ords = (C.c_int * ndev)(*[i % ngpu for i in range(ndev)])
        total = (5 << 20) + 16 * 7 + 3
        m = C.c_void_p()
        p = _lib.ScanParams()
        _lib.check(L.acm_multi_open(L.acsm_tables(a._p), ords, ndev, total // ndev + 4096, C.byref(p), C.byref(m)),
                   "acm_multi_open")
        assert L.acm_multi_devices(m) == ndev
        cuts = []
        rl, lo, hi = C.c_uint64(), C.c_uint64(), C.c_uint64()
        for gdev in range(ndev):
            L.acm_multi_shard(m, total, gdev, C.byref(rl), C.byref(lo), C.byref(hi))
            cuts.append((rl.value, lo.value, hi.value))
        assert cuts[0][1] == 0 and cuts[-1][2] == total and all(c[2] == d[1] for c, d in zip(cuts, cuts[1:]))
        forced = [(c[1] - 300, 40 + k) for k, c in enumerate(cuts[1:])] + [(c[1] - 4, 70 + k) for k, c in enumerate(cuts[1:])]
        plants = synth.Plants([p_ for p_, _ in pats], total, 400, 11, forced)
        whole = synth.stream(total, 11)
        plants.apply_host(whole)
        eo, ep, _, _ = o.search(whole)
        # device-resident shards
        ptrs = (C.c_void_p * ndev)()
        devs = []
        for gdev, (read_lo, lo_, hi_) in enumerate(cuts):
            d = g.Device.__new__(g.Device)
            d.L, d._h, d.ordinal = L, C.c_void_p(L.acm_multi_device(m, gdev)), ords[gdev]
            n = hi_ - read_lo
            ptr = d.alloc(n + 64)
            d.h2d(ptr, whole[read_lo:hi_])
            ptrs[gdev] = ptr
            devs.append((d, ptr))
        cap = int(eo.size) + 16
        keys, owner = g.matcher.pinned_empty(cap * 8)
        keys = keys.view(np.uint64)
        counts = (C.c_uint64 * ndev)()
        res = _lib.ScanResult()
        for _ in range(3):                                   # the workers and their buffers are reused
            keys[:] = 0
            tot = L.acm_multi_scan_device(m, ptrs, total, keys.ctypes.data_as(_lib.u64p), cap, counts, C.byref(res))
            assert tot == eo.size == sum(counts), (tot, eo.size, list(counts))
            goff, gpat = sharded.unpack_keys(keys[:tot].copy())
            assert np.array_equal(goff, eo) and np.array_equal(gpat, ep)
        # a buffer that is too small: the total is still reported, nothing is written beyond cap
        keys[:] = 0
        tot = L.acm_multi_scan_device(m, ptrs, total, keys.ctypes.data_as(_lib.u64p), 10, counts, C.byref(res))
        assert tot == eo.size and not keys[10:].any()
        # host stream
        host, owner2 = g.matcher.pinned_empty(total + 64)
        host[:total] = whole
        off = np.empty(cap, dtype=np.uint64)
        pat = np.empty(cap, dtype=np.uint32)
        tot = L.acm_multi_scan_host(m, C.c_void_p(host.ctypes.data), total, 1000, off.ctypes.data_as(_lib.u64p),
                                    pat.ctypes.data_as(_lib.u32p), cap, C.byref(res))
        assert tot == eo.size
        assert np.array_equal(off[:tot], eo + np.uint64(1000)) and np.array_equal(pat[:tot], ep)
        for d, ptr in devs:
            d.free(ptr)
        L.acm_multi_close(m)
        del owner, owner2
