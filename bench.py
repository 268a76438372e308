#!/usr/bin/env python
"""bench.py -- GB/s scanned by the B200 Aho-Corasick hot path, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one pass of the hot path (scan -> prefix sum -> compaction + sort -> match count;
at N > 1 also the count exchange and the gather of the sorted lists to rank 0) over one batch
of synthetic input.  Workload per GPU (BASELINE.json configs[1]): ClamAV 10000 signatures over
a 1 GiB seeded random byte stream with ~10^5 signatures planted (pure random
bytes contain no match, so parity and emission would be vacuous).  At N GPUs the stream is
N GiB, rank r holds bytes [r, r+1) GiB plus a leading halo of Lmax-1 bytes (weak scaling).

  value      whole-job GB/s with the input already resident in HBM
  e2e        the same stream scanned through the C ABI from a pinned HOST buffer
             (acm_scan_host: chunked H2D overlapped with the scan, D2H of the match list)
  roofline   the scan stage alone (CUDA events inside the library around the streaming filter
             kernel and the kernel that resolves its survivors) against the measured HBM copy
             bandwidth in MEASURED_PEAKS.json; algorithmic traffic = 1 B per input byte
  cpu_baseline  the reference's CPU path (oracle/_ref when present, else the oracle port) on a
             bounded sample of the same stream, all host cores

  clocks     nvidia-smi SM clock / throttle reasons, sampled every 50 ms while the timed steps and
             ~0.6 s of identical untimed steps that follow them run (K steps last milliseconds)

Defaults: N = 1, 50 timed steps after 5 warm-up steps (about 15 s in all, most of it the CPU baseline).
--impl reference times that CPU path as the arm itself.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GIB = 1 << 30
SEED = 2
PLANTS_PER_GIB = 100000
SIGS = 10000
METRIC = "GB/s scanned"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def plants_for(sigs, total_bytes, seed):
    from gpu_pattern_matching_b200 import synth
    return synth.Plants(sigs, total_bytes, int(PLANTS_PER_GIB * total_bytes / GIB), seed)


def cpu_reference_rate(sample_bytes, threads, steps=1, warmup=0):
    """GB/s of the reference's CPU walk (all threads) over the first sample_bytes of the
    workload stream.  Returns (GB/s, kind, matches, seconds per step)."""
    from gpu_pattern_matching_b200 import synth
    from oracle_lib import Oracle, RefAcsm, clamav_signatures, ref_available
    sigs = clamav_signatures(SIGS)
    buf = synth.stream(sample_bytes, SEED)
    plants_for(sigs, GIB, SEED).apply_host(buf)       # same plants as the GPU arm's first GiB
    if ref_available():
        kind, m = "reference", RefAcsm()
    else:
        kind, m = "port", Oracle(256)
    for i, s in enumerate(sigs):
        m.add(s, i)
    m.compile()
    times, found = [], 0
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        found = m.walk_count_mt(buf, threads)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    # the same walk on ONE thread over the first 64 MiB (per-core figure)
    one = buf[:min(sample_bytes, 64 << 20)]
    t0 = time.perf_counter()
    m.walk_count_mt(one, 1)
    cpu_reference_rate.single_thread_gbs = one.size / (time.perf_counter() - t0) / 1e9
    return sample_bytes / sec / 1e9, kind, int(found), sec


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path, host cores only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # bounded sample: ~8 s of wall time per step at ~60 MB/s per core, at most the whole GiB
    sample = int(min(GIB, max(32 << 20, cores * 60e6 * 8 / max(1, args.steps + args.warmup) // (1 << 20) * (1 << 20))))
    rate, kind, found, sec = cpu_reference_rate(sample, cores, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"clamav{SIGS} x seeded random stream, {PLANTS_PER_GIB} planted signatures/GiB",
                   "sample_bytes": sample, "host_threads": cores,
                   "note": "serial DFA walk over the reference-layout int32[states][512] table, "
                           "pthread-sharded with Lmax-1 halo (oracle/_ref = reference acsmx.c compiled here)"
                           if kind == "reference" else "oracle port of the reference walk (oracle/acsm_oracle.c)"},
        "cpu_baseline": {"value": rate, "unit": "GB/s", "cores": cores, "kind": kind,
                         "sample": f"first {sample >> 20} MiB of the workload stream, {found} matches",
                         "single_thread_gbs": cpu_reference_rate.single_thread_gbs},
        "e2e": {"value": rate, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_REAL_STDOUT = None


def guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries (NCCL's version banner, ...) write to
    fd 1 behind Python's back: point fd 1 at stderr for the whole run and keep the real stdout
    for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global SIGS
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes-per-gpu", type=int, default=GIB)
    ap.add_argument("--sigs", type=int, default=SIGS, choices=[2000, 10000, 15000],
                    help="ClamAV signature set (default 10000 = BASELINE configs[1]; 15000 with "
                         "--bytes-per-gpu 4294967296 is configs[2])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    SIGS = args.sigs

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import gpu_pattern_matching_b200 as g
    from gpu_pattern_matching_b200 import sharded
    from oracle_lib import clamav_signatures      # fixture reader only (signature list)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    tdev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=tdev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"

    # ---- automaton (replicated) ----
    sigs = clamav_signatures(SIGS)
    acsm = g.Acsm()
    for i, s in enumerate(sigs):
        acsm.add_pattern(s, i)
    acsm.compile()
    stream = torch.cuda.current_stream()
    dev = g.Device(local_rank, stream=stream.cuda_stream)
    conf_ctx = dev.handle
    acsm.gen_state_table(0, conf_ctx, None)
    lmax = acsm.get_max_pattern_size()

    # ---- this rank's shard of the world*bytes stream, generated in HBM ----
    per = args.bytes_per_gpu
    total = per * world
    read_lo, lo, hi = sharded.shard_window(total, world, rank, lmax)
    n = hi - read_lo
    data = torch.empty(n + 64, dtype=torch.uint8, device=tdev)
    dev.synth_fill(data.data_ptr(), (n + 7) // 8 * 8, SEED, read_lo)
    plants = plants_for(sigs, total, SEED)
    dev.plant(data.data_ptr(), n, read_lo, plants)
    dev.sync()

    emit_lo = lo - read_lo
    key_add = read_lo << sharded.KEY_PAT_BITS
    use_nccl = world > 1 and os.environ.get("BENCH_GATHER", "peer") == "nccl"
    k1_ms, stats = [], {"launches": 0, "fallback": 0, "matches": 0, "mode": 0, "list_bytes": 0}

    if not use_nccl:
        # default: pipelined steps, keys pushed into rank 0's HBM by the step itself (NVLink IPC
        # stores at N > 1), D2H of the gathered list on rank 0's side stream, every step
        pipe = sharded.StepPipeline(dev, acsm.automaton, hi - lo, 1 << 21, rank, world,
                                    scanner_kwargs={"timing": 0 if os.environ.get("BENCH_NO_KERNEL_TIMING") else 2})

        def note(out):
            res, total, keys = out
            k1_ms.append(res.ms_scan)
            stats["launches"] += res.launches
            stats["fallback"] |= res.fallback
            stats["mode"] = res.mode
            if total is not None:
                stats["matches"] = total
                stats["list_bytes"] = int(keys.nbytes)

        def run_steps(k):
            for i in range(k):
                pipe.submit(data.data_ptr(), n, emit_lo, n, key_add)
                if i > 0:
                    note(pipe.complete())
            if k > 0:
                note(pipe.complete())
    else:
        # portable path: NCCL all-gather of the counts + grouped send/recv of the keys, synchronous
        scanner = g.Scanner(dev, acsm.automaton, hi - lo, timing=True)
        pinned_out = torch.empty(1 << 22, dtype=torch.int64).pin_memory() if rank == 0 else None

        def run_steps(k):
            for _ in range(k):
                res = scanner.scan_device(data.data_ptr(), n, emit_lo, n)
                counts = sharded.exchange_counts(res.n_matches, tdev)
                keys = sharded._as_tensor(scanner.keys_ptr(), max(1, int(res.n_matches)), tdev)
                keys = keys[:int(res.n_matches)] + key_add
                out = sharded.gather_keys(keys, counts, 0)
                if rank == 0:
                    pinned_out[:sum(counts)].copy_(out, non_blocking=True)
                    stream.synchronize()
                k1_ms.append(res.ms_scan)
                stats["launches"] += res.launches
                stats["fallback"] |= res.fallback
                stats["mode"] = res.mode
                stats["matches"] = sum(counts)
                stats["list_bytes"] = sum(counts) * 8

    run_steps(args.warmup)
    k1_ms.clear()
    stats["launches"] = 0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    run_steps(args.steps)            # the last complete() has waited for the last step and its D2H
    ev1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    k1_avg = sum(k1_ms) / len(k1_ms)
    launches, matches, fallback, mode = stats["launches"], stats["matches"], stats["fallback"], stats["mode"]
    if world > 1:
        t = torch.tensor([ms, k1_avg], dtype=torch.float64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, k1_avg = float(t[0].item()), float(t[1].item())
        t = torch.tensor([launches], dtype=torch.int64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        launches = int(t.item())
    # K steps of this workload last a few milliseconds, nvidia-smi samples every 50 ms: the same
    # steps keep running, untimed, for ~0.6 s more so that the clock / throttle samples are taken
    # under the load that was just timed (same count on every rank: ms is the max over ranks)
    tail_steps = int(min(20000, max(args.steps, 600.0 / max(ms / args.steps, 1e-3))))
    saved = (list(k1_ms), dict(stats))
    run_steps(tail_steps)
    torch.cuda.synchronize()
    k1_ms[:] = saved[0]
    stats.update(saved[1])
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = (f"the {args.steps} timed steps ({ms:.1f} ms) and {tail_steps} identical untimed steps "
                            f"that follow them, sampled every 50 ms")
    ms_per_step = ms / args.steps
    value = total / (ms_per_step * 1e-3) / 1e9

    # ---- e2e: same shard from pinned host memory through acm_scan_host ----
    e2e = None
    if not args.no_e2e:
        host, owner = g.matcher.pinned_empty(per + 64, dev)      # on the GPU's own NUMA node where allowed
        # identical bytes: copy the kept range of the device shard back once (outside timing)
        g._lib.check(g.lib().acm_memcpy_d2h(dev.handle, C.c_void_p(host.ctypes.data),
                                       C.c_void_p(data.data_ptr() + emit_lo), per), "d2h")
        dev.sync()
        seg = 64 << 20
        hs = g.Scanner(dev, acsm.automaton, seg)
        cap = 1 << 22
        off = np.empty(cap, dtype=np.uint64)
        pat = np.empty(cap, dtype=np.uint32)
        r = g._lib.ScanResult()

        def e2e_step():
            got = g.lib().acm_scan_host(hs._h, C.c_void_p(host.ctypes.data), per, lo,
                                        off.ctypes.data_as(g._lib.u64p), pat.ctypes.data_as(g._lib.u32p),
                                        cap, C.byref(r))
            g._lib.check(got, "acm_scan_host")
            return int(got)
        for _ in range(max(1, args.warmup)):
            e2e_matches = e2e_step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e_steps = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_matches = e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e_steps
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=tdev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        nseg = (per + seg - 1) // seg
        e2e = {"value": total / dt / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": int(per + (nseg - 1) * (lmax - 1)),
               "d2h_bytes_per_step": int(e2e_matches * 8 + nseg * 16),
               "ms_per_step": dt * 1e3, "matches": e2e_matches,
               "api": "acm_scan_host (pinned host buffer, 64 MiB segments, H2D overlapped with the scan)"}
        hs.close()
        del owner

    if not use_nccl:
        pipe.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    achieved = per / (max(k1_avg, 1e-9) * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("k_scan_sampled_dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"clamav{SIGS} x {total / GIB:g} GiB seeded random stream "
                               f"({per / GIB:g} GiB per GPU, halo {lmax - 1} B), {PLANTS_PER_GIB} planted signatures/GiB",
                   "signatures": SIGS, "states": acsm.get_states(), "kernel": g.MODE_NAMES[mode],
                   "bytes_per_gpu": per, "matches": matches, "fallback": int(fallback),
                   "l2_policy": (f"input ({per / GIB:g} GiB per GPU) is larger than L2 (126 MB); no flush needed"
                                 if per > (256 << 20) else
                                 f"WARNING: input ({per >> 20} MiB per GPU) is not much larger than L2 (126 MB)"),
                   "step": ("scan + prefix sum + compaction/sort + NCCL count all-gather + key send/recv to "
                            "rank 0 + D2H of the list" if use_nccl else
                            "scan + prefix sum + compaction/sort + push of the sorted keys into rank 0's gather "
                            "buffer (NVLink IPC stores at N > 1) + D2H of the gathered list to pinned host "
                            "memory, every step; steps are queued two deep (acm_scan_device_async / "
                            "acm_scan_finish), so the host round trip overlaps the next step's scan"),
                   "list_d2h_bytes_per_step": stats["list_bytes"]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic,
                     "kernel": "k_scan_" + g.MODE_NAMES[mode] + (
                         f"<{g.lib().acm_automaton_sample_stride(acsm.automaton)}> + k_resolve_queue "
                         "(scan stage: both launches inside one event pair; the streaming kernel alone "
                         "is ~80 % of it, see profiles/)" if mode == 1 else ""),
                     "kernel_ms": k1_avg, "algorithmic_bytes_per_launch": per, "peak_source": peak_src},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        sample = int(min(GIB, max(32 << 20, int(cores * 60e6 * 8) // (1 << 20) * (1 << 20))))
        rate, kind, found, sec = cpu_reference_rate(sample, cores)
        line["cpu_baseline"] = {"value": rate, "unit": "GB/s", "cores": cores, "kind": kind,
                                "sample": f"first {sample >> 20} MiB of rank 0's stream, "
                                          f"{found} matches, {sec:.2f} s",
                                "single_thread_gbs": cpu_reference_rate.single_thread_gbs}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
