"""One process per GPU: contiguous byte-range shards with a leading halo, a replicated
automaton, and torch.distributed (NCCL on GPUs, gloo in the CPU tests) used only for the
small exchange that follows the scan: an all-gather of the per-rank match counts and a
gather of the sorted key lists to rank 0.

Because every rank keeps only matches that END inside its own range and sorts locally, the
global canonical list is the concatenation of the per-rank lists in rank order: no merge.
Nothing here touches the scan itself; the per-rank scan is libacmatch_b200.so.
"""
import numpy as np
import torch
import torch.distributed as dist

KEY_PAT_BITS = 24
KEY_PAT_MASK = (1 << KEY_PAT_BITS) - 1


def shard_bounds(total_bytes, world, rank, align=16):
    """[lo, hi) of rank's shard; cuts are multiples of `align` (device buffers are 16-byte
    aligned and the scan windows are aligned to the buffer)."""
    def cut(r):
        if r >= world:
            return total_bytes
        return (total_bytes * r // world) // align * align
    return cut(rank), cut(rank + 1)


def shard_window(total_bytes, world, rank, max_pattern_len, align=16):
    """(read_lo, lo, hi): the shard reads [read_lo, hi) and keeps matches ending in [lo, hi).
    read_lo is `lo` minus the halo of Lmax-1 bytes, rounded down to `align`."""
    lo, hi = shard_bounds(total_bytes, world, rank, align)
    halo = max(0, max_pattern_len - 1)
    read_lo = max(0, lo - halo) // align * align
    return read_lo, lo, hi


def exchange_counts(n_local, device):
    """All-gather of the per-rank match counts -> python list (one small collective)."""
    world = dist.get_world_size()
    mine = torch.tensor([int(n_local)], dtype=torch.int64, device=device)
    if device.type == "cuda":
        allc = torch.empty(world, dtype=torch.int64, device=device)
        dist.all_gather_into_tensor(allc, mine)
        return allc.tolist()                   # one synchronisation
    outs = [torch.empty(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(outs, mine)
    return [int(x.item()) for x in outs]


def gather_keys(keys, counts, dst=0):
    """Variable-length gather of the per-rank sorted key tensors (int64, already offset to
    global positions) to rank `dst`.  Returns the concatenated tensor on dst, None elsewhere."""
    rank, world = dist.get_rank(), dist.get_world_size()
    if rank != dst:
        if counts[rank]:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, keys[:counts[rank]].contiguous(), dst)]):
                w.wait()
        return None
    out = torch.empty(sum(counts), dtype=torch.int64, device=keys.device)
    ops, pos = [], 0
    for r in range(world):
        n = counts[r]
        if n:
            if r == dst:
                out[pos:pos + n] = keys[:n]
            else:
                ops.append(dist.P2POp(dist.irecv, out[pos:pos + n], r))
        pos += n
    if ops:
        for w in dist.batch_isend_irecv(ops):      # one grouped NCCL launch
            w.wait()
    return out


def allreduce_histogram(hist):
    """Per-pattern counts over all shards (int64 tensor, summed in place)."""
    dist.all_reduce(hist, op=dist.ReduceOp.SUM)
    return hist


def unpack_keys(keys):
    """int64 key tensor / array -> (end offsets u64, pattern indices u32) numpy arrays."""
    k = keys.cpu().numpy().astype(np.uint64) if isinstance(keys, torch.Tensor) else \
        np.asarray(keys, dtype=np.uint64)
    return k >> np.uint64(KEY_PAT_BITS), (k & np.uint64(KEY_PAT_MASK)).astype(np.uint32)


def pack_keys(off, pat):
    return (np.asarray(off, dtype=np.uint64) << np.uint64(KEY_PAT_BITS)) | np.asarray(pat, dtype=np.uint64)


class ShardedScan:
    """Per-rank state for scanning one shard of a stream that lives in this rank's HBM.

    data_ptr / n describe the device buffer holding stream bytes [read_lo, hi)."""

    def __init__(self, device, automaton, total_bytes, max_pattern_len, scanner_kwargs=None):
        from .matcher import Scanner
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.read_lo, self.lo, self.hi = shard_window(total_bytes, self.world, self.rank, max_pattern_len)
        self.scanner = Scanner(device, automaton, max(1, self.hi - self.lo), **(scanner_kwargs or {}))
        self.device = device
        self.tdev = torch.device("cuda", device.ordinal)

    def scan(self, data_ptr):
        """Scan the shard; returns the ScanResult (matches stay on the device)."""
        n = self.hi - self.read_lo
        return self.scanner.scan_device(data_ptr, n, emit_lo=self.lo - self.read_lo, emit_hi=n)

    def local_keys(self, res):
        """The sorted keys of the last scan as an int64 torch view shifted to stream offsets."""
        n = int(res.n_matches)
        if n == 0:
            return torch.empty(0, dtype=torch.int64, device=self.tdev)
        ptr = self.scanner.keys_ptr()
        arr = _as_tensor(ptr, n, self.tdev)
        return arr + (self.read_lo << KEY_PAT_BITS)

    def gather(self, res, dst=0):
        """Count exchange + key gather.  Returns (offsets, patterns) numpy arrays on dst."""
        counts = exchange_counts(res.n_matches, self.tdev)
        out = gather_keys(self.local_keys(res), counts, dst)
        if out is None:
            return None
        return unpack_keys(out)


class PeerGather:
    """Single-node gather of the per-rank sorted key lists into rank 0's HBM without a
    collective on the critical path: counts travel through a few words of POSIX shared
    memory, keys are stored by each rank's own kernel straight into rank 0's buffer through
    a CUDA IPC mapping (NVLink peer stores), rank 0 copies the result to pinned host memory
    on a side stream so the copy overlaps the next scan.

    Step protocol (step ids 1, 2, ...; counts and gather buffers double-buffered by parity):
      every rank   counts[step & 1][r] = n_r ; gen_counts[r] = step
      rank r       waits for all gen_counts >= step and consumed >= step - 2, pushes its keys
                   into buffer step & 1 at offset sum(counts[:r]), synchronises, done[r] = step
      rank 0       waits for all done >= step; waits for the D2H of step - 1, consumed = step - 1;
                   starts the D2H of this step on the side stream
    gather() on rank 0 returns the PREVIOUS step's list (None for the first); flush() returns
    the last one.
    """
    ROW_CNT0, ROW_CNT1, ROW_GEN, ROW_DONE, ROW_CONSUMED = range(5)

    def __init__(self, device, cap_keys, timeout_s=60.0):
        import ctypes as C
        import os
        import tempfile
        from ._lib import check, lib
        self.C, self.L, self.check = C, lib(), check
        self.device = device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.cap = int(cap_keys)
        self.timeout = timeout_s
        self.step = 0
        self.pending = None          # (buffer parity, total) whose D2H is in flight
        box = [None, None]
        if self.rank == 0:
            fd, path = tempfile.mkstemp(prefix="acm_gather_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            os.ftruncate(fd, 8 * 5 * self.world + 64)
            os.close(fd)
            self.buf = device.alloc(2 * self.cap * 8)
            h = (C.c_ubyte * 64)()
            check(self.L.acm_ipc_export(device.handle, C.c_void_p(self.buf), h), "acm_ipc_export")
            box = [path, bytes(h)]
        dist.broadcast_object_list(box, src=0)
        self.path = box[0]
        self.shm = np.memmap(self.path, dtype=np.int64, mode="r+", shape=(5, self.world))
        if self.rank == 0:
            self.shm[:] = 0
            self.shm.flush()
            self.dst = self.buf
            from .matcher import pinned_empty
            self.host_bytes, self._owner = pinned_empty(2 * self.cap * 8)
            self.host = self.host_bytes.view(np.uint64)
        else:
            p = C.c_void_p()
            h = (C.c_ubyte * 64).from_buffer_copy(box[1])
            check(self.L.acm_ipc_open(device.handle, h, C.byref(p)), "acm_ipc_open")
            self.dst = p.value
        dist.barrier()

    def _wait(self, row, value, ranks):
        import time
        t0 = time.perf_counter()
        while True:
            if all(int(self.shm[row, r]) >= value for r in ranks):
                return
            if time.perf_counter() - t0 > self.timeout:
                raise RuntimeError(f"PeerGather: rank {self.rank} timed out waiting on row {row} >= {value}")

    def _collect(self):
        """rank 0: finish the D2H in flight, release its buffer, return its keys (a view)."""
        if self.pending is None:
            return None
        par, total, st = self.pending
        self.check(self.L.acm_side_sync(self.device.handle), "acm_side_sync")
        self.shm[self.ROW_CONSUMED, 0] = st
        self.pending = None
        return self.host[par * self.cap: par * self.cap + total]

    def gather(self, scanner, n_local, key_add):
        """One step.  Returns (keys of the previous step or None, total of THIS step) on rank 0,
        (None, total) elsewhere; raises if the list exceeds the buffer."""
        C = self.C
        self.step += 1
        st, par = self.step, self.step & 1
        everyone = range(self.world)
        self.shm[par, self.rank] = int(n_local)
        self.shm[self.ROW_GEN, self.rank] = st
        self._wait(self.ROW_GEN, st, everyone)
        counts = [int(self.shm[par, r]) for r in everyone]
        total = sum(counts)
        if total > self.cap:
            raise RuntimeError(f"PeerGather: {total} keys exceed the gather buffer ({self.cap})")
        self._wait(self.ROW_CONSUMED, st - 2, [0])
        if n_local:
            self.check(self.L.acm_scan_push_keys(scanner._h, C.c_void_p(self.dst),
                                                 par * self.cap + sum(counts[:self.rank]), key_add),
                       "acm_scan_push_keys")
        self.device.sync()
        self.shm[self.ROW_DONE, self.rank] = st
        if self.rank != 0:
            return None, total
        self._wait(self.ROW_DONE, st, everyone)
        prev = self._collect()
        if total:
            self.check(self.L.acm_memcpy_d2h_side(
                self.device.handle, C.c_void_p(self.host.ctypes.data + par * self.cap * 8),
                C.c_void_p(self.buf + par * self.cap * 8), total * 8), "acm_memcpy_d2h_side")
        self.pending = (par, total, st)
        return prev, total

    def flush(self):
        """rank 0: the last step's keys (waits for its D2H); None elsewhere."""
        return self._collect() if self.rank == 0 else None

    def close(self):
        import os
        try:
            self.flush()
            dist.barrier()
            if self.rank == 0:
                self.device.free(self.buf)
                os.unlink(self.path)
            else:
                self.L.acm_ipc_close(self.device.handle, self.C.c_void_p(self.dst))
        except Exception:
            pass


class StepPipeline:
    """The multi-GPU step with no host round trip on the critical path.

    Each rank owns two scanners and queues whole steps on its stream: scan -> prefix sum ->
    compaction + sort -> push of the sorted keys (shifted to stream positions) into THIS rank's
    fixed region of rank 0's gather buffer (CUDA IPC mapping: NVLink peer stores; local memory on
    rank 0 itself).  No rank needs another rank's count before it pushes, so nothing waits on a
    peer; the host finishes step i-1 (32-byte status readback, already complete or nearly) while
    the GPU runs step i.  Counts and progress travel through a few words of POSIX shared memory:

        cnt[s % DEPTH][r]   match count of rank r in step s
        done[r]             last step whose keys are complete in rank 0's HBM
        consumed            last step rank 0 has copied to the host (regions of step s are
                            reused by step s + DEPTH, so submit(s + DEPTH) waits for it)

    Rank 0 collects step s as soon as every done[r] >= s: one D2H per rank region on the side
    stream, packed in rank order into pinned host memory -- the global sorted list (ranges are
    disjoint and ordered).  world == 1 needs no torch.distributed.

    Usage per rank:  submit(step 1); submit(2); complete() -> step 1; submit(3); complete() -> 2 ...

    lazy_keys: rank 0's copy of a step's gathered list to the host is issued by complete() but not
    waited for -- the keys it returns are valid after sync_keys() or after the NEXT complete() has
    returned.  At 8 GPUs the copy of a step is 8 regions, ~130 us of a 200-us step: waiting for it
    inside complete() made rank 0's host thread the slowest part of the job.
    """
    DEPTH = 4

    def __init__(self, device, automaton, max_bytes, cap_keys, rank=0, world=1, scanner_kwargs=None,
                 timeout_s=60.0, lazy_keys=False):
        import ctypes as C
        import os
        import tempfile
        from ._lib import check, lib
        from .matcher import Scanner, pinned_empty
        self.C, self.L, self.check = C, lib(), check
        self.device, self.rank, self.world = device, rank, world
        self.cap = int(cap_keys)
        self.timeout = timeout_s
        self.scanners = [Scanner(device, automaton, max_bytes, **(scanner_kwargs or {})) for _ in range(2)]
        self.launched = self.finished = 0
        self.lazy_keys = bool(lazy_keys)
        self.results = {}
        rows = self.DEPTH + 2
        box = [None, None]
        if rank == 0:
            fd, path = tempfile.mkstemp(prefix="acm_steps_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            os.ftruncate(fd, 8 * rows * world)
            os.close(fd)
            self.buf = device.alloc(self.DEPTH * world * self.cap * 8)
            h = (C.c_ubyte * 64)()
            if world > 1:
                check(self.L.acm_ipc_export(device.handle, C.c_void_p(self.buf), h), "acm_ipc_export")
            box = [path, bytes(h)]
        if world > 1:
            dist.broadcast_object_list(box, src=0)
        self.path = box[0]
        self.shm = np.memmap(self.path, dtype=np.int64, mode="r+", shape=(rows, world))
        self.ROW_DONE, self.ROW_CONSUMED = self.DEPTH, self.DEPTH + 1
        if rank == 0:
            self.shm[:] = 0
            self.dst = self.buf
            self.host_bytes, self._owner = pinned_empty(2 * world * self.cap * 8)
            self.host = self.host_bytes.view(np.uint64)
            self._srcs = (C.c_void_p * world)()
            self._lens = (C.c_uint64 * world)()
        else:
            p = C.c_void_p()
            h = (C.c_ubyte * 64).from_buffer_copy(box[1])
            check(self.L.acm_ipc_open(device.handle, h, C.byref(p)), "acm_ipc_open")
            self.dst = p.value
        if world > 1:
            dist.barrier()

    def _region(self, step, r):
        return self.dst + ((step % self.DEPTH) * self.world + r) * self.cap * 8

    def _wait(self, row, value, ranks):
        import time
        shm = self.shm
        t0 = None
        while True:
            if all(shm[row, r] >= value for r in ranks):
                return
            if t0 is None:
                t0 = time.perf_counter()
            elif time.perf_counter() - t0 > self.timeout:
                raise RuntimeError(f"StepPipeline: rank {self.rank} timed out waiting on row {row} >= {value}")

    def submit(self, d_ptr, n, emit_lo, emit_hi, key_add):
        """Queue the next step on this rank's shard; returns its step id."""
        if self.launched - self.finished >= 2:
            raise RuntimeError("StepPipeline: complete() a step before submitting a third")
        s = self.launched + 1
        if s > self.DEPTH:
            self._wait(self.ROW_CONSUMED, s - self.DEPTH, [0])
        self.scanners[s & 1].scan_async(d_ptr, n, emit_lo, emit_hi, push=(self._region(s, self.rank), self.cap, key_add))
        self.launched = s
        return s

    def stream_of(self, step):
        """cudaStream_t (int) the kernels of `step` were queued on."""
        return self.scanners[step & 1].stream()

    def complete(self):
        """Finish the oldest queued step.  Returns (ScanResult of this rank, total matches of the
        step over all ranks or None, keys) -- total and keys (a uint64 view of pinned memory,
        valid until the step after next completes) on rank 0 only."""
        if self.finished >= self.launched:
            raise RuntimeError("StepPipeline: nothing to complete")
        s = self.finished + 1
        res = self.scanners[s & 1].finish()
        self.finished = s
        self.shm[s % self.DEPTH, self.rank] = int(res.n_matches)
        self.shm[self.ROW_DONE, self.rank] = s
        if self.rank != 0:
            return res, None, None
        everyone = range(self.world)
        self._wait(self.ROW_DONE, s, everyone)
        total = 0
        for r in everyone:
            c = int(self.shm[s % self.DEPTH, r])
            self._srcs[r] = self._region(s, r)
            self._lens[r] = c * 8
            total += c
        hbase = (s & 1) * self.world * self.cap
        if self.lazy_keys:
            # the copies of step s - 1 were issued a whole step ago: this wait is over at once, and its
            # regions may be overwritten from now on
            self.check(self.L.acm_side_sync(self.device.handle), "acm_side_sync")
            self.shm[self.ROW_CONSUMED, 0] = s - 1
            self.check(self.L.acm_memcpy_d2h_segments_async(
                self.device.handle, self.C.c_void_p(self.host.ctypes.data + hbase * 8), self._srcs, self._lens,
                self.world), "acm_memcpy_d2h_segments_async")
            return res, total, self.host[hbase:hbase + total]
        self.check(self.L.acm_memcpy_d2h_segments(
            self.device.handle, self.C.c_void_p(self.host.ctypes.data + hbase * 8), self._srcs, self._lens,
            self.world), "acm_memcpy_d2h_segments")
        self.shm[self.ROW_CONSUMED, 0] = s
        return res, total, self.host[hbase:hbase + total]

    def sync_keys(self):
        """lazy_keys: wait for the host copy of every completed step's list (rank 0; no-op elsewhere)."""
        if self.rank == 0 and self.lazy_keys:
            self.check(self.L.acm_side_sync(self.device.handle), "acm_side_sync")
            self.shm[self.ROW_CONSUMED, 0] = self.finished

    def close(self):
        import os
        try:
            while self.finished < self.launched:
                self.complete()
            self.sync_keys()
            if self.world > 1:
                dist.barrier()
            for sc in self.scanners:
                sc.close()
            if self.rank == 0:
                self.device.free(self.buf)
                os.unlink(self.path)
            elif self.dst:
                self.L.acm_ipc_close(self.device.handle, self.C.c_void_p(self.dst))
                self.dst = None
        except Exception:
            pass


class _CudaArrayView:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False),
                                         "version": 3, "strides": None}


def _as_tensor(ptr, n, tdev):
    return torch.as_tensor(_CudaArrayView(ptr, n), device=tdev)
