"""Device, Scanner: the native scan API of include/acm.h from Python.

This is the layer bench.py and the multi-GPU driver use: device-resident streams,
explicit emit windows (halo sharding), pipelined host scans.  The reference-shaped
databuf / worker layer lives in worker.py.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import AcmError, PushTarget, ScanParams, ScanResult, check, lib

MODE_AUTO, MODE_SAMPLED4, MODE_START2, MODE_DFA, MODE_CDFA = 0, 1, 2, 3, 4
MODE_NAMES = {1: "sampled", 2: "start2", 3: "dfa", 4: "cdfa"}
KEY_PAT_BITS = 24


class Device:
    """One GPU (struct acm_device)."""

    def __init__(self, ordinal=0, stream=None):
        self.L = lib()
        h = C.c_void_p()
        check(self.L.acm_device_open(ordinal, C.byref(h)), "acm_device_open")
        self._h = h
        self.ordinal = ordinal
        if stream is not None:
            self.set_stream(stream)

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream):
        """Order all work on an external cudaStream_t (int), e.g. torch's current stream."""
        check(self.L.acm_device_set_stream(self._h, C.c_void_p(cuda_stream)), "acm_device_set_stream")

    def sync(self):
        check(self.L.acm_device_sync(self._h), "acm_device_sync")

    def alloc(self, nbytes):
        p = C.c_void_p()
        check(self.L.acm_dev_alloc(self._h, nbytes, C.byref(p)), "acm_dev_alloc")
        return p.value

    def free(self, ptr):
        self.L.acm_dev_free(self._h, C.c_void_p(ptr))

    def h2d(self, d_ptr, array):
        a = np.ascontiguousarray(array)
        check(self.L.acm_memcpy_h2d(self._h, C.c_void_p(d_ptr), a.ctypes.data_as(C.c_void_p), a.nbytes),
              "acm_memcpy_h2d")
        self.sync()

    def d2h(self, d_ptr, nbytes, dtype=np.uint8):
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        check(self.L.acm_memcpy_d2h(self._h, out.ctypes.data_as(C.c_void_p), C.c_void_p(d_ptr), nbytes),
              "acm_memcpy_d2h")
        self.sync()
        return out

    def synth_fill(self, d_ptr, n, seed, offset=0):
        check(self.L.acm_synth_fill_device(self._h, C.c_void_p(d_ptr), n, seed, offset),
              "acm_synth_fill_device")

    def plant(self, d_ptr, n, buf_offset, plants):
        """plants: synth.Plants (positions are absolute stream offsets)."""
        if plants.count == 0:
            return
        if not plants.disjoint:
            raise ValueError("overlapping plants are order dependent; apply them on the host")
        check(self.L.acm_plant_device(
            self._h, C.c_void_p(d_ptr), n, buf_offset,
            plants.pos.ctypes.data_as(_lib.u64p), plants.blob_off.ctypes.data_as(_lib.u32p),
            plants.length.ctypes.data_as(_lib.u32p), plants.count,
            plants.blob.ctypes.data_as(_lib.u8p), plants.blob.size), "acm_plant_device")

    def close(self):
        if self._h:
            self.L.acm_device_close(self._h)
            self._h = None


class Scanner:
    """struct acm_scanner: scratch + result buffers for scans of up to max_bytes symbols."""

    def __init__(self, device, automaton, max_bytes, mode=MODE_AUTO, bucket_shift=0, bucket_cap=0,
                 timing=False, dfa_chunk=0, own_stream=False):
        self.L = lib()
        self.device = device
        self.automaton = automaton
        p = ScanParams(mode=mode, bucket_shift=bucket_shift, bucket_cap=bucket_cap,
                       timing=int(timing), dfa_chunk=dfa_chunk, own_stream=int(own_stream))
        h = C.c_void_p()
        check(self.L.acm_scanner_create(device.handle, automaton, max_bytes, C.byref(p), C.byref(h)),
              "acm_scanner_create")
        self._h = h
        self.max_bytes = max_bytes
        self.last = None

    def scan_device(self, d_ptr, n, emit_lo=0, emit_hi=None, valid_lo=0):
        """Scan n symbols at device pointer d_ptr; returns the ScanResult."""
        res = ScanResult()
        if emit_hi is None:
            emit_hi = n
        check(self.L.acm_scan_device_ex(self._h, C.c_void_p(d_ptr), n, valid_lo, emit_lo, emit_hi,
                                        C.byref(res)), "acm_scan_device")
        self.last = res
        return res

    def scan_async(self, d_ptr, n, emit_lo=0, emit_hi=None, valid_lo=0, push=None):
        """Queue a scan and return at once (acm_scan_device_async); finish() completes it.
        push = (device pointer of a gather region, capacity in keys, key_add) makes the step
        copy its sorted keys there itself."""
        if emit_hi is None:
            emit_hi = n
        pt = None
        if push is not None:
            pt = C.byref(PushTarget(C.c_void_p(push[0]), push[1], push[2]))
        check(self.L.acm_scan_device_async(self._h, C.c_void_p(d_ptr), n, valid_lo, emit_lo, emit_hi, pt),
              "acm_scan_device_async")

    def finish(self):
        """Wait for the queued scan (acm_scan_finish); returns its ScanResult."""
        res = ScanResult()
        check(self.L.acm_scan_finish(self._h, C.byref(res)), "acm_scan_finish")
        self.last = res
        return res

    def fetch(self, base=0, count=None):
        """(offsets u64, pattern indices u32) of the last scan, canonical order."""
        n = self.last.n_matches if count is None else count
        off = np.empty(n, dtype=np.uint64)
        pat = np.empty(n, dtype=np.uint32)
        if n:
            got = self.L.acm_scan_fetch(self._h, base, off.ctypes.data_as(_lib.u64p),
                                        pat.ctypes.data_as(_lib.u32p), n)
            check(got, "acm_scan_fetch")
            off, pat = off[:got], pat[:got]
        return off, pat

    def keys_ptr(self):
        return self.L.acm_scan_keys(self._h)

    def stream(self):
        """cudaStream_t (int) this scanner's scans are queued on."""
        return self.L.acm_scanner_stream(self._h)

    def histogram_into(self, d_counts_ptr):
        check(self.L.acm_scan_histogram(self._h, C.c_void_p(d_counts_ptr)), "acm_scan_histogram")

    def scan_host(self, data, base=0, cap=None, h_ptr=None, n=None):
        """Pipelined scan of a host buffer (numpy array, or raw pointer + n).  Returns
        (offsets, patterns, ScanResult)."""
        if h_ptr is None:
            a = np.ascontiguousarray(data)
            h_ptr = a.ctypes.data
            n = a.size
        res = ScanResult()
        cap = cap or (1 << 20)
        while True:
            off = np.empty(cap, dtype=np.uint64)
            pat = np.empty(cap, dtype=np.uint32)
            found = self.L.acm_scan_host(self._h, C.c_void_p(h_ptr), n, base,
                                         off.ctypes.data_as(_lib.u64p), pat.ctypes.data_as(_lib.u32p),
                                         cap, C.byref(res))
            check(found, "acm_scan_host")
            if found <= cap:
                return off[:found], pat[:found], res
            cap = int(found)

    def close(self):
        if self._h:
            # acm_scanner_free uses the device (ordinal, streams): after Device.close() that handle is
            # freed memory, so the scanner is abandoned instead (its buffers stay until the process ends)
            if getattr(self.device, "_h", None):
                self.L.acm_scanner_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def pinned_empty(nbytes, device=None):
    """uint8 numpy view over pinned host memory (caller keeps the returned owner alive); with a
    Device, preferably on that GPU's NUMA node."""
    L = lib()
    p = C.c_void_p()
    if device is not None:
        check(L.acm_host_alloc_pinned_near(device.handle, nbytes, C.byref(p)), "acm_host_alloc_pinned_near")
    else:
        check(L.acm_host_alloc_pinned(nbytes, C.byref(p)), "acm_host_alloc_pinned")
    buf = (C.c_ubyte * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.uint8)

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                L.acm_host_free_pinned(C.c_void_p(self.ptr))
            except Exception:
                pass
    return arr, _Owner(p.value)
