"""Parity at BASELINE.json's full sizes.  The whole sorted list cannot be compared with a
single-threaded oracle walk in seconds, so: (1) the match COUNT over the full stream against
the oracle's pthread-sharded walk (same bytes, generated on the host), (2) the full LIST on a
64 MiB prefix, (3) size-independent invariances of the full list -- the same digest from
different kernels, different shardings and different bucket shapes, sortedness, and every
planted signature present at its planted end offset."""
import hashlib
import os

import numpy as np
import pytest

import gpu_pattern_matching_b200 as g
from gpu_pattern_matching_b200 import sharded, synth
from helpers import build_oracle, build_product, clamav_pats

pytestmark = pytest.mark.gpu
GIB = 1 << 30


def _digest(off, pat):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(off, dtype=np.uint64).tobytes())
    h.update(np.ascontiguousarray(pat, dtype=np.uint32).tobytes())
    return h.hexdigest()


def _device_stream(device, pats, n, seed, plants):
    d = device.alloc(n + 64)
    device.synth_fill(d, n, seed)
    pl = synth.Plants([p for p, _ in pats], n, plants, seed)
    device.plant(d, n, 0, pl)
    device.sync()
    return d, pl


def _scan_all(device, a, d, n, **kw):
    sc = g.Scanner(device, a.automaton, n, **kw)
    res = sc.scan_device(d, n)
    off, pat = sc.fetch()
    sc.close()
    return off, pat, res


@pytest.mark.parametrize("nsig,nbytes,plants,seed", [(10000, GIB, 100000, 2), (15000, 4 * GIB, 400000, 3)])
def test_full_size_stream(device, nsig, nbytes, plants, seed):
    pats = clamav_pats(nsig)
    o, a = build_oracle(pats), build_product(pats)
    d, pl = _device_stream(device, pats, nbytes, seed, plants)
    off, pat, res = _scan_all(device, a, d, nbytes)
    assert res.mode == g.MODE_SAMPLED4 and not res.fallback

    # (3a) canonical order, offsets in range
    key = (off.astype(np.uint64) << np.uint64(24)) | pat
    assert np.all(key[1:] > key[:-1]) or np.all(key[1:] >= key[:-1])
    assert off.size >= plants and int(off.max()) < nbytes

    # (3b) every plant is reported at its end offset with its pattern index
    ends = pl.pos + pl.length.astype(np.uint64) - np.uint64(1)
    want = (ends << np.uint64(24)) | pl.pid.astype(np.uint64)
    assert np.isin(want, key).all()

    # (1) count against the reference walk over the same bytes (host copy of the device stream)
    host = device.d2h(d, nbytes)
    assert o.walk_count_mt(host, os.cpu_count() or 4) == off.size

    # (2) the full list on a 64 MiB prefix
    m = 64 << 20
    eo, ep, _, _ = o.search(host[:m])
    k = int(np.searchsorted(off, m))
    assert np.array_equal(off[:k], eo) and np.array_equal(pat[:k], ep)
    del host

    # (3c) invariance: other kernel (on the first GiB), 3 uneven shards with halo, other bucket shape
    full = _digest(off, pat)
    first = min(nbytes, GIB)
    kf = int(np.searchsorted(off, first))
    o2, p2, _ = _scan_all(device, a, d, first, mode=g.MODE_START2)
    assert _digest(o2, p2) == _digest(off[:kf], pat[:kf])
    o3, p3, r3 = _scan_all(device, a, d, nbytes, bucket_shift=18, bucket_cap=4096)
    assert _digest(o3, p3) == full and not r3.fallback
    halo = a.get_max_pattern_size() - 1
    cuts = [0, (nbytes // 3) // 16 * 16 + 16, (nbytes // 2 + 12345) // 16 * 16, nbytes]
    offs, ps = [], []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        start = max(0, lo - halo) // 16 * 16
        sc = g.Scanner(device, a.automaton, hi - lo)
        sc.scan_device(d + start, hi - start, lo - start, hi - start)
        so, sp = sc.fetch(base=start)
        sc.close()
        offs.append(so)
        ps.append(sp)
    assert _digest(np.concatenate(offs), np.concatenate(ps)) == full
    device.free(d)
