"""Synthetic streams shared by tests and bench: a counter-based generator that the C
library reproduces bit for bit on the host (acm_synth_fill_host) and on the device
(k_synth_fill), and disjoint signature plants.

stream byte i = byte (i & 7) of mix64(seed, i >> 3)
"""
import numpy as np

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xD1B54A32D192ED03)
_M3 = np.uint64(0xBF58476D1CE4E5B9)
_M4 = np.uint64(0x94D049BB133111EB)


def mix64(seed, idx):
    """Vectorised splitmix64-style hash; idx is a uint64 array."""
    with np.errstate(over="ignore"):
        z = (idx + np.uint64(1)) * _M1 + np.uint64(seed) * _M2
        z = (z ^ (z >> np.uint64(30))) * _M3
        z = (z ^ (z >> np.uint64(27))) * _M4
        return z ^ (z >> np.uint64(31))


def stream(n, seed, offset=0):
    """Bytes [offset, offset + n) of the random stream `seed` (numpy reference form)."""
    w0 = offset >> 3
    w1 = (offset + n + 7) >> 3
    words = mix64(seed, np.arange(w0, w1, dtype=np.uint64))
    b = words.view(np.uint8)            # little endian
    s = offset - (w0 << 3)
    return b[s:s + n].copy()


class Plants:
    """Disjoint signature plants: slot k of `count` equal slots receives one pattern at a
    pseudo-random offset inside the slot.  Extra `forced` (position, pattern id) plants are
    applied afterwards (tests put them across tile / shard cuts)."""

    def __init__(self, patterns, total_bytes, count, seed, forced=()):
        self.patterns = patterns
        lens = np.array([len(p) for p in patterns], dtype=np.int64)
        maxlen = int(lens.max())
        pos, pid = [], []
        if count > 0:
            slot = total_bytes // count
            if slot < 2 * maxlen + 2:
                raise ValueError("too many plants for this stream")
            k = np.arange(count, dtype=np.uint64)
            pids = (mix64(seed ^ 0x5151, k) % np.uint64(len(patterns))).astype(np.int64)
            room = (slot - lens[pids]).astype(np.uint64)
            offs = (mix64(seed ^ 0xA7A7, k) % room).astype(np.int64)
            pos = (k.astype(np.int64) * slot + offs).tolist()
            pid = pids.tolist()
        for p, i in forced:
            pos.append(int(p))
            pid.append(int(i))
        self.count = len(pos)
        # overlapping plants depend on the order they are applied in: fine on the host
        # (apply_host is sequential), a race in the device kernel (one warp per plant)
        order = np.argsort(np.array(pos, dtype=np.int64), kind="stable") if pos else np.array([], dtype=np.int64)
        ends = np.array([pos[i] + len(patterns[pid[i]]) for i in order], dtype=np.int64)
        starts = np.array([pos[i] for i in order], dtype=np.int64)
        self.disjoint = bool(np.all(ends[:-1] <= starts[1:])) if len(order) > 1 else True
        self.pos = np.array(pos, dtype=np.uint64)
        self.pid = np.array(pid, dtype=np.uint32)
        self.length = np.array([len(patterns[i]) for i in pid], dtype=np.uint32)
        # blob of the distinct patterns used
        used = sorted(set(pid))
        offs, cur = {}, 0
        for i in used:
            offs[i] = cur
            cur += len(patterns[i])
        self.blob = np.frombuffer(b"".join(patterns[i] for i in used) or b"\0", dtype=np.uint8).copy()
        self.blob_off = np.array([offs[i] for i in pid], dtype=np.uint32)

    def apply_host(self, buf, buf_offset=0):
        """Overwrite the numpy byte buffer (stream bytes [buf_offset, buf_offset+len))."""
        n = buf.size
        for p, i in zip(self.pos.tolist(), self.pid.tolist()):
            pat = self.patterns[i]
            lo = p - buf_offset
            for k in range(max(0, -lo), min(len(pat), n - lo)):
                buf[lo + k] = pat[k]
        return buf


def english_like(words, n, seed, zipf=1.0):
    """Space / newline separated words drawn Zipf(zipf) from `words` (list of bytes), n bytes."""
    rng = np.random.default_rng(seed)
    ranks = np.arange(1, len(words) + 1, dtype=np.float64)
    p = ranks ** (-zipf)
    p /= p.sum()
    avg = sum(len(w) for w in words[:200]) / 200 + 1
    out = bytearray()
    while len(out) < n:
        k = int((n - len(out)) / avg) + 64
        idx = rng.choice(len(words), size=k, p=p)
        seps = rng.random(k) < 0.08
        parts = []
        for i, nl in zip(idx.tolist(), seps.tolist()):
            parts.append(words[i])
            parts.append(b"\n" if nl else b" ")
        out += b"".join(parts)
    return np.frombuffer(bytes(out[:n]), dtype=np.uint8).copy()
