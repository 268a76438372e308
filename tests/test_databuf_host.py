"""Host side of databuf (no GPU): the chunk bookkeeping of databuf_add_fd / _add_fp / _add_chunk /
_reset on a caller-built struct databuf -- chunk tables, zero padding, 16-byte line alignment and
the reference's return codes (reference databuf.c:327-528: > 0 more room, 0 end of file, -1 chunk
table full, -2 byte buffer full, -3 chunk too large).  These functions never touch the device."""
import ctypes as C
import os

import numpy as np
import pytest

import gpu_pattern_matching_b200 as g
from gpu_pattern_matching_b200 import _lib


class HostBuf:
    """struct databuf with host arrays only (priv = NULL: nothing here may reach the device)."""

    def __init__(self, max_chunks, chunk_size):
        self.data = np.full(max_chunks * chunk_size + 64, 0xEE, dtype=np.uint8)      # stale bytes on purpose
        self.indices = np.full(max_chunks + 1, -7, dtype=np.int32)
        self.sizes = np.full(max_chunks + 1, -7, dtype=np.int32)
        self.file_ids = np.full(max_chunks + 1, -7, dtype=np.int32)
        self.db = _lib.Databuf()
        self.db.h_data = self.data.ctypes.data_as(_lib.u8p)
        self.db.h_indices = self.indices.ctypes.data_as(_lib.i32p)
        self.db.h_sizes = self.sizes.ctypes.data_as(_lib.i32p)
        self.db.file_ids = self.file_ids.ctypes.data_as(_lib.i32p)
        self.db.max_chunks = max_chunks
        self.db.max_chunk_size = chunk_size
        self.db.size = max_chunks * chunk_size
        self.db.max_results = 16
        self.ref = C.byref(self.db)


@pytest.fixture(scope="module")
def L():
    return g.lib()


def test_add_fd_fixed_chunks_padding_and_codes(L, tmp_path):
    content = bytes(range(256)) * 5 + b"tail!"                  # 1285 bytes
    f = tmp_path / "a.bin"
    f.write_bytes(content)
    hb = HostBuf(max_chunks=8, chunk_size=256)                   # 2048-byte buffer
    rd = C.c_size_t(0)
    fd = os.open(f, os.O_RDONLY)
    try:
        e = L.databuf_add_fd(hb.ref, fd, 3, C.byref(rd))
        assert e == 1285 and rd.value == 1285                    # more room: bytes read
        assert hb.db.chunks == 6 and hb.db.bytes == 6 * 256
        assert hb.indices[:6].tolist() == [0, 256, 512, 768, 1024, 1280]
        assert hb.sizes[:6].tolist() == [256] * 5 + [5] and hb.file_ids[:6].tolist() == [3] * 6
        assert bytes(hb.data[:1285]) == content
        assert not hb.data[1285:1536].any()                      # the partial chunk is zero padded
        assert hb.data[1536] == 0xEE                             # ... and nothing beyond it is touched
        e = L.databuf_add_fd(hb.ref, fd, 3, C.byref(rd))
        assert e == 0 and rd.value == 0                          # end of file
        # a second file goes behind the padded chunk
        os.lseek(fd, 0, os.SEEK_SET)
        e = L.databuf_add_fd(hb.ref, fd, 4, C.byref(rd))
        assert e == -1 and rd.value == 512                       # chunk table full after 2 more chunks
        assert hb.db.chunks == 8 and hb.file_ids[6:8].tolist() == [4, 4]
        assert bytes(hb.data[1536:2048]) == content[:512]
        # full on entry: -1 (the reference would read 0 bytes here and report end of file; its own
        # loop never gets there because it processes and resets the buffer on the first -1)
        assert L.databuf_add_fd(hb.ref, fd, 4, C.byref(rd)) == -1 and rd.value == 0
        L.databuf_reset(hb.ref)
        assert hb.db.chunks == 0 and hb.db.bytes == 0
    finally:
        os.close(fd)
    # a file that fills the whole buffer in one read: -1 (chunks) wins over -2, as in the reference
    big = tmp_path / "b.bin"
    big.write_bytes(b"x" * 4096)
    fd = os.open(big, os.O_RDONLY)
    try:
        e = L.databuf_add_fd(hb.ref, fd, 0, C.byref(rd))
        assert e == -1 and rd.value == 2048 and hb.db.chunks == 8
    finally:
        os.close(fd)


def test_add_chunk_codes_alignment(L):
    hb = HostBuf(max_chunks=4, chunk_size=64)                    # 256-byte buffer
    buf = C.create_string_buffer(b"0123456789abcdefXYZ", 19)
    e = L.databuf_add_chunk(hb.ref, buf, 19, 9, 1)
    assert e == 256 - 32 and hb.db.chunks == 1 and hb.db.bytes == 32            # rounded up to 16
    assert bytes(hb.data[:19]) == b"0123456789abcdefXYZ" and not hb.data[19:32].any()
    assert (hb.indices[0], hb.sizes[0], hb.file_ids[0]) == (0, 19, 9)
    e = L.databuf_add_chunk(hb.ref, buf, 19, 9, 0)                              # not aligned
    assert e == 256 - 51 and hb.db.bytes == 51 and hb.indices[1] == 32
    big = C.create_string_buffer(65)
    assert L.databuf_add_chunk(hb.ref, big, 65, 0, 1) == -3                     # larger than a chunk
    assert L.databuf_add_chunk(hb.ref, big, 64, 0, 1) > 0
    assert L.databuf_add_chunk(hb.ref, big, 64, 0, 1) > 0
    assert hb.db.chunks == 4
    assert L.databuf_add_chunk(hb.ref, buf, 1, 0, 1) == -1                      # chunk table full
    hb2 = HostBuf(max_chunks=64, chunk_size=64)
    hb2.db.size = 100                                                            # byte buffer smaller than the table
    assert L.databuf_add_chunk(hb2.ref, big, 64, 0, 1) > 0
    assert L.databuf_add_chunk(hb2.ref, big, 36, 0, 1) == -2                    # bytes + len >= size


def test_add_fp_one_chunk_per_line(L, tmp_path):
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    lines = [b"short\n", b"x" * 40 + b"\n", b"\n", b"y" * 100 + b"\n", b"last line without newline"]
    f = tmp_path / "t.txt"
    f.write_bytes(b"".join(lines))
    hb = HostBuf(max_chunks=32, chunk_size=64)
    rb, rl = C.c_size_t(0), C.c_size_t(0)
    fp = libc.fopen(str(f).encode(), b"r")
    try:
        e = L.databuf_add_fp(hb.ref, fp, 5, 1, C.byref(rb), C.byref(rl))
    finally:
        libc.fclose(fp)
    # the 101-byte line is split into 63 + 38 (fgets keeps one byte for its NUL), like the reference
    want = [b"short\n", b"x" * 40 + b"\n", b"\n", b"y" * 63, b"y" * 37 + b"\n", b"last line without newline"]
    n = hb.db.chunks
    assert n == len(want) and rl.value == 4 and rb.value == sum(len(l) for l in lines)
    off = 0
    for i, w in enumerate(want):
        assert hb.indices[i] == off and hb.sizes[i] == len(w) and hb.file_ids[i] == 5
        assert bytes(hb.data[off:off + len(w)]) == w
        pad = (len(w) + 15) // 16 * 16
        assert not hb.data[off + len(w):off + pad].any()        # zero padding up to the 16-byte boundary
        off += pad
    assert hb.db.bytes == off and e == hb.db.size - off
    # chunk table full in the middle of a file
    hb = HostBuf(max_chunks=2, chunk_size=64)
    fp = libc.fopen(str(f).encode(), b"r")
    try:
        assert L.databuf_add_fp(hb.ref, fp, 0, 1, C.byref(rb), C.byref(rl)) == -1 and hb.db.chunks == 2
        assert rl.value == 2
        L.databuf_reset(hb.ref)
        assert L.databuf_add_fp(hb.ref, fp, 0, 1, C.byref(rb), C.byref(rl)) == -1     # continues where it stopped
        assert bytes(hb.data[:1]) == b"\n" and bytes(hb.data[16:16 + 63]) == b"y" * 63
    finally:
        libc.fclose(fp)
