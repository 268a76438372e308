"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when it has been
built, the compiled reference builder (oracle/_ref/libacref.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.
"""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
FIXTURES = os.path.join(ROOT, "tests", "fixtures")
GOLDEN = os.path.join(ROOT, "tests", "golden")

_u8p = C.POINTER(C.c_ubyte)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_u16p = C.POINTER(C.c_ushort)


def _ensure_built():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "acsm_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


_lib = None
_ref = None


def oracle_lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_ensure_built())
        L.orc_new.restype = C.c_void_p
        L.orc_new.argtypes = [C.c_int]
        L.orc_add.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.orc_add_syms.argtypes = [C.c_void_p, _u16p, C.c_int, C.c_int]
        L.orc_add_csv.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.orc_compile.argtypes = [C.c_void_p]
        for f in ("orc_num_states", "orc_num_patterns", "orc_max_pattern_len"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int
        L.orc_table_bytes.argtypes = [C.c_void_p]
        L.orc_table_bytes.restype = C.c_size_t
        L.orc_ml_begin.argtypes = [C.c_void_p]
        L.orc_ml_begin.restype = _i64p
        L.orc_ml_index.argtypes = [C.c_void_p]
        L.orc_ml_index.restype = _i32p
        L.orc_pat_len.argtypes = [C.c_void_p, C.c_int]
        L.orc_pat_iid.argtypes = [C.c_void_p, C.c_int]
        L.orc_pat_bytes.argtypes = [C.c_void_p, C.c_int]
        L.orc_pat_bytes.restype = _u8p
        L.orc_pat_syms.argtypes = [C.c_void_p, C.c_int]
        L.orc_pat_syms.restype = _u16p
        L.orc_ref_table.argtypes = [C.c_void_p]
        L.orc_ref_table.restype = _i32p
        L.orc_next_table.argtypes = [C.c_void_p]
        L.orc_next_table.restype = _i32p
        L.orc_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                 C.c_int64, C.c_uint64, _u64p, _u32p, C.c_int64,
                                 _i64p, _i32p]
        L.orc_search.restype = C.c_int64
        L.orc_walk_count.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.orc_walk_count.restype = C.c_int64
        L.orc_walk_count_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.orc_walk_count_mt.restype = C.c_int64
        L.orc_load_pattern_file.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.orc_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def ref_available():
    # ACM_SKIP_REF=1 (tools/run_asan_tests.sh): the compiled reference is not ours to sanitize -- its
    # iacsm_add_pattern copies the pattern into a malloc(sizeof(len)) buffer (reference AC_ushorts/iacsmx.c:398)
    return not os.environ.get("ACM_SKIP_REF") and os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libacref.so"))


def ref_lib():
    """The reference's own acsmx.c/iacsmx.c, compiled by oracle/ref_build."""
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(ORACLE_DIR, "_ref", "libacref.so"))
        L.ref_new.restype = C.c_void_p
        L.ref_add.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        L.ref_compile.argtypes = [C.c_void_p]
        for f in ("ref_num_states", "ref_num_patterns", "ref_max_pattern_len"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int
        L.ref_table_bytes.argtypes = [C.c_void_p]
        L.ref_table_bytes.restype = C.c_size_t
        L.ref_h_trans.argtypes = [C.c_void_p]
        L.ref_h_trans.restype = _i32p
        L.ref_ml_begin.argtypes = [C.c_void_p]
        L.ref_ml_begin.restype = _i64p
        L.ref_ml_index.argtypes = [C.c_void_p]
        L.ref_ml_index.restype = _i32p
        L.ref_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                 C.c_int64, C.c_uint64, _u64p, _u32p, C.c_int64,
                                 _i64p, _i32p]
        L.ref_search.restype = C.c_int64
        L.ref_walk_count.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.ref_walk_count.restype = C.c_int64
        L.ref_walk_count_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.ref_walk_count_mt.restype = C.c_int64
        L.ref_free.argtypes = [C.c_void_p]
        L.iref_new.restype = C.c_void_p
        L.iref_add.argtypes = [C.c_void_p, _u16p, C.c_int, C.c_int]
        L.iref_add_csv.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.iref_compile.argtypes = [C.c_void_p]
        L.iref_num_states.argtypes = [C.c_void_p]
        L.iref_max_pattern_len.argtypes = [C.c_void_p]
        L.iref_h_trans.argtypes = [C.c_void_p]
        L.iref_h_trans.restype = _i32p
        L.iref_ml_begin.argtypes = [C.c_void_p]
        L.iref_ml_begin.restype = _i64p
        L.iref_ml_iid.argtypes = [C.c_void_p]
        L.iref_ml_iid.restype = _i32p
        L.iref_search.argtypes = [C.c_void_p, _u16p, C.c_int64, _u64p, _i32p,
                                  C.c_int64, _i32p]
        L.iref_search.restype = C.c_int64
        L.iref_free.argtypes = [C.c_void_p]
        _ref = L
    return _ref


def _as_u8(data):
    if isinstance(data, (bytes, bytearray)):
        return np.frombuffer(bytes(data), dtype=np.uint8)
    a = np.ascontiguousarray(data)
    assert a.dtype == np.uint8
    return a


class _Searchable:
    """Shared search plumbing: subclasses set self._h and self._search."""

    def search(self, data, start_state=0, emit_from=0, base=0, cap=None):
        """Return (offsets u64, patterns u32) in canonical (offset, index) order,
        plus (hits, final_state)."""
        if self.alpha == 256:
            a = _as_u8(data)
        else:
            a = np.ascontiguousarray(data, dtype=np.uint16)
        n = a.size
        hits = C.c_int64(0)
        fin = C.c_int32(0)
        if cap is None:
            cap = max(1024, n // 8)
        while True:
            off = np.empty(cap, dtype=np.uint64)
            pat = np.empty(cap, dtype=np.uint32)
            found = self._search(self._h, a.ctypes.data_as(C.c_void_p), n, start_state,
                                 emit_from, base, off.ctypes.data_as(_u64p),
                                 pat.ctypes.data_as(_u32p), cap, C.byref(hits),
                                 C.byref(fin))
            if found <= cap:
                break
            cap = int(found)
        off, pat = off[:found], pat[:found]
        order = np.lexsort((pat, off))
        return off[order], pat[order], int(hits.value), int(fin.value)


class Oracle(_Searchable):
    """oracle/acsm_oracle.c: the CPU restatement."""

    def __init__(self, alphabet=256):
        self.L = oracle_lib()
        self.alpha = alphabet
        self._h = C.c_void_p(self.L.orc_new(alphabet))
        self._search = self.L.orc_search
        self.compiled = False

    def add(self, pat, iid=0):
        if self.alpha == 256:
            self.L.orc_add(self._h, bytes(pat), len(pat), iid)
        else:
            a = np.ascontiguousarray(pat, dtype=np.uint16)
            self.L.orc_add_syms(self._h, a.ctypes.data_as(_u16p), a.size, iid)

    def add_csv(self, csv, iid):
        self.L.orc_add_csv(self._h, csv.encode(), iid)

    def load_file(self, path, hex_pat=False, limit=-1):
        return self.L.orc_load_pattern_file(self._h, path.encode(), int(hex_pat), limit)

    def compile(self):
        self.L.orc_compile(self._h)
        self.compiled = True

    @property
    def num_states(self):
        return self.L.orc_num_states(self._h)

    @property
    def num_patterns(self):
        return self.L.orc_num_patterns(self._h)

    @property
    def max_pattern_len(self):
        return self.L.orc_max_pattern_len(self._h)

    @property
    def table_bytes(self):
        return self.L.orc_table_bytes(self._h)

    def pattern(self, idx):
        n = self.L.orc_pat_len(self._h, idx)
        if self.alpha == 256:
            p = self.L.orc_pat_bytes(self._h, idx)
            return bytes(p[:n])
        p = self.L.orc_pat_syms(self._h, idx)
        return list(p[:n])

    def pattern_iid(self, idx):
        return self.L.orc_pat_iid(self._h, idx)

    def patterns(self):
        return [self.pattern(i) for i in range(self.num_patterns)]

    def ml_csr(self):
        ns = self.num_states
        b = np.ctypeslib.as_array(self.L.orc_ml_begin(self._h), shape=(ns + 1,)).copy()
        idx = np.ctypeslib.as_array(self.L.orc_ml_index(self._h), shape=(max(int(b[-1]), 1),))
        return b, idx[:int(b[-1])].copy()

    def ref_table(self):
        ns = self.num_states
        return np.ctypeslib.as_array(self.L.orc_ref_table(self._h),
                                     shape=(ns, 2 * self.alpha))

    def walk_count(self, data, emit_from=0):
        a = _as_u8(data)
        return self.L.orc_walk_count(self._h, a.ctypes.data_as(C.c_void_p), a.size, emit_from)

    def walk_count_mt(self, data, threads):
        a = _as_u8(data)
        return self.L.orc_walk_count_mt(self._h, a.ctypes.data_as(C.c_void_p), a.size, threads)

    def close(self):
        if self._h:
            self.L.orc_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefAcsm(_Searchable):
    """The reference's acsmx.c (bytes), via oracle/_ref/libacref.so."""
    alpha = 256

    def __init__(self):
        self.L = ref_lib()
        self._h = C.c_void_p(self.L.ref_new())
        self._search = self.L.ref_search

    def add(self, pat, iid=0):
        self.L.ref_add(self._h, bytes(pat), len(pat), iid)

    def compile(self):
        self.L.ref_compile(self._h)

    @property
    def num_states(self):
        return self.L.ref_num_states(self._h)

    @property
    def max_pattern_len(self):
        return self.L.ref_max_pattern_len(self._h)

    @property
    def table_bytes(self):
        return self.L.ref_table_bytes(self._h)

    def h_trans(self):
        return np.ctypeslib.as_array(self.L.ref_h_trans(self._h),
                                     shape=(self.num_states, 512))

    def ml_csr(self):
        ns = self.num_states
        b = np.ctypeslib.as_array(self.L.ref_ml_begin(self._h), shape=(ns + 1,)).copy()
        idx = np.ctypeslib.as_array(self.L.ref_ml_index(self._h), shape=(max(int(b[-1]), 1),))
        return b, idx[:int(b[-1])].copy()

    def walk_count(self, data, emit_from=0):
        a = _as_u8(data)
        return self.L.ref_walk_count(self._h, a.ctypes.data_as(C.c_void_p), a.size, emit_from)

    def walk_count_mt(self, data, threads):
        a = _as_u8(data)
        return self.L.ref_walk_count_mt(self._h, a.ctypes.data_as(C.c_void_p), a.size, threads)

    def close(self):
        if self._h:
            self.L.ref_free(self._h)
            self._h = None


class RefIacsm:
    """The reference's AC_ushorts/iacsmx.c (ushort symbols, alphabet 2048)."""
    alpha = 2048

    def __init__(self):
        self.L = ref_lib()
        self._h = C.c_void_p(self.L.iref_new())

    def add(self, items, iid):
        a = np.ascontiguousarray(items, dtype=np.uint16)
        self.L.iref_add(self._h, a.ctypes.data_as(_u16p), a.size, iid)

    def add_csv(self, csv, iid):
        self.L.iref_add_csv(self._h, csv.encode(), iid)

    def compile(self):
        self.L.iref_compile(self._h)

    @property
    def num_states(self):
        return self.L.iref_num_states(self._h)

    @property
    def max_pattern_len(self):
        return self.L.iref_max_pattern_len(self._h)

    def h_trans(self):
        return np.ctypeslib.as_array(self.L.iref_h_trans(self._h),
                                     shape=(self.num_states, 4096))

    def search(self, tokens):
        a = np.ascontiguousarray(tokens, dtype=np.uint16)
        cap = max(1024, a.size * 4)
        off = np.empty(cap, dtype=np.uint64)
        iid = np.empty(cap, dtype=np.int32)
        fin = C.c_int32(0)
        found = self.L.iref_search(self._h, a.ctypes.data_as(_u16p), a.size,
                                   off.ctypes.data_as(_u64p), iid.ctypes.data_as(_i32p),
                                   cap, C.byref(fin))
        assert found <= cap
        off, iid = off[:found], iid[:found]
        order = np.lexsort((iid, off))
        return off[order], iid[order], int(fin.value)


# ----------------------------------------------------------------------------
# fixtures
# ----------------------------------------------------------------------------

def fixture_path(name):
    return os.path.join(FIXTURES, name)


def read_fixture(name):
    p = fixture_path(name)
    if p.endswith(".gz"):
        with gzip.open(p, "rb") as f:
            return f.read()
    with open(p, "rb") as f:
        return f.read()


def materialize(name, tmpdir):
    """Decompress a .gz fixture into tmpdir so C code can fopen() it."""
    p = fixture_path(name)
    if not p.endswith(".gz"):
        return p
    out = os.path.join(str(tmpdir), os.path.basename(p)[:-3])
    if not os.path.exists(out):
        with open(out, "wb") as f:
            f.write(read_fixture(name))
    return out


def clamav_signatures(count):
    """First `count` ClamAV sample signatures as bytes (2000 / 10000 / 15000 are
    prefixes of one list: reference clamav_sample_sigs/*.txt)."""
    lines = read_fixture("clamav_sigs_15000.hex.gz").split(b"\n")
    lines = [l for l in lines if l][:count]
    return [bytes.fromhex(l.decode()) for l in lines]


def parse_pattern_file(data, hex_pat=False, limit=-1):
    """Python restatement of the pattern-file grammar (ocl_worker.c:74-145);
    returns [(pattern bytes, iid)].  Used to cross-check the C parsers."""
    out = []
    categ = False
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    for i, line in enumerate(lines):
        if i == 0:
            fb = -1
            for j, ch in enumerate(line):
                if ch in (0x20, 0x09):
                    fb = j
                    break
            if fb > 0:
                head = line[:fb]
                body = head[1:]
                if (head[:1] in (b"+", b"-") or head[:1].isdigit()) and \
                        (body == b"" or body.isdigit()):
                    categ = True
        if categ:
            j = 0
            while j < len(line) and line[j] in b" \t\n\v\f\r":
                j += 1
            k = j
            if k < len(line) and line[k] in b"+-":
                k += 1
            while k < len(line) and chr(line[k]).isdigit():
                k += 1
            try:
                pid = int(line[j:k])
            except ValueError:
                pid = 0
                k = 0
            while k < len(line) and line[k] in b" \t\n\v\f\r":
                k += 1
            pat = line[k:]
        else:
            pat = line
            pid = i
        if len(pat) >= 1 and pat[:1] == b'"' and pat[-1:] == b'"':
            pat = pat[1:-1] if len(pat) >= 2 else b""
        if hex_pat:
            if limit != -1:
                pat = pat[:limit * 2]
            pat = bytes.fromhex(pat.decode())
        elif limit != -1:
            pat = pat[:limit]
        out.append((pat, pid))
    return out
