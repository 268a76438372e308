"""Python mirror of the reference's builder interface (acsmx.h / iacsmx.h) over the
C ABI.  Method names and argument meaning follow the C functions one to one so the
parity tests read like C caller code; all work happens in libacmatch_b200.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import AcmError, check, lib


class Acsm:
    """acsm_t: byte-alphabet automaton (reference acsmx.h:96-196)."""

    def __init__(self):
        self.L = lib()
        self._p = self.L.acsm_new()
        if not self._p:
            raise AcmError("acsm_new failed")

    # -- reference API -----------------------------------------------------
    def add_pattern(self, pat, iid=0, nocase=0, offset=0, depth=0):
        pat = bytes(pat)
        self.L.acsm_add_pattern(self._p, pat, len(pat), nocase, offset, depth, None, iid)

    def load_pattern_file(self, path, hex_pat=False, pat_size_limit=-1):
        n = self.L.acsm_load_pattern_file(self._p, str(path).encode(), int(hex_pat), pat_size_limit)
        if n < 0:
            raise AcmError(f"pattern file {path}: {_lib.last_error()}")
        return n

    def compile(self):
        self.L.acsm_compile(self._p)
        check(self.status(), "acsm_compile")

    def gen_state_table(self, mapped=0, ctx=None, queue=None):
        self.L.acsm_gen_state_table(self._p, mapped, ctx, queue)
        check(self.status(), "acsm_gen_state_table")

    def get_max_pattern_size(self):
        return self.L.acsm_get_max_pattern_size(self._p)

    def get_min_pattern_size(self):
        return self.L.acsm_get_min_pattern_size(self._p)

    def get_states(self):
        return self.L.acsm_get_states(self._p)

    def get_size(self):
        return self.L.acsm_get_size(self._p)

    def get_patterns_table(self):
        """[(pattern bytes, iid, index)] in index order."""
        n = self._p.contents.num_patterns
        tab = self.L.acsm_get_patterns_table(self._p)
        if not tab:
            return []
        out = [(bytes(tab[i].pattern[:tab[i].n]), tab[i].iid, tab[i].index) for i in range(n)]
        self.L.acsm_free_patterns_table(tab, n)
        return out

    def cleanup(self):
        self.L.acsm_cleanup(self._p)

    def free(self):
        if self._p:
            self.L.acsm_free(self._p)
            self._p = None

    # -- additions ---------------------------------------------------------
    def status(self):
        return self.L.acsm_status(self._p)

    @property
    def num_patterns(self):
        return self._p.contents.num_patterns

    def check_filters(self):
        """Host-side self-check of the scan-filter tables: violations (0 = consistent), -1 = no filter."""
        return self.L.acsm_check_filters(self._p)

    def export_ref_table(self):
        """The table in the reference's layout and numbering, int32[num_states][512]."""
        check(self.L.acsm_export_ref_table(self._p), "acsm_export_ref_table")
        ns = self._p.contents.num_states
        if self._p.contents.d_trans is None:
            ns += 1      # before gen_state_table the field holds the highest id (acsmx.c:615)
        return np.ctypeslib.as_array(self._p.contents.h_trans, shape=(ns, 512))

    @property
    def automaton(self):
        a = self.L.acsm_device_automaton(self._p)
        if not a:
            raise AcmError("automaton is not on the device: call gen_state_table() first")
        return a

    @property
    def ptr(self):
        return self._p

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Iacsm:
    """iacsm_t: ushort-symbol automaton, alphabet 2048 (reference AC_ushorts/iacsmx.h:92-185)."""

    def __init__(self):
        self.L = lib()
        self._p = self.L.iacsm_new()
        if not self._p:
            raise AcmError("iacsm_new failed")

    def add_pattern(self, items, iid=0):
        a = np.ascontiguousarray(items, dtype=np.uint16)
        self.L.iacsm_add_pattern(self._p, a.ctypes.data_as(_lib.u16p), a.size, 0, 0, None, iid)
        check(self.status(), "iacsm_add_pattern")

    def add_fullpattern(self, csv, iid):
        self.L.iacsm_add_fullpattern(self._p, csv.encode(), iid)
        check(self.status(), "iacsm_add_fullpattern")

    def compile(self):
        self.L.iacsm_compile(self._p)
        check(self.status(), "iacsm_compile")

    def gen_state_table(self, mapped=0, ctx=None, queue=None):
        self.L.iacsm_gen_state_table(self._p, mapped, ctx, queue)
        check(self.status(), "iacsm_gen_state_table")

    def get_max_pattern_size(self):
        return self.L.iacsm_get_max_pattern_size(self._p)

    def get_states(self):
        return self.L.iacsm_get_states(self._p)

    def get_size(self):
        return self.L.iacsm_get_size(self._p)

    def status(self):
        return self.L.iacsm_status(self._p)

    def export_ref_table(self):
        check(self.L.iacsm_export_ref_table(self._p), "iacsm_export_ref_table")
        ns = self._p.contents.num_states
        if self._p.contents.d_trans is None:
            ns += 1
        return np.ctypeslib.as_array(self._p.contents.h_trans, shape=(ns, 4096))

    @property
    def automaton(self):
        a = self.L.iacsm_device_automaton(self._p)
        if not a:
            raise AcmError("automaton is not on the device: call gen_state_table() first")
        return a

    @property
    def ptr(self):
        return self._p

    def cleanup(self):
        self.L.iacsm_cleanup(self._p)

    def free(self):
        if self._p:
            self.L.iacsm_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
