"""The C-ABI shared library loads and exports every function include/*.h declares (no GPU, no
compute calls), the Python binding table covers exactly that set, and the struct layouts the
reference's callers touch are mirrored field for field."""
import ctypes as C
import glob
import os
import re

import gpu_pattern_matching_b200 as g
from gpu_pattern_matching_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DECL = re.compile(r"^[A-Za-z_][\w\s\*]*?\b(\w+)\s*\(", re.M)


def declared_functions():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
        src = src.replace('extern "C" {', "")
        src = re.sub(r"\{[^{}]*\}", "", src)            # struct bodies
        for stmt in src.split(";"):
            stmt = stmt.strip()
            if stmt.startswith("typedef") or "(" not in stmt:
                continue
            m = re.search(r"\b([A-Za-z_]\w*)\s*\(", stmt)
            if m:
                names.add(m.group(1))
    return names


def test_library_exports_every_declared_symbol():
    L = C.CDLL(g.LIB_PATH)
    names = declared_functions()
    assert len(names) > 80
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_binding_table_matches_headers():
    names = declared_functions()
    assert set(_lib.SIGNATURES) == names, (sorted(set(_lib.SIGNATURES) - names),
                                           sorted(names - set(_lib.SIGNATURES)))
    g.lib()      # sets argtypes for every entry; AttributeError would mean a missing symbol


def test_reference_entry_points_present():
    """the symbols of libacmatch.a + ocl_worker.o."""
    ref_api = """acsm_new acsm_add_pattern acsm_compile acsm_gen_state_table acsm_get_patterns_table
        acsm_get_max_pattern_size acsm_get_states acsm_get_size acsm_cleanup acsm_free
        iacsm_new iacsm_add_pattern iacsm_add_fullpattern iacsm_compile iacsm_gen_state_table
        iacsm_get_max_pattern_size iacsm_get_states iacsm_get_size iacsm_cleanup iacsm_free
        ocl_aho_match_init ocl_aho_match_close ocl_aho_match
        ocl_worker_ctx_create ocl_worker_ctx_init ocl_worker_ctx_free
        databuf_new databuf_add_fd databuf_add_fp databuf_add_chunk databuf_reset databuf_clear
        databuf_copy_host_to_device databuf_copy_device_to_host databuf_process_results databuf_free
        ocl_prefix_sum_init ocl_prefix_sum_close ocl_prefix_sum
        ocl_compact_array_init ocl_compact_array_close ocl_compact_array
        ocl_bitonic_sort_init ocl_bitonic_sort_close ocl_bitonic_sort clinitctx
        printable_hex_to_bytes gettime""".split()
    L = C.CDLL(g.LIB_PATH)
    assert not [n for n in ref_api if not hasattr(L, n)]


def test_no_cpu_fallback_without_device():
    L = g.lib()
    if L.acm_device_count() > 0:
        return
    h = C.c_void_p()
    assert L.acm_device_open(0, C.byref(h)) == -17 and not h.value
    assert b"no CPU fallback" in L.acm_last_error()
    assert not L.ocl_worker_ctx_create(0)
    conf = _lib.Clconf()
    L.clinitctx(C.byref(conf), 0, -1)
    assert not conf.ctx
    assert not L.databuf_new(16, 4096, 16, 0, C.byref(conf))


def test_struct_layouts(tmp_path):
    """ctypes mirrors == what the C compiler lays out from include/*.h (offsetof, sizeof)."""
    import subprocess
    checks = {
        "acsm_pattern_t": (_lib.AcsmPattern, ["next", "pattern", "n", "id", "iid", "index"]),
        "acsm_t": (_lib.AcsmStruct, ["max_states", "num_states", "max_pattern_len", "size",
                                     "num_patterns", "h_trans", "d_trans", "priv"]),
        "iacsm_t": (_lib.IacsmStruct, ["num_states", "size", "h_trans", "d_trans", "priv"]),
        "struct clconf": (_lib.Clconf, ["ctx", "queue", "kernel_compact_array", "type"]),
        "struct databuf": (_lib.Databuf, ["h_data", "h_indices", "h_sizes", "h_results", "h_results2",
                                          "h_results_comp", "h_results2_comp", "results_comp_size",
                                          "file_ids", "max_results", "last_state", "max_chunks",
                                          "max_chunk_size", "size", "chunks", "bytes", "d_data",
                                          "d_results", "d_results_comp", "cl", "priv"]),
        "struct ocl_worker_ctx": (_lib.WorkerCtx, ["id", "fds", "filenames", "matches_total",
                                                   "matches_reported", "bytes", "lines", "rounds",
                                                   "global_ws", "local_ws", "cl", "db", "acsm",
                                                   "patterns", "patterns_size"]),
        "struct acm_scan_params": (_lib.ScanParams, ["mode", "bucket_cap", "timing", "dfa_chunk"]),
        "struct acm_scan_result": (_lib.ScanResult, ["n_matches", "n_bytes", "mode", "fallback",
                                                     "final_state", "ms_scan", "ms_total", "launches"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "acm.h"', '#include "acsmx.h"',
             '#include "iacsmx.h"', '#include "databuf.h"', '#include "ocl_worker.h"',
             '#include "ocl_aho_match.h"', '#include "ocl_prefix_sum.h"', '#include "ocl_compact_array.h"',
             '#include "ocl_bitonic_sort.h"', '#include "utils.h"', "int main(void){"]
    for ctype, (_, fields) in checks.items():
        lines.append(f'printf("%zu\\n", sizeof({ctype}));')
        for f in fields:
            lines.append(f'printf("%zu\\n", offsetof({ctype}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=gnu11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    out = iter(int(x) for x in subprocess.check_output([str(exe)]).split())
    for ctype, (cls, fields) in checks.items():
        assert C.sizeof(cls) == next(out), ctype
        for f in fields:
            assert getattr(cls, f).offset == next(out), f"{ctype}.{f}"


def test_databuf_read_fd_parallel_reader(tmp_path, monkeypatch):
    """databuf_read_fd (what databuf_add_fd reads with): several pread() threads on a regular
    file give the same bytes and file offset as read(); short files, end of file, appended data,
    pipes and ACM_READ_THREADS=1 take the plain path."""
    import ctypes as C
    import numpy as np
    L = g.lib()
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, size=(40 << 20) + 12345, dtype=np.uint8)
    f = tmp_path / "big.bin"
    data.tofile(f)
    buf = np.zeros(48 << 20, dtype=np.uint8)
    for threads in ("4", "16", "3", "1"):
        monkeypatch.setenv("ACM_READ_THREADS", threads)
        fd = os.open(f, os.O_RDONLY)
        try:
            # a first call that does not reach the end, from an odd offset
            os.lseek(fd, 777, os.SEEK_SET)
            buf[:] = 0
            got = L.databuf_read_fd(fd, buf.ctypes.data_as(C.c_void_p), 16 << 20)
            assert got == 16 << 20 and np.array_equal(buf[:got], data[777:777 + got])
            assert os.lseek(fd, 0, os.SEEK_CUR) == 777 + got
            # the rest: less than asked for
            pos = 777 + got
            got = L.databuf_read_fd(fd, buf.ctypes.data_as(C.c_void_p), buf.size)
            assert got == data.size - pos and np.array_equal(buf[:got], data[pos:])
            assert L.databuf_read_fd(fd, buf.ctypes.data_as(C.c_void_p), buf.size) == 0      # end of file
            # appended data is seen by the next call (follow mode)
            with open(f, "ab") as fh:
                fh.write(b"tail")
            assert L.databuf_read_fd(fd, buf.ctypes.data_as(C.c_void_p), buf.size) == 4
            assert bytes(buf[:4]) == b"tail"
        finally:
            os.close(fd)
            data.tofile(f)
    # a pipe: plain read
    r, w = os.pipe()
    os.write(w, b"hello")
    os.close(w)
    assert L.databuf_read_fd(r, buf.ctypes.data_as(C.c_void_p), buf.size) == 5 and bytes(buf[:5]) == b"hello"
    assert L.databuf_read_fd(r, buf.ctypes.data_as(C.c_void_p), buf.size) == 0
    os.close(r)


def test_reference_style_caller_compiles_and_links(tmp_path):
    """examples/worker_loop.c -- the reference's cpu_worker loop written against include/*.h --
    compiles as C11 with -Wall -Werror and links against the shared library; without a GPU it
    must fail loudly (no CPU path), with the library's error text."""
    import subprocess
    src = os.path.join(ROOT, "examples", "worker_loop.c")
    exe = str(tmp_path / "worker_loop")
    libdir = os.path.join(ROOT, "gpu_pattern_matching_b200")
    subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src,
                    "-L", libdir, "-lacmatch_b200", "-lpthread", f"-Wl,-rpath,{libdir}", "-o", exe], check=True)
    # the CLI tools exist and refuse to run without their arguments
    for tool in ("b200_aho_grep", "b200_flow_grep"):
        p = subprocess.run([os.path.join(ROOT, "cli", tool), "-h"], capture_output=True, timeout=60)
        assert p.returncode != 0 and (b"Usage" in p.stdout + p.stderr or b"usage" in p.stdout + p.stderr)
    if g.lib().acm_device_count() <= 0:
        pf = tmp_path / "p.txt"
        pf.write_text("needle\n")
        df = tmp_path / "d.txt"
        df.write_text("hay needle hay\n")
        p = subprocess.run([exe, str(pf), str(df)], capture_output=True, timeout=60)
        assert p.returncode != 0 and p.stderr.strip(), "a caller without a GPU must fail with an error message"
