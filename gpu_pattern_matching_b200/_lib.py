"""ctypes binding of libacmatch_b200.so (the C ABI declared in include/*.h).

The library is built in-tree by `make -C gpu_pattern_matching_b200/csrc` (see
__graft_entry__.build).  There is no Python or CPU fallback: if the shared
library is missing, importing anything that needs it raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACM_LIB_PATH") or os.path.join(_HERE, "libacmatch_b200.so")   # override: kernel experiments

u8p = C.POINTER(C.c_ubyte)
u16p = C.POINTER(C.c_ushort)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
vp = C.c_void_p


class AcmError(RuntimeError):
    pass


class ScanParams(C.Structure):
    _fields_ = [("mode", C.c_int), ("bucket_shift", C.c_int), ("bucket_cap", C.c_int),
                ("timing", C.c_int), ("dfa_chunk", C.c_int), ("own_stream", C.c_int), ("reserved", C.c_int * 2)]


class ScanResult(C.Structure):
    _fields_ = [("n_matches", C.c_uint64), ("n_bytes", C.c_uint64), ("mode", C.c_int),
                ("fallback", C.c_int), ("final_state", C.c_uint32), ("n_buckets", C.c_uint32),
                ("ms_scan", C.c_float), ("ms_prefix", C.c_float), ("ms_compact", C.c_float),
                ("ms_total", C.c_float), ("launches", C.c_uint32), ("reserved", C.c_uint32)]


class PushTarget(C.Structure):
    _fields_ = [("d_dst", C.c_void_p), ("cap", C.c_uint64), ("key_add", C.c_uint64)]


class AcsmPattern(C.Structure):
    pass


AcsmPattern._fields_ = [("next", C.POINTER(AcsmPattern)), ("pattern", u8p), ("casepattern", u8p),
                        ("n", C.c_int), ("nocase", C.c_int), ("offset", C.c_int), ("depth", C.c_int),
                        ("id", vp), ("iid", C.c_int), ("index", C.c_uint)]


class AcsmStruct(C.Structure):
    _fields_ = [("max_states", C.c_int), ("num_states", C.c_int), ("max_pattern_len", C.c_int),
                ("size", C.c_size_t), ("patterns", vp), ("num_patterns", C.c_int),
                ("state_table", vp), ("h_trans", i32p), ("d_trans", vp), ("priv", vp)]


class IacsmStruct(C.Structure):
    _fields_ = [("max_states", C.c_int), ("num_states", C.c_int), ("max_pattern_len", C.c_int),
                ("size", C.c_size_t), ("patterns", vp), ("state_table", vp), ("h_trans", i32p),
                ("d_trans", vp), ("priv", vp)]


class Clconf(C.Structure):
    _fields_ = [("platform", vp), ("dev", vp), ("ctx", vp), ("queue", vp),
                ("program_aho_match", vp), ("kernel_aho_match", vp), ("program_prefixsum", vp),
                ("kernel_prescan", vp), ("kernel_prescan_store_sum", vp),
                ("kernel_prescan_store_sum_non_power_of_two", vp),
                ("kernel_prescan_non_power_of_two", vp), ("kernel_uniform_add", vp),
                ("program_compact_array", vp), ("kernel_compact_array", vp), ("type", C.c_uint64)]


class Databuf(C.Structure):
    _fields_ = [("h_data", u8p), ("h_indices", i32p), ("h_sizes", i32p), ("h_results", i32p),
                ("h_results2", i32p), ("h_prefixsum", i32p), ("h_results_comp", i32p),
                ("h_results2_comp", i32p), ("results_comp_size", C.c_size_t),
                ("results2_comp_size", C.c_size_t), ("file_ids", i32p), ("mapped", C.c_int),
                ("max_results", C.c_int), ("last_state", C.c_long), ("max_chunks", C.c_size_t),
                ("max_chunk_size", C.c_size_t), ("size", C.c_size_t), ("chunks", C.c_size_t),
                ("bytes", C.c_size_t), ("d_data", vp), ("d_indices", vp), ("d_sizes", vp),
                ("d_results", vp), ("d_results2", vp), ("d_prefixsum", vp), ("d_results_comp", vp),
                ("d_results2_comp", vp), ("p_data", vp), ("p_indices", vp), ("p_sizes", vp),
                ("p_results", vp), ("p_results2", vp), ("p_prefixsum", vp), ("p_results_comp", vp),
                ("p_results2_comp", vp), ("ScanPartialSums", vp), ("ScanPartialSums_size", C.c_uint),
                ("cl", C.POINTER(Clconf)), ("priv", vp)]


class WorkerCtx(C.Structure):
    _fields_ = [("id", C.c_int), ("text_mode", C.c_int), ("follow", C.c_int), ("verbose", C.c_int),
                ("thread_no", C.c_int), ("total_files", C.c_int), ("fds", C.POINTER(C.c_int)),
                ("filenames", C.POINTER(C.c_char_p)), ("matches_total", C.c_size_t),
                ("matches_reported", C.c_size_t), ("bytes", C.c_size_t), ("lines", C.c_size_t),
                ("rounds", C.c_size_t), ("global_ws", C.c_size_t), ("local_ws", C.c_size_t),
                ("cl", Clconf), ("db", C.POINTER(Databuf)), ("acsm", C.POINTER(AcsmStruct)),
                ("patterns", C.POINTER(AcsmPattern)), ("patterns_size", C.c_size_t)]


MATCH_CB = C.CFUNCTYPE(C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp)

# every symbol include/*.h declares: name -> (restype, argtypes)
SIGNATURES = {
    # acm.h
    "acm_last_error": (C.c_char_p, []),
    "acm_device_count": (C.c_int, []),
    "acm_device_open": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "acm_device_close": (None, [vp]),
    "acm_device_ordinal": (C.c_int, [vp]),
    "acm_device_stream": (vp, [vp]),
    "acm_device_set_stream": (C.c_int, [vp, vp]),
    "acm_device_sync": (C.c_int, [vp]),
    "acm_default_device": (vp, []),
    "acm_dev_alloc": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "acm_dev_free": (None, [vp, vp]),
    "acm_host_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "acm_host_alloc_pinned_near": (C.c_int, [vp, C.c_size_t, C.POINTER(vp)]),
    "acm_host_free_pinned": (None, [vp]),
    "acm_dev_memset": (C.c_int, [vp, vp, C.c_int, C.c_size_t]),
    "acm_memcpy_h2d": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "acm_memcpy_d2h": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "acm_memcpy_d2h_side": (C.c_int, [vp, vp, vp, C.c_size_t]),
    "acm_side_sync": (C.c_int, [vp]),
    "acm_memcpy_d2h_segments": (C.c_int, [vp, vp, C.POINTER(vp), u64p, C.c_uint32]),
    "acm_memcpy_d2h_segments_async": (C.c_int, [vp, vp, C.POINTER(vp), u64p, C.c_uint32]),
    "acm_automaton_upload": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "acm_automaton_free": (None, [vp]),
    "acm_automaton_states": (C.c_uint32, [vp]),
    "acm_automaton_patterns": (C.c_uint32, [vp]),
    "acm_automaton_pattern_lengths": (u32p, [vp]),
    "acm_automaton_max_pattern_len": (C.c_int, [vp]),
    "acm_automaton_min_pattern_len": (C.c_int, [vp]),
    "acm_automaton_alphabet": (C.c_int, [vp]),
    "acm_automaton_device_bytes": (C.c_size_t, [vp]),
    "acm_automaton_default_mode": (C.c_int, [vp]),
    "acm_automaton_gram_count": (C.c_uint32, [vp]),
    "acm_automaton_cdfa_classes": (C.c_int, [vp]),
    "acm_automaton_sample_stride": (C.c_int, [vp]),
    "acm_automaton_split_len": (C.c_int, [vp]),
    "acm_scanner_create": (C.c_int, [vp, vp, C.c_uint64, C.POINTER(ScanParams), C.POINTER(vp)]),
    "acm_scanner_free": (None, [vp]),
    "acm_scan_device": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(ScanResult)]),
    "acm_scan_device_ex": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                     C.POINTER(ScanResult)]),
    "acm_scan_device_async": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
                                        C.POINTER(PushTarget)]),
    "acm_scan_finish": (C.c_int, [vp, C.POINTER(ScanResult)]),
    "acm_scan_keys": (vp, [vp]),
    "acm_scanner_stream": (vp, [vp]),
    "acm_scan_fetch": (C.c_int64, [vp, C.c_uint64, u64p, u32p, C.c_uint64]),
    "acm_scan_histogram": (C.c_int, [vp, vp]),
    "acm_scan_trace": (C.c_int, [vp, u64p, C.c_uint32]),
    "acm_ipc_export": (C.c_int, [vp, vp, vp]),
    "acm_ipc_open": (C.c_int, [vp, vp, C.POINTER(vp)]),
    "acm_ipc_close": (C.c_int, [vp, vp]),
    "acm_scan_push_keys": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64]),
    "acm_scan_host": (C.c_int64, [vp, vp, C.c_uint64, C.c_uint64, u64p, u32p, C.c_uint64,
                                  C.POINTER(ScanResult)]),
    "acm_scan_host_ex": (C.c_int64, [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64, u64p, u32p, C.c_uint64,
                                     C.POINTER(ScanResult)]),
    "acm_multi_open": (C.c_int, [vp, C.POINTER(C.c_int), C.c_int, C.c_uint64, C.POINTER(ScanParams), C.POINTER(vp)]),
    "acm_multi_close": (None, [vp]),
    "acm_multi_devices": (C.c_int, [vp]),
    "acm_multi_device": (vp, [vp, C.c_int]),
    "acm_multi_shard": (None, [vp, C.c_uint64, C.c_int, u64p, u64p, u64p]),
    "acm_multi_scan_device": (C.c_int64, [vp, C.POINTER(vp), C.c_uint64, u64p, C.c_uint64, u64p,
                                          C.POINTER(ScanResult)]),
    "acm_multi_scan_host": (C.c_int64, [vp, vp, C.c_uint64, C.c_uint64, u64p, u32p, C.c_uint64,
                                        C.POINTER(ScanResult)]),
    "acm_exclusive_scan_u32": (C.c_int, [vp, vp, vp, C.c_uint32, vp]),
    "acm_compact_columns_i32": (C.c_int, [vp, vp, vp, vp, C.c_int32, C.c_int32, C.c_int64]),
    "acm_radix_sort_u64": (C.c_int, [vp, vp, vp, C.c_uint64, C.c_int, C.c_int, C.c_int]),
    "acm_sort_pairs_u32": (C.c_int, [vp, vp, vp, vp, vp, C.c_uint32, C.c_int]),
    "acm_synth_fill_device": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, C.c_uint64]),
    "acm_synth_fill_host": (None, [vp, C.c_uint64, C.c_uint64, C.c_uint64]),
    "acm_plant_device": (C.c_int, [vp, vp, C.c_uint64, C.c_uint64, u64p, u32p, u32p, C.c_uint32,
                                   u8p, C.c_uint32]),
    # acsmx.h
    "acsm_new": (C.POINTER(AcsmStruct), []),
    "acsm_add_pattern": (None, [C.POINTER(AcsmStruct), C.c_char_p, C.c_int, C.c_int, C.c_int,
                                C.c_int, vp, C.c_int]),
    "acsm_compile": (None, [C.POINTER(AcsmStruct)]),
    "acsm_gen_state_table": (None, [C.POINTER(AcsmStruct), C.c_int, vp, vp]),
    "acsm_get_patterns_table": (C.POINTER(AcsmPattern), [C.POINTER(AcsmStruct)]),
    "acsm_free_patterns_table": (None, [C.POINTER(AcsmPattern), C.c_int]),
    "acsm_get_max_pattern_size": (C.c_int, [C.POINTER(AcsmStruct)]),
    "acsm_get_min_pattern_size": (C.c_int, [C.POINTER(AcsmStruct)]),
    "acsm_get_states": (C.c_int, [C.POINTER(AcsmStruct)]),
    "acsm_get_size": (C.c_size_t, [C.POINTER(AcsmStruct)]),
    "acsm_cleanup": (None, [C.POINTER(AcsmStruct)]),
    "acsm_free": (None, [C.POINTER(AcsmStruct)]),
    "acsm_status": (C.c_int, [C.POINTER(AcsmStruct)]),
    "acsm_export_ref_table": (C.c_int, [C.POINTER(AcsmStruct)]),
    "acsm_check_filters": (C.c_int, [C.POINTER(AcsmStruct)]),
    "acsm_check_cdfa": (C.c_int, [C.POINTER(AcsmStruct), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    "acsm_check_xd": (C.c_int, [C.POINTER(AcsmStruct), C.POINTER(C.c_uint)]),
    "acsm_tables": (vp, [C.POINTER(AcsmStruct)]),
    "acsm_device_automaton": (vp, [C.POINTER(AcsmStruct)]),
    # iacsmx.h
    "iacsm_new": (C.POINTER(IacsmStruct), []),
    "iacsm_add_pattern": (None, [C.POINTER(IacsmStruct), u16p, C.c_int, C.c_int, C.c_int, vp, C.c_int]),
    "iacsm_add_fullpattern": (None, [C.POINTER(IacsmStruct), C.c_char_p, C.c_int]),
    "iacsm_compile": (None, [C.POINTER(IacsmStruct)]),
    "iacsm_gen_state_table": (None, [C.POINTER(IacsmStruct), C.c_int, vp, vp]),
    "iacsm_get_max_pattern_size": (C.c_int, [C.POINTER(IacsmStruct)]),
    "iacsm_get_states": (C.c_int, [C.POINTER(IacsmStruct)]),
    "iacsm_get_size": (C.c_size_t, [C.POINTER(IacsmStruct)]),
    "iacsm_cleanup": (None, [C.POINTER(IacsmStruct)]),
    "iacsm_free": (None, [C.POINTER(IacsmStruct)]),
    "iacsm_status": (C.c_int, [C.POINTER(IacsmStruct)]),
    "iacsm_export_ref_table": (C.c_int, [C.POINTER(IacsmStruct)]),
    "iacsm_check_xd": (C.c_int, [C.POINTER(IacsmStruct), C.POINTER(C.c_uint)]),
    "iacsm_device_automaton": (vp, [C.POINTER(IacsmStruct)]),
    # ocl_context.h
    "clinitctx": (None, [C.POINTER(Clconf), C.c_int, C.c_int]),
    "clfreectx": (None, [C.POINTER(Clconf)]),
    # databuf.h
    "databuf_new": (C.POINTER(Databuf), [C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.POINTER(Clconf)]),
    "databuf_add_fd": (C.c_int, [C.POINTER(Databuf), C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "databuf_add_fp": (C.c_int, [C.POINTER(Databuf), vp, C.c_int, C.c_int, C.POINTER(C.c_size_t),
                                 C.POINTER(C.c_size_t)]),
    "databuf_add_chunk": (C.c_int, [C.POINTER(Databuf), C.c_char_p, C.c_size_t, C.c_int, C.c_char]),
    "databuf_reset": (None, [C.POINTER(Databuf)]),
    "databuf_clear": (None, [C.POINTER(Databuf)]),
    "databuf_copy_host_to_device": (None, [C.POINTER(Databuf), vp]),
    "databuf_copy_device_to_host": (None, [C.POINTER(Databuf), vp]),
    "databuf_process_results": (C.c_int, [C.POINTER(Databuf), MATCH_CB, vp]),
    "databuf_free": (None, [C.POINTER(Databuf), C.c_int, vp]),
    "databuf_status": (C.c_int, [C.POINTER(Databuf)]),
    "databuf_match_count": (C.c_size_t, [C.POINTER(Databuf)]),
    "databuf_alloc_postpass": (C.c_int, [C.POINTER(Databuf)]),
    "databuf_set_file_semantics": (None, [C.POINTER(Databuf), C.c_int]),
    "databuf_read_fd": (C.c_long, [C.c_int, vp, C.c_size_t]),
    # ocl_aho_match.h
    "ocl_aho_match_init": (None, [C.POINTER(Clconf)]),
    "ocl_aho_match_close": (None, [C.POINTER(Clconf)]),
    "ocl_aho_match": (None, [C.POINTER(Clconf), C.POINTER(Databuf), C.POINTER(AcsmStruct),
                             C.c_size_t, C.c_int]),
    "ocl_aho_match_ushort": (None, [C.POINTER(Clconf), C.POINTER(Databuf), C.POINTER(IacsmStruct),
                                    C.c_size_t]),
    # ocl_prefix_sum.h / ocl_compact_array.h / ocl_bitonic_sort.h
    "ocl_prefix_sum_init": (None, [C.POINTER(Clconf)]),
    "ocl_prefix_sum_close": (None, [C.POINTER(Clconf)]),
    "ocl_prefix_sum": (None, [C.POINTER(Clconf), C.POINTER(Databuf), C.c_uint]),
    "ocl_compact_array_init": (None, [C.POINTER(Clconf)]),
    "ocl_compact_array_close": (None, [C.POINTER(Clconf)]),
    "ocl_compact_array": (None, [C.POINTER(Clconf), C.POINTER(Databuf), C.c_size_t]),
    "ocl_bitonic_sort_init": (C.c_int, [C.POINTER(Clconf)]),
    "ocl_bitonic_sort_close": (C.c_int, [C.POINTER(Clconf)]),
    "ocl_bitonic_sort": (C.c_int, [C.POINTER(Clconf), vp, vp, vp, vp, C.c_uint, C.c_uint, C.c_uint]),
    # ocl_worker.h
    "ocl_worker_ctx_create": (C.POINTER(WorkerCtx), [C.c_int]),
    "ocl_worker_ctx_init": (C.c_int, [C.POINTER(WorkerCtx), C.c_int, C.c_size_t, C.c_size_t, C.c_int,
                                      C.c_char_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.POINTER(C.c_int), C.POINTER(C.c_char_p)]),
    "ocl_worker_ctx_free": (None, [C.POINTER(WorkerCtx)]),
    # utils.h
    "printable_hex_to_bytes": (vp, [C.c_char_p]),
    "gettime": (C.c_size_t, []),
    "acsm_load_pattern_file": (C.c_int, [C.POINTER(AcsmStruct), C.c_char_p, C.c_int, C.c_int]),
}

_lib = None


def lib():
    """Load libacmatch_b200.so; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AcmError(
                f"{LIB_PATH} is missing: build it with `make -C gpu_pattern_matching_b200/csrc` "
                "(or __graft_entry__.build()).  There is no fallback implementation.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)      # AttributeError = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error():
    e = lib().acm_last_error()
    return e.decode(errors="replace") if e else ""


def check(rc, what=""):
    if rc is not None and rc < 0:
        raise AcmError(f"{what or 'libacmatch_b200'} failed ({rc}): {last_error()}")
    return rc
