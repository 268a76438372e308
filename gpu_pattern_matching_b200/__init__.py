"""gpu_pattern_matching_b200 -- B200-native Aho-Corasick multi-pattern matching behind
the C host API of gvasilious/gpu_pattern_matching.

The product is libacmatch_b200.so (csrc/: plain-C host side + hand-written sm_100a
kernels, C ABI in include/*.h).  These modules are thin ctypes mirrors of that ABI:

    acsm      Acsm, Iacsm           builder interface (acsmx.h, iacsmx.h)
    matcher   Device, Scanner       native scan API (acm.h)
    sharded   ShardedScan, StepPipeline   one process per GPU, torch.distributed plumbing
    synth     stream, Plants        synthetic inputs for tests and bench
    _lib      lib(), structs        the ctypes prototypes of every entry point of include/*.h,
                                    among them the databuf / ocl_worker / ocl_aho_match path
                                    (databuf.h, common.h), which tests and bench.py call directly
"""
from ._lib import AcmError, LIB_PATH, lib  # noqa: F401
from .acsm import Acsm, Iacsm  # noqa: F401
from .matcher import (Device, Scanner, MODE_AUTO, MODE_CDFA, MODE_DFA, MODE_SAMPLED4,  # noqa: F401
                      MODE_START2, MODE_NAMES)
