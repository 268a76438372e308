#!/bin/bash
# The CPU test suite against the sanitizer build of the library (make -C gpu_pattern_matching_b200/csrc asan):
# AddressSanitizer + UBSan over the builder, the databuf bookkeeping and the C-ABI argument checks.
# Python itself is not instrumented, so the runtime is preloaded and leak checking is off.
set -e
cd "$(dirname "$0")/.."
make -C gpu_pattern_matching_b200/csrc asan >/dev/null
export ACM_LIB_PATH=$PWD/gpu_pattern_matching_b200/libacmatch_b200_asan.so
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=1:protect_shadow_gap=0
# the compiled reference (oracle/_ref) stays out: AddressSanitizer stops in ITS iacsm_add_pattern, a known overflow
export ACM_SKIP_REF=1
export UBSAN_OPTIONS=halt_on_error=1:print_stacktrace=1
exec python -m pytest tests/ -x -q -m "not gpu" "$@"
