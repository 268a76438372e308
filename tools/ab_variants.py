"""Build kernel variants side by side and print the gpurun line that times them.

Compile-time options of the kernels (S2_UNROLL, S4_PF_DIST, DFA_MINB, RQ_MINB, ...) are tried by
building one library per variant HERE (nvcc cross-compiles; the .so files travel with the
snapshot) and selecting them on the GPU box with ACM_LIB_PATH -- one gpurun call times all of
them on the same box instead of one rebuild per call.

    python tools/ab_variants.py base: u2:-DS2_UNROLL=2 dfa4:-DDFA_MINB=4 -- python tools/quick_bench.py 1024 2k 3

builds gpu_pattern_matching_b200/libv_<name>.so for every name:flags pair (empty flags = the
default build) and prints the shell line to pass to gpurun.  `--clean` removes the variants.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gpu_pattern_matching_b200", "csrc")
PKG = os.path.join(ROOT, "gpu_pattern_matching_b200")


def clean():
    for f in os.listdir(PKG):
        if f.startswith("libv_") and f.endswith(".so"):
            os.remove(os.path.join(PKG, f))
    for d in os.listdir(CSRC):
        if d.startswith("build_v_"):
            subprocess.run(["rm", "-rf", os.path.join(CSRC, d)])


def main():
    args = sys.argv[1:]
    if args == ["--clean"]:
        clean()
        return
    if "--" not in args:
        raise SystemExit(__doc__)
    cut = args.index("--")
    variants, cmd = args[:cut], " ".join(args[cut + 1:])
    names = []
    for v in variants:
        name, _, flags = v.partition(":")
        out = f"../libv_{name}.so"
        subprocess.check_call(["make", "-s", "-C", CSRC, f"EXTRA={flags}", f"OUT={out}", f"BUILD=build_v_{name}", out])
        log = open(os.path.join(CSRC, f"build_v_{name}", "ptxas.log")).read()
        spills = [l for l in log.splitlines() if "spill" in l and not l.strip().startswith("0 bytes stack frame, 0 bytes spill stores, 0")]
        print(f"built libv_{name}.so  ({flags or 'default flags'}; {len(spills)} kernels with spills)", file=sys.stderr)
        names.append(name)
    loop = " ".join(names)
    print(f"for v in {loop}; do echo \"== $v\"; ACM_LIB_PATH=$PWD/gpu_pattern_matching_b200/libv_$v.so {cmd}; done")


if __name__ == "__main__":
    main()
