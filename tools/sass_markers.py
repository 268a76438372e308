"""The SASS mnemonics that show which hardware paths the built library uses, per kernel, and the
registers / spills ptxas reported for it:

    python tools/sass_markers.py > profiles/sass_markers.txt

UBLKCP  cp.async.bulk global -> shared (TMA bulk copy)     UBLKPF  cp.async.bulk.prefetch.L2
SYNCS   mbarrier operations                                 LDGSTS  cp.async (per-thread global -> shared)
ACQBULK griddepcontrol.wait, PREEXIT griddepcontrol.launch_dependents (programmatic dependent launch)
MATCH / VOTE / REDUX  warp aggregation of the emitting lanes; LDL / STL  spills (none in a hot loop)
"""
import collections
import os
import re
import subprocess
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gpu_pattern_matching_b200", "libacmatch_b200.so")
LOG = os.path.join(ROOT, "gpu_pattern_matching_b200", "csrc", "build", "ptxas.log")
WANT = ("UBLKCP", "UBLKPF", "SYNCS", "LDGSTS", "ACQBULK", "PREEXIT", "MATCH", "VOTE", "REDUX", "ATOMS", "ATOMG", "RED",
        "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return {m: re.sub(r"^void ", "", re.sub(r"\(.*", "", d)) for m, d in zip(names, out)}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    counts, order, k = collections.defaultdict(collections.Counter), [], None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            k = m.group(1)
            order.append(k)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and k:
            op = m.group(1)
            if op in WANT:
                counts[k][op] += 1
            counts[k]["_instr"] += 1
    regs = {}
    if os.path.exists(LOG):
        cur = None
        for line in open(LOG):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                cur = m.group(1)
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                old = regs.setdefault(cur, {}).get("spill", (0, 0))      # the entry and its out-of-line callees: the sum
                regs[cur]["spill"] = (old[0] + int(m.group(2)), old[1] + int(m.group(3)))
            m = re.search(r"Used (\d+) registers", line)
            if m and cur:
                regs.setdefault(cur, {})["regs"] = int(m.group(1))
    names = demangle(order)
    print(f"# {os.path.basename(SO)}: {os.path.getsize(SO)} bytes, built "
          f"{time.strftime('%Y-%m-%d %H:%M UTC', time.gmtime(os.path.getmtime(SO)))}, nvcc 12.9, -gencode arch=compute_100a,code=sm_100a")
    print("# kernel | SASS instructions | registers | spill stores/loads (bytes) | mnemonic counts")
    for k in sorted(order, key=lambda x: names[x]):
        r = regs.get(k, {})
        c = counts[k]
        ops = " ".join(f"{o}={c[o]}" for o in WANT if c[o])
        print(f"{names[k]:44s} {c['_instr']:6d}  regs {r.get('regs', '?'):>3}  spill {r.get('spill', ('?', '?'))[0]}/{r.get('spill', ('?', '?'))[1]}  {ops}")


if __name__ == "__main__":
    main()
