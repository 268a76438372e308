"""Scratch: file -> match-list throughput of cli/b200_aho_grep (the reference's ocl_aho_grep loop
on this library): ClamAV 10k signatures over files of seeded random bytes with planted
signatures, read from /dev/shm.  Not bench.py.

    python tools/cli_bench.py [MiB per file] [files]
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_pattern_matching_b200 import synth  # noqa: E402
from oracle_lib import clamav_signatures, read_fixture  # noqa: E402

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
nfiles = int(sys.argv[2]) if len(sys.argv) > 2 else 4
base = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
d = tempfile.mkdtemp(prefix="clibench", dir=base)
sigs = clamav_signatures(10000)
pf = os.path.join(d, "sigs.hex")
open(pf, "wb").write(b"\n".join(read_fixture("clamav_sigs_15000.hex.gz").split(b"\n")[:10000]) + b"\n")
indir = os.path.join(d, "in")
os.mkdir(indir)
expect = 0
for i in range(nfiles):
    buf = synth.stream(mib << 20, 30 + i)
    pl = synth.Plants(sigs, mib << 20, mib * 100, 30 + i)
    pl.apply_host(buf)
    expect += pl.count
    buf.tofile(os.path.join(indir, f"f{i}.bin"))
print(f"{nfiles} files x {mib} MiB in {indir}, {expect} planted", flush=True)
cli = os.path.join(ROOT, "cli", "b200_aho_grep")
for extra in sys.argv[3:] or ["-w 1", "-w 2", "-w 4", "-w 4 -G 65536", "-w 4 -G 262144", "-w 8 -G 65536"]:
    p = subprocess.run([cli, "-f", indir, "-p", pf, "-x"] + extra.split(), capture_output=True, timeout=600)
    out = p.stdout.decode()
    g = lambda k: re.search(rf"{k}:\s+([\d.]+)", out)
    if p.returncode != 0 or not g("Matches"):
        print(extra, "FAILED", p.stderr.decode()[-300:])
        continue
    secs = float(g(r"Time \(secs\)").group(1))
    by = float(g("Processed bytes").group(1))
    print(f"{extra:22s} matches {g('Matches').group(1)} time {secs:.3f} s  {by / secs / 1e9:.2f} GB/s  "
          f"launches {g('Kernel launches').group(1)}", flush=True)
    for line in p.stderr.decode().splitlines():        # ACM_DATABUF_STATS=1: where the workers' time went
        if line.startswith("databuf"):
            print("    " + line, flush=True)
subprocess.run(["rm", "-rf", d])
