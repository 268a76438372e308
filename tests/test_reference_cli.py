"""The drop-in claim as a test: the reference's OWN command-line tool -- /root/reference/
ocl_aho_grep.c + file_traverse.c, unmodified, compiled where they lie by oracle/ref_build/Makefile
against this repo's include/ (CL/opencl.h is a shim) and linked with libacmatch_b200.so -- builds,
and on a GPU produces the oracle's matches.  The binary (oracle/_ref/ocl_aho_grep, git-ignored,
travels to the GPU box) is test infrastructure: no reference source is copied into the repo.

Also: the way reference apps/sentiment_analysis.py:188-198 drives the tool (same command line,
same parsing of the -v lines), with cli/b200_aho_grep behind the name ./ocl_aho_grep.
"""
import os
import re
import shlex
import subprocess

import numpy as np
import pytest

from helpers import build_oracle, load_patterns
from oracle_lib import materialize, read_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
REFCLI = os.path.join(ROOT, "oracle", "_ref", "ocl_aho_grep")
CLI = os.path.join(ROOT, "cli", "b200_aho_grep")
# binary signatures are printed with %s: the quoted pattern may hold any byte, newlines included
LINE = re.compile(rb"Pattern (-?\d+) \('(.*?)'\) found in file '([^'\n]*)' at offset (\d+) \[relative: (-?\d+)\]\n", re.S)


def _matches(out):
    return [m.groups() for m in LINE.finditer(out)]


def test_reference_cli_builds_unmodified_against_include():
    """CPU: compile + link the reference's main program against include/ and the library."""
    if not os.path.exists(os.path.join(REF, "ocl_aho_grep.c")):
        pytest.skip("reference tree absent (GPU box): the prebuilt oracle/_ref/ocl_aho_grep is used")
    if os.path.exists(REFCLI):
        os.unlink(REFCLI)
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle", "ref_build"), f"REF={REF}", "refcli"])
    assert os.path.exists(REFCLI)
    # it is the reference's program: its usage text, its option letters
    p = subprocess.run([REFCLI, "-h"], capture_output=True, timeout=60)
    assert b"ocl_aho_grep -f file -p file -B chunk_size -D devpos" in p.stdout
    # every undefined symbol of the reference's objects is resolved by libacmatch_b200.so
    ldd = subprocess.run(["ldd", REFCLI], capture_output=True, text=True).stdout
    assert "libacmatch_b200.so" in ldd and "not found" not in ldd and "OpenCL" not in ldd


@pytest.mark.gpu
def test_reference_cli_output_equals_oracle(tmp_path):
    """GPU: the reference's tests/patterns.txt x tests/input.txt (our fixtures kat_pat_a /
    kat_text_a are byte-identical copies of them): 24 matches, first at offset 85."""
    if not os.path.exists(REFCLI):
        pytest.fail("oracle/_ref/ocl_aho_grep missing: run __graft_entry__.build() where /root/reference exists")
    pf = materialize("kat_pat_a.txt", tmp_path)
    tf = materialize("kat_text_a.txt.gz", tmp_path)
    pats = load_patterns("kat_pat_a.txt")
    text = read_fixture("kat_text_a.txt.gz")
    o = build_oracle(pats)
    eo, ep, _, _ = o.search(text)
    assert eo.size == 24 and int(eo[0]) == 85
    for w in ("1", "2"):
        p = subprocess.run([REFCLI, "-f", tf, "-p", pf, "-B", "4096", "-D", "0", "-G", "8192", "-L", "1024",
                            "-w", w, "-v"], capture_output=True, timeout=300)
        # (the reference declares `void main`: its exit status is whatever is left in the register)
        assert b"ERROR" not in p.stderr and b"-------------- STATS" in p.stdout, p.stderr.decode()
        got = _matches(p.stdout)
        # callback_match (reference ocl_aho_grep.c:272-308) prints iid, pattern, file, end offset + 1
        assert [(int(g[0]), g[1], int(g[3])) for g in got] == \
            [(pats[i][1], pats[i][0], int(e) + 1) for e, i in zip(eo, ep)]
        assert all(g[2] == tf.encode() for g in got)
        assert re.search(rb"Matches:\s+24\n", p.stdout) and re.search(rb"Automaton states:\s+198\n", p.stdout)


@pytest.mark.gpu
def test_reference_cli_hex_signatures_many_buffers(tmp_path):
    """GPU: ClamAV 2000 (-x) over a planted 3 MiB stream in 256 KiB buffers, matches across buffers."""
    if not os.path.exists(REFCLI):
        pytest.fail("oracle/_ref/ocl_aho_grep missing")
    from helpers import clamav_pats, planted_stream
    pats = clamav_pats(2000)
    o = build_oracle(pats)
    pf = tmp_path / "sigs.hex"
    pf.write_bytes(b"\n".join(read_fixture("clamav_sigs_15000.hex.gz").split(b"\n")[:2000]) + b"\n")
    n = 12 * 4096 * 64
    forced = [(4096 * 64 - 20, 31), (2 * 4096 * 64 - 1, 32), (4096 - 3, 33)]
    buf, _ = planted_stream(pats, n, seed=23, plants=300, forced=forced)
    tf = tmp_path / "stream.bin"
    tf.write_bytes(buf.tobytes())
    eo, ep, _, _ = o.search(buf)
    p = subprocess.run([REFCLI, "-f", str(tf), "-p", str(pf), "-x", "-B", "4096", "-D", "0", "-G", "64", "-L", "256",
                        "-w", "1", "-v"], capture_output=True, timeout=300)
    assert b"ERROR" not in p.stderr and b"-------------- STATS" in p.stdout, p.stderr.decode()
    got = _matches(p.stdout)
    size = 4096 * 64
    assert [int(g[0]) for g in got] == [pats[i][1] for i in ep]
    assert [int(g[3]) for g in got] == [int(e) % size + 1 for e in eo]     # offsets are per buffer
    m = re.search(rb"Matches:\s+(\d+)", p.stdout)
    assert m and int(m.group(1)) == eo.size


@pytest.mark.gpu
def test_sentiment_app_drives_the_cli_unchanged(tmp_path):
    """What reference apps/sentiment_analysis.py:188-198 does: write `ID " word "` lines to
    patterns.txt, spawn `./ocl_aho_grep -p patterns.txt -f <dir> -B 4096 -D 0 -L 1024 -G 8192 -w 1
    -v <extra>` and read `Pattern <id> ...` lines.  Here ./ocl_aho_grep is a link to
    cli/b200_aho_grep; the ids parsed the app's way must be the oracle's."""
    cwd = tmp_path
    os.symlink(CLI, cwd / "ocl_aho_grep")
    words = [(b"good", 1), (b"great", 2), (b"bad", -1), (b"awful", -2), (b"love", 3), (b"hate", -3)]
    with open(cwd / "patterns.txt", "wb") as f:
        for w, i in words:
            f.write(str(i).encode() + b' " ' + w + b' "\n')              # the app's own format (:73, :86)
    text = (b"i love this great phone , the camera is good but the battery is bad \n"
            b"awful service , i hate waiting ; great great day \n") * 50
    (cwd / "in").mkdir()
    (cwd / "in" / "tweets.txt").write_bytes(text)
    pats = load_patterns_from(cwd / "patterns.txt")
    o = build_oracle(pats)
    eo, ep, _, _ = o.search(np.frombuffer(text, dtype=np.uint8))
    command = "./ocl_aho_grep -p patterns.txt -f " + "in" + "  -B 4096 -D 0 -L 1024 -G 8192 -w 1  -v " + "-t"
    process = subprocess.Popen(shlex.split(command), stdout=subprocess.PIPE, cwd=cwd)
    pids = []
    while True:
        output = process.stdout.readline()
        if not output and process.poll() is not None:
            break
        output = str(output, "utf-8")
        if output.find("Pattern") == 0:                                     # :195
            pid = output.split()[1]
            pid = pid.replace("#", "")
            pids.append(int(pid))
    assert process.returncode == 0
    # text mode scans line by line (16-byte aligned, zero padded): same matches, same order
    assert pids == [pats[i][1] for i in ep]
    assert any(p < 0 for p in pids) and any(p > 0 for p in pids) and len(pids) == 50 * 7
    # (the script itself lives only in /root/reference, which does not exist on the GPU box, and
    # this container has no GPU: what is tested here is its side of the protocol, line for line)


def load_patterns_from(path):
    from oracle_lib import parse_pattern_file
    return parse_pattern_file(open(path, "rb").read(), False)
