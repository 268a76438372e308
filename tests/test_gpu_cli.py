"""The ocl_aho_grep-compatible CLI (cli/b200_aho_grep) end to end: flags, -v line format,
STATS block, multi-file striping over worker threads, hex patterns, text mode."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import build_oracle, clamav_pats, load_patterns, planted_stream
from oracle_lib import materialize, read_fixture

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cli", "b200_aho_grep")
LINE = re.compile(rb"^Pattern (-?\d+) \('(.*)'\) found in file '(.*)' at offset (\d+) \[relative: (-?\d+)\]$")


def run(args):
    p = subprocess.run([CLI] + args, capture_output=True, timeout=300)
    assert p.returncode == 0, p.stderr.decode()
    return p.stdout


def stats(out):
    d = {}
    for line in out.split(b"\n"):
        m = re.match(rb"^([A-Za-z ()]+):\s+([\d.]+)$", line)
        if m:
            d[m.group(1).decode().strip()] = float(m.group(2))
    return d


def test_cli_plain_patterns_verbose(tmp_path):
    pf = materialize("kat_pat_a.txt", tmp_path)
    tf = materialize("kat_text_a.txt.gz", tmp_path)
    pats = load_patterns("kat_pat_a.txt")
    text = read_fixture("kat_text_a.txt.gz")
    out = run(["-f", tf, "-p", pf, "-v", "-B", "256", "-G", "8", "-w", "1"])     # 2 KiB buffers: many rounds
    o = build_oracle(pats)
    eo, ep, _, _ = o.search(text)
    got = [LINE.match(l).groups() for l in out.split(b"\n") if l.startswith(b"Pattern ")]
    # like the reference, the printed offset is relative to the current BUFFER (here 8 x 256 B),
    # and is the end offset + 1 (reference databuf.c:771, ocl_aho_grep.c:295)
    assert [(int(g[0]), g[1], int(g[3])) for g in got] == \
        [(pats[p][1], pats[p][0], int(e) % 2048 + 1) for e, p in zip(eo, ep)]
    assert all(int(g[4]) == (int(g[3]) - 1) % 256 + 1 for g in got)            # relative to the chunk
    assert all(g[2] == tf.encode() for g in got)
    st = stats(out)
    assert st["Matches"] == eo.size == 24 and st["Matches reported"] == 24
    assert st["Automaton states"] == 198 and st["Processed bytes"] == len(text)
    assert st["Processed files"] == 1 and st["Kernel launches"] >= 4 and st["Throughput (Mbps)"] > 0


def test_cli_hex_patterns_two_workers_many_files(tmp_path):
    pats = clamav_pats(2000)
    o = build_oracle(pats)
    pf = tmp_path / "sigs.hex"
    pf.write_bytes(b"\n".join(read_fixture("clamav_sigs_15000.hex.gz").split(b"\n")[:2000]) + b"\n")
    d = tmp_path / "in"
    d.mkdir()
    expect = 0
    for k in range(5):
        n = (1 << 18) + 4096 * k          # whole chunks: the stream a worker sees is a concatenation
        buf, _ = planted_stream(pats, n, seed=40 + k, plants=40)
        (d / f"f{k}.bin").write_bytes(buf.tobytes())
    # each worker scans its own files as one stream (thread t: files t, t+2, ...), sorted by name
    for t in range(2):
        stream = b"".join((d / f"f{k}.bin").read_bytes() for k in range(t, 5, 2))
        expect += o.search(np.frombuffer(stream, dtype=np.uint8))[0].size
    out = run(["-f", str(d), "-p", str(pf), "-x", "-w", "2", "-B", "4096", "-G", "16"])
    st = stats(out)
    assert st["Matches"] == expect and st["Processed files"] == 5
    assert st["Processed bytes"] == sum((1 << 18) + 4096 * k for k in range(5))


def test_cli_text_mode_categorical(tmp_path):
    pf = materialize("sentiment_categorical.pat.gz", tmp_path)
    tf = materialize("kat_text_readme.txt.gz", tmp_path)
    out = run(["-f", tf, "-p", pf, "-t", "-v", "-B", "512", "-G", "4096", "-w", "1"])
    st = stats(out)
    # every line is its own chunk padded with zeros: a lexicon word cannot span lines, so the
    # count equals the whole-file walk (40 matches) unless a word crosses a newline
    assert st["Matches"] == 40 and st["Processed lines"] == read_fixture("kat_text_readme.txt.gz").count(b"\n")
    ids = [int(LINE.match(l).group(1)) for l in out.split(b"\n") if l.startswith(b"Pattern ")]
    assert len(ids) == 40 and all(i != 0 for i in ids)          # categorical ids (+/- scores), not line numbers
