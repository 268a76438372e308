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


def test_flow_grep_ushort_application(tmp_path):
    """AC_ushorts application layer: signature file `tokens;len;details`, flow files named
    src_sport_dst_dport_proto, one alert per (flow, signature) occurrence, flows independent."""
    from oracle_lib import Oracle, fixture_path
    sig_path = fixture_path("ushort_signatures.txt")
    sig_lines = [l for l in read_fixture("ushort_signatures.txt").decode().splitlines() if l]
    o = Oracle(2048)
    for k, line in enumerate(sig_lines):
        o.add_csv(line.split(";")[0], k)
    o.compile()
    d = tmp_path / "flows"
    d.mkdir()
    flows = {"10.19.1.5_333_152.29.9.15_443_tcp": read_fixture("ushort_flow_333.txt"),
             "10.19.1.5_666_152.29.9.115_443_tcp": read_fixture("ushort_flow_666.txt"),
             # a signature split over the end of one flow and the start of the next must NOT match
             "10.0.0.1_1_10.0.0.2_2_udp": b"5,5,5,666\n",
             "10.0.0.1_1_10.0.0.3_2_udp": b"676,1,2\n3,9\n"}
    expect = {}
    for name, data in flows.items():
        (d / name).write_bytes(data)
        toks = np.array([int(x) for x in re.split(rb"[,\s]+", data) if x], dtype=np.uint16)
        off, pat, _, _ = o.search(toks)
        expect[name] = sorted(pat.tolist())
    out = subprocess.run([os.path.join(ROOT, "cli", "b200_flow_grep"), "-p", sig_path, "-f", str(d), "-v"],
                         capture_output=True, timeout=120)
    assert out.returncode == 0, out.stderr.decode()
    got = {n: [] for n in flows}
    for line in out.stdout.decode().splitlines():
        m = re.match(r"^date: \d{4}-\d\d-\d\d, time: \d\d:\d\d:\d\d, signature id: (\d+), signature pattern: "
                     r"'([\d,]+)', signature length: (\d+), signature details: '(.*)', source ip: (\S+), "
                     r"source port: (\S+), destination ip: (\S+), destination port: (\S+), protocol: (\S+) $", line)
        if m:
            sid = int(m.group(1))
            assert m.group(2) == sig_lines[sid].split(";")[0] and int(m.group(3)) == int(sig_lines[sid].split(";")[1])
            got["_".join(m.group(i) for i in range(5, 10))].append(sid)
    assert {k: sorted(v) for k, v in got.items()} == expect
    assert expect["10.0.0.1_1_10.0.0.2_2_udp"] == [] and expect["10.0.0.1_1_10.0.0.3_2_udp"] == [2]
    assert sum(len(v) for v in expect.values()) >= 3
    assert f"Alerts:              {sum(len(v) for v in expect.values())}".encode() in out.stdout


def test_cli_follow_mode_sees_appended_data(tmp_path):
    """-F (reference ocl_aho_grep.c:60-141): the tool keeps reading what is appended to the file
    until SIGINT; a match whose bytes arrive in two instalments is reported once (the carry
    between buffers is the last Lmax-1 bytes)."""
    import signal
    import time
    pf = materialize("kat_pat_a.txt", tmp_path)
    pats = load_patterns("kat_pat_a.txt")
    text = bytearray(read_fixture("kat_text_a.txt.gz"))
    # an occurrence that straddles the cut; the cut is a whole number of 256-byte chunks -- like the
    # reference (databuf.c:327-400), a partial last chunk is zero-padded before it is scanned, so only
    # a read that ends on a chunk boundary can be continued
    pat = max((p for p, _ in pats), key=len)
    cut = 512
    text[cut - len(pat) // 2:cut - len(pat) // 2 + len(pat)] = pat
    text = bytes(text)
    o = build_oracle(pats)
    eo, ep, _, _ = o.search(text)
    assert any(e >= cut > e + 1 - len(pats[i][0]) for e, i in zip(eo.tolist(), ep.tolist()))
    f = tmp_path / "growing.txt"
    f.write_bytes(text[:cut])
    p = subprocess.Popen([CLI, "-f", str(f), "-p", pf, "-v", "-F", "-w", "1", "-B", "256", "-G", "8"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    def wait_until_read(upto, timeout=120.0):
        """until the tool's file offset on the growing file has reached `upto` (/proc/<pid>/fdinfo)"""
        t0 = time.time()
        while time.time() - t0 < timeout:
            assert p.poll() is None, p.stderr.read().decode()
            try:
                for fd in os.listdir(f"/proc/{p.pid}/fd"):
                    if os.path.realpath(f"/proc/{p.pid}/fd/{fd}") == os.path.realpath(str(f)):
                        pos = int(re.search(r"pos:\s+(\d+)", open(f"/proc/{p.pid}/fdinfo/{fd}").read()).group(1))
                        if pos >= upto:
                            return
            except (OSError, AttributeError):
                pass
            time.sleep(0.05)
        raise AssertionError(f"the tool did not read {upto} bytes within {timeout} s")

    try:
        wait_until_read(cut)                          # device init + first pass over the file
        time.sleep(1.0)                               # ... and its buffer processed (follow sleeps 20 ms per pass)
        with open(f, "ab") as fh:
            fh.write(text[cut:])
        wait_until_read(len(text))
        time.sleep(0.3)
        p.send_signal(signal.SIGINT)
        out, err = p.communicate(timeout=60)
    finally:
        if p.poll() is None:
            p.kill()
    assert p.returncode == 0, err.decode()
    got = sorted((int(m.group(1)), m.group(2)) for m in (LINE.match(l) for l in out.split(b"\n")) if m)
    assert got == sorted((pats[i][1], pats[i][0]) for i in ep)
    st = stats(out)
    assert st["Matches"] == eo.size and st["Processed bytes"] == len(text)


def test_cli_device_list_stripes_workers(tmp_path):
    """-D a,b,...: worker t opens the (t mod n)-th device of the list (reference ocl_aho_grep.c:498-502:
    every worker owns a context on its device).  The same GPU twice on a single-GPU box."""
    import torch
    ngpu = max(1, torch.cuda.device_count())
    pats = clamav_pats(2000)
    o = build_oracle(pats)
    pf = tmp_path / "sigs.hex"
    pf.write_bytes(b"\n".join(read_fixture("clamav_sigs_15000.hex.gz").split(b"\n")[:2000]) + b"\n")
    d = tmp_path / "in"
    d.mkdir()
    expect = 0
    for k in range(4):
        buf, _ = planted_stream(pats, 1 << 18, seed=60 + k, plants=30)
        (d / f"f{k}.bin").write_bytes(buf.tobytes())
        expect += o.search(buf)[0].size                   # per-file semantics: every file on its own
    devs = ",".join(str(i % ngpu) for i in range(2))
    out = run(["-f", str(d), "-p", str(pf), "-x", "-w", "4", "-D", devs, "-B", "4096", "-G", "16"])
    st = stats(out)
    assert st["Matches"] == expect and st["Processed files"] == 4
