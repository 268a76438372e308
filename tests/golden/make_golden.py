"""Generates tests/golden/kat.json from the REFERENCE's own builder (oracle/_ref, compiled
from /root/reference/acsmx.c and AC_ushorts/iacsmx.c by oracle/ref_build) plus the serial
walk in ref_driver.c.  Run in the build container (needs oracle/_ref/libacref.so):

    python tests/golden/make_golden.py

The vectors pin the CPU oracle (tests/test_oracle.py) on machines where the reference
tree, and therefore oracle/_ref, may be absent.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import (RefAcsm, RefIacsm, clamav_signatures, parse_pattern_file,  # noqa: E402
                        read_fixture)

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from gpu_pattern_matching_b200 import synth  # noqa: E402


def digest(off, pat):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(off, dtype=np.uint64).tobytes())
    h.update(np.ascontiguousarray(pat, dtype=np.uint32).tobytes())
    return h.hexdigest()


def table_digest(r):
    """sha256 over the defined cells of h_trans (second half only where the first is negative)."""
    t = r.h_trans()
    a = t[:, :256]
    b = np.where(a < 0, t[:, 256:], 0)
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(a).tobytes())
    h.update(np.ascontiguousarray(b).tobytes())
    return h.hexdigest()


def build(pats):
    r = RefAcsm()
    for p, iid in pats:
        r.add(p, iid)
    r.compile()
    return r


def case(name, pats, text, full_table=True):
    r = build(pats)
    off, pat, hits, fin = r.search(text)
    b, idx = r.ml_csr()
    sizes = np.diff(b)
    out = {
        "name": name, "patterns": len(pats), "states": r.num_states,
        "max_pattern_len": r.max_pattern_len, "table_bytes": int(r.table_bytes),
        "text_bytes": int(len(text)), "hits": hits, "matches": int(off.size),
        "final_state": fin, "digest": digest(off, pat),
        "first": [[int(o), int(p)] for o, p in zip(off[:16], pat[:16])],
        "final_states": int((sizes > 0).sum()), "multi_pattern_finals": int((sizes > 1).sum()),
        "max_list": int(sizes.max()),
    }
    if full_table:
        out["table_digest"] = table_digest(r)
    r.close()
    return out


def main():
    cases = []
    fx = [("kat_pat_a.txt", "kat_text_a.txt.gz"), ("kat_pat_b.txt", "kat_text_b.txt.gz"),
          ("kat_pat_c.txt", "kat_text_a.txt.gz"), ("kat_pat_two_words.txt", "kat_pat_categorical_small.txt"),
          ("sentiment_categorical.pat.gz", "kat_text_a.txt.gz"),
          ("sentiment_categorical.pat.gz", "kat_text_readme.txt.gz")]
    for pf, tf in fx:
        cases.append(case(f"{pf} x {tf}", parse_pattern_file(read_fixture(pf)), read_fixture(tf)))
    hand = [(b"abc", 0), (b"bc", 1), (b"c", 2), (b"abc", 3), (b"xbc", 4), (b"ab", 5)]
    cases.append(case("hand x zabcxbcab", hand, b"zabcxbcab"))
    # ClamAV sets over a seeded planted stream (config 1 stand-in, scaled to 4 MiB)
    for n in (2000, 10000, 15000):
        sigs = clamav_signatures(n)
        pats = [(s, i) for i, s in enumerate(sigs)]
        nbytes = 4 << 20
        buf = synth.stream(nbytes, 7)
        synth.Plants(sigs, nbytes, 512, 7, forced=[(0, 3), (nbytes - len(sigs[5]), 5)]).apply_host(buf)
        cases.append(case(f"clamav{n} x planted 4MiB seed 7", pats, buf))
    # ushort twin
    ir = RefIacsm()
    for k, line in enumerate(read_fixture("ushort_signatures.txt").decode().splitlines()):
        ir.add_csv(line.split(";")[0], k)
    ir.compile()
    toks = np.array([9, 8, 7, 6, 5, 4, 3, 2, 1, 0, 1, 2, 3, 4, 666, 676], dtype=np.uint16)
    off, iid, fin = ir.search(toks)
    ushort = {"states": ir.num_states, "max_pattern_len": ir.max_pattern_len,
              "matches": [[int(o), int(i)] for o, i in zip(off, iid)], "final_state": fin}
    json.dump({"generator": "tests/golden/make_golden.py (oracle/_ref = reference acsmx.c/iacsmx.c)",
               "cases": cases, "ushort": ushort}, open(os.path.join(HERE, "kat.json"), "w"), indent=1)
    for c in cases:
        print(c["name"], c["states"], c["hits"], c["matches"], c["first"][:1])
    print("ushort", ushort)


if __name__ == "__main__":
    main()
