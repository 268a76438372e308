"""Summarise `ncu -i X.ncu-rep --page source --csv`: per-SASS-instruction executed counts,
stall samples and shared-memory conflicts.  usage: ncu_sass.py src.csv [top] [all|hot] [kernel name substring]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# a report with several kernels has one section per kernel: "Kernel Name",<name> / header / instructions
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
want = sys.argv[4] if len(sys.argv) > 4 else None
pick = [i for i in starts if want is None or want in rows[i][1]][-1]
end = min([i for i in starts if i > pick] + [len(rows)])
print(f"kernel: {rows[pick][1]}")
hdr = rows[pick + 1]
ci = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[pick + 2:end] if len(r) >= len(hdr) - 2]
tot_ex = sum(int(r[ci["Instructions Executed"]]) for r in body)
tot_s = sum(int(r[ci["# Samples"]]) for r in body)
print(f"instructions: {len(body)} SASS, {tot_ex} warp-level executed, {tot_s} samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ci[s]]) for r in body) for s in stalls}
print("stall samples:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
print("\n-- SASS in program order (idx, exec, samples, threads/warp, smem conflict, text) --")
for i, r in enumerate(body):
    ex = int(r[ci["Instructions Executed"]])
    sm = int(r[ci["# Samples"]])
    if ex * 200 >= tot_ex or sm * 100 >= tot_s or (len(sys.argv) > 3 and sys.argv[3] == "all"):
        top_st = sorted(((int(r[ci[s]]), s[6:]) for s in stalls), reverse=True)[:2]
        st = " ".join(f"{n}:{v}" for v, n in top_st if v)
        print(f"{i:5d} {ex:10d} {sm:6d} {r[ci['Avg. Threads Executed']]:>5s} "
              f"{r[ci['L1 Conflicts Shared N-Way']]:>5s} {r[ci['Source']].strip():60s} {st}")
