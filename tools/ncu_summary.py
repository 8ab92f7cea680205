#!/usr/bin/env python3
"""Summarise an ncu report offline: key metrics, opcode mix and stall reasons per evaluation,
hottest source lines.   tools/ncu_summary.py REPORT.ncu-rep WARP_EVALS [--lines N]"""
import collections, csv, io, re, subprocess, sys

rep, evals = sys.argv[1], float(sys.argv[2])
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25

def run(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout

raw = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr, vals = raw[0], raw[-1]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__icc_requests_lookup_hit.sum", "sm__icc_requests_lookup_miss.sum"]
for i, h in enumerate(hdr):
    if any(h.startswith(w) for w in want):
        print(f"{h:70s} {vals[i]}")

rows = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "sass"))))
h = rows[1]; H = {k: i for i, k in enumerate(h)}
st = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
byop = collections.Counter(); stall = {"fp64": collections.Counter(), "other": collections.Counter()}; cnt = collections.Counter()
for r in rows[2:]:
    try: n = int(r[H["Instructions Executed"]])
    except Exception: continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[H["Source"]].strip()); op = m.group(2)
    byop[op] += n
    k = "fp64" if op.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP") else "other"
    cnt[k] += n
    for s in st:
        v = r[H[s]]
        if v not in ("", "-"): stall[k][s[6:]] += int(v)
tot = sum(byop.values()); ts = sum(sum(c.values()) for c in stall.values())
print(f"\ninstructions per warp-evaluation: {tot / evals:.0f}  (fp64 {cnt['fp64'] / evals:.0f}, other {cnt['other'] / evals:.0f})")
print("  ".join(f"{o} {n / evals:.1f}" for o, n in byop.most_common(22)))
for k in stall:
    print(f"{k}: {100 * sum(stall[k].values()) / ts:.1f}% of samples: " + "  ".join(f"{s} {100 * v / ts:.1f}" for s, v in stall[k].most_common(8)))

rows = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "cuda,sass"))))
cur = None; agg = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] not in ("", "Line No") and r[2] == "-":
        try: agg.append((cur, int(r[0]), r[1].strip()[:80], int(r[6]), int(r[7])))
        except Exception: pass
tsamp = sum(a[3] for a in agg) or 1
agg.sort(key=lambda a: -a[3])
print()
for f, l, s, sm, ins in agg[:nlines]:
    print(f"{f:18s}:{l:4d} samp {100 * sm / tsamp:5.2f}% instr/eval {ins / evals:7.2f}  {s}")
