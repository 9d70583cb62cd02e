"""Instruction mix of one kernel from an ncu report's source page (per-SASS-instruction executed counts and stall samples):

    ncu -i report.ncu-rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > sass.csv
    python scripts/sass_mix.py sass.csv
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
ops, samples = collections.Counter(), collections.Counter()
tot = tots = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    src = r[ix["Source"]].strip()
    try:
        n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    except ValueError:
        continue
    toks = src.split()
    op = toks[0] if not toks[0].startswith("@") else toks[1]
    op = op.split(".")[0] if not op.startswith("MUFU") else op
    ops[op] += n; samples[op] += s; tot += n; tots += s
print(rows[0][1][:100] if len(rows[0]) > 1 else "")
print("warp instructions executed %d, stall samples %d" % (tot, tots))
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 36):
    print("%-14s %13d %5.1f %%   samples %5.1f %%" % (op, n, 100 * n / tot, 100 * samples[op] / max(tots, 1)))
fp64 = sum(n for op, n in ops.items() if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("FP64-pipe share of the executed instructions: %.1f %%" % (100 * fp64 / tot))
