"""Executed instructions per CUDA source line from an `ncu --set full --import-source on` report.

    ncu -i prof.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:fit_disp_kernel \
        --launch-count 1 > lines.csv
    python scripts/source_lines.py lines.csv "title" > profiles/rNN_source_lines.txt

The cuda,sass view lists an inlined instruction under its own line AND under the line of every call site it was inlined
into, so the shares of a callee line and of its call site overlap (they are not additive across inlining levels); the
opcode mix is therefore computed from the innermost attribution only (the first time an address is seen)."""
import collections
import csv
import sys

FP = ("DFMA", "DMUL", "DADD", "DSETP")


def main(path, title):
    rows = list(csv.reader(open(path)))
    cur_file, hdr, cur = None, None, None
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    seen, ops = set(), collections.Counter()
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) >= 2 and r[0] == "Function Name":
            continue
        if r and r[0] == "Line No":
            hdr = r
            i_e, i_s = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) < len(hdr) - 5:
            continue
        if r[0] != "":
            cur = (cur_file, int(r[0]), r[1].strip()[:90])
            continue
        if cur is None:
            continue
        try:
            e, s = int(r[i_e]), int(r[i_s])
        except ValueError:
            continue
        toks = r[3].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        a = agg[cur]
        a[0] += e; a[1] += s; a[2][op] += e
        if r[2] not in seen:
            seen.add(r[2])
            ops[op] += e
    tot = sum(ops.values())
    tot_s = sum(a[1] for a in agg.values())
    print("# %s" % title)
    print("# executed warp instructions per CUDA source line (ncu --page source --print-source cuda,sass); a line inside an")
    print("# inlined function and the line of its call site both carry the same instructions, so shares overlap across")
    print("# inlining levels.  instr%% is relative to the %d warp instructions the kernel executed." % tot)
    print("#")
    print("# opcode mix: " + ", ".join("%s %.1f%%" % (o, 100 * c / tot) for o, c in ops.most_common(16)))
    print("# FP64 arithmetic (DFMA+DMUL+DADD+DSETP): %.1f%%" % (100 * sum(ops[o] for o in FP) / tot))
    print("#")
    print("%-26s %5s %7s %7s %7s  %s" % ("file", "line", "instr%", "stall%", "nonFP%", "source  [largest non-FP64 opcodes, % of all instructions]"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
        nonfp = sum(c for o, c in a[2].items() if o not in FP)
        top = ", ".join("%s %.1f" % (o, 100 * c / tot) for o, c in a[2].most_common(5) if o not in FP and 100 * c / tot >= 0.05)
        print("%-26s %5d %7.2f %7.2f %7.2f  %s   [%s]" % (k[0], k[1], 100 * a[0] / tot, 100 * a[1] / tot_s, 100 * nonfp / tot, k[2][:70], top))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
