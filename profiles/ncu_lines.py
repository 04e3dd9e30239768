#!/usr/bin/env python
"""Per-source-line and per-opcode instruction shares of one kernel in an .ncu-rep (needs --import-source on).

    python profiles/ncu_lines.py rep.ncu-rep [units] [top]

`units` (optional) divides the instruction counts, e.g. the number of warp-level (i,j) pairs of the launch,
so the table reads "instructions per unit"."""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
lines, ops, samp_ops = [], collections.Counter(), collections.Counter()
fname = hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        ie, iss = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if not hdr or len(r) <= max(ie, iss):
        continue
    try:
        e, s = int(r[ie]), int(r[iss])
    except ValueError:
        continue
    if r[0].isdigit():
        lines.append((fname, int(r[0]), r[1].strip()[:100], e, s))
    elif r[0] == "" and r[3].strip():
        t = r[3].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += e
        samp_ops[op] += s
tot = sum(l[3] for l in lines)
ts = max(1, sum(l[4] for l in lines))
print("instructions executed %d (%.1f per unit), stall samples %d" % (tot, tot / units, ts))
print("\n-- opcodes")
for op, c in ops.most_common(22):
    print("%-10s %7.1f/unit %6.2f%% instr %6.2f%% samples" % (op, c / units, 100.0 * c / tot, 100.0 * samp_ops[op] / ts))
print("\n-- source lines")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%-14s:%4d %7.1f/unit %5.2f%%i %5.2f%%s | %s" % (l[0], l[1], l[3] / units, 100.0 * l[3] / tot, 100.0 * l[4] / ts, l[2]))
