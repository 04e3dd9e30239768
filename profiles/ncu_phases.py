#!/usr/bin/env python
"""Stall samples and instructions of one kernel in an .ncu-rep (--import-source on), bucketed by source-line ranges
("phases"), plus the lines with the most stall samples and their dominant stall reasons.

    python profiles/ncu_phases.py rep.ncu-rep file.cuh name:lo-hi [name:lo-hi ...]
"""
import collections, csv, io, subprocess, sys

rep, srcfile = sys.argv[1], sys.argv[2]
ranges = []
for a in sys.argv[3:]:
    name, r = a.split(":")
    lo, hi = r.split("-")
    ranges.append((name, int(lo), int(hi)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None
lines = []
fname = None


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and r and r[0].isdigit():
        lines.append((fname, int(r[0]), r[1].strip()[:90], dict(zip(hdr, r))))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def phase(f, ln):
    if f != srcfile:
        return "inlined:" + f
    for name, lo, hi in ranges:
        if lo <= ln <= hi:
            return name
    return "other"


tot = sum(num(l[3]["# Samples"]) for l in lines)
ti = sum(num(l[3]["Instructions Executed"]) for l in lines)
ph, phi, phs = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for f, ln, src, d in lines:
    p = phase(f, ln)
    ph[p] += num(d["# Samples"])
    phi[p] += num(d["Instructions Executed"])
    for s in stalls:
        phs[p][s] += num(d[s])
print("samples %d, instructions %d" % (tot, ti))
for k, v in ph.most_common():
    top = ", ".join("%s %.0f%%" % (s[6:], 100.0 * c / max(v, 1)) for s, c in phs[k].most_common(4))
    print("%-22s samples %5.1f%%  instr %5.1f%%   [%s]" % (k, 100.0 * v / tot, 100.0 * phi[k] / ti, top))
print()
for f, ln, src, d in sorted(lines, key=lambda l: -num(l[3]["# Samples"]))[:24]:
    top = ", ".join("%s %d" % (s[6:], num(d[s])) for s in sorted(stalls, key=lambda s: -num(d[s]))[:2])
    print("%-14s:%4d %5.2f%%s %5.2f%%i [%s] | %s" % (f, ln, 100.0 * num(d["# Samples"]) / tot, 100.0 * num(d["Instructions Executed"]) / ti, top, src[:70]))
