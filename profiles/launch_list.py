#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` log.

    python profiles/launch_list.py launches.csv "header line" > launch_list.txt"""
import collections, csv, sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    ms = v / 1e6 if r[iu] in ("ns", "nsecond") else v / 1e3 if r[iu] in ("us", "usecond") else v
    name = r[ik].split("(")[0][:80]
    tot[name] += ms
    cnt[name] += 1
allms = sum(tot.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print("kernel | launches | total ms | share")
for k, v in tot.most_common():
    print("%s | %d | %.3f | %.1f%%" % (k, cnt[k], v, 100 * v / allms))
