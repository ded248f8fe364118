"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total ms, share.
usage: launch_shares.py launches.csv > profiles/rNN_launch_shares.md"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0] != "ID"]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"^void ", "", r[4].split("(")[0]).replace("crgpu::", "")
    name = re.sub(r"\(int\)", "", name)
    c = agg.setdefault(name[:60], [0, 0.0])
    c[0] += 1
    c[1] += float(r[14]) / 1e6
hot = sum(t for n, (c, t) in agg.items() if not n.startswith(("k_int_peak", "at::")))
print("| kernel | launches | total ms | share of hot-path kernel time |")
print("|---|---|---|---|")
for n, (c, t) in agg.items():
    share = "(peak probe)" if n.startswith("k_int_peak") else "%.1f %%" % (100 * t / hot)
    print("| %s | %d | %.3f | %s |" % (n, c, t, share))
