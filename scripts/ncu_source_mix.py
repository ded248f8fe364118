"""Opcode mix (executed warp instructions) per kernel from an `ncu --page source --csv` dump, and the stall reasons per issue
from the matching `--page raw --csv` dump.  usage: ncu_source_mix.py source.csv raw.csv"""
import collections
import csv
import sys

src = list(csv.reader(open(sys.argv[1])))
raw = list(csv.reader(open(sys.argv[2])))
hdr = raw[0]
starts = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
seen = set()
stall_cols = [(i, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
              for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
names_raw = [r[hdr.index("Kernel Name")] for r in raw[2:]]
for bi, st in enumerate(starts):
    name = src[st][1].replace("crgpu::", "").replace("(int)", "").replace("void ", "").split("(")[0]
    if name in seen:
        continue
    seen.add(name)
    en = starts[bi + 1] if bi + 1 < len(starts) else len(src)
    blk = src[st + 2:en]
    ops = collections.Counter()
    tot = 0
    for r in blk:
        ins = r[1].split()
        op = ins[1] if ins and ins[0].startswith("@") and len(ins) > 1 else (ins[0] if ins else "?")
        n = int(r[5])
        ops[op] += n
        tot += n
    mix = ", ".join("%s %.1f %%" % (o, 100.0 * c / tot) for o, c in ops.most_common(8))
    stall = ""
    for k, rn in enumerate(names_raw):
        if rn.replace("crgpu::", "").replace("(int)", "").replace("void ", "").split("(")[0] == name:
            row = raw[2 + k]
            vals = []
            for i, h in stall_cols:
                try:
                    vals.append((float(row[i]), h))
                except ValueError:
                    pass
            s = sum(v for v, _ in vals)
            stall = "; stall reasons (share of warp states): " + ", ".join("%s %.0f %%" % (h, 100 * v / s) for v, h in sorted(vals, reverse=True)[:6])
            break
    print("* `%s`: %.1f M warp instructions; %s%s" % (name, tot / 1e6, mix, stall))
