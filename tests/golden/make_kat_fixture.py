"""Build tests/golden/kat1_merged_reads.json.gz: the reference's end-to-end known-answer input
(tests/test_data/test_L001_R{1,2}_001.fastq.gz, tests/crispresso_tests.py:127-195) after the
paired-end merge CRISPResso delegates to FLASH (restated in oracle/flash_merge.py), de-duplicated
(sequence -> count).  Also tests/golden/qualfilter_subset.json.gz: every 8th record of both files
plus the three records the reference's quality-filter KAT names (tests/crispresso_tests.py:78-88).
Run in the build container only (needs /root/reference).
"""
import collections
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import fastq, flash_merge  # noqa: E402

D = "/root/reference/tests/test_data/"
KAT_IDS = ["M06879:15:000000000-DFF22:1:1101:25894:23776", "M06879:15:000000000-DFF22:1:1101:24046:20708",
           "M06879:15:000000000-DFF22:1:1102:22078:15849"]


def main():
    r1 = fastq.read_fastq(D + "test_L001_R1_001.fastq.gz")
    r2 = fastq.read_fastq(D + "test_L001_R2_001.fastq.gz")
    counts = collections.OrderedDict()
    kinds = collections.Counter()
    for (h1, s1, q1), (h2, s2, q2) in zip(r1, r2):
        m = flash_merge.merge_pair(s1, q1, s2, q2)
        if m:
            counts[m[0]] = counts.get(m[0], 0) + 1
            kinds[m[2]] += 1
    out = dict(source="test_L001_R1/R2_001.fastq.gz merged by oracle/flash_merge.py", n_pairs=len(r1),
               n_merged=sum(counts.values()), kinds=dict(kinds), reads=list(counts.keys()), counts=list(counts.values()))
    with gzip.open(os.path.join(HERE, "kat1_merged_reads.json.gz"), "wt") as f:
        json.dump(out, f)
    print("merged", out["n_merged"], "unique", len(counts), kinds)
    sub = {}
    for name, recs in (("R1", r1), ("R2", r2)):
        keep = [(h.split()[0], q) for i, (h, _s, q) in enumerate(recs) if i % 8 == 0 or h.split()[0] in KAT_IDS]
        sub[name] = dict(ids=[k[0] for k in keep], quals=[k[1] for k in keep])
    with gzip.open(os.path.join(HERE, "qualfilter_subset.json.gz"), "wt") as f:
        json.dump(sub, f)
    for fn in ("kat1_merged_reads.json.gz", "qualfilter_subset.json.gz"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
