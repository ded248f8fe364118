"""Build tests/golden/flash_pairs_subset.json.gz: a subset of the read pairs of the reference's
end-to-end known-answer input (tests/test_data/test_L001_R{1,2}_001.fastq.gz,
tests/crispresso_tests.py:127-195) together with what oracle/flash_merge.py -- the FLASH restatement
that reproduces the reference's golden counts from these files (tests/test_kat_reference.py) -- makes
of each pair: merged sequence, merged qualities, innie/outie, or None.
Subset: every 10th pair, plus the first 250 outies, the first 150 uncombined pairs and every pair whose
best mismatch density is shared by more than one overlap (FLASH's quality tie-break).
Run in the build container only (needs /root/reference).
"""
import gzip
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import fastq, flash_merge  # noqa: E402

D = "/root/reference/tests/test_data/"


def density_ties(s1, q1, s2, q2):
    """number of candidate overlaps sharing the lowest non-zero mismatch density"""
    a = np.frombuffer(s1.encode(), np.uint8)
    b = flash_merge._COMP[np.frombuffer(s2.encode(), np.uint8)][::-1]
    dens = []
    for left, right in ((a, b), (b, a)):
        l1, l2 = len(left), len(right)
        for i in range(max(0, l1 - l2), l1 - 4 + 1):
            n = l1 - i
            x, y = left[i:], right[:n]
            valid = (x != 78) & (y != 78)
            eff = int(valid.sum())
            if eff >= 4:
                dens.append(np.float32(int((valid & (x != y)).sum())) / np.float32(min(eff, 100)))
    if not dens:
        return 0
    m = min(dens)
    return 0 if m == 0 else sum(1 for d in dens if d == m)


def main():
    r1 = fastq.read_fastq(D + "test_L001_R1_001.fastq.gz")
    r2 = fastq.read_fastq(D + "test_L001_R2_001.fastq.gz")
    pairs, n_out, n_none, n_tie = [], 0, 0, 0
    for k, ((h1, s1, q1), (h2, s2, q2)) in enumerate(zip(r1, r2)):
        m = flash_merge.merge_pair(s1, q1, s2, q2)
        take = k % 10 == 0
        if m is None and n_none < 150:
            n_none += 1
            take = True
        if m is not None and m[2] == "outie" and n_out < 250:
            n_out += 1
            take = True
        if not take and k % 3 == 0 and density_ties(s1, q1, s2, q2) > 1:
            n_tie += 1
            take = True
        if take:
            pairs.append(dict(k=k, s1=s1, q1=q1, s2=s2, q2=q2, merged=list(m) if m else None))
    out = dict(source="test_L001_R1/R2_001.fastq.gz, merged by oracle/flash_merge.py", pairs=pairs)
    fn = os.path.join(HERE, "flash_pairs_subset.json.gz")
    with gzip.open(fn, "wt") as f:
        json.dump(out, f)
    print(len(pairs), "pairs", n_out, "outies", n_none, "uncombined", n_tie, "density ties;", os.path.getsize(fn), "bytes")


if __name__ == "__main__":
    main()
