"""FASTQ reading and the phred33 quality filter, restated -- TEST INFRASTRUCTURE (oracle/__init__.py).

get_ids_reads_to_remove / filter_se_fastq_by_qual (CRISPResso/CRISPRessoCORE.py:162-193, 270-310)
use Biopython + np.mean; the same decision in exact integers is sum(phred) < q*len or
min(phred) < s.  Pinned by the reference's own KAT tests/crispresso_tests.py:78-88.
"""
import gzip


def read_fastq(path):
    """-> list of (header_without_@, sequence, quality)"""
    op = gzip.open if path.endswith(".gz") else open
    out = []
    with op(path, "rt") as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip("\n")
            f.readline()
            q = f.readline().rstrip("\n")
            out.append((h.rstrip("\n")[1:], s, q))
    return out


def keep_read(qual, min_bp_quality, min_single_bp_quality):
    ph = [ord(c) - 33 for c in qual]
    if not ph:
        return False
    return sum(ph) >= min_bp_quality * len(ph) and min(ph) >= min_single_bp_quality


def ids_to_remove(records, min_bp_quality=20, min_single_bp_quality=0):
    """record.id of Biopython = first whitespace-delimited token of the header (CORE:191)."""
    return set(h.split()[0] for h, _s, q in records if not keep_read(q, min_bp_quality, min_single_bp_quality))
