"""Host side of seam S1: drop-ins for the reference's quality-filter functions
(CRISPResso/CRISPRessoCORE.py:162-310).  FASTQ parsing and gzip stay on the host; the per-read
decision (mean phred >= q and min phred >= s) runs in k_qualfilter through crgpu_qualfilter.
"""
import gzip

import numpy as np

from . import _lib


def read_fastq(path):
    """-> (headers, seqs, quals) lists of str; header without the leading '@'."""
    op = gzip.open if path.endswith(".gz") else open
    hs, ss, qs = [], [], []
    with op(path, "rt") as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().rstrip("\n")
            f.readline()
            q = f.readline().rstrip("\n")
            hs.append(h.rstrip("\n")[1:])
            ss.append(s)
            qs.append(q)
    return hs, ss, qs


def keep_mask(ctx, quals, min_bp_quality=20, min_single_bp_quality=0):
    """uint8 mask, 1 = read passes (crgpu_qualfilter)."""
    from .aligner import pack_reads
    n = len(quals)
    keep = np.zeros(n, np.uint8)
    if n == 0:
        return keep
    buf, off = pack_reads(quals)
    ctx.check(ctx.lib.crgpu_qualfilter(ctx.handle, _lib.MEM_HOST, _lib.ptr(buf), _lib.ptr(off), n, int(min_bp_quality),
                                       int(min_single_bp_quality), _lib.ptr(keep)))
    return keep


def get_ids_reads_to_remove(ctx, fastq_filename, min_bp_quality=20, min_single_bp_quality=0):
    """CORE:162-193: ids (first header token) of reads below the thresholds, as a set."""
    hs, _ss, qs = read_fastq(fastq_filename)
    keep = keep_mask(ctx, qs, min_bp_quality, min_single_bp_quality)
    return set(h.split()[0] for h, k in zip(hs, keep) if not k)


def _default_out(name):
    return name.replace(".fastq", "").replace(".gz", "") + "_filtered.fastq.gz"


def _write(path, hs, ss, qs, keep):
    with gzip.open(path, "wt") as out:
        for h, s, q, k in zip(hs, ss, qs, keep):
            if k:
                out.write("@%s\n%s\n+\n%s\n" % (h, s, q))


def filter_se_fastq_by_qual(ctx, fastq_filename, output_filename=None, min_bp_quality=20, min_single_bp_quality=0):
    """CORE:270-310.  Returns the output filename."""
    output_filename = output_filename or _default_out(fastq_filename)
    hs, ss, qs = read_fastq(fastq_filename)
    _write(output_filename, hs, ss, qs, keep_mask(ctx, qs, min_bp_quality, min_single_bp_quality))
    return output_filename


def filter_pe_fastq_by_qual(ctx, fastq_r1, fastq_r2, output_filename_r1=None, output_filename_r2=None, min_bp_quality=20,
                            min_single_bp_quality=0):
    """CORE:196-267: a pair is dropped when either mate's id is in the union of failing ids."""
    output_filename_r1 = output_filename_r1 or _default_out(fastq_r1)
    output_filename_r2 = output_filename_r2 or _default_out(fastq_r2)
    h1, s1, q1 = read_fastq(fastq_r1)
    h2, s2, q2 = read_fastq(fastq_r2)
    k1 = keep_mask(ctx, q1, min_bp_quality, min_single_bp_quality)
    k2 = keep_mask(ctx, q2, min_bp_quality, min_single_bp_quality)
    bad = set(h.split()[0] for h, k in zip(h1, k1) if not k) | set(h.split()[0] for h, k in zip(h2, k2) if not k)
    _write(output_filename_r1, h1, s1, q1, [h.split()[0] not in bad for h in h1])
    _write(output_filename_r2, h2, s2, q2, [h.split()[0] not in bad for h in h2])
    return output_filename_r1, output_filename_r2
