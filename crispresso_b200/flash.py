"""Host side of the paired-end merge: the `flash` subprocess of
CRISPResso/CRISPRessoCORE.py:1655-1677 replaced by crgpu_flash_merge.

`flash_merge_files` is the drop-in for the call site: same inputs (the two FASTQ files, the overlap
bounds of args.min/max_paired_end_reads_overlap, the output directory) and the same files FLASH writes
with `-z -d <dir>` that CRISPResso goes on to use: out.extendedFrags.fastq.gz (CORE:1677) and
out.notCombined_{1,2}.fastq.gz (CORE:1674-1675).  FASTQ parsing and gzip stay on the host.
"""
import gzip
import os

import numpy as np

from . import _lib
from .aligner import pack_reads
from .fastq import read_fastq_packed


class MergeResult:
    """kind[n] (0 not combined, 1 innie, 2 outie), pos[n], and the merged reads as the packed
    (seq, offsets) pair the aligner consumes, with their qualities and pair indices."""

    def __init__(self, kind, pos, seq, qual, offsets, index):
        self.kind, self.pos, self.seq, self.qual, self.offsets, self.index = kind, pos, seq, qual, offsets, index

    @property
    def n_merged(self):
        return len(self.index)

    def reads(self):
        s, o = self.seq.tobytes(), self.offsets
        return [s[o[j]:o[j + 1]].decode() for j in range(self.n_merged)]

    def quals(self):
        s, o = self.qual.tobytes(), self.offsets
        return [s[o[j]:o[j + 1]].decode() for j in range(self.n_merged)]


def merge_packed(ctx, seq1, qual1, off1, seq2, qual2, off2, min_overlap=4, max_overlap=100, max_mismatch_density=0.25,
                 allow_outies=True):
    """Packed host arrays in, MergeResult out (crgpu_flash_merge, CRGPU_MEM_HOST)."""
    n = len(off1) - 1
    assert len(off2) - 1 == n, "the two FASTQ files hold different numbers of reads"
    cap = int(off1[-1] + off2[-1]) + 1
    kind = np.zeros(max(n, 1), np.uint8)
    pos = np.zeros(max(n, 1), np.int32)
    seq = np.zeros(cap, np.uint8)
    qual = np.zeros(cap, np.uint8)
    offsets = np.zeros(n + 1, np.int64)
    index = np.zeros(max(n, 1), np.int32)
    prm = _lib.MergeParams(int(min_overlap), int(max_overlap), float(max_mismatch_density), 1 if allow_outies else 0)
    out = _lib.MergeOut()
    out.pos, out.kind, out.seq, out.qual = _lib.ptr(pos), _lib.ptr(kind), _lib.ptr(seq), _lib.ptr(qual)
    out.offsets, out.index, out.cap_bytes, out.cap_reads = _lib.ptr(offsets), _lib.ptr(index), cap, n
    import ctypes
    ctx.check(ctx.lib.crgpu_flash_merge(ctx.handle, _lib.MEM_HOST, _lib.ptr(seq1), _lib.ptr(qual1), _lib.ptr(off1),
                                        _lib.ptr(seq2), _lib.ptr(qual2), _lib.ptr(off2), n, ctypes.byref(prm),
                                        ctypes.byref(out)))
    m = int(out.n_merged)
    return MergeResult(kind[:n], pos[:n], seq[:int(out.bytes)], qual[:int(out.bytes)], offsets[:m + 1], index[:m])


def merge_pairs(ctx, seqs1, quals1, seqs2, quals2, **kw):
    """Lists of str in, MergeResult out."""
    s1, o1 = pack_reads(seqs1)
    q1, _ = pack_reads(quals1)
    s2, o2 = pack_reads(seqs2)
    q2, _ = pack_reads(quals2)
    return merge_packed(ctx, s1, q1, o1, s2, q2, o2, **kw)


def combined_tag(h1, h2):
    """FLASH names the merged read after read 1; when the two tags differ and read 1's has a '/', the
    '/1' mate suffix (everything from the last '/') is dropped."""
    if h1 != h2 and "/" in h1:
        return h1[:h1.rindex("/")]
    return h1


def flash_merge_files(ctx, fastq_r1, fastq_r2, output_directory, min_overlap=4, max_overlap=100, allow_outies=True,
                      max_mismatch_density=0.25):
    """Drop-in for `flash R1 R2 --allow-outies --max-overlap M --min-overlap m -z -d DIR` (CORE:1657-1664).
    Returns (extendedFrags path, notCombined_1 path, notCombined_2 path, MergeResult)."""
    h1, s1, q1, o1 = read_fastq_packed(ctx, fastq_r1)
    h2, s2, q2, o2 = read_fastq_packed(ctx, fastq_r2)
    res = merge_packed(ctx, s1, q1, o1, s2, q2, o2, min_overlap=min_overlap, max_overlap=max_overlap,
                       max_mismatch_density=max_mismatch_density, allow_outies=allow_outies)
    ext = os.path.join(output_directory, "out.extendedFrags.fastq.gz")
    nc1 = os.path.join(output_directory, "out.notCombined_1.fastq.gz")
    nc2 = os.path.join(output_directory, "out.notCombined_2.fastq.gz")
    with gzip.open(ext, "wt") as f:
        for p, s, q in zip(res.index, res.reads(), res.quals()):
            f.write("@%s\n%s\n+\n%s\n" % (combined_tag(h1[p], h2[p]), s, q))

    def rec(h, s, q, o, p):
        return "@%s\n%s\n+\n%s\n" % (h[p], s[o[p]:o[p + 1]].tobytes().decode(), q[o[p]:o[p + 1]].tobytes().decode())

    with gzip.open(nc1, "wt") as f1, gzip.open(nc2, "wt") as f2:
        for p in np.nonzero(res.kind == 0)[0]:
            f1.write(rec(h1, s1, q1, o1, p))
            f2.write(rec(h2, s2, q2, o2, p))
    return ext, nc1, nc2, res
