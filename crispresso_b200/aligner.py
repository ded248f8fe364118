"""Host side of seam S2: the needle subprocess + parse_needle_output
(CRISPResso/CRISPRessoCORE.py:1791-1806, 1707-1786) replaced by crgpu_align."""
import re

import numpy as np

from . import _lib


def pack_reads(reads):
    """list of str/bytes -> (uint8 buffer, int64 offsets[n+1]) as the C ABI wants them."""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    offsets = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(1, np.uint8)
    return buf, offsets


def needle_id(fastq_header):
    """The ID parse_needle_output ends up with for a FASTQ header line (without the leading '@'):
    awk keeps the '@', sed turns every ':' into '_' (CORE:1796-1797), needle names the sequence after
    the first whitespace-delimited token, and the parser turns every '_' back into ':' (CORE:1725) --
    so underscores of the original name come back as colons too (lossy, as in the reference)."""
    return ("@" + fastq_header.split()[0]).replace("_", ":")


def parse_needle_options(options):
    """Pull -gapopen / -gapextend out of --needle_options_string (CORE:4226-4231).  Options the
    GPU aligner cannot honour are refused loudly instead of being ignored."""
    gapopen, gapextend = 10.0, 0.5
    for tok in options.split():
        m = re.match(r"^-+(\w+)(?:=(.*))?$", tok)
        if not m:
            raise ValueError("needle option not understood: %r" % tok)
        key, val = m.group(1).lower(), m.group(2)
        if key == "gapopen":
            gapopen = float(val)
        elif key == "gapextend":
            gapextend = float(val)
        elif key in ("awidth3", "auto", "stdout", "filter"):
            continue
        else:
            raise ValueError("needle option -%s is not supported by the GPU aligner" % key)
    return gapopen, gapextend


def needle_align(ctx, amplicon, reads, gapopen=10.0, gapextend=0.5, want_rows=True):
    """Align reads (list of str, or a (buffer, offsets) pair) to `amplicon`.

    Returns (recs, ref_rows, markup_rows, read_rows): recs is a structured array
    (_lib.ALN_REC); the rows are lists of str (None when want_rows is False).
    """
    buf, offsets = reads if isinstance(reads, tuple) else pack_reads(reads)
    n = len(offsets) - 1
    amp = (amplicon.encode() if isinstance(amplicon, str) else bytes(amplicon)).upper()      # (CORE:1288)
    recs = np.zeros(n, dtype=_lib.ALN_REC)
    if n == 0:
        return recs, [], [], []
    maxlb = int(np.max(np.diff(offsets)))
    slot = len(amp) + maxlb
    rows = [np.zeros(n * slot, np.uint8) for _ in range(3)] if want_rows else [None] * 3
    ctx.check(ctx.lib.crgpu_align(ctx.handle, _lib.MEM_HOST, amp, len(amp), _lib.ptr(buf), _lib.ptr(offsets), n,
                                  float(gapopen), float(gapextend), _lib.ptr(recs),
                                  _lib.ptr(rows[0]), _lib.ptr(rows[1]), _lib.ptr(rows[2]), slot))
    if not want_rows:
        return recs, None, None, None
    out = []
    for arr in rows:
        a2 = arr.reshape(n, slot)
        out.append([a2[i, recs["aln_off"][i]:].tobytes().decode() for i in range(n)])
    return recs, out[0], out[1], out[2]
