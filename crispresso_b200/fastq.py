"""Host side of seam S1 and of the FASTQ ingest: drop-ins for the reference's
quality-filter functions (CRISPResso/CRISPRessoCORE.py:162-310) and FASTQ counters (CORE:313-348).
gzip inflate stays on the host (zlib, on a worker thread so the next chunk inflates while the GPU
indexes the current one); record splitting runs in crgpu_fastq_index, the per-read decision (mean
phred >= q and min phred >= s) in k_qualfilter through crgpu_qualfilter.
"""
import ctypes
import gzip
import threading
import queue
import zlib

import numpy as np

from . import _lib


class FastqBatch:
    """Records of one indexed chunk: packed bases / qualities with shared offsets, and the header
    lines (without the '@') cut lazily out of the text."""

    def __init__(self, text, seq, qual, offsets, name_start, name_len):
        self.text, self.seq, self.qual, self.offsets = text, seq, qual, offsets
        self.name_start, self.name_len = name_start, name_len

    def __len__(self):
        return len(self.offsets) - 1

    def headers(self):
        t = self.text
        return [t[a + 1:a + l].decode() for a, l in zip(self.name_start.tolist(), self.name_len.tolist())]

    def seqs(self):
        b, o = self.seq.tobytes(), self.offsets.tolist()
        return [b[o[i]:o[i + 1]].decode() for i in range(len(o) - 1)]

    def quals(self):
        b, o = self.qual.tobytes(), self.offsets.tolist()
        return [b[o[i]:o[i + 1]].decode() for i in range(len(o) - 1)]


def index_text(ctx, text, final=True, counts_only=False):
    """crgpu_fastq_index over inflated FASTQ bytes.  -> (FastqBatch or None, consumed bytes, n_records, seq_bytes)"""
    text = bytes(text)
    src = np.frombuffer(text, np.uint8) if text else np.zeros(1, np.uint8)
    fo = _lib.FastqOut()
    ctx.check(ctx.lib.crgpu_fastq_index(ctx.handle, _lib.MEM_HOST, _lib.ptr(src), len(text), 1 if final else 0, ctypes.byref(fo)))
    n, total = int(fo.n_records), int(fo.seq_bytes)
    if counts_only or n == 0:
        return None, int(fo.consumed), n, total
    seq, qual = np.zeros(max(total, 1), np.uint8), np.zeros(max(total, 1), np.uint8)
    offsets, ns, nl = np.zeros(n + 1, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int32)
    fo.cap_records, fo.cap_bytes = n, max(total, 1)
    fo.seq, fo.qual, fo.offsets, fo.name_start, fo.name_len = (_lib.ptr(x) for x in (seq, qual, offsets, ns, nl))
    ctx.check(ctx.lib.crgpu_fastq_index(ctx.handle, _lib.MEM_HOST, _lib.ptr(src), len(text), 1 if final else 0, ctypes.byref(fo)))
    return FastqBatch(text, seq[:total], qual[:total], offsets, ns, nl), int(fo.consumed), n, total


def _inflate_chunks(path, chunk_bytes):
    """Generator of inflated byte chunks of a .gz (multi-member aware) or plain file."""
    if not path.endswith(".gz"):
        with open(path, "rb") as f:
            while True:
                b = f.read(chunk_bytes)
                if not b:
                    return
                yield b
    with open(path, "rb") as f:
        d = zlib.decompressobj(zlib.MAX_WBITS | 16)
        pending = b""
        while True:
            raw = pending or f.read(1 << 20)
            pending = b""
            if not raw:
                break
            out = d.decompress(raw, chunk_bytes)
            while True:
                if out:
                    yield out
                if d.eof or not d.unconsumed_tail:      # at eof the rest of the input sits in unused_data
                    break
                out = d.decompress(d.unconsumed_tail, chunk_bytes)
            if d.eof:                                   # next gzip member, if any
                pending = d.unused_data
                d = zlib.decompressobj(zlib.MAX_WBITS | 16)
        tail = d.flush()
        if tail:
            yield tail


def stream_fastq(ctx, path, chunk_bytes=64 << 20):
    """Yield FastqBatch objects for a FASTQ(.gz) file.  A worker thread inflates ahead (zlib releases the
    GIL) while the GPU indexes the current chunk; the incomplete record at the end of a chunk is carried
    over to the next one."""
    q = queue.Queue(maxsize=2)

    def worker():
        try:
            for b in _inflate_chunks(path, chunk_bytes):
                q.put(b)
            q.put(None)
        except BaseException as e:                      # surfaced in the consumer
            q.put(e)

    th = threading.Thread(target=worker, daemon=True)
    th.start()
    carry = b""
    nxt = q.get()
    while nxt is not None:
        if isinstance(nxt, BaseException):
            raise nxt
        cur, nxt = carry + nxt, q.get()
        if isinstance(nxt, BaseException):
            raise nxt
        batch, consumed, _n, _t = index_text(ctx, cur, final=nxt is None)
        carry = cur[consumed:]
        if batch is not None:
            yield batch
    th.join()
    if carry.strip():
        raise ValueError("%s ends inside a FASTQ record" % path)


def read_fastq_gpu(ctx, path, chunk_bytes=64 << 20):
    """-> (headers, seqs, quals) like read_fastq, record splitting on the GPU."""
    hs, ss, qs = [], [], []
    for b in stream_fastq(ctx, path, chunk_bytes):
        hs += b.headers(); ss += b.seqs(); qs += b.quals()
    return hs, ss, qs


def read_fastq_packed(ctx, path, chunk_bytes=64 << 20):
    """-> (headers, seq uint8, qual uint8, offsets int64[n+1]): the packed layout the C ABI consumes."""
    hs, seqs, quals, lens = [], [], [], []
    for b in stream_fastq(ctx, path, chunk_bytes):
        hs += b.headers()
        seqs.append(b.seq); quals.append(b.qual); lens.append(np.diff(b.offsets))
    offsets = np.zeros(len(hs) + 1, np.int64)
    if hs:
        offsets[1:] = np.cumsum(np.concatenate(lens))
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(1, np.uint8)
    return hs, cat(seqs), cat(quals), offsets


def get_n_reads_fastq(ctx, fastq_filename):
    """CORE:335-348 (`wc -l` // 4)."""
    n = 0
    carry = b""
    for chunk in _inflate_chunks(fastq_filename, 64 << 20):
        cur = carry + chunk
        _b, consumed, k, _t = index_text(ctx, cur, final=False, counts_only=True)
        n += k
        carry = cur[consumed:]
    if carry:
        n += index_text(ctx, carry, final=False, counts_only=True)[2]      # wc -l // 4 ignores a ragged end
    return n


def get_average_read_length_fastq(ctx, fastq_filename):
    """CORE:313-332: int(sum of read lengths / number of reads)."""
    n = total = 0
    for b in stream_fastq(ctx, fastq_filename):
        n += len(b)
        total += int(b.offsets[-1])
    return total // n


def read_fastq(path):
    """-> (headers, seqs, quals) lists of str; header without the leading '@'.  Plain host reader, kept
    for tools and tests; the drop-ins below split records on the GPU (read_fastq_gpu)."""
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
    hs, _ss, qs = read_fastq_gpu(ctx, fastq_filename)
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
    hs, ss, qs = read_fastq_gpu(ctx, fastq_filename)
    _write(output_filename, hs, ss, qs, keep_mask(ctx, qs, min_bp_quality, min_single_bp_quality))
    return output_filename


def filter_pe_fastq_by_qual(ctx, fastq_r1, fastq_r2, output_filename_r1=None, output_filename_r2=None, min_bp_quality=20,
                            min_single_bp_quality=0):
    """CORE:196-267: a pair is dropped when either mate's id is in the union of failing ids."""
    output_filename_r1 = output_filename_r1 or _default_out(fastq_r1)
    output_filename_r2 = output_filename_r2 or _default_out(fastq_r2)
    h1, s1, q1 = read_fastq_gpu(ctx, fastq_r1)
    h2, s2, q2 = read_fastq_gpu(ctx, fastq_r2)
    k1 = keep_mask(ctx, q1, min_bp_quality, min_single_bp_quality)
    k2 = keep_mask(ctx, q2, min_bp_quality, min_single_bp_quality)
    bad = set(h.split()[0] for h, k in zip(h1, k1) if not k) | set(h.split()[0] for h, k in zip(h2, k2) if not k)
    _write(output_filename_r1, h1, s1, q1, [h.split()[0] not in bad for h in h1])
    _write(output_filename_r2, h2, s2, q2, [h.split()[0] not in bad for h in h2])
    return output_filename_r1, output_filename_r2
