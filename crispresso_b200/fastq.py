"""Host side of seam S1 and of the FASTQ ingest: drop-ins for the reference's
quality-filter functions (CRISPResso/CRISPRessoCORE.py:162-310) and FASTQ counters (CORE:313-348).
gzip inflate stays on the host (zlib: one worker thread ahead of the GPU, and the members of a multi-member
file -- bgzip / pigz -i / concatenated .gz -- on several threads at once); record splitting runs in crgpu_fastq_index, the per-read decision (mean
phred >= q and min phred >= s) in k_qualfilter through crgpu_qualfilter.
"""
import ctypes
import gzip
import os
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


INFLATE_WORKERS = min(8, os.cpu_count() or 1)


def _inflate_chunks(path, chunk_bytes, parallel=True):
    """Generator of inflated byte chunks of a .gz (multi-member aware) or plain file."""
    if not path.endswith(".gz"):
        with open(path, "rb") as f:
            while True:
                b = f.read(chunk_bytes)
                if not b:
                    return
                yield b
        return
    if parallel and INFLATE_WORKERS > 1:
        yield from _inflate_chunks_parallel(path, chunk_bytes, INFLATE_WORKERS)
        return
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


def _inflate_run(mm, start, stop, piece=1 << 16):
    """Inflate the gzip members that start in [start, stop) of the mapped file, one after the other.  -> (offset after the
    last of them, bytes), or None when no valid member starts at `start` (the magic bytes also occur inside compressed data)."""
    out, pos, n = [], start, len(mm)
    try:
        while pos < stop:
            d = zlib.decompressobj(zlib.MAX_WBITS | 16)
            while not d.eof:
                if pos >= n:
                    return None                            # truncated: not a member (a damaged file is reported by the caller)
                raw = mm[pos:pos + piece]
                out.append(d.decompress(raw))
                pos += len(raw) - len(d.unused_data)
    except zlib.error:
        return None
    return pos, b"".join(out)


def _inflate_chunks_parallel(path, chunk_bytes, workers, group_bytes=4 << 20):
    """_inflate_chunks with the gzip MEMBERS of the file inflated on `workers` threads (zlib releases the GIL): the files
    bgzip / pigz -i / bcl-convert write -- and any concatenation of .gz files -- are sequences of independent members.
    Occurrences of the gzip magic at least `group_bytes` apart are candidate starts of a RUN of members; every run is
    inflated up to the next candidate on a worker, ahead of the stitching point, and a result is used only if the run
    starts exactly where the previous one ended -- a false candidate (magic bytes inside compressed data) or a member that
    spans a candidate costs some wasted work and nothing else.  A file that is one big member (plain `gzip`) gains nothing:
    it is handed to the serial path."""
    import mmap
    from concurrent.futures import ThreadPoolExecutor
    with open(path, "rb") as f:
        try:
            mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:                                 # empty file
            return
        try:
            n = len(mm)
            cands, p = [], mm.find(b"\x1f\x8b\x08")
            while p >= 0:
                if p + 10 <= n and (mm[p + 3] & 0xE0) == 0:                  # FLG: the reserved bits are zero in a real header
                    cands.append(p)
                    p = mm.find(b"\x1f\x8b\x08", p + group_bytes)
                else:
                    p = mm.find(b"\x1f\x8b\x08", p + 1)
            if len(cands) < 3 or cands[0] != 0:
                mm.close()
                yield from _inflate_chunks(path, chunk_bytes, parallel=False)
                return
            stops = cands[1:] + [n]
            with ThreadPoolExecutor(max_workers=workers) as ex:
                futs, submitted = {}, 0
                expect, buf, size, k = 0, [], 0, 0
                while expect < n:
                    while k < len(cands) and cands[k] < expect:              # runs that started inside a consumed member
                        f2 = futs.pop(k, None)
                        if f2 is not None:
                            f2.cancel()
                        k += 1
                    while submitted < len(cands) and len(futs) < 2 * workers:
                        if submitted >= k:
                            futs[submitted] = ex.submit(_inflate_run, mm, cands[submitted], stops[submitted])
                        submitted += 1
                    if k < len(cands) and cands[k] == expect:
                        res = futs.pop(k).result() if k in futs else _inflate_run(mm, expect, stops[k])
                        k += 1
                    else:                                                    # a gap up to the next candidate: inflate it here
                        res = _inflate_run(mm, expect, cands[k] if k < len(cands) else n)
                    if res is None:
                        raise ValueError("%s: damaged gzip member at byte %d" % (path, expect))
                    expect, data = res
                    buf.append(data)
                    size += len(data)
                    if size >= chunk_bytes:
                        yield b"".join(buf)
                        buf, size = [], 0
                if buf:
                    yield b"".join(buf)
        finally:
            try:
                mm.close()
            except (BufferError, ValueError):
                pass


def stream_fastq(ctx, path, chunk_bytes=64 << 20):
    """Yield FastqBatch objects for a FASTQ(.gz) file.  A worker thread inflates ahead (zlib releases the
    GIL; multi-member files on several threads, _inflate_chunks_parallel) while the GPU indexes the current
    chunk; the incomplete record at the end of a chunk is carried over to the next one."""
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
