"""FASTQ record splitting on the B200 (crgpu_fastq_index) against a plain host parse, and the
drop-ins built on it (stream_fastq, get_n_reads_fastq, get_average_read_length_fastq, flash_merge_files)."""
import ctypes
import gzip
import os

import numpy as np
import pytest

from crispresso_b200 import _lib, fastq, flash, synth
from oracle import flash_merge

pytestmark = pytest.mark.gpu


def _records(n, seed, max_len=300, crlf=False, empty=False):
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n):
        L = int(rng.integers(0 if empty else 1, max_len + 1))
        s = synth.random_seq(rng, L)
        q = "".join(chr(int(v)) for v in rng.integers(33, 75, size=L))      # '@' and '+' occur inside qualities
        recs.append(("read%d:%d_x extra%d" % (i, seed, i), s, q))
    nl = "\r\n" if crlf else "\n"
    text = "".join("@%s%s%s%s+%s%s%s" % (h, nl, s, nl, nl, q, nl) for h, s, q in recs)
    return recs, text.encode()


def _same(batch, recs):
    assert len(batch) == len(recs)
    assert batch.headers() == [r[0] for r in recs]
    assert batch.seqs() == [r[1] for r in recs]
    assert batch.quals() == [r[2] for r in recs]


@pytest.mark.parametrize("n,seed,crlf,empty", [(1, 1, False, False), (7, 2, False, True), (5000, 3, False, False),
                                              (3000, 4, True, False), (40000, 5, False, True)])
def test_index_matches_a_host_parse(ctx, n, seed, crlf, empty):
    recs, text = _records(n, seed, crlf=crlf, empty=empty)
    batch, consumed, k, total = fastq.index_text(ctx, text)
    _same(batch, recs)
    assert consumed == len(text) and k == n and total == sum(len(r[1]) for r in recs)
    # a missing final newline is accepted on the last chunk
    batch2, consumed2, _k, _t = fastq.index_text(ctx, text[:-2] if crlf else text[:-1])
    _same(batch2, recs)
    # counts only
    assert fastq.index_text(ctx, text, counts_only=True)[1:] == (len(text), n, total)


def test_chunks_that_end_inside_a_record(ctx):
    recs, text = _records(2000, 9)
    for cut in (len(text) // 3, len(text) // 2 + 17, len(text) - 5, 3):
        b1, consumed, k, _t = fastq.index_text(ctx, text[:cut], final=False)
        assert consumed <= cut and (k == 0 or text[consumed - 1:consumed] == b"\n")
        b2, consumed2, k2, _t2 = fastq.index_text(ctx, text[consumed:], final=True)
        assert k + k2 == len(recs) and consumed + consumed2 == len(text)
        got = (b1.headers() if b1 else []) + b2.headers()
        assert got == [r[0] for r in recs]
        assert (b1.seqs() if b1 else []) + b2.seqs() == [r[1] for r in recs]


def test_malformed_text_fails_loudly(ctx):
    _recs, text = _records(100, 11)
    bad = text.replace(b"\n+\n", b"\n-\n", 1)
    with pytest.raises(_lib.CrgpuError):
        fastq.index_text(ctx, bad)
    with pytest.raises(_lib.CrgpuError):
        fastq.index_text(ctx, b"@r\nACGT\n+\nIII\n")                        # 4 bases, 3 qualities
    with pytest.raises(_lib.CrgpuError):
        fastq.index_text(ctx, b"@r\nACGT\n+\nIIII\n@r2\nAC\n")               # final chunk with half a record
    assert fastq.index_text(ctx, b"")[2] == 0


def test_streaming_a_multi_member_gzip_in_small_chunks(ctx, tmp_path):
    recs, text = _records(6000, 13)
    p = str(tmp_path / "x.fastq.gz")
    with open(p, "wb") as f:
        f.write(gzip.compress(text[:len(text) // 2]) + gzip.compress(text[len(text) // 2:]))
    hs, ss, qs = fastq.read_fastq_gpu(ctx, p, chunk_bytes=100000)
    assert (hs, ss, qs) == ([r[0] for r in recs], [r[1] for r in recs], [r[2] for r in recs])
    assert (hs, ss, qs) == fastq.read_fastq(p)
    assert fastq.get_n_reads_fastq(ctx, p) == 6000
    assert fastq.get_average_read_length_fastq(ctx, p) == sum(len(r[1]) for r in recs) // 6000
    plain = str(tmp_path / "x.fastq")
    with open(plain, "wb") as f:
        f.write(text)
    assert fastq.get_n_reads_fastq(ctx, plain) == 6000


def test_device_memory_index(ctx):
    import torch
    recs, text = _records(20000, 17)
    host, _c, n, total = fastq.index_text(ctx, text)
    d_text = torch.from_numpy(np.frombuffer(text, np.uint8).copy()).cuda()
    d_seq = torch.zeros(total, dtype=torch.uint8, device="cuda"); d_qual = torch.zeros(total, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    d_ns = torch.zeros(n, dtype=torch.int64, device="cuda"); d_nl = torch.zeros(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    fo = _lib.FastqOut()
    fo.cap_records, fo.cap_bytes = n, total
    fo.seq, fo.qual, fo.offsets, fo.name_start, fo.name_len = (t.data_ptr() for t in (d_seq, d_qual, d_off, d_ns, d_nl))
    ctx.check(ctx.lib.crgpu_fastq_index(ctx.handle, _lib.MEM_DEVICE, d_text.data_ptr(), len(text), 1, ctypes.byref(fo)))
    assert (fo.n_records, fo.seq_bytes, fo.consumed) == (n, total, len(text))
    assert np.array_equal(d_seq.cpu().numpy(), host.seq) and np.array_equal(d_qual.cpu().numpy(), host.qual)
    assert np.array_equal(d_off.cpu().numpy(), host.offsets)
    assert np.array_equal(d_ns.cpu().numpy(), host.name_start) and np.array_equal(d_nl.cpu().numpy(), host.name_len)
    # unaligned device text (a view one byte in) still indexes correctly
    d2 = torch.zeros(len(text) + 1, dtype=torch.uint8, device="cuda")
    d2[1:] = d_text
    torch.cuda.synchronize()
    ctx.check(ctx.lib.crgpu_fastq_index(ctx.handle, _lib.MEM_DEVICE, d2.data_ptr() + 1, len(text), 1, ctypes.byref(fo)))
    assert np.array_equal(d_seq.cpu().numpy(), host.seq) and np.array_equal(d_off.cpu().numpy(), host.offsets)


def test_flash_merge_files_drop_in(ctx, tmp_path):
    amp, _g, _c, _h = synth.make_case(23, 250, hdr=False)
    s1, q1, s2, q2 = synth.make_pairs(amp, 600, 150, seed=23)
    names = ["M1:%d:x 1:N:0:1" % i for i in range(600)]
    for fn, ss, qq, mate in (("r1.fastq.gz", s1, q1, "1"), ("r2.fastq.gz", s2, q2, "2")):
        with gzip.open(str(tmp_path / fn), "wt") as f:
            for h, s, q in zip(names, ss, qq):
                f.write("@%s\n%s\n+\n%s\n" % (h.replace(" 1:", " %s:" % mate), s, q))
    ext, nc1, nc2, res = flash.flash_merge_files(ctx, str(tmp_path / "r1.fastq.gz"), str(tmp_path / "r2.fastq.gz"), str(tmp_path))
    assert os.path.basename(ext) == "out.extendedFrags.fastq.gz"
    want = [flash_merge.merge_pair(*t) for t in zip(s1, q1, s2, q2)]
    hs, ss, qs = fastq.read_fastq(ext)
    assert ss == [m[0] for m in want if m] and qs == [m[1] for m in want if m]
    assert hs == [names[i] for i, m in enumerate(want) if m]
    h1, ss1, _q = fastq.read_fastq(nc1)
    h2, ss2, _q = fastq.read_fastq(nc2)
    assert ss1 == [s1[i] for i, m in enumerate(want) if not m] and ss2 == [s2[i] for i, m in enumerate(want) if not m]
    assert fastq.get_n_reads_fastq(ctx, ext) == res.n_merged
