"""Behaviour at the edges of the C ABI: lower-case amplicons, row buffers without a slot, several
contexts (one per device when the box has more than one) with tiles whose profile needs the opt-in shared-memory size,
the caller's current device.  Needs a B200."""
import ctypes

import numpy as np
import pytest
import torch

from crispresso_b200 import Context, _lib, aligner, hotpath, synth
from crispresso_b200._lib import CrgpuError

pytestmark = pytest.mark.gpu


def test_lower_case_amplicon_gives_the_upper_case_answer(ctx):
    """The reference upper-cases the amplicon first (CORE:1288); the C entry points do the same, so a caller that did not
    still gets upper-case reference rows and the same quantification."""
    amp, guide, cut, hdr = synth.make_case(77, 180)
    packed = synth.make_reads(amp, hdr, cut, 800, seed=77, rc_frac=0.05)
    n = len(packed[1]) - 1
    slot = len(amp) + int(np.diff(packed[1]).max())
    outs = []
    for a in (amp, amp.lower()):
        recs = np.zeros(n, dtype=_lib.ALN_REC)
        rows = [np.zeros(n * slot, dtype=np.uint8) for _ in range(3)]
        ctx.check(ctx.lib.crgpu_align(ctx.handle, _lib.MEM_HOST, a.encode(), len(a), _lib.ptr(packed[0]), _lib.ptr(packed[1]), n,
                                      10.0, 0.5, _lib.ptr(recs), _lib.ptr(rows[0]), _lib.ptr(rows[1]), _lib.ptr(rows[2]), slot))
        outs.append((recs, rows))
    assert np.array_equal(outs[0][0], outs[1][0])
    for k in range(3):
        assert np.array_equal(outs[0][1][k], outs[1][1][k])
    inc = hotpath.include_mask(180, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    a = hotpath.run_hot_path(ctx, amp, packed, hdr_amplicon=hdr, flags=hotpath.quant_flags(hdr), inc=inc, want_rows=True)
    b = hotpath.run_hot_path(ctx, amp.lower(), packed, hdr_amplicon=hdr.lower(), flags=hotpath.quant_flags(hdr), inc=inc, want_rows=True)
    assert np.array_equal(a.red.results(), b.red.results()) and np.array_equal(a.recs, b.recs)
    for k in range(3):
        assert np.array_equal(a.rows[k], b.rows[k])


def test_row_buffers_without_a_slot_are_refused_before_any_work(ctx):
    """CRGPU_MEM_DEVICE with row buffers and out->slot == 0: the library must not pick a slot the caller never agreed to."""
    amp, guide, cut, hdr = synth.make_case(78, 150)
    buf, off = synth.make_reads(amp, None, cut, 200, seed=78)
    n = len(off) - 1
    d_buf, d_off = torch.from_numpy(buf).cuda(), torch.from_numpy(off).cuda()
    kept = torch.zeros(n, dtype=torch.uint8, device="cuda")
    aln = torch.zeros(n * _lib.ALN_REC.itemsize, dtype=torch.uint8, device="cuda")
    recs = torch.zeros(n * _lib.READ_REC.itemsize, dtype=torch.uint8, device="cuda")
    rows = [torch.zeros(n * 400, dtype=torch.uint8, device="cuda") for _ in range(3)]
    torch.cuda.synchronize()
    red = hotpath.Reductions(150)
    pp = _lib.PathParams(gapopen=10.0, gapextend=0.5, min_identity_score=60.0, hdr_amplicon=None, hdr_amplicon_len=0, rc_rescue=1)
    inc = np.ones(150, dtype=np.uint8)
    qp = _lib.QuantParams(amplicon_len=150, flags=0, hdr_perfect_alignment_threshold=98.0, include_mask=inc.ctypes.data)
    po = _lib.PathOut()
    po.kept, po.aln, po.recs = kept.data_ptr(), aln.data_ptr(), recs.data_ptr()
    po.ref_rows, po.mark_rows, po.qry_rows = (r.data_ptr() for r in rows)
    po.slot = 0
    po.vectors, po.hist_inframe, po.hist_frameshift = red.vectors.ctypes.data, red.hist_inframe.ctypes.data, red.hist_frameshift.ctypes.data
    po.hist_len, po.hist_zero, po.counters = hotpath.HIST_LEN, hotpath.HIST_ZERO, red.counters.ctypes.data
    rc = ctx.lib.crgpu_align_quantify(ctx.handle, _lib.MEM_DEVICE, amp.encode(), 150, ctypes.byref(pp), ctypes.byref(qp),
                                      d_buf.data_ptr(), d_off.data_ptr(), n, ctypes.byref(po))
    assert rc == _lib.E_ARG
    assert b"slot" in ctx.lib.crgpu_last_error(ctx.handle)
    assert int(rows[0].sum().item()) == 0                      # nothing was written


def test_several_contexts_with_large_tiles():
    """A 500-bp amplicon uses the (16,32) tile, whose profile table (57.6 KB) needs the opt-in dynamic shared-memory size:
    every context -- on every device of the box -- must configure its own launches."""
    amp, guide, cut, hdr = synth.make_case(79, 500)
    packed = synth.make_reads(amp, hdr, cut, 300, seed=79, read_len=500)
    inc = hotpath.include_mask(500, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    devices = list(range(min(torch.cuda.device_count(), 2))) * 2
    ctxs = [Context(d) for d in devices]
    try:
        res = [hotpath.run_hot_path(c, amp, packed, hdr_amplicon=hdr, flags=hotpath.quant_flags(hdr), inc=inc) for c in ctxs]
        for r in res[1:]:
            assert np.array_equal(r.red.results(), res[0].red.results()) and np.array_equal(r.recs, res[0].recs)
    finally:
        for c in ctxs:
            c.close()


def test_the_callers_current_device_is_left_alone():
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU on this box")
    torch.cuda.set_device(0)
    c = Context(1)
    try:
        amp, guide, cut, hdr = synth.make_case(80, 120)
        aligner.needle_align(c, amp, synth.make_reads(amp, None, cut, 50, seed=80))
        assert torch.cuda.current_device() == 0
        x = torch.ones(4, device="cuda")
        assert x.device.index == 0
    finally:
        c.close()


def test_iupac_bases_are_reported_per_read(ctx):
    """A read with a base outside ACGTN(U) must not fail the whole call (include/crgpu.h): it comes back unaligned."""
    amp, guide, cut, hdr = synth.make_case(81, 160)
    reads = [amp, amp[:70] + "R" + amp[71:], amp[:100] + amp[104:], "ACGTYKM" * 20, amp]
    res = hotpath.run_hot_path(ctx, amp, aligner.pack_reads(reads), min_identity_score=60.0, want_rows=True)
    clean = hotpath.run_hot_path(ctx, amp, aligner.pack_reads([reads[0], reads[2], reads[4]]), min_identity_score=60.0, want_rows=True)
    assert (res.kept & 3).tolist() == [1, 0, 1, 0, 1]
    assert res.red.n_total == 3
    assert np.array_equal(res.red.results()[:-2], clean.red.results()[:-2])        # (n_cells counts every read handed in)
    assert res.bad_base.tolist() == [0, 1, 0, 1, 0]
    with pytest.raises(CrgpuError):
        aligner.needle_align(ctx, amp.replace("A", "R", 1), [amp])                # the AMPLICON is still an argument error
