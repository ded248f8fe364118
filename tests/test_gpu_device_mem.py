"""CRGPU_MEM_DEVICE: every entry point gives the same answers with device-resident buffers (torch
tensors only as containers for raw device pointers).  Needs a B200."""
import ctypes

import numpy as np
import pytest
import torch

from crispresso_b200 import _lib, aligner, hotpath, synth

pytestmark = pytest.mark.gpu


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_align_device_memory_matches_host_memory(ctx):
    amp, _g, cut, hdr = synth.make_case(41, 250)
    buf, off = synth.make_reads(amp, hdr, cut, 3000, seed=41)
    n = len(off) - 1
    host = aligner.needle_align(ctx, amp, (buf, off))
    slot = len(amp) + int(np.diff(off).max())
    d_buf, d_off = _dev(buf), _dev(off)
    d_recs = torch.zeros(n * _lib.ALN_REC.itemsize, dtype=torch.uint8, device="cuda")
    rows = [torch.zeros(n * slot, dtype=torch.uint8, device="cuda") for _ in range(3)]
    torch.cuda.synchronize()                         # torch fills on its own stream; the library uses another
    ctx.check(ctx.lib.crgpu_align(ctx.handle, _lib.MEM_DEVICE, amp.encode(), len(amp), d_buf.data_ptr(), d_off.data_ptr(), n,
                                  10.0, 0.5, d_recs.data_ptr(), rows[0].data_ptr(), rows[1].data_ptr(), rows[2].data_ptr(), slot))
    recs = d_recs.cpu().numpy().view(_lib.ALN_REC)
    assert np.array_equal(recs, host[0])
    for k in range(3):
        a2 = rows[k].cpu().numpy().reshape(n, slot)
        got = [a2[i, recs["aln_off"][i]:].tobytes().decode() for i in range(n)]
        assert got == host[1 + k]


def test_fused_path_device_memory_matches_host_memory(ctx):
    amp, guide, cut, hdr = synth.make_case(42, 200)
    buf, off = synth.make_reads(amp, hdr, cut, 2500, seed=42, rc_frac=0.1)
    n = len(off) - 1
    inc = hotpath.include_mask(200, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    flags = hotpath.quant_flags(hdr)
    host = hotpath.run_hot_path(ctx, amp, (buf, off), hdr_amplicon=hdr, flags=flags, inc=inc)
    d_buf, d_off = _dev(buf), _dev(off)
    out = {"kept": torch.zeros(n, dtype=torch.uint8, device="cuda"),
           "aln": torch.zeros(n * _lib.ALN_REC.itemsize, dtype=torch.uint8, device="cuda"),
           "recs": torch.zeros(n * _lib.READ_REC.itemsize, dtype=torch.uint8, device="cuda"),
           "tenths_rep": torch.zeros(n, dtype=torch.int32, device="cuda")}
    red = hotpath.Reductions(200)
    torch.cuda.synchronize()
    hotpath.run_hot_path(ctx, amp, None, hdr_amplicon=hdr, flags=flags, inc=inc, red=red,
                         device_inputs=(d_buf.data_ptr(), d_off.data_ptr(), n, 0, {k: v.data_ptr() for k, v in out.items()}))
    assert np.array_equal(red.results(), host.red.results())
    assert np.array_equal(out["kept"].cpu().numpy(), host.kept)
    assert np.array_equal(out["aln"].cpu().numpy().view(_lib.ALN_REC), host.aln)
    assert np.array_equal(out["recs"].cpu().numpy().view(_lib.READ_REC), host.recs)
    assert np.array_equal(out["tenths_rep"].cpu().numpy(), host.tenths_rep)


def test_qualfilter_device_memory(ctx):
    rng = np.random.default_rng(43)
    quals = ["".join(chr(33 + int(q)) for q in rng.integers(2, 41, size=int(rng.integers(1, 300)))) for _ in range(500)]
    buf, off = aligner.pack_reads(quals)
    keep_h = np.zeros(500, np.uint8)
    ctx.check(ctx.lib.crgpu_qualfilter(ctx.handle, _lib.MEM_HOST, _lib.ptr(buf), _lib.ptr(off), 500, 20, 5, _lib.ptr(keep_h)))
    d_keep = torch.zeros(500, dtype=torch.uint8, device="cuda")
    d_buf, d_off = _dev(buf), _dev(off)              # keep the tensors alive across the call
    torch.cuda.synchronize()
    ctx.check(ctx.lib.crgpu_qualfilter(ctx.handle, _lib.MEM_DEVICE, d_buf.data_ptr(), d_off.data_ptr(), 500, 20, 5, d_keep.data_ptr()))
    assert np.array_equal(d_keep.cpu().numpy(), keep_h)
    ref = [int(sum(ord(c) - 33 for c in q) >= 20 * len(q) and min(ord(c) - 33 for c in q) >= 5) for q in quals]
    assert keep_h.tolist() == ref


def test_two_contexts_and_overlap_toggle_agree(ctx):
    from crispresso_b200 import Context
    amp, guide, cut, hdr = synth.make_case(44, 250)
    packed = synth.make_reads(amp, hdr, cut, 6000, seed=44, read_len=250)
    other = Context(0)
    other.set_traceback_budget(64 << 20)          # many batches -> the three-stream pipeline is exercised
    try:
        a = hotpath.run_hot_path(other, amp, packed, hdr_amplicon=hdr, flags=hotpath.quant_flags(hdr))
        other.set_overlap(False)
        b = hotpath.run_hot_path(other, amp, packed, hdr_amplicon=hdr, flags=hotpath.quant_flags(hdr))
    finally:
        other.close()
    c = hotpath.run_hot_path(ctx, amp, packed, hdr_amplicon=hdr, flags=hotpath.quant_flags(hdr))
    for r in (a, b):
        assert np.array_equal(r.red.results(), c.red.results())
        assert np.array_equal(r.aln, c.aln) and np.array_equal(r.recs, c.recs) and np.array_equal(r.kept, c.kept)
