"""Paired-end merge on the B200 (crgpu_flash_merge) against the FLASH restatement
oracle/flash_merge.py -- the one that turns the reference's paired test FASTQs into its golden counts
(tests/test_kat_reference.py).  Bit-exact: kind, overlap position, merged bases, merged qualities."""
import ctypes
import gzip
import json
import os

import numpy as np
import pytest

from crispresso_b200 import _lib, flash, hotpath, synth
from crispresso_b200.aligner import pack_reads
from oracle import flash_merge

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _check(res, s1, q1, s2, q2, expect=None, **kw):
    reads, quals = res.reads(), res.quals()
    where = {int(p): j for j, p in enumerate(res.index)}
    for p in range(len(s1)):
        m = expect[p] if expect is not None else flash_merge.merge_pair(s1[p], q1[p], s2[p], q2[p], **kw)
        if m is None:
            assert res.kind[p] == 0 and res.pos[p] == -1 and p not in where, p
        else:
            assert p in where, (p, m[2])
            assert ("", "innie", "outie")[res.kind[p]] == m[2], p
            assert reads[where[p]] == m[0], p
            assert quals[where[p]] == m[1], p
    assert list(res.index) == sorted(where)                 # merged reads come out in pair order


def test_reference_pairs_merge_like_the_flash_restatement(ctx):
    with gzip.open(os.path.join(HERE, "golden", "flash_pairs_subset.json.gz"), "rt") as f:
        pairs = json.load(f)["pairs"]
    s1, q1, s2, q2 = ([p[k] for p in pairs] for k in ("s1", "q1", "s2", "q2"))
    res = flash.merge_pairs(ctx, s1, q1, s2, q2)
    _check(res, s1, q1, s2, q2, expect=[tuple(p["merged"]) if p["merged"] else None for p in pairs])
    kinds = np.bincount(res.kind, minlength=3)
    assert kinds[2] >= 250 and kinds[0] >= 150


@pytest.mark.parametrize("seed,read_len,coarse,lowc", [(11, 150, False, 0.0), (12, 100, True, 0.1), (13, 250, True, 0.05),
                                                      (14, 37, True, 0.2), (15, 300, False, 0.0)])
def test_synthetic_pairs(ctx, seed, read_len, coarse, lowc):
    amp, _g, _c, _h = synth.make_case(seed, 280, hdr=False)
    s1, q1, s2, q2 = synth.make_pairs(amp, 400, read_len, seed=seed, coarse_quals=coarse, low_complexity_frac=lowc)
    res = flash.merge_pairs(ctx, s1, q1, s2, q2)
    _check(res, s1, q1, s2, q2)
    assert 0 < res.n_merged < 400


def test_options_and_edge_cases(ctx):
    amp, _g, _c, _h = synth.make_case(21, 200, hdr=False)
    s1, q1, s2, q2 = synth.make_pairs(amp, 250, 120, seed=21, coarse_quals=True, low_complexity_frac=0.1)
    # mates shorter than the minimum overlap, a 1-base mate, all-N mates, identical mates
    s1 += ["ACG", "A", "N" * 50, "ACGTACGTAC", "ACGTTGCA" * 8]
    q1 += ["III", "I", "#" * 50, "IIIIIIIIII", "I" * 64]
    s2 += ["CGT", "T", "N" * 50, "GTACGTACGT", synth.revcomp("ACGTTGCA" * 8)]
    q2 += ["III", "I", "#" * 50, "5555555555", "5" * 64]
    for kw in (dict(), dict(min_overlap=10, max_overlap=60), dict(allow_outies=False), dict(min_overlap=1, max_overlap=300),
               dict(max_mismatch_density=0.1)):
        res = flash.merge_pairs(ctx, s1, q1, s2, q2, **kw)
        _check(res, s1, q1, s2, q2, **kw)
    # no pairs at all
    res = flash.merge_pairs(ctx, [], [], [], [])
    assert res.n_merged == 0


def test_bad_input_fails_loudly(ctx):
    with pytest.raises(_lib.CrgpuError) as e:
        flash.merge_pairs(ctx, ["ACGTXACGT"], ["IIIIIIIII"], ["ACGTACGTA"], ["IIIIIIIII"])
    assert e.value.code == _lib.E_ALIGN
    with pytest.raises(_lib.CrgpuError):
        flash.merge_pairs(ctx, ["A" * 1025], ["I" * 1025], ["T" * 1025], ["I" * 1025])
    with pytest.raises(_lib.CrgpuError):
        flash.merge_pairs(ctx, ["ACGT"], ["IIII"], ["ACGT"], ["IIII"], min_overlap=0)


def test_device_memory_chain_merge_then_align_and_quantify(ctx):
    """CRGPU_MEM_DEVICE: the merged (reads, offsets) stay in HBM and feed crgpu_align_quantify directly."""
    import torch
    amp, guide, cut, _h = synth.make_case(31, 250, hdr=False)
    s1, q1, s2, q2 = synth.make_pairs(amp, 3000, 150, seed=31)
    host = flash.merge_pairs(ctx, s1, q1, s2, q2)
    n = len(s1)
    b1, o1 = pack_reads(s1); c1, _ = pack_reads(q1); b2, o2 = pack_reads(s2); c2, _ = pack_reads(q2)
    dev = [torch.from_numpy(x).cuda() for x in (b1, c1, o1, b2, c2, o2)]
    cap = int(o1[-1] + o2[-1])
    d_pos = torch.zeros(n, dtype=torch.int32, device="cuda"); d_kind = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_seq = torch.zeros(cap, dtype=torch.uint8, device="cuda"); d_qual = torch.zeros(cap, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda"); d_idx = torch.zeros(n, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    prm = _lib.MergeParams(4, 100, 0.25, 1)
    out = _lib.MergeOut()
    out.pos, out.kind, out.seq, out.qual, out.offsets, out.index = (t.data_ptr() for t in (d_pos, d_kind, d_seq, d_qual, d_off, d_idx))
    out.cap_bytes, out.cap_reads = cap, n
    ctx.check(ctx.lib.crgpu_flash_merge(ctx.handle, _lib.MEM_DEVICE, dev[0].data_ptr(), dev[1].data_ptr(), dev[2].data_ptr(),
                                        dev[3].data_ptr(), dev[4].data_ptr(), dev[5].data_ptr(), n, ctypes.byref(prm), ctypes.byref(out)))
    m = int(out.n_merged)
    assert m == host.n_merged and int(out.bytes) == len(host.seq)
    assert np.array_equal(d_kind.cpu().numpy(), host.kind) and np.array_equal(d_pos.cpu().numpy(), host.pos)
    assert np.array_equal(d_seq.cpu().numpy()[:int(out.bytes)], host.seq)
    assert np.array_equal(d_qual.cpu().numpy()[:int(out.bytes)], host.qual)
    assert np.array_equal(d_off.cpu().numpy()[:m + 1], host.offsets) and np.array_equal(d_idx.cpu().numpy()[:m], host.index)
    # merged reads -> alignment + quantification, without leaving the device
    inc = hotpath.include_mask(250, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    ref = hotpath.run_hot_path(ctx, amp, (host.seq, host.offsets), inc=inc, min_identity_score=60.0)
    outd = {"kept": torch.zeros(m, dtype=torch.uint8, device="cuda"),
            "aln": torch.zeros(m * _lib.ALN_REC.itemsize, dtype=torch.uint8, device="cuda"),
            "recs": torch.zeros(m * _lib.READ_REC.itemsize, dtype=torch.uint8, device="cuda"),
            "tenths_rep": torch.zeros(m, dtype=torch.int32, device="cuda")}
    red = hotpath.Reductions(250)
    torch.cuda.synchronize()
    hotpath.run_hot_path(ctx, amp, None, inc=inc, min_identity_score=60.0, red=red,
                         device_inputs=(d_seq.data_ptr(), d_off.data_ptr(), m, 0, {k: v.data_ptr() for k, v in outd.items()}))
    assert np.array_equal(red.results(), ref.red.results())      # (all but n_cells_computed, a work counter)
    assert ref.red.n_total > 0.5 * m
