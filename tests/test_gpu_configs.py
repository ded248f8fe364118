"""Parity cases shaped like BASELINE.json configs[2..4] (reduced sizes, oracle as checker) and
size-independent properties at the full configs[1] size.  Needs a B200."""
import numpy as np
import pytest

from crispresso_b200 import _lib, aligner, hotpath, synth
from oracle import quantify

pytestmark = pytest.mark.gpu


def _check(ctx, amp, packed, guide, hdr="", coding="", window=1):
    L = len(amp)
    inc = hotpath.include_mask(L, hotpath.cut_points_from_guides(amp, guide), window, 15, 15)
    exon = splice = None
    if coding:
        exon, splice = hotpath.exon_masks(amp, coding)
    flags = hotpath.quant_flags(hdr, window_around_sgrna=window, coding_seq=coding)
    res = hotpath.run_hot_path(ctx, amp, packed, hdr_amplicon=hdr or None, flags=flags, inc=inc, exon=exon, splice=splice)
    ora = quantify.hot_path(amp, packed, hdr_amplicon=hdr, opts=quantify.Opts(coding_seq=coding, expected_hdr_amplicon_seq=hdr,
                                                                              window_around_sgrna=window),
                            include=np.nonzero(inc)[0], exon=np.nonzero(exon)[0] if coding else (),
                            splice=np.nonzero(splice)[0] if coding else ())
    red = res.red
    assert red.n_total == ora["n_total"] and red.n_cells == ora["n_cells"]
    assert red.class_counts.tolist() == [ora["classes"][k] for k in ("UNMODIFIED", "NHEJ", "HDR", "MIXED")]
    for k, name in enumerate(hotpath.VECTOR_NAMES):
        assert red.vectors[k].tolist() == ora["vectors"][name].tolist(), name
    assert hotpath.Reductions.hist_dict(red.hist_inframe) == ora["hist_inframe"]
    assert hotpath.Reductions.hist_dict(red.hist_frameshift) == ora["hist_frameshift"]
    fw = [r for r in ora["rows"] if not r["rc"]]
    kept = np.nonzero(res.kept & 1)[0]
    assert kept.tolist() == [r["read"] for r in fw]
    assert (res.aln["tenths"][kept] / 10.0).tolist() == [r["score_ref"] for r in fw]
    return res, ora


def test_cfg3_merged_pe_with_20pct_indels_and_coding_sequence(ctx):
    amp, guide, cut, _ = synth.make_case(303, 300, hdr=False)
    packed = synth.make_reads(amp, None, cut, 4000, seed=303, read_len=300, len_sigma=8.0, p_exact=0.8)
    _res, ora = _check(ctx, amp, packed, guide, coding=amp[cut - 60:cut + 60], window=20)
    assert ora["counters"]["modified_frameshift"] > 50 and ora["counters"]["modified_non_frameshift"] > 10


def test_cfg4_pooled_amplicons(ctx):
    """CRISPRessoPooled hands each amplicon's reads to an independent CRISPResso run
    (CRISPRessoPooledCORE.py:882-908): several amplicons of 150-400 bp, one context, back to back."""
    rng = np.random.default_rng(404)
    for j, L in enumerate(rng.integers(150, 401, size=6)):
        amp, guide, cut, _ = synth.make_case(4040 + j, int(L), hdr=False)
        packed = synth.make_reads(amp, None, cut, 700, seed=4040 + j)
        _check(ctx, amp, packed, guide)


def test_cfg5_600bp_amplicon(ctx):
    amp, guide, cut, _ = synth.make_case(505, 600, hdr=False)
    packed = synth.make_reads(amp, None, cut, 1200, seed=505, read_len=600)
    _check(ctx, amp, packed, guide)


def test_cfg2_full_size_properties(ctx):
    """2^20 reads x (amplicon + HDR amplicon), the bench workload.  The reads are drawn from a pool of
    20000 molecules: (1) every copy of a molecule gets the same records; (2) the pool's records equal
    the oracle's; (3) class counts and n_total are the multiplicity-weighted pool values; (4) the
    reductions of the whole equal the sum over two shards (what multi-GPU sharding relies on)."""
    n, pool_n = 1 << 20, 20000
    amp, guide, cut, hdr = synth.make_case(1234, 250)
    pool = synth.make_reads(amp, hdr, cut, pool_n, seed=1234, read_len=250)
    buf, off = synth.make_reads_fast(amp, hdr, cut, n, seed=1234, read_len=250)
    pick = np.random.default_rng(1235).integers(0, pool_n, size=n)           # the draw make_reads_fast makes
    assert np.array_equal(buf.reshape(n, 250), pool[0].reshape(pool_n, 250)[pick])
    inc = hotpath.include_mask(250, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    flags = hotpath.quant_flags(hdr)
    res = hotpath.run_hot_path(ctx, amp, (buf, off), hdr_amplicon=hdr, flags=flags, inc=inc)
    ora = quantify.hot_path(amp, pool, hdr_amplicon=hdr, opts=quantify.Opts(expected_hdr_amplicon_seq=hdr),
                            include=np.nonzero(inc)[0])
    # (1) + (2): per-read records against the pool molecule's oracle row
    first = np.full(pool_n, -1, np.int64)
    first[pick[::-1]] = np.arange(n)[::-1]
    seen = first >= 0
    for field in ("tenths", "ident", "alnlen", "start1", "start2"):
        assert np.array_equal(res.aln[field], res.aln[field][first[pick]]), field
    for field in ("cls", "n_mutated", "n_inserted", "n_deleted"):
        assert np.array_equal(res.recs[field], res.recs[field][first[pick]]), field
    assert np.array_equal(res.kept, res.kept[first[pick]])
    kept_pool = np.zeros(pool_n, bool)
    cls_pool = np.zeros(pool_n, np.uint8)
    for r, p in zip(ora["rows"], ora["per_row"]):
        assert not r["rc"]
        kept_pool[r["read"]] = True
        cls_pool[r["read"]] = (_lib.C_UNMODIFIED * p["UNMODIFIED"] + _lib.C_NHEJ * p["NHEJ"] + _lib.C_HDR * p["HDR"] + _lib.C_MIXED * p["MIXED"])
    assert np.array_equal((res.kept & 1).astype(bool)[first[seen]], kept_pool[seen])
    assert np.array_equal(res.recs["cls"][first[seen]][kept_pool[seen]], cls_pool[seen][kept_pool[seen]])
    # (3)
    mult = np.bincount(pick, minlength=pool_n)
    assert res.red.n_total == int(mult[kept_pool].sum())
    for slot, bit in enumerate((_lib.C_UNMODIFIED, _lib.C_NHEJ, _lib.C_HDR, _lib.C_MIXED)):
        assert res.red.class_counts[slot] == int(mult[kept_pool & (cls_pool == bit)].sum())
    assert res.red.n_cells == 2 * 250 * 250 * n
    # (4)
    h = n // 2
    red = hotpath.Reductions(250)
    hotpath.run_hot_path(ctx, amp, (buf[:off[h]], off[:h + 1]), hdr_amplicon=hdr, flags=flags, inc=inc, red=red)
    hotpath.run_hot_path(ctx, amp, (buf[off[h]:], off[h:] - off[h]), hdr_amplicon=hdr, flags=flags, inc=inc, red=red)
    assert np.array_equal(red.results(), res.red.results())


def test_chunked_run_equals_single_call_and_large_600bp_batch(ctx):
    """cfg5 shape at a size that needs several traceback batches and chunks: 60000 reads x 600 bp."""
    from crispresso_b200 import distributed
    amp, guide, cut, _ = synth.make_case(606, 600, hdr=False)
    packed = synth.make_reads_fast(amp, None, cut, 60000, seed=606, read_len=600)
    inc = hotpath.include_mask(600, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    whole = hotpath.run_hot_path(ctx, amp, packed, flags=hotpath.quant_flags(), inc=inc).red
    chunked = distributed.run_chunked(ctx, amp, packed, chunk_reads=17000, flags=hotpath.quant_flags(), inc=inc)
    assert np.array_equal(whole.results(), chunked.results())
    assert whole.n_cells == 600 * 600 * 60000 and whole.n_total == 60000
