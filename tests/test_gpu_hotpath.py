"""crgpu_align_quantify (the fused hot path) against oracle.quantify.hot_path.  Needs a B200."""
import numpy as np
import pytest

from crispresso_b200 import _lib, aligner, hotpath, synth
from oracle import quantify

pytestmark = pytest.mark.gpu


def _run_both(ctx, amp, packed, hdr_amp="", window=1, coding="", guide=None, excl=(15, 15), min_id=60.0, hide=False,
              want_rows=True):
    L = len(amp)
    cuts = hotpath.cut_points_from_guides(amp, guide) if guide else []
    inc = hotpath.include_mask(L, cuts, window, excl[0], excl[1])
    exon = splice = None
    if coding:
        exon, splice = hotpath.exon_masks(amp, coding)
    flags = hotpath.quant_flags(hdr_amp, window_around_sgrna=window, hide_mutations_outside_window_NHEJ=hide, coding_seq=coding)
    res = hotpath.run_hot_path(ctx, amp, packed, min_identity_score=min_id, hdr_amplicon=hdr_amp or None, flags=flags,
                               inc=inc, exon=exon, splice=splice, want_rows=want_rows)
    opts = quantify.Opts(coding_seq=coding, expected_hdr_amplicon_seq=hdr_amp, window_around_sgrna=window,
                         hide_mutations_outside_window_NHEJ=hide)
    ora = quantify.hot_path(amp, packed, min_identity_score=min_id, hdr_amplicon=hdr_amp, opts=opts,
                            include=np.nonzero(inc)[0], exon=np.nonzero(exon)[0] if coding else (),
                            splice=np.nonzero(splice)[0] if coding else ())
    return res, ora


def _assert_same(res, ora, names=None, has_hdr=False, amp=None):
    red = res.red
    assert red.n_total == ora["n_total"]
    assert red.n_cells == ora["n_cells"]
    assert red.class_counts.tolist() == [ora["classes"][k] for k in ("UNMODIFIED", "NHEJ", "HDR", "MIXED")]
    for k, name in enumerate(hotpath.VECTOR_NAMES):
        assert red.vectors[k].tolist() == ora["vectors"][name].tolist(), name
    assert hotpath.Reductions.hist_dict(red.hist_inframe) == ora["hist_inframe"]
    assert hotpath.Reductions.hist_dict(red.hist_frameshift) == ora["hist_frameshift"]
    assert {n: red.counter(n) for n in hotpath.COUNTER_NAMES} == ora["counters"]
    if res.rows is not None:
        n = len(res.kept)
        names = names or ["r%d" % i for i in range(n)]
        df = hotpath.build_dataframe(res, names, has_hdr=has_hdr, amplicon=amp)
        assert list(df.index) == [r["ID"] for r in ora["rows"]]
        for col in ("ref_seq", "align_str", "align_seq", "length"):
            assert list(df[col]) == [r[col] for r in ora["rows"]], col
        assert np.array_equal(df["score_ref"].values, np.array([r["score_ref"] for r in ora["rows"]]))
        for col in ("UNMODIFIED", "NHEJ", "HDR", "MIXED", "n_mutated", "n_inserted", "n_deleted"):
            assert [int(x) for x in df[col]] == [int(p[col]) for p in ora["per_row"]], col


def test_cfg2_shape_with_hdr(ctx):
    amp, guide, cut, hdr = synth.make_case(1234, 250)
    packed = synth.make_reads(amp, hdr, cut, 3000, seed=1234, read_len=250)
    res, ora = _run_both(ctx, amp, packed, hdr_amp=hdr, guide=guide)
    _assert_same(res, ora, has_hdr=True, amp=amp)
    assert ora["classes"]["HDR"] > 100 and ora["classes"]["NHEJ"] > 100


def test_cfg3_shape_with_coding_sequence(ctx):
    amp, guide, cut, _ = synth.make_case(77, 300, hdr=False)
    packed = synth.make_reads(amp, None, cut, 2500, seed=77, read_len=300, len_sigma=8.0, p_exact=0.8)
    res, ora = _run_both(ctx, amp, packed, coding=amp[cut - 60:cut + 60], guide=guide, window=20)
    _assert_same(res, ora, amp=amp)
    assert ora["counters"]["modified_frameshift"] > 0


def test_reverse_complement_rescue(ctx):
    amp, guide, cut, hdr = synth.make_case(5, 200)
    packed = synth.make_reads(amp, hdr, cut, 1500, seed=5, rc_frac=0.3)
    for h in ("", hdr):
        res, ora = _run_both(ctx, amp, packed, hdr_amp=h, guide=guide)
        assert any(r["rc"] for r in ora["rows"])
        _assert_same(res, ora, has_hdr=bool(h), amp=amp)


def test_amplicon_with_n_and_junk(ctx):
    rng = np.random.default_rng(6)
    amp, guide, cut, _ = synth.make_case(6, 180, hdr=False)
    a = list(amp); a[cut + 9] = "N"; a[40] = "N"; amp = "".join(a)
    buf, off = synth.make_reads(amp.replace("N", "C"), None, cut, 1200, seed=6, n_rate=0.005)
    reads = [bytes(buf[off[i]:off[i + 1]]).decode() for i in range(len(off) - 1)]
    reads += [synth.random_seq(rng, int(rng.integers(20, 220))) for _ in range(100)]          # dropped by the identity filter
    reads += [amp.replace("N", "A"), amp.replace("N", "G")[:150]]
    from crispresso_b200.aligner import pack_reads
    res, ora = _run_both(ctx, amp, pack_reads(reads), guide=None, window=1, min_id=60.0)
    _assert_same(res, ora, amp=amp)


def test_low_identity_threshold_and_hide(ctx):
    amp, guide, cut, _ = synth.make_case(8, 150, hdr=False)
    packed = synth.make_reads(amp, None, cut, 1500, seed=8, read_len=151)
    res, ora = _run_both(ctx, amp, packed, guide=guide, window=10, min_id=30.0, hide=True, excl=(5, 5))
    _assert_same(res, ora, amp=amp)


def test_reductions_do_not_depend_on_sharding(ctx):
    """What multi-GPU sharding relies on: quantifying two halves and adding == quantifying all."""
    amp, guide, cut, hdr = synth.make_case(9, 250)
    buf, off = synth.make_reads(amp, hdr, cut, 2000, seed=9, read_len=250)
    inc = hotpath.include_mask(250, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    flags = hotpath.quant_flags(hdr)
    whole = hotpath.run_hot_path(ctx, amp, (buf, off), hdr_amplicon=hdr, flags=flags, inc=inc).red
    red = hotpath.Reductions(250)
    h = 1000
    hotpath.run_hot_path(ctx, amp, (buf[:off[h]].copy(), off[:h + 1].copy()), hdr_amplicon=hdr, flags=flags, inc=inc, red=red)
    hotpath.run_hot_path(ctx, amp, (buf[off[h]:].copy(), (off[h:] - off[h]).copy()), hdr_amplicon=hdr, flags=flags, inc=inc, red=red)
    assert np.array_equal(whole.results(), red.results())


def test_shared_dp_prefix_of_the_hdr_pass_changes_nothing(ctx):
    """Same-length HDR amplicon: the HDR pass reuses the DP rows above the first differing base.  Every
    output must equal the two-full-passes result; only the number of evaluated cells drops."""
    from crispresso_b200 import Context
    for La, seed in ((250, 71), (300, 72), (500, 73), (120, 74)):
        amp, guide, cut, hdr = synth.make_case(seed, La)
        packed = synth.make_reads(amp, hdr, cut, 1500, seed=seed, rc_frac=0.05)
        flags = hotpath.quant_flags(hdr)
        other = Context(0)
        try:
            other.set_traceback_budget(32 << 20)
            other.set_band(0)                             # single-pass fill: evaluated cells are exactly the DP cells
            other.set_exact_shortcut(False)               # (... of EVERY read)
            a = hotpath.run_hot_path(other, amp, packed, hdr_amplicon=hdr, flags=flags, want_rows=True)
            other.set_share_prefix(False)
            b = hotpath.run_hot_path(other, amp, packed, hdr_amplicon=hdr, flags=flags, want_rows=True)
        finally:
            other.close()
        assert b.red.n_cells_computed == b.red.n_cells == a.red.n_cells
        if La in (250, 500):
            assert a.red.n_cells_computed < 0.8 * a.red.n_cells           # the edit sits mid-amplicon: ~half of the HDR rows are shared
        a.red.n_cells_computed = b.red.n_cells_computed
        assert np.array_equal(a.red.flat(), b.red.flat())
        assert np.array_equal(a.aln, b.aln) and np.array_equal(a.recs, b.recs) and np.array_equal(a.kept, b.kept)
        assert np.array_equal(a.tenths_rep, b.tenths_rep)
        for k in range(3):
            assert np.array_equal(a.rows[k], b.rows[k])
    # an HDR amplicon that differs in the first bases shares nothing and silently takes two full passes
    amp, guide, cut, _ = synth.make_case(75, 200, hdr=False)
    hdr2 = "T" + amp[1:] if amp[0] != "T" else "G" + amp[1:]
    packed = synth.make_reads(amp, hdr2, cut, 500, seed=75)
    ctx.set_band(0)
    ctx.set_exact_shortcut(False)
    try:
        r = hotpath.run_hot_path(ctx, amp, packed, hdr_amplicon=hdr2, flags=hotpath.quant_flags(hdr2))
    finally:
        ctx.set_band(16)
        ctx.set_exact_shortcut(True)
    assert r.red.n_cells_computed == r.red.n_cells


@pytest.mark.parametrize("La,read_len,sigma,hdr_on", [(250, 250, 0.0, True), (300, 300, 8.0, False), (600, 600, 0.0, True),
                                                      (250, None, 0.0, True), (180, 400, 10.0, False), (700, 700, 0.0, True),
                                                      (60, 60, 2.0, True), (100, 100, 0.0, True)])
def test_banded_two_pass_fill_changes_nothing(La, read_len, sigma, hdr_on):
    """The banded fill (score pass + band pass + re-alignment of reads whose traceback leaves the band) must
    give exactly what the single-pass fill gives, for every band width -- including widths so small that
    most edited reads escape -- with and without the shared HDR prefix."""
    from crispresso_b200 import Context
    seed = 500 + La
    amp, guide, cut, hdr = synth.make_case(seed, La, hdr=hdr_on)
    packed = synth.make_reads(amp, hdr, cut, 1200, seed=seed, read_len=read_len, len_sigma=sigma, rc_frac=0.04, n_rate=0.002)
    flags = hotpath.quant_flags(hdr or "")
    c = Context(0)
    try:
        c.set_traceback_budget(24 << 20)
        c.set_band(0)
        ref = hotpath.run_hot_path(c, amp, packed, hdr_amplicon=hdr, flags=flags, want_rows=True)
        assert c.last_escaped() == (0, 0)
        seen_escape = False
        import os
        for B, share, whole_strips in ((24, True, False), (2, True, False), (40, False, False), (9, True, False), (16, True, True)):
            c.set_band(B)
            c.set_share_prefix(share)
            # the band pass works on half-height sub-strips where the tile allows it; CRGPU_NO_SUBSTRIP keeps whole strips
            os.environ.pop("CRGPU_NO_SUBSTRIP", None)
            if whole_strips:
                os.environ["CRGPU_NO_SUBSTRIP"] = "1"
            try:
                got = hotpath.run_hot_path(c, amp, packed, hdr_amplicon=hdr, flags=flags, want_rows=True)
            finally:
                os.environ.pop("CRGPU_NO_SUBSTRIP", None)
            seen_escape |= c.last_escaped()[0] > 0
            got.red.n_cells_computed = ref.red.n_cells_computed
            assert np.array_equal(got.red.flat(), ref.red.flat()), (B, share, whole_strips)
            assert np.array_equal(got.aln, ref.aln) and np.array_equal(got.recs, ref.recs) and np.array_equal(got.kept, ref.kept)
            assert np.array_equal(got.tenths_rep, ref.tenths_rep)
            for k in range(3):
                assert np.array_equal(got.rows[k], ref.rows[k])
        if La >= 250:
            assert seen_escape                   # the 2-column band cannot hold a read with an indel
    finally:
        c.close()


def test_pipelined_chunks_on_two_contexts_equal_one_call(ctx):
    """run_hot_path_pipelined (read chunks alternating between two contexts, each on its own host thread, so that
    one chunk's PCIe copies overlap the other's kernels) must return what one call over all reads returns: per-read
    records in read order, the RC-rescue list in read order, every reduction."""
    from crispresso_b200 import Context
    amp, guide, cut, hdr = synth.make_case(31, 210)
    packed = synth.make_reads(amp, hdr, cut, 1500, seed=32, read_len=210, rc_frac=0.07, len_sigma=5.0)
    inc = hotpath.include_mask(len(amp), hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    flags = hotpath.quant_flags(hdr)
    one = hotpath.run_hot_path(ctx, amp, packed, hdr_amplicon=hdr, flags=flags, inc=inc)
    other = Context(0)
    try:
        for ctxs, chunk in (([ctx, other], 400), ([ctx], 700), ([ctx, other], 4096)):
            pip = hotpath.run_hot_path_pipelined(ctxs, amp, packed, chunk_reads=chunk, hdr_amplicon=hdr, flags=flags, inc=inc)
            assert np.array_equal(pip.kept, one.kept)
            fields = [f for f in _lib.ALN_REC.names if f != "aln_off"]      # aln_off positions a text row in its call's slot
            for f in fields:
                assert np.array_equal(pip.aln[f], one.aln[f]), f
            assert np.array_equal(pip.tenths_rep, one.tenths_rep)
            assert pip.recs.tobytes() == one.recs.tobytes()
            assert len(one.rc_read) > 0 and np.array_equal(pip.rc_read, one.rc_read)
            for f in fields:
                assert np.array_equal(pip.rc_aln[f], one.rc_aln[f]), f
            assert pip.rc_recs.tobytes() == one.rc_recs.tobytes()
            a, b = pip.red.flat(), one.red.flat()
            assert np.array_equal(a[:-1], b[:-1])          # all but n_cells_computed (the band decision is per chunk)
    finally:
        other.close()


@pytest.mark.parametrize("La,read_len,hdr_on,amp_n", [(250, 250, True, False), (180, 180, False, False), (300, 300, True, False),
                                                        (200, 151, False, False), (160, 160, True, True)])
def test_exact_read_shortcut_changes_nothing(La, read_len, hdr_on, amp_n):
    """Reads identical to the amplicon (case aside) skip the DP: their records, ops, text rows and HDR identities are written
    directly (one representative goes through the DP).  Every output must equal the run with the shortcut off, and most
    unedited reads must take it (none when the amplicon holds an N: the argument needs the maximum score 5 L)."""
    from crispresso_b200 import Context
    seed = 700 + La
    amp, guide, cut, hdr = synth.make_case(seed, La, hdr=hdr_on)
    if amp_n:
        amp = amp[:40] + "N" + amp[41:]
        hdr = hdr[:40] + "N" + hdr[41:] if hdr else hdr
    buf, off = synth.make_reads(amp.replace("N", "A"), hdr.replace("N", "A") if hdr else None, cut, 1600, seed=seed, read_len=read_len,
                                rc_frac=0.03, sub_rate=0.001)
    reads = [bytes(buf[off[i]:off[i + 1]]).decode() for i in range(1600)]
    reads = reads + [r.lower() for r in reads[:40]] + [amp.replace("N", "A")] * 7            # lower-case copies, exact copies
    packed = aligner.pack_reads(reads)
    flags = hotpath.quant_flags(hdr or "")
    inc = hotpath.include_mask(La, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    c = Context(0)
    try:
        c.set_traceback_budget(24 << 20)
        c.set_exact_shortcut(False)
        ref = hotpath.run_hot_path(c, amp, packed, hdr_amplicon=hdr, flags=flags, inc=inc, want_rows=True, alleles=8192, min_identity_score=40.0)
        assert c.last_exact() == 0
        c.set_exact_shortcut(True)
        got = hotpath.run_hot_path(c, amp, packed, hdr_amplicon=hdr, flags=flags, inc=inc, want_rows=True, alleles=8192, min_identity_score=40.0)
        n_same = sum(1 for r in reads if r.upper() == amp)
        assert c.last_exact() == (max(n_same - 1, 0) if not amp_n else 0)
        if read_len == La and not amp_n:
            assert c.last_exact() > 300
        assert np.array_equal(got.red.results(), ref.red.results())
        assert np.array_equal(got.aln, ref.aln) and np.array_equal(got.recs, ref.recs) and np.array_equal(got.kept, ref.kept)
        assert np.array_equal(got.tenths_rep, ref.tenths_rep)
        assert np.array_equal(got.rc_read, ref.rc_read) and np.array_equal(got.rc_aln, ref.rc_aln)
        for k in range(3):
            o = ref.aln["aln_off"]
            for i in range(len(reads)):
                assert np.array_equal(got.rows[k][i, o[i]:], ref.rows[k][i, o[i]:]), (k, i)
        assert sorted(got.allele_count.tolist()) == sorted(ref.allele_count.tolist())
        lean = hotpath.run_hot_path(c, amp, packed, hdr_amplicon=hdr, flags=flags, inc=inc, min_identity_score=40.0)
        assert np.array_equal(lean.red.results(), ref.red.results()) and np.array_equal(lean.recs, ref.recs)
        fields = [f for f in _lib.ALN_REC.names if f != "aln_off"]
        for f in fields:
            assert np.array_equal(lean.aln[f], ref.aln[f]), f
    finally:
        c.close()


def test_staged_pipeline_equals_one_call(ctx):
    """hotpath.run_hot_path_staged -- crgpu_stage_reads (one base per byte, or BAM 4-bit codes unpacked on the device) on the
    copy stream while crgpu_align_quantify_staged works on the previous chunk -- must return what one call over all reads
    returns: per-read records, the RC-rescue list, every reduction, and the same allele table (merged by key)."""
    amp, guide, cut, hdr = synth.make_case(33, 200)
    packed_reads = synth.make_reads(amp, hdr, cut, 2400, seed=34, read_len=200, rc_frac=0.07, n_rate=0.002)
    buf, off = packed_reads
    buf = buf.copy()
    buf[off[17] + 5] = ord("R")                                   # one IUPAC code: reported per read in either format
    inc = hotpath.include_mask(len(amp), hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    flags = hotpath.quant_flags(hdr)
    one = hotpath.run_hot_path(ctx, amp, (buf, off), hdr_amplicon=hdr, flags=flags, inc=inc, want_rows=True, alleles=4096)
    assert one.bad_base[17] == 1 and int(one.bad_base.sum()) == 1
    text = {}
    for k in range(one.allele_n):                                 # allele (text rows + class + counts) -> #Reads
        row = int(one.allele_row[k])
        text[_allele_text(one, row)] = int(one.allele_count[k])
    for chunk, fmt, deferred in ((700, None, False), (1000, "bam4", True), (300, None, True), (4096, "bam4", False)):
        pk = hotpath.pack_bam4(buf) if fmt else None
        got = hotpath.run_hot_path_staged(ctx, amp, (buf, off), chunk_reads=chunk, packed=pk, hdr_amplicon=hdr, flags=flags,
                                          inc=inc, alleles=4096, deferred=deferred)
        assert np.array_equal(got.kept, one.kept)
        fields = [f for f in _lib.ALN_REC.names if f != "aln_off"]
        for f in fields:
            assert np.array_equal(got.aln[f], one.aln[f]), f
        assert np.array_equal(got.tenths_rep, one.tenths_rep) and got.recs.tobytes() == one.recs.tobytes()
        assert len(one.rc_read) > 0 and np.array_equal(got.rc_read, one.rc_read)
        for f in fields:
            assert np.array_equal(got.rc_aln[f], one.rc_aln[f]), f
        assert got.rc_recs.tobytes() == one.rc_recs.tobytes()
        assert np.array_equal(got.red.results(), one.red.results())
        assert got.allele_n == one.allele_n and sorted(got.allele_count.tolist()) == sorted(one.allele_count.tolist())
        # every merged allele's representative row spells an allele of the single call, with that call's count
        for k in range(got.allele_n):
            assert text[_allele_text(one, int(got.allele_row[k]))] == int(got.allele_count[k])


def _allele_text(res, row):
    """Grouping key of the reference's allele table (CORE:2923-2946) for forward row `row` (or RC row row - n) of a
    want_rows result."""
    n = len(res.kept)
    if row < n:
        r = res.recs[row]
        o = int(res.aln["aln_off"][row])
        return (res.rows[0][row, o:].tobytes().decode(), res.rows[2][row, o:].tobytes().decode(), int(r["cls"]), int(r["n_mutated"]),
                int(r["n_inserted"]), int(r["n_deleted"]))
    j = row - n
    r = res.rc_recs[j]
    return (res.rc_rows[0][j], res.rc_rows[2][j], int(r["cls"]), int(r["n_mutated"]), int(r["n_inserted"]), int(r["n_deleted"]))


def test_band_holds_off_after_a_call_whose_reads_mostly_escape(ctx):
    """More than a quarter of the reads leaving the band means the band costs more than it saves: the next calls run
    the single-pass fill (same results, no escapes) until crgpu_set_band is called again."""
    amp, guide, cut, hdr = synth.make_case(41, 250, hdr=False)
    packed = synth.make_reads(amp, None, cut, 800, seed=42, read_len=250, p_exact=0.2, p_hdr=0.0)
    flags = hotpath.quant_flags("")
    ctx.set_band(1)
    ctx.set_exact_shortcut(False)                         # (evaluated cells are compared with the DP cells of every read)
    try:
        first = hotpath.run_hot_path(ctx, amp, packed, flags=flags)
        assert ctx.last_escaped()[0] * 4 > 800
        second = hotpath.run_hot_path(ctx, amp, packed, flags=flags)
        assert ctx.last_escaped() == (0, 0) and second.red.n_cells_computed == second.red.n_cells
        assert np.array_equal(first.aln, second.aln) and np.array_equal(first.recs, second.recs)
        assert np.array_equal(first.red.flat()[:-1], second.red.flat()[:-1])
        ctx.set_band(1)                                   # a new setting is tried again
        hotpath.run_hot_path(ctx, amp, packed, flags=flags)
        assert ctx.last_escaped()[0] * 4 > 800
    finally:
        ctx.set_exact_shortcut(True)
        ctx.set_band(16)


@pytest.mark.parametrize("La,read_len,sigma,hdr_on", [(250, 250, 0.0, True), (300, 300, 8.0, False), (200, 151, 0.0, False),
                                                      (150, 220, 6.0, True), (600, 600, 0.0, True)])
def test_diagonal_shortcut_changes_nothing(La, read_len, sigma, hdr_on):
    """Alignments whose start-cell score equals the substitution-score sum of the diagonal through the start cell are
    emitted right after the score pass, without flags (the traceback is provably that diagonal); only the other read
    pairs go through the band pass and the walk.  Every output must equal the run with the shortcut off -- text rows,
    records, HDR identities, reductions -- and most pairs of an amplicon read set must take the shortcut."""
    from crispresso_b200 import Context
    seed = 900 + La
    amp, guide, cut, hdr = synth.make_case(seed, La, hdr=hdr_on)
    packed = synth.make_reads(amp, hdr, cut, 1500, seed=seed, read_len=read_len, len_sigma=sigma, rc_frac=0.03, n_rate=0.001)
    flags = hotpath.quant_flags(hdr or "")
    c = Context(0)
    try:
        c.set_traceback_budget(24 << 20)
        c.set_exact_shortcut(False)                       # (its reads would leave the plan: the pair counts below are of all reads)
        c.set_diag_shortcut(False)
        ref = hotpath.run_hot_path(c, amp, packed, hdr_amplicon=hdr, flags=flags, want_rows=True, min_identity_score=40.0)
        total, left = c.last_diag()
        assert total == left and total > 0
        for share in (True, False):
            c.set_diag_shortcut(True)
            c.set_share_prefix(share)
            got = hotpath.run_hot_path(c, amp, packed, hdr_amplicon=hdr, flags=flags, want_rows=True, min_identity_score=40.0)
            total, left = c.last_diag()
            if read_len == La and sigma == 0.0 and (share or not hdr_on):
                # the amplicon pass: ~30 % of the reads are edited or HDR reads, i.e. ~half of the pairs hold one (without
                # the shared prefix the HDR amplicon is a pass of its own, where only the HDR reads are diagonal)
                assert left < 0.6 * total, (total, left)
            got.red.n_cells_computed = ref.red.n_cells_computed
            assert np.array_equal(got.red.flat(), ref.red.flat()), share
            assert np.array_equal(got.aln, ref.aln) and np.array_equal(got.recs, ref.recs) and np.array_equal(got.kept, ref.kept)
            assert np.array_equal(got.tenths_rep, ref.tenths_rep)
            assert np.array_equal(got.rc_read, ref.rc_read) and np.array_equal(got.rc_aln, ref.rc_aln)
            for k in range(3):
                assert np.array_equal(got.rows[k], ref.rows[k])
            # without text rows (the benchmark path): records and reductions again
            lean = hotpath.run_hot_path(c, amp, packed, hdr_amplicon=hdr, flags=flags, min_identity_score=40.0)
            lean.red.n_cells_computed = ref.red.n_cells_computed
            assert np.array_equal(lean.red.flat(), ref.red.flat())
            fields = [f for f in _lib.ALN_REC.names if f != "aln_off"]
            for f in fields:
                assert np.array_equal(lean.aln[f], ref.aln[f]), f
            assert np.array_equal(lean.recs, ref.recs) and np.array_equal(lean.tenths_rep, ref.tenths_rep)
    finally:
        c.close()


def test_staged_batches_in_any_call_order(ctx):
    """crgpu_stage_reads only registers a batch; its copy starts inside the next call (or at crgpu_sync, or when the batch itself
    is run first).  Whatever the order of stage / sync / run, every batch must give what the plain call gives, and deferred
    per-read outputs must be in host memory after crgpu_sync."""
    amp, guide, cut, hdr = synth.make_case(91, 180)
    inc = hotpath.include_mask(len(amp), hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    flags = hotpath.quant_flags(hdr)
    sets = [synth.make_reads(amp, hdr, cut, n, seed=92 + k, read_len=180, rc_frac=0.05) for k, n in enumerate((900, 500, 1300))]
    want = [hotpath.run_hot_path(ctx, amp, s, hdr_amplicon=hdr, flags=flags, inc=inc) for s in sets]

    def outputs(n):
        return {"kept": np.zeros(n, np.uint8), "aln": np.zeros(n, _lib.ALN_REC), "recs": np.zeros(n, _lib.READ_REC),
                "tenths_rep": np.zeros(n, np.int32), "rc_read": np.zeros(n, np.int32), "rc_aln": np.zeros(n, _lib.ALN_REC),
                "rc_recs": np.zeros(n, _lib.READ_REC)}

    def same(out, ref):
        assert np.array_equal(out["kept"], ref.kept) and out["recs"].tobytes() == ref.recs.tobytes()
        for f in _lib.ALN_REC.names:
            if f != "aln_off":
                assert np.array_equal(out["aln"][f], ref.aln[f]), f
        assert np.array_equal(out["tenths_rep"], ref.tenths_rep)

    for deferred in (False, True):
        pipe = hotpath.StagedPipeline(ctx, amp, hdr_amplicon=hdr, flags=flags, inc=inc, deferred=deferred)
        outs = [outputs(len(s[1]) - 1) for s in sets]
        # two batches staged back to back, then run: the first one's copy starts when it is run, the second one's inside that call
        pipe.stage(*sets[0]); pipe.stage(*sets[1])
        pipe.run(outs[0]); pipe.run(outs[1])
        # a batch staged, the context synchronised (copy forced), then run; its outputs fetched by the final sync
        pipe.stage(*sets[2])
        ctx.sync()
        pipe.run(outs[2])
        red, _table = pipe.take_results(sync=True)
        for o, w in zip(outs, want):
            same(o, w)
        total = want[0].red.results() + want[1].red.results() + want[2].red.results()
        assert np.array_equal(red.results(), total)


def test_longest_amplicons_quantify_in_shared_memory(ctx):
    """k_quantify keeps the 15 per-position vectors of a CTA in shared memory: above 819 bp that is more than the default 48 KB
    (opt-in size, set per launch).  A 1000-bp amplicon -- close to CRGPU_MAX_AMPLICON -- against the oracle."""
    amp, guide, cut, hdr = synth.make_case(4321, 1000)
    packed = synth.make_reads(amp, hdr, cut, 160, seed=4321, read_len=1000)
    res, ora = _run_both(ctx, amp, packed, hdr_amp=hdr, guide=guide, want_rows=False)
    _assert_same(res, ora, has_hdr=True, amp=amp)
    assert ora["classes"]["NHEJ"] > 5
