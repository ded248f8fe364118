"""Host-side mirror of the reference's geometry / option handling (CPU only)."""
import numpy as np
import pytest

from crispresso_b200 import aligner, hotpath, synth


def _reference_include(L, cuts, w, left, right):
    """CORE:2739-2762, written out as the reference does it."""
    if cuts and w > 0:
        half = max(1, w // 2)
        inc = []
        for c in cuts:
            inc.append(range(max(0, c - half + 1), min(L - 1, c + half + 1)))
    else:
        inc = range(L)
    excl = []
    if left:
        excl += range(left)
    if right:
        excl += range(L)[-right:]
    return set(np.setdiff1d(np.ravel(inc), np.ravel(excl)))


@pytest.mark.parametrize("L,cuts,w,left,right", [(280, [112, 52], 1, 15, 15), (250, [125], 20, 15, 15), (100, [], 1, 0, 0),
                                                 (100, [50], 0, 5, 0), (120, [60], 7, 0, 30)])
def test_include_mask(L, cuts, w, left, right):
    m = hotpath.include_mask(L, cuts, w, left, right)
    assert set(np.nonzero(m)[0]) == _reference_include(L, cuts, w, left, right)


def test_cut_points_and_exons():
    amp, guide, cut, _ = synth.make_case(1, 200, hdr=False)
    assert hotpath.cut_points_from_guides(amp, guide) == [cut - 17 + (-3) + 20 - 1]
    assert hotpath.cut_points_from_guides(amp, synth.revcomp(guide)) == [cut - 17 + 3 - 1]
    exon, splice = hotpath.exon_masks(amp, amp[50:90])
    assert set(np.nonzero(exon)[0]) == set(range(50, 90))
    assert set(np.nonzero(splice)[0]) == {48, 49, 90, 91}
    exon, splice = hotpath.exon_masks(amp, amp[0:10] + "," + amp[190:200])
    assert set(np.nonzero(splice)[0]) == {10, 11, 188, 189}


def test_needle_ids_follow_the_sed_and_parser_round_trip():
    assert aligner.needle_id("M06879:15:000000000-DFF22:1:1101:25894:23776 1:N:0:1") == "@M06879:15:000000000-DFF22:1:1101:25894:23776"
    assert aligner.needle_id("read_7_x extra") == "@read:7:x"          # underscores come back as colons (CORE:1725)


def test_needle_option_parsing():
    assert aligner.parse_needle_options("-gapopen=10 -gapextend=0.5  -awidth3=5000") == (10.0, 0.5)
    assert aligner.parse_needle_options("-gapextend=2 -gapopen=12") == (12.0, 2.0)
    with pytest.raises(ValueError):
        aligner.parse_needle_options("-gapopen=10 -endweight")       # cannot be honoured: refuse loudly
    with pytest.raises(ValueError):
        aligner.parse_needle_options("-datafile=EBLOSUM62")


def test_reductions_round_trip():
    r = hotpath.Reductions(50)
    r.vectors[3, 7] = 5; r.hist_inframe[hotpath.HIST_ZERO - 3] = 2; r.counters[1] = 9; r.class_counts[2] = 4
    r.n_total, r.n_cells = 11, 12345
    s = hotpath.Reductions(50)
    s.load_flat(r.flat())
    assert np.array_equal(s.flat(), r.flat()) and s.n_total == 11
    assert hotpath.Reductions.hist_dict(s.hist_inframe) == {-3: 2}


def test_synthetic_reads_are_deterministic():
    amp, _g, cut, hdr = synth.make_case(1234, 250)
    a = synth.make_reads_fast(amp, hdr, cut, 50000, seed=5, read_len=250)
    b = synth.make_reads_fast(amp, hdr, cut, 50000, seed=5, read_len=250)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert (np.diff(a[1]) == 250).all()
