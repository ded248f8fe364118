"""oracle/needle_oracle.c: known-answer vectors, float32 vs exact-integer
agreement, identity rounding (App. B.3).  CPU only."""
import numpy as np
import pytest

from crispresso_b200 import synth
from oracle import needle

MICRO = [
    ("ACGTACGTTTACGATCGA", "ACGTACGTACGATCGAGG", "ACGTACGTTTACGATCGA--", "|||||||  |||||||||  ", "ACGTACG--TACGATCGAGG", 800, 69.5),
    ("GATTACAAAAAGGCTTCAGT", "GATTACAAAAGGCTTCAGT", "GATTACAAAAAGGCTTCAGT", "|||||| |||||||||||||", "GATTAC-AAAAGGCTTCAGT", 950, 85.0),
    ("GATTACAAAAAGGCTTCAGT", "GATTACAAAAAAGGCTTCAGT", "GATTAC-AAAAAGGCTTCAGT", "|||||| ||||||||||||||", "GATTACAAAAAAGGCTTCAGT", 952, 90.0),
    ("CCGTTAGCATCGATCGGATCTTAGC", "GTTAGCATCGATCGGATCTT", "CCGTTAGCATCGATCGGATCTTAGC", "  ||||||||||||||||||||   ", "--GTTAGCATCGATCGGATCTT---", 800, 100.0),
    ("GTTAGCATCGATCGGATCTT", "CCGTTAGCATCGATCGGATCTTAGC", "--GTTAGCATCGATCGGATCTT---", "  ||||||||||||||||||||   ", "CCGTTAGCATCGATCGGATCTTAGC", 800, 100.0),
    ("ACGTTGCAAGGCTTACGGATCCA", "ACGTTGCATGGCTACGGATCCA", "ACGTTGCAAGGCTTACGGATCCA", "||||||||.||| ||||||||||", "ACGTTGCATGGC-TACGGATCCA", 913, 91.0),
    ("ACGTTGCANGGCTTACGGATCCA", "ACGTTGCAAGGCTTACGGATCCA", "ACGTTGCANGGCTTACGGATCCA", "||||||||.||||||||||||||", "ACGTTGCAAGGCTTACGGATCCA", 957, 108.0),
    ("TTGACCTGAAGGCATCATCATCGGTA", "TTGACCTGAAGGCATCATCGGTA", "TTGACCTGAAGGCATCATCATCGGTA", "||||||||||||   |||||||||||", "TTGACCTGAAGG---CATCATCGGTA", 885, 104.0),
    ("TTGACCTGAAGGCATCATCGGTA", "TTGACCTGAAGGCATCATCATCGGTA", "TTGACCTGAAGG---CATCATCGGTA", "||||||||||||   |||||||||||", "TTGACCTGAAGGCATCATCATCGGTA", 885, 104.0),
]


@pytest.mark.parametrize("use_int", [False, True])
@pytest.mark.parametrize("a,b,ref,mark,qry,tenths,score", MICRO)
def test_micro_known_answers(a, b, ref, mark, qry, tenths, score, use_int):
    res, r, m, q = needle.align_batch(a, [b], use_int=use_int)
    assert (r[0], m[0], q[0]) == (ref, mark, qry)
    assert int(res["tenths"][0]) == tenths and float(res["score"][0]) == score


@pytest.mark.parametrize("gapopen,gapextend", [(10.0, 0.5), (12.0, 2.0), (5.0, 0.0), (10.0, 0.25)])
def test_float_and_integer_forms_agree(gapopen, gapextend):
    amp, _g, cut, hdr = synth.make_case(3, 150)
    packed = synth.make_reads(amp, hdr, cut, 400, seed=3, n_rate=0.01)
    f = needle.align_batch(amp, packed, gapopen, gapextend, use_int=False, nthreads=4)
    i = needle.align_batch(amp, packed, gapopen, gapextend, use_int=True, nthreads=4)
    assert f[1:] == i[1:]
    assert np.array_equal(f[0], i[0])


def test_identity_rounding_is_printf_of_a_float32():
    # "%4.1f" % (float32(100) * ident / len): exact .x5 ties go to even, inexact ones by value
    assert needle.identity_tenths(1, 16) == 62       # 6.25  -> "6.2"
    assert needle.identity_tenths(3, 16) == 188      # 18.75 -> "18.8"
    assert needle.identity_tenths(276, 280) == 986   # the value shown in SURVEY App. B.1
    assert needle.identity_tenths(250, 250) == 1000
    for ident, ln in [(151, 280), (199, 201), (1, 3), (2, 3), (127, 254)]:
        f = np.float32(100.0) * np.float32(ident) / np.float32(ln)
        assert needle.identity_tenths(ident, ln) == int(float("%4.1f" % f) * 10 + 0.5)


def test_threads_do_not_change_results():
    amp, _g, cut, hdr = synth.make_case(4, 120)
    packed = synth.make_reads(amp, hdr, cut, 300, seed=4)
    a = needle.align_batch(amp, packed, nthreads=1)
    b = needle.align_batch(amp, packed, nthreads=8)
    assert a[1:] == b[1:] and np.array_equal(a[0], b[0])


def test_bad_characters_are_rejected():
    with pytest.raises(ValueError):
        needle.align_batch("ACGT", ["ACRT"])
