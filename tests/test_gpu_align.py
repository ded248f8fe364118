"""crgpu_align (k_gotoh_fill + k_traceback_walk) against the oracle: bit-exact rows, identity,
score and start cell.  Needs a B200."""
import numpy as np
import pytest

from crispresso_b200 import aligner, synth
from crispresso_b200._lib import CrgpuError
from oracle import needle

pytestmark = pytest.mark.gpu


def _compare(ctx, amp, packed, gapopen=10.0, gapextend=0.5):
    recs, r, m, q = aligner.needle_align(ctx, amp, packed, gapopen, gapextend)
    ores, orr, om, oq = needle.align_batch(amp, packed, gapopen, gapextend, use_int=True, nthreads=8)
    assert r == orr and m == om and q == oq
    for k in ("tenths", "ident", "alnlen", "start1", "start2"):
        assert np.array_equal(recs[k], ores[k]), k
    assert np.array_equal(recs["score"].astype(np.float64), ores["score"])
    assert np.array_equal(recs["read_len"], np.diff(packed[1]).astype(np.int32))


MICRO = [
    ("ACGTACGTTTACGATCGA", "ACGTACGTACGATCGAGG", "ACGTACGTTTACGATCGA--", "ACGTACG--TACGATCGAGG", 800, 139),
    ("GATTACAAAAAGGCTTCAGT", "GATTACAAAAGGCTTCAGT", "GATTACAAAAAGGCTTCAGT", "GATTAC-AAAAGGCTTCAGT", 950, 170),
    ("GATTACAAAAAGGCTTCAGT", "GATTACAAAAAAGGCTTCAGT", "GATTAC-AAAAAGGCTTCAGT", "GATTACAAAAAAGGCTTCAGT", 952, 180),
    ("CCGTTAGCATCGATCGGATCTTAGC", "GTTAGCATCGATCGGATCTT", "CCGTTAGCATCGATCGGATCTTAGC", "--GTTAGCATCGATCGGATCTT---", 800, 200),
    ("GTTAGCATCGATCGGATCTT", "CCGTTAGCATCGATCGGATCTTAGC", "--GTTAGCATCGATCGGATCTT---", "CCGTTAGCATCGATCGGATCTTAGC", 800, 200),
    ("ACGTTGCAAGGCTTACGGATCCA", "ACGTTGCATGGCTACGGATCCA", "ACGTTGCAAGGCTTACGGATCCA", "ACGTTGCATGGC-TACGGATCCA", 913, 182),
    ("ACGTTGCANGGCTTACGGATCCA", "ACGTTGCAAGGCTTACGGATCCA", "ACGTTGCANGGCTTACGGATCCA", "ACGTTGCAAGGCTTACGGATCCA", 957, 216),
    ("TTGACCTGAAGGCATCATCATCGGTA", "TTGACCTGAAGGCATCATCGGTA", "TTGACCTGAAGGCATCATCATCGGTA", "TTGACCTGAAGG---CATCATCGGTA", 885, 208),
    ("TTGACCTGAAGGCATCATCGGTA", "TTGACCTGAAGGCATCATCATCGGTA", "TTGACCTGAAGG---CATCATCGGTA", "TTGACCTGAAGGCATCATCATCGGTA", 885, 208),
]


@pytest.mark.parametrize("a,b,ref,qry,tenths,score_x2", MICRO)
def test_micro_known_answers(ctx, a, b, ref, qry, tenths, score_x2):
    score = score_x2 / 2.0
    recs, r, m, q = aligner.needle_align(ctx, a, [b])
    assert (r[0], q[0], int(recs["tenths"][0]), float(recs["score"][0])) == (ref, qry, tenths, score)


@pytest.mark.parametrize("La,n,read_len,sigma,n_rate,seed", [
    (250, 3000, 250, 0.0, 0.0, 1),      # cfg2 shape
    (250, 2000, None, 0.0, 0.0, 2),     # ragged lengths (indels change the read length)
    (280, 1500, None, 0.0, 0.01, 3),    # N in reads, tile 8x36
    (300, 1500, 300, 8.0, 0.0, 4),      # cfg3 shape: merged-PE like lengths
    (600, 600, 600, 0.0, 0.0, 5),       # cfg5 shape
    (64, 800, None, 0.0, 0.0, 6),       # smallest tile
    (1024, 100, 1000, 0.0, 0.0, 7),     # largest amplicon
    (150, 1000, 151, 0.0, 0.02, 8),
])
def test_synthetic_reads_match_oracle(ctx, La, n, read_len, sigma, n_rate, seed):
    amp, _g, cut, hdr = synth.make_case(seed, La)
    packed = synth.make_reads(amp, hdr, cut, n, seed=seed, read_len=read_len, len_sigma=sigma, n_rate=n_rate)
    _compare(ctx, amp, packed)


def test_unrelated_and_shifted_reads(ctx):
    """Junk reads, reads shifted along the amplicon (long end gaps, walks along row 0 / column 0),
    reverse-complement reads and very short reads."""
    rng = np.random.default_rng(9)
    amp = synth.random_seq(rng, 200)
    reads = [synth.random_seq(rng, int(rng.integers(2, 260))) for _ in range(300)]
    for _ in range(300):
        s = int(rng.integers(0, 150)); e = int(rng.integers(s + 2, 201))
        reads.append(synth.random_seq(rng, int(rng.integers(0, 40))) + amp[s:e] + synth.random_seq(rng, int(rng.integers(0, 40))))
    reads += [synth.revcomp(amp), amp[:2], amp[-2:], "AC", "NN", "N" * 50, amp.lower()]
    _compare(ctx, amp, aligner.pack_reads(reads))


def test_amplicon_with_n_and_low_complexity(ctx):
    rng = np.random.default_rng(10)
    amp = list(synth.random_seq(rng, 180)); amp[30] = "N"; amp[90] = "N"
    amp = "".join(amp[:100]) + "A" * 25 + "CACACACACACA" + "".join(amp[100:])
    reads = []
    for _ in range(600):
        s = list(amp.replace("N", "A"))
        for _k in range(int(rng.integers(0, 4))):
            p = int(rng.integers(1, len(s) - 1))
            if rng.random() < 0.5:
                del s[p:p + int(rng.integers(1, 12))]
            else:
                s[p:p] = list(s[max(0, p - 6):p])          # tandem duplication: ties in gap placement
        reads.append("".join(s))
    _compare(ctx, amp, aligner.pack_reads(reads))


@pytest.mark.parametrize("gapopen,gapextend", [(10.0, 0.5), (12.0, 2.0), (5.0, 0.0), (10.0, 0.25), (20.0, 1.5)])
def test_gap_penalties(ctx, gapopen, gapextend):
    amp, _g, cut, hdr = synth.make_case(11, 160)
    packed = synth.make_reads(amp, hdr, cut, 600, seed=11)
    _compare(ctx, amp, packed, gapopen, gapextend)


def test_batching_is_invisible(ctx):
    amp, _g, cut, hdr = synth.make_case(12, 250)
    packed = synth.make_reads(amp, hdr, cut, 4000, seed=12, read_len=250)
    ctx.set_traceback_budget(16 << 20)         # forces dozens of batches
    try:
        small = aligner.needle_align(ctx, amp, packed)
    finally:
        ctx.set_traceback_budget(8 << 30)
    big = aligner.needle_align(ctx, amp, packed)
    assert small[1:] == big[1:] and np.array_equal(small[0], big[0])


def test_errors_are_loud(ctx):
    # a read with an IUPAC code outside ACGTN(U) is not aligned, and says so: an empty record, the other reads unaffected
    recs, r, m, q = aligner.needle_align(ctx, "ACGTACGTAC", ["ACGTRYACGT", "ACGTACGTAC"])
    assert (int(recs["alnlen"][0]), r[0], q[0]) == (0, "", "") and (int(recs["tenths"][1]), q[1]) == (1000, "ACGTACGTAC")
    with pytest.raises(CrgpuError):
        aligner.needle_align(ctx, "ACGTRCGTAC", ["ACGTACGTAC"])          # ... in the AMPLICON it is an argument error
    with pytest.raises(CrgpuError):
        aligner.needle_align(ctx, "ACGTACGTAC", ["A"])                   # shorter than CRGPU_MIN_LEN
    with pytest.raises(CrgpuError):
        aligner.needle_align(ctx, "ACGTACGTAC", ["ACGTACGT"], 10.0, 0.3)  # not a dyadic penalty
    recs, r, m, q = aligner.needle_align(ctx, "ACGTACGTAC", [])
    assert len(recs) == 0 and r == []
