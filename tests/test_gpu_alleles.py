"""Device-side allele table (crgpu_path_out.allele_*) against the reference's pandas groupby
(CRISPResso/CRISPRessoCORE.py:2923-2946) and the reference's own golden top-4 counts.  Needs a B200."""
import numpy as np
import pytest

import kat_common as K
from crispresso_b200 import aligner, hotpath, postreduce, synth

pytestmark = pytest.mark.gpu
KEY = ["Aligned_Sequence", "Reference_Sequence", "NHEJ", "UNMODIFIED", "HDR", "n_deleted", "n_inserted", "n_mutated", "#Reads"]


def _as_set(df):
    return sorted(tuple((bool(v) if isinstance(v, (bool, np.bool_)) else (int(v) if not isinstance(v, str) else v)) for v in row)
                  for row in df[KEY].itertuples(index=False, name=None))


def _both(ctx, amp, packed, hdr=None, min_id=60.0, want_rows=False):
    n = len(packed[1]) - 1
    res_rows = hotpath.run_hot_path(ctx, amp, packed, hdr_amplicon=hdr, flags=hotpath.quant_flags(hdr or ""), min_identity_score=min_id,
                                    want_rows=True)
    ref = postreduce.allele_table(hotpath.build_dataframe(res_rows, ["r%d" % i for i in range(n)], has_hdr=bool(hdr), amplicon=amp))
    res = hotpath.run_hot_path(ctx, amp, packed, hdr_amplicon=hdr, flags=hotpath.quant_flags(hdr or ""), min_identity_score=min_id,
                               want_rows=want_rows, alleles=2 * n + 8)
    got = hotpath.allele_table(ctx, res, amp, packed)
    return got, ref, res


def test_reference_kat_allele_counts(ctx):
    packed = aligner.pack_reads(K.merged_reads())
    inc = hotpath.include_mask(len(K.AMPLICON), hotpath.cut_points_from_guides(K.AMPLICON, K.GUIDES), 1, 15, 15)
    res = hotpath.run_hot_path(ctx, K.AMPLICON, packed, inc=inc, flags=hotpath.quant_flags(), alleles=16)
    assert tuple(int(c) for c in res.allele_count[:4]) == K.GOLDEN["alleles"]          # tests/crispresso_tests.py:195
    assert res.allele_n > 1000 and len(res.allele_row) == 16
    full = hotpath.run_hot_path(ctx, K.AMPLICON, packed, inc=inc, flags=hotpath.quant_flags(), alleles=9000)
    assert int(full.allele_count.sum()) == K.GOLDEN["n_total"] and len(full.allele_count) == full.allele_n
    assert (np.diff(full.allele_count) <= 0).all()


@pytest.mark.parametrize("want_rows", [False, True])
def test_matches_pandas_groupby(ctx, want_rows):
    amp, guide, cut, hdr = synth.make_case(61, 220)
    buf, off = synth.make_reads(amp, hdr, cut, 3000, seed=61, rc_frac=0.15, sub_rate=0.001)
    # duplicates (the point of an allele table) + lower-case copies (distinct alleles: align_seq keeps the case)
    reads = [bytes(buf[off[i]:off[i + 1]]).decode() for i in range(3000)]
    reads = reads + reads[:1000] + [r.lower() for r in reads[:50]]
    packed = aligner.pack_reads(reads)
    for h in (None, hdr):
        got, ref, res = _both(ctx, amp, packed, hdr=h, want_rows=want_rows)
        assert res.allele_n == len(ref)
        assert _as_set(got) == _as_set(ref)
        assert list(got["#Reads"]) == sorted(ref["#Reads"], reverse=True)
        assert np.allclose(sorted(got["%Reads"]), sorted(ref["%Reads"]))


def test_forward_and_rc_rows_with_the_same_text_merge(ctx):
    """A read and the reverse complement of the same molecule give identical text rows (when gap
    placement agrees) and identical classes without HDR: pandas puts them in one group."""
    amp, guide, cut, _ = synth.make_case(62, 160, hdr=False)
    fw = [amp, amp[:80] + "T" + amp[81:], amp[:70] + amp[78:]]
    reads = fw * 3 + [synth.revcomp(r) for r in fw] * 2
    got, ref, res = _both(ctx, amp, aligner.pack_reads(reads))
    assert _as_set(got) == _as_set(ref)
    assert res.allele_n == len(ref)
