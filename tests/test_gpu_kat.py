"""The reference's own golden values through the CUDA path.  Needs a B200."""
import gzip
import os

import numpy as np
import pytest

import kat_common as K
from crispresso_b200 import _lib, aligner, fastq as gfastq, hotpath, postreduce

pytestmark = pytest.mark.gpu


def test_end_to_end_known_answer_on_gpu(ctx):
    reads = K.merged_reads()
    cuts = hotpath.cut_points_from_guides(K.AMPLICON, K.GUIDES)
    inc = hotpath.include_mask(len(K.AMPLICON), cuts, 1, 15, 15)
    res = hotpath.run_hot_path(ctx, K.AMPLICON, aligner.pack_reads(reads), inc=inc, flags=hotpath.quant_flags(), want_rows=True)
    G = K.GOLDEN
    red = res.red
    assert red.n_total == G["n_total"]
    assert red.class_counts.tolist() == [G["n_unmodified"], G["n_modified"], G["n_repaired"], G["n_mixed_hdr_nhej"]]
    kept = (res.kept & 1) != 0
    cls, ni, nd, nm = res.recs["cls"][kept], res.recs["n_inserted"][kept], res.recs["n_deleted"][kept], res.recs["n_mutated"][kept]
    ev = postreduce.class_event_counts(cls, ni, nd, nm)
    assert (ev["nhej_inserted"], ev["nhej_deleted"], ev["nhej_mutated"]) == (G["nhej_inserted"], G["nhej_deleted"], G["nhej_mutated"])
    _hl, hd = postreduce.indel_size_histogram(ni, nd, len(K.AMPLICON), cuts, True)
    assert tuple(hd[:4]) == G["indels_fq"]
    (xi, yi), (xd, yd), (xs, ys) = postreduce.event_size_histograms(ni, nd, nm)
    assert tuple(yi[:4]) == G["insertion_fq"] and tuple(yd[:4]) == G["deletion_fq"] and tuple(ys[:4]) == G["substitution_fq"]
    df = hotpath.build_dataframe(res, ["r%d" % i for i in range(len(reads))], amplicon=K.AMPLICON)
    assert tuple(postreduce.allele_table(df)["#Reads"].values[:4]) == G["alleles"]
    combined, avg_ins, avg_del = postreduce.normalise_vectors(red)
    assert combined.shape == (len(K.AMPLICON),) and np.isfinite(avg_del).all()


def test_quality_filter_known_answer_on_gpu(ctx, tmp_path):
    sub = K.qual_subset()
    for (name, q), golden in K.QUAL_GOLDEN.items():
        keep = gfastq.keep_mask(ctx, sub[name]["quals"], q, 0)
        assert set(i for i, k in zip(sub[name]["ids"], keep) if not k) == golden
    # the file-level drop-ins (CORE:196-310)
    ids, quals = sub["R1"]["ids"][:200], sub["R1"]["quals"][:200]
    p = str(tmp_path / "x.fastq.gz")
    with gzip.open(p, "wt") as f:
        for i, ql in zip(ids, quals):
            f.write("@%s 1:N:0:1\n%s\n+\n%s\n" % (i, "A" * len(ql), ql))
    out = gfastq.filter_se_fastq_by_qual(ctx, p, min_bp_quality=30, min_single_bp_quality=5)
    assert out.endswith("x_filtered.fastq.gz") and os.path.exists(out)
    removed = gfastq.get_ids_reads_to_remove(ctx, p, 30, 5)
    kept_ids = [h.split()[0] for h in gfastq.read_fastq(out)[0]]
    assert set(kept_ids) == set(ids) - removed and len(kept_ids) == len(ids) - len(removed)
    o1, o2 = gfastq.filter_pe_fastq_by_qual(ctx, p, p, str(tmp_path / "a.gz"), str(tmp_path / "b.gz"), 30, 5)
    assert [h.split()[0] for h in gfastq.read_fastq(o1)[0]] == kept_ids


def test_quality_filter_edge_cases(ctx):
    quals = ["", "I", "!" * 300, "I" * 299 + "!", "5" * 10]
    keep = gfastq.keep_mask(ctx, quals, 20, 0)
    assert keep.tolist() == [0, 1, 0, 1, 1]            # empty read dropped; mean of the 4th is 39.87
    keep = gfastq.keep_mask(ctx, quals, 20, 1)
    assert keep.tolist() == [0, 1, 0, 0, 1]
    assert gfastq.keep_mask(ctx, [], 20, 0).tolist() == []
