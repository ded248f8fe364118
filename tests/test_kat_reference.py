"""Pins the oracle against the reference's OWN golden values (CPU only):
  * tests/crispresso_tests.py:181-195 -- the end-to-end KAT, reached through the FLASH restatement
    (fixture built by tests/golden/make_kat_fixture.py);
  * tests/crispresso_tests.py:78-88  -- the quality filter's read ids.
"""
import os

import numpy as np
import pandas as pd
import pytest

import kat_common as K
from crispresso_b200 import _lib, hotpath, postreduce
from oracle import fastq, quantify


@pytest.fixture(scope="module")
def kat_result():
    reads = K.merged_reads()
    cuts = hotpath.cut_points_from_guides(K.AMPLICON, K.GUIDES)
    assert cuts == [112, 52]
    inc = hotpath.include_mask(len(K.AMPLICON), cuts, 1, 15, 15)
    return quantify.hot_path(K.AMPLICON, reads, include=np.nonzero(inc)[0], nthreads=os.cpu_count() or 1), cuts


def test_end_to_end_known_answer(kat_result):
    res, cuts = kat_result
    G = K.GOLDEN
    assert res["n_total"] == G["n_total"]
    assert res["classes"]["UNMODIFIED"] == G["n_unmodified"]
    assert res["classes"]["NHEJ"] == G["n_modified"]
    assert res["classes"]["HDR"] == G["n_repaired"] and res["classes"]["MIXED"] == G["n_mixed_hdr_nhej"]
    pr = res["per_row"]
    cls = np.array([(_lib.C_UNMODIFIED if p["UNMODIFIED"] else 0) | (_lib.C_NHEJ if p["NHEJ"] else 0) for p in pr], np.uint8)
    ni = np.array([p["n_inserted"] for p in pr]); nd = np.array([p["n_deleted"] for p in pr]); nm = np.array([p["n_mutated"] for p in pr])
    ev = postreduce.class_event_counts(cls, ni, nd, nm)
    assert (ev["nhej_inserted"], ev["nhej_deleted"], ev["nhej_mutated"]) == (G["nhej_inserted"], G["nhej_deleted"], G["nhej_mutated"])
    _hl, hd = postreduce.indel_size_histogram(ni, nd, len(K.AMPLICON), cuts, True)
    assert tuple(hd[:4]) == G["indels_fq"]
    (xi, yi), (xd, yd), (xs, ys) = postreduce.event_size_histograms(ni, nd, nm)
    assert tuple(yi[:4]) == G["insertion_fq"] and tuple(yd[:4]) == G["deletion_fq"] and tuple(ys[:4]) == G["substitution_fq"]
    df = pd.DataFrame(dict(align_seq=[r["align_seq"] for r in res["rows"]], ref_seq=[r["ref_seq"] for r in res["rows"]],
                           NHEJ=[p["NHEJ"] for p in pr], UNMODIFIED=[p["UNMODIFIED"] for p in pr], HDR=[p["HDR"] for p in pr],
                           n_deleted=nd, n_inserted=ni, n_mutated=nm))
    assert tuple(postreduce.allele_table(df)["#Reads"].values[:4]) == G["alleles"]


def test_quality_filter_known_answer_subset():
    sub = K.qual_subset()
    for (name, q), golden in K.QUAL_GOLDEN.items():
        recs = [(i, "", ql) for i, ql in zip(sub[name]["ids"], sub[name]["quals"])]
        assert fastq.ids_to_remove(recs, q) == golden


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests/test_data"), reason="reference fixtures only exist in the build container")
def test_quality_filter_known_answer_full_files():
    d = "/root/reference/tests/test_data/"
    assert fastq.ids_to_remove(fastq.read_fastq(d + "test_L001_R1_001.fastq.gz"), 23) == K.QUAL_GOLDEN[("R1", 23)]
    assert fastq.ids_to_remove(fastq.read_fastq(d + "test_L001_R2_001.fastq.gz"), 15) == K.QUAL_GOLDEN[("R2", 15)]
    assert len(fastq.read_fastq(d + "test_L001_R1_001.fastq.gz")) == 8906            # tests/crispresso_tests.py:30-33
