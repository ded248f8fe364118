"""The reference runs of the drop-in tests: arguments as tests/crispresso_tests.py builds them."""
import os

import kat_common as K

from . import harness

DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden", "ref_test_data")

# The reads of the fixtures carry no HDR allele: tests/golden/make_dropin_requests.py writes the HDR product's three
# substitutions (outside both guides) into read 1 of every 20th pair of the sub-sampled files (hdr_sub_*).
HDR_POS = (125, 128)
HDR_AMPLICON = (K.AMPLICON[:HDR_POS[0]] + "".join({"A": "C", "C": "G", "G": "T", "T": "A"}[c] for c in K.AMPLICON[HDR_POS[0]:HDR_POS[1]])
                + K.AMPLICON[HDR_POS[1]:])

RUNS = {
    # tests/crispresso_tests.py:127-195 (known-answer test #1)
    "kat1": dict(r1="test_L001_R1_001.fastq.gz", r2="test_L001_R2_001.fastq.gz", guides=K.GUIDES, extra={}),
    # tests/crispresso_tests.py:198-272 (known-answer test #2) WITHOUT Trimmomatic (no java here): the golden values that
    # do not depend on trimming -- n_unmodified 2647, alleles (184, 68, 44, 26), deletion fq[0] 3359 -- are asserted
    "kat2_untrimmed": dict(r1="test1_L001_R1_001.fastq.gz", r2="test1_L001_R2_001.fastq.gz",
                           guides="cgagaagcgactcgacatgg,aaggggctaacttggtccct",
                           extra=dict(window_around_sgrna=23, min_identity_score=30.0)),
    # BASELINE.json configs[0] / SURVEY 8d cfg1: single end, 151-bp reads vs the 280-bp amplicon (identity capped at
    # 53.9 %, hence --min_identity_score 50); first 3000 records
    "cfg1_single_end": dict(r1="sub_test_L001_R1_001.fastq.gz", r2="", guides=K.GUIDES, extra=dict(min_identity_score=50.0)),
    # HDR amplicon + coding sequence on the first 3000 pairs: the repair passes, the reverse-complement rescue with its
    # dying repair-RC needle (CORE:1924-1936), frameshift analysis -- none of which the reference's own tests reach
    "hdr_coding": dict(r1="hdr_sub_test_L001_R1_001.fastq.gz", r2="sub_test_L001_R2_001.fastq.gz", guides=K.GUIDES,
                       extra=dict(expected_hdr_amplicon_seq=HDR_AMPLICON.lower(), coding_seq=K.AMPLICON[80:170].lower(),
                                  min_identity_score=55.0)),
}


def run(name, bin_dir, out_dir, n_processes=1, keep_intermediate=False):
    r = RUNS[name]
    return harness.run_reference(bin_dir, out_dir, os.path.join(DATA, r["r1"]), os.path.join(DATA, r["r2"]) if r["r2"] else "",
                                 r.get("amplicon", K.AMPLICON).lower(), r["guides"], n_processes=n_processes,
                                 keep_intermediate=keep_intermediate, extra=r["extra"])


def summarize(out):
    """The 14-tuple of run_crispresso (CORE:3977-3992) as plain numbers: 9 scalars, the first 8 `fq` values of the four
    histograms, the first 8 allele counts."""
    names = ("n_total", "n_reads_input", "n_unmodified", "n_mixed_hdr_nhej", "n_modified", "n_repaired", "nhej_inserted",
             "nhej_deleted", "nhej_mutated")
    d = {k: int(v) for k, v in zip(names, out[:9])}
    for k, df in zip(("indels_fq", "insertion_fq", "deletion_fq", "substitution_fq"), out[9:13]):
        d[k] = [int(x) for x in df["fq"].values[:8]]
    d["alleles"] = [int(x) for x in out[13]["#Reads"].values[:8]]
    return d
