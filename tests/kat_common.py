"""The reference's end-to-end known-answer test #1 (tests/crispresso_tests.py:127-195): inputs and
golden values, shared by the CPU (oracle) and GPU tests."""
import gzip
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))

AMPLICON = (
    "gtcgcccctcaaatcttacagctgctcactc" "ccctgcagggcaacgcccagggaccaagttag" "ccccttaagcctaggcaaaagaatcccgccca"
    "taatcgagaagcgactcgacatggaggcgatg" "acgagatcacgcgaggaggaaaggagggaggg" "cttcttccaggcccagggcggtccttacaaga"
    "cgggaggcagcagagaactcccataaaggtat" "tgcggcactcccctccccctgcccagaagggt" "gcggccttctctccacctcctccac").upper()
GUIDES = "aatcgagaagcgactcgaca,taaggggctaacttggtccc"

GOLDEN = dict(n_total=7058, n_unmodified=6853, n_mixed_hdr_nhej=0, n_modified=205, n_repaired=0, nhej_inserted=0,
              nhej_deleted=12, nhej_mutated=193, indels_fq=(1, 0, 0, 0), insertion_fq=(7058, 0, 0, 0),
              deletion_fq=(7046, 0, 0, 0), substitution_fq=(6865, 188, 5, 0), alleles=(1098, 346, 19, 17))

QUAL_GOLDEN = {  # tests/crispresso_tests.py:78-88
    ("R1", 23): {"M06879:15:000000000-DFF22:1:1101:25894:23776", "M06879:15:000000000-DFF22:1:1101:24046:20708"},
    ("R2", 15): {"M06879:15:000000000-DFF22:1:1102:22078:15849"},
}


def merged_reads():
    with gzip.open(os.path.join(HERE, "golden", "kat1_merged_reads.json.gz"), "rt") as f:
        d = json.load(f)
    reads = []
    for s, c in zip(d["reads"], d["counts"]):
        reads += [s] * c
    assert len(reads) == d["n_merged"] == 8092
    return reads


def qual_subset():
    with gzip.open(os.path.join(HERE, "golden", "qualfilter_subset.json.gz"), "rt") as f:
        return json.load(f)
