"""crgpu_quantify (k_quantify) against vectors from the UNMODIFIED reference process_df_chunk
(tests/golden/) and the drop-in process_df_chunk 22-tuple.  Needs a B200."""
import argparse

import numpy as np
import pandas as pd
import pytest

from crispresso_b200 import _lib, hotpath
from golden_io import load_process_df_chunk_cases
from oracle import quantify

pytestmark = pytest.mark.gpu
CASES = load_process_df_chunk_cases()


def _masks(case):
    L = case["L"]
    inc = np.zeros(L, np.uint8); inc[case["include"]] = 1
    exon = splice = None
    if case["opts"]["coding_seq"]:
        exon = np.zeros(L, np.uint8); exon[case["exon"]] = 1
        splice = np.zeros(L, np.uint8); splice[case["splice"]] = 1
    return inc, exon, splice


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_kernel_matches_reference_vectors(ctx, case):
    o = case["opts"]
    rows = case["rows"]
    inc, exon, splice = _masks(case)
    flags = hotpath.quant_flags(o["expected_hdr_amplicon_seq"], o["ignore_substitutions"], o["ignore_insertions"],
                                o["ignore_deletions"], o["window_around_sgrna"], o["hide_mutations_outside_window_NHEJ"],
                                o["coding_seq"])
    has_hdr = bool(o["expected_hdr_amplicon_seq"])
    recs, red = hotpath.quantify_rows(
        ctx, [r["ref_seq"] for r in rows], [r["align_str"] for r in rows], [r["align_seq"] for r in rows],
        [r["score_ref"] for r in rows], [r["score_repaired"] for r in rows] if has_hdr else None,
        [r["UNMODIFIED"] for r in rows], case["L"], flags, o["hdr_perfect_alignment_threshold"], inc, exon, splice)
    exp = case["expected"]
    for i, e in enumerate(exp["per_row"]):
        c = int(recs["cls"][i])
        got = dict(UNMODIFIED=bool(c & _lib.C_UNMODIFIED), NHEJ=bool(c & _lib.C_NHEJ), HDR=bool(c & _lib.C_HDR),
                   MIXED=bool(c & _lib.C_MIXED), n_mutated=int(recs["n_mutated"][i]),
                   n_inserted=int(recs["n_inserted"][i]), n_deleted=int(recs["n_deleted"][i]))
        assert got == e, (i, rows[i])
    for k, name in enumerate(hotpath.VECTOR_NAMES):
        assert red.vectors[k].tolist() == exp["vectors"][name], name
    assert {str(k): v for k, v in hotpath.Reductions.hist_dict(red.hist_inframe).items()} == exp["hist_inframe"]
    assert {str(k): v for k, v in hotpath.Reductions.hist_dict(red.hist_frameshift).items()} == exp["hist_frameshift"]
    assert {n: red.counter(n) for n in hotpath.COUNTER_NAMES} == exp["counters"]


def test_process_df_chunk_dropin_returns_reference_tuple(ctx):
    """Same call shape as the reference: process_df_chunk([df, args]) -> 22-tuple (CORE:730-753)."""
    case = next(c for c in CASES if c["name"] == "frameshift_hdr_w0")
    rows = case["rows"]
    df = pd.DataFrame({k: [r[k] for r in rows] for k in ("score_ref", "score_repaired", "ref_seq", "align_str", "align_seq",
                                                         "UNMODIFIED")}, index=["read%d" % i for i in range(len(rows))])
    for col in ("MIXED", "HDR", "NHEJ"):
        df[col] = False
    for col in ("n_mutated", "n_inserted", "n_deleted"):
        df[col] = 0
    args = argparse.Namespace(**case["opts"])
    out = hotpath.process_df_chunk([df, args], ctx, set(case["include"]), case["L"], case["exon"], set(case["splice"]))
    assert len(out) == 22
    exp = case["expected"]
    d = out[0]
    assert list(d.index) == list(df.index)
    assert [bool(x) for x in d["NHEJ"]] == [p["NHEJ"] for p in exp["per_row"]]
    assert [bool(x) for x in d["HDR"]] == [p["HDR"] for p in exp["per_row"]]
    assert [int(x) for x in d["n_deleted"]] == [p["n_deleted"] for p in exp["per_row"]]
    order = list(hotpath.VECTOR_NAMES[:13])
    for k, name in enumerate(order):
        assert out[1 + k].dtype == np.float64 and out[1 + k].tolist() == [float(x) for x in exp["vectors"][name]], name
    assert {str(k): v for k, v in out[14].items()} == exp["hist_inframe"]
    assert {str(k): v for k, v in out[15].items()} == exp["hist_frameshift"]
    assert out[16].tolist() == [float(x) for x in exp["vectors"]["avg_vector_del_all"]]
    assert out[17].tolist() == [float(x) for x in exp["vectors"]["avg_vector_ins_all"]]
    assert (out[18], out[19], out[20], out[21]) == tuple(exp["counters"][n] for n in hotpath.COUNTER_NAMES)


def test_empty_chunk(ctx):
    recs, red = hotpath.quantify_rows(ctx, [], [], [], [], None, [], 50, hotpath.quant_flags(), 98.0, np.ones(50, np.uint8))
    assert len(recs) == 0 and red.vectors.sum() == 0
