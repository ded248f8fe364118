"""The fused hot path (crgpu_align_quantify, or the CPU oracle) on the inputs of a recorded reference run, summarised
like run_crispresso's 14-tuple (CORE:3977-3992) so that it can be compared with what the UNMODIFIED reference
returned for the same run (tests/golden/dropin_requests/expected.json)."""
import json
import os

import numpy as np
import pandas as pd

import kat_common as K
from crispresso_b200 import _lib, aligner, hotpath, postreduce

from . import harness, runs

REQ = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "golden", "dropin_requests")
DEFAULTS = dict(window_around_sgrna=1, min_identity_score=60.0, exclude_bp_from_left=15, exclude_bp_from_right=15,
                hdr_perfect_alignment_threshold=98.0, expected_hdr_amplicon_seq="", coding_seq="")


def expected():
    with open(os.path.join(REQ, "expected.json")) as f:
        return json.load(f)


def run_inputs(name):
    """(amplicon, read names, reads, settings, cut points) of a run: the reads are what the reference piped into its
    first needle call (after its own FASTQ handling / the FLASH step)."""
    exp = expected()[name]
    first = [k for k in exp["requests"] if k.startswith("needle_")][0]
    names, seqs, amp, _argv = harness.request_reads(os.path.join(REQ, first + ".req.json.gz"))
    cfg = dict(DEFAULTS)
    cfg.update(runs.RUNS[name]["extra"])
    cuts = hotpath.cut_points_from_guides(amp.upper(), runs.RUNS[name]["guides"])
    return amp.upper(), names, seqs, cfg, cuts


def _summary(n_total, classes, cls, ni, nd, nm, df, L, cuts):
    ev = postreduce.class_event_counts(cls, ni, nd, nm)
    d = dict(n_total=int(n_total), n_unmodified=int(classes[0]), n_modified=int(classes[1]), n_repaired=int(classes[2]),
             n_mixed_hdr_nhej=int(classes[3]), nhej_inserted=ev["nhej_inserted"], nhej_deleted=ev["nhej_deleted"],
             nhej_mutated=ev["nhej_mutated"])
    _hl, hd = postreduce.indel_size_histogram(ni, nd, L, cuts, True)
    d["indels_fq"] = [int(x) for x in hd[:8]]
    for key, (_x, y) in zip(("insertion_fq", "deletion_fq", "substitution_fq"), postreduce.event_size_histograms(ni, nd, nm)):
        d[key] = [int(v) for v in y[:8]]
    d["alleles"] = [int(x) for x in postreduce.allele_table(df)["#Reads"].values[:8]]
    return d


def _geometry(amp, cfg, cuts):
    inc = hotpath.include_mask(len(amp), cuts, cfg["window_around_sgrna"], cfg["exclude_bp_from_left"], cfg["exclude_bp_from_right"])
    exon = splice = None
    if cfg["coding_seq"]:
        exon, splice = hotpath.exon_masks(amp, cfg["coding_seq"].upper())
    return inc, exon, splice


def summarize_gpu(ctx, name):
    amp, names, seqs, cfg, cuts = run_inputs(name)
    hdr = cfg["expected_hdr_amplicon_seq"].upper()
    inc, exon, splice = _geometry(amp, cfg, cuts)
    flags = hotpath.quant_flags(hdr, window_around_sgrna=cfg["window_around_sgrna"], coding_seq=cfg["coding_seq"])
    res = hotpath.run_hot_path(ctx, amp, aligner.pack_reads(seqs), min_identity_score=cfg["min_identity_score"],
                               hdr_amplicon=hdr or None, flags=flags, hdr_thr=cfg["hdr_perfect_alignment_threshold"],
                               inc=inc, exon=exon, splice=splice, want_rows=True)
    df = hotpath.build_dataframe(res, names, has_hdr=bool(hdr), amplicon=amp)
    cls = ((df["UNMODIFIED"].values * _lib.C_UNMODIFIED) | (df["NHEJ"].values * _lib.C_NHEJ) | (df["HDR"].values * _lib.C_HDR)
           | (df["MIXED"].values * _lib.C_MIXED)).astype(np.uint8)
    d = _summary(res.red.n_total, res.red.class_counts, cls, df["n_inserted"].values, df["n_deleted"].values,
                 df["n_mutated"].values, df, len(amp), cuts)
    return d, res


def summarize_oracle(name):
    from oracle import quantify
    amp, names, seqs, cfg, cuts = run_inputs(name)
    hdr = cfg["expected_hdr_amplicon_seq"].upper()
    inc, exon, splice = _geometry(amp, cfg, cuts)
    opts = quantify.Opts(coding_seq=cfg["coding_seq"], expected_hdr_amplicon_seq=hdr,
                         hdr_perfect_alignment_threshold=cfg["hdr_perfect_alignment_threshold"],
                         window_around_sgrna=cfg["window_around_sgrna"])
    res = quantify.hot_path(amp, seqs, names=names, min_identity_score=cfg["min_identity_score"], hdr_amplicon=hdr, opts=opts,
                            include=np.nonzero(inc)[0], exon=np.nonzero(exon)[0] if exon is not None else (),
                            splice=np.nonzero(splice)[0] if splice is not None else (), nthreads=os.cpu_count() or 1)
    pr = res["per_row"]
    cls = np.array([(_lib.C_UNMODIFIED if p["UNMODIFIED"] else 0) | (_lib.C_NHEJ if p["NHEJ"] else 0)
                    | (_lib.C_HDR if p["HDR"] else 0) | (_lib.C_MIXED if p["MIXED"] else 0) for p in pr], np.uint8)
    ni = np.array([p["n_inserted"] for p in pr], np.int64)
    nd = np.array([p["n_deleted"] for p in pr], np.int64)
    nm = np.array([p["n_mutated"] for p in pr], np.int64)
    df = pd.DataFrame(dict(align_seq=[r["align_seq"] for r in res["rows"]], ref_seq=[r["ref_seq"] for r in res["rows"]],
                           NHEJ=[p["NHEJ"] for p in pr], UNMODIFIED=[p["UNMODIFIED"] for p in pr], HDR=[p["HDR"] for p in pr],
                           n_deleted=nd, n_inserted=ni, n_mutated=nm))
    c = res["classes"]
    return _summary(res["n_total"], [c["UNMODIFIED"], c["NHEJ"], c["HDR"], c["MIXED"]], cls, ni, nd, nm, df, len(amp), cuts)


def assert_matches_reference(name, got):
    exp = expected()[name]
    for k, v in got.items():
        assert exp[k] == v, (name, k, exp[k], v)
    if name == "kat1":                                    # tests/crispresso_tests.py:181-195
        G = K.GOLDEN
        for k in ("n_total", "n_unmodified", "n_mixed_hdr_nhej", "n_modified", "n_repaired", "nhej_inserted", "nhej_deleted", "nhej_mutated"):
            assert got[k] == G[k], k
        for k in ("indels_fq", "insertion_fq", "deletion_fq", "substitution_fq", "alleles"):
            assert tuple(got[k][:4]) == G[k], k
    if name == "kat2_untrimmed":                          # the trimming-independent values of tests/crispresso_tests.py:258-272
        assert got["n_unmodified"] == 2647 and tuple(got["alleles"][:4]) == (184, 68, 44, 26) and got["deletion_fq"][0] == 3359
        assert tuple(got["indels_fq"][:4]) == (2, 4, 5, 5)
