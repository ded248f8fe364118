"""Generate tests/golden/process_df_chunk_cases.json.gz by running the UNMODIFIED reference
CRISPResso/CRISPRessoCORE.py::process_df_chunk (CORE:428-753), imported through tests/ref_shim.py,
on alignment rows produced by the oracle aligner.  Run in the build container only:

    python tests/golden/make_golden.py

The fixture pins oracle/quantify.py::process_rows and the GPU k_quantify kernel.
"""
import argparse
import gzip
import json
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

import ref_shim  # noqa: E402
from oracle import needle, quantify  # noqa: E402

ACGT = "ACGT"


def random_seq(rng, n):
    return "".join(ACGT[i] for i in rng.integers(0, 4, size=n))


def mutate(rng, amp, cut, n_rate=0.0):
    s = list(amp)
    k = int(rng.choice([0, 1, 1, 1, 2, 2, 3]))
    for _ in range(k):
        kind = rng.choice(["sub", "ins", "del"])
        near = rng.random() < 0.7
        pos = int(np.clip(cut + rng.integers(-8, 9), 1, len(s) - 2)) if near else int(rng.integers(1, max(2, len(s) - 1)))
        if kind == "sub":
            for j in range(pos, min(len(s), pos + int(rng.integers(1, 4)))):
                s[j] = ACGT[(ACGT.index(s[j]) + int(rng.integers(1, 4))) % 4] if s[j] in ACGT else "A"
        elif kind == "ins":
            s[pos:pos] = list(random_seq(rng, int(rng.integers(1, 13))))
        else:
            del s[pos:pos + int(rng.integers(1, 31))]
    r = rng.random()
    if r < 0.08:
        s = list(random_seq(rng, int(rng.integers(1, 9)))) + s          # left overhang -> leading insertion
    elif r < 0.16:
        s = s + list(random_seq(rng, int(rng.integers(1, 9))))          # right overhang -> trailing insertion
    elif r < 0.24:
        s = s[int(rng.integers(1, 20)):]                                # truncated -> leading "deletion"
    elif r < 0.32:
        s = s[:len(s) - int(rng.integers(1, 20))]
    if n_rate:
        s = [("N" if rng.random() < n_rate else c) for c in s]
    return "".join(s)


def build_case(name, seed, L=120, nreads=220, hdr=False, coding=False, amp_n=False, opts=None, window=1,
               excl=(15, 15), guides=True, min_id=60.0):
    rng = np.random.default_rng(seed)
    amp = random_seq(rng, L)
    cut = L // 2
    if amp_n:
        a = list(amp); a[cut + 7] = "N"; a[20] = "N"; amp = "".join(a)
    hdr_amp = ""
    if hdr:
        blk = "".join(ACGT[(ACGT.index(c) + 1) % 4] if c in ACGT else c for c in amp[cut - 3:cut + 3])
        hdr_amp = amp[:cut - 3] + blk + amp[cut + 3:]
    reads = []
    for _ in range(nreads):
        src = amp
        u = rng.random()
        if hdr and u < 0.3:
            src = hdr_amp
        reads.append(mutate(rng, src, cut, n_rate=0.01 if amp_n else 0.0) if rng.random() < 0.8 else src)
    o = dict(coding_seq="", ignore_substitutions=False, ignore_deletions=False, ignore_insertions=False,
             expected_hdr_amplicon_seq=hdr_amp, hdr_perfect_alignment_threshold=98.0,
             hide_mutations_outside_window_NHEJ=False, window_around_sgrna=window)
    o.update(opts or {})
    exon, splice = [], []
    if coding:
        st, en = cut - 25, cut + 20
        o["coding_seq"] = amp[st:en]
        exon = list(range(st, en))
        splice = sorted(set([max(0, st - 2), max(0, st - 1), min(L - 1, en), min(L - 1, en + 1)]) - set(exon))
    # INCLUDE_IDXS (CORE:2739-2762)
    if guides and window > 0:
        half = max(1, window // 2)
        inc = list(range(max(0, cut - half + 1), min(L - 1, cut + half + 1)))
    else:
        inc = list(range(L))
    ex = list(range(excl[0])) + (list(range(L))[-excl[1]:] if excl[1] else [])
    inc = sorted(set(inc) - set(ex))

    # alignment rows exactly as run_crispresso prepares them (CORE:1830-2072), via the oracle aligner
    res, ref, mark, qry = needle.align_batch(amp, reads, use_int=True)
    sref = res["tenths"] / 10.0
    if hdr:
        rh, _, _, _ = needle.align_batch(hdr_amp, reads, use_int=True)
        srep = rh["tenths"] / 10.0
    rows = []
    for i in range(nreads):
        if not (sref[i] > min_id or (hdr and srep[i] > min_id)):
            continue
        r = dict(ref_seq=ref[i], align_str=mark[i], align_seq=qry[i], score_ref=float(sref[i]))
        if hdr:
            r["score_repaired"] = float(srep[i])
            if rng.random() < 0.05:
                r["score_repaired"] = float("nan")          # what the reference produces for _RC rows (Q11)
        r["UNMODIFIED"] = bool(r["score_ref"] == 100)
        if "N" in amp:
            r["align_str"], uni = quantify.mask_n(r["ref_seq"], r["align_str"])
            if uni:
                r["UNMODIFIED"] = True
        rows.append(r)
    return dict(name=name, amplicon=amp, L=L, opts=o, include=inc, exon=exon, splice=splice, rows=rows)


def run_reference(case):
    core = ref_shim.load_core()
    rows = case["rows"]
    df = pd.DataFrame({
        "score_ref": [r["score_ref"] for r in rows],
        "ref_seq": [r["ref_seq"] for r in rows],
        "align_str": [r["align_str"] for r in rows],
        "align_seq": [r["align_seq"] for r in rows],
    }, index=["read%d" % i for i in range(len(rows))])
    if case["opts"]["expected_hdr_amplicon_seq"]:
        df["score_repaired"] = [r["score_repaired"] for r in rows]
        df["score_diff"] = df.score_ref - df.score_repaired
    df["UNMODIFIED"] = [r["UNMODIFIED"] for r in rows]
    df["MIXED"] = False
    df["HDR"] = False
    df["NHEJ"] = False
    df["n_mutated"] = 0
    df["n_inserted"] = 0
    df["n_deleted"] = 0
    df["ref_positions"] = df["ref_seq"].apply(lambda s: np.array(quantify.ref_positions(s)))
    core.INCLUDE_IDXS = set(np.array(case["include"], dtype=np.int64))
    core.LEN_AMPLICON = case["L"]
    core.EXON_POSITIONS = sorted(case["exon"])
    core.SPLICING_POSITIONS = set(case["splice"])
    args = argparse.Namespace(**case["opts"])
    out = core.process_df_chunk([df, args])
    d = out[0]
    names = quantify.VECTOR_NAMES
    vec = {}
    order = list(range(1, 14))
    for k, idx in zip(names[:13], order):
        vec[k] = [int(x) for x in out[idx]]
    vec["avg_vector_del_all"] = [int(x) for x in out[16]]
    vec["avg_vector_ins_all"] = [int(x) for x in out[17]]
    exp = dict(
        per_row=[dict(UNMODIFIED=bool(d["UNMODIFIED"].iloc[i]), NHEJ=bool(d["NHEJ"].iloc[i]), HDR=bool(d["HDR"].iloc[i]),
                      MIXED=bool(d["MIXED"].iloc[i]), n_mutated=int(d["n_mutated"].iloc[i]),
                      n_inserted=int(d["n_inserted"].iloc[i]), n_deleted=int(d["n_deleted"].iloc[i]))
                 for i in range(len(rows))],
        vectors=vec,
        hist_inframe={str(k): int(v) for k, v in out[14].items()},
        hist_frameshift={str(k): int(v) for k, v in out[15].items()},
        counters=dict(modified_frameshift=int(out[18]), modified_non_frameshift=int(out[19]),
                      non_modified_non_frameshift=int(out[20]), splicing_sites_modified=int(out[21])))
    return exp


def main():
    cases = [
        build_case("default_window1", 11),
        build_case("no_guides_window1", 12, guides=False),
        build_case("window0", 13, window=0),
        build_case("window20", 14, window=20, excl=(5, 5)),
        build_case("hdr", 15, hdr=True),
        build_case("hdr_hide", 16, hdr=True, opts=dict(hide_mutations_outside_window_NHEJ=True)),
        build_case("hide_outside", 17, opts=dict(hide_mutations_outside_window_NHEJ=True), window=6),
        build_case("frameshift", 18, coding=True, window=10),
        build_case("frameshift_hdr_w0", 19, coding=True, hdr=True, window=0),
        build_case("ignore_subs", 20, opts=dict(ignore_substitutions=True)),
        build_case("ignore_ins", 21, opts=dict(ignore_insertions=True), window=8),
        build_case("ignore_del", 22, opts=dict(ignore_deletions=True), window=8, coding=True),
        build_case("amplicon_with_N", 23, amp_n=True, window=12),
        build_case("low_identity", 24, min_id=30.0, window=4, excl=(0, 0)),
    ]
    for c in cases:
        c["expected"] = run_reference(c)
        for r in c["rows"]:
            if "score_repaired" in r and r["score_repaired"] != r["score_repaired"]:
                r["score_repaired"] = None            # NaN -> null in JSON
        n = len(c["rows"])
        e = c["expected"]
        print("%-20s rows %3d  NHEJ %3d HDR %3d MIXED %3d UNMOD %3d  inframe %s frameshift %s" % (
            c["name"], n, sum(p["NHEJ"] for p in e["per_row"]), sum(p["HDR"] for p in e["per_row"]),
            sum(p["MIXED"] for p in e["per_row"]), sum(p["UNMODIFIED"] for p in e["per_row"]),
            len(e["hist_inframe"]), len(e["hist_frameshift"])))
    path = os.path.join(HERE, "process_df_chunk_cases.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(dict(source="CRISPResso/CRISPRessoCORE.py::process_df_chunk (unmodified reference)", cases=cases), f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
