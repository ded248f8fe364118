"""CPU restatement of the Python half of CRISPResso's hot path -- TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py).

Follows, line by line in meaning but not in code, CRISPResso/CRISPRessoCORE.py:
  * parse_needle_output's columns            CORE:1707-1786
  * merge / identity filter / RC rescue      CORE:1830-2010
  * UNMODIFIED, N-masking, ref_positions     CORE:2014-2072
  * process_df_chunk                         CORE:428-753
  * class counts                             CORE:2866-2869
Pinned against the reference's own process_df_chunk (imported unmodified through
tests/ref_shim.py in the build container; vectors committed under tests/golden/) and, end to end,
against the golden values of the reference's tests/crispresso_tests.py:181-195.
"""
import math
import re

import numpy as np

from . import needle

VECTOR_NAMES = (
    "effect_vector_insertion", "effect_vector_deletion", "effect_vector_mutation", "effect_vector_any",
    "effect_vector_insertion_mixed", "effect_vector_deletion_mixed", "effect_vector_mutation_mixed",
    "effect_vector_insertion_hdr", "effect_vector_deletion_hdr", "effect_vector_mutation_hdr",
    "effect_vector_insertion_noncoding", "effect_vector_deletion_noncoding", "effect_vector_mutation_noncoding",
    "avg_vector_del_all", "avg_vector_ins_all")

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N", "_": "_", "-": "-"}


def reverse_complement(seq):                       # CORE:129-144
    return "".join(_COMP[c] for c in seq.upper()[::-1])


def ref_positions(ref_seq):                        # CORE:2055-2067
    out, idx = [], 0
    for c in ref_seq:
        if c in "ATCGN":
            out.append(idx)
            idx += 1
        else:
            out.append(-1 if idx == 0 else -idx)
    return out


def mask_n(ref_seq, align_str):                    # CORE:2040-2048
    masked = "".join("|" if r == "N" else c for r, c in zip(ref_seq, align_str))
    return masked, len(set(masked)) == 1


def _runs(s, ch):
    """maximal runs of ch in s as (start, end) -- what re.finditer('(-*-)') yields (CORE:474,504,518)."""
    return [m.span() for m in re.finditer("%s+" % re.escape(ch), s)]


class Opts:
    """The fields of `args` process_df_chunk reads (CORE:446,490,503,517,537,541,591,611,645)."""

    def __init__(self, coding_seq="", ignore_substitutions=False, ignore_deletions=False, ignore_insertions=False,
                 expected_hdr_amplicon_seq="", hdr_perfect_alignment_threshold=98.0,
                 hide_mutations_outside_window_NHEJ=False, window_around_sgrna=1):
        self.coding_seq = coding_seq
        self.ignore_substitutions = ignore_substitutions
        self.ignore_deletions = ignore_deletions
        self.ignore_insertions = ignore_insertions
        self.expected_hdr_amplicon_seq = expected_hdr_amplicon_seq
        self.hdr_perfect_alignment_threshold = hdr_perfect_alignment_threshold
        self.hide_mutations_outside_window_NHEJ = hide_mutations_outside_window_NHEJ
        self.window_around_sgrna = window_around_sgrna


def process_rows(rows, opts, include, L, exon=(), splice=()):
    """process_df_chunk over a list of row dicts.

    Each row: ref_seq, align_str, align_seq, score_ref, score_repaired (float or nan), UNMODIFIED.
    Returns (per_row, vectors{name: int64[L]}, hist_inframe, hist_frameshift, counters{name:int});
    per_row[i] = dict(UNMODIFIED, NHEJ, HDR, MIXED, n_mutated, n_inserted, n_deleted).
    """
    include = set(int(i) for i in include)
    exon = set(int(i) for i in exon)
    splice = set(int(i) for i in splice)
    V = {k: np.zeros(L, np.int64) for k in VECTOR_NAMES}
    hist_in, hist_fs = {}, {}
    cnt = dict(modified_frameshift=0, modified_non_frameshift=0, non_modified_non_frameshift=0, splicing_sites_modified=0)
    frameshift = bool(opts.coding_seq)
    has_hdr = bool(opts.expected_hdr_amplicon_seq)
    out = []

    def bump(vec, positions, amount=1):
        # numpy buffered fancy indexing: a repeated (wrapped) index is incremented once
        for p in set(int(x) % L for x in positions):
            vec[p] += amount

    for row in rows:
        res = dict(UNMODIFIED=bool(row["UNMODIFIED"]), NHEJ=False, HDR=False, MIXED=False,
                   n_mutated=0, n_inserted=0, n_deleted=0)
        out.append(res)
        if res["UNMODIFIED"]:
            continue                                                            # CORE:480-481
        pos = ref_positions(row["ref_seq"])
        n = len(pos)
        subs = []
        if not opts.ignore_substitutions:                                       # CORE:489-496
            for st, en in _runs(row["align_str"], "."):
                subs += pos[st:en]
        dels, del_sizes = [], []
        if not opts.ignore_deletions:                                           # CORE:498-510
            for st, en in _runs(row["align_seq"], "-"):
                dels.append(pos[st:en])
                del_sizes.append(en - st)
        del_flat = [p for d in dels for p in d]
        ins, ins_sizes = [], []
        if not opts.ignore_insertions:                                          # CORE:512-533
            for st, en in _runs(row["ref_seq"], "-"):
                ins.append([pos[max(0, st - 1)], pos[min(n - 1, en)]])
                ins_sizes.append(en - st)
        ins_flat = [p for d in ins for p in d]

        hit = bool(include & set(subs)) or bool(include & set(ins_flat)) or bool(include & set(del_flat))
        sr, sp = row["score_ref"], row.get("score_repaired", float("nan"))
        diff_neg = has_hdr and not math.isnan(sp) and (sr - sp) < 0
        if has_hdr and diff_neg and sp >= opts.hdr_perfect_alignment_threshold:    # CORE:540-543
            res["HDR"] = True
        elif has_hdr and diff_neg:                                              # CORE:546-549
            res["MIXED"] = True
        elif hit:
            res["NHEJ"] = True
        else:
            res["UNMODIFIED"] = True

        if res["MIXED"]:                                                        # CORE:579-595
            bump(V["effect_vector_mutation_mixed"], subs); bump(V["effect_vector_deletion_mixed"], del_flat)
            bump(V["effect_vector_insertion_mixed"], ins_flat)
        elif res["HDR"]:
            bump(V["effect_vector_mutation_hdr"], subs); bump(V["effect_vector_deletion_hdr"], del_flat)
            bump(V["effect_vector_insertion_hdr"], ins_flat)
        elif res["NHEJ"] and not opts.hide_mutations_outside_window_NHEJ:
            bump(V["effect_vector_mutation"], subs); bump(V["effect_vector_deletion"], del_flat)
            bump(V["effect_vector_insertion"], ins_flat)
        bump(V["effect_vector_any"], del_flat + ins_flat + subs)               # CORE:597-606

        if res["NHEJ"] and opts.window_around_sgrna:                            # CORE:611-641
            subs = [p for p in set(subs) if p in include]
            kept = [(d, s) for d, s in zip(ins, ins_sizes) if include & set(d)]
            ins, ins_sizes = [k[0] for k in kept], [k[1] for k in kept]
            kept = [(d, s) for d, s in zip(dels, del_sizes) if include & set(d)]
            dels, del_sizes = [k[0] for k in kept], [k[1] for k in kept]
            if dels:                                                            # stale otherwise (Q3)
                del_flat = [p for d in dels for p in d]
        if res["NHEJ"] and opts.hide_mutations_outside_window_NHEJ:             # CORE:643-649
            bump(V["effect_vector_mutation"], subs); bump(V["effect_vector_deletion"], del_flat)
            bump(V["effect_vector_insertion"], ins_flat)

        if not res["UNMODIFIED"]:                                               # CORE:652-725
            res["n_mutated"] = len(subs)
            res["n_inserted"] = int(sum(ins_sizes))
            res["n_deleted"] = int(sum(del_sizes))
            exon_lengths, exons_modified, spliced = [], False, False
            for d, s in zip(ins, ins_sizes):
                bump(V["avg_vector_ins_all"], d, s)
                if frameshift and (exon & set(d)):
                    exon_lengths.append(s)
                    exons_modified = True
            for d, s in zip(dels, del_sizes):
                bump(V["avg_vector_del_all"], d, s)
            if frameshift:
                hit_del = exon & set(del_flat)
                if hit_del:
                    exons_modified = True
                    exon_lengths.append(-len(hit_del))
                if exon & set(subs):
                    exons_modified = True
                if (splice & set(subs)) or (splice & set(del_flat)) or (splice & set(ins_flat)):
                    spliced = True
                if spliced:
                    cnt["splicing_sites_modified"] += 1
                if exons_modified:
                    if not exon_lengths:
                        cnt["modified_non_frameshift"] += 1
                        hist_in[0] = hist_in.get(0, 0) + 1
                    else:
                        eff = sum(exon_lengths)
                        if eff % 3 == 0:
                            cnt["modified_non_frameshift"] += 1
                            hist_in[eff] = hist_in.get(eff, 0) + 1
                        else:
                            cnt["modified_frameshift"] += 1
                            hist_fs[eff] = hist_fs.get(eff, 0) + 1
                else:
                    cnt["non_modified_non_frameshift"] += 1
                    bump(V["effect_vector_insertion_noncoding"], ins_flat)
                    bump(V["effect_vector_deletion_noncoding"], del_flat)
                    bump(V["effect_vector_mutation_noncoding"], subs)
    return out, V, hist_in, hist_fs, cnt


def hot_path(amplicon, reads, names=None, gapopen=10.0, gapextend=0.5, min_identity_score=60.0, hdr_amplicon="",
             opts=None, include=None, exon=(), splice=(), nthreads=8, use_int=True, process=None, timings=None):
    """CORE:1791-2072 + 2773-2869 on the CPU: needle (oracle) -> parse -> merge/filter -> RC rescue ->
    prep -> process_df_chunk.  reads: list of str or (buffer, offsets).  Returns a dict with the
    row list (forward rows in read order, then _RC rows), per-row results and the reductions.
    process: the quantification stage, process_rows (this file's restatement) or ref_quantify.process_rows_reference (the
    reference's own process_df_chunk); timings: optional dict that receives the seconds of the stages."""
    import time as _time
    _t0 = _time.time()
    amplicon = amplicon.upper()
    L = len(amplicon)
    opts = opts or Opts(expected_hdr_amplicon_seq=hdr_amplicon)
    if include is None:
        include = range(L)
    packed = reads if isinstance(reads, tuple) else needle.pack_reads(reads)
    buf, off = packed
    n = len(off) - 1
    names = list(names) if names is not None else ["r%d" % i for i in range(n)]
    res, ref, mark, qry = needle.align_batch(amplicon, packed, gapopen, gapextend, use_int=use_int, nthreads=nthreads)
    score_ref = res["tenths"] / 10.0
    has_hdr = bool(hdr_amplicon)
    if has_hdr:
        res_h, _, _, _ = needle.align_batch(hdr_amplicon.upper(), packed, gapopen, gapextend, use_int=use_int, nthreads=nthreads)
        score_rep = res_h["tenths"] / 10.0
    rows = []
    failed = [i for i in range(n) if score_ref[i] < min_identity_score]                  # CORE:1843-1846, 1865-1867
    for i in range(n):
        keep = score_ref[i] > min_identity_score or (has_hdr and score_rep[i] > min_identity_score)
        if keep:
            rows.append(dict(ID=names[i], read=i, rc=False, score_ref=float(score_ref[i]),
                             score_repaired=float(score_rep[i]) if has_hdr else float("nan"),
                             length=str(int(off[i + 1] - off[i])), ref_seq=ref[i], align_str=mark[i], align_seq=qry[i]))
    cells = int(L * int(off[-1]) * (2 if has_hdr else 1)) if n else 0
    if failed:                                                                          # CORE:1873-2000
        # needle reads its input as an ungapped sequence type: the '-' left in align_seq (CORE:1846)
        # are dropped, i.e. the original read is re-aligned (DESIGN.md "RC rescue input")
        sub = [bytes(buf[off[i]:off[i + 1]]) for i in failed]
        amp_rc = reverse_complement(amplicon)
        res_r, ref_r, mark_r, qry_r = needle.align_batch(amp_rc, sub, gapopen, gapextend, use_int=use_int, nthreads=nthreads)
        cells += sum(L * len(s) for s in sub)
        for j, i in enumerate(failed):
            s = res_r["tenths"][j] / 10.0
            if s > min_identity_score:                                                  # score_repaired is NaN (Q11)
                rows.append(dict(ID=names[i] + "_RC", read=i, rc=True, score_ref=float(s), score_repaired=float("nan"),
                                 length=str(len(sub[j])), ref_seq=reverse_complement(ref_r[j]), align_str=mark_r[j][::-1],
                                 align_seq=reverse_complement(qry_r[j])))
    has_n = "N" in amplicon
    for r in rows:                                                                      # CORE:2014-2052
        r["UNMODIFIED"] = r["score_ref"] == 100
        if has_n:
            r["align_str"], uniform = mask_n(r["ref_seq"], r["align_str"])
            if uniform:
                r["UNMODIFIED"] = True
    _t1 = _time.time()
    per_row, V, hist_in, hist_fs, cnt = (process or process_rows)(rows, opts, include, L, exon, splice)
    if timings is not None:
        timings["align_prepare_s"] = timings.get("align_prepare_s", 0.0) + (_t1 - _t0)
        timings["quantify_s"] = timings.get("quantify_s", 0.0) + (_time.time() - _t1)
    classes = dict(UNMODIFIED=sum(p["UNMODIFIED"] for p in per_row), NHEJ=sum(p["NHEJ"] for p in per_row),
                   HDR=sum(p["HDR"] for p in per_row), MIXED=sum(p["MIXED"] for p in per_row))
    return dict(rows=rows, per_row=per_row, vectors=V, hist_inframe=hist_in, hist_frameshift=hist_fs, counters=cnt,
                classes=classes, n_total=len(rows), n_cells=cells)
