"""CPU baseline only (bench.py: `cpu_baseline` and `--impl reference`) -- TEST / MEASUREMENT INFRASTRUCTURE, never the product.

Quantification stage of the reference path run by the reference's OWN, UNMODIFIED code: process_df_chunk
(CRISPResso/CRISPRessoCORE.py:428-753) driven as run_crispresso drives it (CORE:2764-2864: one call, or a
multiprocessing.Pool over get_chunk's row groups with -p > 1), imported through tests/ref_shim.py from /root/reference or
from its installed copy baseline/_ref.  Input rows are what oracle/quantify.hot_path prepared (CORE:1830-2072 restated).
Same return shape as oracle/quantify.process_rows, so hot_path can use either.
"""
import argparse
import multiprocessing as mp
import os
import sys

import numpy as np
import pandas as pd

from . import quantify

_TESTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")


def _shim():
    if _TESTS not in sys.path:
        sys.path.insert(0, _TESTS)
    import ref_shim
    return ref_shim


def available():
    try:
        return _shim().available()
    except Exception:
        return False


def process_rows_reference(rows, opts, include, L, exon=(), splice=(), n_processes=1):
    """-> (per_row, vectors, hist_inframe, hist_frameshift, counters), computed by the reference's process_df_chunk."""
    core = _shim().load_core()
    has_hdr = bool(opts.expected_hdr_amplicon_seq)
    df = pd.DataFrame({
        "score_ref": [r["score_ref"] for r in rows],
        "ref_seq": [r["ref_seq"] for r in rows],
        "align_str": [r["align_str"] for r in rows],
        "align_seq": [r["align_seq"] for r in rows],
    }, index=[r["ID"] for r in rows])
    if has_hdr:
        df["score_repaired"] = [r["score_repaired"] for r in rows]
        df["score_diff"] = df.score_ref - df.score_repaired
    df["UNMODIFIED"] = [bool(r["UNMODIFIED"]) for r in rows]
    df["MIXED"] = False
    df["HDR"] = False
    df["NHEJ"] = False
    df["n_mutated"] = 0
    df["n_inserted"] = 0
    df["n_deleted"] = 0
    df["ref_positions"] = df["ref_seq"].apply(lambda s: np.array(quantify.ref_positions(s)))          # CORE:2055-2072
    core.INCLUDE_IDXS = set(np.array(list(include), dtype=np.int64))
    core.LEN_AMPLICON = L
    core.EXON_POSITIONS = sorted(exon)
    core.SPLICING_POSITIONS = set(splice)
    args = argparse.Namespace(coding_seq=opts.coding_seq, ignore_substitutions=opts.ignore_substitutions,
                              ignore_deletions=opts.ignore_deletions, ignore_insertions=opts.ignore_insertions,
                              expected_hdr_amplicon_seq=opts.expected_hdr_amplicon_seq,
                              hdr_perfect_alignment_threshold=opts.hdr_perfect_alignment_threshold,
                              hide_mutations_outside_window_NHEJ=opts.hide_mutations_outside_window_NHEJ,
                              window_around_sgrna=opts.window_around_sgrna)
    if n_processes > 1 and len(df) > n_processes:
        processes = min(df.shape[0], n_processes)                     # CORE:2775-2784

        def get_chunk():
            for _, part in df.groupby(np.arange(len(df)) // (len(df) // (processes - 1))):
                yield part, args

        pool = mp.Pool(processes=processes)
        outs = list(pool.imap(core.process_df_chunk, get_chunk()))
        pool.close()
        pool.join()
    else:
        outs = [core.process_df_chunk([df, args])]
    names = quantify.VECTOR_NAMES
    V = {k: np.zeros(L, dtype=np.int64) for k in names}
    hist_in, hist_fs = {}, {}
    cnt = dict(modified_frameshift=0, modified_non_frameshift=0, non_modified_non_frameshift=0, splicing_sites_modified=0)
    per_row = []
    for out in outs:
        d = out[0]
        for k, idx in zip(names[:13], range(1, 14)):
            V[k] += np.asarray(out[idx]).astype(np.int64)
        V["avg_vector_del_all"] += np.asarray(out[16]).astype(np.int64)
        V["avg_vector_ins_all"] += np.asarray(out[17]).astype(np.int64)
        for src, dst in ((out[14], hist_in), (out[15], hist_fs)):
            for k, v in src.items():
                dst[int(k)] = dst.get(int(k), 0) + int(v)
        for k, idx in zip(("modified_frameshift", "modified_non_frameshift", "non_modified_non_frameshift", "splicing_sites_modified"),
                          range(18, 22)):
            cnt[k] += int(out[idx])
        u, nh, hd, mx = d["UNMODIFIED"].values, d["NHEJ"].values, d["HDR"].values, d["MIXED"].values
        nm, ni, nd = d["n_mutated"].values, d["n_inserted"].values, d["n_deleted"].values
        per_row += [dict(UNMODIFIED=bool(u[i]), NHEJ=bool(nh[i]), HDR=bool(hd[i]), MIXED=bool(mx[i]), n_mutated=int(nm[i]),
                         n_inserted=int(ni[i]), n_deleted=int(nd[i])) for i in range(len(d))]
    return per_row, V, hist_in, hist_fs, cnt
