"""Post-reduction scalars and histograms run_crispresso derives from the per-read records and the
reduced vectors (CRISPResso/CRISPRessoCORE.py:2866-2905, 2345-2365, 2960-2973, 3751-3803, 3882-3900).
They are cheap numpy reductions over GPU-produced records and stay on the host.
"""
import numpy as np

from . import _lib


def class_event_counts(cls, n_inserted, n_deleted, n_mutated):
    """nhej_/hdr_/mixed_ inserted/deleted/mutated read counts (CORE:3751-3803)."""
    out = {}
    for name, bit in (("nhej", _lib.C_NHEJ), ("hdr", _lib.C_HDR), ("mixed", _lib.C_MIXED)):
        sel = (cls & bit) != 0
        out[name + "_inserted"] = int(np.sum(n_inserted[sel] > 0))
        out[name + "_deleted"] = int(np.sum(n_deleted[sel] > 0))
        out[name + "_mutated"] = int(np.sum(n_mutated[sel] > 0))
    return out


def normalise_vectors(red):
    """effect_vector_combined and the average-size vectors (CORE:2872-2890)."""
    v = {n: red.vector(n).astype(np.float64) for n in (
        "effect_vector_any", "avg_vector_ins_all", "avg_vector_del_all", "effect_vector_insertion",
        "effect_vector_insertion_hdr", "effect_vector_insertion_mixed", "effect_vector_deletion",
        "effect_vector_deletion_hdr", "effect_vector_deletion_mixed")}
    with np.errstate(divide="ignore", invalid="ignore"):
        combined = 100.0 * v["effect_vector_any"] / float(red.n_total)
        avg_ins = v["avg_vector_ins_all"] / (v["effect_vector_insertion"] + v["effect_vector_insertion_hdr"] + v["effect_vector_insertion_mixed"])
        avg_del = v["avg_vector_del_all"] / (v["effect_vector_deletion"] + v["effect_vector_deletion_hdr"] + v["effect_vector_deletion_mixed"])
    for a in (avg_ins, avg_del):
        a[np.isnan(a)] = 0
        a[np.isinf(a)] = 0
    return combined, avg_ins, avg_del


def indel_size_histogram(n_inserted, n_deleted, amplicon_len, cut_points, has_guides):
    """hlengths, hdensity of effective_len - LEN_AMPLICON (CORE:2903-2973)."""
    L = amplicon_len
    if has_guides:
        xmin, xmax = -min(cut_points), L - max(cut_points)
    else:
        xmin, xmax = -(L // 2), +(L // 2)
    hdensity, hlengths = np.histogram(n_inserted.astype(np.int64) - n_deleted.astype(np.int64), np.arange(xmin, xmax))
    return hlengths[:-1], hdensity


def _range(values):
    nz = values[values > 0]
    try:
        return max(15, int(np.round(np.percentile(nz, 99))))
    except Exception:
        return 15


def event_size_histograms(n_inserted, n_deleted, n_mutated):
    """(ins_size, fq), (del_size, fq), (sub_size, fq) as written to *_histogram.txt (CORE:2345-2365,
    3888-3900).  np.histogram's last bin is closed, so size r-1 is folded into bin r-2."""
    out = []
    for vals, sign in ((n_inserted, 1), (n_deleted, -1), (n_mutated, 1)):
        r = _range(vals)
        y, x = np.histogram(vals, bins=range(0, r))
        out.append((sign * x[:-1], y))
    return out


def allele_table(df):
    """df_alleles (CORE:2923-2946): identical (aligned read, aligned amplicon, class, counts) rows grouped."""
    g = df.groupby(["align_seq", "ref_seq", "NHEJ", "UNMODIFIED", "HDR", "n_deleted", "n_inserted", "n_mutated"]).size()
    g = g.reset_index().rename(columns={0: "#Reads", "align_seq": "Aligned_Sequence", "ref_seq": "Reference_Sequence"})
    g["%Reads"] = g["#Reads"] / g["#Reads"].sum() * 100.0
    return g.sort_values(by="#Reads", ascending=False)
