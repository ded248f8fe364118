"""Host-side mirror of the reference's interface for the hot path.

* ``include_mask`` / ``exon_masks``: the geometry run_crispresso builds on the host
  (INCLUDE_IDXS CORE:2739-2762, EXON_POSITIONS / SPLICING_POSITIONS CORE:1414-1455).
* ``process_df_chunk``: drop-in for CRISPRessoCORE.process_df_chunk (CORE:428-753): same
  argument (``[df, args]``; the four module globals are passed explicitly), same 22-tuple back.
* ``run_hot_path``: CORE:1791-2072 + 2773-2869 through crgpu_align_quantify.
* ``build_dataframe``: the df_needle_alignment the rest of run_crispresso consumes.

Everything here calls libcrgpu through the C ABI; nothing computes on the CPU instead.
"""
import ctypes
from collections import namedtuple

import numpy as np

from . import _lib

HIST_ZERO = 1024            # histogram bin of key 0; keys span [-amplicon_len, +read_len]
HIST_LEN = 1024 + 2048 + 1

VECTOR_NAMES = (
    "effect_vector_insertion", "effect_vector_deletion", "effect_vector_mutation", "effect_vector_any",
    "effect_vector_insertion_mixed", "effect_vector_deletion_mixed", "effect_vector_mutation_mixed",
    "effect_vector_insertion_hdr", "effect_vector_deletion_hdr", "effect_vector_mutation_hdr",
    "effect_vector_insertion_noncoding", "effect_vector_deletion_noncoding", "effect_vector_mutation_noncoding",
    "avg_vector_del_all", "avg_vector_ins_all")
COUNTER_NAMES = ("modified_frameshift", "modified_non_frameshift", "non_modified_non_frameshift",
                 "splicing_sites_modified")


# --------------------------------------------------------------------------------- geometry
def include_mask(amplicon_len, cut_points, window_around_sgrna, exclude_bp_from_left, exclude_bp_from_right):
    """INCLUDE_IDXS as a 0/1 mask (CORE:2739-2762).  Ranges clipped at an amplicon end make the
    reference's np.ravel ragged; that case raises here as well."""
    L = amplicon_len
    if cut_points and window_around_sgrna > 0:
        half = max(1, window_around_sgrna // 2)
        ranges = [range(max(0, c - half + 1), min(L - 1, c + half + 1)) for c in cut_points]
        if len({len(r) for r in ranges}) > 1:
            raise ValueError("INCLUDE_IDXS ranges have unequal lengths (np.ravel of a ragged list, CORE:2759)")
        idx = np.ravel([list(r) for r in ranges]).astype(np.int64)
    else:
        idx = np.arange(L)
    excl = []
    if exclude_bp_from_left:
        excl += list(range(exclude_bp_from_left))
    if exclude_bp_from_right:
        excl += list(range(L)[-exclude_bp_from_right:])
    keep = np.setdiff1d(idx, np.array(excl, dtype=np.int64))
    m = np.zeros(L, np.uint8)
    m[keep] = 1
    return m


def cut_points_from_guides(amplicon, guide_seq, cleavage_offset=-3):
    """cut_points (CORE:1290-1321)."""
    import re

    from .synth import revcomp
    cuts = []
    for g in guide_seq.strip().upper().split(","):
        if not g:
            continue
        off_fw = cleavage_offset + len(g) - 1
        off_rc = (-cleavage_offset) - 1
        cuts += [m.start() + off_fw for m in re.finditer(g, amplicon)]
        cuts += [m.start() + off_rc for m in re.finditer(revcomp(g), amplicon)]
    return cuts


def exon_masks(amplicon, coding_seq):
    """EXON_POSITIONS and SPLICING_POSITIONS as masks (CORE:1414-1455)."""
    L = len(amplicon)
    exon = np.zeros(L, np.uint8)
    splicing = []
    for exon_seq in coding_seq.strip().upper().split(","):
        st = amplicon.find(exon_seq)
        if st < 0:
            raise ValueError("coding subsequence not contained in the amplicon")
        en = st + len(exon_seq)
        exon[st:en] = 1
        splicing += [max(0, st - 2), max(0, st - 1), min(L - 1, en), min(L - 1, en + 1)]
    splice = np.zeros(L, np.uint8)
    for p in set(splicing):
        if not exon[p]:
            splice[p] = 1
    return exon, splice


def quant_flags(expected_hdr_amplicon_seq="", ignore_substitutions=False, ignore_insertions=False,
                ignore_deletions=False, window_around_sgrna=1, hide_mutations_outside_window_NHEJ=False,
                coding_seq="", mask_n=False):
    f = 0
    if expected_hdr_amplicon_seq:
        f |= _lib.Q_HAS_HDR
    if ignore_substitutions:
        f |= _lib.Q_IGNORE_SUBS
    if ignore_insertions:
        f |= _lib.Q_IGNORE_INS
    if ignore_deletions:
        f |= _lib.Q_IGNORE_DEL
    if window_around_sgrna:
        f |= _lib.Q_WINDOW
    if hide_mutations_outside_window_NHEJ:
        f |= _lib.Q_HIDE_OUTSIDE
    if coding_seq:
        f |= _lib.Q_FRAMESHIFT
    if mask_n:
        f |= _lib.Q_MASK_N
    return f


def flags_from_args(args, mask_n=False):
    return quant_flags(getattr(args, "expected_hdr_amplicon_seq", ""), getattr(args, "ignore_substitutions", False),
                       getattr(args, "ignore_insertions", False), getattr(args, "ignore_deletions", False),
                       getattr(args, "window_around_sgrna", 0), getattr(args, "hide_mutations_outside_window_NHEJ", False),
                       getattr(args, "coding_seq", ""), mask_n)


class Reductions:
    """Host accumulators of one amplicon (int64; the reference stores the same integers in float64)."""

    def __init__(self, amplicon_len):
        self.L = amplicon_len
        self.vectors = np.zeros((_lib.NUM_VECTORS, amplicon_len), np.int64)
        self.hist_inframe = np.zeros(HIST_LEN, np.int64)
        self.hist_frameshift = np.zeros(HIST_LEN, np.int64)
        self.counters = np.zeros(_lib.NUM_COUNTERS, np.int64)
        self.class_counts = np.zeros(4, np.int64)      # UNMODIFIED, NHEJ, HDR, MIXED
        self.n_total = 0
        self.n_cells = 0             # La x Lb summed over every alignment made (the GCUPS numerator)
        self.n_cells_computed = 0    # cells actually evaluated (less when the HDR pass shares a DP prefix)

    def flat(self):
        """One int64 vector (for the multi-GPU all-reduce)."""
        return np.concatenate([self.vectors.ravel(), self.hist_inframe, self.hist_frameshift, self.counters,
                               self.class_counts, np.array([self.n_total, self.n_cells, self.n_cells_computed], np.int64)])

    def results(self):
        """flat() without n_cells_computed: the cells the kernels evaluated are a work counter that depends on how the
        reads were batched (band decision, diagonal shortcut), not a result."""
        return self.flat()[:-1]

    def load_flat(self, v):
        v = np.asarray(v, dtype=np.int64)
        o = 0
        for name, shape in (("vectors", self.vectors.shape), ("hist_inframe", (HIST_LEN,)),
                            ("hist_frameshift", (HIST_LEN,)), ("counters", (_lib.NUM_COUNTERS,)), ("class_counts", (4,))):
            k = int(np.prod(shape))
            setattr(self, name, v[o:o + k].reshape(shape).copy())
            o += k
        self.n_total, self.n_cells, self.n_cells_computed = int(v[o]), int(v[o + 1]), int(v[o + 2])

    def vector(self, name):
        return self.vectors[VECTOR_NAMES.index(name)]

    @staticmethod
    def hist_dict(h):
        nz = np.nonzero(h)[0]
        return {int(k - HIST_ZERO): int(h[k]) for k in nz}

    def counter(self, name):
        return int(self.counters[COUNTER_NAMES.index(name)])


def _quant_params(amplicon_len, flags, hdr_thr, inc, exon, splice, keepalive):
    qp = _lib.QuantParams()
    qp.amplicon_len = amplicon_len
    qp.flags = flags
    qp.hdr_perfect_alignment_threshold = float(hdr_thr)
    inc = np.ascontiguousarray(inc, dtype=np.uint8)
    keepalive.append(inc)
    qp.include_mask = inc.ctypes.data
    if exon is not None:
        exon = np.ascontiguousarray(exon, dtype=np.uint8)
        splice = np.ascontiguousarray(splice, dtype=np.uint8)
        keepalive += [exon, splice]
        qp.exon_mask = exon.ctypes.data
        qp.splice_mask = splice.ctypes.data
    return qp


def rows_to_buffer(rows, slot=None):
    """list of equal-role strings -> (uint8 [n, slot] left-aligned, lengths)."""
    lens = np.array([len(r) for r in rows], dtype=np.int32)
    slot = int(slot or max(1, lens.max() if len(lens) else 1))
    buf = np.zeros((len(rows), slot), np.uint8)
    for i, r in enumerate(rows):
        buf[i, :len(r)] = np.frombuffer(r.encode() if isinstance(r, str) else r, dtype=np.uint8)
    return buf, lens


def _score_to_tenths(col):
    a = np.asarray(col, dtype=np.float64)
    out = np.full(len(a), -1, np.int32)
    ok = ~np.isnan(a)
    out[ok] = np.rint(a[ok] * 10.0).astype(np.int32)
    return out


def quantify_rows(ctx, ref_rows, mark_rows, read_rows, score_ref, score_repaired, unmodified, amplicon_len, flags,
                  hdr_thr, inc, exon=None, splice=None, red=None):
    """crgpu_quantify on host strings.  Returns (READ_REC array, Reductions)."""
    n = len(ref_rows)
    red = red or Reductions(amplicon_len)
    recs = np.zeros(n, dtype=_lib.READ_REC)
    if n == 0:
        return recs, red
    rb, lens = rows_to_buffer(ref_rows)
    slot = rb.shape[1]
    mb, _ = rows_to_buffer(mark_rows, slot)
    qb, _ = rows_to_buffer(read_rows, slot)
    tref = _score_to_tenths(score_ref)
    trep = _score_to_tenths(score_repaired) if score_repaired is not None else None
    unmod = np.ascontiguousarray(np.asarray(unmodified, dtype=bool).astype(np.uint8))
    keep = []
    qp = _quant_params(amplicon_len, flags, hdr_thr, inc, exon, splice, keep)
    ctx.check(ctx.lib.crgpu_quantify(
        ctx.handle, _lib.MEM_HOST, ctypes.byref(qp), _lib.ptr(rb), _lib.ptr(mb), _lib.ptr(qb), slot, None,
        _lib.ptr(lens), _lib.ptr(tref), _lib.ptr(trep), _lib.ptr(unmod), n, _lib.ptr(recs), _lib.ptr(red.vectors),
        _lib.ptr(red.hist_inframe), _lib.ptr(red.hist_frameshift), HIST_LEN, HIST_ZERO, _lib.ptr(red.counters)))
    return recs, red


def process_df_chunk(chunk_input, ctx, INCLUDE_IDXS, LEN_AMPLICON, EXON_POSITIONS=None, SPLICING_POSITIONS=None):
    """Drop-in for CRISPRessoCORE.process_df_chunk (CORE:428-753).

    ``chunk_input`` is ``[df_needle_alignment_chunk, args]`` exactly as the reference passes it;
    the module globals it reads (CORE:1261-1264) are explicit parameters.  Returns the same
    22-tuple (CORE:730-753) with the same types: the updated DataFrame, 13 float64 effect
    vectors, two dict histograms, two float64 average-size sums, four ints.
    """
    df, args = chunk_input[0], chunk_input[1]
    L = LEN_AMPLICON
    inc = np.zeros(L, np.uint8)
    inc[[int(i) for i in INCLUDE_IDXS]] = 1
    exon = splice = None
    if getattr(args, "coding_seq", ""):
        exon = np.zeros(L, np.uint8)
        exon[[int(i) for i in EXON_POSITIONS]] = 1
        splice = np.zeros(L, np.uint8)
        splice[[int(i) for i in SPLICING_POSITIONS]] = 1
    has_hdr = bool(getattr(args, "expected_hdr_amplicon_seq", ""))
    recs, red = quantify_rows(
        ctx, list(df["ref_seq"]), list(df["align_str"]), list(df["align_seq"]), df["score_ref"].values,
        df["score_repaired"].values if has_hdr else None, df["UNMODIFIED"].values, L, flags_from_args(args),
        getattr(args, "hdr_perfect_alignment_threshold", 98.0), inc, exon, splice)
    df = df.copy()
    cls = recs["cls"]
    df["UNMODIFIED"] = (cls & _lib.C_UNMODIFIED) != 0
    df["NHEJ"] = (cls & _lib.C_NHEJ) != 0
    df["HDR"] = (cls & _lib.C_HDR) != 0
    df["MIXED"] = (cls & _lib.C_MIXED) != 0
    df["n_mutated"] = recs["n_mutated"].astype(np.int64)
    df["n_inserted"] = recs["n_inserted"].astype(np.int64)
    df["n_deleted"] = recs["n_deleted"].astype(np.int64)
    v = [red.vectors[i].astype(np.float64) for i in range(_lib.NUM_VECTORS)]
    return (df, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12],
            Reductions.hist_dict(red.hist_inframe), Reductions.hist_dict(red.hist_frameshift), v[13], v[14],
            red.counter("modified_frameshift"), red.counter("modified_non_frameshift"),
            red.counter("non_modified_non_frameshift"), red.counter("splicing_sites_modified"))


class HotPathResult(namedtuple("HotPathResult", [
        "kept", "aln", "tenths_rep", "recs", "rows", "slot", "rc_read", "rc_aln", "rc_recs", "rc_rows", "red",
        "allele_row", "allele_count", "allele_n"], defaults=(None, None, 0))):
    __slots__ = ()

    @property
    def bad_base(self):
        """1 where a read had a base outside ACGTN(U): such a read is not aligned and takes part in no count (needle would
        score an IUPAC ambiguity code with EDNAFULL's ambiguity rows; kept bit 3, include/crgpu.h)."""
        return (self.kept >> 3) & 1


def run_hot_path(ctx, amplicon, reads, gapopen=10.0, gapextend=0.5, min_identity_score=60.0, hdr_amplicon=None,
                 flags=None, hdr_thr=98.0, inc=None, exon=None, splice=None, want_rows=False, rc_rescue=True,
                 red=None, device_inputs=None, alleles=0):
    """CORE:1791-2072 + 2773-2869 for one amplicon through crgpu_align_quantify.

    reads: (uint8 buffer, int64 offsets) host arrays, or with ``device_inputs=(ptr_reads,
    ptr_offsets, n, max_len)`` raw device pointers (HBM-resident inputs; per-read outputs are then
    NOT returned -- only the reductions -- unless torch tensors are supplied by the caller).
    Returns HotPathResult; ``red`` carries the vectors / histograms / counters / class counts.
    """
    amp = amplicon.upper().encode()
    L = len(amp)
    red = red or Reductions(L)
    if flags is None:
        flags = quant_flags(expected_hdr_amplicon_seq=hdr_amplicon or "")
    if inc is None:
        inc = np.ones(L, np.uint8)
    keep = []
    qp = _quant_params(L, flags, hdr_thr, inc, exon, splice, keep)
    pp = _lib.PathParams()
    pp.gapopen, pp.gapextend, pp.min_identity_score = float(gapopen), float(gapextend), float(min_identity_score)
    hdr_b = hdr_amplicon.upper().encode() if hdr_amplicon else None
    pp.hdr_amplicon = hdr_b
    pp.hdr_amplicon_len = len(hdr_b) if hdr_b else 0
    pp.rc_rescue = 1 if rc_rescue else 0
    po = _lib.PathOut()
    po.vectors = red.vectors.ctypes.data
    po.hist_inframe = red.hist_inframe.ctypes.data
    po.hist_frameshift = red.hist_frameshift.ctypes.data
    po.hist_len, po.hist_zero = HIST_LEN, HIST_ZERO
    po.counters = red.counters.ctypes.data
    allele_row = allele_count = None
    if alleles:
        allele_row = np.zeros(int(alleles), np.int32)
        allele_count = np.zeros(int(alleles), np.int64)
        po.allele_cap, po.allele_row, po.allele_count = int(alleles), allele_row.ctypes.data, allele_count.ctypes.data

    if device_inputs is not None:
        d_reads, d_off, n, _maxlen, dev_out = device_inputs
        po.kept, po.aln, po.recs = dev_out["kept"], dev_out["aln"], dev_out["recs"]
        po.tenths_rep = dev_out.get("tenths_rep")
        po.slot = 0
        po.rc_cap = 0
        ctx.check(ctx.lib.crgpu_align_quantify(ctx.handle, _lib.MEM_DEVICE, amp, L, ctypes.byref(pp), ctypes.byref(qp),
                                               d_reads, d_off, n, ctypes.byref(po)))
        na = min(int(po.allele_n), int(alleles)) if alleles else 0
        res = HotPathResult(None, None, None, None, None, 0, None, None, None, None, red,
                            allele_row[:na] if alleles else None, allele_count[:na] if alleles else None, int(po.allele_n))
    else:
        buf, offsets = reads
        n = len(offsets) - 1
        kept = np.zeros(n, np.uint8)
        aln = np.zeros(n, _lib.ALN_REC)
        trep = np.full(n, -1, np.int32)
        recs = np.zeros(n, _lib.READ_REC)
        maxlb = int(np.max(np.diff(offsets))) if n else 0
        slot = max(L, len(hdr_b) if hdr_b else 0) + maxlb
        rows = rc_rows = None
        po.kept, po.aln, po.tenths_rep, po.recs = kept.ctypes.data, aln.ctypes.data, trep.ctypes.data, recs.ctypes.data
        po.slot = slot if want_rows else 0
        if want_rows:
            rows = [np.zeros((n, slot), np.uint8) for _ in range(3)]
            po.ref_rows, po.mark_rows, po.qry_rows = (r.ctypes.data for r in rows)
        rc_cap = n
        rc_read = np.zeros(rc_cap, np.int32)
        rc_aln = np.zeros(rc_cap, _lib.ALN_REC)
        rc_recs = np.zeros(rc_cap, _lib.READ_REC)
        po.rc_cap = rc_cap
        po.rc_read, po.rc_aln, po.rc_recs = rc_read.ctypes.data, rc_aln.ctypes.data, rc_recs.ctypes.data
        rc_bufs = None
        if want_rows:
            rc_bufs = [np.zeros((rc_cap, slot), np.uint8) for _ in range(3)]
            po.rc_ref_rows, po.rc_mark_rows, po.rc_qry_rows = (r.ctypes.data for r in rc_bufs)
        ctx.check(ctx.lib.crgpu_align_quantify(ctx.handle, _lib.MEM_HOST, amp, L, ctypes.byref(pp), ctypes.byref(qp),
                                               _lib.ptr(buf), _lib.ptr(offsets), n, ctypes.byref(po)))
        nrc = int(po.rc_n)
        if want_rows:
            # RC rows come back already flipped to the forward strand, left-aligned in their slot
            rc_rows = [[b[j, :rc_aln["alnlen"][j]].tobytes().decode() for j in range(nrc)] for b in rc_bufs]
        na = min(int(po.allele_n), int(alleles)) if alleles else 0
        res = HotPathResult(kept, aln, trep, recs, rows, slot, rc_read[:nrc], rc_aln[:nrc], rc_recs[:nrc], rc_rows, red,
                            allele_row[:na] if alleles else None, allele_count[:na] if alleles else None, int(po.allele_n))
    red.class_counts += np.array(list(po.class_counts), np.int64)
    red.n_total += int(po.n_total)
    red.n_cells += int(po.n_cells)
    red.n_cells_computed += int(po.n_cells_computed)
    return res


def run_hot_path_pipelined(ctxs, amplicon, reads, chunk_reads=1 << 18, gapopen=10.0, gapextend=0.5, min_identity_score=60.0,
                           hdr_amplicon=None, flags=None, hdr_thr=98.0, inc=None, exon=None, splice=None, rc_rescue=True,
                           red=None, out=None):
    """run_hot_path for HOST-resident reads with the PCIe copies hidden behind the kernels.

    The read set is cut into chunks of ``chunk_reads``; chunk c runs as one crgpu_align_quantify call on context
    ``ctxs[c % len(ctxs)]``, each context driven by its own host thread (the call is synchronous and ctypes releases the
    GIL), so the H2D / D2H copies of one chunk overlap the kernels of the others.  Every read is independent and the
    reductions are sums, so the result equals one call over all reads (per-read outputs in read order, RC rows in read
    order; `aln_off`, which positions a text row inside its call's slot, is per chunk).  Text rows and the allele
    table need the single call.

    out: optional dict of preallocated (e.g. pinned) arrays ``kept`` u8[n], ``aln`` ALN_REC[n], ``tenths_rep`` i32[n],
    ``recs`` READ_REC[n], ``rc_read`` i32[n], ``rc_aln`` ALN_REC[n], ``rc_recs`` READ_REC[n], ``offsets`` i64[n + chunks]
    (staging for the chunk-relative offsets)."""
    from concurrent.futures import ThreadPoolExecutor
    buf, offsets = reads
    n = len(offsets) - 1
    amp = amplicon.upper().encode()
    L = len(amp)
    red = red or Reductions(L)
    if flags is None:
        flags = quant_flags(expected_hdr_amplicon_seq=hdr_amplicon or "")
    if inc is None:
        inc = np.ones(L, np.uint8)
    hdr_b = hdr_amplicon.upper().encode() if hdr_amplicon else None
    bounds = [(lo, min(n, lo + chunk_reads)) for lo in range(0, n, chunk_reads)]
    out = out or {}

    def arr(name, dtype, size):
        a = out.get(name)
        return a if a is not None else np.zeros(size, dtype)

    kept, aln, recs = arr("kept", np.uint8, n), arr("aln", _lib.ALN_REC, n), arr("recs", _lib.READ_REC, n)
    trep = out.get("tenths_rep")
    if trep is None:
        trep = np.full(n, -1, np.int32)
    rc_read, rc_aln, rc_recs = arr("rc_read", np.int32, n), arr("rc_aln", _lib.ALN_REC, n), arr("rc_recs", _lib.READ_REC, n)
    off_stage = arr("offsets", np.int64, n + len(bounds))
    buf_addr, isz_aln, isz_rec = _lib.ptr(buf), _lib.ALN_REC.itemsize, _lib.READ_REC.itemsize
    addr = {k: _lib.ptr(v) for k, v in (("kept", kept), ("aln", aln), ("recs", recs), ("trep", trep), ("rc_read", rc_read),
                                        ("rc_aln", rc_aln), ("rc_recs", rc_recs))}

    def worker(w):
        ctx = ctxs[w]
        my = Reductions(L)
        keep = []
        qp = _quant_params(L, flags, hdr_thr, inc, exon, splice, keep)
        pp = _lib.PathParams()
        pp.gapopen, pp.gapextend, pp.min_identity_score = float(gapopen), float(gapextend), float(min_identity_score)
        pp.hdr_amplicon, pp.hdr_amplicon_len, pp.rc_rescue = hdr_b, len(hdr_b) if hdr_b else 0, 1 if rc_rescue else 0
        nrcs = {}
        for c in range(w, len(bounds), len(ctxs)):
            lo, hi = bounds[c]
            m = hi - lo
            offs = off_stage[lo + c:lo + c + m + 1]
            np.subtract(offsets[lo:hi + 1], offsets[lo], out=offs)
            po = _lib.PathOut()
            po.vectors, po.hist_inframe, po.hist_frameshift = my.vectors.ctypes.data, my.hist_inframe.ctypes.data, my.hist_frameshift.ctypes.data
            po.hist_len, po.hist_zero, po.counters = HIST_LEN, HIST_ZERO, my.counters.ctypes.data
            po.kept, po.aln, po.recs = addr["kept"] + lo, addr["aln"] + lo * isz_aln, addr["recs"] + lo * isz_rec
            po.tenths_rep = addr["trep"] + lo * 4
            po.slot, po.rc_cap = 0, m
            po.rc_read, po.rc_aln, po.rc_recs = addr["rc_read"] + lo * 4, addr["rc_aln"] + lo * isz_aln, addr["rc_recs"] + lo * isz_rec
            ctx.check(ctx.lib.crgpu_align_quantify(ctx.handle, _lib.MEM_HOST, amp, L, ctypes.byref(pp), ctypes.byref(qp),
                                                   buf_addr + int(offsets[lo]), _lib.ptr(offs), m, ctypes.byref(po)))
            my.class_counts += np.array(list(po.class_counts), np.int64)
            my.n_total += int(po.n_total)
            my.n_cells += int(po.n_cells)
            my.n_cells_computed += int(po.n_cells_computed)
            nrcs[c] = int(po.rc_n)
        return my, nrcs

    if len(ctxs) == 1:
        results = [worker(0)]
    else:
        with ThreadPoolExecutor(max_workers=len(ctxs)) as ex:
            results = list(ex.map(worker, range(len(ctxs))))
    nrc_of = {}
    for my, nrcs in results:
        red.load_flat(red.flat() + my.flat())
        nrc_of.update(nrcs)
    # RC rows: each chunk left its own compact, read-ordered list at [lo, lo + nrc); close the gaps
    pos = 0
    for c, (lo, hi) in enumerate(bounds):
        k = nrc_of.get(c, 0)
        if k:
            rc_read[pos:pos + k] = rc_read[lo:lo + k] + lo
            if pos != lo:
                rc_aln[pos:pos + k] = rc_aln[lo:lo + k]
                rc_recs[pos:pos + k] = rc_recs[lo:lo + k]
            pos += k
    return HotPathResult(kept, aln, trep, recs, None, 0, rc_read[:pos], rc_aln[:pos], rc_recs[:pos], None, red)


_BAM4 = np.zeros(256, np.uint8)
for _i, _c in enumerate(b"=ACMGRSVTWYHKDBN"):
    _BAM4[_c] = _i
    _BAM4[ord(chr(_c).lower())] = _i


def pack_bam4(buf):
    """One base per byte -> BAM's 4-bit codes, two bases per byte, high nibble first, dense over the whole buffer (base j
    is nibble j): the CRGPU_READS_BAM4 input format of crgpu_stage_reads (include/crgpu.h).  Offsets keep counting bases."""
    codes = _BAM4[np.asarray(buf, dtype=np.uint8)]
    if len(codes) & 1:
        codes = np.concatenate([codes, np.zeros(1, np.uint8)])
    return ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8)


class StagedPipeline:
    """Batches of HOST-resident reads through ONE context with the PCIe copy of the next batch hidden behind the kernels
    of the current one: ``stage`` starts the copy of a batch into one of the context's two staging slots
    (crgpu_stage_reads: asynchronous, on the library's copy stream; keep the buffers pinned), ``run`` waits for the oldest
    staged batch and quantifies it (crgpu_align_quantify_staged).  Reductions accumulate in ``red``; with ``alleles`` the
    per-batch allele tables are merged by their 128-bit keys (``alleles_merged``).  Every read is independent and the
    reductions are sums, so the results equal one call over all reads."""

    def __init__(self, ctx, amplicon, gapopen=10.0, gapextend=0.5, min_identity_score=60.0, hdr_amplicon=None, flags=None,
                 hdr_thr=98.0, inc=None, exon=None, splice=None, rc_rescue=True, red=None, alleles=0, deferred=False):
        self.ctx = ctx
        # deferred: the per-read arrays of a batch arrive behind the next batch's kernels; valid after take_results() / ctx.sync()
        self.deferred = bool(deferred)
        ctx.check(ctx.lib.crgpu_set_deferred_outputs(ctx.handle, 1 if deferred else 0))
        self.amp = amplicon.upper().encode()
        L = len(self.amp)
        self.L = L
        self.red = red or Reductions(L)
        if flags is None:
            flags = quant_flags(expected_hdr_amplicon_seq=hdr_amplicon or "")
        if inc is None:
            inc = np.ones(L, np.uint8)
        self._keep = []
        self.qp = _quant_params(L, flags, hdr_thr, inc, exon, splice, self._keep)
        self.pp = _lib.PathParams()
        self.pp.gapopen, self.pp.gapextend, self.pp.min_identity_score = float(gapopen), float(gapextend), float(min_identity_score)
        self._hdr_b = hdr_amplicon.upper().encode() if hdr_amplicon else None
        self.pp.hdr_amplicon, self.pp.hdr_amplicon_len = self._hdr_b, len(self._hdr_b) if self._hdr_b else 0
        self.pp.rc_rescue = 1 if rc_rescue else 0
        self.alleles = int(alleles)
        if self.alleles:
            self._arow = np.zeros(self.alleles, np.int32)
            self._acnt = np.zeros(self.alleles, np.int64)
            self._akey = np.zeros(2 * self.alleles, np.uint64)
        self._tables = []                    # per batch: (keys[na, 2], counts[na], rows[na], batch index[na])
        # the two staging slots belong to the CONTEXT: pipelines of several amplicons can share one context (pooled runs),
        # each staging its next batch while another one's runs -- first staged, first run
        if not hasattr(ctx, "_staged"):
            ctx._staged, ctx._stage_next = [], 0       # staged, not yet run: (slot, n, owner, buffers kept alive)
        self.batches = 0

    def stage(self, reads, offsets, packed=False):
        """Start copying a batch: ``reads`` u8 (one base per byte, or pack_bam4 output with packed=True), ``offsets`` i64[n+1]
        starting at 0, counting bases.  At most two batches can be staged and not yet run."""
        ctx = self.ctx
        if len(ctx._staged) >= 2:
            raise RuntimeError("both staging slots are in use: run() a batch first")
        n = len(offsets) - 1
        slot = ctx._stage_next
        ctx.check(ctx.lib.crgpu_stage_reads(ctx.handle, slot, _lib.READS_BAM4 if packed else _lib.READS_BYTES,
                                            _lib.ptr(reads), _lib.ptr(offsets), n))
        ctx._staged.append((slot, n, self, (reads, offsets)))
        ctx._stage_next ^= 1
        return n

    def run(self, out):
        """Quantify the oldest staged batch of n reads.  ``out``: arrays (views) ``kept`` u8[n], ``aln`` ALN_REC[n], ``recs``
        READ_REC[n], optional ``tenths_rep`` i32[n], ``rc_read`` i32[n], ``rc_aln`` ALN_REC[n], ``rc_recs`` READ_REC[n].
        Returns (n, number of RC rows)."""
        if not self.ctx._staged or self.ctx._staged[0][2] is not self:
            raise RuntimeError("the oldest staged batch of this context belongs to another pipeline (or nothing is staged)")
        slot, n, _owner, _alive = self.ctx._staged.pop(0)
        self.ctx.check(self.ctx.lib.crgpu_set_deferred_outputs(self.ctx.handle, 1 if self.deferred else 0))
        red = self.red
        po = _lib.PathOut()
        po.vectors, po.hist_inframe, po.hist_frameshift = red.vectors.ctypes.data, red.hist_inframe.ctypes.data, red.hist_frameshift.ctypes.data
        po.hist_len, po.hist_zero, po.counters = HIST_LEN, HIST_ZERO, red.counters.ctypes.data
        po.kept, po.aln, po.recs = _lib.ptr(out["kept"]), _lib.ptr(out["aln"]), _lib.ptr(out["recs"])
        if out.get("tenths_rep") is not None:
            po.tenths_rep = _lib.ptr(out["tenths_rep"])
        po.slot = 0
        if out.get("rc_read") is not None:
            po.rc_cap = n
            po.rc_read, po.rc_aln, po.rc_recs = _lib.ptr(out["rc_read"]), _lib.ptr(out["rc_aln"]), _lib.ptr(out["rc_recs"])
        if self.alleles:
            po.allele_cap, po.allele_row, po.allele_count = self.alleles, self._arow.ctypes.data, self._acnt.ctypes.data
            po.allele_key = self._akey.ctypes.data
        self.ctx.check(self.ctx.lib.crgpu_align_quantify_staged(self.ctx.handle, slot, self.amp, self.L, ctypes.byref(self.pp),
                                                                ctypes.byref(self.qp), ctypes.byref(po)))
        red.class_counts += np.array(list(po.class_counts), np.int64)
        red.n_total += int(po.n_total)
        red.n_cells += int(po.n_cells)
        red.n_cells_computed += int(po.n_cells_computed)
        if self.alleles:
            na = int(po.allele_n)
            if na > self.alleles:
                raise ValueError("a batch holds %d distinct alleles, more than alleles=%d: its table would be truncated" % (na, self.alleles))
            self._tables.append((self._akey[:2 * na].reshape(na, 2).copy(), self._acnt[:na].copy(), self._arow[:na].copy(),
                                 np.full(na, self.batches, np.int32)))
        self.batches += 1
        return n, int(po.rc_n)

    def take_results(self, sync=True):
        """(reductions, merged allele table) of the batches run since the last call; both start afresh afterwards (batches
        already staged stay staged: a stream of read sets can keep the copy of the next set's first batch in flight).
        With deferred outputs the per-read arrays of the last batch are still travelling: sync=True waits for them;
        sync=False leaves them to arrive behind the next batch's kernels (valid after the next-but-one run() or ctx.sync())."""
        if self.deferred and sync:
            self.ctx.sync()
        red, table = self.red, self.alleles_merged()
        self.red = Reductions(self.L)
        self._tables = []
        self.batches = 0
        return red, table

    def alleles_merged(self):
        """(counts i64[A], batch i32[A], row i32[A]) most frequent first: the per-batch tables merged by their two 64-bit
        keys; ``row`` is the representative's row within its batch (read index i for a forward row, n_batch + j for that
        batch's RC row j)."""
        if not self._tables:
            return np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32)
        if len(self._tables) == 1:                          # one batch: the device's table is the answer (most frequent first)
            _k, cnt, row, bat = self._tables[0]
            return cnt.astype(np.int64), bat, row
        keys = np.concatenate([t[0] for t in self._tables])
        cnt = np.concatenate([t[1] for t in self._tables])
        row = np.concatenate([t[2] for t in self._tables])
        bat = np.concatenate([t[3] for t in self._tables])
        _u, first, inv = np.unique(keys[:, 0], return_index=True, return_inverse=True)
        if not np.array_equal(keys[first, 1][inv], keys[:, 1]):
            raise RuntimeError("allele merge: two alleles share their first 64-bit key")
        total = np.zeros(len(first), np.int64)
        np.add.at(total, inv, cnt)
        order = np.argsort(-total, kind="stable")
        return total[order], bat[first][order], row[first][order]


def run_hot_path_staged(ctx, amplicon, reads, chunk_reads=1 << 19, packed=None, out=None, **kw):
    """run_hot_path for HOST-resident reads through StagedPipeline: the read set is cut into chunks of ~``chunk_reads``, the
    copy of chunk c + 1 overlaps the kernels of chunk c.  ``reads`` = (buffer, offsets); ``packed`` = pack_bam4(buffer) to
    send 4 bits per base instead of 8 (chunks then start on even base offsets).  Per-read outputs in read order, RC rows
    in read order, reductions and (with ``alleles=``) the merged allele table as ``(allele_row, allele_count)`` where a row
    >= n is RC row (row - n) of the returned list.  Text rows need the single call (run_hot_path)."""
    buf, offsets = reads
    n = len(offsets) - 1
    pipe = StagedPipeline(ctx, amplicon, **kw)
    out = out or {}

    def arr(name, dtype, size):
        a = out.get(name)
        return a if a is not None else np.zeros(size, dtype)

    kept, aln, recs = arr("kept", np.uint8, n), arr("aln", _lib.ALN_REC, n), arr("recs", _lib.READ_REC, n)
    trep = out.get("tenths_rep")
    if trep is None:
        trep = np.full(n, -1, np.int32)
    rc_read, rc_aln, rc_recs = arr("rc_read", np.int32, n), arr("rc_aln", _lib.ALN_REC, n), arr("rc_recs", _lib.READ_REC, n)
    # chunk boundaries (on even base offsets when the reads travel packed)
    bounds, lo = [], 0
    while lo < n:
        hi = min(n, lo + int(chunk_reads))
        while packed is not None and hi < n and (int(offsets[hi]) & 1):
            hi += 1
        bounds.append((lo, hi))
        lo = hi
    off_stage = arr("offsets", np.int64, n + len(bounds))

    def stage(c):
        lo, hi = bounds[c]
        offs = off_stage[lo + c:lo + c + (hi - lo) + 1]
        np.subtract(offsets[lo:hi + 1], offsets[lo], out=offs)
        b0, b1 = int(offsets[lo]), int(offsets[hi])
        if packed is not None:
            pipe.stage(packed[b0 // 2:(b1 + 1) // 2], offs, packed=True)
        else:
            pipe.stage(buf[b0:b1], offs)

    nrcs = []
    if bounds:
        stage(0)
    for c, (lo, hi) in enumerate(bounds):
        if c + 1 < len(bounds):
            stage(c + 1)
        view = {"kept": kept[lo:hi], "aln": aln[lo:hi], "recs": recs[lo:hi], "tenths_rep": trep[lo:hi],
                "rc_read": rc_read[lo:hi], "rc_aln": rc_aln[lo:hi], "rc_recs": rc_recs[lo:hi]}
        _m, nrc = pipe.run(view)
        nrcs.append(nrc)
    if pipe.deferred:
        ctx.sync()                           # the per-read arrays of the last chunks are still on their way
    # RC rows: each chunk left its own compact, read-ordered list at [lo, lo + nrc); close the gaps
    pos, rc_base = 0, []
    for c, (lo, hi) in enumerate(bounds):
        k = nrcs[c]
        rc_base.append(pos)
        if k:
            rc_read[pos:pos + k] = rc_read[lo:lo + k] + lo
            if pos != lo:
                rc_aln[pos:pos + k] = rc_aln[lo:lo + k]
                rc_recs[pos:pos + k] = rc_recs[lo:lo + k]
            pos += k
    allele_row = allele_count = None
    allele_n = 0
    if pipe.alleles:
        allele_count, bat, row = pipe.alleles_merged()
        allele_n = len(allele_count)
        los = np.array([lo for lo, _hi in bounds], np.int64)[bat]
        sizes = np.array([hi - lo for lo, hi in bounds], np.int64)[bat]
        rcb = np.array(rc_base, np.int64)[bat]
        allele_row = np.where(row < sizes, los + row, n + rcb + (row - sizes))
    return HotPathResult(kept, aln, trep, recs, None, 0, rc_read[:pos], rc_aln[:pos], rc_recs[:pos], None, pipe.red,
                         allele_row, allele_count, allele_n)


def build_dataframe(res, read_names, has_hdr=False, amplicon=None):
    """df_needle_alignment as run_crispresso holds it after CORE:2072 and the quantification
    (CORE:2864): index ID, columns score_ref [score_repaired score_diff] length ref_seq align_str
    align_seq UNMODIFIED MIXED HDR NHEJ n_mutated n_inserted n_deleted, forward rows in read
    order followed by the ``_RC`` rows (CORE:1993-1998).  Needs want_rows=True."""
    import pandas as pd
    if res.rows is None:
        raise ValueError("run_hot_path(want_rows=True) is required to build the DataFrame")
    mask_n = amplicon is not None and "N" in amplicon.upper()

    def frame(ids, aln, trep, recs, ref, mark, qry, rc):
        d = {"score_ref": aln["tenths"] / 10.0}
        if has_hdr:
            d["score_repaired"] = np.where(trep >= 0, trep / 10.0, np.nan) if not rc else np.full(len(ids), np.nan)
        d["length"] = [str(int(x)) for x in aln["read_len"]]
        if mask_n:
            mark = ["".join("|" if rch == "N" else c for rch, c in zip(r_, m_)) for r_, m_ in zip(ref, mark)]
        d["ref_seq"], d["align_str"], d["align_seq"] = ref, mark, qry
        if has_hdr:
            d["score_diff"] = d["score_ref"] - d["score_repaired"]
        cls = recs["cls"]
        d["UNMODIFIED"] = (cls & _lib.C_UNMODIFIED) != 0
        d["MIXED"] = (cls & _lib.C_MIXED) != 0
        d["HDR"] = (cls & _lib.C_HDR) != 0
        d["NHEJ"] = (cls & _lib.C_NHEJ) != 0
        d["n_mutated"] = recs["n_mutated"].astype(np.int64)
        d["n_inserted"] = recs["n_inserted"].astype(np.int64)
        d["n_deleted"] = recs["n_deleted"].astype(np.int64)
        return pd.DataFrame(d, index=pd.Index(ids, name="ID"))

    fw = np.nonzero(res.kept & 1)[0]
    rows = []
    for arr in res.rows:
        rows.append([arr[i, res.aln["aln_off"][i]:res.aln["aln_off"][i] + res.aln["alnlen"][i]].tobytes().decode() for i in fw])
    names = np.asarray(read_names, dtype=object)
    df = frame(list(names[fw]), res.aln[fw], res.tenths_rep[fw], res.recs[fw], rows[0], rows[1], rows[2], False)
    if len(res.rc_read):
        sel = np.nonzero((res.kept[res.rc_read] & 2) != 0)[0]
        if len(sel):
            rr = [[res.rc_rows[k][j] for j in sel] for k in range(3)]
            ids = [names[res.rc_read[j]] + "_RC" for j in sel]
            df_rc = frame(ids, res.rc_aln[sel], None, res.rc_recs[sel], rr[0], rr[1], rr[2], True)
            df = pd.concat([df, df_rc])
    if df.shape[0] != df.index.unique().shape[0]:                      # CORE:2002-2010
        raise ValueError("The .fastq file/s contain/s duplicate sequence IDs (DuplicateSequenceIdException)")
    return df


def allele_table(ctx, res, amplicon, reads, gapopen=10.0, gapextend=0.5):
    """df_alleles (CORE:2923-2946) from the device-side grouping of run_hot_path(..., alleles=K): one
    row per allele, most frequent first, columns Aligned_Sequence, Reference_Sequence, NHEJ,
    UNMODIFIED, HDR, n_deleted, n_inserted, n_mutated, #Reads, %Reads.  Only the representatives'
    text rows are materialised (a second, small alignment call when run_hot_path did not keep rows)."""
    import pandas as pd

    from .aligner import needle_align
    from .synth import revcomp
    buf, offsets = reads
    n = len(offsets) - 1
    rows, counts = res.allele_row, res.allele_count
    fw = [(k, int(r)) for k, r in enumerate(rows) if r < n]
    rc = [(k, int(r) - n) for k, r in enumerate(rows) if r >= n]
    aligned = [None] * len(rows)
    refseq = [None] * len(rows)
    recs = [None] * len(rows)

    def sub(idx):
        lens = (offsets[1:] - offsets[:-1])[idx]
        off = np.zeros(len(idx) + 1, np.int64)
        off[1:] = np.cumsum(lens)
        out = np.empty(int(off[-1]), np.uint8)
        for j, i in enumerate(idx):
            out[off[j]:off[j + 1]] = buf[offsets[i]:offsets[i + 1]]
        return out, off

    if fw:
        idx = np.array([r for _k, r in fw])
        if res.rows is not None:
            for (k, r) in fw:
                o, ln = res.aln["aln_off"][r], res.aln["alnlen"][r]
                refseq[k] = res.rows[0][r, o:o + ln].tobytes().decode()
                aligned[k] = res.rows[2][r, o:o + ln].tobytes().decode()
        else:
            _r, a_ref, _m, a_qry = needle_align(ctx, amplicon.upper(), sub(idx), gapopen, gapextend)
            for j, (k, _r2) in enumerate(fw):
                refseq[k], aligned[k] = a_ref[j], a_qry[j]
        for (k, r) in fw:
            recs[k] = res.recs[r]
    if rc:
        reads_idx = np.array([int(res.rc_read[j]) for _k, j in rc])
        if res.rc_rows is not None:
            for (k, j) in rc:
                refseq[k], aligned[k] = res.rc_rows[0][j], res.rc_rows[2][j]
        else:
            _r, a_ref, _m, a_qry = needle_align(ctx, revcomp(amplicon.upper()), sub(reads_idx), gapopen, gapextend)
            for j2, (k, _j) in enumerate(rc):
                refseq[k], aligned[k] = revcomp(a_ref[j2].upper()), revcomp(a_qry[j2].upper())
        for (k, j) in rc:
            recs[k] = res.rc_recs[j]
    cls = np.array([r["cls"] for r in recs], np.uint8) if recs else np.zeros(0, np.uint8)
    df = pd.DataFrame({
        "Aligned_Sequence": aligned, "Reference_Sequence": refseq,
        "NHEJ": (cls & _lib.C_NHEJ) != 0, "UNMODIFIED": (cls & _lib.C_UNMODIFIED) != 0, "HDR": (cls & _lib.C_HDR) != 0,
        "n_deleted": [int(r["n_deleted"]) for r in recs], "n_inserted": [int(r["n_inserted"]) for r in recs],
        "n_mutated": [int(r["n_mutated"]) for r in recs], "#Reads": counts.astype(np.int64)})
    df["%Reads"] = df["#Reads"] / float(res.red.n_total) * 100.0
    return df
