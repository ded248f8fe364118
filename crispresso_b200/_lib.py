"""ctypes binding of libcrgpu.so (include/crgpu.h).  There is no CPU fallback: if the library is
missing or no B200 is visible, every entry point raises."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CRGPU_LIB") or os.path.join(HERE, "libcrgpu.so")     # CRGPU_LIB: experiment builds (scripts/)

MEM_HOST, MEM_DEVICE = 0, 1
READS_BYTES, READS_BAM4 = 0, 1
E_CUDA, E_ARG, E_ALIGN, E_NOMEM = 1, 2, 3, 4

ALN_REC = np.dtype([("score", "<f4"), ("alnlen", "<i4"), ("ident", "<i4"), ("tenths", "<i4"),
                    ("aln_off", "<i4"), ("start1", "<i4"), ("start2", "<i4"), ("read_len", "<i4")])
READ_REC = np.dtype([("cls", "u1"), ("pad", "u1", (3,)), ("n_mutated", "<i4"), ("n_inserted", "<i4"),
                     ("n_deleted", "<i4")])
NUM_VECTORS = 15
NUM_COUNTERS = 4
Q_HAS_HDR, Q_IGNORE_SUBS, Q_IGNORE_INS, Q_IGNORE_DEL, Q_WINDOW, Q_HIDE_OUTSIDE, Q_FRAMESHIFT, Q_MASK_N = (
    1, 2, 4, 8, 16, 32, 64, 128)
C_UNMODIFIED, C_NHEJ, C_HDR, C_MIXED = 1, 2, 4, 8


class QuantParams(ctypes.Structure):
    _fields_ = [("amplicon_len", ctypes.c_int32), ("flags", ctypes.c_int32),
                ("hdr_perfect_alignment_threshold", ctypes.c_double),
                ("include_mask", ctypes.c_void_p), ("exon_mask", ctypes.c_void_p), ("splice_mask", ctypes.c_void_p)]


class PathParams(ctypes.Structure):
    _fields_ = [("gapopen", ctypes.c_double), ("gapextend", ctypes.c_double), ("min_identity_score", ctypes.c_double),
                ("hdr_amplicon", ctypes.c_char_p), ("hdr_amplicon_len", ctypes.c_int32), ("rc_rescue", ctypes.c_int32)]


class PathOut(ctypes.Structure):
    _fields_ = [("kept", ctypes.c_void_p), ("aln", ctypes.c_void_p), ("tenths_rep", ctypes.c_void_p),
                ("recs", ctypes.c_void_p), ("ref_rows", ctypes.c_void_p), ("mark_rows", ctypes.c_void_p),
                ("qry_rows", ctypes.c_void_p), ("slot", ctypes.c_int64),
                ("rc_cap", ctypes.c_int64), ("rc_n", ctypes.c_int64), ("rc_read", ctypes.c_void_p),
                ("rc_aln", ctypes.c_void_p), ("rc_recs", ctypes.c_void_p), ("rc_ref_rows", ctypes.c_void_p),
                ("rc_mark_rows", ctypes.c_void_p), ("rc_qry_rows", ctypes.c_void_p),
                ("vectors", ctypes.c_void_p), ("hist_inframe", ctypes.c_void_p), ("hist_frameshift", ctypes.c_void_p),
                ("hist_len", ctypes.c_int32), ("hist_zero", ctypes.c_int32), ("counters", ctypes.c_void_p),
                ("class_counts", ctypes.c_int64 * 4), ("n_total", ctypes.c_int64), ("n_cells", ctypes.c_int64),
                ("allele_cap", ctypes.c_int64), ("allele_n", ctypes.c_int64), ("allele_row", ctypes.c_void_p),
                ("allele_count", ctypes.c_void_p), ("n_cells_computed", ctypes.c_int64), ("allele_key", ctypes.c_void_p)]


class MergeParams(ctypes.Structure):
    _fields_ = [("min_overlap", ctypes.c_int32), ("max_overlap", ctypes.c_int32),
                ("max_mismatch_density", ctypes.c_float), ("allow_outies", ctypes.c_int32)]


class MergeOut(ctypes.Structure):
    _fields_ = [("pos", ctypes.c_void_p), ("kind", ctypes.c_void_p), ("cap_bytes", ctypes.c_int64),
                ("cap_reads", ctypes.c_int64), ("seq", ctypes.c_void_p), ("qual", ctypes.c_void_p),
                ("offsets", ctypes.c_void_p), ("index", ctypes.c_void_p), ("n_merged", ctypes.c_int64),
                ("n_innie", ctypes.c_int64), ("n_outie", ctypes.c_int64), ("bytes", ctypes.c_int64)]


class FastqOut(ctypes.Structure):
    _fields_ = [("n_records", ctypes.c_int64), ("consumed", ctypes.c_int64), ("seq_bytes", ctypes.c_int64),
                ("cap_records", ctypes.c_int64), ("cap_bytes", ctypes.c_int64), ("seq", ctypes.c_void_p),
                ("qual", ctypes.c_void_p), ("offsets", ctypes.c_void_p), ("name_start", ctypes.c_void_p),
                ("name_len", ctypes.c_void_p)]


class CrgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libcrgpu error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load():
    """Load libcrgpu.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CRGPU_LIB", LIB_PATH)         # (a differently built libcrgpu for A/B measurements)
    if not os.path.exists(path):
        raise ImportError("crispresso_b200: %s is missing -- run `python -m crispresso_b200.build` "
                          "(there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
    lib.crgpu_abi_version.restype = i32
    lib.crgpu_create.argtypes = [ctypes.POINTER(vp), i32]
    lib.crgpu_destroy.argtypes = [vp]
    lib.crgpu_destroy.restype = None
    lib.crgpu_last_error.argtypes = [vp]
    lib.crgpu_last_error.restype = ctypes.c_char_p
    lib.crgpu_set_traceback_budget.argtypes = [vp, ctypes.c_size_t]
    lib.crgpu_last_timing.argtypes = [vp, vp, vp]
    lib.crgpu_last_fill_breakdown.argtypes = [vp, vp, vp, vp]
    lib.crgpu_set_overlap.argtypes = [vp, i32]
    lib.crgpu_set_share_prefix.argtypes = [vp, i32]
    lib.crgpu_set_band.argtypes = [vp, i32]
    lib.crgpu_get_band.argtypes = [vp]
    lib.crgpu_get_band.restype = i32
    lib.crgpu_last_escaped.argtypes = [vp, ctypes.POINTER(ctypes.c_int * 2)]
    lib.crgpu_set_diag_shortcut.argtypes = [vp, i32]
    lib.crgpu_last_diag.argtypes = [vp, ctypes.POINTER(ctypes.c_int64 * 2)]
    lib.crgpu_sync.argtypes = [vp]
    lib.crgpu_stream.argtypes = [vp]
    lib.crgpu_stream.restype = vp
    lib.crgpu_qualfilter.argtypes = [vp, i32, vp, vp, i64, i32, i32, vp]
    lib.crgpu_align.argtypes = [vp, i32, ctypes.c_char_p, i32, vp, vp, i64, dbl, dbl, vp, vp, vp, vp, i64]
    lib.crgpu_quantify.argtypes = [vp, i32, ctypes.POINTER(QuantParams), vp, vp, vp, i64, vp, vp, vp, vp, vp, i64,
                                   vp, vp, vp, vp, ctypes.c_int32, ctypes.c_int32, vp]
    lib.crgpu_align_quantify.argtypes = [vp, i32, ctypes.c_char_p, i32, ctypes.POINTER(PathParams),
                                         ctypes.POINTER(QuantParams), vp, vp, i64, ctypes.POINTER(PathOut)]
    lib.crgpu_int_peak.argtypes = [vp, i32, ctypes.POINTER(dbl)]
    lib.crgpu_stage_reads.argtypes = [vp, i32, i32, vp, vp, i64]
    lib.crgpu_set_deferred_outputs.argtypes = [vp, i32]
    lib.crgpu_set_exact_shortcut.argtypes = [vp, i32]
    lib.crgpu_last_exact.argtypes = [vp]
    lib.crgpu_last_exact.restype = i64
    lib.crgpu_align_quantify_staged.argtypes = [vp, i32, ctypes.c_char_p, i32, ctypes.POINTER(PathParams), ctypes.POINTER(QuantParams),
                                                ctypes.POINTER(PathOut)]
    lib.crgpu_fastq_index.argtypes = [vp, i32, vp, i64, i32, ctypes.POINTER(FastqOut)]
    lib.crgpu_flash_merge.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, i64, ctypes.POINTER(MergeParams),
                                      ctypes.POINTER(MergeOut)]
    for name in ("crgpu_create", "crgpu_set_diag_shortcut", "crgpu_last_diag", "crgpu_set_overlap", "crgpu_set_share_prefix", "crgpu_set_band", "crgpu_last_escaped", "crgpu_set_traceback_budget", "crgpu_last_timing", "crgpu_last_fill_breakdown", "crgpu_sync", "crgpu_qualfilter",
                 "crgpu_align", "crgpu_quantify", "crgpu_align_quantify", "crgpu_int_peak", "crgpu_flash_merge", "crgpu_fastq_index",
                 "crgpu_stage_reads", "crgpu_align_quantify_staged", "crgpu_set_deferred_outputs", "crgpu_set_exact_shortcut"):
        getattr(lib, name).restype = i32
    _lib = lib
    return lib


def ptr(x):
    """Raw address of a numpy array, a torch tensor, an int address, or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


class Context:
    """One per GPU (crgpu_create).  Not thread-safe."""

    TIMING_NAMES = ("encode", "fill", "walk", "quantify", "qualfilter", "other")

    def __init__(self, device=0):
        self.lib = load()
        self.handle = ctypes.c_void_p()
        rc = self.lib.crgpu_create(ctypes.byref(self.handle), int(device))
        if rc:
            raise CrgpuError(rc, "crgpu_create(device=%d) failed: no usable sm_100 GPU (no CPU fallback exists)" % device)
        self.device = device

    def close(self):
        if self.handle:
            self.lib.crgpu_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc:
            raise CrgpuError(rc, self.lib.crgpu_last_error(self.handle).decode())

    def set_traceback_budget(self, nbytes):
        self.check(self.lib.crgpu_set_traceback_budget(self.handle, int(nbytes)))

    def set_overlap(self, on):
        self.check(self.lib.crgpu_set_overlap(self.handle, 1 if on else 0))

    def set_share_prefix(self, on):
        self.check(self.lib.crgpu_set_share_prefix(self.handle, 1 if on else 0))

    def set_band(self, half_width):
        """Half-width (read columns) of the banded two-pass fill; 0 = single-pass fill."""
        self.check(self.lib.crgpu_set_band(self.handle, int(half_width)))

    def band(self):
        return int(self.lib.crgpu_get_band(self.handle))

    def last_escaped(self):
        """(amplicon pass, HDR pass) reads of the last fused call that left the band and were re-aligned."""
        out = (ctypes.c_int * 2)()
        self.check(self.lib.crgpu_last_escaped(self.handle, ctypes.byref(out)))
        return int(out[0]), int(out[1])

    def set_exact_shortcut(self, on):
        """Reads identical to the amplicon skip the DP (default on); results do not depend on it."""
        self.check(self.lib.crgpu_set_exact_shortcut(self.handle, 1 if on else 0))

    def last_exact(self):
        return int(self.lib.crgpu_last_exact(self.handle))

    def set_diag_shortcut(self, on):
        """Diagonal shortcut of the banded fill (default on); results never depend on it."""
        self.check(self.lib.crgpu_set_diag_shortcut(self.handle, 1 if on else 0))

    def last_diag(self):
        """(read pairs of the last fused call's banded passes, pairs that still needed band pass + walk)."""
        out = (ctypes.c_int64 * 2)()
        self.check(self.lib.crgpu_last_diag(self.handle, ctypes.byref(out)))
        return int(out[0]), int(out[1])

    def last_fill_breakdown(self):
        """{'full' | 'score' | 'band': (ms, launches, DP cells evaluated)} of the last call's fill kernels."""
        ms = (ctypes.c_double * 3)()
        ln = (ctypes.c_int64 * 3)()
        cells = (ctypes.c_int64 * 3)()
        self.check(self.lib.crgpu_last_fill_breakdown(self.handle, ms, ln, cells))
        return {k: (float(ms[i]), int(ln[i]), int(cells[i])) for i, k in enumerate(("full", "score", "band"))}

    def last_timing(self):
        ms = (ctypes.c_float * 6)()
        ln = (ctypes.c_int64 * 6)()
        self.check(self.lib.crgpu_last_timing(self.handle, ms, ln))
        return ({k: float(ms[i]) for i, k in enumerate(self.TIMING_NAMES)},
                {k: int(ln[i]) for i, k in enumerate(self.TIMING_NAMES)})

    def stream_ptr(self):
        return int(self.lib.crgpu_stream(self.handle) or 0)

    def sync(self):
        self.check(self.lib.crgpu_sync(self.handle))

    def int_peak(self, which=0):
        v = ctypes.c_double()
        self.check(self.lib.crgpu_int_peak(self.handle, int(which), ctypes.byref(v)))
        return v.value
