"""ctypes front end of oracle/needle_oracle.c (restatement of EMBOSS needle 6.6.0 as
CRISPResso calls it, CRISPResso/CRISPRessoCORE.py:1791-1806).

TEST INFRASTRUCTURE -- see oracle/__init__.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OracleResult(ctypes.Structure):
    _fields_ = [
        ("alnlen", ctypes.c_int),
        ("ident", ctypes.c_int),
        ("gaps", ctypes.c_int),
        ("start1", ctypes.c_int),
        ("start2", ctypes.c_int),
        ("tenths", ctypes.c_int),
        ("score", ctypes.c_double),
    ]


RESULT_DTYPE = np.dtype(
    [("alnlen", "<i4"), ("ident", "<i4"), ("gaps", "<i4"), ("start1", "<i4"),
     ("start2", "<i4"), ("tenths", "<i4"), ("score", "<f8")], align=True)


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    src = os.path.join(_HERE, "needle_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []) + ["all"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.needle_align_batch.restype = ctypes.c_int
        _LIB.needle_align_batch.argtypes = [
            ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _LIB.oracle_identity_tenths.restype = ctypes.c_int
        _LIB.oracle_identity_tenths.argtypes = [ctypes.c_int, ctypes.c_int]
    return _LIB


def identity_tenths(ident, length):
    return lib().oracle_identity_tenths(int(ident), int(length))


def pack_reads(reads):
    """list of str/bytes -> (uint8 buffer, int64 offsets[n+1])"""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    offsets = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return buf, offsets


def align_batch(amplicon, reads, gapopen=10.0, gapextend=0.5, use_int=False, nthreads=1):
    """Align every read to `amplicon` as needle would.

    Returns (records, ref_seqs, markups, align_seqs): a structured array (RESULT_DTYPE)
    and three lists of str -- the three srspair rows parse_needle_output reads
    (CORE:1746-1754).
    """
    amp = amplicon.encode() if isinstance(amplicon, str) else bytes(amplicon)
    if isinstance(reads, tuple):
        buf, offsets = reads
    else:
        buf, offsets = pack_reads(reads)
    n = len(offsets) - 1
    maxlb = int(np.max(np.diff(offsets))) if n else 0
    slot = len(amp) + maxlb + 1
    ref = np.zeros(n * slot, np.uint8)
    mark = np.zeros(n * slot, np.uint8)
    qry = np.zeros(n * slot, np.uint8)
    res = np.zeros(n, dtype=RESULT_DTYPE)
    assert RESULT_DTYPE.itemsize == ctypes.sizeof(OracleResult)
    if n:
        rc = lib().needle_align_batch(
            amp, len(amp), buf.ctypes.data, offsets.ctypes.data, n, float(gapopen), float(gapextend),
            int(bool(use_int)), int(nthreads), ref.ctypes.data, mark.ctypes.data, qry.ctypes.data, slot,
            res.ctypes.data)
        if rc:
            raise ValueError("oracle: needle_align_batch failed rc=%d (bad character or empty sequence)" % rc)

    def rows(arr):
        out = []
        for i in range(n):
            L = int(res["alnlen"][i])
            out.append(arr[i * slot:i * slot + L].tobytes().decode())
        return out

    return res, rows(ref), rows(mark), rows(qry)
