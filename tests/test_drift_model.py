"""Host model of the score pass's arithmetic (crispresso_b200/csrc/gotoh_score2.cu).

k_gotoh_score2 evaluates the Gotoh recurrences in drift coordinates v' = v + ext * (row + column), where a gap extension
leaves a value unchanged.  This file restates, in plain Python integers, (a) the recurrences in the form the kernels use
(H3 = max(m, ix, iy) as the opening source, needle's zero end-gap penalties on the last amplicon row / last read column) and
(b) the same in drift coordinates with the kernel's constants, and checks cell by cell that (b) minus the drift is (a), and
that (a) reproduces the oracle's score and start cell (the oracle restates needle itself).  No GPU."""
import numpy as np
import pytest

from oracle import needle

SUB = {True: 5, False: -4}


def _sub(a, b, scale):
    if a == "N" and b == "N":
        return -1 * scale
    if a == "N" or b == "N":
        return -2 * scale
    return SUB[a == b] * scale


def plain_dp(amp, read, open_s, ext_s, scale):
    """m, ix, iy, h3 of every cell, kernel formulation (gotoh_fill.cu column_step), exact scaled integers."""
    La, Lb = len(amp), len(read)
    NEG = None
    m = np.zeros((La, Lb), np.int64); ix = np.zeros_like(m); iy = np.zeros_like(m); h3 = np.zeros_like(m)
    for x in range(Lb):
        for y in range(La):
            h3_diag = h3[y - 1, x - 1] if y > 0 and x > 0 else 0          # free boundary: max3 = 0 above row 0 / left of column 0
            h3_left = h3[y, x - 1] if x > 0 else 0
            ix_left = ix[y, x - 1] if x > 0 else -open_s
            m_left = m[y, x - 1] if x > 0 else 0
            h3_up = h3[y - 1, x] if y > 0 else 0
            iy_up = iy[y - 1, x] if y > 0 else -open_s
            m_up = m[y - 1, x] if y > 0 else 0
            m[y, x] = _sub(amp[y], read[x], scale) + h3_diag
            if y == La - 1:                                               # last amplicon row: opens from m only, zero penalties
                ix[y, x] = max(m_left, ix_left)
            else:
                ix[y, x] = max(h3_left - open_s, ix_left - ext_s)
            if x == Lb - 1:                                               # last read column likewise
                iy[y, x] = max(m_up, iy_up)
            else:
                iy[y, x] = max(h3_up - open_s, iy_up - ext_s)
            h3[y, x] = max(m[y, x], ix[y, x], iy[y, x])
    return m, ix, iy, h3


def drift_dp(amp, read, open_s, ext_s, scale):
    """The same matrices in drift coordinates, with the constants of gotoh_score2.cu: S' = S + 2 ext, c = ext - open, and
    `+ ext` terms on the zero-penalty row / column; boundary values carry the drift of their (virtual) cell."""
    La, Lb = len(amp), len(read)
    e, c = ext_s, ext_s - open_s
    d = lambda y, x: e * (y + x)
    m = np.zeros((La, Lb), np.int64); ix = np.zeros_like(m); iy = np.zeros_like(m); h3 = np.zeros_like(m)
    for x in range(Lb):
        for y in range(La):
            h3_diag = h3[y - 1, x - 1] if y > 0 and x > 0 else 0 + d(y - 1, x - 1)
            h3_left = h3[y, x - 1] if x > 0 else 0 + d(y, x - 1)
            ix_left = ix[y, x - 1] if x > 0 else -open_s + d(y, x - 1)
            m_left = m[y, x - 1] if x > 0 else 0 + d(y, x - 1)
            h3_up = h3[y - 1, x] if y > 0 else 0 + d(y - 1, x)
            iy_up = iy[y - 1, x] if y > 0 else -open_s + d(y - 1, x)
            m_up = m[y - 1, x] if y > 0 else 0 + d(y - 1, x)
            m[y, x] = (_sub(amp[y], read[x], scale) + 2 * e) + h3_diag
            ix[y, x] = max(m_left + e, ix_left + e) if y == La - 1 else max(h3_left + c, ix_left)
            iy[y, x] = max(m_up + e, iy_up + e) if x == Lb - 1 else max(h3_up + c, iy_up)
            h3[y, x] = max(m[y, x], ix[y, x], iy[y, x])
    return m, ix, iy, h3


def start_cell(h3):
    """needle's start-cell scan (App. A.4): last row left to right, then last column top to bottom, strict '>'."""
    La, Lb = h3.shape
    best, s1, s2 = None, La - 1, Lb - 1
    for x in range(Lb):
        if best is None or h3[La - 1, x] > best:
            best, s1, s2 = h3[La - 1, x], La - 1, x
    for y in range(La):
        if h3[y, Lb - 1] > best:
            best, s1, s2 = h3[y, Lb - 1], y, Lb - 1
    return int(best), s1, s2


def _cases():
    rng = np.random.default_rng(5)
    out = []
    for _ in range(40):
        La = int(rng.integers(2, 40))
        amp = "".join(rng.choice(list("ACGTN"), La, p=[.24, .24, .24, .24, .04]))
        kind = rng.random()
        if kind < 0.4:                                                    # an edited copy of the amplicon
            cut = int(rng.integers(0, La))
            k = int(rng.integers(0, 6))
            read = amp[:cut] + ("".join(rng.choice(list("ACGT"), k)) if rng.random() < 0.5 else "") + amp[min(La, cut + (k if rng.random() < 0.5 else 0)):]
            read = read.replace("N", "A") or "A"
        else:
            read = "".join(rng.choice(list("ACGTN"), int(rng.integers(1, 50)), p=[.24, .24, .24, .24, .04]))
        if len(read) < 2:
            read += "AC"
        gapopen, gapextend = [(10.0, 0.5), (10.0, 1.0), (5.0, 5.0), (12.5, 0.25), (4.0, 0.0)][int(rng.integers(0, 5))]
        out.append((amp, read, gapopen, gapextend))
    return out


@pytest.mark.parametrize("amp,read,gapopen,gapextend", _cases())
def test_drift_coordinates_are_an_exact_change_of_variables(amp, read, gapopen, gapextend):
    scale = 4                                                             # makes 12.5 / 0.25 integral
    open_s, ext_s = int(gapopen * scale), int(gapextend * scale)
    plain = plain_dp(amp, read, open_s, ext_s, scale)
    drift = drift_dp(amp, read, open_s, ext_s, scale)
    yy, xx = np.meshgrid(np.arange(len(amp)), np.arange(len(read)), indexing="ij")
    for p, q, name in zip(plain, drift, ("m", "ix", "iy", "max3")):
        assert np.array_equal(q - ext_s * (yy + xx), p), name
    # and the plain form is needle: score and start cell of the oracle's exact integer restatement
    best, s1, s2 = start_cell(plain[3])
    res, _, _, _ = needle.align_batch(amp, [read], gapopen, gapextend, use_int=True)
    assert res["score"][0] * scale == best
    assert (int(res["start1"][0]), int(res["start2"][0])) == (s1, s2)
