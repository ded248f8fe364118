"""Synthetic amplicon-sequencing reads of the shapes BASELINE.json / SURVEY.md 8(d) name."""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N", "-": "-"}


def random_seq(rng, n):
    return _ACGT[rng.integers(0, 4, size=n)].tobytes().decode()


def revcomp(s):
    return "".join(_COMP[c] for c in reversed(s))


def make_case(seed=1234, amplicon_len=250, hdr=True):
    """Amplicon, guide (20-mer ending 3 bp left of the cut), cut index, HDR amplicon (6-bp
    substitution block at the cut)."""
    rng = np.random.default_rng(seed)
    amp = random_seq(rng, amplicon_len)
    cut = amplicon_len // 2
    guide = amp[cut - 17:cut + 3]
    hdr_amp = None
    if hdr:
        block = "".join({"A": "C", "C": "G", "G": "T", "T": "A"}[c] for c in amp[cut - 3:cut + 3])
        hdr_amp = amp[:cut - 3] + block + amp[cut + 3:]
    return amp, guide, cut, hdr_amp


def make_reads(amp, hdr_amp, cut, n, seed=1234, read_len=None, p_exact=0.70, p_hdr=0.10, sub_rate=0.002,
               len_sigma=0.0, n_rate=0.0, rc_frac=0.0):
    """Read model of SURVEY 8(d) cfg2: 70 % unedited, 10 % HDR, 20 % one indel (size ~ geometric(0.2)
    capped 40, half insertions, centred within +-5 bp of the cut), per-base substitution errors.
    read_len: reads are cut/extended to this length when given (fixed-cycle sequencing); with
    len_sigma > 0 lengths are N(read_len, sigma) (merged paired-end).  Returns (uint8 buffer,
    int64 offsets)."""
    rng = np.random.default_rng(seed)
    L = len(amp)
    a = np.frombuffer(amp.encode(), dtype=np.uint8)
    h = np.frombuffer((hdr_amp or amp).encode(), dtype=np.uint8)
    kind = rng.random(n)
    if hdr_amp is None:
        p_hdr = 0.0
    pieces = []
    lens = np.empty(n, dtype=np.int64)
    flank = _ACGT[rng.integers(0, 4, size=64)]
    for i in range(n):
        if kind[i] < p_exact:
            r = a
        elif kind[i] < p_exact + p_hdr:
            r = h
        else:
            size = int(min(40, rng.geometric(0.2)))
            pos = int(cut + rng.integers(-5, 6))
            if rng.random() < 0.5:
                r = np.concatenate([a[:pos], _ACGT[rng.integers(0, 4, size=size)], a[pos:]])
            else:
                lo = max(0, pos - size // 2)
                r = np.concatenate([a[:lo], a[lo + size:]])
        if sub_rate > 0:
            m = rng.random(len(r)) < sub_rate
            if m.any():
                r = r.copy()
                r[m] = _ACGT[rng.integers(0, 4, size=int(m.sum()))]
        if n_rate > 0:
            m = rng.random(len(r)) < n_rate
            if m.any():
                r = r.copy()
                r[m] = ord("N")
        if read_len is not None:
            tgt = read_len if len_sigma <= 0 else int(np.clip(round(rng.normal(read_len, len_sigma)), read_len - 40, read_len + 40))
            if len(r) >= tgt:
                r = r[:tgt]
            else:
                r = np.concatenate([r, flank[:tgt - len(r)]])
        if rc_frac > 0 and rng.random() < rc_frac:
            r = np.frombuffer(revcomp(r.tobytes().decode()).encode(), dtype=np.uint8)
        pieces.append(r)
        lens[i] = len(r)
    offsets = np.zeros(n + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(lens)
    return np.concatenate(pieces).astype(np.uint8), offsets


def make_reads_fast(amp, hdr_amp, cut, n, seed=1234, read_len=None, **kw):
    """Large read sets: draw a pool of distinct molecules with make_reads and sample reads from it
    (amplicon sequencing is highly redundant), so that 10^6-10^8 reads are generated in seconds."""
    pool_n = min(n, 20000)
    buf, off = make_reads(amp, hdr_amp, cut, pool_n, seed=seed, read_len=read_len, **kw)
    if pool_n == n:
        return buf, off
    rng = np.random.default_rng(seed + 1)
    pick = rng.integers(0, pool_n, size=n)
    lens = np.diff(off)[pick]
    offsets = np.zeros(n + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(lens)
    if read_len is not None and kw.get("len_sigma", 0.0) <= 0:
        out = buf.reshape(pool_n, read_len)[pick].reshape(-1)
    else:
        out = np.empty(int(offsets[-1]), dtype=np.uint8)
        starts = off[:-1][pick]
        idx = np.repeat(starts - offsets[:-1], lens) + np.arange(int(offsets[-1]))
        out[:] = buf[idx]
    return np.ascontiguousarray(out), offsets


def make_pairs(amp, n, read_len=150, seed=1234, sub_rate=0.01, n_rate=0.002, short_frac=0.08, junk_frac=0.03,
               low_complexity_frac=0.0, coarse_quals=False):
    """Paired-end reads of one amplicon (the input of the FLASH step, CORE:1655-1664): read 1 = the first
    read_len bases of the fragment, read 2 = reverse complement of its last read_len bases; sequencing
    errors and N calls grow towards the 3' ends, where qualities drop.  short_frac of the fragments are
    shorter than a read (the mates then run into the adapter on both sides: an "outie"), junk_frac of the
    pairs are unrelated sequences.  coarse_quals draws qualities from 4 bins (as current instruments
    do), which makes equal-quality and equal-density ties common.  Returns lists (s1, q1, s2, q2) of str."""
    rng = np.random.default_rng(seed)
    a = np.frombuffer(amp.encode(), dtype=np.uint8)
    adapter1, adapter2 = _ACGT[rng.integers(0, 4, size=read_len)], _ACGT[rng.integers(0, 4, size=read_len)]
    bins = np.array([2, 12, 23, 37])
    s1, q1, s2, q2 = [], [], [], []

    def sequence(frag, adapter, L):
        r = np.concatenate([frag, adapter])[:L].copy()
        ramp = np.linspace(0.3, 3.0, len(r))
        m = rng.random(len(r)) < sub_rate * ramp
        r[m] = _ACGT[rng.integers(0, 4, size=int(m.sum()))]
        m2 = rng.random(len(r)) < n_rate * ramp
        r[m2] = ord("N")
        q = np.clip(np.round(rng.normal(36, 3, size=len(r)) - 6 * ramp * rng.random()), 2, 41).astype(np.int64)
        q[m] = np.minimum(q[m], rng.integers(2, 30, size=int(m.sum())))
        q[m2] = 2
        if coarse_quals:
            q = bins[np.abs(q[:, None] - bins[None, :]).argmin(axis=1)]
        return r.tobytes().decode(), "".join(chr(int(v) + 33) for v in q)

    for _ in range(n):
        u = rng.random()
        L1 = read_len if rng.random() < 0.9 else int(rng.integers(max(1, read_len // 3), read_len + 1))
        L2 = read_len if rng.random() < 0.9 else int(rng.integers(max(1, read_len // 3), read_len + 1))
        if u < junk_frac:
            f1, f2 = _ACGT[rng.integers(0, 4, size=L1)], _ACGT[rng.integers(0, 4, size=L2)]
            r1, qa = sequence(f1, adapter1, L1)
            r2, qb = sequence(f2, adapter2, L2)
        else:
            if u < junk_frac + low_complexity_frac:
                unit = _ACGT[rng.integers(0, 4, size=int(rng.integers(1, 4)))]
                frag = np.tile(unit, 2 * read_len)[:int(rng.integers(read_len, 2 * read_len))]
            elif u < junk_frac + low_complexity_frac + short_frac:
                lo = int(rng.integers(0, len(a) // 2))
                frag = a[lo:lo + int(rng.integers(max(8, read_len // 4), read_len))]
            else:
                frag = a
                if rng.random() < 0.2:                           # an indel somewhere in the middle
                    pos, size = int(rng.integers(len(a) // 3, 2 * len(a) // 3)), int(min(40, rng.geometric(0.2)))
                    frag = np.concatenate([a[:pos], a[pos + size:]]) if rng.random() < 0.5 else \
                        np.concatenate([a[:pos], _ACGT[rng.integers(0, 4, size=size)], a[pos:]])
            r1, qa = sequence(frag, adapter1, L1)
            rc = np.frombuffer(revcomp(frag.tobytes().decode()).encode(), dtype=np.uint8)
            r2, qb = sequence(rc, adapter2, L2)
        s1.append(r1); q1.append(qa); s2.append(r2); q2.append(qb)
    return s1, q1, s2, q2
