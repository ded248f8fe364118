"""CPU models of three pieces of device bit logic, transcribed line by line from the CUDA sources, against plain
restatements: the quantifier's match-run skip (quantify.cu: match_run), the unaligned 4-byte load (crgpu_common.cuh: load4,
with k_any_lower's edge masks from alleles.cu) and the position-keyed row hash of the allele table (alleles.cu:
k_hash_rows).  The kernels themselves are checked against the oracle on the GPU (tests/test_gpu_*.py); these tests pin the
formulas where no GPU is needed."""
import numpy as np

M32 = 0xFFFFFFFF
M64 = 0xFFFFFFFFFFFFFFFF


# ---- quantify.cu: match_run ---------------------------------------------------------------------------------------------
def _pack_ops(ops_in_storage_order):
    words = [0] * ((len(ops_in_storage_order) + 15) // 16 + 1)
    for j, op in enumerate(ops_in_storage_order):
        words[j >> 4] |= int(op) << ((j & 15) * 2)
    return words


def _clz32(x):
    return 32 - int(x).bit_length()


def _ffs32(x):
    return (int(x) & -int(x)).bit_length()


def match_run_model(words, n, rev, c):
    """Columns from c on that are matches, up to the end of the op word column c lies in (0: column c is not a match)."""
    if rev:
        j = n - 1 - c
        k = j & 15
        word = words[j >> 4]
        m = word if k == 15 else (word & ((1 << (2 * k + 2)) - 1))
        if m == 0:
            return k + 1
        top = (31 - _clz32(m)) // 2
        return k - top if top < k else 0
    k = c & 15
    avail = min(16 - k, n - c)
    m = (words[c >> 4] >> (2 * k)) & M32
    if avail < 16:
        m &= (1 << (2 * avail)) - 1
    if m == 0:
        return avail
    return (_ffs32(m) - 1) // 2


def test_match_run_skips_exactly_the_matches_of_the_current_word():
    rng = np.random.default_rng(5)
    for trial in range(400):
        n = int(rng.integers(1, 120))
        # mostly matches, some runs of the other three ops
        fwd = np.zeros(n, np.int64)
        for _ in range(int(rng.integers(0, 6))):
            a = int(rng.integers(0, n))
            fwd[a:a + int(rng.integers(1, 20))] = int(rng.integers(1, 4))
        for rev in (False, True):
            stored = fwd[::-1] if rev else fwd                   # the walker stores forward rows last column first
            words = _pack_ops(stored)
            # garbage above the row's last op must not matter
            words[(n - 1) >> 4] |= (0xFFFFFFFF << (2 * ((n - 1) & 15) + 2)) & M32 if (n - 1) & 15 != 15 else 0
            c = 0
            visited = []
            while c < n:
                run = match_run_model(words, n, rev, c)
                # plain restatement: matches from c on, cut at the boundary of the 16-op word column c is stored in
                j = n - 1 - c if rev else c
                room = (j & 15) + 1 if rev else min(16 - (j & 15), n - c)
                want = 0
                while want < room and fwd[c + want] == 0:
                    want += 1
                assert run == want, (trial, rev, c, run, want)
                if run:
                    assert np.all(fwd[c:c + run] == 0)
                    c += run
                else:
                    visited.append(c)
                    c += 1
            assert visited == [int(i) for i in np.nonzero(fwd)[0]]      # every non-match column is looked at, no match column is


# ---- crgpu_common.cuh: load4, alleles.cu: k_any_lower's masks ------------------------------------------------------------
def load4_model(mem, base, i, length):
    """mem: bytes of the whole allocation; the string is mem[base : base + length]."""
    p = base + i
    w = p & ~3
    mis = p & 3
    rem = length - i
    lo = int.from_bytes(mem[w:w + 4], "little")
    hi = 0
    touched = [w]
    if mis and rem > 4 - mis:
        hi = int.from_bytes(mem[w + 4:w + 8], "little")
        touched.append(w + 4)
    v = ((lo | (hi << 32)) >> (8 * mis)) & M32                     # __funnelshift_r(lo, hi, 8 * mis)
    if rem < 4:
        v &= (1 << (8 * rem)) - 1
    return v, touched


def test_load4_reads_the_string_and_nothing_past_its_last_word():
    rng = np.random.default_rng(6)
    for trial in range(300):
        length = int(rng.integers(1, 70))
        base = int(rng.integers(0, 9))
        mem = bytes(rng.integers(1, 256, base + length + 16, dtype=np.uint8))
        s = mem[base:base + length]
        last_word_of_string = (base + length - 1) & ~3
        for i in range(0, length, 4):
            v, touched = load4_model(mem, base, i, length)
            want = int.from_bytes(s[i:i + 4].ljust(4, b"\0"), "little")
            assert v == want
            assert max(touched) <= last_word_of_string                # no aligned word beyond the one holding the last byte


def any_lower_model(mem, p0, p1):
    a0 = p0 & ~3
    nwords = (((p1 + 3) & ~3) - a0) >> 2
    acc = 0
    for i in range(nwords):
        wa = a0 + 4 * i
        w = int.from_bytes(mem[wa:wa + 4].ljust(4, b"\0"), "little")
        if wa < p0:
            w &= (M32 << (8 * (p0 - wa))) & M32
        if wa + 4 > p1:
            w &= M32 >> (8 * (wa + 4 - p1))
        acc |= w
    return (acc & 0x20202020) != 0


def test_any_lower_looks_at_the_bytes_of_the_batch_only():
    rng = np.random.default_rng(7)
    for trial in range(300):
        p0 = int(rng.integers(0, 9))
        n = int(rng.integers(1, 60))
        body = bytearray(rng.choice(np.frombuffer(b"ACGTN", np.uint8), n).tobytes())
        has = bool(rng.integers(0, 2))
        if has:
            k = int(rng.integers(0, n))
            body[k] = body[k] | 0x20
        mem = b"a" * p0 + bytes(body) + b"zzzzzzzz"                   # lower case right before and after the batch
        assert any_lower_model(mem, p0, p0 + n) == has


# ---- alleles.cu: k_hash_rows -------------------------------------------------------------------------------------------
SEED1, SEED2 = 0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F


def fmix64(x):
    x &= M64
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & M64
    x ^= x >> 33
    x = (x * 0xC4CEB9FE1A85EC53) & M64
    x ^= x >> 33
    return x


_COMP = {ord("A"): ord("T"), ord("C"): ord("G"), ord("G"): ord("C"), ord("T"): ord("A"), ord("U"): ord("A"), ord("N"): ord("N")}


def comp_up(c):
    return _COMP.get(c & 0xDF if chr(c).isalpha() else c, c)


def row_hash_model(read, ops_stored, rc, fields):
    """read: bytes as given; ops_stored: 2-bit ops in STORAGE order (forward rows: last column first; RC rows: forward
    order); fields = (cls, n_mutated, n_inserted, n_deleted).  The order of the additions is irrelevant (lanes)."""
    length, ncol = len(read), len(ops_stored)
    h1 = h2 = 0
    for w in range((length + 3) >> 2):
        c4 = 0
        for q in range(4):
            i = 4 * w + q
            if i < length:
                c = comp_up(read[length - 1 - i]) if rc else read[i]
                c4 |= c << (8 * q)
        x = c4 | ((w + 1) << 32)
        h1 = (h1 + fmix64(x ^ SEED1)) & M64
        h2 = (h2 + fmix64(((x + SEED2) & M64) * SEED1)) & M64
    for j, op in enumerate(ops_stored):
        if op >= 2:
            col = j if rc else ncol - 1 - j
            x = col | (op << 32) | (1 << 62)
            h1 = (h1 + fmix64(x ^ SEED1)) & M64
            h2 = (h2 + fmix64(((x + SEED2) & M64) * SEED1)) & M64
    cls, nm, ni, nd = fields
    f = cls | (nm << 8) | (ni << 24) | (nd << 44)
    g = ncol | (length << 32)
    h1 = fmix64(h1 + fmix64(f ^ SEED2) + fmix64(g + SEED1))
    h2 = fmix64(h2 ^ fmix64(f + SEED1) ^ fmix64((g * SEED2 + 1) & M64))
    return h1, h2


def _revcomp(b):
    return bytes(comp_up(c) for c in reversed(b))


def test_a_forward_row_and_an_rc_row_with_the_same_text_hash_alike():
    """A read aligned forward and the reverse complement of that read rescued on the other strand spell the same three text
    rows (CORE:1873-2000): the allele table must count them as one allele."""
    rng = np.random.default_rng(8)
    seen = {}
    for trial in range(200):
        n = int(rng.integers(20, 90))
        read = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), n).tobytes())
        # forward-order ops of the alignment: n read bases (match / mismatch / insertion) + some deletions
        fwd = []
        for _ in range(n):
            fwd.append(int(rng.choice([0, 0, 0, 0, 1, 2])))
            if rng.random() < 0.05:
                fwd.extend([3] * int(rng.integers(1, 6)))
        fields = (int(rng.integers(0, 4)), int(rng.integers(0, 9)), int(rng.integers(0, 9)), int(rng.integers(0, 30)))
        fw = row_hash_model(read, fwd[::-1], False, fields)             # forward rows are stored last column first
        rc = row_hash_model(_revcomp(read), fwd, True, fields)          # the RC row of the reverse-complemented read
        assert fw == rc
        key = (read, tuple(fwd), fields)
        assert seen.setdefault(fw, key) == key                          # and different rows do not collide
        # one gap moved by a column, one base changed, one field changed: different hashes
        if 3 in fwd:
            k = fwd.index(3)
            moved = fwd[:]
            moved[k], moved[k - 1 if k else k + 1] = moved[k - 1 if k else k + 1], moved[k]
            if moved != fwd:
                assert row_hash_model(read, moved[::-1], False, fields) != fw
        other = bytes([read[0] ^ 0x06]) + read[1:]
        assert row_hash_model(other, fwd[::-1], False, fields) != fw
        assert row_hash_model(read, fwd[::-1], False, (fields[0], fields[1] + 1, fields[2], fields[3])) != fw
        # a lower-case copy of a forward read is another allele (align_seq keeps the case); RC rows are upper-cased
        assert row_hash_model(read.lower(), fwd[::-1], False, fields) != fw
        assert row_hash_model(_revcomp(read).lower(), fwd, True, fields) == rc


# ---- traceback_walk.cu: the walk's three-way decision (App. A.4) ------------------------------------------------------------
F_NM, F_NX, F_NY = 1, 2, 4


def direction_if_chain(prev, contL, contD, f):
    """The decision as SURVEY App. A.4 states it (and as the walker had it before it became branch-free)."""
    if prev == 1 and contL:
        return 1
    if prev == 2 and contD:
        return 2
    if not (f & F_NM):
        if prev == 1 and not (f & F_NX):
            return 1
        if prev == 2 and not (f & F_NY):
            return 2
        return 0
    if not (f & F_NX):
        return 1
    return 2


def direction_selects(prev, contL, contD, f):
    """traceback_walk.cu: the same decision as selects."""
    nm, nx, ny = (f & F_NM) != 0, (f & F_NX) != 0, (f & F_NY) != 0
    wasL, wasD = prev == 1, prev == 2
    dir_m = 1 if (wasL and not nx) else 2 if (wasD and not ny) else 0
    dir_g = 2 if nx else 1
    d = dir_g if nm else dir_m
    d = 2 if (wasD and contD) else d
    d = 1 if (wasL and contL) else d
    return d


def test_branch_free_direction_equals_the_if_chain_on_every_input():
    for prev in (0, 1, 2):
        for contL in (False, True):
            for contD in (False, True):
                for f in range(32):
                    assert direction_selects(prev, contL, contD, f) == direction_if_chain(prev, contL, contD, f)


def test_case_folded_identity_equals_the_base_code_compare():
    """fold_base(x) == fold_base(y) <=> base_code(x) == base_code(y) on the accepted alphabet (A C G T U N, either case)."""
    code = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3, "N": 4}

    def fold(c):
        c &= 0xDF
        return ord("T") if c == ord("U") else c

    alphabet = [ord(ch) for ch in "ACGTUNacgtun"]
    for x in alphabet:
        for y in alphabet:
            assert (fold(x) == fold(y)) == (code[chr(x).upper()] == code[chr(y).upper()])
