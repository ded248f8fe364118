"""Restatement of FLASH 1.2.11 paired-end merging as CRISPResso runs it
(`flash R1 R2 --allow-outies --max-overlap 100 --min-overlap 4`, CRISPResso/CRISPRessoCORE.py:1655-1664)
-- TEST INFRASTRUCTURE (oracle/__init__.py).  FLASH is a third-party binary absent from
/root/reference; this follows SURVEY.md App. D and exists only to reach the reference's end-to-end
known-answer test (tests/crispresso_tests.py:127-195), which starts from paired-end FASTQ files.
"""
import numpy as np

_COMP = np.zeros(256, np.uint8)
for a, b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[a] = b


def _scan(s1, q1, s2, q2, min_overlap, max_overlap, best):
    """Slide read 2 along read 1 (overlap starts at i in read 1); best = [density, qscore, i] updated
    in place with FLASH's rule: lower mismatch density wins, ties go to the lower mismatch-quality
    score, the first (= longest) overlap wins exact ties."""
    l1, l2 = len(s1), len(s2)
    found = None
    for i in range(max(0, l1 - l2), l1 - min_overlap + 1):
        n = l1 - i
        a, b = s1[i:], s2[:n]
        valid = (a != ord("N")) & (b != ord("N"))
        eff = int(valid.sum())
        if eff < min_overlap:
            continue
        mism = valid & (a != b)
        score_len = np.float32(min(eff, max_overlap))
        nm = int(mism.sum())
        density = np.float32(nm) / score_len
        qscore = np.float32(np.minimum(q1[i:], q2[:n])[mism].sum()) / score_len if nm else np.float32(0.0)
        if density <= best[0] and (density < best[0] or qscore < best[1]):
            best[0], best[1] = density, qscore
            found = i
    return found


def merge_pair(seq1, qual1, seq2, qual2, min_overlap=4, max_overlap=100, max_mismatch_density=0.25, allow_outies=True):
    """-> (merged_seq, merged_qual, kind) or None.  kind: 'innie' / 'outie'."""
    s1 = np.frombuffer(seq1.encode(), np.uint8)
    q1 = np.frombuffer(qual1.encode(), np.uint8).astype(np.int32) - 33
    s2 = _COMP[np.frombuffer(seq2.encode(), np.uint8)][::-1]
    q2 = (np.frombuffer(qual2.encode(), np.uint8).astype(np.int32) - 33)[::-1]
    best = [np.float32(max_mismatch_density + 1.0), np.float32(0.0)]
    pos_in = _scan(s1, q1, s2, q2, min_overlap, max_overlap, best)
    pos_out = _scan(s2, q2, s1, q1, min_overlap, max_overlap, best) if allow_outies else None
    if best[0] > np.float32(max_mismatch_density):
        return None
    if pos_out is not None:
        left_s, left_q, right_s, right_q, i, kind = s2, q2, s1, q1, pos_out, "outie"
    elif pos_in is not None:
        left_s, left_q, right_s, right_q, i, kind = s1, q1, s2, q2, pos_in, "innie"
    else:
        return None
    n = len(left_s) - i
    a, qa, b, qb = left_s[i:], left_q[i:], right_s[:n], right_q[:n]
    # overlap: equal -> that base (max quality); different -> base of higher quality; equal quality ->
    # the second read's base unless it is N
    ov = np.where(a == b, a, np.where(qa > qb, a, np.where(qb > qa, b, np.where(b != ord("N"), b, a))))
    oq = np.where(a == b, np.maximum(qa, qb), np.where(qa > qb, qa, np.where(qb > qa, qb, np.where(b != ord("N"), qb, qa))))
    # FLASH's combine_reads works in read-1 / read-2 terms; for an outie the two are exchanged first
    if kind == "innie":
        ms = np.concatenate([left_s[:i], ov, right_s[n:]])
        mq = np.concatenate([left_q[:i], oq, right_q[n:]])
    else:
        # outie: the overhangs on both sides are adapter read-through and are dropped
        ms, mq = ov, oq
    return ms.tobytes().decode(), "".join(chr(int(v) + 33) for v in mq), kind
