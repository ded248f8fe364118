"""Randomised differential test of crgpu_align against the oracle (run under gpurun)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from crispresso_b200 import Context, aligner, synth
from oracle import needle

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
ctx = Context(0)
t0 = time.time()
case = 0
bad = 0
PEN = [(10.0, 0.5), (10.0, 0.5), (10.0, 0.5), (12.0, 2.0), (5.0, 0.0), (8.0, 0.25), (4.0, 4.0), (16.0, 1.5), (10.0, 1.0), (6.5, 0.125)]
while time.time() - t0 < budget:
    rng = np.random.default_rng(seed0 * 100000 + case)
    kind = rng.integers(0, 6)
    La = int(rng.choice([2, 3, 5, 17, 31, 33, 63, 64, 65, 100, 127, 128, 129, 159, 161, 200, 255, 256, 257, 300, 320, 321, 383, 385, 500, 513, 640, 641, 769, 1000, 1024])) if kind < 4 else int(rng.integers(2, 1025))
    amp = synth.random_seq(rng, La)
    if rng.random() < 0.3:
        a = list(amp)
        for _ in range(int(rng.integers(1, 4))):
            a[int(rng.integers(0, La))] = "N"
        amp = "".join(a)
    n = int(rng.integers(1, 60)) if La > 400 else int(rng.integers(1, 300))
    reads = []
    for _ in range(n):
        u = rng.random()
        if u < 0.15:
            r = synth.random_seq(rng, int(rng.integers(2, min(2048, 3 * La + 10))))
        elif u < 0.3:
            s_ = int(rng.integers(0, La)); e_ = int(rng.integers(s_, La + 1))
            r = synth.random_seq(rng, int(rng.integers(0, 30))) + amp[s_:e_].replace("N", "A") + synth.random_seq(rng, int(rng.integers(0, 30)))
        else:
            s = list(amp.replace("N", "ACGT"[int(rng.integers(0, 4))]))
            for _k in range(int(rng.integers(0, 5))):
                p = int(rng.integers(0, len(s) + 1))
                w = rng.random()
                if w < 0.33 and len(s) > 3:
                    del s[p:p + int(rng.integers(1, 25))]
                elif w < 0.66:
                    s[p:p] = list(synth.random_seq(rng, int(rng.integers(1, 25))))
                elif len(s):
                    q = min(p, len(s) - 1); s[q] = "ACGTN"[int(rng.integers(0, 5))]
            r = "".join(s)
        if len(r) < 2:
            r = r + "AC"
        reads.append(r[:2048])
    go, ge = PEN[int(rng.integers(0, len(PEN)))]
    packed = aligner.pack_reads(reads)
    try:
        recs, r, m, q = aligner.needle_align(ctx, amp, packed, go, ge)
    except Exception as e:
        # range errors are legitimate for big scale factors on long sequences
        if "int16 range" in str(e):
            case += 1
            continue
        raise
    ores, orr, om, oq = needle.align_batch(amp, packed, go, ge, use_int=True, nthreads=8)
    ok = (r == orr and m == om and q == oq and np.array_equal(recs["tenths"], ores["tenths"]) and np.array_equal(recs["ident"], ores["ident"])
          and np.array_equal(recs["start1"], ores["start1"]) and np.array_equal(recs["start2"], ores["start2"])
          and np.array_equal(recs["score"].astype(np.float64), ores["score"]))
    if not ok:
        bad += 1
        idx = [i for i in range(n) if r[i] != orr[i] or q[i] != oq[i] or m[i] != om[i] or recs["start1"][i] != ores["start1"][i] or recs["start2"][i] != ores["start2"][i] or float(recs["score"][i]) != ores["score"][i]]
        print("MISMATCH case", case, "La", La, "pen", go, ge, "n", n, "bad reads", idx[:5])
        i = idx[0] if idx else 0
        print(" amp", amp); print(" read", reads[i]); print(" gpu", r[i], q[i], recs[i]); print(" ora", orr[i], oq[i], ores[i])
        if bad > 3:
            break
    case += 1
print("fuzz seed %d: %d cases in %.0f s, %d mismatching cases" % (seed0, case, time.time() - t0, bad))
sys.exit(1 if bad else 0)
