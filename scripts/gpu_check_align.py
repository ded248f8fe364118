"""Ad-hoc GPU check: crgpu_align vs the oracle on synthetic reads (run under gpurun)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from crispresso_b200 import Context, aligner, synth
from oracle import needle

def check(La, n, read_len, seed, len_sigma=0.0, n_rate=0.0):
    amp, guide, cut, hdr = synth.make_case(seed, La)
    buf, off = synth.make_reads(amp, hdr, cut, n, seed=seed, read_len=read_len, len_sigma=len_sigma, n_rate=n_rate)
    ctx = Context(0)
    t0 = time.time()
    recs, r, m, q = aligner.needle_align(ctx, amp, (buf, off))
    t1 = time.time()
    ores, orr, om, oq = needle.align_batch(amp, (buf, off), use_int=True, nthreads=8)
    t2 = time.time()
    bad = 0
    for i in range(n):
        ok = (r[i] == orr[i] and m[i] == om[i] and q[i] == oq[i] and recs["tenths"][i] == ores["tenths"][i]
              and recs["ident"][i] == ores["ident"][i] and recs["alnlen"][i] == ores["alnlen"][i]
              and float(recs["score"][i]) == ores["score"][i] and recs["start1"][i] == ores["start1"][i]
              and recs["start2"][i] == ores["start2"][i])
        if not ok:
            bad += 1
            if bad <= 3:
                print("MISMATCH read", i, "len", off[i+1]-off[i])
                print(" gpu", r[i]); print("    ", m[i]); print("    ", q[i], recs[i])
                print(" ora", orr[i]); print("    ", om[i]); print("    ", oq[i], ores[i])
    print("La=%d n=%d read_len=%s sigma=%g: mismatches=%d  gpu %.3fs oracle %.3fs timing=%s" % (
        La, n, read_len, len_sigma, bad, t1 - t0, t2 - t1, ctx.last_timing()))
    return bad

if __name__ == "__main__":
    bad = 0
    bad += check(250, 2000, 250, 1)
    bad += check(250, 2000, None, 2)
    bad += check(280, 1000, None, 3, n_rate=0.01)
    bad += check(300, 1000, 300, 4, len_sigma=8.0)
    bad += check(600, 500, 600, 5)
    bad += check(100, 500, None, 6)
    bad += check(1000, 200, 1000, 7)
    ctx = Context(0)
    for w in range(5):
        print("int peak which=%d: %.3f Tlane-op/s" % (w, ctx.int_peak(w) / 1e12))
    print("TOTAL MISMATCHES", bad)
    sys.exit(1 if bad else 0)
