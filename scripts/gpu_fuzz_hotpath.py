"""Randomised differential test of crgpu_align_quantify against oracle.quantify.hot_path."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from crispresso_b200 import Context, aligner, hotpath, synth
from oracle import quantify

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 120.0
ctx = Context(0)
t0 = time.time()
case = bad = 0
while time.time() - t0 < budget:
    rng = np.random.default_rng(seed0 * 100000 + case)
    La = int(rng.integers(40, 420))
    amp, guide, cut, hdr = synth.make_case(int(rng.integers(1, 1 << 30)), La, hdr=True)
    if rng.random() < 0.25:
        a = list(amp); a[int(rng.integers(0, La))] = "N"; a[min(La - 1, cut + 2)] = "N"; amp = "".join(a)
        hdr = hdr[:min(La - 1, cut + 2)] + "N" + hdr[min(La - 1, cut + 2) + 1:]
    use_hdr = rng.random() < 0.5
    n = int(rng.integers(1, 400))
    buf, off = synth.make_reads(amp.replace("N", "A"), hdr.replace("N", "A") if use_hdr else None, cut, n, seed=int(rng.integers(1, 1 << 30)),
                                read_len=None if rng.random() < 0.5 else La, p_exact=float(rng.uniform(0.2, 0.8)),
                                sub_rate=float(rng.choice([0.0, 0.002, 0.02])), n_rate=float(rng.choice([0.0, 0.0, 0.01])),
                                rc_frac=float(rng.choice([0.0, 0.0, 0.2])), len_sigma=0.0)
    reads = [bytes(buf[off[i]:off[i + 1]]).decode() for i in range(n)]
    reads += [synth.random_seq(rng, int(rng.integers(2, 2 * La))) for _ in range(int(rng.integers(0, 10)))]
    packed = aligner.pack_reads(reads)
    window = int(rng.choice([0, 1, 1, 2, 7, 20]))
    has_guides = rng.random() < 0.8
    el, er = (int(rng.integers(0, 20)), int(rng.integers(0, 20))) if rng.random() < 0.5 else (15, 15)
    if el + er >= La:
        el = er = 0
    cuts = hotpath.cut_points_from_guides(amp, guide) if has_guides else []
    if "N" in guide:
        cuts = []
    try:
        inc = hotpath.include_mask(La, cuts, window, el, er)
    except ValueError:
        case += 1
        continue
    coding = ""
    exon = splice = None
    if rng.random() < 0.4 and La > 80:
        st = int(rng.integers(5, La // 2)); en = int(rng.integers(st + 10, La - 3))
        coding = amp[st:en]
        if amp.find(coding) != st:
            coding = ""
        else:
            exon, splice = hotpath.exon_masks(amp, coding)
    opt = dict(ignore_substitutions=bool(rng.random() < 0.15), ignore_insertions=bool(rng.random() < 0.15),
               ignore_deletions=bool(rng.random() < 0.15), hide_mutations_outside_window_NHEJ=bool(rng.random() < 0.3))
    thr = float(rng.choice([98.0, 98.0, 90.0, 99.5]))
    min_id = float(rng.choice([60.0, 60.0, 30.0, 75.5, 0.0]))
    flags = hotpath.quant_flags(hdr if use_hdr else "", opt["ignore_substitutions"], opt["ignore_insertions"], opt["ignore_deletions"],
                                window, opt["hide_mutations_outside_window_NHEJ"], coding)
    res = hotpath.run_hot_path(ctx, amp, packed, min_identity_score=min_id, hdr_amplicon=hdr if use_hdr else None, flags=flags,
                               hdr_thr=thr, inc=inc, exon=exon, splice=splice, want_rows=True)
    ora = quantify.hot_path(amp, packed, min_identity_score=min_id, hdr_amplicon=hdr if use_hdr else "",
                            opts=quantify.Opts(coding_seq=coding, expected_hdr_amplicon_seq=hdr if use_hdr else "",
                                               hdr_perfect_alignment_threshold=thr, window_around_sgrna=window, **opt),
                            include=np.nonzero(inc)[0], exon=np.nonzero(exon)[0] if coding else (),
                            splice=np.nonzero(splice)[0] if coding else ())
    red = res.red
    ok = red.n_total == ora["n_total"] and red.n_cells == ora["n_cells"]
    ok = ok and red.class_counts.tolist() == [ora["classes"][k] for k in ("UNMODIFIED", "NHEJ", "HDR", "MIXED")]
    for k, name in enumerate(hotpath.VECTOR_NAMES):
        ok = ok and red.vectors[k].tolist() == ora["vectors"][name].tolist()
    ok = ok and hotpath.Reductions.hist_dict(red.hist_inframe) == ora["hist_inframe"]
    ok = ok and hotpath.Reductions.hist_dict(red.hist_frameshift) == ora["hist_frameshift"]
    ok = ok and {nm: red.counter(nm) for nm in hotpath.COUNTER_NAMES} == ora["counters"]
    if ok and ora["n_total"]:
        df = hotpath.build_dataframe(res, ["r%d" % i for i in range(len(reads))], has_hdr=use_hdr, amplicon=amp)
        ok = ok and list(df.index) == [r["ID"] for r in ora["rows"]]
        for col in ("ref_seq", "align_str", "align_seq"):
            ok = ok and list(df[col]) == [r[col] for r in ora["rows"]]
        for col in ("UNMODIFIED", "NHEJ", "HDR", "MIXED", "n_mutated", "n_inserted", "n_deleted"):
            ok = ok and [int(x) for x in df[col]] == [int(p[col]) for p in ora["per_row"]]
    if not ok:
        bad += 1
        print("MISMATCH case", case, dict(La=La, n=len(reads), hdr=use_hdr, window=window, guides=has_guides, excl=(el, er), coding=bool(coding),
                                           thr=thr, min_id=min_id, **opt), "gpu", red.n_total, red.class_counts.tolist(), "ora", ora["n_total"], ora["classes"])
        if bad > 3:
            break
    case += 1
print("hot-path fuzz seed %d: %d cases in %.0f s, %d mismatching cases" % (seed0, case, time.time() - t0, bad))
sys.exit(1 if bad else 0)
