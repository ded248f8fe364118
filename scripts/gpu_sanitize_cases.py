"""A small, deterministic tour of the hot path for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
banded two-pass fill with the shared HDR prefix, diagonal shortcut, band escapes, RC rescue, several traceback batches with
the walk/fill overlap on, odd read lengths, the K = 40 tiles, the staged pipeline with packed reads and deferred outputs,
the allele table, the paired-end merge and the FASTQ index -- every result checked against the oracle or the single call.
usage: compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/gpu_sanitize_cases.py"""
import sys
sys.path.insert(0, ".")
import numpy as np
from crispresso_b200 import Context, aligner, fastq, flash, hotpath, synth
from oracle import quantify

ctx = Context(0)
ctx.set_traceback_budget(16 << 20)         # several batches per call: the two scratch sets and three streams are all in play
n_checked = 0
for seed, La, rl, sigma, hdr_on, nreads in ((1, 250, 250, 0.0, True, 900), (2, 211, 211, 4.0, True, 700), (3, 300, 300, 8.0, False, 600),
                                            (4, 600, 600, 0.0, True, 260), (5, 97, 150, 3.0, False, 500), (6, 64, 64, 0.0, True, 400)):
    amp, guide, cut, hdr = synth.make_case(seed, La, hdr=hdr_on)
    packed = synth.make_reads(amp, hdr, cut, nreads, seed=seed, read_len=rl, len_sigma=sigma, rc_frac=0.06, n_rate=0.002)
    reads = [synth.random_seq(np.random.default_rng(seed), 40 + 7 * i) for i in range(5)]          # junk: escapes the band
    buf = np.concatenate([packed[0], np.frombuffer("".join(reads).encode(), np.uint8)])
    off = np.concatenate([packed[1], packed[1][-1] + np.cumsum([len(r) for r in reads])]).astype(np.int64)
    inc = hotpath.include_mask(La, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    flags = hotpath.quant_flags(hdr or "")
    res = hotpath.run_hot_path(ctx, amp, (buf, off), hdr_amplicon=hdr, flags=flags, inc=inc, want_rows=True, alleles=4096, min_identity_score=50.0)
    ora = quantify.hot_path(amp, (buf, off), hdr_amplicon=hdr or "", opts=quantify.Opts(expected_hdr_amplicon_seq=hdr or ""),
                            include=np.nonzero(inc)[0], min_identity_score=50.0)
    df = hotpath.build_dataframe(res, ["r%d" % i for i in range(len(off) - 1)], has_hdr=bool(hdr), amplicon=amp)
    assert list(df["align_seq"]) == [r["align_seq"] for r in ora["rows"]] and list(df["ref_seq"]) == [r["ref_seq"] for r in ora["rows"]]
    for k, name in enumerate(hotpath.VECTOR_NAMES):
        assert res.red.vectors[k].tolist() == ora["vectors"][name].tolist(), name
    assert res.red.class_counts.tolist() == [ora["classes"][k] for k in ("UNMODIFIED", "NHEJ", "HDR", "MIXED")]
    st = hotpath.run_hot_path_staged(ctx, amp, (buf, off), chunk_reads=300, packed=hotpath.pack_bam4(buf), hdr_amplicon=hdr, flags=flags,
                                     inc=inc, alleles=4096, deferred=True, min_identity_score=50.0)
    assert np.array_equal(st.red.results(), res.red.results()) and st.recs.tobytes() == res.recs.tobytes()
    assert sorted(st.allele_count.tolist()) == sorted(res.allele_count.tolist())
    # race detector of last resort (compute-sanitizer is closed on this pool): the same call again and again, stream
    # overlap on and off, must give the same bytes -- two scratch sets, three streams and the copy stream are all in play
    for rep in range(4):
        ctx.set_overlap(rep % 2 == 0)
        again = hotpath.run_hot_path(ctx, amp, (buf, off), hdr_amplicon=hdr, flags=flags, inc=inc, want_rows=True, alleles=4096, min_identity_score=50.0)
        assert again.recs.tobytes() == res.recs.tobytes() and again.aln.tobytes() == res.aln.tobytes() and np.array_equal(again.kept, res.kept)
        assert np.array_equal(again.red.results(), res.red.results()) and np.array_equal(again.tenths_rep, res.tenths_rep)
        for k in range(3):
            assert np.array_equal(again.rows[k], res.rows[k])
    ctx.set_overlap(True)
    n_checked += len(off) - 1
# S1 / S0 / FASTQ ingest
amp, guide, cut, _ = synth.make_case(9, 180, hdr=False)
s1, q1, s2, q2 = synth.make_pairs(amp, 400, read_len=120, seed=9)
m = flash.merge_pairs(ctx, s1, q1, s2, q2)
assert len(m.index) > 300
text = "".join("@r%d\n%s\n+\n%s\n" % (i, s1[i], q1[i]) for i in range(len(s1))).encode()
batch, consumed, nrec, _tot = fastq.index_text(ctx, text)
assert nrec == len(s1) and consumed == len(text)
keep = fastq.keep_mask(ctx, q1, 30, 5)
assert 0 < int(keep.sum()) <= len(q1)
print("sanitizer tour ok: %d reads checked against the oracle" % n_checked)
ctx.close()
