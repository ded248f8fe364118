#!/usr/bin/env python
"""bench.py -- headline benchmark of the CRISPResso hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads R]

A "step" = one pass of the hot path (CORE:1791-2072 + 2773-2869: alignment to the amplicon AND the
HDR amplicon with needle semantics, reverse-complement rescue, classification and all
histograms) over one batch of synthetic reads.  Workload at any N: BASELINE.json configs[1]
per GPU -- 2^20 single-end 250-bp reads vs a 250-bp amplicon + HDR amplicon (weak scaling: reads
shard across ranks, one NCCL all-reduce of the int64 histogram block per step).

Prints ONE JSON line (rank 0).  `value` = reads/s with inputs resident in HBM, `e2e` = the same
through the C ABI with pinned HOST buffers (H2D of the reads and D2H of the per-read records
inside the timed region).  `--impl reference` times the CPU restatement of the reference path
(oracle/: C needle port on all host cores + the Python quantification loop) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OPS_PER_CELL = 13          # scalar integer ops per DP cell (continuity figure)
ALU_INSTR_PER_CELL = 1.5   # integer-ALU-pipe lane-instructions per DP cell that no formulation avoids
AMPLICON_LEN = 250
READ_LEN = 250
SEED = 1234


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=1 << 20, help="reads per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=200000, help="reads in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=0, help="reads per staged chunk of the end-to-end arm (default: reads / 2)")
    ap.add_argument("--e2e-alleles", type=int, default=1 << 16, help="capacity of the per-chunk allele table of the end-to-end arm")
    return ap.parse_args()


def workload(n_reads, rank):
    from crispresso_b200 import hotpath, synth
    amp, guide, cut, hdr = synth.make_case(SEED, AMPLICON_LEN)
    buf, off = synth.make_reads_fast(amp, hdr, cut, n_reads, seed=SEED + 17 * rank, read_len=READ_LEN)
    inc = hotpath.include_mask(AMPLICON_LEN, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    return amp, guide, cut, hdr, buf, off, inc


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = max(mx, int(float(f[1])))
                if t0 <= t <= t1 + 0.1:
                    sm.append(int(float(f[0])))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": 0}
        return {"sm_mhz": int(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(amp, hdr, inc, buf, off, nthreads):
    """One pass of the reference path restated on the CPU (oracle/): needle port on `nthreads`
    host threads for the amplicon and the HDR amplicon + RC rescue + the Python per-read loop."""
    from oracle import quantify
    t0 = time.time()
    res = quantify.hot_path(amp, (buf, off), hdr_amplicon=hdr, opts=quantify.Opts(expected_hdr_amplicon_seq=hdr),
                            include=np.nonzero(inc)[0], nthreads=nthreads, use_int=False)
    return time.time() - t0, res


def cpu_baseline(args, amp, hdr, inc, buf, off):
    n = min(args.cpu_sample, len(off) - 1)
    sb, so = buf[:off[n]].copy(), off[:n + 1].copy()
    cores = os.cpu_count() or 1
    dt, res = cpu_reference_step(amp, hdr, inc, sb, so, cores)
    return {"value": n / dt, "unit": "reads/s", "cores": cores, "kind": "port",
            "gcups": res["n_cells"] / dt / 1e9,
            "sample": "first %d reads of the rank-0 workload, both amplicons + RC rescue + quantification, "
                      "C needle port on %d pthreads + single-process Python quantifier, %.1f s" % (n, cores, dt)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    amp, guide, cut, hdr, buf, off, inc = workload(max(args.cpu_sample, 1), 0)
    n = min(args.cpu_sample, len(off) - 1)
    sb, so = buf[:off[n]].copy(), off[:n + 1].copy()
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_reference_step(amp, hdr, inc, sb[:so[200]].copy(), so[:201].copy(), cores)
    t0 = time.time()
    cells = 0
    for _ in range(args.steps):
        _dt, res = cpu_reference_step(amp, hdr, inc, sb, so, cores)
        cells += res["n_cells"]
    dt = time.time() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "aligned_reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "gcups": cells / dt / 1e9,
        "config": {"workload": "cfg2: single-end %d-bp reads vs %d-bp amplicon + HDR amplicon; CPU arm runs a bounded "
                               "sample of %d reads per step" % (READ_LEN, AMPLICON_LEN, n), "reads_per_step": n},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port",
                         "sample": "%d reads/step x %d steps; oracle C needle port (float32, as needle) on %d pthreads + "
                                   "Python quantifier (the reference's own needle is a missing third-party binary)" % (
                                       n, args.steps, cores)},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    from crispresso_b200 import Context, _lib, hotpath

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    ext_stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=torch.device("cuda", local))

    n = args.reads
    amp, guide, cut, hdr, buf, off, inc = workload(n, rank)
    L = len(amp)
    flags = hotpath.quant_flags(hdr)

    # integer issue peaks, measured live: (i) both integer pipes -- dependency-free IADD chains that
    # ptxas splits 1:1 over the alu (IADD3) and fma (IMAD.IADD) pipes, and a VIMNMX.S16x2 + IMAD 1:1 mix; (ii) the
    # integer-ALU pipe alone -- VIMNMX / VIADDMNMX / VIMNMX3 .S16x2 issue nowhere else (profiles/r01_notes.md)
    int_peak = max(ctx.int_peak(0), ctx.int_peak(4))
    alu_peak = max(ctx.int_peak(2), ctx.int_peak(3), ctx.int_peak(5))

    # ---- device-resident arm ------------------------------------------------------------------
    d_buf = torch.from_numpy(buf).cuda()
    d_off = torch.from_numpy(off).cuda()
    dev_out = {
        "kept": torch.zeros(n, dtype=torch.uint8, device="cuda"),
        "aln": torch.zeros(n * _lib.ALN_REC.itemsize, dtype=torch.uint8, device="cuda"),
        "recs": torch.zeros(n * _lib.READ_REC.itemsize, dtype=torch.uint8, device="cuda"),
        "tenths_rep": torch.zeros(n, dtype=torch.int32, device="cuda"),
    }
    dev_ptrs = {k: v.data_ptr() for k, v in dev_out.items()}
    torch.cuda.synchronize()          # torch initialises these on its own stream; the library uses another

    from crispresso_b200.distributed import allreduce_reductions

    def allreduce(red):
        allreduce_reductions(red, device=torch.device("cuda", local))

    fam_ms = {k: 0.0 for k in Context.TIMING_NAMES}
    fam_launch = {k: 0 for k in Context.TIMING_NAMES}

    def step_device(record=False):
        red = hotpath.Reductions(L)
        hotpath.run_hot_path(ctx, amp, None, hdr_amplicon=hdr, flags=flags, inc=inc, red=red,
                             device_inputs=(d_buf.data_ptr(), d_off.data_ptr(), n, READ_LEN, dev_ptrs))
        if record:
            ms, ln = ctx.last_timing()
            for k in ms:
                fam_ms[k] += ms[k]
                fam_launch[k] += ln[k]
        allreduce(red)
        return red

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        red = step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record(ext_stream)
    for _ in range(args.steps):
        red = step_device(record=True)
    ev1.record(ext_stream)
    barrier()
    t_wall1 = time.time()
    dev_ms = ev0.elapsed_time(ev1)
    # one extra, untimed step with the walk/fill overlap off: per-kernel times without a co-running kernel
    iso_ms = {k: 0.0 for k in Context.TIMING_NAMES}
    iso_ln = {k: 0 for k in Context.TIMING_NAMES}
    ctx.set_overlap(False)
    iso_red = hotpath.Reductions(L)
    hotpath.run_hot_path(ctx, amp, None, hdr_amplicon=hdr, flags=flags, inc=inc, red=iso_red,
                         device_inputs=(d_buf.data_ptr(), d_off.data_ptr(), n, READ_LEN, dev_ptrs))
    ms_, ln_ = ctx.last_timing()
    iso_kinds = ctx.last_fill_breakdown()
    for k in ms_:
        iso_ms[k] += ms_[k]
        iso_ln[k] += ln_[k]
    ctx.set_overlap(True)
    torch.cuda.synchronize()
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    cells_step = red.n_cells if world == 1 else None
    n_total = red.n_total

    # ---- end-to-end arm: pinned HOST buffers through the public host API -------------------------
    # hotpath.StagedPipeline on ONE context: every step's reads travel as BAM 4-bit codes (crgpu_stage_reads, asynchronous
    # H2D on the library's copy stream + unpack on the device) in chunks; the copy of the next chunk -- of this step or of the
    # next one, as in a stream of batches -- runs while crgpu_align_quantify_staged works on the current chunk.  Every step
    # brings back what CORE:2892-3992 consumes: per-read records, the RC-rescue rows, all reductions and the allele table.
    if args.e2e_chunk <= 0:
        args.e2e_chunk = n            # (measured: one chunk per step 37.9 ms, two 44.7, four 58.1 -- per-call costs)
    bounds = [(lo, min(n, lo + args.e2e_chunk)) for lo in range(0, n, args.e2e_chunk)]
    assert all(int(off[lo]) % 2 == 0 for lo, _hi in bounds), "packed chunks start on even base offsets"
    n_chunks = len(bounds)
    p_packed = torch.from_numpy(hotpath.pack_bam4(buf)).pin_memory()
    p_offs = [torch.from_numpy((off[lo:hi + 1] - off[lo]).astype(np.int64)).pin_memory() for lo, hi in bounds]
    pinned = {
        "kept": torch.zeros(n, dtype=torch.uint8).pin_memory(),
        "aln": torch.zeros(n * _lib.ALN_REC.itemsize, dtype=torch.uint8).pin_memory(),
        "recs": torch.zeros(n * _lib.READ_REC.itemsize, dtype=torch.uint8).pin_memory(),
        "tenths_rep": torch.zeros(n, dtype=torch.int32).pin_memory(),
        "rc_read": torch.zeros(n, dtype=torch.int32).pin_memory(),
        "rc_aln": torch.zeros(n * _lib.ALN_REC.itemsize, dtype=torch.uint8).pin_memory(),
        "rc_recs": torch.zeros(n * _lib.READ_REC.itemsize, dtype=torch.uint8).pin_memory(),
    }
    outs = {k: v.numpy() for k, v in pinned.items()}
    for k, dt in (("aln", _lib.ALN_REC), ("rc_aln", _lib.ALN_REC), ("recs", _lib.READ_REC), ("rc_recs", _lib.READ_REC)):
        outs[k] = outs[k].view(dt)
    packed_np = p_packed.numpy()
    pipe = hotpath.StagedPipeline(ctx, amp, hdr_amplicon=hdr, flags=flags, inc=inc, alleles=args.e2e_alleles, deferred=True)

    def stage_chunk(c):
        lo, hi = bounds[c]
        pipe.stage(packed_np[int(off[lo]) // 2:(int(off[hi]) + 1) // 2], p_offs[c].numpy(), packed=True)

    def run_host_steps(k_steps):
        """k_steps passes over the read set as one stream of chunks; returns the last step's (reductions, allele table)."""
        last = None
        stage_chunk(0)
        for st in range(k_steps):
            for c, (lo, hi) in enumerate(bounds):
                if c + 1 < n_chunks:
                    stage_chunk(c + 1)
                elif st + 1 < k_steps:
                    stage_chunk(0)                                   # the next step's first chunk
                pipe.run({k: v[lo:hi] for k, v in outs.items()})
            # (the per-read arrays of this step's last chunk travel behind the next step's kernels)
            red_s, table = pipe.take_results(sync=False)
            allreduce(red_s)
            last = (red_s, table)
        ctx.sync()                                                   # ... and are all in host memory here
        return last

    run_host_steps(max(1, args.warmup - 1))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext_stream)
    red_h, alleles_h = run_host_steps(args.steps)
    e1.record(ext_stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d_bytes = int(p_packed.numel() + sum(o.numel() * 8 for o in p_offs))
    d2h_bytes = int(n * (1 + _lib.ALN_REC.itemsize + _lib.READ_REC.itemsize + 4) + red_h.flat().nbytes * n_chunks +
                    24 * len(alleles_h[0]))
    same = bool(np.array_equal(red_h.flat()[:-1], red.flat()[:-1]))     # every reduction (n_cells_computed aside: bookkeeping)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_reads = n * world * args.steps
    value = total_reads / (dev_ms * 1e-3)
    e2e_value = total_reads / (e2e_ms * 1e-3)
    cells_per_rank_step = 2.0 * L * float(off[-1])          # amplicon + HDR amplicon, full DP matrices (RC rescue cells are extra)
    # cells the fill launches of one step actually evaluate (score pass: every cell once, the HDR pass only the rows
    # below the shared prefix; band pass: the band columns again, with flags; escapes and the RC rescue: single-pass
    # fill) -- the numerators of the kernel rooflines
    computed_per_step = float(iso_red.n_cells_computed)
    peak = int_peak / 1e12
    alu_pk = alu_peak / 1e12
    # Roofline model: per packed cell pair the recurrences need 1 add + 2 VIADDMNMX.S16x2 + 1
    # VIMNMX3.S16x2; the three min/max instructions issue only on the integer-ALU pipe => 1.5 ALU lane-instructions
    # per cell, peak = the measured single-pipe rate.  Per-kernel durations come from one extra step right after the
    # timed region (same data) with the stream overlap off, so every launch runs alone between its CUDA events.
    kinds = {}
    for kname, label in (("score", "k_gotoh_score2<G,K> (score pass: drift coordinates, two read columns per step, no flags)"),
                         ("band", "k_gotoh_band<G,K> (band pass, flags)"),
                         ("full", "k_gotoh_fill<G,K> (single pass with flags: band escapes + RC rescue)")):
        ms_k, ln_k, cells_k = iso_kinds[kname]
        if ln_k == 0 or ms_k <= 0:
            continue
        tc = cells_k / (ms_k * 1e-3) / 1e12
        kinds[kname] = {"kernel": label, "launches_per_step": ln_k, "ms_per_step": ms_k, "ms_per_launch": ms_k / ln_k,
                        "cells_per_step": cells_k, "cells_per_launch": cells_k / ln_k, "tcups": tc,
                        "achieved": tc * ALU_INSTR_PER_CELL, "frac": tc * ALU_INSTR_PER_CELL / alu_pk,
                        "survey_13ops_tiops": tc * OPS_PER_CELL, "survey_13ops_frac_of_dual_pipe_peak": tc * OPS_PER_CELL / peak}
    dom = max(kinds, key=lambda k: kinds[k]["ms_per_step"])
    iso_fill_total = sum(v["ms_per_step"] for v in kinds.values())
    blended_tcups = computed_per_step / (iso_fill_total * 1e-3) / 1e12
    # inside the timed region the launches overlap each other (tail back-fill) and the walks, so their individual
    # durations are not additive; the whole step's effective rate is reported instead
    step_tcups_evaluated = computed_per_step / (dev_ms / args.steps * 1e-3) / 1e12
    step_tcups_useful = cells_per_rank_step / (dev_ms / args.steps * 1e-3) / 1e12
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = 6650.0
    traffic = None
    traffic_all = None
    try:
        traffic_all = json.load(open(os.path.join(ROOT, "profiles", "fill_traffic.json")))
        traffic = traffic_all[dom]["dram_bytes_per_cell"] * kinds[dom]["cells_per_launch"]
    except Exception:
        pass
    line = {
        "metric": "aligned_reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16x2", "data": "synthetic",
        "gcups": (cells_per_rank_step * world * args.steps) / (dev_ms * 1e-3) / 1e9,
        "gcups_note": "La x Lb of every alignment made (2 per read) / time; `gcups_evaluated` counts only the DP cells the "
                      "kernels evaluate (the HDR pass reuses the rows it shares with the amplicon pass; the band pass evaluates the "
                      "band columns a second time, with flags; bit-identical results)",
        "gcups_evaluated": (computed_per_step * world * args.steps) / (dev_ms * 1e-3) / 1e9,
        "config": {"workload": "cfg2: %d single-end %d-bp reads per GPU vs %d-bp amplicon + HDR amplicon (needle "
                               "gapopen 10 / gapextend 0.5), RC rescue, classification + histograms" % (n, READ_LEN, L),
                   "reads_per_gpu_per_step": n, "alignments_per_read": 2, "l2": "inputs and traceback exceed L2 (reads %d MB, "
                   "traceback scratch 8 GB per batch)" % (buf.nbytes >> 20), "parallelism": "reads sharded x%d" % world},
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_ms / args.steps, "results_equal_device_arm": same,
                "distinct_alleles": len(alleles_h[0]),
                "api": "hotpath.StagedPipeline on one context: crgpu_stage_reads (BAM 4-bit reads from pinned memory, asynchronous "
                       "H2D on the library's copy stream, unpacked on the device; the next %d-read chunk is staged while the current "
                       "one runs) + crgpu_align_quantify_staged (per-read records, RC-rescue rows, reductions and the allele table "
                       "back in host memory: what CORE:2892-3992 consumes)" % args.e2e_chunk},
        "gpu_launches": int(sum(fam_launch.values())),
        "kernel_ms_per_step": {k: fam_ms[k] / args.steps for k in fam_ms},
        "roofline": {"bound": "int_alu", "kernel": kinds[dom]["kernel"], "achieved": kinds[dom]["achieved"], "peak": alu_pk,
                     "unit": "T lane-instr/s on the integer-ALU pipe", "frac": kinds[dom]["frac"], "traffic": traffic,
                     "alu_instr_per_cell": ALU_INSTR_PER_CELL, "cells_per_launch": kinds[dom]["cells_per_launch"],
                     "ms_per_launch": kinds[dom]["ms_per_launch"], "tcups": kinds[dom]["tcups"],
                     "peak_source": "measured live: crgpu_int_peak, best of the VIMNMX / VIADDMNMX / VIMNMX3 .S16x2 probes (one "
                                    "pipe; the dual-pipe IADD3+IMAD rate is %.2f)" % peak,
                     "peak_tcups": alu_pk / ALU_INSTR_PER_CELL,
                     "model": "per packed cell pair: 1 add (m) + 2 VIADDMNMX.S16x2 (ix, iy; the gap extension is carried as a "
                              "coordinate drift) + 1 VIMNMX3.S16x2; the 3 min/max instructions issue only on the integer-ALU "
                              "pipe => 1.5 ALU lane-instructions per DP cell; flags, boundary hand-over and the walk are overhead",
                     "by_kernel": kinds,
                     "fill_blended": {"tcups": blended_tcups, "frac": blended_tcups * ALU_INSTR_PER_CELL / alu_pk,
                                      "cells_evaluated_per_step": computed_per_step, "ms_per_step": iso_fill_total},
                     "survey_8d": {"ops_per_cell": OPS_PER_CELL, "peak_tiops": peak, "peak_theoretical_tiops": 148 * 128 * 1.965e9 / 1e12,
                                   "achieved_tiops": blended_tcups * OPS_PER_CELL, "frac": blended_tcups * OPS_PER_CELL / peak,
                                   "note": "SURVEY 8(d)'s accounting (13 scalar ops per cell against the 32-bit lane-op rate of both "
                                           "pipes): one packed .S16x2 lane-instruction serves two cells and VIADDMNMX / VIMNMX3 fuse two "
                                           "scalar ops each, so this ratio can exceed 1 and is kept only for continuity with the round-1 "
                                           "lines; `frac` above is the bound that holds"},
                     "in_timed_region": {"note": "fill launches of consecutive batches and the traceback walks overlap on three streams; "
                                                 "per-launch durations there are not additive",
                                         "fill_ms_per_step_concurrent": fam_ms["fill"] / args.steps,
                                         "step_tcups_evaluated": step_tcups_evaluated, "step_tcups_useful": step_tcups_useful,
                                         "step_frac_evaluated": step_tcups_evaluated * ALU_INSTR_PER_CELL / alu_pk},
                     "isolated_kernel_ms_per_step": iso_ms},
        "roofline_hbm": {"bound": "hbm", "kernel": "k_gotoh_band<G,K> (1 flag byte per band cell)",
                         "achieved": (kinds["band"]["tcups"] * 1e3 if "band" in kinds else None), "peak": hbm_peak, "unit": "GB/s",
                         "frac": (kinds["band"]["tcups"] * 1e3 / hbm_peak if "band" in kinds else None), "bytes_per_cell": 1.0,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else "fallback"},
        "band": {"half_width": ctx.band(), "escaped_amplicon_hdr": list(ctx.last_escaped())},
        "clocks": clocks,
        "classes": {"n_total": int(n_total), "unmodified": int(red.class_counts[0]), "nhej": int(red.class_counts[1]),
                    "hdr": int(red.class_counts[2]), "mixed": int(red.class_counts[3])},
    }
    if not args.no_cpu_baseline and world == 1:          # the CPU leg is timed at N = 1 only (bounded sample, rank 0)
        line["cpu_baseline"] = cpu_baseline(args, amp, hdr, inc, buf, off)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
