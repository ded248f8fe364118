#!/usr/bin/env python
"""bench.py -- headline benchmark of the CRISPResso hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads R] [--config cfg2|cfg3|cfg4|cfg5]
                    [--scaling weak|strong]

A "step" = one pass of the hot path (CORE:1791-2072 + 2773-2869: alignment to the amplicon AND the
HDR amplicon with needle semantics, reverse-complement rescue, classification and all
histograms) over one batch of synthetic reads.  Workload at any N: BASELINE.json configs[1]
per GPU -- 2^20 single-end 250-bp reads vs a 250-bp amplicon + HDR amplicon (weak scaling: reads
shard across ranks, one NCCL all-reduce of the int64 histogram block per step).

--config selects another BASELINE.json workload: cfg3 = merged-PE-like reads N(300, 8) vs a 300-bp amplicon
with a 120-bp coding sequence (frameshift analysis), no HDR; cfg4 = pooled amplicons (lengths uniform 150-400, one
crgpu_align_quantify call per amplicon, as CRISPRessoPooled runs one CRISPResso per amplicon); cfg5 = 600-bp amplicon,
600-bp reads, no HDR.  --scaling strong keeps the TOTAL reads per step fixed (--reads) and splits them over the ranks.

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
    ap.add_argument("--reads", type=int, default=1 << 20, help="reads per GPU per step (--scaling strong: per step over all GPUs)")
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"], help="BASELINE.json workload (headline: cfg2)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--pool-amplicons", type=int, default=16, help="cfg4: amplicons per GPU per step (reads are split evenly)")
    ap.add_argument("--cpu-sample", type=int, default=100000, help="reads in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=0, help="reads per staged chunk of the end-to-end arm (default: reads / 2)")
    ap.add_argument("--e2e-alleles", type=int, default=1 << 16, help="capacity of the per-chunk allele table of the end-to-end arm")
    return ap.parse_args()


class Job:
    """One crgpu_align_quantify call of a step: an amplicon (+ HDR amplicon), its analysis options and its reads."""

    def __init__(self, amp, guide, hdr, buf, off, window=1, coding=""):
        from crispresso_b200 import hotpath
        self.amp, self.hdr, self.buf, self.off = amp, hdr, buf, off
        self.L = len(amp)
        self.n = len(off) - 1
        self.inc = hotpath.include_mask(self.L, hotpath.cut_points_from_guides(amp, guide), window, 15, 15)
        self.exon = self.splice = None
        if coding:
            self.exon, self.splice = hotpath.exon_masks(amp, coding)
        self.window, self.coding = window, coding
        self.flags = hotpath.quant_flags(hdr or "", window_around_sgrna=window, coding_seq=coding)
        self.kw = dict(hdr_amplicon=hdr, flags=self.flags, inc=self.inc, exon=self.exon, splice=self.splice)


def workload(config, n_reads, rank, pool_amplicons=16):
    """-> (jobs of one step on this rank, description)"""
    from crispresso_b200 import synth
    if config == "cfg2":
        amp, guide, cut, hdr = synth.make_case(SEED, AMPLICON_LEN)
        buf, off = synth.make_reads_fast(amp, hdr, cut, n_reads, seed=SEED + 17 * rank, read_len=READ_LEN)
        return [Job(amp, guide, hdr, buf, off)], (
            "cfg2: %d single-end %d-bp reads per GPU vs %d-bp amplicon + HDR amplicon (needle gapopen 10 / gapextend 0.5), "
            "RC rescue, classification + histograms" % (n_reads, READ_LEN, AMPLICON_LEN))
    if config == "cfg3":
        amp, guide, cut, _ = synth.make_case(SEED + 3, 300, hdr=False)
        buf, off = synth.make_reads_fast(amp, None, cut, n_reads, seed=SEED + 3 + 17 * rank, read_len=300, len_sigma=8.0, p_exact=0.8)
        return [Job(amp, guide, None, buf, off, window=20, coding=amp[cut - 60:cut + 60])], (
            "cfg3: %d merged-PE-like reads per GPU (lengths N(300, 8) clipped to [260, 340], 20 %% with an indel) vs a 300-bp amplicon, "
            "120-bp coding sequence around the cut (frameshift analysis), window 20, no HDR amplicon" % n_reads)
    if config == "cfg4":
        rng = np.random.default_rng(SEED + 4)
        lens = rng.integers(150, 401, size=200)                       # the 200 amplicons of the pooled run
        mine = [int(x) for x in lens[(rank * pool_amplicons) % 200:][:pool_amplicons]]
        while len(mine) < pool_amplicons:
            mine += [int(x) for x in lens[:pool_amplicons - len(mine)]]
        per = max(1, n_reads // pool_amplicons)
        jobs = []
        for j, L in enumerate(mine):
            amp, guide, cut, _ = synth.make_case(SEED + 400 + 1000 * rank + j, L, hdr=False)
            buf, off = synth.make_reads_fast(amp, None, cut, per, seed=SEED + 400 + 1000 * rank + j, read_len=L)
            jobs.append(Job(amp, guide, None, buf, off))
        return jobs, ("cfg4: %d pooled amplicons per GPU per step (lengths uniform 150-400: %s), %d reads each, one "
                      "crgpu_align_quantify call per amplicon, no HDR amplicon" % (pool_amplicons, mine, per))
    amp, guide, cut, _ = synth.make_case(SEED + 5, 600, hdr=False)
    buf, off = synth.make_reads_fast(amp, None, cut, n_reads, seed=SEED + 5 + 17 * rank, read_len=600)
    return [Job(amp, guide, None, buf, off)], "cfg5: %d single-end 600-bp reads per GPU vs a 600-bp amplicon, no HDR amplicon" % n_reads


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
def cpu_reference_step(job, sb, so, nthreads, use_reference=True):
    """One pass of the reference path on the host.  Stage 1 -- needle for the amplicon and the HDR
    amplicon + the RC rescue: the reference shells out to EMBOSS needle, a third-party binary that is not in this image, so
    its C restatement (oracle/needle_oracle.c, float32 as needle) runs on `nthreads` pthreads.  Stage 2 -- quantification:
    the reference's OWN unmodified process_df_chunk (CORE:428-753) through multiprocessing.Pool as run_crispresso drives it
    with -p <cores> (CORE:2773-2838), imported from baseline/_ref (pip-installed copy) or /root/reference; when neither is
    there, the oracle's Python restatement of it.  -> (seconds, result, stage seconds, kind of the quantifier)."""
    from oracle import quantify, ref_quantify
    process, qkind = None, "port (oracle/quantify.py)"
    if use_reference and ref_quantify.available():
        qkind = "reference (CRISPRessoCORE.process_df_chunk, mp.Pool of %d)" % nthreads

        def process(*a):
            return ref_quantify.process_rows_reference(*a, n_processes=nthreads)
    stages = {}
    opts = quantify.Opts(expected_hdr_amplicon_seq=job.hdr or "", coding_seq=job.coding, window_around_sgrna=job.window)
    t0 = time.time()
    res = quantify.hot_path(job.amp, (sb, so), hdr_amplicon=job.hdr or "", opts=opts, include=np.nonzero(job.inc)[0],
                            exon=np.nonzero(job.exon)[0] if job.exon is not None else (),
                            splice=np.nonzero(job.splice)[0] if job.splice is not None else (),
                            nthreads=nthreads, use_int=False, process=process, timings=stages)
    return time.time() - t0, res, stages, qkind


def cpu_baseline_dict(n, steps, dt, cells, stages, qkind, cores):
    al, qu = stages.get("align_prepare_s", 0.0), stages.get("quantify_s", 0.0)
    return {"value": n * steps / dt, "unit": "reads/s", "cores": cores,
            "kind": "reference" if qkind.startswith("reference") else "port",
            "gcups": cells / dt / 1e9,
            "stages": {"needle": {"kind": "port (oracle/needle_oracle.c, float32 as EMBOSS needle 6.6.0; the binary is not in the image)",
                                  "threads": cores, "reads_per_s": n * steps / al if al > 0 else None},
                       "quantify": {"kind": qkind, "reads_per_s": n * steps / qu if qu > 0 else None}},
            "sample": "%d reads/step x %d step(s) of the rank-0 workload: needle port on %d pthreads for every amplicon + RC rescue, "
                      "then the quantification stage (%s), %.1f s" % (n, steps, cores, qkind, dt)}


def cpu_baseline(args, job):
    n = min(args.cpu_sample, job.n)
    sb, so = job.buf[:job.off[n]].copy(), job.off[:n + 1].copy()
    cores = os.cpu_count() or 1
    dt, res, stages, qkind = cpu_reference_step(job, sb, so, cores)
    return cpu_baseline_dict(n, 1, dt, res["n_cells"], stages, qkind, cores)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    jobs, workload_text = workload(args.config, max(args.cpu_sample, 1), 0, 1 if args.config == "cfg4" else args.pool_amplicons)
    job = jobs[0]
    n = min(args.cpu_sample, job.n)
    sb, so = job.buf[:job.off[n]].copy(), job.off[:n + 1].copy()
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_reference_step(job, sb[:so[200]].copy(), so[:201].copy(), cores)
    t0 = time.time()
    cells, stages, qkind = 0, {}, ""
    for _ in range(args.steps):
        _dt, res, st, qkind = cpu_reference_step(job, sb, so, cores)
        cells += res["n_cells"]
        for k, v in st.items():
            stages[k] = stages.get(k, 0.0) + v
    dt = time.time() - t0
    value = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "aligned_reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "gcups": cells / dt / 1e9,
        "config": {"workload": workload_text + "; the CPU arm runs a bounded sample of %d reads per step" % n, "name": args.config,
                   "reads_per_step": n},
        "cpu_baseline": cpu_baseline_dict(n, args.steps, dt, cells, stages, qkind, cores),
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bind_to_gpu_numa_node(torch, local):
    """One process per GPU: run on the cores of the NUMA node the GPU's PCIe root hangs off, so that the pinned staging
    buffers (first touch) and the library's mailbox live in that node's memory -- eight ranks copying ~200 MB per step each
    through one node's memory controller is what the end-to-end arm scaled against.  Best effort: returns the node or None."""
    try:
        props = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


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
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    ext_stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=torch.device("cuda", local))

    # reads of one step on this rank: --reads per GPU (weak scaling), or --reads over all ranks (strong scaling)
    n = args.reads if args.scaling == "weak" else max(1, args.reads // world)
    jobs, workload_text = workload(args.config, n, rank, args.pool_amplicons)
    n = sum(j.n for j in jobs)
    n_alignments = sum(j.n * (2 if j.hdr else 1) for j in jobs)

    # integer issue peaks, measured live: (i) both integer pipes -- dependency-free IADD chains that
    # ptxas splits 1:1 over the alu (IADD3) and fma (IMAD.IADD) pipes, and a VIMNMX.S16x2 + IMAD 1:1 mix; (ii) the
    # integer-ALU pipe alone -- VIMNMX / VIADDMNMX / VIMNMX3 .S16x2 issue nowhere else (profiles/r01_notes.md)
    int_peak = max(ctx.int_peak(0), ctx.int_peak(4))
    alu_peak = max(ctx.int_peak(2), ctx.int_peak(3), ctx.int_peak(5))

    # ---- device-resident arm ------------------------------------------------------------------
    for j in jobs:
        j.d_buf = torch.from_numpy(j.buf).cuda()
        j.d_off = torch.from_numpy(j.off).cuda()
        j.dev_out = {
            "kept": torch.zeros(j.n, dtype=torch.uint8, device="cuda"),
            "aln": torch.zeros(j.n * _lib.ALN_REC.itemsize, dtype=torch.uint8, device="cuda"),
            "recs": torch.zeros(j.n * _lib.READ_REC.itemsize, dtype=torch.uint8, device="cuda"),
            "tenths_rep": torch.zeros(j.n, dtype=torch.int32, device="cuda"),
        }
        j.dev_ptrs = {k: v.data_ptr() for k, v in j.dev_out.items()}
    torch.cuda.synchronize()          # torch initialises these on its own stream; the library uses another

    from crispresso_b200.distributed import allreduce_reductions

    def allreduce(reds):
        for red in reds:
            allreduce_reductions(red, device=torch.device("cuda", local))

    fam_ms = {k: 0.0 for k in Context.TIMING_NAMES}
    fam_launch = {k: 0 for k in Context.TIMING_NAMES}

    def step_device(record=False):
        reds = []
        for j in jobs:
            red = hotpath.Reductions(j.L)
            hotpath.run_hot_path(ctx, j.amp, None, red=red, device_inputs=(j.d_buf.data_ptr(), j.d_off.data_ptr(), j.n, 0, j.dev_ptrs),
                                 **j.kw)
            if record:
                ms, ln = ctx.last_timing()
                for k in ms:
                    fam_ms[k] += ms[k]
                    fam_launch[k] += ln[k]
            reds.append(red)
        if args.config != "cfg4":          # (pooled amplicons: every amplicon is a result set of its own, nothing to reduce)
            allreduce(reds)
        return reds

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        reds = step_device()
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
        reds = step_device(record=True)
    ev1.record(ext_stream)
    barrier()
    t_wall1 = time.time()
    dev_ms = ev0.elapsed_time(ev1)
    # one extra, untimed step with the walk/fill overlap off: per-kernel times without a co-running kernel
    iso_ms = {k: 0.0 for k in Context.TIMING_NAMES}
    iso_kinds = {k: [0.0, 0, 0] for k in ("score", "band", "full")}
    iso_computed = 0
    ctx.set_overlap(False)
    for j in jobs:
        iso_red = hotpath.Reductions(j.L)
        hotpath.run_hot_path(ctx, j.amp, None, red=iso_red, device_inputs=(j.d_buf.data_ptr(), j.d_off.data_ptr(), j.n, 0, j.dev_ptrs), **j.kw)
        ms_, _ln = ctx.last_timing()
        for k in ms_:
            iso_ms[k] += ms_[k]
        for k, (ms_k, ln_k, cells_k) in ctx.last_fill_breakdown().items():
            iso_kinds[k][0] += ms_k; iso_kinds[k][1] += ln_k; iso_kinds[k][2] += cells_k
        iso_computed += iso_red.n_cells_computed
    ctx.set_overlap(True)
    torch.cuda.synchronize()
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    red = reds[0]
    n_total = sum(r.n_total for r in reds)

    # ---- end-to-end arm: pinned HOST buffers through the public host API -------------------------
    # hotpath.StagedPipeline on ONE context: every step's reads travel as BAM 4-bit codes (crgpu_stage_reads, asynchronous
    # H2D on the library's copy stream + unpack on the device) in chunks; the copy of the next chunk -- of this job, the next
    # job or the next step, as in a stream of batches -- runs while crgpu_align_quantify_staged works on the current chunk.
    # Every job brings back what CORE:2892-3992 consumes: per-read records, the RC-rescue rows, all reductions and the allele
    # table (the per-read arrays travel behind the next chunk's kernels; all of them are in host memory before the clock stops).
    items = []                               # the chunks of one step, in order: (job, lo, hi, packed bytes view, offsets, outputs)
    for j in jobs:
        chunk = args.e2e_chunk if args.e2e_chunk > 0 else j.n      # (measured: one chunk per call 37.9 ms, two 44.7, four 58.1)
        j.p_packed = torch.from_numpy(hotpath.pack_bam4(j.buf)).pin_memory()
        packed_np = j.p_packed.numpy()
        j.pinned = {
            "kept": torch.zeros(j.n, dtype=torch.uint8).pin_memory(),
            "aln": torch.zeros(j.n * _lib.ALN_REC.itemsize, dtype=torch.uint8).pin_memory(),
            "recs": torch.zeros(j.n * _lib.READ_REC.itemsize, dtype=torch.uint8).pin_memory(),
            "tenths_rep": torch.zeros(j.n, dtype=torch.int32).pin_memory(),
            "rc_read": torch.zeros(j.n, dtype=torch.int32).pin_memory(),
            "rc_aln": torch.zeros(j.n * _lib.ALN_REC.itemsize, dtype=torch.uint8).pin_memory(),
            "rc_recs": torch.zeros(j.n * _lib.READ_REC.itemsize, dtype=torch.uint8).pin_memory(),
        }
        outs = {k: v.numpy() for k, v in j.pinned.items()}
        for k, dt in (("aln", _lib.ALN_REC), ("rc_aln", _lib.ALN_REC), ("recs", _lib.READ_REC), ("rc_recs", _lib.READ_REC)):
            outs[k] = outs[k].view(dt)
        j.pipe = hotpath.StagedPipeline(ctx, j.amp, alleles=args.e2e_alleles, deferred=True, **j.kw)
        j.p_offs = []
        lo = 0
        while lo < j.n:
            hi = min(j.n, lo + chunk)
            while hi < j.n and (int(j.off[hi]) & 1):                  # packed chunks start on even base offsets
                hi += 1
            po = torch.from_numpy((j.off[lo:hi + 1] - j.off[lo]).astype(np.int64)).pin_memory()
            j.p_offs.append(po)
            items.append((j, packed_np[int(j.off[lo]) // 2:(int(j.off[hi]) + 1) // 2], po.numpy(), {k: v[lo:hi] for k, v in outs.items()},
                          hi == j.n))
            lo = hi
    n_chunks = len(items)

    def run_host_steps(k_steps):
        """k_steps passes over the step's chunks, as a section of an endless stream of steps: on entry the first chunk is
        already staged (by the previous section), every chunk's run is preceded by the staging of the chunk after it -- the
        last chunk stages the next step's first one -- so a section of k steps holds k copies of the step's reads and k runs.
        Returns the last step's [(reductions, allele table)] per job."""
        last = []
        for st in range(k_steps):
            last = []
            for c, (j, _pk, _po, out, job_done) in enumerate(items):
                nxt = items[(c + 1) % n_chunks]
                nxt[0].pipe.stage(nxt[1], nxt[2], packed=True)
                j.pipe.run(out)
                if job_done:
                    # (the per-read arrays of this job's last chunk travel behind the next chunk's kernels)
                    last.append(j.pipe.take_results(sync=False))
            if args.config != "cfg4":
                allreduce([r for r, _t in last])
        ctx.sync()                                                   # ... and are all in host memory here
        return last

    items[0][0].pipe.stage(items[0][1], items[0][2], packed=True)    # the stream's very first chunk
    run_host_steps(max(1, args.warmup - 1))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext_stream)
    host_res = run_host_steps(args.steps)
    e1.record(ext_stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    h2d_bytes = int(sum(j.p_packed.numel() + sum(o.numel() * 8 for o in j.p_offs) for j in jobs))
    n_alleles = int(sum(len(tb[0]) for _r, tb in host_res))
    d2h_bytes = int(n * (1 + _lib.ALN_REC.itemsize + _lib.READ_REC.itemsize + 4) + sum(r.flat().nbytes for r, _t in host_res) + 24 * n_alleles)
    same = all(bool(np.array_equal(rh.flat()[:-1], rd.flat()[:-1])) for (rh, _t), rd in zip(host_res, reds))   # (n_cells_computed aside)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_reads = n * world * args.steps
    value = total_reads / (dev_ms * 1e-3)
    e2e_value = total_reads / (e2e_ms * 1e-3)
    # full DP matrices of every alignment made: amplicon (+ HDR amplicon) x read (RC rescue cells are extra)
    cells_per_rank_step = float(sum((2 if j.hdr else 1) * j.L * float(j.off[-1]) for j in jobs))
    # cells the fill launches of one step actually evaluate (score pass: every cell once, the HDR pass only the rows
    # below the shared prefix; band pass: the band columns again, with flags; escapes and the RC rescue: single-pass
    # fill) -- the numerators of the kernel rooflines
    computed_per_step = float(iso_computed)
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
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "int16x2", "data": "synthetic",
        "gcups": (cells_per_rank_step * world * args.steps) / (dev_ms * 1e-3) / 1e9,
        "gcups_note": "La x Lb of every alignment delivered (2 per read with an HDR amplicon; reads identical to the amplicon are delivered "
                      "without DP) / time; `gcups_evaluated` counts only the DP cells the "
                      "kernels evaluate (the HDR pass reuses the rows it shares with the amplicon pass; the band pass evaluates the "
                      "band columns a second time, with flags; bit-identical results)",
        "gcups_evaluated": (computed_per_step * world * args.steps) / (dev_ms * 1e-3) / 1e9,
        "config": {"workload": workload_text, "name": args.config, "reads_per_gpu_per_step": n,
                   "alignments_per_read": n_alignments / max(n, 1), "calls_per_step": len(jobs),
                   "l2": "inputs and traceback exceed L2 (reads %d MB per GPU per step, traceback scratch up to 8 GB per batch)" % (
                       sum(j.buf.nbytes for j in jobs) >> 20),
                   "parallelism": "reads sharded x%d (%s scaling)" % (world, args.scaling),
                   "numa_node_of_rank0": numa},
        "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_ms / args.steps, "results_equal_device_arm": same,
                "distinct_alleles": n_alleles,
                "stream": "the timed K steps are a section of a stream of steps: each step's run is preceded by the staging (H2D copy) "
                          "of the next step's reads, so the section holds K copies of a step's reads, K runs and K sets of results in host "
                          "memory (synchronised at both ends); the first timed step's reads were copied during the last warm-up step",
                "api": "hotpath.StagedPipeline on one context: crgpu_stage_reads (BAM 4-bit reads from pinned memory, asynchronous "
                       "H2D on the library's copy stream, unpacked on the device; the next %d-read chunk is staged while the current "
                       "one runs) + crgpu_align_quantify_staged (per-read records, RC-rescue rows, reductions and the allele table "
                       "back in host memory: what CORE:2892-3992 consumes)" % (args.e2e_chunk if args.e2e_chunk > 0 else max(j.n for j in jobs))},
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
        # work the last call of the step did NOT have to do (bit-identical results either way): reads that
        # are the amplicon itself skip the DP (one representative stays); amplicon alignments whose traceback is provably the
        # diagonal skip the band pass and the walk
        "shortcuts": {"exact_reads_of_last_call": ctx.last_exact(), "pairs_in_plan_and_pairs_needing_the_amplicon_band": list(ctx.last_diag())},
        "clocks": clocks,
        "classes": {"n_total": int(n_total), "unmodified": int(sum(r.class_counts[0] for r in reds)),
                    "nhej": int(sum(r.class_counts[1] for r in reds)), "hdr": int(sum(r.class_counts[2] for r in reds)),
                    "mixed": int(sum(r.class_counts[3] for r in reds))},
    }
    if not args.no_cpu_baseline and world == 1:          # the CPU leg is timed at N = 1 only (bounded sample, rank 0)
        line["cpu_baseline"] = cpu_baseline(args, jobs[0])
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
