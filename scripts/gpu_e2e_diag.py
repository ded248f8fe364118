"""Where the end-to-end step's time over the device-resident step goes (cfg2, one B200): the staged pipeline with and
without the allele table / deferred outputs, host time inside stage() / run() / take_results(), and the event time of the
'other' kernel family (which holds the allele kernels) with the stream overlap off.  Run on the GPU box:
    python scripts/gpu_e2e_diag.py [reads]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from crispresso_b200 import Context, _lib, hotpath  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    steps = 6
    ctx = Context(0)
    jobs, _text = bench.workload("cfg2", n, 0)
    j = jobs[0]
    d_buf, d_off = torch.from_numpy(j.buf).cuda(), torch.from_numpy(j.off).cuda()
    dev = {"kept": torch.zeros(j.n, dtype=torch.uint8, device="cuda"),
           "aln": torch.zeros(j.n * _lib.ALN_REC.itemsize, dtype=torch.uint8, device="cuda"),
           "recs": torch.zeros(j.n * _lib.READ_REC.itemsize, dtype=torch.uint8, device="cuda"),
           "tenths_rep": torch.zeros(j.n, dtype=torch.int32, device="cuda")}
    ptrs = {k: v.data_ptr() for k, v in dev.items()}
    torch.cuda.synchronize()

    def dev_step():
        red = hotpath.Reductions(j.L)
        hotpath.run_hot_path(ctx, j.amp, None, red=red, device_inputs=(d_buf.data_ptr(), d_off.data_ptr(), j.n, 0, ptrs), **j.kw)

    for _ in range(3):
        dev_step()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        dev_step()
    ctx.sync()
    print("device-resident        %.2f ms/step" % ((time.perf_counter() - t0) / steps * 1e3), flush=True)

    packed = torch.from_numpy(hotpath.pack_bam4(j.buf)).pin_memory()
    offs = torch.from_numpy(j.off.astype(np.int64)).pin_memory()
    pinned = {"kept": torch.zeros(j.n, dtype=torch.uint8).pin_memory(),
              "aln": torch.zeros(j.n * _lib.ALN_REC.itemsize, dtype=torch.uint8).pin_memory(),
              "recs": torch.zeros(j.n * _lib.READ_REC.itemsize, dtype=torch.uint8).pin_memory(),
              "tenths_rep": torch.zeros(j.n, dtype=torch.int32).pin_memory(),
              "rc_read": torch.zeros(j.n, dtype=torch.int32).pin_memory(),
              "rc_aln": torch.zeros(j.n * _lib.ALN_REC.itemsize, dtype=torch.uint8).pin_memory(),
              "rc_recs": torch.zeros(j.n * _lib.READ_REC.itemsize, dtype=torch.uint8).pin_memory()}
    outs = {k: v.numpy() for k, v in pinned.items()}
    for k, dt in (("aln", _lib.ALN_REC), ("rc_aln", _lib.ALN_REC), ("recs", _lib.READ_REC), ("rc_recs", _lib.READ_REC)):
        outs[k] = outs[k].view(dt)
    minimal = {k: outs[k] for k in ("kept", "aln", "recs")}

    for alleles, deferred, out, label in ((0, True, minimal, "no table, deferred, no RC/trep outputs"),
                                          (0, True, outs, "no table, deferred"),
                                          (0, False, outs, "no table, outputs in the call"),
                                          (1 << 16, True, outs, "table, deferred (bench)"),
                                          (1 << 16, False, outs, "table, outputs in the call")):
        pipe = hotpath.StagedPipeline(ctx, j.amp, alleles=alleles, deferred=deferred, **j.kw)
        host = {"stage": 0.0, "run": 0.0, "take": 0.0}

        def run_steps(k, rec):
            pipe.stage(packed.numpy(), offs.numpy(), packed=True)
            for st in range(k):
                a = time.perf_counter()
                if st + 1 < k:
                    pipe.stage(packed.numpy(), offs.numpy(), packed=True)
                b = time.perf_counter()
                pipe.run(out)
                c = time.perf_counter()
                pipe.take_results(sync=False)
                d = time.perf_counter()
                if rec:
                    host["stage"] += b - a; host["run"] += c - b; host["take"] += d - c
            ctx.sync()

        run_steps(2, False)
        t0 = time.perf_counter()
        run_steps(steps, True)
        dt = (time.perf_counter() - t0) / steps * 1e3
        print("%-42s %.2f ms/step   host: stage %.2f run %.2f take %.2f" % (
            label, dt, host["stage"] / steps * 1e3, host["run"] / steps * 1e3, host["take"] / steps * 1e3), flush=True)

    # event time of the kernel families with the overlap off, with and without the table
    ctx.set_overlap(False)
    for alleles in (0, 1 << 16):
        pipe = hotpath.StagedPipeline(ctx, j.amp, alleles=alleles, deferred=False, **j.kw)
        pipe.stage(packed.numpy(), offs.numpy(), packed=True)
        pipe.run(outs)
        ms, ln = ctx.last_timing()
        print("alleles=%d  families ms: %s" % (alleles, {k: round(v, 3) for k, v in ms.items()}), flush=True)
    ctx.set_overlap(True)


if __name__ == "__main__":
    main()
