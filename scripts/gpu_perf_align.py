"""Ad-hoc throughput probe of crgpu_align with device-resident inputs."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from crispresso_b200 import Context, synth, _lib

def run(La, n, read_len, want_rows, budget_gb=8):
    amp, guide, cut, hdr = synth.make_case(1234, La)
    buf, off = synth.make_reads_fast(amp, hdr, cut, n, seed=1234, read_len=read_len)
    ctx = Context(0)
    ctx.set_traceback_budget(budget_gb << 30)
    d_buf = torch.from_numpy(buf).cuda(); d_off = torch.from_numpy(off).cuda()
    d_recs = torch.zeros(n * 8, dtype=torch.int32, device="cuda")
    slot = La + int(np.diff(off).max())
    rows = [torch.zeros(n * slot, dtype=torch.uint8, device="cuda") for _ in range(3)] if want_rows else [None] * 3
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        ctx.check(ctx.lib.crgpu_align(ctx.handle, _lib.MEM_DEVICE, amp.encode(), La, d_buf.data_ptr(), d_off.data_ptr(), n,
                                      10.0, 0.5, d_recs.data_ptr(), *[_lib.ptr(r) for r in rows], slot))
        torch.cuda.synchronize(); dt = time.time() - t0
        ms, ln = ctx.last_timing()
        cells = float(La) * float(off[-1])
        print("La=%d n=%d rows=%s wall %.1f ms  fill %.2f ms (%.1f GCUPS) walk %.2f ms encode %.2f launches=%d  e2e %.1f GCUPS" % (
            La, n, want_rows, dt * 1e3, ms["fill"], cells / ms["fill"] / 1e6, ms["walk"], ms["encode"], ln["fill"], cells / dt / 1e9))

run(250, 1 << 20, 250, False)
run(250, 1 << 20, 250, True)
run(600, 1 << 18, 600, False)
run(300, 1 << 19, 300, False)
