"""Fill-kernel throughput probe for one shape (device-resident, no rows)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from crispresso_b200 import Context, synth, _lib
La, n, rl = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sigma = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
amp, guide, cut, hdr = synth.make_case(1234, La)
buf, off = synth.make_reads_fast(amp, hdr, cut, n, seed=1234, read_len=rl, len_sigma=sigma)
ctx = Context(0)
d_buf = torch.from_numpy(buf).cuda(); d_off = torch.from_numpy(off).cuda()
d_recs = torch.zeros(n * 8, dtype=torch.int32, device="cuda")
best = None
for it in range(3):
    ctx.check(ctx.lib.crgpu_align(ctx.handle, _lib.MEM_DEVICE, amp.encode(), La, d_buf.data_ptr(), d_off.data_ptr(), n,
                                  10.0, 0.5, d_recs.data_ptr(), None, None, None, La + int(np.diff(off).max())))
    ms, ln = ctx.last_timing()
    best = ms if best is None or ms["fill"] < best["fill"] else best
cells = float(La) * float(off[-1])
print("La=%d n=%d len=%d sigma=%g: fill %.2f ms  %.1f GCUPS   walk %.2f ms" % (La, n, rl, sigma, best["fill"], cells / best["fill"] / 1e6, best["walk"]))
