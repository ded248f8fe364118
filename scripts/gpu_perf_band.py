"""Fused hot path, band on/off, kernels serialised (CRGPU_NO_OVERLAP=1 recommended).  usage: gpu_perf_band.py [La] [n] [B...]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from crispresso_b200 import Context, _lib, hotpath, synth
La = int(sys.argv[1]) if len(sys.argv) > 1 else 250
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 19
Bs = [int(x) for x in sys.argv[3:]] or [0, 24]
amp, guide, cut, hdr = synth.make_case(1234, La)
buf, off = synth.make_reads_fast(amp, hdr, cut, n, seed=1234, read_len=La)
inc = hotpath.include_mask(La, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
flags = hotpath.quant_flags(hdr)
ctx = Context(0)
d_buf = torch.from_numpy(buf).cuda(); d_off = torch.from_numpy(off).cuda()
out = {"kept": torch.zeros(n, dtype=torch.uint8, device="cuda"), "aln": torch.zeros(n * 32, dtype=torch.uint8, device="cuda"),
       "recs": torch.zeros(n * 16, dtype=torch.uint8, device="cuda"), "tenths_rep": torch.zeros(n, dtype=torch.int32, device="cuda")}
ptrs = {k: v.data_ptr() for k, v in out.items()}
torch.cuda.synchronize()
for hdr_on in (True, False):
    for B in Bs:
        ctx.set_band(B)
        best = None
        for it in range(3):
            hotpath.run_hot_path(ctx, amp, None, hdr_amplicon=hdr if hdr_on else None, flags=flags if hdr_on else hotpath.quant_flags(""), inc=inc,
                                 device_inputs=(d_buf.data_ptr(), d_off.data_ptr(), n, La, ptrs))
            ms, ln = ctx.last_timing()
            best = ms if best is None or ms["fill"] < best["fill"] else best
        print("La=%d n=%d hdr=%s B=%d: fill %.2f ms walk %.2f ms escaped %s launches %s" % (La, n, hdr_on, B, best["fill"], best["walk"], ctx.last_escaped(), ln["fill"]))
