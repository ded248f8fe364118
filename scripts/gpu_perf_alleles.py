"""Cost of the device-side allele grouping on the bench workload."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from crispresso_b200 import Context, _lib, hotpath, synth
n = 1 << 20
amp, guide, cut, hdr = synth.make_case(1234, 250)
buf, off = synth.make_reads_fast(amp, hdr, cut, n, seed=1234, read_len=250)
inc = hotpath.include_mask(250, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
flags = hotpath.quant_flags(hdr)
ctx = Context(0)
d_buf = torch.from_numpy(buf).cuda(); d_off = torch.from_numpy(off).cuda()
out = {"kept": torch.zeros(n, dtype=torch.uint8, device="cuda"), "aln": torch.zeros(n * 32, dtype=torch.uint8, device="cuda"),
       "recs": torch.zeros(n * 16, dtype=torch.uint8, device="cuda"), "tenths_rep": torch.zeros(n, dtype=torch.int32, device="cuda")}
ptrs = {k: v.data_ptr() for k, v in out.items()}
torch.cuda.synchronize()
for alleles in (0, 0, 0, 4096, 4096, 4096, 50000):
    t0 = time.time()
    res = hotpath.run_hot_path(ctx, amp, None, hdr_amplicon=hdr, flags=flags, inc=inc, device_inputs=(d_buf.data_ptr(), d_off.data_ptr(), n, 250, ptrs), alleles=alleles)
    dt = time.time() - t0
    ms, ln = ctx.last_timing()
    print("alleles=%d: %.1f ms  other %.2f ms  allele_n=%s top=%s" % (alleles, dt * 1e3, ms["other"], res.allele_n, None if not alleles else res.allele_count[:4].tolist()))
