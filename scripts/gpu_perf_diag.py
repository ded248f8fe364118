"""Fused hot path with the diagonal shortcut on / off (CRGPU_NO_OVERLAP=1 recommended).  usage: gpu_perf_diag.py [La] [n]"""
import sys
sys.path.insert(0, ".")
import time
import numpy as np, torch
from crispresso_b200 import Context, _lib, hotpath, synth
args = [a for a in sys.argv[1:] if not a.startswith("--")]
La = int(args[0]) if len(args) > 0 else 250
n = int(args[1]) if len(args) > 1 else 1 << 20
amp, guide, cut, hdr = synth.make_case(1234, La)
buf, off = synth.make_reads_fast(amp, hdr, cut, n, seed=1234, read_len=La)
inc = hotpath.include_mask(La, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
flags = hotpath.quant_flags(hdr)
ctx = Context(0)
for a_ in sys.argv:
    if a_.startswith("--budget="):
        ctx.set_traceback_budget(int(float(a_.split("=")[1]) * (1 << 30)))
d_buf = torch.from_numpy(buf).cuda(); d_off = torch.from_numpy(off).cuda()
out = {"kept": torch.zeros(n, dtype=torch.uint8, device="cuda"), "aln": torch.zeros(n * 32, dtype=torch.uint8, device="cuda"),
       "recs": torch.zeros(n * 16, dtype=torch.uint8, device="cuda"), "tenths_rep": torch.zeros(n, dtype=torch.int32, device="cuda")}
ptrs = {k: v.data_ptr() for k, v in out.items()}
torch.cuda.synchronize()
for hdr_on in ((True,) if "--hdr-only" in sys.argv else (True, False)):
    for diag in ((True,) if "--diag-only" in sys.argv else (False, True) if "--diag" in sys.argv else (False,)):
        ctx.set_diag_shortcut(diag)
        best = None
        wall = 1e9
        for it in range(1 if "--once" in sys.argv else 4):
            t0 = time.perf_counter()
            hotpath.run_hot_path(ctx, amp, None, hdr_amplicon=hdr if hdr_on else None, flags=flags if hdr_on else hotpath.quant_flags(""), inc=inc,
                                 device_inputs=(d_buf.data_ptr(), d_off.data_ptr(), n, La, ptrs))
            wall = min(wall, (time.perf_counter() - t0) * 1e3)
            ms, ln = ctx.last_timing()
            tot = sum(ms.values())
            best = ms if best is None or tot < sum(best.values()) else best
        print("La=%d n=%d hdr=%s diag=%s: %s total %.2f ms, wall %.2f ms; pairs total/left %s escaped %s" % (
            La, n, hdr_on, diag, {k: round(v, 2) for k, v in best.items()}, sum(best.values()), wall, ctx.last_diag(), ctx.last_escaped()))
