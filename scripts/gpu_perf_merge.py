"""Throughput of crgpu_flash_merge on one B200: N pairs of 2 x L bp sampled from a pool of distinct
synthetic pairs, inputs resident in HBM (CRGPU_MEM_DEVICE).  usage: gpu_perf_merge.py [N] [L]"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crispresso_b200 import Context, _lib, synth  # noqa: E402
from crispresso_b200.aligner import pack_reads  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
amp, _g, _c, _h = synth.make_case(5, 250, hdr=False)
pool = 4096
s1, q1, s2, q2 = synth.make_pairs(amp, pool, L, seed=5)
rng = np.random.default_rng(1)
pick = rng.integers(0, pool, size=N)
arrs = []
for lst in (s1, q1, s2, q2):
    lens = np.array([len(x) for x in lst])[pick]
    off = np.zeros(N + 1, np.int64); off[1:] = np.cumsum(lens)
    buf, poff = pack_reads(lst)
    idx = np.repeat(poff[:-1][pick] - off[:-1], lens) + np.arange(int(off[-1]))
    arrs.append((buf[idx], off))
ctx = Context(0)
d = [torch.from_numpy(a).cuda() for a in (arrs[0][0], arrs[1][0], arrs[0][1], arrs[2][0], arrs[3][0], arrs[2][1])]
cap = int(arrs[0][1][-1] + arrs[2][1][-1])
o_pos = torch.zeros(N, dtype=torch.int32, device="cuda"); o_kind = torch.zeros(N, dtype=torch.uint8, device="cuda")
o_seq = torch.zeros(cap, dtype=torch.uint8, device="cuda"); o_qual = torch.zeros(cap, dtype=torch.uint8, device="cuda")
o_off = torch.zeros(N + 1, dtype=torch.int64, device="cuda"); o_idx = torch.zeros(N, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
prm = _lib.MergeParams(4, 100, 0.25, 1)
out = _lib.MergeOut()
out.pos, out.kind, out.seq, out.qual, out.offsets, out.index = (t.data_ptr() for t in (o_pos, o_kind, o_seq, o_qual, o_off, o_idx))
out.cap_bytes, out.cap_reads = cap, N
best = 1e9
for it in range(5):
    t0 = time.time()
    ctx.check(ctx.lib.crgpu_flash_merge(ctx.handle, _lib.MEM_DEVICE, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(),
                                        d[3].data_ptr(), d[4].data_ptr(), d[5].data_ptr(), N, ctypes.byref(prm), ctypes.byref(out)))
    wall = time.time() - t0
    ms, ln = ctx.last_timing()
    best = min(best, ms["other"])
    print("iter %d: kernels %.3f ms (%d launches), wall %.3f ms, merged %d of %d" % (it, ms["other"], ln["other"], wall * 1e3, out.n_merged, N))
in_bytes = 2 * cap
print("pairs/s %.1f M; input %.1f MB + output %.1f MB -> %.1f GB/s of algorithmic traffic" % (
    N / best / 1e3, in_bytes / 1e6, 2 * out.bytes / 1e6, (in_bytes + 2 * out.bytes) / best / 1e6))
