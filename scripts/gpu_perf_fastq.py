"""Throughput of crgpu_fastq_index on one B200 with the text resident in HBM.  usage: gpu_perf_fastq.py [N] [L]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crispresso_b200 import Context, _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
L = int(sys.argv[2]) if len(sys.argv) > 2 else 250
rng = np.random.default_rng(1)
rec = np.zeros((N, 49 + 2 * L), np.uint8)              # "@" + 43-char name + \n + L + \n + "+" + \n + L + \n
name = np.frombuffer(b"@M06879:15:000000000-DFF22:1:1101:25894:23776 ", np.uint8)
rec[:, :len(name)] = name
rec[:, len(name):44] = ord("x")
rec[:, 44] = 10
rec[:, 45:45 + L] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=(N, L))]
rec[:, 45 + L] = 10; rec[:, 46 + L] = ord("+"); rec[:, 47 + L] = 10
rec[:, 48 + L:48 + 2 * L] = rng.integers(35, 74, size=(N, L))
rec[:, 48 + 2 * L] = 10
text = rec.reshape(-1)
ctx = Context(0)
d_text = torch.from_numpy(text).cuda()
total = N * L
d_seq = torch.zeros(total, dtype=torch.uint8, device="cuda"); d_qual = torch.zeros(total, dtype=torch.uint8, device="cuda")
d_off = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
d_ns = torch.zeros(N, dtype=torch.int64, device="cuda"); d_nl = torch.zeros(N, dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
fo = _lib.FastqOut()
fo.cap_records, fo.cap_bytes = N, total
fo.seq, fo.qual, fo.offsets, fo.name_start, fo.name_len = (t.data_ptr() for t in (d_seq, d_qual, d_off, d_ns, d_nl))
best = 1e9
for it in range(5):
    ctx.check(ctx.lib.crgpu_fastq_index(ctx.handle, _lib.MEM_DEVICE, d_text.data_ptr(), len(text), 1, ctypes.byref(fo)))
    ms, ln = ctx.last_timing()
    best = min(best, ms["other"])
    print("iter %d: kernels %.3f ms (%d launches), records %d, bases %d" % (it, ms["other"], ln["other"], fo.n_records, fo.seq_bytes))
B = len(text)
alg = 2 * B + 4 * total + 72 * N
print("text %.1f MB: %.1f M records/s, %.1f GB/s of text, algorithmic traffic %.1f MB -> %.1f GB/s" % (
    B / 1e6, N / best / 1e3, B / best / 1e6, alg / 1e6, alg / best / 1e6))
