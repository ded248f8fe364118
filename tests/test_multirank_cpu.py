"""The N>1 path on the CPU: read sharding + the all-reduce of the histogram block, with
torch.distributed (gloo, world_size 2).  The per-rank 'compute' is the oracle (this is a test of the
host-side reduction logic, not of the kernels)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crispresso_b200 import hotpath, synth
from oracle import quantify


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _to_reductions(res, L):
    red = hotpath.Reductions(L)
    for k, name in enumerate(hotpath.VECTOR_NAMES):
        red.vectors[k] = res["vectors"][name]
    for key, v in res["hist_inframe"].items():
        red.hist_inframe[key + hotpath.HIST_ZERO] = v
    for key, v in res["hist_frameshift"].items():
        red.hist_frameshift[key + hotpath.HIST_ZERO] = v
    red.counters[:] = [res["counters"][n] for n in hotpath.COUNTER_NAMES]
    red.class_counts[:] = [res["classes"][k] for k in ("UNMODIFIED", "NHEJ", "HDR", "MIXED")]
    red.n_total, red.n_cells = res["n_total"], res["n_cells"]
    return red


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    amp, guide, cut, hdr = synth.make_case(21, 120)
    buf, off = synth.make_reads(amp, hdr, cut, 400, seed=21)
    from crispresso_b200 import distributed
    sbuf, soff = distributed.shard_reads(buf, off, rank, world)     # contiguous, equal-count shards
    inc = hotpath.include_mask(120, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    res = quantify.hot_path(amp, (sbuf.copy(), soff.copy()), hdr_amplicon=hdr,
                            opts=quantify.Opts(expected_hdr_amplicon_seq=hdr, coding_seq=amp[40:80]),
                            include=np.nonzero(inc)[0], exon=range(40, 80), splice=[38, 39, 80, 81], nthreads=1)
    red = distributed.allreduce_reductions(_to_reductions(res, 120))
    if rank == 0:
        np.save(out, red.flat())
    dist.destroy_process_group()


def test_sharded_reduction_equals_single_rank(tmp_path):
    out = str(tmp_path / "flat.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    amp, guide, cut, hdr = synth.make_case(21, 120)
    packed = synth.make_reads(amp, hdr, cut, 400, seed=21)
    inc = hotpath.include_mask(120, hotpath.cut_points_from_guides(amp, guide), 1, 15, 15)
    whole = quantify.hot_path(amp, packed, hdr_amplicon=hdr, opts=quantify.Opts(expected_hdr_amplicon_seq=hdr, coding_seq=amp[40:80]),
                              include=np.nonzero(inc)[0], exon=range(40, 80), splice=[38, 39, 80, 81], nthreads=1)
    assert np.array_equal(np.load(out), _to_reductions(whole, 120).flat())
