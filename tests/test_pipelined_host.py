"""Host logic of hotpath.run_hot_path_pipelined without a GPU: a stand-in for the C entry point computes per-read
records and reductions that depend only on each read's bytes, so chunking / pointer offsets / the RC-list compaction /
the reduction merge can be checked against one call over all reads."""
import ctypes
import zlib

import numpy as np

from crispresso_b200 import _lib, hotpath


class _FakeLib:
    """crgpu_align_quantify(CRGPU_MEM_HOST) whose outputs are a pure function of every read's bytes."""

    def __init__(self):
        self.calls = []
        self.staged = {}

    def crgpu_set_deferred_outputs(self, handle, on):
        return 0

    def crgpu_stage_reads(self, handle, slot, fmt, reads_addr, offs_addr, n):
        """Copies the batch (as the library would, onto the device) so that a caller that drops its buffers early is caught."""
        offs = np.ctypeslib.as_array(ctypes.cast(offs_addr, ctypes.POINTER(ctypes.c_int64)), (n + 1,)).copy()
        assert offs[0] == 0
        total = int(offs[n])
        nbytes = (total + 1) // 2 if fmt == _lib.READS_BAM4 else total
        raw = np.ctypeslib.as_array(ctypes.cast(reads_addr, ctypes.POINTER(ctypes.c_uint8)), (nbytes,)).copy()
        if fmt == _lib.READS_BAM4:
            lut = np.frombuffer(b"=ACMGRSVTWYHKDBN", np.uint8)
            both = np.empty(2 * nbytes, np.uint8)
            both[0::2], both[1::2] = lut[raw >> 4], lut[raw & 15]
            raw = both[:total].copy()
        assert slot in (0, 1) and slot not in self.staged, "slot staged twice without a run in between"
        self.staged[slot] = (raw, offs, n)
        return 0

    def crgpu_align_quantify_staged(self, handle, slot, amp, L, pp, qp, po):
        raw, offs, n = self.staged.pop(slot)
        self._keep = (raw, offs)
        return self.crgpu_align_quantify(handle, _lib.MEM_HOST, amp, L, pp, qp, raw.ctypes.data, offs.ctypes.data, n, po)

    def crgpu_align_quantify(self, handle, mem, amp, L, pp, qp, reads_addr, offs_addr, n, po):
        assert mem == _lib.MEM_HOST
        po = po._obj
        offs = np.ctypeslib.as_array(ctypes.cast(offs_addr, ctypes.POINTER(ctypes.c_int64)), (n + 1,))
        assert offs[0] == 0                                       # chunk-relative offsets
        buf = np.ctypeslib.as_array(ctypes.cast(reads_addr, ctypes.POINTER(ctypes.c_uint8)), (int(offs[n]),))
        self.calls.append((handle, int(n)))

        def view(addr, dtype, count):
            raw = (ctypes.c_uint8 * (count * np.dtype(dtype).itemsize)).from_address(addr)
            return np.frombuffer(raw, dtype=dtype, count=count)

        kept, aln, recs = view(po.kept, np.uint8, n), view(po.aln, _lib.ALN_REC, n), view(po.recs, _lib.READ_REC, n)
        trep = view(po.tenths_rep, np.int32, n)
        rc_read, rc_aln = view(po.rc_read, np.int32, po.rc_cap), view(po.rc_aln, _lib.ALN_REC, po.rc_cap)
        rc_recs = view(po.rc_recs, _lib.READ_REC, po.rc_cap)
        vectors = view(po.vectors, np.int64, _lib.NUM_VECTORS * L).reshape(_lib.NUM_VECTORS, L)
        hist = view(po.hist_inframe, np.int64, po.hist_len)
        counters = view(po.counters, np.int64, _lib.NUM_COUNTERS)
        nrc = 0
        for i in range(n):
            seq = bytes(buf[offs[i]:offs[i + 1]])
            h = zlib.crc32(seq)
            kept[i] = h & 1
            aln[i]["alnlen"], aln[i]["ident"], aln[i]["read_len"] = len(seq) + (h & 7), h % 251, len(seq)
            aln[i]["score"], aln[i]["tenths"] = float(h % 1000), h % 1001
            trep[i] = h % 997
            recs[i]["cls"], recs[i]["n_mutated"] = h & 3, (h >> 4) & 15
            vectors[h % _lib.NUM_VECTORS, h % L] += 1
            hist[h % po.hist_len] += 1
            counters[h % _lib.NUM_COUNTERS] += 1
            po.class_counts[h & 3] += 1
            if h % 5 == 0:
                rc_read[nrc] = i
                rc_aln[nrc]["alnlen"], rc_recs[nrc]["n_deleted"] = len(seq), h % 17
                nrc += 1
        if po.allele_cap:
            # alleles of the stand-in: kept rows grouped by crc32 % 13, keys = (group, group * 7 + 1)
            groups = {}
            for i in range(n):
                if kept[i]:
                    groups.setdefault(zlib.crc32(bytes(buf[offs[i]:offs[i + 1]])) % 13, []).append(i)
            table = sorted(groups.items(), key=lambda kv: -len(kv[1]))
            arow, acnt = view(po.allele_row, np.int32, po.allele_cap), view(po.allele_count, np.int64, po.allele_cap)
            akey = view(po.allele_key, np.uint64, 2 * po.allele_cap)
            for k, (g, members) in enumerate(table[:po.allele_cap]):
                arow[k], acnt[k], akey[2 * k], akey[2 * k + 1] = members[0], len(members), g, g * 7 + 1
            po.allele_n = len(table)
        po.rc_n, po.n_total, po.n_cells = nrc, po.n_total + n, po.n_cells + int(offs[n]) * L
        po.n_cells_computed += int(offs[n]) * L
        return 0


class _FakeCtx:
    def __init__(self, lib, handle):
        self.lib, self.handle = lib, handle

    def sync(self):
        pass

    def check(self, rc):
        assert rc == 0


def _reads(n, seed):
    rng = np.random.default_rng(seed)
    lens = rng.integers(5, 40, n)
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum(lens)
    buf = rng.choice(np.frombuffer(b"ACGT", np.uint8), int(off[-1])).astype(np.uint8)
    return buf, off


def test_chunks_on_several_contexts_equal_one_call():
    amp = "ACGT" * 10
    reads = _reads(1000, 3)
    lib1 = _FakeLib()
    one = hotpath.run_hot_path_pipelined([_FakeCtx(lib1, 1)], amp, reads, chunk_reads=1 << 20)
    assert lib1.calls == [(1, 1000)]
    for nctx, chunk in ((2, 170), (3, 64), (1, 333), (2, 1000)):
        lib = _FakeLib()
        got = hotpath.run_hot_path_pipelined([_FakeCtx(lib, h) for h in range(nctx)], amp, reads, chunk_reads=chunk)
        assert sum(c[1] for c in lib.calls) == 1000 and len(lib.calls) == -(-1000 // chunk)
        assert {c[0] for c in lib.calls} == set(range(min(nctx, len(lib.calls))))          # every context got work
        for f in ("kept", "tenths_rep", "rc_read"):
            assert np.array_equal(getattr(got, f), getattr(one, f)), f
        for f in ("aln", "recs", "rc_aln", "rc_recs"):
            assert getattr(got, f).tobytes() == getattr(one, f).tobytes(), f
        assert np.array_equal(got.red.flat(), one.red.flat())
    assert len(one.rc_read) > 100 and np.all(np.diff(one.rc_read) > 0)                    # RC rows in read order


def test_preallocated_outputs_are_used_in_place():
    amp = "ACGT" * 10
    reads = _reads(300, 4)
    n = 300
    out = {"kept": np.zeros(n, np.uint8), "aln": np.zeros(n, _lib.ALN_REC), "recs": np.zeros(n, _lib.READ_REC),
           "tenths_rep": np.zeros(n, np.int32), "rc_read": np.zeros(n, np.int32), "rc_aln": np.zeros(n, _lib.ALN_REC),
           "rc_recs": np.zeros(n, _lib.READ_REC), "offsets": np.zeros(n + 3, np.int64)}
    lib = _FakeLib()
    red = hotpath.Reductions(len(amp))
    got = hotpath.run_hot_path_pipelined([_FakeCtx(lib, 0), _FakeCtx(lib, 1)], amp, reads, chunk_reads=100, red=red, out=out)
    assert got.kept is out["kept"] and got.aln is out["aln"] and got.red is red
    assert red.n_total == n and int(red.class_counts.sum()) == n


def test_staged_pipeline_equals_one_call_bytes_and_packed():
    """hotpath.run_hot_path_staged (one context, the next chunk staged while the current one runs): chunk slicing, the
    even-offset rule of the 4-bit format, RC-list compaction, reductions and the allele tables merged by key."""
    amp = "ACGT" * 10
    reads = _reads(1000, 5)
    lib1 = _FakeLib()
    one = hotpath.run_hot_path_staged(_FakeCtx(lib1, 1), amp, reads, chunk_reads=1 << 20, alleles=64)
    assert lib1.calls == [(1, 1000)] and one.allele_n == 13 and int(one.allele_count.sum()) == int((one.kept & 1).sum())
    packed = hotpath.pack_bam4(reads[0])
    for chunk, pk in ((170, None), (64, None), (333, packed), (50, packed), (1000, packed)):
        lib = _FakeLib()
        got = hotpath.run_hot_path_staged(_FakeCtx(lib, 0), amp, reads, chunk_reads=chunk, packed=pk, alleles=64)
        assert sum(c[1] for c in lib.calls) == 1000 and not lib.staged
        for f in ("kept", "tenths_rep", "rc_read"):
            assert np.array_equal(getattr(got, f), getattr(one, f)), f
        for f in ("aln", "recs", "rc_aln", "rc_recs"):
            assert getattr(got, f).tobytes() == getattr(one, f).tobytes(), f
        assert np.array_equal(got.red.flat(), one.red.flat())
        assert sorted(got.allele_count.tolist()) == sorted(one.allele_count.tolist()) and got.allele_n == one.allele_n
        # a representative row belongs to its allele: same group key as the single call's representative with that count
        def group(res, row):
            b, o = reads
            return zlib.crc32(bytes(b[o[row]:o[row + 1]])) % 13
        assert {group(got, int(r)): int(c) for r, c in zip(got.allele_row, got.allele_count)} == \
               {group(one, int(r)): int(c) for r, c in zip(one.allele_row, one.allele_count)}
