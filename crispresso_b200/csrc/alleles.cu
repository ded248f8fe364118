// alleles.cu -- device side of the allele table.
//
// The reference groups df_needle_alignment by (align_seq, ref_seq, NHEJ, UNMODIFIED, HDR, n_deleted,
// n_inserted, n_mutated) and counts the rows (CRISPResso/CRISPRessoCORE.py:2923-2946).  For a fixed
// amplicon that key is a function of the read's bases as they appear in align_seq, the alignment's
// gap / mismatch pattern (the walker's 2-bit ops) and the four record fields, so rows are hashed over
// exactly those, radix-sorted (cub), run-length encoded and ordered by count.  Only one
// representative row per allele has to be materialised as text afterwards.
// A second, independent hash guards the grouping: rows that collide in the 64-bit sort key but differ
// in the check hash make the call fail instead of merging two alleles.
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "crgpu_common.cuh"
#include "../../include/crgpu.h"

namespace crgpu {

// murmur3's 64-bit finaliser: every (position, content) element of a row is mixed on its own and the row's hash is the SUM of
// its elements' mixes, so the 32 lanes of a warp hash one row together, in any order
__device__ __forceinline__ uint64_t fmix64(uint64_t x)
{
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

__device__ __forceinline__ uint8_t comp_up(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': return 'T';
    case 'C': case 'c': return 'G';
    case 'G': case 'g': return 'C';
    case 'T': case 't': case 'U': case 'u': return 'A';
    case 'N': case 'n': return 'N';
    default: return c;
    }
}

constexpr uint64_t SEED1 = 0x9e3779b97f4a7c15ull, SEED2 = 0xc2b2ae3d27d4eb4full;

// does any read of the batch hold a lower-case letter?  (align_seq keeps the read's case, CORE:141-144 upper-cases RC rows
// only: without one -- the usual FASTQ -- every forward row of identity 100.0 is the same allele and needs no look at its bases)
__global__ void k_any_lower(const uint8_t *__restrict__ reads, const int64_t *__restrict__ offsets, int64_t n, int *flag)
{
    const uintptr_t p0 = reinterpret_cast<uintptr_t>(reads) + (uintptr_t)offsets[0];
    const uintptr_t p1 = reinterpret_cast<uintptr_t>(reads) + (uintptr_t)offsets[n];
    const uintptr_t a0 = p0 & ~(uintptr_t)3;
    const int64_t nwords = (int64_t)(((p1 + 3) & ~(uintptr_t)3) - a0) >> 2;
    uint32_t acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t w = reinterpret_cast<const uint32_t *>(a0)[i];
        const uintptr_t wa = a0 + (uintptr_t)i * 4;
        if (wa < p0) w &= 0xffffffffu << (8 * (unsigned)(p0 - wa));
        if (wa + 4 > p1) w &= 0xffffffffu >> (8 * (unsigned)(wa + 4 - p1));
        acc |= w;
    }
    // letters only carry bit 5 as "lower case" (digits / punctuation have it too: they send the batch the long way, no more)
    if (__any_sync(0xffffffffu, (acc & 0x20202020u) != 0) && (threadIdx.x & 31) == 0) *flag = 1;
}

// rows [0, n): forward row of read i (valid iff kept[i] & 1); rows [n, n + nrc): RC row j of read
// rc_read[j] (valid iff kept[rc_read[j]] & 2).  Invalid rows get key = ~0 and sort last.
// A CTA takes 256 rows.  First one thread per row: rows that are not kept, and rows of identity 100.0 -- their three strings
// are the amplicon itself (CORE:2014), whatever the strand: one allele, by far the most frequent one, with a constant key --
// are done at once (a forward row only when the batch has no lower-case letter: a read with one is another allele and
// takes the long way).  The other rows are gathered and hashed by eight lanes each, four rows per warp.
// Elements of a row: the bases as they appear in align_seq, four per element, keyed by their position (raw for forward rows,
// upper-cased reverse complement for RC rows); the gap columns (op 2 / 3), keyed by their FORWARD column -- the walker stores
// forward rows' ops last column first -- so that a forward row and an RC row with the same text rows agree (matches and
// mismatches follow from the bases and the gaps); the alignment length, the read length and the four record fields.
constexpr int HASH_THREADS = 256;
__global__ void __launch_bounds__(HASH_THREADS) k_hash_rows(const uint8_t *__restrict__ reads, const int64_t *__restrict__ offsets, int64_t n,
                            const uint8_t *__restrict__ kept, const int32_t *__restrict__ rc_read, int64_t nrc,
                            const uint32_t *__restrict__ ops_fw, const uint32_t *__restrict__ ops_rc, int64_t ops_stride,
                            const crgpu_aln_rec *__restrict__ aln_fw, const crgpu_aln_rec *__restrict__ aln_rc,
                            const crgpu_read_rec *__restrict__ rec_fw, const crgpu_read_rec *__restrict__ rec_rc,
                            const int *__restrict__ any_lower,
                            uint64_t *__restrict__ keys, uint64_t *__restrict__ chk, int32_t *__restrict__ rows)
{
    __shared__ int heavy[HASH_THREADS];
    __shared__ int nheavy;
    if (threadIdx.x == 0) nheavy = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * HASH_THREADS;
    {
        const int64_t r = base + threadIdx.x;
        if (r < n + nrc) {
            rows[r] = (int32_t)r;
            const bool rc = r >= n;
            const int64_t read = rc ? rc_read[r - n] : r;
            const uint8_t kp = kept[read];
            const int tenths = rc ? aln_rc[r - n].tenths : aln_fw[read].tenths;
            const bool valid = rc ? (kp & 2) != 0 : (kp & 1) != 0;
            if (!valid) { keys[r] = ~0ull; chk[r] = 0; }
            else if (tenths == 1000 && (rc || !*any_lower)) { keys[r] = 0x243f6a8885a308d3ull; chk[r] = 0x13198a2e03707344ull; }
            else heavy[atomicAdd(&nheavy, 1)] = threadIdx.x;
        }
    }
    __syncthreads();
    const int sub = threadIdx.x & 7;
    for (int k0 = 0; k0 < nheavy; k0 += HASH_THREADS / 8) {           // (uniform trip count: the shuffles below see whole warps)
        const int k = k0 + (threadIdx.x >> 3);
        const bool in = k < nheavy;
        const int64_t r = base + (in ? heavy[k] : 0);
        const bool rc = in && r >= n;
        const int64_t read = !in ? 0 : rc ? rc_read[r - n] : r;
        const int64_t o0 = offsets[read], o1 = offsets[read + 1];
        const int tenths = rc ? aln_rc[r - n].tenths : aln_fw[read].tenths;
        const int ncol = rc ? aln_rc[r - n].alnlen : aln_fw[read].alnlen;
        const uint8_t *b = reads + o0;
        const int len = (int)(o1 - o0);
        // a forward row of identity 100.0 in a batch with lower-case letters: constant key unless THIS read has one
        uint32_t any = 0;
        const bool t1000 = in && tenths == 1000;
        if (t1000) for (int i = sub * 4; i < len; i += 32) any |= load4(b, i, len);
        any |= __shfl_xor_sync(0xffffffffu, any, 1);
        any |= __shfl_xor_sync(0xffffffffu, any, 2);
        any |= __shfl_xor_sync(0xffffffffu, any, 4);
        const bool fast = t1000 && !(any & 0x20202020u);
        uint64_t h1 = 0, h2 = 0;
        if (in && !fast) {
            const int nwords = (len + 3) >> 2;
            for (int w = sub; w < nwords; w += 8) {
                uint32_t c4 = 0;
                if (!rc) c4 = load4(b, 4 * w, len);
                else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int i = 4 * w + q;
                        if (i < len) c4 |= (uint32_t)comp_up(b[len - 1 - i]) << (8 * q);
                    }
                }
                const uint64_t x = (uint64_t)c4 | ((uint64_t)(w + 1) << 32);
                h1 += fmix64(x ^ SEED1);
                h2 += fmix64((x + SEED2) * SEED1);
            }
            const uint32_t *ops = rc ? ops_rc + (r - n) * ops_stride : ops_fw + read * ops_stride;
            const int nopw = (ncol + 15) >> 4;
            for (int w = sub; w < nopw; w += 8) {
                const uint32_t ow = ops[w];
                uint32_t g = ow & 0xaaaaaaaau;                              // ops 2 and 3
                while (g) {
                    const int bit = __ffs(g) - 1;
                    g &= g - 1;
                    const int j = w * 16 + (bit >> 1);
                    if (j >= ncol) break;
                    const int col = rc ? j : ncol - 1 - j;
                    const uint64_t x = (uint64_t)(uint32_t)col | ((uint64_t)((ow >> (bit - 1)) & 3u) << 32) | (1ull << 62);
                    h1 += fmix64(x ^ SEED1);
                    h2 += fmix64((x + SEED2) * SEED1);
                }
            }
        }
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            h1 += __shfl_xor_sync(0xffffffffu, h1, d);
            h2 += __shfl_xor_sync(0xffffffffu, h2, d);
        }
        if (!in || sub != 0) continue;
        if (fast) { keys[r] = 0x243f6a8885a308d3ull; chk[r] = 0x13198a2e03707344ull; continue; }
        const crgpu_read_rec q = rc ? rec_rc[r - n] : rec_fw[read];
        const uint64_t f = (uint64_t)q.cls | ((uint64_t)(uint32_t)q.n_mutated << 8) | ((uint64_t)(uint32_t)q.n_inserted << 24) |
                           ((uint64_t)(uint32_t)q.n_deleted << 44);
        const uint64_t g = (uint64_t)(uint32_t)ncol | ((uint64_t)(uint32_t)len << 32);
        h1 = fmix64(h1 + fmix64(f ^ SEED2) + fmix64(g + SEED1));
        h2 = fmix64(h2 ^ fmix64(f + SEED1) ^ fmix64(g * SEED2 + 1));
        if (h1 == ~0ull) h1 = 0x5bd1e995u;
        keys[r] = h1;
        chk[r] = h2;
    }
}

// sorted order: rows with equal key must carry equal check hashes
__global__ void k_check_groups(const uint64_t *__restrict__ skeys, const int32_t *__restrict__ srows,
                               const uint64_t *__restrict__ chk, int64_t m, int *err)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < 1 || p >= m) return;
    if (skeys[p] != ~0ull && skeys[p] == skeys[p - 1] && chk[srows[p]] != chk[srows[p - 1]]) atomicOr(err, 1);
}

// representative (first sorted row) of every run; the run of invalid rows gets count 0; slots past
// the last run are neutral (count 0 from the memset, representative -1)
__global__ void k_group_reps(const uint64_t *__restrict__ ukeys, const int32_t *__restrict__ starts, const int32_t *__restrict__ srows,
                             int32_t *__restrict__ counts, int *__restrict__ nruns, int32_t *__restrict__ rep,
                             int32_t *__restrict__ gid, int64_t m)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= m) return;
    gid[g] = (int32_t)g;
    if (g >= *nruns) { rep[g] = -1; return; }
    rep[g] = srows[starts[g]];
    if (ukeys[g] == ~0ull) { counts[g] = 0; nruns[2] = 1; }      // (nruns[2]: one of the runs is the rows that were not kept)
}

__global__ void k_gather_i32(const int32_t *__restrict__ src, const int32_t *__restrict__ idx, int32_t *__restrict__ out, int64_t m)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) out[i] = src[idx[i]];
}

// the two hashes of the k-th most frequent allele: what a caller that quantifies a read set in several calls merges the
// per-call allele tables by
__global__ void k_allele_keys(const uint64_t *__restrict__ ukeys, const uint64_t *__restrict__ chk, const int32_t *__restrict__ rep,
                              const int32_t *__restrict__ sgid, const int *__restrict__ nruns, uint64_t *__restrict__ out, int64_t m)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const int g = sgid[k];
    const bool ok = g < *nruns && rep[g] >= 0;
    out[2 * k] = ok ? ukeys[g] : 0;
    out[2 * k + 1] = ok ? chk[rep[g]] : 0;
}

// Group the rows; returns the number of alleles (runs with a non-zero count) and, sorted by count
// descending, up to `cap` (representative row, count) pairs in host arrays.
// tmp: caller-provided scratch allocator callback is avoided: all scratch comes in through `scratch`
// (bytes >= allele_scratch_bytes(m)).
size_t allele_scratch_bytes(int64_t m)
{
    size_t a = 0, b = 0, c = 0, d = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (uint64_t *)nullptr, (uint64_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, (int)m);
    cub::DeviceRunLengthEncode::Encode(nullptr, b, (uint64_t *)nullptr, (uint64_t *)nullptr, (int32_t *)nullptr, (int *)nullptr, (int)m);
    cub::DeviceScan::ExclusiveSum(nullptr, c, (int32_t *)nullptr, (int32_t *)nullptr, (int)m);
    cub::DeviceRadixSort::SortPairsDescending(nullptr, d, (int32_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, (int)m);
    size_t t = a;
    if (b > t) t = b;
    if (c > t) t = c;
    if (d > t) t = d;
    t = (t + 255) & ~(size_t)255;
    // keys, chk, skeys, ukeys (8 B) + rows, srows, counts, starts, rep, gid, scounts, sgid (4 B) + key pairs (16 B) + nruns/err
    return t + (size_t)m * (4 * 8 + 8 * 4 + 16) + 1024;
}

cudaError_t allele_groups(const uint8_t *reads, const int64_t *offsets, int64_t n, const uint8_t *kept, const int32_t *rc_read,
                          int64_t nrc, const uint32_t *ops_fw, const uint32_t *ops_rc, int64_t ops_stride,
                          const crgpu_aln_rec *aln_fw, const crgpu_aln_rec *aln_rc, const crgpu_read_rec *rec_fw,
                          const crgpu_read_rec *rec_rc, int reads_upper, void *scratch, size_t scratch_bytes, cudaStream_t s,
                          int32_t **d_rep_sorted, int32_t **d_count_sorted, int **d_nruns, int **d_err, uint64_t **d_key_pairs)
{
    const int64_t m = n + nrc;
    size_t tmp_bytes = scratch_bytes - ((size_t)m * (4 * 8 + 8 * 4 + 16) + 1024);
    uint8_t *p = reinterpret_cast<uint8_t *>(scratch);
    void *tmp = p; p += tmp_bytes;
    uint64_t *keys = reinterpret_cast<uint64_t *>(p); p += m * 8;
    uint64_t *chk = reinterpret_cast<uint64_t *>(p); p += m * 8;
    uint64_t *skeys = reinterpret_cast<uint64_t *>(p); p += m * 8;
    uint64_t *ukeys = reinterpret_cast<uint64_t *>(p); p += m * 8;
    int32_t *rows = reinterpret_cast<int32_t *>(p); p += m * 4;
    int32_t *srows = reinterpret_cast<int32_t *>(p); p += m * 4;
    int32_t *counts = reinterpret_cast<int32_t *>(p); p += m * 4;
    int32_t *starts = reinterpret_cast<int32_t *>(p); p += m * 4;
    int32_t *rep = reinterpret_cast<int32_t *>(p); p += m * 4;
    int32_t *gid = reinterpret_cast<int32_t *>(p); p += m * 4;
    int32_t *scounts = reinterpret_cast<int32_t *>(p); p += m * 4;
    int32_t *sgid = reinterpret_cast<int32_t *>(p); p += m * 4;
    uint64_t *kpairs = reinterpret_cast<uint64_t *>(p); p += m * 16;
    int *nruns = reinterpret_cast<int *>(p); int *err = nruns + 1;
    cudaError_t e;
    if ((e = cudaMemsetAsync(nruns, 0, 16, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(counts, 0, (size_t)m * 4, s)) != cudaSuccess) return e;
    const unsigned grid = (unsigned)((m + 127) / 128);
    // nruns[3]: "some read of the batch has a lower-case letter" (reads_upper: the caller knows there is none)
    if (!reads_upper) k_any_lower<<<1184, 256, 0, s>>>(reads, offsets, n, nruns + 3);
    k_hash_rows<<<(unsigned)((m + HASH_THREADS - 1) / HASH_THREADS), HASH_THREADS, 0, s>>>(reads, offsets, n, kept, rc_read, nrc, ops_fw, ops_rc,
                                     ops_stride, aln_fw, aln_rc, rec_fw, rec_rc, nruns + 3, keys, chk, rows);
    size_t tb = tmp_bytes;
    if ((e = cub::DeviceRadixSort::SortPairs(tmp, tb, keys, skeys, rows, srows, (int)m, 0, 64, s)) != cudaSuccess) return e;
    k_check_groups<<<grid, 128, 0, s>>>(skeys, srows, chk, m, err);
    tb = tmp_bytes;
    if ((e = cub::DeviceRunLengthEncode::Encode(tmp, tb, skeys, ukeys, counts, nruns, (int)m, s)) != cudaSuccess) return e;
    tb = tmp_bytes;
    if ((e = cub::DeviceScan::ExclusiveSum(tmp, tb, counts, starts, (int)m, s)) != cudaSuccess) return e;
    k_group_reps<<<grid, 128, 0, s>>>(ukeys, starts, srows, counts, nruns, rep, gid, m);
    tb = tmp_bytes;
    // all m slots are sorted (slots past the number of runs hold count 0 and sort last)
    if ((e = cub::DeviceRadixSort::SortPairsDescending(tmp, tb, counts, scounts, gid, sgid, (int)m, 0, 32, s)) != cudaSuccess) return e;
    // representatives in the sorted order (the exclusive-sum buffer is free again)
    k_gather_i32<<<grid, 128, 0, s>>>(rep, sgid, starts, m);
    k_allele_keys<<<grid, 128, 0, s>>>(ukeys, chk, rep, sgid, nruns, kpairs, m);
    *d_key_pairs = kpairs;
    *d_rep_sorted = starts; *d_count_sorted = scounts; *d_nruns = nruns; *d_err = err;
    return cudaGetLastError();
}


// ---- stream compaction of the reads that go to the reverse-complement rescue (read order kept) ----
struct FlagSet {
    const uint8_t *flags;
    int bit;
    __device__ __forceinline__ bool operator()(const int32_t &i) const { return (flags[i] & bit) != 0; }
};

size_t select_scratch_bytes(int64_t n)
{
    size_t t = 0;
    thrust::counting_iterator<int32_t> it(0);
    cub::DeviceSelect::If(nullptr, t, it, (int32_t *)nullptr, (int *)nullptr, (int)n, FlagSet{nullptr, 0});
    return t + 256;
}

cudaError_t select_flagged(const uint8_t *flags, int64_t n, int bit, int32_t *out_idx, int *d_count, void *tmp, size_t tmp_bytes,
                           cudaStream_t s)
{
    thrust::counting_iterator<int32_t> it(0);
    return cub::DeviceSelect::If(tmp, tmp_bytes, it, out_idx, d_count, (int)n, FlagSet{flags, bit}, s);
}

}  // namespace crgpu
