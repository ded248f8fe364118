// flash_merge.cu -- paired-end merge on the device: k_merge_decide, k_merge_emit, crgpu_flash_merge.
//
// Replaces the `flash R1 R2 --allow-outies --max-overlap M --min-overlap m` subprocess of
// CRISPResso/CRISPRessoCORE.py:1655-1664 (FLASH 1.2.11, a third-party binary that is not part of the
// reference tree; algorithm as restated in SURVEY.md App. D, pinned through the reference's end-to-end
// known-answer test tests/crispresso_tests.py:127-195, which starts from paired FASTQ files).
//
// Per pair: read 2 is reverse-complemented; every overlap of read 1's tail with read 2's head (an
// "innie", overlap starting at i in read 1) and -- with allow_outies -- of read 2's tail with read 1's
// head (an "outie") is scored: columns where either base is N are skipped, eff = remaining columns,
// score_len = min(eff, max_overlap), density = mismatches / score_len (float32), qscore =
// sum over mismatches of min(q1, q2) / score_len (float32).  FLASH scans the candidates in order and
// keeps a candidate when `density <= best && (density < best || qscore < best_q)`; that is the first
// lexicographic minimum of (density, qscore) in scan order (innies by increasing i, then outies).
//
// Device form: one warp per pair.  Both reads become three bit planes (code bit 0, code bit 1, N) built
// with ballots; a candidate is three funnel shifts + XOR/OR/AND + two POPCs per 32 columns, lanes own
// consecutive candidates (so the shifted words are a shared-memory broadcast).  The warp reduces
// (density bits, scan order) with a 64-bit min.  qscore is only ever needed when two candidates tie on
// a non-zero density: that rare case re-scans and sums qualities cooperatively, exactly as FLASH orders it.
#include "crgpu_internal.h"

#include <cub/cub.cuh>

namespace crgpu {

constexpr int MERGE_MAXLEN = 1024;                 // longest mate
constexpr int MERGE_W = MERGE_MAXLEN / 32;         // words per bit plane
constexpr int MERGE_PW = MERGE_W + 2;              // + zero words read by the funnel shift past the end
constexpr int MERGE_WARPS = 4;

struct MergeArgs {
    const uint8_t *s1, *q1; const int64_t *off1;
    const uint8_t *s2, *q2; const int64_t *off2;
    int64_t n;
    int min_overlap, max_overlap;
    float max_density;
    int allow_outies;
    int32_t *pos;        // [n] overlap start in the left read (-1: not combined)
    uint8_t *kind;       // [n] 0 not combined, 1 innie, 2 outie
    int64_t *mlen;       // [n] merged length, 0 when not combined
    int32_t *flag;       // [n] 1 when combined
    int *err;
};

__device__ __forceinline__ int base_code(uint8_t c)        // 0..3 ACGT, 4 N, -1 anything else
{
    switch (c) {
    case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'N': return 4;
    default: return -1;
    }
}

__device__ __forceinline__ uint8_t comp_base(uint8_t c)
{
    switch (c) {
    case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
    default: return c;
    }
}

struct Planes { uint32_t lo[MERGE_PW], hi[MERGE_PW], nn[MERGE_PW]; };

// eff / mismatches of the overlap L[i .. i+n) vs R[0 .. n)
__device__ __forceinline__ void overlap_counts(const Planes &L, const Planes &R, int i, int n, int &eff, int &nm)
{
    const int w0 = i >> 5, sh = i & 31;
    uint32_t lo0 = L.lo[w0], hi0 = L.hi[w0], nn0 = L.nn[w0];
    eff = 0; nm = 0;
    const int nw = (n + 31) >> 5;
    for (int w = 0; w < nw; ++w) {
        const uint32_t lo1 = L.lo[w0 + w + 1], hi1 = L.hi[w0 + w + 1], nn1 = L.nn[w0 + w + 1];
        const uint32_t lo = __funnelshift_r(lo0, lo1, sh), hi = __funnelshift_r(hi0, hi1, sh), nn = __funnelshift_r(nn0, nn1, sh);
        const int left = n - (w << 5);
        const uint32_t mask = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
        const uint32_t valid = ~(nn | R.nn[w]) & mask;
        const uint32_t mm = ((lo ^ R.lo[w]) | (hi ^ R.hi[w])) & valid;
        eff += __popc(valid);
        nm += __popc(mm);
        lo0 = lo1; hi0 = hi1; nn0 = nn1;
    }
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d);
        v = o < v ? o : v;
    }
    return v;
}

__global__ void __launch_bounds__(MERGE_WARPS * 32) k_merge_decide(const MergeArgs a)
{
    __shared__ Planes planes[MERGE_WARPS][2];          // [warp][0 = read 1, 1 = revcomp(read 2)]
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    Planes &P1 = planes[wib][0], &P2 = planes[wib][1];
    const int64_t warp = (int64_t)blockIdx.x * MERGE_WARPS + wib;
    const int64_t nwarps = (int64_t)gridDim.x * MERGE_WARPS;
    for (int64_t p = warp; p < a.n; p += nwarps) {
        const int64_t o1 = a.off1[p], o2 = a.off2[p];
        const int l1 = (int)(a.off1[p + 1] - o1), l2 = (int)(a.off2[p + 1] - o2);
        bool bad = l1 > MERGE_MAXLEN || l2 > MERGE_MAXLEN;
        __syncwarp();
        if (!bad) {
            // ---- bit planes (positions past the end read as N, so they are never valid columns) ----
            const int W1 = (l1 + 31) >> 5, W2 = (l2 + 31) >> 5;
            for (int w = 0; w < W1; ++w) {
                const int x = (w << 5) + lane;
                const int c = x < l1 ? base_code(a.s1[o1 + x]) : 4;
                bad |= c < 0;
                const uint32_t lo = __ballot_sync(0xffffffffu, c & 1), hi = __ballot_sync(0xffffffffu, c & 2),
                               nn = __ballot_sync(0xffffffffu, c == 4);
                if (lane == 0) { P1.lo[w] = lo; P1.hi[w] = hi; P1.nn[w] = nn; }
            }
            for (int w = 0; w < W2; ++w) {
                const int x = (w << 5) + lane;
                int c = x < l2 ? base_code(a.s2[o2 + (l2 - 1 - x)]) : 4;
                bad |= c < 0;
                if (c >= 0 && c < 4) c = 3 - c;                       // complement
                const uint32_t lo = __ballot_sync(0xffffffffu, c & 1), hi = __ballot_sync(0xffffffffu, c & 2),
                               nn = __ballot_sync(0xffffffffu, c == 4);
                if (lane == 0) { P2.lo[w] = lo; P2.hi[w] = hi; P2.nn[w] = nn; }
            }
            for (int w = W1 + lane; w < MERGE_PW; w += 32) { P1.lo[w] = 0; P1.hi[w] = 0; P1.nn[w] = 0xffffffffu; }
            for (int w = W2 + lane; w < MERGE_PW; w += 32) { P2.lo[w] = 0; P2.hi[w] = 0; P2.nn[w] = 0xffffffffu; }
        }
        bad = __any_sync(0xffffffffu, bad);
        __syncwarp();
        if (bad) {
            if (lane == 0) { atomicExch(a.err, 1); a.pos[p] = -1; a.kind[p] = 0; a.mlen[p] = 0; a.flag[p] = 0; }
            continue;
        }
        // ---- candidates in FLASH's scan order: innies i = i_in0 .. l1 - min_overlap, then outies ----
        const int i_in0 = max(0, l1 - l2), n_in = max(0, l1 - a.min_overlap + 1 - i_in0);
        const int i_out0 = max(0, l2 - l1), n_out = a.allow_outies ? max(0, l2 - a.min_overlap + 1 - i_out0) : 0;
        const int ncand = n_in + n_out;
        const unsigned long long NONE = ~0ull;
        unsigned long long best = NONE;                // (density bits << 32) | scan order
        int ties = 0;                                  // candidates of this lane sharing its best density
        for (int c = lane; c < ncand; c += 32) {
            int eff, nm;
            if (c < n_in) { const int i = i_in0 + c; overlap_counts(P1, P2, i, l1 - i, eff, nm); }
            else { const int i = i_out0 + (c - n_in); overlap_counts(P2, P1, i, l2 - i, eff, nm); }
            if (eff < a.min_overlap) continue;
            const float density = __fdiv_rn((float)nm, (float)min(eff, a.max_overlap));
            const unsigned long long key = ((unsigned long long)__float_as_uint(density) << 32) | (unsigned)c;
            if ((key >> 32) < (best >> 32)) { best = key; ties = 1; }
            else if ((key >> 32) == (best >> 32)) ++ties;          // later in scan order: best keeps the first
        }
        const unsigned long long wbest = warp_min_u64(best);
        int win = -1;                                  // winning candidate (scan order)
        if (wbest != NONE && __uint_as_float((unsigned)(wbest >> 32)) <= a.max_density) {
            int t = ((best >> 32) == (wbest >> 32)) ? ties : 0;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
            win = (int)(unsigned)wbest;
            if (t > 1 && (wbest >> 32) != 0) {
                // several candidates share the lowest (non-zero) density: FLASH's quality tie-break.
                // Re-scan in order; the warp sums min(q1,q2)-33 over each tied candidate's mismatches.
                float best_q = 0.f;
                bool have = false;
                for (int c0 = 0; c0 < ncand; c0 += 32) {
                    const int c = c0 + lane;
                    int eff = 0, nm = 0;
                    bool tied = false;
                    if (c < ncand) {
                        if (c < n_in) { const int i = i_in0 + c; overlap_counts(P1, P2, i, l1 - i, eff, nm); }
                        else { const int i = i_out0 + (c - n_in); overlap_counts(P2, P1, i, l2 - i, eff, nm); }
                        if (eff >= a.min_overlap)
                            tied = __float_as_uint(__fdiv_rn((float)nm, (float)min(eff, a.max_overlap))) == (unsigned)(wbest >> 32);
                    }
                    unsigned m = __ballot_sync(0xffffffffu, tied);
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const int cc = c0 + src;
                        const int ceff = __shfl_sync(0xffffffffu, eff, src);
                        const bool innie = cc < n_in;
                        const int i = innie ? i_in0 + cc : i_out0 + (cc - n_in);
                        const int n = innie ? l1 - i : l2 - i;
                        int sum = 0;
                        for (int k = lane; k < n; k += 32) {
                            // left read position i + k, right read position k
                            uint8_t x1, y1, x2, y2;              // (base, quality) of read 1 / revcomp(read 2)
                            const int k1 = innie ? i + k : k, k2 = innie ? k : i + k;
                            x1 = a.s1[o1 + k1]; y1 = a.q1[o1 + k1];
                            x2 = comp_base(a.s2[o2 + (l2 - 1 - k2)]); y2 = a.q2[o2 + (l2 - 1 - k2)];
                            if (x1 != 'N' && x2 != 'N' && x1 != x2) sum += (int)min(y1, y2) - 33;
                        }
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
                        const float q = __fdiv_rn((float)sum, (float)min(ceff, a.max_overlap));
                        if (!have || q < best_q) { have = true; best_q = q; win = cc; }
                    }
                }
            }
        }
        if (lane == 0) {
            if (win < 0) { a.pos[p] = -1; a.kind[p] = 0; a.mlen[p] = 0; a.flag[p] = 0; }
            else if (win < n_in) {
                const int i = i_in0 + win;
                a.pos[p] = i; a.kind[p] = 1; a.mlen[p] = (int64_t)i + l2; a.flag[p] = 1;
            } else {
                const int i = i_out0 + (win - n_in);
                a.pos[p] = i; a.kind[p] = 2; a.mlen[p] = (int64_t)(l2 - i); a.flag[p] = 1;
                atomicAdd(a.err + 1, 1);                         // outies (a.err[1]); innies = merged - outies
            }
        }
    }
}

struct EmitArgs {
    const uint8_t *s1, *q1; const int64_t *off1;
    const uint8_t *s2, *q2; const int64_t *off2;
    int64_t n;
    const int32_t *pos; const uint8_t *kind;
    const int64_t *boff;     // [n] exclusive scan of mlen
    const int32_t *cidx;     // [n] exclusive scan of flag
    uint8_t *out_seq, *out_qual;
    int64_t *out_off;        // [n_merged + 1]
    int32_t *out_index;      // [n_merged]
};

// One warp per combined pair: the merged read = left overhang + overlap + right overhang (innie) or the
// overlap alone (outie: both overhangs are adapter read-through).  Overlap column: equal bases -> that
// base with the higher quality; different -> the base of higher quality; equal quality -> the right
// read's base unless it is N.
__global__ void __launch_bounds__(128) k_merge_emit(const EmitArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < a.n; p += nwarps) {
        const int kind = a.kind[p];
        if (!kind) continue;
        const int64_t o1 = a.off1[p], o2 = a.off2[p];
        const int l1 = (int)(a.off1[p + 1] - o1), l2 = (int)(a.off2[p + 1] - o2);
        const int i = a.pos[p];
        const int64_t ob = a.boff[p];
        const int j = a.cidx[p];
        const int mlen = kind == 1 ? i + l2 : l2 - i;
        if (lane == 0) { a.out_off[j] = ob; a.out_index[j] = (int32_t)p; }
        for (int t = lane; t < mlen; t += 32) {
            // positions in read 1 (k1) and revcomp(read 2) (k2) covering output column t; -1 = none
            int k1, k2;
            if (kind == 1) { k1 = t < l1 ? t : -1; k2 = t >= i ? t - i : -1; }
            else { k1 = t; k2 = i + t; }
            uint8_t b1 = 0, q1 = 0, b2 = 0, q2 = 0;
            if (k1 >= 0) { b1 = a.s1[o1 + k1]; q1 = a.q1[o1 + k1]; }
            if (k2 >= 0) { b2 = comp_base(a.s2[o2 + (l2 - 1 - k2)]); q2 = a.q2[o2 + (l2 - 1 - k2)]; }
            uint8_t b, q;
            if (k2 < 0) { b = b1; q = q1; }
            else if (k1 < 0) { b = b2; q = q2; }
            else {
                // left/right roles: innie left = read 1, outie left = revcomp(read 2)
                const uint8_t lb = kind == 1 ? b1 : b2, lq = kind == 1 ? q1 : q2;
                const uint8_t rb = kind == 1 ? b2 : b1, rq = kind == 1 ? q2 : q1;
                if (lb == rb) { b = lb; q = lq > rq ? lq : rq; }
                else if (lq > rq) { b = lb; q = lq; }
                else if (rq > lq) { b = rb; q = rq; }
                else if (rb != 'N') { b = rb; q = rq; }
                else { b = lb; q = lq; }
            }
            a.out_seq[ob + t] = b;
            a.out_qual[ob + t] = q;
        }
    }
}

}  // namespace crgpu

using namespace crgpu;

static int flash_merge_impl(crgpu_ctx *ctx, int mem, const uint8_t *seq1, const uint8_t *qual1, const int64_t *off1,
                                 const uint8_t *seq2, const uint8_t *qual2, const int64_t *off2, int64_t n,
                                 const crgpu_merge_params *prm, crgpu_merge_out *out)
{
    if (!ctx) return CRGPU_E_ARG;
    if (!prm || !out || n < 0 || (n > 0 && (!seq1 || !qual1 || !off1 || !seq2 || !qual2 || !off2)))
        return fail(ctx, CRGPU_E_ARG, "crgpu_flash_merge: bad argument");
    if (prm->min_overlap < 1 || prm->max_overlap < prm->min_overlap)
        return fail(ctx, CRGPU_E_ARG, "crgpu_flash_merge: need 1 <= min_overlap <= max_overlap");
    if (mem != CRGPU_MEM_HOST && mem != CRGPU_MEM_DEVICE) return fail(ctx, CRGPU_E_ARG, "bad mem");
    if (n > 0 && (!out->pos || !out->kind || !out->seq || !out->qual || !out->offsets || !out->index))
        return fail(ctx, CRGPU_E_ARG, "crgpu_flash_merge: missing output buffer");
    timing_reset(ctx);
    out->n_merged = 0; out->n_innie = 0; out->n_outie = 0; out->bytes = 0;
    if (n == 0) return CRGPU_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;

    const uint8_t *d_s1 = seq1, *d_q1 = qual1, *d_s2 = seq2, *d_q2 = qual2;
    const int64_t *d_o1 = off1, *d_o2 = off2;
    int64_t tot1 = 0, tot2 = 0;
    if (mem == CRGPU_MEM_HOST) {
        tot1 = off1[n]; tot2 = off2[n];
        CK(ctx->q_in[0].reserve((size_t)std::max<int64_t>(tot1, 1))); CK(ctx->q_in[1].reserve((size_t)std::max<int64_t>(tot1, 1)));
        CK(ctx->q_in[2].reserve((size_t)std::max<int64_t>(tot2, 1))); CK(ctx->q_in[3].reserve((size_t)std::max<int64_t>(tot2, 1)));
        CK(ctx->q_in[4].reserve((size_t)(n + 1) * 8)); CK(ctx->q_in[5].reserve((size_t)(n + 1) * 8));
        CK(cudaMemcpyAsync(ctx->q_in[0].p, seq1, (size_t)tot1, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->q_in[1].p, qual1, (size_t)tot1, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->q_in[2].p, seq2, (size_t)tot2, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->q_in[3].p, qual2, (size_t)tot2, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->q_in[4].p, off1, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->q_in[5].p, off2, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, s));
        d_s1 = ctx->q_in[0].as<uint8_t>(); d_q1 = ctx->q_in[1].as<uint8_t>();
        d_s2 = ctx->q_in[2].as<uint8_t>(); d_q2 = ctx->q_in[3].as<uint8_t>();
        d_o1 = ctx->q_in[4].as<int64_t>(); d_o2 = ctx->q_in[5].as<int64_t>();
    }

    // per-pair decision arrays + scans
    CK(ctx->aux[0].reserve((size_t)n * 4));            // pos
    CK(ctx->aux[1].reserve((size_t)n));                // kind
    CK(ctx->aux[2].reserve((size_t)(n + 1) * 8));      // mlen, then its exclusive scan (+ total)
    CK(ctx->aux[3].reserve((size_t)(n + 1) * 4));      // flag, then its exclusive scan (+ total)
    CK(ctx->aux[4].reserve((size_t)(n + 1) * 8));
    CK(ctx->aux[5].reserve((size_t)(n + 1) * 4));
    CK(ctx->errflag.reserve(16));
    int32_t *d_pos = mem == CRGPU_MEM_DEVICE ? out->pos : ctx->aux[0].as<int32_t>();
    uint8_t *d_kind = mem == CRGPU_MEM_DEVICE ? out->kind : ctx->aux[1].as<uint8_t>();
    int64_t *d_mlen = ctx->aux[2].as<int64_t>(), *d_boff = ctx->aux[4].as<int64_t>();
    int32_t *d_flag = ctx->aux[3].as<int32_t>(), *d_cidx = ctx->aux[5].as<int32_t>();
    int *d_err = ctx->errflag.as<int>();
    CK(cudaMemsetAsync(d_err, 0, 16, s));
    CK(cudaMemsetAsync(d_mlen + n, 0, 8, s));
    CK(cudaMemsetAsync(d_flag + n, 0, 4, s));

    MergeArgs ma;
    ma.s1 = d_s1; ma.q1 = d_q1; ma.off1 = d_o1; ma.s2 = d_s2; ma.q2 = d_q2; ma.off2 = d_o2; ma.n = n;
    ma.min_overlap = prm->min_overlap; ma.max_overlap = prm->max_overlap; ma.max_density = prm->max_mismatch_density;
    ma.allow_outies = prm->allow_outies ? 1 : 0;
    ma.pos = d_pos; ma.kind = d_kind; ma.mlen = d_mlen; ma.flag = d_flag; ma.err = d_err;
    int64_t grid = (n + MERGE_WARPS - 1) / MERGE_WARPS;
    grid = std::min<int64_t>(grid, (int64_t)ctx->num_sms * 16);
    span_begin(ctx, T_OTHER);
    k_merge_decide<<<(int)grid, MERGE_WARPS * 32, 0, s>>>(ma);
    CK(cudaGetLastError());
    size_t tmp1 = 0, tmp2 = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp1, d_mlen, d_boff, n + 1, s));
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, d_flag, d_cidx, n + 1, s));
    CK(ctx->aux[6].reserve(std::max(tmp1, tmp2)));
    CK(cub::DeviceScan::ExclusiveSum(ctx->aux[6].p, tmp1, d_mlen, d_boff, n + 1, s));
    CK(cub::DeviceScan::ExclusiveSum(ctx->aux[6].p, tmp2, d_flag, d_cidx, n + 1, s));
    span_end(ctx, 5);
    int64_t total = 0;
    int32_t nm = 0;
    int herr2[2] = {0, 0};
    CK(cudaMemcpyAsync(&total, d_boff + n, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&nm, d_cidx + n, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(herr2, d_err, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int herr = herr2[0];
    if (herr) return fail(ctx, CRGPU_E_ALIGN, "crgpu_flash_merge: a mate is longer than %d bases or holds a base outside ACGTN", MERGE_MAXLEN);
    if (total > out->cap_bytes || nm > out->cap_reads)
        return fail(ctx, CRGPU_E_ARG, "crgpu_flash_merge: output capacity too small (need %lld bytes, %d reads)", (long long)total, nm);

    uint8_t *d_oseq = out->seq, *d_oqual = out->qual;
    int64_t *d_ooff = out->offsets;
    int32_t *d_oidx = out->index;
    if (mem == CRGPU_MEM_HOST) {
        CK(ctx->q_out[0].reserve((size_t)std::max<int64_t>(total, 1))); CK(ctx->q_out[1].reserve((size_t)std::max<int64_t>(total, 1)));
        CK(ctx->q_out[2].reserve((size_t)(nm + 1) * 8)); CK(ctx->q_out[3].reserve((size_t)(nm + 1) * 4));
        d_oseq = ctx->q_out[0].as<uint8_t>(); d_oqual = ctx->q_out[1].as<uint8_t>();
        d_ooff = ctx->q_out[2].as<int64_t>(); d_oidx = ctx->q_out[3].as<int32_t>();
    }
    EmitArgs ea;
    ea.s1 = d_s1; ea.q1 = d_q1; ea.off1 = d_o1; ea.s2 = d_s2; ea.q2 = d_q2; ea.off2 = d_o2; ea.n = n;
    ea.pos = d_pos; ea.kind = d_kind; ea.boff = d_boff; ea.cidx = d_cidx;
    ea.out_seq = d_oseq; ea.out_qual = d_oqual; ea.out_off = d_ooff; ea.out_index = d_oidx;
    int64_t grid2 = std::min<int64_t>((n + 3) / 4, (int64_t)ctx->num_sms * 16);
    span_begin(ctx, T_OTHER);
    k_merge_emit<<<(int)grid2, 128, 0, s>>>(ea);
    CK(cudaGetLastError());
    span_end(ctx, 1);
    CK(cudaMemcpyAsync(d_ooff + nm, d_boff + n, 8, cudaMemcpyDeviceToDevice, s));
    if (mem == CRGPU_MEM_HOST) {
        CK(cudaMemcpyAsync(out->pos, d_pos, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(out->kind, d_kind, (size_t)n, cudaMemcpyDeviceToHost, s));
        if (total > 0) {
            CK(cudaMemcpyAsync(out->seq, d_oseq, (size_t)total, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(out->qual, d_oqual, (size_t)total, cudaMemcpyDeviceToHost, s));
        }
        CK(cudaMemcpyAsync(out->offsets, d_ooff, (size_t)(nm + 1) * 8, cudaMemcpyDeviceToHost, s));
        if (nm > 0) CK(cudaMemcpyAsync(out->index, d_oidx, (size_t)nm * 4, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    out->n_merged = nm;
    out->n_outie = herr2[1];
    out->n_innie = (int64_t)nm - herr2[1];
    out->bytes = total;
    timing_collect(ctx);
    return CRGPU_OK;
}

// the exported entry point: device guard + "no work of a failed call is left running" (ApiGuard, crgpu_internal.h)
extern "C" int crgpu_flash_merge(crgpu_ctx *ctx, int mem, const uint8_t *seq1, const uint8_t *qual1, const int64_t *off1,
                                 const uint8_t *seq2, const uint8_t *qual2, const int64_t *off2, int64_t n,
                                 const crgpu_merge_params *prm, crgpu_merge_out *out)
{
    if (!ctx) return CRGPU_E_ARG;
    ApiGuard guard(ctx);
    return guard.done(flash_merge_impl(ctx, mem, seq1, qual1, off1, seq2, qual2, off2, n, prm, out));
}
