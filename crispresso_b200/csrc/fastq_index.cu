// fastq_index.cu -- FASTQ record splitting on the device: k_count_newlines, k_line_starts,
// k_fastq_records, k_fastq_gather, crgpu_fastq_index.
//
// Replaces the text plumbing either side of the hot path in CRISPResso/CRISPRessoCORE.py:
//   * `gunzip | awk 'NR % 4 == 1 {print ">" $0} NR % 4 == 2 {print $0}' | sed 's/:/_/g'` feeding needle
//     (CORE:1793-1797, 1813-1817, 1911-1915): lines 4r+1 / 4r+2 of the inflated text are the name / bases;
//   * get_n_reads_fastq (`wc -l` // 4, CORE:335-348) and get_average_read_length_fastq
//     (awk: int(sum of line-2 lengths / n), CORE:313-332);
//   * the Biopython record iterator inside the quality filter (CORE:176-190, 216-230, 290-306).
// gzip inflate stays on the host (a deflate stream is sequential); the inflated bytes are indexed here.
//
// HBM-bound byte work, four small passes over a text of B bytes holding n records of s bases:
//   A  k_count_newlines   read B            -> newline count per 8 KiB tile
//      cub ExclusiveSum over the tile counts
//   B  k_line_starts      read B            -> start offset of every line (8 B per line)
//   C  k_fastq_records    read 40 B/record  -> validation ('@', '+', |bases| == |qualities|), lengths
//      cub ExclusiveSum over the lengths    -> offsets[n+1]
//   D  k_fastq_gather     read 2s, write 2s -> packed bases and qualities (the layout crgpu_align,
//                                              crgpu_qualfilter and crgpu_flash_merge consume)
// Algorithmic traffic = 2B + 4s*n + 72n bytes.
#include "crgpu_internal.h"

#include <cub/cub.cuh>

namespace crgpu {

constexpr int FQ_TILE = 8192;          // bytes per CTA tile
constexpr int FQ_THREADS = 128;        // 64 bytes per thread: four 16-byte loads

__device__ __forceinline__ unsigned nl_mask4(uint32_t w)      // 0xff in every byte of w that is '\n'
{
    return __vcmpeq4(w, 0x0a0a0a0au);
}

__device__ __forceinline__ int count16(const uint4 v)
{
    return (__popc(nl_mask4(v.x)) + __popc(nl_mask4(v.y)) + __popc(nl_mask4(v.z)) + __popc(nl_mask4(v.w))) >> 3;
}

// tile bytes of thread t: [tile*FQ_TILE + t*64, +64); text is padded to a multiple of 16 bytes by the
// caller (device scratch) or read bytewise at the ragged end.
__device__ __forceinline__ uint4 load16(const uint8_t *text, int64_t at, int64_t nbytes)
{
    if (at + 16 <= nbytes && ((reinterpret_cast<uintptr_t>(text + at) & 15) == 0))
        return *reinterpret_cast<const uint4 *>(text + at);
    uint32_t w[4] = {0, 0, 0, 0};
    for (int k = 0; k < 16; ++k)
        if (at + k < nbytes) w[k >> 2] |= (uint32_t)text[at + k] << ((k & 3) * 8);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(FQ_THREADS) k_count_newlines(const uint8_t *__restrict__ text, int64_t nbytes,
                                                               int64_t ntiles, int64_t *__restrict__ tile_count)
{
    __shared__ int warp_sum[FQ_THREADS / 32];
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * FQ_TILE + (int64_t)threadIdx.x * 64;
        int c = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (base + 16 * k < nbytes) c += count16(load16(text, base + 16 * k, nbytes));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            int s = 0;
            for (int w = 0; w < FQ_THREADS / 32; ++w) s += warp_sum[w];
            tile_count[tile] = s;
        }
        __syncthreads();
    }
}

// line_start[g + 1] = offset of the byte after the g-th newline; line_start[0] = 0 is written by the host code.
__global__ void __launch_bounds__(FQ_THREADS) k_line_starts(const uint8_t *__restrict__ text, int64_t nbytes, int64_t ntiles,
                                                            const int64_t *__restrict__ tile_base,
                                                            int64_t *__restrict__ line_start)
{
    typedef cub::BlockScan<int, FQ_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * FQ_TILE + (int64_t)threadIdx.x * 64;
        uint4 v[4];
        int c = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[k] = base + 16 * k < nbytes ? load16(text, base + 16 * k, nbytes) : make_uint4(0, 0, 0, 0);
            c += count16(v[k]);
        }
        int before;
        Scan(tmp).ExclusiveSum(c, before);
        int64_t g = tile_base[tile] + before;
        if (c) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    unsigned m = nl_mask4(w[j]) & 0x01010101u;
                    while (m) {
                        const int b = (__ffs(m) - 1) >> 3;
                        m &= m - 1;
                        line_start[++g] = base + 16 * k + 4 * j + b + 1;
                    }
                }
            }
        }
        __syncthreads();
    }
}

struct FastqRecArgs {
    const uint8_t *text;
    int64_t nbytes;
    const int64_t *line_start;     // [nlines + 1]; line l = text[line_start[l] .. line_start[l+1] - 1)
    int64_t nrec;
    int64_t *len;                  // [nrec + 1] bases per record (scanned into offsets afterwards)
    int64_t *name_start;           // [nrec] or null
    int32_t *name_len;             // [nrec] or null
    int *err;                      // err[0] = flag, err[1] = first bad record (min)
};

__global__ void k_fastq_records(const FastqRecArgs a)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.nrec) return;
    const int64_t l0 = a.line_start[4 * r], l1 = a.line_start[4 * r + 1], l2 = a.line_start[4 * r + 2],
                  l3 = a.line_start[4 * r + 3], l4 = a.line_start[4 * r + 4];
    // a line ends one byte before the next line starts (its '\n'); a Windows '\r' before it is not a base
    int64_t slen = l2 - 1 - l1, qlen = l4 - 1 - l3, hlen = l1 - 1 - l0;
    if (slen > 0 && a.text[l2 - 2] == '\r') --slen;
    if (qlen > 0 && a.text[l4 - 2] == '\r') --qlen;
    if (hlen > 0 && a.text[l1 - 2] == '\r') --hlen;
    const bool ok = a.text[l0] == '@' && a.text[l2] == '+' && slen == qlen;
    if (!ok) { atomicExch(&a.err[0], 1); atomicMin(&a.err[1], (int)min(r, (int64_t)0x7fffffff)); }
    a.len[r] = ok ? slen : 0;
    if (a.name_start) a.name_start[r] = l0;
    if (a.name_len) a.name_len[r] = (int32_t)hlen;
}

// one warp per record: bases and qualities copied into the packed buffers
__global__ void __launch_bounds__(128) k_fastq_gather(const uint8_t *__restrict__ text, const int64_t *__restrict__ line_start,
                                                      const int64_t *__restrict__ offsets, int64_t nrec,
                                                      uint8_t *__restrict__ seq, uint8_t *__restrict__ qual)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < nrec; r += nwarps) {
        const int64_t o = offsets[r];
        const int len = (int)(offsets[r + 1] - o);
        const uint8_t *s = text + line_start[4 * r + 1], *q = text + line_start[4 * r + 3];
        for (int x = lane; x < len; x += 32) {
            seq[o + x] = s[x];
            qual[o + x] = q[x];
        }
    }
}

}  // namespace crgpu

using namespace crgpu;

static int fastq_index_impl(crgpu_ctx *ctx, int mem, const uint8_t *text, int64_t nbytes, int final_chunk,
                                 crgpu_fastq_out *out)
{
    if (!ctx) return CRGPU_E_ARG;
    if (!out || nbytes < 0 || (nbytes > 0 && !text)) return fail(ctx, CRGPU_E_ARG, "crgpu_fastq_index: bad argument");
    if (mem != CRGPU_MEM_HOST && mem != CRGPU_MEM_DEVICE) return fail(ctx, CRGPU_E_ARG, "bad mem");
    timing_reset(ctx);
    out->n_records = 0; out->consumed = 0; out->seq_bytes = 0;
    if (nbytes == 0) return CRGPU_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;

    // a final chunk may lack the trailing newline: index it as if it were there
    int64_t nb = nbytes;
    bool add_nl = false;
    const uint8_t *d_text = text;
    if (mem == CRGPU_MEM_HOST) {
        add_nl = final_chunk && text[nbytes - 1] != '\n';
        CK(ctx->q_in[0].reserve((size_t)nbytes + 16));
        CK(cudaMemcpyAsync(ctx->q_in[0].p, text, (size_t)nbytes, cudaMemcpyHostToDevice, s));
        if (add_nl) { CK(cudaMemsetAsync(ctx->q_in[0].as<uint8_t>() + nbytes, '\n', 1, s)); nb = nbytes + 1; }
        d_text = ctx->q_in[0].as<uint8_t>();
    } else if (final_chunk) {
        uint8_t last = 0;
        CK(cudaMemcpyAsync(&last, text + nbytes - 1, 1, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (last != '\n') {          // device text is the caller's: index a private copy with the newline added
            CK(ctx->q_in[0].reserve((size_t)nbytes + 16));
            CK(cudaMemcpyAsync(ctx->q_in[0].p, text, (size_t)nbytes, cudaMemcpyDeviceToDevice, s));
            CK(cudaMemsetAsync(ctx->q_in[0].as<uint8_t>() + nbytes, '\n', 1, s));
            d_text = ctx->q_in[0].as<uint8_t>(); nb = nbytes + 1; add_nl = true;
        }
    }

    const int64_t ntiles = (nb + FQ_TILE - 1) / FQ_TILE;
    CK(ctx->aux[0].reserve((size_t)(ntiles + 1) * 8));      // tile counts
    CK(ctx->aux[1].reserve((size_t)(ntiles + 1) * 8));      // their exclusive scan (+ total)
    int64_t *d_tc = ctx->aux[0].as<int64_t>(), *d_tb = ctx->aux[1].as<int64_t>();
    CK(cudaMemsetAsync(d_tc + ntiles, 0, 8, s));
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctx->num_sms * 16);
    span_begin(ctx, T_OTHER);
    k_count_newlines<<<grid, FQ_THREADS, 0, s>>>(d_text, nb, ntiles, d_tc);
    CK(cudaGetLastError());
    size_t tmp = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_tc, d_tb, ntiles + 1, s));
    CK(ctx->aux[6].reserve(tmp));
    CK(cub::DeviceScan::ExclusiveSum(ctx->aux[6].p, tmp, d_tc, d_tb, ntiles + 1, s));
    span_end(ctx, 3);
    int64_t nlines = 0;
    CK(cudaMemcpyAsync(&nlines, d_tb + ntiles, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const int64_t nrec = nlines / 4;
    if (final_chunk && nlines % 4 != 0)
        return fail(ctx, CRGPU_E_ARG, "crgpu_fastq_index: %lld lines is not a whole number of 4-line FASTQ records", (long long)nlines);
    out->n_records = nrec;
    if (nrec == 0) return CRGPU_OK;

    CK(ctx->aux[2].reserve((size_t)(nlines + 1) * 8));      // line starts
    int64_t *d_ls = ctx->aux[2].as<int64_t>();
    CK(cudaMemsetAsync(d_ls, 0, 8, s));
    span_begin(ctx, T_OTHER);
    k_line_starts<<<grid, FQ_THREADS, 0, s>>>(d_text, nb, ntiles, d_tb, d_ls);
    CK(cudaGetLastError());
    span_end(ctx, 1);

    const bool want = out->seq != nullptr || out->qual != nullptr || out->offsets != nullptr;
    CK(ctx->aux[3].reserve((size_t)(nrec + 1) * 8));        // lengths
    CK(ctx->aux[4].reserve((size_t)(nrec + 1) * 8));        // offsets
    CK(ctx->errflag.reserve(16));
    int64_t *d_len = ctx->aux[3].as<int64_t>(), *d_off = ctx->aux[4].as<int64_t>();
    int *d_err = ctx->errflag.as<int>();
    const int errinit[2] = {0, 0x7fffffff};
    CK(cudaMemcpyAsync(d_err, errinit, 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(d_len + nrec, 0, 8, s));
    int64_t *d_ns = nullptr;
    int32_t *d_nl = nullptr;
    if (out->name_start && out->name_len) {
        if (out->cap_records < nrec)
            return fail(ctx, CRGPU_E_ARG, "crgpu_fastq_index: cap_records %lld < %lld records", (long long)out->cap_records, (long long)nrec);
        if (mem == CRGPU_MEM_DEVICE) { d_ns = out->name_start; d_nl = out->name_len; }
        else {
            CK(ctx->q_out[2].reserve((size_t)nrec * 8)); CK(ctx->q_out[3].reserve((size_t)nrec * 4));
            d_ns = ctx->q_out[2].as<int64_t>(); d_nl = ctx->q_out[3].as<int32_t>();
        }
    }
    FastqRecArgs ra;
    ra.text = d_text; ra.nbytes = nb; ra.line_start = d_ls; ra.nrec = nrec; ra.len = d_len;
    ra.name_start = d_ns; ra.name_len = d_nl; ra.err = d_err;
    span_begin(ctx, T_OTHER);
    k_fastq_records<<<(int)((nrec + 255) / 256), 256, 0, s>>>(ra);
    CK(cudaGetLastError());
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_len, d_off, nrec + 1, s));
    CK(ctx->aux[6].reserve(tmp));
    CK(cub::DeviceScan::ExclusiveSum(ctx->aux[6].p, tmp, d_len, d_off, nrec + 1, s));
    span_end(ctx, 3);
    int64_t total = 0, consumed = 0;
    int herr[2] = {0, 0};
    CK(cudaMemcpyAsync(&total, d_off + nrec, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&consumed, d_ls + 4 * nrec, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(herr, d_err, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (herr[0])
        return fail(ctx, CRGPU_E_ARG, "crgpu_fastq_index: record %d is not FASTQ (line 1 must start with '@', line 3 with '+', "
                                      "bases and qualities must have the same length)", herr[1]);
    out->seq_bytes = total;
    out->consumed = std::min(consumed, nbytes);             // the newline added to a final chunk is not the caller's
    if (!want) { timing_collect(ctx); return CRGPU_OK; }
    if (!out->seq || !out->qual || !out->offsets)
        return fail(ctx, CRGPU_E_ARG, "crgpu_fastq_index: seq, qual and offsets go together");
    if (out->cap_records < nrec || out->cap_bytes < total)
        return fail(ctx, CRGPU_E_ARG, "crgpu_fastq_index: capacity too small (need %lld records, %lld bytes)", (long long)nrec, (long long)total);

    uint8_t *d_seq = out->seq, *d_qual = out->qual;
    if (mem == CRGPU_MEM_HOST) {
        CK(ctx->q_out[0].reserve((size_t)std::max<int64_t>(total, 1))); CK(ctx->q_out[1].reserve((size_t)std::max<int64_t>(total, 1)));
        d_seq = ctx->q_out[0].as<uint8_t>(); d_qual = ctx->q_out[1].as<uint8_t>();
    }
    span_begin(ctx, T_OTHER);
    k_fastq_gather<<<(int)std::min<int64_t>((nrec + 3) / 4, (int64_t)ctx->num_sms * 32), 128, 0, s>>>(d_text, d_ls, d_off, nrec, d_seq, d_qual);
    CK(cudaGetLastError());
    span_end(ctx, 1);
    if (mem == CRGPU_MEM_HOST) {
        if (total > 0) {
            CK(cudaMemcpyAsync(out->seq, d_seq, (size_t)total, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(out->qual, d_qual, (size_t)total, cudaMemcpyDeviceToHost, s));
        }
        CK(cudaMemcpyAsync(out->offsets, d_off, (size_t)(nrec + 1) * 8, cudaMemcpyDeviceToHost, s));
        if (d_ns) {
            CK(cudaMemcpyAsync(out->name_start, d_ns, (size_t)nrec * 8, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(out->name_len, d_nl, (size_t)nrec * 4, cudaMemcpyDeviceToHost, s));
        }
    } else {
        CK(cudaMemcpyAsync(out->offsets, d_off, (size_t)(nrec + 1) * 8, cudaMemcpyDeviceToDevice, s));
    }
    CK(cudaStreamSynchronize(s));
    timing_collect(ctx);
    return CRGPU_OK;
}

// the exported entry point: device guard + "no work of a failed call is left running" (ApiGuard, crgpu_internal.h)
extern "C" int crgpu_fastq_index(crgpu_ctx *ctx, int mem, const uint8_t *text, int64_t nbytes, int final_chunk,
                                 crgpu_fastq_out *out)
{
    if (!ctx) return CRGPU_E_ARG;
    ApiGuard guard(ctx);
    return guard.done(fastq_index_impl(ctx, mem, text, nbytes, final_chunk, out));
}
