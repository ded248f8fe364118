// gotoh_tile.cuh -- pieces shared by the fill kernels (gotoh_fill.cu: single pass with flags + band pass;
// gotoh_score2.cu: score pass of the banded fill).
#pragma once
#include "crgpu_common.cuh"

namespace crgpu {

__device__ __forceinline__ uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }      // VIMNMX.U16x2
__device__ __forceinline__ uint32_t vmin2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
__device__ __forceinline__ uint32_t vaddmax2(uint32_t a, uint32_t b, uint32_t c)                   // VIADDMNMX.S16x2
{
    return __viaddmax_s16x2(a, b, c);
}

// a + b issued on the FMA pipe (IMAD): `one` is a register holding 1 that ptxas cannot see through,
// so the multiply-add is not turned back into an IADD3.  Used to take plain adds off the integer-ALU
// pipe, which also has to run every VIMNMX / VIADDMNMX of the cell (profiles/r01_notes.md).
__device__ __forceinline__ uint32_t fma_add(uint32_t a, uint32_t b, uint32_t one)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int K>
struct Strip {
    uint32_t H3[K];   // max(m,ix,iy)[row, x-1]
    uint32_t IX[K];   // ix[row, x-1]
    uint32_t mlast;   // m[row K-1, x-1]; only meaningful in the lane that owns amplicon row La-1
};

// 128-thread CTAs, as many per SM as registers / shared memory allow (3 for the K = 32 strips).
// The register cap per strip height was swept on a B200 (profiles/r01_notes.md): K = 32 runs
// 2007 GCUPS at 136 registers vs 1845 at ptxas' own choice (133) and 1513 at 128 (4 CTAs/SM but
// spills + extra moves); one 12-warp CTA per SM under __launch_bounds__(384) is 25 % slower.
// K = 40: 1605 GCUPS at 144 registers (3 CTAs/SM), 1855 at 184 (2 CTAs/SM): ILP beats occupancy here.
#ifdef FILL_MAXNREG
template <int K> constexpr int fill_maxnreg() { return FILL_MAXNREG; }
#else
template <int K> constexpr int fill_maxnreg() { return K >= 48 ? 255 : (K >= 36 ? 184 : (K >= 32 ? 136 : 128)); }
#endif

enum { FILL_FULL = 0, FILL_SCORE = 1 };

// stage the pair profile with one TMA bulk copy (all threads of the CTA call this)
__device__ __forceinline__ void stage_profile(int32_t *sprof, uint64_t *mbar, const int32_t *prof, uint32_t prof_bytes)
{
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(prof_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sprof)), "l"(prof), "r"(prof_bytes), "r"(smem_u32(mbar)) : "memory");
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(mbar)), "r"(0u) : "memory");
    }
}

// Boundary rows in memory are 16 bytes per column, (max3, iy, m, 0).  A lane writes consecutive columns on
// consecutive steps, so the two halves of a 32-byte sector meet in L2 long before the sector is evicted (a
// half-written sector would cost HBM a read-modify-write: measured 2.6x on the whole kernel when every lane
// leaves its sectors half-written).  First column of pair p in a per-pair array: even, and the ranges of
// consecutive pairs do not overlap.
__device__ __forceinline__ int64_t top_base_col(int64_t pco_rel, int pair_rel) { return (pco_rel + pair_rel + 1) & ~(int64_t)1; }


// (max3, iy, m) of a saved boundary row: 8 + 4 bytes, NOT one 16-byte load -- ptxas recycles the register of the unused
// fourth word as a temporary in the middle of the step, and every such write then waits for the load in flight (a third
// of the HDR pass's stall samples, ncu source page)
__device__ __forceinline__ uint4 load_top3(const uint4 *q)
{
    const uint2 a = *reinterpret_cast<const uint2 *>(q);
    const uint32_t c = reinterpret_cast<const uint32_t *>(q)[2];
    return make_uint4(a.x, a.y, c, 0u);
}

// launch of the score pass + the last-row scan (gotoh_score2.cu; needs an odd band left edge: P + B even)
cudaError_t launch_score2(int G, int K, const FillArgs &a, int num_sms, cudaStream_t stream);

}  // namespace crgpu
