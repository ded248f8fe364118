// int_peak.cu -- integer issue-rate micro-benchmark ("measure it").
// Dependency-free chains (16 independent accumulators per thread) of one instruction kind,
// 1024 threads x 2 CTAs per SM.  The op count reported is lane-ops: one SASS instruction on
// one lane.  `cuobjdump -sass` of this file must show the unrolled body made of exactly the
// named instruction (checked in profiles/r01_sass_notes.md).
#include "crgpu_common.cuh"
#include <cuda_fp16.h>

namespace crgpu {

constexpr int PEAK_CHAINS = 16;
constexpr int PEAK_UNROLL = 8;

// Every op takes another accumulator as its second operand, so ptxas cannot fold a chain into
// a*k+b or merge two adds into one IADD3: one source op == one SASS instruction (checked with
// cuobjdump, profiles/r01_sass_notes.md).
template <int WHICH>
__global__ void __launch_bounds__(1024) k_int_peak(int iters, unsigned *sink, unsigned b, unsigned c)
{
    unsigned a[PEAK_CHAINS];
#pragma unroll
    for (int j = 0; j < PEAK_CHAINS; ++j) a[j] = threadIdx.x * 2654435761u + j * 40503u + blockIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < PEAK_UNROLL; ++u) {
#pragma unroll
            for (int j = 0; j < PEAK_CHAINS; ++j) {
                const unsigned o = a[(j + 1 + u) % PEAK_CHAINS];
                if (WHICH == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[j]) : "r"(o));                        // IADD3 / VIADD (alu)
                else if (WHICH == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(o));     // IMAD (fma)
                else if (WHICH == 2) a[j] = (u & 1) ? __vmaxs2(a[j], o) : __vmins2(a[j], o);                        // VIMNMX.S16x2
                else if (WHICH == 3) a[j] = __viaddmax_s16x2(a[j], b, o);                                           // VIADDMNMX.S16x2
                else if (WHICH == 4) {                                                                              // 1:1 alu + fma
                    if (j & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(o));
                    else a[j] = (u & 1) ? __vmaxs2(a[j], o) : __vmins2(a[j], o);
                }
                else if (WHICH == 5) a[j] = __vimax3_s16x2(a[j], o, c);                                             // VIMNMX3.S16x2
                else if (WHICH == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(o), "r"(c)); // LOP3
                else if (WHICH == 7) a[j] = __vminu2(a[j] - o, 0x00010001u) + c;                                    // sub + VIMNMX imm + add
                else if (WHICH == 8) a[j] = __hne2_mask(*reinterpret_cast<__half2 *>(&a[j]), *reinterpret_cast<const __half2 *>(&o)) ^ c;   // HSET2.NE + LOP3
                else if (WHICH == 9) {                                                                              // HSET2.NE only (result kept live through a xor chain every 8th)
                    const unsigned m_ = __hne2_mask(*reinterpret_cast<__half2 *>(&a[j]), *reinterpret_cast<const __half2 *>(&o));
                    a[j] = m_;
                }
            }
        }
    }
    unsigned r = 0;
#pragma unroll
    for (int j = 0; j < PEAK_CHAINS; ++j) r ^= a[j];
    if (r == 0x12345678u) sink[0] = r;   // practically never: keeps the chains alive
}

cudaError_t launch_int_peak(int which, int num_sms, int iters, unsigned *sink, cudaStream_t s, double *lane_ops)
{
    const int grid = num_sms * 2, block = 1024;
    const unsigned b = 3u + (unsigned)(iters & 1), c = 0x00050007u;
    switch (which) {
    case 0: k_int_peak<0><<<grid, block, 0, s>>>(iters, sink, b, c); break;
    case 1: k_int_peak<1><<<grid, block, 0, s>>>(iters, sink, b, c); break;
    case 2: k_int_peak<2><<<grid, block, 0, s>>>(iters, sink, b, c); break;
    case 3: k_int_peak<3><<<grid, block, 0, s>>>(iters, sink, b, c); break;
    case 4: k_int_peak<4><<<grid, block, 0, s>>>(iters, sink, b, c); break;
    case 5: k_int_peak<5><<<grid, block, 0, s>>>(iters, sink, b, c); break;
    case 6: k_int_peak<6><<<grid, block, 0, s>>>(iters, sink, b, c); break;
    case 7: k_int_peak<7><<<grid, block, 0, s>>>(iters, sink, b, c); break;
    case 8: k_int_peak<8><<<grid, block, 0, s>>>(iters, sink, b, c); break;
    case 9: k_int_peak<9><<<grid, block, 0, s>>>(iters, sink, b, c); break;
    default: return cudaErrorInvalidValue;
    }
    *lane_ops = (double)grid * block * (double)iters * PEAK_UNROLL * PEAK_CHAINS * (which == 7 ? 3.0 : (which == 8 ? 2.0 : 1.0));
    return cudaGetLastError();
}

}  // namespace crgpu
