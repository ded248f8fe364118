// quant_args.cuh -- argument block of k_quantify, shared by quantify.cu and hotpath.cu.
#pragma once
#include "../../include/crgpu.h"
#include <cuda_runtime.h>
namespace crgpu {
struct QuantArgs {
    // per-column alignment ops, 2 bits each (0 match, 1 mismatch, 2 gap in the amplicon row = insertion,
    // 3 gap in the read row = deletion), 16 per word, row i at ops + i*ops_stride.  k_traceback_walk
    // emits them in walk order: ops_reversed = 1 means entry 0 is the LAST alignment column.
    const uint32_t *ops;
    int64_t ops_stride;
    int ops_reversed;
    const uint8_t *amp;                // amplicon (forward strand, upper case): N positions for CRGPU_Q_MASK_N
    const int32_t *alnlen;
    const int32_t *tenths_ref, *tenths_rep;   // tenths_rep may be null (no HDR -> NaN)
    const uint8_t *unmod_in;           // UNMODIFIED column on entry
    const uint8_t *active;             // optional row filter (null: all rows)
    int active_bit;                    // row i is processed iff active[i] & active_bit
    int64_t n;
    crgpu_read_rec *recs;
    int L, W, flags;
    double hdr_thr;
    const uint32_t *inc, *exon, *splice;      // W words each (device)
    unsigned long long *vectors;              // [CRGPU_NUM_VECTORS][L]
    unsigned long long *hist_in, *hist_fs;    // [hist_len] or null
    int hist_len, hist_zero;
    unsigned long long *counters;             // [CRGPU_NUM_COUNTERS + 4 class counts + 1 rows]
};

}  // namespace crgpu
