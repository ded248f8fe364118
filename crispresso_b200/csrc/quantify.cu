// quantify.cu -- placeholder translation unit; the kernel lands in the next commit.
#include "../../include/crgpu.h"
#include "crgpu_common.cuh"
namespace crgpu {
int quantify_device(crgpu_ctx *, const crgpu_quant_params *, const uint8_t *, const uint8_t *, const uint8_t *, int64_t,
                    const int32_t *, const int32_t *, const int32_t *, const int32_t *, const uint8_t *, int64_t,
                    crgpu_read_rec *, int64_t *, int64_t *, int64_t *, int32_t, int32_t, int64_t *) { return CRGPU_E_ARG; }
}
extern "C" {
int crgpu_quantify(crgpu_ctx *, int, const crgpu_quant_params *, const uint8_t *, const uint8_t *, const uint8_t *, int64_t,
                   const int32_t *, const int32_t *, const int32_t *, const int32_t *, const uint8_t *, int64_t,
                   crgpu_read_rec *, int64_t *, int64_t *, int64_t *, int32_t, int32_t, int64_t *) { return CRGPU_E_ARG; }
int crgpu_align_quantify(crgpu_ctx *, int, const char *, int, const crgpu_path_params *, const crgpu_quant_params *,
                         const uint8_t *, const int64_t *, int64_t, uint8_t *, crgpu_aln_rec *, int32_t *, crgpu_read_rec *,
                         uint8_t *, uint8_t *, uint8_t *, int64_t, int64_t *, int64_t *, int64_t *, int32_t, int32_t,
                         int64_t *, int64_t *) { return CRGPU_E_ARG; }
}
