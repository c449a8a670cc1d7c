/*
 * zl_b200_test.h — unit-test and measurement hooks of the B200 detector.  NOT part of the drop-in boundary: these symbols
 * live in libzl_b200_test.so (the engine's objects + the hooks), which only tests/ and scripts/ load; the product
 * library libzl_b200.so (include/zl_b200.h) does not export them.
 */
#ifndef ZL_B200_TEST_H_
#define ZL_B200_TEST_H_

#include "zl_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One convolution through the engine's conv kernels, host tensors in/out.
 * x: [n,h,w,cin] fp32 NHWC; wgt: [cout,kh,kw,cin] fp32; bias: [cout];
 * res: optional [n,ho,wo,cout]; y: [n,ho,wo,cout] fp32.  impl: 0 = fp32 SIMT,
 * 1 = tcgen05 (A via software gather), 2 = tcgen05 (A via TMA where possible), 3 = persistent halo kernel (3x3 s1). */
ZL_API int32_t zl_test_conv(int32_t device, int32_t impl, const float* x, int32_t n, int32_t h, int32_t w,
                            int32_t cin, const float* wgt, const float* bias, int32_t cout,
                            int32_t k, int32_t stride, int32_t act, const float* res, float* y);

/* Measurement hook: cycles for `count` tcgen05.mma (M=128) of width N with the given swizzle / group stride /
 * accumulator rotation; sizes the conv tiles (DESIGN.md "UMMA probe"). */
/* SPPF max-pools (5/9/13 windows, -inf padding) of x [n,h,w,c] into the concat buffer [n,h,w,4c] = x | p1 | p2 | p3, as fp32.
 * dtype: 0 fp32, 1 bf16, 2 fp16 (x is rounded to the format first). */
ZL_API int32_t zl_test_sppf_pool(int32_t device, int32_t dtype, const float* x, int32_t n, int32_t h, int32_t w, int32_t c, float* cat);

ZL_API int32_t zl_probe_umma(int32_t device, int32_t N, int32_t swz, int32_t sbo_a, int32_t nacc, int32_t count,
                             int32_t shift_rows, int32_t ksteps, int32_t grid, int64_t* issue_cycles, int64_t* total_cycles);

/* Measurement hook: one 4-D tiled TMA load (optional element stride) dumped from shared memory. */
ZL_API int32_t zl_probe_tma(int32_t device, const uint16_t* x, int32_t n, int32_t h, int32_t w, int32_t c,
                            int32_t box_c, int32_t box_w, int32_t box_h, int32_t estride, int32_t swizzle_bytes,
                            int32_t c0, int32_t c1, int32_t c2, int32_t c3, uint32_t expect_bytes,
                            uint8_t* dump, uint32_t dump_bytes, int32_t* completed);

#ifdef __cplusplus
}
#endif
#endif /* ZL_B200_TEST_H_ */
