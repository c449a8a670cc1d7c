// head_math.cuh — the arithmetic of the Detect tail and of the candidate key, shared by every kernel
// that decodes the head (postprocess.cu: dfl_decode_kernel / decode_filter_kernel; head_fused.cu: head_decode_kernel) so
// that all of them produce the same bits for the same logits: ONE source expression per formula, fixed operation order.
#pragma once
#include <stdint.h>

#include <type_traits>

#include "common.h"

namespace zl {

// DFL softmax expectation over the 16 bins of one box side (ultralytics DFL == the reference graph's Softmax + 1x1 conv,
// inside Ort::Session::Run, src/inference/onnx_engine.cpp:577-585).  PRECISE = fp64 (exact mode), else fp32 fast math.
template <bool PRECISE>
__device__ __forceinline__ float dfl_expect(const float (&z)[16])
{
    using T = typename std::conditional<PRECISE, double, float>::type;
    float m = z[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) m = fmaxf(m, z[i]);
    T se = (T)0, sw = (T)0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const T e = PRECISE ? (T)exp((double)z[i] - (double)m) : (T)__expf(z[i] - m);
        se += e;
        sw += e * (T)i;
    }
    return PRECISE ? (float)(sw / se) : __fdividef((float)sw, (float)se);
}

// dist2bbox (xywh) x stride for the anchor at grid cell (x, y).
template <bool PRECISE>
__device__ __forceinline__ float4 dfl_box(float dl, float dt, float dr, float db, int x, int y, int stride)
{
    using T = typename std::conditional<PRECISE, double, float>::type;
    const T ax = (T)x + (T)0.5, ay = (T)y + (T)0.5, s = (T)stride;
    const T x1 = ax - (T)dl, y1 = ay - (T)dt, x2 = ax + (T)dr, y2 = ay + (T)db;
    return make_float4((float)((x1 + x2) * (T)0.5 * s), (float)((y1 + y2) * (T)0.5 * s), (float)((x2 - x1) * s), (float)((y2 - y1) * s));
}

template <bool PRECISE>
__device__ __forceinline__ float cls_score(float z)
{
    return PRECISE ? (float)(1.0 / (1.0 + exp(-(double)z))) : __fdividef(1.0f, 1.0f + __expf(-z));
}

// key = class[12] | (~confidence bits)[32] | anchor[20]: ascending key order == (class asc, confidence desc, anchor asc).
// Confidence is > 0 here, so its IEEE bit pattern is monotone.
__device__ __forceinline__ uint64_t make_key(int cls, float conf, int anchor) {
    return ((uint64_t)(uint32_t)cls << 52) | ((uint64_t)(~__float_as_uint(conf)) << kKeyAnchorBits) | (uint64_t)(uint32_t)anchor;
}
__device__ __forceinline__ int key_class(uint64_t k) { return (int)(k >> 52); }
__device__ __forceinline__ float key_conf(uint64_t k) { return __uint_as_float(~(uint32_t)(k >> kKeyAnchorBits)); }
__device__ __forceinline__ int key_anchor(uint64_t k) { return (int)(k & ((1u << kKeyAnchorBits) - 1)); }

}  // namespace zl
