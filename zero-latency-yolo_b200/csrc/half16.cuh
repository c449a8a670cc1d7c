// half16.cuh — the two 16-bit storage formats (bf16 / fp16) behind one runtime flag.
// Buffers are typed __nv_bfloat16* ("h16") throughout; when the engine runs in fp16 mode the same
// 16-bit slots hold IEEE half values and only these conversion helpers (and the UMMA instruction
// descriptor) change.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace zl {

__device__ __forceinline__ uint32_t pack2_16(float a, float b, bool f16) {
    if (f16) { __half2 h = __floats2half2_rn(a, b); return *reinterpret_cast<uint32_t*>(&h); }
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack2_16(uint32_t w, bool f16, float& a, float& b) {
    if (f16) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w)); a = f.x; b = f.y; return; }
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w);
    a = __bfloat162float(h.x); b = __bfloat162float(h.y);
}
__device__ __forceinline__ uint16_t pack1_16(float a, bool f16) {
    if (f16) { __half h = __float2half_rn(a); return *reinterpret_cast<uint16_t*>(&h); }
    __nv_bfloat16 h = __float2bfloat16_rn(a);
    return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ float unpack1_16(uint16_t v, bool f16) {
    if (f16) return __half2float(*reinterpret_cast<const __half*>(&v));
    return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&v));
}

// SiLU in the conv epilogues.  Two forms behind one runtime flag (HaloParams / conv_tc Params `silu_tanh`):
//   exp : x * rcp(1 + ex2(-x*log2e))                      two MUFU operations per element, ~2^-22 relative error
//   tanh: h + h*tanh.approx(h), h = x/2                    ONE MUFU operation per element; tanh.approx.f32 has a maximum
//         relative error of 2^-11, i.e. the result is good to the storage format (fp16: 2^-11, bf16: 2^-8) and no better.
// The SFU pipe issues 16 lanes per clock per SM: on the narrow high-resolution layers the exp form alone costs as much
// as the layer's MMAs.
__device__ __forceinline__ float silu_exp(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
__device__ __forceinline__ float silu_tanh(float v) {
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

}  // namespace zl
