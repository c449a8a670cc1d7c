// conv_simt.cu — CUDA-core convolutions.
//
// (1) conv_simt_kernel: fp32 implicit GEMM used by the exact ("fp32") mode, the
//     mode that must match the reference's ONNX Runtime CPU session within 1e-3
//     abs on raw head outputs (BASELINE.json north_star).  fp32 operands and
//     fp32 activation storage like the reference, but products are accumulated
//     in fp64 and the bias/SiLU/residual epilogue runs in fp64 before the single
//     rounding to fp32: two fp32-accumulating implementations of this 25-layer
//     stack differ by ~3e-3 px on the box rows from summation order alone
//     (measured: torch-CPU fp32 vs this kernel with fp32 accumulators), so the
//     1e-3 gate is only meaningful against a summation-order-free result.
//     Replaces the same Ort::Session::Run nodes as conv_tc.cu (onnx_engine.cpp:577-585).
// (2) conv0_direct_kernel: the bf16 path's first layer (3->c1, 3x3 s2), whose
//     K = 27 is too small for an MMA tile and which is purely HBM-bound.
#include "half16.cuh"
#include "kernels.h"

namespace zl {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;

struct SimtParams {
    const float* x; const float* w; const float* bias; const float* res; float* y;
    int H, W, Cin, xpitch, Ho, Wo, Cout, cout_pad, ypitch, rpitch;
    int k, stride, pad, act, ktot, m_total, vec;
};

__global__ void __launch_bounds__(256)
conv_simt_kernel(const SimtParams p)
{
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN];
    const int t = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    // A-load role: one row, 4 consecutive k
    const int arow = t >> 2, akq = (t & 3) * 4;
    const int am = m0 + arow;
    const bool arow_ok = am < p.m_total;
    int an = 0, aoy = 0, aox = 0;
    if (arow_ok) {
        an = am / (p.Ho * p.Wo);
        const int rem = am - an * (p.Ho * p.Wo);
        aoy = rem / p.Wo;
        aox = rem - aoy * p.Wo;
    }
    const int iy0 = aoy * p.stride - p.pad, ix0 = aox * p.stride - p.pad;
    // B-load role: one k, 4 consecutive n
    const int bk = t >> 4, bn4 = (t & 15) * 4;
    // compute role
    const int tx = t & 15, ty = t >> 4;
    double acc[4][4] = {};

    for (int kk = 0; kk < p.ktot; kk += BK) {
        float av[4] = {0.f, 0.f, 0.f, 0.f};
        const int kbase = kk + akq;
        if (arow_ok && kbase < p.ktot) {
            if (p.vec) {
                const int tap = kbase / p.Cin, c = kbase - tap * p.Cin;
                const int r = tap / p.k, s = tap - r * p.k;
                const int iy = iy0 + r, ix = ix0 + s;
                if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(p.x + ((size_t)(an * p.H + iy) * p.W + ix) * p.xpitch + c));
                    av[0] = v.x; av[1] = v.y; av[2] = v.z; av[3] = v.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int kidx = kbase + i;
                    if (kidx < p.ktot) {
                        const int tap = kidx / p.Cin, c = kidx - tap * p.Cin;
                        const int r = tap / p.k, s = tap - r * p.k;
                        const int iy = iy0 + r, ix = ix0 + s;
                        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
                            av[i] = __ldg(p.x + ((size_t)(an * p.H + iy) * p.W + ix) * p.xpitch + c);
                    }
                }
            }
        }
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kk + bk < p.ktot && n0 + bn4 < p.cout_pad)
            bv = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)(kk + bk) * p.cout_pad + n0 + bn4));
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) As[akq + i][arow] = av[i];
        *reinterpret_cast<float4*>(&Bs[bk][bn4]) = bv;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float aa[4] = {a.x, a.y, a.z, a.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma((double)aa[i], (double)bb[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= p.m_total) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = n0 + tx * 4 + j;
            if (c >= p.Cout) continue;
            double v = acc[i][j] + (double)__ldg(p.bias + c);
            if (p.act) v = v / (1.0 + exp(-v));
            if (p.res) v += (double)__ldg(p.res + (size_t)m * p.rpitch + c);
            p.y[(size_t)m * p.ypitch + c] = (float)v;
        }
    }
}

// First layer of the 16-bit path: x = [n,H,W,4] 16-bit (R,G,B,0), 3x3 stride 2 pad 1.
// One thread = 2 consecutive output pixels x 16 channels (each weight float4 from smem feeds 8 FMAs; 4 pixels
// per thread was measured slower: 168 registers, 12 warps/SM), inputs are 8-byte pixel loads.  FMA order per output = k ascending from the bias.
__global__ void __launch_bounds__(128, 5)
conv0_direct_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ y, int N, int H, int W, int Ho, int Wo, int Cout, int cout_pad, int ypitch, int f16)
{
    extern __shared__ float ws[];            // [27][cout_pad] + bias[cout_pad]
    for (int i = threadIdx.x; i < 27 * cout_pad; i += blockDim.x) ws[i] = w[i];
    for (int i = threadIdx.x; i < cout_pad; i += blockDim.x) ws[27 * cout_pad + i] = bias[i];
    __syncthreads();
    const int wq = (Wo + 1) >> 1;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= wq * Ho) return;
    const int n = blockIdx.y;
    const int oy = item / wq, ox0 = (item - oy * wq) * 2;
    __nv_bfloat16* yp = y + (((size_t)n * Ho + oy) * Wo + ox0) * ypitch;
    for (int c0 = 0; c0 < cout_pad; c0 += 16) {
        float acc[2][16];
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[px][j] = ws[27 * cout_pad + c0 + j];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int iy = 2 * oy - 1 + r;
            float in[15];                       // 5 input columns x (R,G,B)
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const int ix = 2 * ox0 - 1 + j;
                float a = 0.f, b = 0.f, c = 0.f;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
                    const uint2 v = __ldg(reinterpret_cast<const uint2*>(x + ((size_t)(n * H + iy) * W + ix) * 4));
                    float pad_;
                    unpack2_16(v.x, f16, a, b);
                    unpack2_16(v.y, f16, c, pad_);
                }
                in[3 * j + 0] = a; in[3 * j + 1] = b; in[3 * j + 2] = c;
            }
#pragma unroll
            for (int s = 0; s < 3; ++s) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const int k = (r * 3 + s) * 3 + ch;
                    const float4* wr = reinterpret_cast<const float4*>(&ws[k * cout_pad + c0]);
                    const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
                    const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
                    for (int px = 0; px < 2; ++px) {
                        const float a = in[3 * (2 * px + s) + ch];
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[px][j] = fmaf(a, wv[j], acc[px][j]);
                    }
                }
            }
        }
#pragma unroll
        for (int px = 0; px < 2; ++px) {
            if (ox0 + px >= Wo) break;
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a = acc[px][2 * j], b = acc[px][2 * j + 1];
                pk[j] = pack2_16(__fdividef(a, 1.0f + __expf(-a)), __fdividef(b, 1.0f + __expf(-b)), f16);
            }
            __nv_bfloat16* o = yp + (size_t)px * ypitch + c0;
            if (c0 + 16 <= Cout) {
                reinterpret_cast<uint4*>(o)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                reinterpret_cast<uint4*>(o)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else {
                for (int j = 0; j < 16 && c0 + j < Cout; ++j) o[j] = reinterpret_cast<const __nv_bfloat16*>(pk)[j];
            }
        }
    }
}

// P1 + layer 0 fused (16-bit modes): nearest-stretch sampling of the u8 BGR frame (exactly preProcess,
// onnx_engine.cpp:649-700), the 16-bit rounding of the preprocessed value (through a 256-entry table, so the
// numbers equal the unfused path bit for bit) and the 3x3/s2 first conv + bias + SiLU, without ever writing the
// preprocessed image.  One thread = 4 consecutive output pixels x 16 output channels: the frame bytes are read
// once from L1/L2, weights come from smem as float4 broadcasts (16 FMAs per LDS.128), 128 B stored per thread.
__global__ void __launch_bounds__(128)
pre_conv0_kernel(const uint8_t* __restrict__ staging, const FrameDesc* __restrict__ descs, const float* __restrict__ w,
                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ y, int mw, int mh, int Ho, int Wo,
                 int Cout, int cout_pad, int ypitch, int f16)
{
    extern __shared__ float ws[];            // [27][cout_pad] weights | bias[cout_pad] | lut[256]
    float* bs = ws + 27 * cout_pad;
    float* lut = bs + cout_pad;
    for (int i = threadIdx.x; i < 27 * cout_pad; i += blockDim.x) ws[i] = w[i];
    for (int i = threadIdx.x; i < cout_pad; i += blockDim.x) bs[i] = bias[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = unpack1_16(pack1_16(__fdiv_rn((float)i, 255.0f), f16), f16);
    __syncthreads();
    const int wq = (Wo + 3) >> 2;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= wq * Ho) return;
    const int f = blockIdx.y;
    const int oy = item / wq, ox0 = (item - oy * wq) * 4;
    const FrameDesc d = descs[f];
    if (d.w <= 0 || d.h <= 0) return;
    const uint8_t* __restrict__ img = staging + d.offset;
    const float scale_w = __fdiv_rn((float)d.w, (float)mw);
    const float scale_h = __fdiv_rn((float)d.h, (float)mh);
    int sx[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const int mx = 2 * ox0 - 1 + j;
        sx[j] = (mx >= 0 && mx < mw) ? min(__float2int_rz(__fmul_rn((float)mx, scale_w)), d.w - 1) * 3 : -1;
    }
    __nv_bfloat16* yp = y + (((size_t)f * Ho + oy) * Wo + ox0) * ypitch;
    for (int c0 = 0; c0 < cout_pad; c0 += 16) {
        float acc[4][16];
#pragma unroll
        for (int px = 0; px < 4; ++px)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[px][j] = bs[c0 + j];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int my = 2 * oy - 1 + r;
            float in[27];                       // 9 columns x (R,G,B) of this input row
            if (my >= 0 && my < mh) {
                const int sy = min(__float2int_rz(__fmul_rn((float)my, scale_h)), d.h - 1);
                const uint8_t* __restrict__ row = img + (size_t)sy * d.w * 3;
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    if (sx[j] >= 0) {
                        in[3 * j + 0] = lut[__ldg(row + sx[j] + 2)];     // R: the reference reads byte 2-c for channel c
                        in[3 * j + 1] = lut[__ldg(row + sx[j] + 1)];
                        in[3 * j + 2] = lut[__ldg(row + sx[j] + 0)];
                    } else {
                        in[3 * j + 0] = in[3 * j + 1] = in[3 * j + 2] = 0.0f;
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 27; ++j) in[j] = 0.0f;
            }
#pragma unroll
            for (int s = 0; s < 3; ++s) {
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const int k = (r * 3 + s) * 3 + ch;
                    const float4* wr = reinterpret_cast<const float4*>(&ws[k * cout_pad + c0]);
                    const float4 w0 = wr[0], w1 = wr[1], w2 = wr[2], w3 = wr[3];
                    const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
                    for (int px = 0; px < 4; ++px) {
                        const float a = in[3 * (2 * px + s) + ch];
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[px][j] = fmaf(a, wv[j], acc[px][j]);
                    }
                }
            }
        }
#pragma unroll
        for (int px = 0; px < 4; ++px) {
            if (ox0 + px >= Wo) break;
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a = acc[px][2 * j], b = acc[px][2 * j + 1];
                pk[j] = pack2_16(__fdividef(a, 1.0f + __expf(-a)), __fdividef(b, 1.0f + __expf(-b)), f16);
            }
            __nv_bfloat16* o = yp + (size_t)px * ypitch + c0;
            if (c0 + 16 <= Cout) {
                reinterpret_cast<uint4*>(o)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                reinterpret_cast<uint4*>(o)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else {
                for (int j = 0; j < 16 && c0 + j < Cout; ++j) o[j] = reinterpret_cast<const __nv_bfloat16*>(pk)[j];
            }
        }
    }
}

}  // namespace

int32_t launch_pre_conv0(cudaStream_t st, const uint8_t* staging, const FrameDesc* descs, int32_t n, int32_t mw, int32_t mh,
                         const ConvWeights& w, const View& y)
{
    if (w.cin != 3 || w.k != 3 || w.stride != 2 || !y.is16() || (y.pitch % 8) != 0 || (mw & 1) || (mh & 1))
        ZL_FAIL(ZL_INVALID_ARGUMENT, "pre_conv0: expects the 3->c 3x3 s2 first layer with 16-bit output");
    const int Ho = mh / 2, Wo = mw / 2;
    if (y.h != Ho || y.w != Wo || y.c != w.cout) ZL_FAIL(ZL_INVALID_ARGUMENT, "pre_conv0: output view mismatch");
    const size_t smem = ((size_t)28 * w.cout_pad + 256) * sizeof(float);
    dim3 grid(ceil_div(ceil_div(Wo, 4) * Ho, 128), n);
    pre_conv0_kernel<<<grid, 128, smem, st>>>(staging, descs, w.w_simt, w.bias, (__nv_bfloat16*)y.ptr, mw, mh, Ho, Wo, w.cout, w.cout_pad, y.pitch,
                                              y.dtype == DT_F16 ? 1 : 0);
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

int32_t launch_conv_simt(cudaStream_t st, const ConvWeights& w, const View& x, const View& y, const View* res)
{
    if (x.dtype != DT_F32 || y.dtype != DT_F32 || (res && res->dtype != DT_F32))
        ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_simt: fp32 views required");
    const int pad = w.k / 2;
    const int Ho = (x.h + 2 * pad - w.k) / w.stride + 1, Wo = (x.w + 2 * pad - w.k) / w.stride + 1;
    if (x.c != w.cin || y.h != Ho || y.w != Wo || y.n != x.n || y.c != w.cout)
        ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_simt: view mismatch (" + w.name + ")");
    SimtParams p;
    p.x = (const float*)x.ptr; p.w = w.w_simt; p.bias = w.bias; p.res = res ? (const float*)res->ptr : nullptr; p.y = (float*)y.ptr;
    p.H = x.h; p.W = x.w; p.Cin = w.cin; p.xpitch = x.pitch; p.Ho = Ho; p.Wo = Wo; p.Cout = w.cout; p.cout_pad = w.cout_pad;
    p.ypitch = y.pitch; p.rpitch = res ? res->pitch : 0;
    p.k = w.k; p.stride = w.stride; p.pad = pad; p.act = w.act; p.ktot = w.ktot; p.m_total = x.n * Ho * Wo;
    p.vec = (w.cin % 4 == 0 && x.pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(x.ptr) & 15) == 0) ? 1 : 0;
    dim3 grid(ceil_div(p.m_total, BM), ceil_div(w.cout, BN));
    conv_simt_kernel<<<grid, 256, 0, st>>>(p);
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

int32_t launch_conv0_direct(cudaStream_t st, const ConvWeights& w, const View& x, const View& y)
{
    if (w.cin != 3 || w.k != 3 || w.stride != 2 || x.pitch != 4 || !x.is16() || y.dtype != x.dtype)
        ZL_FAIL(ZL_INVALID_ARGUMENT, "conv0_direct: expects the 3->c 3x3 s2 first layer on NHWC4 bf16");
    const int Ho = (x.h + 2 - 3) / 2 + 1, Wo = (x.w + 2 - 3) / 2 + 1;
    if (y.h != Ho || y.w != Wo || y.c != w.cout || (y.pitch % 8) != 0) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv0_direct: output view mismatch");
    const size_t smem = (size_t)28 * w.cout_pad * sizeof(float);
    dim3 grid(ceil_div(ceil_div(Wo, 2) * Ho, 128), x.n);
    conv0_direct_kernel<<<grid, 128, smem, st>>>((const __nv_bfloat16*)x.ptr, w.w_simt, w.bias, (__nv_bfloat16*)y.ptr,
                                                 x.n, x.h, x.w, Ho, Wo, w.cout, w.cout_pad, y.pitch, x.dtype == DT_F16 ? 1 : 0);
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

}  // namespace zl
