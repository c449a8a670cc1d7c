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

// First layer of the bf16 path: x = [n,H,W,4] bf16 (R,G,B,0), 3x3 stride 2 pad 1.
__global__ void __launch_bounds__(128)
conv0_direct_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ y, int N, int H, int W, int Ho, int Wo, int Cout, int cout_pad, int ypitch, int f16)
{
    extern __shared__ float ws[];            // [27][cout_pad] + bias[cout_pad]
    for (int i = threadIdx.x; i < 27 * cout_pad; i += blockDim.x) ws[i] = w[i];
    for (int i = threadIdx.x; i < cout_pad; i += blockDim.x) ws[27 * cout_pad + i] = bias[i];
    __syncthreads();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= N * Ho * Wo) return;
    const int n = m / (Ho * Wo);
    const int rem = m - n * (Ho * Wo);
    const int oy = rem / Wo, ox = rem - oy * Wo;
    float in[27];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int iy = oy * 2 - 1 + r, ix = ox * 2 - 1 + s;
            float a = 0.f, b = 0.f, c = 0.f;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(x + ((size_t)(n * H + iy) * W + ix) * 4));
                float pad_;
                unpack2_16(v.x, f16, a, b);
                unpack2_16(v.y, f16, c, pad_);
            }
            in[(r * 3 + s) * 3 + 0] = a; in[(r * 3 + s) * 3 + 1] = b; in[(r * 3 + s) * 3 + 2] = c;
        }
    }
    __nv_bfloat16* yp = y + (size_t)m * ypitch;
    for (int c0 = 0; c0 < cout_pad; c0 += 16) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = ws[27 * cout_pad + c0 + j];
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            const float4* wr = reinterpret_cast<const float4*>(&ws[k * cout_pad + c0]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 wv = wr[q];
                acc[4 * q + 0] = fmaf(in[k], wv.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(in[k], wv.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(in[k], wv.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(in[k], wv.w, acc[4 * q + 3]);
            }
        }
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float a = acc[2 * j], b = acc[2 * j + 1];
            pk[j] = pack2_16(a / (1.0f + __expf(-a)), b / (1.0f + __expf(-b)), f16);
        }
        if (c0 + 16 <= Cout) {
            reinterpret_cast<uint4*>(yp + c0)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            reinterpret_cast<uint4*>(yp + c0)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else {
            for (int j = 0; j < 16 && c0 + j < Cout; ++j)
                yp[c0 + j] = reinterpret_cast<const __nv_bfloat16*>(pk)[j];
        }
    }
}

}  // namespace

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
    const int m_total = x.n * Ho * Wo;
    const size_t smem = (size_t)28 * w.cout_pad * sizeof(float);
    conv0_direct_kernel<<<ceil_div(m_total, 128), 128, smem, st>>>((const __nv_bfloat16*)x.ptr, w.w_simt, w.bias, (__nv_bfloat16*)y.ptr,
                                                                  x.n, x.h, x.w, Ho, Wo, w.cout, w.cout_pad, y.pitch, x.dtype == DT_F16 ? 1 : 0);
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

}  // namespace zl
