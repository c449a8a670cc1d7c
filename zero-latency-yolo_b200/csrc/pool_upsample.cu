// pool_upsample.cu — SPPF max-pools and nearest 2x upsample (NHWC, channel-sliced views).
//
// Part of the graph the reference runs inside Ort::Session::Run
// (src/inference/onnx_engine.cpp:577-585; SURVEY.md Appendix A: SPPF = three
// chained 5x5/s1/p2 max-pools == 5x5, 9x9, 13x13 windows of the same input
// with -inf padding; Upsample = nearest, scale 2).  Both are pure HBM traffic:
// 16-byte vector loads/stores, results written straight into the concat buffer.
#include <cfloat>

#include "half16.cuh"
#include "kernels.h"

namespace zl {
namespace {

template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    using V = float4;
    __device__ static void load(const float* p, float (&f)[4]) { float4 v = __ldg(reinterpret_cast<const float4*>(p)); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
    __device__ static void store(float* p, const float (&f)[4]) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
};
template <bool F16> struct H16 { uint16_t v; };      // 16-bit element tagged with its format
template <bool F16> struct Vec<H16<F16>> {
    static constexpr int N = 8;
    __device__ static void load(const H16<F16>* p, float (&f)[8]) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) unpack2_16(w[i], F16, f[2 * i], f[2 * i + 1]);
    }
    __device__ static void store(H16<F16>* p, const float (&f)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = pack2_16(f[2 * i], f[2 * i + 1], F16);
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// One thread: one pixel x one channel vector; windows 5/9/13 share the scan.
template <typename T>
__global__ void __launch_bounds__(256)
sppf_pool_kernel(const T* __restrict__ a, T* __restrict__ p1, T* __restrict__ p2, T* __restrict__ p3,
                 int N, int H, int W, int C, int apitch, int p1pitch, int p2pitch, int p3pitch)
{
    constexpr int VN = Vec<T>::N;
    const int cv = C / VN;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)N * H * W * cv) return;
    const int c = (int)(idx % cv) * VN;
    const long long pix = idx / cv;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    float m5[VN], m9[VN], m13[VN];
#pragma unroll
    for (int i = 0; i < VN; ++i) m5[i] = m9[i] = m13[i] = -FLT_MAX;
    for (int dy = -6; dy <= 6; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -6; dx <= 6; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            float v[VN];
            Vec<T>::load(a + ((size_t)(n * H + yy) * W + xx) * apitch + c, v);
            const bool in9 = (dy >= -4 && dy <= 4 && dx >= -4 && dx <= 4);
            const bool in5 = (dy >= -2 && dy <= 2 && dx >= -2 && dx <= 2);
#pragma unroll
            for (int i = 0; i < VN; ++i) {
                m13[i] = fmaxf(m13[i], v[i]);
                if (in9) m9[i] = fmaxf(m9[i], v[i]);
                if (in5) m5[i] = fmaxf(m5[i], v[i]);
            }
        }
    }
    const size_t o = (size_t)pix;
    Vec<T>::store(p1 + o * p1pitch + c, m5);
    Vec<T>::store(p2 + o * p2pitch + c, m9);
    Vec<T>::store(p3 + o * p3pitch + c, m13);
}

// SPPF pools through shared memory: one CTA owns one image x 16 channels, keeps the whole (small) plane in smem
// and runs the three chained 5x5/s1/p2 max-pools separably (5 + 5 reads per output instead of a 13x13 scan).
template <typename T>
__global__ void __launch_bounds__(256)
sppf_pool_smem_kernel(const T* __restrict__ a, T* __restrict__ p1, T* __restrict__ p2, T* __restrict__ p3,
                      int H, int W, int apitch, int p1pitch, int p2pitch, int p3pitch)
{
    // the next kernel of the stream (a tcgen05 conv launched with programmatic stream serialization) may start its prologue
    // now; it still waits for this grid to complete (griddepcontrol.wait) before it touches this kernel's output
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr int VN = Vec<T>::N, CG = 16, VPP = CG / VN;      // vectors per pixel handled by this CTA
    extern __shared__ float pool_sm[];
    float* A = pool_sm;
    float* B = pool_sm + (size_t)H * W * CG;
    const int n = blockIdx.y, cg = blockIdx.x * CG, tid = threadIdx.x;
    const int npix = H * W;
    for (int idx = tid; idx < npix * VPP; idx += 256) {
        const int pix = idx / VPP, v = idx - pix * VPP;
        float f[VN];
        Vec<T>::load(a + ((size_t)n * npix + pix) * apitch + cg + v * VN, f);
#pragma unroll
        for (int i = 0; i < VN; ++i) A[pix * CG + v * VN + i] = f[i];
    }
    __syncthreads();
    T* outs[3] = {p1, p2, p3};
    const int pitches[3] = {p1pitch, p2pitch, p3pitch};
    for (int pass = 0; pass < 3; ++pass) {
        // (x, y) of a thread's pixel advance by 16 pixels per iteration without any division: this kernel was issue-bound
        // on its index arithmetic (ncu: issue slots 70 % busy, 56 us for 6.5 MB)
        int x = (tid >> 4) % W, y = (tid >> 4) / W;
        for (int idx = tid; idx < npix * CG; idx += 256) {          // row max: B = max over x-2..x+2 of A
            const int c = idx & (CG - 1);
            float m = -FLT_MAX;
#pragma unroll
            for (int d = -2; d <= 2; ++d) {
                const int xx = x + d;
                if (xx >= 0 && xx < W) m = fmaxf(m, A[(y * W + xx) * CG + c]);
            }
            B[idx] = m;
            x += 16;
            while (x >= W) { x -= W; ++y; }
        }
        __syncthreads();
        x = (tid >> 4) % W; y = (tid >> 4) / W;
        for (int idx = tid; idx < npix * CG; idx += 256) {          // column max: A = max over y-2..y+2 of B
            const int c = idx & (CG - 1);
            float m = -FLT_MAX;
#pragma unroll
            for (int d = -2; d <= 2; ++d) {
                const int yy = y + d;
                if (yy >= 0 && yy < H) m = fmaxf(m, B[(yy * W + x) * CG + c]);
            }
            A[idx] = m;
            x += 16;
            while (x >= W) { x -= W; ++y; }
        }
        __syncthreads();
        T* o = outs[pass];
        const int op = pitches[pass];
        for (int idx = tid; idx < npix * VPP; idx += 256) {
            const int pix = idx / VPP, v = idx - pix * VPP;
            float f[VN];
#pragma unroll
            for (int i = 0; i < VN; ++i) f[i] = A[pix * CG + v * VN + i];
            Vec<T>::store(o + ((size_t)n * npix + pix) * op + cg + v * VN, f);
        }
    }
}

// 16-bit variant of the kernel above: max() is exact in any format, so the planes stay PACKED (one uint4 = 8 channels of
// one pixel) and every step is a 128-bit shared-memory access plus four packed max instructions — a fifth of the
// instructions of the fp32-staged version (round-2 profile: 50 us for 6.5 MB in, 19.7 MB out; this form: HBM-bound).
template <bool F16>
__device__ __forceinline__ uint32_t max2_16(uint32_t a, uint32_t b)
{
    if (F16) { const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b)); return *reinterpret_cast<const uint32_t*>(&r); }
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}
template <bool F16>
__device__ __forceinline__ uint4 vmax16(const uint4 a, const uint4 b)
{
    return make_uint4(max2_16<F16>(a.x, b.x), max2_16<F16>(a.y, b.y), max2_16<F16>(a.z, b.z), max2_16<F16>(a.w, b.w));
}

template <bool F16>
__global__ void __launch_bounds__(256)
sppf_pool16_kernel(const uint16_t* __restrict__ a, uint16_t* __restrict__ p1, uint16_t* __restrict__ p2, uint16_t* __restrict__ p3,
                   int H, int W, int apitch, int p1pitch, int p2pitch, int p3pitch)
{
    // the next kernel of the stream (a tcgen05 conv launched with programmatic stream serialization) may start its prologue
    // now; it still waits for this grid to complete (griddepcontrol.wait) before it touches this kernel's output
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ uint4 pool16_sm[];
    const int npix = H * W, nv = npix * 2;                     // 16 channels per CTA = two vectors per pixel
    uint4* A = pool16_sm;
    uint4* B = pool16_sm + nv;
    const int n = blockIdx.y, cg = blockIdx.x * 16, tid = threadIdx.x;
    for (int idx = tid; idx < nv; idx += 256)
        A[idx] = __ldg(reinterpret_cast<const uint4*>(a + ((size_t)n * npix + (idx >> 1)) * apitch + cg + (idx & 1) * 8));
    __syncthreads();
    uint16_t* outs[3] = {p1, p2, p3};
    const int pitches[3] = {p1pitch, p2pitch, p3pitch};
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
        for (int idx = tid; idx < nv; idx += 256) {            // row max: B = max over x-2..x+2 of A
            const int pix = idx >> 1, x = pix % W;
            uint4 m = A[idx];
            if (x >= 2) m = vmax16<F16>(m, A[idx - 4]);
            if (x >= 1) m = vmax16<F16>(m, A[idx - 2]);
            if (x + 1 < W) m = vmax16<F16>(m, A[idx + 2]);
            if (x + 2 < W) m = vmax16<F16>(m, A[idx + 4]);
            B[idx] = m;
        }
        __syncthreads();
        uint16_t* o = outs[pass];
        const int op = pitches[pass];
        for (int idx = tid; idx < nv; idx += 256) {            // column max: A = max over y-2..y+2 of B, written out as it is produced
            const int pix = idx >> 1, y = pix / W;
            uint4 m = B[idx];
            if (y >= 2) m = vmax16<F16>(m, B[idx - 4 * W]);
            if (y >= 1) m = vmax16<F16>(m, B[idx - 2 * W]);
            if (y + 1 < H) m = vmax16<F16>(m, B[idx + 2 * W]);
            if (y + 2 < H) m = vmax16<F16>(m, B[idx + 4 * W]);
            A[idx] = m;
            *reinterpret_cast<uint4*>(o + ((size_t)n * npix + pix) * op + cg + (idx & 1) * 8) = m;
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int xpitch, int ypitch)
{
    // the next kernel of the stream (a tcgen05 conv launched with programmatic stream serialization) may start its prologue
    // now; it still waits for this grid to complete (griddepcontrol.wait) before it touches this kernel's output
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // one thread copies one source vector to its four destination pixels: 32-bit index math, one load, four stores
    constexpr int VN = Vec<T>::N;
    const int cv = C / VN;
    const int Wo = 2 * W;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;        // over (row = n*H + y, x, channel vector) of the SOURCE
    if (idx >= H * W * cv) return;
    const int n = blockIdx.y;
    const int c = (idx % cv) * VN;
    const int pix = idx / cv;
    const int sx = pix % W, sy = pix / W;
    float v[VN];
    Vec<T>::load(x + ((size_t)(n * H + sy) * W + sx) * xpitch + c, v);
    T* o = y + (((size_t)n * 2 * H + 2 * sy) * Wo + 2 * sx) * ypitch + c;
    Vec<T>::store(o, v);
    Vec<T>::store(o + ypitch, v);
    Vec<T>::store(o + (size_t)Wo * ypitch, v);
    Vec<T>::store(o + (size_t)Wo * ypitch + ypitch, v);
}

bool aligned16(const View& v) { return (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0 && (v.pitch * v.esize()) % 16 == 0; }

}  // namespace

int32_t launch_sppf_pool(cudaStream_t st, const View& a, const View& p1, const View& p2, const View& p3)
{
    const int vn = a.dtype == DT_F32 ? 4 : 8;
    if (a.c % vn || !aligned16(a) || !aligned16(p1) || !aligned16(p2) || !aligned16(p3))
        ZL_FAIL(ZL_INVALID_ARGUMENT, "sppf_pool: views must be 16-B aligned");
    const size_t smem16 = (size_t)a.h * a.w * 2 * sizeof(uint4) * 2;
    if (a.dtype != DT_F32 && (a.c % 16) == 0 && smem16 <= 200 * 1024) {
        static thread_local int last_dev16 = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev != last_dev16) {
            ZL_CUDA(cudaFuncSetAttribute(sppf_pool16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ZL_CUDA(cudaFuncSetAttribute(sppf_pool16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            last_dev16 = dev;
        }
        dim3 g(a.c / 16, a.n);
        if (a.dtype == DT_F16)
            sppf_pool16_kernel<true><<<g, 256, smem16, st>>>((const uint16_t*)a.ptr, (uint16_t*)p1.ptr, (uint16_t*)p2.ptr, (uint16_t*)p3.ptr, a.h, a.w,
                                                              a.pitch, p1.pitch, p2.pitch, p3.pitch);
        else
            sppf_pool16_kernel<false><<<g, 256, smem16, st>>>((const uint16_t*)a.ptr, (uint16_t*)p1.ptr, (uint16_t*)p2.ptr, (uint16_t*)p3.ptr, a.h, a.w,
                                                               a.pitch, p1.pitch, p2.pitch, p3.pitch);
        ZL_CUDA(cudaGetLastError());
        return ZL_OK;
    }
    const size_t smem = (size_t)a.h * a.w * 16 * sizeof(float) * 2;
    if ((a.c % 16) == 0 && smem <= 200 * 1024) {
        static thread_local int last_dev = -1;
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev != last_dev) {
            ZL_CUDA(cudaFuncSetAttribute(sppf_pool_smem_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ZL_CUDA(cudaFuncSetAttribute(sppf_pool_smem_kernel<H16<true>>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ZL_CUDA(cudaFuncSetAttribute(sppf_pool_smem_kernel<H16<false>>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            last_dev = dev;
        }
        dim3 g(a.c / 16, a.n);
        if (a.dtype == DT_F32)
            sppf_pool_smem_kernel<float><<<g, 256, smem, st>>>((const float*)a.ptr, (float*)p1.ptr, (float*)p2.ptr, (float*)p3.ptr, a.h, a.w,
                                                                a.pitch, p1.pitch, p2.pitch, p3.pitch);
        else if (a.dtype == DT_F16)
            sppf_pool_smem_kernel<H16<true>><<<g, 256, smem, st>>>((const H16<true>*)a.ptr, (H16<true>*)p1.ptr, (H16<true>*)p2.ptr, (H16<true>*)p3.ptr,
                                                                    a.h, a.w, a.pitch, p1.pitch, p2.pitch, p3.pitch);
        else
            sppf_pool_smem_kernel<H16<false>><<<g, 256, smem, st>>>((const H16<false>*)a.ptr, (H16<false>*)p1.ptr, (H16<false>*)p2.ptr, (H16<false>*)p3.ptr,
                                                                     a.h, a.w, a.pitch, p1.pitch, p2.pitch, p3.pitch);
        ZL_CUDA(cudaGetLastError());
        return ZL_OK;
    }
    const long long total = (long long)a.pixels() * (a.c / vn);
    const int grid = (int)((total + 255) / 256);
    if (a.dtype == DT_F32)
        sppf_pool_kernel<float><<<grid, 256, 0, st>>>((const float*)a.ptr, (float*)p1.ptr, (float*)p2.ptr, (float*)p3.ptr,
                                                      a.n, a.h, a.w, a.c, a.pitch, p1.pitch, p2.pitch, p3.pitch);
    else if (a.dtype == DT_F16)
        sppf_pool_kernel<H16<true>><<<grid, 256, 0, st>>>((const H16<true>*)a.ptr, (H16<true>*)p1.ptr, (H16<true>*)p2.ptr,
                                                          (H16<true>*)p3.ptr, a.n, a.h, a.w, a.c, a.pitch, p1.pitch, p2.pitch, p3.pitch);
    else
        sppf_pool_kernel<H16<false>><<<grid, 256, 0, st>>>((const H16<false>*)a.ptr, (H16<false>*)p1.ptr, (H16<false>*)p2.ptr,
                                                           (H16<false>*)p3.ptr, a.n, a.h, a.w, a.c, a.pitch, p1.pitch, p2.pitch, p3.pitch);
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

int32_t launch_upsample2x(cudaStream_t st, const View& x, const View& y)
{
    const int vn = x.dtype == DT_F32 ? 4 : 8;
    if (x.c % vn || y.h != 2 * x.h || y.w != 2 * x.w || y.c != x.c || !aligned16(x) || !aligned16(y))
        ZL_FAIL(ZL_INVALID_ARGUMENT, "upsample2x: view mismatch");
    dim3 grid(ceil_div(x.h * x.w * (x.c / vn), 256), x.n);
    if (x.dtype == DT_F32)
        upsample2x_kernel<float><<<grid, 256, 0, st>>>((const float*)x.ptr, (float*)y.ptr, x.n, x.h, x.w, x.c, x.pitch, y.pitch);
    else
        upsample2x_kernel<H16<false>><<<grid, 256, 0, st>>>((const H16<false>*)x.ptr, (H16<false>*)y.ptr, x.n, x.h, x.w, x.c, x.pitch, y.pitch);   // pure copy: format-agnostic
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

}  // namespace zl
