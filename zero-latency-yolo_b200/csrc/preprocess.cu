// preprocess.cu — P1: fused resize + BGR->RGB + /255 + layout change, one pass.
//
// Replaces OnnxInferenceEngine::preProcess (reference:
// src/inference/onnx_engine.cpp:649-700; identical body at :703-755).  Parity
// mode reproduces the reference exactly: nearest-neighbour STRETCH with
//   scale = float(src)/dst (fp32 divide), src_idx = min(int(i*scale), src-1)
// (fp32 multiply, truncation), channel read at 2-c, value/255.0f (fp32 divide,
// round-to-nearest).  No FMA contraction is possible in these expressions but
// the intrinsics pin the rounding anyway.
//
// HBM-bound: each thread produces 4 consecutive output pixels of one row and
// issues one 16/32-byte store (NHWC4) or three 16-byte stores (NCHW planes).
#include "half16.cuh"
#include "kernels.h"

namespace zl {
namespace {

__device__ __forceinline__ int src_index(int i, float scale, int src_dim) {
    int s = __float2int_rz(__fmul_rn((float)i, scale));
    return min(s, src_dim - 1);
}

// Where model pixel (x, y) samples the frame.  Parity mode: the reference's independent stretch of both axes.
// Letterbox: one gain, centred, border = 114 (returns false: the pixel is padding).
struct Sampler {
    float scale_w, scale_h;
    int w, h, pad_x, pad_y, nw, nh;
    bool letterbox;
    __device__ __forceinline__ Sampler(const FrameDesc& d, int mw, int mh, bool lb) : w(d.w), h(d.h), letterbox(lb) {
        if (!lb) {
            scale_w = __fdiv_rn((float)d.w, (float)mw);
            scale_h = __fdiv_rn((float)d.h, (float)mh);
            pad_x = pad_y = 0; nw = mw; nh = mh;
        } else {
            const LetterboxMap m = letterbox_map(d.w, d.h, mw, mh);
            pad_x = m.pad_x; pad_y = m.pad_y; nw = m.nw; nh = m.nh;
            scale_w = __fdiv_rn((float)d.w, (float)m.nw);
            scale_h = __fdiv_rn((float)d.h, (float)m.nh);
        }
    }
    __device__ __forceinline__ bool row(int y, int* sy) const {
        if (letterbox && (y < pad_y || y >= pad_y + nh)) return false;
        *sy = src_index(y - pad_y, scale_h, h);
        return true;
    }
    __device__ __forceinline__ bool col(int x, int* sx) const {
        if (letterbox && (x < pad_x || x >= pad_x + nw)) return false;
        *sx = src_index(x - pad_x, scale_w, w);
        return true;
    }
};
constexpr float kLetterboxFill = 114.0f;

template <int LAYOUT>
__global__ void __launch_bounds__(256)
preprocess_kernel(const uint8_t* __restrict__ staging, const FrameDesc* __restrict__ descs,
                  int mw, int mh, void* __restrict__ out, int letterbox)
{
    const int f = blockIdx.y;
    const int wq = (mw + 3) >> 2;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= wq * mh) return;
    const int y = item / wq;
    const int x0 = (item - y * wq) * 4;
    const FrameDesc d = descs[f];
    if (d.w <= 0 || d.h <= 0) return;                 // unused batch slot
    const uint8_t* __restrict__ img = staging + d.offset;
    const Sampler sm(d, mw, mh, letterbox != 0);
    int sy = 0;
    const bool row_ok = sm.row(y, &sy);
    const uint8_t* __restrict__ row = img + (size_t)sy * d.w * 3;

    float r[4], g[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = min(x0 + i, mw - 1);
        int sx = 0;
        if (row_ok && sm.col(x, &sx)) {
            const uint8_t* px = row + (size_t)sx * 3;
            // source is BGR; the reference reads channel 2-c for output channel c
            b[i] = __fdiv_rn((float)__ldg(px + 0), 255.0f);
            g[i] = __fdiv_rn((float)__ldg(px + 1), 255.0f);
            r[i] = __fdiv_rn((float)__ldg(px + 2), 255.0f);
        } else {
            b[i] = g[i] = r[i] = __fdiv_rn(kLetterboxFill, 255.0f);
        }
    }
    const int nvalid = min(4, mw - x0);
    if (LAYOUT == PRE_NCHW_F32) {
        float* o = reinterpret_cast<float*>(out) + (size_t)f * 3 * mh * mw + (size_t)y * mw + x0;
        const size_t plane = (size_t)mh * mw;
        if (nvalid == 4 && (mw & 3) == 0) {
            *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
            *reinterpret_cast<float4*>(o + plane) = make_float4(g[0], g[1], g[2], g[3]);
            *reinterpret_cast<float4*>(o + 2 * plane) = make_float4(b[0], b[1], b[2], b[3]);
        } else {
            for (int i = 0; i < nvalid; ++i) { o[i] = r[i]; o[plane + i] = g[i]; o[2 * plane + i] = b[i]; }
        }
    } else if (LAYOUT == PRE_NHWC4_F32) {
        float4* o = reinterpret_cast<float4*>(out) + ((size_t)f * mh + y) * mw + x0;
        for (int i = 0; i < nvalid; ++i) o[i] = make_float4(r[i], g[i], b[i], 0.0f);
    } else {
        uint2* o = reinterpret_cast<uint2*>(out) + ((size_t)f * mh + y) * mw + x0;
        uint2 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[i] = make_uint2(pack2_16(r[i], g[i], LAYOUT == PRE_NHWC4_F16), pack2_16(b[i], 0.0f, LAYOUT == PRE_NHWC4_F16));
        }
        if (nvalid == 4 && (mw & 3) == 0) {
            reinterpret_cast<uint4*>(o)[0] = make_uint4(v[0].x, v[0].y, v[1].x, v[1].y);
            reinterpret_cast<uint4*>(o)[1] = make_uint4(v[2].x, v[2].y, v[3].x, v[3].y);
        } else {
            for (int i = 0; i < nvalid; ++i) o[i] = v[i];
        }
    }
}

// Space-to-depth variant for the tensor-core first layer: one thread = one 2x2 block of model pixels -> 16 halves
// (k = (dy*2+dx)*3 + channel, channels R,G,B; 12..15 = 0), one 32-byte store.  Same sampling and rounding as above.
template <bool F16>
__global__ void __launch_bounds__(256)
preprocess_s2d_kernel(const uint8_t* __restrict__ staging, const FrameDesc* __restrict__ descs, int mw, int mh, uint4* __restrict__ out, int letterbox)
{
    // the next kernel of the stream (a tcgen05 conv launched with programmatic stream serialization) may start its prologue
    // now; it still waits for this grid to complete (griddepcontrol.wait) before it touches this kernel's output
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int f = blockIdx.y;
    const int W2 = mw >> 1, H2 = mh >> 1;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= W2 * H2) return;
    const FrameDesc d = descs[f];
    if (d.w <= 0 || d.h <= 0) return;
    const int Y = item / W2, X = item - Y * W2;
    const uint8_t* __restrict__ img = staging + d.offset;
    const Sampler sm(d, mw, mh, letterbox != 0);
    float v[12];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        int sy = 0;
        const bool row_ok = sm.row(2 * Y + dy, &sy);
        const uint8_t* __restrict__ row = img + (size_t)sy * d.w * 3;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            int sx = 0;
            if (row_ok && sm.col(2 * X + dx, &sx)) {
                const uint8_t* px = row + (size_t)sx * 3;
                v[(dy * 2 + dx) * 3 + 0] = __fdiv_rn((float)__ldg(px + 2), 255.0f);     // R = byte 2
                v[(dy * 2 + dx) * 3 + 1] = __fdiv_rn((float)__ldg(px + 1), 255.0f);
                v[(dy * 2 + dx) * 3 + 2] = __fdiv_rn((float)__ldg(px + 0), 255.0f);
            } else {
                v[(dy * 2 + dx) * 3 + 0] = v[(dy * 2 + dx) * 3 + 1] = v[(dy * 2 + dx) * 3 + 2] = __fdiv_rn(kLetterboxFill, 255.0f);
            }
        }
    }
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 6; ++i) w[i] = pack2_16(v[2 * i], v[2 * i + 1], F16);
    w[6] = 0u; w[7] = 0u;
    uint4* o = out + ((size_t)f * H2 * W2 + item) * 2;
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// Fast path of the space-to-depth variant for frames that already have the model's size (the reference's default client
// sends 416x416 to a 416x416 model; the throughput configs send 640x640): no resampling, so a thread takes FOUR 2x2 blocks
// of one block row = 24 contiguous bytes from each of two frame rows (three 8-byte loads per row instead of 24 single-byte
// loads) and turns bytes into 16-bit values through a 256-entry table of the correctly rounded value/255.0f (built with
// __fdiv_rn, so the result is bit-identical to the general kernel's).  One thread = 128 contiguous output bytes.
template <bool F16>
__global__ void __launch_bounds__(256)
preprocess_s2d_same_size_kernel(const uint8_t* __restrict__ staging, const FrameDesc* __restrict__ descs, int mw, int mh, uint4* __restrict__ out)
{
    // the next kernel of the stream (a tcgen05 conv launched with programmatic stream serialization) may start its prologue
    // now; it still waits for this grid to complete (griddepcontrol.wait) before it touches this kernel's output
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ uint16_t lut[256];
    {
        const float q = __fdiv_rn((float)threadIdx.x, 255.0f);
        lut[threadIdx.x] = pack1_16(q, F16);
    }
    __syncthreads();
    const int f = blockIdx.y;
    const int W2 = mw >> 1, H2 = mh >> 1, WQ = W2 >> 2;                // WQ groups of four blocks per block row
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= WQ * H2) return;
    const FrameDesc d = descs[f];
    const int Y = item / WQ, XQ = item - Y * WQ;
    const uint8_t* __restrict__ img = staging + d.offset;
    uint32_t b[2][6];                                                   // two rows x 24 bytes
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const uint2* rp = reinterpret_cast<const uint2*>(img + ((size_t)(2 * Y + dy) * mw + (size_t)XQ * 8) * 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) { const uint2 t = __ldg(rp + k); b[dy][2 * k] = t.x; b[dy][2 * k + 1] = t.y; }
    }
    auto byte_at = [&](int dy, int i) -> uint32_t { return (b[dy][i >> 2] >> ((i & 3) * 8)) & 0xffu; };
    // A thread owns 128 contiguous output bytes; stored directly, a warp-wide 16-byte store would touch 32 different
    // 128-byte lines.  The eight 16-byte chunks go through shared memory (XOR-swizzled: conflict-free both ways) so that
    // every store instruction of a warp writes 512 contiguous bytes.
    __shared__ uint4 tr[8][32 * 8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint4* tw = tr[warp];
#pragma unroll
    for (int j = 0; j < 4; ++j) {                                       // block j of the four: pixels 2j, 2j+1 of both rows
        uint32_t v[12];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int px = (2 * j + dx) * 3;                         // BGR in the frame, RGB out (channel c reads byte 2 - c)
                v[(dy * 2 + dx) * 3 + 0] = lut[byte_at(dy, px + 2)];
                v[(dy * 2 + dx) * 3 + 1] = lut[byte_at(dy, px + 1)];
                v[(dy * 2 + dx) * 3 + 2] = lut[byte_at(dy, px + 0)];
            }
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 6; ++i) w[i] = v[2 * i] | (v[2 * i + 1] << 16);
        w[6] = 0u; w[7] = 0u;
        tw[lane * 8 + ((2 * j) ^ (lane & 7))] = make_uint4(w[0], w[1], w[2], w[3]);
        tw[lane * 8 + ((2 * j + 1) ^ (lane & 7))] = make_uint4(w[4], w[5], w[6], w[7]);
    }
    __syncwarp();
    // items of a warp are consecutive (the grid is sized so that whole warps are in range or out), so its 4 KB are contiguous
    const int item0 = item - lane;
    uint4* o = out + ((size_t)f * H2 * W2 + (size_t)item0 * 4) * 2;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = k * 32 + lane, t = c >> 3, j = c & 7;              // chunk c of the warp belongs to thread t, its j-th chunk
        o[c] = tw[t * 8 + (j ^ (t & 7))];
    }
}

// The kept boxes of a letterboxed batch, mapped back to the request frame: the filter normalised model-pixel boxes by the
// frame's width / height like the reference does; undo that, remove the padding and the gain, normalise again.
__global__ void letterbox_unmap_kernel(int n, int maxn, const uint32_t* __restrict__ header, DevDet* __restrict__ dets,
                                       const FrameDesc* __restrict__ descs, int mw, int mh, uint32_t cap)
{
    const int f = blockIdx.x;
    if (f >= n) return;
    const uint32_t cnt = header[4 + f], off = header[4 + maxn + f];
    const FrameDesc d = descs[f];
    const LetterboxMap m = letterbox_map(d.w, d.h, mw, mh);
    const float fw = (float)d.w, fh = (float)d.h;
    for (uint32_t i = threadIdx.x; i < cnt && off + i < cap; i += blockDim.x) {
        DevDet t = dets[off + i];
        t.x = ((t.x * fw - (float)m.pad_x) * m.inv_gain) / fw;
        t.y = ((t.y * fh - (float)m.pad_y) * m.inv_gain) / fh;
        t.w = (t.w * fw * m.inv_gain) / fw;
        t.h = (t.h * fh * m.inv_gain) / fh;
        dets[off + i] = t;
    }
}

}  // namespace

int32_t launch_letterbox_unmap(cudaStream_t st, int32_t n, int32_t maxn, const uint32_t* header, DevDet* dets, const FrameDesc* descs,
                               int32_t mw, int32_t mh, uint32_t cap)
{
    if (n <= 0) return ZL_OK;
    letterbox_unmap_kernel<<<n, 128, 0, st>>>(n, maxn, header, dets, descs, mw, mh, cap);
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

int32_t launch_preprocess(cudaStream_t st, const uint8_t* staging, const FrameDesc* descs, int32_t n,
                          int32_t mw, int32_t mh, int32_t layout, void* out, int32_t letterbox, int32_t same_size)
{
    if (n <= 0) return ZL_OK;
    if (layout == PRE_S2D16_BF16 || layout == PRE_S2D16_F16) {
        if ((mw & 1) || (mh & 1)) ZL_FAIL(ZL_INVALID_ARGUMENT, "s2d preprocess needs even model dims");
        if (same_size && !letterbox && (mw % 8) == 0 && ((mw / 8) * (mh / 2)) % 32 == 0) {      // every frame of the batch already has the model's size
            dim3 gs(ceil_div((mw / 8) * (mh / 2), 256), n);
            if (layout == PRE_S2D16_F16) preprocess_s2d_same_size_kernel<true><<<gs, 256, 0, st>>>(staging, descs, mw, mh, (uint4*)out);
            else preprocess_s2d_same_size_kernel<false><<<gs, 256, 0, st>>>(staging, descs, mw, mh, (uint4*)out);
            ZL_CUDA(cudaGetLastError());
            return ZL_OK;
        }
        dim3 g(ceil_div((mw / 2) * (mh / 2), 256), n);
        if (layout == PRE_S2D16_F16) preprocess_s2d_kernel<true><<<g, 256, 0, st>>>(staging, descs, mw, mh, (uint4*)out, letterbox);
        else preprocess_s2d_kernel<false><<<g, 256, 0, st>>>(staging, descs, mw, mh, (uint4*)out, letterbox);
        ZL_CUDA(cudaGetLastError());
        return ZL_OK;
    }
    const int threads = 256;
    dim3 grid(ceil_div(ceil_div(mw, 4) * mh, threads), n);
    switch (layout) {
        case PRE_NCHW_F32: preprocess_kernel<PRE_NCHW_F32><<<grid, threads, 0, st>>>(staging, descs, mw, mh, out, letterbox); break;
        case PRE_NHWC4_F32: preprocess_kernel<PRE_NHWC4_F32><<<grid, threads, 0, st>>>(staging, descs, mw, mh, out, letterbox); break;
        case PRE_NHWC4_BF16: preprocess_kernel<PRE_NHWC4_BF16><<<grid, threads, 0, st>>>(staging, descs, mw, mh, out, letterbox); break;
        case PRE_NHWC4_F16: preprocess_kernel<PRE_NHWC4_F16><<<grid, threads, 0, st>>>(staging, descs, mw, mh, out, letterbox); break;
        default: ZL_FAIL(ZL_INVALID_ARGUMENT, "bad preprocess layout");
    }
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

}  // namespace zl
