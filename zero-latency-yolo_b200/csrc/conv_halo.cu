// conv_halo.cu — persistent 3x3 / stride-1 convolution on tcgen05 with halo reuse (sm_100a only).
//
// Same reference nodes as conv_tc.cu (Conv+BN+SiLU inside Ort::Session::Run,
// src/inference/onnx_engine.cpp:577-585), but for the layers that dominate HBM
// traffic — 3x3 stride-1 convs on the high-resolution maps — the nine filter
// taps are NOT gathered nine times.  Per output tile of 8 (w) x 16 (h) pixels:
//   * ONE 4-D TMA load per channel chunk brings the (8+2) x (16+2) input patch
//     into shared memory (OOB zero fill == conv padding, image borders, ragged
//     tiles), written in the 32/64/128-byte swizzle the tensor core expects;
//   * the A operand of tap (r,s) is the SAME patch read through a UMMA
//     descriptor whose start address is shifted by (r*10+s) pixels and whose
//     8-row groups (one tile row = 8 pixels) are strided by the patch row pitch
//     (SBO = 10 pixels).  The swizzle XOR is a function of the absolute smem
//     address, so a shifted view stays consistent with what TMA wrote;
//   * all 9 x Cin/16 MMAs of the tile accumulate into one of two TMEM buffers
//     while the four epilogue warps drain the other one.
// The CTA is persistent (grid = #SMs) and keeps the whole weight tensor
// (9 x Cout x Cin 16-bit values) resident in shared memory, so per tile only the
// patch (1.4x the tile's input) is read and the output written once.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "half16.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace zl {
namespace {

using namespace tc;

constexpr int kTW = 8, kTH = 16;            // output tile
constexpr int kPW = kTW + 2, kPH = kTH + 2;  // input patch
constexpr int kThreads = 192;
constexpr int kMaxPatchStages = 12;

struct HaloParams {
    void* y;
    const __nv_bfloat16* res;
    const float* bias;
    int32_t N, H, W, Cin, Cout, ntile;
    int32_t ypitch, rpitch, y_f32, y_vec, r_vec, f16, act;
    int32_t kc, cchunks, stages;
    int32_t tiles_x, tiles_y, num_tiles;
    uint32_t wtile_bytes, wtile_alloc, patch_bytes, patch_alloc, tmem_cols;
};

__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x, const HaloParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_pfull = base;                               // kMaxPatchStages x 8
    const uint32_t bar_pempty = base + 8u * kMaxPatchStages;       // kMaxPatchStages x 8
    const uint32_t bar_wfull = base + 16u * kMaxPatchStages;
    const uint32_t bar_tfull = bar_wfull + 8u;                     // 2 x 8
    const uint32_t bar_tempty = bar_tfull + 16u;                   // 2 x 8
    const uint32_t tmem_slot = bar_tempty + 16u;
    const uint32_t wbase = base + 1024u;
    const uint32_t nwt = 9u * (uint32_t)p.cchunks;
    const uint32_t pbase = wbase + nwt * p.wtile_alloc;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int stages = p.stages;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bar_pfull + 8u * s, 1u);
            mbar_init(bar_pempty + 8u * s, 1u);
        }
        mbar_init(bar_wfull, 1u);
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8u * a, 1u);
            mbar_init(bar_tempty + 8u * a, 4u);          // one arrival per epilogue warp
        }
        fence_barrier_init();
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        // ===== TMA producer: weights once, then one patch per (tile, channel chunk) =====
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_wfull, nwt * p.wtile_bytes);
            for (uint32_t t = 0; t < nwt; ++t) {
                const int tap = (int)t / p.cchunks, cc = (int)t - tap * p.cchunks;
                tma_load_2d(&tmap_w, bar_wfull, wbase + t * p.wtile_alloc, tap * p.Cin + cc * p.kc, 0);
            }
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int n = tile / tiles_per_img;
                const int rem = tile - n * tiles_per_img;
                const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
                for (int cc = 0; cc < p.cchunks; ++cc, ++it) {
                    const uint32_t s = it % (uint32_t)stages, ph = (it / (uint32_t)stages) & 1u;
                    mbar_wait(bar_pempty + 8u * s, ph ^ 1u);
                    mbar_arrive_expect_tx(bar_pfull + 8u * s, p.patch_bytes);
                    tma_load_4d(&tmap_x, bar_pfull + 8u * s, pbase + s * p.patch_alloc, cc * p.kc, tx * kTW - 1, ty * kTH - 1, n);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t swz = (uint32_t)p.kc * 2u;
        const uint32_t fmt = p.f16 ? 0u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.ntile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int ksteps = p.kc / 16;
        const uint32_t sbo_a = (uint32_t)kPW * swz;          // one tile row (8 pixels) per 8-row group, groups strided by the patch row
        mbar_wait(bar_wfull, 0u);
        tc_fence_after();
        uint32_t it = 0, tl = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
            const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
            mbar_wait(bar_tempty + 8u * acc, aph ^ 1u);       // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * (p.tmem_cols >> 1);
            for (int cc = 0; cc < p.cchunks; ++cc, ++it) {
                const uint32_t s = it % (uint32_t)stages, ph = (it / (uint32_t)stages) & 1u;
                mbar_wait(bar_pfull + 8u * s, ph);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t patch = pbase + s * p.patch_alloc;
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        const int r = tap / 3, sft = tap - r * 3;
                        const uint64_t adesc = make_smem_desc_sbo(patch + (uint32_t)(r * kPW + sft) * swz, swz, sbo_a);
                        const uint64_t bdesc = make_smem_desc_sbo(wbase + (uint32_t)(tap * p.cchunks + cc) * p.wtile_alloc, swz, 8u * swz);
                        for (int k = 0; k < ksteps; ++k)
                            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (cc | tap | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(bar_pempty + 8u * s);
                    if (cc == p.cchunks - 1) umma_commit(bar_tfull + 8u * acc);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue warps 2..5 =====
        const uint32_t q = warp & 3u;
        const int row = (int)(q * 32u + lane);               // A row == TMEM lane: h = row / 8, w = row % 8
        const int th = row >> 3, tw = row & 7;
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
            const uint32_t acc = tl & 1u, aph = (tl >> 1) & 1u;
            const int n = tile / tiles_per_img;
            const int rem = tile - n * tiles_per_img;
            const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
            const int oy = ty * kTH + th, ox = tx * kTW + tw;
            const bool ok = oy < p.H && ox < p.W;
            const size_t m = ((size_t)n * p.H + oy) * p.W + ox;
            mbar_wait(bar_tfull + 8u * acc, aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * (p.tmem_cols >> 1);
            for (int c0 = 0; c0 < p.ntile; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + (uint32_t)c0, v);
                tmem_ld_wait();
                if (!ok || c0 >= p.Cout) continue;
                float f[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float a = __uint_as_float(v[i]) + __ldg(p.bias + c0 + i);
                    f[i] = p.act ? silu(a) : a;
                }
                const bool full = (c0 + 16 <= p.Cout);
                if (p.res != nullptr) {
                    const __nv_bfloat16* rp = p.res + m * p.rpitch + c0;
                    if (full && p.r_vec) {
                        const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rp));
                        const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
                        const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float a, b;
                            unpack2_16(rw[i], p.f16, a, b);
                            f[2 * i] += a;
                            f[2 * i + 1] += b;
                        }
                    } else {
                        for (int i = 0; i < 16 && c0 + i < p.Cout; ++i) f[i] += unpack1_16(reinterpret_cast<const uint16_t*>(rp)[i], p.f16);
                    }
                }
                if (p.y_f32) {
                    float* yp = reinterpret_cast<float*>(p.y) + m * p.ypitch + c0;
                    if (full && p.y_vec) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            reinterpret_cast<float4*>(yp)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                    } else {
                        for (int i = 0; i < 16 && c0 + i < p.Cout; ++i) yp[i] = f[i];
                    }
                } else {
                    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + m * p.ypitch + c0;
                    if (full && p.y_vec) {
                        uint32_t w[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) w[i] = pack2_16(f[2 * i], f[2 * i + 1], p.f16);
                        reinterpret_cast<uint4*>(yp)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                        reinterpret_cast<uint4*>(yp)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    } else {
                        for (int i = 0; i < 16 && c0 + i < p.Cout; ++i) reinterpret_cast<uint16_t*>(yp)[i] = pack1_16(f[i], p.f16);
                    }
                }
            }
            // this warp is done reading the accumulator: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8u * acc);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int32_t make_tmap_nhwc(CUtensorMap* map, const View& x, int kc, int swz, bool f16)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            ZL_FAIL(ZL_SYSTEM_ERROR, "cuTensorMapEncodeTiled entry point not available");
        fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    cuuint64_t dims[4] = {(cuuint64_t)x.c, (cuuint64_t)x.w, (cuuint64_t)x.h, (cuuint64_t)x.n};
    cuuint64_t strides[3] = {(cuuint64_t)x.pitch * 2, (cuuint64_t)x.w * x.pitch * 2, (cuuint64_t)x.h * x.w * x.pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)kPW, (cuuint32_t)kPH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x.ptr, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) ZL_FAIL(ZL_SYSTEM_ERROR, "cuTensorMapEncodeTiled(4d) failed, CUresult " + std::to_string((int)r));
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && (uint64_t)x.n * strides[2] < 131072ull) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);   // see make_tmap_2d_16
    return ZL_OK;
}

}  // namespace

bool conv_halo_supported(const ConvWeights& w, const View& x, const View& y, int* smem_out)
{
    if (w.k != 3 || w.stride != 1 || !x.is16() || (w.cin % 16) != 0 || w.cout_pad > 128) return false;
    if ((x.pitch % 8) != 0 || (reinterpret_cast<uintptr_t>(x.ptr) & 15)) return false;
    if (y.is16() && y.dtype != x.dtype) return false;
    const int kc = (w.cin % 64 == 0) ? 64 : (w.cin % 32 == 0 ? 32 : 16);
    const int cchunks = w.cin / kc;
    const uint32_t wtile_alloc = ((uint32_t)w.cout_pad * kc * 2 + 1023u) & ~1023u;
    const uint32_t patch_alloc = ((uint32_t)kPW * kPH * kc * 2 + 1023u) & ~1023u;
    const uint32_t fixed = 2048u + 9u * cchunks * wtile_alloc;
    const uint32_t budget = 227u * 1024u;
    if (fixed + (uint32_t)(cchunks + 1) * patch_alloc > budget) return false;      // need more than one tile's patches in flight
    if (smem_out) {
        int stages = (int)((budget - fixed) / patch_alloc);
        if (stages > kMaxPatchStages) stages = kMaxPatchStages;
        if (stages > 3 * cchunks) stages = 3 * cchunks;
        *smem_out = (int)(fixed + (uint32_t)stages * patch_alloc);
    }
    return true;
}

int32_t conv_halo_prepare(const ConvWeights& w, const View& x, const View& y, const View* res, ConvHaloOp* op)
{
    int smem = 0;
    if (!conv_halo_supported(w, x, y, &smem)) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_halo: layer not supported (" + w.name + ")");
    if (y.h != x.h || y.w != x.w || y.n != x.n || y.c != w.cout || x.c != w.cin) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_halo: view mismatch (" + w.name + ")");
    if (res && (res->dtype != x.dtype || res->c != w.cout)) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_halo: residual view mismatch");
    ConvHaloOp& o = *op;
    o.y = y.ptr;
    o.res = res ? reinterpret_cast<const __nv_bfloat16*>(res->ptr) : nullptr;
    o.bias = w.bias;
    o.N = x.n; o.H = x.h; o.W = x.w; o.Cin = w.cin; o.Cout = w.cout; o.ntile = w.cout_pad;
    o.ypitch = y.pitch; o.rpitch = res ? res->pitch : 0;
    o.y_f32 = y.dtype == DT_F32;
    o.f16 = x.dtype == DT_F16 ? 1 : 0;
    o.act = w.act;
    o.y_vec = ((y.pitch * y.esize()) % 16 == 0 && (reinterpret_cast<uintptr_t>(y.ptr) & 15) == 0) ? 1 : 0;
    o.r_vec = (res && (res->pitch % 8) == 0 && (reinterpret_cast<uintptr_t>(res->ptr) & 15) == 0) ? 1 : 0;
    o.kc = (w.cin % 64 == 0) ? 64 : (w.cin % 32 == 0 ? 32 : 16);
    o.cchunks = w.cin / o.kc;
    o.tiles_x = ceil_div(x.w, kTW); o.tiles_y = ceil_div(x.h, kTH);
    o.num_tiles = o.tiles_x * o.tiles_y * x.n;
    o.wtile_bytes = (uint32_t)w.cout_pad * o.kc * 2;
    o.wtile_alloc = (o.wtile_bytes + 1023u) & ~1023u;
    o.patch_bytes = (uint32_t)kPW * kPH * o.kc * 2;
    o.patch_alloc = (o.patch_bytes + 1023u) & ~1023u;
    o.smem_bytes = smem;
    o.stages = (int)((smem - 2048 - 9 * o.cchunks * (int)o.wtile_alloc) / (int)o.patch_alloc);
    int cols = 32;
    while (cols < 2 * w.cout_pad) cols <<= 1;
    o.tmem_cols = cols;
    ZL_TRY(make_tmap_2d_16(&o.tmap_w, w.w_tc, (uint64_t)w.ktot, (uint64_t)w.cout_pad, (uint64_t)w.ktot * 2, o.kc, w.cout_pad, o.kc * 2, o.f16));
    ZL_TRY(make_tmap_nhwc(&o.tmap_x, x, o.kc, o.kc * 2, o.f16));
    o.flops = 2.0 * (double)y.pixels() * w.cout * w.ktot;
    o.bytes = (double)x.pixels() * w.cin * 2 + (double)y.pixels() * w.cout * (o.y_f32 ? 4 : 2) + (double)w.cout * w.ktot * 2 +
              (res ? (double)y.pixels() * w.cout * 2 : 0.0);
    return ZL_OK;
}

int32_t conv_halo_launch(cudaStream_t st, const ConvHaloOp& o, int num_sms)
{
    static thread_local int last_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != last_dev) {
        ZL_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        last_dev = dev;
    }
    HaloParams p;
    p.y = o.y; p.res = o.res; p.bias = o.bias;
    p.N = o.N; p.H = o.H; p.W = o.W; p.Cin = o.Cin; p.Cout = o.Cout; p.ntile = o.ntile;
    p.ypitch = o.ypitch; p.rpitch = o.rpitch; p.y_f32 = o.y_f32; p.y_vec = o.y_vec; p.r_vec = o.r_vec; p.f16 = o.f16; p.act = o.act;
    p.kc = o.kc; p.cchunks = o.cchunks; p.stages = o.stages;
    p.tiles_x = o.tiles_x; p.tiles_y = o.tiles_y; p.num_tiles = o.num_tiles;
    p.wtile_bytes = o.wtile_bytes; p.wtile_alloc = o.wtile_alloc; p.patch_bytes = o.patch_bytes; p.patch_alloc = o.patch_alloc;
    p.tmem_cols = o.tmem_cols;
    const int grid = o.num_tiles < num_sms ? o.num_tiles : num_sms;
    conv_halo_kernel<<<grid, kThreads, o.smem_bytes, st>>>(o.tmap_w, o.tmap_x, p);
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

}  // namespace zl
