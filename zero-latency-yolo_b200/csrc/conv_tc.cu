// conv_tc.cu — implicit-GEMM convolution on tcgen05 tensor cores (sm_100a only).
//
// Replaces the Conv+BN+SiLU nodes ONNX Runtime executes inside
// Ort::Session::Run (reference: src/inference/onnx_engine.cpp:577-585; graph
// restated in SURVEY.md Appendix A).  BN is folded, so each node is
//     y = act(W (*) x + b) [+ residual], written at a channel offset of a wider
// NHWC buffer (concat fusion).
//
// GEMM view: M = N*Ho*Wo output pixels (128 per CTA), N = Cout (<=256 per CTA),
// K = k*k*Cin walked in K-blocks of (one filter tap) x (kc = 16/32/64 channels).
//   warp 0      : TMA producer — weights tile [ntile][kc] (K-major, 32/64/128B
//                 swizzle) and, for 1x1 convs, the activation tile [128][kc].
//   warp 1      : allocates TMEM, issues tcgen05.mma (M=128, N=ntile, K=16) from
//                 one lane, commits stage-empty / accumulator-full mbarriers.
//   warps 2..5  : for 3x3 convs gather the shifted NHWC rows of the current tap
//                 into the same swizzled UMMA layout (zero-filled padding), then
//                 run the epilogue: tcgen05.ld -> +bias -> SiLU -> +residual ->
//                 bf16/fp32 NHWC store.
// Accumulators live in TMEM (fp32), never in registers.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include "half16.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace zl {

namespace {

constexpr int kTileM = 128;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;

struct Params {
    const __nv_bfloat16* x;
    void* y;
    const __nv_bfloat16* res;
    const float* bias;
    int32_t H, W, Cin, xpitch;
    int32_t Ho, Wo, Cout, ypitch, rpitch, y_f32;
    int32_t k, stride, pad, act;
    int32_t kc, nkb, cchunks;
    int32_t ntile, m_total, a_tma, stages, y_vec, r_vec, f16, silu_tanh;
    uint32_t a_bytes, b_bytes, stage_bytes, tmem_cols;
};

using namespace tc;

// ---------------------------------------------------------------- kernel
__global__ void __launch_bounds__(kThreads)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_a, const Params p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;

    // carve: [barriers | tmem ptr] in the first 1 KB-aligned block, then the stage ring
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_full = base;                       // kMaxStages x 8 B
    const uint32_t bar_empty = base + 8u * kMaxStages;    // kMaxStages x 8 B
    const uint32_t bar_accum = base + 16u * kMaxStages;
    const uint32_t tmem_slot = bar_accum + 8u;
    const uint32_t tiles = base + 1024u;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int m0 = blockIdx.x * kTileM;
    const int n0 = blockIdx.y * p.ntile;
    const int stages = p.stages;

    if (threadIdx.x == 0) {
        const uint32_t full_count = p.a_tma ? 1u : 1u + 128u;
        for (int s = 0; s < stages; ++s) {
            mbar_init(bar_full + 8u * s, full_count);
            mbar_init(bar_empty + 8u * s, 1u);
        }
        mbar_init(bar_accum, 1u);
        fence_barrier_init();
        tma_prefetch_desc(&tmap_w);
        if (p.a_tma) tma_prefetch_desc(&tmap_a);
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, p.tmem_cols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // programmatic dependent launch: the next kernel may start its prologue under this one's tail; nothing the previous
    // kernel produced is touched before griddepcontrol.wait (weights and bias are constants)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ===== TMA producer =====
        // K-blocks run channel-chunk-major, tap-minor — the same accumulation order as the persistent kernel
        // (conv_halo.cu), so a layer gives bit-identical results whichever kernel the batch size selects.
        if (elect_one()) {
            const uint32_t tx = p.b_bytes + (p.a_tma ? p.a_bytes : 0u);
            const int taps = p.k * p.k;
            uint32_t s = 0, ph = 0;
            for (int cc = 0; cc < p.cchunks; ++cc) {
                for (int tap = 0; tap < taps; ++tap) {
                    mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                    const uint32_t sa = tiles + s * p.stage_bytes;
                    const uint32_t sb = sa + p.a_bytes;
                    mbar_arrive_expect_tx(bar_full + 8u * s, tx);
                    tma_load_2d(&tmap_w, bar_full + 8u * s, sb, tap * p.Cin + cc * p.kc, n0);
                    if (p.a_tma) tma_load_2d(&tmap_a, bar_full + 8u * s, sa, cc * p.kc, m0);
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t swz = (uint32_t)p.kc * 2u;
        // kind::f16 instruction descriptor: D=f32, A=B=bf16 (format 1) or fp16 (format 0), K-major both, N>>3, M>>4
        const uint32_t fmt = p.f16 ? 0u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.ntile >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        const int ksteps = p.kc / 16;
        const uint64_t adesc0 = make_smem_desc(tiles, swz);
        const uint64_t bdesc0 = make_smem_desc(tiles + p.a_bytes, swz);
        const uint32_t stage16 = p.stage_bytes >> 4;
        uint32_t s = 0, ph = 0;
        for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(bar_full + 8u * s, ph);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t adesc = adesc0 + (uint64_t)(s * stage16);
                const uint64_t bdesc = bdesc0 + (uint64_t)(s * stage16);
                // advance 16 elements (32 B) along K inside the swizzle atom: +2 in 16-B units
                if (ksteps == 4) {
                    umma_bf16(tmem_base, adesc, bdesc, idesc, kb != 0 ? 1u : 0u);
                    umma_bf16(tmem_base, adesc + 2, bdesc + 2, idesc, 1u);
                    umma_bf16(tmem_base, adesc + 4, bdesc + 4, idesc, 1u);
                    umma_bf16(tmem_base, adesc + 6, bdesc + 6, idesc, 1u);
                } else if (ksteps == 2) {
                    umma_bf16(tmem_base, adesc, bdesc, idesc, kb != 0 ? 1u : 0u);
                    umma_bf16(tmem_base, adesc + 2, bdesc + 2, idesc, 1u);
                } else {
                    umma_bf16(tmem_base, adesc, bdesc, idesc, kb != 0 ? 1u : 0u);
                }
                umma_commit(bar_empty + 8u * s);
                if (kb == p.nkb - 1) umma_commit(bar_accum);
            }
            __syncwarp();
            if (++s == (uint32_t)stages) { s = 0; ph ^= 1u; }
        }
    } else {
        // ===== warps 2..5: A gather (3x3 / strided) then epilogue =====
        const int t = (int)threadIdx.x - 64;   // 0..127 : row of the A tile this thread fills
        if (!p.a_tma) {
            const int m = m0 + t;
            const bool row_ok = m < p.m_total;
            int n_img = 0, oy = 0, ox = 0;
            if (row_ok) {
                n_img = m / (p.Ho * p.Wo);
                const int rem = m - n_img * (p.Ho * p.Wo);
                oy = rem / p.Wo;
                ox = rem - oy * p.Wo;
            }
            const int iy0 = oy * p.stride - p.pad, ix0 = ox * p.stride - p.pad;
            const int nchunk = p.kc / 8;                       // 16-B chunks per row: 2, 4 or 8
            const uint32_t rowbytes = (uint32_t)p.kc * 2u;
            // software image of the hardware swizzle (Swizzle<B,4,3> on byte addresses, tiles 1 KB aligned)
            const uint32_t xr = nchunk == 8 ? ((uint32_t)t & 7u) : (nchunk == 4 ? (((uint32_t)t >> 1) & 3u) : (((uint32_t)t >> 2) & 1u));
            const int taps = p.k * p.k;
            uint32_t s = 0, ph = 0;
            for (int cc = 0; cc < p.cchunks; ++cc) {
                int r = 0, sft = 0;
                for (int tap = 0; tap < taps; ++tap) {
                    const int iy = iy0 + r, ix = ix0 + sft;
                    const bool ok = row_ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                    uint4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = make_uint4(0u, 0u, 0u, 0u);
                    if (ok) {
                        const uint4* src = reinterpret_cast<const uint4*>(p.x + ((size_t)(n_img * p.H + iy) * p.W + ix) * p.xpitch + cc * p.kc);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j < nchunk) v[j] = __ldg(src + j);
                    }
                    mbar_wait(bar_empty + 8u * s, ph ^ 1u);
                    const uint32_t rowbase = tiles + s * p.stage_bytes + (uint32_t)t * rowbytes;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < nchunk) {
                            const uint32_t dst = rowbase + ((((uint32_t)j) ^ xr) << 4);
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(v[j].x), "r"(v[j].y), "r"(v[j].z), "r"(v[j].w) : "memory");
                        }
                    }
                    fence_proxy_async();                 // generic-proxy writes -> visible to the tensor core (async proxy)
                    mbar_arrive(bar_full + 8u * s);
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1u; }
                    if (++sft == p.k) { sft = 0; ++r; }
                }
            }
        }
        // ----- epilogue -----
        mbar_wait(bar_accum, 0u);
        tc_fence_after();
        const uint32_t q = warp & 3u;                      // TMEM lane quarter this warp may read
        const int row = (int)(q * 32u + lane);
        const int m = m0 + row;
        const bool m_ok = m < p.m_total;
        const uint32_t taddr = tmem_base + ((q * 32u) << 16);
        for (int c0 = 0; c0 < p.ntile; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + (uint32_t)c0, v);
            tmem_ld_wait();
            if (!m_ok) continue;
            const int cg = n0 + c0;                        // first output channel of this group
            if (cg >= p.Cout) continue;
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float a = __uint_as_float(v[i]) + __ldg(p.bias + cg + i);
                f[i] = p.act ? (p.silu_tanh ? silu_tanh(a) : silu_exp(a)) : a;      // same formula as conv_halo.cu: results do not depend on the kernel chosen
            }
            const bool full = (cg + 16 <= p.Cout);
            if (p.res != nullptr) {
                const __nv_bfloat16* rp = p.res + (size_t)m * p.rpitch + cg;
                if (full && p.r_vec) {
                    uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rp));
                    uint4 r1 = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
                    const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float a, b;
                        unpack2_16(rw[i], p.f16, a, b);
                        f[2 * i] += a;
                        f[2 * i + 1] += b;
                    }
                } else {
                    for (int i = 0; i < 16 && cg + i < p.Cout; ++i) f[i] += unpack1_16(reinterpret_cast<const uint16_t*>(rp)[i], p.f16);
                }
            }
            if (p.y_f32) {
                float* yp = reinterpret_cast<float*>(p.y) + (size_t)m * p.ypitch + cg;
                if (full && p.y_vec) {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        reinterpret_cast<float4*>(yp)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                } else {
                    for (int i = 0; i < 16 && cg + i < p.Cout; ++i) yp[i] = f[i];
                }
            } else {
                __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + (size_t)m * p.ypitch + cg;
                if (full && p.y_vec) {
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = pack2_16(f[2 * i], f[2 * i + 1], p.f16);
                    reinterpret_cast<uint4*>(yp)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                    reinterpret_cast<uint4*>(yp)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                } else {
                    for (int i = 0; i < 16 && cg + i < p.Cout; ++i) reinterpret_cast<uint16_t*>(yp)[i] = pack1_16(f[i], p.f16);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

}  // namespace

// Channels per K chunk of a 16-bit conv.  A function of the LAYER only (never of the batch or the tile count): both
// tcgen05 kernels accumulate in (chunk, tap, 16-wide k-step) order, so a frame's result does not depend on how it was
// batched or on which kernel / launch plan ran (tests: test_16bit_results_do_not_depend_on_batch).
//   * the widest of 64 / 32 / 16 that divides Cin (128 / 64 / 32-byte swizzle rows);
//   * 3x3 stride 2: at most 32 (a stage holds four parity sub-patches);
//   * 3x3: halved until two pipeline stages of the weight-STREAMING plan fit at an output width of min(Cout, 128)
//     (conv_halo.cu; an MMA gets no cheaper per column beyond N = 128), so the deep layers never fall back to a narrow N split.
int32_t conv_kc(const ConvWeights& w, bool y_f32)
{
    int kc = (w.cin % 64 == 0) ? 64 : (w.cin % 32 == 0 ? 32 : 16);
    if (w.k == 3 && w.stride == 2 && kc == 64) kc = 32;
    if (w.k == 3 && w.cin >= 16 && w.cin % 16 == 0) {
        const int s2 = w.stride == 2;
        const uint32_t fixed = 3072u + 16u * (y_f32 ? 4096u : 2048u);
        const int nt = w.cout_pad < 128 ? w.cout_pad : 128;     // an MMA gets no cheaper per column beyond N = 128 (umma_probe)
        while (kc > 16) {
            const uint32_t patch = (((uint32_t)(s2 ? 9 : 10) * (16 + (s2 ? 1 : 2)) * kc * 2 + 1023u) & ~1023u) * (s2 ? 4u : 1u);
            const uint32_t wchunk = (9u * nt * kc * 2 + 1023u) & ~1023u;
            const bool resident_fits = fixed + (uint32_t)(w.cin / kc) * wchunk + 3u * patch <= 227u * 1024u;
            if (resident_fits || fixed + 2u * (patch + wchunk) <= 227u * 1024u) break;
            kc /= 2;
        }
    }
    return kc;
}

// Which SiLU form the 16-bit conv epilogues use (half16.cuh).  ZL_SILU=exp|tanh overrides; read once per process.
bool conv_silu_tanh(bool f16)
{
    static const int forced = [] {
        const char* e = getenv("ZL_SILU");
        if (!e) return -1;
        return (e[0] == 't' || e[0] == 'T' || e[0] == '1') ? 1 : 0;
    }();
    if (forced >= 0) return forced == 1;
    return !f16 ? kSiluTanhDefaultBf16 : kSiluTanhDefaultF16;
}

int32_t make_tmap_2d_16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer, int32_t swizzle_bytes, bool f16)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) ZL_FAIL(ZL_SYSTEM_ERROR, "cuTensorMapEncodeTiled entry point not available");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (outer_stride_bytes & 15))
        ZL_FAIL(ZL_INVALID_ARGUMENT, "TMA base / stride must be 16-byte aligned");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {outer_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32  ? CU_TENSOR_MAP_SWIZZLE_32B
                                                 : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) ZL_FAIL(ZL_SYSTEM_ERROR, "cuTensorMapEncodeTiled failed, CUresult " + std::to_string((int)r));
    // Drivers up to 13.1 mis-encode maps of tensors smaller than 128 KiB; CUTLASS
    // (cute/atom/copy_traits_sm90_tma.hpp) clears this descriptor bit for them.
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && outer * outer_stride_bytes < 131072ull)
        reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
    return ZL_OK;
}

int32_t conv_tc_prepare(const ConvWeights& w, const View& x, const View& y, const View* res,
                        bool allow_tma_a, int32_t ntile_hint, ConvTcOp* op)
{
    if (!x.is16()) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_tc: input must be bf16 or fp16");
    if (y.is16() && y.dtype != x.dtype) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_tc: 16-bit output must use the input format");
    if (x.c != w.cin || (w.cin % 16) != 0) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_tc: Cin must be a multiple of 16 (" + w.name + ")");
    if ((x.pitch % 8) != 0 || (reinterpret_cast<uintptr_t>(x.ptr) & 15)) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_tc: input slice not 16-B aligned");
    if (w.k != 1 && w.k != 3) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_tc: k must be 1 or 3");
    const int pad = w.k / 2;
    const int Ho = (x.h + 2 * pad - w.k) / w.stride + 1, Wo = (x.w + 2 * pad - w.k) / w.stride + 1;
    if (y.h != Ho || y.w != Wo || y.n != x.n || y.c != w.cout) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_tc: output view mismatch (" + w.name + ")");
    if (res && (res->dtype != x.dtype || res->c != w.cout)) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_tc: residual view mismatch");

    ConvTcOp& o = *op;
    o.x = reinterpret_cast<const __nv_bfloat16*>(x.ptr);
    o.y = y.ptr;
    o.res = res ? reinterpret_cast<const __nv_bfloat16*>(res->ptr) : nullptr;
    o.bias = w.bias;
    o.N = x.n; o.H = x.h; o.W = x.w; o.Cin = w.cin; o.xpitch = x.pitch;
    o.Ho = Ho; o.Wo = Wo; o.Cout = w.cout; o.ypitch = y.pitch; o.rpitch = res ? res->pitch : 0;
    o.y_f32 = y.dtype == DT_F32;
    o.f16 = x.dtype == DT_F16 ? 1 : 0;
    o.silu_tanh = conv_silu_tanh(o.f16 != 0) ? 1 : 0;
    // 16-byte vector stores / residual loads need aligned slices; otherwise the epilogue goes scalar
    o.y_vec = ((y.pitch * y.esize()) % 16 == 0 && (reinterpret_cast<uintptr_t>(y.ptr) & 15) == 0) ? 1 : 0;
    o.r_vec = (res && (res->pitch % 8) == 0 && (reinterpret_cast<uintptr_t>(res->ptr) & 15) == 0) ? 1 : 0;
    o.k = w.k; o.stride = w.stride; o.pad = pad; o.act = w.act;
    o.kc = conv_kc(w, y.dtype == DT_F32);                       // same K chunking as the persistent kernel: identical accumulation order
    o.swz = o.kc * 2;
    o.cchunks = w.cin / o.kc;
    o.nkb = w.k * w.k * o.cchunks;
    o.m_total = x.n * Ho * Wo;
    o.a_tma = (allow_tma_a && w.k == 1 && w.stride == 1) ? 1 : 0;

    // N tile: whole Cout when it fits one UMMA (<=256), else the largest multiple-of-16 divisor <= 256
    int ntile = w.cout_pad;
    if (ntile_hint > 0 && ntile_hint % 16 == 0 && w.cout_pad % ntile_hint == 0) ntile = ntile_hint;
    if (ntile > 256) {
        ntile = 0;
        for (int c = 256; c >= 16; c -= 16) if (w.cout_pad % c == 0) { ntile = c; break; }
    }
    o.ntile = ntile;
    o.ngrid = w.cout_pad / ntile;

    const uint32_t a_bytes = kTileM * o.kc * 2;
    const uint32_t b_bytes = (uint32_t)ntile * o.kc * 2;
    const uint32_t b_alloc = (b_bytes + 1023u) & ~1023u;
    const uint32_t stage_bytes = a_bytes + b_alloc;
    int stages = 4;
    while (stages < kMaxStages && (uint32_t)(stages + 1) * stage_bytes <= 96u * 1024u) ++stages;   // small tiles: deeper ring, still 2 CTAs/SM
    while (stages > 2 && (uint32_t)stages * stage_bytes > 200u * 1024u) --stages;
    if (stages > o.nkb) stages = o.nkb;
    o.stages = stages;
    o.smem_bytes = 2048 + stages * stage_bytes;   // 1 KB alignment slack + 1 KB barrier block
    int cols = 32;
    while (cols < ntile) cols <<= 1;
    o.tmem_cols = cols;

    ZL_TRY(make_tmap_2d_16(&o.tmap_w, w.w_tc, (uint64_t)w.ktot, (uint64_t)w.cout_pad, (uint64_t)w.ktot * 2, o.kc, ntile, o.swz, o.f16));
    if (o.a_tma) {
        ZL_TRY(make_tmap_2d_16(&o.tmap_a, x.ptr, (uint64_t)w.cin, (uint64_t)o.m_total, (uint64_t)x.pitch * 2, o.kc, kTileM, o.swz, o.f16));
    } else {
        o.tmap_a = o.tmap_w;
    }
    o.flops = 2.0 * o.m_total * (double)w.cout * w.ktot;
    o.bytes = (double)x.pixels() * w.cin * 2 + (double)o.m_total * w.cout * (o.y_f32 ? 4 : 2) +
              (double)w.cout * w.ktot * 2 + (res ? (double)o.m_total * w.cout * 2 : 0.0);
    return ZL_OK;
}

int32_t conv_tc_launch(cudaStream_t st, const ConvTcOp& o)
{
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    // the attribute is per device: set it again cheaply when the current device changes
    static thread_local int last_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != last_dev) {
        ZL_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        last_dev = dev;
    }
    if (attr_err != cudaSuccess) ZL_FAIL(ZL_SYSTEM_ERROR, "cudaFuncSetAttribute(conv_tc_kernel) failed");

    Params p;
    p.x = o.x; p.y = o.y; p.res = o.res; p.bias = o.bias;
    p.H = o.H; p.W = o.W; p.Cin = o.Cin; p.xpitch = o.xpitch;
    p.Ho = o.Ho; p.Wo = o.Wo; p.Cout = o.Cout; p.ypitch = o.ypitch; p.rpitch = o.rpitch; p.y_f32 = o.y_f32;
    p.k = o.k; p.stride = o.stride; p.pad = o.pad; p.act = o.act;
    p.kc = o.kc; p.nkb = o.nkb; p.cchunks = o.cchunks;
    p.ntile = o.ntile; p.m_total = o.m_total; p.a_tma = o.a_tma; p.stages = o.stages; p.y_vec = o.y_vec; p.r_vec = o.r_vec; p.f16 = o.f16; p.silu_tanh = o.silu_tanh;
    p.a_bytes = kTileM * o.kc * 2;
    p.b_bytes = (uint32_t)o.ntile * o.kc * 2;
    p.stage_bytes = p.a_bytes + ((p.b_bytes + 1023u) & ~1023u);
    p.tmem_cols = o.tmem_cols;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ceil_div(o.m_total, kTileM), o.ngrid); cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = (size_t)o.smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    ZL_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel, o.tmap_w, o.tmap_a, p));
    return ZL_OK;
}

}  // namespace zl
