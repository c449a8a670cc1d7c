// conv_halo.cu — the persistent tcgen05 convolution kernel behind every 16-bit conv layer (sm_100a only).
//
// Same reference nodes as conv_tc.cu (Conv+BN+SiLU inside Ort::Session::Run,
// src/inference/onnx_engine.cpp:577-585).  MODE 9 = 3x3 stride 1, MODE 2 = 3x3 stride 2, MODE 1 = 1x1, MODE 4 = layer 0 as a
// 2x2 conv over the space-to-depth image (Geo<MODE> below).  The nine filter taps are NOT gathered nine times.  Per output
// tile of 8 (w) x 16*sub (h) pixels:
//   * ONE 4-D TMA load per channel chunk brings the (8+2) x (16*sub+2) input patch into shared memory (OOB zero fill ==
//     conv padding, image borders, ragged tiles), written in the 32/64/128-byte swizzle the tensor core expects;
//   * the A operand of tap (r,s) is the SAME patch read through a UMMA descriptor whose start address is shifted by
//     (r*10+s) pixels and whose 8-row groups (one tile row = 8 pixels) are strided by the patch row pitch (SBO = 10 pixels).
//     The swizzle XOR is a function of the absolute smem address, so a shifted view stays consistent with what TMA wrote;
//   * all taps x Cin/16 MMAs of the tile accumulate into one slot of a ring of up to 8 TMEM accumulators while sixteen
//     epilogue warps (two sets of eight on alternate tiles) drain earlier ones: +bias, SiLU, +residual, 16-bit / fp32 NHWC
//     through swizzled staging tiles and bulk tensor stores (or 256-bit global stores on the 1x1 layers).
// The CTA is persistent (grid = #SMs, 640 threads: 3 TMA producer warps, 1 MMA warp, 16 epilogue warps).  Weights are
// either RESIDENT in shared memory for the whole launch (9 x Cout x Cin 16-bit values, or an N-split slice of them) or
// STREAMED with the patches (each stage carries its channel chunk's [taps][nt][kc] weights), whichever persist_plan prices
// cheaper; per tile only the patch (1.4x the tile's input) is read and the output written once.
// Measured behaviour and the reasons for this shape: DESIGN.md 4.1, profiles/README_r02.md.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "half16.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace zl {
namespace {

using namespace tc;

constexpr int kTW = 8, kTH = 16;            // one MMA sub-tile: 8 (w) x 16 (h) output pixels = 128 GEMM rows
// TAPS == 9: 3x3 conv, patch = tile + 1-pixel halo (width 10, height 16*sub + 2)
// TAPS == 1: 1x1 conv, patch = tile (width 8, height 16*sub); the pixel list is viewed as an 8-wide image
// MODE == 2: 3x3 stride-2 conv.  The input region of a tile is fetched as FOUR parity sub-patches (even/odd rows x
//            even/odd columns) by TMA loads with element stride 2; tap (r,s) then reads sub-patch
//            (r != 1, s != 1) shifted by (r >= 1, s >= 1) — the same shifted-view trick, patch width 9.
// MODE == 4: layer 0 (3->c 3x3 s2) as a 2x2 stride-1 conv over the SPACE-TO-DEPTH image the preprocess kernel writes
//            ([B, H/2, W/2, 16] = the 2x2 pixel block's 12 values + 4 zeros): taps at offsets {-1, 0}^2, patch = tile
//            + one halo row/column on the top/left, so the first layer runs on TMA + tcgen05 like every other conv.
template <int MODE> struct Geo {
    static constexpr int TAPS = MODE == 1 ? 1 : (MODE == 4 ? 4 : 9);
    static constexpr int PW = MODE == 9 ? kTW + 2 : ((MODE == 2 || MODE == 4) ? kTW + 1 : kTW);
    static constexpr int HALO = (MODE == 9 || MODE == 4) ? 1 : 0;
    static constexpr int NPATCH = MODE == 2 ? 4 : 1;
};
constexpr int kEpiWarps = 16;                // 4 per TMEM lane quarter, each owning a share of the accumulator columns
// Measured on B200 (probes/tma_rate.cu): ONE thread completes a pipeline stage (empty-wait, expect_tx, TMA issue) every
// ~500 cycles + ~130 per TMA instruction whatever the box size, while issuers in different warps scale linearly (8 warps
// reach 100 B/clk/SM).  Stages here are 5-25 KB, so a single producer thread caps a CTA at 10-40 B/clk: the stage
// sequence is dealt round-robin to kProducers warps (warp 0 and the warps after the epilogue warps).
constexpr int kProducers = 3;
constexpr int kThreads = 64 + 32 * kEpiWarps + 32 * (kProducers - 1);
constexpr int kMaxPatchStages = 12;
constexpr int kMaxAcc = 8;                   // TMEM accumulator ring (512 columns / (sub * nt))

struct HaloParams {
    void* y;
    const __nv_bfloat16* res;
    const float* bias;
    int32_t N, H, W, Cin, Cout, ntile;
    int32_t ypitch, rpitch, y_f32, y_vec, r_vec, f16, act, silu_tanh;
    int32_t kc, cchunks, stages, sub, y_tma, nsplit, nt, ostage;
    int32_t tiles_x, tiles_y, num_tiles;
    int32_t epi_variant, esets, direct_store;                      // epilogue_role specialisation (0..5 fast, 6 generic); warp sets (1 or 2)
    int32_t wstream, nacc, nacc_log2, acc_cols;      // weights streamed with the patches (1) or resident (0); accumulator ring
    uint32_t wtile_bytes, wchunk_bytes, wchunk_alloc, patch_bytes, patch_alloc, subpatch_alloc, chunk_stride, stage_stride, tmem_cols;
    int32_t cps, nst;                                // channel chunks per pipeline stage; stages per tile (= cchunks / cps)
    unsigned long long* stats;      // STATS instantiation only: kHaloStatSlots cycle counters summed over CTAs (zl_engine_profile_stalls)
};

// Issue all MMAs of one (tile, channel chunk): 9 taps x SUB sub-tiles x KSTEPS k-steps, fully unrolled.
// Measured on B200 (umma_probe.cu): a tcgen05.mma with N <= 64 occupies the tensor pipe for ~48 cycles, so
// the single issuing thread must spend far less than that per instruction: both descriptors are the
// stage's base descriptor plus a compile-time constant in the 14-bit address field.
template <int KSTEPS, int SUB, int MODE>
__device__ __forceinline__ void issue_chunk(uint32_t tmem_d, uint32_t ntile, uint64_t adesc0, uint64_t bdesc0, uint32_t subpatch16,
                                            uint32_t btap_stride16, uint32_t idesc, bool first_chunk)
{
    constexpr uint32_t swz16 = (uint32_t)KSTEPS * 2u;        // bytes per pixel row / 16
    constexpr int kPW = Geo<MODE>::PW, TAPS = Geo<MODE>::TAPS;
#pragma unroll
    for (int tap = 0; tap < TAPS; ++tap) {
        const int r = tap / 3, sf = tap % 3;
        uint32_t aoff = 0u;
        if (MODE == 9) aoff = (uint32_t)(r * kPW + sf) * swz16;
        if (MODE == 4) aoff = (uint32_t)((tap >> 1) * kPW + (tap & 1)) * swz16;
        if (MODE == 2) aoff = (uint32_t)((r != 1 ? 2 : 0) + (sf != 1 ? 1 : 0)) * subpatch16 + (uint32_t)((r >= 1 ? kPW : 0) + (sf >= 1 ? 1 : 0)) * swz16;
        const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)tap * btap_stride16);
#pragma unroll
        for (int j = 0; j < SUB; ++j) {
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
                const uint64_t ad = adesc0 + (uint64_t)(aoff + (uint32_t)(j * kTH * kPW) * swz16 + 2u * (uint32_t)k);
                umma_bf16(tmem_d + (uint32_t)j * ntile, ad, bdesc + (uint64_t)(2 * k), idesc, (first_chunk && tap == 0 && k == 0) ? 0u : 1u);
            }
        }
    }
}

// Division-free walk over a CTA's tiles: tile -> (image n, tile row ty, tile column tx), advanced by a fixed step.
struct TileWalk { int n, ty, tx, dn, dty, dtx; };
__device__ __forceinline__ void walk_init(TileWalk& t, int tile, int step, int tiles_x, int tiles_y)
{
    const int per_img = tiles_x * tiles_y;
    t.n = tile / per_img;
    int rem = tile - t.n * per_img;
    t.ty = rem / tiles_x; t.tx = rem - t.ty * tiles_x;
    t.dn = step / per_img;
    rem = step - t.dn * per_img;
    t.dty = rem / tiles_x; t.dtx = rem - t.dty * tiles_x;
}
__device__ __forceinline__ void walk_next(TileWalk& t, int tiles_x, int tiles_y)
{
    t.tx += t.dtx;
    if (t.tx >= tiles_x) { t.tx -= tiles_x; ++t.ty; }
    t.ty += t.dty;
    if (t.ty >= tiles_y) { t.ty -= tiles_y; ++t.n; }
    t.n += t.dn;
}


// ---------------------------------------------------------------------------------------------------------------------
// Epilogue role.  VAR 0..5 are the specialised fast variants (format = VAR / 3: 0 bf16, 1 fp16; kind = VAR % 3: 0 SiLU ->
// 16-bit, 1 SiLU + residual -> 16-bit, 2 no activation -> fp32), chosen on the host when the output goes through bulk
// tensor stores and every 16-channel chunk is whole; VAR 6 is the generic body with run-time flags (ragged channel
// counts, unaligned rows, plain stores).  Round-2 ncu: the kernel is ISSUE-bound in these warps (issue slots ~60 % busy,
// ~350 instructions per 32 px x 16 ch item of which 64 are the arithmetic), so the hot variants carry no run-time format
// / activation / residual branches, no 64-bit index arithmetic per item, and waiting warps sleep in mbarrier.try_wait.
// * The 16 warps work as TWO SETS of 8 on alternate tiles; a warp's items of one tile are processed in pairs with both
//   tcgen05.ld in flight.  No integer division in the loop.
// * Register arrays are indexed with compile-time constants only: a[] / v[] must never be demoted to local memory.
struct EpiCtx {
    const CUtensorMap* tmap_y;
    const float* bias_s;
    uint32_t tmem_base, bar_tfull, bar_tempty, obase;
    int tile0, tile_step, n_off, ntile, cout_l;
    const __nv_bfloat16* res_g;
    void* y_g;
    uint32_t warp, lane;
};

template <bool STATS, int VAR>
__device__ __forceinline__ bool epilogue_role(const HaloParams& p, const EpiCtx& c)
{
    constexpr bool FAST = VAR < 6;
    constexpr bool F16C = (VAR / 3) == 1;                    // compile-time format of the fast variants
    constexpr int KIND = VAR % 3;
    long long st_a = 0, st_b = 0, st_c = 0, st_d = 0, st_e = 0, st_f = 0;
#define ZL_ST_BEGIN(t) long long t = 0; if (STATS) t = clock64()
#define ZL_ST_END(t, acc) if (STATS) acc += clock64() - t
    const uint32_t warp = c.warp, lane = c.lane;
    const uint32_t q = warp & 3u;                            // TMEM lane quarter this warp may read
    const int cgp = (int)(warp - 2u) >> 2;                   // 0..3 within the quarter
    // esets == 2: two sets of 8 warps on alternate tiles; esets == 1 (a CTA with a single tile: the latency path): all 16 on every tile
    const int wq = 4 / p.esets;                              // warps of this quarter serving one tile
    const int set = cgp / wq, c2 = cgp - set * wq;           // tile residue this warp serves; its share of the quarter's items
    const int row = (int)(q * 32u + lane);                   // A row == TMEM lane: h = row / 8, w = row % 8
    const int th = row >> 3, tw = row & 7;
    const int nchunk = c.ntile >> 4;
    const int items = p.sub * nchunk;
    const bool leader = elect_one();                        // the one lane that owns this warp's TMA-store bulk groups
    const uint32_t stage_out = c.obase + (warp - 2u) * 2u * (uint32_t)p.ostage;   // this warp's two staging blocks ([32 px][16 ch]; 1 KB 16-bit / 2 KB fp32)
    const bool f16 = FAST ? F16C : (p.f16 != 0);
    const bool y_f32 = FAST ? (KIND == 2) : (p.y_f32 != 0);
    const bool act = FAST ? (KIND != 2) : (p.act != 0);
    const bool has_res = FAST ? (KIND == 1) : (c.res_g != nullptr);
    const bool res_fast = has_res && (FAST || p.r_vec);      // residual rows are 16-byte aligned: two 128-bit loads per thread
    const int cout_l = c.cout_l;
    const __nv_bfloat16* res_g = c.res_g;
    const float* bias_s = c.bias_s;
    uint32_t nstore = 0;
    // first item of this warp in every tile: item index c2 -> (sub-tile j0, chunk k0)
    int j0 = 0, k0 = c2;
    while (k0 >= nchunk) { k0 -= nchunk; ++j0; }
    const int res_jstride = kTH * p.W * p.rpitch;            // residual elements between sub-tiles (fits 32 bits: one image row block)
    const int y_jstride = kTH * p.W * p.ypitch;              // output elements between sub-tiles
    // tile walk: this set takes tiles tile0 + set*step, then every 2*step-th; (n, ty, tx) advance by a fixed decomposition
    TileWalk tw_;
    walk_init(tw_, c.tile0 + set * c.tile_step, p.esets * c.tile_step, p.tiles_x, p.tiles_y);
    // the residual (and, transitively, everything this kernel overwrites) belongs to earlier kernels of the stream
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // The residual of this warp's first item of the NEXT tile is prefetched into L1 one tile ahead (no registers held: the
    // kernel is at its 96-register cap, and a register prefetch spilled and doubled the time of the residual layers): on
    // the narrow residual layers a warp has one item per tile and no earlier point inside the tile to ask from.
    TileWalk twn = tw_;                                      // the walk, one tile ahead
    auto prefetch_first = [&](const TileWalk& t) {
        if (!(res_fast && c2 < items)) return;
        const int oxq = t.tx * kTW + tw, oyq = t.ty * p.sub * kTH + th + j0 * kTH;
        if (oyq < p.H && oxq < p.W)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(res_g + (((size_t)t.n * p.H + oyq) * p.W + oxq) * p.rpitch + (k0 << 4)));
    };
    for (int tile = c.tile0 + set * c.tile_step, tl = set; tile < p.num_tiles; tile += p.esets * c.tile_step, tl += p.esets, walk_next(tw_, p.tiles_x, p.tiles_y)) {
        const uint32_t acc = (uint32_t)tl & (uint32_t)(p.nacc - 1), aph = ((uint32_t)tl >> p.nacc_log2) & 1u;
        const int n = tw_.n, ty = tw_.ty, tx = tw_.tx;
        const int ox = tx * kTW + tw;
        const int oy_base = ty * p.sub * kTH + th;           // this thread's row in sub-tile 0
        // residual pixel of sub-tile 0 (64-bit once per tile; per item only 32-bit offsets are added)
        const __nv_bfloat16* res_px = has_res ? res_g + (((size_t)n * p.H + oy_base) * p.W + ox) * p.rpitch : nullptr;
        char* y_px = reinterpret_cast<char*>(c.y_g) + (((size_t)n * p.H + oy_base) * p.W + ox) * p.ypitch * (y_f32 ? 4 : 2);   // direct-store variant
        // residual of this warp's FIRST item of the tile: requested before the accumulator wait, so its latency hides under the
        // MMAs (and it was prefetched into L1 one tile ago); the residuals of the following items are requested one item ahead
        uint4 rn0 = make_uint4(0u, 0u, 0u, 0u), rn1 = rn0;
        bool rn_have = false;
        if (res_fast && c2 < items) {
            const int oy = oy_base + j0 * kTH;
            if (oy < p.H && ox < p.W && (FAST || (k0 << 4) + 16 <= cout_l)) {
                const uint4* rp = reinterpret_cast<const uint4*>(res_px + j0 * res_jstride + (k0 << 4));
                rn0 = __ldg(rp);
                rn1 = __ldg(rp + 1);
                rn_have = true;
            }
        }
        walk_next(twn, p.tiles_x, p.tiles_y);
        if (tile + p.esets * c.tile_step < p.num_tiles) prefetch_first(twn);
        {
            ZL_ST_BEGIN(t0);
            mbar_wait(c.bar_tfull + 8u * acc, aph, 5);
            ZL_ST_END(t0, st_a);
        }
        tc_fence_after();
        ZL_ST_BEGIN(t_epi);
        const uint32_t taddr = c.tmem_base + ((q * 32u) << 16) + acc * (uint32_t)p.acc_cols;
        // One item (32 px x 16 ch) per iteration, ROLLED: the body is ~300 instructions and has to stay inside the ~6 KB L0
        // instruction cache of the scheduler (round-2 ncu: with two items unrolled per iteration stall_no_instruction was
        // 2-4 per issue on the epilogue-bound layers).  The next item's accumulator columns (tcgen05.ld) and residual are
        // requested before the current item is processed, so both latencies stay hidden as in the unrolled form.
        int j = j0, k = k0;                                  // (sub-tile, chunk) of the current item
        uint32_t vn[16];
        if (c2 < items) tmem_ld16(taddr + (uint32_t)(j * p.nt + (k << 4)), vn);
#pragma unroll 1
        for (int item = c2; item < items; item += wq) {
            uint32_t vc[16];
            {
                ZL_ST_BEGIN(t0);
                tmem_ld_wait();
                ZL_ST_END(t0, st_c);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) vc[i] = vn[i];
            const uint4 r0 = rn0, r1 = rn1;
            const bool r_have = rn_have;
            int jn = j, kn = k + wq;                          // the next item of this warp in this tile
            while (kn >= nchunk) { kn -= nchunk; ++jn; }
            rn_have = false;
            if (item + wq < items) {
                tmem_ld16(taddr + (uint32_t)(jn * p.nt + (kn << 4)), vn);
                if (res_fast) {
                    const int oyn = oy_base + jn * kTH;
                    if (oyn < p.H && ox < p.W && (FAST || (kn << 4) + 16 <= cout_l)) {
                        const uint4* rp = reinterpret_cast<const uint4*>(res_px + jn * res_jstride + (kn << 4));
                        rn0 = __ldg(rp);
                        rn1 = __ldg(rp + 1);
                        rn_have = true;
                    }
                }
            }
            {
                const int jj = j, c0 = k << 4;
                const int oy = oy_base + jj * kTH;
                const bool in_px = oy < p.H && ox < p.W;
                const bool in_img = in_px && c0 < cout_l;
                float a[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + 4 * i);
                    a[4 * i + 0] = __uint_as_float(vc[4 * i + 0]) + b4.x; a[4 * i + 1] = __uint_as_float(vc[4 * i + 1]) + b4.y;
                    a[4 * i + 2] = __uint_as_float(vc[4 * i + 2]) + b4.z; a[4 * i + 3] = __uint_as_float(vc[4 * i + 3]) + b4.w;
                }
                if (act) {
                    if (FAST || p.silu_tanh) {                   // the fast variants are tanh-only: the two-MUFU form (an A/B switch) takes the generic body
#pragma unroll
                        for (int i = 0; i < 16; ++i) a[i] = silu_tanh(a[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) a[i] = silu_exp(a[i]);
                    }
                }
                if (has_res) {
                    if (r_have) {
                        const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                        if (f16) {                               // one format branch per item, not per pair of values
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rw[i]));
                                a[2 * i] += f.x; a[2 * i + 1] += f.y;
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                a[2 * i] += __uint_as_float(rw[i] << 16);            // bf16 -> fp32 is a shift
                                a[2 * i + 1] += __uint_as_float(rw[i] & 0xffff0000u);
                            }
                        }
                    } else if (!FAST && in_img) {                // unaligned or ragged residual rows: element by element, statically indexed
                        const uint16_t* rp = reinterpret_cast<const uint16_t*>(res_px + jj * res_jstride + c0);
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < cout_l) a[i] += unpack1_16(rp[i], f16);
                    }
                }
                if (FAST && p.direct_store) {
                    // ---- 256-bit global stores straight from registers (st.global.v8.b32, sm_100): one full 32-byte sector
                    //      per thread (two for fp32), no staging tile, no proxy fence, no bulk-store instruction — the TMA unit
                    //      and ~12 % of the tile's shared-memory traffic are left to the patches and the MMAs
                    if (in_px) {
                        if (y_f32) {
                            float* yp = reinterpret_cast<float*>(y_px) + jj * y_jstride + c0;
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(yp), "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])),
                                         "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])), "r"(__float_as_uint(a[4])), "r"(__float_as_uint(a[5])),
                                         "r"(__float_as_uint(a[6])), "r"(__float_as_uint(a[7])) : "memory");
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(yp + 8), "r"(__float_as_uint(a[8])), "r"(__float_as_uint(a[9])),
                                         "r"(__float_as_uint(a[10])), "r"(__float_as_uint(a[11])), "r"(__float_as_uint(a[12])), "r"(__float_as_uint(a[13])),
                                         "r"(__float_as_uint(a[14])), "r"(__float_as_uint(a[15])) : "memory");
                        } else {
                            uint32_t w[8];
                            if (f16) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) { const __half2 hh = __floats2half2_rn(a[2 * i], a[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&hh); }
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) { const __nv_bfloat162 hh = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&hh); }
                            }
                            uint16_t* yp = reinterpret_cast<uint16_t*>(y_px) + jj * y_jstride + c0;
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(yp), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                                         "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
                        }
                    }
                } else if (FAST || p.y_tma) {
                    // ---- stage [32 px][16 ch] in smem, one TMA store per warp (image borders and ragged channel counts are
                    //      clipped by the hardware against the tensor map's extents)
                    const uint32_t sbuf = stage_out + (nstore & 1u) * (uint32_t)p.ostage;
                    {
                        ZL_ST_BEGIN(t0);
                        if (nstore >= 2u) { if (leader) tma_store_wait_read<1>(); __syncwarp(); }    // the store that last used this block has read it
                        ZL_ST_END(t0, st_d);
                    }
                    ZL_ST_BEGIN(t_st);
                    if (y_f32) {
                        const uint32_t xr = (lane >> 1) & 3u;                                       // 64-B swizzle: chunk ^= address bits 7..8
                        const uint32_t rowb = sbuf + lane * 64u;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + ((((uint32_t)i) ^ xr) << 4)), "r"(__float_as_uint(a[4 * i])),
                                         "r"(__float_as_uint(a[4 * i + 1])), "r"(__float_as_uint(a[4 * i + 2])), "r"(__float_as_uint(a[4 * i + 3])) : "memory");
                    } else {
                        uint32_t w[8];
                        if (f16) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) { const __half2 hh = __floats2half2_rn(a[2 * i], a[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&hh); }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) { const __nv_bfloat162 hh = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&hh); }
                        }
                        const uint32_t xr = (lane >> 2) & 1u;                                       // 32-B swizzle: chunk ^= address bit 7
                        const uint32_t rowb = sbuf + lane * 32u;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + ((0u ^ xr) << 4)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowb + ((1u ^ xr) << 4)), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (leader) {
                        tma_store_4d(c.tmap_y, sbuf, c.n_off + c0, tx * kTW, (ty * p.sub + jj) * kTH + (int)q * 4, n);
                        tma_store_commit();
                    }
                    ZL_ST_END(t_st, st_e);
                    ++nstore;
                } else if (in_img) {
                    // ---- outputs whose rows are not 16-byte aligned (e.g. nc = 2 class maps): plain stores, statically indexed
                    const size_t m = ((size_t)n * p.H + oy) * p.W + ox;
                    const bool full = (c0 + 16 <= cout_l);
                    if (y_f32) {
                        float* yp = reinterpret_cast<float*>(c.y_g) + m * p.ypitch + c0;
                        if (full && p.y_vec) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                reinterpret_cast<float4*>(yp)[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < cout_l) yp[i] = a[i];
                        }
                    } else {
                        uint16_t* yp = reinterpret_cast<uint16_t*>(c.y_g) + m * p.ypitch + c0;
                        if (full && p.y_vec) {
                            uint32_t w[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) w[i] = pack2_16(a[2 * i], a[2 * i + 1], f16);
                            reinterpret_cast<uint4*>(yp)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                            reinterpret_cast<uint4*>(yp)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                if (c0 + i < cout_l) yp[i] = pack1_16(a[i], f16);
                        }
                    }
                }
            }
            j = jn; k = kn;
        }
        // this warp is done reading the accumulator: hand it back to the MMA warp
        {
            ZL_ST_BEGIN(t0);
            tc_fence_before();
            __syncwarp();
            if (leader) mbar_arrive(c.bar_tempty + 8u * acc);
            ZL_ST_END(t0, st_f);
        }
        ZL_ST_END(t_epi, st_b);
    }
    if (STATS && warp == 2 && lane == 0) {
        atomicAdd(p.stats + 4, (unsigned long long)st_a);      // epilogue warp 2 waiting for a full accumulator
        atomicAdd(p.stats + 5, (unsigned long long)st_b);      // ... busy
        atomicAdd(p.stats + 12, (unsigned long long)st_c);     // ...... of which: tcgen05.ld + wait::ld
        atomicAdd(p.stats + 13, (unsigned long long)st_d);     // ...... waiting for the staging block's previous bulk store to be read
        atomicAdd(p.stats + 14, (unsigned long long)st_e);     // ...... st.shared + proxy fence + bulk store issue
        atomicAdd(p.stats + 15, (unsigned long long)st_f);     // ...... handing the accumulator back
    }
#undef ZL_ST_BEGIN
#undef ZL_ST_END
    return leader;
}

template <int MODE, bool STATS>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_x,
                 const __grid_constant__ CUtensorMap tmap_y, const HaloParams p)
{
    extern __shared__ uint8_t smem_raw[];
    constexpr int kPW = Geo<MODE>::PW, kHalo = Geo<MODE>::HALO, TAPS = Geo<MODE>::TAPS, NPATCH = Geo<MODE>::NPATCH;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    // STATS: where each role's cycles go (slot meanings in kernels.h, HaloStat).  Compiled out of the production kernel.
    long long st_a = 0, st_b = 0, st_c = 0, st_d = 0, st_e = 0, st_f = 0, t_begin = 0, t_pro = 0;
    if (STATS) t_begin = clock64();
#define ZL_ST_BEGIN(t) long long t = 0; if (STATS) t = clock64()
#define ZL_ST_END(t, acc) if (STATS) acc += clock64() - t

    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_pfull = base;                               // kMaxPatchStages x 8
    const uint32_t bar_pempty = base + 8u * kMaxPatchStages;       // kMaxPatchStages x 8
    const uint32_t bar_wfull = base + 16u * kMaxPatchStages;
    const uint32_t bar_tfull = bar_wfull + 8u;                     // kMaxAcc x 8
    const uint32_t bar_tempty = bar_tfull + 8u * kMaxAcc;          // kMaxAcc x 8
    const uint32_t tmem_slot = bar_tempty + 8u * kMaxAcc;
    const uint32_t bias_off = 1024u;                               // fp32 bias[ntile <= 256] in its own 1 KB block
    const uint32_t wbase = base + 2048u;
    // resident weights: [chunk][tap][nt][kc] in front of the stage ring; streamed weights: the second half of every stage
    const uint32_t pbase = wbase + (p.wstream ? 0u : (uint32_t)p.cchunks * p.wchunk_alloc);
    const uint32_t obase = pbase + (uint32_t)p.stages * p.stage_stride;   // per-epilogue-warp output staging: 2 x 1 KB each
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int stages = p.stages;
    // N-split: CTA c owns output-channel slice (c % nsplit) for good, so its weight slice stays resident
    const int slice = (int)blockIdx.x % p.nsplit;
    const int n_off = slice * p.nt;
    const int ntile = min(p.nt, p.ntile - n_off);            // this CTA's UMMA N (multiple of 16)
    const int cout_l = p.Cout - n_off;                       // valid output channels left in this slice (may be <= 0)
    const int tile0 = (int)blockIdx.x / p.nsplit, tile_step = (int)gridDim.x / p.nsplit;
    const float* bias_g = p.bias + n_off;
    const size_t esz_y = p.y_f32 ? 4 : 2;
    void* y_g = reinterpret_cast<char*>(p.y) + (size_t)n_off * esz_y;
    const __nv_bfloat16* res_g = p.res ? p.res + n_off : nullptr;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bar_pfull + 8u * s, 1u);
            mbar_init(bar_pempty + 8u * s, 1u);
        }
        mbar_init(bar_wfull, (uint32_t)kProducers);
        for (int a = 0; a < p.nacc; ++a) {
            mbar_init(bar_tfull + 8u * a, 1u);
            mbar_init(bar_tempty + 8u * a, (uint32_t)kEpiWarps / (uint32_t)p.esets);   // one arrival per epilogue warp of the set that drains it
        }
        fence_barrier_init();
        tma_prefetch_desc(&tmap_w);
        tma_prefetch_desc(&tmap_x);
        if (p.y_tma) tma_prefetch_desc(&tmap_y);
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, p.tmem_cols);
        tmem_relinquish();
    }
    float* bias_s = reinterpret_cast<float*>(smem_raw + (base + bias_off - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < ntile; i += (int)blockDim.x) bias_s[i] = bias_g[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // Programmatic dependent launch: let the next kernel's CTAs take this SM as soon as this CTA exits and run their
    // own prologue (barrier init, TMEM alloc, weight TMA) under this kernel's tail ...
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (STATS) t_pro = clock64();

    const int tiles_per_img = p.tiles_x * p.tiles_y;
    bool store_leader = false;                               // set on the one lane per epilogue warp that issues TMA stores

    if (warp == 0 || warp >= 2u + kEpiWarps) {
        // ===== TMA producers: (resident mode) the weights once, then one stage per (tile, channel chunk), dealt round-robin =====
        const uint32_t pi = warp == 0 ? 0u : warp - (1u + kEpiWarps);            // producer index 0 .. kProducers-1
        if (elect_one()) {
            if (!p.wstream) {
                // chunk cc of the weight slice = ONE 3-D box [tap][nt][kc] (tensor map dims: channel-in-tap, Cout, tap)
                uint32_t mine = 0;
                for (uint32_t cc = pi; cc < (uint32_t)p.cchunks; cc += kProducers) ++mine;
                mbar_arrive_expect_tx(bar_wfull, mine * p.wchunk_bytes);
                for (uint32_t cc = pi; cc < (uint32_t)p.cchunks; cc += kProducers)
                    tma_load_3d(&tmap_w, bar_wfull, wbase + cc * p.wchunk_alloc, (int)cc * p.kc, n_off, 0);
            }
            // ... but nothing produced by the previous kernel is read (and nothing it may still read is overwritten)
            // before it has completed: weights and bias above are constants, activations start here
            asm volatile("griddepcontrol.wait;" ::: "memory");
            // Parity waits alias when a producer runs two phases ahead of a stage's barrier.  Producer pi reaches item i only
            // after item i - np - stages was consumed, so the stage's phase is at most one behind as long as np <= stages.
            const uint32_t np = (uint32_t)stages < (uint32_t)kProducers ? (uint32_t)stages : (uint32_t)kProducers;
            uint32_t s = 0, ph = 0, turn = 0;
            const uint32_t stage_tx = (p.patch_bytes * (uint32_t)NPATCH + (p.wstream ? p.wchunk_bytes : 0u)) * (uint32_t)p.cps;
            TileWalk tw_;
            walk_init(tw_, tile0, tile_step, p.tiles_x, p.tiles_y);
            for (int tile = tile0; tile < p.num_tiles; tile += tile_step, walk_next(tw_, p.tiles_x, p.tiles_y)) {
                const int n = tw_.n, ty = tw_.ty, tx = tw_.tx;
                for (int cs = 0; cs < p.nst; ++cs) {        // one pipeline stage = cps consecutive channel chunks under ONE barrier
                    if (turn == pi) {
                        ZL_ST_BEGIN(t0);
                        mbar_wait(bar_pempty + 8u * s, ph ^ 1u, 1);
                        ZL_ST_END(t0, st_a);
                        mbar_arrive_expect_tx(bar_pfull + 8u * s, stage_tx);
                        for (int g = 0; g < p.cps; ++g) {
                            const int cc = cs * p.cps + g;
                            const uint32_t stage = pbase + s * p.stage_stride + (uint32_t)g * p.chunk_stride;
                            if (p.wstream) tma_load_3d(&tmap_w, bar_pfull + 8u * s, stage + p.patch_alloc, cc * p.kc, n_off, 0);
                            if (MODE == 2) {
#pragma unroll
                                for (int pp = 0; pp < 4; ++pp)       // parity sub-patch pp = 2*(row parity) + (column parity)
                                    tma_load_4d(&tmap_x, bar_pfull + 8u * s, stage + (uint32_t)pp * p.subpatch_alloc, cc * p.kc,
                                                2 * (tx * kTW - 1) + (pp & 1), 2 * (ty * kTH * p.sub - 1) + (pp >> 1), n);
                            } else {
                                tma_load_4d(&tmap_x, bar_pfull + 8u * s, stage, cc * p.kc, tx * kTW - kHalo, ty * kTH * p.sub - kHalo, n);
                            }
                        }
                    }
                    if (++turn == np) turn = 0;
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1u; }
                }
            }
            if (STATS && pi == 0) atomicAdd(p.stats + 0, (unsigned long long)st_a);
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t swz = (uint32_t)p.kc * 2u;
        const uint32_t fmt = p.f16 ? 0u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(ntile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int ksteps = p.kc / 16;
        const uint32_t sbo_a = (uint32_t)kPW * swz;          // one tile row (8 pixels) per 8-row group, groups strided by the patch row
        if (!p.wstream) {
            ZL_ST_BEGIN(t0);
            mbar_wait(bar_wfull, 0u, 2);
            ZL_ST_END(t0, st_d);
        }
        tc_fence_after();
        uint32_t s = 0, ph = 0, tl = 0;
        const uint64_t adesc_base = make_smem_desc_sbo(pbase, swz, sbo_a);
        const uint64_t bdesc_base = make_smem_desc_sbo(p.wstream ? pbase + p.patch_alloc : wbase, swz, 8u * swz);
        const uint32_t stage16 = p.stage_stride >> 4, chunk16 = p.chunk_stride >> 4, wchunk16 = p.wchunk_alloc >> 4, subpatch16 = p.subpatch_alloc >> 4;
        const uint32_t btap16 = p.wtile_bytes >> 4;                          // a chunk's weight tiles are packed [tap][nt][kc]
        const int sel = (ksteps == 4 ? 0 : (ksteps == 2 ? 3 : 6)) + (p.sub == 1 ? 0 : (p.sub == 2 ? 1 : 2));
        for (int tile = tile0; tile < p.num_tiles; tile += tile_step, ++tl) {
            const uint32_t acc = tl & (uint32_t)(p.nacc - 1), aph = (tl >> p.nacc_log2) & 1u;
            {
                ZL_ST_BEGIN(t0);
                mbar_wait(bar_tempty + 8u * acc, aph ^ 1u, 3);    // epilogue has drained this accumulator
                ZL_ST_END(t0, st_b);
            }
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.acc_cols;
            for (int cs = 0; cs < p.nst; ++cs) {
                {
                    ZL_ST_BEGIN(t0);
                    mbar_wait(bar_pfull + 8u * s, ph, 4);
                    ZL_ST_END(t0, st_a);
                }
                tc_fence_after();
                ZL_ST_BEGIN(t_issue);
                if (elect_one()) {
                    const uint32_t nt = (uint32_t)p.nt;                 // accumulator column stride between sub-tiles
                    for (int g = 0; g < p.cps; ++g) {
                        const int cc = cs * p.cps + g;
                        const uint64_t ad = adesc_base + (uint64_t)(s * stage16 + (uint32_t)g * chunk16);
                        const uint64_t bd = bdesc_base + (uint64_t)(p.wstream ? s * stage16 + (uint32_t)g * chunk16 : (uint32_t)cc * wchunk16);
                        const bool first = cc == 0;
                        switch (sel) {
                            case 0: issue_chunk<4, 1, MODE>(tmem_d, nt, ad, bd, subpatch16, btap16, idesc, first); break;
                            case 1: issue_chunk<4, 2, MODE>(tmem_d, nt, ad, bd, subpatch16, btap16, idesc, first); break;
                            case 2: issue_chunk<4, 4, MODE>(tmem_d, nt, ad, bd, subpatch16, btap16, idesc, first); break;
                            case 3: issue_chunk<2, 1, MODE>(tmem_d, nt, ad, bd, subpatch16, btap16, idesc, first); break;
                            case 4: issue_chunk<2, 2, MODE>(tmem_d, nt, ad, bd, subpatch16, btap16, idesc, first); break;
                            case 5: issue_chunk<2, 4, MODE>(tmem_d, nt, ad, bd, subpatch16, btap16, idesc, first); break;
                            case 6: issue_chunk<1, 1, MODE>(tmem_d, nt, ad, bd, subpatch16, btap16, idesc, first); break;
                            case 7: issue_chunk<1, 2, MODE>(tmem_d, nt, ad, bd, subpatch16, btap16, idesc, first); break;
                            default: issue_chunk<1, 4, MODE>(tmem_d, nt, ad, bd, subpatch16, btap16, idesc, first); break;
                        }
                    }
                    umma_commit(bar_pempty + 8u * s);
                    if (cs == p.nst - 1) umma_commit(bar_tfull + 8u * acc);
                }
                __syncwarp();
                ZL_ST_END(t_issue, st_c);
                if (++s == (uint32_t)stages) { s = 0; ph ^= 1u; }
            }
        }
        if (STATS && lane == 0) {
            atomicAdd(p.stats + 1, (unsigned long long)st_a);      // MMA warp waiting for patches
            atomicAdd(p.stats + 2, (unsigned long long)st_b);      // ... for a free accumulator
            atomicAdd(p.stats + 3, (unsigned long long)st_c);      // ... issuing
            atomicAdd(p.stats + 8, (unsigned long long)st_d);      // ... for the weights
        }
    } else {
        // ===== epilogue warps: TMEM -> +bias -> SiLU -> +residual -> 16-bit / fp32 NHWC (epilogue_role, specialised per variant) =====
        EpiCtx ec;
        ec.tmap_y = &tmap_y; ec.bias_s = bias_s; ec.tmem_base = tmem_base; ec.bar_tfull = bar_tfull; ec.bar_tempty = bar_tempty; ec.obase = obase;
        ec.tile0 = tile0; ec.tile_step = tile_step; ec.n_off = n_off; ec.ntile = ntile; ec.cout_l = cout_l;
        ec.res_g = res_g; ec.y_g = y_g; ec.warp = warp; ec.lane = lane;
        switch (p.epi_variant) {
            case 0: store_leader = epilogue_role<STATS, 0>(p, ec); break;
            case 1: store_leader = epilogue_role<STATS, 1>(p, ec); break;
            case 2: store_leader = epilogue_role<STATS, 2>(p, ec); break;
            case 3: store_leader = epilogue_role<STATS, 3>(p, ec); break;
            case 4: store_leader = epilogue_role<STATS, 4>(p, ec); break;
            case 5: store_leader = epilogue_role<STATS, 5>(p, ec); break;
            default: store_leader = epilogue_role<STATS, 6>(p, ec); break;
        }
    }

    if (store_leader && p.y_tma) tma_store_wait_all();               // smem must outlive the bulk stores that read it
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
    if (STATS && threadIdx.x == 0) {
        const long long t_end = clock64();
        atomicAdd(p.stats + 6, (unsigned long long)(t_end - t_begin));   // CTA lifetime
        atomicAdd(p.stats + 7, (unsigned long long)(t_pro - t_begin));   // prologue
        atomicMax(p.stats + 9, (unsigned long long)(t_end - t_begin));   // slowest CTA
        atomicAdd(p.stats + 10, 1ull);                                    // CTAs
        if (blockIdx.x == 0)                                              // the launch plan, packed
            p.stats[11] = (unsigned long long)p.kc | ((unsigned long long)p.cchunks << 8) | ((unsigned long long)p.sub << 16) |
                          ((unsigned long long)p.nsplit << 20) | ((unsigned long long)p.stages << 24) | ((unsigned long long)p.nt << 32) |
                          (((unsigned long long)p.num_tiles & 0x3fffull) << 42) | ((unsigned long long)p.nacc << 56) | ((unsigned long long)p.wstream << 62);
    }
#undef ZL_ST_BEGIN
#undef ZL_ST_END
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int32_t make_tmap_nhwc(CUtensorMap* map, const View& x, int kc, int swz, bool f16, int patch_w, int patch_h, int estride)
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
    // with a traversal stride the box spans patch*stride source elements and ceil(box/stride) of them are loaded
    cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)(patch_w * estride), (cuuint32_t)(patch_h * estride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    const CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x.ptr, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) ZL_FAIL(ZL_SYSTEM_ERROR, "cuTensorMapEncodeTiled(4d) failed, CUresult " + std::to_string((int)r));
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && (uint64_t)x.n * strides[2] < 131072ull) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);   // see make_tmap_2d_16
    return ZL_OK;
}

int32_t make_tmap_out(CUtensorMap* map, const View& y, bool f16)   // 16-bit (32-B swizzle) or fp32 (64-B swizzle) output tiles
{
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        ZL_FAIL(ZL_SYSTEM_ERROR, "cuTensorMapEncodeTiled entry point not available");
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ptr);
    cuuint64_t dims[4] = {(cuuint64_t)y.c, (cuuint64_t)y.w, (cuuint64_t)y.h, (cuuint64_t)y.n};
    const uint64_t es = y.esize();
    cuuint64_t strides[3] = {(cuuint64_t)y.pitch * es, (cuuint64_t)y.w * y.pitch * es, (cuuint64_t)y.h * y.w * y.pitch * es};
    cuuint32_t box[4] = {16, (cuuint32_t)kTW, 4, 1};       // one epilogue warp: 16 channels x 8 px x 4 rows
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const bool f32 = y.dtype == DT_F32;
    CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 4, y.ptr, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, f32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) ZL_FAIL(ZL_SYSTEM_ERROR, "cuTensorMapEncodeTiled(out) failed, CUresult " + std::to_string((int)r));
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && (uint64_t)y.n * strides[2] < 131072ull) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);
    return ZL_OK;
}

}  // namespace

// Weight tensor map: w_tc is [cout_pad][ktot] with k = tap * cin + channel.  Seen as a 3-D tensor {channel-in-tap, Cout, tap}
// (strides ktot*2 and cin*2 bytes: TMA strides need not be monotonic), a box {kc, nt, taps} lands in shared memory as
// [tap][nt][kc] — every tap's [nt][kc] K-major tile packed behind the previous one — with ONE instruction per channel chunk.
int32_t make_tmap_w3d(CUtensorMap* map, const void* w, int cin, int cout_pad, int taps, int ktot, int kc, int nt, bool f16)
{
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        ZL_FAIL(ZL_SYSTEM_ERROR, "cuTensorMapEncodeTiled entry point not available");
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ptr);
    cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)cout_pad, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)ktot * 2, (cuuint64_t)cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)nt, (cuuint32_t)taps};
    cuuint32_t estr[3] = {1, 1, 1};
    const int swz = kc * 2;
    const CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) ZL_FAIL(ZL_SYSTEM_ERROR, "cuTensorMapEncodeTiled(weights 3d) failed, CUresult " + std::to_string((int)r));
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && (uint64_t)cout_pad * ktot * 2 < 131072ull) reinterpret_cast<uint64_t*>(map)[1] &= ~(1ull << 21);   // see make_tmap_2d_16
    return ZL_OK;
}

// Sub-tiles per tile: narrow layers stack several 8x16 sub-tiles along h so one barrier round trip, one
// patch and one epilogue pass cover more pixels.  An MMA with N <= 64 costs ~48 tensor-pipe cycles whatever
// N is (umma_probe), so stacking only pays for per-tile overheads; wide layers keep one sub-tile.
static int halo_sub(const ConvWeights& w, int rows)
{
    int sub = w.cout_pad <= 16 ? 4 : (w.cout_pad <= 32 ? 2 : 1);
    while (sub > 1 && kTH * sub > round_up(rows, kTH)) --sub;     // do not make tiles taller than the image
    return sub;
}

// 1x1 convs see their input as an 8-pixel-wide image of the flattened pixel list, so a tile is 128*sub
// consecutive pixels and no rows are wasted on small maps.
static bool persist_views(const ConvWeights& w, const View& x, const View& y, View* xv, View* yv)
{
    *xv = x; *yv = y;
    if (w.k == 3) return true;
    const size_t m = x.pixels();
    if (m % 8 != 0) return true;          // odd pixel counts (13x13 maps at b=1): keep the native image view, tiles 8 wide x 16 high
    xv->n = 1; xv->w = 8; xv->h = (int32_t)(m / 8);
    yv->n = 1; yv->w = 8; yv->h = (int32_t)(m / 8);
    return true;
}

struct PersistPlan {
    int taps, pw, prow_extra, npatch, kc, cchunks, sub, nsplit, nt, stages, smem, tiles, wstream, nacc, cps;
    uint32_t wtile_bytes, wchunk_bytes, wchunk_alloc, patch_bytes, patch_alloc, subpatch_alloc, chunk_stride, stage_stride;
    View xv, yv;
    double cost;
};

// Tensor-pipe cycles of one tcgen05.mma (M = 128, K = 16, both operands in shared memory), measured (umma_probe.cu):
// 46 / 46 / 48 / 64 / 127 for N = 16 / 32 / 64 / 128 / 256 — the operand reads (4 KB of A + 32*N bytes of B at 128 B/clk)
// below N = 128, the math above.
static double mma_cycles(int n) { return std::max(46.0, std::max(32.0 + n / 4.0, n / 2.0)); }

// Geometry + shared-memory plan.  Two ways to hold the weights:
//   resident : the CTA's slice [taps][nt][Cin] stays in shared memory for the whole launch next to >= 2-3 patch stages; if it
//              does not fit, Cout is split over `nsplit` CTAs per tile (each re-reads the patch from L2);
//   streamed : (3x3 only) every stage carries the patch chunk AND that chunk's weights [taps][nt][kc], so nt can stay at the
//              full Cout (<= 256) however large taps*Cin*Cout is — the deep 20x20 / 40x40 layers, where a resident slice
//              forces nt down to 16-48 and every MMA below N = 64 costs the same 46 cycles.
// kc is fixed per layer (conv_kc).  Every candidate (nsplit, mode) that fits is priced with a small cycle model and the cheapest one wins.
static bool persist_plan(const ConvWeights& w, const View& x, const View& y, int num_sms, PersistPlan* best)
{
    const bool s2 = w.k == 3 && w.stride == 2;
    if (!(((w.k == 3 || w.k == 1) && w.stride == 1) || s2) || !x.is16() || (w.cin % 16) != 0) return false;
    if (s2 && ((x.h & 1) || (x.w & 1))) return false;
    if ((x.pitch % 8) != 0 || (reinterpret_cast<uintptr_t>(x.ptr) & 15)) return false;
    if (y.is16() && y.dtype != x.dtype) return false;
    PersistPlan base{};
    if (!persist_views(w, x, y, &base.xv, &base.yv)) return false;
    base.taps = w.k * w.k;
    base.pw = s2 ? kTW + 1 : (w.k == 3 ? kTW + 2 : kTW);
    base.prow_extra = s2 ? 1 : (w.k == 3 ? 2 : 0);
    base.npatch = s2 ? 4 : 1;
    const int sub_default = halo_sub(w, base.yv.h);
    static const int sub_search = [] { const char* e = getenv("ZL_SUB_SEARCH"); return e ? atoi(e) : 0; }();   // 1 = let the cost model pick sub too (measured: no gain, so the rule stays "by Cout only")
    const uint32_t budget = 227u * 1024u;
    const uint32_t fixed0 = 3072u + (uint32_t)kEpiWarps * (y.dtype == DT_F32 ? 4096u : 2048u);
    static const int force_stream = [] { const char* e = getenv("ZL_WSTREAM"); return e ? atoi(e) : -1; }();   // A/B: 0 never, 1 whenever possible
    bool found = false;
    const int kc = conv_kc(w, y.dtype == DT_F32);              // fixed per layer: the accumulation order must not depend on the plan
    for (int sub : {1, 2, 4}) {                              // sub-tiles per tile: the per-tile overheads against shared memory and TMEM
        if (sub_search ? (sub > 1 && kTH * sub > round_up(base.yv.h, kTH)) : sub != sub_default) continue;
        PersistPlan pl = base;
        pl.sub = sub;
        pl.tiles = ceil_div(base.yv.w, kTW) * ceil_div(base.yv.h, kTH * sub) * base.yv.n;
        pl.kc = kc;
        pl.cchunks = w.cin / kc;
        pl.patch_bytes = (uint32_t)pl.pw * (kTH * pl.sub + pl.prow_extra) * kc * 2;      // one (sub-)patch
        pl.subpatch_alloc = (pl.patch_bytes + 1023u) & ~1023u;
        pl.patch_alloc = pl.subpatch_alloc * (uint32_t)pl.npatch;
        for (int mode = 0; mode < 2; ++mode) {                // 0 resident, 1 streamed
            if (mode == 1 && (w.k != 3 || force_stream == 0)) continue;
            if (mode == 0 && force_stream == 1 && w.k == 3) continue;
            for (int nsplit = 1; nsplit <= 16; ++nsplit) {
                const int nt = round_up(ceil_div(w.cout_pad, nsplit), 16);
                if (nsplit > 1 && nt * (nsplit - 1) >= w.cout_pad) continue;          // this split count adds nothing
                if (nt > 256 || 2 * pl.sub * nt > 512) continue;                      // UMMA N limit, >= two accumulators in TMEM
                pl.nsplit = nsplit; pl.nt = nt; pl.wstream = mode;
                pl.wtile_bytes = (uint32_t)nt * kc * 2;
                pl.wchunk_bytes = (uint32_t)pl.taps * pl.wtile_bytes;
                pl.wchunk_alloc = (pl.wchunk_bytes + 1023u) & ~1023u;
                pl.chunk_stride = pl.patch_alloc + (mode ? pl.wchunk_alloc : 0u);
                const uint32_t fixed = fixed0 + (mode ? 0u : (uint32_t)pl.cchunks * pl.wchunk_alloc);
                // chunks per stage: one barrier round trip (producer ~500 cycles, MMA warp ~250) then covers cps chunks
                for (int cps = 1; cps <= pl.cchunks; ++cps) {
                    if (pl.cchunks % cps) continue;
                    pl.cps = cps;
                    pl.stage_stride = pl.chunk_stride * (uint32_t)cps;
                    const int nst = pl.cchunks / cps;
                    const int min_stages = mode ? 2 : (nst + 1 < 3 ? nst + 1 : 3);
                    if (fixed + (uint32_t)min_stages * pl.stage_stride > budget) continue;
                    int stages = (int)((budget - fixed) / pl.stage_stride);
                    if (stages > kMaxPatchStages) stages = kMaxPatchStages;
                    pl.stages = stages;
                    pl.smem = (int)(fixed + (uint32_t)stages * pl.stage_stride);
                    int nacc = 2;
                    while (nacc * 2 <= kMaxAcc && nacc * 2 * pl.sub * nt <= 512) nacc *= 2;
                    pl.nacc = nacc;
                    // ---- price it (cycles per CTA): waves of (tile, slice) units, each the slower of its MMAs, its stage
                    //      traffic (L2 -> SM at ~48 B/clk sustained) and its barrier round trips
                    const double units = (double)pl.tiles * nsplit;
                    const int ctas = units < num_sms ? (int)units : (num_sms / nsplit) * nsplit;
                    const double waves = std::ceil(units / std::max(ctas, 1));
                    const double mma = (double)pl.sub * pl.taps * (w.cin / 16) * mma_cycles(nt) + 250.0 * nst + 600.0;   // + per-tile hand-over (accumulator wait, commits)
                    const double stage_bytes = (double)(pl.patch_bytes * pl.npatch + (mode ? pl.wchunk_bytes : 0u)) * cps;
                    const int np = std::min(kProducers, stages);
                    const double stage_cyc = std::max(stage_bytes / 48.0, (500.0 + 130.0 * cps * (pl.npatch + mode)) / np);
                    const double load = nst * stage_cyc;
                    const double epi = 500.0 + 250.0 * pl.sub * (nt / 16) / 4.0;            // per-tile epilogue: ~1000 cycles per item and warp, 8 warps per tile
                    const double fill = stage_cyc * 0.5;                                      // the first stage of a CTA is exposed
                    const double prologue = mode ? 1500.0 : 1500.0 + (double)pl.cchunks * pl.wchunk_bytes / 48.0;
                    pl.cost = prologue + fill + waves * std::max(std::max(mma, load), epi);
                    if (!found || pl.cost < best->cost * 0.97) { *best = pl; found = true; }   // ties: keep the earlier (resident, fewer splits, fewer chunks per stage)
                }
            }
        }
    }
    return found;
}

bool conv_halo_supported(const ConvWeights& w, const View& x, const View& y, int num_sms, int* work_units)
{
    PersistPlan pl;
    if (!persist_plan(w, x, y, num_sms, &pl)) return false;
    if (work_units) *work_units = pl.tiles * pl.nsplit;
    return true;
}

static void fill_ring(ConvHaloOp& o)
{
    o.acc_cols = o.sub * o.nt;
    int nacc = 2;
    while (nacc * 2 <= kMaxAcc && nacc * 2 * o.acc_cols <= 512) nacc *= 2;
    o.nacc = nacc;
    int cols = 32;
    while (cols < o.nacc * o.acc_cols) cols <<= 1;
    o.tmem_cols = cols;
}

int32_t conv_halo_prepare(const ConvWeights& w, const View& x, const View& y, const View* res, int num_sms, ConvHaloOp* op)
{
    PersistPlan pl;
    if (!persist_plan(w, x, y, num_sms, &pl)) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_persist: layer not supported (" + w.name + ")");
    if (y.h != x.h / w.stride || y.w != x.w / w.stride || y.n != x.n || y.c != w.cout || x.c != w.cin) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_persist: view mismatch (" + w.name + ")");
    if (res && (res->dtype != x.dtype || res->c != w.cout)) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_persist: residual view mismatch");
    ConvHaloOp& o = *op;
    o = ConvHaloOp{};
    o.taps = pl.taps;
    o.mode = (w.k == 1) ? 1 : (w.stride == 2 ? 2 : 9);
    o.y = y.ptr;
    o.res = res ? reinterpret_cast<const __nv_bfloat16*>(res->ptr) : nullptr;
    o.bias = w.bias;
    o.N = pl.yv.n; o.H = pl.yv.h; o.W = pl.yv.w;      // output geometry (tiles, clipping)
    o.Cin = w.cin; o.Cout = w.cout; o.ntile = w.cout_pad;
    o.ypitch = y.pitch; o.rpitch = res ? res->pitch : 0;
    o.y_f32 = y.dtype == DT_F32;
    o.f16 = x.dtype == DT_F16 ? 1 : 0;
    o.silu_tanh = conv_silu_tanh(o.f16 != 0) ? 1 : 0;
    o.act = w.act;
    o.y_vec = ((y.pitch * y.esize()) % 16 == 0 && (reinterpret_cast<uintptr_t>(y.ptr) & 15) == 0) ? 1 : 0;
    o.r_vec = (res && (res->pitch % 8) == 0 && (reinterpret_cast<uintptr_t>(res->ptr) & 15) == 0) ? 1 : 0;
    o.kc = pl.kc; o.cchunks = pl.cchunks; o.sub = pl.sub; o.nsplit = pl.nsplit; o.nt = pl.nt;
    o.tiles_x = ceil_div(pl.yv.w, kTW); o.tiles_y = ceil_div(pl.yv.h, kTH * o.sub);
    o.num_tiles = pl.tiles;
    o.wstream = pl.wstream;
    o.wtile_bytes = pl.wtile_bytes; o.wchunk_bytes = pl.wchunk_bytes; o.wchunk_alloc = pl.wchunk_alloc;
    o.patch_bytes = pl.patch_bytes; o.patch_alloc = pl.patch_alloc; o.subpatch_alloc = pl.subpatch_alloc; o.chunk_stride = pl.chunk_stride; o.stage_stride = pl.stage_stride;
    o.cps = pl.cps; o.nst = pl.cchunks / pl.cps;
    o.smem_bytes = pl.smem; o.stages = pl.stages;
    fill_ring(o);
    static const bool plan_debug = [] { const char* e = getenv("ZL_PLAN_DEBUG"); return e && e[0] == '1'; }();
    if (plan_debug)
        fprintf(stderr, "plan %-26s k%d s%d cin %d cout %d | %s kc %d chunks %d cps %d sub %d nsplit %d nt %d stages %d nacc %d tiles %d smem %d stage %u wchunk %u cost %.0f\n",
                w.name.c_str(), w.k, w.stride, w.cin, w.cout, pl.wstream ? "stream" : "resident", pl.kc, pl.cchunks, pl.cps, pl.sub, pl.nsplit, pl.nt, pl.stages, o.nacc,
                pl.tiles, pl.smem, pl.stage_stride, pl.wchunk_bytes, pl.cost);
    // weight box = one channel chunk of one slice of nt output channels, all taps (rows past Cout_pad are zero-filled by TMA)
    ZL_TRY(make_tmap_w3d(&o.tmap_w, w.w_tc, w.cin, w.cout_pad, pl.taps, w.ktot, o.kc, o.nt, o.f16));
    ZL_TRY(make_tmap_nhwc(&o.tmap_x, pl.xv, o.kc, o.kc * 2, o.f16, pl.pw, kTH * o.sub + pl.prow_extra, w.stride));
    o.y_tma = ((y.pitch * y.esize()) % 16 == 0 && (reinterpret_cast<uintptr_t>(y.ptr) & 15) == 0) ? 1 : 0;
    o.ostage = y.dtype == DT_F32 ? 2048 : 1024;
    if (o.y_tma) ZL_TRY(make_tmap_out(&o.tmap_y, pl.yv, o.f16)); else o.tmap_y = o.tmap_x;
    o.flops = 2.0 * (double)y.pixels() * w.cout * w.ktot;
    o.bytes = (double)x.pixels() * w.cin * 2 + (double)y.pixels() * w.cout * (o.y_f32 ? 4 : 2) + (double)w.cout * w.ktot * 2 +
              (res ? (double)y.pixels() * w.cout * 2 : 0.0);
    return ZL_OK;
}

// Layer 0 on the space-to-depth image (MODE 4).  x = [B, H/2, W/2, 16] 16-bit written by the preprocess kernel
// (PRE_S2D16), w.w_tc = [cout_pad][4 taps x 16] (engine.cpp builds it from the 3x3 stride-2 filter).
int32_t conv_s2d_prepare(const ConvWeights& w, const View& x, const View& y, int num_sms, ConvHaloOp* op)
{
    (void)num_sms;
    if (w.cin != 3 || w.k != 3 || w.stride != 2 || !w.w_tc || !x.is16() || y.dtype != x.dtype || x.c != 16 || x.pitch != 16 || w.cout_pad > 128)
        ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_s2d: expects the 3->c 3x3 s2 first layer on the 16-channel space-to-depth image");
    if (y.h != x.h || y.w != x.w || y.n != x.n || y.c != w.cout) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_s2d: view mismatch");
    ConvHaloOp& o = *op;
    o = ConvHaloOp{};
    o.taps = 4; o.mode = 4;
    o.y = y.ptr; o.res = nullptr; o.bias = w.bias;
    o.N = y.n; o.H = y.h; o.W = y.w; o.Cin = 16; o.Cout = w.cout; o.ntile = w.cout_pad;
    o.ypitch = y.pitch; o.rpitch = 0; o.y_f32 = 0; o.f16 = y.dtype == DT_F16 ? 1 : 0; o.act = w.act; o.silu_tanh = conv_silu_tanh(o.f16 != 0) ? 1 : 0;
    o.y_vec = 1; o.r_vec = 0;
    o.kc = 16; o.cchunks = 1; o.nsplit = 1; o.nt = w.cout_pad;
    o.sub = halo_sub(w, y.h);
    o.tiles_x = ceil_div(y.w, kTW); o.tiles_y = ceil_div(y.h, kTH * o.sub);
    o.num_tiles = o.tiles_x * o.tiles_y * y.n;
    o.wstream = 0;
    o.wtile_bytes = (uint32_t)o.nt * 32u; o.wchunk_bytes = 4u * o.wtile_bytes; o.wchunk_alloc = (o.wchunk_bytes + 1023u) & ~1023u;
    o.patch_bytes = (uint32_t)(kTW + 1) * (kTH * o.sub + 1) * 32u;
    o.subpatch_alloc = (o.patch_bytes + 1023u) & ~1023u; o.patch_alloc = o.subpatch_alloc; o.chunk_stride = o.patch_alloc; o.stage_stride = o.patch_alloc;
    o.cps = 1; o.nst = 1;
    const uint32_t fixed = 3072u + o.wchunk_alloc + (uint32_t)kEpiWarps * 2048u;
    int stages = (int)((227u * 1024u - fixed) / o.stage_stride);
    if (stages > kMaxPatchStages) stages = kMaxPatchStages;
    o.stages = stages;
    o.smem_bytes = (int)(fixed + (uint32_t)stages * o.stage_stride);
    o.ostage = 1024;
    fill_ring(o);
    ZL_TRY(make_tmap_w3d(&o.tmap_w, w.w_tc, 16, w.cout_pad, 4, 64, 16, o.nt, o.f16));
    ZL_TRY(make_tmap_nhwc(&o.tmap_x, x, 16, 32, o.f16, kTW + 1, kTH * o.sub + 1, 1));
    o.y_tma = ((y.pitch % 8) == 0 && (reinterpret_cast<uintptr_t>(y.ptr) & 15) == 0) ? 1 : 0;
    if (!o.y_tma) ZL_FAIL(ZL_INVALID_ARGUMENT, "conv_s2d: output must be 16-byte aligned");
    ZL_TRY(make_tmap_out(&o.tmap_y, y, o.f16));
    o.flops = 2.0 * (double)y.pixels() * w.cout * 27;
    o.bytes = (double)x.pixels() * 32 + (double)y.pixels() * w.cout * 2;
    return ZL_OK;
}

static const bool g_use_pdl = [] { const char* e = getenv("ZL_DISABLE_PDL"); return !(e && e[0] == '1'); }();

static int grid_for(const ConvHaloOp& o, int num_sms)
{
    int grid = o.num_tiles * o.nsplit;
    if (grid > num_sms) grid = (num_sms / o.nsplit) * o.nsplit;       // every CTA keeps one slice: grid is a multiple of nsplit
    return grid;
}

int32_t conv_halo_launch(cudaStream_t st, const ConvHaloOp& o, int num_sms, unsigned long long* stats)
{
    static thread_local int last_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != last_dev) {
        ZL_CUDA(cudaFuncSetAttribute(conv_halo_kernel<9, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ZL_CUDA(cudaFuncSetAttribute(conv_halo_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ZL_CUDA(cudaFuncSetAttribute(conv_halo_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ZL_CUDA(cudaFuncSetAttribute(conv_halo_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ZL_CUDA(cudaFuncSetAttribute(conv_halo_kernel<9, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ZL_CUDA(cudaFuncSetAttribute(conv_halo_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ZL_CUDA(cudaFuncSetAttribute(conv_halo_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ZL_CUDA(cudaFuncSetAttribute(conv_halo_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        last_dev = dev;
    }
    HaloParams p;
    p.y = o.y; p.res = o.res; p.bias = o.bias;
    p.N = o.N; p.H = o.H; p.W = o.W; p.Cin = o.Cin; p.Cout = o.Cout; p.ntile = o.ntile;
    p.ypitch = o.ypitch; p.rpitch = o.rpitch; p.y_f32 = o.y_f32; p.y_vec = o.y_vec; p.r_vec = o.r_vec; p.f16 = o.f16; p.act = o.act; p.silu_tanh = o.silu_tanh;
    p.kc = o.kc; p.cchunks = o.cchunks; p.stages = o.stages; p.sub = o.sub; p.y_tma = o.y_tma; p.nsplit = o.nsplit; p.nt = o.nt; p.ostage = o.ostage;
    p.tiles_x = o.tiles_x; p.tiles_y = o.tiles_y; p.num_tiles = o.num_tiles;
    p.wtile_bytes = o.wtile_bytes; p.wchunk_bytes = o.wchunk_bytes; p.wchunk_alloc = o.wchunk_alloc;
    p.patch_bytes = o.patch_bytes; p.patch_alloc = o.patch_alloc; p.subpatch_alloc = o.subpatch_alloc; p.chunk_stride = o.chunk_stride; p.stage_stride = o.stage_stride;
    p.cps = o.cps; p.nst = o.nst;
    p.wstream = o.wstream; p.nacc = o.nacc; p.acc_cols = o.acc_cols;
    p.esets = (o.num_tiles * o.nsplit > grid_for(o, num_sms)) ? 2 : 1;      // more than one tile per CTA: two alternating sets
    // 256-bit global stores instead of staged bulk stores.  Measured (profiles/README_r02.md): a few us faster on the 1x1 layers
    // with 16-bit outputs, slower on fp32 outputs and on layer 0; and bulk stores of 32-byte rows into a concat slice that starts
    // INSIDE a 128-byte line are 1.3x slower than the same stores into an aligned slice (YOLOv8m, c = 96: 509 vs 386 us), which
    // direct stores are not.  Default 3 = 16-bit outputs of 1x1 layers and of slices off a 128-byte boundary; 1 = 1x1 layers
    // only; 2 = wherever aligned to 32 bytes; 0 = never.
    static const int direct_env = [] { const char* e = getenv("ZL_EPI_DIRECT"); return e ? atoi(e) : 3; }();
    const bool direct_ok = (o.ypitch * (o.y_f32 ? 4 : 2)) % 32 == 0 && (reinterpret_cast<uintptr_t>(o.y) & 31) == 0;
    const bool slice_off_line = (reinterpret_cast<uintptr_t>(o.y) & 127) != 0;     // output slice starts inside a 128-byte line
    p.direct_store = (direct_ok && (direct_env == 2 || (direct_env == 1 && o.mode == 1 && !o.y_f32) ||
                                    (direct_env == 3 && !o.y_f32 && (o.mode == 1 || slice_off_line)))) ? 1 : 0;
    p.epi_variant = 6;
    if (o.y_tma && o.Cout % 16 == 0 && !getenv("ZL_EPI_GENERIC") && (o.silu_tanh || !o.act)) {   // ZL_SILU=exp: generic body (it keeps both SiLU forms; 2 KB of never-run code sat in every fast loop)
        const int fmt = o.f16 ? 3 : 0;
        if (!o.y_f32 && o.act && !o.res) p.epi_variant = fmt + 0;
        else if (!o.y_f32 && o.act && o.res && o.r_vec) p.epi_variant = fmt + 1;
        else if (o.y_f32 && !o.act && !o.res) p.epi_variant = fmt + 2;
    }
    p.nacc_log2 = o.nacc == 8 ? 3 : (o.nacc == 4 ? 2 : (o.nacc == 2 ? 1 : 0));
    p.tmem_cols = o.tmem_cols;
    p.stats = stats;
    const int grid = grid_for(o, num_sms);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = (size_t)o.smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_use_pdl ? 1 : 0;
    if (stats) {
        if (o.mode == 4) ZL_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<4, true>, o.tmap_w, o.tmap_x, o.tmap_y, p));
        else if (o.mode == 9) ZL_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<9, true>, o.tmap_w, o.tmap_x, o.tmap_y, p));
        else if (o.mode == 2) ZL_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<2, true>, o.tmap_w, o.tmap_x, o.tmap_y, p));
        else ZL_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, true>, o.tmap_w, o.tmap_x, o.tmap_y, p));
        return ZL_OK;
    }
    if (o.mode == 4) ZL_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<4, false>, o.tmap_w, o.tmap_x, o.tmap_y, p));
    else if (o.mode == 9) ZL_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<9, false>, o.tmap_w, o.tmap_x, o.tmap_y, p));
    else if (o.mode == 2) ZL_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<2, false>, o.tmap_w, o.tmap_x, o.tmap_y, p));
    else ZL_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, false>, o.tmap_w, o.tmap_x, o.tmap_y, p));
    return ZL_OK;
}

}  // namespace zl
