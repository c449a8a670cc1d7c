// head_fused.cu — the last two 1x1 convolutions of the Detect head (model.22.cv2.l.2: 64 -> 64 DFL logits, model.22.cv3.l.2:
// c3 -> nc class logits), the Detect tail (D1) and the decode / threshold loop (F1) as ONE persistent tcgen05 kernel for
// all three levels (sm_100a only).
//
// Reference nodes: the two Conv nodes and the DFL / sigmoid / dist2bbox tail inside Ort::Session::Run
// (src/inference/onnx_engine.cpp:577-585) followed by the decode loop of postProcess (onnx_engine.cpp:773-819).
// In the unfused chain the two convs write their fp32 logits ([pixels][64] and [pixels][nc]) and decode_filter_kernel reads
// them back; here the logits never leave the SM:
//   * a tile = 128 consecutive pixels of one level's flattened [n*h*w] pixel list; the tiles of the three levels form one
//     list that is dealt round-robin to the CTAs (every CTA then holds all three levels' weights; when they do not fit,
//     CTAs are dedicated to one level in proportion to the tile counts).  Two TMA loads bring the tile's rows of the two
//     branch inputs (HB2_l / HC2_l, 16-bit NHWC) into shared memory in the swizzled K-major UMMA layout — a 3-D box
//     {kc, 128, K/kc} lands as [chunk][128][kc], one instruction per branch;
//   * the MMA warp accumulates [128 x 64] and [128 x nc_pad] into one TMEM slot (columns 0..63 and 64..), with the k-steps
//     in ascending order == the accumulation order of the standalone conv kernels, so the logits are the same bits;
//   * per TMEM lane quarter two warps drain the slot: the SCAN warp reads its rows' class logits (tcgen05.ld 32x32b), adds
//     the bias and finds the best class exactly like decode_filter_kernel; the DFL warp reads the 64 box logits and computes
//     the four expected distances of EVERY row (head_math.cuh) into shared memory; the scan warp appends the candidates
//     with the same key / box arithmetic as the unfused chain.
// Algorithmic HBM bytes: (64 + c3) x 2 per anchor read, nothing written but the candidates (vs. + 2 x (64 + nc) x 4).
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cstdio>
#include <cstdlib>

#include "head_math.cuh"
#include "kernels.h"
#include "tc_ptx.cuh"

namespace zl {
namespace {

using namespace tc;

constexpr int kHfEpiWarps = 8;               // per TMEM lane quarter: one SCAN warp (classes, candidates) + one DFL warp (box distances)
constexpr int kHfThreads = 320;              // warps: 0 producer, 1 MMA, 2..5 scan, 6..9 DFL
constexpr int kHfMaxStages = 8;
constexpr int kHfTile = 128;
constexpr uint32_t kHfBiasOff = 512u;         // fp32 bias [3 levels][256]: 64 box | nc_pad class
constexpr uint32_t kHfDistOff = 4096u;        // float4 box distances [2 buffers][128 rows], DFL warps -> scan warps
constexpr uint32_t kHfHeader = 8192u;         // barriers | bias | box distances, in front of the weights

struct HfLevel {
    int32_t npix, hw, w, stride, a0, ntiles, g0, cta0, ctas;      // g0 = index of the level's first tile in the global tile list
    const float* bias_box;
    const float* bias_cls;
};
struct HfParams {
    HfLevel lv[3];
    int32_t nc, nc_pad, A, key_pitch, f16, all_levels, total_tiles;
    int32_t kcb, nchb, kcc, nchc;                        // K chunking of the box (K = 64) and class (K = c3) branches
    int32_t stages, nacc, nacc_log2, acc_cols;
    uint32_t tmem_cols, wb_bytes, wc_bytes, wc_off, w_alloc, xb_bytes, xc_bytes, xc_off, stage_stride;
    const FrameDesc* descs;
    float conf_thr;
    const float* class_weights;
    uint64_t* keys;
    float4* box_by_anchor;
    uint32_t* cand_count;
    unsigned long long* stats;                           // STATS instantiation only (slot meanings: kernels.h, conv_halo_launch)
};
struct HfMaps { CUtensorMap m[12]; };        // [level * 4 + {0: box input, 1: class input, 2: box weights, 3: class weights}]

__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// ---- class scan of one tile row (the thread's anchor).  Same decisions as decode_filter_kernel:
//   zmax over the biased logits; only classes within 0.05 of min(zmax, 8) can hold the largest score; among them the
//   scores are compared in ascending class order with strict '>' (onnx_engine.cpp:787-796).
// Round-2 measurements behind this shape (profiles/README_r02.md):
//   * a tcgen05.ld + wait::ld round trip costs ~200 cycles next to the running MMAs, so the common class counts keep the
//     whole row in registers after ONE round trip (scan_classes_reg<NCH>: 16 or 80 columns);
//   * a first version with every pass of every class count unrolled was INSTRUCTION-FETCH bound (80 KB of straight-line
//     code per tile, stall_no_instruction on top, IPC 0.2): the L1.5 instruction cache is 32 KB.  So only pass 1 + 2 of
//     the register form are unrolled; everything rare is a rolled loop that re-reads TMEM.
//   * pass 1 = bias add + 3-input max.  Pass 2 (only warps where some anchor can reach the threshold) = count the classes
//     above the cut and remember the lowest one, branch-free.  A single class above the cut IS the class of zmax, so its
//     score is cls_score(zmax); only warps where some anchor has two or more classes above the cut run the comparison loop.
//   Padded classes carry a bias of -FLT_MAX: never the maximum, never above the cut.
__device__ __forceinline__ void scan_compare_loop(uint32_t tcls, const float* __restrict__ bias_c, int nch, float zcut, float& best, int& best_id)
{
    best = 0.0f; best_id = -1;
#pragma unroll 1
    for (int ch = 0; ch < nch; ++ch) {
        uint32_t va[16];
        tmem_ld16(tcls + (uint32_t)(ch << 4), va);
        tmem_ld_wait();
        const float* bc = bias_c + (ch << 4);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float z = __uint_as_float(va[i]) + bc[i];
            if (z >= zcut) {
                const float s = cls_score<false>(z);
                if (s > best) { best = s; best_id = (ch << 4) + i; }   // strict '>': first maximum wins (onnx_engine.cpp:792)
            }
        }
    }
}

template <int NCH>
__device__ __forceinline__ void scan_classes_reg(uint32_t tcls, const float* __restrict__ bias_c, bool in_range, float conf_thr, float& best, int& best_id)
{
    uint32_t v[NCH][16];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) tmem_ld16(tcls + (uint32_t)(ch << 4), v[ch]);
    tmem_ld_wait();
    float zm0 = -FLT_MAX, zm1 = -FLT_MAX;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias_c + (ch << 4) + (q4 << 2));
            const float z0 = __uint_as_float(v[ch][(q4 << 2) + 0]) + b4.x, z1 = __uint_as_float(v[ch][(q4 << 2) + 1]) + b4.y;
            const float z2 = __uint_as_float(v[ch][(q4 << 2) + 2]) + b4.z, z3 = __uint_as_float(v[ch][(q4 << 2) + 3]) + b4.w;
            v[ch][(q4 << 2) + 0] = __float_as_uint(z0); v[ch][(q4 << 2) + 1] = __float_as_uint(z1);
            v[ch][(q4 << 2) + 2] = __float_as_uint(z2); v[ch][(q4 << 2) + 3] = __float_as_uint(z3);
            zm0 = fmax3(zm0, z0, z1);
            zm1 = fmax3(zm1, z2, z3);
        }
    }
    const float zmax = fmaxf(zm0, zm1);
    // no class of this warp's 32 anchors can reach the threshold: done (the fast sigmoid is monotone to within a few ulp;
    // 1e-5 relative is a wide margin on the safe side)
    const float smax = cls_score<false>(zmax);
    const bool possible = in_range && smax >= conf_thr * (1.0f - 1e-5f);
    if (!__any_sync(0xffffffffu, possible)) return;
    const float zcut = fminf(zmax, 8.0f) - 0.05f;
    int cnt0 = 0, cnt1 = 0, lo0 = 1 << 20, lo1 = 1 << 20;
#pragma unroll
    for (int ch = NCH - 1; ch >= 0; --ch) {                           // descending: the last hit written is the lowest class
#pragma unroll
        for (int i = 15; i >= 1; i -= 2) {
            const bool h1 = __uint_as_float(v[ch][i]) >= zcut, h0 = __uint_as_float(v[ch][i - 1]) >= zcut;
            cnt1 += h1 ? 1 : 0; lo1 = h1 ? (ch << 4) + i : lo1;
            cnt0 += h0 ? 1 : 0; lo0 = h0 ? (ch << 4) + i - 1 : lo0;
        }
    }
    if (smax > 0.0f) { best = smax; best_id = min(lo0, lo1); }         // exactly one class above the cut: the class of zmax
    if (__any_sync(0xffffffffu, possible && cnt0 + cnt1 > 1)) scan_compare_loop(tcls, bias_c, NCH, zcut, best, best_id);
}

// Any class count: rolled loops, two 16-column chunks per TMEM round trip.
__device__ __forceinline__ void scan_classes_loop(uint32_t tcls, const float* __restrict__ bias_c, int nch, bool in_range, float conf_thr,
                                                  float& best, int& best_id)
{
    float zm0 = -FLT_MAX, zm1 = -FLT_MAX;
#pragma unroll 1
    for (int ch = 0; ch < nch; ch += 2) {
        uint32_t va[16], vb[16];
        const bool two = ch + 1 < nch;                                 // warp-uniform
        tmem_ld16(tcls + (uint32_t)(ch << 4), va);
        if (two) tmem_ld16(tcls + (uint32_t)((ch + 1) << 4), vb);
        tmem_ld_wait();
        const float* bc = bias_c + (ch << 4);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            const float4 b4 = *reinterpret_cast<const float4*>(bc + (q4 << 2));
            zm0 = fmax3(zm0, __uint_as_float(va[(q4 << 2) + 0]) + b4.x, __uint_as_float(va[(q4 << 2) + 1]) + b4.y);
            zm1 = fmax3(zm1, __uint_as_float(va[(q4 << 2) + 2]) + b4.z, __uint_as_float(va[(q4 << 2) + 3]) + b4.w);
        }
        if (two) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bc + 16 + (q4 << 2));
                zm0 = fmax3(zm0, __uint_as_float(vb[(q4 << 2) + 0]) + b4.x, __uint_as_float(vb[(q4 << 2) + 1]) + b4.y);
                zm1 = fmax3(zm1, __uint_as_float(vb[(q4 << 2) + 2]) + b4.z, __uint_as_float(vb[(q4 << 2) + 3]) + b4.w);
            }
        }
    }
    const float zmax = fmaxf(zm0, zm1);
    const float smax = cls_score<false>(zmax);
    const bool possible = in_range && smax >= conf_thr * (1.0f - 1e-5f);
    if (!__any_sync(0xffffffffu, possible)) return;
    const float zcut = fminf(zmax, 8.0f) - 0.05f;
    int cnt = 0, first = 1 << 20;
#pragma unroll 1
    for (int ch = (nch - 1) & ~1; ch >= 0; ch -= 2) {                  // descending: the last hit written is the lowest class
        uint32_t va[16], vb[16];
        const bool two = ch + 1 < nch;
        tmem_ld16(tcls + (uint32_t)(ch << 4), va);
        if (two) tmem_ld16(tcls + (uint32_t)((ch + 1) << 4), vb);
        tmem_ld_wait();
        const float* bc = bias_c + (ch << 4);
        if (two) {
#pragma unroll
            for (int i = 15; i >= 0; --i) {
                const bool hit = __uint_as_float(vb[i]) + bc[16 + i] >= zcut;
                cnt += hit ? 1 : 0;
                first = hit ? (ch << 4) + 16 + i : first;
            }
        }
#pragma unroll
        for (int i = 15; i >= 0; --i) {
            const bool hit = __uint_as_float(va[i]) + bc[i] >= zcut;
            cnt += hit ? 1 : 0;
            first = hit ? (ch << 4) + i : first;
        }
    }
    if (smax > 0.0f) { best = smax; best_id = first; }
    if (__any_sync(0xffffffffu, possible && cnt > 1)) scan_compare_loop(tcls, bias_c, nch, zcut, best, best_id);
}

// All MMAs of one branch of one tile, fully unrolled: NCH channel chunks x KS k-steps, ascending k.
template <int NCH, int KS>
__device__ __forceinline__ void issue_branch(uint32_t tmem_d, uint64_t a0, uint64_t b0, uint32_t a_chunk16, uint32_t b_chunk16, uint32_t idesc)
{
#pragma unroll
    for (int g = 0; g < NCH; ++g)
#pragma unroll
        for (int k = 0; k < KS; ++k)
            umma_bf16(tmem_d, a0 + (uint64_t)((uint32_t)g * a_chunk16 + 2u * (uint32_t)k), b0 + (uint64_t)((uint32_t)g * b_chunk16 + 2u * (uint32_t)k), idesc,
                      (g | k) ? 1u : 0u);
}
__device__ __forceinline__ void issue_branch_any(int nch, int ks, uint32_t tmem_d, uint64_t a0, uint64_t b0, uint32_t a_chunk16, uint32_t b_chunk16, uint32_t idesc)
{
    const int sel = nch * 8 + ks;
    switch (sel) {
        case 1 * 8 + 4: issue_branch<1, 4>(tmem_d, a0, b0, a_chunk16, b_chunk16, idesc); break;    // K = 64
        case 2 * 8 + 4: issue_branch<2, 4>(tmem_d, a0, b0, a_chunk16, b_chunk16, idesc); break;    // K = 128 (yolov8s classes)
        case 3 * 8 + 4: issue_branch<3, 4>(tmem_d, a0, b0, a_chunk16, b_chunk16, idesc); break;    // K = 192 (yolov8m)
        case 5 * 8 + 1: issue_branch<5, 1>(tmem_d, a0, b0, a_chunk16, b_chunk16, idesc); break;    // K = 80 (yolov8n, nc = 80)
        default:
            for (int g = 0; g < nch; ++g)
                for (int k = 0; k < ks; ++k)
                    umma_bf16(tmem_d, a0 + (uint64_t)((uint32_t)g * a_chunk16 + 2u * (uint32_t)k), b0 + (uint64_t)((uint32_t)g * b_chunk16 + 2u * (uint32_t)k), idesc,
                              (g | k) ? 1u : 0u);
    }
}

// Tile g of the global list -> its level.
__device__ __forceinline__ int level_of(const HfParams& p, int g) { return g >= p.lv[2].g0 ? 2 : (g >= p.lv[1].g0 ? 1 : 0); }

template <bool STATS>
__global__ void __launch_bounds__(kHfThreads, 1)
head_decode_kernel(const __grid_constant__ HfMaps maps, const HfParams p)
{
    extern __shared__ uint8_t smem_raw[];
    long long st_a = 0, st_b = 0, st_c = 0, st_d = 0, st_e = 0, t_begin = 0, t_pro = 0;
    if (STATS) t_begin = clock64();
#define ZL_ST_BEGIN(t) long long t = 0; if (STATS) t = clock64()
#define ZL_ST_END(t, acc) if (STATS) acc += clock64() - t
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_pfull = base, bar_pempty = base + 8u * kHfMaxStages, bar_wfull = base + 16u * kHfMaxStages;
    const uint32_t bar_tfull = bar_wfull + 8u, bar_tempty = bar_tfull + 32u, tmem_slot = bar_tempty + 32u;
    const uint32_t wbase = base + kHfHeader;
    const int nslots = p.all_levels ? 3 : 1;                         // weight / bias sets resident in this CTA
    const uint32_t pbase = wbase + (uint32_t)nslots * p.w_alloc;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* bias_s = reinterpret_cast<float*>(smem_raw + (base + kHfBiasOff - smem_u32(smem_raw)));

    // this CTA's tiles: g = g_first, g_first + g_step, ... < g_end in the global tile list (level 0 | level 1 | level 2)
    const int bid = (int)blockIdx.x;
    int g_first, g_step, g_end, own = 0;
    if (p.all_levels) { g_first = bid; g_step = (int)gridDim.x; g_end = p.total_tiles; }
    else {
        own = bid >= p.lv[2].cta0 ? 2 : (bid >= p.lv[1].cta0 ? 1 : 0);
        g_first = p.lv[own].g0 + bid - p.lv[own].cta0; g_step = p.lv[own].ctas; g_end = p.lv[own].g0 + p.lv[own].ntiles;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(bar_pfull + 8u * s, 1u); mbar_init(bar_pempty + 8u * s, 1u); }
        mbar_init(bar_wfull, 1u);
        for (int a = 0; a < p.nacc; ++a) { mbar_init(bar_tfull + 8u * a, 1u); mbar_init(bar_tempty + 8u * a, (uint32_t)kHfEpiWarps); }
        fence_barrier_init();
        for (int i = 0; i < 12; ++i) tma_prefetch_desc(&maps.m[i]);
    }
    if (warp == 1) { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
    for (int i = threadIdx.x; i < nslots * 256; i += (int)blockDim.x) {      // padded classes: -FLT_MAX (see scan_classes_reg)
        const int sl = i >> 8, c = i & 255;
        const HfLevel& lv = p.lv[p.all_levels ? sl : own];
        float b = 0.0f;
        if (c < 64) b = lv.bias_box[c];
        else if (c - 64 < p.nc) b = lv.bias_cls[c - 64];
        else b = -FLT_MAX;
        bias_s[i] = b;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (STATS) t_pro = clock64();

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            // the weights, once (constants: before the dependency wait)
            mbar_arrive_expect_tx(bar_wfull, (uint32_t)nslots * (p.wb_bytes + p.wc_bytes));
            for (int sl = 0; sl < nslots; ++sl) {
                const int l = p.all_levels ? sl : own;
                tma_load_3d(&maps.m[l * 4 + 2], bar_wfull, wbase + (uint32_t)sl * p.w_alloc, 0, 0, 0);
                tma_load_3d(&maps.m[l * 4 + 3], bar_wfull, wbase + (uint32_t)sl * p.w_alloc + p.wc_off, 0, 0, 0);
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            uint32_t s = 0, ph = 0;
            for (int g = g_first; g < g_end; g += g_step) {
                const int l = level_of(p, g);
                const int tile = g - p.lv[l].g0;
                { ZL_ST_BEGIN(t0); mbar_wait(bar_pempty + 8u * s, ph ^ 1u, 11); ZL_ST_END(t0, st_a); }
                mbar_arrive_expect_tx(bar_pfull + 8u * s, p.xb_bytes + p.xc_bytes);
                const uint32_t stage = pbase + s * p.stage_stride;
                tma_load_3d(&maps.m[l * 4 + 0], bar_pfull + 8u * s, stage, 0, tile * kHfTile, 0);
                tma_load_3d(&maps.m[l * 4 + 1], bar_pfull + 8u * s, stage + p.xc_off, 0, tile * kHfTile, 0);
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
            }
            if (STATS) atomicAdd(p.stats + 0, (unsigned long long)st_a);
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t fmt = p.f16 ? 0u : 1u;
        const uint32_t idesc_b = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_c = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.nc_pad >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t swzb = (uint32_t)p.kcb * 2u, swzc = (uint32_t)p.kcc * 2u;
        const uint32_t xb_chunk16 = (uint32_t)(kHfTile * p.kcb * 2) >> 4, xc_chunk16 = (uint32_t)(kHfTile * p.kcc * 2) >> 4;
        const uint32_t wb_chunk16 = (uint32_t)(64 * p.kcb * 2) >> 4, wc_chunk16 = (uint32_t)(p.nc_pad * p.kcc * 2) >> 4;
        const uint64_t ab0 = make_smem_desc(pbase, swzb), ac0 = make_smem_desc(pbase + p.xc_off, swzc);
        const uint64_t bb0 = make_smem_desc(wbase, swzb), bc0 = make_smem_desc(wbase + p.wc_off, swzc);
        const uint32_t stage16 = p.stage_stride >> 4, wslot16 = p.w_alloc >> 4;
        const int ksb = p.kcb / 16, ksc = p.kcc / 16;
        { ZL_ST_BEGIN(t0); mbar_wait(bar_wfull, 0u, 12); ZL_ST_END(t0, st_d); }
        tc_fence_after();
        uint32_t s = 0, ph = 0, tl = 0;
        for (int g = g_first; g < g_end; g += g_step, ++tl) {
            const uint32_t wsl = p.all_levels ? (uint32_t)level_of(p, g) : 0u;
            const uint32_t acc = tl & (uint32_t)(p.nacc - 1), aph = (tl >> p.nacc_log2) & 1u;
            { ZL_ST_BEGIN(t0); mbar_wait(bar_tempty + 8u * acc, aph ^ 1u, 13); ZL_ST_END(t0, st_b); }
            { ZL_ST_BEGIN(t0); mbar_wait(bar_pfull + 8u * s, ph, 14); ZL_ST_END(t0, st_a); }
            tc_fence_after();
            ZL_ST_BEGIN(t_issue);
            if (elect_one()) {
                const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.acc_cols;
                issue_branch_any(p.nchb, ksb, tmem_d, ab0 + (uint64_t)(s * stage16), bb0 + (uint64_t)(wsl * wslot16), xb_chunk16, wb_chunk16, idesc_b);
                issue_branch_any(p.nchc, ksc, tmem_d + 64u, ac0 + (uint64_t)(s * stage16), bc0 + (uint64_t)(wsl * wslot16), xc_chunk16, wc_chunk16, idesc_c);
                umma_commit(bar_pempty + 8u * s);
                umma_commit(bar_tfull + 8u * acc);
            }
            __syncwarp();
            ZL_ST_END(t_issue, st_c);
            if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
        }
        if (STATS && lane == 0) {
            atomicAdd(p.stats + 1, (unsigned long long)st_a);
            atomicAdd(p.stats + 2, (unsigned long long)st_b);
            atomicAdd(p.stats + 3, (unsigned long long)st_c);
            atomicAdd(p.stats + 8, (unsigned long long)st_d);
        }
    } else if (warp < 6u) {
        // ===== scan warps (one per TMEM lane quarter): class logits (TMEM) -> best class -> threshold -> candidate =====
        const uint32_t q = warp & 3u;                                 // TMEM lane quarter this warp may read
        const int row = (int)(q * 32u + lane);
        const bool leader = elect_one();
        const int nch = p.nc_pad >> 4;
        const float* cw = p.class_weights;
        const float conf_thr = p.conf_thr;
        const float4* dist_s = reinterpret_cast<const float4*>(smem_raw + (base + kHfDistOff - smem_u32(smem_raw)));
        asm volatile("griddepcontrol.wait;" ::: "memory");            // cand_count / keys belong to earlier work of the stream
        int tl = 0;
#pragma unroll 1
        for (int g = g_first; g < g_end; g += g_step, ++tl) {
            const int l = level_of(p, g);
            const HfLevel& lv = p.lv[l];
            const float* bias_c = bias_s + (p.all_levels ? l : 0) * 256 + 64;
            const uint32_t acc = (uint32_t)tl & (uint32_t)(p.nacc - 1), aph = ((uint32_t)tl >> p.nacc_log2) & 1u;
            const int pix = (g - lv.g0) * kHfTile + row;
            const bool in_range = pix < lv.npix;
            { ZL_ST_BEGIN(t0); mbar_wait(bar_tfull + 8u * acc, aph, 15); ZL_ST_END(t0, st_a); }
            tc_fence_after();
            ZL_ST_BEGIN(t_epi);
            ZL_ST_BEGIN(t_scan);
            const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * (uint32_t)p.acc_cols;
            float best = 0.0f;
            int best_id = -1;
            if (cw) {
                // weighted scores: every class is scored (onnx_engine.cpp:787-796 with the class weight applied)
#pragma unroll 1
                for (int ch = 0; ch < nch; ++ch) {
                    uint32_t v[16];
                    tmem_ld16(taddr + 64u + (uint32_t)(ch << 4), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = (ch << 4) + i;
                        if (c < p.nc) {
                            const float s = __fmul_rn(cls_score<false>(__uint_as_float(v[i]) + bias_c[c]), __ldg(cw + c));
                            if (s > best) { best = s; best_id = c; }
                        }
                    }
                }
            } else if (nch == 5) {
                scan_classes_reg<5>(taddr + 64u, bias_c, in_range, conf_thr, best, best_id);
            } else if (nch == 1) {
                scan_classes_reg<1>(taddr + 64u, bias_c, in_range, conf_thr, best, best_id);
            } else {
                scan_classes_loop(taddr + 64u, bias_c, nch, in_range, conf_thr, best, best_id);
            }
            ZL_ST_END(t_scan, st_c);
            // this warp is done with the accumulator slot
            tc_fence_before();
            __syncwarp();
            if (leader) mbar_arrive(bar_tempty + 8u * acc);
            const bool keep = in_range && (best >= conf_thr) && (best_id >= 0);          // onnx_engine.cpp:799
            const unsigned kb = __ballot_sync(0xffffffffu, keep);
            { ZL_ST_BEGIN(t0); asm volatile("bar.sync %0, 64;" ::"r"(1u + q) : "memory"); ZL_ST_END(t0, st_d); }   // this tile's box distances are in dist_s[tl & 1]
            ZL_ST_BEGIN(t_emit);
            if (keep) {
                const float4 d = dist_s[((tl & 1) << 7) + row];
                const int f = pix / lv.hw, idx = pix - f * lv.hw;
                const int y = idx / lv.w, x = idx - y * lv.w;
                const float4 bx = dfl_box<false>(d.x, d.y, d.z, d.w, x, y, lv.stride);
                // one atomic per (warp, frame): the keepers of a frame elect the lowest lane
                const unsigned peers = __match_any_sync(kb, f);
                const int lead = __ffs(peers) - 1;
                uint32_t slot0 = 0;
                if ((int)lane == lead) slot0 = atomicAdd(p.cand_count + f, (uint32_t)__popc(peers));
                slot0 = __shfl_sync(peers, slot0, lead);
                const uint32_t slot = slot0 + (uint32_t)__popc(peers & ((1u << lane) - 1u));
                const int a = lv.a0 + idx;
                const float fw = (float)p.descs[f].w, fh = (float)p.descs[f].h;
                p.keys[(size_t)f * p.key_pitch + slot] = make_key(best_id, best, a);
                p.box_by_anchor[(size_t)f * p.A + a] = make_float4(__fdiv_rn(bx.x, fw), __fdiv_rn(bx.y, fh), __fdiv_rn(bx.z, fw), __fdiv_rn(bx.w, fh));
            }
            ZL_ST_END(t_emit, st_e);
            ZL_ST_END(t_epi, st_b);
        }
        if (STATS && warp == 2 && lane == 0) {
            atomicAdd(p.stats + 4, (unsigned long long)st_a);      // scan warp 2 waiting for a full accumulator
            atomicAdd(p.stats + 5, (unsigned long long)st_b);      // ... busy
            atomicAdd(p.stats + 12, (unsigned long long)st_c);     // ...... class scan
            atomicAdd(p.stats + 13, (unsigned long long)st_d);     // ...... waiting for the DFL warp's box distances
            atomicAdd(p.stats + 14, (unsigned long long)st_e);     // ...... candidate write
        }
    } else {
        // ===== DFL warps (one per TMEM lane quarter): the 64 box logits of every row -> four expected distances =====
        // Decoded for every anchor, not only for candidates: the work runs on warps that have nothing else to do, under the
        // HBM floor of a tile, and the scan warps never pay a conditional second TMEM pass.
        const uint32_t q = warp & 3u;
        const int row = (int)(q * 32u + lane);
        const bool leader = elect_one();
        float4* dist_s = reinterpret_cast<float4*>(smem_raw + (base + kHfDistOff - smem_u32(smem_raw)));
        int tl = 0;
#pragma unroll 1
        for (int g = g_first; g < g_end; g += g_step, ++tl) {
            const float* bias_b = bias_s + (p.all_levels ? level_of(p, g) : 0) * 256;
            const uint32_t acc = (uint32_t)tl & (uint32_t)(p.nacc - 1), aph = ((uint32_t)tl >> p.nacc_log2) & 1u;
            { ZL_ST_BEGIN(t0); mbar_wait(bar_tfull + 8u * acc, aph, 16); ZL_ST_END(t0, st_a); }
            tc_fence_after();
            ZL_ST_BEGIN(t_dfl);
            const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * (uint32_t)p.acc_cols;
            float d[4];
#pragma unroll 1
            for (int sd = 0; sd < 4; sd += 2) {                        // two sides per iteration: compact code (see scan_classes_reg)
                uint32_t ba[16], bb[16];
                tmem_ld16(taddr + (uint32_t)(sd << 4), ba);
                tmem_ld16(taddr + (uint32_t)((sd + 1) << 4), bb);
                tmem_ld_wait();
                if (sd == 2) {                                         // all 64 box columns have been read: release the slot
                    tc_fence_before();
                    __syncwarp();
                    if (leader) mbar_arrive(bar_tempty + 8u * acc);
                }
                float za[16], zb[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias_b + (sd << 4) + (q4 << 2));
                    const float4 c4 = *reinterpret_cast<const float4*>(bias_b + (sd << 4) + 16 + (q4 << 2));
                    za[(q4 << 2) + 0] = __uint_as_float(ba[(q4 << 2) + 0]) + b4.x; za[(q4 << 2) + 1] = __uint_as_float(ba[(q4 << 2) + 1]) + b4.y;
                    za[(q4 << 2) + 2] = __uint_as_float(ba[(q4 << 2) + 2]) + b4.z; za[(q4 << 2) + 3] = __uint_as_float(ba[(q4 << 2) + 3]) + b4.w;
                    zb[(q4 << 2) + 0] = __uint_as_float(bb[(q4 << 2) + 0]) + c4.x; zb[(q4 << 2) + 1] = __uint_as_float(bb[(q4 << 2) + 1]) + c4.y;
                    zb[(q4 << 2) + 2] = __uint_as_float(bb[(q4 << 2) + 2]) + c4.z; zb[(q4 << 2) + 3] = __uint_as_float(bb[(q4 << 2) + 3]) + c4.w;
                }
                const float da = dfl_expect<false>(za), db = dfl_expect<false>(zb);
                if (sd == 0) { d[0] = da; d[1] = db; } else { d[2] = da; d[3] = db; }
            }
            dist_s[((tl & 1) << 7) + row] = make_float4(d[0], d[1], d[2], d[3]);
            ZL_ST_END(t_dfl, st_b);
            asm volatile("bar.sync %0, 64;" ::"r"(1u + q) : "memory");
        }
        if (STATS && warp == 6 && lane == 0) atomicAdd(p.stats + 15, (unsigned long long)st_b);   // DFL warp 6 busy
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
    if (STATS && threadIdx.x == 0) {
        const long long t_end = clock64();
        atomicAdd(p.stats + 6, (unsigned long long)(t_end - t_begin));
        atomicAdd(p.stats + 7, (unsigned long long)(t_pro - t_begin));
        atomicMax(p.stats + 9, (unsigned long long)(t_end - t_begin));
        atomicAdd(p.stats + 10, 1ull);
        if (blockIdx.x == 0)
            p.stats[11] = (unsigned long long)p.kcc | ((unsigned long long)p.nchc << 8) | (1ull << 16) | (1ull << 20) | ((unsigned long long)p.stages << 24) |
                          ((unsigned long long)p.nc_pad << 32) | (((unsigned long long)p.total_tiles & 0x3fffull) << 42) | ((unsigned long long)p.nacc << 56);
    }
#undef ZL_ST_BEGIN
#undef ZL_ST_END
}

}  // namespace

struct HeadFusedImpl {
    HfMaps maps;
    HfParams p;
};
static_assert(sizeof(HeadFusedImpl) <= sizeof(((HeadFusedOp*)nullptr)->blob), "HeadFusedOp::blob too small");

bool head_fused_supported(const ConvWeights* const wb[3], const ConvWeights* const wc[3], const View xb[3], const View xc[3], int nc)
{
    static const bool enabled = [] { const char* e = getenv("ZL_FUSE_HEAD"); return !(e && e[0] == '0'); }();
    if (!enabled) return false;
    for (int l = 0; l < 3; ++l) {
        if (!wb[l] || !wc[l] || !wb[l]->w_tc || !wc[l]->w_tc) return false;
        if (wb[l]->k != 1 || wc[l]->k != 1 || wb[l]->act || wc[l]->act) return false;
        if (wb[l]->cin != 64 || wb[l]->cout != 64 || wc[l]->cout != nc || wc[l]->cin % 16 != 0) return false;
        if (wc[l]->cin != wc[0]->cin || wc[l]->cout_pad != wc[0]->cout_pad) return false;
        if (!xb[l].is16() || !xc[l].is16() || xb[l].dtype != xc[l].dtype) return false;
        if (xb[l].c != 64 || xc[l].c != wc[l]->cin || (xb[l].pitch % 8) || (xc[l].pitch % 8)) return false;
        if ((reinterpret_cast<uintptr_t>(xb[l].ptr) & 15) || (reinterpret_cast<uintptr_t>(xc[l].ptr) & 15)) return false;
        if (xb[l].pixels() != xc[l].pixels()) return false;
    }
    const int nc_pad = wc[0]->cout_pad;
    if (64 + nc_pad > 256) return false;                               // one TMEM slot holds both accumulators
    // shared memory: weights + at least two stages
    const int c3 = wc[0]->cin;
    const uint32_t w_alloc = (((uint32_t)64 * 64 * 2 + 1023u) & ~1023u) + (((uint32_t)nc_pad * c3 * 2 + 1023u) & ~1023u);
    const uint32_t stage = (uint32_t)kHfTile * (64 + c3) * 2;
    return kHfHeader + 1024u + w_alloc + 2u * stage <= 227u * 1024u;
}

int32_t head_fused_prepare(const ConvWeights* const wb[3], const ConvWeights* const wc[3], const View xb[3], const View xc[3],
                           const HeadLevel lvl[3], int nc, int A, int num_sms, HeadFusedOp* op)
{
    if (!head_fused_supported(wb, wc, xb, xc, nc)) ZL_FAIL(ZL_INVALID_ARGUMENT, "head_fused: configuration not supported");
    HeadFusedImpl& im = *reinterpret_cast<HeadFusedImpl*>(op->blob);
    im = HeadFusedImpl{};
    HfParams& p = im.p;
    const bool f16 = xb[0].dtype == DT_F16;
    const int c3 = wc[0]->cin, nc_pad = wc[0]->cout_pad;
    p.nc = nc; p.nc_pad = nc_pad; p.A = A; p.f16 = f16 ? 1 : 0;
    p.kcb = conv_kc(*wb[0], true); p.nchb = 64 / p.kcb;
    p.kcc = conv_kc(*wc[0], true); p.nchc = c3 / p.kcc;
    p.wb_bytes = 64u * 64u * 2u; p.wc_bytes = (uint32_t)nc_pad * c3 * 2u;
    p.wc_off = (p.wb_bytes + 1023u) & ~1023u;
    p.w_alloc = p.wc_off + ((p.wc_bytes + 1023u) & ~1023u);
    p.xb_bytes = (uint32_t)kHfTile * 64u * 2u; p.xc_bytes = (uint32_t)kHfTile * c3 * 2u;
    p.xc_off = p.xb_bytes;                                             // 16 KB: 1024-aligned
    p.stage_stride = (p.xb_bytes + p.xc_bytes + 1023u) & ~1023u;
    // all three levels' weights resident (tiles of all levels dealt round-robin: balanced whatever the levels cost) when
    // that leaves at least three stages; otherwise CTAs are dedicated to one level
    static const int force_levels = [] { const char* e = getenv("ZL_HEAD_ALL_LEVELS"); return e ? atoi(e) : -1; }();   // A/B: 0 per-level CTAs, 1 all levels
    const uint32_t budget = 227u * 1024u - kHfHeader - 1024u;
    p.all_levels = (3u * p.w_alloc + 3u * p.stage_stride <= budget) ? 1 : 0;
    if (force_levels == 0) p.all_levels = 0;
    const uint32_t w_total = (p.all_levels ? 3u : 1u) * p.w_alloc;
    int stages = (int)((budget - w_total) / p.stage_stride);
    if (stages > kHfMaxStages) stages = kHfMaxStages;
    p.stages = stages;
    op->smem_bytes = (int)(kHfHeader + 1024u + w_total + (uint32_t)stages * p.stage_stride);
    p.acc_cols = 64 + nc_pad <= 128 ? 128 : 256;
    p.nacc = 512 / p.acc_cols; p.nacc_log2 = p.nacc == 4 ? 2 : 1;
    p.tmem_cols = 512;
    // CTAs per level in proportion to the tile counts (every CTA keeps one level's weights)
    int tiles[3], total = 0;
    for (int l = 0; l < 3; ++l) { tiles[l] = ceil_div((int)xb[l].pixels(), kHfTile); total += tiles[l]; }
    int ctas[3];
    if (total <= num_sms) {
        for (int l = 0; l < 3; ++l) ctas[l] = tiles[l];
    } else {
        int used = 0;
        for (int l = 0; l < 3; ++l) { ctas[l] = std::max(1, (int)((long long)num_sms * tiles[l] / total)); if (ctas[l] > tiles[l]) ctas[l] = tiles[l]; used += ctas[l]; }
        while (used < num_sms) {                                       // leftovers to the level with the most tiles per CTA
            int bl = -1; double bw = 0;
            for (int l = 0; l < 3; ++l) { const double wgt = (double)tiles[l] / ctas[l]; if (ctas[l] < tiles[l] && wgt > bw) { bw = wgt; bl = l; } }
            if (bl < 0) break;
            ++ctas[bl]; ++used;
        }
        while (used > num_sms) {                                       // the max(1, .) floor can overshoot on tiny levels
            int bl = -1; double bw = 1e30;
            for (int l = 0; l < 3; ++l) { const double wgt = (double)tiles[l] / ctas[l]; if (ctas[l] > 1 && wgt < bw) { bw = wgt; bl = l; } }
            if (bl < 0) break;
            --ctas[bl]; --used;
        }
    }
    int cta0 = 0, g0 = 0;
    double bytes = 0, flops = 0;
    p.total_tiles = total;
    for (int l = 0; l < 3; ++l) {
        HfLevel& h = p.lv[l];
        h.g0 = g0; g0 += tiles[l];
        h.npix = (int32_t)xb[l].pixels(); h.hw = lvl[l].h * lvl[l].w; h.w = lvl[l].w; h.stride = lvl[l].stride; h.a0 = lvl[l].a0;
        h.ntiles = tiles[l]; h.cta0 = cta0; h.ctas = ctas[l];
        h.bias_box = wb[l]->bias; h.bias_cls = wc[l]->bias;
        cta0 += ctas[l];
        // input rows seen as {channel-in-chunk, pixel, chunk}: a box {kc, 128, K/kc} lands as [chunk][128][kc]
        ZL_TRY(make_tmap_w3d(&im.maps.m[l * 4 + 0], xb[l].ptr, p.kcb, h.npix, p.nchb, xb[l].pitch, p.kcb, kHfTile, f16));
        ZL_TRY(make_tmap_w3d(&im.maps.m[l * 4 + 1], xc[l].ptr, p.kcc, h.npix, p.nchc, xc[l].pitch, p.kcc, kHfTile, f16));
        ZL_TRY(make_tmap_w3d(&im.maps.m[l * 4 + 2], wb[l]->w_tc, p.kcb, 64, p.nchb, 64, p.kcb, 64, f16));
        ZL_TRY(make_tmap_w3d(&im.maps.m[l * 4 + 3], wc[l]->w_tc, p.kcc, nc_pad, p.nchc, c3, p.kcc, nc_pad, f16));
        bytes += (double)h.npix * (64 + c3) * 2;
        flops += 2.0 * h.npix * (64.0 * 64 + (double)nc * c3);
    }
    op->grid = p.all_levels ? std::min(total, num_sms) : cta0;
    op->bytes = bytes; op->flops = flops;
    return ZL_OK;
}

int32_t head_fused_launch(cudaStream_t st, const HeadFusedOp& op, const FrameDesc* descs, float conf_thr, const float* class_weights, const PostBuffers& pb,
                          unsigned long long* stats)
{
    static thread_local int last_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != last_dev) {
        ZL_CUDA(cudaFuncSetAttribute(head_decode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        ZL_CUDA(cudaFuncSetAttribute(head_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        last_dev = dev;
    }
    const HeadFusedImpl& im = *reinterpret_cast<const HeadFusedImpl*>(op.blob);
    HfParams p = im.p;
    p.descs = descs; p.conf_thr = conf_thr; p.class_weights = class_weights;
    p.keys = pb.keys; p.key_pitch = pb.key_pitch; p.box_by_anchor = pb.box_by_anchor; p.cand_count = pb.cand_count;
    static const bool use_pdl = [] { const char* e = getenv("ZL_DISABLE_PDL"); return !(e && e[0] == '1'); }();
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(op.grid); cfg.blockDim = dim3(kHfThreads); cfg.dynamicSmemBytes = (size_t)op.smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
    p.stats = stats;
    if (stats) ZL_CUDA(cudaLaunchKernelEx(&cfg, head_decode_kernel<true>, im.maps, p));
    else ZL_CUDA(cudaLaunchKernelEx(&cfg, head_decode_kernel<false>, im.maps, p));
    return ZL_OK;
}

}  // namespace zl
