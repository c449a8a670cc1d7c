// head_fused.cu — the last two 1x1 convolutions of the Detect head (model.22.cv2.l.2: 64 -> 64 DFL logits, model.22.cv3.l.2:
// c3 -> nc class logits), the Detect tail (D1) and the decode / threshold loop (F1) as ONE persistent tcgen05 kernel for
// all three levels (sm_100a only).
//
// Reference nodes: the two Conv nodes and the DFL / sigmoid / dist2bbox tail inside Ort::Session::Run
// (src/inference/onnx_engine.cpp:577-585) followed by the decode loop of postProcess (onnx_engine.cpp:773-819).
// In the unfused chain the two convs write their fp32 logits ([pixels][64] and [pixels][nc]: 2.3 KB per anchor row of
// 128) and decode_filter_kernel reads them back; here the logits never leave the SM:
//   * a tile = 128 consecutive pixels of one level's flattened [n*h*w] pixel list.  Two TMA loads bring the tile's rows
//     of the two branch inputs (HB2_l / HC2_l, 16-bit NHWC) into shared memory in the swizzled K-major UMMA layout —
//     a 3-D box {kc, 128, K/kc} lands as [chunk][128][kc], one instruction per branch;
//   * the MMA warp accumulates [128 x 64] and [128 x nc_pad] into one TMEM slot (columns 0..63 and 64..), with the k-steps
//     in ascending order == the accumulation order of the standalone conv kernels, so the logits are the same bits;
//   * the epilogue thread of pixel row r reads ITS row of class logits from TMEM (tcgen05.ld 32x32b), adds the bias,
//     finds the best class exactly like decode_filter_kernel, and only warps that hold a candidate read the 64 DFL columns;
//     candidates are appended to the frame's key list with the same key / box arithmetic (head_math.cuh).
//   * every CTA serves ONE level (CTAs are dealt to the levels in proportion to their tile counts), so it keeps only that
//     level's weights resident.
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

constexpr int kHfEpiWarps = 8;               // two sets of four (one warp per TMEM lane quarter) on alternate tiles
constexpr int kHfProducers = 2;              // stages dealt round-robin (one thread completes a stage every ~500 + 130/TMA cycles)
constexpr int kHfThreads = 64 + 32 * kHfEpiWarps + 32 * (kHfProducers - 1);
constexpr int kHfMaxStages = 8;
constexpr int kHfTile = 128;

struct HfLevel {
    int32_t npix, hw, w, stride, a0, ntiles, cta0, ctas;
    const float* bias_box;
    const float* bias_cls;
};
struct HfParams {
    HfLevel lv[3];
    int32_t nc, nc_pad, A, key_pitch, f16;
    int32_t kcb, nchb, kcc, nchc;                        // K chunking of the box (K = 64) and class (K = c3) branches
    int32_t stages, nacc, nacc_log2, acc_cols;
    uint32_t tmem_cols, wb_bytes, wc_bytes, wc_off, w_alloc, xb_bytes, xc_bytes, xc_off, stage_stride;
    const FrameDesc* descs;
    float conf_thr;
    const float* class_weights;
    uint64_t* keys;
    float4* box_by_anchor;
    uint32_t* cand_count;
};
struct HfMaps { CUtensorMap m[12]; };        // [level * 4 + {0: box input, 1: class input, 2: box weights, 3: class weights}]

__global__ void __launch_bounds__(kHfThreads, 1)
head_decode_kernel(const __grid_constant__ HfMaps maps, const HfParams p)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_pfull = base, bar_pempty = base + 8u * kHfMaxStages, bar_wfull = base + 16u * kHfMaxStages;
    const uint32_t bar_tfull = bar_wfull + 8u, bar_tempty = bar_tfull + 32u, tmem_slot = bar_tempty + 32u;
    const uint32_t bias_off = 512u;                        // fp32 bias: [64 box | nc_pad class] (<= 256 floats) up to +1536
    const uint32_t wbase = base + 2048u;
    const uint32_t pbase = wbase + p.w_alloc;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* bias_s = reinterpret_cast<float*>(smem_raw + (base + bias_off - smem_u32(smem_raw)));

    const int bid = (int)blockIdx.x;
    const int l = bid >= p.lv[2].cta0 ? 2 : (bid >= p.lv[1].cta0 ? 1 : 0);
    const HfLevel& lv = p.lv[l];
    const int tile0 = bid - lv.cta0, tile_step = lv.ctas;
    const CUtensorMap* map_xb = &maps.m[l * 4 + 0];
    const CUtensorMap* map_xc = &maps.m[l * 4 + 1];
    const CUtensorMap* map_wb = &maps.m[l * 4 + 2];
    const CUtensorMap* map_wc = &maps.m[l * 4 + 3];

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(bar_pfull + 8u * s, 1u); mbar_init(bar_pempty + 8u * s, 1u); }
        mbar_init(bar_wfull, 1u);
        for (int a = 0; a < p.nacc; ++a) { mbar_init(bar_tfull + 8u * a, 1u); mbar_init(bar_tempty + 8u * a, 4u); }
        fence_barrier_init();
        tma_prefetch_desc(map_xb); tma_prefetch_desc(map_xc); tma_prefetch_desc(map_wb); tma_prefetch_desc(map_wc);
    }
    if (warp == 1) { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
    for (int i = threadIdx.x; i < 64 + p.nc_pad; i += (int)blockDim.x) bias_s[i] = i < 64 ? lv.bias_box[i] : lv.bias_cls[i - 64];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0 || warp >= 2u + kHfEpiWarps) {
        // ===== TMA producers =====
        const uint32_t pi = warp == 0 ? 0u : warp - (1u + kHfEpiWarps);
        if (elect_one()) {
            if (pi == 0) {                                 // this level's weights, once (constants: before the dependency wait)
                mbar_arrive_expect_tx(bar_wfull, p.wb_bytes + p.wc_bytes);
                tma_load_3d(map_wb, bar_wfull, wbase, 0, 0, 0);
                tma_load_3d(map_wc, bar_wfull, wbase + p.wc_off, 0, 0, 0);
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            const uint32_t np = (uint32_t)p.stages < (uint32_t)kHfProducers ? (uint32_t)p.stages : (uint32_t)kHfProducers;
            uint32_t s = 0, ph = 0, turn = 0;
            for (int tile = tile0; tile < lv.ntiles; tile += tile_step) {
                if (turn == pi) {
                    mbar_wait(bar_pempty + 8u * s, ph ^ 1u, 11);
                    mbar_arrive_expect_tx(bar_pfull + 8u * s, p.xb_bytes + p.xc_bytes);
                    const uint32_t stage = pbase + s * p.stage_stride;
                    tma_load_3d(map_xb, bar_pfull + 8u * s, stage, 0, tile * kHfTile, 0);
                    tma_load_3d(map_xc, bar_pfull + 8u * s, stage + p.xc_off, 0, tile * kHfTile, 0);
                }
                if (++turn == np) turn = 0;
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
            }
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
        const uint32_t stage16 = p.stage_stride >> 4;
        const int ksb = p.kcb / 16, ksc = p.kcc / 16;
        mbar_wait(bar_wfull, 0u, 12);
        tc_fence_after();
        uint32_t s = 0, ph = 0, tl = 0;
        for (int tile = tile0; tile < lv.ntiles; tile += tile_step, ++tl) {
            const uint32_t acc = tl & (uint32_t)(p.nacc - 1), aph = (tl >> p.nacc_log2) & 1u;
            mbar_wait(bar_tempty + 8u * acc, aph ^ 1u, 13);
            mbar_wait(bar_pfull + 8u * s, ph, 14);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t tmem_d = tmem_base + acc * (uint32_t)p.acc_cols;
                for (int g = 0; g < p.nchb; ++g)
                    for (int k = 0; k < ksb; ++k)
                        umma_bf16(tmem_d, ab0 + (uint64_t)(s * stage16 + (uint32_t)g * xb_chunk16 + 2u * (uint32_t)k),
                                  bb0 + (uint64_t)((uint32_t)g * wb_chunk16 + 2u * (uint32_t)k), idesc_b, (g | k) ? 1u : 0u);
                for (int g = 0; g < p.nchc; ++g)
                    for (int k = 0; k < ksc; ++k)
                        umma_bf16(tmem_d + 64u, ac0 + (uint64_t)(s * stage16 + (uint32_t)g * xc_chunk16 + 2u * (uint32_t)k),
                                  bc0 + (uint64_t)((uint32_t)g * wc_chunk16 + 2u * (uint32_t)k), idesc_c, (g | k) ? 1u : 0u);
                umma_commit(bar_pempty + 8u * s);
                umma_commit(bar_tfull + 8u * acc);
            }
            __syncwarp();
            if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1u; }
        }
    } else {
        // ===== epilogue: logits (TMEM) -> best class -> threshold -> DFL box -> candidate =====
        const uint32_t q = warp & 3u;                                 // TMEM lane quarter this warp may read
        const int set = (int)(warp - 2u) >> 2;                        // tile residue this warp serves
        const int row = (int)(q * 32u + lane);
        const bool leader = elect_one();
        const int nch = p.nc_pad >> 4;
        const float* bias_c = bias_s + 64;
        const float* cw = p.class_weights;
        const float conf_thr = p.conf_thr;
        asm volatile("griddepcontrol.wait;" ::: "memory");            // cand_count / keys belong to earlier work of the stream
        for (int tile = tile0 + set * tile_step, tl = set; tile < lv.ntiles; tile += 2 * tile_step, tl += 2) {
            const uint32_t acc = (uint32_t)tl & (uint32_t)(p.nacc - 1), aph = ((uint32_t)tl >> p.nacc_log2) & 1u;
            const int pix = tile * kHfTile + row;
            const bool in_range = pix < lv.npix;
            mbar_wait(bar_tfull + 8u * acc, aph, 15);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * (uint32_t)p.acc_cols;
            float best = 0.0f;
            int best_id = -1;
            if (cw) {
                // weighted scores: every class is scored (onnx_engine.cpp:787-796 with the class weight applied)
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
            } else {
                // pass 1: the largest logit.  The sigmoid is monotone, so only classes whose LOGIT is near the largest one
                // can hold the largest score (same rule, same constants as decode_filter_kernel)
                float zmax = -FLT_MAX;
                for (int ch = 0; ch < nch; ch += 2) {
                    uint32_t v[2][16];
                    const bool two = ch + 1 < nch;
                    tmem_ld16(taddr + 64u + (uint32_t)(ch << 4), v[0]);
                    if (two) tmem_ld16(taddr + 64u + (uint32_t)((ch + 1) << 4), v[1]);
                    tmem_ld_wait();
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (h == 1 && !two) break;
                        const int c0 = (ch + h) << 4;
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < p.nc) zmax = fmaxf(zmax, __uint_as_float(v[h][i]) + bias_c[c0 + i]);
                    }
                }
                // no class of this warp's 32 anchors can reach the threshold: skip the scoring pass (the fast sigmoid is
                // monotone to within a few ulp; 1e-5 relative is a wide margin on the safe side)
                const bool possible = in_range && cls_score<false>(zmax) >= conf_thr * (1.0f - 1e-5f);
                if (__any_sync(0xffffffffu, possible)) {
                    const float zcut = fminf(zmax, 8.0f) - 0.05f;
                    for (int ch = 0; ch < nch; ch += 2) {
                        uint32_t v[2][16];
                        const bool two = ch + 1 < nch;
                        tmem_ld16(taddr + 64u + (uint32_t)(ch << 4), v[0]);
                        if (two) tmem_ld16(taddr + 64u + (uint32_t)((ch + 1) << 4), v[1]);
                        tmem_ld_wait();
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (h == 1 && !two) break;
                            const int c0 = (ch + h) << 4;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int c = c0 + i;
                                const float z = __uint_as_float(v[h][i]) + bias_c[c];
                                if (c < p.nc && z >= zcut) {
                                    const float s = cls_score<false>(z);
                                    if (s > best) { best = s; best_id = c; }      // strict '>': first maximum wins (onnx_engine.cpp:792)
                                }
                            }
                        }
                    }
                }
            }
            const bool keep = in_range && (best >= conf_thr) && (best_id >= 0);          // onnx_engine.cpp:799
            const unsigned kb = __ballot_sync(0xffffffffu, keep);
            uint32_t b[4][16];
            if (kb != 0u) {
#pragma unroll
                for (int sd = 0; sd < 4; ++sd) tmem_ld16(taddr + (uint32_t)(sd << 4), b[sd]);
                tmem_ld_wait();
            }
            // this warp is done with the accumulator slot
            tc_fence_before();
            __syncwarp();
            if (leader) mbar_arrive(bar_tempty + 8u * acc);
            if (keep) {
                float d[4];
#pragma unroll
                for (int sd = 0; sd < 4; ++sd) {
                    float z[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) z[i] = __uint_as_float(b[sd][i]) + bias_s[(sd << 4) + i];
                    d[sd] = dfl_expect<false>(z);
                }
                const int f = pix / lv.hw, idx = pix - f * lv.hw;
                const int y = idx / lv.w, x = idx - y * lv.w;
                const float4 bx = dfl_box<false>(d[0], d[1], d[2], d[3], x, y, lv.stride);
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
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
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
    return 3072u + w_alloc + 2u * stage <= 227u * 1024u;
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
    int stages = (int)((227u * 1024u - 3072u - p.w_alloc) / p.stage_stride);
    if (stages > kHfMaxStages) stages = kHfMaxStages;
    p.stages = stages;
    op->smem_bytes = (int)(3072u + p.w_alloc + (uint32_t)stages * p.stage_stride);
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
    int cta0 = 0;
    double bytes = 0, flops = 0;
    for (int l = 0; l < 3; ++l) {
        HfLevel& h = p.lv[l];
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
    op->grid = cta0;
    op->bytes = bytes; op->flops = flops;
    return ZL_OK;
}

int32_t head_fused_launch(cudaStream_t st, const HeadFusedOp& op, const FrameDesc* descs, float conf_thr, const float* class_weights, const PostBuffers& pb)
{
    static thread_local int last_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != last_dev) {
        ZL_CUDA(cudaFuncSetAttribute(head_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
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
    ZL_CUDA(cudaLaunchKernelEx(&cfg, head_decode_kernel, im.maps, p));
    return ZL_OK;
}

}  // namespace zl
