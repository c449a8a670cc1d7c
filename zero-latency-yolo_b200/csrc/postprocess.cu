// postprocess.cu — D1 (Detect tail), F1 (decode/threshold), N1 (class-aware NMS).
//
// D1 is the tail of the ONNX graph (DFL softmax expectation, dist2bbox, x stride,
// sigmoid) and produces the reference's `output0`
// [n, 4+nc, A] (src/inference/onnx_engine.cpp:50,767-774).
// F1 restates the decode loop of postProcess (onnx_engine.cpp:773-819):
//   per-anchor argmax with strict '>' from 0.0f, keep on '>= threshold',
//   box / REQUEST-frame width,height.
// N1 restates applyNMS + calculateIoU (onnx_engine.cpp:837-909): sort by
//   (class asc, confidence desc) — completed to a total order with the anchor
//   index ascending, because the reference's std::sort is unstable — then a
//   greedy per-class sweep that suppresses on IoU > threshold.
// All fp32 arithmetic that decides an outcome uses the *_rn intrinsics so no
// FMA contraction can make a comparison differ from the IEEE CPU oracle.
#include <cfloat>
#include <type_traits>

#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "head_math.cuh"
#include "kernels.h"

namespace zl {
namespace {

// ============================================================= D1: DFL decode
constexpr int kDflAnchors = 32;    // anchors per CTA
constexpr int kDflThreads = 256;   // 8 warps, 4 anchors each

// Detect tail, four lanes per anchor: lane `side` (0..3 = left, top, right, bottom) reads its 16
// DFL bins with four float4 loads and does the softmax expectation in-thread, in a fixed order shared by the raw-head
// kernel and the fused hot-path kernel (so both give the same bits); two xor-shuffles gather the four sides.
template <bool PRECISE>
__device__ __forceinline__ float dfl_side(const float* __restrict__ bins)
{
    float z[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(bins) + q);
        z[4 * q] = v.x; z[4 * q + 1] = v.y; z[4 * q + 2] = v.z; z[4 * q + 3] = v.w;
    }
    return dfl_expect<PRECISE>(z);
}

// All four lanes of the anchor's quad call this with the same (lv, pix, x, y); every lane returns the box.
template <bool PRECISE>
__device__ __forceinline__ float4 dfl_box4(const HeadLevel& lv, size_t pix, int x, int y, int side, unsigned quad_mask)
{
    const float d = dfl_side<PRECISE>(lv.box + pix * 64 + side * 16);
    const int lane = threadIdx.x & 31, q0 = lane & ~3;
    const float dl = __shfl_sync(quad_mask, d, q0 + 0), dt = __shfl_sync(quad_mask, d, q0 + 1);
    const float dr = __shfl_sync(quad_mask, d, q0 + 2), db = __shfl_sync(quad_mask, d, q0 + 3);
    return dfl_box<PRECISE>(dl, dt, dr, db, x, y, lv.stride);
}

// PRECISE = fp64 softmax / sigmoid / box arithmetic (exact mode); otherwise fp32 (bf16 mode).
template <bool PRECISE>
__global__ void __launch_bounds__(kDflThreads)
dfl_decode_kernel(const HeadLevel l0, const HeadLevel l1, const HeadLevel l2, int nc, int A, float* __restrict__ raw)
{
    extern __shared__ float stage[];                 // [(4+nc)][33]
    const int f = blockIdx.y;
    const int a0 = blockIdx.x * kDflAnchors;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows = 4 + nc;

    // boxes: this CTA's 32 anchors = 128 (anchor, side) pairs, one per thread of warps 0..3
    if (threadIdx.x < 4 * kDflAnchors) {
        const int la = threadIdx.x >> 2, side = threadIdx.x & 3;
        const int a = a0 + la;
        const unsigned quad_mask = 0xfu << (lane & ~3);
        if (a < A) {                                     // uniform per quad
            const HeadLevel& lv = (a >= l2.a0) ? l2 : (a >= l1.a0 ? l1 : l0);
            const int idx = a - lv.a0;
            const int y = idx / lv.w, x = idx - y * lv.w;
            const size_t pix = (size_t)(f * lv.h + y) * lv.w + x;
            const float4 bx = dfl_box4<PRECISE>(lv, pix, x, y, side, quad_mask);
            if (side == 0) {
                stage[0 * 33 + la] = bx.x; stage[1 * 33 + la] = bx.y; stage[2 * 33 + la] = bx.z; stage[3 * 33 + la] = bx.w;
            }
        }
    }
    // classes: sigmoid, transposed through smem (warp w handles anchors 4w..4w+3)
    for (int i = 0; i < kDflAnchors / 8; ++i) {
        const int la = warp * (kDflAnchors / 8) + i;
        const int a = a0 + la;
        if (a >= A) break;                           // warp-uniform
        const HeadLevel& lv = (a >= l2.a0) ? l2 : (a >= l1.a0 ? l1 : l0);
        const int idx = a - lv.a0;
        const int y = idx / lv.w, x = idx - y * lv.w;
        const size_t pix = (size_t)(f * lv.h + y) * lv.w + x;
        const float* cp = lv.cls + pix * lv.cls_pitch;
        for (int c = lane; c < nc; c += 32) {
            const float z = __ldg(cp + c);
            stage[(4 + c) * 33 + la] = cls_score<PRECISE>(z);
        }
    }
    __syncthreads();
    float* out = raw + (size_t)f * rows * A;
    for (int i = threadIdx.x; i < rows * kDflAnchors; i += kDflThreads) {
        const int r = i / kDflAnchors, la = i - r * kDflAnchors;
        if (a0 + la < A) out[(size_t)r * A + a0 + la] = stage[r * 33 + la];
    }
}

// ============================================================= F1: filter
// candidate sort key: head_math.cuh (make_key / key_class / key_conf / key_anchor)
__global__ void __launch_bounds__(256)
filter_kernel(const float* __restrict__ raw, int nc, int A, const FrameDesc* __restrict__ descs,
              const int32_t* __restrict__ img_wh, float conf_thr, const float* __restrict__ class_weights,
              uint64_t* __restrict__ keys, int key_pitch, float4* __restrict__ box_by_anchor, uint32_t* __restrict__ cand_count)
{
    const int f = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const float* r = raw + (size_t)f * (4 + nc) * A;
    bool keep = false;
    float max_conf = 0.0f;
    int max_id = -1;
    if (a < A) {
        for (int j = 0; j < nc; ++j) {
            float s = __ldg(r + (size_t)(4 + j) * A + a);
            if (class_weights) s = __fmul_rn(s, __ldg(class_weights + j));
            if (s > max_conf) { max_conf = s; max_id = j; }
        }
        keep = (max_conf >= conf_thr) && (max_id >= 0);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    if (ballot == 0u) return;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(cand_count + f, (uint32_t)__popc(ballot));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) {
        int iw, ih;
        if (img_wh) { iw = img_wh[2 * f]; ih = img_wh[2 * f + 1]; } else { iw = descs[f].w; ih = descs[f].h; }
        const float fw = (float)iw, fh = (float)ih;
        const float cx = __ldg(r + a), cy = __ldg(r + (size_t)A + a), w = __ldg(r + 2 * (size_t)A + a), h = __ldg(r + 3 * (size_t)A + a);
        const uint32_t slot = base + (uint32_t)__popc(ballot & ((1u << lane) - 1u));
        keys[(size_t)f * key_pitch + slot] = make_key(max_id, max_conf, a);
        box_by_anchor[(size_t)f * A + a] = make_float4(__fdiv_rn(cx, fw), __fdiv_rn(cy, fh), __fdiv_rn(w, fw), __fdiv_rn(h, fh));
    }
}

// D1 + F1 fused (engine hot path): the raw head tensor [n,4+nc,A] is never materialised.  One warp per anchor:
// lanes score the classes (same sigmoid as dfl_decode_kernel), a shuffle reduction finds (max score, lowest class
// index) == the reference's strict-'>' scan (onnx_engine.cpp:787-796), and ONLY anchors that pass the threshold
// (:799) pay for the DFL box decode.  Emits the same keys / boxes as filter_kernel, bit for bit.
template <bool PRECISE, bool STAGED>
__global__ void __launch_bounds__(128)
decode_filter_kernel(const HeadLevel l0, const HeadLevel l1, const HeadLevel l2, int nc, int A, const FrameDesc* __restrict__ descs,
                     float conf_thr, const float* __restrict__ class_weights, uint64_t* __restrict__ keys, int key_pitch,
                     float4* __restrict__ box_by_anchor, uint32_t* __restrict__ cand_count)
{
    extern __shared__ float df_sm[];                                // STAGED: [4 warps][32 anchors][nc + 1]
    const int f = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;            // one thread scans the classes of one anchor
    float best = 0.0f;
    int best_id = -1;
    int lx = 0, ly = 0, lvl = 0;
    size_t pix = 0;
    if (a < A) {
        lvl = (a >= l2.a0) ? 2 : (a >= l1.a0 ? 1 : 0);
        const HeadLevel& lv = lvl == 2 ? l2 : (lvl == 1 ? l1 : l0);
        const int idx = a - lv.a0;
        ly = idx / lv.w; lx = idx - ly * lv.w;
        pix = (size_t)(f * lv.h + ly) * lv.w + lx;
    }
    const float* cp = nullptr;
    if (a < A) {
        const HeadLevel& lv = lvl == 2 ? l2 : (lvl == 1 ? l1 : l0);
        cp = lv.cls + pix * lv.cls_pitch;                            // cls_pitch is a multiple of 4 floats: 16-byte aligned rows
    }
    if (STAGED) {
        // the warp copies its 32 anchors' class rows into shared memory (row stride nc+1: conflict-free per-lane scans)
        // so every logit crosses L1 exactly once with fully coalesced reads
        float* wsm = df_sm + (size_t)(threadIdx.x >> 5) * 32 * (nc + 1);
        const int a_last = a - lane + 31;
        const int lvl_first = __shfl_sync(0xffffffffu, lvl, 0), lvl_last = __shfl_sync(0xffffffffu, lvl, 31);
        if (a_last < A && lvl_first == lvl_last) {
            // common case: 32 anchors of one level = one contiguous run of 32 x cls_pitch floats.  Flat float4 copy,
            // four independent loads in flight per lane before any store
            const HeadLevel& lv = lvl_first == 2 ? l2 : (lvl_first == 1 ? l1 : l0);
            const int pitch = lv.cls_pitch;
            const float4* src = reinterpret_cast<const float4*>(__shfl_sync(0xffffffffu, (unsigned long long)cp, 0));
            const int n4 = 8 * pitch;                                  // 32 * pitch / 4
            for (int i0 = lane; i0 < n4; i0 += 128) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int i = i0 + 32 * u; v[u] = i < n4 ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + 32 * u;
                    if (i < n4) {
                        const int e = 4 * i, j = e / pitch, c = e - j * pitch;
                        float* d = wsm + j * (nc + 1) + c;
                        if (c + 0 < nc) d[0] = v[u].x;
                        if (c + 1 < nc) d[1] = v[u].y;
                        if (c + 2 < nc) d[2] = v[u].z;
                        if (c + 3 < nc) d[3] = v[u].w;
                    }
                }
            }
        } else {
            for (int j = 0; j < 32; ++j) {                             // level boundary / tail: row by row
                const float* rp = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)cp, j));
                if (rp != nullptr)
                    for (int c = lane; c < nc; c += 32) wsm[j * (nc + 1) + c] = __ldg(rp + c);
            }
        }
        __syncwarp();
        cp = wsm + lane * (nc + 1);
    }
    if (a < A) {
        if (class_weights) {
            for (int c = 0; c < nc; ++c) {
                const float s = __fmul_rn(cls_score<PRECISE>(cp[c]), __ldg(class_weights + c));
                if (s > best) { best = s; best_id = c; }             // strict '>': first maximum wins (onnx_engine.cpp:792)
            }
        } else {
            // the sigmoid is monotone, so only classes whose LOGIT is near the largest one can hold the largest score.
            // "Near" = within 0.05 of min(zmax, 8): below 8 a logit gap of 0.05 moves the score by >= 1.7e-5, far above
            // the error of the fast exp; above 8 scores start to collide in fp32 (ties go to the lowest class index),
            // so every saturated class is scored.  Identical to scoring all nc classes, ~1 sigmoid per anchor.
            float zmax = -FLT_MAX;
            for (int c = 0; c < nc; ++c) zmax = fmaxf(zmax, cp[c]);
            const float zcut = fminf(zmax, 8.0f) - 0.05f;
            for (int c = 0; c < nc; ++c) {
                const float z = cp[c];
                if (z >= zcut) {
                    const float s = cls_score<PRECISE>(z);
                    if (s > best) { best = s; best_id = c; }
                }
            }
        }
    }
    const bool keep = (best >= conf_thr) && (best_id >= 0);         // onnx_engine.cpp:799
    unsigned todo = __ballot_sync(0xffffffffu, keep);
    if (todo == 0u) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(cand_count + f, (uint32_t)__popc(todo));
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned all = todo;
    // the anchors that pass pay for the DFL box decode: eight at a time, four lanes (one per box side) each
    const int quad = lane >> 2, side = lane & 3;
    const unsigned quad_mask = 0xfu << (lane & ~3);
    while (todo != 0u) {
        // the quad-th set bit of `todo` is this quad's anchor (if any)
        unsigned t = todo;
        for (int i = 0; i < quad; ++i) t &= t - 1u;
        const bool have = t != 0u;
        const int src = have ? __ffs(t) - 1 : 0;
        const int k_lvl = __shfl_sync(0xffffffffu, lvl, src);
        const int kx = __shfl_sync(0xffffffffu, lx, src), ky = __shfl_sync(0xffffffffu, ly, src);
        const unsigned long long kpix = __shfl_sync(0xffffffffu, (unsigned long long)pix, src);
        const float kbest = __shfl_sync(0xffffffffu, best, src);
        const int kid = __shfl_sync(0xffffffffu, best_id, src);
        if (have) {
            const HeadLevel& lv = k_lvl == 2 ? l2 : (k_lvl == 1 ? l1 : l0);
            const float4 bx = dfl_box4<PRECISE>(lv, (size_t)kpix, kx, ky, side, quad_mask);
            if (side == 0) {
                const int ka = a - lane + src;
                const float fw = (float)descs[f].w, fh = (float)descs[f].h;
                const uint32_t slot = base + (uint32_t)__popc(all & ((1u << src) - 1u));
                keys[(size_t)f * key_pitch + slot] = make_key(kid, kbest, ka);
                box_by_anchor[(size_t)f * A + ka] = make_float4(__fdiv_rn(bx.x, fw), __fdiv_rn(bx.y, fh), __fdiv_rn(bx.z, fw), __fdiv_rn(bx.w, fh));
            }
        }
        for (int i = 0; i < 8 && todo != 0u; ++i) todo &= todo - 1u;     // eight anchors done
    }
}

// ============================================================= N1: NMS
constexpr int kNmsThreads = 1024;
constexpr int kMaxLargeSeg = 512;
#ifndef ZL_NMS_W
#define ZL_NMS_W 4
#endif
constexpr int kSweepW = ZL_NMS_W;     // IoU tests in flight per candidate in the block sweeps
constexpr int kNmsSplitMin = 256;     // frames with fewer candidates are not split over a cluster's CTAs
constexpr int kWholeCtaSeg = 512;     // class segments larger than this are swept by all 1024 threads, one segment at a time     // queue slots for class segments with more than 32 candidates

// calculateIoU (onnx_engine.cpp:881-909) with IEEE single ops in the reference's order.
__device__ __forceinline__ float iou_ref(const float4 a, const float4 b) {
    const float ahw = __fmul_rn(a.z, 0.5f), ahh = __fmul_rn(a.w, 0.5f);
    const float bhw = __fmul_rn(b.z, 0.5f), bhh = __fmul_rn(b.w, 0.5f);
    const float x1_min = __fsub_rn(a.x, ahw), y1_min = __fsub_rn(a.y, ahh);
    const float x1_max = __fadd_rn(a.x, ahw), y1_max = __fadd_rn(a.y, ahh);
    const float x2_min = __fsub_rn(b.x, bhw), y2_min = __fsub_rn(b.y, bhh);
    const float x2_max = __fadd_rn(b.x, bhw), y2_max = __fadd_rn(b.y, bhh);
    const float xo = fmaxf(0.0f, __fsub_rn(fminf(x1_max, x2_max), fmaxf(x1_min, x2_min)));
    const float yo = fmaxf(0.0f, __fsub_rn(fminf(y1_max, y2_max), fmaxf(y1_min, y2_min)));
    const float inter = __fmul_rn(xo, yo);
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(a.z, a.w), __fmul_rn(b.z, b.w)), inter);
    return uni > 0.0f ? __fdiv_rn(inter, uni) : 0.0f;
}

// Phase timestamps of the slowest-looking CTA (debug aid, ZL_NMS_DEBUG=1 prints them): written by thread 0 of the CTA
// whose frame index is g_nms_dbg_frame.
__device__ long long g_nms_dbg[8][10];      // [cluster rank][0..5 phase stamps, 6 entry, 7 candidates of the rank, 8/9 globaltimer at entry / exit]
__device__ int g_nms_dbg_frame = -1;
__device__ long long g_nms_dbg2[64];        // fine stamps inside the large-segment phase (group 0 of rank 0, in program order)
__device__ __forceinline__ long long nms_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#ifndef ZL_NMS_STAMPS
#define ZL_NMS_STAMPS 1
#endif
#if ZL_NMS_STAMPS
#define ZL_NMS_STAMP(k) do { if (tid == 0 && f == g_nms_dbg_frame) g_nms_dbg[rank][k] = clock64(); } while (0)
#else
#define ZL_NMS_STAMP(k) do { } while (0)
#endif

// `calculateIoU(a, b) > thr` (onnx_engine.cpp:871,881-909), decided EXACTLY but usually without the IEEE division: boxes that
// do not overlap have IoU 0 (never > thr for thr >= 0); otherwise a reciprocal estimate of inter/union settles every case
// that is not within 1e-5 of the threshold, and only the knife edges pay for __fdiv_rn.  Same operation order as iou_ref.
__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {
    const float ahw = __fmul_rn(a.z, 0.5f), ahh = __fmul_rn(a.w, 0.5f);
    const float bhw = __fmul_rn(b.z, 0.5f), bhh = __fmul_rn(b.w, 0.5f);
    const float xo = fmaxf(0.0f, __fsub_rn(fminf(__fadd_rn(a.x, ahw), __fadd_rn(b.x, bhw)), fmaxf(__fsub_rn(a.x, ahw), __fsub_rn(b.x, bhw))));
    const float yo = fmaxf(0.0f, __fsub_rn(fminf(__fadd_rn(a.y, ahh), __fadd_rn(b.y, bhh)), fmaxf(__fsub_rn(a.y, ahh), __fsub_rn(b.y, bhh))));
    const float inter = __fmul_rn(xo, yo);
    if (!(inter > 0.0f) && thr >= 0.0f) return false;            // IoU is 0 (or the union guard returns 0): not > thr
    const float uni = __fsub_rn(__fadd_rn(__fmul_rn(a.z, a.w), __fmul_rn(b.z, b.w)), inter);
    if (!(uni > 0.0f)) return 0.0f > thr;
    const float q = inter * __frcp_rn(uni);                      // within a few ulp of the correctly rounded quotient
    if (q > thr + 1e-5f) return true;
    if (q < thr - 1e-5f) return false;
    return __fdiv_rn(inter, uni) > thr;
}

// The same decision for SEVERAL independent pairs at once.  One test is a chain of ~40 dependent instructions (~350
// cycles of latency on the round-2 stamps: a sweep against 13 kept boxes cost 4500 cycles), so the bulk sweeps call the
// branch-free form below four or eight times in a row — the compiler interleaves the chains — and only the pairs it could
// not settle (quotient within 1e-5 of the threshold, degenerate unions) go through the exact, out-of-line form.
struct BoxC { float x0, x1, y0, y1, area; };
__device__ __forceinline__ BoxC box_corners(const float4 b) {
    const float hw = __fmul_rn(b.z, 0.5f), hh = __fmul_rn(b.w, 0.5f);
    BoxC c;
    c.x0 = __fsub_rn(b.x, hw); c.x1 = __fadd_rn(b.x, hw); c.y0 = __fsub_rn(b.y, hh); c.y1 = __fadd_rn(b.y, hh);
    c.area = __fmul_rn(b.z, b.w);
    return c;
}
__device__ __forceinline__ bool iou_gt_fast(const BoxC& a, const BoxC& b, float thr, bool& undecided) {
    const float xo = fmaxf(0.0f, __fsub_rn(fminf(a.x1, b.x1), fmaxf(a.x0, b.x0)));
    const float yo = fmaxf(0.0f, __fsub_rn(fminf(a.y1, b.y1), fmaxf(a.y0, b.y0)));
    const float inter = __fmul_rn(xo, yo);
    const float uni = __fsub_rn(__fadd_rn(a.area, b.area), inter);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(uni));      // 1 ulp: far inside the 1e-5 guard band
    const float q = inter * r;
    const bool zero = !(inter > 0.0f) && thr >= 0.0f;            // IoU is 0: not > thr
    const bool degenerate = !(uni > 1e-30f);                     // union guard of the reference / reciprocal out of range: exact form decides
    const bool gt = q > thr + 1e-5f, lt = q < thr - 1e-5f;
    undecided = !zero && (degenerate || !(gt || lt));
    return !zero && gt;
}
__device__ __noinline__ bool iou_gt_exact(const float4 a, const float4 b, float thr) { return iou_ref(a, b) > thr; }

// The reference's greedy chain over <= 32 candidates in sorted order (applyNMS, onnx_engine.cpp:856-875), given for every
// candidate i (= lane) the mask `row` of EARLIER candidates whose IoU with it exceeds the threshold and the mask `und` of the
// candidates still alive.  "i is kept iff it is alive and no kept t < i suppresses it" has one solution; instead of walking
// the kept boxes one by one (a dependent shuffle + two bit operations per link, ~130 cycles) the warp settles, per round,
// every candidate whose earlier suppressors are all settled: removed if one of them is kept, kept if none is left
// undecided.  The first undecided candidate always settles, so the loop ends; real segments take 2-4 rounds.
__device__ __forceinline__ uint32_t resolve_rows(uint32_t und, uint32_t row, int lane) {
    uint32_t kept = 0u;
    while (und != 0u) {
        const bool me = (und >> lane) & 1u;
        const bool rem = me && (row & kept) != 0u;
        const bool kp = me && !rem && (row & und) == 0u;
        const unsigned kb = __ballot_sync(0xffffffffu, kp), rb = __ballot_sync(0xffffffffu, rem);
        kept |= kb;
        und &= ~(kb | rb);
    }
    return kept;
}
// One CTA per frame, or — launched as thread-block CLUSTERS of S CTAs (launch_nms) — S CTAs per frame, each owning a
// contiguous range of classes.  Classes never interact in applyNMS (onnx_engine.cpp:856-875 compares class ids before any
// IoU), so the ranks sort, sweep and compact their own candidates independently; they only exchange their kept counts
// (distributed shared memory) so that the frame's detections land contiguously in (class asc, confidence desc) order.
// The class ranges are cut where the running candidate count crosses multiples of n/S: balanced up to one class.
// Dynamic smem: keys[P] (P = pow2 >= n) | removed bitmask | sorted boxes (the class histogram of the split lives there first).
// CL = false is the instantiation of the engine's step (one CTA per frame, no cluster code at all); CL = true the one of
// the stand-alone decode + NMS call at small batches.
template <bool CL>
__global__ void __launch_bounds__(kNmsThreads)
nms_kernel(int A, float iou_thr, int key_cap_smem, int box_cap_smem, int key_pitch, uint64_t* __restrict__ keys_g, const float4* __restrict__ box_by_anchor,
           float4* __restrict__ sorted_box, const uint32_t* __restrict__ cand_count, uint32_t* __restrict__ header,
           DevDet* __restrict__ dets, int maxn, uint32_t cap, uint32_t* __restrict__ host_hdr)
{
    extern __shared__ __align__(16) uint8_t nms_smem[];
    __shared__ uint32_t s_warp_tot[32];
    __shared__ uint32_t s_base, s_mine, s_lower;
    __shared__ uint32_t s_xkept[8], s_xbase;                  // written by the other ranks of the cluster
    __shared__ int s_nlarge, s_qhead, s_gq[8], s_nk[8];
    __shared__ float4 s_kbox[8][32];                           // kept boxes of the block a team is sweeping with
    __shared__ int s_large[2 * kMaxLargeSeg];
    __shared__ uint32_t s_col[2 * 8][32];                      // suppression-matrix columns: two buffers per team
    namespace cg = cooperative_groups;
    int S = 1, rank = 0;
    if constexpr (CL) { S = (int)cg::this_cluster().num_blocks(); rank = (int)cg::this_cluster().block_rank(); }
    const int f = blockIdx.x / S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#if ZL_NMS_STAMPS
    const bool dbg_on = f == g_nms_dbg_frame && rank == 0;
    int dbg_i = 0;
    bool dbg_seg = false;                                    // this thread leads the group that drew the first queued segment
#define ZL_NMS_FINE(tag) do { if (dbg_seg && dbg_i < 31) { g_nms_dbg2[2 * dbg_i] = (tag); g_nms_dbg2[2 * dbg_i + 1] = clock64(); ++dbg_i; } } while (0)
#define ZL_NMS_FINE_ARM(cond) dbg_seg = dbg_on && (cond)
    if (tid == 0 && f == g_nms_dbg_frame) { g_nms_dbg[rank][6] = clock64(); g_nms_dbg[rank][8] = nms_gtime(); }
#else
#define ZL_NMS_FINE(tag) do { } while (0)
#define ZL_NMS_FINE_ARM(cond) do { } while (0)
#endif
    // launched with programmatic stream serialization behind the decode kernel: everything above ran under its tail
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // The count is read with a volatile asm load: cand_count is `const __restrict__`, and nvcc hoists such (invariant,
    // LDG.CONSTANT) loads ABOVE the asm statement of griddepcontrol.wait whatever its memory clobber — seen in the SASS of
    // this kernel, where the NMS then started on a count the head kernel had not finished writing.
    uint32_t n_raw;
    asm volatile("ld.global.u32 %0, [%1];" : "=r"(n_raw) : "l"(cand_count + f) : "memory");
    const int n_all = (int)min(n_raw, (uint32_t)A);
    uint32_t* h_total = header;
    uint32_t* h_cnt = header + 4;
    uint32_t* h_off = header + 4 + maxn;
    // few candidates: the split (histogram + two cluster barriers) costs more than it saves; rank 0 does the frame alone.
    // n_all is the same in every rank, so the whole cluster takes the same branch (no rank waits on a barrier alone).
    const bool split = CL && S > 1 && n_all > kNmsSplitMin;
    if (n_all == 0 || (!split && rank != 0)) {
        if (tid == 0 && rank == 0) {
            h_cnt[f] = 0; h_off[f] = 0;
            if (host_hdr) { host_hdr[0] = 0u; host_hdr[4] = 0u; host_hdr[4 + maxn] = 0u; }      // single-frame launch (see below)
        }
        return;
    }
    uint64_t* gkeys = keys_g + (size_t)f * key_pitch;
    uint64_t* skeys = reinterpret_cast<uint64_t*>(nms_smem);
    volatile uint32_t* removed = reinterpret_cast<volatile uint32_t*>(nms_smem + (size_t)key_cap_smem * 8);
    int n = n_all;
    if (split) {
        // every rank must be running before anyone writes into its shared memory: arrive now, wait right before the exchange
        if constexpr (CL) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        // launch_nms only forms clusters when every frame's keys fit the smem sort buffer (key_cap_smem >= pow2(A))
        uint32_t* hist = reinterpret_cast<uint32_t*>(nms_smem + (size_t)key_cap_smem * 8 + (size_t)(((((A + 31) >> 5) * 4) + 15) & ~15));
        for (int i = tid; i < kMaxClasses; i += kNmsThreads) hist[i] = 0u;
        if (tid == 0) { s_mine = 0u; s_lower = 0u; }
        __syncthreads();
        for (int i = tid; i < n_all; i += kNmsThreads) atomicAdd(&hist[key_class(gkeys[i])], 1u);
        __syncthreads();
        // exclusive prefix over the class bins, in place (thread t owns bins 4t .. 4t+3)
        static_assert(kMaxClasses == 4 * kNmsThreads, "one uint4 of class bins per thread");
        const uint4 hc = reinterpret_cast<const uint4*>(hist)[tid];
        const uint32_t mine4 = hc.x + hc.y + hc.z + hc.w;
        uint32_t x = mine4;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) s_warp_tot[warp] = x;
        __syncthreads();
        uint32_t woff = 0;
        for (int w = 0; w < warp; ++w) woff += s_warp_tot[w];
        const uint32_t e0 = woff + x - mine4;
        reinterpret_cast<uint4*>(hist)[tid] = make_uint4(e0, e0 + hc.x, e0 + hc.x + hc.y, e0 + hc.x + hc.y + hc.z);
        __syncthreads();
        // this rank's candidates -> smem, any order (the sort follows)
        for (int i0 = 0; i0 < n_all; i0 += kNmsThreads) {
            const int i = i0 + tid;
            uint64_t k = 0;
            bool mine = false, lower = false;
            if (i < n_all) {
                k = gkeys[i];
                const int r = min(S - 1, (int)((hist[key_class(k)] * (uint32_t)S) / (uint32_t)n_all));   // owner of the class: where its first candidate falls
                mine = r == rank;
                lower = r < rank;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, mine), ball = __ballot_sync(0xffffffffu, lower);
            uint32_t wbase = 0;
            if (lane == 0 && ball != 0u) atomicAdd(&s_lower, (uint32_t)__popc(ball));      // candidates of the ranks below: this rank's slot in the global box scratch
            if (lane == 0 && bal != 0u) wbase = atomicAdd(&s_mine, (uint32_t)__popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (mine) skeys[wbase + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = k;
        }
        __syncthreads();
        n = (int)s_mine;
    }
    int P = 1;
    while (P < n) P <<= 1;
    // keys live in smem when they fit, else they are sorted in place in global memory
    // (each frame's slice holds key_pitch = pow2 >= A slots, so the +inf padding is real)
    const bool in_smem = P <= key_cap_smem;
    const int nwords = (n + 31) >> 5;

    if (split) {
        for (int i = n + tid; i < P; i += kNmsThreads) skeys[i] = ~0ull;
    } else if (in_smem) {
        for (int i = tid; i < P; i += kNmsThreads) skeys[i] = i < n ? gkeys[i] : ~0ull;
    } else {
        for (int i = n + tid; i < P; i += kNmsThreads) gkeys[i] = ~0ull;
    }
    for (int i = tid; i < nwords; i += kNmsThreads) removed[i] = 0u;
    __syncthreads();

    ZL_NMS_STAMP(0);
    // ---- bitonic sort, ascending.  The five innermost strides (16..1) of every merge stay inside aligned 32-key groups, so
    // a warp runs them in registers with shuffles (no memory traffic, no barrier), and the first five merges (k = 2..32)
    // are one register pass.  Strides >= 32 go through memory: every thread owns whole comparator PAIRS (no idle half), and
    // two strides (j, j/2) are taken per barrier on quads {i, i|j/2, i|j, i|j|j/2} while both are >= 32:
    // 16 block barriers + 8 register passes for 4096 keys (it was 28 + 12, with half of the warps idle in each step).
    if (n > 1) {
        uint64_t* kk = in_smem ? skeys : gkeys;
        auto cmpx = [](uint64_t& a, uint64_t& b, bool up) { if ((a > b) == up) { const uint64_t t = a; a = b; b = t; } };
        auto warp_steps = [&](uint64_t x, int i, int k, int jfirst) {
            const bool up = (i & k) == 0;
            for (int jj = jfirst; jj >= 1; jj >>= 1) {
                const uint64_t y = __shfl_xor_sync(0xffffffffu, x, jj);
                const bool lower = (lane & jj) == 0;                          // this lane keeps the smaller key of the pair when ascending
                const bool take_min = lower == up;
                x = take_min ? (x < y ? x : y) : (x > y ? x : y);
            }
            return x;
        };
        if (P >= 32) {
            for (int base = warp * 32; base < P; base += kNmsThreads) {      // each warp owns whole 32-key groups
                const int i = base + lane;
                uint64_t x = kk[i];
                for (int k = 2; k <= 32; k <<= 1) x = warp_steps(x, i, k, k >> 1);
                kk[i] = x;
            }
            __syncthreads();
            for (int k = 64; k <= P; k <<= 1) {
                int j = k >> 1;
                while (j >= 64) {
                    const int h = j >> 1, lowm = h - 1;
                    for (int q = tid; q < (P >> 2); q += kNmsThreads) {
                        const int i = (q & lowm) | ((q & ~lowm) << 2);          // bits h and j clear
                        const bool up = (i & k) == 0;
                        uint64_t x0 = kk[i], x1 = kk[i | h], x2 = kk[i | j], x3 = kk[i | j | h];
                        cmpx(x0, x2, up); cmpx(x1, x3, up);                     // stride j
                        cmpx(x0, x1, up); cmpx(x2, x3, up);                     // stride j/2
                        kk[i] = x0; kk[i | h] = x1; kk[i | j] = x2; kk[i | j | h] = x3;
                    }
                    __syncthreads();
                    j >>= 2;
                }
                if (j == 32) {
                    for (int q = tid; q < (P >> 1); q += kNmsThreads) {
                        const int i = (q & 31) | ((q & ~31) << 1);              // bit 32 clear
                        const bool up = (i & k) == 0;
                        uint64_t a = kk[i], b = kk[i | 32];
                        if ((a > b) == up) { kk[i] = b; kk[i | 32] = a; }
                    }
                    __syncthreads();
                }
                for (int base = warp * 32; base < P; base += kNmsThreads) {
                    const int i = base + lane;
                    kk[i] = warp_steps(kk[i], i, k, 16);
                }
                __syncthreads();
            }
        } else {
            for (int k = 2; k <= P; k <<= 1) {
                for (int j = k >> 1; j >= 1; j >>= 1) {                           // tiny P (< 32): plain steps
                    for (int i = tid; i < P; i += kNmsThreads) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const uint64_t a = kk[i], b = kk[ixj];
                            const bool up = (i & k) == 0;
                            if ((a > b) == up) { kk[i] = b; kk[ixj] = a; }
                        }
                    }
                    __syncthreads();
                }
            }
        }
    }
    const uint64_t* K = in_smem ? skeys : gkeys;

    ZL_NMS_STAMP(1);
    // ---- gather boxes into sorted order: shared memory when they fit (the greedy sweep below is a chain of
    // dependent box reads, so its latency is the box read latency), else the global scratch
    float4* sb = (n <= box_cap_smem)
                     ? reinterpret_cast<float4*>(nms_smem + (size_t)key_cap_smem * 8 + (size_t)(((((A + 31) >> 5) * 4) + 15) & ~15))
                     : sorted_box + (size_t)f * A + (split ? s_lower : 0u);
    for (int i = tid; i < n; i += kNmsThreads) sb[i] = box_by_anchor[(size_t)f * A + key_anchor(K[i])];
    __syncthreads();

    ZL_NMS_STAMP(2);
    // ---- greedy per-class suppression (applyNMS, onnx_engine.cpp:856-875), exact, in two phases.
    // A class segment is a run of equal class ids in the sorted list.  The reference's loop is a serial chain over the
    // KEPT candidates of a segment; each link tests the still-alive later candidates, which is the parallel part.
    //  phase 1: segments of <= 32 candidates — one warp each: every lane builds its row of the suppression matrix (eight
    //           independent IoU tests in flight), then the chain is settled in rounds of two ballots (resolve_rows);
    //  phase 2: larger segments in blocks of 32 by teams of 128 ... 1024 threads (below).
    if (n > 1) {
        if (tid == 0) { s_nlarge = 0; s_qhead = 0; }
        __syncthreads();
        for (int base = warp * 32; base < n; base += kNmsThreads) {      // a warp looks at 32 sorted positions at once
            const int pos = base + lane;
            const int cls = pos < n ? key_class(K[pos]) : -1;
            int prev = __shfl_up_sync(0xffffffffu, cls, 1);
            if (lane == 0) prev = pos > 0 && pos < n ? key_class(K[pos - 1]) : -2;
            unsigned heads = __ballot_sync(0xffffffffu, pos < n && cls != prev);
            while (heads != 0u) {
                const int hl = __ffs(heads) - 1;
                heads &= heads - 1u;
                const int head = base + hl;
                const int hcls = __shfl_sync(0xffffffffu, cls, hl);
                // the next 32 positions: where the class changes is the end of a small segment
                const int pp = head + 1 + lane;
                const unsigned db = __ballot_sync(0xffffffffu, pp >= n || key_class(K[pp]) != hcls);
                if (db == 0u) {                                  // more than 32 candidates: find the end, queue for phase 2
                    int lo = head + 33, hi = n;                  // first index with a larger class
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (key_class(K[mid]) > hcls) hi = mid; else lo = mid + 1;
                    }
                    if (lane == 0) { const int q = atomicAdd(&s_nlarge, 1); if (q < kMaxLargeSeg) { s_large[2 * q] = head; s_large[2 * q + 1] = lo; } }
                    continue;
                }
                const int m = __ffs(db);                         // candidates head .. head + m - 1
                if (m == 1) continue;
                // each candidate's row of the suppression matrix, eight independent tests in flight, then the chain in rounds
                const int i = head + lane;
                const bool valid = lane < m;
                const float4 bi = sb[min(i, n - 1)];
                const BoxC ci = box_corners(bi);
                uint32_t row = 0u;
                for (int t0 = 0; t0 < m - 1; t0 += 8) {
                    bool r[8], u[8], any_u = false;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        r[k] = iou_gt_fast(box_corners(sb[min(head + t0 + k, n - 1)]), ci, iou_thr, u[k]);
                        any_u |= u[k];
                    }
                    if (any_u) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (u[k]) r[k] = iou_gt_exact(sb[min(head + t0 + k, n - 1)], bi, iou_thr);
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (r[k] && t0 + k < lane) row |= 1u << (t0 + k);       // strict '>' (onnx_engine.cpp:871); t < i only
                }
                const uint32_t all = m == 32 ? 0xffffffffu : ((1u << m) - 1u);
                const uint32_t kept = resolve_rows(all, valid ? (row & all) : 0u, lane);
                if (valid && !((kept >> lane) & 1u)) atomicOr(const_cast<uint32_t*>(&removed[i >> 5]), 1u << (i & 31));
            }
        }
        __syncthreads();
        ZL_NMS_STAMP(3);
        const int nlarge = min(s_nlarge, kMaxLargeSeg);
        const bool overflow = s_nlarge > kMaxLargeSeg;                      // more large segments than queue slots: handled below
        // Blocked greedy sweep with bitmask suppression, exact.  A segment is walked in blocks of 32 candidates by a TEAM of
        // 128 ... 1024 threads.  By the time a block is reached every candidate in it has been tested against the kept boxes
        // of ALL earlier blocks, so
        //  (a) the block resolves internally from its 32 x 32 suppression matrix: column t = the ballot of "box t suppresses
        //      box i" over the lanes i > t.  The matrix does not depend on what earlier blocks removed: while the team's first
        //      warp resolves block b (the chain over its kept boxes), its other warps already compute the columns of block b + 1;
        //  (b) all threads of the team test the later candidates against the block's kept boxes in parallel.
        // Every bulk IoU test runs kSweepW independent tests in flight (iou_gt_fast): one test alone is ~350 cycles of latency.
        // Team size: the whole CTA, one segment at a time, for very large segments (one class holding hundreds of candidates);
        // for the rest the CTA splits into as few teams as there are segments (1 / 2 / 4 / 8): the b=1 latency frame has four
        // classes, and 128-thread teams left half of the CTA idle.
        auto block_columns = [&](int b0, int e0, int t_first, int t_stride, uint32_t* col) {   // columns t_first, t_first + t_stride, ... < 32
            const int i = b0 + lane;
            const bool valid = i < e0;
            const float4 bi = sb[min(i, e0 - 1)];
            const BoxC ci = box_corners(bi);
            for (int tb = t_first; tb < 32; tb += kSweepW * t_stride) {
                bool r[kSweepW], u[kSweepW], any_u = false;
#pragma unroll
                for (int k = 0; k < kSweepW; ++k) {                          // independent chains, interleaved
                    r[k] = iou_gt_fast(box_corners(sb[min(b0 + tb + k * t_stride, e0 - 1)]), ci, iou_thr, u[k]);
                    any_u |= u[k];
                }
                if (any_u) {
#pragma unroll
                    for (int k = 0; k < kSweepW; ++k)
                        if (u[k]) r[k] = iou_gt_exact(sb[min(b0 + tb + k * t_stride, e0 - 1)], bi, iou_thr);
                }
#pragma unroll
                for (int k = 0; k < kSweepW; ++k) {
                    const int t = tb + k * t_stride;
                    if (t < 32) {                                            // warp-uniform
                        const bool sup = valid && b0 + t < e0 && t < lane && r[k];   // strict '>' (onnx_engine.cpp:871)
                        const unsigned bal = __ballot_sync(0xffffffffu, sup);
                        if (lane == 0) col[t] = bal;
                    }
                }
            }
        };
        auto block_resolve = [&](int b0, int e0, const uint32_t* col, float4* kbox, int* nk_out) {      // one warp
            const int i = b0 + lane;
            const bool valid = i < e0;
            const bool alive = valid && (((removed[i >> 5] >> (i & 31)) & 1u) == 0u);
            uint32_t cur = __ballot_sync(0xffffffffu, alive);                // alive, not yet decided
            const uint32_t mycol = col[lane];
            const float4 bi = sb[min(i, e0 - 1)];
            uint32_t kept = 0u;
            while (cur != 0u) {                                              // the reference's chain over the kept boxes: one shuffle per link
                const int t = __ffs(cur) - 1;                                // (measured: ~130 cycles per link; transposing the columns for
                kept |= 1u << t;                                             //  resolve_rows costs more than the 8-13 links of a block)
                cur &= ~(1u << t);
                cur &= ~__shfl_sync(0xffffffffu, mycol, t);                  // everything box t suppresses
            }
            if (alive && !((kept >> lane) & 1u)) atomicOr(const_cast<uint32_t*>(&removed[i >> 5]), 1u << (i & 31));
            if ((kept >> lane) & 1u) kbox[__popc(kept & ((1u << lane) - 1u))] = bi;
            if (lane == 0) *nk_out = __popc(kept);
        };
        // later candidates of the segment against the block's kept boxes.  When fewer candidates are left than the team has
        // threads, 2 / 4 / 8 threads share a candidate and split the kept boxes between them (setting a flag twice is harmless).
        auto block_sweep = [&](int j_first, int e0, int gtid, int T, const float4* kbox, int nk) {
            const int left = e0 - j_first;
            int psh = 0;                                                     // log2(threads per candidate)
            while (psh < 3 && (left << (psh + 1)) <= T && (kSweepW << psh) < nk) ++psh;
            const int part = gtid & ((1 << psh) - 1);
            for (int j = j_first + (gtid >> psh); j < e0; j += T >> psh) {
                if (((removed[j >> 5] >> (j & 31)) & 1u) != 0u) continue;
                const float4 bj = sb[j];
                const BoxC cj = box_corners(bj);
                bool sup = false;
                for (int t = part * kSweepW; t < nk && !sup; t += kSweepW << psh) {
                    bool r[kSweepW], u[kSweepW], any_u = false;
#pragma unroll
                    for (int k = 0; k < kSweepW; ++k) {                      // past the end: the last kept box again (same answer)
                        r[k] = iou_gt_fast(box_corners(kbox[min(t + k, nk - 1)]), cj, iou_thr, u[k]);
                        any_u |= u[k];
                    }
                    if (any_u) {
#pragma unroll
                        for (int k = 0; k < kSweepW; ++k)
                            if (u[k]) r[k] = iou_gt_exact(kbox[min(t + k, nk - 1)], bj, iou_thr);
                    }
#pragma unroll
                    for (int k = 0; k < kSweepW; ++k) sup = sup || r[k];
                }
                if (sup) atomicOr(const_cast<uint32_t*>(&removed[j >> 5]), 1u << (j & 31));
            }
        };
        // one segment by one team of T threads (team index grp, thread gtid of the team, named barrier grp + 1)
        auto run_segment = [&](int s0, int e0, int T, int grp, int gtid, bool stamp) {
            const int TW = T >> 5, gw = gtid >> 5;
            uint32_t* gcol = s_col[2 * grp];                               // the team's two column buffers
            ZL_NMS_FINE_ARM(stamp);
            ZL_NMS_FINE(1000 + (e0 - s0));
            block_columns(s0, e0, gw, TW, gcol);
            asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(T) : "memory");
            ZL_NMS_FINE(1);
            int buf = 0;
            for (int b0 = s0; b0 < e0; b0 += 32, buf ^= 1) {
                if (gw == 0) block_resolve(b0, e0, gcol + 32 * buf, s_kbox[grp], &s_nk[grp]);
                else if (b0 + 32 < e0) block_columns(b0 + 32, e0, gw - 1, TW - 1, gcol + 32 * (buf ^ 1));
                ZL_NMS_FINE(2);
                asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(T) : "memory");
                const int nk = s_nk[grp];
                ZL_NMS_FINE(300 + nk);
                if (nk > 0 && b0 + 32 < e0) block_sweep(b0 + 32, e0, gtid, T, s_kbox[grp], nk);
                ZL_NMS_FINE(4);
                asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(T) : "memory");       // flags and kept boxes settled before the next block
                ZL_NMS_FINE(6);
            }
        };
        int nrest = 0;                                                      // segments the teams of pass 2 share
        for (int q = 0; q < nlarge; ++q) {
            const int s0 = s_large[2 * q], e0 = s_large[2 * q + 1];
            if (e0 - s0 <= kWholeCtaSeg) { ++nrest; continue; }
            run_segment(s0, e0, kNmsThreads, 0, tid, false);
        }
        if (nrest > 0) {
            const int T = nrest <= 1 ? 1024 : (nrest <= 2 ? 512 : (nrest <= 4 ? 256 : 128));
            const int grp = tid / T, gtid = tid - grp * T;
            for (;;) {
                // the team's next segment: drawn by its first thread, broadcast through smem + the team's named barrier
                if (gtid == 0) s_gq[grp] = atomicAdd(&s_qhead, 1);
                asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(T) : "memory");
                const int q = s_gq[grp];
                asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(T) : "memory");
                if (q >= nlarge) break;
                const int s0 = s_large[2 * q], e0 = s_large[2 * q + 1];
                if (e0 - s0 > kWholeCtaSeg) continue;                      // done above by the whole CTA
                run_segment(s0, e0, T, grp, gtid, gtid == 0 && q == 0);
            }
        }
        if (overflow) {
            // pathological: > kMaxLargeSeg large segments (needs > 64 classes with > 32 candidates each AND queue overflow);
            // fall back to the plain sequential sweep for the segments that did not get a slot
            __syncthreads();
            for (int head = warp; head < n; head += kNmsThreads / 32) {
                const bool is_head = head == 0 || key_class(K[head]) != key_class(K[head - 1]);
                if (!is_head) continue;
                const int cls = key_class(K[head]);
                int lo = head + 1, hi = n;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (key_class(K[mid]) > cls) hi = mid; else lo = mid + 1; }
                const int end = lo;
                if (end - head <= 32) continue;
                bool queued = false;
                for (int q = 0; q < kMaxLargeSeg; ++q) if (s_large[2 * q] == head) { queued = true; break; }
                if (queued) continue;
                for (int i = head; i < end - 1; ++i) {
                    if ((removed[i >> 5] >> (i & 31)) & 1u) continue;
                    const float4 bi = sb[i];
                    for (int j0 = (i + 1) & ~31; j0 < end; j0 += 32) {
                        const int j = j0 + lane;
                        bool sup = false;
                        if (j > i && j < end && !((removed[j >> 5] >> (j & 31)) & 1u)) sup = iou_gt(bi, sb[j], iou_thr);
                        const unsigned bal = __ballot_sync(0xffffffffu, sup);
                        if (bal != 0u && lane == 0) atomicOr(const_cast<uint32_t*>(&removed[j0 >> 5]), bal);
                    }
                    __syncwarp();
                }
            }
        }
    }
    __syncthreads();

    ZL_NMS_STAMP(4);
    // ---- ordered compaction of survivors
    uint32_t kept_total = 0;
    for (int i0 = 0; i0 < n; i0 += kNmsThreads) {
        const int i = i0 + tid;
        const bool keep = i < n && !((removed[i >> 5] >> (i & 31)) & 1u);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp_tot[warp] = __popc(bal);
        __syncthreads();
        uint32_t woff = 0, tot = 0;
        for (int w = 0; w < kNmsThreads / 32; ++w) { const uint32_t c = s_warp_tot[w]; if (w < warp) woff += c; tot += c; }
        kept_total += tot;
        __syncthreads();
    }
    if (!split) {
        if (tid == 0) {
            s_base = atomicAdd(h_total, kept_total);
            h_cnt[f] = kept_total;
            h_off[f] = s_base;
            // host_hdr != nullptr: a ONE-frame launch of the engine's b=1 step.  The CTA writes the result block (header +
            // records, the layout of the device block) straight into the lane's pinned host buffer: the device-to-host copy
            // node behind the NMS (~4 us on the latency path) is not needed.
            if (host_hdr) { host_hdr[0] = s_base + kept_total; host_hdr[4] = kept_total; host_hdr[4 + maxn] = s_base; }
        }
        __syncthreads();
    } else if constexpr (CL) {
        // every rank tells every rank how many it kept; rank 0 reserves the frame's slice and tells everyone where it starts
        cg::cluster_group cl = cg::this_cluster();
        asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
        if (tid < S) *cl.map_shared_rank(&s_xkept[rank], tid) = kept_total;
        cl.sync();
        if (rank == 0 && tid == 0) {
            uint32_t total = 0;
            for (int r = 0; r < S; ++r) total += s_xkept[r];
            const uint32_t b = atomicAdd(h_total, total);
            h_cnt[f] = total;
            h_off[f] = b;
            for (int r = 0; r < S; ++r) *cl.map_shared_rank(&s_xbase, r) = b;
        }
        cl.sync();                                            // no remote access after this barrier: ranks may exit independently
        if (tid == 0) {
            uint32_t b = s_xbase;
            for (int r = 0; r < rank; ++r) b += s_xkept[r];
            s_base = b;
        }
        __syncthreads();
    }
    const uint32_t base = s_base;
    uint32_t running = 0;
    for (int i0 = 0; i0 < n; i0 += kNmsThreads) {
        const int i = i0 + tid;
        const bool keep = i < n && !((removed[i >> 5] >> (i & 31)) & 1u);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp_tot[warp] = __popc(bal);
        __syncthreads();
        uint32_t woff = 0, tot = 0;
        for (int w = 0; w < kNmsThreads / 32; ++w) { const uint32_t c = s_warp_tot[w]; if (w < warp) woff += c; tot += c; }
        if (keep) {
            const uint32_t slot = base + running + woff + (uint32_t)__popc(bal & ((1u << lane) - 1u));
            if (slot < cap) {
                const uint64_t k = K[i];
                const float4 b = sb[i];
                DevDet d;
                d.x = b.x; d.y = b.y; d.w = b.z; d.h = b.w; d.conf = key_conf(k); d.cls = key_class(k);
                dets[slot] = d;
                if (host_hdr) reinterpret_cast<DevDet*>(host_hdr + 4 + 2 * maxn)[slot] = d;
            }
        }
        running += tot;
        __syncthreads();
    }
    ZL_NMS_STAMP(5);
#if ZL_NMS_STAMPS
    if (tid == 0 && f == g_nms_dbg_frame) { g_nms_dbg[rank][7] = n; g_nms_dbg[rank][9] = nms_gtime(); }
    if (dbg_i > 0) g_nms_dbg2[62] = dbg_i;
#endif
#undef ZL_NMS_FINE
#undef ZL_NMS_FINE_ARM
}

int g_nms_smem_keys = 0;   // key capacity (elements) of the smem sort buffer

// ---- N3: the reference's result wire layout, written by the device -----------------------------------------------------
// One CTA per batch.  Phase 1: exclusive scan of the per-frame block lengths (14 + 40 * count) in FRAME order, so the
// blocks are laid out deterministically whatever order the NMS CTAs finished in.  Phase 2: one warp per frame writes the
// 14-byte header and the records with 16-bit stores (a block starts at an even, otherwise unaligned, offset).
__device__ __forceinline__ void st16(uint8_t* p, uint32_t v) { *reinterpret_cast<uint16_t*>(p) = (uint16_t)v; }
__device__ __forceinline__ void st32u(uint8_t* p, uint32_t v) { st16(p, v & 0xffffu); st16(p + 2, v >> 16); }

__global__ void __launch_bounds__(256)
wire_pack_kernel(int n, int maxn, const uint32_t* __restrict__ header, const DevDet* __restrict__ dets, const WireMeta* __restrict__ meta,
                 uint8_t* __restrict__ wire, uint32_t wire_cap, uint32_t* __restrict__ wire_off)
{
    __shared__ uint32_t s_off[257];
    const int tid = threadIdx.x;
    const uint32_t* h_cnt = header + 4;
    const uint32_t* h_off = header + 4 + maxn;
    if (tid == 0) {
        uint32_t run = 0;
        for (int i = 0; i < n; ++i) { s_off[i] = run; run += (uint32_t)kWireHeader + (uint32_t)kWireDet * h_cnt[i]; }
        s_off[n] = run;
    }
    __syncthreads();
    for (int i = tid; i <= n; i += blockDim.x) wire_off[i] = s_off[i];
    const uint64_t det_ts = meta[maxn].timestamp;
    const int warp = tid >> 5, lane = tid & 31;
    for (int f = warp; f < n; f += (int)blockDim.x >> 5) {
        const uint32_t cnt = h_cnt[f], src = h_off[f], base = s_off[f];
        if (lane == 0 && base + kWireHeader <= wire_cap) {
            uint8_t* o = wire + base;
            st32u(o, meta[f].frame_id);
            st32u(o + 4, (uint32_t)meta[f].timestamp); st32u(o + 8, (uint32_t)(meta[f].timestamp >> 32));
            st16(o + 12, cnt);                                       // static_cast<uint16_t>(detections.size())
        }
        for (uint32_t j = lane; j < cnt; j += 32) {
            const uint32_t at = base + kWireHeader + j * kWireDet;
            if (at + kWireDet > wire_cap) break;
            const DevDet d = dets[src + j];
            uint8_t* o = wire + at;
            st32u(o, __float_as_uint(d.x)); st32u(o + 4, __float_as_uint(d.y)); st32u(o + 8, __float_as_uint(d.w)); st32u(o + 12, __float_as_uint(d.h));
            st32u(o + 16, __float_as_uint(d.conf)); st32u(o + 20, (uint32_t)d.cls);
            st32u(o + 24, 0u);                                       // track_id = 0 (onnx_engine.cpp:812)
            st32u(o + 28, 0u);                                       // struct padding
            st32u(o + 32, (uint32_t)det_ts); st32u(o + 36, (uint32_t)(det_ts >> 32));
        }
    }
}

}  // namespace

int32_t launch_wire_pack(cudaStream_t st, int32_t n, const PostBuffers& pb, const WireMeta* meta, uint8_t* wire, uint32_t wire_cap, uint32_t* wire_off)
{
    if (n < 1 || n > 256) ZL_FAIL(ZL_INVALID_ARGUMENT, "wire_pack: batch must be 1..256");
    wire_pack_kernel<<<1, 256, 0, st>>>(n, pb.maxn, pb.header, pb.dets, meta, wire, wire_cap, wire_off);
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

int32_t launch_dfl_decode(cudaStream_t st, const HeadLevel lv[3], int32_t n, int32_t nc, int32_t A, float* raw, bool precise)
{
    dim3 grid(ceil_div(A, kDflAnchors), n);
    const size_t smem = (size_t)(4 + nc) * 33 * sizeof(float);
    if (smem > 48 * 1024) ZL_FAIL(ZL_INVALID_ARGUMENT, "dfl_decode: nc too large");
    if (precise) dfl_decode_kernel<true><<<grid, kDflThreads, smem, st>>>(lv[0], lv[1], lv[2], nc, A, raw);
    else dfl_decode_kernel<false><<<grid, kDflThreads, smem, st>>>(lv[0], lv[1], lv[2], nc, A, raw);
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

int32_t launch_filter(cudaStream_t st, const float* raw, int32_t n, int32_t nc, int32_t A,
                      const FrameDesc* descs, const int32_t* img_wh, float conf_thr,
                      const float* class_weights, const PostBuffers& pb)
{
    if (A > kMaxAnchors || nc > kMaxClasses) ZL_FAIL(ZL_INVALID_ARGUMENT, "filter: A or nc beyond key range");
    dim3 grid(ceil_div(A, 256), n);
    filter_kernel<<<grid, 256, 0, st>>>(raw, nc, A, descs, img_wh, conf_thr, class_weights, pb.keys, pb.key_pitch, pb.box_by_anchor, pb.cand_count);
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

int32_t launch_decode_filter(cudaStream_t st, const HeadLevel lv[3], int32_t n, int32_t nc, int32_t A, const FrameDesc* descs,
                             float conf_thr, const float* class_weights, const PostBuffers& pb, bool precise)
{
    if (A > kMaxAnchors || nc > kMaxClasses) ZL_FAIL(ZL_INVALID_ARGUMENT, "decode_filter: A or nc beyond key range");
    dim3 grid(ceil_div(A, 128), n);
    const size_t smem = (size_t)4 * 32 * (nc + 1) * sizeof(float);
    if (smem <= 48 * 1024) {
        if (precise) decode_filter_kernel<true, true><<<grid, 128, smem, st>>>(lv[0], lv[1], lv[2], nc, A, descs, conf_thr, class_weights, pb.keys, pb.key_pitch, pb.box_by_anchor, pb.cand_count);
        else decode_filter_kernel<false, true><<<grid, 128, smem, st>>>(lv[0], lv[1], lv[2], nc, A, descs, conf_thr, class_weights, pb.keys, pb.key_pitch, pb.box_by_anchor, pb.cand_count);
    } else if (precise) {
        decode_filter_kernel<true, false><<<grid, 128, 0, st>>>(lv[0], lv[1], lv[2], nc, A, descs, conf_thr, class_weights, pb.keys, pb.key_pitch, pb.box_by_anchor, pb.cand_count);
    } else {
        decode_filter_kernel<false, false><<<grid, 128, 0, st>>>(lv[0], lv[1], lv[2], nc, A, descs, conf_thr, class_weights, pb.keys, pb.key_pitch, pb.box_by_anchor, pb.cand_count);
    }
    ZL_CUDA(cudaGetLastError());
    return ZL_OK;
}

int32_t nms_configure()
{
    // 16384 keys (128 KB) + removed bitmask for up to 2^20 candidates would not fit; the
    // bitmask is sized for kMaxAnchors only when keys spill to global.  Budget: 200 KB.
    ZL_CUDA(cudaFuncSetAttribute(nms_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    ZL_CUDA(cudaFuncSetAttribute(nms_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    g_nms_smem_keys = 16384;
    return ZL_OK;
}

// CTAs per frame (cluster size) for a batch of n frames: as many as keep every cluster resident at once.  One CTA owns
// an SM (200 KB of shared memory), a cluster lives inside one GPC (16+ usable SMs on B200).  ZL_NMS_SPLIT=1/2/4/8 overrides.
static int nms_split_for(int n, bool keys_fit_smem, bool allow_cluster)
{
    if (!keys_fit_smem) return 1;                  // global-memory sort works in place on the frame's key slice: one CTA only
    static const int forced = [] { const char* e = getenv("ZL_NMS_SPLIT"); return e ? atoi(e) : 0; }();
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return forced;
    if (!allow_cluster) return 1;
    return n <= 16 ? 8 : (n <= 32 ? 4 : (n <= 72 ? 2 : 1));
}

int32_t launch_nms(cudaStream_t st, int32_t n, int32_t A, float iou_thr, const PostBuffers& pb, bool allow_cluster, uint32_t* host_result)
{
    if (host_result && n != 1) ZL_FAIL(ZL_INVALID_ARGUMENT, "nms: the direct host result block is for one-frame launches");
    // smem plan: keys (only as many as can ever be needed) + removed bitmask (A bits)
    int key_cap = 1;
    while (key_cap < A) key_cap <<= 1;
    if (key_cap > 16384) key_cap = 0;              // too many to sort in smem -> sort in global memory
    const int split = host_result ? 1 : nms_split_for(n, key_cap != 0, allow_cluster);
    const size_t mask_bytes = ((size_t)ceil_div(A, 32) * 4 + 15) & ~(size_t)15;
    size_t smem = (size_t)key_cap * 8 + mask_bytes;
    if (smem > 200 * 1024) ZL_FAIL(ZL_INVALID_ARGUMENT, "nms: anchor count too large for the suppression bitmask");
    int box_cap = (int)((200 * 1024 - smem) / 16);
    if (box_cap > A) box_cap = A;
    size_t tail = (size_t)box_cap * 16;
    if (split > 1 && tail < (size_t)kMaxClasses * 4) tail = (size_t)kMaxClasses * 4;      // the class histogram of the split
    smem += tail;
    static thread_local int last_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != last_dev) { ZL_TRY(nms_configure()); last_dev = dev; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n * split); cfg.blockDim = dim3(kNmsThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    static const bool use_pdl = [] { const char* e = getenv("ZL_DISABLE_PDL"); return !(e && e[0] == '1'); }();
    // Cluster launches keep plain stream order (they only serve the stand-alone call, where nothing is gained from an early
    // start); ZL_NMS_CLUSTER_PDL=1 makes them programmatic dependents as well (results identical: tests/test_gpu_headfused.py).
    static const bool cluster_pdl = [] { const char* e = getenv("ZL_NMS_CLUSTER_PDL"); return e && e[0] == '1'; }();
    if (use_pdl && (split == 1 || cluster_pdl)) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (split > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = (unsigned)split; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    ZL_CUDA(cudaLaunchKernelEx(&cfg, split > 1 ? nms_kernel<true> : nms_kernel<false>, A, iou_thr, key_cap, box_cap, pb.key_pitch, pb.keys, (const float4*)pb.box_by_anchor, pb.sorted_box,
                               (const uint32_t*)pb.cand_count, pb.header, pb.dets, pb.maxn, pb.cap, host_result));
    static const char* dbg = getenv("ZL_NMS_DEBUG");
    if (dbg) {
        // debug aid: phase stamps of frame <ZL_NMS_DEBUG> of the PREVIOUS launch on this stream (the frame index is armed below)
        long long h[8][10];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_nms_dbg, sizeof(h));
        const int fr = atoi(dbg);
        int cur = -1;
        cudaMemcpyFromSymbol(&cur, g_nms_dbg_frame, sizeof(int));
        if (cur == fr)
            for (int r = 0; r < split; ++r)
                fprintf(stderr, "nms frame %d rank %d/%d: cand %lld | cycles: load/split %lld sort %lld gather %lld small-seg %lld large-seg %lld compact+exchange %lld | wall %lld ns, start +%lld ns\n",
                        fr, r, split, h[r][7], h[r][0] - h[r][6], h[r][1] - h[r][0], h[r][2] - h[r][1], h[r][3] - h[r][2], h[r][4] - h[r][3], h[r][5] - h[r][4],
                        h[r][9] - h[r][8], h[r][8] - h[0][8]);
        if (cur == fr) {
            long long h2[64];
            cudaMemcpyFromSymbol(h2, g_nms_dbg2, sizeof(h2));
            fprintf(stderr, "  large-segment fine stamps (tag:+cycles; 1000+m segment of m, 1 columns, 2 resolve, 300+nk after barrier, 4 sweep, 5 next columns, 6 barrier):");
            for (int i = 1; i < (int)h2[62] && i < 31; ++i) fprintf(stderr, " %lld:+%lld", h2[2 * i], h2[2 * i + 1] - h2[2 * i - 1]);
            fprintf(stderr, "\n");
        }
        cudaMemcpyToSymbol(g_nms_dbg_frame, &fr, sizeof(int));
    }
    return ZL_OK;
}

}  // namespace zl
