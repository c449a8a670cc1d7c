// kernels.h — launchers of the sm_100a kernels (one translation unit per family).
#pragma once
#include <vector>

#include "common.h"

namespace zl {

// ---------------------------------------------------------------- weights
// One convolution's parameters, resident on one device.  BN is already folded
// (SURVEY.md Appendix A), so a conv is y = act(W*x + b) (+ residual).
struct ConvWeights {
    std::string name;
    int32_t cin = 0, cout = 0, k = 1, stride = 1, act = 1;
    int32_t cout_pad = 0;          // multiple of 16 (tcgen05 N granularity)
    int32_t ktot = 0;              // k*k*cin
    float* w_simt = nullptr;       // fp32 [ktot][cout_pad]  (k index = (r*k+s)*cin + c)
    __nv_bfloat16* w_tc = nullptr; // bf16 [cout_pad][ktot]  K-major rows for UMMA B
    float* bias = nullptr;         // fp32 [cout_pad]
    ConvWeights() = default;
    ConvWeights(const ConvWeights&) = delete;
    ConvWeights& operator=(const ConvWeights&) = delete;
    ~ConvWeights() {               // owns its device buffers (load_weights error paths, hot reload, engine teardown)
        if (w_simt) cudaFree(w_simt);
        if (w_tc) cudaFree(w_tc);
        if (bias) cudaFree(bias);
    }
};

// ---------------------------------------------------------------- P1
enum PreLayout : int32_t {
    PRE_NCHW_F32 = 0,    // [n,3,mh,mw] fp32 planar: the tensor the reference feeds ORT (onnx_engine.cpp:560)
    PRE_NHWC4_F32 = 1,   // [n,mh,mw,4] fp32 (R,G,B,0): input of the fp32 conv path
    PRE_NHWC4_BF16 = 2,  // [n,mh,mw,4] bf16 (R,G,B,0): input of the bf16 conv path
    PRE_NHWC4_F16 = 3,   // same, IEEE half
    PRE_S2D16_BF16 = 4,  // [n,mh/2,mw/2,16] bf16: the 2x2 pixel block (dy,dx) x (R,G,B) = 12 values + 4 zeros (layer 0 on tensor cores)
    PRE_S2D16_F16 = 5
};
int32_t launch_preprocess(cudaStream_t st, const uint8_t* staging, const FrameDesc* descs, int32_t n,
                          int32_t mw, int32_t mh, int32_t layout, void* out, int32_t letterbox = 0, int32_t same_size = 0);   // same_size: every frame is mw x mh (host-checked)
// ZL_PRE_LETTERBOX: the kept boxes of a batch mapped back from the letterboxed model frame to the request frame.
int32_t launch_letterbox_unmap(cudaStream_t st, int32_t n, int32_t maxn, const uint32_t* header, DevDet* dets, const FrameDesc* descs,
                               int32_t mw, int32_t mh, uint32_t cap);

// ---------------------------------------------------------------- convs
// fp32 CUDA-core implicit GEMM (exact mode) — any cin/cout, k in {1,3}, stride in {1,2}.
int32_t launch_conv_simt(cudaStream_t st, const ConvWeights& w, const View& x, const View& y, const View* res);
// bf16 first layer (cin=3 padded to 4): direct conv on CUDA cores, bf16 NHWC out.
int32_t launch_conv0_direct(cudaStream_t st, const ConvWeights& w, const View& x, const View& y);
// P1 + layer 0 fused for the 16-bit modes: u8 frames in, 16-bit NHWC activations out (conv_simt.cu).
int32_t launch_pre_conv0(cudaStream_t st, const uint8_t* staging, const FrameDesc* descs, int32_t n, int32_t mw, int32_t mh,
                         const ConvWeights& w, const View& y);

// SiLU form of the 16-bit conv epilogues (half16.cuh): true = one-MUFU tanh form.  ZL_SILU=exp|tanh overrides the defaults.
constexpr bool kSiluTanhDefaultBf16 = true, kSiluTanhDefaultF16 = true;
bool conv_silu_tanh(bool f16);
// Channels per K chunk of a 16-bit conv: a function of the layer only, shared by both tcgen05 kernels (conv_tc.cu).
int32_t conv_kc(const ConvWeights& w, bool y_f32);

// tcgen05 implicit GEMM.  A operand staged either by TMA (1x1 convs: plain 2-D
// tiled map over [pixels][cin]) or by producer warps gathering NHWC rows into
// the swizzled UMMA layout (3x3, any stride); B (weights) always by TMA.
struct ConvTcOp {
    CUtensorMap tmap_w;
    CUtensorMap tmap_a;
    const __nv_bfloat16* x;
    void* y;
    const __nv_bfloat16* res;
    const float* bias;
    int32_t N, H, W, Cin, xpitch;
    int32_t Ho, Wo, Cout, ypitch, rpitch, y_f32;
    int32_t k, stride, pad, act;
    int32_t kc, swz, nkb, cchunks;     // channels per K-block, swizzle bytes, #K-blocks, chunks per tap
    int32_t ntile, ngrid;              // N tile (<=256, multiple of 16) and number of N tiles
    int32_t m_total, a_tma, y_vec, r_vec, f16, silu_tanh;
    int32_t stages, smem_bytes, tmem_cols;
    double flops, bytes;
};
int32_t conv_tc_prepare(const ConvWeights& w, const View& x, const View& y, const View* res,
                        bool allow_tma_a, int32_t ntile_hint, ConvTcOp* op);
int32_t conv_tc_launch(cudaStream_t st, const ConvTcOp& op);

// Persistent 3x3 stride-1 conv with halo reuse (conv_halo.cu): weights resident in smem, one TMA patch
// load per tile and channel chunk, nine shifted UMMA views of the same patch.
struct ConvHaloOp {
    CUtensorMap tmap_w;
    CUtensorMap tmap_x;
    CUtensorMap tmap_y;
    void* y;
    const __nv_bfloat16* res;
    const float* bias;
    int32_t N, H, W, Cin, Cout, ntile;
    int32_t ypitch, rpitch, y_f32, y_vec, r_vec, f16, act, silu_tanh;
    int32_t kc, cchunks, stages, sub, y_tma, taps, nsplit, nt, mode, ostage;
    int32_t tiles_x, tiles_y, num_tiles;
    int32_t wstream, nacc, acc_cols;     // weights streamed with the patches / resident; TMEM accumulator ring
    uint32_t wtile_bytes, wchunk_bytes, wchunk_alloc, patch_bytes, patch_alloc, subpatch_alloc, chunk_stride, stage_stride, tmem_cols;
    int32_t cps, nst;                    // channel chunks per pipeline stage, stages per tile
    int32_t smem_bytes;
    double flops, bytes;
};
// Layer 0 as a 2x2 conv over the space-to-depth image written by the preprocess kernel (conv_halo.cu, MODE 4).
int32_t conv_s2d_prepare(const ConvWeights& w, const View& x, const View& y, int num_sms, ConvHaloOp* op);
bool conv_halo_supported(const ConvWeights& w, const View& x, const View& y, int num_sms, int* work_units);
int32_t conv_halo_prepare(const ConvWeights& w, const View& x, const View& y, const View* res, int num_sms, ConvHaloOp* op);
// stats != nullptr launches the instrumented instantiation: kHaloStatSlots cycle counters summed over CTAs
//   0 producer waiting for a free patch stage   1 MMA warp waiting for a patch   2 MMA warp waiting for a drained accumulator
//   3 MMA warp issuing   4 epilogue warp 2 waiting for an accumulator   5 epilogue warp 2 busy   6 CTA lifetime (sum)
//   7 prologue (sum)   8 MMA warp waiting for the weights   9 slowest CTA   10 CTAs   11 plan (packed)
//   12-15 epilogue warp 2: tcgen05.ld wait / waiting for the staging block / st.shared + fence + store issue / accumulator hand-back
constexpr int kHaloStatSlots = 16;
int32_t conv_halo_launch(cudaStream_t st, const ConvHaloOp& op, int num_sms, unsigned long long* stats = nullptr);

// ---------------------------------------------------------------- pool / upsample
// SPPF: p1 = max5(a), p2 = max5(p1) = max9(a), p3 = max13(a) in one pass.
int32_t launch_sppf_pool(cudaStream_t st, const View& a, const View& p1, const View& p2, const View& p3);
int32_t launch_upsample2x(cudaStream_t st, const View& x, const View& y);

// ---------------------------------------------------------------- D1 / F1 / N1
struct HeadLevel {
    const float* box;   // [n,h,w,64] fp32
    const float* cls;   // [n,h,w,cls_pitch] fp32 (nc valid)
    int32_t h, w, stride, cls_pitch, a0;   // a0 = first anchor index of the level
};
// Detect tail -> raw head output [n, 4+nc, A] fp32 (the reference's output0).
int32_t launch_dfl_decode(cudaStream_t st, const HeadLevel lv[3], int32_t n, int32_t nc, int32_t A, float* raw, bool precise);

struct PostBuffers {
    uint64_t* keys;        // [n][key_pitch] candidate sort keys (key_pitch = pow2 >= A)
    int32_t key_pitch;
    float4* box_by_anchor; // [n][A]
    float4* sorted_box;    // [n][A] scratch
    uint32_t* cand_count;  // [n]
    uint32_t* header;      // [4 + 2*maxn]: total, pad[3], cnt[maxn], off[maxn]
    DevDet* dets;          // [cap]
    int32_t maxn;
    uint32_t cap;
};
// postProcess decode part (onnx_engine.cpp:773-819): argmax / threshold / normalise, ballot compaction.
int32_t launch_filter(cudaStream_t st, const float* raw, int32_t n, int32_t nc, int32_t A,
                      const FrameDesc* descs, const int32_t* img_wh, float conf_thr,
                      const float* class_weights, const PostBuffers& pb);
// D1 + F1 fused: head maps -> candidates without materialising the raw head tensor (engine hot path).
int32_t launch_decode_filter(cudaStream_t st, const HeadLevel lv[3], int32_t n, int32_t nc, int32_t A, const FrameDesc* descs,
                             float conf_thr, const float* class_weights, const PostBuffers& pb, bool precise);
// applyNMS (onnx_engine.cpp:837-878): per-frame key sort + per-class greedy bitmask suppression.
// allow_cluster: batches of up to 72 frames may deal a frame's classes to the 2 / 4 / 8 CTAs of a thread-block cluster (the
// stand-alone decode + NMS call: 2-6x shorter at small batches).  The engine's graph-captured step passes false: measured
// on B200 the split is neutral there (the step's throughput depends on the NMS's SM-time, not on its span).
// host_result (one-frame launches only): pinned host copy of the result block {header, records} the CTA writes itself, so
// that the b=1 step needs no device-to-host copy node.
int32_t launch_nms(cudaStream_t st, int32_t n, int32_t A, float iou_thr, const PostBuffers& pb, bool allow_cluster = false,
                   uint32_t* host_result = nullptr);
int32_t nms_configure();   // one-time cudaFuncSetAttribute calls

// model.22.cv2.l.2 + model.22.cv3.l.2 (the two last 1x1 convs of every level) + D1 + F1 in one persistent tcgen05 kernel
// (head_fused.cu): the fp32 logits stay in TMEM, only the candidates are written.  Same candidates, bit for bit, as the
// conv kernels followed by launch_decode_filter.  ZL_FUSE_HEAD=0 turns it off (A/B).
struct HeadFusedOp {
    alignas(64) unsigned char blob[2560];   // tensor maps + kernel parameters (head_fused.cu: HeadFusedImpl)
    int32_t grid, smem_bytes;
    double flops, bytes;
};
bool head_fused_supported(const ConvWeights* const wb[3], const ConvWeights* const wc[3], const View xb[3], const View xc[3], int nc);
int32_t head_fused_prepare(const ConvWeights* const wb[3], const ConvWeights* const wc[3], const View xb[3], const View xc[3],
                           const HeadLevel lvl[3], int nc, int A, int num_sms, HeadFusedOp* op);
int32_t head_fused_launch(cudaStream_t st, const HeadFusedOp& op, const FrameDesc* descs, float conf_thr, const float* class_weights, const PostBuffers& pb,
                          unsigned long long* stats = nullptr);     // stats: the instrumented instantiation (same slots as conv_halo_launch; 12-14 = scan pass 1 / pass 2 / box + emit)

// Result wire layout on the device: per frame {frame_id u32, timestamp u64, count u16} + count x 40-byte
// Detection, packed back to back in batch order (src/common/protocol.h:541-567, src/common/types.h:20-26).
struct WireMeta { uint32_t frame_id; uint32_t pad; uint64_t timestamp; };     // meta[n] per frame; meta[maxn].timestamp = Detection::timestamp of the batch
constexpr int kWireHeader = 14, kWireDet = 40;
int32_t launch_wire_pack(cudaStream_t st, int32_t n, const PostBuffers& pb, const WireMeta* meta, uint8_t* wire, uint32_t wire_cap, uint32_t* wire_off);

// ---------------------------------------------------------------- TMA helper
// 3-D map {cin, cout_pad, taps} with strides {ktot*2, cin*2} bytes and box {kc, nt, taps}: lands as [taps][nt][kc] (conv_halo.cu)
int32_t make_tmap_w3d(CUtensorMap* map, const void* w, int cin, int cout_pad, int taps, int ktot, int kc, int nt, bool f16);
int32_t make_tmap_2d_16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer,
                        uint64_t outer_stride_bytes, uint32_t box_inner, uint32_t box_outer, int32_t swizzle_bytes, bool f16);

}  // namespace zl
