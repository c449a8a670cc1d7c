// engine.h — host-side engine behind the C-ABI (include/zl_b200.h).
//
// One Engine == one CUDA device.  It owns the folded weights, `num_lanes`
// independent pipelines (stream + activation buffers + CUDA graphs) and the
// bounded request queue with its worker threads.  The reference's equivalent
// is OnnxInferenceEngine (src/inference/onnx_engine.{h,cpp}); what differs by
// design is listed in DESIGN.md (true batching, a callback for every frame,
// no simulation mode, no CPU fallback).
#pragma once
#include <atomic>
#include <condition_variable>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "kernels.h"
#include "weights.h"

namespace zl {

struct ModelDef {
    int scale = 0, nc = 0;
    int c[5] = {0, 0, 0, 0, 0};       // widths of the five stages
    int n[4] = {1, 2, 2, 1};          // backbone C2f repeats
    int nh = 1;                       // head C2f repeats
    int cb = 64, cc = 64;             // Detect box / class branch widths (as in the model file)
    int ccd = 64;                     // class branch width ON THE DEVICE: cc, padded to a multiple of 16 in the 16-bit modes (zero channels)
};

struct Op {
    enum Kind : int32_t { PRE = 0, CONV_TC = 1, CONV_SIMT = 2, CONV0 = 3, POOL = 4, UPSAMPLE = 5, DECODE = 6, FILTER = 7, NMS = 8, CONV_HALO = 9, PRE_CONV0 = 10, DECODE_FILTER = 11, HEAD_FUSED = 12 };
    int32_t kind = PRE;
    std::string name;
    const ConvWeights* w = nullptr;
    ConvTcOp tc;
    ConvHaloOp halo;
    HeadFusedOp hf;
    // which passes run this op: 0 all; 1 raw-head passes only (zl_forward_raw); 2 hot path when the head is NOT fused;
    // 3 raw passes and the unfused hot path (the last 1x1 convs of the head); 4 hot path with the fused head kernel
    int32_t path = 0;
    // head branches (small batches: the Detect head of a level runs on side streams next to the rest of the neck):
    // side = 0 main stream, 1..6 side stream; wait_ev / record_ev index Lane::ev_dep (-1 none); join = wait for every side stream first
    int32_t side = 0, wait_ev = -1, record_ev = -1, join = 0;
    View x, y, res, p1, p2, p3;
    bool has_res = false;
    double flops = 0, bytes = 0;
};

struct Request {
    uint32_t client_id, frame_id;
    uint64_t timestamp;
    int32_t w, h;
    int32_t slot;                 // pinned slot holding the frame bytes
    std::chrono::steady_clock::time_point t_submit;
};

struct WireReq { const uint32_t* frame_ids; const uint64_t* timestamps; uint64_t det_ts; };   // per-frame header fields + the batch's Detection::timestamp

struct Lane {
    int id = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;                              // one batch at a time per lane
    // device memory
    char* arena = nullptr; size_t arena_bytes = 0;
    uint8_t* staging = nullptr;                 // [(1 + resident sets) * max_batch] frame slots
    size_t staging_slots = 0;
    FrameDesc* d_descs = nullptr;               // [max_batch]
    FrameDesc* d_res_descs = nullptr;           // [4][max_batch] descriptors of the resident sets
    float* raw = nullptr;                       // [max_batch][4+nc][A]
    PostBuffers pb{};
    std::map<std::string, View> bufs;
    HeadLevel levels[3];
    // pinned host memory
    FrameDesc* h_descs = nullptr;
    uint32_t* nms_host_out = nullptr;            // set by run_ops: the NMS of a one-frame step writes the result block here itself
    uint64_t batches_run = 0;                   // batches this lane has launched (the device-time statistic samples every 256th)
    std::vector<FrameDesc> descs_on_device;     // what d_descs holds (empty = unknown): a batch with the same geometry as the last one skips the copy
    uint8_t* h_result = nullptr; size_t h_result_bytes = 0;
    uint8_t* h_frames = nullptr;                // staging for non-pinned sync inputs: [max_batch] slots
    // N3 (cfg.emit_wire): result wire blocks written by the device
    uint8_t* d_wire = nullptr; uint32_t* d_wire_off = nullptr; WireMeta* d_wmeta = nullptr;
    uint8_t* h_wire = nullptr; uint32_t* h_wire_off = nullptr; WireMeta* h_wmeta = nullptr;
    size_t wire_cap = 0;
    // per batch size
    std::map<int, std::vector<Op>> ops;
    std::map<int, cudaGraphExec_t> graphs;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t side[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // head branches: level l -> side[2l] (stem, box branch), side[2l+1] (class branch)
    cudaEvent_t ev_dep[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // 0..2: level l's input map is ready (main stream); 3..5: level l's stem is done
    cudaEvent_t ev_join[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int resident_n[4] = {0, 0, 0, 0};
    bool resident_same[4] = {false, false, false, false};
    bool same_size = false;                     // the batch being launched: every frame already has the model's size (fast preprocess kernel)
};

class Engine {
public:
    explicit Engine(const zl_config& cfg);
    ~Engine();
    int32_t init();
    int32_t load_weights(const void* blob, size_t len);        // prepare + commit
    int32_t prepare_weights(const void* blob, size_t len);     // parse, validate, upload next to the live set (serving continues)
    int32_t commit_weights();                                  // atomic swap under the lanes' locks
    int32_t discard_weights();
    int32_t warmup(int iters);
    int32_t start_workers();
    void stop_workers();

    int32_t submit(uint32_t client_id, uint32_t frame_id, uint64_t ts, int w, int h, const uint8_t* bgr, size_t len);
    int32_t drain();
    size_t queue_size() const;
    void get_stats(zl_stats* out) const;

    int32_t infer_batch(const uint8_t* const* frames, const int32_t* ws, const int32_t* hs, int n,
                        zl_det* out, int cap, int32_t* counts, int32_t* offsets, float* raw_out);
    int32_t infer_batch_wire(const uint8_t* const* frames, const int32_t* ws, const int32_t* hs, int n, const uint32_t* frame_ids,
                             const uint64_t* timestamps, uint64_t det_ts, uint8_t* out, size_t cap, uint32_t* offsets);
    int32_t preprocess_one(const uint8_t* bgr, int w, int h, size_t len, float* out_chw);
    int32_t decode_nms(const float* raw, int n, int nc, int A, const int32_t* iw, const int32_t* ih, float conf, float iou,
                       zl_det* out, int cap, int32_t* counts, int32_t* offsets,
                       int iters, float* ms_filter, float* ms_nms, int64_t* kept);
    int32_t upload_resident(int set, const uint8_t* const* frames, const int32_t* ws, const int32_t* hs, int n);
    int32_t run_resident(int n_sets, int steps, float* total_ms, int64_t* launches, int64_t* total_dets);
    int32_t profile(int set, int iters, zl_op_profile* out, int cap, int32_t* n_out);
    int32_t profile_stalls(int set, uint64_t* out, int cap_ops, int32_t* n_out);
    int32_t bench_preprocess(int w, int h, int n, int iters, float* ms, double* bytes);
    int32_t bench_latency(const uint8_t* bgr, int w, int h, int warm, int iters, float* ms_out);
    int32_t bench_e2e(const uint8_t* const* batches, int threads, int n, int w, int h, int steps_total, double* seconds, int64_t* dets_last);
    int32_t bench_h2d(size_t bytes, int iters, double* gbs);

    zl_config cfg;
    zl_result_fn cb = nullptr;
    void* cb_user = nullptr;
    zl_wire_fn wire_cb = nullptr;   // replaces cb when set (needs cfg.emit_wire)
    void* wire_user = nullptr;
    int num_anchors = 0;
    int num_sms = 148;
    bool use_halo = true;
    int persist_min_units = 0;     // a layer goes to the persistent kernel when it has at least this many (tile, N-slice) work units ...
    bool deep_k_persist = false;   // ... or (experiment) when its K loop is deep
    bool fuse_stems = true;    // Detect stems cv2.l.0 + cv3.l.0 as one conv of width cb + cc (16-bit modes)
    bool use_stem = true;      // layer 0 on tensor cores (2x2 conv over the space-to-depth image the preprocess kernel writes)
    bool fuse_pre = false;     // P1 fused into layer 0 (bit-identical, measured slower than the two-kernel path: scattered byte loads)
    bool weights_loaded = false;

private:
    int32_t build_model_def();
    int32_t alloc_lane(Lane& L);
    void free_lane(Lane& L);
    int32_t build_ops(Lane& L, int B);
    int32_t run_ops(Lane& L, int B, bool with_d2h, bool want_raw = false);
    int32_t launch_op(Lane& L, int B, const Op& op, cudaStream_t on = nullptr);
    int32_t ensure_graph(Lane& L, int B);
    int32_t launch_batch(Lane& L, int B, bool want_raw = false);   // graph if enabled, else direct
    int graph_batch_for(int n) const;
    int32_t run_lane_batch(Lane& L, const uint8_t* const* frames, const int32_t* ws, const int32_t* hs, int n,
                           bool frames_pinned, std::vector<zl_det>* dets, int32_t* counts, bool want_raw = false, const WireReq* wr = nullptr);
    void worker_main(int lane_id);
    size_t slot_bytes() const { return (size_t)cfg.max_frame_w * cfg.max_frame_h * 3; }
    int inline_dets(int B) const;
    size_t wire_inline_bytes(int B) const;

    ModelDef md;
    std::mutex load_mu;                            // one load_weights at a time
    std::vector<std::unique_ptr<ConvWeights>> convs;
    std::map<std::string, ConvWeights*> conv_by_name;
    std::vector<std::unique_ptr<ConvWeights>> pending_convs;   // prepared, not yet committed
    std::map<std::string, ConvWeights*> pending_by_name;
    float* d_class_weights = nullptr;
    std::vector<std::unique_ptr<Lane>> lanes;

    // async path
    mutable std::mutex qmu;
    std::condition_variable qcv, done_cv, order_cv;
    std::deque<Request> queue;
    std::vector<int> free_slots;
    uint8_t* h_slots = nullptr;                    // pinned: queue_depth frame slots
    std::vector<std::thread> workers;
    std::atomic<bool> running{false};
    bool stopping = false;
    uint64_t next_seq = 0, deliver_seq = 0;
    uint64_t in_flight = 0;

    // stats
    mutable std::mutex smu;
    uint64_t st_count = 0, st_errors = 0, st_dropped = 0, st_hwm = 0, st_batches = 0;
    std::deque<double> lat_ms;
    double dev_ms_sum = 0; uint64_t dev_ms_n = 0;
    int graph_captured = 0;
    std::atomic<uint32_t> sync_rr{0};
};

void register_pinned_range(const void* p, size_t bytes);     // zl_host_alloc / zl_host_free keep the pinned-range table exact
void unregister_pinned_range(const void* p);

}  // namespace zl
