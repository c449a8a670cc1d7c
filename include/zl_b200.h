/*
 * zl_b200.h — C-ABI of the B200-native YOLOv8 detector (libzl_b200.so).
 *
 * This is the seam under the reference's C++ detector interface.  Every entry
 * point names the reference interface it replaces; paths are relative to the
 * reference tree (yynps737/zero-latency-yolo):
 *
 *   IInferenceEngine                 src/inference/inference_engine.h:33-43
 *   InferenceRequest                 src/inference/inference_engine.h:16-29
 *   OnnxInferenceEngine::runInference   src/inference/onnx_engine.cpp:518-646
 *   OnnxInferenceEngine::preProcess     src/inference/onnx_engine.cpp:649-700
 *   Ort::Session::Run                   src/inference/onnx_engine.cpp:577-585
 *   OnnxInferenceEngine::postProcess    src/inference/onnx_engine.cpp:758-834
 *   OnnxInferenceEngine::applyNMS       src/inference/onnx_engine.cpp:837-878
 *   OnnxInferenceEngine::calculateIoU   src/inference/onnx_engine.cpp:881-909
 *   ErrorCode                           src/common/result.h:14-48
 *   Detection / BoundingBox             src/common/types.h:16-26
 *
 * Plain C: opaque handle, plain pointers and sizes, int32 status codes that
 * carry the reference's ErrorCode numeric values.  No exceptions cross this
 * boundary, nothing aborts.  There is no CPU fallback: without a CUDA device
 * zl_engine_create() fails with ZL_INSUFFICIENT_RESOURCES.
 */
#ifndef ZL_B200_H_
#define ZL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define ZL_API
#else
#define ZL_API __attribute__((visibility("default")))
#endif

/* ---- status codes: numeric values of zero_latency::ErrorCode (src/common/result.h:14-48) */
enum {
    ZL_OK = 0,
    ZL_UNKNOWN_ERROR = 1,
    ZL_INVALID_ARGUMENT = 2,
    ZL_NOT_INITIALIZED = 3,      /* submit before warm-up / after shutdown (onnx_engine.cpp:224-226) */
    ZL_TIMEOUT = 4,
    ZL_INFERENCE_ERROR = 200,    /* also "queue full, frame dropped" (network_server.cpp:213-215) */
    ZL_MODEL_NOT_FOUND = 201,
    ZL_MODEL_LOAD_FAILED = 202,
    ZL_INVALID_INPUT = 203,      /* frame length != w*h*3 (onnx_engine.cpp:659-665) */
    ZL_SYSTEM_ERROR = 300,
    ZL_INSUFFICIENT_RESOURCES = 303
};

/* ---- enums ---- */
enum { ZL_SCALE_N = 0, ZL_SCALE_S = 1, ZL_SCALE_M = 2 };          /* YOLOv8 n / s / m */
enum { ZL_PRECISION_FP32 = 0, ZL_PRECISION_BF16 = 1, ZL_PRECISION_FP16 = 2 };   /* fp32 = exact CUDA-core mode; bf16 / fp16 = tcgen05, fp32 accumulate */
enum { ZL_PRE_STRETCH_NEAREST = 0, ZL_PRE_LETTERBOX = 1 };        /* 0 = the reference's behaviour (parity mode) */

/* One detection as the device emits it: the first 24 bytes of the reference's
 * 40-byte Detection (src/common/types.h:20-26); track_id=0 and timestamp are
 * filled by the host adapter (onnx_engine.cpp:812-815).  Box is centre-x,
 * centre-y, width, height divided by the REQUEST frame's width/height
 * (onnx_engine.cpp:802-805). */
typedef struct zl_det {
    float x, y, w, h;
    float confidence;
    int32_t class_id;
} zl_det;

/* Engine configuration.  Mirrors the ServerConfig fields the path consumes
 * (src/server/config.h:305-345, :110-128) plus the B200-only knobs. */
typedef struct zl_config {
    int32_t device;             /* CUDA device ordinal */
    int32_t model_w, model_h;   /* detection.model_width/height; multiples of 32 */
    int32_t num_classes;        /* nc: output0 is [B, 4+nc, A] (onnx_engine.cpp:773) */
    int32_t scale;              /* ZL_SCALE_* */
    int32_t precision;          /* ZL_PRECISION_* */
    float conf_threshold;       /* confidence_threshold, kept on >= (onnx_engine.cpp:799) */
    float iou_threshold;        /* nms_threshold, suppressed on > (onnx_engine.cpp:871) */
    const float* class_weights; /* nc floats or NULL (= all 1.0, the reference's effective behaviour) */
    int32_t max_batch;          /* frames per launch, 1..256 */
    int32_t max_frame_w, max_frame_h; /* largest request frame accepted (staging size) */
    int32_t preprocess_mode;    /* ZL_PRE_* */
    int32_t queue_depth;        /* bound of the async queue; full queue -> ZL_INFERENCE_ERROR */
    int32_t num_lanes;          /* concurrent pipelines (streams + buffers) for the async path, 1..8 */
    int32_t use_graph;          /* 1 = capture each batch size in a CUDA graph */
    int32_t batch_window_us;    /* async path: wait this long to coalesce a batch (0 = take what is queued) */
    int32_t emit_wire;          /* 1 = also emit every frame's result in the reference's wire layout on the device (zl_*_wire entry points) */
    int32_t cpu_core_id;        /* ServerConfig::use_cpu_affinity / cpu_core_id: worker i is pinned to core cpu_core_id + i; -1 = not pinned */
    int32_t high_priority;      /* ServerConfig::use_high_priority: raise the workers' scheduling priority (best effort) */
    int32_t reserved[5];
} zl_config;

typedef struct zl_stats {
    uint64_t inference_count;
    uint64_t inference_errors;
    uint64_t dropped_frames;
    uint64_t queue_size;
    uint64_t queue_high_water_mark;
    uint64_t batches;
    double avg_inference_time_ms;      /* submit -> callback, mean of last 1000 */
    double p99_inference_time_ms;
    double avg_preprocessing_time_ms;  /* device time of P1 per batch (profile mode only, else 0) */
    double avg_postprocessing_time_ms;
    double avg_device_time_ms;         /* device time per batch (CUDA events around a SAMPLE of the batches: every 256th of a lane) */
    int32_t graph_captured;
    int32_t device;
    int32_t precision;
    int32_t running;
} zl_stats;

/* Per-kernel timing record returned by zl_engine_profile(). */
typedef struct zl_op_profile {
    char name[48];
    int32_t kind;          /* 0 pre, 1 conv(tcgen05), 2 conv(fp32 simt), 3 conv0 direct, 4 pool, 5 upsample, 6 decode, 7 filter, 8 nms */
    int32_t launches;
    float ms;              /* mean device time per launch */
    double flops;          /* algorithmic FLOPs per launch (2*MAC) */
    double bytes;          /* algorithmic bytes per launch (each operand once) */
} zl_op_profile;

typedef struct zl_engine zl_engine;

/* Result callback of the async path — replaces InferenceCallback
 * (inference_engine.h:31).  Fired from an engine-owned thread, once for EVERY
 * accepted frame, in submission order per engine.  `dets` is valid only for
 * the duration of the call.  status != ZL_OK means the frame failed. */
typedef void (*zl_result_fn)(void* user, uint32_t client_id, uint32_t frame_id,
                             uint64_t timestamp, int32_t status,
                             const zl_det* dets, int32_t n);

/* Result wire layout.  The body of the reference's DetectionResultPacket
 * (src/common/protocol.h:541-567) for one frame, byte for byte:
 *     uint32 frame_id | uint64 timestamp | uint16 count | count x Detection
 * with Detection the 40-byte record of src/common/types.h:20-26 ({x,y,w,h,confidence f32; class_id i32; track_id u32 = 0;
 * 4 padding bytes = 0; timestamp u64 = the batch's wall-clock ms, onnx_engine.cpp:812-815).  Packed, unaligned (14-byte
 * header).  The device writes these blocks itself, frame after frame in batch order, and they come back in ONE copy:
 * the adapter hands the bytes through (or memcpy's the records into GameState::detections) without touching a detection. */
#define ZL_WIRE_HEADER_BYTES 14
#define ZL_WIRE_DET_BYTES 40
typedef void (*zl_wire_fn)(void* user, uint32_t client_id, uint32_t frame_id, uint64_t timestamp, int32_t status,
                           const uint8_t* packet_body, size_t body_bytes);

/* ---- lifecycle (IInferenceEngine::initialize / shutdown, onnx_engine.cpp:67-221) ---- */
ZL_API void    zl_config_default(zl_config* cfg);
ZL_API int32_t zl_engine_create(const zl_config* cfg, zl_engine** out);
ZL_API int32_t zl_engine_destroy(zl_engine* e);
/* loadModel (onnx_engine.cpp:957-1062): a ZLW1 weights container or an ultralytics YOLOv8 ONNX export
 * (the file the reference loads, start.sh:122-125; BN fused, fp32 or fp16 initialisers).  Callable while serving:
 * the swap is atomic (hot reload, onnx_engine.cpp:473-515). */
ZL_API int32_t zl_engine_load_weights(zl_engine* e, const char* path);
ZL_API int32_t zl_engine_load_weights_mem(zl_engine* e, const void* blob, size_t len);
/* The two halves of a load, for hosts that serve several devices: prepare builds the new weight set on the device next
 * to the live one (serving continues), commit swaps atomically, discard drops a prepared set.  Reload that is atomic
 * ACROSS devices = prepare on every engine, then commit on all of them or discard on all of them. */
ZL_API int32_t zl_engine_prepare_weights(zl_engine* e, const char* path);
ZL_API int32_t zl_engine_commit_weights(zl_engine* e);
ZL_API int32_t zl_engine_discard_weights(zl_engine* e);
/* warmupModel (onnx_engine.cpp:919-954): `iters` runs on an all-128 frame of model size; captures graphs. */
ZL_API int32_t zl_engine_warmup(zl_engine* e, int32_t iters);

/* ---- async path (IInferenceEngine::setCallback / submitInference / getQueueSize) ---- */
ZL_API int32_t zl_engine_set_callback(zl_engine* e, zl_result_fn fn, void* user);
/* Copies the frame before returning (the reference copies too, onnx_engine.cpp:235). Non-blocking. */
ZL_API int32_t zl_engine_submit(zl_engine* e, uint32_t client_id, uint32_t frame_id,
                                uint64_t timestamp, int32_t width, int32_t height,
                                const uint8_t* bgr, size_t len, int32_t is_keyframe);
/* Async results as wire blocks (needs cfg.emit_wire = 1); replaces the zl_result_fn callback when set. */
ZL_API int32_t zl_engine_set_wire_callback(zl_engine* e, zl_wire_fn fn, void* user);
ZL_API size_t  zl_engine_queue_size(const zl_engine* e);
/* Blocks until every accepted frame has had its callback. */
ZL_API int32_t zl_engine_drain(zl_engine* e);
ZL_API int32_t zl_engine_get_stats(const zl_engine* e, zl_stats* out);

/* ---- synchronous entry points (tests, bench, one-shot users) ---- */
/* runInference over n frames (onnx_engine.cpp:518-646).  frames[i] is a HOST
 * buffer of widths[i]*heights[i]*3 BGR bytes (pinned memory from
 * zl_host_alloc is copied without staging).  dets_out has room for
 * det_capacity records; frame i's detections are dets_out[offsets[i] ..
 * offsets[i]+counts[i]).  Returns ZL_INSUFFICIENT_RESOURCES if capacity is
 * too small (counts are still valid). */
ZL_API int32_t zl_infer_batch(zl_engine* e, const uint8_t* const* frames,
                              const int32_t* widths, const int32_t* heights, int32_t n,
                              zl_det* dets_out, int32_t det_capacity,
                              int32_t* counts, int32_t* offsets);
/* runInference over n frames with the results in the wire layout (needs cfg.emit_wire = 1): frame i's packet body is
 * out[offsets[i] .. offsets[i+1]); offsets has n+1 entries.  det_timestamp_ms is the value every Detection::timestamp
 * of the batch carries (the reference stamps wall-clock ms at post-processing time). */
ZL_API int32_t zl_infer_batch_wire(zl_engine* e, const uint8_t* const* frames, const int32_t* widths, const int32_t* heights, int32_t n,
                                   const uint32_t* frame_ids, const uint64_t* timestamps, uint64_t det_timestamp_ms,
                                   uint8_t* out, size_t out_capacity, uint32_t* offsets);
/* preProcess alone: out = [3, model_h, model_w] fp32 NCHW, exactly the tensor
 * the reference hands to ORT (onnx_engine.cpp:560-569). */
ZL_API int32_t zl_preprocess(zl_engine* e, const uint8_t* bgr, int32_t width, int32_t height,
                             size_t len, float* out_chw);
/* Session::Run alone: n frames -> raw head output [n, 4+nc, A] fp32 (host). */
ZL_API int32_t zl_forward_raw(zl_engine* e, const uint8_t* const* frames,
                              const int32_t* widths, const int32_t* heights, int32_t n,
                              float* raw_out);
/* postProcess+applyNMS alone on a HOST raw head tensor [n, 4+nc, A] (nc and A
 * need not match the engine's model: this is the decode/NMS stress entry). */
ZL_API int32_t zl_decode_nms(zl_engine* e, const float* raw, int32_t n, int32_t nc, int32_t A,
                             const int32_t* img_w, const int32_t* img_h,
                             float conf_thr, float iou_thr,
                             zl_det* dets_out, int32_t det_capacity,
                             int32_t* counts, int32_t* offsets);
ZL_API int32_t zl_engine_num_anchors(const zl_engine* e);

/* ---- measurement helpers (bench.py) ---- */
/* Upload n frames once into resident input set `set` (0..3). */
ZL_API int32_t zl_engine_upload_resident(zl_engine* e, int32_t set, const uint8_t* const* frames,
                                         const int32_t* widths, const int32_t* heights, int32_t n);
/* Run `steps` passes of the whole path over resident sets (cycled), timed with
 * CUDA events on the engine's stream; returns total ms and kernel launches. */
ZL_API int32_t zl_engine_run_resident(zl_engine* e, int32_t n_sets, int32_t steps,
                                      float* total_ms, int64_t* launches, int64_t* total_dets);
/* Same pass, un-captured, one CUDA-event pair per kernel; fills up to cap records. */
ZL_API int32_t zl_engine_profile(zl_engine* e, int32_t set, int32_t iters,
                                 zl_op_profile* out, int32_t cap, int32_t* n_out);
/* Where the persistent conv kernel's cycles go: one un-captured pass with its instrumented instantiation.  out holds
 * ZL_STALL_SLOTS uint64 per op, in zl_engine_profile's op order (all zero for ops that are not persistent convs):
 * 0 producer waiting for a free patch stage, 1 MMA warp waiting for a patch, 2 MMA warp waiting for a drained
 * accumulator, 3 MMA warp issuing, 4 epilogue warp waiting for an accumulator, 5 epilogue warp busy, 6 CTA lifetime
 * (sum over CTAs), 7 prologue (sum), 8 MMA warp waiting for the weights, 9 slowest CTA, 10 CTAs, 11 launch plan (packed),
 * 12-15 epilogue warp detail: tcgen05.ld wait, staging-block wait, st.shared + fence + bulk-store issue, accumulator hand-back. */
#define ZL_STALL_SLOTS 16
ZL_API int32_t zl_engine_profile_stalls(zl_engine* e, int32_t set, uint64_t* out, int32_t cap_ops, int32_t* n_out);
/* b=1 latency loop in C (no interpreter in the timed path): `iters` synchronous runInference calls on one
 * HOST frame (pinned or not), each timed with steady_clock from call to detections-on-host; ms_out[iters]. */
ZL_API int32_t zl_bench_latency(zl_engine* e, const uint8_t* bgr, int32_t width, int32_t height,
                                int32_t warmup, int32_t iters, float* ms_out);
/* End-to-end throughput loop in C: `threads` host threads (one engine lane each), thread t owning the batch of n
 * equal-size frames stored back to back at batches[t] (pinned host memory), call zl_infer_batch steps_total times in
 * all: every step copies its frames host->device, runs the whole path and reads the detections back.  seconds = wall
 * time until the slowest thread is done. */
ZL_API int32_t zl_bench_e2e(zl_engine* e, const uint8_t* const* batches, int32_t threads, int32_t n, int32_t width, int32_t height,
                            int32_t steps_total, double* seconds, int64_t* dets_last_step);
/* Pinned host -> device copy bandwidth on the engine's device (GB/s): the ceiling of the end-to-end leg. */
ZL_API int32_t zl_bench_h2d(zl_engine* e, size_t bytes, int32_t iters, double* gbs);
/* Stand-alone kernels with device-resident synthetic data, for roofline lines. */
ZL_API int32_t zl_bench_preprocess(zl_engine* e, int32_t width, int32_t height, int32_t n,
                                   int32_t iters, float* ms_per_launch, double* bytes_per_launch);
ZL_API int32_t zl_bench_decode_nms(zl_engine* e, const float* raw, int32_t n, int32_t nc, int32_t A,
                                   float conf_thr, float iou_thr, int32_t iters,
                                   float* ms_filter, float* ms_nms, int64_t* kept);

/* Host-only (no CUDA) probe of a model file: which YOLOv8 scale / class count it holds, how many conv tensors, and
 * an FNV-1a checksum over their names, shapes and values (equal for ZLW1 and ONNX files with the same parameters). */
ZL_API int32_t zl_model_probe(const void* blob, size_t len, int32_t* scale, int32_t* num_classes, int32_t* n_tensors, uint64_t* checksum);

/* ---- host memory + errors ---- */
ZL_API void*   zl_host_alloc(size_t bytes);   /* pinned */
ZL_API void    zl_host_free(void* p);
ZL_API const char* zl_last_error(void);       /* thread-local message of the last failing call */
ZL_API const char* zl_version(void);
ZL_API int32_t zl_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ZL_B200_H_ */
