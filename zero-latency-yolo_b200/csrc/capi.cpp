// capi.cpp — the extern "C" surface declared in include/zl_b200.h.
// No exception may cross this boundary (the reference wraps every engine call in
// try/catch and returns Result::error, onnx_engine.cpp:165-169,621-645).
#include <cstdlib>
#include <cstring>

#include <cuda_fp16.h>
#include <fstream>
#include <new>
#include <vector>

#include "engine.h"

struct zl_engine {
    zl::Engine* impl;
};

using zl::Engine;

#define ZL_GUARD_BEGIN try {
#define ZL_GUARD_END                                                        \
    } catch (const std::bad_alloc&) {                                       \
        zl::set_error("out of host memory");                                \
        return ZL_INSUFFICIENT_RESOURCES;                                   \
    } catch (const std::exception& ex) {                                    \
        zl::set_error(std::string("exception: ") + ex.what());              \
        return ZL_UNKNOWN_ERROR;                                            \
    } catch (...) {                                                         \
        zl::set_error("unknown exception");                                 \
        return ZL_UNKNOWN_ERROR;                                            \
    }

#define ZL_CHECK_ENGINE(e)                                                  \
    if (!(e) || !(e)->impl) {                                               \
        zl::set_error("null engine handle");                                \
        return ZL_INVALID_ARGUMENT;                                         \
    }

extern "C" {

void zl_config_default(zl_config* c)
{
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->device = 0;
    c->model_w = 416; c->model_h = 416;          // constants::DEFAULT_MODEL_WIDTH/HEIGHT (src/common/constants.h:26-27)
    c->num_classes = 4;                          // constants::cs16::CLASS_COUNT (:39)
    c->scale = ZL_SCALE_N;
    c->precision = ZL_PRECISION_BF16;
    c->conf_threshold = 0.5f;                    // DEFAULT_CONF_THRESHOLD (:28)
    c->iou_threshold = 0.45f;                    // DEFAULT_NMS_THRESHOLD (:29)
    c->class_weights = nullptr;
    c->max_batch = 1;
    c->max_frame_w = 0; c->max_frame_h = 0;
    c->preprocess_mode = ZL_PRE_STRETCH_NEAREST;
    c->queue_depth = 8;                          // INFERENCE_QUEUE_SIZE
    c->num_lanes = 1;
    c->use_graph = 1;
    c->batch_window_us = 0;
    c->emit_wire = 0;
    c->cpu_core_id = -1;
    c->high_priority = 0;
}

int32_t zl_engine_create(const zl_config* cfg, zl_engine** out)
{
    ZL_GUARD_BEGIN
    if (!cfg || !out) { zl::set_error("null argument"); return ZL_INVALID_ARGUMENT; }
    *out = nullptr;
    std::unique_ptr<Engine> e(new Engine(*cfg));
    int32_t rc = e->init();
    if (rc != ZL_OK) return rc;
    zl_engine* h = new zl_engine{e.release()};
    *out = h;
    return ZL_OK;
    ZL_GUARD_END
}

int32_t zl_engine_destroy(zl_engine* e)
{
    ZL_GUARD_BEGIN
    if (!e) return ZL_OK;
    delete e->impl;
    delete e;
    return ZL_OK;
    ZL_GUARD_END
}

int32_t zl_engine_load_weights_mem(zl_engine* e, const void* blob, size_t len)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    if (!blob) { zl::set_error("null blob"); return ZL_INVALID_ARGUMENT; }
    return e->impl->load_weights(blob, len);
    ZL_GUARD_END
}

int32_t zl_engine_load_weights(zl_engine* e, const char* path)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    if (!path) { zl::set_error("null path"); return ZL_INVALID_ARGUMENT; }
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { zl::set_error(std::string("Model file not found: ") + path); return ZL_MODEL_NOT_FOUND; }   // onnx_engine.cpp:961-966
    const std::streamsize n = f.tellg();
    f.seekg(0);
    std::vector<char> buf((size_t)n);
    if (!f.read(buf.data(), n)) { zl::set_error(std::string("cannot read ") + path); return ZL_MODEL_LOAD_FAILED; }
    return e->impl->load_weights(buf.data(), buf.size());
    ZL_GUARD_END
}

int32_t zl_engine_prepare_weights(zl_engine* e, const char* path)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    if (!path) { zl::set_error("null path"); return ZL_INVALID_ARGUMENT; }
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) { zl::set_error(std::string("Model file not found: ") + path); return ZL_MODEL_NOT_FOUND; }
    const std::streamsize n = f.tellg();
    f.seekg(0);
    std::vector<char> buf((size_t)n);
    if (!f.read(buf.data(), n)) { zl::set_error(std::string("cannot read ") + path); return ZL_MODEL_LOAD_FAILED; }
    return e->impl->prepare_weights(buf.data(), buf.size());
    ZL_GUARD_END
}

int32_t zl_engine_commit_weights(zl_engine* e)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->commit_weights();
    ZL_GUARD_END
}

int32_t zl_engine_discard_weights(zl_engine* e)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->discard_weights();
    ZL_GUARD_END
}

int32_t zl_engine_warmup(zl_engine* e, int32_t iters)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->warmup(iters);
    ZL_GUARD_END
}

int32_t zl_engine_set_callback(zl_engine* e, zl_result_fn fn, void* user)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    e->impl->cb = fn;
    e->impl->cb_user = user;
    return ZL_OK;
    ZL_GUARD_END
}

int32_t zl_engine_submit(zl_engine* e, uint32_t client_id, uint32_t frame_id, uint64_t timestamp,
                         int32_t width, int32_t height, const uint8_t* bgr, size_t len, int32_t /*is_keyframe*/)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->submit(client_id, frame_id, timestamp, width, height, bgr, len);
    ZL_GUARD_END
}

size_t zl_engine_queue_size(const zl_engine* e)
{
    try { return (e && e->impl) ? e->impl->queue_size() : 0; } catch (...) { return 0; }
}

int32_t zl_engine_drain(zl_engine* e)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->drain();
    ZL_GUARD_END
}

int32_t zl_engine_get_stats(const zl_engine* e, zl_stats* out)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    if (!out) { zl::set_error("null out"); return ZL_INVALID_ARGUMENT; }
    e->impl->get_stats(out);
    return ZL_OK;
    ZL_GUARD_END
}

int32_t zl_infer_batch(zl_engine* e, const uint8_t* const* frames, const int32_t* widths, const int32_t* heights, int32_t n,
                       zl_det* dets_out, int32_t det_capacity, int32_t* counts, int32_t* offsets)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->infer_batch(frames, widths, heights, n, dets_out, det_capacity, counts, offsets, nullptr);
    ZL_GUARD_END
}

int32_t zl_preprocess(zl_engine* e, const uint8_t* bgr, int32_t width, int32_t height, size_t len, float* out_chw)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->preprocess_one(bgr, width, height, len, out_chw);
    ZL_GUARD_END
}

int32_t zl_forward_raw(zl_engine* e, const uint8_t* const* frames, const int32_t* widths, const int32_t* heights, int32_t n, float* raw_out)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    if (!raw_out) { zl::set_error("null raw_out"); return ZL_INVALID_ARGUMENT; }
    return e->impl->infer_batch(frames, widths, heights, n, nullptr, 0, nullptr, nullptr, raw_out);
    ZL_GUARD_END
}

int32_t zl_decode_nms(zl_engine* e, const float* raw, int32_t n, int32_t nc, int32_t A, const int32_t* img_w, const int32_t* img_h,
                      float conf_thr, float iou_thr, zl_det* dets_out, int32_t det_capacity, int32_t* counts, int32_t* offsets)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    if (!dets_out) { zl::set_error("null dets_out"); return ZL_INVALID_ARGUMENT; }
    return e->impl->decode_nms(raw, n, nc, A, img_w, img_h, conf_thr, iou_thr, dets_out, det_capacity, counts, offsets, 1, nullptr, nullptr, nullptr);
    ZL_GUARD_END
}

int32_t zl_engine_num_anchors(const zl_engine* e) { return (e && e->impl) ? e->impl->num_anchors : 0; }

int32_t zl_engine_upload_resident(zl_engine* e, int32_t set, const uint8_t* const* frames, const int32_t* widths, const int32_t* heights, int32_t n)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->upload_resident(set, frames, widths, heights, n);
    ZL_GUARD_END
}

int32_t zl_engine_run_resident(zl_engine* e, int32_t n_sets, int32_t steps, float* total_ms, int64_t* launches, int64_t* total_dets)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->run_resident(n_sets, steps, total_ms, launches, total_dets);
    ZL_GUARD_END
}

int32_t zl_engine_profile(zl_engine* e, int32_t set, int32_t iters, zl_op_profile* out, int32_t cap, int32_t* n_out)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->profile(set, iters, out, cap, n_out);
    ZL_GUARD_END
}

int32_t zl_engine_set_wire_callback(zl_engine* e, zl_wire_fn fn, void* user)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    if (fn && !e->impl->cfg.emit_wire) { zl::set_error("engine was created without emit_wire"); return ZL_INVALID_ARGUMENT; }
    e->impl->wire_cb = fn; e->impl->wire_user = user;
    return ZL_OK;
    ZL_GUARD_END
}

int32_t zl_infer_batch_wire(zl_engine* e, const uint8_t* const* frames, const int32_t* widths, const int32_t* heights, int32_t n,
                            const uint32_t* frame_ids, const uint64_t* timestamps, uint64_t det_timestamp_ms,
                            uint8_t* out, size_t out_capacity, uint32_t* offsets)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->infer_batch_wire(frames, widths, heights, n, frame_ids, timestamps, det_timestamp_ms, out, out_capacity, offsets);
    ZL_GUARD_END
}

int32_t zl_bench_e2e(zl_engine* e, const uint8_t* const* batches, int32_t threads, int32_t n, int32_t width, int32_t height,
                     int32_t steps_total, double* seconds, int64_t* dets_last_step)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->bench_e2e(batches, threads, n, width, height, steps_total, seconds, dets_last_step);
    ZL_GUARD_END
}

int32_t zl_bench_h2d(zl_engine* e, size_t bytes, int32_t iters, double* gbs)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->bench_h2d(bytes, iters, gbs);
    ZL_GUARD_END
}

int32_t zl_engine_profile_stalls(zl_engine* e, int32_t set, uint64_t* out, int32_t cap_ops, int32_t* n_out)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->profile_stalls(set, out, cap_ops, n_out);
    ZL_GUARD_END
}

int32_t zl_bench_latency(zl_engine* e, const uint8_t* bgr, int32_t width, int32_t height, int32_t warmup, int32_t iters, float* ms_out)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->bench_latency(bgr, width, height, warmup, iters, ms_out);
    ZL_GUARD_END
}

int32_t zl_bench_preprocess(zl_engine* e, int32_t width, int32_t height, int32_t n, int32_t iters, float* ms_per_launch, double* bytes_per_launch)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    return e->impl->bench_preprocess(width, height, n, iters, ms_per_launch, bytes_per_launch);
    ZL_GUARD_END
}

int32_t zl_bench_decode_nms(zl_engine* e, const float* raw, int32_t n, int32_t nc, int32_t A, float conf_thr, float iou_thr,
                            int32_t iters, float* ms_filter, float* ms_nms, int64_t* kept)
{
    ZL_GUARD_BEGIN
    ZL_CHECK_ENGINE(e)
    std::vector<int32_t> iw(n, 640), ih(n, 640);
    return e->impl->decode_nms(raw, n, nc, A, iw.data(), ih.data(), conf_thr, iou_thr, nullptr, 0, nullptr, nullptr, iters + 1, ms_filter, ms_nms, kept);
    ZL_GUARD_END
}

// Host-only model probe (no CUDA): container kind agnostic summary, used by tests and by operators to check a file.
int32_t zl_model_probe(const void* blob, size_t len, int32_t* scale, int32_t* num_classes, int32_t* n_tensors, uint64_t* checksum)
{
    ZL_GUARD_BEGIN
    zl::ParsedModel pm;
    int32_t rc = zl::parse_model(blob, len, &pm);
    if (rc != ZL_OK) return rc;
    if (scale) *scale = pm.scale;
    if (num_classes) *num_classes = pm.nc;
    if (n_tensors) *n_tensors = (int32_t)pm.tensors.size();
    if (checksum) *checksum = zl::model_checksum(pm);
    return ZL_OK;
    ZL_GUARD_END
}

void* zl_host_alloc(size_t bytes)
{
    void* p = nullptr;
    // Frame staging memory is written by the CPU and read only by the copy engine.  ZL_PINNED_WC=1 makes it write-combined
    // (no cache snooping on the DMA reads) — an A/B switch for hosts whose pinned-copy bandwidth, not the GPU, bounds e2e.
    static const bool wc = [] { const char* e = getenv("ZL_PINNED_WC"); return e && e[0] == '1'; }();
    if (cudaHostAlloc(&p, bytes, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); zl::set_error("cudaHostAlloc failed"); return nullptr; }
    zl::register_pinned_range(p, bytes);
    return p;
}
void zl_host_free(void* p) { if (p) { zl::unregister_pinned_range(p); cudaFreeHost(p); } }
const char* zl_last_error(void) { return zl::get_error(); }
const char* zl_version(void) { return "zl_b200 0.1 (sm_100a)"; }
int32_t zl_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---- unit-test hook: one convolution, host tensors in/out ----

}  // extern "C"
