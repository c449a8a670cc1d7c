// b200_engine.cpp — see b200_engine.h.  No exception leaves this class (the reference wraps every
// engine entry point in try/catch -> Result::error, onnx_engine.cpp:165-169,621-645).
#include "b200_engine.h"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <fstream>

namespace zero_latency {

namespace {
int precisionOf(const std::string& s) { return s == "fp32" ? ZL_PRECISION_FP32 : (s == "bf16" ? ZL_PRECISION_BF16 : ZL_PRECISION_FP16); }
int scaleOf(const std::string& s) { return s == "s" ? ZL_SCALE_S : (s == "m" ? ZL_SCALE_M : ZL_SCALE_N); }
std::string lastError() { const char* e = zl_last_error(); return e ? std::string(e) : std::string(); }

// SHA-256 of the weights file: the status map's "model_hash" and the hot-reload watcher's change detector, like the
// reference's calculateModelHash (onnx_engine.cpp:1087-1124, which uses OpenSSL).  FIPS 180-4, written out here so the
// adapter needs no libcrypto.
struct Sha256 {
    uint32_t h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    uint8_t buf[64];
    size_t fill = 0;
    uint64_t total = 0;
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void block(const uint8_t* p) {
        static const uint32_t K[64] = {
            0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u,
            0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
            0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u,
            0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
            0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u,
            0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
        uint32_t w[64];
        for (int i = 0; i < 16; ++i) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; ++i) {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; ++i) {
            const uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const uint8_t* p, size_t n) {
        total += n;
        while (n) {
            const size_t k = std::min(n, sizeof(buf) - fill);
            std::memcpy(buf + fill, p, k);
            fill += k; p += k; n -= k;
            if (fill == 64) { block(buf); fill = 0; }
        }
    }
    std::string hex() {
        const uint64_t bits = total * 8;
        const uint8_t one = 0x80, zero = 0;
        update(&one, 1);
        while (fill != 56) update(&zero, 1);
        uint8_t len[8];
        for (int i = 0; i < 8; ++i) len[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(len, 8);
        char out[65];
        for (int i = 0; i < 8; ++i) snprintf(out + 8 * i, 9, "%08x", h[i]);
        return std::string(out, 64);
    }
};

std::string fileHash(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return "";
    Sha256 s;
    char buf[8192];                                   // the reference reads 8 KB blocks too (onnx_engine.cpp:1101)
    while (f.good()) {
        f.read(buf, sizeof(buf));
        if (f.gcount() > 0) s.update(reinterpret_cast<const uint8_t*>(buf), (size_t)f.gcount());
    }
    return s.hex();
}
}  // namespace

B200InferenceEngine::B200InferenceEngine(const ServerConfig& config) : config_(config) {}

B200InferenceEngine::~B200InferenceEngine() { shutdown(); }

Result<void> B200InferenceEngine::initialize() {
    try {
        if (running_) return Result<void>::ok();
        if (config_.b200.devices.empty()) return Result<void>::error(ErrorCode::INVALID_ARGUMENT, "b200.devices is empty");
        // No simulation mode and no CPU fallback: a missing model is an error here, where the reference
        // silently fabricates random boxes (onnx_engine.cpp:70-75,1133-1177).
        {
            std::ifstream probe(config_.model_path, std::ios::binary);
            if (!probe) return Result<void>::error(ErrorCode::MODEL_NOT_FOUND, "Model file not found: " + config_.model_path);
        }
        model_hash_ = fileHash(config_.model_path);
        for (int dev : config_.b200.devices) {
            zl_config c;
            zl_config_default(&c);
            c.device = dev;
            c.model_w = config_.detection.model_width;
            c.model_h = config_.detection.model_height;
            c.num_classes = config_.b200.num_classes;
            c.scale = scaleOf(config_.b200.scale);
            c.precision = precisionOf(config_.b200.precision);
            c.conf_threshold = config_.confidence_threshold;
            c.iou_threshold = config_.nms_threshold;
            c.class_weights = (config_.b200.use_class_weights && (int)config_.b200.class_weights.size() == c.num_classes)
                                  ? config_.b200.class_weights.data() : nullptr;
            c.max_batch = config_.b200.max_batch;
            c.max_frame_w = config_.b200.max_frame_width;
            c.max_frame_h = config_.b200.max_frame_height;
            c.queue_depth = (int32_t)std::max<size_t>(config_.max_queue_size, 1);
            c.num_lanes = config_.b200.num_lanes;
            c.batch_window_us = config_.b200.batch_window_us;
            c.emit_wire = config_.b200.wire_results ? 1 : 0;
            c.cpu_core_id = config_.use_cpu_affinity ? config_.cpu_core_id + (int)engines_.size() * std::max(config_.b200.num_lanes, 1) : -1;
            c.high_priority = config_.use_high_priority ? 1 : 0;
            zl_engine* e = nullptr;
            int32_t rc = zl_engine_create(&c, &e);
            if (rc == ZL_OK) rc = zl_engine_load_weights(e, config_.model_path.c_str());
            if (rc == ZL_OK) rc = config_.b200.wire_results ? zl_engine_set_wire_callback(e, &B200InferenceEngine::onWire, this)
                                                            : zl_engine_set_callback(e, &B200InferenceEngine::onResult, this);
            if (rc == ZL_OK) rc = zl_engine_warmup(e, 3);            // warmupModel: 3 runs (onnx_engine.cpp:939)
            if (rc != ZL_OK) {
                const std::string msg = "Failed to initialize B200 engine on device " + std::to_string(dev) + ": " + lastError();
                if (e) zl_engine_destroy(e);
                for (zl_engine* o : engines_) zl_engine_destroy(o);
                engines_.clear();
                return Result<void>::error(toErrorCode(rc), msg);
            }
            engines_.push_back(e);
        }
        running_ = true;
        if (config_.b200.use_model_monitor) monitor_thread_ = std::thread(&B200InferenceEngine::modelMonitorThreadFunc, this);
        Event event(events::SYSTEM_STARTUP);             // onnx_engine.cpp:160-162
        event.setSource("B200InferenceEngine");
        publishEvent(event);
        return Result<void>::ok();
    } catch (const std::exception& ex) {
        return Result<void>::error(ErrorCode::INFERENCE_ERROR, std::string("Failed to initialize B200 engine: ") + ex.what());
    }
}

Result<void> B200InferenceEngine::shutdown() {
    try {
        if (running_.exchange(false)) {
            { std::lock_guard<std::mutex> g(monitor_mu_); }
            monitor_cv_.notify_all();
            if (monitor_thread_.joinable()) monitor_thread_.join();
            for (zl_engine* e : engines_) { zl_engine_drain(e); zl_engine_destroy(e); }
            engines_.clear();
            Event event(events::SYSTEM_SHUTDOWN);        // onnx_engine.cpp:214-216
            event.setSource("B200InferenceEngine");
            publishEvent(event);
        }
        return Result<void>::ok();
    } catch (const std::exception& ex) {
        return Result<void>::error(ErrorCode::SYSTEM_ERROR, ex.what());
    }
}

Result<void> B200InferenceEngine::submitInference(const InferenceRequest& request) {
    if (!running_) return Result<void>::error(ErrorCode::NOT_INITIALIZED, "Inference engine not running");   // onnx_engine.cpp:224-226
    EventBus::getInstance().publishInferenceEvent(events::INFERENCE_REQUESTED, request.client_id, request.frame_id);   // onnx_engine.cpp:229-231
    zl_engine* e = engines_[request.client_id % engines_.size()];
    const int32_t rc = zl_engine_submit(e, request.client_id, request.frame_id, request.timestamp, request.width, request.height,
                                        request.data.data(), request.data.size(), request.is_keyframe ? 1 : 0);
    if (rc != ZL_OK) return Result<void>::error(toErrorCode(rc), lastError());
    return Result<void>::ok();
}

// Model hot reload (onnx_engine.cpp:473-515): re-hash the model file periodically; on change build the new weight set on
// EVERY device next to the live one (zl_engine_prepare_weights: serving continues), and only when all of them succeeded
// swap them all (zl_engine_commit_weights: atomic under the lanes' locks, frames in flight finish on the old weights) and
// recapture the CUDA graphs.  If any device fails, every prepared set is discarded: no device ever serves a model the
// others do not, and the failure is visible in getStatus ("model_reload_failures", "model_reload_error").
void B200InferenceEngine::modelMonitorThreadFunc() {
    std::string last_hash = model_hash_;
    while (running_) {
        {
            std::unique_lock<std::mutex> lk(monitor_mu_);
            monitor_cv_.wait_for(lk, std::chrono::milliseconds(std::max(config_.b200.model_check_interval_ms, 10)), [&] { return !running_.load(); });
        }
        if (!running_) break;
        const std::string h = fileHash(config_.model_path);
        if (h.empty() || h == last_hash) continue;              // missing file: keep serving (the reference only logs a warning)
        last_hash = h;                                          // do not retry a bad file every interval
        bool ok = true;
        std::string err;
        for (zl_engine* e : engines_)
            if (zl_engine_prepare_weights(e, config_.model_path.c_str()) != ZL_OK) { ok = false; err = lastError(); break; }
        if (ok) {
            for (zl_engine* e : engines_)
                if (zl_engine_commit_weights(e) != ZL_OK) { ok = false; err = "commit: " + lastError(); }   // cannot fail once prepared, short of a device fault
            for (zl_engine* e : engines_) zl_engine_warmup(e, 1);
        } else {
            for (zl_engine* e : engines_) zl_engine_discard_weights(e);
        }
        std::lock_guard<std::mutex> g(monitor_mu_);
        if (ok) {
            model_hash_ = h;
            model_version_++;
            reload_error_.clear();
            Event event("MODEL_UPDATED");                       // onnx_engine.cpp:503-507
            event.setSource("B200InferenceEngine");
            event.setData("model_path", config_.model_path);
            event.setData("model_hash", h);
            publishEvent(event);
        } else {
            reload_failures_++;
            reload_error_ = err;
        }
    }
}

void B200InferenceEngine::setCallback(InferenceCallback callback) { callback_ = std::move(callback); }

size_t B200InferenceEngine::getQueueSize() const {
    size_t n = 0;
    for (zl_engine* e : engines_) n += zl_engine_queue_size(e);
    return n;
}

void B200InferenceEngine::onResult(void* user, uint32_t client_id, uint32_t frame_id, uint64_t timestamp, int32_t status,
                                   const zl_det* dets, int32_t n) {
    auto* self = static_cast<B200InferenceEngine*>(user);
    if (!self->callback_) return;
    try {
        GameState state;
        state.frame_id = frame_id;
        state.timestamp = timestamp;
        if (status == ZL_OK) {
            const uint64_t now_ms = (uint64_t)std::chrono::duration_cast<std::chrono::milliseconds>(
                                        std::chrono::system_clock::now().time_since_epoch()).count();
            state.detections.resize((size_t)n);
            for (int32_t i = 0; i < n; ++i) {
                Detection& d = state.detections[i];
                d.box = BoundingBox{dets[i].x, dets[i].y, dets[i].w, dets[i].h};
                d.confidence = dets[i].confidence;
                d.class_id = dets[i].class_id;
                d.track_id = 0;              // onnx_engine.cpp:812: tracking belongs to the game adapter
                d.timestamp = now_ms;        // onnx_engine.cpp:813-815
            }
        }
        self->callback_(client_id, state);
        if (status == ZL_OK) EventBus::getInstance().publishInferenceEvent(events::INFERENCE_COMPLETED, client_id, frame_id);   // onnx_engine.cpp:359-363
    } catch (...) {
        self->callback_errors_++;
    }
}

// the device already wrote the frame's result as the reference's wire records, so building the GameState
// is one memcpy of count x 40 bytes; no detection is touched here.
void B200InferenceEngine::onWire(void* user, uint32_t client_id, uint32_t frame_id, uint64_t timestamp, int32_t status,
                                 const uint8_t* body, size_t bytes) {
    auto* self = static_cast<B200InferenceEngine*>(user);
    if (!self->callback_) return;
    try {
        GameState state;
        state.frame_id = frame_id;
        state.timestamp = timestamp;
        if (status == ZL_OK && body && bytes >= ZL_WIRE_HEADER_BYTES) {
            uint16_t count = 0;
            std::memcpy(&count, body + 12, 2);
            const size_t n = std::min<size_t>(count, (bytes - ZL_WIRE_HEADER_BYTES) / ZL_WIRE_DET_BYTES);
            state.detections.resize(n);
            static_assert(sizeof(Detection) == ZL_WIRE_DET_BYTES, "Detection is the 40-byte wire record");
            if (n) std::memcpy(static_cast<void*>(state.detections.data()), body + ZL_WIRE_HEADER_BYTES, n * ZL_WIRE_DET_BYTES);
        }
        self->callback_(client_id, state);
        if (status == ZL_OK) EventBus::getInstance().publishInferenceEvent(events::INFERENCE_COMPLETED, client_id, frame_id);   // onnx_engine.cpp:359-363
    } catch (...) {
        self->callback_errors_++;
    }
}

std::unordered_map<std::string, std::string> B200InferenceEngine::getStatus() const {
    std::unordered_map<std::string, std::string> s;
    zl_stats acc{};
    double lat_sum = 0, p99 = 0, dev_ms = 0;
    for (zl_engine* e : engines_) {
        zl_stats st{};
        if (zl_engine_get_stats(e, &st) != ZL_OK) continue;
        acc.inference_count += st.inference_count; acc.inference_errors += st.inference_errors; acc.dropped_frames += st.dropped_frames;
        acc.queue_size += st.queue_size; acc.queue_high_water_mark = std::max(acc.queue_high_water_mark, st.queue_high_water_mark);
        acc.batches += st.batches; acc.graph_captured += st.graph_captured;
        lat_sum += st.avg_inference_time_ms; p99 = std::max(p99, st.p99_inference_time_ms); dev_ms += st.avg_device_time_ms;
    }
    const double ne = engines_.empty() ? 1.0 : (double)engines_.size();
    // the reference's keys (onnx_engine.cpp:279-312) ...
    s["name"] = getName();
    s["simulation_mode"] = "false";
    s["running"] = running_ ? "true" : "false";
    s["model_path"] = config_.model_path;
    s["model_version"] = std::to_string(model_version_.load());
    {
        std::lock_guard<std::mutex> g(monitor_mu_);
        s["model_hash"] = model_hash_;
        s["model_reload_failures"] = std::to_string(reload_failures_);
        s["model_reload_error"] = reload_error_;
    }
    s["wire_results"] = config_.b200.wire_results ? "device" : "host";
    s["queue_size"] = std::to_string(acc.queue_size);
    s["queue_high_water_mark"] = std::to_string(acc.queue_high_water_mark);
    s["inference_count"] = std::to_string(acc.inference_count);
    s["inference_errors"] = std::to_string(acc.inference_errors + callback_errors_.load());
    s["dropped_frames"] = std::to_string(acc.dropped_frames);
    s["int8_quantization"] = "disabled";
    s["zero_copy"] = "enabled";
    s["dynamic_batching"] = config_.b200.max_batch > 1 ? "enabled" : "disabled";
    s["avg_inference_time_ms"] = std::to_string(lat_sum / ne);
    s["p99_inference_time_ms"] = std::to_string(p99);
    s["avg_preprocessing_time_ms"] = "0";
    s["avg_postprocessing_time_ms"] = "0";
    s["worker_threads"] = std::to_string(engines_.size() * (size_t)std::max(config_.b200.num_lanes, 1));
    // ... plus the B200 extras
    s["devices"] = std::to_string(engines_.size());
    s["fp_mode"] = config_.b200.precision;
    s["graph_captured"] = std::to_string(acc.graph_captured);
    s["batches"] = std::to_string(acc.batches);
    s["avg_device_time_ms"] = std::to_string(dev_ms / ne);
    return s;
}

REGISTER_INFERENCE_ENGINE(B200InferenceEngineFactory)

}  // namespace zero_latency
