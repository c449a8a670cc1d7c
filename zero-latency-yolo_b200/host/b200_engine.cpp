// b200_engine.cpp — see b200_engine.h.  No exception leaves this class (the reference wraps every
// engine entry point in try/catch -> Result::error, onnx_engine.cpp:165-169,621-645).
#include "b200_engine.h"

#include <chrono>
#include <fstream>

namespace zero_latency {

namespace {
int precisionOf(const std::string& s) { return s == "fp32" ? ZL_PRECISION_FP32 : (s == "bf16" ? ZL_PRECISION_BF16 : ZL_PRECISION_FP16); }
int scaleOf(const std::string& s) { return s == "s" ? ZL_SCALE_S : (s == "m" ? ZL_SCALE_M : ZL_SCALE_N); }
std::string lastError() { const char* e = zl_last_error(); return e ? std::string(e) : std::string(); }

// FNV-1a of the weights file: the status map's "model_hash" (the reference uses SHA-256 for its
// hot-reload watcher, onnx_engine.cpp:1087-1124; reload is a "next" row).
std::string fileHash(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return "";
    uint64_t h = 1469598103934665603ull;
    char buf[65536];
    while (f.read(buf, sizeof(buf)) || f.gcount() > 0) {
        for (std::streamsize i = 0; i < f.gcount(); ++i) { h ^= (uint8_t)buf[i]; h *= 1099511628211ull; }
        if (f.eof()) break;
    }
    char out[17];
    snprintf(out, sizeof(out), "%016llx", (unsigned long long)h);
    return out;
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
            zl_engine* e = nullptr;
            int32_t rc = zl_engine_create(&c, &e);
            if (rc == ZL_OK) rc = zl_engine_load_weights(e, config_.model_path.c_str());
            if (rc == ZL_OK) rc = zl_engine_set_callback(e, &B200InferenceEngine::onResult, this);
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
        }
        return Result<void>::ok();
    } catch (const std::exception& ex) {
        return Result<void>::error(ErrorCode::SYSTEM_ERROR, ex.what());
    }
}

Result<void> B200InferenceEngine::submitInference(const InferenceRequest& request) {
    if (!running_) return Result<void>::error(ErrorCode::NOT_INITIALIZED, "Inference engine not running");   // onnx_engine.cpp:224-226
    zl_engine* e = engines_[request.client_id % engines_.size()];
    const int32_t rc = zl_engine_submit(e, request.client_id, request.frame_id, request.timestamp, request.width, request.height,
                                        request.data.data(), request.data.size(), request.is_keyframe ? 1 : 0);
    if (rc != ZL_OK) return Result<void>::error(toErrorCode(rc), lastError());
    return Result<void>::ok();
}

// Model hot reload (onnx_engine.cpp:473-515): re-hash the model file periodically; on change load the new weights into
// every device's engine (the swap is atomic under the lanes' locks: frames in flight finish on the old weights) and
// recapture the CUDA graphs.  A file that fails to load leaves the running model untouched.
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
        bool ok = true;
        for (zl_engine* e : engines_) {
            if (zl_engine_load_weights(e, config_.model_path.c_str()) != ZL_OK) { ok = false; break; }
            if (zl_engine_warmup(e, 1) != ZL_OK) { ok = false; break; }
        }
        last_hash = h;                                          // do not retry a bad file every interval
        if (ok) {
            std::lock_guard<std::mutex> g(monitor_mu_);
            model_hash_ = h;
            model_version_++;
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
    { std::lock_guard<std::mutex> g(monitor_mu_); s["model_hash"] = model_hash_; }
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
