// zl_iface.h — the slice of the reference's C++ detector interface this adapter implements.
//
// In the reference tree these declarations live in
//   src/inference/inference_engine.h:16-103  (InferenceRequest, InferenceCallback, IInferenceEngine,
//                                             IInferenceEngineFactory, InferenceEngineManager, REGISTER_INFERENCE_ENGINE)
//   src/common/result.h:14-221               (ErrorCode, Error, Result<T>)
//   src/common/types.h:16-40                 (BoundingBox, Detection, GameState)
//   src/server/config.h:110-149,305-345      (DetectionConfig, ServerConfig: only the fields the path reads)
// The reference headers do not compile as shipped (two conflicting `enum class ErrorCode`, SURVEY.md §0
// fact 5) and pull in the whole server, so a standalone build uses this freshly written, minimal
// restatement with the SAME names, signatures and numeric values.  A maintainer building inside the
// reference tree defines ZL_USE_REFERENCE_HEADERS and gets the real ones.
#pragma once
#ifdef ZL_USE_REFERENCE_HEADERS
#include "inference/inference_engine.h"
#include "common/event_bus.h"
#else
#include <cstdint>
#include <functional>
#include <chrono>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace zero_latency {

enum class ErrorCode : int {
    OK = 0, UNKNOWN_ERROR = 1, INVALID_ARGUMENT = 2, NOT_INITIALIZED = 3, TIMEOUT = 4,
    INFERENCE_ERROR = 200, MODEL_NOT_FOUND = 201, MODEL_LOAD_FAILED = 202, INVALID_INPUT = 203, INFERENCE_TIMEOUT = 204,
    SYSTEM_ERROR = 300, FILE_NOT_FOUND = 301, FILE_ACCESS_DENIED = 302, INSUFFICIENT_RESOURCES = 303,
    CONFIG_ERROR = 400
};

struct Error {
    ErrorCode code = ErrorCode::OK;
    std::string message;
    bool isOk() const { return code == ErrorCode::OK; }
    std::string toString() const { return "Error " + std::to_string(static_cast<int>(code)) + ": " + message; }
};

template <typename T = void> class Result;

template <> class Result<void> {
public:
    static Result ok() { return Result(); }
    static Result error(ErrorCode c, const std::string& m) { Result r; r.err_ = Error{c, m}; return r; }
    bool isOk() const { return err_.isOk(); }
    bool hasError() const { return !err_.isOk(); }
    const Error& error() const { return err_; }
private:
    Error err_;
};

template <typename T> class Result {
public:
    static Result ok(T v) { Result r; r.val_ = std::move(v); return r; }
    static Result error(ErrorCode c, const std::string& m) { Result r; r.err_ = Error{c, m}; return r; }
    bool isOk() const { return err_.isOk(); }
    bool hasError() const { return !err_.isOk(); }
    const Error& error() const { return err_; }
    const T& value() const { return val_; }
private:
    T val_{};
    Error err_;
};

struct BoundingBox { float x, y, width, height; };                     // centre x, centre y, w, h
struct Detection { BoundingBox box; float confidence; int class_id; uint32_t track_id; uint64_t timestamp; };
static_assert(sizeof(Detection) == 40, "Detection must keep the reference's 40-byte layout");
struct GameState { uint32_t frame_id; uint64_t timestamp; std::vector<Detection> detections; };

struct InferenceRequest {
    uint32_t client_id = 0;
    uint32_t frame_id = 0;
    uint64_t timestamp = 0;
    uint16_t width = 0, height = 0;
    std::vector<uint8_t> data;      // width*height*3 bytes, HWC, BGR
    bool is_keyframe = false;
};

using InferenceCallback = std::function<void(uint32_t client_id, const GameState& state)>;

struct DetectionConfig {
    uint16_t model_width = 416, model_height = 416;
    std::unordered_map<std::string, float> class_weights;   // configured by the reference but never read
};

// B200-only knobs; in configs/server.json they live in a "b200" sub-object the reference ignores.
struct B200Config {
    std::vector<int> devices{0};
    std::string precision = "fp16";     // "fp32" | "bf16" | "fp16"
    std::string scale = "n";            // "n" | "s" | "m"
    int num_classes = 4;
    int max_batch = 8;
    int max_frame_width = 1920, max_frame_height = 1080;   // constants.h:10 maximum client frame
    int num_lanes = 2;
    int batch_window_us = 0;
    bool use_class_weights = false;     // false == the reference's effective behaviour
    bool wire_results = true;           // detections arrive as the reference's 40-byte wire records, written by the device
    bool use_model_monitor = false;     // hot reload: the reference's optimization.use_model_monitor (onnx_engine.cpp:38,145)
    int model_check_interval_ms = 10000; // the reference checks every 10 s (onnx_engine.cpp:480)
    std::vector<float> class_weights;
};

struct ServerConfig {
    std::string model_path = "models/yolo_nano_cs16.onnx";
    std::string inference_engine = "onnx";
    uint32_t target_fps = 60;
    float confidence_threshold = 0.5f;
    float nms_threshold = 0.45f;
    size_t max_queue_size = 8;
    bool use_cpu_affinity = true;
    int cpu_core_id = 0;
    bool use_high_priority = true;
    uint8_t worker_threads = 1;
    DetectionConfig detection;
    B200Config b200;
};

class IInferenceEngine {
public:
    virtual ~IInferenceEngine() = default;
    virtual Result<void> initialize() = 0;
    virtual Result<void> shutdown() = 0;
    virtual Result<void> submitInference(const InferenceRequest& request) = 0;
    virtual void setCallback(InferenceCallback callback) = 0;
    virtual size_t getQueueSize() const = 0;
    virtual std::string getName() const = 0;
    virtual std::unordered_map<std::string, std::string> getStatus() const = 0;
};

class IInferenceEngineFactory {
public:
    virtual ~IInferenceEngineFactory() = default;
    virtual std::unique_ptr<IInferenceEngine> createEngine(const ServerConfig& config) = 0;
    virtual std::string getName() const = 0;
};

class InferenceEngineManager {
public:
    static InferenceEngineManager& getInstance() { static InferenceEngineManager m; return m; }
    void registerFactory(std::shared_ptr<IInferenceEngineFactory> f) { if (f) factories_[f->getName()] = std::move(f); }
    std::unique_ptr<IInferenceEngine> createEngine(const std::string& name, const ServerConfig& cfg) {
        auto it = factories_.find(name);
        return it == factories_.end() ? nullptr : it->second->createEngine(cfg);
    }
    bool isEngineAvailable(const std::string& name) const { return factories_.count(name) != 0; }
    std::vector<std::string> getAvailableEngines() const {
        std::vector<std::string> v;
        for (const auto& kv : factories_) v.push_back(kv.first);
        return v;
    }
private:
    std::map<std::string, std::shared_ptr<IInferenceEngineFactory>> factories_;
};

// ---- the slice of src/common/event_bus.h:15-176 the engine publishes to (same names and signatures) ----
using EventType = std::string;
namespace events {
constexpr const char* SYSTEM_STARTUP = "SYSTEM_STARTUP";            // event_bus.h:17
constexpr const char* SYSTEM_SHUTDOWN = "SYSTEM_SHUTDOWN";          // event_bus.h:18
constexpr const char* INFERENCE_REQUESTED = "INFERENCE_REQUESTED";  // event_bus.h:25
constexpr const char* INFERENCE_COMPLETED = "INFERENCE_COMPLETED";  // event_bus.h:26
}  // namespace events

class Event {
public:
    explicit Event(EventType type) : type_(std::move(type)), timestamp_(std::chrono::system_clock::now()) {}
    virtual ~Event() = default;
    const EventType& getType() const { return type_; }
    const std::chrono::system_clock::time_point& getTimestamp() const { return timestamp_; }
    void setSource(const std::string& s) { source_ = s; }
    const std::string& getSource() const { return source_; }
    void setData(const std::string& k, const std::string& v) { data_[k] = v; }
    std::string getData(const std::string& k) const { auto it = data_.find(k); return it == data_.end() ? std::string() : it->second; }
private:
    EventType type_;
    std::chrono::system_clock::time_point timestamp_;
    std::string source_;
    std::unordered_map<std::string, std::string> data_;
};

class InferenceEvent : public Event {
public:
    InferenceEvent(const EventType& type, uint32_t client_id, uint32_t frame_id) : Event(type), client_id_(client_id), frame_id_(frame_id) {}
    uint32_t getClientId() const { return client_id_; }
    uint32_t getFrameId() const { return frame_id_; }
private:
    uint32_t client_id_, frame_id_;
};

using EventHandler = std::function<void(const Event&)>;

class EventBus {
public:
    static EventBus& getInstance() { static EventBus b; return b; }
    void subscribe(const EventType& type, EventHandler handler) { std::lock_guard<std::mutex> g(mu_); handlers_[type].push_back(std::move(handler)); }
    void publish(const Event& event) {
        std::vector<EventHandler> copy;
        { std::lock_guard<std::mutex> g(mu_); auto it = handlers_.find(event.getType()); if (it != handlers_.end()) copy = it->second; }
        for (const auto& h : copy) { try { h(event); } catch (...) {} }
    }
    void publishInferenceEvent(const EventType& type, uint32_t client_id, uint32_t frame_id) { InferenceEvent e(type, client_id, frame_id); publish(e); }
private:
    std::mutex mu_;
    std::unordered_map<EventType, std::vector<EventHandler>> handlers_;
};
inline void publishEvent(const Event& e) { EventBus::getInstance().publish(e); }

#define REGISTER_INFERENCE_ENGINE(factory_class)                                                        \
    namespace {                                                                                         \
    struct Register##factory_class {                                                                    \
        Register##factory_class() {                                                                     \
            zero_latency::InferenceEngineManager::getInstance().registerFactory(std::make_shared<factory_class>()); \
        }                                                                                               \
    };                                                                                                  \
    static Register##factory_class register_##factory_class;                                            \
    }

}  // namespace zero_latency
#endif  // ZL_USE_REFERENCE_HEADERS
