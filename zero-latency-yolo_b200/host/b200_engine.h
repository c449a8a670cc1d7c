// b200_engine.h — B200InferenceEngine: the drop-in for OnnxInferenceEngine behind IInferenceEngine.
//
// Replaces src/inference/onnx_engine.{h,cpp} (reference) for the server-side detector path only.
// It is a thin C++17 adapter over the C-ABI (include/zl_b200.h): one zl_engine per configured CUDA
// device, frames routed by client_id % devices (frames never share state, so there is no collective
// and per-client callback order is preserved without a reorder buffer).
// Selected with  "inference_engine": "b200"  — src/server/main.cpp:228-230 already routes any
// non-"onnx" name through InferenceEngineManager.
#pragma once
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "../../include/zl_b200.h"
#include "zl_iface.h"

namespace zero_latency {

class B200InferenceEngine final : public IInferenceEngine {
public:
    explicit B200InferenceEngine(const ServerConfig& config);
    ~B200InferenceEngine() override;

    Result<void> initialize() override;                                     // onnx_engine.cpp:67-170
    Result<void> shutdown() override;                                       // onnx_engine.cpp:173-221
    Result<void> submitInference(const InferenceRequest& request) override; // onnx_engine.cpp:223-261
    void setCallback(InferenceCallback callback) override;                  // onnx_engine.cpp:263-265
    size_t getQueueSize() const override;                                   // onnx_engine.cpp:268-270
    std::string getName() const override { return "b200"; }
    std::unordered_map<std::string, std::string> getStatus() const override; // onnx_engine.cpp:279-312

private:
    static void onResult(void* user, uint32_t client_id, uint32_t frame_id, uint64_t timestamp, int32_t status,
                         const zl_det* dets, int32_t n);
    static void onWire(void* user, uint32_t client_id, uint32_t frame_id, uint64_t timestamp, int32_t status,
                       const uint8_t* body, size_t bytes);
    static ErrorCode toErrorCode(int32_t rc) { return static_cast<ErrorCode>(rc); }
    void modelMonitorThreadFunc();                                          // onnx_engine.cpp:473-515

    ServerConfig config_;
    std::vector<zl_engine*> engines_;
    InferenceCallback callback_;
    std::atomic<bool> running_{false};
    std::atomic<uint64_t> callback_errors_{0};
    std::string model_hash_;
    std::atomic<uint32_t> model_version_{1};
    uint64_t reload_failures_ = 0;          // under monitor_mu_
    std::string reload_error_;
    mutable std::mutex monitor_mu_;
    std::condition_variable monitor_cv_;
    std::thread monitor_thread_;
};

class B200InferenceEngineFactory final : public IInferenceEngineFactory {
public:
    std::unique_ptr<IInferenceEngine> createEngine(const ServerConfig& config) override {
        return std::make_unique<B200InferenceEngine>(config);
    }
    std::string getName() const override { return "b200"; }
};

}  // namespace zero_latency
