// host_test.cpp — exercises the adapter the way the reference server does (src/server/main.cpp:224-242,
// src/network/network_server.cpp:184-283): pick the engine by name, initialize, setCallback, submit frames
// from a "receive thread", read getStatus from a "monitor thread", shutdown.
//   host_test cpu                      : checks that need no GPU (registry, error mapping, no-fallback)
//   host_test gpu <weights.zlw> <nc> <precision> <out.bin> <frames.bin> <n> <w> <h>
//                                      : full run; writes per-frame detections for the python test to compare
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <mutex>
#include <thread>

#include "b200_engine.h"

using namespace zero_latency;

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) { std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

static int run_cpu() {
    auto& mgr = InferenceEngineManager::getInstance();
    CHECK(mgr.isEngineAvailable("b200"));                       // static-init registration (inference_engine.h:94-103)
    CHECK(!mgr.isEngineAvailable("onnx"));
    CHECK(mgr.createEngine("nope", ServerConfig{}) == nullptr);
    ServerConfig cfg;
    cfg.inference_engine = "b200";
    cfg.model_path = "/nonexistent/model.zlw";
    auto eng = mgr.createEngine(cfg.inference_engine, cfg);
    CHECK(eng != nullptr);
    CHECK(eng->getName() == "b200");
    // submit before initialize -> NOT_INITIALIZED (onnx_engine.cpp:224-226)
    InferenceRequest rq;
    rq.width = 2; rq.height = 2; rq.data.assign(12, 0);
    auto r = eng->submitInference(rq);
    CHECK(r.hasError() && r.error().code == ErrorCode::NOT_INITIALIZED);
    // missing model: an error, NOT the reference's simulation mode
    auto ri = eng->initialize();
    CHECK(ri.hasError() && ri.error().code == ErrorCode::MODEL_NOT_FOUND);
    auto st = eng->getStatus();
    for (const char* k : {"name", "simulation_mode", "running", "model_path", "model_hash", "queue_size", "queue_high_water_mark",
                          "inference_count", "inference_errors", "dropped_frames", "avg_inference_time_ms", "p99_inference_time_ms",
                          "avg_preprocessing_time_ms", "avg_postprocessing_time_ms", "worker_threads", "fp_mode", "graph_captured"})
        CHECK(st.count(k) == 1);
    CHECK(st["running"] == "false" && st["simulation_mode"] == "false");
    CHECK(st.count("model_reload_failures") == 1 && st.count("wire_results") == 1);
    CHECK(eng->getQueueSize() == 0);
    // model_hash is SHA-256 like the reference's calculateModelHash (onnx_engine.cpp:1087-1124): FIPS 180-4 vectors
    {
        const char* tmp = "/tmp/zl_host_test_sha.bin";
        { std::ofstream o(tmp, std::ios::binary | std::ios::trunc); o << "abc"; }
        ServerConfig hc; hc.model_path = tmp;
        auto he = mgr.createEngine("b200", hc);
        (void)he->initialize();                                      // fails (not a model / no GPU) but has hashed the file
        CHECK(he->getStatus()["model_hash"] == "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad");
        { std::ofstream o(tmp, std::ios::binary | std::ios::trunc); o << std::string(1000000, 'a'); }
        auto he2 = mgr.createEngine("b200", hc);
        (void)he2->initialize();
        CHECK(he2->getStatus()["model_hash"] == "cdc76e5c9914fb9281a1c7e284d73e67f1809a48a497200e046d39ccc7112cd0");
        std::remove(tmp);
    }
    CHECK(eng->shutdown().isOk());
    // an existing file that is not a model + no GPU: still a clean error, never a fallback
    cfg.model_path = "/proc/self/cmdline";
    auto eng2 = mgr.createEngine("b200", cfg);
    auto r2 = eng2->initialize();
    CHECK(r2.hasError());
    const int c = (int)r2.error().code;
    CHECK(c == 303 || c == 202 || c == 2);      // INSUFFICIENT_RESOURCES (no device) / MODEL_LOAD_FAILED (on a GPU box)
    std::printf("host_test cpu: ok (%s)\n", r2.error().toString().c_str());
    return 0;
}

static int run_gpu(int argc, char** argv) {
    CHECK(argc >= 10);
    ServerConfig cfg;
    cfg.inference_engine = "b200";
    cfg.model_path = argv[2];
    cfg.b200.num_classes = std::atoi(argv[3]);
    cfg.b200.precision = argv[4];
    const char* out_path = argv[5];
    const char* frames_path = argv[6];
    const int n = std::atoi(argv[7]), w = std::atoi(argv[8]), h = std::atoi(argv[9]);
    cfg.b200.max_batch = 4;
    cfg.b200.max_frame_width = w; cfg.b200.max_frame_height = h;
    cfg.max_queue_size = 64;
    cfg.b200.wire_results = !(argc >= 11 && std::strcmp(argv[10], "host") == 0);   // default: 40-byte records written by the device
    std::atomic<int> ev_start{0}, ev_stop{0}, ev_req{0}, ev_done{0};
    std::atomic<bool> ev_ids_ok{true};
    EventBus::getInstance().subscribe(events::SYSTEM_STARTUP, [&](const Event& e) { if (e.getSource() == "B200InferenceEngine") ev_start++; });
    EventBus::getInstance().subscribe(events::SYSTEM_SHUTDOWN, [&](const Event& e) { if (e.getSource() == "B200InferenceEngine") ev_stop++; });
    EventBus::getInstance().subscribe(events::INFERENCE_REQUESTED, [&](const Event& e) {
        ev_req++;
        if (static_cast<const InferenceEvent&>(e).getClientId() != 42) ev_ids_ok = false;
    });
    EventBus::getInstance().subscribe(events::INFERENCE_COMPLETED, [&](const Event& e) {
        ev_done++;
        if (static_cast<const InferenceEvent&>(e).getClientId() != 42) ev_ids_ok = false;
    });
    std::vector<uint8_t> frames((size_t)n * w * h * 3);
    { std::ifstream f(frames_path, std::ios::binary); CHECK(f.read((char*)frames.data(), frames.size())); }

    auto eng = InferenceEngineManager::getInstance().createEngine(cfg.inference_engine, cfg);
    CHECK(eng != nullptr);
    auto ri = eng->initialize();
    if (ri.hasError()) { std::fprintf(stderr, "initialize: %s\n", ri.error().toString().c_str()); return 1; }
    CHECK(ev_start == 1);                                                // SYSTEM_STARTUP (onnx_engine.cpp:160-162)
    CHECK(eng->getStatus()["wire_results"] == (cfg.b200.wire_results ? "device" : "host"));

    std::mutex mu;
    std::condition_variable cv;
    std::vector<GameState> got;
    std::vector<uint32_t> got_client;
    std::thread::id cb_thread;
    eng->setCallback([&](uint32_t client_id, const GameState& st) {
        std::lock_guard<std::mutex> g(mu);
        got.push_back(st);
        got_client.push_back(client_id);
        cb_thread = std::this_thread::get_id();
        cv.notify_all();
    });
    const int rounds = 3, total = rounds * n;
    std::atomic<bool> stop{false};
    std::thread monitor([&] { while (!stop) { auto s = eng->getStatus(); (void)s; (void)eng->getQueueSize(); std::this_thread::sleep_for(std::chrono::milliseconds(1)); } });
    std::thread::id submit_thread;
    std::thread receiver([&] {
        submit_thread = std::this_thread::get_id();
        for (int i = 0; i < total; ++i) {
            InferenceRequest rq;
            rq.client_id = 42; rq.frame_id = (uint32_t)i; rq.timestamp = 1000u + i;
            rq.width = (uint16_t)w; rq.height = (uint16_t)h;
            const uint8_t* src = frames.data() + (size_t)(i % n) * w * h * 3;
            rq.data.assign(src, src + (size_t)w * h * 3);
            for (;;) {
                auto r = eng->submitInference(rq);
                if (r.isOk()) break;
                if (r.error().code != ErrorCode::INFERENCE_ERROR) { std::fprintf(stderr, "submit: %s\n", r.error().toString().c_str()); std::abort(); }
                std::this_thread::sleep_for(std::chrono::microseconds(200));     // queue full -> the caller drops / retries
            }
        }
    });
    receiver.join();
    {
        std::unique_lock<std::mutex> lk(mu);
        CHECK(cv.wait_for(lk, std::chrono::seconds(60), [&] { return (int)got.size() == total; }));
    }
    stop = true;
    monitor.join();
    // bad frame length -> INVALID_INPUT (onnx_engine.cpp:659-665)
    InferenceRequest bad;
    bad.client_id = 42; bad.width = (uint16_t)w; bad.height = (uint16_t)h; bad.data.assign(10, 0);
    auto rb = eng->submitInference(bad);
    CHECK(rb.hasError() && rb.error().code == ErrorCode::INVALID_INPUT);
    CHECK(cb_thread != submit_thread);                                   // callbacks come from an engine-owned thread
    for (int i = 0; i < total; ++i) {
        CHECK(got[i].frame_id == (uint32_t)i && got[i].timestamp == 1000u + i && got_client[i] == 42);   // every frame, in order
        for (const auto& d : got[i].detections) CHECK(d.track_id == 0 && d.timestamp > 0);
    }
    auto st = eng->getStatus();
    CHECK(st["running"] == "true" && std::stoull(st["inference_count"]) >= (unsigned long long)total);
    std::ofstream out(out_path, std::ios::binary);
    for (int i = 0; i < n; ++i) {           // first round only: [u32 count][count x 40-byte Detection]
        const uint32_t c = (uint32_t)got[i].detections.size();
        out.write((const char*)&c, 4);
        out.write((const char*)got[i].detections.data(), (std::streamsize)c * sizeof(Detection));
        // later rounds must repeat the first bit for bit (timestamps aside)
        for (int r = 1; r < rounds; ++r) {
            const auto& o = got[r * n + i].detections;
            CHECK(o.size() == c);
            for (uint32_t k = 0; k < c; ++k) CHECK(std::memcmp(&o[k], &got[i].detections[k], 24) == 0);
        }
    }
    CHECK(eng->shutdown().isOk());
    CHECK(eng->submitInference(bad).error().code == ErrorCode::NOT_INITIALIZED);
    // the four events of the reference engine (onnx_engine.cpp:160-162,214-216,229-231,359-363); the rejected frame was
    // announced too, like in the reference, which publishes before it looks at the request
    CHECK(ev_stop == 1 && ev_req >= total && ev_done == total && ev_ids_ok);
    std::printf("host_test gpu: ok, %d frames, avg latency %s ms, graphs %s\n", total, st["avg_inference_time_ms"].c_str(), st["graph_captured"].c_str());
    return 0;
}

// host_test reload <weightsA> <weightsB> <nc> <frame.bin> <w> <h>: serve with A, overwrite the model file with B, wait for
// the monitor thread to pick it up, check that detections change and the status map reports the new version/hash.
static int run_reload(int argc, char** argv) {
    CHECK(argc >= 8);
    const std::string live = std::string(argv[2]) + ".live";
    auto copy = [&](const char* src) { std::ifstream i(src, std::ios::binary); std::ofstream o(live, std::ios::binary | std::ios::trunc); o << i.rdbuf(); };
    copy(argv[2]);
    ServerConfig cfg;
    cfg.inference_engine = "b200"; cfg.model_path = live;
    cfg.b200.num_classes = std::atoi(argv[4]); cfg.b200.max_batch = 2;
    cfg.b200.use_model_monitor = true; cfg.b200.model_check_interval_ms = 100;
    const int w = std::atoi(argv[6]), h = std::atoi(argv[7]);
    cfg.b200.max_frame_width = w; cfg.b200.max_frame_height = h;
    std::vector<uint8_t> frame((size_t)w * h * 3);
    { std::ifstream f(argv[5], std::ios::binary); CHECK(f.read((char*)frame.data(), frame.size())); }
    auto eng = InferenceEngineManager::getInstance().createEngine("b200", cfg);
    CHECK(eng->initialize().isOk());
    std::mutex mu; std::condition_variable cv; std::vector<GameState> got;
    eng->setCallback([&](uint32_t, const GameState& st) { std::lock_guard<std::mutex> g(mu); got.push_back(st); cv.notify_all(); });
    auto infer = [&](uint32_t id) -> GameState {
        InferenceRequest rq; rq.client_id = 1; rq.frame_id = id; rq.width = (uint16_t)w; rq.height = (uint16_t)h; rq.data = frame;
        while (!eng->submitInference(rq).isOk()) std::this_thread::sleep_for(std::chrono::milliseconds(1));
        std::unique_lock<std::mutex> lk(mu);
        cv.wait_for(lk, std::chrono::seconds(30), [&] { return !got.empty() && got.back().frame_id == id; });
        return got.back();
    };
    const GameState a = infer(1);
    const std::string hash_a = eng->getStatus()["model_hash"];
    copy(argv[3]);                                              // the operator drops a new model file in place
    bool changed = false;
    for (int i = 0; i < 100 && !changed; ++i) {                 // frames keep flowing while the reload happens
        std::this_thread::sleep_for(std::chrono::milliseconds(50));
        (void)infer(100 + i);
        changed = eng->getStatus()["model_version"] == "2";
    }
    CHECK(changed);
    const GameState b = infer(999);
    CHECK(eng->getStatus()["model_hash"] != hash_a);
    bool differ = a.detections.size() != b.detections.size();
    for (size_t i = 0; !differ && i < a.detections.size(); ++i) differ = std::memcmp(&a.detections[i], &b.detections[i], 24) != 0;
    CHECK(differ);                                              // different weights -> different detections
    // a corrupt file must not take the engine down
    { std::ofstream o(live, std::ios::binary | std::ios::trunc); o << "garbage"; }
    std::this_thread::sleep_for(std::chrono::milliseconds(400));
    const GameState c = infer(1000);
    CHECK(c.detections.size() == b.detections.size());
    CHECK(eng->getStatus()["model_version"] == "2");
    CHECK(eng->getStatus()["model_reload_failures"] == "1" && !eng->getStatus()["model_reload_error"].empty());   // visible, not silent
    CHECK(eng->getStatus()["model_hash"] != hash_a);
    CHECK(eng->shutdown().isOk());
    std::remove(live.c_str());
    std::printf("host_test reload: ok (%zu -> %zu detections)\n", a.detections.size(), b.detections.size());
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && std::strcmp(argv[1], "gpu") == 0) return run_gpu(argc, argv);
    if (argc >= 2 && std::strcmp(argv[1], "reload") == 0) return run_reload(argc, argv);
    return run_cpu();
}
