// engine.cpp — model graph, buffers, CUDA graphs, batching queue (see engine.h).
#include "engine.h"
#include "weights.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <pthread.h>
#include <sched.h>
#include <unistd.h>

#include <cuda_fp16.h>

namespace zl {

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
void set_error(const std::string& m) { g_err = m; }
const char* get_error() { return g_err.c_str(); }

// ------------------------------------------------------------------ helpers
static inline uint16_t f2bf(float f) {          // round-to-nearest-even, NaN preserved
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

static inline uint16_t f2h(float f) { const __half h = __float2half_rn(f); uint16_t u; std::memcpy(&u, &h, 2); return u; }

static int ch(int c, float width, int maxc) { return (int)std::ceil(std::min(c, maxc) * width / 8.0) * 8; }
static int rep(int n, float depth) { return std::max((int)std::lround(n * depth), 1); }   // ties: none occur for n in {3,6}

Engine::Engine(const zl_config& c) : cfg(c) {}

Engine::~Engine()
{
    stop_workers();
    cudaSetDevice(cfg.device);
    for (auto& L : lanes) free_lane(*L);
    convs.clear();
    if (d_class_weights) cudaFree(d_class_weights);
    if (h_slots) cudaFreeHost(h_slots);
}

int32_t Engine::build_model_def()
{
    static const float depth[3] = {0.33f, 0.33f, 0.67f}, width[3] = {0.25f, 0.50f, 0.75f};
    static const int maxc[3] = {1024, 1024, 768};
    if (cfg.scale < 0 || cfg.scale > 2) ZL_FAIL(ZL_INVALID_ARGUMENT, "scale must be n/s/m");
    md.scale = cfg.scale;
    md.nc = cfg.num_classes;
    const int base[5] = {64, 128, 256, 512, 1024};
    for (int i = 0; i < 5; ++i) md.c[i] = ch(base[i], width[cfg.scale], maxc[cfg.scale]);
    md.n[0] = rep(3, depth[cfg.scale]); md.n[1] = rep(6, depth[cfg.scale]);
    md.n[2] = rep(6, depth[cfg.scale]); md.n[3] = rep(3, depth[cfg.scale]);
    md.nh = rep(3, depth[cfg.scale]);
    md.cb = std::max(16, std::max(md.c[2] / 4, 64));
    md.cc = std::max(md.c[2], std::min(md.nc, 100));
    // 16-bit modes: the tensor-core kernels take channel counts in multiples of 16 (UMMA K step, 16-byte rows), and cc = 100
    // for every model with >= 100 classes.  The class branch is widened with ZERO channels on the device: zero filter rows
    // and bias give SiLU(0) = 0, zero filter columns in the consumer add exact zeros — the results do not change.
    md.ccd = cfg.precision != ZL_PRECISION_FP32 ? round_up(md.cc, 16) : md.cc;
    return ZL_OK;
}

int32_t Engine::init()
{
    if (cfg.model_w <= 0 || cfg.model_h <= 0 || cfg.model_w % 32 || cfg.model_h % 32)
        ZL_FAIL(ZL_INVALID_ARGUMENT, "model_w/model_h must be positive multiples of 32");
    if (cfg.num_classes < 1 || cfg.num_classes > kMaxClasses) ZL_FAIL(ZL_INVALID_ARGUMENT, "num_classes out of range");
    if (cfg.max_batch < 1 || cfg.max_batch > 256) ZL_FAIL(ZL_INVALID_ARGUMENT, "max_batch must be 1..256");
    if (cfg.precision < ZL_PRECISION_FP32 || cfg.precision > ZL_PRECISION_FP16) ZL_FAIL(ZL_INVALID_ARGUMENT, "bad precision");
    if (cfg.preprocess_mode != ZL_PRE_STRETCH_NEAREST && cfg.preprocess_mode != ZL_PRE_LETTERBOX) ZL_FAIL(ZL_INVALID_ARGUMENT, "bad preprocess_mode");
    if (cfg.max_frame_w <= 0) cfg.max_frame_w = cfg.model_w;
    if (cfg.max_frame_h <= 0) cfg.max_frame_h = cfg.model_h;
    if (cfg.num_lanes < 1) cfg.num_lanes = 1;
    if (cfg.num_lanes > 8) cfg.num_lanes = 8;
    if (cfg.queue_depth < 1) cfg.queue_depth = 8;     // constants::INFERENCE_QUEUE_SIZE (src/common/constants.h)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        ZL_FAIL(ZL_INSUFFICIENT_RESOURCES, "no CUDA device: this engine has no CPU fallback");
    if (cfg.device < 0 || cfg.device >= ndev) ZL_FAIL(ZL_INVALID_ARGUMENT, "device ordinal out of range");
    ZL_CUDA(cudaSetDevice(cfg.device));
    cudaDeviceProp prop;
    ZL_CUDA(cudaGetDeviceProperties(&prop, cfg.device));
    if (prop.major != 10) ZL_FAIL(ZL_INSUFFICIENT_RESOURCES, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", this build is sm_100a only");
    num_sms = prop.multiProcessorCount;
    if (const char* ev = getenv("ZL_DISABLE_HALO")) use_halo = !(ev[0] == '1');
    if (const char* ev = getenv("ZL_FUSE_PRE")) fuse_pre = (ev[0] == '1');
    if (const char* ev = getenv("ZL_DISABLE_STEM")) use_stem = !(ev[0] == '1');
    if (const char* ev = getenv("ZL_FUSE_STEMS")) fuse_stems = (ev[0] == '1');
    if (cfg.preprocess_mode == ZL_PRE_LETTERBOX) fuse_pre = false;     // the fused A/B kernel only knows the parity sampling
    persist_min_units = 0;      // measured: the persistent kernel wins even for b=1 (p50 0.77 -> 0.44 ms), so it always runs when it can
    if (const char* ev = getenv("ZL_PERSIST_MIN_UNITS")) persist_min_units = atoi(ev);
    if (const char* ev = getenv("ZL_DEEP_K_PERSIST")) deep_k_persist = (ev[0] == '1');
    ZL_TRY(build_model_def());
    num_anchors = 0;
    for (int s : {8, 16, 32}) num_anchors += (cfg.model_h / s) * (cfg.model_w / s);
    if (num_anchors > kMaxAnchors) ZL_FAIL(ZL_INVALID_ARGUMENT, "model input too large (anchor count beyond key range)");
    if (cfg.class_weights) {
        ZL_CUDA(cudaMalloc(&d_class_weights, sizeof(float) * cfg.num_classes));
        ZL_CUDA(cudaMemcpy(d_class_weights, cfg.class_weights, sizeof(float) * cfg.num_classes, cudaMemcpyHostToDevice));
        cfg.class_weights = nullptr;   // caller's pointer is not retained
    }
    for (int i = 0; i < cfg.num_lanes; ++i) {
        lanes.emplace_back(new Lane());
        lanes.back()->id = i;
        ZL_TRY(alloc_lane(*lanes.back()));
    }
    return ZL_OK;
}

// ------------------------------------------------------------------ weights
// load = prepare + commit.  The two halves are separate entry points so that a host serving several devices can build the
// new weight set on EVERY device first and only then swap them all (or none): hot reload that is atomic across devices.
int32_t Engine::load_weights(const void* blob, size_t len)
{
    ZL_TRY(prepare_weights(blob, len));
    return commit_weights();
}

int32_t Engine::discard_weights()
{
    std::lock_guard<std::mutex> load_guard(load_mu);
    cudaSetDevice(cfg.device);
    pending_convs.clear();                      // ~ConvWeights frees the device buffers
    pending_by_name.clear();
    return ZL_OK;
}

int32_t Engine::commit_weights()
{
    std::lock_guard<std::mutex> load_guard(load_mu);
    if (pending_convs.empty()) ZL_FAIL(ZL_NOT_INITIALIZED, "no prepared weights to commit");
    ZL_CUDA(cudaSetDevice(cfg.device));
    {
        // swap under every lane's lock: no batch is in flight, every cached op list / graph is stale
        std::vector<std::unique_lock<std::mutex>> locks;
        for (auto& L : lanes) locks.emplace_back(L->mu);
        ZL_CUDA(cudaDeviceSynchronize());
        convs.swap(pending_convs);
        conv_by_name.swap(pending_by_name);
        for (auto& L : lanes) {
            for (auto& kv : L->graphs) cudaGraphExecDestroy(kv.second);
            L->graphs.clear();
            L->ops.clear();
        }
        weights_loaded = true;
    }
    pending_convs.clear();                      // the previous set: ~ConvWeights frees its device buffers
    pending_by_name.clear();
    return ZL_OK;
}

int32_t Engine::prepare_weights(const void* blob, size_t len)
{
    ZL_CUDA(cudaSetDevice(cfg.device));
    ParsedModel pm;
    ZL_TRY(parse_model(blob, len, &pm));              // ZLW1 container or an ultralytics ONNX export (BN fused)
    if (pm.scale != cfg.scale || pm.nc != cfg.num_classes)
        ZL_FAIL(ZL_MODEL_LOAD_FAILED, "weights are for scale " + std::to_string(pm.scale) + " nc " + std::to_string(pm.nc) + ", engine configured otherwise");
    std::lock_guard<std::mutex> load_guard(load_mu);      // concurrent loads (API call vs the adapter's model monitor) are serialised
    const std::map<std::string, HostTensor>& host_w = pm.tensors;

    // the conv list of YOLOv8, names as ultralytics exports them
    struct Spec { std::string name; int cin, cout, k, s, act, cin_dev, cout_dev; };     // *_dev: widths on the device (>= the file's: zero padding)
    std::vector<Spec> specs;
    auto conv = [&](const std::string& n, int ci, int co, int k, int s, int act = 1) { specs.push_back({n, ci, co, k, s, act, ci, co}); };
    auto c2f = [&](int idx, int ci, int co, int n) {
        const int c = co / 2;
        const std::string b = "model." + std::to_string(idx);
        conv(b + ".cv1.conv", ci, 2 * c, 1, 1);
        for (int j = 0; j < n; ++j) {
            conv(b + ".m." + std::to_string(j) + ".cv1.conv", c, c, 3, 1);
            conv(b + ".m." + std::to_string(j) + ".cv2.conv", c, c, 3, 1);
        }
        conv(b + ".cv2.conv", (2 + n) * c, co, 1, 1);
    };
    const int* c = md.c;
    conv("model.0.conv", 3, c[0], 3, 2);
    conv("model.1.conv", c[0], c[1], 3, 2);
    c2f(2, c[1], c[1], md.n[0]);
    conv("model.3.conv", c[1], c[2], 3, 2);
    c2f(4, c[2], c[2], md.n[1]);
    conv("model.5.conv", c[2], c[3], 3, 2);
    c2f(6, c[3], c[3], md.n[2]);
    conv("model.7.conv", c[3], c[4], 3, 2);
    c2f(8, c[4], c[4], md.n[3]);
    conv("model.9.cv1.conv", c[4], c[4] / 2, 1, 1);
    conv("model.9.cv2.conv", c[4] * 2, c[4], 1, 1);
    c2f(12, c[4] + c[3], c[3], md.nh);
    c2f(15, c[3] + c[2], c[2], md.nh);
    conv("model.16.conv", c[2], c[2], 3, 2);
    c2f(18, c[2] + c[3], c[3], md.nh);
    conv("model.19.conv", c[3], c[3], 3, 2);
    c2f(21, c[3] + c[4], c[4], md.nh);
    for (int l = 0; l < 3; ++l) {
        const std::string b = "model.22.cv2." + std::to_string(l);
        conv(b + ".0.conv", c[2 + l], md.cb, 3, 1);
        conv(b + ".1.conv", md.cb, md.cb, 3, 1);
        conv(b + ".2", md.cb, 64, 1, 1, 0);
    }
    for (int l = 0; l < 3; ++l) {
        const std::string b = "model.22.cv3." + std::to_string(l);
        conv(b + ".0.conv", c[2 + l], md.cc, 3, 1);
        specs.back().cout_dev = md.ccd;
        conv(b + ".1.conv", md.cc, md.cc, 3, 1);
        specs.back().cin_dev = md.ccd; specs.back().cout_dev = md.ccd;
        conv(b + ".2", md.cc, md.nc, 1, 1, 0);
        specs.back().cin_dev = md.ccd;
    }

    // validate every tensor BEFORE the first device allocation: a bad or partially written file costs nothing
    for (const Spec& s : specs) {
        auto wi = host_w.find(s.name + ".weight"), bi = host_w.find(s.name + ".bias");
        if (wi == host_w.end() || bi == host_w.end()) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "missing tensor " + s.name);
        const HostTensor& W = wi->second;
        if (W.dims.size() != 4 || (int)W.dims[0] != s.cout || (int)W.dims[1] != s.cin || (int)W.dims[2] != s.k || (int)W.dims[3] != s.k ||
            W.data.size() != (size_t)s.cout * s.cin * s.k * s.k || bi->second.data.size() != (size_t)s.cout)
            ZL_FAIL(ZL_MODEL_LOAD_FAILED, "shape mismatch for " + s.name);
    }

    // build the new weight set on the side; it replaces the live one atomically at the end (hot reload while serving).
    // ConvWeights frees its device buffers in its destructor, so every early return below releases what was uploaded.
    std::vector<std::unique_ptr<ConvWeights>> new_convs;
    std::map<std::string, ConvWeights*> new_by_name;
    const bool bf16 = cfg.precision != ZL_PRECISION_FP32;   // "bf16" == any 16-bit tensor-core mode
    const bool f16 = cfg.precision == ZL_PRECISION_FP16; (void)f16;
    for (const Spec& s : specs) {
        auto wi = host_w.find(s.name + ".weight"), bi = host_w.find(s.name + ".bias");
        if (wi == host_w.end() || bi == host_w.end()) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "missing tensor " + s.name);
        const HostTensor& W = wi->second;
        if (W.dims.size() != 4 || (int)W.dims[0] != s.cout || (int)W.dims[1] != s.cin || (int)W.dims[2] != s.k || (int)W.dims[3] != s.k ||
            bi->second.data.size() != (size_t)s.cout)
            ZL_FAIL(ZL_MODEL_LOAD_FAILED, "shape mismatch for " + s.name);
        std::unique_ptr<ConvWeights> cw(new ConvWeights());
        cw->name = s.name; cw->cin = s.cin_dev; cw->cout = s.cout_dev; cw->k = s.k; cw->stride = s.s; cw->act = s.act;
        cw->cout_pad = round_up(s.cout_dev, 16);
        cw->ktot = s.k * s.k * s.cin_dev;
        std::vector<float> bias(cw->cout_pad, 0.f);
        std::copy(bi->second.data.begin(), bi->second.data.end(), bias.begin());
        ZL_CUDA(cudaMalloc(&cw->bias, bias.size() * 4));
        ZL_CUDA(cudaMemcpy(cw->bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
        const bool need_simt = !bf16 || s.cin == 3;
        if (need_simt) {
            std::vector<float> ws((size_t)cw->ktot * cw->cout_pad, 0.f);
            for (int o = 0; o < s.cout; ++o)
                for (int ci = 0; ci < s.cin; ++ci)
                    for (int r = 0; r < s.k; ++r)
                        for (int q = 0; q < s.k; ++q)
                            ws[((size_t)(r * s.k + q) * s.cin_dev + ci) * cw->cout_pad + o] = W.data[(((size_t)o * s.cin + ci) * s.k + r) * s.k + q];
            ZL_CUDA(cudaMalloc(&cw->w_simt, ws.size() * 4));
            ZL_CUDA(cudaMemcpy(cw->w_simt, ws.data(), ws.size() * 4, cudaMemcpyHostToDevice));
        }
        if (bf16 && s.cin == 3) {
            // tensor-core first layer on the space-to-depth image: the 3x3 stride-2 filter becomes a 2x2 filter over
            // 2x2 pixel blocks.  K = 4 taps x 16: k = (ty*2+tx)*16 + (dy*2+dx)*3 + channel, where block tap (ty,tx) sits at
            // block offset (ty-1, tx-1) and filter row r = 2*ty + dy - 1, column q = 2*tx + dx - 1 (outside 0..2 -> 0).
            std::vector<uint16_t> wt((size_t)cw->cout_pad * 64, 0);
            for (int o = 0; o < s.cout; ++o)
                for (int ci = 0; ci < 3; ++ci)
                    for (int ty = 0; ty < 2; ++ty) for (int dy = 0; dy < 2; ++dy)
                        for (int tx = 0; tx < 2; ++tx) for (int dx = 0; dx < 2; ++dx) {
                            const int r = 2 * ty + dy - 1, q = 2 * tx + dx - 1;
                            if (r < 0 || q < 0) continue;
                            const float v = W.data[(((size_t)o * 3 + ci) * 3 + r) * 3 + q];
                            wt[(size_t)o * 64 + (ty * 2 + tx) * 16 + (dy * 2 + dx) * 3 + ci] = f16 ? f2h(v) : f2bf(v);
                        }
            ZL_CUDA(cudaMalloc(&cw->w_tc, wt.size() * 2));
            ZL_CUDA(cudaMemcpy(cw->w_tc, wt.data(), wt.size() * 2, cudaMemcpyHostToDevice));
        }
        if (bf16 && s.cin != 3) {
            std::vector<uint16_t> wt((size_t)cw->cout_pad * cw->ktot, 0);
            for (int o = 0; o < s.cout; ++o)
                for (int ci = 0; ci < s.cin; ++ci)
                    for (int r = 0; r < s.k; ++r)
                        for (int q = 0; q < s.k; ++q)
                            wt[(size_t)o * cw->ktot + (size_t)(r * s.k + q) * s.cin_dev + ci] = f16 ? f2h(W.data[(((size_t)o * s.cin + ci) * s.k + r) * s.k + q]) : f2bf(W.data[(((size_t)o * s.cin + ci) * s.k + r) * s.k + q]);
            ZL_CUDA(cudaMalloc(&cw->w_tc, wt.size() * 2));
            ZL_CUDA(cudaMemcpy(cw->w_tc, wt.data(), wt.size() * 2, cudaMemcpyHostToDevice));
        }
        new_by_name[s.name] = cw.get();
        new_convs.push_back(std::move(cw));
    }
    // fused Detect stems (16-bit modes): rows of cv2.l.0 followed by rows of cv3.l.0, one [cb + cc][9 * cin] tensor
    if (bf16 && md.cb + md.ccd <= 256) {
        for (int l = 0; l < 3; ++l) {
            const std::string sl = std::to_string(l);
            const ConvWeights* a = new_by_name["model.22.cv2." + sl + ".0.conv"];
            const ConvWeights* b = new_by_name["model.22.cv3." + sl + ".0.conv"];
            if (!a || !b || !a->w_tc || !b->w_tc || a->ktot != b->ktot) continue;
            std::unique_ptr<ConvWeights> cw(new ConvWeights());
            cw->name = "model.22.stem." + sl; cw->cin = a->cin; cw->cout = a->cout + b->cout; cw->k = 3; cw->stride = 1; cw->act = 1;
            cw->cout_pad = round_up(cw->cout, 16); cw->ktot = a->ktot;
            ZL_CUDA(cudaMalloc(&cw->w_tc, (size_t)cw->cout_pad * cw->ktot * 2));
            ZL_CUDA(cudaMemset(cw->w_tc, 0, (size_t)cw->cout_pad * cw->ktot * 2));
            ZL_CUDA(cudaMemcpy(cw->w_tc, a->w_tc, (size_t)a->cout * a->ktot * 2, cudaMemcpyDeviceToDevice));
            ZL_CUDA(cudaMemcpy(reinterpret_cast<char*>(cw->w_tc) + (size_t)a->cout * a->ktot * 2, b->w_tc, (size_t)b->cout * b->ktot * 2, cudaMemcpyDeviceToDevice));
            ZL_CUDA(cudaMalloc(&cw->bias, (size_t)cw->cout_pad * 4));
            ZL_CUDA(cudaMemset(cw->bias, 0, (size_t)cw->cout_pad * 4));
            ZL_CUDA(cudaMemcpy(cw->bias, a->bias, (size_t)a->cout * 4, cudaMemcpyDeviceToDevice));
            ZL_CUDA(cudaMemcpy(cw->bias + a->cout, b->bias, (size_t)b->cout * 4, cudaMemcpyDeviceToDevice));
            new_by_name[cw->name] = cw.get();
            new_convs.push_back(std::move(cw));
        }
    }
    ZL_CUDA(cudaDeviceSynchronize());           // uploads done: the set is complete
    pending_convs.swap(new_convs);              // a previously prepared, never committed set is dropped here
    pending_by_name.swap(new_by_name);
    return ZL_OK;
}

// ------------------------------------------------------------------ lane buffers
int Engine::inline_dets(int B) const { return std::min<int>(B * 64, B * num_anchors); }
size_t Engine::wire_inline_bytes(int B) const { return (size_t)kWireHeader * B + (size_t)kWireDet * inline_dets(B); }

int32_t Engine::alloc_lane(Lane& L)
{
    ZL_CUDA(cudaSetDevice(cfg.device));
    ZL_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
    ZL_CUDA(cudaEventCreate(&L.ev0));
    ZL_CUDA(cudaEventCreate(&L.ev1));
    for (int i = 0; i < 6; ++i) {
        ZL_CUDA(cudaStreamCreateWithFlags(&L.side[i], cudaStreamNonBlocking));
        ZL_CUDA(cudaEventCreateWithFlags(&L.ev_dep[i], cudaEventDisableTiming));
        ZL_CUDA(cudaEventCreateWithFlags(&L.ev_join[i], cudaEventDisableTiming));
    }
    const int MB = cfg.max_batch, H = cfg.model_h, W = cfg.model_w, nc = md.nc, A = num_anchors;
    const int adt = cfg.precision == ZL_PRECISION_BF16 ? DT_BF16 : (cfg.precision == ZL_PRECISION_FP16 ? DT_F16 : DT_F32);
    const size_t es = adt == DT_F32 ? 4 : 2;

    struct Req { std::string name; int h, w, c, dtype; };
    std::vector<Req> reqs;
    auto add = [&](const std::string& n, int h, int w, int c, int dt) { reqs.push_back({n, h, w, c, dt}); };
    const int* c = md.c;
    add("X0", H, W, 4, adt);
    if (adt != DT_F32) add("X0S", H / 2, W / 2, 16, adt);       // space-to-depth image for the tensor-core first layer
    add("A0", H / 2, W / 2, c[0], adt);
    add("A1", H / 4, W / 4, c[1], adt);
    add("CAT2", H / 4, W / 4, (2 + md.n[0]) * c[1] / 2, adt);  add("T2", H / 4, W / 4, c[1] / 2, adt);  add("A2", H / 4, W / 4, c[1], adt);
    add("A3", H / 8, W / 8, c[2], adt);
    add("CAT4", H / 8, W / 8, (2 + md.n[1]) * c[2] / 2, adt);  add("T4", H / 8, W / 8, c[2] / 2, adt);
    add("CAT14", H / 8, W / 8, c[3] + c[2], adt);              // [ up(h12) | p3 ]
    add("A5", H / 16, W / 16, c[3], adt);
    add("CAT6", H / 16, W / 16, (2 + md.n[2]) * c[3] / 2, adt); add("T6", H / 16, W / 16, c[3] / 2, adt);
    add("CAT11", H / 16, W / 16, c[4] + c[3], adt);            // [ up(p5) | p4 ]
    add("A7", H / 32, W / 32, c[4], adt);
    add("CAT8", H / 32, W / 32, (2 + md.n[3]) * c[4] / 2, adt); add("T8", H / 32, W / 32, c[4] / 2, adt);  add("A8", H / 32, W / 32, c[4], adt);
    add("CAT9", H / 32, W / 32, 2 * c[4], adt);
    add("CAT20", H / 32, W / 32, c[3] + c[4], adt);            // [ conv19 | p5 ]
    add("CAT12", H / 16, W / 16, (2 + md.nh) * c[3] / 2, adt); add("T12", H / 16, W / 16, c[3] / 2, adt);
    add("CAT17", H / 16, W / 16, c[2] + c[3], adt);            // [ conv16 | h12 ]
    add("CAT15", H / 8, W / 8, (2 + md.nh) * c[2] / 2, adt);   add("T15", H / 8, W / 8, c[2] / 2, adt);   add("O3", H / 8, W / 8, c[2], adt);
    add("CAT18", H / 16, W / 16, (2 + md.nh) * c[3] / 2, adt); add("T18", H / 16, W / 16, c[3] / 2, adt); add("O4", H / 16, W / 16, c[3], adt);
    add("CAT21", H / 32, W / 32, (2 + md.nh) * c[4] / 2, adt); add("T21", H / 32, W / 32, c[4] / 2, adt); add("O5", H / 32, W / 32, c[4], adt);
    const int ncp = round_up(nc, 4);
    for (int l = 0; l < 3; ++l) {
        const int hh = H / (8 << l), ww = W / (8 << l);
        const std::string s = std::to_string(l);
        // HBC1 = [ cv2.l.0 output (cb) | cv3.l.0 output (cc) ]: the two Detect stems read the same map, so in the 16-bit modes
        // they run as ONE conv of width cb + cc (one wide MMA costs far less than two narrow ones: umma_probe)
        add("HBC1_" + s, hh, ww, md.cb + md.ccd, adt); add("HB2_" + s, hh, ww, md.cb, adt); add("BOX_" + s, hh, ww, 64, DT_F32);
        add("HC2_" + s, hh, ww, md.ccd, adt); add("CLS_" + s, hh, ww, ncp, DT_F32);
    }
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t total = 0;
    for (auto& r : reqs) total += al((size_t)MB * r.h * r.w * r.c * (r.dtype == DT_F32 ? 4 : es));
    int key_pitch = 1;
    while (key_pitch < A) key_pitch <<= 1;
    const uint32_t det_cap = (uint32_t)MB * A;
    const size_t raw_b = al((size_t)MB * (4 + nc) * A * 4), keys_b = al((size_t)MB * key_pitch * 8), box_b = al((size_t)MB * A * 16),
                 cnt_b = al((size_t)MB * 4), hdr_b = al((size_t)(4 + 2 * MB) * 4), det_b = al((size_t)det_cap * sizeof(DevDet)),
                 desc_b = al((size_t)MB * sizeof(FrameDesc)), rdesc_b = al((size_t)4 * MB * sizeof(FrameDesc));
    total += raw_b + keys_b + 2 * box_b + cnt_b + hdr_b + det_b + desc_b + rdesc_b;
    ZL_CUDA(cudaMalloc(&L.arena, total));
    ZL_CUDA(cudaMemset(L.arena, 0, total));
    L.arena_bytes = total;
    char* p = L.arena;
    for (auto& r : reqs) {
        View v;
        v.ptr = p; v.n = MB; v.h = r.h; v.w = r.w; v.c = r.c; v.pitch = r.c; v.dtype = r.dtype;
        L.bufs[r.name] = v;
        p += al((size_t)MB * r.h * r.w * r.c * (r.dtype == DT_F32 ? 4 : es));
    }
    L.bufs["X0"].c = 3;   // three channels live in a pitch-4 pixel
    L.raw = (float*)p; p += raw_b;
    L.pb.keys = (uint64_t*)p; p += keys_b;
    L.pb.key_pitch = key_pitch;
    L.pb.box_by_anchor = (float4*)p; p += box_b;
    L.pb.sorted_box = (float4*)p; p += box_b;
    L.pb.cand_count = (uint32_t*)p; p += cnt_b;
    // result block: header {total, pad[3], cnt[MB], off[MB]} immediately followed by the records, so that the header and the
    // inline window of records leave in ONE device-to-host copy (one graph node less on the latency path)
    L.pb.header = (uint32_t*)p;
    L.pb.dets = (DevDet*)(p + (size_t)(4 + 2 * MB) * 4);
    p += hdr_b + det_b;
    L.pb.maxn = MB;
    L.pb.cap = det_cap;
    L.d_descs = (FrameDesc*)p; p += desc_b;
    L.d_res_descs = (FrameDesc*)p; p += rdesc_b;

    int a0 = 0;
    for (int l = 0; l < 3; ++l) {
        const std::string s = std::to_string(l);
        HeadLevel& hl = L.levels[l];
        hl.box = (const float*)L.bufs["BOX_" + s].ptr;
        hl.cls = (const float*)L.bufs["CLS_" + s].ptr;
        hl.h = H / (8 << l); hl.w = W / (8 << l); hl.stride = 8 << l; hl.cls_pitch = ncp; hl.a0 = a0;
        a0 += hl.h * hl.w;
    }
    L.staging_slots = MB;
    ZL_CUDA(cudaMalloc(&L.staging, L.staging_slots * slot_bytes()));
    ZL_CUDA(cudaMemset(L.staging, 128, L.staging_slots * slot_bytes()));
    ZL_CUDA(cudaHostAlloc(&L.h_descs, sizeof(FrameDesc) * MB * 5, cudaHostAllocDefault));
    L.h_result_bytes = (size_t)(4 + 2 * MB) * 4 + (size_t)det_cap * sizeof(DevDet);
    ZL_CUDA(cudaHostAlloc(&L.h_result, L.h_result_bytes, cudaHostAllocDefault));
    ZL_CUDA(cudaHostAlloc(&L.h_frames, (size_t)MB * slot_bytes(), cudaHostAllocDefault));
    if (cfg.emit_wire) {
        // N3: the reference's result wire layout, written by the device (postprocess.cu wire_pack_kernel)
        L.wire_cap = (size_t)kWireHeader * MB + (size_t)kWireDet * det_cap;
        ZL_CUDA(cudaMalloc(&L.d_wire, L.wire_cap));
        ZL_CUDA(cudaMalloc(&L.d_wire_off, sizeof(uint32_t) * (MB + 1)));
        ZL_CUDA(cudaMalloc(&L.d_wmeta, sizeof(WireMeta) * (MB + 1)));
        ZL_CUDA(cudaMemset(L.d_wmeta, 0, sizeof(WireMeta) * (MB + 1)));
        ZL_CUDA(cudaHostAlloc(&L.h_wire, L.wire_cap, cudaHostAllocDefault));
        ZL_CUDA(cudaHostAlloc(&L.h_wire_off, sizeof(uint32_t) * (MB + 1), cudaHostAllocDefault));
        ZL_CUDA(cudaHostAlloc(&L.h_wmeta, sizeof(WireMeta) * (MB + 1), cudaHostAllocDefault));
        std::memset(L.h_wmeta, 0, sizeof(WireMeta) * (MB + 1));
    }
    return ZL_OK;
}

void Engine::free_lane(Lane& L)
{
    for (auto& kv : L.graphs) cudaGraphExecDestroy(kv.second);
    L.graphs.clear();
    if (L.arena) cudaFree(L.arena);
    if (L.staging) cudaFree(L.staging);
    if (L.h_descs) cudaFreeHost(L.h_descs);
    if (L.h_result) cudaFreeHost(L.h_result);
    if (L.h_frames) cudaFreeHost(L.h_frames);
    if (L.d_wire) cudaFree(L.d_wire);
    if (L.d_wire_off) cudaFree(L.d_wire_off);
    if (L.d_wmeta) cudaFree(L.d_wmeta);
    if (L.h_wire) cudaFreeHost(L.h_wire);
    if (L.h_wire_off) cudaFreeHost(L.h_wire_off);
    if (L.h_wmeta) cudaFreeHost(L.h_wmeta);
    L.d_wire = nullptr; L.d_wire_off = nullptr; L.d_wmeta = nullptr; L.h_wire = nullptr; L.h_wire_off = nullptr; L.h_wmeta = nullptr;
    if (L.ev0) cudaEventDestroy(L.ev0);
    if (L.ev1) cudaEventDestroy(L.ev1);
    for (int i = 0; i < 6; ++i) {
        if (L.side[i]) cudaStreamDestroy(L.side[i]);
        if (L.ev_dep[i]) cudaEventDestroy(L.ev_dep[i]);
        if (L.ev_join[i]) cudaEventDestroy(L.ev_join[i]);
        L.side[i] = nullptr; L.ev_dep[i] = nullptr; L.ev_join[i] = nullptr;
    }
    if (L.stream) cudaStreamDestroy(L.stream);
    L.arena = nullptr; L.staging = nullptr; L.h_descs = nullptr; L.h_result = nullptr; L.h_frames = nullptr;
}

// ------------------------------------------------------------------ op list for batch size B
int32_t Engine::build_ops(Lane& L, int B)
{
    if (!weights_loaded) ZL_FAIL(ZL_NOT_INITIALIZED, "weights not loaded");
    std::vector<Op> ops;
    const bool bf16 = cfg.precision != ZL_PRECISION_FP32;   // "bf16" == any 16-bit tensor-core mode
    const bool f16 = cfg.precision == ZL_PRECISION_FP16; (void)f16;
    auto buf = [&](const std::string& n) { return L.bufs.at(n).with_n(B); };
    int32_t rc = ZL_OK;

    auto conv = [&](const std::string& name, const View& x, const View& y, const View* res) {
        if (rc != ZL_OK) return;
        auto it = conv_by_name.find(name);
        if (it == conv_by_name.end()) { set_error("no weights for " + name); rc = ZL_MODEL_LOAD_FAILED; return; }
        Op op;
        op.name = name; op.w = it->second; op.x = x; op.y = y;
        if (res) { op.res = *res; op.has_res = true; }
        const ConvWeights& w = *it->second;
        op.flops = 2.0 * (double)y.pixels() * w.cout * w.ktot;
        op.bytes = (double)x.pixels() * w.cin * x.esize() + (double)y.pixels() * w.cout * y.esize() +
                   (double)w.cout * w.ktot * (bf16 ? 2 : 4) + (res ? (double)y.pixels() * w.cout * res->esize() : 0.0);
        if (!bf16) op.kind = Op::CONV_SIMT;
        else if (w.cin == 3) op.kind = Op::CONV0;
        else if (int units = 0; use_halo && conv_halo_supported(w, x, y, num_sms, &units) &&
                 (units >= persist_min_units || (deep_k_persist && w.k * w.k * (w.cin / 16) >= 72))) {
            // stride-1 layers with at least one work unit per SM: persistent weights-resident kernel
            // (3x3: input read ~1.4x instead of 9x; 1x1: flattened pixel tiles)
            op.kind = Op::CONV_HALO;
            rc = conv_halo_prepare(w, x, y, res, num_sms, &op.halo);
        } else {
            op.kind = Op::CONV_TC;
            // small problems (latency path): split Cout over more CTAs so more than a handful of SMs work
            int hint = 0;
            const int mtiles = ceil_div((int)y.pixels(), 128);
            if (mtiles < 64 && w.cout_pad >= 64) hint = w.cout_pad % 32 == 0 ? 32 : 16;
            rc = conv_tc_prepare(w, x, y, res, true, hint, &op.tc);
        }
        ops.push_back(std::move(op));
    };
    auto c2f = [&](int idx, const View& x, const std::string& cat_n, const std::string& tmp_n, const View& out, int n, bool shortcut) {
        const std::string b = "model." + std::to_string(idx);
        View cat = buf(cat_n), tmp = buf(tmp_n);
        const int cw = tmp.c;
        conv(b + ".cv1.conv", x, cat.slice(0, 2 * cw), nullptr);
        for (int j = 0; j < n; ++j) {
            View in = cat.slice((j + 1) * cw, cw), o = cat.slice((j + 2) * cw, cw);
            conv(b + ".m." + std::to_string(j) + ".cv1.conv", in, tmp, nullptr);
            conv(b + ".m." + std::to_string(j) + ".cv2.conv", tmp, o, shortcut ? &in : nullptr);
        }
        conv(b + ".cv2.conv", cat, out, nullptr);
    };
    const int* c = md.c;

    if (bf16 && use_halo && use_stem && conv_by_name.count("model.0.conv") && conv_by_name["model.0.conv"]->cout_pad <= 64) {
        // 16-bit modes: the preprocess kernel writes the space-to-depth image and layer 0 runs as a 2x2 conv on it
        // through the persistent TMA + tcgen05 kernel, like every other layer
        { Op op; op.kind = Op::PRE; op.name = "preprocess"; op.y = buf("X0S");
          op.bytes = (double)B * cfg.model_w * cfg.model_h * (3 + 8); ops.push_back(op); }
        Op op; op.kind = Op::CONV_HALO; op.name = "model.0.conv"; op.w = conv_by_name["model.0.conv"]; op.x = buf("X0S"); op.y = buf("A0");
        ZL_TRY(conv_s2d_prepare(*op.w, op.x, op.y, num_sms, &op.halo));
        op.flops = op.halo.flops; op.bytes = op.halo.bytes;
        ops.push_back(op);
    } else if (bf16 && fuse_pre) {
        // 16-bit modes: preprocessing is fused into layer 0 (the preprocessed image is never written)
        auto it0 = conv_by_name.find("model.0.conv");
        if (it0 == conv_by_name.end()) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "no weights for model.0.conv");
        Op op; op.kind = Op::PRE_CONV0; op.name = "preprocess+model.0.conv"; op.w = it0->second; op.y = buf("A0");
        op.flops = 2.0 * (double)op.y.pixels() * op.w->cout * 27;
        op.bytes = (double)B * cfg.model_w * cfg.model_h * 3 + (double)op.y.pixels() * op.w->cout * 2;
        ops.push_back(op);
    } else {
        Op op; op.kind = Op::PRE; op.name = "preprocess"; op.y = buf("X0");
        op.bytes = (double)B * cfg.model_w * cfg.model_h * (3 + 4 * (bf16 ? 2 : 4)); ops.push_back(op);
        conv("model.0.conv", buf("X0"), buf("A0"), nullptr);
    }
    conv("model.1.conv", buf("A0"), buf("A1"), nullptr);
    c2f(2, buf("A1"), "CAT2", "T2", buf("A2"), md.n[0], true);
    conv("model.3.conv", buf("A2"), buf("A3"), nullptr);
    View p3 = buf("CAT14").slice(c[3], c[2]);
    c2f(4, buf("A3"), "CAT4", "T4", p3, md.n[1], true);
    conv("model.5.conv", p3, buf("A5"), nullptr);
    View p4 = buf("CAT11").slice(c[4], c[3]);
    c2f(6, buf("A5"), "CAT6", "T6", p4, md.n[2], true);
    conv("model.7.conv", p4, buf("A7"), nullptr);
    c2f(8, buf("A7"), "CAT8", "T8", buf("A8"), md.n[3], true);
    {   // SPPF
        View cat = buf("CAT9");
        const int h = c[4] / 2;
        conv("model.9.cv1.conv", buf("A8"), cat.slice(0, h), nullptr);
        Op op; op.kind = Op::POOL; op.name = "model.9.pool";
        op.x = cat.slice(0, h); op.p1 = cat.slice(h, h); op.p2 = cat.slice(2 * h, h); op.p3 = cat.slice(3 * h, h);
        op.bytes = (double)op.x.pixels() * h * op.x.esize() * 4;
        ops.push_back(op);
        View p5 = buf("CAT20").slice(c[3], c[4]);
        conv("model.9.cv2.conv", cat, p5, nullptr);
        Op up; up.kind = Op::UPSAMPLE; up.name = "model.10.upsample"; up.x = p5; up.y = buf("CAT11").slice(0, c[4]);
        up.bytes = (double)up.y.pixels() * c[4] * up.y.esize() * 1.25;
        ops.push_back(up);
    }
    View h12 = buf("CAT17").slice(c[2], c[3]);
    c2f(12, buf("CAT11"), "CAT12", "T12", h12, md.nh, false);
    { Op up; up.kind = Op::UPSAMPLE; up.name = "model.13.upsample"; up.x = h12; up.y = buf("CAT14").slice(0, c[3]);
      up.bytes = (double)up.y.pixels() * c[3] * up.y.esize() * 1.25; ops.push_back(up); }
    c2f(15, buf("CAT14"), "CAT15", "T15", buf("O3"), md.nh, false);
    if (rc == ZL_OK) ops.back().record_ev = 0;                       // O3 ready: level 0's head may start
    conv("model.16.conv", buf("O3"), buf("CAT17").slice(0, c[2]), nullptr);
    c2f(18, buf("CAT17"), "CAT18", "T18", buf("O4"), md.nh, false);
    if (rc == ZL_OK) ops.back().record_ev = 1;
    conv("model.19.conv", buf("O4"), buf("CAT20").slice(0, c[3]), nullptr);
    c2f(21, buf("CAT20"), "CAT21", "T21", buf("O5"), md.nh, false);
    if (rc == ZL_OK) ops.back().record_ev = 2;
    const char* outs[3] = {"O3", "O4", "O5"};
    for (int l = 0; l < 3; ++l) {
        const std::string s = std::to_string(l), b = "model.22.cv2." + s;
        const bool fused = bf16 && fuse_stems && conv_by_name.count("model.22.stem." + s) != 0;
        if (fused) conv("model.22.stem." + s, buf(outs[l]), buf("HBC1_" + s), nullptr);
        else conv(b + ".0.conv", buf(outs[l]), buf("HBC1_" + s).slice(0, md.cb), nullptr);
        // head branches (fused stems only): stem + box branch on side stream 2l, class branch on 2l+1 after the stem
        if (rc == ZL_OK && fused) { Op& o = ops.back(); o.side = 1 + 2 * l; o.wait_ev = l; o.record_ev = 3 + l; }
        conv(b + ".1.conv", buf("HBC1_" + s).slice(0, md.cb), buf("HB2_" + s), nullptr);
        if (rc == ZL_OK && fused) ops.back().side = 1 + 2 * l;
        conv(b + ".2", buf("HB2_" + s), buf("BOX_" + s), nullptr);
        if (rc == ZL_OK) { ops.back().path = 3; if (fused) ops.back().side = 1 + 2 * l; }
    }
    for (int l = 0; l < 3; ++l) {
        const std::string s = std::to_string(l), b = "model.22.cv3." + s;
        const bool fused = bf16 && fuse_stems && conv_by_name.count("model.22.stem." + s) != 0;
        if (!fused) conv(b + ".0.conv", buf(outs[l]), buf("HBC1_" + s).slice(md.cb, md.ccd), nullptr);
        conv(b + ".1.conv", buf("HBC1_" + s).slice(md.cb, md.ccd), buf("HC2_" + s), nullptr);
        if (rc == ZL_OK && fused) { Op& o = ops.back(); o.side = 2 + 2 * l; o.wait_ev = 3 + l; }
        View cls = buf("CLS_" + s); cls.c = md.nc;
        conv(b + ".2", buf("HC2_" + s), cls, nullptr);
        if (rc == ZL_OK) { ops.back().path = 3; if (fused) ops.back().side = 2 + 2 * l; }
    }
    if (rc != ZL_OK) return rc;
    if (bf16) {
        // the last 1x1 convs of both branches + Detect tail + decode/threshold as one kernel (head_fused.cu)
        const ConvWeights* wb[3] = {nullptr, nullptr, nullptr};
        const ConvWeights* wc[3] = {nullptr, nullptr, nullptr};
        View xb[3], xc[3];
        for (int l = 0; l < 3; ++l) {
            const std::string s = std::to_string(l);
            auto ib = conv_by_name.find("model.22.cv2." + s + ".2"), ic = conv_by_name.find("model.22.cv3." + s + ".2");
            if (ib != conv_by_name.end()) wb[l] = ib->second;
            if (ic != conv_by_name.end()) wc[l] = ic->second;
            xb[l] = buf("HB2_" + s); xc[l] = buf("HC2_" + s);
        }
        if (head_fused_supported(wb, wc, xb, xc, md.nc)) {
            Op op; op.kind = Op::HEAD_FUSED; op.name = "head.2+decode+filter"; op.path = 4; op.join = 1;
            ZL_TRY(head_fused_prepare(wb, wc, xb, xc, L.levels, md.nc, num_anchors, num_sms, &op.hf));
            op.flops = op.hf.flops; op.bytes = op.hf.bytes;
            ops.push_back(op);
        }
    }
    const double rawb = (double)B * (4 + md.nc) * num_anchors * 4;
    { Op op; op.kind = Op::DECODE; op.name = "dfl_decode"; op.path = 1; op.join = 1; op.bytes = (double)B * num_anchors * (64 + md.nc) * 4 + rawb; ops.push_back(op); }
    { Op op; op.kind = Op::FILTER; op.name = "filter"; op.path = 1; op.bytes = rawb; ops.push_back(op); }
    { Op op; op.kind = Op::DECODE_FILTER; op.name = "decode+filter"; op.path = 2; op.join = 1; op.bytes = (double)B * num_anchors * md.nc * 4; ops.push_back(op); }
    { Op op; op.kind = Op::NMS; op.name = "nms"; op.bytes = 0; ops.push_back(op); }
    L.ops[B] = std::move(ops);
    return ZL_OK;
}

// ------------------------------------------------------------------ execution
int32_t Engine::launch_op(Lane& L, int B, const Op& op, cudaStream_t on)
{
    cudaStream_t st = on ? on : L.stream;
    const bool bf16 = cfg.precision != ZL_PRECISION_FP32;   // any 16-bit tensor-core mode
    const bool f16 = cfg.precision == ZL_PRECISION_FP16;
    switch (op.kind) {
        case Op::PRE:
            if (op.y.c == 16) return launch_preprocess(st, L.staging, L.d_descs, B, cfg.model_w, cfg.model_h, f16 ? PRE_S2D16_F16 : PRE_S2D16_BF16, op.y.ptr, cfg.preprocess_mode == ZL_PRE_LETTERBOX, L.same_size);
            return launch_preprocess(st, L.staging, L.d_descs, B, cfg.model_w, cfg.model_h, bf16 ? (f16 ? PRE_NHWC4_F16 : PRE_NHWC4_BF16) : PRE_NHWC4_F32, op.y.ptr, cfg.preprocess_mode == ZL_PRE_LETTERBOX);
        case Op::CONV_TC: return conv_tc_launch(st, op.tc);
        case Op::CONV_HALO: return conv_halo_launch(st, op.halo, num_sms);
        case Op::CONV_SIMT: return launch_conv_simt(st, *op.w, op.x, op.y, op.has_res ? &op.res : nullptr);
        case Op::CONV0: return launch_conv0_direct(st, *op.w, op.x, op.y);
        case Op::PRE_CONV0: return launch_pre_conv0(st, L.staging, L.d_descs, B, cfg.model_w, cfg.model_h, *op.w, op.y);
        case Op::POOL: return launch_sppf_pool(st, op.x, op.p1, op.p2, op.p3);
        case Op::UPSAMPLE: return launch_upsample2x(st, op.x, op.y);
        case Op::DECODE: return launch_dfl_decode(st, L.levels, B, md.nc, num_anchors, L.raw, !bf16);
        case Op::FILTER: return launch_filter(st, L.raw, B, md.nc, num_anchors, L.d_descs, nullptr, cfg.conf_threshold, d_class_weights, L.pb);
        case Op::DECODE_FILTER:
            return launch_decode_filter(st, L.levels, B, md.nc, num_anchors, L.d_descs, cfg.conf_threshold, d_class_weights, L.pb, !bf16);
        case Op::HEAD_FUSED: return head_fused_launch(st, op.hf, L.d_descs, cfg.conf_threshold, d_class_weights, L.pb);
        case Op::NMS: return launch_nms(st, B, num_anchors, cfg.iou_threshold, L.pb, false, L.nms_host_out);
    }
    return ZL_OK;
}

// Which ops of the list a pass runs (Op::path).
static bool head_is_fused(const std::vector<Op>& ops)
{
    for (const Op& op : ops) if (op.kind == Op::HEAD_FUSED) return true;
    return false;
}
static bool op_selected(const Op& op, bool want_raw, bool fused)
{
    switch (op.path) {
        case 1: return want_raw;
        case 2: return !want_raw && !fused;
        case 3: return want_raw || !fused;
        case 4: return !want_raw && fused;
        default: return true;
    }
}
static std::vector<Op> hot_ops(const std::vector<Op>& all)
{
    const bool fused = head_is_fused(all);
    std::vector<Op> ops;
    for (const Op& op : all) if (op_selected(op, false, fused)) ops.push_back(op);
    return ops;
}

// want_raw: materialise the raw head tensor (zl_forward_raw) with the two-kernel D1, F1 path; otherwise the fused
// decode+filter kernel runs and L.raw is not written.  Both produce the same candidates bit for bit.
int32_t Engine::run_ops(Lane& L, int B, bool with_d2h, bool want_raw)
{
    cudaStream_t st = L.stream;
    auto it = L.ops.find(B);
    if (it == L.ops.end()) { ZL_TRY(build_ops(L, B)); it = L.ops.find(B); }
    // candidate counts and the result total in ONE memset: the header sits right behind the counts (alloc_lane), and every
    // node in front of the first kernel is ~2 us on the b=1 path
    ZL_CUDA(cudaMemsetAsync(L.pb.cand_count, 0, (size_t)(reinterpret_cast<char*>(L.pb.header) - reinterpret_cast<char*>(L.pb.cand_count)) + sizeof(uint32_t) * 4, st));
    // b=1, plain results: the NMS CTA writes the result block into the pinned host buffer itself (no copy node at the end)
    static const bool host_direct = [] { const char* e = getenv("ZL_NMS_HOST_DIRECT"); return !(e && e[0] == '0'); }();
    L.nms_host_out = (with_d2h && B == 1 && host_direct && !cfg.emit_wire && cfg.preprocess_mode != ZL_PRE_LETTERBOX) ? reinterpret_cast<uint32_t*>(L.h_result) : nullptr;
    const bool fused = head_is_fused(it->second);
    // Small batches (the latency path): a level's Detect head — stem, box branch, class branch — runs on side streams as soon
    // as the level's map exists, next to the rest of the neck, instead of queueing behind it: at b=1 every kernel is a few
    // CTAs, so the branches really overlap (7 kernels leave the critical path).  Captured, this becomes a graph with forks.
    static const int branch_max_batch = [] { const char* e = getenv("ZL_HEAD_BRANCHES"); return e ? atoi(e) : 8; }();
    const bool branches = B <= branch_max_batch;
    bool dirty[6] = {false, false, false, false, false, false};
    for (const Op& op : it->second) {
        if (!op_selected(op, want_raw, fused)) continue;
        cudaStream_t s = (branches && op.side > 0) ? L.side[op.side - 1] : st;
        if (branches && op.join) {
            for (int k = 0; k < 6; ++k)
                if (dirty[k]) { ZL_CUDA(cudaEventRecord(L.ev_join[k], L.side[k])); ZL_CUDA(cudaStreamWaitEvent(st, L.ev_join[k], 0)); dirty[k] = false; }
        }
        if (branches && op.side > 0 && op.wait_ev >= 0) ZL_CUDA(cudaStreamWaitEvent(s, L.ev_dep[op.wait_ev], 0));
        ZL_TRY(launch_op(L, B, op, s));
        if (branches && op.side > 0) dirty[op.side - 1] = true;
        if (branches && op.record_ev >= 0) ZL_CUDA(cudaEventRecord(L.ev_dep[op.record_ev], s));
    }
    for (int k = 0; k < 6; ++k)      // nothing may be left un-joined (a list without a join op)
        if (dirty[k]) { ZL_CUDA(cudaEventRecord(L.ev_join[k], L.side[k])); ZL_CUDA(cudaStreamWaitEvent(st, L.ev_join[k], 0)); }
    if (cfg.preprocess_mode == ZL_PRE_LETTERBOX)      // non-parity mode: boxes back from the letterboxed model frame to the request frame
        ZL_TRY(launch_letterbox_unmap(st, B, L.pb.maxn, L.pb.header, L.pb.dets, L.d_descs, cfg.model_w, cfg.model_h, L.pb.cap));
    if (cfg.emit_wire) ZL_TRY(launch_wire_pack(st, B, L.pb, L.d_wmeta, L.d_wire, (uint32_t)std::min<size_t>(L.wire_cap, 0xffffffffu), L.d_wire_off));
    if (with_d2h && !L.nms_host_out) {
        // header (total, cnt[], off[]) and the first inline_dets records in one fixed-size copy
        const size_t hdr = (size_t)(4 + 2 * L.pb.maxn) * 4;
        ZL_CUDA(cudaMemcpyAsync(L.h_result, L.pb.header, hdr + (size_t)inline_dets(B) * sizeof(DevDet), cudaMemcpyDeviceToHost, st));   // dets follow the header (alloc_lane)
        if (cfg.emit_wire) {           // the wire blocks: offsets and the inline window, fixed sizes (graph-capturable)
            ZL_CUDA(cudaMemcpyAsync(L.h_wire_off, L.d_wire_off, sizeof(uint32_t) * (B + 1), cudaMemcpyDeviceToHost, st));
            ZL_CUDA(cudaMemcpyAsync(L.h_wire, L.d_wire, wire_inline_bytes(B), cudaMemcpyDeviceToHost, st));
        }
    }
    return ZL_OK;
}

int Engine::graph_batch_for(int n) const
{
    int b = 1;
    while (b < n) b <<= 1;
    return std::min(b, cfg.max_batch);
}

// One graph per (batch size, "every frame already has the model's size"): the second half picks the preprocess kernel.
static inline int graph_key(int B, bool same_size) { return B | (same_size ? 1 << 20 : 0); }

int32_t Engine::ensure_graph(Lane& L, int B)
{
    if (L.graphs.count(graph_key(B, L.same_size))) return ZL_OK;
    // one un-captured pass first: lazy module load, cudaFuncSetAttribute and op building stay out of the capture
    ZL_TRY(run_ops(L, B, true));
    ZL_CUDA(cudaStreamSynchronize(L.stream));
    cudaGraph_t g = nullptr;
    ZL_CUDA(cudaStreamBeginCapture(L.stream, cudaStreamCaptureModeThreadLocal));
    int32_t rc = run_ops(L, B, true);
    cudaError_t ce = cudaStreamEndCapture(L.stream, &g);
    if (rc != ZL_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (ce != cudaSuccess) ZL_FAIL(ZL_INFERENCE_ERROR, std::string("graph capture failed: ") + cudaGetErrorString(ce));
    cudaGraphExec_t ge = nullptr;
    ce = cudaGraphInstantiate(&ge, g, 0);
    cudaGraphDestroy(g);
    if (ce != cudaSuccess) ZL_FAIL(ZL_INFERENCE_ERROR, std::string("graph instantiate failed: ") + cudaGetErrorString(ce));
    L.graphs[graph_key(B, L.same_size)] = ge;
    { std::lock_guard<std::mutex> g2(smu); graph_captured++; }
    return ZL_OK;
}

int32_t Engine::launch_batch(Lane& L, int B, bool want_raw)
{
    if (want_raw) return run_ops(L, B, true, true);
    if (cfg.use_graph) {
        ZL_TRY(ensure_graph(L, B));
        ZL_CUDA(cudaGraphLaunch(L.graphs[graph_key(B, L.same_size)], L.stream));
        return ZL_OK;
    }
    return run_ops(L, B, true);
}

// Runs n (<= max_batch) frames on lane L.  Caller holds L.mu.
int32_t Engine::run_lane_batch(Lane& L, const uint8_t* const* frames, const int32_t* ws, const int32_t* hs, int n,
                               bool frames_pinned, std::vector<zl_det>* dets, int32_t* counts, bool want_raw, const WireReq* wr)
{
    if (wr && !cfg.emit_wire) ZL_FAIL(ZL_INVALID_ARGUMENT, "engine was created without emit_wire");
    ZL_CUDA(cudaSetDevice(cfg.device));
    const int B = graph_batch_for(n);
    const size_t slot = slot_bytes();
    // Frames that sit back to back in pinned host memory and fill their slots exactly (the bench's batches, the async
    // path's slot ring when consecutive) go up in ONE copy; anything else is one copy per frame.
    bool one_copy = frames_pinned && n > 1;
    bool same = true;
    for (int i = 0; i < n; ++i) {
        const size_t bytes = (size_t)ws[i] * hs[i] * 3;
        if (ws[i] <= 0 || hs[i] <= 0 || bytes > slot) ZL_FAIL(ZL_INVALID_INPUT, "frame larger than max_frame_w x max_frame_h");
        if (bytes != slot || (i > 0 && frames[i] != frames[i - 1] + slot)) one_copy = false;
        if (ws[i] != cfg.model_w || hs[i] != cfg.model_h) same = false;
    }
    L.same_size = same && (slot % 8) == 0;          // frames start at multiples of the slot size: the fast kernel's 8-byte loads stay aligned
    if (one_copy) ZL_CUDA(cudaMemcpyAsync(L.staging, frames[0], (size_t)n * slot, cudaMemcpyHostToDevice, L.stream));
    for (int i = 0; i < n; ++i) {
        const size_t bytes = (size_t)ws[i] * hs[i] * 3;
        if (!one_copy) {
            const uint8_t* src = frames[i];
            if (!frames_pinned) {
                std::memcpy(L.h_frames + (size_t)i * slot, frames[i], bytes);
                src = L.h_frames + (size_t)i * slot;
            }
            ZL_CUDA(cudaMemcpyAsync(L.staging + (size_t)i * slot, src, bytes, cudaMemcpyHostToDevice, L.stream));
        }
        L.h_descs[i] = FrameDesc{(uint64_t)i * slot, ws[i], hs[i]};
    }
    for (int i = n; i < B; ++i) L.h_descs[i] = L.h_descs[0];     // padding frames repeat frame 0; their results are ignored
    // the descriptors only change with the geometry of the batch: a stream of equal-sized frames (the serving case, and the
    // b=1 latency path, where every stream operation in front of the graph is ~2-3 us) sends them once
    if (L.descs_on_device.size() != (size_t)B || std::memcmp(L.descs_on_device.data(), L.h_descs, sizeof(FrameDesc) * B) != 0) {
        ZL_CUDA(cudaMemcpyAsync(L.d_descs, L.h_descs, sizeof(FrameDesc) * B, cudaMemcpyHostToDevice, L.stream));
        L.descs_on_device.assign(L.h_descs, L.h_descs + B);
    }
    if (cfg.emit_wire) {
        for (int i = 0; i < B; ++i) {
            const int k = i < n ? i : 0;
            L.h_wmeta[i] = WireMeta{wr && wr->frame_ids ? wr->frame_ids[k] : 0u, 0u, wr && wr->timestamps ? wr->timestamps[k] : 0ull};
        }
        L.h_wmeta[L.pb.maxn] = WireMeta{0u, 0u, wr ? wr->det_ts : 0ull};
        ZL_CUDA(cudaMemcpyAsync(L.d_wmeta, L.h_wmeta, sizeof(WireMeta) * (L.pb.maxn + 1), cudaMemcpyHostToDevice, L.stream));
    }
    // Device time of the step (getStatus: avg_device_time_ms) from a SAMPLE of the batches: the two event records around the
    // graph launch cost 9 us per call on the b=1 path (measured: p50 0.3546 -> 0.3457 ms without them), so only every
    // 256th batch of a lane (and its second) is timed.  ZL_TIMING_EVERY=1 times every batch.
    static const int timing_every = [] { const char* e = getenv("ZL_TIMING_EVERY"); const int v = e ? atoi(e) : 256; return v > 0 ? v : 256; }();
    const uint64_t kth = L.batches_run++;                 // the lane's first batch (graph capture inside) is never the sample
    const bool timed = kth == 1 || (kth >= (uint64_t)timing_every && kth % (uint64_t)timing_every == 0);
    if (timed) ZL_CUDA(cudaEventRecord(L.ev0, L.stream));
    ZL_TRY(launch_batch(L, B, want_raw));
    if (timed) ZL_CUDA(cudaEventRecord(L.ev1, L.stream));
    ZL_CUDA(cudaStreamSynchronize(L.stream));
    {
        float ms = 0;
        if (timed) cudaEventElapsedTime(&ms, L.ev0, L.ev1);
        std::lock_guard<std::mutex> g(smu);
        if (timed) { dev_ms_sum += ms; dev_ms_n++; }
        st_batches++;
    }

    const uint32_t* hdr = (const uint32_t*)L.h_result;
    const uint32_t total = hdr[0];
    const uint32_t* cnt = hdr + 4;
    const uint32_t* off = hdr + 4 + L.pb.maxn;
    const size_t hdr_b = (size_t)(4 + 2 * L.pb.maxn) * 4;
    if (total > L.pb.cap) ZL_FAIL(ZL_INFERENCE_ERROR, "detection buffer overflow");
    if (total > (uint32_t)inline_dets(B)) {       // rare: more survivors than the inline window
        ZL_CUDA(cudaMemcpyAsync(L.h_result + hdr_b, L.pb.dets, (size_t)total * sizeof(DevDet), cudaMemcpyDeviceToHost, L.stream));
        ZL_CUDA(cudaStreamSynchronize(L.stream));
    }
    if (wr) {
        const uint32_t wbytes = L.h_wire_off[n];                 // frames n..B-1 are padding: their blocks come after
        if (wbytes > L.wire_cap) ZL_FAIL(ZL_INFERENCE_ERROR, "wire buffer overflow");
        if (wbytes > wire_inline_bytes(B)) {
            ZL_CUDA(cudaMemcpyAsync(L.h_wire, L.d_wire, wbytes, cudaMemcpyDeviceToHost, L.stream));
            ZL_CUDA(cudaStreamSynchronize(L.stream));
        }
    }
    const zl_det* all = (const zl_det*)(L.h_result + hdr_b);
    static_assert(sizeof(zl_det) == sizeof(DevDet), "layout");
    dets->clear();
    for (int i = 0; i < n; ++i) {
        counts[i] = (int32_t)cnt[i];
        dets->insert(dets->end(), all + off[i], all + off[i] + cnt[i]);
    }
    return ZL_OK;
}

static bool is_pinned_query(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// Pinned-ness of a caller's frame buffers, remembered per thread: a serving loop hands the same few buffers in again and
// again, and cudaPointerGetAttributes costs microseconds per frame (64 of them per batch in round 1).  A batch whose first
// and last frame lie inside one remembered pinned range is pinned; ranges come from zl_host_alloc (registered exactly)
// or are learnt one frame at a time.
namespace {
struct PinRange { const uint8_t* lo; const uint8_t* hi; };
std::mutex g_pin_mu;
std::vector<PinRange> g_pin_ranges;
}
void register_pinned_range(const void* p, size_t bytes)
{
    std::lock_guard<std::mutex> g(g_pin_mu);
    g_pin_ranges.push_back(PinRange{(const uint8_t*)p, (const uint8_t*)p + bytes});
}
void unregister_pinned_range(const void* p)
{
    std::lock_guard<std::mutex> g(g_pin_mu);
    for (size_t i = 0; i < g_pin_ranges.size(); ++i)
        if (g_pin_ranges[i].lo == (const uint8_t*)p) { g_pin_ranges.erase(g_pin_ranges.begin() + i); return; }
}
static bool in_registered_range(const uint8_t* p, size_t bytes)
{
    std::lock_guard<std::mutex> g(g_pin_mu);
    for (const PinRange& r : g_pin_ranges)
        if (p >= r.lo && p + bytes <= r.hi) return true;
    return false;
}
static bool is_pinned(const void* p, size_t bytes)
{
    if (in_registered_range((const uint8_t*)p, bytes)) return true;
    return is_pinned_query(p);
}

int32_t Engine::infer_batch(const uint8_t* const* frames, const int32_t* ws, const int32_t* hs, int n,
                            zl_det* out, int cap, int32_t* counts, int32_t* offsets, float* raw_out)
{
    if (!weights_loaded) ZL_FAIL(ZL_NOT_INITIALIZED, "weights not loaded");
    if (n < 0 || (n > 0 && (!frames || !ws || !hs))) ZL_FAIL(ZL_INVALID_ARGUMENT, "null argument");
    // concurrent callers get different lanes (stream + buffers), so one caller's H2D copies overlap another's kernels
    Lane* Lp = nullptr;
    std::unique_lock<std::mutex> g;
    for (auto& cand : lanes) {
        std::unique_lock<std::mutex> t(cand->mu, std::try_to_lock);
        if (t.owns_lock()) { Lp = cand.get(); g = std::move(t); break; }
    }
    if (!Lp) {
        Lp = lanes[sync_rr.fetch_add(1) % lanes.size()].get();
        g = std::unique_lock<std::mutex>(Lp->mu);
    }
    Lane& L = *Lp;
    std::vector<zl_det> dets;
    int total = 0;
    bool overflow = false;
    for (int i0 = 0; i0 < n; i0 += cfg.max_batch) {
        const int nb = std::min(cfg.max_batch, n - i0);
        bool pinned = true;
        for (int i = 0; i < nb && pinned; ++i) pinned = is_pinned(frames[i0 + i], (size_t)ws[i0 + i] * hs[i0 + i] * 3);
        std::vector<int32_t> cnt(nb);
        ZL_TRY(run_lane_batch(L, frames + i0, ws + i0, hs + i0, nb, pinned, &dets, cnt.data(), raw_out != nullptr));
        if (raw_out) {
            const size_t per = (size_t)(4 + md.nc) * num_anchors;
            ZL_CUDA(cudaMemcpy(raw_out + (size_t)i0 * per, L.raw, per * nb * 4, cudaMemcpyDeviceToHost));
        }
        size_t k = 0;
        for (int i = 0; i < nb; ++i) {
            if (counts) counts[i0 + i] = cnt[i];
            if (offsets) offsets[i0 + i] = total;
            for (int j = 0; j < cnt[i]; ++j, ++k) {
                if (out && total < cap) out[total] = dets[k]; else if (out) overflow = true;
                total++;
            }
        }
    }
    { std::lock_guard<std::mutex> g2(smu); st_count += n; }
    if (overflow) ZL_FAIL(ZL_INSUFFICIENT_RESOURCES, "dets_out capacity too small");
    return ZL_OK;
}

int32_t Engine::infer_batch_wire(const uint8_t* const* frames, const int32_t* ws, const int32_t* hs, int n, const uint32_t* frame_ids,
                                 const uint64_t* timestamps, uint64_t det_ts, uint8_t* out, size_t cap, uint32_t* offsets)
{
    if (!weights_loaded) ZL_FAIL(ZL_NOT_INITIALIZED, "weights not loaded");
    if (!cfg.emit_wire) ZL_FAIL(ZL_INVALID_ARGUMENT, "engine was created without emit_wire");
    if (n < 0 || (n > 0 && (!frames || !ws || !hs)) || !offsets || (!out && cap)) ZL_FAIL(ZL_INVALID_ARGUMENT, "null argument");
    Lane* Lp = nullptr;
    std::unique_lock<std::mutex> g;
    for (auto& cand : lanes) {
        std::unique_lock<std::mutex> t(cand->mu, std::try_to_lock);
        if (t.owns_lock()) { Lp = cand.get(); g = std::move(t); break; }
    }
    if (!Lp) {
        Lp = lanes[sync_rr.fetch_add(1) % lanes.size()].get();
        g = std::unique_lock<std::mutex>(Lp->mu);
    }
    Lane& L = *Lp;
    std::vector<zl_det> dets;
    size_t total = 0;
    bool overflow = false;
    offsets[0] = 0;
    for (int i0 = 0; i0 < n; i0 += cfg.max_batch) {
        const int nb = std::min(cfg.max_batch, n - i0);
        bool pinned = true;
        for (int i = 0; i < nb && pinned; ++i) pinned = is_pinned(frames[i0 + i], (size_t)ws[i0 + i] * hs[i0 + i] * 3);
        std::vector<int32_t> cnt(nb);
        WireReq wr{frame_ids ? frame_ids + i0 : nullptr, timestamps ? timestamps + i0 : nullptr, det_ts};
        ZL_TRY(run_lane_batch(L, frames + i0, ws + i0, hs + i0, nb, pinned, &dets, cnt.data(), false, &wr));
        const uint32_t bytes = L.h_wire_off[nb];
        if (total + bytes <= cap) std::memcpy(out + total, L.h_wire, bytes); else overflow = true;
        for (int i = 0; i < nb; ++i) offsets[i0 + i + 1] = (uint32_t)(total + L.h_wire_off[i + 1]);
        total += bytes;
    }
    { std::lock_guard<std::mutex> g2(smu); st_count += n; }
    if (overflow) ZL_FAIL(ZL_INSUFFICIENT_RESOURCES, "wire output capacity too small");
    return ZL_OK;
}

int32_t Engine::preprocess_one(const uint8_t* bgr, int w, int h, size_t len, float* out_chw)
{
    if (!bgr || !out_chw) ZL_FAIL(ZL_INVALID_ARGUMENT, "null argument");
    if (w <= 0 || h <= 0 || len != (size_t)w * h * 3)      // onnx_engine.cpp:659-665
        ZL_FAIL(ZL_INVALID_INPUT, "Invalid image data size: expected " + std::to_string((size_t)w * h * 3) + ", got " + std::to_string(len));
    ZL_CUDA(cudaSetDevice(cfg.device));
    Lane& L = *lanes[0];
    std::lock_guard<std::mutex> g(L.mu);
    uint8_t* d_img = nullptr; float* d_out = nullptr; FrameDesc* d_desc = nullptr;
    const size_t ob = (size_t)3 * cfg.model_w * cfg.model_h * 4;
    ZL_CUDA(cudaMalloc(&d_img, len));
    ZL_CUDA(cudaMalloc(&d_out, ob));
    ZL_CUDA(cudaMalloc(&d_desc, sizeof(FrameDesc)));
    FrameDesc fd{0, w, h};
    int32_t rc = ZL_OK;
    cudaMemcpyAsync(d_img, bgr, len, cudaMemcpyHostToDevice, L.stream);
    cudaMemcpyAsync(d_desc, &fd, sizeof(fd), cudaMemcpyHostToDevice, L.stream);
    rc = launch_preprocess(L.stream, d_img, d_desc, 1, cfg.model_w, cfg.model_h, PRE_NCHW_F32, d_out, cfg.preprocess_mode == ZL_PRE_LETTERBOX);
    cudaError_t ce = cudaMemcpyAsync(out_chw, d_out, ob, cudaMemcpyDeviceToHost, L.stream);
    cudaError_t cs = cudaStreamSynchronize(L.stream);
    cudaFree(d_img); cudaFree(d_out); cudaFree(d_desc);
    if (rc != ZL_OK) return rc;
    if (ce != cudaSuccess || cs != cudaSuccess) ZL_FAIL(ZL_INFERENCE_ERROR, std::string("preprocess: ") + cudaGetErrorString(cs != cudaSuccess ? cs : ce));
    return ZL_OK;
}

int32_t Engine::decode_nms(const float* raw, int n, int nc, int A, const int32_t* iw, const int32_t* ih, float conf, float iou,
                           zl_det* out, int cap, int32_t* counts, int32_t* offsets,
                           int iters, float* ms_filter, float* ms_nms, int64_t* kept)
{
    if (!raw || n <= 0 || nc <= 0 || A <= 0 || !iw || !ih) ZL_FAIL(ZL_INVALID_ARGUMENT, "bad argument");
    if (A > kMaxAnchors || nc > kMaxClasses) ZL_FAIL(ZL_INVALID_ARGUMENT, "A or nc beyond supported range");
    ZL_CUDA(cudaSetDevice(cfg.device));
    Lane& L = *lanes[0];
    std::lock_guard<std::mutex> g(L.mu);
    int key_pitch = 1;
    while (key_pitch < A) key_pitch <<= 1;
    const size_t raw_b = (size_t)n * (4 + nc) * A * 4;
    float* d_raw = nullptr; int32_t* d_wh = nullptr; char* scratch = nullptr;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t keys_b = al((size_t)n * key_pitch * 8), box_b = al((size_t)n * A * 16), cnt_b = al((size_t)n * 4),
                 hdr_b = al((size_t)(4 + 2 * n) * 4), det_b = al((size_t)n * A * sizeof(DevDet));
    cudaError_t e1 = cudaMalloc(&d_raw, raw_b), e2 = cudaMalloc(&d_wh, (size_t)n * 8),
                e3 = cudaMalloc(&scratch, keys_b + 2 * box_b + cnt_b + hdr_b + det_b);
    auto cleanup = [&] { cudaFree(d_raw); cudaFree(d_wh); cudaFree(scratch); };
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { cleanup(); ZL_FAIL(ZL_INSUFFICIENT_RESOURCES, "decode_nms: out of device memory"); }
    PostBuffers pb{};
    char* p = scratch;
    pb.keys = (uint64_t*)p; p += keys_b; pb.key_pitch = key_pitch;
    pb.box_by_anchor = (float4*)p; p += box_b; pb.sorted_box = (float4*)p; p += box_b;
    pb.cand_count = (uint32_t*)p; p += cnt_b; pb.header = (uint32_t*)p; p += hdr_b; pb.dets = (DevDet*)p;
    pb.maxn = n; pb.cap = (uint32_t)n * A;
    std::vector<int32_t> wh(2 * n);
    for (int i = 0; i < n; ++i) { wh[2 * i] = iw[i]; wh[2 * i + 1] = ih[i]; }
    cudaStream_t st = L.stream;
    int32_t rc = ZL_OK;
    cudaMemcpyAsync(d_raw, raw, raw_b, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_wh, wh.data(), wh.size() * 4, cudaMemcpyHostToDevice, st);
    cudaEvent_t ev[3];
    for (auto& e : ev) cudaEventCreate(&e);
    float tf = 0, tn = 0;
    const int reps = std::max(1, iters);
    for (int it = 0; it < reps && rc == ZL_OK; ++it) {
        cudaMemsetAsync(pb.cand_count, 0, (size_t)n * 4, st);
        cudaMemsetAsync(pb.header, 0, 16, st);
        cudaEventRecord(ev[0], st);
        rc = launch_filter(st, d_raw, n, nc, A, nullptr, d_wh, conf, nullptr, pb);
        cudaEventRecord(ev[1], st);
        if (rc == ZL_OK) rc = launch_nms(st, n, A, iou, pb, true);
        cudaEventRecord(ev[2], st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { set_error(std::string("decode_nms: ") + cudaGetErrorString(cudaGetLastError())); rc = ZL_INFERENCE_ERROR; break; }
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, ev[0], ev[1]); cudaEventElapsedTime(&b, ev[1], ev[2]);
        if (it > 0 || reps == 1) { tf += a; tn += b; }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (rc == ZL_OK) {
        std::vector<uint32_t> hdr(4 + 2 * n);
        cudaMemcpy(hdr.data(), pb.header, hdr.size() * 4, cudaMemcpyDeviceToHost);
        const uint32_t total = hdr[0];
        if (kept) *kept = total;
        if (ms_filter) *ms_filter = tf / std::max(1, reps - (reps > 1 ? 1 : 0));
        if (ms_nms) *ms_nms = tn / std::max(1, reps - (reps > 1 ? 1 : 0));
        if (out) {
            std::vector<DevDet> all(total);
            if (total) cudaMemcpy(all.data(), pb.dets, (size_t)total * sizeof(DevDet), cudaMemcpyDeviceToHost);
            int run = 0;
            for (int i = 0; i < n; ++i) {
                const uint32_t c = hdr[4 + i], o = hdr[4 + n + i];
                if (counts) counts[i] = (int32_t)c;
                if (offsets) offsets[i] = run;
                for (uint32_t j = 0; j < c; ++j) {
                    if (run < cap) std::memcpy(&out[run], &all[o + j], sizeof(zl_det)); else rc = ZL_INSUFFICIENT_RESOURCES;
                    run++;
                }
            }
            if (rc == ZL_INSUFFICIENT_RESOURCES) set_error("dets_out capacity too small");
        }
    }
    cleanup();
    return rc;
}

// ------------------------------------------------------------------ measurement
int32_t Engine::upload_resident(int set, const uint8_t* const* frames, const int32_t* ws, const int32_t* hs, int n)
{
    if (set < 0 || set > 3 || n < 1 || n > cfg.max_batch) ZL_FAIL(ZL_INVALID_ARGUMENT, "bad resident set / count");
    ZL_CUDA(cudaSetDevice(cfg.device));
    const size_t slot = slot_bytes();
    for (auto& Lp : lanes) {                                // every lane gets its own copy: resident steps can run on all lanes
        Lane& L = *Lp;
        std::lock_guard<std::mutex> g(L.mu);
        if (L.staging_slots < (size_t)cfg.max_batch * 5) {      // grow once: [live | set0 | set1 | set2 | set3]
            uint8_t* ns = nullptr;
            ZL_CUDA(cudaMalloc(&ns, (size_t)cfg.max_batch * 5 * slot));
            ZL_CUDA(cudaMemset(ns, 128, (size_t)cfg.max_batch * 5 * slot));
            ZL_CUDA(cudaDeviceSynchronize());
            cudaFree(L.staging);
            L.staging = ns;
            L.staging_slots = (size_t)cfg.max_batch * 5;
            for (auto& kv : L.graphs) cudaGraphExecDestroy(kv.second);   // graphs and op lists captured the old staging pointer
            L.graphs.clear();
            L.ops.clear();
        }
        FrameDesc* hd = L.h_descs + (size_t)cfg.max_batch * (1 + set);
        for (int i = 0; i < n; ++i) {
            const size_t bytes = (size_t)ws[i] * hs[i] * 3;
            if (bytes > slot) ZL_FAIL(ZL_INVALID_INPUT, "frame larger than max_frame");
            const size_t off = ((size_t)(1 + set) * cfg.max_batch + i) * slot;
            ZL_CUDA(cudaMemcpy(L.staging + off, frames[i], bytes, cudaMemcpyHostToDevice));
            hd[i] = FrameDesc{off, ws[i], hs[i]};
        }
        ZL_CUDA(cudaMemcpy(L.d_res_descs + (size_t)set * cfg.max_batch, hd, sizeof(FrameDesc) * n, cudaMemcpyHostToDevice));
        L.resident_n[set] = n;
        bool same = (slot % 8) == 0;
        for (int i = 0; i < n; ++i) same = same && ws[i] == cfg.model_w && hs[i] == cfg.model_h;
        L.resident_same[set] = same;
    }
    return ZL_OK;
}

// `steps` passes over the resident sets, distributed round-robin over all lanes (streams): lane 0's stream forks the
// others with an event and joins them before the closing event, so the CUDA-event time covers all of them.
int32_t Engine::run_resident(int n_sets, int steps, float* total_ms, int64_t* launches, int64_t* total_dets)
{
    if (n_sets < 1 || n_sets > 4 || steps < 1) ZL_FAIL(ZL_INVALID_ARGUMENT, "bad n_sets / steps");
    ZL_CUDA(cudaSetDevice(cfg.device));
    std::vector<std::unique_lock<std::mutex>> locks;
    for (auto& Lp : lanes) locks.emplace_back(Lp->mu);
    Lane& L0 = *lanes[0];
    const int n = L0.resident_n[0];
    for (int s = 0; s < n_sets; ++s) if (L0.resident_n[s] != n || n == 0) ZL_FAIL(ZL_INVALID_ARGUMENT, "resident sets not uploaded / unequal");
    const int B = graph_batch_for(n);
    if (B != n) ZL_FAIL(ZL_INVALID_ARGUMENT, "resident batch must be a power of two or max_batch");
    const int nl = (int)lanes.size();
    bool rsame = true;
    for (int s = 0; s < n_sets; ++s) rsame = rsame && L0.resident_same[s];
    for (auto& Lp : lanes) Lp->same_size = rsame;
    if (cfg.use_graph) for (auto& Lp : lanes) ZL_TRY(ensure_graph(*Lp, B));
    ZL_CUDA(cudaDeviceSynchronize());
    ZL_CUDA(cudaEventRecord(L0.ev0, L0.stream));
    for (int l = 1; l < nl; ++l) ZL_CUDA(cudaStreamWaitEvent(lanes[l]->stream, L0.ev0, 0));
    for (int s = 0; s < steps; ++s) {
        Lane& L = *lanes[s % nl];
        ZL_CUDA(cudaMemcpyAsync(L.d_descs, L.d_res_descs + (size_t)(s % n_sets) * cfg.max_batch, sizeof(FrameDesc) * B, cudaMemcpyDeviceToDevice, L.stream));
        L.descs_on_device.clear();
        ZL_TRY(launch_batch(L, B, false));
    }
    for (int l = 1; l < nl; ++l) {
        ZL_CUDA(cudaEventRecord(lanes[l]->ev1, lanes[l]->stream));
        ZL_CUDA(cudaStreamWaitEvent(L0.stream, lanes[l]->ev1, 0));
    }
    ZL_CUDA(cudaEventRecord(L0.ev1, L0.stream));
    ZL_CUDA(cudaStreamSynchronize(L0.stream));
    float ms = 0;
    ZL_CUDA(cudaEventElapsedTime(&ms, L0.ev0, L0.ev1));
    if (total_ms) *total_ms = ms;
    if (launches) *launches = (int64_t)steps * (int64_t)hot_ops(L0.ops[B]).size();     // the ops of the hot path (raw-mode alternates excluded)
    if (total_dets) *total_dets = ((const uint32_t*)L0.h_result)[0];
    return ZL_OK;
}

int32_t Engine::profile(int set, int iters, zl_op_profile* out, int cap, int32_t* n_out)
{
    if (set < 0 || set > 3 || iters < 1 || !out || !n_out) ZL_FAIL(ZL_INVALID_ARGUMENT, "bad argument");
    ZL_CUDA(cudaSetDevice(cfg.device));
    Lane& L = *lanes[0];
    std::lock_guard<std::mutex> g(L.mu);
    const int n = L.resident_n[set];
    if (n == 0) ZL_FAIL(ZL_INVALID_ARGUMENT, "resident set not uploaded");
    L.same_size = L.resident_same[set];
    const int B = graph_batch_for(n);
    if (!L.ops.count(B)) ZL_TRY(build_ops(L, B));
    ZL_CUDA(cudaMemcpyAsync(L.d_descs, L.d_res_descs + (size_t)set * cfg.max_batch, sizeof(FrameDesc) * B, cudaMemcpyDeviceToDevice, L.stream));
    L.descs_on_device.clear();
    ZL_TRY(run_ops(L, B, false));                         // warm
    ZL_CUDA(cudaStreamSynchronize(L.stream));
    std::vector<Op> ops = hot_ops(L.ops[B]);
    std::vector<cudaEvent_t> ev(ops.size() + 1);
    for (auto& e : ev) cudaEventCreate(&e);
    std::vector<double> acc(ops.size(), 0.0);
    cudaStream_t st = L.stream;
    int32_t rc = ZL_OK;
    for (int it = 0; it < iters && rc == ZL_OK; ++it) {
        cudaMemcpyAsync(L.d_descs, L.d_res_descs + (size_t)set * cfg.max_batch, sizeof(FrameDesc) * B, cudaMemcpyDeviceToDevice, st);
        L.descs_on_device.clear();
        cudaMemsetAsync(L.pb.cand_count, 0, sizeof(uint32_t) * B, st);
        cudaMemsetAsync(L.pb.header, 0, 16, st);
        for (size_t i = 0; i < ops.size() && rc == ZL_OK; ++i) {
            cudaEventRecord(ev[i], st);
            rc = launch_op(L, B, ops[i]);
        }
        cudaEventRecord(ev[ops.size()], st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { set_error(std::string("profile: ") + cudaGetErrorString(cudaGetLastError())); rc = ZL_INFERENCE_ERROR; }
        for (size_t i = 0; i < ops.size() && rc == ZL_OK; ++i) { float ms = 0; cudaEventElapsedTime(&ms, ev[i], ev[i + 1]); acc[i] += ms; }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (rc != ZL_OK) return rc;
    int k = 0;
    for (size_t i = 0; i < ops.size() && k < cap; ++i, ++k) {
        zl_op_profile& r = out[k];
        std::memset(&r, 0, sizeof(r));
        std::strncpy(r.name, ops[i].name.c_str(), sizeof(r.name) - 1);
        r.kind = ops[i].kind; r.launches = 1; r.ms = (float)(acc[i] / iters);
        r.flops = ops[i].flops; r.bytes = ops[i].bytes;
    }
    *n_out = k;
    return ZL_OK;
}

// One un-captured pass with the instrumented persistent conv kernel: kHaloStatSlots cycle counters per op (zero for the
// ops that are not conv_halo launches), op order == zl_engine_profile's.
int32_t Engine::profile_stalls(int set, uint64_t* out, int cap_ops, int32_t* n_out)
{
    if (set < 0 || set > 3 || !out || !n_out) ZL_FAIL(ZL_INVALID_ARGUMENT, "bad argument");
    ZL_CUDA(cudaSetDevice(cfg.device));
    Lane& L = *lanes[0];
    std::lock_guard<std::mutex> g(L.mu);
    const int n = L.resident_n[set];
    if (n == 0) ZL_FAIL(ZL_INVALID_ARGUMENT, "resident set not uploaded");
    L.same_size = L.resident_same[set];
    const int B = graph_batch_for(n);
    if (!L.ops.count(B)) ZL_TRY(build_ops(L, B));
    std::vector<Op> ops = hot_ops(L.ops[B]);
    if ((int)ops.size() > cap_ops) ZL_FAIL(ZL_INSUFFICIENT_RESOURCES, "stall buffer too small");
    unsigned long long* d = nullptr;
    const size_t bytes = ops.size() * kHaloStatSlots * sizeof(unsigned long long);
    ZL_CUDA(cudaMalloc(&d, bytes));
    cudaMemsetAsync(d, 0, bytes, L.stream);
    cudaMemcpyAsync(L.d_descs, L.d_res_descs + (size_t)set * cfg.max_batch, sizeof(FrameDesc) * B, cudaMemcpyDeviceToDevice, L.stream);
    L.descs_on_device.clear();
    cudaMemsetAsync(L.pb.cand_count, 0, sizeof(uint32_t) * B, L.stream);
    cudaMemsetAsync(L.pb.header, 0, 16, L.stream);
    int32_t rc = ZL_OK;
    for (size_t i = 0; i < ops.size() && rc == ZL_OK; ++i)
        rc = ops[i].kind == Op::CONV_HALO ? conv_halo_launch(L.stream, ops[i].halo, num_sms, d + i * kHaloStatSlots)
             : ops[i].kind == Op::HEAD_FUSED ? head_fused_launch(L.stream, ops[i].hf, L.d_descs, cfg.conf_threshold, d_class_weights, L.pb, d + i * kHaloStatSlots)
                                             : launch_op(L, B, ops[i]);
    cudaError_t ce = cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, L.stream);
    cudaError_t cs = cudaStreamSynchronize(L.stream);
    cudaFree(d);
    if (rc != ZL_OK) return rc;
    if (ce != cudaSuccess || cs != cudaSuccess) ZL_FAIL(ZL_INFERENCE_ERROR, std::string("profile_stalls: ") + cudaGetErrorString(cs != cudaSuccess ? cs : ce));
    *n_out = (int32_t)ops.size();
    return ZL_OK;
}

int32_t Engine::bench_preprocess(int w, int h, int n, int iters, float* ms, double* bytes)
{
    if (w <= 0 || h <= 0 || n < 1 || iters < 1) ZL_FAIL(ZL_INVALID_ARGUMENT, "bad argument");
    ZL_CUDA(cudaSetDevice(cfg.device));
    Lane& L = *lanes[0];
    std::lock_guard<std::mutex> g(L.mu);
    const bool bf16 = cfg.precision != ZL_PRECISION_FP32;   // "bf16" == any 16-bit tensor-core mode
    const bool f16 = cfg.precision == ZL_PRECISION_FP16; (void)f16;
    const size_t fb = (size_t)w * h * 3, ob = (size_t)cfg.model_w * cfg.model_h * 4 * (bf16 ? 2 : 4);
    uint8_t* d_in = nullptr; void* d_out = nullptr; FrameDesc* d_desc = nullptr;
    ZL_CUDA(cudaMalloc(&d_in, fb * n));
    ZL_CUDA(cudaMalloc(&d_out, ob * n));
    ZL_CUDA(cudaMalloc(&d_desc, sizeof(FrameDesc) * n));
    ZL_CUDA(cudaMemset(d_in, 77, fb * n));
    std::vector<FrameDesc> hd(n);
    for (int i = 0; i < n; ++i) hd[i] = FrameDesc{(uint64_t)i * fb, w, h};
    ZL_CUDA(cudaMemcpy(d_desc, hd.data(), sizeof(FrameDesc) * n, cudaMemcpyHostToDevice));
    int32_t rc = ZL_OK;
    for (int i = 0; i < 3 && rc == ZL_OK; ++i) rc = launch_preprocess(L.stream, d_in, d_desc, n, cfg.model_w, cfg.model_h, bf16 ? (f16 ? PRE_NHWC4_F16 : PRE_NHWC4_BF16) : PRE_NHWC4_F32, d_out);
    cudaEventRecord(L.ev0, L.stream);
    for (int i = 0; i < iters && rc == ZL_OK; ++i) rc = launch_preprocess(L.stream, d_in, d_desc, n, cfg.model_w, cfg.model_h, bf16 ? (f16 ? PRE_NHWC4_F16 : PRE_NHWC4_BF16) : PRE_NHWC4_F32, d_out);
    cudaEventRecord(L.ev1, L.stream);
    cudaError_t cs = cudaStreamSynchronize(L.stream);
    float t = 0;
    cudaEventElapsedTime(&t, L.ev0, L.ev1);
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_desc);
    if (rc != ZL_OK) return rc;
    if (cs != cudaSuccess) ZL_FAIL(ZL_INFERENCE_ERROR, std::string("bench_preprocess: ") + cudaGetErrorString(cs));
    if (ms) *ms = t / iters;
    // bytes actually sampled (min(src, dst) pixels x 3) + output written (3 channels x element size)
    const double sampled = (double)std::min((size_t)w * h, (size_t)cfg.model_w * cfg.model_h) * 3;
    if (bytes) *bytes = n * (sampled + (double)cfg.model_w * cfg.model_h * 3 * (bf16 ? 2 : 4));
    return ZL_OK;
}

int32_t Engine::bench_latency(const uint8_t* bgr, int w, int h, int warm, int iters, float* ms_out)
{
    if (!weights_loaded) ZL_FAIL(ZL_NOT_INITIALIZED, "weights not loaded");
    if (!bgr || !ms_out || iters < 1) ZL_FAIL(ZL_INVALID_ARGUMENT, "bad argument");
    Lane& L = *lanes[0];
    std::lock_guard<std::mutex> g(L.mu);
    const bool pinned = is_pinned(bgr, (size_t)w * h * 3);
    std::vector<zl_det> dets;
    int32_t cnt = 0;
    const uint8_t* fr[1] = {bgr};
    for (int i = 0; i < warm + iters; ++i) {
        const auto t0 = std::chrono::steady_clock::now();
        ZL_TRY(run_lane_batch(L, fr, &w, &h, 1, pinned, &dets, &cnt));
        const auto t1 = std::chrono::steady_clock::now();
        if (i >= warm) ms_out[i - warm] = std::chrono::duration<float, std::milli>(t1 - t0).count();
    }
    return ZL_OK;
}

// End-to-end throughput loop in C (no interpreter in the timed path): `threads` host threads, each owning one batch of
// `n` equal-size frames that sit back to back in (pinned) host memory at batches[t], call infer_batch `steps_per_thread`
// times: H2D of the frames, the whole device path, D2H of the detections, every step.  Returns the wall time of the
// slowest thread and the detections of the last step.
int32_t Engine::bench_e2e(const uint8_t* const* batches, int threads, int n, int w, int h, int steps_total, double* seconds, int64_t* dets_last)
{
    if (!weights_loaded) ZL_FAIL(ZL_NOT_INITIALIZED, "weights not loaded");
    if (!batches || threads < 1 || threads > 16 || n < 1 || steps_total < 1 || !seconds) ZL_FAIL(ZL_INVALID_ARGUMENT, "bad argument");
    const size_t fb = (size_t)w * h * 3;
    std::vector<int32_t> rcs(threads, ZL_OK);
    std::vector<std::string> errs(threads);
    std::vector<int64_t> nd(threads, 0);
    std::vector<std::thread> th;
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < threads; ++t) {
        const int my_steps = steps_total / threads + (t < steps_total % threads ? 1 : 0);
        th.emplace_back([&, t, my_steps] {
            cudaSetDevice(cfg.device);
            std::vector<const uint8_t*> fr(n);
            std::vector<int32_t> ws(n, w), hs(n, h), cnt(n), off(n);
            for (int i = 0; i < n; ++i) fr[i] = batches[t] + (size_t)i * fb;
            std::vector<zl_det> out((size_t)n * 512);
            for (int s = 0; s < my_steps; ++s) {
                int32_t rc = infer_batch(fr.data(), ws.data(), hs.data(), n, out.data(), (int)out.size(), cnt.data(), off.data(), nullptr);
                if (rc == ZL_INSUFFICIENT_RESOURCES) rc = ZL_OK;          // more detections than the scratch holds: counts are still valid
                if (rc != ZL_OK) { rcs[t] = rc; errs[t] = get_error(); return; }
                int64_t k = 0;
                for (int i = 0; i < n; ++i) k += cnt[i];
                nd[t] = k;
            }
        });
    }
    for (auto& x : th) x.join();
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int t = 0; t < threads; ++t) if (rcs[t] != ZL_OK) { set_error(errs[t]); return rcs[t]; }
    if (dets_last) *dets_last = nd[0];
    return ZL_OK;
}

// Pinned host -> device copy bandwidth on this engine's device (the e2e leg's ceiling): `iters` copies of `bytes`.
int32_t Engine::bench_h2d(size_t bytes, int iters, double* gbs)
{
    if (bytes == 0 || iters < 1 || !gbs) ZL_FAIL(ZL_INVALID_ARGUMENT, "bad argument");
    ZL_CUDA(cudaSetDevice(cfg.device));
    void* h = nullptr; void* d = nullptr;
    static const bool wc = [] { const char* e = getenv("ZL_PINNED_WC"); return e && e[0] == '1'; }();
    ZL_CUDA(cudaHostAlloc(&h, bytes, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    if (cudaMalloc(&d, bytes) != cudaSuccess) { cudaFreeHost(h); ZL_FAIL(ZL_INSUFFICIENT_RESOURCES, "bench_h2d: out of device memory"); }
    std::memset(h, 1, bytes);
    Lane& L = *lanes[0];
    std::lock_guard<std::mutex> g(L.mu);
    cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, L.stream);
    cudaEventRecord(L.ev0, L.stream);
    for (int i = 0; i < iters; ++i) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, L.stream);
    cudaEventRecord(L.ev1, L.stream);
    const cudaError_t cs = cudaStreamSynchronize(L.stream);
    float ms = 0;
    cudaEventElapsedTime(&ms, L.ev0, L.ev1);
    cudaFree(d); cudaFreeHost(h);
    if (cs != cudaSuccess) ZL_FAIL(ZL_INFERENCE_ERROR, std::string("bench_h2d: ") + cudaGetErrorString(cs));
    *gbs = (double)bytes * iters / (ms * 1e-3) / 1e9;
    return ZL_OK;
}

int32_t Engine::warmup(int iters)
{
    if (!weights_loaded) ZL_FAIL(ZL_NOT_INITIALIZED, "weights not loaded");
    // warmupModel (onnx_engine.cpp:919-954): all-128 frame of model size, 3 runs by default
    const size_t fb = (size_t)cfg.model_w * cfg.model_h * 3;
    if (fb > slot_bytes()) ZL_FAIL(ZL_INVALID_ARGUMENT, "max_frame smaller than the model input");
    std::vector<uint8_t> grey(fb, 128);
    for (auto& Lp : lanes) {
        Lane& L = *Lp;
        std::lock_guard<std::mutex> g(L.mu);
        std::vector<zl_det> dets;
        for (int B : {1, cfg.max_batch}) {
            std::vector<const uint8_t*> fr(B, grey.data());
            std::vector<int32_t> ws(B, cfg.model_w), hs(B, cfg.model_h), cnt(B);
            for (int i = 0; i < std::max(1, iters); ++i)
                ZL_TRY(run_lane_batch(L, fr.data(), ws.data(), hs.data(), B, false, &dets, cnt.data()));
            if (cfg.max_batch == 1) break;
        }
    }
    return start_workers();
}

// ------------------------------------------------------------------ async path
int32_t Engine::start_workers()
{
    if (running.load()) return ZL_OK;
    if (!h_slots) {
        ZL_CUDA(cudaSetDevice(cfg.device));
        ZL_CUDA(cudaHostAlloc(&h_slots, (size_t)cfg.queue_depth * slot_bytes(), cudaHostAllocDefault));
        free_slots.clear();
        for (int i = cfg.queue_depth - 1; i >= 0; --i) free_slots.push_back(i);
    }
    stopping = false;
    running.store(true);
    for (int i = 0; i < (int)lanes.size(); ++i) workers.emplace_back(&Engine::worker_main, this, i);
    return ZL_OK;
}

void Engine::stop_workers()
{
    if (!running.load()) return;
    { std::lock_guard<std::mutex> g(qmu); stopping = true; }
    qcv.notify_all();
    order_cv.notify_all();
    for (auto& t : workers) if (t.joinable()) t.join();
    workers.clear();
    running.store(false);
}

int32_t Engine::submit(uint32_t client_id, uint32_t frame_id, uint64_t ts, int w, int h, const uint8_t* bgr, size_t len)
{
    if (!running.load()) ZL_FAIL(ZL_NOT_INITIALIZED, "Inference engine not running");          // onnx_engine.cpp:224-226
    if (!bgr || w <= 0 || h <= 0 || len != (size_t)w * h * 3)                                   // onnx_engine.cpp:659-665
        ZL_FAIL(ZL_INVALID_INPUT, "Invalid image data size: expected " + std::to_string((size_t)std::max(w, 0) * std::max(h, 0) * 3) + ", got " + std::to_string(len));
    if (len > slot_bytes()) ZL_FAIL(ZL_INVALID_INPUT, "frame larger than max_frame_w x max_frame_h");
    int slot;
    {
        std::lock_guard<std::mutex> g(qmu);
        if (free_slots.empty()) {
            std::lock_guard<std::mutex> g2(smu);
            st_dropped++;
            ZL_FAIL(ZL_INFERENCE_ERROR, "inference queue full, frame dropped");               // network_server.cpp:213-215
        }
        slot = free_slots.back();
        free_slots.pop_back();
        in_flight++;
    }
    std::memcpy(h_slots + (size_t)slot * slot_bytes(), bgr, len);       // the reference copies the frame too (onnx_engine.cpp:235)
    Request r{client_id, frame_id, ts, w, h, slot, std::chrono::steady_clock::now()};
    {
        std::lock_guard<std::mutex> g(qmu);
        queue.push_back(r);
        std::lock_guard<std::mutex> g2(smu);
        st_hwm = std::max<uint64_t>(st_hwm, queue.size());
    }
    qcv.notify_one();
    return ZL_OK;
}

void Engine::worker_main(int lane_id)
{
    cudaSetDevice(cfg.device);
    // ServerConfig::use_cpu_affinity / cpu_core_id / use_high_priority (the reference pins and raises its inference thread,
    // onnx_engine.cpp:318-330): worker i of this engine goes to core cpu_core_id + i; failures are not errors
    if (cfg.cpu_core_id >= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET((cfg.cpu_core_id + lane_id) % std::max(1, (int)std::thread::hardware_concurrency()), &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    if (cfg.high_priority) {
        sched_param sp{};
        sp.sched_priority = 0;
        if (nice(-5) == -1) { /* not permitted: keep the default */ }
        (void)sp;
    }
    Lane& L = *lanes[lane_id];
    std::vector<Request> batch;
    std::vector<const uint8_t*> fr;
    std::vector<int32_t> ws, hs, cnt;
    std::vector<zl_det> dets;
    while (true) {
        uint64_t seq;
        batch.clear();
        {
            std::unique_lock<std::mutex> lk(qmu);
            qcv.wait(lk, [&] { return stopping || !queue.empty(); });
            if (queue.empty() && stopping) return;
            if (cfg.batch_window_us > 0 && (int)queue.size() < cfg.max_batch)
                qcv.wait_for(lk, std::chrono::microseconds(cfg.batch_window_us), [&] { return stopping || (int)queue.size() >= cfg.max_batch; });
            while (!queue.empty() && (int)batch.size() < cfg.max_batch) { batch.push_back(queue.front()); queue.pop_front(); }
            if (batch.empty()) continue;        // another lane took the frames while this one sat in the batch window
            seq = next_seq++;
        }
        const int n = (int)batch.size();
        fr.resize(n); ws.resize(n); hs.resize(n); cnt.assign(n, 0);
        for (int i = 0; i < n; ++i) { fr[i] = h_slots + (size_t)batch[i].slot * slot_bytes(); ws[i] = batch[i].w; hs[i] = batch[i].h; }
        int32_t rc;
        std::string err;
        std::vector<uint8_t> wire;
        std::vector<uint32_t> woff;
        {
            std::lock_guard<std::mutex> g(L.mu);
            if (wire_cb) {
                std::vector<uint32_t> ids(n);
                std::vector<uint64_t> tss(n);
                for (int i = 0; i < n; ++i) { ids[i] = batch[i].frame_id; tss[i] = batch[i].timestamp; }
                const uint64_t now_ms = (uint64_t)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
                WireReq wr{ids.data(), tss.data(), now_ms};        // Detection::timestamp = wall-clock ms (onnx_engine.cpp:813-815)
                rc = run_lane_batch(L, fr.data(), ws.data(), hs.data(), n, true, &dets, cnt.data(), false, &wr);
                if (rc == ZL_OK) { woff.assign(L.h_wire_off, L.h_wire_off + n + 1); wire.assign(L.h_wire, L.h_wire + woff[n]); }
            } else {
                rc = run_lane_batch(L, fr.data(), ws.data(), hs.data(), n, true, &dets, cnt.data());
            }
            if (rc != ZL_OK) err = get_error();
        }
        // deliver in pop order across lanes
        {
            std::unique_lock<std::mutex> lk(qmu);
            order_cv.wait(lk, [&] { return deliver_seq == seq; });
        }
        const auto now = std::chrono::steady_clock::now();
        size_t k = 0;
        for (int i = 0; i < n; ++i) {
            const int c = rc == ZL_OK ? cnt[i] : 0;
            if (wire_cb) wire_cb(wire_user, batch[i].client_id, batch[i].frame_id, batch[i].timestamp, rc,
                                 rc == ZL_OK ? wire.data() + woff[i] : nullptr, rc == ZL_OK ? (size_t)(woff[i + 1] - woff[i]) : 0);
            else if (cb) cb(cb_user, batch[i].client_id, batch[i].frame_id, batch[i].timestamp, rc, c ? dets.data() + k : nullptr, c);
            k += c;
        }
        {
            std::lock_guard<std::mutex> g(smu);
            for (int i = 0; i < n; ++i) {
                lat_ms.push_back(std::chrono::duration<double, std::milli>(now - batch[i].t_submit).count());
                if (lat_ms.size() > 1000) lat_ms.pop_front();
            }
            if (rc == ZL_OK) st_count += n; else st_errors += n;
        }
        {
            std::lock_guard<std::mutex> g(qmu);
            for (int i = 0; i < n; ++i) free_slots.push_back(batch[i].slot);
            in_flight -= n;
            deliver_seq++;
        }
        order_cv.notify_all();
        done_cv.notify_all();
    }
}

int32_t Engine::drain()
{
    std::unique_lock<std::mutex> lk(qmu);
    done_cv.wait(lk, [&] { return in_flight == 0 || !running.load(); });
    return ZL_OK;
}

size_t Engine::queue_size() const
{
    std::lock_guard<std::mutex> g(qmu);
    return queue.size();
}

void Engine::get_stats(zl_stats* o) const
{
    std::memset(o, 0, sizeof(*o));
    { std::lock_guard<std::mutex> g(qmu); o->queue_size = queue.size(); }
    std::lock_guard<std::mutex> g(smu);
    o->inference_count = st_count; o->inference_errors = st_errors; o->dropped_frames = st_dropped;
    o->queue_high_water_mark = st_hwm; o->batches = st_batches;
    if (!lat_ms.empty()) {
        std::vector<double> v(lat_ms.begin(), lat_ms.end());
        double s = 0; for (double x : v) s += x;
        o->avg_inference_time_ms = s / v.size();
        std::sort(v.begin(), v.end());
        o->p99_inference_time_ms = v[std::min(v.size() - 1, (size_t)(v.size() * 0.99))];
    }
    o->avg_device_time_ms = dev_ms_n ? dev_ms_sum / dev_ms_n : 0.0;
    o->graph_captured = graph_captured; o->device = cfg.device; o->precision = cfg.precision; o->running = running.load() ? 1 : 0;
}

}  // namespace zl
