// test_hooks.cpp — unit-test and measurement hooks (zl_test_conv here, zl_probe_umma / zl_probe_tma in umma_probe.cu).
// NOT part of the product: built into lib/libzl_b200_test.so only (csrc/build.sh); libzl_b200.so does not export them.
// Declared in include/zl_b200_test.h.
#include <cstring>

#include <cuda_fp16.h>
#include <new>
#include <vector>

#include "../../include/zl_b200_test.h"
#include "engine.h"

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

extern "C" {

static inline uint16_t f2bf_host(float f) {
    uint32_t u; std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float bf2f_host(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; std::memcpy(&f, &u, 4); return f; }
static inline uint16_t f2h_host(float f) { const __half h = __float2half_rn(f); uint16_t u; std::memcpy(&u, &h, 2); return u; }
static inline float h2f_host(uint16_t v) { const __half_raw r{v}; return __half2float(__half(r)); }

int32_t zl_test_conv(int32_t device, int32_t impl, const float* x, int32_t n, int32_t h, int32_t w, int32_t cin,
                     const float* wgt, const float* bias, int32_t cout, int32_t k, int32_t stride, int32_t act_flags,
                     const float* res, float* y)
{
    ZL_GUARD_BEGIN
    using namespace zl;
    if (!x || !wgt || !bias || !y) { set_error("null argument"); return ZL_INVALID_ARGUMENT; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) { cudaGetLastError(); set_error("no such CUDA device"); return ZL_INSUFFICIENT_RESOURCES; }
    ZL_CUDA(cudaSetDevice(device));
    const int act = act_flags & 1, out_f32 = (act_flags >> 1) & 1, f16 = (act_flags >> 2) & 1, hint = (act_flags >> 8) & 0x1ff;
    auto cvt = [&](float v) { return f16 ? f2h_host(v) : f2bf_host(v); };
    const int dt16 = f16 ? DT_F16 : DT_BF16;
    const int pad = k / 2, ho = (h + 2 * pad - k) / stride + 1, wo = (w + 2 * pad - k) / stride + 1;
    ConvWeights cw;
    cw.name = "test"; cw.cin = cin; cw.cout = cout; cw.k = k; cw.stride = stride; cw.act = act;
    cw.cout_pad = round_up(cout, 16); cw.ktot = k * k * cin;
    std::vector<float> b(cw.cout_pad, 0.f);
    std::copy(bias, bias + cout, b.begin());
    const size_t nx = (size_t)n * h * w * cin, ny = (size_t)n * ho * wo * cout;
    std::vector<void*> frees;
    auto dalloc = [&](size_t bytes) -> void* { void* p = nullptr; if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr; frees.push_back(p); return p; };
    // cw's buffers are tracked in `frees` like the rest: detach them so ~ConvWeights does not free them again
    auto cleanup = [&] { for (void* p : frees) cudaFree(p); frees.clear(); cw.bias = nullptr; cw.w_simt = nullptr; cw.w_tc = nullptr; };
    cw.bias = (float*)dalloc(b.size() * 4);
    if (!cw.bias) { cleanup(); set_error("oom"); return ZL_INSUFFICIENT_RESOURCES; }
    cudaMemcpy(cw.bias, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    int32_t rc = ZL_OK;
    cudaStream_t st = nullptr;
    cudaStreamCreate(&st);
    if (impl == 0) {
        std::vector<float> ws((size_t)cw.ktot * cw.cout_pad, 0.f);
        for (int o = 0; o < cout; ++o) for (int t = 0; t < k * k; ++t) for (int c = 0; c < cin; ++c)
            ws[((size_t)t * cin + c) * cw.cout_pad + o] = wgt[((size_t)o * k * k + t) * cin + c];
        cw.w_simt = (float*)dalloc(ws.size() * 4);
        float* dx = (float*)dalloc(nx * 4); float* dy = (float*)dalloc(ny * 4); float* dr = res ? (float*)dalloc(ny * 4) : nullptr;
        if (!cw.w_simt || !dx || !dy || (res && !dr)) { cleanup(); set_error("oom"); return ZL_INSUFFICIENT_RESOURCES; }
        cudaMemcpy(cw.w_simt, ws.data(), ws.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dx, x, nx * 4, cudaMemcpyHostToDevice);
        if (res) cudaMemcpy(dr, res, ny * 4, cudaMemcpyHostToDevice);
        View vx{dx, n, h, w, cin, cin, DT_F32}, vy{dy, n, ho, wo, cout, cout, DT_F32}, vr{dr, n, ho, wo, cout, cout, DT_F32};
        rc = launch_conv_simt(st, cw, vx, vy, res ? &vr : nullptr);
        if (rc == ZL_OK && cudaStreamSynchronize(st) != cudaSuccess) { set_error(std::string("conv_simt: ") + cudaGetErrorString(cudaGetLastError())); rc = ZL_INFERENCE_ERROR; }
        if (rc == ZL_OK) cudaMemcpy(y, dy, ny * 4, cudaMemcpyDeviceToHost);
    } else {
        std::vector<uint16_t> wt((size_t)cw.cout_pad * cw.ktot, 0), hx(nx), hr(res ? ny : 0);
        for (int o = 0; o < cout; ++o) for (int t = 0; t < cw.ktot; ++t) wt[(size_t)o * cw.ktot + t] = cvt(wgt[(size_t)o * cw.ktot + t]);
        for (size_t i = 0; i < nx; ++i) hx[i] = cvt(x[i]);
        for (size_t i = 0; i < hr.size(); ++i) hr[i] = cvt(res[i]);
        cw.w_tc = (__nv_bfloat16*)dalloc(wt.size() * 2);
        void* dx = dalloc(nx * 2); void* dy = dalloc(ny * (out_f32 ? 4 : 2)); void* dr = res ? dalloc(ny * 2) : nullptr;
        if (!cw.w_tc || !dx || !dy || (res && !dr)) { cleanup(); set_error("oom"); return ZL_INSUFFICIENT_RESOURCES; }
        cudaMemcpy(cw.w_tc, wt.data(), wt.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dx, hx.data(), nx * 2, cudaMemcpyHostToDevice);
        if (res) cudaMemcpy(dr, hr.data(), ny * 2, cudaMemcpyHostToDevice);
        cudaMemset(dy, 0xff, ny * (out_f32 ? 4 : 2));
        View vx{dx, n, h, w, cin, cin, dt16}, vy{dy, n, ho, wo, cout, cout, out_f32 ? DT_F32 : dt16}, vr{dr, n, ho, wo, cout, cout, dt16};
        if (impl == 3) {
            ConvHaloOp hop;
            cudaDeviceProp prop;
            cudaGetDeviceProperties(&prop, device);
            rc = conv_halo_prepare(cw, vx, vy, res ? &vr : nullptr, hint > 0 ? hint : prop.multiProcessorCount, &hop);
            if (rc == ZL_OK) rc = conv_halo_launch(st, hop, hint > 0 ? hint : prop.multiProcessorCount);   // hint = CTA count (forces several tiles per CTA)
        } else {
            ConvTcOp op;
            rc = conv_tc_prepare(cw, vx, vy, res ? &vr : nullptr, impl == 2, hint, &op);
            if (rc == ZL_OK) rc = conv_tc_launch(st, op);
        }
        if (rc == ZL_OK && cudaStreamSynchronize(st) != cudaSuccess) { set_error(std::string("conv_tc: ") + cudaGetErrorString(cudaGetLastError())); rc = ZL_INFERENCE_ERROR; }
        if (rc == ZL_OK) {
            if (out_f32) cudaMemcpy(y, dy, ny * 4, cudaMemcpyDeviceToHost);
            else {
                std::vector<uint16_t> hy(ny);
                cudaMemcpy(hy.data(), dy, ny * 2, cudaMemcpyDeviceToHost);
                for (size_t i = 0; i < ny; ++i) y[i] = f16 ? h2f_host(hy[i]) : bf2f_host(hy[i]);
            }
        }
    }
    cudaStreamDestroy(st);
    cleanup();
    return rc;
    ZL_GUARD_END
}

// SPPF pools (pool_upsample.cu) on a 16-bit or fp32 NHWC map: x [n,h,w,c] -> p1 | p2 | p3 written into ONE concat buffer
// [n,h,w,4c] next to a copy of x (the layout the engine uses), returned as fp32.  dtype: 0 fp32, 1 bf16, 2 fp16.
int32_t zl_test_sppf_pool(int32_t device, int32_t dtype, const float* x, int32_t n, int32_t h, int32_t w, int32_t c, float* cat)
{
    ZL_GUARD_BEGIN
    using namespace zl;
    if (!x || !cat || dtype < 0 || dtype > 2) { set_error("bad argument"); return ZL_INVALID_ARGUMENT; }
    ZL_CUDA(cudaSetDevice(device));
    const size_t npix = (size_t)n * h * w, nel = npix * 4 * c;
    const int es = dtype == 0 ? 4 : 2;
    std::vector<uint8_t> host(nel * es, 0);
    for (size_t p = 0; p < npix; ++p)
        for (int k = 0; k < c; ++k) {
            const float v = x[p * c + k];
            const size_t o = p * 4 * c + k;
            if (dtype == 0) std::memcpy(&host[o * 4], &v, 4);
            else { const uint16_t q = dtype == 2 ? f2h_host(v) : f2bf_host(v); std::memcpy(&host[o * 2], &q, 2); }
        }
    void* d = nullptr;
    ZL_CUDA(cudaMalloc(&d, nel * es));
    cudaMemcpy(d, host.data(), nel * es, cudaMemcpyHostToDevice);
    const int dt = dtype == 0 ? DT_F32 : (dtype == 2 ? DT_F16 : DT_BF16);
    View cv{d, n, h, w, 4 * c, 4 * c, dt};
    int32_t rc = launch_sppf_pool(nullptr, cv.slice(0, c), cv.slice(c, c), cv.slice(2 * c, c), cv.slice(3 * c, c));
    if (rc == ZL_OK && cudaDeviceSynchronize() != cudaSuccess) { set_error(std::string("sppf_pool: ") + cudaGetErrorString(cudaGetLastError())); rc = ZL_INFERENCE_ERROR; }
    if (rc == ZL_OK) {
        cudaMemcpy(host.data(), d, nel * es, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < nel; ++i) {
            if (dtype == 0) std::memcpy(&cat[i], &host[i * 4], 4);
            else { uint16_t q; std::memcpy(&q, &host[i * 2], 2); cat[i] = dtype == 2 ? h2f_host(q) : bf2f_host(q); }
        }
    }
    cudaFree(d);
    return rc;
    ZL_GUARD_END
}

}  // extern "C"
