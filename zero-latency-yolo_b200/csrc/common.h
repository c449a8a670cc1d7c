// common.h — shared types of the B200 detector (host + device).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/zl_b200.h"

namespace zl {

// ---- error plumbing -------------------------------------------------------
void set_error(const std::string& msg);          // thread-local, read by zl_last_error()
const char* get_error();

struct Status {
    int32_t code = ZL_OK;
    bool ok() const { return code == ZL_OK; }
};

#define ZL_FAIL(code_, msg_)                                   \
    do {                                                       \
        ::zl::set_error(std::string(msg_));                    \
        return (code_);                                        \
    } while (0)

#define ZL_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t err__ = (expr);                                                        \
        if (err__ != cudaSuccess) {                                                        \
            ::zl::set_error(std::string(#expr) + ": " + cudaGetErrorString(err__) +        \
                            " (" __FILE__ ":" + std::to_string(__LINE__) + ")");           \
            return (err__ == cudaErrorMemoryAllocation) ? ZL_INSUFFICIENT_RESOURCES        \
                                                        : ZL_INFERENCE_ERROR;              \
        }                                                                                  \
    } while (0)

#define ZL_TRY(expr)                      \
    do {                                  \
        int32_t rc__ = (expr);            \
        if (rc__ != ZL_OK) return rc__;   \
    } while (0)

// ---- tensor view: NHWC slice of a (possibly wider) buffer -----------------
enum DType : int32_t { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };   // both 16-bit formats share the h16 buffers

struct View {
    void* ptr = nullptr;   // already offset to the first channel of the slice
    int32_t n = 0, h = 0, w = 0, c = 0;
    int32_t pitch = 0;     // elements between consecutive pixels of the underlying buffer
    int32_t dtype = DT_BF16;
    size_t esize() const { return dtype == DT_F32 ? 4 : 2; }
    bool is16() const { return dtype != DT_F32; }
    size_t pixels() const { return (size_t)n * h * w; }
    View slice(int32_t c0, int32_t cn) const {
        View v = *this;
        v.ptr = (char*)ptr + (size_t)c0 * esize();
        v.c = cn;
        return v;
    }
    View with_n(int32_t nn) const { View v = *this; v.n = nn; return v; }
};

// One candidate / detection as stored on the device (== zl_det, 24 B).
struct DevDet {
    float x, y, w, h, conf;
    int32_t cls;
};

// Frame descriptor consumed by the preprocess kernel.
struct FrameDesc {
    uint64_t offset;   // byte offset of the frame inside the staging buffer
    int32_t w, h;
};

// Letterbox preprocessing (ZL_PRE_LETTERBOX; north_star (1) — NOT a parity mode: the reference stretches,
// onnx_engine.cpp:673-693).  The frame is resized with ONE gain (aspect preserved, nearest sampling like the parity mode),
// centred, and the border is filled with 114/255 (the ultralytics convention).  The same mapping, inverted, takes the
// kept boxes back to the request frame.
struct LetterboxMap {
    float gain, inv_gain;
    int32_t nw, nh, pad_x, pad_y;
};
__host__ __device__ inline LetterboxMap letterbox_map(int w, int h, int mw, int mh)
{
    LetterboxMap m;
    const float gw = (float)mw / (float)w, gh = (float)mh / (float)h;
    m.gain = gw < gh ? gw : gh;
    m.nw = (int)((float)w * m.gain + 0.5f); m.nh = (int)((float)h * m.gain + 0.5f);
    if (m.nw > mw) m.nw = mw;
    if (m.nh > mh) m.nh = mh;
    if (m.nw < 1) m.nw = 1;
    if (m.nh < 1) m.nh = 1;
    m.pad_x = (mw - m.nw) / 2; m.pad_y = (mh - m.nh) / 2;
    m.inv_gain = 1.0f / m.gain;
    return m;
}

constexpr int kKeyAnchorBits = 20;   // sort key: class[12] | ~conf_bits[32] | anchor[20]
constexpr int kMaxAnchors = 1 << kKeyAnchorBits;
constexpr int kMaxClasses = 1 << 12;

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

}  // namespace zl
