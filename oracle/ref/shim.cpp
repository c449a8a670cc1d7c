// shim.cpp — TEST INFRASTRUCTURE ONLY.  The smallest environment in which the reference's OWN text of the hot path
// compiles unmodified: preProcess / postProcess / applyNMS / calculateIoU are cut out of
// /root/reference/src/inference/onnx_engine.cpp by line range (oracle/ref/extract.py, checksummed) into
// oracle/_ref/ref_engine.inc and #included below; BoundingBox / Detection come the same way from
// src/common/types.h:16-26.
// Everything in THIS file is scaffolding written for the oracle: the class declaration restates only the member
// signatures of src/inference/onnx_engine.h:162-190 that the cut text defines, `config_` carries the four fields the
// text reads, ReusableBuffer is the three std::vector forwards the text calls (src/common/memory_pool.h:216-247, a
// header that does not compile on its own), and Ort::Value is a view over a float array with a shape
// (onnx_engine.cpp:767-774 calls GetTensorTypeAndShapeInfo().GetShape() and GetTensorData<float>()).  Result<T> is a
// minimal stand-in with the factory / accessor names of src/common/result.h:76-160 and ErrorCode carries that
// header's numeric values (result.h:14-48): the header itself cannot be used, `Result<T>::ok(lvalue)` — exactly what
// onnx_engine.cpp:699 and :825 call — deduces U = T& and fails to instantiate (one of the reference's hard compile
// errors; types.h:84 and result.h:14 also both define ErrorCode).
//
// The whole reference cannot be built here (ONNX Runtime and OpenCV are absent; types.h and result.h both define
// ErrorCode) — this is the part of the path that is self-contained C++.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>
#include <vector>


namespace zero_latency {
enum class ErrorCode { OK = 0, INFERENCE_ERROR = 200, INVALID_INPUT = 203 };      // values of src/common/result.h:14-48
struct Error { ErrorCode code = ErrorCode::OK; std::string message; };
template <typename T> class Result {
public:
    static Result ok(const T& v) { Result r; r.value_ = v; return r; }
    static Result error(ErrorCode c, const std::string& m) { Result r; r.err_ = Error{c, m}; r.bad_ = true; return r; }
    bool hasError() const { return bad_; }
    bool isOk() const { return !bad_; }
    const T& value() const { return value_; }
    const Error& error() const { return err_; }
private:
    T value_{};
    Error err_;
    bool bad_ = false;
};
}  // namespace zero_latency

namespace Ort {
struct ShapeInfo {
    std::vector<int64_t> shape;
    std::vector<int64_t> GetShape() const { return shape; }
};
struct Value {
    const float* data = nullptr;
    std::vector<int64_t> shape;
    ShapeInfo GetTensorTypeAndShapeInfo() const { return ShapeInfo{shape}; }
    template <typename T> const T* GetTensorData() const { return reinterpret_cast<const T*>(data); }
};
}  // namespace Ort

namespace zero_latency {

#include "ref_types.inc"            // BoundingBox, Detection

template <typename T> class ReusableBuffer {
public:
    void reset() { buffer_.clear(); }
    void resize(size_t n) { buffer_.resize(n); }
    std::vector<T>& getBuffer() { return buffer_; }
private:
    std::vector<T> buffer_;
};

struct ShimConfig {
    struct { int model_width = 0, model_height = 0; } detection;
    float confidence_threshold = 0.5f;
    float nms_threshold = 0.45f;
};

class OnnxInferenceEngine {
public:
    ShimConfig config_;
    std::atomic<uint64_t> inference_errors_{0};
    Result<std::vector<float>> preProcess(const std::vector<uint8_t>& image_data, int width, int height, ReusableBuffer<float>& buffer);
    Result<std::vector<Detection>> postProcess(Ort::Value& output_tensor, int img_width, int img_height);
    std::vector<Detection> applyNMS(std::vector<Detection>& detections, float iou_threshold);
    float calculateIoU(const BoundingBox& box1, const BoundingBox& box2);
};

#include "ref_engine.inc"           // the four member functions, verbatim

}  // namespace zero_latency

using namespace zero_latency;

struct zlr_det { float x, y, w, h, confidence; int32_t class_id; };   // == zl_det / oracle Det: first 24 bytes of Detection
static_assert(sizeof(Detection) == 40, "Detection layout (SURVEY.md 8a T2)");

static void put(const Detection& d, zlr_det* o) { o->x = d.box.x; o->y = d.box.y; o->w = d.box.width; o->h = d.box.height; o->confidence = d.confidence; o->class_id = d.class_id; }

extern "C" {

// returns the numeric ErrorCode (0 = OK, 203 = INVALID_INPUT)
int zlr_preprocess(const uint8_t* img, size_t len, int width, int height, int mw, int mh, float* out_chw)
{
    OnnxInferenceEngine e;
    e.config_.detection.model_width = mw;
    e.config_.detection.model_height = mh;
    std::vector<uint8_t> data(img, img + len);
    ReusableBuffer<float> buf;
    auto r = e.preProcess(data, width, height, buf);
    if (r.hasError()) return (int)r.error().code;
    std::memcpy(out_chw, r.value().data(), r.value().size() * sizeof(float));
    return 0;
}

// raw: [4+nc][A] fp32; returns the kept count (or -ErrorCode), detections in the reference's output order
int zlr_postprocess(const float* raw, int nc, int A, int img_w, int img_h, float conf_thr, float iou_thr, zlr_det* out)
{
    OnnxInferenceEngine e;
    e.config_.confidence_threshold = conf_thr;
    e.config_.nms_threshold = iou_thr;
    Ort::Value v;
    v.data = raw;
    v.shape = {1, (int64_t)(4 + nc), (int64_t)A};
    auto r = e.postProcess(v, img_w, img_h);
    if (r.hasError()) return -(int)r.error().code;
    const std::vector<Detection>& d = r.value();
    for (size_t i = 0; i < d.size(); ++i) put(d[i], out + i);
    return (int)d.size();
}

int zlr_nms(const zlr_det* in, int n, float iou_thr, zlr_det* out)
{
    OnnxInferenceEngine e;
    std::vector<Detection> d(n);
    for (int i = 0; i < n; ++i) {
        d[i].box = BoundingBox{in[i].x, in[i].y, in[i].w, in[i].h};
        d[i].confidence = in[i].confidence; d[i].class_id = in[i].class_id; d[i].track_id = 0; d[i].timestamp = 0;
    }
    std::vector<Detection> k = e.applyNMS(d, iou_thr);
    for (size_t i = 0; i < k.size(); ++i) put(k[i], out + i);
    return (int)k.size();
}

float zlr_iou(const float* a, const float* b)
{
    OnnxInferenceEngine e;
    return e.calculateIoU(BoundingBox{a[0], a[1], a[2], a[3]}, BoundingBox{b[0], b[1], b[2], b[3]});
}

int zlr_sizeof_detection(void) { return (int)sizeof(Detection); }

}  // extern "C"
