// weights.cpp — ZLW1 and ONNX-initialiser readers (see weights.h).
//
// ONNX is protobuf; only four message types matter here and they are walked with a ~60-line wire-format reader
// (no protobuf library offline):
//   ModelProto   : field 7  = graph (GraphProto)
//   GraphProto   : field 5  = initializer (repeated TensorProto)
//   TensorProto  : field 1  = dims (repeated int64, packed or not), 2 = data_type (1 = FLOAT, 10 = FLOAT16),
//                  4 = float_data (packed), 8 = name, 9 = raw_data
// An ultralytics export has BN already fused into the convs, and keeps the module names
// ("model.0.conv.weight" ... "model.22.cv3.2.2.bias"), which is exactly the naming of the ZLW1 container.
#include "weights.h"

#include <cstring>

#include <cuda_fp16.h>

namespace zl {
namespace {

// Element count of a tensor whose dims come from the FILE: every dim in 1..2^31, and the running product never beyond
// `limit` (what the container could possibly hold), so the byte-size comparisons below cannot wrap.
bool checked_count(const std::vector<uint32_t>& dims, size_t limit, size_t* out)
{
    size_t cnt = 1;
    for (uint32_t d : dims) {
        if (d == 0 || d > 0x7fffffffu) return false;
        if (cnt > limit / d) return false;
        cnt *= d;
    }
    *out = cnt;
    return true;
}

struct Reader {
    const uint8_t* p; const uint8_t* end; bool ok = true;
    bool more() const { return ok && p < end; }
    uint64_t varint() {
        uint64_t v = 0; int shift = 0;
        while (p < end && shift < 64) {
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
            shift += 7;
        }
        ok = false; return 0;
    }
    // returns field number, sets wire type; for length-delimited fields sets [sub, sub_end)
    uint32_t field(uint32_t* wt, const uint8_t** sub, const uint8_t** sub_end, uint64_t* val) {
        const uint64_t key = varint();
        *wt = (uint32_t)(key & 7);
        switch (*wt) {
            case 0: *val = varint(); break;
            case 1: if (end - p < 8) { ok = false; break; } std::memcpy(val, p, 8); p += 8; break;
            case 5: if (end - p < 4) { ok = false; break; } { uint32_t v32; std::memcpy(&v32, p, 4); *val = v32; } p += 4; break;
            case 2: { const uint64_t n = varint(); if (!ok || (uint64_t)(end - p) < n) { ok = false; break; } *sub = p; *sub_end = p + n; p += n; break; }
            default: ok = false;
        }
        return (uint32_t)(key >> 3);
    }
};

bool parse_tensor(const uint8_t* b, const uint8_t* e, std::string* name, HostTensor* t)
{
    Reader r{b, e};
    int data_type = 0;
    const uint8_t *raw = nullptr, *raw_end = nullptr, *fd = nullptr, *fd_end = nullptr;
    std::vector<float> unpacked;
    while (r.more()) {
        uint32_t wt; const uint8_t *s = nullptr, *se = nullptr; uint64_t v = 0;
        const uint32_t f = r.field(&wt, &s, &se, &v);
        if (!r.ok) return false;
        if (f == 1) {
            // ONNX dims are int64: anything that does not fit a positive 31-bit value is rejected, never truncated
            if (wt == 0) { if (v == 0 || v > 0x7fffffffull) return false; t->dims.push_back((uint32_t)v); }
            else if (wt == 2) {
                Reader d{s, se};
                while (d.more()) { const uint64_t dv = d.varint(); if (!d.ok || dv == 0 || dv > 0x7fffffffull) return false; t->dims.push_back((uint32_t)dv); }
                if (!d.ok) return false;
            }
            if (t->dims.size() > 8) return false;
        } else if (f == 2 && wt == 0) data_type = (int)v;
        else if (f == 8 && wt == 2) name->assign((const char*)s, (size_t)(se - s));
        else if (f == 9 && wt == 2) { raw = s; raw_end = se; }
        else if (f == 4 && wt == 2) { fd = s; fd_end = se; }
        else if (f == 4 && wt == 5) { float x; const uint32_t u = (uint32_t)v; std::memcpy(&x, &u, 4); unpacked.push_back(x); }
    }
    size_t cnt = 0;
    if (!checked_count(t->dims, (size_t)(e - b), &cnt)) {     // a tensor cannot have more elements than its message has bytes
        if (data_type == 1 || data_type == 10) return false;
        t->data.clear();
        return true;
    }
    if (data_type == 1) {
        if (raw && (size_t)(raw_end - raw) == cnt * 4) { t->data.resize(cnt); std::memcpy(t->data.data(), raw, cnt * 4); }
        else if (fd && (size_t)(fd_end - fd) == cnt * 4) { t->data.resize(cnt); std::memcpy(t->data.data(), fd, cnt * 4); }
        else if (unpacked.size() == cnt) t->data = unpacked;
        else return false;
    } else if (data_type == 10) {                 // FLOAT16 export (half=True)
        if (!raw || (size_t)(raw_end - raw) != cnt * 2) return false;
        t->data.resize(cnt);
        for (size_t i = 0; i < cnt; ++i) { __half_raw hr; std::memcpy(&hr.x, raw + 2 * i, 2); t->data[i] = __half2float(__half(hr)); }
    } else {
        t->data.clear();                          // other initialisers (shapes, int64 constants) are not weights
    }
    return true;
}

int32_t parse_onnx(const uint8_t* p, size_t len, ParsedModel* out)
{
    Reader m{p, p + len};
    bool saw_graph = false;
    while (m.more()) {
        uint32_t wt; const uint8_t *s = nullptr, *se = nullptr; uint64_t v = 0;
        const uint32_t f = m.field(&wt, &s, &se, &v);
        if (!m.ok) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "not a ZLW1 container and not parseable as ONNX");
        if (f != 7 || wt != 2) continue;
        saw_graph = true;
        Reader g{s, se};
        while (g.more()) {
            uint32_t gwt; const uint8_t *gs = nullptr, *gse = nullptr; uint64_t gv = 0;
            const uint32_t gf = g.field(&gwt, &gs, &gse, &gv);
            if (!g.ok) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "corrupt ONNX graph");
            if (gf != 5 || gwt != 2) continue;
            std::string name; HostTensor t;
            if (!parse_tensor(gs, gse, &name, &t)) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "corrupt ONNX initializer");
            if (!t.data.empty() && name.rfind("model.", 0) == 0 && name.find(".dfl.") == std::string::npos)
                out->tensors[name] = std::move(t);
        }
    }
    if (!saw_graph || out->tensors.empty()) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "not a ZLW1 container and no YOLOv8 initializers found in the ONNX graph");
    auto w0 = out->tensors.find("model.0.conv.weight"), wc = out->tensors.find("model.22.cv3.0.2.weight");
    if (w0 == out->tensors.end() || wc == out->tensors.end() || w0->second.dims.size() != 4 || wc->second.dims.size() != 4)
        ZL_FAIL(ZL_MODEL_LOAD_FAILED, "ONNX model does not look like an ultralytics YOLOv8 detect export (BN must be fused)");
    const uint32_t c1 = w0->second.dims[0];
    out->scale = c1 == 16 ? ZL_SCALE_N : (c1 == 32 ? ZL_SCALE_S : (c1 == 48 ? ZL_SCALE_M : -1));
    if (out->scale < 0) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "unsupported YOLOv8 scale (first conv has " + std::to_string(c1) + " channels; n/s/m supported)");
    out->nc = (int)wc->second.dims[0];
    return ZL_OK;
}

int32_t parse_zlw(const uint8_t* p, size_t len, ParsedModel* out)
{
    uint32_t hdr[5];
    std::memcpy(hdr, p + 4, 20);
    if (hdr[0] != 1) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "unsupported ZLW version");
    out->scale = (int)hdr[1]; out->nc = (int)hdr[2];
    size_t off = 24;
    for (uint32_t i = 0; i < hdr[3]; ++i) {
        if (off + 4 > len) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "truncated weights container");
        uint32_t nl; std::memcpy(&nl, p + off, 4); off += 4;
        if (off + nl > len || nl > 4096) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "truncated weights container");
        std::string name((const char*)p + off, nl); off += nl + ((4 - nl % 4) % 4);
        if (off + 4 > len) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "truncated weights container");
        uint32_t nd; std::memcpy(&nd, p + off, 4); off += 4;
        if (nd > 8 || off + 4ull * nd > len) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "bad tensor rank");
        HostTensor t; t.dims.resize(nd);
        for (uint32_t d = 0; d < nd; ++d) { std::memcpy(&t.dims[d], p + off, 4); off += 4; }
        size_t cnt = 0;
        if (!checked_count(t.dims, (len - off) / 4, &cnt)) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "truncated or oversized tensor " + name);
        t.data.resize(cnt);
        std::memcpy(t.data.data(), p + off, cnt * 4); off += cnt * 4;
        out->tensors[name] = std::move(t);
    }
    return ZL_OK;
}

}  // namespace

int32_t parse_model(const void* blob, size_t len, ParsedModel* out)
{
    const uint8_t* p = (const uint8_t*)blob;
    if (!p || len < 24) ZL_FAIL(ZL_MODEL_LOAD_FAILED, "model file too small");
    if (std::memcmp(p, "ZLW1", 4) == 0) return parse_zlw(p, len, out);
    return parse_onnx(p, len, out);
}

uint64_t model_checksum(const ParsedModel& m)
{
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* d, size_t n) { const uint8_t* b = (const uint8_t*)d; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
    for (const auto& kv : m.tensors) {
        mix(kv.first.data(), kv.first.size());
        mix(kv.second.dims.data(), kv.second.dims.size() * 4);
        mix(kv.second.data.data(), kv.second.data.size() * 4);
    }
    return h;
}

}  // namespace zl
