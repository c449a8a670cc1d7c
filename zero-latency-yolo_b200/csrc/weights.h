// weights.h — model containers the engine can read (host only, no CUDA): the repo's ZLW1 container and the
// initialisers of an ultralytics YOLOv8 ONNX export (the file the reference loads:
// start.sh:122-125, src/inference/onnx_engine.cpp:977).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.h"

namespace zl {

struct HostTensor {
    std::vector<uint32_t> dims;
    std::vector<float> data;
};

struct ParsedModel {
    int scale = -1;        // ZL_SCALE_* (ZLW1: from the header; ONNX: from model.0.conv.weight's Cout)
    int nc = -1;           // classes (ONNX: Cout of model.22.cv3.0.2.weight)
    std::map<std::string, HostTensor> tensors;
};

// Detects the container by its magic ("ZLW1") and otherwise parses ONNX protobuf.  Returns ZL_OK or
// ZL_MODEL_LOAD_FAILED with zl_last_error() set.
int32_t parse_model(const void* blob, size_t len, ParsedModel* out);
// FNV-1a over (name, dims, data) of every conv weight/bias tensor in name order: equal for a ZLW1 file and an
// ONNX file holding the same parameters.
uint64_t model_checksum(const ParsedModel& m);

}  // namespace zl
