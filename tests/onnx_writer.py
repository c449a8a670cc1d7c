"""A minimal ONNX (protobuf) WRITER for the tests of the engine's ONNX initialiser import.
Only what an ultralytics YOLOv8 export needs to be recognised: ModelProto{ir_version, graph{initializer*}}.
Field numbers: ModelProto.ir_version=1, .graph=7; GraphProto.name=2, .initializer=5;
TensorProto.dims=1, .data_type=2, .float_data=4, .int64_data=7, .name=8, .raw_data=9."""
import struct

import numpy as np


def _varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field, wire):
    return _varint((field << 3) | wire)


def _ld(field, payload):
    return _key(field, 2) + _varint(len(payload)) + payload


def tensor(name, arr, mode="raw", packed_dims=False):
    """mode: raw (fp32 raw_data) | float_data (packed field 4) | fp16 (FLOAT16 raw_data) | int64 (a shape-like initialiser)."""
    arr = np.asarray(arr)
    body = b""
    if packed_dims:
        body += _ld(1, b"".join(_varint(int(d)) for d in arr.shape))
    else:
        for d in arr.shape:
            body += _key(1, 0) + _varint(int(d))
    if mode == "raw":
        body += _key(2, 0) + _varint(1) + _ld(8, name.encode()) + _ld(9, arr.astype("<f4").tobytes())
    elif mode == "float_data":
        body += _key(2, 0) + _varint(1) + _ld(4, arr.astype("<f4").tobytes()) + _ld(8, name.encode())
    elif mode == "fp16":
        body += _key(2, 0) + _varint(10) + _ld(8, name.encode()) + _ld(9, arr.astype("<f2").tobytes())
    elif mode == "int64":
        body += _key(2, 0) + _varint(7) + _ld(8, name.encode()) + _ld(7, b"".join(_varint(int(v)) for v in arr.reshape(-1)))
    else:
        raise ValueError(mode)
    return body


def model(tensors, mode="raw", extras=True):
    """tensors: {name: ndarray} (the ZLW1 tensor dict).  Returns the bytes of a ModelProto."""
    g = _ld(2, b"torch_jit")
    for i, (name, arr) in enumerate(sorted(tensors.items())):
        m = mode if mode != "mixed" else ("raw", "float_data")[i % 2]
        g += _ld(5, tensor(name, arr, m, packed_dims=(i % 3 == 0)))
    if extras:   # what a real export also carries and the importer must ignore
        g += _ld(5, tensor("model.22.dfl.conv.weight", np.arange(16, dtype=np.float32).reshape(1, 16, 1, 1)))
        g += _ld(5, tensor("/model.22/Constant_output_0", np.array([1, 4, 16, -1], np.int64), "int64"))
        g += _ld(5, tensor("onnx::Reshape_123", np.zeros(3, np.float32)))
    return _key(1, 0) + _varint(8) + _ld(2, b"pytorch") + _ld(7, g) + _ld(8, _ld(1, b"") + _key(2, 0) + _varint(17))
