#!/usr/bin/env bash
# Builds libzl_b200.so (sm_100a only) in-tree.  Usage: csrc/build.sh [-j N]
set -euo pipefail
cd "$(dirname "$0")"
OUT=../lib
mkdir -p "$OUT" .obj
NVCC=${NVCC:-nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xcompiler -Wall)
SRCS=(conv_tc.cu conv_halo.cu umma_probe.cu conv_simt.cu preprocess.cu pool_upsample.cu postprocess.cu engine.cpp weights.cpp capi.cpp)
pids=()
for s in "${SRCS[@]}"; do
  o=.obj/${s%.*}.o
  if [[ ! -f $o || $s -nt $o || kernels.h -nt $o || common.h -nt $o || tc_ptx.cuh -nt $o || half16.cuh -nt $o || engine.h -nt $o || weights.h -nt $o || ../../include/zl_b200.h -nt $o ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$OUT/libzl_b200.so" .obj/*.o -lpthread -ldl -lrt
echo "built $OUT/libzl_b200.so"
