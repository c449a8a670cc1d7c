#!/usr/bin/env bash
# Builds libzl_b200.so (sm_100a only) in-tree.  Usage: csrc/build.sh [-j N]
set -euo pipefail
cd "$(dirname "$0")"
OUT=../lib
mkdir -p "$OUT" .obj
NVCC=${NVCC:-nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xcompiler -Wall)
SRCS=(conv_tc.cu conv_halo.cu head_fused.cu conv_simt.cu preprocess.cu pool_upsample.cu postprocess.cu engine.cpp weights.cpp capi.cpp)
TEST_SRCS=(umma_probe.cu test_hooks.cpp)     # unit-test / measurement hooks: libzl_b200_test.so only
pids=()
for s in "${SRCS[@]}" "${TEST_SRCS[@]}"; do
  o=.obj/${s%.*}.o
  if [[ ! -f $o || $s -nt $o || kernels.h -nt $o || common.h -nt $o || tc_ptx.cuh -nt $o || head_math.cuh -nt $o || half16.cuh -nt $o || engine.h -nt $o || weights.h -nt $o || ../../include/zl_b200.h -nt $o || ../../include/zl_b200_test.h -nt $o ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$s" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
PROD=(); for s in "${SRCS[@]}"; do PROD+=(.obj/${s%.*}.o); done
TEST=(); for s in "${TEST_SRCS[@]}"; do TEST+=(.obj/${s%.*}.o); done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$OUT/libzl_b200.so" "${PROD[@]}" -lpthread -ldl -lrt
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$OUT/libzl_b200_test.so" "${PROD[@]}" "${TEST[@]}" -lpthread -ldl -lrt
echo "built $OUT/libzl_b200.so $OUT/libzl_b200_test.so"
