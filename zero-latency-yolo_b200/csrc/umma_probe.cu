// umma_probe.cu — measurement hook: cycles per tcgen05.mma for a given (N, swizzle, SBO, accumulator
// rotation, A-shift) pattern.  Not on the product path; used to size the conv kernels' tiles.
#include <cuda_runtime.h>

#include "common.h"
#include "tc_ptx.cuh"

namespace zl {
namespace {
using namespace tc;

__global__ void __launch_bounds__(64, 1)
umma_probe_kernel(int N, int swz, int sbo_a, int nacc, int count, int shift_rows, int ksteps, long long* out)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar = base, tmem_slot = base + 16u;
    const uint32_t a_base = base + 1024u, b_base = a_base + 64u * 1024u;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    for (uint32_t i = threadIdx.x; i < (100u * 1024u) / 4u; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem_raw + (a_base - smem_u32(smem_raw)))[i] = 0u;
    if (threadIdx.x == 0) { mbar_init(bar, 1u); fence_barrier_init(); }
    if (threadIdx.x < 32) { tmem_alloc(tmem_slot, 512u); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        // descriptors are built outside the timed loop; the loop body is an add + the MMA itself
        const uint64_t adesc0 = make_smem_desc_sbo(a_base + (uint32_t)shift_rows * (uint32_t)swz, (uint32_t)swz, (uint32_t)sbo_a);
        const uint64_t bdesc0 = make_smem_desc_sbo(b_base, (uint32_t)swz, 8u * (uint32_t)swz);
        const uint32_t d0 = tmem_base, d1 = tmem_base + (uint32_t)(nacc > 1 ? N : 0);
        const uint64_t kadv = ksteps > 1 ? 2ull : 0ull;
        const long long t0 = clock64();
        for (int i = 0; i < count; i += 4) {
            umma_bf16(d0, adesc0, bdesc0, idesc, i > 0 ? 1u : 0u);
            umma_bf16(d1, adesc0 + kadv, bdesc0 + kadv, idesc, i > 0 ? 1u : 0u);
            umma_bf16(d0, adesc0 + 2 * kadv, bdesc0 + 2 * kadv, idesc, 1u);
            umma_bf16(d1, adesc0 + 3 * kadv, bdesc0 + 3 * kadv, idesc, 1u);
        }
        const long long t1 = clock64();
        umma_commit(bar);
        mbar_wait(bar, 0u);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem_base, 512u);
}
}  // namespace
}  // namespace zl

extern "C" ZL_API int32_t zl_probe_umma(int32_t device, int32_t N, int32_t swz, int32_t sbo_a, int32_t nacc, int32_t count,
                                        int32_t shift_rows, int32_t ksteps, int32_t grid, int64_t* issue_cycles, int64_t* total_cycles)
{
    using namespace zl;
    if (N < 16 || N > 256 || N % 16 || nacc < 1 || nacc * N > 512 || count < 1 || ksteps < 1) ZL_FAIL(ZL_INVALID_ARGUMENT, "bad probe argument");
    ZL_CUDA(cudaSetDevice(device));
    long long* d = nullptr;
    ZL_CUDA(cudaMalloc(&d, 16));
    ZL_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int rep = 0; rep < 2; ++rep) umma_probe_kernel<<<grid, 64, 110 * 1024>>>(N, swz, sbo_a, nacc, count, shift_rows, ksteps, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) ZL_FAIL(ZL_INFERENCE_ERROR, std::string("umma probe: ") + cudaGetErrorString(e));
    *issue_cycles = h[0];
    *total_cycles = h[1];
    return ZL_OK;
}
