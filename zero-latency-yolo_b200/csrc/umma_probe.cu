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

// ---- TMA probe: one 4-D tiled load (optionally with element strides) into smem, dumped back to the host.
namespace zl {
namespace {
__global__ void tma_probe_kernel(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, int c3, uint32_t expect_bytes,
                                 uint32_t dump_bytes, uint8_t* out, int* status)
{
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* data = smem_raw + (base + 1024u - tc::smem_u32(smem_raw));
    for (uint32_t i = threadIdx.x; i < dump_bytes; i += blockDim.x) data[i] = 0xEE;
    if (threadIdx.x == 0) { tc::mbar_init(base, 1u); tc::fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        tc::fence_proxy_async();
        tc::mbar_arrive_expect_tx(base, expect_bytes);
        tc::tma_load_4d(&tmap, base, base + 1024u, c0, c1, c2, c3);
        int done = 0;
        for (int spin = 0; spin < 2000000; ++spin) if (tc::mbar_try_wait(base, 0u)) { done = 1; break; }
        *status = done;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < dump_bytes; i += blockDim.x) out[i] = data[i];
}
}  // namespace
}  // namespace zl

extern "C" ZL_API int32_t zl_probe_tma(int32_t device, const uint16_t* x, int32_t n, int32_t h, int32_t w, int32_t c,
                                       int32_t box_c, int32_t box_w, int32_t box_h, int32_t estride, int32_t swizzle_bytes,
                                       int32_t c0, int32_t c1, int32_t c2, int32_t c3, uint32_t expect_bytes,
                                       uint8_t* dump, uint32_t dump_bytes, int32_t* completed)
{
    using namespace zl;
    ZL_CUDA(cudaSetDevice(device));
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        ZL_FAIL(ZL_SYSTEM_ERROR, "no cuTensorMapEncodeTiled");
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    uint16_t* dx = nullptr; uint8_t* dout = nullptr; int* dstat = nullptr;
    const size_t nx = (size_t)n * h * w * c;
    ZL_CUDA(cudaMalloc(&dx, nx * 2)); ZL_CUDA(cudaMalloc(&dout, dump_bytes)); ZL_CUDA(cudaMalloc(&dstat, 4));
    ZL_CUDA(cudaMemcpy(dx, x, nx * 2, cudaMemcpyHostToDevice));
    ZL_CUDA(cudaMemset(dstat, 0, 4));
    CUtensorMap map;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = reinterpret_cast<Fn>(fnp)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { cudaFree(dx); cudaFree(dout); cudaFree(dstat); ZL_FAIL(ZL_SYSTEM_ERROR, "encode failed " + std::to_string((int)r)); }
    tma_probe_kernel<<<1, 128, dump_bytes + 4096>>>(map, c0, c1, c2, c3, expect_bytes, dump_bytes, dout, dstat);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) { cudaMemcpy(dump, dout, dump_bytes, cudaMemcpyDeviceToHost); cudaMemcpy(completed, dstat, 4, cudaMemcpyDeviceToHost); }
    cudaFree(dx); cudaFree(dout); cudaFree(dstat);
    if (e != cudaSuccess) ZL_FAIL(ZL_INFERENCE_ERROR, std::string("tma probe: ") + cudaGetErrorString(e));
    return ZL_OK;
}
