// tma_rate.cu — measurement probe (NOT part of libzl_b200.so): what bounds ONE SM's TMA load rate?  tma_bw.cu showed a
// constant ~631 cycles per 4-D box from a single issuing thread, whatever the box size.  This probe varies the tensor
// map rank (2-D / 3-D / 4-D views of the same bytes), the number of issuing warps (each with its own stage ring), the
// L2 promotion mode and whether TMA stores run next to the loads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ../lib/tma_rate tma_rate.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try(bar, parity)) if (++spins > 100000000u) { printf("probe: mbarrier timeout block %d\n", (int)blockIdx.x); __trap(); }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// rank 2: [rows][C] box {bc, brows};  rank 3: [H][8][C] box {bc, 8, bh};  rank 4: [N][H][8][C] box {bc, 8, bh, 1};  rank 1: cp.async.bulk of box_bytes
// `nissue` warps issue independently (own barriers + own smem), `stages` loads in flight each.  store_warps extra warps
// stream 1 KB 2-D TMA stores concurrently.
__global__ void __launch_bounds__(512, 1) rate_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap smap, const uint8_t* raw,
                                                      int rank, int nissue, int store_warps, int stages, uint32_t box_bytes, uint32_t alloc,
                                                      int brows, int total_boxes, int iters, int prefetch, int bps, int lane_mode)
{
    extern __shared__ uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane_mode) {                       // issuers are lanes 0..nissue-1 of warp 0 instead of lane 0 of nissue warps
        if (warp != 0 || lane >= nissue) return;
        warp = lane;
    } else if (lane != 0) return;
    if (warp < nissue) {
        const uint32_t bars = base + 128u * warp, data = base + 2048u + (uint32_t)warp * stages * alloc;
        for (int s = 0; s < stages; ++s) mbar_init(bars + 8u * s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (prefetch) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map)) : "memory");
        int box = (blockIdx.x * nissue + warp) % total_boxes;
        for (int it = 0; it < iters; ++it) {
            const int s = it % stages, round = it / stages;
            if (round > 0) mbar_wait(bars + 8u * s, (uint32_t)(round - 1) & 1u);
            mbar_expect(bars + 8u * s, box_bytes * (uint32_t)bps);
            for (int b = 0; b < bps; ++b) {
                const uint32_t dst = data + s * alloc + (uint32_t)b * box_bytes;
                if (rank == 1) bulk_load_1d(dst, raw + (size_t)box * box_bytes, box_bytes, bars + 8u * s);
                else if (rank == 2) tma_load_2d(&map, bars + 8u * s, dst, 0, box * brows);
                else if (rank == 3) tma_load_3d(&map, bars + 8u * s, dst, 0, 0, box * brows);
                else tma_load_4d(&map, bars + 8u * s, dst, 0, 0, box * brows, 0);
                box += gridDim.x * nissue;
                if (box >= total_boxes) box -= total_boxes;
            }
        }
        for (int s = 0; s < stages && s < iters; ++s) {
            const int uses = (iters - s + stages - 1) / stages;
            mbar_wait(bars + 8u * s, (uint32_t)(uses - 1) & 1u);
        }
    } else if (warp < nissue + store_warps) {
        const int sw = warp - nissue;
        const uint32_t src = base + 2048u + (uint32_t)nissue * stages * alloc + (uint32_t)sw * 4096u;
        int row = (blockIdx.x * store_warps + sw) * 32;
        // run for roughly as long as the loaders: 12 stores per load iteration
        for (int it = 0; it < iters * 12; ++it) {
            tma_store_2d(&smap, src + (uint32_t)(it & 3) * 1024u, 0, row);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            row += gridDim.x * store_warps * 32;
            if (row >= 1 << 20) row -= 1 << 20;
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode;

static void make_map(CUtensorMap* m, void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, int swz, int promo)
{
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d (rank %d)\n", (int)r, rank); exit(1); }
}

int main()
{
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    g_encode = (EncodeTiledFn)fp;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const size_t cap = (size_t)1 << 30;
    uint8_t* buf; CK(cudaMalloc(&buf, cap)); CK(cudaMemset(buf, 1, cap));
    uint8_t* sbuf; CK(cudaMalloc(&sbuf, (size_t)(1 << 20) * 32)); 
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CUtensorMap smap;
    { cuuint64_t d[2] = {16, 1 << 20}; cuuint64_t st[1] = {32}; cuuint32_t b[2] = {16, 32}; make_map(&smap, sbuf, 2, d, st, b, 32, 0); }

    printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, clk);
    printf("%-4s %4s %5s %6s %6s %6s %4s %5s %4s %8s | %9s %9s %8s\n", "rank", "C", "rows", "grid", "issue", "stages", "bps", "lanes", "stw", "box B", "GB/s", "B/clk/SM", "clk/iter");
    struct Cfg { int rank, C, brows, grid, nissue, stages, promo, pref, stw, bps, lanes; };
    const Cfg cfgs[] = {
        // rank C rows grid issue stages promo pref stw bps lanes
        {2, 64, 64, 1, 1, 4, 1, 0, 0, 1, 0}, {2, 64, 64, 1, 1, 4, 1, 0, 0, 2, 0}, {2, 64, 64, 1, 1, 4, 1, 0, 0, 4, 0}, {2, 64, 32, 1, 1, 4, 1, 0, 0, 8, 0},
        {4, 64, 8, 1, 1, 4, 1, 0, 0, 1, 0}, {4, 64, 8, 1, 1, 4, 1, 0, 0, 2, 0}, {4, 64, 8, 1, 1, 4, 1, 0, 0, 4, 0},
        {2, 64, 64, 1, 2, 4, 1, 0, 0, 1, 1}, {2, 64, 64, 1, 4, 4, 1, 0, 0, 1, 1}, {2, 64, 64, 1, 8, 2, 1, 0, 0, 1, 1},
        {2, 64, 64, 1, 2, 4, 1, 0, 0, 1, 0}, {2, 64, 64, 1, 4, 4, 1, 0, 0, 1, 0}, {2, 64, 64, 1, 8, 2, 1, 0, 0, 1, 0},
        {2, 64, 64, 1, 1, 1, 1, 0, 0, 1, 0}, {2, 64, 64, 1, 1, 2, 1, 0, 0, 1, 0}, {2, 64, 64, 1, 1, 8, 1, 0, 0, 1, 0}, {2, 64, 64, 1, 1, 16, 1, 0, 0, 1, 0},
        {2, 64, 64, 148, 1, 8, 1, 0, 0, 1, 0}, {2, 64, 64, 148, 1, 8, 1, 0, 0, 2, 0}, {2, 64, 64, 148, 4, 4, 1, 0, 0, 1, 0}, {2, 64, 64, 148, 4, 4, 1, 0, 0, 1, 1},
        {4, 64, 8, 1, 1, 4, 1, 0, 2, 1, 0}, {4, 64, 8, 1, 1, 4, 1, 0, 8, 1, 0},
    };
    for (const Cfg& c : cfgs) {
        // the same bytes under every rank: pixel rows of C channels, an "image" is 8 pixels wide
        const uint32_t box_bytes = (uint32_t)c.C * 2 * (c.rank >= 3 ? 8 * c.brows : c.brows);
        const int rows_per_box = c.rank >= 3 ? 8 * c.brows : c.brows;
        const size_t tensor_rows = c.grid == 1 ? (size_t)(4 << 20) / (c.C * 2) : cap / (c.C * 2);    // grid 1: 4 MB, L2-resident
        const int total_boxes = (int)(tensor_rows / rows_per_box);
        CUtensorMap m;
        if (c.rank == 2 || c.rank == 1) { cuuint64_t d[2] = {(cuuint64_t)c.C, tensor_rows}; cuuint64_t st[1] = {(cuuint64_t)c.C * 2}; cuuint32_t b[2] = {(cuuint32_t)c.C, (cuuint32_t)c.brows}; make_map(&m, buf, 2, d, st, b, c.C * 2, c.promo); }
        else if (c.rank == 3) { cuuint64_t d[3] = {(cuuint64_t)c.C, 8, tensor_rows / 8}; cuuint64_t st[2] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.C * 16}; cuuint32_t b[3] = {(cuuint32_t)c.C, 8, (cuuint32_t)c.brows}; make_map(&m, buf, 3, d, st, b, c.C * 2, c.promo); }
        else { cuuint64_t d[4] = {(cuuint64_t)c.C, 8, tensor_rows / 8, 1}; cuuint64_t st[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.C * 16, (cuuint64_t)tensor_rows * c.C * 2}; cuuint32_t b[4] = {(cuuint32_t)c.C, 8, (cuuint32_t)c.brows, 1}; make_map(&m, buf, 4, d, st, b, c.C * 2, c.promo); }
        const uint32_t alloc = (box_bytes * (uint32_t)c.bps + 1023u) & ~1023u;
        const size_t smem = 1024 + 2048 + (size_t)c.nissue * c.stages * alloc + (size_t)c.stw * 4096;
        if (smem > 227 * 1024) { printf("skip (smem)\n"); continue; }
        const int iters = c.grid == 1 ? 4000 : 3000;
        const int brows_coord = c.rank >= 3 ? c.brows : c.brows;
        const int threads = ((c.lanes ? 1 : c.nissue) + c.stw) * 32;
        rate_kernel<<<c.grid, threads, smem>>>(m, smap, buf, c.rank, c.nissue, c.stw, c.stages, box_bytes, alloc, brows_coord, total_boxes, iters, c.pref, c.bps, c.lanes);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        rate_kernel<<<c.grid, threads, smem>>>(m, smap, buf, c.rank, c.nissue, c.stw, c.stages, box_bytes, alloc, brows_coord, total_boxes, iters, c.pref, c.bps, c.lanes);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double bytes = (double)c.grid * c.nissue * iters * box_bytes * c.bps;
        printf("%-4d %4d %5d %6d %6d %6d %4d %5d %4d %8u | %9.1f %9.2f %8.0f\n", c.rank, c.C, rows_per_box, c.grid, c.nissue, c.stages, c.bps, c.lanes, c.stw, box_bytes,
               bytes / ms / 1e6, bytes / c.grid / (ms * 1e-3 * clk * 1e3), ms * 1e-3 * clk * 1e3 / iters);
    }
    printf("done\n");
    return 0;
}
