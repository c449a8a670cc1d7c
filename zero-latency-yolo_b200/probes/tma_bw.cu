// tma_bw.cu — measurement probe (NOT part of libzl_b200.so): chip-wide TMA load / store throughput as a function of
// the box geometry the persistent conv kernel uses (inner row bytes = channels x 2, rows per box, element stride),
// next to plain vectorised global stores.  Answers "is the conv kernel TMA-request-bound on narrow layers?".
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ../lib/tma_bw tma_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct Geo { int tiles_x, tiles_y, nimg, step_w, step_h; };

// One issuing thread per CTA, `stages` loads in flight, nobody consumes the data.
__global__ void __launch_bounds__(128, 1) load_kernel(const __grid_constant__ CUtensorMap map, Geo g, int stages, uint32_t box_bytes, uint32_t alloc, int iters, int cstep, int nchunk)
{
    extern __shared__ uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t bars = base, data = base + 1024u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(bars + 8u * s, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int total = g.tiles_x * g.tiles_y * g.nimg;
        int tile = blockIdx.x;
        for (int it = 0; it < iters; ++it) {
            const int s = it % stages, round = it / stages;
            if (round > 0) mbar_wait(bars + 8u * s, (uint32_t)(round - 1) & 1u);
            const int n = tile / (g.tiles_x * g.tiles_y), rem = tile - n * (g.tiles_x * g.tiles_y);
            const int ty = rem / g.tiles_x, tx = rem - ty * g.tiles_x;
            mbar_expect(bars + 8u * s, box_bytes);
            tma_load_4d(&map, bars + 8u * s, data + s * alloc, (it % nchunk) * cstep, tx * g.step_w - 1, ty * g.step_h - 1, n);
            if ((it % nchunk) == nchunk - 1) { tile += gridDim.x; if (tile >= total) tile -= total; }
        }
        for (int s = 0; s < stages && s < iters; ++s) {        // drain: the last use of every stage
            const int uses = (iters - s + stages - 1) / stages;
            mbar_wait(bars + 8u * s, (uint32_t)(uses - 1) & 1u);
        }
    }
}

// `nwarps` warps per CTA each own a stream of TMA stores from their own smem block, 4 bulk groups in flight per warp.
__global__ void __launch_bounds__(512, 1) store_kernel(const __grid_constant__ CUtensorMap map, Geo g, uint32_t alloc, int iters, int cstep, int nchunk)
{
    extern __shared__ uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    if (lane == 0) {
        const int total = g.tiles_x * g.tiles_y * g.nimg;
        int tile = (blockIdx.x * nwarps + warp) % total;
        for (int it = 0; it < iters; ++it) {
            const int n = tile / (g.tiles_x * g.tiles_y), rem = tile - n * (g.tiles_x * g.tiles_y);
            const int ty = rem / g.tiles_x, tx = rem - ty * g.tiles_x;
            tma_store_4d(&map, base + (uint32_t)(warp * 4 + (it & 3)) * alloc, (it % nchunk) * cstep, tx * g.step_w, ty * g.step_h, n);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            if ((it % nchunk) == nchunk - 1) { tile += gridDim.x * nwarps; if (tile >= total) tile -= total; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// Plain stores: every thread writes `vec16` x 16 bytes of "its pixel" (row pitch `pitch` bytes), warps walk the buffer.
__global__ void __launch_bounds__(512, 1) direct_store_kernel(uint8_t* y, size_t pixels, int pitch, int vec16, int iters)
{
    const size_t gthreads = (size_t)gridDim.x * blockDim.x;
    size_t px = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint4 v = make_uint4(threadIdx.x, 2, 3, 4);
    for (int it = 0; it < iters; ++it) {
        uint4* p = reinterpret_cast<uint4*>(y + px * (size_t)pitch);
        for (int k = 0; k < vec16; ++k) p[k] = v;
        px += gthreads;
        if (px >= pixels) px -= pixels;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode;

static void make_map(CUtensorMap* m, void* ptr, int N, int H, int W, int C, int pitchC, int bc, int bw, int bh, int estride, int swz)
{
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)pitchC * 2, (cuuint64_t)W * pitchC * 2, (cuuint64_t)H * W * pitchC * 2};
    cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)(bw * estride), (cuuint32_t)(bh * estride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    CUtensorMapSwizzle sw = swz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swz == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : swz == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main(int argc, char** argv)
{
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    g_encode = (EncodeTiledFn)fp;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, clk);
    const size_t cap = (size_t)1 << 30;   // 1 GiB scratch tensor
    uint8_t* buf; CK(cudaMalloc(&buf, cap)); CK(cudaMemset(buf, 1, cap));
    CK(cudaFuncSetAttribute(load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    struct LoadCfg { const char* name; int C, pitchC, HW, N, bc, bw, bh, estride, swz, stepw, steph, nchunk; };
    const LoadCfg lc[] = {
        // name                      C  pitch HW  N   bc  bw  bh es swz sw  sh  chunks
        {"ld 32B x10x66 (c16 3x3)", 16, 16, 320, 64, 16, 10, 66, 1, 32, 8, 64, 1},
        {"ld 32B x10x18",           16, 16, 320, 64, 16, 10, 18, 1, 32, 8, 16, 1},
        {"ld 64B x10x34 (c32 3x3)", 32, 32, 320, 32, 32, 10, 34, 1, 64, 8, 32, 1},
        {"ld 128B x10x18 (c64 3x3)", 64, 64, 160, 64, 64, 10, 18, 1, 128, 8, 16, 1},
        {"ld 128B x8x16 (c64 1x1)", 64, 64, 160, 64, 64, 8, 16, 1, 128, 8, 16, 1},
        {"ld 64B x8x32 (c32 1x1)",  32, 32, 320, 32, 32, 8, 32, 1, 64, 8, 32, 1},
        {"ld 32B x8x64 (c16 1x1)",  16, 16, 320, 64, 16, 8, 64, 1, 32, 8, 64, 1},
        {"ld 128B of c128 2 chunks", 128, 128, 160, 32, 64, 10, 18, 1, 128, 8, 16, 2},
        {"ld 32B of c80 5 chunks",  80, 80, 160, 32, 16, 10, 18, 1, 32, 8, 16, 5},
        {"ld 32B slice pitch48",    16, 48, 320, 32, 16, 10, 66, 1, 32, 8, 64, 1},
        {"ld s2 64B x9x17 es2",     32, 32, 320, 32, 32, 9, 17, 2, 64, 16, 32, 1},
        {"ld s2 32B x9x33 es2",     16, 16, 320, 64, 16, 9, 33, 2, 32, 16, 64, 1},
        {"ld s2 128B x9x17 es2",    64, 64, 160, 64, 64, 9, 17, 2, 128, 16, 32, 1},
        {"ld 128B x10x18 L2-res",   64, 64, 80, 32, 64, 10, 18, 1, 128, 8, 16, 1},
        {"ld 32B x10x66 L2-res",    16, 16, 160, 32, 16, 10, 66, 1, 32, 8, 64, 1},
    };
    printf("\n%-28s %5s %6s %8s %9s %9s %8s %9s\n", "TMA loads (1 thread/CTA)", "grid", "stages", "box B", "us", "GB/s", "B/clk/SM", "clk/box");
    for (const LoadCfg& c : lc) {
        for (int grid : {1, sms}) {
            for (int stages : {1, 2, 4, 8}) {
                if (grid == sms && (stages == 1 || stages == 2)) continue;
                CUtensorMap m;
                make_map(&m, buf, c.N, c.HW, c.HW, c.C, c.pitchC, c.bc, c.bw, c.bh, c.estride, c.swz);
                const uint32_t box_bytes = (uint32_t)c.bc * 2 * c.bw * c.bh;
                const uint32_t alloc = (box_bytes + 1023u) & ~1023u;
                if (1024u + 1024u + (uint32_t)stages * alloc > 227u * 1024u) continue;
                // grid 1: walk only the first image's tiles (L2-resident after the warm-up pass)
                Geo g{c.HW / c.stepw, c.HW / c.steph, grid == 1 ? 1 : c.N, c.stepw, c.steph};
                const size_t total_bytes_tensor = (size_t)c.N * c.HW * c.HW * c.pitchC * 2;
                if (total_bytes_tensor > cap) { printf("%s: tensor too large\n", c.name); continue; }
                const size_t smem = 2048 + 1024 + (size_t)stages * alloc;
                int iters = (int)((size_t)3 * 1024 * 1024 * 1024 / ((size_t)sms * box_bytes));
                if (iters > 20000) iters = 20000;
                if (grid == 1) iters = 4000;
                iters = iters / c.nchunk * c.nchunk;
                load_kernel<<<grid, 128, smem>>>(m, g, stages, box_bytes, alloc, grid == 1 ? iters : iters / 8 + c.nchunk, c.bc, c.nchunk);
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                load_kernel<<<grid, 128, smem>>>(m, g, stages, box_bytes, alloc, iters, c.bc, c.nchunk);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                const double bytes = (double)grid * iters * box_bytes;
                printf("%-28s %5d %6d %8u %9.1f %9.1f %8.2f %9.0f\n", c.name, grid, stages, box_bytes, ms * 1e3, bytes / ms / 1e6, bytes / grid / (ms * 1e-3 * clk * 1e3),
                       ms * 1e-3 * clk * 1e3 / iters);
            }
        }
    }

    struct StoreCfg { const char* name; int C, pitchC, HW, N, bc, bw, bh, swz, warps, nchunk; };
    const StoreCfg sc[] = {
        {"st 32B 8x4 (1KB) 16w",    16, 16, 320, 64, 16, 8, 4, 32, 16, 1},
        {"st 32B 8x4 (1KB) 4w",     16, 16, 320, 64, 16, 8, 4, 32, 4, 1},
        {"st 32B 8x4 of c64 4chunks 16w", 64, 64, 160, 64, 16, 8, 4, 32, 16, 4},
        {"st 32B 8x4 of c32 2chunks 16w", 32, 32, 320, 32, 16, 8, 4, 32, 16, 2},
        {"st 64B 8x4 (2KB) 16w",    32, 32, 320, 32, 32, 8, 4, 64, 16, 1},
        {"st 64B 8x4 (2KB) 4w",     32, 32, 320, 32, 32, 8, 4, 64, 4, 1},
        {"st 128B 8x4 (4KB) 4w",    64, 64, 160, 64, 64, 8, 4, 128, 4, 1},
        {"st 128B 8x4 (4KB) 8w",    64, 64, 160, 64, 64, 8, 4, 128, 8, 1},
        {"st 128B 8x16 (16KB) 2w",  64, 64, 160, 64, 64, 8, 16, 128, 2, 1},
        {"st 32B 8x16 (4KB) 4w",    16, 16, 320, 64, 16, 8, 16, 32, 4, 1},
        {"st 32B 8x64 (16KB) 2w",   16, 16, 320, 64, 16, 8, 64, 32, 2, 1},
        {"st 64B 8x32 (16KB) 2w",   32, 32, 320, 32, 32, 8, 32, 64, 2, 1},
    };
    printf("\n%-32s %8s %9s %9s %8s\n", "TMA stores", "box B", "us", "GB/s", "B/clk/SM");
    for (const StoreCfg& c : sc) {
        CUtensorMap m;
        make_map(&m, buf, c.N, c.HW, c.HW, c.C, c.pitchC, c.bc, c.bw, c.bh, 1, c.swz);
        const uint32_t box_bytes = (uint32_t)c.bc * 2 * c.bw * c.bh;
        const uint32_t alloc = (box_bytes + 1023u) & ~1023u;
        const size_t smem = 2048 + (size_t)c.warps * 4 * alloc;
        if (smem > 227 * 1024) { printf("%s: smem too large\n", c.name); continue; }
        Geo g{c.HW / c.bw, c.HW / c.bh, c.N, c.bw, c.bh};
        int iters = (int)((size_t)2 * 1024 * 1024 * 1024 / ((size_t)sms * c.warps * box_bytes));
        if (iters > 40000) iters = 40000;
        iters = iters / c.nchunk * c.nchunk;
        store_kernel<<<sms, c.warps * 32, smem>>>(m, g, alloc, iters / 8 + c.nchunk, c.bc, c.nchunk);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        store_kernel<<<sms, c.warps * 32, smem>>>(m, g, alloc, iters, c.bc, c.nchunk);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double bytes = (double)sms * c.warps * iters * box_bytes;
        printf("%-32s %8u %9.1f %9.1f %8.2f\n", c.name, box_bytes, ms * 1e3, bytes / ms / 1e6, bytes / sms / (ms * 1e-3 * clk * 1e3));
    }

    printf("\n%-32s %9s %9s\n", "plain st.global.v4", "us", "GB/s");
    struct DCfg { const char* name; int pitch, vec16; };
    const DCfg dc[] = {{"32B/thread pitch 32", 32, 2}, {"32B/thread pitch 128", 128, 2}, {"32B/thread pitch 64", 64, 2}, {"128B/thread pitch 128", 128, 8}, {"64B/thread pitch 64", 64, 4}};
    for (const DCfg& c : dc) {
        const size_t pixels = cap / c.pitch / 2;
        const int iters = 64;
        direct_store_kernel<<<sms, 512>>>(buf, pixels, c.pitch, c.vec16, 8);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        direct_store_kernel<<<sms * 4, 512>>>(buf, pixels, c.pitch, c.vec16, iters);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double bytes = (double)sms * 4 * 512 * iters * c.vec16 * 16;
        printf("%-32s %9.1f %9.1f\n", c.name, ms * 1e3, bytes / ms / 1e6);
    }
    printf("done\n");
    return 0;
}
