// tma_read_lab.cu -- what caps TMA read throughput on the config-4a pattern (rows of the box 4 MB apart)?
// Read-only kernels: tiles land in shared memory and are dropped.  Not part of the library.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
extern __shared__ __align__(16) unsigned char dyn_smem[];

struct P { const float* in; float* sink; int n_i, n_j, n_k, box_i, box_j, box_k, stages, kind; };

// One thread per block produces AND consumes (waits for the tile, then reuses the stage): pure TMA throughput.
// kind 0: tensor TMA; kind 1: 1-D bulk copies of box_i*4 bytes per (j,k) row.
__global__ void __launch_bounds__(32) read_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ P p) {
    const uint32_t bars = smem_u32(dyn_smem);
    const uint32_t ring = (bars + 256u + 1023u) & ~1023u;
    const uint32_t tile_bytes = p.box_i * p.box_j * p.box_k * 4;
    const uint32_t ti_n = p.n_i / p.box_i, tj_n = p.n_j / p.box_j, tk_n = p.n_k / p.box_k;
    const uint32_t tiles = ti_n * tj_n * tk_n;
    if (threadIdx.x != 0) return;
    for (int s = 0; s < p.stages; ++s) mbar_init(bars + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    int ps = 0, pph = 0, cs = 0, cph = 0;
    uint32_t issued = 0, done = 0;
    uint32_t t = blockIdx.x;
    const uint32_t mine = (tiles > blockIdx.x) ? (tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    while (done < mine) {
        while (issued < mine && issued - done < (uint32_t)p.stages) {
            uint32_t b = t;
            const uint32_t ti = b % ti_n; b /= ti_n;
            const uint32_t tk = b % tk_n; b /= tk_n;
            const uint32_t tj = b;
            mbar_expect(bars + 8 * ps, tile_bytes);
            if (p.kind == 0) {
                tma_load_3d(ring + ps * tile_bytes, &tm, ti * p.box_i, tj * p.box_j, tk * p.box_k, bars + 8 * ps);
            } else {
                uint32_t dst = ring + ps * tile_bytes;
                for (int k = 0; k < p.box_k; ++k)
                    for (int j = 0; j < p.box_j; ++j) {
                        bulk_load(dst, p.in + ((size_t)(tk * p.box_k + k) * p.n_j + tj * p.box_j + j) * p.n_i + ti * p.box_i, p.box_i * 4, bars + 8 * ps);
                        dst += p.box_i * 4;
                    }
            }
            if (++ps == p.stages) { ps = 0; pph ^= 1; }
            ++issued; t += gridDim.x;
        }
        mbar_wait(bars + 8 * cs, cph);
        if (++cs == p.stages) { cs = 0; cph ^= 1; }
        ++done;
    }
}

// LDG comparison: warp reads 128B..512B contiguous per row, rows 4 MB apart, 4 x 16 B in flight per thread
__global__ void __launch_bounds__(256) ldg_kernel(const float4* __restrict__ in, float* sink, int n_i4, int n_j, int n_k, int lanes_i) {
    // work item = (k, j, i-chunk of lanes_i float4); a warp covers 32/lanes_i consecutive k at one (j, chunk)
    const int lane = threadIdx.x & 31;
    const size_t chunks_i = n_i4 / lanes_i;
    const size_t groups_k = n_k / (32 / lanes_i * 4);
    const size_t total = chunks_i * n_j * groups_k;
    float acc = 0;
    for (size_t w = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / 32; w < total; w += (size_t)gridDim.x * blockDim.x / 32) {
        size_t b = w;
        const size_t ci = b % chunks_i; b /= chunks_i;
        const size_t gk = b % groups_k; b /= groups_k;
        const size_t j = b;
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t k = (gk * 4 + u) * (32 / lanes_i) + lane / lanes_i;
            v[u] = in[(k * n_j + j) * n_i4 + ci * lanes_i + lane % lanes_i];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) *sink = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn encoder() {
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    return (EncodeFn)f;
}

static void run(float* in, float* sink, int n_i, int n_j, int n_k, int bi, int bj, int bk, int stages, int bps, int kind, int swz) {
    CUtensorMap tm;
    cuuint64_t dims[3] = {(cuuint64_t)n_i, (cuuint64_t)n_j, (cuuint64_t)n_k};
    cuuint64_t strides[2] = {(cuuint64_t)n_i * 4, (cuuint64_t)n_i * n_j * 4};
    cuuint32_t box[3] = {(cuuint32_t)bi, (cuuint32_t)bj, (cuuint32_t)bk}, es[3] = {1, 1, 1};
    CUresult r = encoder()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           (CUtensorMapSwizzle)swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode failed %d (bi=%d swz=%d)\n", (int)r, bi, swz); return; }
    P p{in, sink, n_i, n_j, n_k, bi, bj, bk, stages, kind};
    const unsigned smem = stages * bi * bj * bk * 4 + 2048;
    if (smem > 227 * 1024) { printf("skip smem\n"); return; }
    CK(cudaFuncSetAttribute(read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) read_kernel<<<148 * bps, 32, smem>>>(tm, p);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int w = 0; w < 5; ++w) read_kernel<<<148 * bps, 32, smem>>>(tm, p);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
    printf("kind=%d box=(%4d i,%2d j,%3d k) %6d B/tile stages=%d bps=%d swz=%d  %.3f ms  %7.1f GB/s\n", kind, bi, bj, bk, bi * bj * bk * 4,
           stages, bps, swz, ms, 4.0 * n_i * n_j * n_k / ms / 1e6);
    fflush(stdout);
}

int main() {
    const int n_i = 1024, n_j = 1024, n_k = 256;
    const size_t n = (size_t)n_i * n_j * n_k;
    float *in, *sink;
    CK(cudaMalloc(&in, n * 4)); CK(cudaMalloc(&sink, 4096));
    CK(cudaMemset(in, 0, n * 4));
    // tensor TMA, the tiler's box, swizzle modes 0 (none) .. 3 (128B)
    for (int swz = 0; swz < 4; ++swz) run(in, sink, n_i, n_j, n_k, 32, 1, 128, 6, 2, 0, swz);
    // wider rows without swizzle
    run(in, sink, n_i, n_j, n_k, 64, 1, 64, 6, 2, 0, 0);
    run(in, sink, n_i, n_j, n_k, 128, 1, 32, 6, 2, 0, 0);
    run(in, sink, n_i, n_j, n_k, 256, 1, 16, 6, 2, 0, 0);
    run(in, sink, n_i, n_j, n_k, 256, 1, 64, 3, 1, 0, 0);
    // fully contiguous boxes (k extent 1): is the cap about row separation?
    run(in, sink, n_i, n_j, n_k, 256, 16, 1, 6, 2, 0, 0);
    run(in, sink, n_i, n_j, n_k, 32, 128, 1, 6, 2, 0, 3);
    run(in, sink, n_i, n_j, n_k, 32, 16, 8, 6, 2, 0, 3);
    // more blocks / deeper rings
    run(in, sink, n_i, n_j, n_k, 32, 1, 128, 3, 4, 0, 3);
    run(in, sink, n_i, n_j, n_k, 32, 1, 128, 12, 1, 0, 3);
    run(in, sink, n_i, n_j, n_k, 32, 1, 256, 6, 1, 0, 3);
    run(in, sink, n_i, n_j, n_k, 32, 1, 64, 6, 4, 0, 3);
    // 1-D bulk copies
    run(in, sink, n_i, n_j, n_k, 32, 1, 128, 6, 2, 1, 0);
    run(in, sink, n_i, n_j, n_k, 256, 1, 16, 6, 2, 1, 0);
    run(in, sink, n_i, n_j, n_k, 1024, 1, 4, 6, 2, 1, 0);
    run(in, sink, n_i, n_j, n_k, 1024, 4, 1, 6, 2, 1, 0);
    // LDG comparison
    for (int lanes_i = 8; lanes_i <= 32; lanes_i *= 2) {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int w = 0; w < 2; ++w) ldg_kernel<<<148 * 8, 256>>>((const float4*)in, sink, n_i / 4, n_j, n_k, lanes_i);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int w = 0; w < 5; ++w) ldg_kernel<<<148 * 8, 256>>>((const float4*)in, sink, n_i / 4, n_j, n_k, lanes_i);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 5;
        printf("LDG.128 rows of %4d B, 4 MB apart: %.3f ms  %7.1f GB/s\n", lanes_i * 16, ms, 4.0 * n / ms / 1e6);
    }
    return 0;
}
