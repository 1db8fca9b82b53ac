// regblock_lab.cu -- register-block transposing copy for config 4a (out[i][j][k] = in[k][j][i]):
// each lane loads CH 16-byte vectors along i (rows k..k+3) and stores 4 vectors along k (rows i..i+3);
// the transpose is register renaming, no shared memory.  Development probe, not part of the library.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// LI lanes along i (x4 floats), 32/LI lanes along k (x4 floats); UN units per thread stacked along k.
// block = 256 threads = 8 warps stacked along k.
template <int LI, int UN, bool EXP>
__global__ void __launch_bounds__(256) regblock(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ v,
                                                int n_i, int n_j, int n_k, int order) {
    constexpr int LK = 32 / LI;
    constexpr int WI = LI * 4;              // i extent of a warp unit
    constexpr int WK = LK * 4;              // k extent of a warp unit
    constexpr int BK = WK * UN * 8;         // k extent of a block tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tiles_i = n_i / WI, tiles_k = n_k / BK;
    const uint32_t tiles = tiles_i * tiles_k * n_j;
    for (uint32_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        uint32_t b = t, ti, tk, j;
        if (order == 0) { ti = b % tiles_i; b /= tiles_i; tk = b % tiles_k; b /= tiles_k; j = b; }
        else { tk = b % tiles_k; b /= tiles_k; ti = b % tiles_i; b /= tiles_i; j = b; }
        const int i0 = ti * WI + (lane % LI) * 4;
        float4 q[UN][4];
#pragma unroll
        for (int un = 0; un < UN; ++un) {
            const int k0 = tk * BK + (warp * UN + un) * WK + (lane / LI) * 4;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                q[un][kk] = *reinterpret_cast<const float4*>(in + ((size_t)(k0 + kk) * n_j + j) * n_i + i0);
        }
#pragma unroll
        for (int un = 0; un < UN; ++un) {
            const int k0 = tk * BK + (warp * UN + un) * WK + (lane / LI) * 4;
            float4 vv = make_float4(0, 0, 0, 0);
            if (EXP) vv = *reinterpret_cast<const float4*>(v + k0);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float4 w;
                w.x = (&q[un][0].x)[u]; w.y = (&q[un][1].x)[u]; w.z = (&q[un][2].x)[u]; w.w = (&q[un][3].x)[u];
                if (EXP) { w.x = expf(w.x) + vv.x; w.y = expf(w.y) + vv.y; w.z = expf(w.z) + vv.z; w.w = expf(w.w) + vv.w; }
                *reinterpret_cast<float4*>(out + ((size_t)(i0 + u) * n_j + j) * n_k + k0) = w;
            }
        }
    }
}

template <int LI, int UN, bool EXP>
static void run(const float* in, float* out, const float* v, int n_i, int n_j, int n_k, int bps, int order) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int grid = 148 * bps;
    for (int w = 0; w < 3; ++w) regblock<LI, UN, EXP><<<grid, 256>>>(in, out, v, n_i, n_j, n_k, order);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int w = 0; w < 10; ++w) regblock<LI, UN, EXP><<<grid, 256>>>(in, out, v, n_i, n_j, n_k, order);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 10;
    printf("regblock LI=%d UN=%d exp=%d bps=%2d order=%d  %.3f ms  %7.1f GB/s\n", LI, UN, (int)EXP, bps, order, ms, 8.0 * n_i * n_j * n_k / ms / 1e6);
    if (!EXP) {
        float a, b2; int bad = 0;
        for (int s = 0; s < 50; ++s) {
            size_t i = (s * 7919) % n_i, j = (s * 104729) % n_j, k = (s * 31) % n_k;
            CK(cudaMemcpy(&a, in + (k * n_j + j) * n_i + i, 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&b2, out + (i * n_j + j) * n_k + k, 4, cudaMemcpyDeviceToHost));
            if (a != b2) ++bad;
        }
        if (bad) printf("   MISMATCH %d\n", bad);
    }
    fflush(stdout);
}

__global__ void fill(float* p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = (float)(i % 1000003) * 1e-6f;
}

int main() {
    const int n_i = 1024, n_j = 1024, n_k = 256;
    const size_t n = (size_t)n_i * n_j * n_k;
    float *in, *out, *v;
    CK(cudaMalloc(&in, n * 4)); CK(cudaMalloc(&out, n * 4)); CK(cudaMalloc(&v, 4096));
    fill<<<1184, 256>>>(in, n); fill<<<1, 256>>>(v, 1024);
    CK(cudaDeviceSynchronize());
    for (int order = 0; order < 2; ++order)
        for (int bps : {4, 8, 16, 32}) {
            run<8, 1, false>(in, out, v, n_i, n_j, n_k, bps, order);
            run<4, 1, false>(in, out, v, n_i, n_j, n_k, bps, order);
            run<8, 2, false>(in, out, v, n_i, n_j, n_k, bps, order);
            run<4, 2, false>(in, out, v, n_i, n_j, n_k, bps, order);
        }
    for (int bps : {4, 8, 16}) {
        run<8, 1, true>(in, out, v, n_i, n_j, n_k, bps, 0);
        run<8, 2, true>(in, out, v, n_i, n_j, n_k, bps, 0);
        run<4, 2, true>(in, out, v, n_i, n_j, n_k, bps, 0);
    }
    return 0;
}
