// Lab driver for b200/scan_pipe.cuh: times the pipelined flat scan in several configurations against
// cub::DeviceScan (toolkit CCCL) and a plain copy, and checks results bit-exactly / within tolerance.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I../cupy_b200/csrc -I../cupy_b200/csrc/include
//        -I../include scan_lab2.cu -o scan_lab2.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cub/cub.cuh>
#include "common.h"
#include "include/b200/scan_pipe.cuh"
#include "tma_host.h"

namespace b200 {
thread_local char g_err[512];
int fail(int code, const char* fmt, ...) { (void)fmt; return code; }
}
using namespace b200;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <class Cfg, class Op, int MODE>
__global__ void __launch_bounds__(kPipeThreads, 1) pipe_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                               const __grid_constant__ CUtensorMap tm_out,
                                                               const typename Cfg::in_t* x, typename Cfg::out_t* y,
                                                               int64_t n_main, int64_t n,
                                                               typename PipeSlot<sizeof(typename Cfg::acc_t)>::storage_t* slots) {
    scan_pipe_body<Cfg, Op, MODE>(&tm_in, &tm_out, x, y, n_main, n, slots);
}

template <class Cfg, int MODE>
float run_pipe(const typename Cfg::in_t* x, typename Cfg::out_t* y, int64_t n, void* ws, int sm, int iters, int grid_override = 0) {
    typedef typename Cfg::in_t In; typedef typename Cfg::out_t Out;
    typedef typename PipeSlot<sizeof(typename Cfg::acc_t)>::storage_t slot_t;
    auto kern = pipe_kernel<Cfg, ScanSum, MODE>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    const int64_t n_main = n / Cfg::GRANULE * Cfg::GRANULE;
    CUtensorMap tin, tout;
    const uint64_t din[2] = {uint64_t(128 / sizeof(In)), uint64_t(n_main * sizeof(In) / 128)};
    const uint64_t dout[2] = {uint64_t(128 / sizeof(Out)), uint64_t(n_main * sizeof(Out) / 128)};
    const uint64_t strides[1] = {128};
    const uint32_t bin[2] = {uint32_t(128 / sizeof(In)), uint32_t(Cfg::IN_BOX)};
    const uint32_t bout[2] = {uint32_t(128 / sizeof(Out)), uint32_t(Cfg::OUT_BOX)};
    if (make_tensor_map(&tin, sizeof(In), x, 2, din, strides, bin, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("tmap in failed\n"); exit(1); }
    if (make_tensor_map(&tout, sizeof(Out), y, 2, dout, strides, bout, CU_TENSOR_MAP_SWIZZLE_128B)) { printf("tmap out failed\n"); exit(1); }
    const int64_t tiles = (n_main + Cfg::TILE - 1) / Cfg::TILE;
    int grid = int(std::min<int64_t>(tiles, grid_override ? grid_override : sm));
    slot_t* slots = static_cast<slot_t*>(ws);
    void* args[] = {&tin, &tout, &x, &y, (void*)&n_main, &n, &slots};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<float> ts;
    for (int i = 0; i < iters + 2; ++i) {
        CK(cudaMemsetAsync(ws, 0, size_t(kPipeRing) * grid * sizeof(slot_t), 0));
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kPipeThreads), args, Cfg::SMEM, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (i >= 2) ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    return ts[ts.size() / 2];
}

template <class T> __global__ void fill_rand(T* p, int64_t n, uint32_t seed, int lo, int hi) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
        uint32_t h = uint32_t(i) * 2654435761u ^ seed ^ uint32_t(i >> 32) * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        p[i] = T(lo + int(h % uint32_t(hi - lo)));
    }
}
template <class A> __device__ bool same_bits(const A& a, const A& b) {
    const unsigned char* p = reinterpret_cast<const unsigned char*>(&a);
    const unsigned char* q = reinterpret_cast<const unsigned char*>(&b);
    for (int k = 0; k < int(sizeof(A)); ++k) if (p[k] != q[k]) return false;
    return true;
}
template <class A, class B> __global__ void compare(const A* a, const B* b, int64_t n, unsigned long long* bad) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x)
        if (!same_bits(a[i], b[i])) atomicAdd(bad, 1ull);
}
template <class A, class B> __global__ void cast_copy(const A* a, B* b, int64_t n) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) b[i] = B(a[i]);
}

template <class In, class Out>
void reference(const In* x, Out* yref, int64_t n, void* tmp, size_t tmp_bytes, Out* xcast) {
    cast_copy<<<2048, 256>>>(x, xcast, n);
    size_t need = 0;
    cub::DeviceScan::InclusiveSum(nullptr, need, xcast, yref, n);
    if (need > tmp_bytes) { printf("cub tmp too small\n"); exit(1); }
    CK(cub::DeviceScan::InclusiveSum(tmp, need, xcast, yref, n));
}

template <class Cfg, int MODE = 0>
void bench(const char* name, int64_t n, const typename Cfg::in_t* x, typename Cfg::out_t* y, typename Cfg::out_t* yref, void* ws,
           int sm, double bytes_per_item, int grid_override = 0) {
    CK(cudaMemset(y, 0xff, size_t(n) * sizeof(typename Cfg::out_t)));
    float ms = run_pipe<Cfg, MODE>(x, y, n, ws, sm, 10, grid_override);
    unsigned long long* bad; CK(cudaMallocManaged(&bad, 8)); *bad = 0;
    if (MODE == 0) { compare<<<2048, 256>>>(y, yref, n, bad); CK(cudaDeviceSynchronize()); }
    printf("%-58s n=%lld  %8.3f ms  %8.1f GB/s  %s\n", name, (long long)n, ms, bytes_per_item * n / ms / 1e6,
           MODE == 0 ? (*bad ? "MISMATCH" : "exact") : "(not checked)");
    if (*bad) printf("    mismatches: %llu\n", *bad);
    fflush(stdout);
    CK(cudaFree(bad));
}

int main(int argc, char** argv) {
    int64_t n = (int64_t(1) << 28) + (argc > 1 ? atoll(argv[1]) : 0);
    int sm; CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0));
    void *ws, *tmp; CK(cudaMalloc(&ws, 1 << 20)); size_t tmp_bytes = 64 << 20; CK(cudaMalloc(&tmp, tmp_bytes));
    {   // ---------------- int64 -> int64
        typedef long long T;
        T *x, *y, *yref; CK(cudaMalloc(&x, n * 8)); CK(cudaMalloc(&y, n * 8)); CK(cudaMalloc(&yref, n * 8));
        fill_rand<<<2048, 256>>>(x, n, 1u, -(1 << 20), 1 << 20);
        size_t need = 0; cub::DeviceScan::InclusiveSum(nullptr, need, x, yref, n);
        CK(cub::DeviceScan::InclusiveSum(tmp, need, x, yref, n));
        if (argc > 2) {      // profiling mode: one configuration only
            bench<ScanPipeCfg<T, T, T, 8, 4, 2, 3>>("i64 pipe IPT8 SI4 SO2 LAG3", n, x, y, yref, ws, sm, 16);
            return 0;
        }
        bench<ScanPipeCfg<T, T, T, 8, 4, 2, 1>, 2>("i64 pipe copy-only  IPT8 SI4 SO2", n, x, y, yref, ws, sm, 16);
        bench<ScanPipeCfg<T, T, T, 8, 4, 2, 2>, 1>("i64 pipe no-exchange IPT8 SI4 SO2 LAG2", n, x, y, yref, ws, sm, 16);
        bench<ScanPipeCfg<T, T, T, 8, 4, 2, 2>>("i64 pipe IPT8 SI4 SO2 LAG2", n, x, y, yref, ws, sm, 16);
        bench<ScanPipeCfg<T, T, T, 8, 4, 2, 3>>("i64 pipe IPT8 SI4 SO2 LAG3", n, x, y, yref, ws, sm, 16);
        bench<ScanPipeCfg<T, T, T, 8, 4, 2, 4>>("i64 pipe IPT8 SI4 SO2 LAG4", n, x, y, yref, ws, sm, 16);
        bench<ScanPipeCfg<T, T, T, 8, 4, 3, 3>>("i64 pipe IPT8 SI4 SO3 LAG3", n, x, y, yref, ws, sm, 16);
        bench<ScanPipeCfg<T, T, T, 8, 3, 2, 3>>("i64 pipe IPT8 SI3 SO2 LAG3", n, x, y, yref, ws, sm, 16);
        for (int64_t m : {n - 1, n - 77, (int64_t(1) << 20) + 3, int64_t(4096 * 148 * 3 + 5)}) {
            CK(cub::DeviceScan::InclusiveSum(tmp, need, x, yref, m));
            bench<ScanPipeCfg<T, T, T, 8, 4, 2, 3>>("i64 pipe IPT8 SI4 SO2 LAG3 (ragged)", m, x, y, yref, ws, sm, 16);
        }
        CK(cudaFree(x)); CK(cudaFree(y)); CK(cudaFree(yref));
    }
    {   // ---------------- int32 -> int32 (same size, 4 bytes)
        typedef int T;
        T *x, *y, *yref; CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&y, n * 4)); CK(cudaMalloc(&yref, n * 4));
        fill_rand<<<2048, 256>>>(x, n, 2u, -3, 4);
        reference(x, yref, n, tmp, tmp_bytes, y);
        bench<ScanPipeCfg<T, T, T, 16, 4, 2, 2>>("i32 pipe IPT16 SI4 SO2 LAG2", n, x, y, yref, ws, sm, 8);
        bench<ScanPipeCfg<T, T, T, 16, 4, 2, 3>>("i32 pipe IPT16 SI4 SO2 LAG3", n, x, y, yref, ws, sm, 8);
        bench<ScanPipeCfg<T, T, T, 32, 2, 1, 2>>("i32 pipe IPT32 SI2 SO1 LAG2 (64 KB stages)", n, x, y, yref, ws, sm, 8);
        CK(cudaFree(x)); CK(cudaFree(y)); CK(cudaFree(yref));
    }
    {   // ---------------- int32 -> int64 (casting)
        int* x; long long *y, *yref, *xc;
        CK(cudaMalloc(&x, n * 4)); CK(cudaMalloc(&y, n * 8)); CK(cudaMalloc(&yref, n * 8)); CK(cudaMalloc(&xc, n * 8));
        fill_rand<<<2048, 256>>>(x, n, 3u, -1000, 1000);
        reference(x, yref, n, tmp, tmp_bytes, xc);
        bench<ScanPipeCfg<int, long long, long long, 8, 6, 2, 3>>("i32->i64 pipe IPT8 SI6 SO2 LAG3", n, x, y, yref, ws, sm, 12);
        bench<ScanPipeCfg<int, long long, long long, 16, 3, 2, 2>>("i32->i64 pipe IPT16 SI3 SO2 LAG2", n, x, y, yref, ws, sm, 12);
        bench<ScanPipeCfg<int, long long, long long, 16, 3, 2, 3>>("i32->i64 pipe IPT16 SI3 SO2 LAG3", n, x, y, yref, ws, sm, 12);
        CK(cudaFree(x));
        // ---------------- bool -> int64
        bool* xb; CK(cudaMalloc(&xb, n));
        fill_rand<<<2048, 256>>>(reinterpret_cast<unsigned char*>(xb), n, 4u, 0, 2);
        reference(reinterpret_cast<unsigned char*>(xb), yref, n, tmp, tmp_bytes, xc);
        bench<ScanPipeCfg<bool, long long, long long, 8, 8, 2, 3>>("bool->i64 pipe IPT8 SI8 SO2 LAG3", n, xb, y, yref, ws, sm, 9);
        bench<ScanPipeCfg<bool, long long, long long, 16, 6, 2, 3>>("bool->i64 pipe IPT16 SI6 SO2 LAG3", n, xb, y, yref, ws, sm, 9);
        bench<ScanPipeCfg<bool, long long, long long, 16, 4, 3, 3>>("bool->i64 pipe IPT16 SI4 SO3 LAG3", n, xb, y, yref, ws, sm, 9);
        bench<ScanPipeCfg<bool, long long, long long, 16, 4, 2, 2>>("bool->i64 pipe IPT16 SI4 SO2 LAG2", n, xb, y, yref, ws, sm, 9);
        CK(cudaFree(xb)); CK(cudaFree(y)); CK(cudaFree(yref)); CK(cudaFree(xc));
    }
    {   // ---------------- float16 -> float16 with a float accumulator: checked against the same kernel without
        //                  the cross-block exchange is not possible; compare with a float cub scan rounded to half
        __half *x, *y; float *xf, *yf; __half* yref;
        CK(cudaMalloc(&x, n * 2)); CK(cudaMalloc(&y, n * 2)); CK(cudaMalloc(&yref, n * 2)); CK(cudaMalloc(&xf, n * 4)); CK(cudaMalloc(&yf, n * 4));
        fill_rand<<<2048, 256>>>(xf, n, 5u, -2, 3);      // small integers: float sums are exact until they exceed 2^24
        cast_copy<<<2048, 256>>>(xf, x, n);
        size_t need = 0; cub::DeviceScan::InclusiveSum(nullptr, need, xf, yf, n);
        CK(cub::DeviceScan::InclusiveSum(tmp, need, xf, yf, n));
        cast_copy<<<2048, 256>>>(yf, yref, n);
        bench<ScanPipeCfg<float16, float, float16, 32, 4, 2, 3>>("f16 (float acc) pipe IPT32 SI4 SO2 LAG3", n, reinterpret_cast<float16*>(x),
                                                                 reinterpret_cast<float16*>(y), reinterpret_cast<float16*>(yref), ws, sm, 4);
        bench<ScanPipeCfg<float16, float, float16, 32, 4, 2, 2>>("f16 (float acc) pipe IPT32 SI4 SO2 LAG2", n, reinterpret_cast<float16*>(x),
                                                                 reinterpret_cast<float16*>(y), reinterpret_cast<float16*>(yref), ws, sm, 4);
        bench<ScanPipeCfg<float16, float, float16, 64, 2, 1, 2>>("f16 (float acc) pipe IPT64 SI2 SO1 LAG2", n, reinterpret_cast<float16*>(x),
                                                                 reinterpret_cast<float16*>(y), reinterpret_cast<float16*>(yref), ws, sm, 4);
    }
    printf("done\n");
    return 0;
}
