// scan.cu -- cumsum / cumprod entry points on the look-back scan of b200/scan.cuh.
//
// Replaces cub_device_scan / cub_device_scan_get_workspace_size
// (cupy/cuda/cupy_cub.cu:1163-1185, called from cupy/cuda/cub.pyx:276-306) and,
// because the input dtype is converted on load, the `astype(order='C')` pre-pass
// of cupy/_core/_routines_math.pyx:726-727.  No `int num_items` limit: n is 64-bit.
#include "common.h"
#include "include/b200/scan.cuh"

namespace b200 {

template <class In, class Acc, class Out, class Op>
__global__ void __launch_bounds__(kScanThreads) scan_kernel(const In* x, Out* y, int64_t n, ScanWorkspace<Acc> ws) {
    scan_body<In, Acc, Out, Op>(x, y, n, ws);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static inline int64_t scan_tiles(int64_t n) { return (n + kScanTile - 1) / kScanTile; }

// [counter:16 B][flags: 4*tiles][aggregate: acc*tiles][inclusive: acc*tiles]
static size_t scan_ws_bytes(int64_t n, size_t acc_size) {
    const size_t tiles = size_t(scan_tiles(n));
    return 16 + align_up(4 * tiles, 16) + 2 * align_up(acc_size * tiles, 16);
}

template <class In, class Acc, class Out>
static int run(int op, const void* x, void* y, int64_t n, void* wsp, size_t ws_bytes, cudaStream_t stream) {
    if (reinterpret_cast<uintptr_t>(x) % 16 || reinterpret_cast<uintptr_t>(y) % 16)
        return fail(B200_E_UNSUPPORTED, "scan needs 16-byte aligned x and y (the host stages misaligned views)");
    const size_t tiles = size_t(scan_tiles(n));
    const size_t need = scan_ws_bytes(n, sizeof(Acc));
    if (ws_bytes < need) return fail(B200_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, need);
    char* base = static_cast<char*>(wsp);
    ScanWorkspace<Acc> ws;
    ws.counter = reinterpret_cast<uint32_t*>(base);
    ws.flags = reinterpret_cast<uint32_t*>(base + 16);
    ws.aggregate = reinterpret_cast<Acc*>(base + 16 + align_up(4 * tiles, 16));
    ws.inclusive = reinterpret_cast<Acc*>(base + 16 + align_up(4 * tiles, 16) + align_up(sizeof(Acc) * tiles, 16));
    B200_CUDA_TRY(cudaMemsetAsync(base, 0, 16 + align_up(4 * tiles, 16), stream));
    const In* xi = static_cast<const In*>(x);
    Out* yo = static_cast<Out*>(y);
    if (op == B200_OP_CUMSUM)
        scan_kernel<In, Acc, Out, ScanSum><<<unsigned(tiles), kScanThreads, 0, stream>>>(xi, yo, n, ws);
    else
        scan_kernel<In, Acc, Out, ScanProd><<<unsigned(tiles), kScanThreads, 0, stream>>>(xi, yo, n, ws);
    B200_CUDA_TRY(cudaPeekAtLastError());
    return 0;
}

// (in dtype, out dtype) pairs with a prebuilt kernel.  Result dtype rules:
// cupy/_core/_routines_math.pyx:704-714 (bool/int -> int64, uint -> uint64, else same).
#define B200_SCAN_TABLE(X)                                              \
    X(B200_TYPE_INT64, B200_TYPE_INT64, long long, long long, long long) \
    X(B200_TYPE_INT32, B200_TYPE_INT64, int32_t, long long, long long)   \
    X(B200_TYPE_INT32, B200_TYPE_INT32, int32_t, int32_t, int32_t)       \
    X(B200_TYPE_INT16, B200_TYPE_INT64, int16_t, long long, long long)   \
    X(B200_TYPE_INT8, B200_TYPE_INT64, int8_t, long long, long long)     \
    X(B200_TYPE_INT8, B200_TYPE_INT8, int8_t, int32_t, int8_t)            \
    X(B200_TYPE_BOOL, B200_TYPE_INT64, bool, long long, long long)       \
    X(B200_TYPE_UINT8, B200_TYPE_UINT64, uint8_t, unsigned long long, unsigned long long)   \
    X(B200_TYPE_UINT16, B200_TYPE_UINT64, uint16_t, unsigned long long, unsigned long long) \
    X(B200_TYPE_UINT32, B200_TYPE_UINT64, uint32_t, unsigned long long, unsigned long long) \
    X(B200_TYPE_UINT64, B200_TYPE_UINT64, unsigned long long, unsigned long long, unsigned long long) \
    X(B200_TYPE_FLOAT32, B200_TYPE_FLOAT32, float, float, float)         \
    X(B200_TYPE_FLOAT64, B200_TYPE_FLOAT64, double, double, double)      \
    X(B200_TYPE_FLOAT16, B200_TYPE_FLOAT16, float16, float, float16)     \
    X(B200_TYPE_FLOAT16, B200_TYPE_FLOAT32, float16, float, float)       \
    X(B200_TYPE_FLOAT32, B200_TYPE_FLOAT64, float, double, double)

static size_t acc_size_for(int in_dtype, int out_dtype) {
#define X(I, O, TI, TA, TO) if (in_dtype == I && out_dtype == O) return sizeof(TA);
    B200_SCAN_TABLE(X)
#undef X
    return 0;
}

}  // namespace b200

using namespace b200;

extern "C" __attribute__((visibility("default"))) int b200_scan_supported(int op, int in_dtype, int out_dtype) {
    if (op != B200_OP_CUMSUM && op != B200_OP_CUMPROD) return 0;
    return acc_size_for(in_dtype, out_dtype) ? 1 : 0;
}

extern "C" __attribute__((visibility("default"))) int b200_scan_workspace_bytes(int64_t n, int out_dtype, size_t* bytes) {
    if (!bytes || n < 0) return fail(B200_E_INVALID, "bad argument");
    (void)out_dtype;
    *bytes = scan_ws_bytes(n, 8);   // widest accumulator
    return 0;
}

extern "C" __attribute__((visibility("default"))) int b200_scan_run(int op, int in_dtype, int out_dtype, const void* x, void* y, int64_t n,
                             void* workspace, size_t workspace_bytes, void* stream) {
    if (op != B200_OP_CUMSUM && op != B200_OP_CUMPROD) return fail(B200_E_INVALID, "op code %d is not a scan", op);
    if (n < 0) return fail(B200_E_INVALID, "negative size");
    if (n == 0) return 0;
    if (!x || !y || !workspace) return fail(B200_E_INVALID, "null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define X(I, O, TI, TA, TO) \
    if (in_dtype == I && out_dtype == O) return run<TI, TA, TO>(op, x, y, n, workspace, workspace_bytes, s);
    B200_SCAN_TABLE(X)
#undef X
    return fail(B200_E_UNSUPPORTED, "no prebuilt scan for dtypes %d -> %d", in_dtype, out_dtype);
}
