// scan.cu -- cumsum / cumprod entry points on the look-back scan of b200/scan.cuh.
//
// Replaces cub_device_scan / cub_device_scan_get_workspace_size
// (cupy/cuda/cupy_cub.cu:1163-1185, called from cupy/cuda/cub.pyx:276-306) and,
// because the input dtype is converted on load, the `astype(order='C')` pre-pass
// of cupy/_core/_routines_math.pyx:726-727.  No `int num_items` limit: n is 64-bit.
#include <algorithm>
#include <cstdlib>

#include "common.h"
#include "include/b200/scan.cuh"
#include "include/b200/scan_tma.cuh"
#include "tma_host.h"
#include "scan_table.h"

namespace b200 {

template <class In, class Acc, class Out, class Op>
__global__ void __launch_bounds__(kScanThreads) scan_kernel(const In* x, Out* y, int64_t n, ScanWorkspace<Acc> ws) {
    scan_body<In, Acc, Out, Op>(x, y, n, ws);
}

template <class T, class Op, int THREADS, int STAGES, int LAG>
__global__ void __launch_bounds__(THREADS) scan_tma_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                           const __grid_constant__ CUtensorMap tm_out,
                                                           const T* x, T* y, int64_t n,
                                                           typename LookbackSlot<sizeof(T)>::storage_t* slots) {
    scan_tma_body<T, Op, THREADS, STAGES, LAG>(&tm_in, &tm_out, x, y, n, slots);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- TMA-pipelined path (same-size 4 / 8 byte types, large n) -----------------------------
constexpr int64_t kScanTmaMinN = int64_t(1) << 20;
constexpr int kScanTmaMinThreads = 128;     // smallest tile = 128 rows of 128 bytes

static size_t scan_tma_ws_bytes(int64_t n) {
    // worst case over element sizes: 8-byte items, 16 per row, 16-byte slots
    const int64_t rows = (n + 15) / 16;
    const int64_t tiles = (rows + kScanTmaMinThreads - 1) / kScanTmaMinThreads;
    return 16 + 2 * size_t(tiles) * 16;     // per-tile aggregates + per-wave inclusive prefixes
}

template <class T, class Op, int THREADS, int STAGES, int LAG>
static int launch_tma(const CUtensorMap& tin, const CUtensorMap& tout, const T* x, T* y, int64_t n, void* wsp,
                      int sm_count, cudaStream_t stream) {
    typedef typename LookbackSlot<sizeof(T)>::storage_t slot_t;
    constexpr int ITEMS = 128 / int(sizeof(T));
    constexpr int smem = ScanTmaSmem<THREADS, STAGES>::kBytes;
    auto kern = scan_tma_kernel<T, Op, THREADS, STAGES, LAG>;
    static int occ = 0;      // per instantiation; benign race
    if (!occ) {
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int o = 0;
        B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, THREADS, smem));
        occ = o > 0 ? o : 1;
    }
    const int64_t rows = (n + ITEMS - 1) / ITEMS;
    const int64_t tiles = (rows + THREADS - 1) / THREADS;
    char* base = static_cast<char*>(wsp);
    B200_CUDA_TRY(cudaMemsetAsync(base, 0, 16 + 2 * size_t(tiles) * sizeof(slot_t), stream));
    // every block gathers a whole wave with <= 4 slots per thread: G <= 4 * THREADS
    const unsigned grid = unsigned(std::min<int64_t>(std::min<int64_t>(tiles, int64_t(sm_count) * occ), 4 * THREADS));
    slot_t* slots = reinterpret_cast<slot_t*>(base + 16);
    CUtensorMap tin_v = tin, tout_v = tout;
    void* args[] = {&tin_v, &tout_v, &x, &y, &n, &slots};
    // cooperative: the look-back spins on tiles of other blocks, so every block must be resident
    B200_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(THREADS), args,
                                              smem, stream));
    return 0;
}

// returns B200_E_UNSUPPORTED when the call does not qualify (caller uses the register path)
template <class T>
static int run_tma(int op, const void* x, void* y, int64_t n, void* wsp, size_t ws_bytes, cudaStream_t stream) {
    constexpr int ITEMS = 128 / int(sizeof(T));
    if (n < kScanTmaMinN || n / ITEMS >= (int64_t(1) << 31)) return B200_E_UNSUPPORTED;
    if (ws_bytes < scan_tma_ws_bytes(n)) return B200_E_UNSUPPORTED;
    static const int cfg_env = [] { const char* e = getenv("B200_SCAN_CFG"); return e ? atoi(e) : 0; }();
    int cfg = cfg_env;
    if (cfg < 0) return B200_E_UNSUPPORTED;
    // default by item size (B200 sweep, profiles/r01_scan_probe_s7.log): 4-byte items run best as two
    // 128-row blocks per SM (f32 2^28: 5683 vs 5364 GB/s), 8-byte items as one 256-row block
    if (cfg == 0) cfg = sizeof(T) == 4 ? 163 : 263;
    DeviceInfo di;
    int st = device_info(&di);
    if (st) return st;
    const int threads = (cfg / 100) ? (cfg / 100) * 128 : 256;
    CUtensorMap tin, tout;
    const uint64_t dims[2] = {uint64_t(ITEMS), uint64_t(n / ITEMS)};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {uint32_t(ITEMS), uint32_t(threads)};
    st = make_tensor_map(&tin, sizeof(T), x, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (st) return st;
    st = make_tensor_map(&tout, sizeof(T), y, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (st) return st;
    const T* xi = static_cast<const T*>(x);
    T* yo = static_cast<T*>(y);
#define B200_TMA_CASE(CODE, TH, ST, LG)                                                                     \
    if (cfg == CODE) {                                                                                      \
        if (op == B200_OP_CUMSUM) return launch_tma<T, ScanSum, TH, ST, LG>(tin, tout, xi, yo, n, wsp, di.sm_count, stream); \
        return launch_tma<T, ScanProd, TH, ST, LG>(tin, tout, xi, yo, n, wsp, di.sm_count, stream);         \
    }
    B200_TMA_CASE(162, 128, 6, 2)
    B200_TMA_CASE(163, 128, 6, 3)
    B200_TMA_CASE(252, 256, 5, 2)
    B200_TMA_CASE(262, 256, 6, 2)
    B200_TMA_CASE(263, 256, 6, 3)
    B200_TMA_CASE(231, 256, 3, 1)      // 96 KB: two blocks (16 warps) per SM
    B200_TMA_CASE(131, 128, 3, 1)      // 48 KB: four blocks per SM
    B200_TMA_CASE(142, 128, 4, 2)      // 64 KB: three blocks per SM
    B200_TMA_CASE(431, 512, 3, 1)      // one block of 16 warps per SM
#undef B200_TMA_CASE
    return B200_E_UNSUPPORTED;
}

template <class In, class Acc, class Out> struct tma_eligible { static constexpr bool value = false; };
template <> struct tma_eligible<long long, long long, long long> { static constexpr bool value = true; };
template <> struct tma_eligible<unsigned long long, unsigned long long, unsigned long long> { static constexpr bool value = true; };
template <> struct tma_eligible<double, double, double> { static constexpr bool value = true; };
template <> struct tma_eligible<float, float, float> { static constexpr bool value = true; };
template <> struct tma_eligible<int32_t, int32_t, int32_t> { static constexpr bool value = true; };

static inline int64_t scan_tiles(int64_t n) { return (n + kScanTile - 1) / kScanTile; }

// [counter:16 B][flags: 4*tiles][aggregate: acc*tiles][inclusive: acc*tiles]
static size_t scan_ws_bytes(int64_t n, size_t acc_size) {
    const size_t tiles = size_t(scan_tiles(n));
    return 16 + align_up(4 * tiles, 16) + 2 * align_up(acc_size * tiles, 16);
}

// scan_axis.cu: flat scan as (line totals, carry scan, line scans) for the dtype pairs the TMA scan skips
template <class In, class Acc, class Out, class Op>
int scan_flat_lines(const In* x, Out* y, int64_t n, void* ws, size_t ws_bytes, int sm_count, cudaStream_t s);
constexpr int64_t kFlatLinesMinN = int64_t(1) << 18;

template <class In, class Acc, class Out>
static int run(int op, const void* x, void* y, int64_t n, void* wsp, size_t ws_bytes, cudaStream_t stream) {
    if (reinterpret_cast<uintptr_t>(x) % 16 || reinterpret_cast<uintptr_t>(y) % 16)
        return fail(B200_E_UNSUPPORTED, "scan needs 16-byte aligned x and y (the host stages misaligned views)");
    if constexpr (tma_eligible<In, Acc, Out>::value) {
        const int st = run_tma<In>(op, x, y, n, wsp, ws_bytes, stream);
        if (st != B200_E_UNSUPPORTED) return st;
    }
    static const bool force_lookback = getenv("B200_SCAN_LOOKBACK") != nullptr;     // read once, not per call
    if (n >= kFlatLinesMinN && !force_lookback) {
        DeviceInfo di;
        int st = device_info(&di);
        if (st) return st;
        const In* xi = static_cast<const In*>(x);
        Out* yo = static_cast<Out*>(y);
        st = op == B200_OP_CUMSUM ? scan_flat_lines<In, Acc, Out, ScanSum>(xi, yo, n, wsp, ws_bytes, di.sm_count, stream)
                                  : scan_flat_lines<In, Acc, Out, ScanProd>(xi, yo, n, wsp, ws_bytes, di.sm_count, stream);
        if (st == 0) { B200_CUDA_TRY(cudaPeekAtLastError()); return 0; }
        if (st != B200_E_WORKSPACE) return st;       // too little workspace: the look-back kernel below
    }
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
    *bytes = std::max(scan_ws_bytes(n, 8), scan_tma_ws_bytes(n));   // widest accumulator, either path
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
