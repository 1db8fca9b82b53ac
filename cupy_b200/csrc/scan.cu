// scan.cu -- cumsum / cumprod entry points: the TMA-pipelined scan of b200/scan_pipe.cuh for large arrays
// (every dtype pair of the table, conversions fused), the look-back scan of b200/scan.cuh for small ones.
//
// Replaces cub_device_scan / cub_device_scan_get_workspace_size
// (cupy/cuda/cupy_cub.cu:1163-1185, called from cupy/cuda/cub.pyx:276-306) and,
// because the input dtype is converted on load, the `astype(order='C')` pre-pass
// of cupy/_core/_routines_math.pyx:726-727.  No `int num_items` limit: n is 64-bit.
#include <algorithm>
#include <cstdlib>

#include "common.h"
#include "include/b200/scan.cuh"
#include "include/b200/scan_pipe.cuh"
#include "tma_host.h"
#include "scan_table.h"

namespace b200 {

template <class In, class Acc, class Out, class Op>
__global__ void __launch_bounds__(kScanThreads) scan_kernel(const In* x, Out* y, int64_t n, ScanWorkspace<Acc> ws) {
    scan_body<In, Acc, Out, Op>(x, y, n, ws);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- the pipelined path (b200/scan_pipe.cuh): every dtype pair of the table, large n ----------------
constexpr int64_t kScanPipeMinN = int64_t(1) << 20;

// per dtype pair: items per thread, input / output ring depths, LAG (B200 sweep: profiles/r02_scan_lab.log).
// A thread's span is 64 bytes when input and result have one size, 128 bytes of result otherwise, i.e. 32 / 64 KB
// stages; LAG 3 is where the cross-block exchange stops costing (int64: LAG 1 / 2 / 3 = 4900 / 5490 / 6123 GB/s).
template <class In, class Acc, class Out> struct pipe_cfg {
    static constexpr int SIN = int(sizeof(In)), SOUT = int(sizeof(Out));
    static constexpr int IPT = SIN == SOUT ? 64 / SIN : 128 / SOUT;            // 8 .. 64 | 16 (-> 8 B), 32 (-> 4 B)
    static constexpr int IN_STAGE = kPipeThreads * IPT * SIN, OUT_STAGE = kPipeThreads * IPT * SOUT;
    static constexpr int SO = (OUT_STAGE >= 65536 && IN_STAGE <= 8192) ? 3 : 2;
    static constexpr int SI_FIT = (232448 - 1024 - 64 - SO * OUT_STAGE) / IN_STAGE;
    static constexpr int SI = SI_FIT > 4 ? 4 : SI_FIT;
    typedef ScanPipeCfg<In, Acc, Out, IPT, SI, SO, 3> type;
};

template <class Cfg, class Op>
__global__ void __launch_bounds__(kPipeThreads, 1) scan_pipe_kernel(
        const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out,
        const typename Cfg::in_t* x, typename Cfg::out_t* y, int64_t n_main, int64_t n,
        typename PipeSlot<sizeof(typename Cfg::acc_t)>::storage_t* slots) {
    scan_pipe_body<Cfg, Op>(&tm_in, &tm_out, x, y, n_main, n, slots);
}

static size_t scan_pipe_ws_bytes() { return size_t(kPipeRing) * kPipeThreads * 16; }    // slot ring, G <= 512 blocks

// returns B200_E_UNSUPPORTED when the call does not qualify (the caller uses the look-back kernel)
template <class In, class Acc, class Out, class Op>
static int run_pipe(const In* x, Out* y, int64_t n, void* wsp, size_t ws_bytes, cudaStream_t stream) {
    typedef typename pipe_cfg<In, Acc, Out>::type Cfg;
    typedef typename PipeSlot<sizeof(Acc)>::storage_t slot_t;
    int64_t n_main = n / Cfg::GRANULE * Cfg::GRANULE;
    const int64_t rows_in = n_main * int64_t(sizeof(In)) / 128, rows_out = n_main * int64_t(sizeof(Out)) / 128;
    if (n < kScanPipeMinN || rows_out >= (int64_t(1) << 31) || ws_bytes < scan_pipe_ws_bytes()) return B200_E_UNSUPPORTED;
    DeviceInfo di;
    int st = device_info(&di);
    if (st) return st;
    auto kern = scan_pipe_kernel<Cfg, Op>;
    static int occ = 0;      // per instantiation; benign race
    if (!occ) {
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        int o = 0;
        B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, kPipeThreads, Cfg::SMEM));
        occ = o > 0 ? o : 1;
    }
    CUtensorMap tin, tout;
    const uint64_t din[2] = {uint64_t(128 / sizeof(In)), uint64_t(rows_in)};
    const uint64_t dout[2] = {uint64_t(128 / sizeof(Out)), uint64_t(rows_out)};
    const uint64_t strides[1] = {128};
    const uint32_t bin[2] = {uint32_t(128 / sizeof(In)), uint32_t(Cfg::IN_BOX)};
    const uint32_t bout[2] = {uint32_t(128 / sizeof(Out)), uint32_t(Cfg::OUT_BOX)};
    st = make_tensor_map(&tin, sizeof(In), x, 2, din, strides, bin, CU_TENSOR_MAP_SWIZZLE_128B);
    if (st) return st;
    st = make_tensor_map(&tout, sizeof(Out), y, 2, dout, strides, bout, CU_TENSOR_MAP_SWIZZLE_128B);
    if (st) return st;
    const int64_t tiles = (n_main + Cfg::TILE - 1) / Cfg::TILE;
    // every block gathers a whole wave with one slot per thread: G <= 512
    const unsigned grid = unsigned(std::min<int64_t>(std::min<int64_t>(tiles, int64_t(di.sm_count) * occ), kPipeThreads));
    slot_t* slots = static_cast<slot_t*>(wsp);
    B200_CUDA_TRY(cudaMemsetAsync(slots, 0, size_t(kPipeRing) * grid * sizeof(slot_t), stream));
    void* args[] = {&tin, &tout, &x, &y, &n_main, &n, &slots};
    // cooperative: blocks spin on slots of other blocks, so every block must be resident
    B200_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(kPipeThreads), args,
                                              Cfg::SMEM, stream));
    return 0;
}

static inline int64_t scan_tiles(int64_t n) { return (n + kScanTile - 1) / kScanTile; }

// [counter:16 B][flags: 4*tiles][aggregate: acc*tiles][inclusive: acc*tiles]
static size_t scan_ws_bytes(int64_t n, size_t acc_size) {
    // the look-back kernel only serves n below the pipelined path's threshold (or the A/B knob)
    const size_t tiles = size_t(scan_tiles(n));
    return 16 + align_up(4 * tiles, 16) + 2 * align_up(acc_size * tiles, 16);
}

template <class In, class Acc, class Out>
static int run(int op, const void* x, void* y, int64_t n, void* wsp, size_t ws_bytes, cudaStream_t stream) {
    if (reinterpret_cast<uintptr_t>(x) % 16 || reinterpret_cast<uintptr_t>(y) % 16)
        return fail(B200_E_UNSUPPORTED, "scan needs 16-byte aligned x and y (the host stages misaligned views)");
    static const bool force_lookback = getenv("B200_SCAN_LOOKBACK") != nullptr;     // A/B knob, read once
    const In* xi = static_cast<const In*>(x);
    Out* yo = static_cast<Out*>(y);
    if (!force_lookback) {
        const int st = op == B200_OP_CUMSUM ? run_pipe<In, Acc, Out, ScanSum>(xi, yo, n, wsp, ws_bytes, stream)
                                            : run_pipe<In, Acc, Out, ScanProd>(xi, yo, n, wsp, ws_bytes, stream);
        if (st != B200_E_UNSUPPORTED) return st;
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
    *bytes = std::max(scan_ws_bytes(n, 8), scan_pipe_ws_bytes());   // widest accumulator, either path
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
