// scan_axis.cu -- cumsum / cumprod along ONE axis of a dense array viewed as
// x[outer][n][inner] (C order), scanning n.  No transposes, no `astype` pre-pass: the input
// dtype is converted on load and the result is written in place of the output layout.
//
// Replaces `_proc_as_batch` + `_batch_scan_op` (cupy/_core/_routines_math.pyx:499-699), which
// rolls the axis to the end (a transposing copy each way) and runs a log-step Hillis-Steele
// kernel per doubling stride over the whole array (log2(n) passes over HBM).  Here it is one
// pass: read once, write once.
//
//   inner == 1 (scan along the contiguous axis): LINES kernels.  A line is walked in tiles of
//       GROUP x ITEMS consecutive elements by a GROUP of threads (a block for long lines, a
//       warp for short ones): 16-byte loads, thread-serial scan of its ITEMS, shuffle scan
//       across the warp, shared-memory scan across the warps, running carry in a register.
//   inner > 1 (scan along a strided axis): COLS kernel.  A thread owns V adjacent columns and
//       walks down n with U rows of loads in flight; a warp instruction reads 32*V adjacent
//       elements of one row, so every access is a full line; the running sums stay in registers.
#include <algorithm>
#include <cstdlib>

#include "common.h"
#include "include/b200/scan.cuh"
#include "include/b200/scan_march.cuh"
#include "include/b200/scan_narrow.cuh"
#include "tma_host.h"
#include "scan_table.h"

namespace b200 {

// ------------------------------------------------------------------------------------------
// carry_in (optional): exclusive prefix each line starts from; total (>= 0): the lines are consecutive
// chunks of a FLAT array of `total` elements, the last one ragged.
// W256 (compile time): 32-byte packs (8-byte items) move with one 256-bit access; the host guarantees
// 32-byte aligned lines for them.
template <class In, class Acc, class Out, class Op, int GROUP, int ITEMS, bool W256 = false>
__global__ void __launch_bounds__(256) scan_lines_kernel(const In* __restrict__ x, Out* __restrict__ y, int64_t lines,
                                                         int64_t n_line, int in_vec, int out_vec,
                                                         const Acc* __restrict__ carry_in, int64_t total) {
    constexpr int THREADS = 256;
    constexpr int GROUPS = THREADS / GROUP;
    constexpr int GWARPS = GROUP / 32;
    __shared__ Acc warp_tot[THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = tid / GROUP, gt = tid % GROUP, gw = gt >> 5;
    const int64_t rounds = (lines + int64_t(gridDim.x) * GROUPS - 1) / (int64_t(gridDim.x) * GROUPS);
    for (int64_t r = 0; r < rounds; ++r) {
        const int64_t line = (r * gridDim.x + blockIdx.x) * GROUPS + g;
        const bool live = line < lines;                 // dead groups still reach the barriers
        const In* xl = x + (live ? line : 0) * n_line;
        Out* yl = y + (live ? line : 0) * n_line;
        int64_t n = n_line;
        if (total >= 0 && live) { const int64_t left = total - line * n_line; n = left < n_line ? left : n_line; }
        Acc carry = (carry_in && live) ? carry_in[line] : Op::template identity<Acc>();
        for (int64_t base = 0; base < n_line; base += int64_t(GROUP) * ITEMS) {
            const int64_t i0 = base + int64_t(gt) * ITEMS;
            Acc item[ITEMS];
            if (live && in_vec && i0 + ITEMS <= n) {
                Pack<In, ITEMS> v;
                if constexpr (W256 && sizeof(In) * ITEMS == 32) load_pack256(v, xl + i0);
                else load_pack(v, xl + i0);
#pragma unroll
                for (int j = 0; j < ITEMS; ++j) item[j] = static_cast<Acc>(v[j]);
            } else {
#pragma unroll
                for (int j = 0; j < ITEMS; ++j)
                    item[j] = (live && i0 + j < n) ? static_cast<Acc>(xl[i0 + j]) : Op::template identity<Acc>();
            }
#pragma unroll
            for (int j = 1; j < ITEMS; ++j) item[j] = Op::combine(item[j - 1], item[j]);
            Acc incl = item[ITEMS - 1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                Acc t = shfl_up_any(incl, d);
                if (lane >= d) incl = Op::combine(t, incl);
            }
            Acc excl = shfl_up_any(incl, 1);
            if (lane == 0) excl = Op::template identity<Acc>();
            Acc tile_tot = shfl_any(incl, 31);
            if (GWARPS > 1) {
                if (lane == 31) warp_tot[warp] = incl;
                __syncthreads();
                Acc before = Op::template identity<Acc>();
                tile_tot = Op::template identity<Acc>();
#pragma unroll
                for (int w = 0; w < GWARPS; ++w) {
                    const Acc t = warp_tot[g * GWARPS + w];
                    if (w < gw) before = Op::combine(before, t);
                    tile_tot = Op::combine(tile_tot, t);
                }
                excl = Op::combine(before, excl);
                __syncthreads();
            }
            const Acc pre = Op::combine(carry, excl);
            if (live && out_vec && i0 + ITEMS <= n) {
                Pack<Out, ITEMS> o;
#pragma unroll
                for (int j = 0; j < ITEMS; ++j) o[j] = static_cast<Out>(Op::combine(pre, item[j]));
                if constexpr (W256 && sizeof(Out) * ITEMS == 32) store_pack256(yl + i0, o);
                else store_pack(yl + i0, o);
            } else if (live) {
#pragma unroll
                for (int j = 0; j < ITEMS; ++j)
                    if (i0 + j < n) yl[i0 + j] = static_cast<Out>(Op::combine(pre, item[j]));
            }
            carry = Op::combine(carry, tile_tot);
        }
    }
}

// ------------------------------------------------------------------------------------------
// MODE 0: plain scan.  When outer * inner / V threads cannot fill the GPU, n is cut into S
// segments of `seg` rows that are scanned independently: MODE 1 = segment totals only
// (tot[o][s][inner], no output), then the totals are scanned along s in place (MODE 0 on
// tot), then MODE 2 = scan of every segment starting from the inclusive total of the
// segments before it.  That reads x twice (12 instead of 8 bytes per float32 element) but keeps
// every SM busy; a column-parallel single pass over 4096 columns runs at 10 % of peak.
template <class In, class Acc, class Out, class Op, int V, int U, int MODE>
__global__ void __launch_bounds__(256) scan_cols_kernel(const In* __restrict__ x, Out* __restrict__ y, int64_t outer,
                                                        int64_t n, int64_t inner, int64_t S, int64_t seg,
                                                        Acc* __restrict__ tot) {
    const int64_t chunks = inner / V;
    const int64_t total = outer * S * chunks;
    for (int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; c < total; c += int64_t(gridDim.x) * blockDim.x) {
        const int64_t ic = c % chunks, os = c / chunks;
        const int64_t sg = os % S, o = os / S;
        const int64_t j0 = sg * seg, j1 = (j0 + seg < n) ? j0 + seg : n;
        const In* xp = x + o * n * inner + ic * V;
        Out* yp = y + o * n * inner + ic * V;
        Acc acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = Op::template identity<Acc>();
        if (MODE == 2 && sg > 0) {
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] = tot[((o * S + sg - 1) * inner) + ic * V + k];
        }
        int64_t j = j0;
        for (; j + U <= j1; j += U) {
            Pack<In, V> v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) load_pack(v[u], xp + (j + u) * inner);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                Pack<Out, V> w;
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    acc[k] = Op::combine(acc[k], static_cast<Acc>(v[u][k]));
                    w[k] = static_cast<Out>(acc[k]);
                }
                if (MODE != 1) store_pack(yp + (j + u) * inner, w);
            }
        }
        for (; j < j1; ++j) {
            Pack<In, V> v;
            load_pack(v, xp + j * inner);
            Pack<Out, V> w;
#pragma unroll
            for (int k = 0; k < V; ++k) {
                acc[k] = Op::combine(acc[k], static_cast<Acc>(v[k]));
                w[k] = static_cast<Out>(acc[k]);
            }
            if (MODE != 1) store_pack(yp + j * inner, w);
        }
        if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < V; ++k) tot[(o * S + sg) * inner + ic * V + k] = acc[k];
        }
    }
}

// ---- the column-march kernel (b200/scan_march.cuh): one persistent block per column strip, TMA-pipelined --------
template <class Cfg, class Op>
__global__ void __launch_bounds__(Cfg::THREADS, 1) scan_march_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                                    const __grid_constant__ CUtensorMap tm_out,
                                                                    int64_t outer, int64_t n, int64_t inner) {
    scan_march_body<Cfg, Op>(&tm_in, &tm_out, outer, n, inner);
}

template <class In, class Acc, class Out, class Op, int W_BYTES>
static int launch_march(const In* x, Out* y, int64_t outer, int64_t n, int64_t inner, int sm_count, cudaStream_t s) {
    typedef ScanMarchCfg<In, Acc, Out, W_BYTES, 4, 2> Cfg;
    auto kern = scan_march_kernel<Cfg, Op>;
    static bool attr_set = false;       // per instantiation; benign race
    if (!attr_set) {
        B200_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_set = true;
    }
    CUtensorMap tin, tout;
    const uint64_t dims[3] = {uint64_t(inner), uint64_t(n), uint64_t(outer)};
    const uint64_t sin[2] = {uint64_t(inner) * sizeof(In), uint64_t(n) * uint64_t(inner) * sizeof(In)};
    const uint64_t sout[2] = {uint64_t(inner) * sizeof(Out), uint64_t(n) * uint64_t(inner) * sizeof(Out)};
    const uint32_t box[3] = {uint32_t(Cfg::W), uint32_t(Cfg::R), 1u};
    int st = make_tensor_map(&tin, sizeof(In), x, 3, dims, sin, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (st) return st;
    st = make_tensor_map(&tout, sizeof(Out), y, 3, dims, sout, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (st) return st;
    const int64_t insts = outer * ((inner + Cfg::W - 1) / Cfg::W);
    const unsigned grid = unsigned(std::min<int64_t>(insts, sm_count));
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(tin, tout, outer, n, inner);
    B200_CUDA_TRY(cudaPeekAtLastError());
    return 0;
}

// ---- tall, narrow matrices (b200/scan_narrow.cuh): flat-stream tiles, reduce-then-scan over one segment per block ----
template <class In, class Acc, class Out, class Op>
__global__ void __launch_bounds__(kNarrowScanThreads) scan_narrow_totals_kernel(const In* x, int64_t n, int cols, int active,
                                                                                int64_t seg_tiles, Acc* tot) {
    scan_narrow_totals_body<In, Acc, Out, Op>(x, n, cols, active, seg_tiles, tot);
}
template <class In, class Acc, class Out, class Op>
__global__ void __launch_bounds__(kNarrowScanThreads, 4) scan_narrow_kernel(const In* x, Out* y, int64_t n, int cols, int active,
                                                                         int64_t seg_tiles, const Acc* tot) {
    scan_narrow_body<In, Acc, Out, Op>(x, y, n, cols, active, seg_tiles, tot);
}

template <class In, class Acc, class Out, class Op>
__global__ void __launch_bounds__(kNarrowScanThreads, 4) scan_short_rows_kernel(const In* x, Out* y, int64_t rows, int len, int active) {
    scan_short_rows_body<In, Acc, Out, Op>(x, y, rows, len, active);
}

constexpr int kShortRowsMax = 64;
static bool short_rows_shape(int64_t outer, int64_t n) {
    static const bool off = getenv("B200_SCAN_NO_NARROW") != nullptr;          // A/B knob
    return !off && n >= 2 && n <= kShortRowsMax && outer * n >= 65536;
}

static int narrow_scan_active(int cols, int vec) {
    int a = cols, b = vec;
    while (b) { const int r = a % b; a = b; b = r; }                  // a = gcd(cols, vec)
    const int q = cols / a;
    return kNarrowScanThreads / q * q;                                // active * vec is a multiple of cols
}

template <class In, class Acc, class Out, class Op>
static int launch_short_rows(const In* x, Out* y, int64_t rows, int len, int sm_count, cudaStream_t s) {
    typedef ScanNarrowCfg<In, Acc, Out> Cfg;
    const int active = narrow_scan_active(len, Cfg::VEC);
    const int64_t tile_elems = int64_t(active) * Cfg::VEC * Cfg::U;
    const int64_t tiles = (rows * len + tile_elems - 1) / tile_elems;
    const unsigned grid = unsigned(std::min<int64_t>(tiles, int64_t(sm_count) * 4));
    scan_short_rows_kernel<In, Acc, Out, Op><<<grid, kNarrowScanThreads, 0, s>>>(x, y, rows, len, active);
    B200_CUDA_TRY(cudaPeekAtLastError());
    return 0;
}

static bool narrow_scan_shape(int64_t outer, int64_t n, int64_t inner) {
    static const bool off = getenv("B200_SCAN_NO_NARROW") != nullptr;          // A/B knob
    return !off && outer == 1 && inner >= 2 && inner <= kNarrowScanMaxCols && n * inner >= 65536;
}
constexpr size_t kNarrowScanWorkspace = size_t(296) * 8 * kNarrowScanMaxCols * 8;     // segments x columns x widest accumulator

template <class In, class Acc, class Out, class Op>
static int launch_narrow(const In* x, Out* y, int64_t n, int cols, void* ws, size_t ws_bytes, int sm_count, cudaStream_t s) {
    typedef ScanNarrowCfg<In, Acc, Out> Cfg;
    const int active = narrow_scan_active(cols, Cfg::VEC);
    const int64_t tile_elems = int64_t(active) * Cfg::VEC * Cfg::U;
    const int64_t tiles = (n * cols + tile_elems - 1) / tile_elems;
    static int occ = 0;                                               // per instantiation; benign race
    if (!occ) {
        int o = 0;
        const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, scan_narrow_kernel<In, Acc, Out, Op>, kNarrowScanThreads, 0);
        occ = (e == cudaSuccess && o > 0) ? std::min(o, 8) : 4;
    }
    int64_t S = std::min<int64_t>(tiles, int64_t(sm_count) * occ);
    const int64_t seg_tiles = (tiles + S - 1) / S;
    S = (tiles + seg_tiles - 1) / seg_tiles;
    Acc* tot = static_cast<Acc*>(ws);
    if (S > 1) {
        if (!ws || ws_bytes < size_t(S) * cols * sizeof(Acc))
            return fail(B200_E_WORKSPACE, "scan_axis workspace %zu < %zu", ws_bytes, size_t(S) * cols * sizeof(Acc));
        scan_narrow_totals_kernel<In, Acc, Out, Op><<<unsigned(S), kNarrowScanThreads, 0, s>>>(x, n, cols, active, seg_tiles, tot);
    }
    scan_narrow_kernel<In, Acc, Out, Op><<<unsigned(S), kNarrowScanThreads, 0, s>>>(x, y, n, cols, active, seg_tiles, tot);
    B200_CUDA_TRY(cudaPeekAtLastError());
    return 0;
}

// Strip width for the march: the widest of 512 / 256 / 128 bytes that still gives (nearly) one strip per SM; 0 when
// even 128-byte strips leave more than half of the SMs idle (narrow matrices keep the split scheme).
static int march_width(int64_t outer, int64_t inner_bytes, int sm_count) {
    const int widths[3] = {512, 256, 128};
    for (int w : widths)
        if (outer * ((inner_bytes + w - 1) / w) * 5 >= int64_t(sm_count) * 4) return w;
    return (outer * ((inner_bytes + 127) / 128) * 2 >= sm_count) ? 128 : 0;
}

// every pair of the scan table whose wider side has >= 2 bytes (a 512-byte strip of 1-byte items would exceed the
// 256-element TMA box)
template <class In, class Acc, class Out> struct march_eligible {
    static constexpr bool value = (sizeof(In) > sizeof(Out) ? sizeof(In) : sizeof(Out)) >= 2;
};

// segments to cut n into so that outer * S * inner / V threads fill the GPU (1 = no split)
static int64_t axis_split(int64_t outer, int64_t n, int64_t inner, int sm_count) {
    if (inner <= 1) return 1;
    const int64_t cols = outer * inner;            // upper bound on threads (V = 1)
    const int64_t target = int64_t(sm_count) * 2048;
    if (cols * 4 >= target * 4 / 2 || n < 256) return 1;
    int64_t S = (target + cols - 1) / cols;
    S = std::min<int64_t>(S, n / 64);
    return std::max<int64_t>(S, 1);
}

template <class In, class Out> constexpr int axis_vec() {
    return (16 / int(sizeof(In) > sizeof(Out) ? sizeof(In) : sizeof(Out))) < 1
               ? 1 : (16 / int(sizeof(In) > sizeof(Out) ? sizeof(In) : sizeof(Out)));
}

template <class In, class Acc, class Out, class Op>
static int run_axis(const void* xv, void* yv, int64_t outer, int64_t n, int64_t inner, void* ws, size_t ws_bytes,
                    int sm_count, cudaStream_t s) {
    const In* x = static_cast<const In*>(xv);
    Out* y = static_cast<Out*>(yv);
    const uintptr_t xa = reinterpret_cast<uintptr_t>(xv), ya = reinterpret_cast<uintptr_t>(yv);
    if (inner == 1 && short_rows_shape(outer, n) && xa % 16 == 0 && ya % 16 == 0) {
        return launch_short_rows<In, Acc, Out, Op>(x, y, outer, int(n), sm_count, s);
    } else if (inner == 1) {
        constexpr int ITEMS = (16 / int(sizeof(In))) < 4 ? 4 : (16 / int(sizeof(In)));
        constexpr int in_al = int(sizeof(In)) * ITEMS >= 16 ? 16 : int(sizeof(In)) * ITEMS;
        constexpr int out_al = int(sizeof(Out)) * ITEMS >= 16 ? 16 : int(sizeof(Out)) * ITEMS;
        const int in_vec = (xa % in_al == 0) && ((n * int64_t(sizeof(In))) % in_al == 0);
        const int out_vec = (ya % out_al == 0) && ((n * int64_t(sizeof(Out))) % out_al == 0);
        constexpr bool k256 = sizeof(In) * ITEMS == 32 || sizeof(Out) * ITEMS == 32;
        const bool w256 = k256 && in_vec && out_vec &&
                          (sizeof(In) * ITEMS != 32 || (xa % 32 == 0 && (n * int64_t(sizeof(In))) % 32 == 0)) &&
                          (sizeof(Out) * ITEMS != 32 || (ya % 32 == 0 && (n * int64_t(sizeof(Out))) % 32 == 0));
        if (n > 32 * ITEMS * 2 && w256) {
            const unsigned grid = unsigned(std::min<int64_t>(outer, int64_t(sm_count) * 8));
            scan_lines_kernel<In, Acc, Out, Op, 256, ITEMS, k256><<<grid, 256, 0, s>>>(x, y, outer, n, 1, 1, nullptr, -1);
        } else if (n > 32 * ITEMS * 2) {       // long lines: a block per line
            const unsigned grid = unsigned(std::min<int64_t>(outer, int64_t(sm_count) * 8));
            scan_lines_kernel<In, Acc, Out, Op, 256, ITEMS><<<grid, 256, 0, s>>>(x, y, outer, n, in_vec, out_vec, nullptr, -1);
        } else {                        // short lines: a warp per line
            const unsigned grid = unsigned(std::min<int64_t>((outer + 7) / 8, int64_t(sm_count) * 8));
            scan_lines_kernel<In, Acc, Out, Op, 32, ITEMS><<<grid, 256, 0, s>>>(x, y, outer, n, in_vec, out_vec, nullptr, -1);
        }
    } else if (narrow_scan_shape(outer, n, inner) && xa % 16 == 0 && ya % 16 == 0) {
        return launch_narrow<In, Acc, Out, Op>(x, y, n, int(inner), ws, ws_bytes, sm_count, s);
    } else {
        constexpr int V = axis_vec<In, Out>();
        constexpr int in_al = int(sizeof(In)) * V >= 16 ? 16 : int(sizeof(In)) * V;
        constexpr int out_al = int(sizeof(Out)) * V >= 16 ? 16 : int(sizeof(Out)) * V;
        const bool vec = V > 1 && inner % V == 0 && xa % in_al == 0 && ya % out_al == 0;
        const int64_t S = axis_split(outer, n, inner, sm_count);
        const int64_t seg = (n + S - 1) / S;
        const int64_t cols = outer * S * (vec ? inner / V : inner);
        const unsigned grid = unsigned(std::max<int64_t>(1, std::min<int64_t>((cols + 255) / 256, int64_t(sm_count) * 8)));
        Acc* tot = static_cast<Acc*>(ws);
        if constexpr (march_eligible<In, Acc, Out>::value) {
            static const bool force_split = getenv("B200_SCAN_AXIS_SPLIT") != nullptr;      // A/B knob: the two-read scheme
            constexpr int SW = int(sizeof(In) > sizeof(Out) ? sizeof(In) : sizeof(Out));
            const int64_t inner_bytes = inner * int64_t(SW);        // strip widths are counted on the wider side
            const int wb = (S > 1 && !force_split && (inner * int64_t(sizeof(In))) % 16 == 0 &&
                            (inner * int64_t(sizeof(Out))) % 16 == 0 && xa % 16 == 0 && ya % 16 == 0 &&
                            inner < (int64_t(1) << 31) && n < (int64_t(1) << 31) && outer < (int64_t(1) << 31))
                               ? march_width(outer, inner_bytes, sm_count) : 0;
            if (wb == 512) return launch_march<In, Acc, Out, Op, 512>(x, y, outer, n, inner, sm_count, s);
            if (wb == 256) return launch_march<In, Acc, Out, Op, 256>(x, y, outer, n, inner, sm_count, s);
            if (wb == 128) return launch_march<In, Acc, Out, Op, 128>(x, y, outer, n, inner, sm_count, s);
        }
        if (S > 1) {
            if (ws_bytes < size_t(outer * S * inner) * sizeof(Acc) || !ws)
                return fail(B200_E_WORKSPACE, "scan_axis workspace %zu < %zu", ws_bytes, size_t(outer * S * inner) * sizeof(Acc));
            // tot is Acc-aligned (caller workspace is 256-byte aligned); vector access to it needs V * sizeof(Acc) <= 16
            if (vec) scan_cols_kernel<In, Acc, Out, Op, V, 8, 1><<<grid, 256, 0, s>>>(x, y, outer, n, inner, S, seg, tot);
            else scan_cols_kernel<In, Acc, Out, Op, 1, 8, 1><<<grid, 256, 0, s>>>(x, y, outer, n, inner, S, seg, tot);
            const int64_t tcols = outer * inner;
            const unsigned tgrid = unsigned(std::max<int64_t>(1, std::min<int64_t>((tcols + 255) / 256, int64_t(sm_count) * 8)));
            scan_cols_kernel<Acc, Acc, Acc, Op, 1, 8, 0><<<tgrid, 256, 0, s>>>(tot, tot, outer, S, inner, 1, S, nullptr);
            if (vec) scan_cols_kernel<In, Acc, Out, Op, V, 8, 2><<<grid, 256, 0, s>>>(x, y, outer, n, inner, S, seg, tot);
            else scan_cols_kernel<In, Acc, Out, Op, 1, 8, 2><<<grid, 256, 0, s>>>(x, y, outer, n, inner, S, seg, tot);
        } else {
            if (vec) scan_cols_kernel<In, Acc, Out, Op, V, 8, 0><<<grid, 256, 0, s>>>(x, y, outer, n, inner, 1, n, nullptr);
            else scan_cols_kernel<In, Acc, Out, Op, 1, 8, 0><<<grid, 256, 0, s>>>(x, y, outer, n, inner, 1, n, nullptr);
        }
    }
    B200_CUDA_TRY(cudaPeekAtLastError());
    return 0;
}

}  // namespace b200

using namespace b200;

extern "C" __attribute__((visibility("default"))) int b200_scan_axis_workspace_bytes(int64_t outer, int64_t n, int64_t inner,
                                                                                      size_t* bytes) {
    if (!bytes || outer < 0 || n < 0 || inner < 0) return fail(B200_E_INVALID, "bad argument");
    DeviceInfo di;
    int st = device_info(&di);
    if (st) return st;
    const int64_t S = (outer == 0 || n == 0 || inner == 0) ? 1 : axis_split(outer, n, inner, di.sm_count);
    *bytes = S > 1 ? size_t(outer * S * inner) * 8 : 0;      // widest accumulator
    if (outer > 0 && n > 0 && inner > 0 && narrow_scan_shape(outer, n, inner)) *bytes = std::max(*bytes, kNarrowScanWorkspace);
    return 0;
}

extern "C" __attribute__((visibility("default"))) int b200_scan_axis_run(int op, int in_dtype, int out_dtype, const void* x, void* y,
                                                                          int64_t outer, int64_t n, int64_t inner,
                                                                          void* workspace, size_t workspace_bytes, void* stream) {
    if (op != B200_OP_CUMSUM && op != B200_OP_CUMPROD) return fail(B200_E_INVALID, "op code %d is not a scan", op);
    if (outer < 0 || n < 0 || inner < 0) return fail(B200_E_INVALID, "negative extent");
    if (outer == 0 || n == 0 || inner == 0) return 0;
    if (!x || !y) return fail(B200_E_INVALID, "null pointer");
    DeviceInfo di;
    int st = device_info(&di);
    if (st) return st;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define X(I, O, TI, TA, TO)                                                                        \
    if (in_dtype == I && out_dtype == O)                                                           \
        return op == B200_OP_CUMSUM                                                                \
                   ? run_axis<TI, TA, TO, ScanSum>(x, y, outer, n, inner, workspace, workspace_bytes, di.sm_count, s)  \
                   : run_axis<TI, TA, TO, ScanProd>(x, y, outer, n, inner, workspace, workspace_bytes, di.sm_count, s);
    B200_SCAN_TABLE(X)
#undef X
    return fail(B200_E_UNSUPPORTED, "no prebuilt scan for dtypes %d -> %d", in_dtype, out_dtype);
}
