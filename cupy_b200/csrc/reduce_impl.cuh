// reduce_impl.cuh -- prebuilt reductions (sum/prod/min/max/argmin/argmax/mean/var) on
// the skeleton of b200/reduce.cuh, their launch geometry and dtype dispatch.
//
// Entry points replace cupy/cuda/cupy_cub.cu:1083-1159 (cub_device_reduce,
// cub_device_segmented_reduce + workspace queries) and the generic launch of
// cupy/_core/_reduction.pyx:239-253, 481-508; op / dtype codes are the
// reference's (cupy_cub.h:4-11, type_dispatcher.cuh:15-28).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#pragma once
#include "common.h"
#include "include/b200/reduce.cuh"
#include "include/b200/reduce_ops.cuh"

namespace b200 {

constexpr int kRedThreads = 256;
// Workspace header: atomic tickets of the single-pass combines.  Zero when the
// workspace is first handed in; every kernel leaves it zeroed.
constexpr size_t kTicketBytes = 16384;

template <class Op, int VEC, int UNROLL>
__global__ void __launch_bounds__(kRedThreads) reduce_full_kernel(
        Op op, const typename Op::in_t* x, typename Op::out_t* y, int64_t n,
        typename Op::acc_t* partials, uint32_t* ticket) {
    reduce_full_body<Op, VEC, UNROLL, kRedThreads>(op, x, y, n, partials, ticket);
}

// the sharded variant (cross-GPU combine in the last block) is a separate instantiation: its extra
// registers and shared memory would cost the plain kernel a resident block per SM (sum 2^28: -7 %)
template <class Op, int VEC, int UNROLL>
__global__ void __launch_bounds__(kRedThreads) reduce_full_sharded_kernel(
        Op op, const typename Op::in_t* x, typename Op::out_t* y, int64_t n,
        typename Op::acc_t* partials, uint32_t* ticket, const __grid_constant__ PeerEx ex) {
    reduce_full_body<Op, VEC, UNROLL, kRedThreads>(op, x, y, n, partials, ticket, &ex);
}

template <class Op, int VEC, int UNROLL, int GROUP>
__global__ void __launch_bounds__(kRedThreads) reduce_rows_kernel(
        Op op, const typename Op::in_t* x, typename Op::out_t* y, int64_t rows, int64_t n) {
    reduce_rows_body<Op, VEC, UNROLL, kRedThreads, GROUP>(op, x, y, rows, n);
}

template <class Op, int VEC, int RU, int WC>
__global__ void __launch_bounds__(kRedThreads) reduce_cols_kernel(
        Op op, const typename Op::in_t* x, typename Op::out_t* y, int64_t n, int64_t cols,
        typename Op::acc_t* partials, uint32_t* tickets) {
    reduce_cols_body<Op, VEC, RU, WC>(op, x, y, n, cols, partials, tickets);
}

template <class Op, int VEC, int UNROLL>
__global__ void __launch_bounds__(kRedThreads) reduce_narrow_kernel(
        Op op, const typename Op::in_t* x, typename Op::out_t* y, int64_t n, int cols, int active,
        typename Op::acc_t* partials, uint32_t* ticket) {
    reduce_narrow_body<Op, VEC, UNROLL, kRedThreads>(op, x, y, n, cols, active, partials, ticket);
}

template <class Op, int VEC, int U>
__global__ void __launch_bounds__(kRedThreads) reduce_short_rows_kernel(
        Op op, const typename Op::in_t* x, typename Op::out_t* y, int64_t rows, int len, int active) {
    reduce_short_rows_body<Op, VEC, U, kRedThreads>(op, x, y, rows, len, active);
}

// ---- geometry ---------------------------------------------------------------
struct Geometry {
    int vec;                 // chosen vector width
    unsigned gx, gy, gz;
    int group;               // ROWS
    size_t partial_count;    // accumulators in the workspace
    size_t ticket_count;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <int FULLVEC>
static int pick_vec(const void* x, int64_t inner, int itemsize) {
    int vec = FULLVEC;
    for (; vec > 1; vec >>= 1)
        if (reinterpret_cast<uintptr_t>(x) % (uintptr_t(vec) * itemsize) == 0 && inner % vec == 0) break;
    return vec;
}

// FULL reductions are persistent: exactly as many blocks as the device holds at once (occupancy of the very
// instantiation x SMs), each striding over an equal share -- one wave, no tail (a fixed 8 blocks per SM left
// a partially filled second wave for every functor whose registers allow only 5 or 6 resident blocks)
static int full_grid(int64_t n, int vec, int unroll, int sm, int blocks_per_sm) {
    const int64_t tile = int64_t(kRedThreads) * vec * unroll;
    const int64_t tiles = (n + tile - 1) / tile;
    return int(std::max<int64_t>(1, std::min<int64_t>(tiles, int64_t(sm) * blocks_per_sm)));
}

template <class Op, int VEC, int UNROLL, bool SHARDED>
static int full_blocks_per_sm() {
    static int occ = 0;      // per instantiation; benign race
    if (!occ) {
        int o = 0;
        cudaError_t e;
        if constexpr (SHARDED) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, reduce_full_sharded_kernel<Op, VEC, UNROLL>, kRedThreads, 0);
        else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, reduce_full_kernel<Op, VEC, UNROLL>, kRedThreads, 0);
        occ = (e == cudaSuccess && o > 0) ? std::min(o, 8) : 8;
        static const int forced = [] { const char* v = getenv("B200_FULL_BLOCKS_PER_SM"); return v ? atoi(v) : 0; }();   // A/B knob
        if (forced > 0 && forced < occ) occ = forced;
    }
    return occ;
}

// threads per row: the largest group whose unrolled tile (group * vec * unroll) still fits the row,
// so that the row is read with full vector batches and not through the scalar tail loop
static int rows_group(int64_t n, int vec, int unroll, bool heavy = false, int64_t rows = 0, int sm = 148,
                      bool lane_state = false) {
    static const int forced = [] { const char* e = getenv("B200_ROWS_GROUP"); return e ? atoi(e) : 0; }();   // A/B knob
    if (forced == 256 || forced == 32 || forced == 8 || forced == 1) return forced;
    // arg-reductions and moments on rows under 512 bytes: a thread per row.  Eight lanes per row leave each lane one
    // or two vectors and then pay the lane merges (index / NaN bookkeeping, Chan) through shuffles for every row:
    // rows of 32 float32 195 -> 65 us, of 64 110 -> 65 us, of 100 139 -> 88 us (profiles/r02_short_rows_probe.log)
    if (lane_state && n < int64_t(32) * vec) return 1;
    // functors with a costly per-thread epilogue (arg-reductions: V lane merges with index and NaN handling;
    // moments: V Chan merges) amortise it over 8x more elements with a warp per row, when there are enough rows
    // to keep every warp of the device on its own row
    if (heavy && n >= 2048 && rows >= int64_t(sm) * 64) return 32;
    if (n >= 2048) return kRedThreads;
    if (n >= int64_t(32) * vec * unroll) return 32;     // unrolled batches of a warp
    if (n >= int64_t(8) * vec) return 8;                // unrolled batches or single-vector steps of 8 lanes
    return 1;                                           // a thread per row
}

// wide matrices and plain functors: the 8 warps of a block stand side by side across the columns
// (reduce_cols_body<WC = 8>, 4 KB of every row per visit; WC = 4 measured no gain); functors with
// their own lane state keep one strip per block
static int cols_wc(int64_t cols, int vec, bool heavy) { return (!heavy && cols >= int64_t(8) * 32 * vec * 2) ? 8 : 1; }

static void cols_geometry(int64_t batch, int64_t n, int64_t cols, int vec, int sm, bool heavy, Geometry* g) {
    const int64_t block_cols = int64_t(32) * vec * cols_wc(cols, vec, heavy);
    const int64_t tiles = (cols + block_cols - 1) / block_cols;
    // enough blocks for ~8 per SM, at least 64 rows per split
    int64_t want = (int64_t(sm) * 8 + tiles * batch - 1) / (tiles * batch);
    static const int min_rows = [] { const char* e = getenv("B200_COLS_MIN_ROWS"); return e && atoi(e) > 0 ? atoi(e) : 64; }();   // A/B knob
    int64_t nsplit = std::max<int64_t>(1, std::min<int64_t>(want, n / min_rows));
    // one strip per block (8 split lanes in the fold): the fold by the tile's last block costs ~ nsplit / 8 batched
    // L2 round trips, the main loop ~ n / (nsplit * 8 * unroll); past nsplit ~ sqrt(n) the fold is the longer of the
    // two (tall, narrow matrices: (65536, 256) var went 48 -> see profiles/r02_cols_probe.log)
    static const bool sqrt_cap = getenv("B200_COLS_NO_SQRT_CAP") == nullptr;      // A/B knob
    if (sqrt_cap && cols_wc(cols, vec, heavy) == 1 && nsplit > 1) {
        const int64_t cap = std::max<int64_t>(int64_t(std::sqrt(double(n))), (2 * int64_t(sm) + tiles * batch - 1) / (tiles * batch));
        nsplit = std::max<int64_t>(1, std::min<int64_t>(nsplit, cap));
    }
    nsplit = std::min<int64_t>(nsplit, 65535);
    g->gx = unsigned(tiles);
    g->gy = unsigned(nsplit);
    g->gz = unsigned(batch);
    g->partial_count = nsplit > 1 ? size_t(batch) * nsplit * cols : 0;
    g->ticket_count = nsplit > 1 ? size_t(batch) * tiles : 0;
}

// short ROWS (reduce_short_rows_body): rows of 2..64 elements, enough of them to fill the device ...
constexpr int kShortRowsMax = 64;
// ... and where the group kernels are at their worst (profiles/r02_short_rows_probe.log): rows of exactly one
// 16-byte vector (one small load in flight per thread: 3.0 -> 3.8 TB/s, argmax 1.75 -> 3.4), and -- for functors
// without lane state -- rows of 96..128 bytes, which fall between a thread per row and eight lanes per row
static bool short_rows_shape(const b200_reduce_desc_t* d, int itemsize, bool lane_state) {
    static const bool off = getenv("B200_ROWS_NO_SHORT") != nullptr;           // A/B knob
    static const bool all = getenv("B200_ROWS_ALL_SHORT") != nullptr;          // A/B knob: every row length up to the cap
    if (off || d->n_reduce < 2 || d->n_reduce > kShortRowsMax || d->n_out * d->n_reduce < 65536) return false;
    const int64_t row_bytes = d->n_reduce * itemsize;
    return all || row_bytes == 16 || (!lane_state && row_bytes >= 96 && row_bytes <= 128);
}

// narrow COLS (reduce_narrow_body): rows of at most 64 elements, one batch, enough rows to be worth a
// persistent grid.  Threads that take part per block: the largest count <= 256 whose vectors tile whole rows.
constexpr int kNarrowMaxCols = 64;       // past it the strip kernel has most lanes busy and fewer partials to fold
static bool narrow_shape(const b200_reduce_desc_t* d) {
    static const bool off = getenv("B200_COLS_NO_NARROW") != nullptr;          // A/B knob
    return !off && d->batch == 1 && d->n_out >= 1 && d->n_out <= kNarrowMaxCols && d->n_reduce * d->n_out >= 32768
        && d->n_reduce < (int64_t(1) << 31);
}
static int narrow_active(int cols, int vec) {
    int a = cols, b = vec;
    while (b) { const int r = a % b; a = b; b = r; }            // a = gcd(cols, vec)
    const int q = cols / a;
    return kRedThreads / q * q;
}
// widest vector (elements) whose lane accumulators fit 32 KiB of shared memory
template <class acc_t> constexpr int narrow_vec(int fullvec) {
    return (fullvec > 1 && size_t(kRedThreads) * fullvec * sizeof(acc_t) > 32768) ? narrow_vec<acc_t>(fullvec / 2) : fullvec;
}
template <class Op, int VEC, int UNROLL>
static int narrow_blocks_per_sm() {
    static int occ = 0;      // per instantiation; benign race
    if (!occ) {
        int o = 0;
        const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, reduce_narrow_kernel<Op, VEC, UNROLL>, kRedThreads, 0);
        occ = (e == cudaSuccess && o > 0) ? std::min(o, 8) : 4;
    }
    return occ;
}

// which functors may be folded across GPUs by value: everything except the arg-reductions, whose
// accumulators carry shard-local indices
template <class Op> struct peer_exchangeable { static constexpr bool value = true; };
template <class T, class I, bool kMax> struct peer_exchangeable<ArgOp<T, I, kMax>> { static constexpr bool value = false; };

// ---- typed launch -------------------------------------------------------------
template <class Op, int FULLVEC>
static int run_typed(const Op& op, const b200_reduce_desc_t* d, const void* xv, void* yv,
                     void* ws, size_t ws_bytes, cudaStream_t stream, bool query, size_t* need,
                     const b200_peer_exchange_t* pex = nullptr) {
    typedef typename Op::in_t in_t;
    typedef typename Op::out_t out_t;
    typedef typename Op::acc_t acc_t;
    DeviceInfo di = {148, 10, 0, 0};
    if (!query) {
        int st = device_info(&di);
        if (st) return st;
    }
    const in_t* x = static_cast<const in_t*>(xv);
    out_t* y = static_cast<out_t*>(yv);
    constexpr int U = 4;    // 16-byte loads in flight per thread (lane-state cost is the functor's business)
    // COLS keeps 8 warps x 32*VEC accumulators in shared memory: cap it at 32 KiB
    constexpr int CV = (8 * 32 * FULLVEC * int(sizeof(acc_t)) > 32768) ? FULLVEC / 2 : FULLVEC;
    constexpr int CU = 4;

    if (pex != nullptr && !query) {
        if (d->layout != B200_RED_FULL) return fail(B200_E_UNSUPPORTED, "the cross-GPU combine is fused into FULL reductions only");
        if (Op::kWideIndex || sizeof(acc_t) > 4 * kExWords || !peer_exchangeable<Op>::value)
            return fail(B200_E_UNSUPPORTED, "this reduction cannot be combined across GPUs in-kernel");
        if (pex->nranks < 1 || pex->nranks > kMaxPeers || pex->rank < 0 || pex->rank >= pex->nranks || pex->tag == 0)
            return fail(B200_E_INVALID, "bad peer exchange descriptor");
    }
    PeerEx ex;
    std::memset(&ex, 0, sizeof(ex));
    if (pex != nullptr && !query && pex->nranks > 1) {
        ex.rank = pex->rank; ex.nranks = pex->nranks; ex.tag = pex->tag; ex.n_total = pex->n_total;
        for (int r = 0; r < pex->nranks; ++r) {
            if (!pex->slots[r]) return fail(B200_E_INVALID, "peer exchange: null slot pointer for rank %d", r);
            ex.slots[r] = static_cast<uint64_t*>(pex->slots[r]);
        }
    }
    if (d->layout == B200_RED_FULL) {
        // workspace = [tickets: kTicketBytes][partials]; sized for the widest grid
        const size_t partial_bytes = size_t(di.sm_count) * 8 * sizeof(acc_t);
        if (query) { *need = kTicketBytes + align_up(size_t(296) * 8 * sizeof(acc_t), 16); return 0; }
        const int vec = pick_vec<FULLVEC>(x, d->n_reduce, sizeof(in_t));
        const bool fullvec = vec == FULLVEC;
        uint32_t* ticket = static_cast<uint32_t*>(ws);
        acc_t* partials = reinterpret_cast<acc_t*>(static_cast<char*>(ws) + kTicketBytes);
        if constexpr (peer_exchangeable<Op>::value && !Op::kWideIndex && sizeof(acc_t) <= 4 * kExWords) {
            if (ex.nranks > 1) {
                // never more resident blocks than the plain instantiation runs with: the sharded one needs fewer
                // registers, but a sixth block per SM costs the moments kernel 11 us at 2^29 (profiles/r02_sharded_probe.log)
                const int bps = fullvec ? std::min(full_blocks_per_sm<Op, FULLVEC, U, true>(), full_blocks_per_sm<Op, FULLVEC, U, false>())
                                        : std::min(full_blocks_per_sm<Op, 1, U, true>(), full_blocks_per_sm<Op, 1, U, false>());
                const int grid = full_grid(d->n_reduce, fullvec ? FULLVEC : 1, U, di.sm_count, bps);
                if (grid > 1 && ws_bytes < kTicketBytes + partial_bytes)
                    return fail(B200_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, kTicketBytes + partial_bytes);
                if (fullvec)
                    reduce_full_sharded_kernel<Op, FULLVEC, U><<<grid, kRedThreads, 0, stream>>>(op, x, y, d->n_reduce, partials, ticket, ex);
                else
                    reduce_full_sharded_kernel<Op, 1, U><<<grid, kRedThreads, 0, stream>>>(op, x, y, d->n_reduce, partials, ticket, ex);
                B200_CUDA_TRY(cudaPeekAtLastError());
                return 0;
            }
        }
        const int bps = fullvec ? full_blocks_per_sm<Op, FULLVEC, U, false>() : full_blocks_per_sm<Op, 1, U, false>();
        const int grid = full_grid(d->n_reduce, fullvec ? FULLVEC : 1, U, di.sm_count, bps);
        if (grid > 1 && ws_bytes < kTicketBytes + partial_bytes)
            return fail(B200_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, kTicketBytes + partial_bytes);
        if (fullvec)
            reduce_full_kernel<Op, FULLVEC, U><<<grid, kRedThreads, 0, stream>>>(op, x, y, d->n_reduce, partials, ticket);
        else
            reduce_full_kernel<Op, 1, U><<<grid, kRedThreads, 0, stream>>>(op, x, y, d->n_reduce, partials, ticket);
    } else if constexpr (Op::kWideIndex) {
        // (value, 64-bit index) pairs are only built for FULL; the host routes
        // rows/cols with >= 2^31 reduced elements elsewhere
        return fail(B200_E_UNSUPPORTED, "reduced extent >= 2^31 is only prebuilt for the FULL layout");
    } else if (d->layout == B200_RED_ROWS) {
        if (query) { *need = 0; return 0; }
        constexpr int SV = FULLVEC > 8 ? 8 : FULLVEC;                   // elements per load
        constexpr int SU = (4096 / (kRedThreads * SV)) < 1 ? 1 : 4096 / (kRedThreads * SV);     // 4096 staged elements per tile
        if (SV > 1 && short_rows_shape(d, int(sizeof(in_t)), fast_lanes<Op>::value) && reinterpret_cast<uintptr_t>(x) % 16 == 0) {
            const int len = int(d->n_reduce);
            const int active = narrow_active(len, SV);
            const int64_t tile_elems = int64_t(active) * SV * SU;
            const int64_t tiles = (d->n_out * len + tile_elems - 1) / tile_elems;
            const unsigned grid = unsigned(std::min<int64_t>(tiles, int64_t(di.sm_count) * 6));
            reduce_short_rows_kernel<Op, SV, SU><<<grid, kRedThreads, 0, stream>>>(op, x, y, d->n_out, len, active);
            B200_CUDA_TRY(cudaPeekAtLastError());
            return 0;
        }
        const int vec = pick_vec<FULLVEC>(x, d->n_reduce, sizeof(in_t));
        const int v = (vec == FULLVEC) ? FULLVEC : 1;
        // (measured at 32768^2: float16 argmax 84 -> 98 %, var 86 -> 96 % of peak with a warp per row; float32 loses 3-5 %)
        const int group = rows_group(d->n_reduce, v, U, fast_lanes<Op>::value && sizeof(in_t) <= 2, d->n_out, di.sm_count,
                                     fast_lanes<Op>::value);
        const int64_t rows_per_block = kRedThreads / group;
        const int64_t blocks = (d->n_out + rows_per_block - 1) / rows_per_block;
        const unsigned grid = unsigned(std::max<int64_t>(1, std::min<int64_t>(blocks, int64_t(di.sm_count) * 64)));
#define B200_ROWS(V, G) reduce_rows_kernel<Op, V, U, G><<<grid, kRedThreads, 0, stream>>>(op, x, y, d->n_out, d->n_reduce)
        if (v == FULLVEC) {
            if (group == kRedThreads) B200_ROWS(FULLVEC, kRedThreads);
            else if (group == 32) B200_ROWS(FULLVEC, 32);
            else if (group == 8) B200_ROWS(FULLVEC, 8);
            else B200_ROWS(FULLVEC, 1);
        } else {
            if (group == kRedThreads) B200_ROWS(1, kRedThreads);
            else if (group == 32) B200_ROWS(1, 32);
            else if (group == 8) B200_ROWS(1, 8);
            else B200_ROWS(1, 1);
        }
#undef B200_ROWS
    } else if (d->layout == B200_RED_COLS) {
        Geometry g = {};
        if (query) {
            // worst case over vector widths (tickets live in the fixed header)
            cols_geometry(d->batch, d->n_reduce, d->n_out, 1, 296, fast_lanes<Op>::value, &g);
            Geometry g2 = {};
            cols_geometry(d->batch, d->n_reduce, d->n_out, CV, 296, fast_lanes<Op>::value, &g2);
            size_t pc = std::max(g.partial_count, g2.partial_count);
            if (narrow_shape(d)) pc = std::max(pc, size_t(296) * 8 * size_t(d->n_out));
            *need = pc ? kTicketBytes + align_up(pc * sizeof(acc_t), 16) : 0;
            return 0;
        }
        constexpr int NV = narrow_vec<acc_t>(FULLVEC);
        if (NV > 1 && narrow_shape(d) && reinterpret_cast<uintptr_t>(x) % (uintptr_t(NV) * sizeof(in_t)) == 0) {
            const int cols = int(d->n_out);
            const int active = narrow_active(cols, NV);
            const int64_t chunks = d->n_reduce / (int64_t(active) * NV / cols);
            // the last block folds grid * cols partials with 256 / cols threads per column: wider rows get fewer
            // blocks (never under four per SM) so that this fold stays a few L2 round trips long
            const int64_t resident = int64_t(di.sm_count) * narrow_blocks_per_sm<Op, NV, U>();
            const int64_t fold_cap = std::max<int64_t>(4 * int64_t(di.sm_count), 32768 / cols);
            const int grid = int(std::max<int64_t>(1, std::min<int64_t>((chunks + U - 1) / U, std::min(resident, fold_cap))));
            const size_t pbytes = size_t(grid) * cols * sizeof(acc_t);
            if (grid > 1 && ws_bytes < kTicketBytes + pbytes)
                return fail(B200_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, kTicketBytes + pbytes);
            if (grid > 1 && (reinterpret_cast<uintptr_t>(ws) & 15))
                return fail(B200_E_INVALID, "workspace must be 16-byte aligned");
            reduce_narrow_kernel<Op, NV, U><<<grid, kRedThreads, 0, stream>>>(
                op, x, y, d->n_reduce, cols, active, reinterpret_cast<acc_t*>(static_cast<char*>(ws) + kTicketBytes),
                static_cast<uint32_t*>(ws));
            B200_CUDA_TRY(cudaPeekAtLastError());
            return 0;
        }
        int vec = pick_vec<CV>(x, d->n_out, sizeof(in_t));
        if (vec != CV) vec = 1;
        cols_geometry(d->batch, d->n_reduce, d->n_out, vec, di.sm_count, fast_lanes<Op>::value, &g);
        if (g.ticket_count * sizeof(uint32_t) > kTicketBytes) {   // cannot happen: splits only when tiles*batch < 8*SMs
            g.gy = 1; g.partial_count = 0; g.ticket_count = 0;
        }
        const size_t pbytes = g.partial_count * sizeof(acc_t);
        if (g.partial_count && ws_bytes < kTicketBytes + pbytes)
            return fail(B200_E_WORKSPACE, "workspace %zu < %zu", ws_bytes, kTicketBytes + pbytes);
        if (g.partial_count && (reinterpret_cast<uintptr_t>(ws) & 15))
            return fail(B200_E_INVALID, "workspace must be 16-byte aligned (vector loads of the split partials)");
        uint32_t* tickets = static_cast<uint32_t*>(ws);
        acc_t* partials = reinterpret_cast<acc_t*>(static_cast<char*>(ws) + kTicketBytes);
        const dim3 grid(g.gx, g.gy, g.gz);
        const int wc = cols_wc(d->n_out, vec, fast_lanes<Op>::value);
#define B200_COLS(V, R, W) reduce_cols_kernel<Op, V, R, W><<<grid, kRedThreads, 0, stream>>>(op, x, y, d->n_reduce, d->n_out, partials, tickets)
        if (vec == CV) { if (wc == 8) B200_COLS(CV, CU, 8); else B200_COLS(CV, CU, 1); }
        else { if (wc == 8) B200_COLS(1, 4, 8); else B200_COLS(1, 4, 1); }
#undef B200_COLS
    } else {
        return fail(B200_E_INVALID, "bad reduction layout %d", d->layout);
    }
    B200_CUDA_TRY(cudaPeekAtLastError());
    return 0;
}

// ---- dtype dispatch -----------------------------------------------------------
// accumulator / result rules: cupy/_core/_routines_math.pyx:762-807 (sum, prod),
// cupy/_core/_routines_statistics.pyx:128-146 (mean), :556-600 (var)
template <class T> struct sum_acc { typedef T type; };
template <> struct sum_acc<bool> { typedef long long type; };
template <> struct sum_acc<int8_t> { typedef long long type; };
template <> struct sum_acc<int16_t> { typedef long long type; };
template <> struct sum_acc<int32_t> { typedef long long type; };
template <> struct sum_acc<uint8_t> { typedef unsigned long long type; };
template <> struct sum_acc<uint16_t> { typedef unsigned long long type; };
template <> struct sum_acc<uint32_t> { typedef unsigned long long type; };
template <> struct sum_acc<float16> { typedef float type; };

template <class T> struct sum_out { typedef typename sum_acc<T>::type type; };
template <> struct sum_out<float16> { typedef float16 type; };

template <class T> struct mom_float { typedef double type; };   // ints: float64
template <> struct mom_float<float> { typedef float type; };
template <> struct mom_float<float16> { typedef float type; };
template <class T> struct mom_out { typedef double type; };
template <> struct mom_out<float> { typedef float type; };
template <> struct mom_out<float16> { typedef float16 type; };

template <class T> struct out_id;
template <> struct out_id<long long> { static constexpr int v = B200_TYPE_INT64; };
template <> struct out_id<unsigned long long> { static constexpr int v = B200_TYPE_UINT64; };
template <> struct out_id<float16> { static constexpr int v = B200_TYPE_FLOAT16; };
template <> struct out_id<float> { static constexpr int v = B200_TYPE_FLOAT32; };
template <> struct out_id<double> { static constexpr int v = B200_TYPE_FLOAT64; };
template <> struct out_id<int32_t> { static constexpr int v = B200_TYPE_INT32; };
template <> struct out_id<int8_t> { static constexpr int v = B200_TYPE_INT8; };
template <> struct out_id<uint8_t> { static constexpr int v = B200_TYPE_UINT8; };
template <> struct out_id<int16_t> { static constexpr int v = B200_TYPE_INT16; };
template <> struct out_id<uint16_t> { static constexpr int v = B200_TYPE_UINT16; };
template <> struct out_id<uint32_t> { static constexpr int v = B200_TYPE_UINT32; };
template <> struct out_id<bool> { static constexpr int v = B200_TYPE_BOOL; };

#define B200_REQUIRE_OUT(T)                                                                         \
    if (d->out_dtype != out_id<T>::v)                                                               \
        return fail(B200_E_UNSUPPORTED, "op %d in dtype %d: prebuilt result dtype is %d, asked %d", \
                    d->op, d->in_dtype, out_id<T>::v, d->out_dtype)

template <class T>
static int run_for_type(const b200_reduce_desc_t* d, const void* x, void* y, void* ws, size_t wsb,
                        cudaStream_t s, bool query, size_t* need, const b200_peer_exchange_t* pex) {
    constexpr int FV = (16 / int(sizeof(T))) > 8 ? 8 : (16 / int(sizeof(T)));
    // index type of arg-reductions: 32 bit whenever the reduced extent allows
    const bool j32 = d->n_reduce < (int64_t(1) << 31);
    switch (d->op) {
        case B200_OP_SUM: {
            typedef typename sum_acc<T>::type A; typedef typename sum_out<T>::type O;
            B200_REQUIRE_OUT(O);
            return run_typed<SumOp<T, A, O>, FV>(SumOp<T, A, O>(), d, x, y, ws, wsb, s, query, need, pex);
        }
        case B200_OP_PROD: {
            typedef typename sum_acc<T>::type A; typedef typename sum_out<T>::type O;
            B200_REQUIRE_OUT(O);
            return run_typed<ProdOp<T, A, O>, FV>(ProdOp<T, A, O>(), d, x, y, ws, wsb, s, query, need, pex);
        }
        case B200_OP_MIN:
            B200_REQUIRE_OUT(T);
            return run_typed<MinMaxOp<T, false>, FV>(MinMaxOp<T, false>(), d, x, y, ws, wsb, s, query, need, pex);
        case B200_OP_MAX:
            B200_REQUIRE_OUT(T);
            return run_typed<MinMaxOp<T, true>, FV>(MinMaxOp<T, true>(), d, x, y, ws, wsb, s, query, need, pex);
        case B200_OP_ARGMIN:
            B200_REQUIRE_OUT(long long);
            if (j32) return run_typed<ArgOp<T, int, false>, FV>(ArgOp<T, int, false>(), d, x, y, ws, wsb, s, query, need, pex);
            return run_typed<ArgOp<T, long long, false>, FV>(ArgOp<T, long long, false>(), d, x, y, ws, wsb, s, query, need, pex);
        case B200_OP_ARGMAX:
            B200_REQUIRE_OUT(long long);
            if (j32) return run_typed<ArgOp<T, int, true>, FV>(ArgOp<T, int, true>(), d, x, y, ws, wsb, s, query, need, pex);
            return run_typed<ArgOp<T, long long, true>, FV>(ArgOp<T, long long, true>(), d, x, y, ws, wsb, s, query, need, pex);
        case B200_OP_MEAN: {
            typedef typename mom_float<T>::type F; typedef typename mom_out<T>::type O;
            B200_REQUIRE_OUT(O);
            return run_typed<MeanOp<T, F, O>, FV>(MeanOp<T, F, O>(), d, x, y, ws, wsb, s, query, need, pex);
        }
        case B200_OP_VAR: {
            typedef typename mom_float<T>::type F; typedef typename mom_out<T>::type O;
            B200_REQUIRE_OUT(O);
            MomentsOp<T, F, O, kMomVar> op;
            op.ddof = F(d->param);
            return run_typed<MomentsOp<T, F, O, kMomVar>, FV>(op, d, x, y, ws, wsb, s, query, need, pex);
        }
        case B200_OP_MOMENTS: {
            // y holds one (n, mean, M2) triple of doubles
            typedef typename mom_float<T>::type F;
            B200_REQUIRE_OUT(double);
            if (d->layout != B200_RED_FULL) return fail(B200_E_UNSUPPORTED, "B200_OP_MOMENTS is a full reduction");
            MomentsOp<T, F, F, kMomPair> op;
            op.ddof = F(0);
            return run_typed<MomentsOp<T, F, F, kMomPair>, FV>(op, d, x, y, ws, wsb, s, query, need, pex);
        }
        default:
            return fail(B200_E_INVALID, "op code %d is not a reduction", d->op);
    }
}

}  // namespace b200
