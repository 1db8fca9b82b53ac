// b200/reduce.cuh -- the reduction skeleton: single-pass full, row (segmented)
// and column kernels over a small functor interface.  The same kernels are
// instantiated (a) by nvcc for the built-in sum/prod/min/max/arg*/mean/var
// functors of reduce_ops.cuh and (b) by NVRTC for functors generated from
// ReductionKernel / create_reduction_func code strings.
//
// Functor interface (`Op`):
//     typedef ... in_t;      element type in memory
//     typedef ... acc_t;     the reference's `_type_reduce`
//     typedef ... out_t;     element type of the result in memory
//     typedef ... index_t;   type of the reduce-axis index `_J` (int or long long)
//     typedef ... ctx_t;     per-step context shared by the lanes of a thread
//     acc_t identity() const;
//     ctx_t step(int count) const;              // count = elements each lane state holds after this step
//     void  accumulate(acc_t&, const ctx_t&, const in_t&, index_t j) const;   // fold ONE element (j increases per state)
//     acc_t single(const in_t&, index_t j) const;                             // acc_t of one element (tails)
//     acc_t combine(const acc_t& a, const acc_t& b) const;                    // a precedes b in index order
//     out_t post(const acc_t&, long long n_reduce) const;
//
// Replaces: the shared-memory tree of cupy/_core/_reduction.pyx:59-112, the CUB
// block-reduce template of cupy/_core/_cub_reduction.pyx:67-215 (two-pass for
// full reductions) and the CUB device calls of cupy/cuda/cupy_cub.cu:749-1148.
#pragma once
#include "base.cuh"

namespace b200 {

// The input "pointer" of a skeleton: `const in_t*` unless the functor brings its own type
// (Op::ptr_t).  NVRTC-generated functors of reductions over SEVERAL arrays of one common layout
// use a struct of pointers whose in_t is the tuple of their elements: it only has to offer
// `p + offset`, `p[i]` and a `load_pack(Pack<in_t, N>&, p)` overload.
template <class T> struct void_of { typedef void type; };
template <class Op, class = void> struct in_ptr { typedef const typename Op::in_t* __restrict__ type; };
template <class Op> struct in_ptr<Op, typename void_of<typename Op::ptr_t>::type> { typedef typename Op::ptr_t type; };

// Pointer steps of the ROWS / COLS skeletons, as hooks: a pointer struct whose members are not all
// laid out like x (operands broadcast along the reduced or the kept axes: `x - mean`, weights) brings
// its own overloads.
template <class P> B200_DEVICE P row_ptr(const P& x, int64_t row, int64_t n) { return x + row * n; }
template <class P> B200_DEVICE P cols_ptr(const P& x, int64_t b, int64_t n, int64_t cols, int64_t c0) {
    return x + b * n * cols + c0;
}
template <class P> B200_DEVICE P cols_row(const P& xb, int64_t r, int64_t cols) { return xb + r * cols; }

template <class T>
B200_DEVICE T load_volatile(const T* p) {
    // accumulators published by other blocks: read around L1, in words when possible
    typedef typename RawVec<(sizeof(T) % 4 == 0) ? 4 : 1>::type W;
    constexpr int words = sizeof(T) / sizeof(W);
    union U { T t; W w[words]; B200_DEVICE U() {} };
    U u;
    const volatile W* s = reinterpret_cast<const volatile W*>(p);
#pragma unroll
    for (int i = 0; i < words; ++i) u.w[i] = s[i];
    return u.t;
}

// The same for data whose publication the caller has already ordered (ticket + fences): L2 loads (ld.global.cg,
// around the non-coherent L1) in the widest words the type allows, NOT volatile, so independent loads can be
// batched ahead of their uses.  `p` must be aligned to that word (16 / 8 / 4 bytes).
template <class T>
B200_DEVICE T load_cg(const T* p) {
    constexpr int wbytes = (sizeof(T) % 16 == 0) ? 16 : (sizeof(T) % 8 == 0) ? 8 : (sizeof(T) % 4 == 0) ? 4 : 1;
    typedef typename RawVec<wbytes>::type W;
    constexpr int words = sizeof(T) / sizeof(W);
    union U { T t; W w[words]; B200_DEVICE U() {} };
    U u;
    const W* s = reinterpret_cast<const W*>(p);
#pragma unroll
    for (int i = 0; i < words; ++i) u.w[i] = __ldcg(s + i);
    return u.t;
}

// Sub-warp combine over GROUP consecutive lanes; result valid in the group's lane 0.
template <int GROUP, class Op, class T>
B200_DEVICE T group_combine(const Op& op, T v) {
#pragma unroll
    for (int d = GROUP / 2; d > 0; d >>= 1) {
        T o = shfl_down_any(v, d);
        v = op.combine(v, o);
    }
    return v;
}

// Merge a thread's lane states in a fixed order.
template <class Op, int U, int V>
B200_DEVICE typename Op::acc_t merge_lanes(const Op& op, typename Op::acc_t (&acc)[U][V]) {
    typename Op::acc_t r = acc[0][0];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < V; ++k)
            if (u | k) r = op.combine(r, acc[u][k]);
    return r;
}

// ---------------------------------------------------------------------------
// ThreadAcc: what a thread keeps while it streams over its share of the reduced
// axis.  The generic form holds U*V independent lane states of Op::acc_t and
// folds one element into each per batch.  A functor can replace it with a
// cheaper representation (fast_lanes<Op>: e.g. arg-reductions keep the batch
// start index instead of a per-element index and treat NaN out of line) as long
// as it offers the same four members.
//   fold(v, j0, us, ks)  one full batch; element (u, k) has reduce-axis index j0 + u*us + k*ks
//   fold_one(x, j)       one stray element (tails)
//   result()             everything merged (FULL / ROWS)
//   result_lane(k)       merged over u only: lane k is its own output column (COLS)
// ---------------------------------------------------------------------------
template <class Op> struct fast_lanes { static constexpr bool value = false; };

template <class Op, int U, int V, bool FAST = fast_lanes<Op>::value>
struct ThreadAcc {
    typedef typename Op::acc_t acc_t;
    typedef typename Op::index_t index_t;
    const Op& op;
    acc_t acc[U][V];
    int count;
    B200_DEVICE explicit ThreadAcc(const Op& op_) : op(op_), count(0) {
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < V; ++k) acc[u][k] = op.identity();
    }
    B200_DEVICE void fold(const Pack<typename Op::in_t, V> (&v)[U], index_t j0, index_t us, index_t ks) {
        const typename Op::ctx_t ctx = op.step(++count);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < V; ++k) op.accumulate(acc[u][k], ctx, v[u][k], j0 + u * us + k * ks);
    }
    B200_DEVICE void fold_one(const typename Op::in_t& x, index_t j) {
        acc[0][0] = op.combine(acc[0][0], op.single(x, j));
    }
    B200_DEVICE void fold_one_lane(int k, const typename Op::in_t& x, index_t j) {
        acc[0][k] = op.combine(acc[0][k], op.single(x, j));
    }
    B200_DEVICE acc_t result() { return merge_lanes<Op, U, V>(op, acc); }
    B200_DEVICE acc_t result_lane(int k) {
        acc_t a = acc[0][k];
#pragma unroll
        for (int u = 1; u < U; ++u) a = op.combine(a, acc[u][k]);
        return a;
    }
};

// ---------------------------------------------------------------------------
// FULL: x[n] -> y[0].  Persistent grid; each block folds its tiles into
// U*V lane states per thread, combines through shuffles + shared memory,
// publishes one partial, and the LAST block to arrive (atomic ticket) folds the
// partials in a fixed order -- one launch, one read of x, deterministic result.
// workspace: gridDim.x partials + a zeroed uint32 ticket (left zeroed).
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Cross-GPU combine fused into the last block of a FULL reduction (sharded sum / mean / var of
// BASELINE config 5).  Every rank owns a small exchange buffer that all ranks of the node have mapped
// (symmetric memory over NVLink / NVSwitch peer access).  The last block of rank r
//   1. stores its partial accumulator, as 64-bit words {tag, 32 payload bits}, into slot [parity][r] of
//      EVERY rank's buffer (plain st.global on the mapped peer pointers -- NVLink stores),
//   2. spins on its OWN buffer until the words of all ranks carry this call's tag,
//   3. folds the partials in rank order (bit-identical result on every rank) and applies post().
// A word is written by one atomic 8-byte store, so a matching tag implies the payload: no fences, no
// second launch, no NCCL small-message latency (replaces the ncclAllReduce(count=1) of
// cupyx/distributed/_nccl_comm.py:312-320 -> cupy_backends/cuda/libs/nccl.pyx:470-477 for this path).
// Slots are double-buffered by tag parity: a rank can only start call e+2 after every rank finished
// call e (it needed their words of call e+1), so call e's words are never overwritten while in use.
// ---------------------------------------------------------------------------
constexpr int kMaxPeers = 16;        // == B200_MAX_PEERS
constexpr int kExWords = 8;          // accumulators up to 32 bytes
struct PeerEx {
    int32_t   rank, nranks;          // nranks <= 1: no exchange
    uint32_t  tag, pad_;             // same on every rank, different from the previous call's, never 0
    long long n_total;               // elements over all ranks (post() divides by it for mean / var)
    uint64_t* slots[kMaxPeers];      // slots[r] = rank r's buffer: [2][kMaxPeers][kExWords] words, zeroed once
};

// __noinline__: run once, by the last block; keeping it out of line leaves the register allocation and load
// batching of the streaming loop exactly as in the plain kernel
template <class Op, int THREADS>
__device__ __noinline__ typename Op::acc_t peer_allreduce(const Op& op, const typename Op::acc_t& mine,
                                                          const PeerEx& ex) {
    typedef typename Op::acc_t acc_t;
    constexpr int W = (int(sizeof(acc_t)) + 3) / 4;
    static_assert(W <= kExWords, "accumulator too large for the peer exchange");
    static_assert(THREADS >= kMaxPeers * kExWords, "one thread per (rank, word)");
    __shared__ uint32_t ex_mine[kExWords];
    __shared__ uint32_t ex_words[kMaxPeers][kExWords];
    if (threadIdx.x == 0) {
        union U { acc_t a; uint32_t w[W]; B200_DEVICE U() {} } u;
#pragma unroll
        for (int w = 0; w < W; ++w) u.w[w] = 0u;
        u.a = mine;
#pragma unroll
        for (int w = 0; w < W; ++w) ex_mine[w] = u.w[w];
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t < ex.nranks * W) {
        const int p = t / W, w = t % W;
        const size_t base = size_t(ex.tag & 1u) * kMaxPeers * kExWords;
        const uint64_t word = (uint64_t(ex.tag) << 32) | ex_mine[w];
        uint64_t* dst = ex.slots[p] + base + size_t(ex.rank) * kExWords + w;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
        const uint64_t* src = ex.slots[ex.rank] + base + size_t(p) * kExWords + w;
        uint64_t v;
        uint32_t polls = 0;
        do {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
            // a rank that never shows up (crashed peer, mismatched call order) must not hang the GPU: give up
            // after ~2^26 polls (tens of seconds) with a trap, which the host sees as a launch failure
            if (++polls == (1u << 26)) __trap();
        } while (static_cast<uint32_t>(v >> 32) != ex.tag);
        ex_words[p][w] = static_cast<uint32_t>(v);
    }
    __syncthreads();
    union U { acc_t a; uint32_t w[W]; B200_DEVICE U() {} } u;
#pragma unroll
    for (int w = 0; w < W; ++w) u.w[w] = ex_words[0][w];
    acc_t r = u.a;
    for (int p = 1; p < ex.nranks; ++p) {
#pragma unroll
        for (int w = 0; w < W; ++w) u.w[w] = ex_words[p][w];
        r = op.combine(r, u.a);
    }
    return r;
}

template <class Op, int VEC, int UNROLL, int THREADS>
__device__ __forceinline__ void reduce_full_body(
        const Op& op, typename in_ptr<Op>::type x, typename Op::out_t* __restrict__ y,
        int64_t n, typename Op::acc_t* partials, uint32_t* ticket, const PeerEx* ex = nullptr) {
    typedef typename Op::acc_t acc_t;
    typedef typename Op::index_t index_t;
    constexpr int64_t kTile = int64_t(THREADS) * VEC * UNROLL;
    // raw storage: acc_t may have user constructors (the reference's min_max_st)
    __shared__ __align__(16) char smem_raw[(THREADS / 32) * sizeof(acc_t)];
    acc_t* smem = reinterpret_cast<acc_t*>(smem_raw);
    __shared__ bool is_last;

    ThreadAcc<Op, UNROLL, VEC> ta(op);
    for (int64_t base = int64_t(blockIdx.x) * kTile; base < n; base += int64_t(gridDim.x) * kTile) {
        if (base + kTile <= n) {
            Pack<typename Op::in_t, VEC> v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                load_pack(v[u], x + base + (int64_t(u) * THREADS + threadIdx.x) * VEC);
            ta.fold(v, static_cast<index_t>(base + int64_t(threadIdx.x) * VEC), static_cast<index_t>(THREADS * VEC),
                    static_cast<index_t>(1));
        } else {
            for (int64_t i = base + threadIdx.x; i < n; i += THREADS) ta.fold_one(x[i], static_cast<index_t>(i));
        }
    }
    acc_t r = ta.result();
    r = block_combine(op, r, smem);

    if (gridDim.x == 1) {
        if (ex != nullptr && ex->nranks > 1) {
            r = peer_allreduce<Op, THREADS>(op, r, *ex);
            n = ex->n_total;
        }
        if (threadIdx.x == 0) y[0] = op.post(r, n);
        return;
    }
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = r;
        __threadfence();
        // self-resetting ticket: atomicInc wraps to 0 on the last arrival, so the workspace is left zeroed
        // without a separate store (one launch = exactly gridDim.x arrivals)
        const uint32_t t = atomicInc(ticket, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    r = op.identity();
    for (int i = threadIdx.x; i < int(gridDim.x); i += THREADS)
        r = op.combine(r, load_volatile(partials + i));
    r = block_combine(op, r, smem);
    if (ex != nullptr && ex->nranks > 1) {         // sharded: fold in the other GPUs' partials (block-uniform branch)
        r = peer_allreduce<Op, THREADS>(op, r, *ex);
        n = ex->n_total;
    }
    if (threadIdx.x == 0) y[0] = op.post(r, n);
}

// ---------------------------------------------------------------------------
// ROWS: x[rows][n] (n contiguous) -> y[rows].  GROUP threads cooperate on a row:
// GROUP == THREADS (a block per row) for long rows, 32 / 8 / 1 for short ones so
// that small segments still fill the machine.  Grid-stride over rows.
// Host guarantees for VEC > 1: n % VEC == 0 and 16-byte style alignment of x.
// ---------------------------------------------------------------------------
template <class Op, int VEC, int UNROLL, int THREADS, int GROUP>
__device__ __forceinline__ void reduce_rows_body(
        const Op& op, typename in_ptr<Op>::type x, typename Op::out_t* __restrict__ y,
        int64_t rows, int64_t n) {
    typedef typename Op::acc_t acc_t;
    typedef typename Op::index_t index_t;
    constexpr int kRowsPerBlock = THREADS / GROUP;
    constexpr int64_t kTile = int64_t(GROUP) * VEC * UNROLL;
    // raw storage: acc_t may have user constructors (the reference's min_max_st)
    __shared__ __align__(16) char smem_raw[(THREADS / 32) * sizeof(acc_t)];
    acc_t* smem = reinterpret_cast<acc_t*>(smem_raw);
    const int g = threadIdx.x % GROUP;       // lane within the row group
    const int gi = threadIdx.x / GROUP;      // which row of the block

    for (int64_t row0 = int64_t(blockIdx.x) * kRowsPerBlock; row0 < rows;
         row0 += int64_t(gridDim.x) * kRowsPerBlock) {
        const int64_t row = row0 + gi;
        const bool live = row < rows;
        const typename in_ptr<Op>::type xr = row_ptr(x, live ? row : 0, n);
        ThreadAcc<Op, UNROLL, VEC> ta(op);
        if (live) {
            int64_t base = 0;
            for (; base + kTile <= n; base += kTile) {
                Pack<typename Op::in_t, VEC> v[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) load_pack(v[u], xr + base + (int64_t(u) * GROUP + g) * VEC);
                ta.fold(v, static_cast<index_t>(base + int64_t(g) * VEC), static_cast<index_t>(GROUP * VEC),
                        static_cast<index_t>(1));
            }
            // whole vectors that do not fill an unrolled batch: one 16-byte load per lane and step
            // (rows shorter than GROUP * VEC * UNROLL live here entirely).  Small groups only: in the
            // long-row instantiations the extra lane states cost registers (fp16 var -15 %) for nothing.
            if (VEC > 1 && GROUP <= 8) {
                // fewer than UNROLL steps are left: issue all their loads before folding any (rows of 64-127
                // elements live here entirely; one load in flight per lane ran them at 72-80 % of peak)
                constexpr int64_t kStep = int64_t(GROUP) * VEC;
                Pack<typename Op::in_t, VEC> v1[UNROLL - 1];
                bool ok[UNROLL - 1];
#pragma unroll
                for (int u = 0; u < UNROLL - 1; ++u) {
                    ok[u] = base + (u + 1) * kStep <= n;
                    if (ok[u]) load_pack(v1[u], xr + base + u * kStep + int64_t(g) * VEC);
                }
#pragma unroll
                for (int u = 0; u < UNROLL - 1; ++u) {
                    if (ok[u]) {
#pragma unroll
                        for (int k = 0; k < VEC; ++k)
                            ta.fold_one_lane(k, v1[u][k], static_cast<index_t>(base + u * kStep + int64_t(g) * VEC + k));
                    }
                }
#pragma unroll
                for (int u = 0; u < UNROLL - 1; ++u) base += ok[u] ? kStep : 0;
            }
            for (int64_t i = base + g; i < n; i += GROUP) ta.fold_one(xr[i], static_cast<index_t>(i));
        }
        acc_t r = ta.result();
        if (GROUP == THREADS) {
            r = block_combine(op, r, smem);
            if (threadIdx.x == 0) y[row] = op.post(r, n);
        } else {
            if (GROUP > 1) r = group_combine<(GROUP > 32 ? 32 : GROUP)>(op, r);
            if (g == 0 && live) y[row] = op.post(r, n);
        }
    }
}

// The last block of a column tile folds the row splits (reduce_cols_body).  __noinline__: it runs once per tile;
// kept out of line, its batched loads do not raise the register count of the streaming loop (inlined, the
// float16 moments kernel went from 80 to 95 registers and from 3 to 2 resident blocks: 95 % -> 71 % of peak).
template <class Op, int VEC, int WC>
__device__ __noinline__ void cols_fold_splits(const Op& op, const typename Op::acc_t* partials,
                                              typename Op::out_t* __restrict__ y, int64_t b, int nsplit, int64_t cols,
                                              int64_t tile_c0, int64_t n, typename Op::acc_t* smem_flat) {
    typedef typename Op::acc_t acc_t;
    constexpr int kWarps = 8;
    constexpr int kTileCols = 32 * VEC;
    constexpr int kBlockCols = WC * kTileCols;
    acc_t (*smem)[kBlockCols] = reinterpret_cast<acc_t (*)[kBlockCols]>(smem_flat);
    // The last block of the tile folds the splits.  Threads are laid out as (pack of VEC columns) x (kLanes
    // split lanes): each thread folds every kLanes-th split of its pack with vector L2 loads, several
    // independent loads in flight -- a serial walk would be one dependent L2 round trip per split, which is what
    // bounds L2-resident and tall-narrow matrices -- and the split lanes meet through shared memory.
    constexpr int kQuads = kBlockCols / VEC;              // WC * 32 packs across the tile
    constexpr int kLanes = kWarps * 32 / kQuads;          // == kWarpRows
    struct Q { acc_t v[VEC]; };
    constexpr int kBatch = sizeof(Q) <= 16 ? 8 : (sizeof(Q) <= 32 ? 2 : 1);  // more in flight costs the whole kernel registers
    const int q = threadIdx.x % kQuads, g = threadIdx.x / kQuads;
    const int64_t c = tile_c0 + int64_t(q) * VEC;
    const bool live = c < cols && g < nsplit;             // VEC > 1 => cols % VEC == 0 => whole pack in range
    Q a;
    if (live) {
        const acc_t* col = partials + (b * nsplit) * cols + c;
        a = load_cg(reinterpret_cast<const Q*>(col + int64_t(g) * cols));
        int s = g + kLanes;
        for (; s + (kBatch - 1) * kLanes < nsplit; s += kBatch * kLanes) {
            Q v[kBatch];
#pragma unroll
            for (int k = 0; k < kBatch; ++k) v[k] = load_cg(reinterpret_cast<const Q*>(col + int64_t(s + k * kLanes) * cols));
#pragma unroll
            for (int k = 0; k < kBatch; ++k) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) a.v[e] = op.combine(a.v[e], v[k].v[e]);
            }
        }
        for (; s < nsplit; s += kLanes) {
            const Q v = load_cg(reinterpret_cast<const Q*>(col + int64_t(s) * cols));
#pragma unroll
            for (int e = 0; e < VEC; ++e) a.v[e] = op.combine(a.v[e], v.v[e]);
        }
    }
    if (kLanes == 1) {
        if (live) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) y[b * cols + c + e] = op.post(a.v[e], n);
        }
        return;
    }
    if (live) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) smem[g][q * VEC + e] = a.v[e];
    }
    __syncthreads();
    const int groups = nsplit < kLanes ? nsplit : kLanes;
    for (int t = threadIdx.x; t < kBlockCols; t += blockDim.x) {
        if (tile_c0 + t >= cols) continue;
        acc_t r = smem[0][t];
        for (int w = 1; w < groups; ++w) r = op.combine(r, smem[w][t]);
        y[b * cols + tile_c0 + t] = op.post(r, n);
    }
}

// ---------------------------------------------------------------------------
// COLS: x[batch][n][cols] (cols contiguous) -> y[batch][cols]: the reduced axis
// is the STRIDED one.  Block = 8 warps; a warp reads 32*VEC consecutive columns
// of one row per instruction (full 128-byte lines) and walks down the rows with
// RU independent loads in flight; the 8 warps' states meet in shared memory.
// gridDim = (column tiles, row splits, batch).  With more than one row split the
// split partials go to the workspace and the last block of the column tile
// (atomic ticket) folds them in row order.
// workspace: batch*nsplit*cols partials, then batch*gridDim.x zeroed tickets.
// ---------------------------------------------------------------------------
// WC = warps standing side by side across the columns (1 or 8); the other 8 / WC warp rows split the
// reduced rows and meet in shared memory.  WC = 8 makes a block read 8 * 32 * VEC contiguous elements
// of every row (4 KB for float32: whole DRAM pages instead of 512-byte strips): 94 % -> 109 % of the
// copy peak on sum / max over axis 0 of a 32768^2 float32 array (WC = 4, 2 KB per visit: no gain).
// Functors with their own lane state (arg-reductions, moments) keep WC = 1: fewer column tiles mean more
// row splits, and their split partials are expensive to merge (measured 2x slower at WC = 8).
template <class Op, int VEC, int RU, int WC = 1>
__device__ __forceinline__ void reduce_cols_body(
        const Op& op, typename in_ptr<Op>::type x, typename Op::out_t* __restrict__ y,
        int64_t n, int64_t cols, typename Op::acc_t* partials, uint32_t* tickets) {
    typedef typename Op::acc_t acc_t;
    typedef typename Op::index_t index_t;
    constexpr int kWarps = 8;
    constexpr int kWarpRows = kWarps / WC;            // warps sharing a column strip
    constexpr int kTileCols = 32 * VEC;
    constexpr int kBlockCols = WC * kTileCols;
    __shared__ __align__(16) char smem_raw[kWarps * kTileCols * sizeof(acc_t)];
    acc_t (*smem)[kBlockCols] = reinterpret_cast<acc_t (*)[kBlockCols]>(smem_raw);
    __shared__ bool is_last;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wc = warp % WC, wr = warp / WC;
    const int64_t b = blockIdx.z;
    const int nsplit = gridDim.y, split = blockIdx.y;
    const int64_t c0 = ((int64_t(blockIdx.x) * WC + wc) * 32 + lane) * VEC;
    const bool col_ok = c0 < cols;     // VEC > 1 => cols % VEC == 0 => whole pack in range
    const int64_t rows_per_split = (n + nsplit - 1) / nsplit;
    const int64_t r_begin = int64_t(split) * rows_per_split;
    const int64_t r_end = (r_begin + rows_per_split < n) ? r_begin + rows_per_split : n;
    const typename in_ptr<Op>::type xb = cols_ptr(x, b, n, cols, c0);

    ThreadAcc<Op, RU, VEC> ta(op);
    if (col_ok) {
        int64_t r = r_begin + wr;
        for (; r + int64_t(RU - 1) * kWarpRows < r_end; r += int64_t(RU) * kWarpRows) {
            Pack<typename Op::in_t, VEC> v[RU];
#pragma unroll
            for (int u = 0; u < RU; ++u) load_pack(v[u], cols_row(xb, r + int64_t(u) * kWarpRows, cols));
            ta.fold(v, static_cast<index_t>(r), static_cast<index_t>(kWarpRows), static_cast<index_t>(0));
        }
        for (; r < r_end; r += kWarpRows) {
            Pack<typename Op::in_t, VEC> v;
            load_pack(v, cols_row(xb, r, cols));
#pragma unroll
            for (int k = 0; k < VEC; ++k) ta.fold_one_lane(k, v[k], static_cast<index_t>(r));
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) smem[wr][(wc * 32 + lane) * VEC + k] = ta.result_lane(k);
    __syncthreads();

    const int64_t tile_c0 = int64_t(blockIdx.x) * kBlockCols;
    for (int t = threadIdx.x; t < kBlockCols; t += blockDim.x) {
        if (tile_c0 + t >= cols) continue;
        acc_t a = smem[0][t];
#pragma unroll
        for (int w = 1; w < kWarpRows; ++w) a = op.combine(a, smem[w][t]);
        if (nsplit == 1) y[b * cols + tile_c0 + t] = op.post(a, n);
        else partials[(b * nsplit + split) * cols + tile_c0 + t] = a;
    }
    if (nsplit == 1) return;

    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t t = atomicInc(&tickets[b * gridDim.x + blockIdx.x], uint32_t(nsplit) - 1);   // self-resetting
        is_last = (t == uint32_t(nsplit) - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    cols_fold_splits<Op, VEC, WC>(op, partials, y, b, nsplit, cols, tile_c0, n, &smem[0][0]);
}

// ---------------------------------------------------------------------------
// COLS, narrow: x[n][cols] with short rows (cols <= 64 elements: point clouds, feature tables).  The strip kernel above leaves most lanes of a warp without a column there; this one reads
// the matrix as the flat stream it is -- grid-stride 16-byte loads, the FULL reduction's access pattern -- with
// `active` threads per block, chosen by the host so that active * VEC is a multiple of cols: the k-th element of
// a thread's vector then falls in the same column, (t * VEC + k) % cols, in every chunk, so the thread's VEC
// lane accumulators ARE column accumulators.  A chunk is active * VEC / cols whole rows; the row handed to the
// functor is the chunk's first row, and the lane's constant row inside a chunk, (t * VEC + k) / cols, is added
// to arg-reduction results afterwards (shift_index).  Lanes meet by column through shared memory, blocks
// through the workspace (gridDim.x * cols partials + the self-resetting ticket); every fold order is fixed.
// ---------------------------------------------------------------------------
template <class A> B200_DEVICE void shift_index(A&, long long) {}     // overloaded for (value, index) accumulators

// Column totals of `count` row-major rows of `cols` accumulators starting at `rows`: the threads of a block are laid
// out as (column, group); group gi folds rows gi, gi + G, ...; the groups' results meet in `groups`.  The value
// returned to thread t < cols is the total of column t (other threads get a don't-care).  Contains a barrier.
template <class Op, int THREADS, bool FROM_GLOBAL>
B200_DEVICE typename Op::acc_t fold_by_column(const Op& op, const typename Op::acc_t* rows, int count, int cols,
                                              typename Op::acc_t* groups) {
    typedef typename Op::acc_t acc_t;
    const int t = threadIdx.x, G = THREADS / cols;
    const int c = t % cols, gi = t / cols;
    const bool has = gi < G && gi < count;
    if (has) {
        const acc_t* p = rows + c;
        acc_t a = FROM_GLOBAL ? load_cg(p + int64_t(gi) * cols) : p[int64_t(gi) * cols];
        int r = gi + G;
        if (FROM_GLOBAL) {
            constexpr int kBatch = sizeof(acc_t) <= 8 ? 8 : 4;       // independent L2 loads in flight
            for (; r + (kBatch - 1) * G < count; r += kBatch * G) {
                acc_t v[kBatch];
#pragma unroll
                for (int k = 0; k < kBatch; ++k) v[k] = load_cg(p + int64_t(r + k * G) * cols);
#pragma unroll
                for (int k = 0; k < kBatch; ++k) a = op.combine(a, v[k]);
            }
        }
        for (; r < count; r += G) a = op.combine(a, FROM_GLOBAL ? load_cg(p + int64_t(r) * cols) : p[int64_t(r) * cols]);
        groups[gi * cols + c] = a;
    }
    __syncthreads();
    acc_t r = groups[t < cols ? t : 0];
    if (t < cols) {
        const int used = count < G ? count : G;
        for (int w = 1; w < used; ++w) r = op.combine(r, groups[w * cols + t]);
    }
    return r;
}

template <class Op, int VEC, int UNROLL, int THREADS>
__device__ __forceinline__ void reduce_narrow_body(
        const Op& op, typename in_ptr<Op>::type x, typename Op::out_t* __restrict__ y,
        int64_t n, int cols, int active, typename Op::acc_t* partials, uint32_t* ticket) {
    typedef typename Op::acc_t acc_t;
    typedef typename Op::index_t index_t;
    // raw storage: acc_t may have user constructors
    __shared__ __align__(16) char lanes_raw[THREADS * VEC * sizeof(acc_t)];
    __shared__ __align__(16) char groups_raw[THREADS * sizeof(acc_t)];
    acc_t* lanes = reinterpret_cast<acc_t*>(lanes_raw);
    acc_t* groups = reinterpret_cast<acc_t*>(groups_raw);
    __shared__ bool is_last;

    const int t = threadIdx.x;
    const int chunk_rows = active * VEC / cols;
    const int64_t chunk_elems = int64_t(active) * VEC;
    const int64_t chunks = n / chunk_rows;
    const bool on = t < active;

    if (on) {
        ThreadAcc<Op, UNROLL, VEC> ta(op);
        const typename in_ptr<Op>::type xt = x + int64_t(t) * VEC;
        const int64_t g = gridDim.x;
        int64_t c = blockIdx.x;
        for (; c + int64_t(UNROLL - 1) * g < chunks; c += int64_t(UNROLL) * g) {
            Pack<typename Op::in_t, VEC> v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) load_pack(v[u], xt + (c + int64_t(u) * g) * chunk_elems);
            ta.fold(v, static_cast<index_t>(c * chunk_rows), static_cast<index_t>(g * chunk_rows), static_cast<index_t>(0));
        }
        for (; c < chunks; c += g) {
            Pack<typename Op::in_t, VEC> v;
            load_pack(v, xt + c * chunk_elems);
#pragma unroll
            for (int k = 0; k < VEC; ++k) ta.fold_one_lane(k, v[k], static_cast<index_t>(c * chunk_rows));
        }
        if (blockIdx.x == gridDim.x - 1) {                 // the rows past the last whole chunk
            const int64_t e0 = chunks * chunk_elems + int64_t(t) * VEC, total = n * cols;
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                if (e0 + k < total) ta.fold_one_lane(k, x[e0 + k], static_cast<index_t>(chunks * chunk_rows));
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            acc_t a = ta.result_lane(k);
            shift_index(a, (t * VEC + k) / cols);
            lanes[t * VEC + k] = a;
        }
    }
    __syncthreads();
    const bool mine = t < cols;
    acc_t r = fold_by_column<Op, THREADS, false>(op, lanes, chunk_rows, cols, groups);
    if (gridDim.x == 1) {
        if (mine) y[t] = op.post(r, n);
        return;
    }
    if (mine) partials[int64_t(blockIdx.x) * cols + t] = r;
    __threadfence();
    __syncthreads();
    if (t == 0) {
        const uint32_t k = atomicInc(ticket, gridDim.x - 1);      // self-resetting
        is_last = (k == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    r = fold_by_column<Op, THREADS, true>(op, partials, int(gridDim.x), cols, groups);
    if (mine) y[t] = op.post(r, n);
}

// ---------------------------------------------------------------------------
// ROWS, short: x[rows][len] with len <= 64 (per-point norms, argmax over a few classes).  The group kernels above
// give a row to 1 / 8 / 32 lanes and end up with one small load in flight per lane; here a tile of whole rows is
// staged in shared memory with the flat stream's coalesced 16-byte loads (the next tile's loads already in flight),
// and a thread then folds a whole row out of shared memory (the padded layout keeps the 32 walks of a warp on
// different banks).  Results leave coalesced: consecutive threads own consecutive rows.
// ---------------------------------------------------------------------------
B200_DEVICE int short_pad(int i) { return i + (i >> 5); }

template <class Op, int VEC, int U, int THREADS>
__device__ __forceinline__ void reduce_short_rows_body(
        const Op& op, const typename Op::in_t* __restrict__ x, typename Op::out_t* __restrict__ y,
        int64_t rows, int len, int active) {
    typedef typename Op::in_t in_t;
    typedef typename Op::acc_t acc_t;
    typedef typename Op::index_t index_t;
    constexpr int SLOTS = THREADS * VEC * U;
    __shared__ __align__(16) char tile_raw[(SLOTS + SLOTS / 32 + 1) * sizeof(in_t)];
    in_t* tile = reinterpret_cast<in_t*>(tile_raw);
    const int t = threadIdx.x;
    const int chunk_elems = active * VEC;
    const int tile_elems = chunk_elems * U, tile_rows = tile_elems / len;
    const int64_t total = rows * len;
    const int64_t tiles = (total + tile_elems - 1) / tile_elems;
    Pack<in_t, VEC> pre[U];
    auto fetch = [&](int64_t tl) {
        if (t < active && tl < tiles) {
            const int64_t e0 = tl * tile_elems;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t ge = e0 + u * chunk_elems + t * VEC;
                if (ge + VEC <= total) {
                    load_pack(pre[u], x + ge);
                } else {
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        if (ge + k < total) pre[u][k] = x[ge + k];
                }
            }
        }
    };
    fetch(blockIdx.x);
    for (int64_t tl = blockIdx.x; tl < tiles; tl += gridDim.x) {
        if (t < active) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int f = u * chunk_elems + t * VEC;
#pragma unroll
                for (int k = 0; k < VEC; ++k) tile[short_pad(f + k)] = pre[u][k];     // past the end: never read
            }
        }
        fetch(tl + gridDim.x);
        __syncthreads();
        const int64_t row0 = tl * tile_rows;
        const int64_t left = rows - row0;
        const int here = left < tile_rows ? int(left) : tile_rows;
        for (int r = t; r < here; r += THREADS) {
            acc_t a = op.identity();
            const int base = r * len;
            for (int j = 0; j < len; ++j) op.accumulate(a, op.step(j + 1), tile[short_pad(base + j)], static_cast<index_t>(j));
            y[row0 + r] = op.post(a, len);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// GENERIC: any strides, any number of (broadcast) inputs and outputs -- the
// fallback for layouts that are none of FULL / ROWS / COLS and for user
// ReductionKernels with several array operands (e.g. the reference's
// _var_core kernels: x, broadcast mean, alpha).  GROUP threads per output
// element; every element's offsets come from an index decomposition, so this
// is correct for everything and fast for nothing in particular.
// The functor provides map_at / post_at, which read / write through pointers.
// ---------------------------------------------------------------------------
template <int NIN, int NOUT>
struct GenericReduceParams {
    int32_t out_ndim, red_ndim;
    int64_t out_size, red_size;
    int64_t out_shape[kMaxNdim], red_shape[kMaxNdim];
    int64_t in_out_strides[NIN][kMaxNdim];    // input strides along the out dims (bytes)
    int64_t in_red_strides[NIN][kMaxNdim];    // input strides along the reduced dims
    int64_t out_strides[NOUT][kMaxNdim];
    char*   in_ptr[NIN];
    char*   out_ptr[NOUT];
};

template <class Op, int NIN, int NOUT, int THREADS, int GROUP>
__device__ __forceinline__ void reduce_generic_body(const Op& op, const GenericReduceParams<NIN, NOUT>& p) {
    typedef typename Op::acc_t acc_t;
    typedef typename Op::index_t index_t;
    constexpr int kPerBlock = THREADS / GROUP;
    // raw storage: acc_t may have user constructors (the reference's min_max_st)
    __shared__ __align__(16) char smem_raw[(THREADS / 32) * sizeof(acc_t)];
    acc_t* smem = reinterpret_cast<acc_t*>(smem_raw);
    const int g = threadIdx.x % GROUP, gi = threadIdx.x / GROUP;

    for (int64_t o0 = int64_t(blockIdx.x) * kPerBlock; o0 < p.out_size; o0 += int64_t(gridDim.x) * kPerBlock) {
        const int64_t o = o0 + gi;
        const bool live = o < p.out_size;
        const char* ibase[NIN];
        char* obase[NOUT];
#pragma unroll
        for (int a = 0; a < NIN; ++a) ibase[a] = p.in_ptr[a];
#pragma unroll
        for (int a = 0; a < NOUT; ++a) obase[a] = p.out_ptr[a];
        if (live) {
            int64_t rest = o;
#pragma unroll 1
            for (int d = p.out_ndim - 1; d >= 0; --d) {
                const int64_t q = rest / p.out_shape[d], r = rest - q * p.out_shape[d];
                rest = q;
#pragma unroll
                for (int a = 0; a < NIN; ++a) ibase[a] += r * p.in_out_strides[a][d];
#pragma unroll
                for (int a = 0; a < NOUT; ++a) obase[a] += r * p.out_strides[a][d];
            }
        }
        acc_t acc = op.identity();
        if (live) {
            for (int64_t j = g; j < p.red_size; j += GROUP) {
                const char* ptrs[NIN];
#pragma unroll
                for (int a = 0; a < NIN; ++a) ptrs[a] = ibase[a];
                int64_t rest = j;
#pragma unroll 1
                for (int d = p.red_ndim - 1; d >= 0; --d) {
                    const int64_t q = rest / p.red_shape[d], r = rest - q * p.red_shape[d];
                    rest = q;
#pragma unroll
                    for (int a = 0; a < NIN; ++a) ptrs[a] += r * p.in_red_strides[a][d];
                }
                acc = op.combine(acc, op.map_at(ptrs, static_cast<index_t>(j), static_cast<ptrdiff_t>(o)));
            }
        }
        if (GROUP == THREADS) {
            acc = block_combine(op, acc, smem);
            if (threadIdx.x == 0) op.post_at(obase, acc, static_cast<ptrdiff_t>(o));
        } else {
            if (GROUP > 1) acc = group_combine<(GROUP > 32 ? 32 : GROUP)>(op, acc);
            if (g == 0 && live) op.post_at(obase, acc, static_cast<ptrdiff_t>(o));
        }
    }
}

}  // namespace b200
