// b200/elementwise.cuh -- the elementwise skeleton.
//
// One loop body ("glue"), three tilers.  A tiler decides which elements a thread
// owns in the current tile and how operands travel between global memory and the
// thread's register packs; the glue (prebuilt template in csrc/elementwise.cu, or
// text generated around a user's operation string and compiled by NVRTC) is the
// same for all three:
//
//     Tiler t(p);
//     for (; t.valid(); t.next())            // FULL = the tile lies wholly inside the array:
//         tile<FULL = t.is_full()>:          // no predicates, no tail code on that path
//             Pack<T0,V> a[U]; t.load<FULL>(0, a);  ...            // inputs
//             Pack<TO,V> o[U];                                      // outputs
//             for u, k:  if (t.in_range<FULL>(u,k)) { i = t.index(u,k); o[u][k] = f(a[u][k]...); }
//             t.store<FULL>(NIN, o);
//
// Replaces the reference's one-element-per-thread CUPY_FOR loop
// (cupy/_core/_kernel.pyx:86-97, cupy/_core/include/cupy/carray.cuh:68-72) and
// its per-element CIndexer div/mod (carray.cuh:588-615).
#pragma once
#include "base.cuh"
#include "tma.cuh"

namespace b200 {

struct EwArg {
    union {
        char*   ptr;          // array operand: base pointer
        int64_t scalar[2];    // scalar operand: raw value bytes
    };
    int64_t strides[kMaxNdim];   // bytes, collapsed dims
};

struct EwParams {
    int64_t  size;                 // loop elements
    int32_t  ndim;                 // collapsed rank
    int32_t  tile_axis;            // TILED only
    uint32_t staged_mask;          // TILED only
    uint32_t scalar_mask;          // bit k set = operand k is a by-value scalar
    int32_t  tma_stages;           // TILED_TMA only: depth of the shared-memory tile ring
    int32_t  tma_pad_;
    int64_t  shape[kMaxNdim];
    int64_t  cstride[kMaxNdim];    // C-order element strides of `shape` (for the linear index `i`)
    FastDiv  fdiv[kMaxNdim];       // fast division by shape[d] (valid when size < 2^31)
    FastDiv  fdiv_chunks;          // ROWWISE: fast division by shape[ndim-1] / vec (vectors per row)
    EwArg    arg[kMaxArgs];
};

// TILED_TMA: one opaque 128-byte CUtensorMap per staged operand (in operand order)
constexpr int kMaxStaged = 4;
struct alignas(64) TensorMapBlob { uint64_t v[16]; };
struct TileMaps { TensorMapBlob m[kMaxStaged]; };

struct true_t { static constexpr bool value = true; };
struct false_t { static constexpr bool value = false; };

template <class T>
B200_DEVICE T scalar_arg(const EwParams& p, int a) {
    return *reinterpret_cast<const T*>(&p.arg[a].scalar[0]);
}

// ---------------------------------------------------------------------------
// FLAT: every array operand is dense with one common layout -> 1-D.
// Tile = THREADS*VEC*UNROLL elements; unroll step u covers a contiguous span of
// THREADS*VEC elements so each warp instruction touches 32*VEC*sizeof(T)
// consecutive bytes (128-bit accesses when VEC*sizeof(T) >= 16).  Persistent
// grid-stride over tiles.
// ---------------------------------------------------------------------------
// PERIODIC (compile time: the plain form must not carry its branches, they break the compiler's
// load batching): some operands are row vectors broadcast over the rows of a dense array.
template <int NARGS, int VEC, int UNROLL, int THREADS, bool PERIODIC = false>
struct FlatTiler {
    static constexpr int kV = VEC, kU = UNROLL;
    static constexpr int64_t kTile = int64_t(THREADS) * VEC * UNROLL;
    const EwParams& p;
    int64_t base;
    bool full;

    int poff;     // periodic operands: this thread's element offset inside the period (same for every tile and u)

    B200_DEVICE explicit FlatTiler(const EwParams& p_) : p(p_) {
        base = int64_t(blockIdx.x) * kTile;
        full = base + kTile <= p.size;
        // FLAT with periodic operands (p.staged_mask, period p.tile_axis elements): a row vector
        // broadcast over a dense array.  The host guarantees THREADS * VEC % period == 0, so the
        // element (base + (u * THREADS + tid) * VEC) % period does not depend on the tile or on u.
        poff = PERIODIC ? int((int64_t(threadIdx.x) * VEC) % p.tile_axis) : 0;
    }
    B200_DEVICE bool valid() const { return base < p.size; }
    B200_DEVICE void next() {
        base += int64_t(gridDim.x) * kTile;
        full = base + kTile <= p.size;
    }
    B200_DEVICE int64_t index(int u, int k) const {
        return base + (int64_t(u) * THREADS + threadIdx.x) * VEC + k;
    }
    B200_DEVICE bool is_full() const { return full; }
    template <bool FULL>
    B200_DEVICE bool in_range(int u, int k) const { return FULL || index(u, k) < p.size; }

    template <bool FULL, class T>
    B200_DEVICE void load(int a, Pack<T, VEC> (&r)[UNROLL]) const {
        const T* __restrict__ ptr = reinterpret_cast<const T*>(p.arg[a].ptr);
        if (PERIODIC && ((p.staged_mask >> a) & 1u)) {
            // periodic operand: one vector serves every unroll step (period % VEC == 0: never straddles)
            Pack<T, VEC> v;
            load_pack(v, ptr + poff);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) r[u] = v;
            return;
        }
        if (FULL) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) load_pack(r[u], ptr + index(u, 0));
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t i0 = index(u, 0);
                if (i0 + VEC <= p.size) {
                    load_pack(r[u], ptr + i0);
                } else {
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        if (i0 + k < p.size) r[u][k] = ptr[i0 + k];
                }
            }
        }
    }
    template <bool FULL, class T>
    B200_DEVICE void store(int a, const Pack<T, VEC> (&r)[UNROLL]) const {
        T* __restrict__ ptr = reinterpret_cast<T*>(p.arg[a].ptr);
        if (FULL) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) store_pack(ptr + index(u, 0), r[u]);
        } else {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int64_t i0 = index(u, 0);
                if (i0 + VEC <= p.size) {
                    store_pack(ptr + i0, r[u]);
                } else {
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        if (i0 + k < p.size) ptr[i0 + k] = r[u][k];
                }
            }
        }
    }
};

// ---------------------------------------------------------------------------
// ROWWISE: N-D strided / broadcast operands.  A work item is VEC consecutive
// elements of the innermost dim; its outer coordinates are decomposed ONCE (fast
// division, 32-bit when IDX32) into a byte offset per operand.  Per operand the
// innermost stride selects the access: == sizeof(T) -> one vector load,
// 0 -> one scalar load broadcast to the VEC lanes, otherwise VEC strided loads.
// Host guarantees for VEC > 1: shape[ndim-1] % VEC == 0 and VEC*sizeof(T)
// alignment of every unit-stride operand (base and outer strides).
// ---------------------------------------------------------------------------
template <bool IDX32> struct row_offset { typedef int64_t type; };
template <> struct row_offset<true> { typedef int32_t type; };      // the planner guarantees every span < 2^31

// SPEC: two bits per operand fixing its innermost-stride kind at compile time (NVRTC kernels know
// it; 0 = decide at run time): 1 = unit stride (vector access), 2 = stride 0 (one scalar, broadcast
// to the lanes), 3 = other (scalar accesses).  Without the run-time branches the compiler batches all
// loads of a tile ahead of the arithmetic.
template <int NARGS, int VEC, int UNROLL, int THREADS, bool IDX32, uint64_t SPEC = 0>
struct RowTiler {
    static constexpr int kV = VEC, kU = UNROLL;
    typedef typename row_offset<IDX32>::type off_t;
    const EwParams& p;
    int64_t inner, chunks, total, wbase;
    off_t off[UNROLL][NARGS];
    int64_t lin[UNROLL];
    bool ok[UNROLL];

    B200_DEVICE explicit RowTiler(const EwParams& p_) : p(p_) {
        inner = p.shape[p.ndim - 1];
        chunks = inner / VEC;   // host guarantees inner % VEC == 0
        total = (p.size / inner) * chunks;
        wbase = int64_t(blockIdx.x) * (THREADS * UNROLL);
        locate();
    }
    B200_DEVICE void locate() {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t w = wbase + int64_t(u) * THREADS + threadIdx.x;
            ok[u] = w < total;
#pragma unroll
            for (int a = 0; a < NARGS; ++a) off[u][a] = 0;
            lin[u] = 0;
            if (!ok[u]) continue;
            if (IDX32) {
                // all 32-bit: one multiply-high division per dim, 32-bit offsets per operand
                uint32_t rest = static_cast<uint32_t>(w), c, q;
                p.fdiv_chunks.divmod(rest, q, c);     // innermost: chunk index
                rest = q;
                const uint32_t e0 = c * VEC;
                lin[u] = e0;
#pragma unroll
                for (int a = 0; a < NARGS; ++a)
                    off[u][a] = static_cast<off_t>(e0 * static_cast<uint32_t>(p.arg[a].strides[p.ndim - 1]));
#pragma unroll 1
                for (int d = p.ndim - 2; d >= 0; --d) {
                    uint32_t r;
                    p.fdiv[d].divmod(rest, q, r);
                    rest = q;
                    lin[u] += int64_t(r) * p.cstride[d];
#pragma unroll
                    for (int a = 0; a < NARGS; ++a)
                        off[u][a] += static_cast<off_t>(r * static_cast<uint32_t>(p.arg[a].strides[d]));
                }
            } else {
                int64_t rest = w;
                const int64_t c = rest % chunks;
                rest /= chunks;
                const int64_t e0 = c * VEC;
                lin[u] = e0;
#pragma unroll
                for (int a = 0; a < NARGS; ++a) off[u][a] = e0 * p.arg[a].strides[p.ndim - 1];
#pragma unroll 1
                for (int d = p.ndim - 2; d >= 0; --d) {
                    const int64_t r = rest % p.shape[d];
                    rest /= p.shape[d];
                    lin[u] += r * p.cstride[d];
#pragma unroll
                    for (int a = 0; a < NARGS; ++a) off[u][a] += r * p.arg[a].strides[d];
                }
            }
        }
    }
    B200_DEVICE bool valid() const { return wbase < total; }
    B200_DEVICE void next() {
        wbase += int64_t(gridDim.x) * (THREADS * UNROLL);
        locate();
    }
    B200_DEVICE int64_t index(int u, int k) const { return lin[u] + k; }
    B200_DEVICE bool is_full() const { return wbase + int64_t(THREADS) * UNROLL <= total; }
    template <bool FULL>
    B200_DEVICE bool in_range(int u, int) const { return FULL || ok[u]; }

    template <bool FULL, class T>
    B200_DEVICE void load(int a, Pack<T, VEC> (&r)[UNROLL]) const {
        const uint32_t kind = uint32_t(SPEC >> (2 * a)) & 3u;
        const int64_t si = kind == 1 ? int64_t(sizeof(T)) : kind == 2 ? 0 : p.arg[a].strides[p.ndim - 1];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (!FULL && !ok[u]) continue;
            const char* base = p.arg[a].ptr + off[u][a];
            if (VEC > 1 && (kind == 1 || (kind == 0 && si == int64_t(sizeof(T))))) {
                load_pack(r[u], reinterpret_cast<const T*>(base));
            } else if (VEC > 1 && (kind == 2 || (kind == 0 && si == 0))) {
                const T v = *reinterpret_cast<const T*>(base);
#pragma unroll
                for (int k = 0; k < VEC; ++k) r[u][k] = v;
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) r[u][k] = *reinterpret_cast<const T*>(base + k * si);
            }
        }
    }
    template <bool FULL, class T>
    B200_DEVICE void store(int a, const Pack<T, VEC> (&r)[UNROLL]) const {
        const uint32_t kind = uint32_t(SPEC >> (2 * a)) & 3u;
        const int64_t si = kind == 1 ? int64_t(sizeof(T)) : p.arg[a].strides[p.ndim - 1];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (!FULL && !ok[u]) continue;
            char* base = p.arg[a].ptr + off[u][a];
            if (VEC > 1 && (kind == 1 || (kind == 0 && si == int64_t(sizeof(T))))) {
                store_pack(reinterpret_cast<T*>(base), r[u]);
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k) *reinterpret_cast<T*>(base + k * si) = r[u][k];
            }
        }
    }
};

// ---------------------------------------------------------------------------
// TILED: some input is unit-stride along dim `tile_axis` (call it I) while the
// loop's innermost dim (O) is where the outputs are unit-stride -- a transpose.
// One 32x32 (I x O) tile per block; block = 32x8 threads, 4 elements per thread.
// Staged operands are read with the warp running along I (128-byte rows of the
// operand), parked in a 32x33 shared-memory tile, and picked up transposed so
// that compute and every direct operand / output run along O: full 128-byte
// lines on both sides, no bank conflicts.
// ---------------------------------------------------------------------------
template <int NARGS>
struct TileTiler {
    static constexpr int kV = 1, kU = 4;
    static constexpr int kTile = 32;
    const EwParams& p;
    int64_t n_i, n_o, i0, o0, lin0;
    int64_t boff[NARGS];
    int tx, ty;
    bool live;

    B200_DEVICE explicit TileTiler(const EwParams& p_) : p(p_) {
        tx = threadIdx.x & 31;
        ty = threadIdx.x >> 5;
        const int ax_i = p.tile_axis, ax_o = p.ndim - 1;
        n_i = p.shape[ax_i];
        n_o = p.shape[ax_o];
        const uint32_t tiles_i = static_cast<uint32_t>((n_i + kTile - 1) / kTile);
        const uint32_t tiles_o = static_cast<uint32_t>((n_o + kTile - 1) / kTile);
        uint32_t b = blockIdx.x;
        const uint32_t ti = b % tiles_i;
        b /= tiles_i;
        const uint32_t to = b % tiles_o;
        b /= tiles_o;
        i0 = int64_t(ti) * kTile;
        o0 = int64_t(to) * kTile;
        lin0 = 0;
#pragma unroll
        for (int a = 0; a < NARGS; ++a) boff[a] = 0;
#pragma unroll 1
        for (int d = p.ndim - 2; d >= 0; --d) {
            if (d == ax_i) continue;
            const uint32_t s = static_cast<uint32_t>(p.shape[d]);
            const uint32_t r = b % s;
            b /= s;
            lin0 += int64_t(r) * p.cstride[d];
#pragma unroll
            for (int a = 0; a < NARGS; ++a) boff[a] += int64_t(r) * p.arg[a].strides[d];
        }
        live = true;
    }
    B200_DEVICE bool valid() const { return live; }
    B200_DEVICE void next() { live = false; }
    B200_DEVICE bool is_full() const { return i0 + kTile <= n_i && o0 + kTile <= n_o; }
    // compute-phase ownership: element (i = i0 + ty + 8u, o = o0 + tx)
    template <bool FULL>
    B200_DEVICE bool in_range(int u, int) const { return FULL || ((i0 + ty + 8 * u) < n_i && (o0 + tx) < n_o); }
    B200_DEVICE int64_t index(int u, int) const {
        return lin0 + (i0 + ty + 8 * u) * p.cstride[p.tile_axis] + (o0 + tx);
    }

    template <bool FULL, class T>
    B200_DEVICE void load(int a, Pack<T, 1> (&r)[4]) const {
        const int64_t s_i = p.arg[a].strides[p.tile_axis];
        const int64_t s_o = p.arg[a].strides[p.ndim - 1];
        const char* base = p.arg[a].ptr + boff[a];
        if ((p.staged_mask >> a) & 1u) {
            typedef typename RawVec<sizeof(T)>::type W;
            __shared__ W tile[kTile][kTile + 1];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t o = o0 + ty + 8 * u, i = i0 + tx;
                if (FULL || (o < n_o && i < n_i))
                    tile[ty + 8 * u][tx] = *reinterpret_cast<const W*>(base + i * s_i + o * s_o);
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                W w = tile[tx][ty + 8 * u];
                r[u][0] = *reinterpret_cast<T*>(&w);
            }
            __syncthreads();
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (in_range<FULL>(u, 0))
                    r[u][0] = *reinterpret_cast<const T*>(base + (i0 + ty + 8 * u) * s_i + (o0 + tx) * s_o);
        }
    }
    template <bool FULL, class T>
    B200_DEVICE void store(int a, const Pack<T, 1> (&r)[4]) const {
        const int64_t s_i = p.arg[a].strides[p.tile_axis];
        const int64_t s_o = p.arg[a].strides[p.ndim - 1];
        char* base = p.arg[a].ptr + boff[a];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (in_range<FULL>(u, 0))
                *reinterpret_cast<T*>(base + (i0 + ty + 8 * u) * s_i + (o0 + tx) * s_o) = r[u][0];
    }
};

// ---------------------------------------------------------------------------
// TILED_TMA: the same transposing call as TILED, on the sm_100a copy engine.
// Every array operand has the same item size ESZ (2, 4 or 8 bytes).
//
//   tile     = kTI x kTO elements (I x O); a staged operand's tile is kTO rows
//              of 128 bytes (kTI elements along I), fetched by ONE TMA tensor
//              copy (5-D map: I, O, batch dims) with 128-byte swizzle.
//   block    = 8 consumer warps + 1 producer warp (288 threads), persistent
//              over tiles.  The producer warp lives entirely inside the
//              constructor: it keeps `tma_stages` tiles in flight through a
//              full/empty mbarrier ring, so the bytes in flight per SM are set
//              by the ring depth and not by occupancy or register count.
//   consumer = each lane reads kV 16-byte chunks (rows o..o+kV-1, one chunk of
//              kCH elements along I) with conflict-free LDS.128 and owns the
//              kCH x kV register block they form: the transpose itself is
//              register renaming, r[u][k] = chunk[k][u].  Compute, direct
//              operands and outputs then run along O with 128-bit accesses.
// ---------------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char b200_dyn_smem[];

template <int NARGS, int ESZ>
struct TmaTileTiler {
    static constexpr int kCH = 16 / ESZ;                  // elements per 16-byte chunk
    static constexpr int kV = kCH, kU = kCH;
    static constexpr int kTI = 128 / ESZ;                 // 8 chunks
    static constexpr int kLanesI = (kCH == 4) ? 4 : 8;    // lanes along I (chunks); see bank note below
    static constexpr int kLanesO = 32 / kLanesI;
    static constexpr int kWarpsI = 8 / kLanesI;
    static constexpr int kWarpsO = 8 / kWarpsI;
    static constexpr int kTO = kWarpsO * kLanesO * kV;    // 128 (4 B), 256 (2 B), 64 (8 B)
    static constexpr int kTileBytes = kTO * 128;
    static constexpr int kThreads = 288;
    static constexpr int kMaxStages = 8;
    // Bank note: a quarter-warp (8 lanes) must hit 8 distinct 16-byte bank groups.
    // Physical chunk = chunk ^ (row & 7).  With 8 lanes along I the row is shared
    // and the chunks differ; with 4 lanes along I (kV == 4) the two rows of a
    // quarter-warp differ by 4, which flips bit 2 of the XOR and separates them.

    const EwParams& p;
    const TileMaps& tm;
    int64_t n_i, n_o, i0, o0, lin0;
    int64_t boff[NARGS];
    uint32_t tiles_i, tiles_o, tiles, t;
    uint32_t bars, ring;          // shared-window addresses: barriers, first stage
    int stage, phase, last_staged;
    int li, lo, chunk;            // lane's element offsets inside the tile, chunk index
    bool live;

    B200_DEVICE uint32_t full_bar(int s) const { return bars + 8u * s; }
    B200_DEVICE uint32_t empty_bar(int s) const { return bars + 8u * (kMaxStages + s); }

    B200_DEVICE TmaTileTiler(const EwParams& p_, const TileMaps& tm_) : p(p_), tm(tm_) {
        const int ax_i = p.tile_axis, ax_o = p.ndim - 1;
        n_i = p.shape[ax_i];
        n_o = p.shape[ax_o];
        tiles_i = static_cast<uint32_t>((n_i + kTI - 1) / kTI);
        tiles_o = static_cast<uint32_t>((n_o + kTO - 1) / kTO);
        tiles = static_cast<uint32_t>((p.size / (n_i * n_o)) * tiles_i * tiles_o);
        const int nstaged = __popc(p.staged_mask);
        last_staged = 31 - __clz(p.staged_mask);
        const int S = p.tma_stages;
        bars = smem_u32(b200_dyn_smem);
        ring = (bars + 16u * kMaxStages + 1023u) & ~1023u;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (threadIdx.x == 0) {
            for (int s = 0; s < S; ++s) {
                mbar_init_a(full_bar(s), 1);
                mbar_init_a(empty_bar(s), 8);
            }
            mbar_fence_init();
        }
        __syncthreads();
        t = blockIdx.x;
        stage = 0;
        phase = 0;
        if (warp == 8) {
            // ---- producer: runs to completion here, then leaves the loop idle
            if (lane == 0) {
                for (int a = 0; a < NARGS; ++a)
                    if ((p.staged_mask >> a) & 1u) tma_prefetch_desc(&tm.m[__popc(p.staged_mask & ((1u << a) - 1u))]);
                uint32_t k = 0;
                for (; t < tiles; t += gridDim.x, ++k) {
                    if (k >= uint32_t(S)) mbar_wait_a(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx_a(full_bar(stage), uint32_t(nstaged) * kTileBytes);
                    uint32_t b = t;
                    int32_t c[5] = {0, 0, 0, 0, 0};
                    c[0] = int32_t((b % tiles_i) * kTI);
                    b /= tiles_i;
                    c[1] = int32_t((b % tiles_o) * kTO);
                    b /= tiles_o;
                    int n = 2;
#pragma unroll 1
                    for (int d = p.ndim - 2; d >= 0; --d) {
                        if (d == ax_i) continue;
                        const uint32_t s = static_cast<uint32_t>(p.shape[d]);
                        c[n++] = int32_t(b % s);
                        b /= s;
                    }
                    uint32_t dst = ring + uint32_t(stage) * uint32_t(nstaged) * kTileBytes;
                    for (int a = 0; a < NARGS; ++a) {
                        if (!((p.staged_mask >> a) & 1u)) continue;
                        const int slot = __popc(p.staged_mask & ((1u << a) - 1u));
                        tma_load_5d(dst, &tm.m[slot], c[0], c[1], c[2], c[3], c[4], full_bar(stage));
                        dst += kTileBytes;
                    }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
            live = false;
            return;
        }
        // ---- consumers
        const int wi = warp % kWarpsI, wo = warp / kWarpsI;
        chunk = wi * kLanesI + (lane % kLanesI);
        li = chunk * kCH;
        lo = (wo * kLanesO + lane / kLanesI) * kV;
        live = t < tiles;
        if (live) locate();
    }

    B200_DEVICE void locate() {
        const int ax_i = p.tile_axis;
        uint32_t b = t;
        i0 = int64_t(b % tiles_i) * kTI;
        b /= tiles_i;
        o0 = int64_t(b % tiles_o) * kTO;
        b /= tiles_o;
        lin0 = 0;
#pragma unroll
        for (int a = 0; a < NARGS; ++a) boff[a] = 0;
#pragma unroll 1
        for (int d = p.ndim - 2; d >= 0; --d) {
            if (d == ax_i) continue;
            const uint32_t s = static_cast<uint32_t>(p.shape[d]);
            const uint32_t r = b % s;
            b /= s;
            lin0 += int64_t(r) * p.cstride[d];
#pragma unroll
            for (int a = 0; a < NARGS; ++a) boff[a] += int64_t(r) * p.arg[a].strides[d];
        }
    }
    B200_DEVICE bool valid() const { return live; }
    B200_DEVICE void next() {
        t += gridDim.x;
        if (++stage == p.tma_stages) { stage = 0; phase ^= 1; }
        live = t < tiles;
        if (live) locate();
    }
    B200_DEVICE bool is_full() const { return i0 + kTI <= n_i && o0 + kTO <= n_o; }
    template <bool FULL>
    B200_DEVICE bool in_range(int u, int k) const { return FULL || ((i0 + li + u) < n_i && (o0 + lo + k) < n_o); }
    B200_DEVICE int64_t index(int u, int k) const {
        return lin0 + (i0 + li + u) * p.cstride[p.tile_axis] + (o0 + lo + k);
    }

    template <bool FULL, class T>
    B200_DEVICE void load(int a, Pack<T, kV> (&r)[kU]) const {
        static_assert(sizeof(T) == ESZ, "TILED_TMA: every array operand has the tile's item size");
        if ((p.staged_mask >> a) & 1u) {
            const int nstaged = __popc(p.staged_mask);
            const int slot = __popc(p.staged_mask & ((1u << a) - 1u));
            mbar_wait_a(full_bar(stage), uint32_t(phase));
            const uint32_t tb = ring + (uint32_t(stage) * uint32_t(nstaged) + uint32_t(slot)) * kTileBytes;
            uint4 q[kV];
#pragma unroll
            for (int k = 0; k < kV; ++k) q[k] = ld_shared_v4(tb + swz128(uint32_t(lo + k), uint32_t(chunk)));
#pragma unroll
            for (int u = 0; u < kU; ++u)
#pragma unroll
                for (int k = 0; k < kV; ++k) r[u][k] = reinterpret_cast<const T*>(&q[k])[u];
            if (a == last_staged) {
                __syncwarp();
                if ((threadIdx.x & 31) == 0) mbar_arrive_a(empty_bar(stage));
            }
        } else {
            const int64_t s_i = p.arg[a].strides[p.tile_axis];
            const int64_t s_o = p.arg[a].strides[p.ndim - 1];
            const char* base = p.arg[a].ptr + boff[a] + (i0 + li) * s_i + (o0 + lo) * s_o;
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (!FULL && !in_range<false>(u, 0)) continue;
                const char* q = base + u * s_i;
                if (s_o == int64_t(sizeof(T))) {
                    load_pack(r[u], reinterpret_cast<const T*>(q));
                } else if (s_o == 0) {
                    const T v = *reinterpret_cast<const T*>(q);
#pragma unroll
                    for (int k = 0; k < kV; ++k) r[u][k] = v;
                } else {
#pragma unroll
                    for (int k = 0; k < kV; ++k) r[u][k] = *reinterpret_cast<const T*>(q + k * s_o);
                }
            }
        }
    }
    template <bool FULL, class T>
    B200_DEVICE void store(int a, const Pack<T, kV> (&r)[kU]) const {
        const int64_t s_i = p.arg[a].strides[p.tile_axis];
        char* base = p.arg[a].ptr + boff[a] + (i0 + li) * s_i + (o0 + lo) * int64_t(sizeof(T));
#pragma unroll
        for (int u = 0; u < kU; ++u)
            if (FULL || in_range<false>(u, 0)) store_pack(reinterpret_cast<T*>(base + u * s_i), r[u]);
    }
};

// ---------------------------------------------------------------------------
// TILED_REG: register-block transpose, the default for transposing calls whose
// array operands share one item size ESZ (2, 4 or 8 bytes) and are 16-byte
// aligned.  No shared memory, no barriers:
//
//   lane     = a CH x CH block of elements (CH = 16 / ESZ).  A staged operand is
//              read as CH 16-byte vectors along I (rows o .. o+CH-1); direct
//              operands and outputs are CH vectors along O (rows i .. i+CH-1);
//              between the two the block is only renamed, r[a][k] = q[k][a].
//   warp     = 8 lanes along I x 4 lanes along O: every load instruction covers
//              4 rows of 128 contiguous bytes, every store 8 rows of 64 bytes
//              (whole 32-byte sectors on both sides).
//   thread   = UN such blocks (8 vector loads in flight), block = 8 warps; the
//              8*UN warp units of a tile are laid along O first (long output
//              rows), the rest along I.  Persistent grid-stride over tiles.
//
// Measured on config 4a (scripts/regblock_lab.cu, tma_read_lab.cu): 5.57 TB/s
// against 4.75 TB/s for the TMA ring (TILED_TMA), whose tensor copies are bound
// by address translation at ~16 B/clk/SM when every 128-byte box row lies in a
// different 2 MB page; plain LDG.128 on the same rows reads at 7.0 TB/s.
// ---------------------------------------------------------------------------
// blocks per thread: 8 vector loads of the staged operand in flight
B200_HD constexpr int reg_tile_unroll(int esz) { return esz == 4 ? 2 : esz == 2 ? 1 : 4; }
B200_HD int reg_tile_units_o(int64_t n_o, int unit_o, int units) {
    const int64_t need = (n_o + unit_o - 1) / unit_o;
    int u = 1;
    while (u < units && u < need) u <<= 1;
    return u;
}

// SPEC: three bits per operand fixing its access at compile time (the prebuilt unary
// table and every JIT kernel know the strides when the kernel is chosen):
//   bits 0-1: 1 = staged (unit-stride along I), 2 = direct, unit-stride along O,
//             3 = direct, broadcast or strided along O (scalar accesses);
//   bit 2   : direct operand with stride 0 along I -> one row serves the whole block.
template <int NARGS, int ESZ, int UN, uint64_t SPEC>
struct RegTileTiler {
    static constexpr int kCH = 16 / ESZ;
    static constexpr int kV = kCH, kU = kCH * UN;
    static constexpr int kLanesI = 8, kLanesO = 4;
    static constexpr int kUnitI = kLanesI * kCH;      // elements along I of one warp unit (128 bytes)
    static constexpr int kUnitO = kLanesO * kCH;      // elements along O of one warp unit (64 bytes)
    static constexpr int kUnits = 8 * UN;             // warp units per tile

    const EwParams& p;
    int64_t lin0;
    char* tbase[NARGS];           // operand pointers at the tile origin
    uint32_t tiles_i, tiles_o, tiles, t;
    int n_i, n_o, i0, o0, tile_i, tile_o;
    int li[UN], lo[UN];           // lane's element offsets inside the tile, per unit
    bool live;

    B200_DEVICE explicit RegTileTiler(const EwParams& p_) : p(p_) {
        n_i = int(p.shape[p.tile_axis]);
        n_o = int(p.shape[p.ndim - 1]);
        const int units_o = reg_tile_units_o(n_o, kUnitO, kUnits);
        tile_o = units_o * kUnitO;
        tile_i = (kUnits / units_o) * kUnitI;
        tiles_i = static_cast<uint32_t>((n_i + tile_i - 1) / tile_i);
        tiles_o = static_cast<uint32_t>((n_o + tile_o - 1) / tile_o);
        tiles = static_cast<uint32_t>((p.size / (int64_t(n_i) * n_o)) * tiles_i * tiles_o);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int un = 0; un < UN; ++un) {
            const int w = warp * UN + un;
            li[un] = (w / units_o) * kUnitI + (lane % kLanesI) * kCH;
            lo[un] = (w % units_o) * kUnitO + (lane / kLanesI) * kCH;
        }
        t = blockIdx.x;
        live = t < tiles;
        if (live) locate();
    }
    B200_DEVICE void locate() {
        const int ax_i = p.tile_axis, ax_o = p.ndim - 1;
        uint32_t b = t;
        i0 = int(b % tiles_i) * tile_i;
        b /= tiles_i;
        o0 = int(b % tiles_o) * tile_o;
        b /= tiles_o;
        lin0 = int64_t(i0) * p.cstride[ax_i] + o0;
#pragma unroll
        for (int a = 0; a < NARGS; ++a)
            tbase[a] = p.arg[a].ptr + int64_t(i0) * p.arg[a].strides[ax_i] + int64_t(o0) * p.arg[a].strides[ax_o];
#pragma unroll 1
        for (int d = p.ndim - 2; d >= 0; --d) {
            if (d == ax_i) continue;
            const uint32_t s = static_cast<uint32_t>(p.shape[d]);
            const uint32_t r = b % s;
            b /= s;
            lin0 += int64_t(r) * p.cstride[d];
#pragma unroll
            for (int a = 0; a < NARGS; ++a) tbase[a] += int64_t(r) * p.arg[a].strides[d];
        }
    }
    B200_DEVICE bool valid() const { return live; }
    B200_DEVICE void next() {
        t += gridDim.x;
        live = t < tiles;
        if (live) locate();
    }
    // One code path: the range test is one compare pair per 16-byte-square block, so a separate
    // unpredicated instantiation would only double the code and the register allocation.
    B200_DEVICE constexpr bool is_full() const { return false; }
    // n_i and n_o are multiples of CH (planner), so a lane's block is wholly inside or wholly outside
    B200_DEVICE bool unit_in(int un) const { return (i0 + li[un]) < n_i && (o0 + lo[un]) < n_o; }
    template <bool FULL>
    B200_DEVICE bool in_range(int u, int) const { return FULL || unit_in(u / kCH); }
    B200_DEVICE int64_t index(int u, int k) const {
        return lin0 + int64_t(li[u / kCH] + (u % kCH)) * p.cstride[p.tile_axis] + (lo[u / kCH] + k);
    }

    template <bool FULL, class T>
    B200_DEVICE void load(int a, Pack<T, kV> (&r)[kU]) const {
        static_assert(sizeof(T) == ESZ, "TILED_REG: every array operand has the tile's item size");
        const uint32_t kind = uint32_t(SPEC >> (3 * a)) & 3u;
        const bool bcast_i = (SPEC >> (3 * a + 2)) & 1u;
        const int64_t s_o = (kind == 2) ? int64_t(sizeof(T)) : p.arg[a].strides[p.ndim - 1];
        if (kind == 1) {
            // unit-stride along I: CH vectors along I, one per row o .. o+CH-1
#pragma unroll
            for (int un = 0; un < UN; ++un) {
                if (!FULL && !unit_in(un)) continue;
                const char* b0 = tbase[a] + li[un] * int(sizeof(T)) + lo[un] * s_o;
                Pack<T, kCH> q[kCH];
#pragma unroll
                for (int k = 0; k < kCH; ++k) load_pack(q[k], reinterpret_cast<const T*>(b0 + k * s_o));
#pragma unroll
                for (int x = 0; x < kCH; ++x)
#pragma unroll
                    for (int k = 0; k < kCH; ++k) r[un * kCH + x][k] = q[k][x];
            }
        } else {
            const int64_t s_i = bcast_i ? 0 : p.arg[a].strides[p.tile_axis];
#pragma unroll
            for (int un = 0; un < UN; ++un) {
                if (!FULL && !unit_in(un)) continue;
                const char* b0 = tbase[a] + li[un] * s_i + lo[un] * s_o;
#pragma unroll
                for (int x = 0; x < (bcast_i ? 1 : kCH); ++x) {
                    const char* q = b0 + x * s_i;
                    if (kind == 2) {
                        load_pack(r[un * kCH + x], reinterpret_cast<const T*>(q));
                    } else if (s_o == 0) {
                        const T v = *reinterpret_cast<const T*>(q);
#pragma unroll
                        for (int k = 0; k < kV; ++k) r[un * kCH + x][k] = v;
                    } else {
#pragma unroll
                        for (int k = 0; k < kV; ++k) r[un * kCH + x][k] = *reinterpret_cast<const T*>(q + k * s_o);
                    }
                }
                if (bcast_i) {
#pragma unroll
                    for (int x = 1; x < kCH; ++x) r[un * kCH + x] = r[un * kCH];
                }
            }
        }
    }
    template <bool FULL, class T>
    B200_DEVICE void store(int a, const Pack<T, kV> (&r)[kU]) const {
        const int64_t s_i = p.arg[a].strides[p.tile_axis];
#pragma unroll
        for (int un = 0; un < UN; ++un) {
            if (!FULL && !unit_in(un)) continue;
            char* b0 = tbase[a] + li[un] * s_i + lo[un] * int(sizeof(T));
#pragma unroll
            for (int x = 0; x < kCH; ++x) store_pack(reinterpret_cast<T*>(b0 + x * s_i), r[un * kCH + x]);
        }
    }
};

}  // namespace b200
