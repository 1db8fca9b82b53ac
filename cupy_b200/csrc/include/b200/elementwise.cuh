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
    int64_t  shape[kMaxNdim];
    int64_t  cstride[kMaxNdim];    // C-order element strides of `shape` (for the linear index `i`)
    FastDiv  fdiv[kMaxNdim];       // fast division by shape[d] (valid when size < 2^31)
    EwArg    arg[kMaxArgs];
};

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
template <int NARGS, int VEC, int UNROLL, int THREADS>
struct FlatTiler {
    static constexpr int kV = VEC, kU = UNROLL;
    static constexpr int64_t kTile = int64_t(THREADS) * VEC * UNROLL;
    const EwParams& p;
    int64_t base;
    bool full;

    B200_DEVICE explicit FlatTiler(const EwParams& p_) : p(p_) {
        base = int64_t(blockIdx.x) * kTile;
        full = base + kTile <= p.size;
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
template <int NARGS, int VEC, int UNROLL, int THREADS, bool IDX32>
struct RowTiler {
    static constexpr int kV = VEC, kU = UNROLL;
    const EwParams& p;
    int64_t inner, chunks, total, wbase;
    int64_t off[UNROLL][NARGS];
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
                uint32_t rest = static_cast<uint32_t>(w), c;
                // innermost: chunk index
                {
                    uint32_t q = rest / static_cast<uint32_t>(chunks);
                    c = rest - q * static_cast<uint32_t>(chunks);
                    rest = q;
                }
                const int64_t e0 = int64_t(c) * VEC;
                lin[u] = e0;
#pragma unroll
                for (int a = 0; a < NARGS; ++a) off[u][a] = e0 * p.arg[a].strides[p.ndim - 1];
#pragma unroll 1
                for (int d = p.ndim - 2; d >= 0; --d) {
                    uint32_t q, r;
                    p.fdiv[d].divmod(rest, q, r);
                    rest = q;
                    lin[u] += int64_t(r) * p.cstride[d];
#pragma unroll
                    for (int a = 0; a < NARGS; ++a) off[u][a] += int64_t(r) * p.arg[a].strides[d];
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
        const int64_t si = p.arg[a].strides[p.ndim - 1];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (!FULL && !ok[u]) continue;
            const char* base = p.arg[a].ptr + off[u][a];
            if (VEC > 1 && si == int64_t(sizeof(T))) {
                load_pack(r[u], reinterpret_cast<const T*>(base));
            } else if (VEC > 1 && si == 0) {
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
        const int64_t si = p.arg[a].strides[p.ndim - 1];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (!FULL && !ok[u]) continue;
            char* base = p.arg[a].ptr + off[u][a];
            if (VEC > 1 && si == int64_t(sizeof(T))) {
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

}  // namespace b200
