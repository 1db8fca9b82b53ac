// b200/scan_narrow.cuh -- cumsum / cumprod along axis 0 of a tall, narrow matrix x[n][cols] (cols <= 128: point
// clouds, feature tables), where the strip kernels (scan_march.cuh, scan_cols_kernel) find no 128-byte strip to
// give a warp and fall back to one element per thread: every 4-byte load drags a 32-byte sector, ~100 GB/s.
//
// Here the matrix is read as the flat stream it is.  `active` threads per block (<= 256, chosen by the host so that
// active * VEC is a multiple of cols) load 16-byte vectors of consecutive elements: a CHUNK is active * VEC elements
// = chunk_rows whole rows, and a TILE is U chunks staged in shared memory as accumulators.  Inside a tile
//   P1  coalesced vector loads -> shared memory (converted to the accumulator type),
//   P2  thread (column c, group g) scans rows [g * L, (g + 1) * L) of column c in place (stride-cols walk),
//   P3  one warp per column (a thread per column when there are at most 32 groups) scans the G group totals, adds
//       the running carry of the column, leaves each group's offset behind and advances the carry,
//   P4  every thread re-reads ITS OWN flat elements, adds the offset of the group they fell in (indices that are
//       the same for every tile: kept in registers) and stores 16..64-byte vectors, coalesced.
// Across blocks it is reduce-then-scan, because that keeps floating-point results independent of timing: the
// matrix is cut into one contiguous segment of whole tiles per block; kernel 1 leaves every segment's column totals
// in the workspace, kernel 2 scans the segments, each block first folding the totals of the segments before its own
// (fixed order).  12 instead of 8 bytes of traffic per 4-byte element, all of it in full sectors.
#pragma once
#include "scan.cuh"
#include "scan_pipe.cuh"      // PipeCvt

namespace b200 {

constexpr int kNarrowScanThreads = 256;
constexpr int kNarrowScanMaxCols = 128;

template <class In, class Acc, class Out>
struct ScanNarrowCfg {
    static constexpr int T = kNarrowScanThreads;
    static constexpr int VEC = (16 / int(sizeof(In))) > 8 ? 8 : (16 / int(sizeof(In)));     // elements per load
    static constexpr int ENTRIES = 16384 / int(sizeof(Acc));                               // accumulators per tile
    static constexpr int U = ENTRIES / (T * VEC) < 1 ? 1 : ENTRIES / (T * VEC);            // chunks per tile
    static constexpr int SLOTS = T * VEC * U;
};

B200_DEVICE int narrow_pad(int i) { return i + (i >> 5); }       // one spare word per 32: column walks spread over banks

// Column totals of `count` row-major rows of `cols` accumulators in shared memory -> `out[c]`, c < cols.
// Threads are (column, group); group g folds rows g, g + G, ...; thread c then folds the groups.  Fixed order.
template <class Acc, class Op, int T>
B200_DEVICE void narrow_fold_columns(const Acc* rows, int count, int cols, Acc* groups, Acc* out) {
    const int t = threadIdx.x, G = T / cols, c = t % cols, g = t / cols;
    if (g < G) {
        Acc a = Op::template identity<Acc>();
        for (int r = g; r < count; r += G) a = Op::combine(a, rows[r * cols + c]);
        groups[g * cols + c] = a;
    }
    __syncthreads();
    if (t < cols) {
        Acc a = groups[t];
        for (int w = 1; w < G; ++w) a = Op::combine(a, groups[w * cols + t]);
        out[t] = a;
    }
}

// ---- kernel 1: column totals of every segment ------------------------------------------------------------------
template <class In, class Acc, class Out, class Op>
__device__ __forceinline__ void scan_narrow_totals_body(const In* __restrict__ x, int64_t n, int cols, int active,
                                                        int64_t seg_tiles, Acc* __restrict__ tot) {
    typedef ScanNarrowCfg<In, Acc, Out> Cfg;
    constexpr int T = Cfg::T, VEC = Cfg::VEC, U = Cfg::U;
    __shared__ Acc lanes[T * VEC];
    __shared__ Acc groups[T];
    const int t = threadIdx.x;
    const int64_t chunk_elems = int64_t(active) * VEC;
    const int64_t total = n * cols;
    const int64_t e_begin = int64_t(blockIdx.x) * seg_tiles * U * chunk_elems;
    int64_t e_end = e_begin + seg_tiles * U * chunk_elems;
    if (e_end > total) e_end = total;
    Acc acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = Op::template identity<Acc>();
    if (t < active) {
        int64_t e = e_begin + int64_t(t) * VEC;
        for (; e + 3 * chunk_elems + VEC <= e_end; e += 4 * chunk_elems) {       // four loads in flight
            Pack<In, VEC> v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) load_pack(v[u], x + e + u * chunk_elems);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[k] = Op::combine(acc[k], PipeCvt<In, Acc>::in(v[u][k]));
        }
        for (; e < e_end; e += chunk_elems) {
            if (e + VEC <= e_end) {
                Pack<In, VEC> v;
                load_pack(v, x + e);
#pragma unroll
                for (int k = 0; k < VEC; ++k) acc[k] = Op::combine(acc[k], PipeCvt<In, Acc>::in(v[k]));
            } else {
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    if (e + k < e_end) acc[k] = Op::combine(acc[k], PipeCvt<In, Acc>::in(x[e + k]));
            }
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) lanes[t * VEC + k] = acc[k];
    }
    __syncthreads();
    narrow_fold_columns<Acc, Op, T>(lanes, active * VEC / cols, cols, groups, tot + int64_t(blockIdx.x) * cols);
}

// ---- kernel 2: the scan of every segment, starting from the totals of the segments before it -------------------
template <class In, class Acc, class Out, class Op>
__device__ __forceinline__ void scan_narrow_body(const In* __restrict__ x, Out* __restrict__ y, int64_t n, int cols,
                                                 int active, int64_t seg_tiles, const Acc* __restrict__ tot) {
    typedef ScanNarrowCfg<In, Acc, Out> Cfg;
    constexpr int T = Cfg::T, VEC = Cfg::VEC, U = Cfg::U, SLOTS = Cfg::SLOTS;
    __shared__ __align__(16) Acc tile[SLOTS + SLOTS / 32 + 1];
    __shared__ Acc groups[T];          // P2: group totals; P3: group offsets (carry included)
    __shared__ Acc carry[kNarrowScanMaxCols];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int chunk_rows = active * VEC / cols;
    const int chunk_elems = active * VEC;
    const int tile_elems = chunk_elems * U, tile_rows = chunk_rows * U;
    const int64_t total = n * cols;
    const int G = T / cols, c = t % cols, g = t / cols;
    const int L = (tile_rows + G - 1) / G;             // rows per group
    const int J = (G + 31) / 32;                       // groups per lane in P3

    // exclusive prefix of this block's segment: the totals of the segments before it, folded in a fixed order
    {
        const int before = int(blockIdx.x);
        if (g < G) {
            Acc a = Op::template identity<Acc>();
            int s = g;
            for (; s + 3 * G < before; s += 4 * G) {
                Acc v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = tot[int64_t(s + k * G) * cols + c];
#pragma unroll
                for (int k = 0; k < 4; ++k) a = Op::combine(a, v[k]);
            }
            for (; s < before; s += G) a = Op::combine(a, tot[int64_t(s) * cols + c]);
            groups[g * cols + c] = a;
        }
        __syncthreads();
        if (t < cols) {
            Acc a = groups[t];
            for (int w = 1; w < G; ++w) a = Op::combine(a, groups[w * cols + t]);
            carry[t] = a;
        }
        __syncthreads();
    }

    // where each of this thread's elements sits in a tile: flat slot and the slot of its group's offset
    int off_slot[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            const int f = u * chunk_elems + t * VEC + k;
            off_slot[u][k] = ((f / cols) / L) * cols + f % cols;
        }

    const int64_t tile_first = int64_t(blockIdx.x) * seg_tiles;
    // the loads of tile i + 1 are issued before tile i is scanned and wait in registers: a block alone keeps
    // U vectors per thread in flight through all four phases (three or four blocks fit an SM, each of which would
    // otherwise spend two thirds of its time with nothing outstanding)
    Pack<In, VEC> pre[U];
    auto fetch = [&](int64_t tl) {
        const int64_t e0 = tl * tile_elems;
        if (t < active && e0 < total) {
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
    fetch(tile_first);
    for (int64_t tl = tile_first; tl < tile_first + seg_tiles; ++tl) {
        const int64_t e0 = tl * tile_elems;
        if (e0 >= total) break;                          // block-uniform
        const int64_t rows_left = n - tl * tile_rows;
        const int rows_here = rows_left < tile_rows ? int(rows_left) : tile_rows;
        // P1
        if (t < active) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int f = u * chunk_elems + t * VEC;
                const int64_t ge = e0 + f;
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    tile[narrow_pad(f + k)] = ge + k < total ? PipeCvt<In, Acc>::in(pre[u][k]) : Op::template identity<Acc>();
            }
        }
        if (tl + 1 < tile_first + seg_tiles) fetch(tl + 1);
        __syncthreads();
        // P2
        if (g < G) {
            Acc a = Op::template identity<Acc>();
            const int r0 = g * L;
            const int r1 = r0 + L < rows_here ? r0 + L : rows_here;
            for (int r = r0; r < r1; ++r) {
                const int i = narrow_pad(r * cols + c);
                a = Op::combine(a, tile[i]);
                tile[i] = a;
            }
            groups[g * cols + c] = a;
        }
        __syncthreads();
        // P3
        if (G <= 32) {                                   // few groups per column: a thread per column walks them
            if (t < cols) {
                Acc run = carry[t];
                for (int gi = 0; gi < G; ++gi) {
                    const Acc v = groups[gi * cols + t];
                    groups[gi * cols + t] = run;
                    run = Op::combine(run, v);
                }
                carry[t] = run;
            }
        } else
        for (int cc = warp; cc < cols; cc += T / 32) {
            Acc v[8];                                     // J <= 8: G <= 256
            Acc loc = Op::template identity<Acc>();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int gi = lane * J + j;
                v[j] = (j < J && gi < G) ? groups[gi * cols + cc] : Op::template identity<Acc>();
                loc = Op::combine(loc, v[j]);
            }
            Acc incl = loc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const Acc o = shfl_up_any(incl, d);
                if (lane >= d) incl = Op::combine(o, incl);
            }
            Acc excl = shfl_up_any(incl, 1);
            if (lane == 0) excl = Op::template identity<Acc>();
            const Acc agg = shfl_any(incl, 31);
            const Acc base = carry[cc];
            Acc run = Op::combine(base, excl);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int gi = lane * J + j;
                if (j < J && gi < G) groups[gi * cols + cc] = run;
                run = Op::combine(run, v[j]);
            }
            __syncwarp();
            if (lane == 0) carry[cc] = Op::combine(base, agg);
        }
        __syncthreads();
        // P4
        if (t < active) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int f = u * chunk_elems + t * VEC;
                const int64_t ge = e0 + f;
                Pack<Out, VEC> o;
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    o[k] = static_cast<Out>(Op::combine(groups[off_slot[u][k]], tile[narrow_pad(f + k)]));
                if (ge + VEC <= total) {
                    store_pack(y + ge, o);
                } else {
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        if (ge + k < total) y[ge + k] = o[k];
                }
            }
        }
        __syncthreads();
    }
}

// ---- the same tiles for scans ALONG short rows (x[rows][len], len <= 64, scanning len) ------------------------------
// A warp per row (scan_lines_kernel) leaves most lanes idle below 128 elements per row.  Here a tile is a whole number
// of rows staged in shared memory with coalesced vector loads; a thread then walks whole rows (the padded layout keeps
// the walks of a warp on different banks), and the tile goes back out coalesced.  Rows never cross a tile, so there
// is no carry and no second pass: 8 bytes of traffic per 4-byte element, one launch.
template <class In, class Acc, class Out, class Op>
__device__ __forceinline__ void scan_short_rows_body(const In* __restrict__ x, Out* __restrict__ y, int64_t rows, int len,
                                                     int active) {
    typedef ScanNarrowCfg<In, Acc, Out> Cfg;
    constexpr int T = Cfg::T, VEC = Cfg::VEC, U = Cfg::U, SLOTS = Cfg::SLOTS;
    __shared__ __align__(16) Acc tile[SLOTS + SLOTS / 32 + 1];
    const int t = threadIdx.x;
    const int chunk_elems = active * VEC;
    const int tile_elems = chunk_elems * U, tile_rows = tile_elems / len;
    const int64_t total = rows * len;
    const int64_t tiles = (total + tile_elems - 1) / tile_elems;
    Pack<In, VEC> pre[U];
    auto fetch = [&](int64_t tl) {
        const int64_t e0 = tl * tile_elems;
        if (t < active && tl < tiles) {
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
        const int64_t e0 = tl * tile_elems;
        if (t < active) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int f = u * chunk_elems + t * VEC;
                const int64_t ge = e0 + f;
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    tile[narrow_pad(f + k)] = ge + k < total ? PipeCvt<In, Acc>::in(pre[u][k]) : Op::template identity<Acc>();
            }
        }
        fetch(tl + gridDim.x);
        __syncthreads();
        for (int r = t; r < tile_rows; r += T) {
            Acc a = Op::template identity<Acc>();
            const int base = r * len;
            for (int j = 0; j < len; ++j) {
                const int i = narrow_pad(base + j);
                a = Op::combine(a, tile[i]);
                tile[i] = a;
            }
        }
        __syncthreads();
        if (t < active) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int f = u * chunk_elems + t * VEC;
                const int64_t ge = e0 + f;
                Pack<Out, VEC> o;
#pragma unroll
                for (int k = 0; k < VEC; ++k) o[k] = static_cast<Out>(tile[narrow_pad(f + k)]);
                if (ge + VEC <= total) {
                    store_pack(y + ge, o);
                } else {
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        if (ge + k < total) y[ge + k] = o[k];
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace b200
