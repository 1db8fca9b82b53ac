// b200/scan_tma.cuh -- the sm_100a scan: persistent blocks, TMA-pipelined tiles,
// fence-free decoupled look-back.
//
//   * The array is viewed as rows of 128 bytes.  A tile is THREADS rows; one TMA
//     tensor copy (128-byte swizzle) brings it into a shared-memory stage, one TMA
//     tensor store writes the scanned tile back from the same stage.  STAGES
//     tiles per block are in flight, so HBM stays busy while a tile waits for
//     its prefix -- the latency that a load/compute/look-back/store block of the
//     classic design exposes.
//   * Thread t owns row t: 128 / sizeof(T) consecutive items, read from the
//     swizzled stage with conflict-free 128-bit accesses straight into registers.
//   * Tiles are dealt round-robin to a co-resident (cooperatively launched) grid of G
//     blocks: block b owns tiles b, b+G, ...; the tiles of iteration k form "wave" k.
//   * No look-back chain.  Every block publishes its tile aggregate, and every block
//     reads ALL aggregates of a wave (G slots, a few KB from L2) to get both its own
//     exclusive prefix inside the wave and the wave total, which it adds to a running
//     prefix kept in registers.  Nothing a block waits for depends on another
//     block's wait, so there is no serial hand-off latency per wave, and the sums
//     are taken in a fixed order (float scans are run-to-run deterministic).
//   * Software pipeline inside a block: iteration k scans tile k locally (phase A:
//     publish aggregate, park the locally scanned row back in its stage) and finishes
//     tile k-LAG (phase B: gather wave k-LAG, add prefix, TMA store).  The aggregates
//     consumed in phase B were published LAG iterations earlier: the gather finds
//     them ready, and a block may run up to LAG waves ahead of the slowest one.
//   * Slots carry status and value in words that are written atomically ({status, 32
//     value bits} per 64-bit word), so no fences are needed: a reader that sees
//     matching statuses has the value.
//
// Same results as scan.cuh (integers: bit-exact).  T is the element type of input,
// accumulator and output (4 or 8 bytes); other dtype combinations use scan.cuh.
// Replaces cub::DeviceScan as called from cupy/cuda/cupy_cub.cu:991-1013.
#pragma once
#include "scan.cuh"
#include "tma.cuh"

namespace b200 {

// ---- look-back slots ---------------------------------------------------------
// status: 0 = empty, 1 = tile aggregate, 2 = inclusive prefix
template <int BYTES> struct LookbackSlot;

template <> struct LookbackSlot<4> {
    typedef uint64_t storage_t;
    B200_DEVICE static void publish(storage_t* p, uint32_t status, uint32_t bits) {
        const uint64_t w = (uint64_t(status) << 32) | bits;
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    }
    // returns status (0 = not ready)
    B200_DEVICE static uint32_t peek(const storage_t* p, uint32_t& bits) {
        uint64_t w;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
        bits = static_cast<uint32_t>(w);
        return static_cast<uint32_t>(w >> 32);
    }
};

template <> struct LookbackSlot<8> {
    struct __align__(16) storage_t { uint64_t w0, w1; };
    B200_DEVICE static void publish(storage_t* p, uint32_t status, uint64_t bits) {
        const uint64_t w0 = (uint64_t(status) << 32) | (bits & 0xffffffffull);
        const uint64_t w1 = (uint64_t(status) << 32) | (bits >> 32);
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
    }
    B200_DEVICE static uint32_t peek(const storage_t* p, uint64_t& bits) {
        uint64_t w0, w1;
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
        const uint32_t s0 = static_cast<uint32_t>(w0 >> 32), s1 = static_cast<uint32_t>(w1 >> 32);
        bits = (w0 & 0xffffffffull) | (w1 << 32);
        return s0 == s1 ? s0 : 0u;     // torn read (halves of different generations): retry
    }
};

template <class T> struct bits_of;
template <> struct bits_of<float> { typedef uint32_t type; };
template <> struct bits_of<int32_t> { typedef uint32_t type; };
template <> struct bits_of<uint32_t> { typedef uint32_t type; };
template <> struct bits_of<double> { typedef uint64_t type; };
template <> struct bits_of<long long> { typedef uint64_t type; };
template <> struct bits_of<unsigned long long> { typedef uint64_t type; };

template <class T>
B200_DEVICE typename bits_of<T>::type to_bits(const T& v) {
    union { T t; typename bits_of<T>::type b; } u;
    u.t = v;
    return u.b;
}
template <class T>
B200_DEVICE T from_bits(typename bits_of<T>::type b) {
    union { T t; typename bits_of<T>::type b; } u;
    u.b = b;
    return u.t;
}

template <int THREADS, int STAGES>
struct ScanTmaSmem {
    static constexpr int kStageBytes = THREADS * 128;
    // [stages (1024-aligned)] [mbarriers | warp totals | gather partials]
    static constexpr int kBytes = 1024 + STAGES * kStageBytes + 512;
};

// DBG (lab builds only, scripts/scan_lab.cu): 1 = skip the cross-block gather, 2 = pure TMA copy
template <class T, class Op, int THREADS, int STAGES, int LAG = 1, int DBG = 0>
__device__ __forceinline__ void scan_tma_body(const void* tm_in, const void* tm_out,
                                              const T* __restrict__ x, T* __restrict__ y, int64_t n,
                                              typename LookbackSlot<sizeof(T)>::storage_t* slots) {
    typedef LookbackSlot<sizeof(T)> Slot;
    typedef typename bits_of<T>::type bits_t;
    static_assert(STAGES >= LAG + 2, "tiles it-LAG..it hold a stage each; the rest are loads in flight");
    constexpr int ITEMS = 128 / int(sizeof(T));      // items per row = per thread
    constexpr int EPC = 16 / int(sizeof(T));         // items per 16-byte chunk
    constexpr int NWARPS = THREADS / 32;
    constexpr int kStageBytes = THREADS * 128;
    constexpr int kMaxGather = 4;                    // slots per thread per wave: G <= 4 * THREADS

    extern __shared__ uint8_t smem_raw[];
    uint8_t* stage0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(stage0 + STAGES * kStageBytes);      // [STAGES] <= 64 B
    T* warp_total = reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(full) + 64);      // [NWARPS] <= 64 B
    T* g_before = reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(full) + 128);       // [NWARPS]
    T* g_all = reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(full) + 192);          // [NWARPS]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t full_rows = n / ITEMS;                        // rows the tensor maps cover
    const int64_t rows = (n + ITEMS - 1) / ITEMS;
    const int64_t tiles = (rows + THREADS - 1) / THREADS;
    const int64_t G = gridDim.x;
    const int64_t bid = blockIdx.x;
    const int my_tiles = bid < tiles ? int((tiles - bid + G - 1) / G) : 0;
    const bool dma = (tid == THREADS - 32);          // issues TMA; lane 0 of the last warp

    auto request = [&](int64_t k) {                  // my k-th tile -> stage k % STAGES
        if (k < my_tiles) {
            const int s = int(k % STAGES);
            mbar_expect_tx(full + s, kStageBytes);
            tma_load_2d(stage0 + s * kStageBytes, tm_in, 0, int32_t((k * G + bid) * THREADS), full + s);
        }
    };

    if (dma) {
        tma_prefetch_desc(tm_in);
        tma_prefetch_desc(tm_out);
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(full + s, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (dma) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) request(s);
    }

    T running = Op::template identity<T>();          // inclusive prefix of all waves before the one in phase B

    for (int it = 0; it < my_tiles + LAG; ++it) {
        // ---- issue the gather of wave it-LAG now; its L2 latency hides behind phase A
        uint32_t pf_status[kMaxGather];
        bits_t pf_bits[kMaxGather];
        const int64_t wave0 = int64_t(it - LAG) * G;
        const int count = (it >= LAG && DBG == 0) ? int((tiles - wave0) < G ? (tiles - wave0) : G) : 0;
#pragma unroll
        for (int r = 0; r < kMaxGather; ++r) {
            pf_status[r] = 1u;
            pf_bits[r] = to_bits(Op::template identity<T>());
        }
#pragma unroll
        for (int r = 0; r < kMaxGather; ++r) {
            if (r * THREADS >= count) break;          // block-uniform: G <= THREADS needs one round only
            const int bp = tid + r * THREADS;
            if (bp < count) pf_status[r] = Slot::peek(slots + wave0 + bp, pf_bits[r]);
        }
        // ================= phase A: local scan of my tile `it` =================
        const bool do_a = (it < my_tiles) && (DBG != 2);
        uint8_t* const st_a = stage0 + (it % STAGES) * kStageBytes;
        const int64_t tile_a = int64_t(it) * G + bid;
        T item[ITEMS];
        T lane_excl = Op::template identity<T>();
        if (it < my_tiles) mbar_wait(full + (it % STAGES), (it / STAGES) & 1);
        if (do_a) {
            const int64_t row = tile_a * THREADS + tid;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const Pack<T, EPC> v = *reinterpret_cast<const Pack<T, EPC>*>(st_a + swz128(tid, c));
#pragma unroll
                for (int k = 0; k < EPC; ++k) item[c * EPC + k] = v[k];
            }
            // the ragged last row lies outside the tensor map (TMA zero-filled it): one thread of the
            // whole grid reads it directly -- through its stage, so the hot path carries no predicates
            if (__builtin_expect(row == full_rows && full_rows != rows, 0)) {
                const int cnt = int(n - full_rows * ITEMS);
                T* srow = reinterpret_cast<T*>(st_a);
#pragma unroll 1
                for (int j = 0; j < ITEMS; ++j) {
                    const uint32_t off = swz128(tid, j / EPC) + (j % EPC) * sizeof(T);
                    *reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(srow) + off) =
                        (j < cnt) ? x[row * ITEMS + j] : Op::template identity<T>();
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const Pack<T, EPC> v = *reinterpret_cast<const Pack<T, EPC>*>(st_a + swz128(tid, c));
#pragma unroll
                    for (int k = 0; k < EPC; ++k) item[c * EPC + k] = v[k];
                }
            }
#pragma unroll
            for (int j = 1; j < ITEMS; ++j) item[j] = Op::combine(item[j - 1], item[j]);
            T incl = item[ITEMS - 1];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                T t = shfl_up_any(incl, d);
                if (lane >= d) incl = Op::combine(t, incl);
            }
            lane_excl = shfl_up_any(incl, 1);
            if (lane == 0) lane_excl = Op::template identity<T>();
            if (lane == 31) warp_total[warp] = incl;
        }
        // ---- wave it-LAG: all aggregates (wave total) and those before me (my offset); the
        // warp partials ride on barrier (A)
        if (count > 0) {
            T before = Op::template identity<T>(), all = Op::template identity<T>();
#pragma unroll
            for (int r = 0; r < kMaxGather; ++r) {
                if (r * THREADS >= count) break;
                const int bp = tid + r * THREADS;
                if (bp < count) {
                    while (pf_status[r] == 0u) pf_status[r] = Slot::peek(slots + wave0 + bp, pf_bits[r]);
                    const T v = from_bits<T>(pf_bits[r]);
                    all = Op::combine(all, v);
                    if (bp < bid) before = Op::combine(before, v);
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                before = Op::combine(before, shfl_down_any(before, d));
                all = Op::combine(all, shfl_down_any(all, d));
            }
            if (lane == 0) { g_before[warp] = before; g_all[warp] = all; }
        }
        __syncthreads();                                                         // (A)
        if (do_a) {
            T warp_excl = Op::template identity<T>();
            T block_agg = Op::template identity<T>();
#pragma unroll
            for (int w = 0; w < NWARPS; ++w) {
                const T t = warp_total[w];
                if (w < warp) warp_excl = Op::combine(warp_excl, t);
                block_agg = Op::combine(block_agg, t);
            }
            if (tid == 0 && DBG == 0) Slot::publish(slots + tile_a, 1u, to_bits(block_agg));
            // park the tile-locally scanned row in its stage until phase B
            const T pre = Op::combine(warp_excl, lane_excl);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                Pack<T, EPC> v;
#pragma unroll
                for (int k = 0; k < EPC; ++k) v[k] = Op::combine(pre, item[c * EPC + k]);
                *reinterpret_cast<Pack<T, EPC>*>(st_a + swz128(tid, c)) = v;
            }
        }
        // ================= phase B: finish my tile `it - LAG` =================
        if (it >= LAG) {
            const int k = it - LAG;
            const int s = k % STAGES;
            uint8_t* st = stage0 + s * kStageBytes;
            const int64_t tile = int64_t(k) * G + bid;
            T prefix = running;
            if (DBG == 0) {
                T sb = Op::template identity<T>(), sa = Op::template identity<T>();
#pragma unroll
                for (int w = 0; w < NWARPS; ++w) {
                    sb = Op::combine(sb, g_before[w]);
                    sa = Op::combine(sa, g_all[w]);
                }
                prefix = Op::combine(running, sb);
                running = Op::combine(running, sa);
            }
            if (DBG != 2) {
                const int64_t row = tile * THREADS + tid;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    Pack<T, EPC>* p = reinterpret_cast<Pack<T, EPC>*>(st + swz128(tid, c));
                    Pack<T, EPC> v = *p;
#pragma unroll
                    for (int e = 0; e < EPC; ++e) v[e] = Op::combine(prefix, v[e]);
                    *p = v;
                }
                // the ragged last row is clipped by the TMA store: its owner writes it directly
                if (__builtin_expect((row == full_rows) && (full_rows != rows), 0)) {
                    const int cnt = int(n - full_rows * ITEMS);
#pragma unroll 1
                    for (int j = 0; j < cnt; ++j) {
                        const uint32_t off = swz128(tid, j / EPC) + (j % EPC) * sizeof(T);
                        y[row * ITEMS + j] = *reinterpret_cast<const T*>(st + off);
                    }
                }
            }
            fence_proxy_async_smem();
            __syncthreads();                                                     // (C)
            if (dma) {
                tma_store_2d(tm_out, 0, int32_t(tile * THREADS), st);     // rows past the end are clipped
                tma_commit_group();
                // refill the stage whose store was issued one iteration ago, once that store has
                // finished READING the stage
                if (k >= 1) {
                    tma_wait_group_read<1>();
                    request(int64_t(k) - 1 + STAGES);
                }
            }
        }
    }
    if (dma) tma_wait_group<0>();    // stores must complete before the block retires its smem
}

}  // namespace b200
