// b200/scan_pipe.cuh -- the sm_100a flat scan (cumsum / cumprod), second generation.
//
// One persistent 512-thread block per SM.  HBM <-> shared memory traffic is all TMA
// (128-byte rows, 128-byte swizzle, one tensor copy per tile and direction); threads only
// touch shared memory and registers:
//
//   in-ring (SI stages)  --LDS-->  registers: the thread's IPT RAW input items, held for LAG
//   iterations; only their sum enters the warp / block scan  --(the tile's exclusive prefix
//   becomes known)-->  running prefix + thread offset, then the items are converted, scanned
//   serially and  --STS-->  out-ring (SO stages)  --TMA store--> y
//
//   * every element crosses shared memory exactly twice (one LDS, one STS); round 1's design
//     parked the locally scanned tile in its stage and re-read it (four crossings), at 8 warps
//     per SM with a 16-item dependent chain per thread: latency-bound at 74 % of the copy peak.
//     Here 16 warps hide the chains, an input stage is free for its next TMA load as soon as it
//     has been read, and what waits in registers is the INPUT (16 words per thread and tile for
//     every dtype pair), so casting scans hold 16-32 items per thread without spilling.
//   * thread t owns the bytes [t * IPT * sizeof(In), ...) of the input tile and [t * IPT *
//     sizeof(Out), ...) of the output tile: consecutive threads read consecutive 16-byte chunks,
//     which the 128-byte swizzle spreads over all 32 banks for every chunk count per thread
//     (1, 2, 4 or 8), so In and Out may differ in size: int32 -> int64, bool -> int64, float16
//     with a float accumulator run on the same kernel -- one read of x, no `astype` pre-pass
//     (cupy/_core/_routines_math.pyx:726-727) and no second read (round 1's line-totals scheme).
//   * cross-block: tiles are dealt round-robin to the G co-resident blocks; the tiles of
//     iteration k form wave k.  Every block publishes its tile aggregate into a small ring of
//     tagged slots and, LAG iterations later, reads ALL aggregates of that wave (G <= 512 slots,
//     one per thread) to get its own offset inside the wave and the wave total for its running
//     prefix.  No look-back chain, fixed summation order (float scans are deterministic), and a
//     slot is {tag = wave + 1, value} written with one atomic 8/16-byte store: no fences.
//   * the ragged end (n not a multiple of the row granule) is scanned by one thread of the block
//     that owns the last tile, directly in global memory (< 128 items).
//
// Replaces cub::DeviceScan as called from cupy/cuda/cupy_cub.cu:991-1013.
#pragma once
#include "scan.cuh"
#include "tma.cuh"

namespace b200 {

template <int V> struct IntTag { static constexpr int value = V; };

// input element -> accumulator (bool bytes are normalised: anything non-zero counts as one)
template <class In, class Acc> struct PipeCvt {
    B200_DEVICE static Acc in(const In& v) { return static_cast<Acc>(v); }
};
template <class Acc> struct PipeCvt<bool, Acc> {
    B200_DEVICE static Acc in(const bool& v) { return *reinterpret_cast<const uint8_t*>(&v) != 0 ? Acc(1) : Acc(0); }
};

constexpr int kPipeThreads = 512;
constexpr int kPipeRing = 16;          // waves of slots kept; needs > 2 * LAG + 1

template <int BYTES> struct PipeSlot;
template <> struct PipeSlot<4> {
    typedef uint64_t storage_t;
    typedef uint32_t bits_t;
    B200_DEVICE static void publish(storage_t* p, uint32_t tag, uint32_t bits) {
        const uint64_t w = (uint64_t(tag) << 32) | bits;
        asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
    }
    B200_DEVICE static uint32_t peek(const storage_t* p, uint32_t& bits) {
        uint64_t w;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
        bits = static_cast<uint32_t>(w);
        return static_cast<uint32_t>(w >> 32);
    }
};
template <> struct PipeSlot<8> {
    struct __align__(16) storage_t { uint64_t w0, w1; };
    typedef uint64_t bits_t;
    B200_DEVICE static void publish(storage_t* p, uint32_t tag, uint64_t bits) {
        const uint64_t w0 = (uint64_t(tag) << 32) | (bits & 0xffffffffull);
        const uint64_t w1 = (uint64_t(tag) << 32) | (bits >> 32);
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
    }
    // returns the tag; halves of different generations (a torn read) return 0
    B200_DEVICE static uint32_t peek(const storage_t* p, uint64_t& bits) {
        uint64_t w0, w1;
        asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
        const uint32_t s0 = static_cast<uint32_t>(w0 >> 32), s1 = static_cast<uint32_t>(w1 >> 32);
        bits = (w0 & 0xffffffffull) | (w1 << 32);
        return s0 == s1 ? s0 : 0u;
    }
};

template <class T, int BYTES = sizeof(T)> struct PipeBits;
template <class T> struct PipeBits<T, 4> {
    B200_DEVICE static uint32_t to(const T& v) { union { T t; uint32_t b; } u; u.t = v; return u.b; }
    B200_DEVICE static T from(uint32_t b) { union { T t; uint32_t b; } u; u.b = b; return u.t; }
};
template <class T> struct PipeBits<T, 8> {
    B200_DEVICE static uint64_t to(const T& v) { union { T t; uint64_t b; } u; u.t = v; return u.b; }
    B200_DEVICE static T from(uint64_t b) { union { T t; uint64_t b; } u; u.b = b; return u.t; }
};

// ---- shared-memory access in the shared window (32-bit addresses; LDS / STS, never generic) ----
B200_DEVICE void st_shared_v4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
B200_DEVICE uint2 ld_shared_v2(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
// byte offset inside a 128B-swizzled stage of linear byte offset `b` (the stage base is 1024-aligned)
B200_DEVICE uint32_t pipe_swz(uint32_t b) { return b ^ ((b >> 3) & 0x70u); }

template <class In, class Acc, class Out, int IPT_, int SI_, int SO_, int LAG_>
struct ScanPipeCfg {
    typedef In in_t; typedef Acc acc_t; typedef Out out_t;
    static constexpr int IPT = IPT_, SI = SI_, SO = SO_, LAG = LAG_;
    static constexpr int THREADS = kPipeThreads;
    static constexpr int IN_T = IPT * int(sizeof(In));       // bytes a thread reads per tile
    static constexpr int OUT_T = IPT * int(sizeof(Out));     // bytes a thread writes per tile
    static constexpr int IN_STAGE = THREADS * IN_T, OUT_STAGE = THREADS * OUT_T;
    static constexpr int IN_ROWS = IN_STAGE / 128, OUT_ROWS = OUT_STAGE / 128;
    static constexpr int IN_BOX = IN_ROWS > 256 ? 256 : IN_ROWS, OUT_BOX = OUT_ROWS > 256 ? 256 : OUT_ROWS;   // TMA box rows
    static constexpr int TILE = THREADS * IPT;
    static constexpr int GRANULE = 128 / int(sizeof(In) < sizeof(Out) ? sizeof(In) : sizeof(Out));   // items per widest row
    static constexpr int SMEM = 1024 + SI * IN_STAGE + SO * OUT_STAGE + 8 * SI;
    // what waits in registers for the tile prefix: the raw input items (rescanned serially when the prefix is
    // known; 16 words per thread and tile for every casting pair), or -- when input and accumulator are the same
    // type anyway -- the thread-locally scanned values, which then only need one independent add each
    static constexpr bool HOLD_RAW = !(sizeof(In) == sizeof(Acc) && sizeof(Acc) == sizeof(Out));
    static_assert(IN_T == 8 || IN_T % 16 == 0, "a thread reads 8 bytes or whole 16-byte chunks");
    static_assert(OUT_T % 16 == 0 && OUT_T <= 128 && IN_T <= 128, "per-thread spans stay inside one 128-byte row");
    static_assert(IN_ROWS >= 8 && OUT_ROWS >= 8 && IN_ROWS % IN_BOX == 0 && OUT_ROWS % OUT_BOX == 0, "TMA box rows");
    static_assert(kPipeRing > 2 * LAG + 1, "slot ring too short for this LAG");
    static_assert(SMEM <= 232448, "stages exceed the 227 KB of shared memory a block can have");
};

// MODE (lab builds): 0 = scan, 1 = scan without the cross-block exchange, 2 = convert-copy only
template <class Cfg, class Op, int MODE = 0>
__device__ __forceinline__ void scan_pipe_body(const void* tm_in, const void* tm_out,
                                               const typename Cfg::in_t* __restrict__ x,
                                               typename Cfg::out_t* __restrict__ y, int64_t n_main, int64_t n,
                                               typename PipeSlot<sizeof(typename Cfg::acc_t)>::storage_t* slots) {
    typedef typename Cfg::in_t In;
    typedef typename Cfg::acc_t Acc;
    typedef typename Cfg::out_t Out;
    typedef PipeSlot<sizeof(Acc)> Slot;
    typedef typename Slot::bits_t bits_t;
    constexpr int IPT = Cfg::IPT, SI = Cfg::SI, SO = Cfg::SO, LAG = Cfg::LAG, THREADS = Cfg::THREADS;
    constexpr int NWARPS = THREADS / 32;
    constexpr int RW = Cfg::IN_T / 4;                // raw 32-bit words a thread holds per tile
    static_assert(NWARPS == 16, "the block scan below is written for 16 warps");

    extern __shared__ uint8_t pipe_smem_raw[];
    __shared__ Acc warp_total[NWARPS];
    __shared__ Acc g_before[2][NWARPS], g_all[2][NWARPS];    // double-buffered by wave parity
    const uint32_t base = (smem_u32(pipe_smem_raw) + 1023u) & ~1023u;
    const uint32_t in0 = base, out0 = base + SI * Cfg::IN_STAGE, bar0 = out0 + SO * Cfg::OUT_STAGE;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tiles = (n_main + Cfg::TILE - 1) / Cfg::TILE;
    const int64_t G = gridDim.x, bid = blockIdx.x;
    const int my_tiles = bid < tiles ? int((tiles - bid + G - 1) / G) : 0;
    const bool dma = (tid == THREADS - 32);          // lane 0 of the last warp issues all TMA traffic
    const Acc ident = Op::template identity<Acc>();

    auto request = [&](int k) {                      // my k-th tile -> in-stage k % SI
        if (k < my_tiles) {
            const int s = k % SI;
            mbar_expect_tx_a(bar0 + 8 * s, Cfg::IN_STAGE);
#pragma unroll
            for (int r0 = 0; r0 < Cfg::IN_ROWS; r0 += Cfg::IN_BOX)
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                    ::"r"(in0 + s * Cfg::IN_STAGE + r0 * 128), "l"(tm_in), "r"(0),
                      "r"(int32_t((int64_t(k) * G + bid) * Cfg::IN_ROWS + r0)), "r"(bar0 + 8 * s) : "memory");
        }
    };

    if (dma) {
        tma_prefetch_desc(tm_in);
        tma_prefetch_desc(tm_out);
#pragma unroll
        for (int s = 0; s < SI; ++s) mbar_init_a(bar0 + 8 * s, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (dma) {
#pragma unroll
        for (int s = 0; s < SI; ++s) request(s);
    }

    constexpr int U = LAG + 1;                       // register sets: tiles it-LAG .. it
    uint32_t raw[U][RW];                             // the thread's raw input items of those tiles
    Acc texcl[U];                                    // ... and its exclusive prefix inside each of them
    Acc running = ident;                             // inclusive prefix of all waves before the one being finished
#pragma unroll
    for (int l = 0; l < U; ++l) {
        texcl[l] = ident;
#pragma unroll
        for (int j = 0; j < RW; ++j) raw[l][j] = 0u;
    }

    // One iteration.  SET = it % U is a compile-time register-set index (the loop below is unrolled by U):
    // phase A fills set SET with tile `it`; the tile finished here, it-LAG, sits in set (SET + 1) % U.
    auto iteration = [&](int it, auto set_tag) {
        constexpr int SET = decltype(set_tag)::value;
        constexpr int OLD = (SET + 1) % U;
        // ================= phase A: my tile `it` -> raw registers, thread sums -> warp scan =================
        const bool do_a = it < my_tiles;
        Acc lane_excl = ident;
        if (do_a) {
            const int s = it % SI;
            mbar_wait_a(bar0 + 8 * s, (it / SI) & 1);
            const uint32_t st = in0 + s * Cfg::IN_STAGE;
            if constexpr (Cfg::IN_T == 8) {
                const uint2 v = ld_shared_v2(st + pipe_swz(uint32_t(tid) * 8u));
                raw[SET][0] = v.x; raw[SET][1] = v.y;
            } else {
#pragma unroll
                for (int c = 0; c < Cfg::IN_T / 16; ++c) {
                    const uint4 v = ld_shared_v4(st + pipe_swz(uint32_t(tid) * Cfg::IN_T + c * 16u));
                    raw[SET][4 * c] = v.x; raw[SET][4 * c + 1] = v.y; raw[SET][4 * c + 2] = v.z; raw[SET][4 * c + 3] = v.w;
                }
            }
            if (MODE != 2) {
                const int64_t g0 = (int64_t(it) * G + bid) * Cfg::TILE + int64_t(tid) * IPT;
                const bool ragged = g0 + IPT > n_main;   // only in the last tile: items past n_main do not count
                Acc tsum = ident;
                if constexpr (Cfg::HOLD_RAW) {
                    const In* e = reinterpret_cast<const In*>(raw[SET]);
                    if (__builtin_expect(!ragged, 1)) {          // no per-item predicates on the streaming path
#pragma unroll
                        for (int j = 0; j < IPT; ++j) tsum = Op::combine(tsum, PipeCvt<In, Acc>::in(e[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < IPT; ++j)
                            if (g0 + j < n_main) tsum = Op::combine(tsum, PipeCvt<In, Acc>::in(e[j]));
                    }
                } else {
                    Acc* v = reinterpret_cast<Acc*>(raw[SET]);      // In == Acc == Out: scan in place
                    if (__builtin_expect(ragged, 0)) {
#pragma unroll
                        for (int j = 0; j < IPT; ++j)
                            if (g0 + j >= n_main) v[j] = ident;
                    }
#pragma unroll
                    for (int j = 1; j < IPT; ++j) v[j] = Op::combine(v[j - 1], v[j]);
                    tsum = v[IPT - 1];
                }
                Acc incl = tsum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const Acc t = shfl_up_any(incl, d);
                    if (lane >= d) incl = Op::combine(t, incl);
                }
                if (lane == 31) warp_total[warp] = incl;
                lane_excl = shfl_up_any(incl, 1);
                if (lane == 0) lane_excl = ident;
            }
        }
        // ---- prefetch the aggregates of wave it+1-LAG (needed by the NEXT iteration's finish); the L2 round
        // trip hides behind the barrier, the block scan and this iteration's finish
        const int kn = it + 1 - LAG;
        const int64_t nwave0 = int64_t(kn) * G;
        const int ncount = (kn >= 0 && kn < my_tiles && MODE == 0) ? int((tiles - nwave0) < G ? (tiles - nwave0) : G) : 0;
        uint32_t pf_tag = 0;
        bits_t pf_bits = 0;
        const typename Slot::storage_t* my_slot = slots + (kn & (kPipeRing - 1)) * G + tid;
        if (tid < ncount) pf_tag = Slot::peek(my_slot, pf_bits);
        __syncthreads();                                                         // (A)
        if (dma && do_a) request(it + SI);           // the in-stage of tile `it` has been read by everyone
        if (do_a && MODE != 2) {
            // block scan of the 16 warp totals, redundantly in every warp (lanes 0..15)
            Acc wt = lane < NWARPS ? warp_total[lane] : ident;
#pragma unroll
            for (int d = 1; d < NWARPS; d <<= 1) {
                const Acc t = shfl_up_any(wt, d);
                if (lane >= d) wt = Op::combine(t, wt);
            }
            const Acc block_agg = shfl_any(wt, NWARPS - 1);
            Acc warp_excl = shfl_any(wt, warp > 0 ? warp - 1 : 0);
            if (warp == 0) warp_excl = ident;
            if (tid == 0 && MODE == 0)
                Slot::publish(slots + (it & (kPipeRing - 1)) * G + bid, uint32_t(it) + 1u, PipeBits<Acc>::to(block_agg));
            texcl[SET] = Op::combine(warp_excl, lane_excl);
        }
        // ================= finish my tile it-LAG: its wave was gathered in the previous iteration ==========
        const int kf = it - LAG;
        const bool do_f = kf >= 0 && kf < my_tiles;
        if (do_f) {
            Acc prefix = running;
            if (MODE == 0) {
                const int64_t wave0 = int64_t(kf) * G;
                const int count = int((tiles - wave0) < G ? (tiles - wave0) : G);
                const int nw = (count + 31) >> 5;    // warps that wrote a partial
                Acc sb = ident, sa = ident;
#pragma unroll
                for (int w = 0; w < NWARPS; ++w) {
                    if (w < nw) {
                        sb = Op::combine(sb, g_before[kf & 1][w]);
                        sa = Op::combine(sa, g_all[kf & 1][w]);
                    }
                }
                prefix = Op::combine(running, sb);
                running = Op::combine(running, sa);
            }
            const int64_t tile = int64_t(kf) * G + bid;
            const int64_t g0 = tile * Cfg::TILE + int64_t(tid) * IPT;
            const bool ragged = g0 + IPT > n_main;
            Acc acc = MODE == 2 ? ident : Op::combine(prefix, texcl[OLD]);
            Out o[IPT];
            if constexpr (Cfg::HOLD_RAW) {
                const In* e = reinterpret_cast<const In*>(raw[OLD]);
                if (__builtin_expect(!ragged, 1)) {
#pragma unroll
                    for (int j = 0; j < IPT; ++j) {
                        const Acc v = PipeCvt<In, Acc>::in(e[j]);
                        acc = MODE == 2 ? v : Op::combine(acc, v);
                        o[j] = static_cast<Out>(acc);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < IPT; ++j) {
                        const Acc v = PipeCvt<In, Acc>::in(e[j]);
                        if (MODE == 2) acc = v;
                        else if (g0 + j < n_main) acc = Op::combine(acc, v);
                        o[j] = static_cast<Out>(acc);
                    }
                }
            } else {
                const Acc* v = reinterpret_cast<const Acc*>(raw[OLD]);
#pragma unroll
                for (int j = 0; j < IPT; ++j) o[j] = MODE == 2 ? v[j] : Op::combine(acc, v[j]);
                acc = o[IPT - 1];                    // inclusive total through this thread (the ragged-end owner needs it)
            }
            const uint32_t st = out0 + (kf % SO) * Cfg::OUT_STAGE;
#pragma unroll
            for (int c = 0; c < Cfg::OUT_T / 16; ++c)
                st_shared_v4(st + pipe_swz(uint32_t(tid) * Cfg::OUT_T + c * 16u), reinterpret_cast<const uint4*>(o)[c]);
            fence_proxy_async_smem();
            // ragged end: the owner of the very last item of the last tile scans the < GRANULE leftover items
            if (__builtin_expect(tile == tiles - 1 && tid == THREADS - 1 && n_main < n, 0)) {
#pragma unroll 1
                for (int64_t i = n_main; i < n; ++i) {
                    acc = Op::combine(acc, PipeCvt<In, Acc>::in(x[i]));
                    y[i] = static_cast<Out>(acc);
                }
            }
        }
        // ---- wave it+1-LAG: the aggregates before mine (my offset) and all of them (wave total)
        if (ncount > 0 && warp * 32 < ncount) {
            Acc before = ident, all = ident;
            if (tid < ncount) {
                const uint32_t want = uint32_t(kn) + 1u;
                while (pf_tag != want) pf_tag = Slot::peek(my_slot, pf_bits);
                const Acc v = PipeBits<Acc>::from(pf_bits);
                all = v;
                if (tid < bid) before = v;
            }
            // fixed-order tree: deterministic for floats
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                before = Op::combine(before, shfl_down_any(before, d));
                all = Op::combine(all, shfl_down_any(all, d));
            }
            if (lane == 0) { g_before[kn & 1][warp] = before; g_all[kn & 1][warp] = all; }
        }
        __syncthreads();                                                         // (C)
        if (dma && do_f) {
            const int so = kf % SO;
#pragma unroll
            for (int r0 = 0; r0 < Cfg::OUT_ROWS; r0 += Cfg::OUT_BOX)
                asm volatile(
                    "cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                    ::"l"(tm_out), "r"(0), "r"(int32_t((int64_t(kf) * G + bid) * Cfg::OUT_ROWS + r0)),
                      "r"(out0 + so * Cfg::OUT_STAGE + r0 * 128) : "memory");
            tma_commit_group();
            tma_wait_group_read<SO - 1>();           // the out-stage written next iteration is free again
        }
    };

    // my_tiles + LAG iterations, rounded up to a multiple of U (the extra ones only pass the barriers), so that
    // the body exists exactly U times in the code
    const int iters = (my_tiles + LAG + U - 1) / U * U;
    for (int it = 0; it < iters; it += U) {
        iteration(it, IntTag<0>());
        if constexpr (U >= 2) iteration(it + 1, IntTag<1>());
        if constexpr (U >= 3) iteration(it + 2, IntTag<2>());
        if constexpr (U >= 4) iteration(it + 3, IntTag<3>());
        if constexpr (U >= 5) iteration(it + 4, IntTag<4>());
        static_assert(U <= 5, "LAG <= 4");
    }
    if (dma) tma_wait_group<0>();    // stores must complete before the block retires its shared memory
}

}  // namespace b200
