// b200/scan.cuh -- single-pass inclusive scan (cumsum / cumprod) with decoupled
// look-back.  One read of x, one write of y; the input dtype is converted while
// loading, so the reference's separate `astype` pass
// (cupy/_core/_routines_math.pyx:726-727) disappears.
//
// Tile = 256 threads x 16 items.  Global traffic is 128-bit and fully coalesced
// (a warp instruction moves 512 consecutive bytes); a warp-private, padded
// shared-memory exchange (no block barrier, no bank conflicts) turns that
// striped order into 16 consecutive items per lane, which are scanned in
// registers; lane totals are scanned with shuffles, warp totals through shared
// memory, and tile prefixes through the look-back chain in global memory.
//
// Replaces cub::DeviceScan as called from cupy/cuda/cupy_cub.cu:991-1013 and the
// three-phase fallback cupy/_core/_routines_math.pyx:160-496.
#pragma once
#include "base.cuh"

namespace b200 {

struct ScanSum {
    template <class T> B200_DEVICE static T identity() { return T(0); }
    template <class T> B200_DEVICE static T combine(const T& a, const T& b) { return a + b; }
};
struct ScanProd {
    template <class T> B200_DEVICE static T identity() { return T(1); }
    template <class T> B200_DEVICE static T combine(const T& a, const T& b) { return a * b; }
};

constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

// workspace layout (host mirrors this in scan.cu)
template <class Acc>
struct ScanWorkspace {
    uint32_t* counter;   // dynamic tile ids            (zeroed per launch)
    uint32_t* flags;     // 0 none, 1 aggregate, 2 inclusive (zeroed per launch)
    Acc*      aggregate;
    Acc*      inclusive;
};

template <class T>
B200_DEVICE void store_cg(T* p, const T& v) {
    // partial prefixes are consumed by other SMs: keep them out of L1
    constexpr int words = sizeof(T) / 4;
    union U { T t; uint32_t w[words]; B200_DEVICE U() {} };
    U u;
    u.t = v;
    volatile uint32_t* d = reinterpret_cast<volatile uint32_t*>(p);
#pragma unroll
    for (int i = 0; i < words; ++i) d[i] = u.w[i];
}

template <class T>
B200_DEVICE T load_cg(const T* p) {
    constexpr int words = sizeof(T) / 4;
    union U { T t; uint32_t w[words]; B200_DEVICE U() {} };
    U u;
    const volatile uint32_t* s = reinterpret_cast<const volatile uint32_t*>(p);
#pragma unroll
    for (int i = 0; i < words; ++i) u.w[i] = s[i];
    return u.t;
}

template <class In, class Acc, class Out, class Op>
__device__ __forceinline__ void scan_body(const In* __restrict__ x, Out* __restrict__ y, int64_t n,
                                          ScanWorkspace<Acc> ws) {
    constexpr int ITEMS = kScanItems;
    constexpr int LB_IN = ITEMS * int(sizeof(In));
    constexpr int LB_OUT = ITEMS * int(sizeof(Out));
    constexpr int PITCH = (LB_IN > LB_OUT ? LB_IN : LB_OUT) + 16;
    constexpr int NWARPS = kScanThreads / 32;
    __shared__ __align__(16) char sm[NWARPS][32 * PITCH];
    __shared__ Acc warp_total[NWARPS];
    __shared__ Acc tile_prefix_s;
    __shared__ uint32_t tile_id_s;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) tile_id_s = atomicAdd(ws.counter, 1u);
    __syncthreads();
    const int64_t tile = tile_id_s;
    const int64_t warp_base = tile * kScanTile + int64_t(warp) * 32 * ITEMS;
    const int64_t rem = n - warp_base;               // valid elements in this warp's segment
    char* wsm = sm[warp];

    // ---- load: coalesced 16-byte chunks -> padded shared -> 16 consecutive items per lane
    Acc item[ITEMS];
    {
        constexpr int NCH = LB_IN / 16;
        constexpr int EPC = 16 / int(sizeof(In));    // elements per chunk
        const In* src = x + warp_base;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int off = (c * 32 + lane) * 16;
            const int e = off / int(sizeof(In));
            Pack<In, EPC> v;
            if (e + EPC <= rem) {
                load_pack(v, src + e);
            } else {
#pragma unroll
                for (int k = 0; k < EPC; ++k) v[k] = (e + k < rem) ? src[e + k] : In();
            }
            const int L = off / LB_IN, w = off % LB_IN;
            *reinterpret_cast<Pack<In, EPC>*>(wsm + L * PITCH + w) = v;
        }
        __syncwarp();
        const int64_t e0 = int64_t(lane) * ITEMS;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            Pack<In, EPC> v = *reinterpret_cast<const Pack<In, EPC>*>(wsm + lane * PITCH + c * 16);
#pragma unroll
            for (int k = 0; k < EPC; ++k)
                item[c * EPC + k] = (e0 + c * EPC + k < rem) ? static_cast<Acc>(v[k])
                                                             : Op::template identity<Acc>();
        }
        __syncwarp();
    }

    // ---- registers: inclusive scan of the lane's items
#pragma unroll
    for (int j = 1; j < ITEMS; ++j) item[j] = Op::combine(item[j - 1], item[j]);

    // ---- warp: scan of lane totals
    Acc incl = item[ITEMS - 1];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Acc t = shfl_up_any(incl, d);
        if (lane >= d) incl = Op::combine(t, incl);
    }
    Acc lane_excl = shfl_up_any(incl, 1);
    if (lane == 0) lane_excl = Op::template identity<Acc>();
    if (lane == 31) warp_total[warp] = incl;
    __syncthreads();

    Acc warp_excl = Op::template identity<Acc>();
    Acc block_agg = Op::template identity<Acc>();
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) {
        const Acc t = warp_total[w];
        if (w < warp) warp_excl = Op::combine(warp_excl, t);
        block_agg = Op::combine(block_agg, t);
    }

    // ---- grid: decoupled look-back (warp 0)
    if (warp == 0) {
        Acc prefix = Op::template identity<Acc>();
        if (tile == 0) {
            if (lane == 0) {
                store_cg(ws.inclusive + 0, block_agg);
                __threadfence();
                st_release_u32(ws.flags + 0, 2u);
            }
        } else {
            if (lane == 0) {
                store_cg(ws.aggregate + tile, block_agg);
                __threadfence();
                st_release_u32(ws.flags + tile, 1u);
            }
            int64_t look = tile - 1;          // lane 0 inspects `look`, lane k inspects look-k
            while (true) {
                const int64_t t = look - lane;
                uint32_t f = 2u;              // tiles before 0 behave as "inclusive = identity"
                if (t >= 0) {
                    do { f = ld_acquire_u32(ws.flags + t); } while (f == 0u);
                }
                Acc v = Op::template identity<Acc>();
                if (t >= 0) v = (f == 2u) ? load_cg(ws.inclusive + t) : load_cg(ws.aggregate + t);
                const uint32_t done = __ballot_sync(0xffffffffu, f == 2u);
                const int first = done ? (__ffs(done) - 1) : 31;    // closest tile holding an inclusive value
                if (lane > first) v = Op::template identity<Acc>();
                // fold lanes first..0 so that farther tiles stay on the left
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    Acc o = shfl_down_any(v, d);
                    if (lane + d < 32) v = Op::combine(o, v);
                }
                if (lane == 0) prefix = Op::combine(v, prefix);
                if (done) break;
                look -= 32;
            }
            if (lane == 0) {
                store_cg(ws.inclusive + tile, Op::combine(prefix, block_agg));
                __threadfence();
                st_release_u32(ws.flags + tile, 2u);
            }
        }
        if (lane == 0) tile_prefix_s = prefix;
    }
    __syncthreads();

    // ---- apply prefixes, exchange back, coalesced store
    const Acc pre = Op::combine(Op::combine(tile_prefix_s, warp_excl), lane_excl);
    {
        constexpr int NCH = LB_OUT / 16;
        constexpr int EPC = 16 / int(sizeof(Out));
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            Pack<Out, EPC> v;
#pragma unroll
            for (int k = 0; k < EPC; ++k) v[k] = static_cast<Out>(Op::combine(pre, item[c * EPC + k]));
            *reinterpret_cast<Pack<Out, EPC>*>(wsm + lane * PITCH + c * 16) = v;
        }
        __syncwarp();
        Out* dst = y + warp_base;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const int off = (c * 32 + lane) * 16;
            const int e = off / int(sizeof(Out));
            const int L = off / LB_OUT, w = off % LB_OUT;
            const Pack<Out, EPC> v = *reinterpret_cast<const Pack<Out, EPC>*>(wsm + L * PITCH + w);
            if (e + EPC <= rem) {
                store_pack(dst + e, v);
            } else {
#pragma unroll
                for (int k = 0; k < EPC; ++k)
                    if (e + k < rem) dst[e + k] = v[k];
            }
        }
    }
}

}  // namespace b200
