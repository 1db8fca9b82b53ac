// b200/scan_march.cuh -- cumsum / cumprod down the rows of x[outer][n][inner] for matrices of a few
// thousand to ~150 thousand columns: too few column vectors to fill the GPU with one thread per column
// (scan_cols), too many rows to waste a second read on segment totals (the split scheme, 12 B / float32).
//
// The columns are cut into strips of W_BYTES (128 / 256 / 512 bytes per row, so that there are about as many
// strips as SMs; the width is that of the wider of input and result, so casting scans -- int32 -> int64,
// bool -> int64, float16 with a float accumulator -- run here too); ONE persistent 512-thread block marches down each strip, tile by tile (R rows x W_BYTES),
// carrying the running column sums in registers.  Nothing is exchanged between blocks: the parallelism is
// across strips, the bytes in flight come from the TMA ring (input tiles of 32 KB, 4 stages deep, one tensor
// copy each; output tiles leave through a second ring by TMA stores).  A warp owns RW * RPT consecutive rows of
// the tile and all its columns (a lane: one 16-byte column chunk of RPT rows), scans them in registers; the
// row-group totals meet by shuffles inside the warp and the 16 warp totals through shared memory.
// x is read once, y written once.  Replaces _proc_as_batch + _batch_scan_op (cupy/_core/_routines_math.pyx:499-699).
#pragma once
#include "scan.cuh"
#include "scan_pipe.cuh"
#include "tma.cuh"

namespace b200 {

template <class In, class Acc, class Out, int W_BYTES_, int SI_, int SO_>
struct ScanMarchCfg {
    typedef In in_t; typedef Acc acc_t; typedef Out out_t;
    static constexpr int THREADS = 512, NWARPS = 16, RPT = 4;
    static constexpr int SWIDE = int(sizeof(In) > sizeof(Out) ? sizeof(In) : sizeof(Out));
    static constexpr int W_BYTES = W_BYTES_;                 // bytes of a tile row on its WIDER side: 128, 256 or 512
    static constexpr int LPR = W_BYTES / 16;                 // lanes per row: 8, 16, 32
    static constexpr int RW = 32 / LPR;                      // rows a warp covers per load instruction: 4, 2, 1
    static constexpr int R = NWARPS * RW * RPT;              // tile rows: 256, 128, 64
    static constexpr int V = 16 / SWIDE;                     // columns per lane (16 bytes of the wider type)
    static constexpr int W = LPR * V;                        // tile columns
    static constexpr int IN_LANE = V * int(sizeof(In)), OUT_LANE = V * int(sizeof(Out));   // bytes a lane reads / writes per row
    static constexpr int IN_ROW = W * int(sizeof(In)), OUT_ROW = W * int(sizeof(Out));
    static constexpr int IN_STAGE = R * IN_ROW, OUT_STAGE = R * OUT_ROW;
    static constexpr int SI = SI_, SO = SO_;
    static constexpr int SMEM = 1024 + SI * ((IN_STAGE + 1023) / 1024 * 1024) + SO * OUT_STAGE + 8 * SI;
    static constexpr int IN_PITCH = (IN_STAGE + 1023) / 1024 * 1024;     // stage pitch (TMA destinations stay 128-byte aligned)
    static_assert(W_BYTES == 128 || W_BYTES == 256 || W_BYTES == 512, "strip width");
    static_assert(R <= 256 && W <= 256 && IN_ROW % 16 == 0 && OUT_ROW % 16 == 0, "TMA box");
    static_assert(IN_LANE == 16 || IN_LANE == 8 || IN_LANE == 4 || IN_LANE == 2, "lane load width");
};

template <int BYTES> struct MarchLoad;
template <> struct MarchLoad<16> {
    B200_DEVICE static void ld(uint32_t addr, void* dst) { *reinterpret_cast<uint4*>(dst) = ld_shared_v4(addr); }
};
template <> struct MarchLoad<8> {
    B200_DEVICE static void ld(uint32_t addr, void* dst) { *reinterpret_cast<uint2*>(dst) = ld_shared_v2(addr); }
};
template <> struct MarchLoad<4> {
    B200_DEVICE static void ld(uint32_t addr, void* dst) {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
        *reinterpret_cast<uint32_t*>(dst) = v;
    }
};
template <> struct MarchLoad<2> {
    B200_DEVICE static void ld(uint32_t addr, void* dst) {
        uint16_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
        *reinterpret_cast<uint16_t*>(dst) = v;
    }
};

template <class Cfg, class Op>
__device__ __forceinline__ void scan_march_body(const void* tm_in, const void* tm_out, int64_t outer, int64_t n,
                                                int64_t inner) {
    typedef typename Cfg::in_t In;
    typedef typename Cfg::acc_t T;                          // accumulator; the tile is converted on load
    typedef typename Cfg::out_t Out;
    constexpr int THREADS = Cfg::THREADS, NWARPS = Cfg::NWARPS, RPT = Cfg::RPT, LPR = Cfg::LPR, RW = Cfg::RW;
    constexpr int V = Cfg::V, W = Cfg::W, R = Cfg::R, SI = Cfg::SI, SO = Cfg::SO;

    extern __shared__ uint8_t march_smem_raw[];
    __shared__ __align__(16) T wslab[NWARPS][W];            // per-warp column totals of the current tile
    __shared__ __align__(16) T tile_tot[W];                 // column totals of the whole tile
    const uint32_t base = (smem_u32(march_smem_raw) + 1023u) & ~1023u;
    const uint32_t in0 = base, out0 = base + SI * Cfg::IN_PITCH, bar0 = out0 + SO * Cfg::OUT_STAGE;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = lane % LPR, rg = lane / LPR;               // 16-byte column chunk, row group inside the warp
    const bool dma = (tid == THREADS - 32);
    const T ident = Op::template identity<T>();

    const int64_t nstrips = (inner + W - 1) / W;
    const int64_t ntr = (n + R - 1) / R;                     // row tiles per strip
    const int64_t insts = outer * nstrips;                   // independent strips (one per outer slice and column strip)
    const int64_t G = gridDim.x, bid = blockIdx.x;
    const int64_t my_insts = bid < insts ? (insts - bid + G - 1) / G : 0;
    const int64_t my_tiles = my_insts * ntr;

    auto coords = [&](int64_t t, int64_t& o, int64_t& strip, int64_t& seg) {
        const int64_t inst = bid + (t / ntr) * G;
        seg = t % ntr;
        o = inst / nstrips;
        strip = inst % nstrips;
    };
    auto request = [&](int64_t t) {
        if (t < my_tiles) {
            int64_t o, strip, seg;
            coords(t, o, strip, seg);
            const int s = int(t % SI);
            mbar_expect_tx_a(bar0 + 8 * s, Cfg::IN_STAGE);
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                ::"r"(in0 + s * Cfg::IN_PITCH), "l"(tm_in), "r"(int32_t(strip * W)), "r"(int32_t(seg * R)), "r"(int32_t(o)),
                  "r"(bar0 + 8 * s) : "memory");
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

    T running[V];                                            // column sums of all tiles above the current one
#pragma unroll
    for (int k = 0; k < V; ++k) running[k] = ident;
    const uint32_t my_row = uint32_t(warp * RW * RPT + rg * RPT);                             // first of my RPT rows
    const uint32_t my_in = my_row * Cfg::IN_ROW + c * Cfg::IN_LANE, my_out = my_row * Cfg::OUT_ROW + c * Cfg::OUT_LANE;

    for (int64_t t = 0; t < my_tiles; ++t) {
        int64_t o, strip, seg;
        coords(t, o, strip, seg);
        if (seg == 0) {
#pragma unroll
            for (int k = 0; k < V; ++k) running[k] = ident;
        }
        const int s = int(t % SI);
        mbar_wait_a(bar0 + 8 * s, uint32_t((t / SI) & 1));
        const uint32_t st = in0 + s * Cfg::IN_PITCH + my_in;
        T a[RPT][V];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            In e[V];
            MarchLoad<Cfg::IN_LANE>::ld(st + r * Cfg::IN_ROW, e);
#pragma unroll
            for (int k = 0; k < V; ++k) a[r][k] = PipeCvt<In, T>::in(e[k]);
        }
        // rows past n and columns past inner were zero-filled by TMA: cumprod needs ones there (the stores are clipped)
        const int64_t row0 = seg * R + (warp * RW * RPT + rg * RPT);
        const int64_t col0 = strip * W + int64_t(c) * V;
        if (__builtin_expect(row0 + RPT > n || col0 + V > inner, 0)) {
#pragma unroll
            for (int r = 0; r < RPT; ++r)
#pragma unroll
                for (int k = 0; k < V; ++k)
                    if (row0 + r >= n || col0 + k >= inner) a[r][k] = ident;
        }
        // scan down my RPT rows
#pragma unroll
        for (int k = 0; k < V; ++k)
#pragma unroll
            for (int r = 1; r < RPT; ++r) a[r][k] = Op::combine(a[r - 1][k], a[r][k]);
        // row groups of the warp (lanes LPR apart hold the same columns, consecutive row groups)
        T incl[V], excl[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            incl[k] = a[RPT - 1][k];
#pragma unroll
            for (int d = LPR; d < 32; d <<= 1) {
                const T u = shfl_up_any(incl[k], d);
                if (lane >= d) incl[k] = Op::combine(u, incl[k]);
            }
            excl[k] = shfl_up_any(incl[k], LPR);
            if (rg == 0) excl[k] = ident;
        }
        if (rg == RW - 1) {
#pragma unroll
            for (int k = 0; k < V; ++k) wslab[warp][c * V + k] = incl[k];
        }
        __syncthreads();                                                         // (A)
        if (dma) request(t + SI);                    // the in-stage has been read by everyone
        // warps above mine
        T before[V];
#pragma unroll
        for (int k = 0; k < V; ++k) before[k] = ident;
#pragma unroll
        for (int w = 0; w < NWARPS - 1; ++w) {
            if (w < warp) {
#pragma unroll
                for (int k = 0; k < V; ++k) before[k] = Op::combine(before[k], wslab[w][c * V + k]);
            }
        }
        if (warp == NWARPS - 1 && rg == RW - 1) {
#pragma unroll
            for (int k = 0; k < V; ++k) tile_tot[c * V + k] = Op::combine(before[k], incl[k]);
        }
        T pre[V];
#pragma unroll
        for (int k = 0; k < V; ++k) pre[k] = Op::combine(running[k], Op::combine(before[k], excl[k]));
        const uint32_t so = out0 + uint32_t(t % SO) * Cfg::OUT_STAGE + my_out;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            Out o4[V];
#pragma unroll
            for (int k = 0; k < V; ++k) o4[k] = static_cast<Out>(Op::combine(pre[k], a[r][k]));
            if constexpr (Cfg::OUT_LANE == 16) {
                st_shared_v4(so + r * Cfg::OUT_ROW, *reinterpret_cast<const uint4*>(o4));
            } else {
                static_assert(Cfg::OUT_LANE == 8, "a lane writes 16 bytes, or 8 when the input is the wider side");
                const uint2 q = *reinterpret_cast<const uint2*>(o4);
                asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(so + r * Cfg::OUT_ROW), "r"(q.x), "r"(q.y) : "memory");
            }
        }
        fence_proxy_async_smem();
        __syncthreads();                                                         // (C)
        if (dma) {
            asm volatile(
                "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                ::"l"(tm_out), "r"(int32_t(strip * W)), "r"(int32_t(seg * R)), "r"(int32_t(o)),
                  "r"(out0 + uint32_t(t % SO) * Cfg::OUT_STAGE) : "memory");
            tma_commit_group();
            tma_wait_group_read<SO - 1>();
        }
#pragma unroll
        for (int k = 0; k < V; ++k) running[k] = Op::combine(running[k], tile_tot[c * V + k]);
    }
    if (dma) tma_wait_group<0>();
}

}  // namespace b200
