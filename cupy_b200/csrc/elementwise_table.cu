// elementwise_table.cu -- prebuilt sm_100a kernels for the built-in ufunc table.  One glue template (`ew_kernel`) is instantiated over the three
// tilers of b200/elementwise.cuh for every (ufunc, dtype loop) in the table; a
// call whose operand dtypes are not in the table returns B200_E_UNSUPPORTED and
// the Python host compiles the same glue around the routine string with NVRTC.
//
// Replaces cupy/_core/_kernel.pyx:1024-1100 (_get_ufunc_kernel: JIT per dtype /
// ndim / contiguity) + cupy/cuda/function.pyx:153-171 (linear_launch, 128-thread
// blocks, one element per thread).
#include "common.h"
#include "elementwise_registry.h"
#include "include/b200/ufunc_ops.cuh"

namespace b200 {


template <bool FULL, class Tiler, class F>
__device__ __forceinline__ void ew_tile(const Tiler& t, const EwParams& p) {
    constexpr int V = Tiler::kV, U = Tiler::kU;
    typedef typename F::in0_t T0;
    typedef typename F::in1_t T1;
    typedef typename F::in2_t T2;
    typedef typename F::out_t TO;
    Pack<T0, V> a0[U];
    Pack<T1, V> a1[U];
    Pack<T2, V> a2[U];
    Pack<TO, V> o[U];
    if (p.scalar_mask & 1u) {
        const T0 s = scalar_arg<T0>(p, 0);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < V; ++k) a0[u][k] = s;
    } else {
        t.template load<FULL>(0, a0);
    }
    if (F::nin >= 2) {
        if (p.scalar_mask & 2u) {
            const T1 s = scalar_arg<T1>(p, 1);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < V; ++k) a1[u][k] = s;
        } else {
            t.template load<FULL>(1, a1);
        }
    }
    if (F::nin >= 3) {
        if (p.scalar_mask & 4u) {
            const T2 s = scalar_arg<T2>(p, 2);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < V; ++k) a2[u][k] = s;
        } else {
            t.template load<FULL>(2, a2);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < V; ++k)
            o[u][k] = F::apply(a0[u][k], F::nin >= 2 ? a1[u][k] : T1(), F::nin >= 3 ? a2[u][k] : T2());
    t.template store<FULL>(F::nin, o);
}

template <class Tiler, class F, int MIN_BLOCKS = 1>
__global__ void __launch_bounds__(kEwThreads, MIN_BLOCKS) ew_kernel(const __grid_constant__ EwParams p) {
    Tiler t(p);
    for (; t.valid(); t.next()) {
        if (t.is_full()) ew_tile<true, Tiler, F>(t, p);
        else ew_tile<false, Tiler, F>(t, p);
    }
}

template <class Tiler, class F>
__global__ void __launch_bounds__(Tiler::kThreads) ew_tma_kernel(const __grid_constant__ EwParams p,
                                                                 const __grid_constant__ TileMaps tm) {
    Tiler t(p, tm);
    for (; t.valid(); t.next()) {
        if (t.is_full()) ew_tile<true, Tiler, F>(t, p);
        else ew_tile<false, Tiler, F>(t, p);
    }
}

template <class F, bool OK> struct TmaKernelOf { static const void* get() { return nullptr; } };
template <class F> struct TmaKernelOf<F, true> {
    static const void* get() {
        return reinterpret_cast<const void*>(&ew_tma_kernel<TmaTileTiler<F::nin + 1, int(sizeof(typename F::out_t))>, F>);
    }
};
// TILED_REG is prebuilt for unary ufuncs only (input staged, output unit-stride: SPEC known);
// calls with more operands go through NVRTC, where the generator fixes each operand's access.
template <class F, bool OK> struct RegKernelOf { static const void* get() { return nullptr; } };
template <class F> struct RegKernelOf<F, true> {
    static const void* get() {
        constexpr int e = int(sizeof(typename F::out_t));
        return reinterpret_cast<const void*>(&ew_kernel<RegTileTiler<2, e, reg_tile_unroll(e), (1ull | (2ull << 3))>, F, (e == 2 ? 2 : 3)>);
    }
};
template <class F> constexpr bool tma_eligible() {
    constexpr int so = int(sizeof(typename F::out_t));
    return (so == 2 || so == 4 || so == 8) && int(sizeof(typename F::in0_t)) == so &&
           (F::nin < 2 || int(sizeof(typename F::in1_t)) == so) && (F::nin < 3 || int(sizeof(typename F::in2_t)) == so);
}

template <class T> constexpr int max_size2(int a) { return int(sizeof(T)) > a ? int(sizeof(T)) : a; }

template <class F>
static EwKernels make_kernels() {
    constexpr int N = F::nin + 1;
    constexpr int maxsz = max_size2<typename F::out_t>(int(sizeof(typename F::in0_t)));
    constexpr int V = (16 / maxsz) < 1 ? 1 : (16 / maxsz);
    constexpr int UF = 4, UR = 2;   // ROWWISE: B200 sweep in profiles/r01_row_probe.log (2 beats 1 and 4 once the index math is 32-bit)
    EwKernels k;
    k.flat_v = reinterpret_cast<const void*>(&ew_kernel<FlatTiler<N, V, UF, kEwThreads>, F>);
    k.flat_1 = reinterpret_cast<const void*>(&ew_kernel<FlatTiler<N, 1, UF, kEwThreads>, F>);
    // vector ROWWISE kernels are prebuilt for the all-unit-stride case (padded / sliced rows); calls with
    // broadcast or strided operands are specialised by NVRTC (run-time stride branches cost 10-17 %)
    constexpr uint64_t kAllUnit = N == 2 ? 0x5ull : N == 3 ? 0x15ull : 0x55ull;
    k.row_v32 = reinterpret_cast<const void*>(&ew_kernel<RowTiler<N, V, UR, kEwThreads, true, kAllUnit>, F>);
    k.row_132 = reinterpret_cast<const void*>(&ew_kernel<RowTiler<N, 1, UR, kEwThreads, true>, F>);
    k.row_164 = reinterpret_cast<const void*>(&ew_kernel<RowTiler<N, 1, UR, kEwThreads, false>, F>);
    k.row_v64 = reinterpret_cast<const void*>(&ew_kernel<RowTiler<N, V, UR, kEwThreads, false, kAllUnit>, F>);
    k.tiled = reinterpret_cast<const void*>(&ew_kernel<TileTiler<N>, F>);
    k.tiled_tma = TmaKernelOf<F, tma_eligible<F>()>::get();
    k.tiled_reg = RegKernelOf<F, tma_eligible<F>() && F::nin == 1>::get();
    k.vec = V;
    k.unroll_flat = UF;
    k.unroll_row = UR;
    return k;
}

template <class T> struct dtype_id;
template <> struct dtype_id<int8_t> { static constexpr int v = B200_TYPE_INT8; };
template <> struct dtype_id<uint8_t> { static constexpr int v = B200_TYPE_UINT8; };
template <> struct dtype_id<int16_t> { static constexpr int v = B200_TYPE_INT16; };
template <> struct dtype_id<uint16_t> { static constexpr int v = B200_TYPE_UINT16; };
template <> struct dtype_id<int32_t> { static constexpr int v = B200_TYPE_INT32; };
template <> struct dtype_id<uint32_t> { static constexpr int v = B200_TYPE_UINT32; };
template <> struct dtype_id<long long> { static constexpr int v = B200_TYPE_INT64; };
template <> struct dtype_id<unsigned long long> { static constexpr int v = B200_TYPE_UINT64; };
template <> struct dtype_id<float16> { static constexpr int v = B200_TYPE_FLOAT16; };
template <> struct dtype_id<float> { static constexpr int v = B200_TYPE_FLOAT32; };
template <> struct dtype_id<double> { static constexpr int v = B200_TYPE_FLOAT64; };
template <> struct dtype_id<bool> { static constexpr int v = B200_TYPE_BOOL; };

template <class F>
static void reg(int ufunc) {
    registry()[Key(ufunc, dtype_id<typename F::in0_t>::v, dtype_id<typename F::out_t>::v)] = make_kernels<F>();
}

template <template <class, class> class F>
static void reg_arith(int ufunc) {   // the loops the configs and their neighbours use
    reg<F<int32_t, int32_t>>(ufunc);
    reg<F<long long, long long>>(ufunc);
    reg<F<float16, float16>>(ufunc);
    reg<F<float, float>>(ufunc);
    reg<F<double, double>>(ufunc);
}
template <template <class, class> class F>
static void reg_float(int ufunc) {
    reg<F<float16, float16>>(ufunc);
    reg<F<float, float>>(ufunc);
    reg<F<double, double>>(ufunc);
}
template <class TI>
static void reg_copy_from() {
    reg<CopyF<TI, int32_t>>(B200_UF_COPY);
    reg<CopyF<TI, long long>>(B200_UF_COPY);
    reg<CopyF<TI, float16>>(B200_UF_COPY);
    reg<CopyF<TI, float>>(B200_UF_COPY);
    reg<CopyF<TI, double>>(B200_UF_COPY);
}

// The table is instantiated in independent groups (one translation unit each,
// -DB200_EW_GROUP=k) so that the groups compile in parallel.
#if B200_EW_GROUP == 0
void register_ew_group0() {
    reg_arith<AddF>(B200_UF_ADD);
    reg<AddF<bool, bool>>(B200_UF_ADD);
    reg_arith<SubtractF>(B200_UF_SUBTRACT);
}
#elif B200_EW_GROUP == 1
void register_ew_group1() {
    reg_arith<MultiplyF>(B200_UF_MULTIPLY);
    reg<MultiplyF<bool, bool>>(B200_UF_MULTIPLY);
    reg_float<TrueDivideF>(B200_UF_TRUE_DIVIDE);
    reg_float<FmaF>(B200_UF_FMA);
}
#elif B200_EW_GROUP == 2
void register_ew_group2() {
    reg_arith<NegativeF>(B200_UF_NEGATIVE);
    reg_arith<AbsoluteF>(B200_UF_ABSOLUTE);
    reg_arith<SquareF>(B200_UF_SQUARE);
}
#elif B200_EW_GROUP == 3
void register_ew_group3() {
    reg_float<SqrtF>(B200_UF_SQRT);
    reg_float<ExpF>(B200_UF_EXP);
    reg_float<LogF>(B200_UF_LOG);
    reg_arith<MaximumF>(B200_UF_MAXIMUM);
}
#elif B200_EW_GROUP == 4
void register_ew_group4() {
    reg_arith<MinimumF>(B200_UF_MINIMUM);
    // dtype-casting copy (elementwise_copy / astype): all pairs of the five
    // main dtypes ...
    reg_copy_from<int32_t>();
    reg_copy_from<long long>();
}
#elif B200_EW_GROUP == 5
void register_ew_group5() {
    reg_copy_from<float16>();
    reg_copy_from<float>();
    reg_copy_from<double>();
}
#elif B200_EW_GROUP == 6
void register_ew_group6() {
    // ... plus same-dtype copies of every other width and the casts scans/sums use
    reg<CopyF<int8_t, int8_t>>(B200_UF_COPY);
    reg<CopyF<uint8_t, uint8_t>>(B200_UF_COPY);
    reg<CopyF<int16_t, int16_t>>(B200_UF_COPY);
    reg<CopyF<uint16_t, uint16_t>>(B200_UF_COPY);
    reg<CopyF<uint32_t, uint32_t>>(B200_UF_COPY);
    reg<CopyF<unsigned long long, unsigned long long>>(B200_UF_COPY);
    reg<CopyF<bool, bool>>(B200_UF_COPY);
    reg<CopyF<bool, long long>>(B200_UF_COPY);
    reg<CopyF<int8_t, long long>>(B200_UF_COPY);
    reg<CopyF<uint8_t, long long>>(B200_UF_COPY);
}
#endif

}  // namespace b200
