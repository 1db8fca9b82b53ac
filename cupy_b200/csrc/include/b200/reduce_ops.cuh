// b200/reduce_ops.cuh -- built-in reduction functors for the skeleton in
// reduce.cuh.  Semantics follow the reference's routine tables:
//   sum / prod      cupy/_core/_routines_math.pyx:777-807, 851-866
//                   (ints accumulate in 64 bit, float16 accumulates in float)
//   min / max       cupy/_core/_routines_statistics.pyx:279-307 (NaN propagates)
//   argmin / argmax cupy/_core/_routines_statistics.pyx:255-276, 342-353
//                   (ties -> lowest index; NaN wins; among NaNs the FIRST one,
//                   which is NumPy's rule and one of the reference's outcomes)
//   mean            cupy/_core/_routines_statistics.pyx:647-655
//   var             cupy/_core/_routines_statistics.pyx:556-643 (the reference runs
//                   two passes; here one pass: per-lane Welford + Chan merge)
#pragma once
#include <limits>

#include "base.cuh"
#include "reduce.cuh"

namespace b200 {

struct NoCtx {};

template <class T> struct is_floating { static constexpr bool value = false; };
template <> struct is_floating<float> { static constexpr bool value = true; };
template <> struct is_floating<double> { static constexpr bool value = true; };
template <> struct is_floating<float16> { static constexpr bool value = true; };

template <class T> B200_DEVICE bool is_nan(const T&) { return false; }
template <> B200_DEVICE bool is_nan<float>(const float& v) { return v != v; }
template <> B200_DEVICE bool is_nan<double>(const double& v) { return v != v; }
template <> B200_DEVICE bool is_nan<float16>(const float16& v) { return __hisnan(v.raw()); }

// comparison domain: float16 compares as float, everything else as itself
template <class T> struct cmp_type { typedef T type; };
template <> struct cmp_type<float16> { typedef float type; };
template <> struct cmp_type<bool> { typedef int type; };

template <class In, class Acc, class Out>
struct SumOp {
    typedef In in_t; typedef Acc acc_t; typedef Out out_t; typedef long long index_t; typedef NoCtx ctx_t;
    static constexpr bool kWideIndex = false;
    B200_DEVICE acc_t identity() const { return acc_t(0); }
    B200_DEVICE ctx_t step(int) const { return ctx_t(); }
    B200_DEVICE void accumulate(acc_t& a, const ctx_t&, const in_t& v, index_t) const { a = a + static_cast<acc_t>(v); }
    B200_DEVICE acc_t single(const in_t& v, index_t) const { return static_cast<acc_t>(v); }
    B200_DEVICE acc_t combine(const acc_t& a, const acc_t& b) const { return a + b; }
    B200_DEVICE out_t post(const acc_t& a, long long) const { return static_cast<out_t>(a); }
};

template <class In, class Acc, class Out>
struct ProdOp : SumOp<In, Acc, Out> {
    typedef Acc acc_t; typedef In in_t; typedef NoCtx ctx_t; typedef long long index_t;
    B200_DEVICE acc_t identity() const { return acc_t(1); }
    B200_DEVICE void accumulate(acc_t& a, const ctx_t&, const in_t& v, index_t) const { a = a * static_cast<acc_t>(v); }
    B200_DEVICE acc_t combine(const acc_t& a, const acc_t& b) const { return a * b; }
};

template <class In, class Acc, class Out>
struct MeanOp : SumOp<In, Acc, Out> {
    typedef Acc acc_t; typedef Out out_t;
    B200_DEVICE out_t post(const acc_t& a, long long n) const { return static_cast<out_t>(a / static_cast<acc_t>(n)); }
};

// min / max carry a validity flag in the sign of `idx` so that no sentinel value
// is needed (works for every dtype, NaN included).
template <class V>
struct ValIdx32 { V value; int index; };
template <class V>
struct ValIdx64 { V value; long long index; };
template <class V> B200_DEVICE void shift_index(ValIdx32<V>& a, long long d) { if (a.index >= 0) a.index += static_cast<int>(d); }
template <class V> B200_DEVICE void shift_index(ValIdx64<V>& a, long long d) { if (a.index >= 0) a.index += d; }
template <class V, class I> struct val_idx;
template <class V> struct val_idx<V, int> { typedef ValIdx32<V> type; };
template <class V> struct val_idx<V, long long> { typedef ValIdx64<V> type; };

// kMax: true = max/argmax, false = min/argmin.  kArg: result is the index.
template <class In, class Out, class Index, bool kMax, bool kArg>
struct ExtremumOp {
    typedef In in_t; typedef Out out_t; typedef Index index_t; typedef NoCtx ctx_t;
    typedef typename cmp_type<In>::type C;
    typedef typename val_idx<C, Index>::type acc_t;
    static constexpr bool kWideIndex = sizeof(Index) > 4;

    // does candidate v (at a later index) replace the current value c?
    B200_DEVICE static bool beats(const C& v, const C& c) {
        if (is_floating<In>::value) {
            if (is_nan(c)) return false;
            if (is_nan(v)) return true;
        }
        return kMax ? (v > c) : (v < c);
    }
    B200_DEVICE acc_t identity() const { acc_t a; a.value = C(); a.index = -1; return a; }
    B200_DEVICE ctx_t step(int) const { return ctx_t(); }
    B200_DEVICE void accumulate(acc_t& a, const ctx_t&, const in_t& v, index_t j) const {
        const C c = static_cast<C>(v);
        if (a.index < 0 || beats(c, a.value)) { a.value = c; a.index = j; }
    }
    B200_DEVICE acc_t single(const in_t& v, index_t j) const {
        acc_t a; a.value = static_cast<C>(v); a.index = j; return a;
    }
    B200_DEVICE acc_t combine(const acc_t& a, const acc_t& b) const {
        if (a.index < 0) return b;
        if (b.index < 0) return a;
        const bool nan_a = is_floating<In>::value && is_nan(a.value);
        const bool nan_b = is_floating<In>::value && is_nan(b.value);
        const bool same = (nan_a && nan_b) || (!nan_a && !nan_b && a.value == b.value);
        if (same) return (a.index <= b.index) ? a : b;
        if (nan_a) return a;
        if (nan_b) return b;
        return (kMax ? (a.value > b.value) : (a.value < b.value)) ? a : b;
    }
    B200_DEVICE out_t post(const acc_t& a, long long) const {
        if (kArg) return static_cast<out_t>(a.index);
        return static_cast<out_t>(static_cast<In>(a.value));
    }
};

// ---------------------------------------------------------------------------
// min / max without an index: one NaN-propagating min/max instruction per element
// (FMNMX.NAN / HMNMX2.NAN / IMNMX).  +-inf (floats) and the extreme integer are
// exact identities, so no validity flag is needed.
// ---------------------------------------------------------------------------
template <class T> struct ext_type { typedef T type; };            // register type of a running extremum
template <> struct ext_type<float16> { typedef __half type; };

template <bool kMax> B200_DEVICE float nan_ext(float a, float b) {
    float r;
    if (kMax) asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    else asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
template <bool kMax> B200_DEVICE __half nan_ext(__half a, __half b) { return kMax ? __hmax_nan(a, b) : __hmin_nan(a, b); }
template <bool kMax> B200_DEVICE double nan_ext(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    return kMax ? (a > b ? a : b) : (a < b ? a : b);
}
template <bool kMax> B200_DEVICE bool nan_ext(bool a, bool b) { return kMax ? (a || b) : (a && b); }
template <bool kMax, class T> B200_DEVICE T nan_ext(T a, T b) { return kMax ? (a > b ? a : b) : (a < b ? a : b); }

template <class T, bool kMax> struct ext_identity {
    B200_DEVICE static T get() { return kMax ? std::numeric_limits<T>::lowest() : std::numeric_limits<T>::max(); }
};
template <bool kMax> struct ext_identity<float, kMax> {
    B200_DEVICE static float get() { return kMax ? -__int_as_float(0x7f800000) : __int_as_float(0x7f800000); }
};
template <bool kMax> struct ext_identity<double, kMax> {
    B200_DEVICE static double get() { return kMax ? -__longlong_as_double(0x7ff0000000000000LL) : __longlong_as_double(0x7ff0000000000000LL); }
};
template <bool kMax> struct ext_identity<__half, kMax> {
    B200_DEVICE static __half get() { return __ushort_as_half(kMax ? (unsigned short)0xfc00 : (unsigned short)0x7c00); }
};
template <bool kMax> struct ext_identity<bool, kMax> {
    B200_DEVICE static bool get() { return !kMax; }
};

template <class T> B200_DEVICE typename ext_type<T>::type to_ext(const T& v) { return v; }
template <> B200_DEVICE __half to_ext<float16>(const float16& v) { return v.raw(); }
template <class T> B200_DEVICE T from_ext(const typename ext_type<T>::type& v) { return v; }
template <> B200_DEVICE float16 from_ext<float16>(const __half& v) { return float16(v); }

template <class In, bool kMax>
struct MinMaxOp {
    typedef In in_t; typedef In out_t; typedef long long index_t; typedef NoCtx ctx_t;
    typedef typename ext_type<In>::type acc_t;
    static constexpr bool kWideIndex = false;
    B200_DEVICE acc_t identity() const { return ext_identity<acc_t, kMax>::get(); }
    B200_DEVICE ctx_t step(int) const { return ctx_t(); }
    B200_DEVICE void accumulate(acc_t& a, const ctx_t&, const in_t& v, index_t) const { a = nan_ext<kMax>(a, to_ext(v)); }
    B200_DEVICE acc_t single(const in_t& v, index_t) const { return to_ext(v); }
    B200_DEVICE acc_t combine(const acc_t& a, const acc_t& b) const { return nan_ext<kMax>(a, b); }
    B200_DEVICE out_t post(const acc_t& a, long long) const { return from_ext<In>(a); }
};

// ---------------------------------------------------------------------------
// argmin / argmax.  Across threads a (value, index) pair travels (ExtremumOp's
// combine); inside a thread the streaming state is cheaper (ArgLanes below).
// ---------------------------------------------------------------------------
template <class In, class Index, bool kMax>
struct ArgOp : ExtremumOp<In, long long, Index, kMax, true> {};

template <class T> B200_DEVICE bool ext_is_nan(const T&) { return false; }
template <> B200_DEVICE bool ext_is_nan<float>(const float& v) { return v != v; }
template <> B200_DEVICE bool ext_is_nan<double>(const double& v) { return v != v; }
template <> B200_DEVICE bool ext_is_nan<__half>(const __half& v) { return __hisnan(v); }

// strict "v replaces c"; false when either is NaN
template <bool kMax, class T> B200_DEVICE bool ext_better(const T& v, const T& c) { return kMax ? (v > c) : (v < c); }
template <bool kMax> B200_DEVICE bool ext_better(const __half& v, const __half& c) { return kMax ? __hgt(v, c) : __hlt(v, c); }

template <class T> B200_DEVICE typename cmp_type<T>::type ext_to_cmp(const typename ext_type<T>::type& v) { return v; }
template <> B200_DEVICE float ext_to_cmp<float16>(const __half& v) { return __half2float(v); }
template <> B200_DEVICE int ext_to_cmp<bool>(const bool& v) { return v ? 1 : 0; }

// Streaming state of an arg-reduction inside a thread: V lanes (one per vector element;
// in the COLS layout each lane is its own output column), each a running value and the index
// that set it.  The U elements a lane receives per batch are folded in index order with a
// strict comparison, so the first occurrence wins and the per-element cost is
// compare + 2 selects.  The +-inf / extreme-integer start value never wins a comparison
// against itself: a lane that was never updated reports its first index.  NaN never wins a
// strict comparison either, so it is detected per batch with a NaN-propagating max
// (FMNMX3.NAN) and located out of line -- first NaN of each lane.
template <class T> struct is_f16 { static constexpr bool value = false; };
template <> struct is_f16<float16> { static constexpr bool value = true; };

template <class In, class Index, bool kMax, int U, int V>
struct ArgLanes {
    typedef ArgOp<In, Index, kMax> Op;
    typedef typename Op::acc_t acc_t;
    typedef typename ext_type<In>::type E;
    static constexpr bool kFloat = is_floating<In>::value;
    // float16: the running values live in half2 pairs, so a pair costs ONE packed compare
    // (HSETP2, two predicates), ONE packed max (HMNMX2) and two index selects -- 2 instructions
    // per element instead of ~4 (scalar selects of 16-bit values need PRMT packing).  max(val, e)
    // equals "e > val ? e : val" here because val is never NaN and a NaN e loses both ways.
    static constexpr bool kPacked = is_f16<In>::value && (V % 2 == 0);
    static constexpr int kPairs = kPacked ? V / 2 : 1;
    const Op& op;
    E val[kPacked ? 1 : V];
    __half2 val2[kPairs];
    Index jst[V];        // index (without the lane offset k*ks) of the element that set val; -1 = never updated
    Index nan_j[V];      // first NaN of the lane (full index); -1 = none
    Index j_first, ks_;  // first index this thread saw (lane 0, without k*ks); -1 = saw nothing

    B200_DEVICE E get_val(int k) const {
        if constexpr (kPacked) return (k & 1) ? __high2half(val2[k >> 1]) : __low2half(val2[k >> 1]);
        else return val[k];
    }
    B200_DEVICE void set_val(int k, const E& e) {
        if constexpr (kPacked) {
            if (k & 1) val2[k >> 1] = __halves2half2(__low2half(val2[k >> 1]), e);
            else val2[k >> 1] = __halves2half2(e, __high2half(val2[k >> 1]));
        } else {
            val[k] = e;
        }
    }

    B200_DEVICE explicit ArgLanes(const Op& op_) : op(op_), j_first(-1), ks_(0) {
#pragma unroll
        for (int k = 0; k < V; ++k) { set_val(k, ext_identity<E, kMax>::get()); jst[k] = -1; nan_j[k] = -1; }
    }
    B200_DEVICE void fold(const Pack<In, V> (&v)[U], Index j0, Index us, Index ks) {
        ks_ = ks;
        j_first = j_first < 0 ? j0 : j_first;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const Index ju = j0 + u * us;
            if constexpr (kPacked) {
#pragma unroll
                for (int q = 0; q < kPairs; ++q) {
                    const __half2 e2 = __halves2half2(to_ext(v[u][2 * q]), to_ext(v[u][2 * q + 1]));
                    const bool p0 = ext_better<kMax>(__low2half(e2), __low2half(val2[q]));
                    const bool p1 = ext_better<kMax>(__high2half(e2), __high2half(val2[q]));
                    val2[q] = kMax ? __hmax2(val2[q], e2) : __hmin2(val2[q], e2);
                    jst[2 * q] = p0 ? ju : jst[2 * q];
                    jst[2 * q + 1] = p1 ? ju : jst[2 * q + 1];
                }
            } else {
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    const E e = to_ext(v[u][k]);
                    const bool p = ext_better<kMax>(e, val[k]);
                    val[k] = p ? e : val[k];
                    jst[k] = p ? ju : jst[k];
                }
            }
        }
        if (kFloat) {
            // one NaN test per batch; locating it is rare and kept off the streaming path
            E m = to_ext(v[0][0]);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < V; ++k)
                    if (u | k) m = nan_ext<true>(m, to_ext(v[u][k]));
            if (__builtin_expect(ext_is_nan(m), 0)) {
#pragma unroll
                for (int k = 0; k < V; ++k) {
                    if (nan_j[k] < 0) {
#pragma unroll
                        for (int u = U - 1; u >= 0; --u)
                            if (ext_is_nan(to_ext(v[u][k]))) nan_j[k] = j0 + u * us + k * ks;
                    }
                }
            }
        }
    }
    // stray elements come after every batch of the lane (larger indices)
    B200_DEVICE void fold_one_lane(int k, const In& x, Index j) {
        const E e = to_ext(x);
        if (k == 0) j_first = j_first < 0 ? j : j_first;
        const bool p = ext_better<kMax>(e, get_val(k));
        if (p) set_val(k, e);
        jst[k] = p ? (j - k * ks_) : jst[k];
        if (kFloat && ext_is_nan(e) && nan_j[k] < 0) nan_j[k] = j;
    }
    B200_DEVICE void fold_one(const In& x, Index j) { fold_one_lane(0, x, j); }
    B200_DEVICE acc_t result_lane(int k) {
        acc_t r = op.identity();
        // A lane that was never updated holds only start-value elements: its first one counts.
        // (A lane k > 0 that saw nothing because the thread only took stray elements yields a
        // start-value candidate too; it can only tie with real start-value elements, and the
        // lowest index among those is always a real one.)
        if (j_first >= 0) {
            r.value = ext_to_cmp<In>(get_val(k));
            r.index = (jst[k] >= 0 ? jst[k] : j_first) + k * ks_;
        }
        if (kFloat && nan_j[k] >= 0) {
            acc_t c;
            c.value = ext_to_cmp<In>(to_ext(In(__int_as_float(0x7fc00000))));
            c.index = nan_j[k];
            r = op.combine(r, c);
        }
        return r;
    }
    B200_DEVICE acc_t result() {
        acc_t r = result_lane(0);
#pragma unroll
        for (int k = 1; k < V; ++k) r = op.combine(r, result_lane(k));
        return r;
    }
};

template <class In, class Index, bool kMax>
struct fast_lanes<ArgOp<In, Index, kMax>> { static constexpr bool value = true; };

template <class In, class Index, bool kMax, int U, int V>
struct ThreadAcc<ArgOp<In, Index, kMax>, U, V, true> : ArgLanes<In, Index, kMax, U, V> {
    B200_DEVICE explicit ThreadAcc(const ArgOp<In, Index, kMax>& op_) : ArgLanes<In, Index, kMax, U, V>(op_) {}
};

// Single-pass variance.  Each lane state holds (count, mean, M2) of the
// elements it has seen; a step folds one element into every lane state, and the
// 1/count it needs is computed once per step for all lanes (ctx).  States meet
// through Chan's pairwise merge.  F = float (fp16/fp32 inputs) or double.
template <class F>
struct Moments { F n, mean, m2; };
// what B200_OP_MOMENTS writes per output element: the triple a caller needs to merge shards,
// always in double so that it can go straight into an all-gather
struct MomentTriple { double n, mean, m2; };
// kMode of MomentsOp
constexpr int kMomMean = 0, kMomVar = 1, kMomPair = 2;
template <class F, class Out, int kMode> struct moments_out { typedef Out type; };
template <class F, class Out> struct moments_out<F, Out, kMomPair> { typedef MomentTriple type; };

template <class In, class F, class Out, int kMode>
struct MomentsOp {
    typedef In in_t; typedef Moments<F> acc_t; typedef typename moments_out<F, Out, kMode>::type out_t; typedef long long index_t;
    struct ctx_t { F rcp; F cnt; };
    static constexpr bool kWideIndex = false;
    F ddof;

    B200_DEVICE acc_t identity() const { acc_t a; a.n = F(0); a.mean = F(0); a.m2 = F(0); return a; }
    B200_DEVICE ctx_t step(int count) const { ctx_t c; c.cnt = F(count); c.rcp = F(1) / F(count); return c; }
    B200_DEVICE void accumulate(acc_t& a, const ctx_t& c, const in_t& v, index_t) const {
        const F x = static_cast<F>(v);
        const F d = x - a.mean;
        a.mean = a.mean + d * c.rcp;
        a.m2 = a.m2 + d * (x - a.mean);
        a.n = c.cnt;
    }
    B200_DEVICE acc_t single(const in_t& v, index_t) const {
        acc_t a; a.n = F(1); a.mean = static_cast<F>(v); a.m2 = F(0); return a;
    }
    B200_DEVICE acc_t combine(const acc_t& a, const acc_t& b) const {
        if (a.n == F(0)) return b;
        if (b.n == F(0)) return a;
        acc_t r;
        r.n = a.n + b.n;
        const F d = b.mean - a.mean;
        const F w = b.n / r.n;
        r.mean = a.mean + d * w;
        r.m2 = a.m2 + b.m2 + d * d * a.n * w;
        return r;
    }
    B200_DEVICE out_t post(const acc_t& a, long long n) const {
        if constexpr (kMode == kMomPair) {
            out_t r; r.n = double(n); r.mean = double(a.mean); r.m2 = double(a.m2); return r;
        } else if constexpr (kMode == kMomMean) {
            return static_cast<out_t>(a.mean);
        } else {
            const F div = F(n) - ddof;
            // alpha = 1/max(n-ddof,0), NaN when empty: cupy/_core/_routines_statistics.pyx:585-586
            return static_cast<out_t>(div > F(0) ? a.m2 / div : (a.m2 / F(0)) * F(0));
        }
    }
};

// ---------------------------------------------------------------------------
// Streaming state of the single-pass variance inside a thread: V lanes of (mean, M2) that
// share one element count.  A batch hands every lane U elements; they are folded together:
//     d_u = x_u - mean        S1 = sum d_u        S2 = sum d_u^2
//     n' = n + U    delta = S1 / n'    mean' = mean + delta    M2' = M2 + S2 - delta * S1
// (exactly Chan's merge of the running state with the U-element batch, written around the
// running mean so nothing large is ever squared).  ~3 flops per element, 2V registers of state
// whatever U is -- so U can follow the memory system (bytes in flight), not the register file.
// The very first batch is centred on its own first element.
// ---------------------------------------------------------------------------
// x - c in the accumulation type.  float16 input with a float accumulator uses the sm_100a
// mixed-precision subtract (sub.f32.f16 -> FHADD): the half -> float conversion is free.
template <class In, class F>
B200_DEVICE F centred(const In& x, const F& c) { return static_cast<F>(x) - c; }
template <>
B200_DEVICE float centred<float16, float>(const float16& x, const float& c) {
    float d;
    asm("sub.f32.f16 %0, %1, %2;" : "=f"(d) : "h"(__half_as_ushort(x.raw())), "f"(c));
    return d;
}

template <class In, class F, class Out, int kMode, int U, int V>
struct MomentLanes {
    typedef MomentsOp<In, F, Out, kMode> Op;
    typedef typename Op::acc_t acc_t;
    typedef typename Op::index_t index_t;
    const Op& op;
    F mean[V], m2[V];
    int batches;             // whole batches folded (n = batches * U per lane)
    acc_t tail[V];           // stray elements
    bool any_tail;

    B200_DEVICE explicit MomentLanes(const Op& op_) : op(op_), batches(0), any_tail(false) {
#pragma unroll
        for (int k = 0; k < V; ++k) { mean[k] = F(0); m2[k] = F(0); tail[k] = op.identity(); }
    }
    B200_DEVICE void fold(const Pack<In, V> (&v)[U], index_t, index_t, index_t) {
        const bool first = batches == 0;
        ++batches;
        const F rcp = F(1) / F(batches * U);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const F c = first ? static_cast<F>(v[0][k]) : mean[k];
            F s1 = F(0), s2 = F(0);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const F d = centred<In, F>(v[u][k], c);
                s1 += d;
                s2 += d * d;
            }
            const F delta = s1 * rcp;
            mean[k] = c + delta;
            m2[k] = (m2[k] + s2) - delta * s1;
        }
    }
    B200_DEVICE void fold_one_lane(int k, const In& x, index_t j) {
        any_tail = true;
        tail[k] = op.combine(tail[k], op.single(x, j));
    }
    B200_DEVICE void fold_one(const In& x, index_t j) { fold_one_lane(0, x, j); }
    B200_DEVICE acc_t result_lane(int k) {
        acc_t a;
        a.n = F(batches * U);
        a.mean = mean[k];
        a.m2 = m2[k];
        return any_tail ? op.combine(a, tail[k]) : a;
    }
    B200_DEVICE acc_t result() {
        acc_t r = result_lane(0);
#pragma unroll
        for (int k = 1; k < V; ++k) r = op.combine(r, result_lane(k));
        return r;
    }
};

template <class In, class F, class Out, int kMode>
struct fast_lanes<MomentsOp<In, F, Out, kMode>> { static constexpr bool value = true; };

template <class In, class F, class Out, int kMode, int U, int V>
struct ThreadAcc<MomentsOp<In, F, Out, kMode>, U, V, true> : MomentLanes<In, F, Out, kMode, U, V> {
    B200_DEVICE explicit ThreadAcc(const MomentsOp<In, F, Out, kMode>& op_) : MomentLanes<In, F, Out, kMode, U, V>(op_) {}
};

}  // namespace b200
