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
#include "base.cuh"

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

// Single-pass variance.  Each lane state holds (count, mean, M2) of the
// elements it has seen; a step folds one element into every lane state, and the
// 1/count it needs is computed once per step for all lanes (ctx).  States meet
// through Chan's pairwise merge.  F = float (fp16/fp32 inputs) or double.
template <class F>
struct Moments { F n, mean, m2; };

template <class In, class F, class Out, bool kVar>
struct MomentsOp {
    typedef In in_t; typedef Moments<F> acc_t; typedef Out out_t; typedef long long index_t;
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
        if (!kVar) return static_cast<out_t>(a.mean);
        const F div = F(n) - ddof;
        // alpha = 1/max(n-ddof,0), NaN when empty: cupy/_core/_routines_statistics.pyx:585-586
        return static_cast<out_t>(div > F(0) ? a.m2 / div : (a.m2 / F(0)) * F(0));
    }
};

}  // namespace b200
