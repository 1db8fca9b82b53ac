// b200/ufunc_ops.cuh -- functors of the prebuilt ufunc table.  Each mirrors the
// routine string of the reference's ufunc loop of the same name
// (cupy/_core/_routines_math.pyx:878-1178 arithmetic, cupy/_math/explog.py,
// cupy/_math/misc.py, cupy/_core/_ufuncs.py:7-13 elementwise_copy): same result
// type, float16 computed in float and rounded once, NaN-propagating
// maximum/minimum.
#pragma once
#include "base.cuh"

namespace b200 {

// computation domain of a storage type
template <class T> struct compute_type { typedef T type; };
template <> struct compute_type<float16> { typedef float type; };

#define B200_UNARY(NAME, EXPR)                                                    \
    template <class TI, class TO = TI>                                            \
    struct NAME {                                                                 \
        static constexpr int nin = 1;                                             \
        typedef TI in0_t; typedef TI in1_t; typedef TI in2_t; typedef TO out_t;  \
        B200_DEVICE static TO apply(const TI& a0, const TI&, const TI&) {         \
            typedef typename compute_type<TI>::type C;                            \
            const C x = static_cast<C>(a0);                                       \
            return static_cast<TO>(EXPR);                                         \
        }                                                                         \
    };

#define B200_BINARY(NAME, EXPR)                                                   \
    template <class TI, class TO = TI>                                            \
    struct NAME {                                                                 \
        static constexpr int nin = 2;                                             \
        typedef TI in0_t; typedef TI in1_t; typedef TI in2_t; typedef TO out_t;  \
        B200_DEVICE static TO apply(const TI& a0, const TI& a1, const TI&) {      \
            typedef typename compute_type<TI>::type C;                            \
            const C x = static_cast<C>(a0), y = static_cast<C>(a1);               \
            return static_cast<TO>(EXPR);                                         \
        }                                                                         \
    };

template <class C> B200_DEVICE C abs_of(C x) { return x < C(0) ? C(-x) : x; }
template <> B200_DEVICE float abs_of<float>(float x) { return fabsf(x); }
template <> B200_DEVICE double abs_of<double>(double x) { return fabs(x); }

template <class C> B200_DEVICE C max_nan(C x, C y) { return x > y ? x : y; }
template <class C> B200_DEVICE C min_nan(C x, C y) { return x < y ? x : y; }
// out0 = (isnan(in0) | isnan(in1)) ? in0 + in1 : max(in0, in1)   (reference: _math/misc.py maximum)
template <> B200_DEVICE float max_nan<float>(float x, float y) { return (x != x || y != y) ? x + y : fmaxf(x, y); }
template <> B200_DEVICE double max_nan<double>(double x, double y) { return (x != x || y != y) ? x + y : fmax(x, y); }
template <> B200_DEVICE float min_nan<float>(float x, float y) { return (x != x || y != y) ? x + y : fminf(x, y); }
template <> B200_DEVICE double min_nan<double>(double x, double y) { return (x != x || y != y) ? x + y : fmin(x, y); }

template <class C> B200_DEVICE C sqrt_of(C x);
template <> B200_DEVICE float sqrt_of<float>(float x) { return sqrtf(x); }
template <> B200_DEVICE double sqrt_of<double>(double x) { return sqrt(x); }
template <class C> B200_DEVICE C exp_of(C x);
template <> B200_DEVICE float exp_of<float>(float x) { return expf(x); }
template <> B200_DEVICE double exp_of<double>(double x) { return exp(x); }
template <class C> B200_DEVICE C log_of(C x);
template <> B200_DEVICE float log_of<float>(float x) { return logf(x); }
template <> B200_DEVICE double log_of<double>(double x) { return log(x); }

B200_UNARY(CopyF, x)
B200_UNARY(NegativeF, -x)
B200_UNARY(AbsoluteF, abs_of<C>(x))
B200_UNARY(SquareF, x * x)
B200_UNARY(SqrtF, sqrt_of<C>(x))
B200_UNARY(ExpF, exp_of<C>(x))
B200_UNARY(LogF, log_of<C>(x))
B200_BINARY(AddF, x + y)
B200_BINARY(SubtractF, x - y)
B200_BINARY(MultiplyF, x * y)
B200_BINARY(TrueDivideF, x / y)
B200_BINARY(MaximumF, max_nan<C>(x, y))
B200_BINARY(MinimumF, min_nan<C>(x, y))

template <class C> B200_DEVICE C fma_of(C x, C y, C z);
template <> B200_DEVICE float fma_of<float>(float x, float y, float z) { return fmaf(x, y, z); }
template <> B200_DEVICE double fma_of<double>(double x, double y, double z) { return fma(x, y, z); }

template <class TI, class TO = TI>
struct FmaF {
    static constexpr int nin = 3;
    typedef TI in0_t; typedef TI in1_t; typedef TI in2_t; typedef TO out_t;
    B200_DEVICE static TO apply(const TI& a0, const TI& a1, const TI& a2) {
        typedef typename compute_type<TI>::type C;
        return static_cast<TO>(fma_of<C>(static_cast<C>(a0), static_cast<C>(a1), static_cast<C>(a2)));
    }
};

// bool addition is logical or, bool multiplication logical and (reference:
// create_arithmetic('add', '+', '|'), ('multiply', '*', '&'))
template <> struct AddF<bool, bool> {
    static constexpr int nin = 2;
    typedef bool in0_t; typedef bool in1_t; typedef bool in2_t; typedef bool out_t;
    B200_DEVICE static bool apply(const bool& a, const bool& b, const bool&) { return a | b; }
};
template <> struct MultiplyF<bool, bool> {
    static constexpr int nin = 2;
    typedef bool in0_t; typedef bool in1_t; typedef bool in2_t; typedef bool out_t;
    B200_DEVICE static bool apply(const bool& a, const bool& b, const bool&) { return a & b; }
};

}  // namespace b200
