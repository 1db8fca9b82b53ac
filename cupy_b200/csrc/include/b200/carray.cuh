// b200/carray.cuh -- the objects user code strings see inside ElementwiseKernel /
// ReductionKernel bodies: `raw` arrays (CArray) and the loop indexer (`_ind`).
// Source-compatible with the names and methods documented for the reference's
// kernels (cupy/_core/include/cupy/carray.cuh:228-458 CArray, :518-620 CIndexer)
// -- size(), shape(), strides(), operator[] by linear index or by index array --
// but written from scratch on top of the launcher's RawView descriptor.
#pragma once
#include "base.cuh"

namespace b200 {

struct RawView {
    char*   data;
    int64_t size;
    int32_t ndim;
    int32_t pad_;
    int64_t shape[kMaxNdim];
    int64_t strides[kMaxNdim];   // bytes
};

// the `raw` operands of a kernel (one view each, in operand order), passed as the second kernel parameter;
// generated kernels declare exactly as many as they use (+ one trailing shape-only view carrying the
// un-collapsed loop shape when the user code reads `_ind` of a reduce_dims=False kernel)
template <int N> struct RawPackN { RawView v[N > 0 ? N : 1]; };
typedef RawPackN<4> RawPack;

}  // namespace b200

// The template signature mirrors the reference so user preambles that spell the
// type out (rare) still compile; c_contiguous / use_32bit are hints only.
template <typename T, int _ndim, bool _c_contiguous = false, bool _use_32bit = false>
class CArray {
public:
    static const int ndim = _ndim;
    typedef ptrdiff_t index_t;

private:
    T* data_;
    ptrdiff_t size_;
    ptrdiff_t shape_[_ndim > 0 ? _ndim : 1];
    ptrdiff_t strides_[_ndim > 0 ? _ndim : 1];

public:
    __device__ explicit CArray(const b200::RawView& v)
        : data_(reinterpret_cast<T*>(v.data)), size_(v.size) {
#pragma unroll
        for (int d = 0; d < _ndim; ++d) {
            shape_[d] = v.shape[d];
            strides_[d] = v.strides[d];
        }
    }
    __device__ ptrdiff_t size() const { return size_; }
    __device__ const ptrdiff_t* shape() const { return shape_; }
    __device__ const ptrdiff_t* strides() const { return strides_; }
    __device__ T* data() const { return data_; }

    template <typename Int>
    __device__ T& operator[](const Int (&idx)[_ndim > 0 ? _ndim : 1]) {
        return const_cast<T&>(const_cast<const CArray&>(*this)[idx]);
    }
    template <typename Int>
    __device__ const T& operator[](const Int (&idx)[_ndim > 0 ? _ndim : 1]) const {
        const char* p = reinterpret_cast<const char*>(data_);
#pragma unroll
        for (int d = 0; d < _ndim; ++d) p += ptrdiff_t(idx[d]) * strides_[d];
        return *reinterpret_cast<const T*>(p);
    }
    __device__ T& operator[](const ptrdiff_t* idx) {
        return const_cast<T&>(const_cast<const CArray&>(*this)[idx]);
    }
    __device__ const T& operator[](const ptrdiff_t* idx) const {
        const char* p = reinterpret_cast<const char*>(data_);
#pragma unroll
        for (int d = 0; d < _ndim; ++d) p += idx[d] * strides_[d];
        return *reinterpret_cast<const T*>(p);
    }
    // linear C-order index
    __device__ T& operator[](ptrdiff_t i) {
        return const_cast<T&>(const_cast<const CArray&>(*this)[i]);
    }
    __device__ const T& operator[](ptrdiff_t i) const {
        if (_c_contiguous) return data_[i];
        const char* p = reinterpret_cast<const char*>(data_);
#pragma unroll
        for (int d = _ndim - 1; d > 0; --d) {
            const ptrdiff_t q = i / shape_[d];
            p += (i - q * shape_[d]) * strides_[d];
            i = q;
        }
        if (_ndim > 0) p += i * strides_[0];
        return *reinterpret_cast<const T*>(p);
    }
};

template <int _ndim, bool _use_32bit = false>
class CIndexer {
public:
    static const int ndim = _ndim;

private:
    ptrdiff_t size_;
    ptrdiff_t shape_[_ndim > 0 ? _ndim : 1];
    ptrdiff_t index_[_ndim > 0 ? _ndim : 1];

public:
    __device__ CIndexer(ptrdiff_t size, const int64_t* shape) : size_(size) {
#pragma unroll
        for (int d = 0; d < _ndim; ++d) {
            shape_[d] = shape[d];
            index_[d] = 0;
        }
    }
    __device__ ptrdiff_t size() const { return size_; }
    __device__ const ptrdiff_t* shape() const { return shape_; }
    __device__ const ptrdiff_t* get() const { return index_; }
    __device__ void set(ptrdiff_t i) {
#pragma unroll
        for (int d = _ndim - 1; d > 0; --d) {
            const ptrdiff_t q = i / shape_[d];
            index_[d] = i - q * shape_[d];
            i = q;
        }
        if (_ndim > 0) index_[0] = i;
    }
};

// Size-only stand-in for `_in_ind` / `_out_ind` inside reduction bodies
// (user expressions only ever call .size() on them).
struct CSizeIndexer {
    ptrdiff_t size_;
    __device__ ptrdiff_t size() const { return size_; }
};

// math.h constants user code strings spell (NVRTC has no host <cmath>; the reference gets them from
// cupy/_core/include/cupy/math_constants.h through its own headers)
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_E
#define M_E 2.7182818284590452354
#endif

// `_floor_divide(x, y)`: the helper the reference's `floor_divide` / `remainder` / `divmod` routine strings
// (and user kernels) call -- cupy/_core/include/cupy/carray.cuh:671-700.  Integers round toward minus
// infinity and a zero divisor yields 0; floats are floor(x / y).
template <class I>
__device__ inline I _b200_floor_div_signed(I x, I y) {
    if (y == 0) return 0;
    const I q = x / y;
    const I r = x - q * y;
    return (r != 0 && ((r < 0) != (y < 0))) ? q - 1 : q;
}
__device__ inline int _floor_divide(int x, int y) { return _b200_floor_div_signed<int>(x, y); }
__device__ inline long long _floor_divide(long long x, long long y) { return _b200_floor_div_signed<long long>(x, y); }
__device__ inline unsigned _floor_divide(unsigned x, unsigned y) { return y == 0 ? 0u : x / y; }
__device__ inline unsigned long long _floor_divide(unsigned long long x, unsigned long long y) {
    return y == 0 ? 0ull : x / y;
}
__device__ inline float _floor_divide(float x, float y) { return floorf(x / y); }
__device__ inline double _floor_divide(double x, double y) { return floor(x / y); }

#ifndef CUPY_FOR
#define CUPY_FOR(i, n)                                                        \
    for (ptrdiff_t i = static_cast<ptrdiff_t>(blockIdx.x) * blockDim.x + threadIdx.x; \
         i < (n); i += static_cast<ptrdiff_t>(blockDim.x) * gridDim.x)
#endif
