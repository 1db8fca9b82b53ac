// b200/base.cuh -- device-side building blocks shared by the prebuilt kernels
// (compiled by nvcc into libcupy_b200.so) and by the NVRTC-compiled user kernels.
// sm_100a only.  Nothing here is copied from the reference; the reference's
// counterpart is cupy/_core/include/cupy/carray.cuh (grid-stride macro, CArray)
// and cupy/_core/include/cupy/float16.cuh.
#pragma once

#ifndef __CUDACC_RTC__
#include <cstddef>
#include <cstdint>
#else
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#endif
#include <cuda_fp16.h>

#define B200_DEVICE __device__ __forceinline__
#define B200_HD __host__ __device__ __forceinline__

namespace b200 {

constexpr int kMaxNdim = 10;   // == B200_MAX_NDIM
constexpr int kMaxArgs = 12;   // == B200_MAX_ARGS
constexpr int kWarp = 32;

// ---------------------------------------------------------------------------
// float16: the C++ type user code strings see for dtype float16 (the reference
// names it `float16`, cupy/_core/_scalar.pyx:_typenames).  Storage is __half,
// arithmetic goes through float (fp32 compute, one rounding on store).
// ---------------------------------------------------------------------------
class float16 {
    __half h_;
public:
    float16() = default;
    B200_DEVICE float16(float v) : h_(__float2half_rn(v)) {}
    B200_DEVICE float16(double v) : h_(__float2half_rn(static_cast<float>(v))) {}
    B200_DEVICE float16(int v) : h_(__float2half_rn(static_cast<float>(v))) {}
    B200_DEVICE float16(unsigned int v) : h_(__float2half_rn(static_cast<float>(v))) {}
    B200_DEVICE float16(long long v) : h_(__float2half_rn(static_cast<float>(v))) {}
    B200_DEVICE float16(unsigned long long v) : h_(__float2half_rn(static_cast<float>(v))) {}
    B200_DEVICE float16(bool v) : h_(__float2half_rn(v ? 1.0f : 0.0f)) {}
    B200_DEVICE explicit float16(__half h) : h_(h) {}
    B200_DEVICE operator float() const { return __half2float(h_); }
    B200_DEVICE __half raw() const { return h_; }
    B200_DEVICE float16& operator+=(float v) { *this = float16(float(*this) + v); return *this; }
    B200_DEVICE float16& operator-=(float v) { *this = float16(float(*this) - v); return *this; }
    B200_DEVICE float16& operator*=(float v) { *this = float16(float(*this) * v); return *this; }
    B200_DEVICE float16& operator/=(float v) { *this = float16(float(*this) / v); return *this; }
};
static_assert(sizeof(float16) == 2, "float16 must be 2 bytes");

B200_DEVICE bool isnan(float16 x) { return __hisnan(x.raw()); }
B200_DEVICE bool isinf(float16 x) { return __hisinf(x.raw()) != 0; }

// ---------------------------------------------------------------------------
// Register packs and 128-bit global memory access.
// ---------------------------------------------------------------------------
template <int BYTES> struct RawVec;
template <> struct RawVec<1>  { typedef uint8_t  type; };
template <> struct RawVec<2>  { typedef uint16_t type; };
template <> struct RawVec<4>  { typedef uint32_t type; };
template <> struct RawVec<8>  { typedef uint2    type; };
template <> struct RawVec<16> { typedef uint4    type; };

template <class T, int N>
struct alignas((sizeof(T) * N >= 16) ? 16 : sizeof(T) * N) Pack {
    T e[N];
    B200_DEVICE T& operator[](int i) { return e[i]; }
    B200_DEVICE const T& operator[](int i) const { return e[i]; }
};

// Loads N consecutive elements; `p` must be aligned to min(16, N*sizeof(T)).
template <class T, int N>
B200_DEVICE void load_pack(Pack<T, N>& dst, const T* __restrict__ p) {
    constexpr int total = int(sizeof(T)) * N;
    constexpr int chunk = total >= 16 ? 16 : total;
    typedef typename RawVec<chunk>::type V;
    static_assert(total % chunk == 0, "pack size must be a multiple of the vector width");
#pragma unroll
    for (int c = 0; c < total / chunk; ++c)
        reinterpret_cast<V*>(&dst)[c] = reinterpret_cast<const V*>(p)[c];
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256) for packs of exactly 32 bytes whose address is
// 32-byte aligned: a thread that owns 32 contiguous bytes touches whole sectors with ONE instruction
// instead of two half-sector ones.
template <class T, int N>
B200_DEVICE void load_pack256(Pack<T, N>& dst, const T* __restrict__ p) {
    static_assert(sizeof(T) * N == 32, "load_pack256 moves exactly 32 bytes");
    uint64_t a, b, c, d;
    asm("ld.global.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
    uint64_t* w = reinterpret_cast<uint64_t*>(&dst);
    w[0] = a; w[1] = b; w[2] = c; w[3] = d;
}
template <class T, int N>
B200_DEVICE void store_pack256(T* __restrict__ p, const Pack<T, N>& src) {
    static_assert(sizeof(T) * N == 32, "store_pack256 moves exactly 32 bytes");
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&src);
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(w[0]), "l"(w[1]), "l"(w[2]), "l"(w[3]) : "memory");
}

template <class T, int N>
B200_DEVICE void store_pack(T* __restrict__ p, const Pack<T, N>& src) {
    constexpr int total = int(sizeof(T)) * N;
    constexpr int chunk = total >= 16 ? 16 : total;
    typedef typename RawVec<chunk>::type V;
#pragma unroll
    for (int c = 0; c < total / chunk; ++c)
        reinterpret_cast<V*>(p)[c] = reinterpret_cast<const V*>(&src)[c];
}

// ---------------------------------------------------------------------------
// Division by a run-time constant without a divide instruction: the launcher
// decomposes ONE linear index per vector (not per element), and this keeps even
// that off the slow 64-bit IDIV path when sizes fit 32 bits.
// ---------------------------------------------------------------------------
struct FastDiv {
    uint32_t d, magic, shift;
    FastDiv() = default;
    B200_HD explicit FastDiv(uint32_t div) : d(div) {
        if (div <= 1) { magic = 0; shift = 0; d = div ? div : 1; return; }
        uint32_t s = 0;
        while ((1ull << s) < div) ++s;
        shift = s;
        uint64_t m = ((1ull << 32) * ((1ull << s) - div)) / div + 1;
        magic = static_cast<uint32_t>(m);
    }
    B200_DEVICE uint32_t div(uint32_t n) const {
        if (d == 1) return n;
#ifdef __CUDA_ARCH__
        uint32_t t = __umulhi(n, magic);
#else
        uint32_t t = static_cast<uint32_t>((uint64_t(n) * magic) >> 32);
#endif
        return static_cast<uint32_t>((uint64_t(t) + n) >> shift);
    }
    B200_DEVICE void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
        q = div(n);
        r = n - q * d;
    }
};

// ---------------------------------------------------------------------------
// Warp / block combine.  `T` can be any trivially copyable struct (value/index
// pairs, Welford triples): it is shuffled in 32-bit words.
// ---------------------------------------------------------------------------
template <class T>
B200_DEVICE T shfl_down_any(const T& v, int delta) {
    constexpr int words = (sizeof(T) + 3) / 4;
    union U { T t; uint32_t w[words]; B200_DEVICE U() {} };
    U in, out;
    in.t = v;
#pragma unroll
    for (int i = 0; i < words; ++i) out.w[i] = __shfl_down_sync(0xffffffffu, in.w[i], delta);
    return out.t;
}

template <class T>
B200_DEVICE T shfl_up_any(const T& v, int delta) {
    constexpr int words = (sizeof(T) + 3) / 4;
    union U { T t; uint32_t w[words]; B200_DEVICE U() {} };
    U in, out;
    in.t = v;
#pragma unroll
    for (int i = 0; i < words; ++i) out.w[i] = __shfl_up_sync(0xffffffffu, in.w[i], delta);
    return out.t;
}

template <class T>
B200_DEVICE T shfl_any(const T& v, int lane) {
    constexpr int words = (sizeof(T) + 3) / 4;
    union U { T t; uint32_t w[words]; B200_DEVICE U() {} };
    U in, out;
    in.t = v;
#pragma unroll
    for (int i = 0; i < words; ++i) out.w[i] = __shfl_sync(0xffffffffu, in.w[i], lane);
    return out.t;
}

// Result valid in lane 0.  Lower lanes are always the left operand, so
// order-sensitive ops (first-NaN / lowest-index ties) stay deterministic.
template <class Op, class T>
B200_DEVICE T warp_combine(const Op& op, T v) {
#pragma unroll
    for (int d = kWarp / 2; d > 0; d >>= 1) {
        T o = shfl_down_any(v, d);
        v = op.combine(v, o);
    }
    return v;
}

// Block-wide combine through shared memory; result valid in thread 0.
// `smem` must hold (blockDim.x / 32) elements of T.  Contains __syncthreads().
template <class Op, class T>
B200_DEVICE T block_combine(const Op& op, T v, T* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    v = warp_combine(op, v);
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        T w = lane < nwarps ? smem[lane] : op.identity();
        w = warp_combine(op, w);
        v = w;
    }
    __syncthreads();
    return v;
}

// ---------------------------------------------------------------------------
// Release / acquire flags for single-pass grid protocols (ticket combine,
// decoupled look-back).
// ---------------------------------------------------------------------------
B200_DEVICE void st_release_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
B200_DEVICE uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
B200_DEVICE void st_release_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
B200_DEVICE uint64_t ld_acquire_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

}  // namespace b200
