// reduce.cu -- C ABI of the prebuilt reductions: validation and dtype dispatch.
// The kernels live in reduce_impl.cuh and are instantiated once per input dtype
// by reduce_types.cu (compiled with -DB200_T=... so the types build in parallel).
#include "common.h"

namespace b200 {

#define B200_DECL(NAME) \
    int reduce_run_##NAME(const b200_reduce_desc_t*, const void*, void*, void*, size_t, cudaStream_t, bool, size_t*, \
                          const b200_peer_exchange_t*);
B200_DECL(f32) B200_DECL(f16) B200_DECL(f64) B200_DECL(i32) B200_DECL(i64) B200_DECL(i8) B200_DECL(u8) B200_DECL(b1)
#undef B200_DECL

static int dispatch(const b200_reduce_desc_t* d, const void* x, void* y, void* ws, size_t wsb,
                    cudaStream_t s, bool query, size_t* need, const b200_peer_exchange_t* pex = nullptr) {
    if (!d) return fail(B200_E_INVALID, "null descriptor");
    if (d->n_reduce <= 0 || d->n_out <= 0 || d->batch <= 0)
        return fail(B200_E_INVALID, "empty reduction: the host handles zero-size arrays");
    if (d->layout != B200_RED_COLS && d->batch != 1) return fail(B200_E_INVALID, "batch is only defined for the COLS layout");
    if (d->layout == B200_RED_FULL && d->n_out != 1) return fail(B200_E_INVALID, "FULL layout has exactly one output");
    switch (d->in_dtype) {
        case B200_TYPE_FLOAT32: return reduce_run_f32(d, x, y, ws, wsb, s, query, need, pex);
        case B200_TYPE_FLOAT16: return reduce_run_f16(d, x, y, ws, wsb, s, query, need, pex);
        case B200_TYPE_FLOAT64: return reduce_run_f64(d, x, y, ws, wsb, s, query, need, pex);
        case B200_TYPE_INT32:   return reduce_run_i32(d, x, y, ws, wsb, s, query, need, pex);
        case B200_TYPE_INT64:   return reduce_run_i64(d, x, y, ws, wsb, s, query, need, pex);
        case B200_TYPE_INT8:    return reduce_run_i8(d, x, y, ws, wsb, s, query, need, pex);
        case B200_TYPE_UINT8:   return reduce_run_u8(d, x, y, ws, wsb, s, query, need, pex);
        case B200_TYPE_BOOL:    return reduce_run_b1(d, x, y, ws, wsb, s, query, need, pex);
        default:
            return fail(B200_E_UNSUPPORTED, "no prebuilt reduction for input dtype %d", d->in_dtype);
    }
}

// Chan merge of `count` (n, mean, M2) triples in index order (rank order: every rank of a
// sharded variance computes the bit-identical result); one thread, the data is 24*count bytes.
__global__ void moments_merge_kernel(const double* __restrict__ t, int count, double ddof, double* __restrict__ out) {
    double n = t[0], mean = t[1], m2 = t[2];
    for (int k = 1; k < count; ++k) {
        const double nb = t[3 * k], mb = t[3 * k + 1], m2b = t[3 * k + 2];
        const double tot = n + nb;
        if (nb == 0.0) continue;
        if (n == 0.0) { n = nb; mean = mb; m2 = m2b; continue; }
        const double d = mb - mean, w = nb / tot;
        mean += d * w;
        m2 += m2b + d * d * n * w;
        n = tot;
    }
    const double div = n - ddof;
    out[0] = div > 0.0 ? m2 / div : (m2 / 0.0) * 0.0;
    out[1] = mean;
    out[2] = n;
}

}  // namespace b200

using namespace b200;

extern "C" __attribute__((visibility("default"))) int b200_moments_merge(const double* triples, int count, double ddof,
                                                                          double* out, void* stream) {
    if (!triples || !out) return fail(B200_E_INVALID, "null data pointer");
    if (count < 1) return fail(B200_E_INVALID, "moments_merge: count %d", count);
    moments_merge_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(triples, count, ddof, out);
    B200_CUDA_TRY(cudaPeekAtLastError());
    return 0;
}

extern "C" __attribute__((visibility("default"))) int b200_reduce_workspace_bytes(const b200_reduce_desc_t* d, size_t* bytes) {
    if (!bytes) return fail(B200_E_INVALID, "null argument");
    *bytes = 0;
    return dispatch(d, nullptr, nullptr, nullptr, 0, nullptr, true, bytes);
}

extern "C" __attribute__((visibility("default"))) int b200_reduce_supported(const b200_reduce_desc_t* d) {
    size_t b = 0;
    return dispatch(d, nullptr, nullptr, nullptr, 0, nullptr, true, &b) == 0 ? 1 : 0;
}

extern "C" __attribute__((visibility("default"))) int b200_reduce_run(const b200_reduce_desc_t* d, const void* x, void* y,
                               void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !y) return fail(B200_E_INVALID, "null data pointer");
    size_t need = 0;
    return dispatch(d, x, y, workspace, workspace_bytes, static_cast<cudaStream_t>(stream), false, &need);
}

extern "C" __attribute__((visibility("default"))) int b200_reduce_run_sharded(const b200_reduce_desc_t* d, const void* x, void* y,
                               void* workspace, size_t workspace_bytes, const b200_peer_exchange_t* exchange,
                               void* stream) {
    if (!x || !y || !exchange) return fail(B200_E_INVALID, "null pointer");
    size_t need = 0;
    return dispatch(d, x, y, workspace, workspace_bytes, static_cast<cudaStream_t>(stream), false, &need, exchange);
}
