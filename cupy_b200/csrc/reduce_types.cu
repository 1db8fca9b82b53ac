// reduce_types.cu -- instantiates the prebuilt reductions for ONE input dtype;
// compiled once per dtype: nvcc -DB200_T=float -DB200_TNAME=f32 ...
#include "reduce_impl.cuh"

namespace b200 {
#define B200_CAT2(a, b) a##b
#define B200_CAT(a, b) B200_CAT2(a, b)
int B200_CAT(reduce_run_, B200_TNAME)(const b200_reduce_desc_t* d, const void* x, void* y, void* ws, size_t wsb,
                                      cudaStream_t s, bool query, size_t* need, const b200_peer_exchange_t* pex) {
    return run_for_type<B200_T>(d, x, y, ws, wsb, s, query, need, pex);
}
}  // namespace b200
