// elementwise.cu -- dispatch of the prebuilt ufunc table.  One glue template
// (`ew_kernel`, elementwise_table.cu) is instantiated over the three
// tilers of b200/elementwise.cuh for every (ufunc, dtype loop) in the table; a
// call whose operand dtypes are not in the table returns B200_E_UNSUPPORTED and
// the Python host compiles the same glue around the routine string with NVRTC.
//
// Replaces cupy/_core/_kernel.pyx:1024-1100 (_get_ufunc_kernel: JIT per dtype /
// ndim / contiguity) + cupy/cuda/function.pyx:153-171 (linear_launch, 128-thread
// blocks, one element per thread).
#include <algorithm>
#include "common.h"
#include "elementwise_registry.h"

namespace b200 {

std::map<Key, EwKernels>& registry() {
    static std::map<Key, EwKernels> r;
    return r;
}

static void build_registry() {
    register_ew_group0(); register_ew_group1(); register_ew_group2(); register_ew_group3();
    register_ew_group4(); register_ew_group5(); register_ew_group6();
}

static const EwKernels* find_kernels(int ufunc, int in_dtype, int out_dtype) {
    static const bool once = (build_registry(), true);
    (void)once;
    auto it = registry().find(Key(ufunc, in_dtype, out_dtype));
    return it == registry().end() ? nullptr : &it->second;
}

static int ufunc_nin(int ufunc) {
    switch (ufunc) {
        case B200_UF_COPY: case B200_UF_NEGATIVE: case B200_UF_ABSOLUTE: case B200_UF_SQUARE:
        case B200_UF_SQRT: case B200_UF_EXP: case B200_UF_LOG: return 1;
        case B200_UF_FMA: return 3;
        default: return (ufunc > 0 && ufunc < B200_UF_COUNT) ? 2 : -1;
    }
}

}  // namespace b200

using namespace b200;

extern "C" __attribute__((visibility("default"))) int b200_ufunc_supported(int ufunc, int nin, const int32_t* in_dtypes, int32_t out_dtype) {
    if (ufunc < 0 || ufunc >= B200_UF_COUNT || !in_dtypes) return 0;
    if (nin != ufunc_nin(ufunc)) return 0;
    for (int i = 1; i < nin; ++i)
        if (in_dtypes[i] != in_dtypes[0]) return 0;
    return find_kernels(ufunc, in_dtypes[0], out_dtype) ? 1 : 0;
}

extern "C" __attribute__((visibility("default"))) int b200_ufunc_launch(int ufunc, const b200_ew_plan_t* plan, int nargs,
                                 const b200_operand_t* args, void* stream) {
    if (!plan || !args) return fail(B200_E_INVALID, "null argument");
    const int nin = ufunc_nin(ufunc);
    if (nin < 0) return fail(B200_E_INVALID, "unknown ufunc id %d", ufunc);
    if (nargs != nin + 1) return fail(B200_E_INVALID, "ufunc %d takes %d operands, got %d", ufunc, nin + 1, nargs);
    for (int i = 1; i < nin; ++i)
        if (args[i].dtype != args[0].dtype)
            return fail(B200_E_UNSUPPORTED, "mixed input dtypes have no prebuilt kernel");
    if (args[nin].kind != B200_KIND_ARRAY || !args[nin].is_output)
        return fail(B200_E_INVALID, "last operand must be the output array");
    const EwKernels* k = find_kernels(ufunc, args[0].dtype, args[nin].dtype);
    if (!k) return fail(B200_E_UNSUPPORTED, "no prebuilt kernel for ufunc %d dtypes %d->%d", ufunc, args[0].dtype, args[nin].dtype);
    if (plan->size == 0) return 0;

    EwParams p;
    int st = fill_ew_params(plan, nargs, args, &p);
    if (st) return st;
    DeviceInfo di;
    st = device_info(&di);
    if (st) return st;

    const void* fn = nullptr;
    b200_ew_plan_t eff = *plan;
    int unroll = 1;
    switch (plan->variant) {
        case B200_EW_FLAT:
            // periodic operands (row vector over a dense array) are compiled per call shape by NVRTC
            if (plan->staged_mask) return fail(B200_E_UNSUPPORTED, "FLAT plans with periodic operands have no prebuilt kernel");
            if (plan->vec >= k->vec) { fn = k->flat_v; eff.vec = k->vec; }
            else { fn = k->flat_1; eff.vec = 1; }
            unroll = k->unroll_flat;
            break;
        case B200_EW_ROWWISE: {
            bool all_unit = true;
            for (int a = 0; a < nargs; ++a) {
                if (args[a].kind == B200_KIND_SCALAR) all_unit = false;      // by-value operands: NVRTC folds them
                else if (plan->strides[a][plan->ndim - 1] != b200_dtype_itemsize(args[a].dtype)) all_unit = false;
            }
            if (!all_unit)
                return fail(B200_E_UNSUPPORTED, "ROWWISE plans with broadcast / strided / scalar operands have no prebuilt kernel");
            if (plan->vec >= k->vec && k->vec > 1) { fn = plan->idx32 ? k->row_v32 : k->row_v64; eff.vec = k->vec; }
            else { fn = plan->idx32 ? k->row_132 : k->row_164; eff.vec = 1; }
            unroll = k->unroll_row;
            p.fdiv_chunks = FastDiv(uint32_t(std::min<int64_t>(plan->shape[plan->ndim - 1] / eff.vec, 0xfffffffe)));
            break;
        }
        case B200_EW_TILED:
            fn = k->tiled;
            break;
        case B200_EW_TILED_REG:
            if (!k->tiled_reg) return fail(B200_E_UNSUPPORTED, "no register-tiled prebuilt kernel for ufunc %d", ufunc);
            fn = k->tiled_reg;
            break;
        case B200_EW_TILED_TMA: {
            if (!k->tiled_tma) return fail(B200_E_UNSUPPORTED, "no TMA-tiled prebuilt kernel for ufunc %d", ufunc);
            TileMaps tm;
            st = build_tile_maps(plan, args, &tm);
            if (st) return st;
            int stages;
            unsigned smem;
            tma_ring(plan, &stages, &smem);
            p.tma_stages = stages;
            B200_CUDA_TRY(cudaFuncSetAttribute(k->tiled_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            const unsigned g = ew_grid(plan, 288, 1, di.sm_count);
            void* targs[] = {&p, &tm};
            B200_CUDA_TRY(cudaLaunchKernel(k->tiled_tma, dim3(g), dim3(288), targs, smem, static_cast<cudaStream_t>(stream)));
            return 0;
        }
        default:
            return fail(B200_E_INVALID, "bad plan variant %d", plan->variant);
    }
    const unsigned grid = ew_grid(&eff, kEwThreads, unroll, di.sm_count);
    void* kargs[] = {&p};
    B200_CUDA_TRY(cudaLaunchKernel(fn, dim3(grid), dim3(kEwThreads), kargs, 0, static_cast<cudaStream_t>(stream)));
    return 0;
}
