// launcher.cu -- host side of the elementwise launcher: collapse dims, classify
// the call by stride pattern, choose vector width / index width, and build the
// kernel parameter block.  Pure host code (no kernels): the plan is testable
// without a GPU.
//
// Reference behaviour being replaced: cupy/_core/_kernel.pyx:360-461
// (_reduce_dims / _reduced_view_core: merge adjacent dims; all-contiguous -> 1-D;
// a 2-D call with any non-contiguous operand is left alone) and the per-kernel
// template flags (ndim, c_contiguous, index_32_bits) of _ArgInfo (:189-338).
// Here collapsing always runs (2-D included), size-1 dims are dropped, and the
// result is one of three kernel variants instead of a family of JIT templates.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.h"

namespace b200 {

std::string& last_error() {
    static thread_local std::string s;
    return s;
}

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

int device_info(DeviceInfo* out) {
    static std::mutex mu;
    static DeviceInfo cache[64];
    static bool have[64] = {false};
    int dev = 0;
    B200_CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 0 || dev >= 64) return fail(B200_E_INVALID, "device id %d out of range", dev);
    if (!have[dev]) {
        int v = 0;
        B200_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        cache[dev].sm_count = v;
        B200_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
        cache[dev].cc_major = v;
        B200_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
        cache[dev].cc_minor = v;
        B200_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev));
        cache[dev].l2_bytes = size_t(v);
        have[dev] = true;
    }
    *out = cache[dev];
    return 0;
}

static inline bool is_array(const b200_operand_t& a) { return a.kind == B200_KIND_ARRAY; }

int fill_ew_params(const b200_ew_plan_t* plan, int nargs, const b200_operand_t* args, EwParams* out) {
    if (nargs != plan->nargs || nargs > kMaxArgs) return fail(B200_E_INVALID, "operand count does not match the plan");
    EwParams& p = *out;
    std::memset(&p, 0, sizeof(p));
    p.size = plan->size;
    p.ndim = plan->ndim;
    p.tile_axis = plan->tile_axis;
    p.staged_mask = plan->staged_mask;
    int64_t cs = 1;
    for (int d = plan->ndim - 1; d >= 0; --d) {
        p.shape[d] = plan->shape[d];
        p.cstride[d] = cs;
        cs *= plan->shape[d];
        const uint64_t s = uint64_t(plan->shape[d]);
        p.fdiv[d] = FastDiv(s < 0xffffffffull ? uint32_t(s) : 1u);
    }
    for (int a = 0; a < nargs; ++a) {
        if (args[a].kind == B200_KIND_SCALAR) {
            p.scalar_mask |= 1u << a;
            p.arg[a].scalar[0] = args[a].scalar[0];
            p.arg[a].scalar[1] = args[a].scalar[1];
        } else {
            p.arg[a].ptr = static_cast<char*>(args[a].data);
            if (is_array(args[a]))
                for (int d = 0; d < plan->ndim; ++d) p.arg[a].strides[d] = plan->strides[a][d];
        }
    }
    return 0;
}

unsigned ew_grid(const b200_ew_plan_t* plan, int threads, int unroll, int sm_count) {
    int64_t blocks;
    if (plan->variant == B200_EW_TILED) {
        const int64_t ni = plan->shape[plan->tile_axis], no = plan->shape[plan->ndim - 1];
        blocks = ((ni + 31) / 32) * ((no + 31) / 32) * (plan->size / (ni * no));
        return unsigned(std::min<int64_t>(blocks, 0x7fffffff));
    }
    const int64_t work = plan->size / std::max(1, plan->vec);          // vectors
    const int64_t per_block = int64_t(threads) * unroll;
    blocks = (work + per_block - 1) / per_block;
    // persistent: enough resident blocks to cover HBM latency, then grid-stride.
    // plan->reserved bits 8..23: blocks per SM override (tuning sweeps)
    const int per_sm = ((plan->reserved >> 8) & 0xffff) ? int((plan->reserved >> 8) & 0xffff) : (2048 / threads) * 4;
    const int64_t cap = int64_t(sm_count) * per_sm;
    return unsigned(std::max<int64_t>(1, std::min(blocks, cap)));
}

}  // namespace b200

using namespace b200;

extern "C" __attribute__((visibility("default"))) int b200_abi_version(void) { return B200_ABI_VERSION; }

extern "C" __attribute__((visibility("default"))) const char* b200_last_error_string(void) { return last_error().c_str(); }

extern "C" __attribute__((visibility("default"))) int b200_dtype_itemsize(int dtype) { return dtype_size(dtype); }

extern "C" __attribute__((visibility("default"))) int b200_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes) {
    DeviceInfo di;
    int st = device_info(&di);
    if (st) return st;
    if (sm_count) *sm_count = di.sm_count;
    if (cc_major) *cc_major = di.cc_major;
    if (cc_minor) *cc_minor = di.cc_minor;
    if (l2_bytes) *l2_bytes = di.l2_bytes;
    return 0;
}

extern "C" __attribute__((visibility("default"))) int b200_ew_plan(int nargs, const b200_operand_t* args, b200_ew_plan_t* plan) {
    if (!args || !plan) return fail(B200_E_INVALID, "null argument");
    if (nargs <= 0 || nargs > B200_MAX_ARGS) return fail(B200_E_INVALID, "operand count %d not in 1..%d", nargs, B200_MAX_ARGS);
    std::memset(plan, 0, sizeof(*plan));
    plan->nargs = nargs;

    // ---- the loop shape: every ARRAY operand carries it (host broadcast them)
    int first = -1;
    for (int a = 0; a < nargs; ++a) {
        if (dtype_size(args[a].dtype) == 0) return fail(B200_E_INVALID, "operand %d: bad dtype id %d", a, args[a].dtype);
        if (!is_array(args[a])) continue;
        if (args[a].ndim < 0) return fail(B200_E_INVALID, "operand %d: negative ndim", a);
        if (first < 0) first = a;
        else if (args[a].ndim != args[first].ndim) return fail(B200_E_INVALID, "operand %d: ndim differs from operand %d", a, first);
    }
    if (first < 0) return fail(B200_E_INVALID, "no array operand: loop size is undecided");
    const int nd0 = args[first].ndim;
    int64_t size = 1;
    for (int d = 0; d < nd0 && d < 64; ++d) {
        if (d < B200_MAX_NDIM) {
            const int64_t s = args[first].shape[d];
            if (s < 0) return fail(B200_E_INVALID, "negative extent");
            for (int a = 0; a < nargs; ++a)
                if (is_array(args[a]) && args[a].shape[d] != s)
                    return fail(B200_E_INVALID, "operand %d: shape differs at dim %d (operands must be broadcast by the host)", a, d);
            size *= s;
        }
    }
    if (nd0 > B200_MAX_NDIM) return fail(B200_E_UNSUPPORTED, "rank %d exceeds %d (collapse on the host first)", nd0, B200_MAX_NDIM);
    plan->size = size;
    if (size == 0) { plan->variant = B200_EW_FLAT; plan->ndim = 1; plan->vec = 1; plan->idx32 = 1; return 0; }

    // ---- drop extent-1 dims, then merge (d, d+1) whenever EVERY array operand
    //      satisfies stride[d] == stride[d+1] * shape[d+1]  (0 == 0*n covers broadcasts)
    int64_t shape[B200_MAX_NDIM];
    int64_t st[B200_MAX_ARGS][B200_MAX_NDIM];
    int nd = 0;
    for (int d = 0; d < nd0; ++d) {
        if (args[first].shape[d] == 1) continue;
        shape[nd] = args[first].shape[d];
        for (int a = 0; a < nargs; ++a) st[a][nd] = is_array(args[a]) ? args[a].strides[d] : 0;
        ++nd;
    }
    if (nd == 0) {  // a single element
        nd = 1; shape[0] = 1;
        for (int a = 0; a < nargs; ++a) st[a][0] = is_array(args[a]) ? dtype_size(args[a].dtype) : 0;
    }
    int w = 0;  // write cursor: dims [0..w] are final so far
    for (int d = 1; d < nd; ++d) {
        bool merge = true;
        for (int a = 0; a < nargs && merge; ++a)
            if (is_array(args[a]) && st[a][w] != st[a][d] * shape[d]) merge = false;
        if (merge) {
            shape[w] *= shape[d];
            for (int a = 0; a < nargs; ++a) st[a][w] = st[a][d];
        } else {
            ++w;
            shape[w] = shape[d];
            for (int a = 0; a < nargs; ++a) st[a][w] = st[a][d];
        }
    }
    nd = w + 1;
    plan->ndim = nd;
    for (int d = 0; d < nd; ++d) {
        plan->shape[d] = shape[d];
        for (int a = 0; a < nargs; ++a) plan->strides[a][d] = st[a][d];
    }

    // ---- index width
    bool idx32 = size < (int64_t(1) << 31);
    int max_item = 1;
    for (int a = 0; a < nargs; ++a) {
        if (!is_array(args[a])) continue;
        max_item = std::max(max_item, dtype_size(args[a].dtype));
        int64_t span = 0;
        for (int d = 0; d < nd; ++d) span += std::llabs(st[a][d]) * (shape[d] - 1);
        if (span >= (int64_t(1) << 31)) idx32 = false;
    }
    plan->idx32 = idx32 ? 1 : 0;

    // ---- classify
    const int last = nd - 1;
    bool flat = (nd == 1);
    for (int a = 0; a < nargs && flat; ++a)
        if (is_array(args[a]) && st[a][0] != dtype_size(args[a].dtype)) flat = false;

    int vec = std::max(1, 16 / max_item);
    if (flat) {
        plan->variant = B200_EW_FLAT;
        for (; vec > 1; vec >>= 1) {
            bool ok = true;
            for (int a = 0; a < nargs && ok; ++a)
                if (is_array(args[a]) && (reinterpret_cast<uintptr_t>(args[a].data) % (uintptr_t(vec) * dtype_size(args[a].dtype))) != 0) ok = false;
            if (ok) break;
        }
        plan->vec = vec;
        return 0;
    }

    // TILED: an input that is unit-stride along another dim while every output
    // is unit-stride along the innermost dim
    if (nd >= 2) {
        bool outs_ok = true;
        for (int a = 0; a < nargs; ++a)
            if (is_array(args[a]) && args[a].is_output && st[a][last] != dtype_size(args[a].dtype)) outs_ok = false;
        int axis = -1;
        if (outs_ok && shape[last] >= 16) {
            for (int a = 0; a < nargs && axis < 0; ++a) {
                if (!is_array(args[a]) || args[a].is_output) continue;
                const int isz = dtype_size(args[a].dtype);
                if (st[a][last] == isz || st[a][last] == 0) continue;
                for (int d = 0; d < last; ++d)
                    if (st[a][d] == isz && shape[d] >= 16) { axis = d; break; }
            }
        }
        if (axis >= 0) {
            uint32_t mask = 0;
            for (int a = 0; a < nargs; ++a) {
                if (!is_array(args[a]) || args[a].is_output) continue;
                const int isz = dtype_size(args[a].dtype);
                if (st[a][axis] == isz && st[a][last] != isz && st[a][last] != 0) mask |= 1u << a;
            }
            plan->variant = B200_EW_TILED;
            plan->tile_axis = axis;
            plan->staged_mask = mask;
            plan->vec = 1;
            return 0;
        }
    }

    plan->variant = B200_EW_ROWWISE;
    for (; vec > 1; vec >>= 1) {
        bool ok = (shape[last] % vec) == 0;
        for (int a = 0; a < nargs && ok; ++a) {
            if (!is_array(args[a])) continue;
            const int isz = dtype_size(args[a].dtype);
            if (st[a][last] != isz) continue;      // broadcast / strided operands use scalar accesses
            const uintptr_t al = uintptr_t(vec) * isz;
            if (reinterpret_cast<uintptr_t>(args[a].data) % al) ok = false;
            for (int d = 0; d < last && ok; ++d)
                if (std::llabs(st[a][d]) % int64_t(al)) ok = false;
        }
        if (ok) break;
    }
    plan->vec = vec;
    return 0;
}
