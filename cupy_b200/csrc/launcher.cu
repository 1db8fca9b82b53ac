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
#include "tma_host.h"

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
    {
        const uint64_t chunks = uint64_t(plan->shape[plan->ndim - 1]) / uint64_t(std::max(1, plan->vec));
        p.fdiv_chunks = FastDiv(chunks && chunks < 0xffffffffull ? uint32_t(chunks) : 1u);
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

// ---- TILED_TMA geometry (mirrors TmaTileTiler in b200/elementwise.cuh)
void tma_tile_geometry(const b200_ew_plan_t* plan, int* esz, int* tile_i, int* tile_o) {
    const int e = int(plan->reserved >> 24);          // item size, set by the planner
    *esz = e;
    *tile_i = 128 / e;
    *tile_o = 32 * (16 / e);
}

static int tma_nstaged(const b200_ew_plan_t* plan) { return __builtin_popcount(plan->staged_mask); }

int tma_blocks_per_sm(const b200_ew_plan_t* plan) {
    int e, ti, to;
    tma_tile_geometry(plan, &e, &ti, &to);
    const int tile_bytes = to * 128 * tma_nstaged(plan);
    return tile_bytes * 3 * 2 + 4096 <= 200 * 1024 ? 2 : 1;
}

// ring depth and dynamic shared memory of one block
void tma_ring(const b200_ew_plan_t* plan, int* stages, unsigned* smem_bytes) {
    int e, ti, to;
    tma_tile_geometry(plan, &e, &ti, &to);
    const int tile_bytes = to * 128 * tma_nstaged(plan);
    const int budget = 200 * 1024 / tma_blocks_per_sm(plan) - 2048;
    int s = (plan->reserved & 0xff) ? int(plan->reserved & 0xff) : budget / tile_bytes;
    s = std::max(2, std::min(s, 8));
    while (s > 2 && s * tile_bytes > 220 * 1024) --s;
    *stages = s;
    *smem_bytes = unsigned(s * tile_bytes + 2048);
}

// One 5-D tensor map per staged operand: dims (tile axis, innermost loop dim,
// then the batch dims innermost-first), box = one tile, 128-byte swizzle.
int build_tile_maps(const b200_ew_plan_t* plan, const b200_operand_t* args, TileMaps* out) {
    int e, ti, to;
    tma_tile_geometry(plan, &e, &ti, &to);
    const int nd = plan->ndim, last = nd - 1, ax = plan->tile_axis;
    int slot = 0;
    for (int a = 0; a < plan->nargs; ++a) {
        if (!((plan->staged_mask >> a) & 1u)) continue;
        uint64_t dims[5] = {1, 1, 1, 1, 1}, strides[4] = {16, 16, 16, 16};
        uint32_t box[5] = {uint32_t(ti), uint32_t(to), 1, 1, 1};
        dims[0] = uint64_t(plan->shape[ax]);
        dims[1] = uint64_t(plan->shape[last]);
        strides[0] = uint64_t(plan->strides[a][last]);
        int n = 2;
        for (int d = nd - 2; d >= 0; --d) {
            if (d == ax) continue;
            dims[n] = uint64_t(plan->shape[d]);
            strides[n - 1] = uint64_t(plan->strides[a][d]);
            ++n;
        }
        static_assert(sizeof(CUtensorMap) == sizeof(TensorMapBlob), "tensor map blob size");
        int st = make_tensor_map(reinterpret_cast<CUtensorMap*>(&out->m[slot]), e, args[a].data, 5, dims, strides, box,
                                 CU_TENSOR_MAP_SWIZZLE_128B);
        if (st) return st;
        ++slot;
    }
    return 0;
}

unsigned ew_grid(const b200_ew_plan_t* plan, int threads, int unroll, int sm_count) {
    int64_t blocks;
    if (plan->variant == B200_EW_TILED_TMA) {
        int esz, ti, to;
        tma_tile_geometry(plan, &esz, &ti, &to);
        const int64_t ni = plan->shape[plan->tile_axis], no = plan->shape[plan->ndim - 1];
        blocks = ((ni + ti - 1) / ti) * ((no + to - 1) / to) * (plan->size / (ni * no));
        const int per_sm = ((plan->reserved >> 8) & 0xffff) ? int((plan->reserved >> 8) & 0xffff) : tma_blocks_per_sm(plan);
        return unsigned(std::max<int64_t>(1, std::min<int64_t>(blocks, int64_t(sm_count) * per_sm)));
    }
    if (plan->variant == B200_EW_TILED_REG) {
        // mirrors RegTileTiler: 8*UN warp units of (128 B along I) x (64 B along O), O first
        const int e = int(plan->reserved >> 24), ch = 16 / e, un = (plan->reserved & 0xff) ? int(plan->reserved & 0xff) : reg_tile_unroll(e), units = 8 * un;
        const int64_t ni = plan->shape[plan->tile_axis], no = plan->shape[plan->ndim - 1];
        const int units_o = reg_tile_units_o(no, 4 * ch, units);
        const int64_t to = int64_t(units_o) * 4 * ch, ti = int64_t(units / units_o) * 8 * ch;
        blocks = ((ni + ti - 1) / ti) * ((no + to - 1) / to) * (plan->size / (ni * no));
        const int per_sm = ((plan->reserved >> 8) & 0xffff) ? int((plan->reserved >> 8) & 0xffff) : 32;
        return unsigned(std::max<int64_t>(1, std::min<int64_t>(blocks, int64_t(sm_count) * per_sm)));
    }
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
    const int per_sm = ((plan->reserved >> 8) & 0xffff) ? int((plan->reserved >> 8) & 0xffff)
                       : (2048 / threads) * (plan->variant == B200_EW_ROWWISE ? 8 : 4);
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

extern "C" __attribute__((visibility("default"))) int b200_workspace_init(void* workspace, size_t bytes, void* stream) {
    if (!workspace && bytes) return fail(B200_E_INVALID, "null workspace");
    if (bytes) B200_CUDA_TRY(cudaMemsetAsync(workspace, 0, bytes, static_cast<cudaStream_t>(stream)));
    return 0;
}

static int plan_impl(int nargs, const b200_operand_t* args, uint32_t flags, b200_ew_plan_t* plan);

extern "C" __attribute__((visibility("default"))) int b200_ew_plan(int nargs, const b200_operand_t* args, b200_ew_plan_t* plan) {
    return plan_impl(nargs, args, B200_PLAN_KEEP_ORDER, plan);
}

extern "C" __attribute__((visibility("default"))) int b200_ew_plan_ex(int nargs, const b200_operand_t* args, uint32_t flags,
                                                                       b200_ew_plan_t* plan) {
    return plan_impl(nargs, args, flags, plan);
}

static int plan_impl(int nargs, const b200_operand_t* args, uint32_t flags, b200_ew_plan_t* plan) {
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
    // ---- loop order: a kernel that never observes the C-order linear index may walk the dims in any
    //      order, so sort them by the (first) output's stride, largest first -- F-ordered / permuted
    //      operands then collapse and vectorise exactly like C-ordered ones
    if (!(flags & B200_PLAN_KEEP_ORDER) && nd > 1) {
        int ref = -1;
        for (int a = 0; a < nargs && ref < 0; ++a)
            if (is_array(args[a]) && args[a].is_output) ref = a;
        if (ref < 0) ref = first;
        int order[B200_MAX_NDIM];
        for (int d = 0; d < nd; ++d) order[d] = d;
        auto key2 = [&](int d) {
            int64_t m = 0;
            for (int a = 0; a < nargs; ++a)
                if (is_array(args[a])) m = std::max<int64_t>(m, std::llabs(st[a][d]));
            return m;
        };
        std::stable_sort(order, order + nd, [&](int x, int y) {
            const int64_t ax = std::llabs(st[ref][x]), ay = std::llabs(st[ref][y]);
            if (ax != ay) return ax > ay;
            return key2(x) > key2(y);
        });
        int64_t shape2[B200_MAX_NDIM];
        int64_t st2[B200_MAX_ARGS][B200_MAX_NDIM];
        for (int d = 0; d < nd; ++d) {
            shape2[d] = shape[order[d]];
            for (int a = 0; a < nargs; ++a) st2[a][d] = st[a][order[d]];
        }
        for (int d = 0; d < nd; ++d) {
            shape[d] = shape2[d];
            for (int a = 0; a < nargs; ++a) st[a][d] = st2[a][d];
        }
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
    // ---- FLAT with periodic operands: a dense (rows, inner) loop whose other operands are row vectors
    //      broadcast over the rows (stride 0, unit stride).  The dense operands are walked as 1-D; a
    //      periodic operand is indexed by (i % inner), which is constant per thread when 256 * vec is a
    //      multiple of inner (FlatTiler).  Requires full vector alignment everywhere.
    if (nd == 2) {
        uint32_t pmask = 0;
        bool ok = shape[1] % vec == 0 && (int64_t(256) * vec) % shape[1] == 0;
        for (int a = 0; a < nargs && ok; ++a) {
            if (!is_array(args[a])) continue;
            const int isz = dtype_size(args[a].dtype);
            if (reinterpret_cast<uintptr_t>(args[a].data) % (uintptr_t(vec) * isz)) ok = false;
            if (st[a][1] != isz) { ok = false; break; }
            if (st[a][0] == 0 && !args[a].is_output) pmask |= 1u << a;
            else if (st[a][0] != shape[1] * isz) ok = false;
        }
        if (ok && pmask) {
            plan->variant = B200_EW_FLAT;
            plan->ndim = 1;
            plan->shape[0] = size;
            for (int a = 0; a < nargs; ++a) plan->strides[a][0] = is_array(args[a]) ? dtype_size(args[a].dtype) : 0;
            plan->vec = vec;
            plan->staged_mask = pmask;
            plan->tile_axis = int32_t(shape[1]);
            return 0;
        }
    }
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
            // Vector forms (TILED_REG by default, TILED_TMA on request): one item size (2/4/8 B)
            // across the array operands and 16-byte alignment of every base and every stride that
            // a vector access (or a tensor map) steps by; both tile extents multiples of 16/item.
            const char* mode_env = getenv("B200_EW_TILED_MODE");
            const int mode = !mode_env ? 0 : !strcmp(mode_env, "tma") ? 1 : !strcmp(mode_env, "smem") ? 2 : 0;
            int esz = 0;
            bool vecok = mask != 0 && mode != 2;
            bool tma = nd <= 5 && __builtin_popcount(mask) <= kMaxStaged;
            for (int a = 0; a < nargs && vecok; ++a) {
                if (!is_array(args[a])) continue;
                const int isz = dtype_size(args[a].dtype);
                if (esz == 0) esz = isz;
                if (isz != esz || (isz != 2 && isz != 4 && isz != 8)) { vecok = false; break; }
                const bool staged = (mask >> a) & 1u;
                const bool unit = st[a][last] == isz;
                if (!staged && !unit) continue;              // scalar accesses: no alignment needed
                if (reinterpret_cast<uintptr_t>(args[a].data) % 16) vecok = false;
                for (int d = 0; d < nd && vecok; ++d) {
                    if (staged ? d == axis : d == last) continue;
                    if (st[a][d] % 16) vecok = false;
                    if (staged && (st[a][d] <= 0 || st[a][d] >= (int64_t(1) << 40))) tma = false;
                }
            }
            for (int d = 0; d < nd; ++d)
                if (shape[d] >= (int64_t(1) << 31)) vecok = false;
            if (vecok && esz && shape[last] % (16 / esz) == 0 && shape[axis] % (16 / esz) == 0 &&
                size / (16 / esz) / (16 / esz) < (int64_t(1) << 31)) {
                plan->variant = (mode == 1 && tma) ? B200_EW_TILED_TMA : B200_EW_TILED_REG;
                plan->vec = 16 / esz;
                plan->reserved = uint32_t(esz) << 24;
            }
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
