// jit.cu -- NVRTC compilation of user code strings and launch of the resulting
// kernels through the CUDA driver API.  libnvrtc and libcuda are opened at run
// time (dlopen) so that the library loads -- and the plan / workspace entry
// points work -- on a machine without a GPU driver.
//
// Replaces cupy/cuda/compiler.py:346-403 (compile_using_nvrtc), :655-790
// (_compile_with_cache_cuda, minus the disk cache which lives in the Python
// host), cupy/cuda/function.pyx:92-232 (Module / Function / _launch) and
// cupy_backends/cuda/api/driver.pyx:273-286 (launchKernel).
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.h"
#include "include/b200/carray.cuh"

namespace b200 {

// ---- minimal NVRTC / driver API surface (resolved with dlsym) ---------------
typedef struct _nvrtcProgram* nvrtcProgram;
typedef int nvrtcResult;
typedef int CUresult;
typedef struct CUmod_st* CUmodule;
typedef struct CUfunc_st* CUfunction;
typedef struct CUstream_st* CUstream;

struct Nvrtc {
    void* handle = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    nvrtcResult (*DestroyProgram)(nvrtcProgram*);
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
    const char* (*GetErrorString)(nvrtcResult);
};

struct Driver {
    void* handle = nullptr;
    CUresult (*ModuleLoadData)(CUmodule*, const void*);
    CUresult (*ModuleUnload)(CUmodule);
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                             unsigned, CUstream, void**, void**);
    CUresult (*GetErrorString)(CUresult, const char**);
    CUresult (*FuncSetAttribute)(CUfunction, int, int);
};

static void* open_first(const char* const* names) {
    for (; *names; ++names) {
        void* h = dlopen(*names, RTLD_NOW | RTLD_GLOBAL);
        if (h) return h;
    }
    return nullptr;
}

static Nvrtc* nvrtc() {
    static Nvrtc lib;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("CUPY_B200_NVRTC");
        const char* names[] = {env ? env : "libnvrtc.so.12", "libnvrtc.so.12", "libnvrtc.so",
                               "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so", nullptr};
        void* h = open_first(names);
        if (!h) return;
#define SYM(field, name) *reinterpret_cast<void**>(&lib.field) = dlsym(h, name)
        SYM(CreateProgram, "nvrtcCreateProgram");
        SYM(DestroyProgram, "nvrtcDestroyProgram");
        SYM(CompileProgram, "nvrtcCompileProgram");
        SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
        SYM(GetProgramLog, "nvrtcGetProgramLog");
        SYM(GetCUBINSize, "nvrtcGetCUBINSize");
        SYM(GetCUBIN, "nvrtcGetCUBIN");
        SYM(GetErrorString, "nvrtcGetErrorString");
        if (lib.CreateProgram && lib.CompileProgram && lib.GetCUBIN) lib.handle = h;
    });
    return lib.handle ? &lib : nullptr;
}

static Driver* driver() {
    static Driver lib;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libcuda.so.1", "libcuda.so", nullptr};
        void* h = open_first(names);
        if (!h) return;
        SYM(ModuleLoadData, "cuModuleLoadData");
        SYM(ModuleUnload, "cuModuleUnload");
        SYM(ModuleGetFunction, "cuModuleGetFunction");
        SYM(LaunchKernel, "cuLaunchKernel");
        SYM(GetErrorString, "cuGetErrorString");
        SYM(FuncSetAttribute, "cuFuncSetAttribute");
#undef SYM
        if (lib.ModuleLoadData && lib.LaunchKernel) lib.handle = h;
    });
    return lib.handle ? &lib : nullptr;
}

static std::string& jit_log() {
    static thread_local std::string s;
    return s;
}

static int cu_fail(Driver* d, CUresult r, const char* what) {
    const char* msg = nullptr;
    if (d->GetErrorString) d->GetErrorString(r, &msg);
    return fail(int(r), "%s: %s", what, msg ? msg : "CUDA driver error");
}

}  // namespace b200

using namespace b200;

extern "C" __attribute__((visibility("default"))) const char* b200_jit_last_log(void) { return jit_log().c_str(); }

extern "C" __attribute__((visibility("default"))) int b200_jit_compile(const char* source, const char* name, int n_options,
                                const char* const* options, void** image, size_t* image_bytes) {
    if (!source || !image || !image_bytes) return fail(B200_E_INVALID, "null argument");
    Nvrtc* rt = nvrtc();
    if (!rt) return fail(B200_E_NOLIB, "libnvrtc could not be loaded (set CUPY_B200_NVRTC)");
    jit_log().clear();
    nvrtcProgram prog = nullptr;
    nvrtcResult r = rt->CreateProgram(&prog, source, name ? name : "kernel.cu", 0, nullptr, nullptr);
    if (r) return fail(int(r), "nvrtcCreateProgram: %s", rt->GetErrorString ? rt->GetErrorString(r) : "?");
    r = rt->CompileProgram(prog, n_options, options);
    size_t log_size = 0;
    if (rt->GetProgramLogSize(prog, &log_size) == 0 && log_size > 1) {
        jit_log().resize(log_size);
        rt->GetProgramLog(prog, &jit_log()[0]);
    }
    if (r) {
        rt->DestroyProgram(&prog);
        return fail(B200_E_COMPILE, "NVRTC compilation failed (%s); see b200_jit_last_log()",
                    rt->GetErrorString ? rt->GetErrorString(r) : "?");
    }
    size_t size = 0;
    r = rt->GetCUBINSize(prog, &size);
    if (r || size == 0) {
        rt->DestroyProgram(&prog);
        return fail(r ? int(r) : B200_E_COMPILE, "nvrtcGetCUBINSize failed: pass --gpu-architecture=sm_100a (a real arch)");
    }
    char* buf = static_cast<char*>(malloc(size));
    if (!buf) { rt->DestroyProgram(&prog); return fail(B200_E_INVALID, "out of host memory"); }
    r = rt->GetCUBIN(prog, buf);
    rt->DestroyProgram(&prog);
    if (r) { free(buf); return fail(int(r), "nvrtcGetCUBIN failed"); }
    *image = buf;
    *image_bytes = size;
    return 0;
}

extern "C" __attribute__((visibility("default"))) void b200_jit_free_image(void* image) { free(image); }

extern "C" __attribute__((visibility("default"))) int b200_module_load(const void* image, void** module) {
    if (!image || !module) return fail(B200_E_INVALID, "null argument");
    Driver* d = driver();
    if (!d) return fail(B200_E_NOLIB, "libcuda could not be loaded");
    // make sure the primary context of the runtime's current device is current
    B200_CUDA_TRY(cudaFree(nullptr));
    CUmodule m = nullptr;
    CUresult r = d->ModuleLoadData(&m, image);
    if (r) return cu_fail(d, r, "cuModuleLoadData");
    *module = m;
    return 0;
}

extern "C" __attribute__((visibility("default"))) int b200_module_unload(void* module) {
    Driver* d = driver();
    if (!d) return fail(B200_E_NOLIB, "libcuda could not be loaded");
    CUresult r = d->ModuleUnload(static_cast<CUmodule>(module));
    return r ? cu_fail(d, r, "cuModuleUnload") : 0;
}

extern "C" __attribute__((visibility("default"))) int b200_module_get_function(void* module, const char* name, void** function) {
    if (!module || !name || !function) return fail(B200_E_INVALID, "null argument");
    Driver* d = driver();
    if (!d) return fail(B200_E_NOLIB, "libcuda could not be loaded");
    CUfunction f = nullptr;
    CUresult r = d->ModuleGetFunction(&f, static_cast<CUmodule>(module), name);
    if (r) return cu_fail(d, r, "cuModuleGetFunction");
    *function = f;
    return 0;
}

extern "C" __attribute__((visibility("default"))) int b200_jit_launch(void* function, unsigned gx, unsigned gy, unsigned gz, unsigned bx,
                               unsigned shared_bytes, const void* params, size_t params_bytes, void* stream) {
    if (!function) return fail(B200_E_INVALID, "null function");
    Driver* d = driver();
    if (!d) return fail(B200_E_NOLIB, "libcuda could not be loaded");
    // single by-value parameter block: the generated kernels take one struct
    void* kargs[] = {const_cast<void*>(params)};
    (void)params_bytes;
    if (shared_bytes > 48 * 1024) {
        CUresult ra = d->FuncSetAttribute(static_cast<CUfunction>(function), 8 /*MAX_DYNAMIC_SHARED_SIZE_BYTES*/, int(shared_bytes));
        if (ra) return cu_fail(d, ra, "cuFuncSetAttribute");
    }
    CUresult r = d->LaunchKernel(static_cast<CUfunction>(function), gx, gy, gz, bx, 1, 1, shared_bytes,
                                 static_cast<CUstream>(stream), kargs, nullptr);
    return r ? cu_fail(d, r, "cuLaunchKernel") : 0;
}

extern "C" __attribute__((visibility("default"))) int b200_jit_ew_launch(void* function, const b200_ew_plan_t* plan, int nargs,
                                  const b200_operand_t* args, int block_size, void* stream) {
    return b200_jit_ew_launch_ex(function, plan, nargs, args, block_size, 0, nullptr, stream);
}

extern "C" __attribute__((visibility("default"))) int b200_jit_ew_launch_ex(void* function, const b200_ew_plan_t* plan, int nargs,
                                  const b200_operand_t* args, int block_size, int ind_ndim, const int64_t* ind_shape,
                                  void* stream) {
    if (!function || !plan || !args) return fail(B200_E_INVALID, "null argument");
    if (ind_ndim < 0 || ind_ndim > kMaxNdim || (ind_ndim > 0 && !ind_shape)) return fail(B200_E_INVALID, "bad ind_ndim");
    if (plan->size == 0) return 0;
    Driver* d = driver();
    if (!d) return fail(B200_E_NOLIB, "libcuda could not be loaded");
    EwParams p;
    int st = fill_ew_params(plan, nargs, args, &p);
    if (st) return st;
    // the kernel declares RawPackN<n> for exactly the views it uses; the driver copies that many from here
    RawPackN<kMaxArgs + 1> raws;
    memset(&raws, 0, sizeof(raws));
    int nraw = 0;
    for (int a = 0; a < nargs; ++a) {
        if (args[a].kind != B200_KIND_RAW) continue;
        if (args[a].ndim > kMaxNdim) return fail(B200_E_UNSUPPORTED, "raw operand rank %d exceeds %d", args[a].ndim, kMaxNdim);
        RawView& v = raws.v[nraw++];
        v.data = static_cast<char*>(args[a].data);
        v.ndim = args[a].ndim;
        v.size = 1;
        for (int k = 0; k < args[a].ndim; ++k) {
            v.shape[k] = args[a].shape[k];
            v.strides[k] = args[a].strides[k];
            v.size *= args[a].shape[k];
        }
    }
    if (ind_ndim > 0) {          // shape-only view: the un-collapsed loop shape user code sees through `_ind`
        RawView& v = raws.v[nraw++];
        v.ndim = ind_ndim;
        v.size = 1;
        for (int k = 0; k < ind_ndim; ++k) {
            v.shape[k] = ind_shape[k];
            v.size *= ind_shape[k];
        }
    }
    DeviceInfo di;
    st = device_info(&di);
    if (st) return st;
    // block_size: threads per block the kernel was generated for; unroll is
    // encoded by the generator in the plan's reserved field
    if (plan->variant == B200_EW_TILED_TMA) {
        TileMaps tm;
        st = build_tile_maps(plan, args, &tm);
        if (st) return st;
        int stages;
        unsigned smem;
        tma_ring(plan, &stages, &smem);
        p.tma_stages = stages;
        CUresult ra = d->FuncSetAttribute(static_cast<CUfunction>(function), 8 /*MAX_DYNAMIC_SHARED_SIZE_BYTES*/, 227 * 1024);
        if (ra) return cu_fail(d, ra, "cuFuncSetAttribute");
        const unsigned g = ew_grid(plan, 288, 1, di.sm_count);
        void* targs[] = {&p, &raws, &tm};
        CUresult rt = d->LaunchKernel(static_cast<CUfunction>(function), g, 1, 1, 288u, 1, 1, smem,
                                      static_cast<CUstream>(stream), targs, nullptr);
        return rt ? cu_fail(d, rt, "cuLaunchKernel") : 0;
    }
    const int unroll = (plan->reserved & 0xff) ? int(plan->reserved & 0xff) : 1;
    const int threads = (plan->variant == B200_EW_TILED || plan->variant == B200_EW_TILED_REG) ? 256 : block_size;
    const unsigned grid = ew_grid(plan, threads, unroll, di.sm_count);
    void* kargs[] = {&p, &raws};
    CUresult r = d->LaunchKernel(static_cast<CUfunction>(function), grid, 1, 1, unsigned(threads), 1, 1, 0,
                                 static_cast<CUstream>(stream), kargs, nullptr);
    return r ? cu_fail(d, r, "cuLaunchKernel") : 0;
}
