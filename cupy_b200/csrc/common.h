// Host-side helpers shared by the translation units of libcupy_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/cupy_b200.h"
#include "include/b200/elementwise.cuh"

namespace b200 {

// thread-local error text behind b200_last_error_string()
std::string& last_error();
int fail(int code, const char* fmt, ...);

#define B200_CUDA_TRY(expr)                                                           \
    do {                                                                              \
        cudaError_t e__ = (expr);                                                     \
        if (e__ != cudaSuccess)                                                       \
            return ::b200::fail(int(e__), "%s: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

struct DeviceInfo {
    int sm_count, cc_major, cc_minor;
    size_t l2_bytes;
};
// cached per device; returns non-zero status on failure
int device_info(DeviceInfo* out);

inline int dtype_size(int dtype) {
    static const int sz[B200_NUM_TYPES] = {1, 1, 2, 2, 4, 4, 8, 8, 2, 4, 8, 8, 16, 1};
    return (dtype >= 0 && dtype < B200_NUM_TYPES) ? sz[dtype] : 0;
}

// Build the by-value kernel parameter block from a plan and its operands.
int fill_ew_params(const b200_ew_plan_t* plan, int nargs, const b200_operand_t* args, EwParams* out);

// TILED_TMA: tile geometry, ring depth + dynamic shared memory, tensor maps of the staged operands
void tma_tile_geometry(const b200_ew_plan_t* plan, int* esz, int* tile_i, int* tile_o);
int tma_blocks_per_sm(const b200_ew_plan_t* plan);
void tma_ring(const b200_ew_plan_t* plan, int* stages, unsigned* smem_bytes);
int build_tile_maps(const b200_ew_plan_t* plan, const b200_operand_t* args, TileMaps* out);

// Grid size for a persistent elementwise launch.
unsigned ew_grid(const b200_ew_plan_t* plan, int threads, int unroll, int sm_count);

}  // namespace b200
