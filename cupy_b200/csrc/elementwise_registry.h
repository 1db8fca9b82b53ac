// Registry of prebuilt elementwise kernels shared by elementwise.cu (dispatch)
// and elementwise_table.cu (instantiation groups).
#pragma once
#include <map>
#include <tuple>
namespace b200 {
constexpr int kEwThreads = 256;
// ---- registry -------------------------------------------------------------
struct EwKernels {
    const void* flat_v;      // FLAT, full vector width
    const void* flat_1;      // FLAT, scalar access (misaligned views)
    const void* row_v32;     // ROWWISE, full vector width, 32-bit index
    const void* row_132;     // ROWWISE, scalar, 32-bit index
    const void* row_164;     // ROWWISE, scalar, 64-bit index
    const void* row_v64;     // ROWWISE, vector, 64-bit index
    const void* tiled;
    const void* tiled_reg;   // TILED_REG (same eligibility as tiled_tma)
    const void* tiled_tma;   // TILED_TMA (nullptr unless all operand item sizes are equal and 2/4/8 bytes)
    int vec;                 // "full" vector width of this instantiation
    int unroll_flat, unroll_row;
};

typedef std::tuple<int, int, int> Key;   // (ufunc, in dtype, out dtype)

std::map<Key, EwKernels>& registry();

void register_ew_group0(); void register_ew_group1(); void register_ew_group2(); void register_ew_group3();
void register_ew_group4(); void register_ew_group5(); void register_ew_group6();
}  // namespace b200
