// Dtype pairs of the prebuilt scans, shared by scan.cu (flat) and scan_axis.cu (along an axis).
#pragma once
// (in dtype, out dtype) pairs with a prebuilt kernel.  Result dtype rules:
// cupy/_core/_routines_math.pyx:704-714 (bool/int -> int64, uint -> uint64, else same).
#define B200_SCAN_TABLE(X)                                              \
    X(B200_TYPE_INT64, B200_TYPE_INT64, long long, long long, long long) \
    X(B200_TYPE_INT32, B200_TYPE_INT64, int32_t, long long, long long)   \
    X(B200_TYPE_INT32, B200_TYPE_INT32, int32_t, int32_t, int32_t)       \
    X(B200_TYPE_INT16, B200_TYPE_INT64, int16_t, long long, long long)   \
    X(B200_TYPE_INT8, B200_TYPE_INT64, int8_t, long long, long long)     \
    X(B200_TYPE_INT8, B200_TYPE_INT8, int8_t, int32_t, int8_t)            \
    X(B200_TYPE_BOOL, B200_TYPE_INT64, bool, long long, long long)       \
    X(B200_TYPE_BOOL, B200_TYPE_INT32, bool, int32_t, int32_t)           \
    X(B200_TYPE_UINT8, B200_TYPE_UINT64, uint8_t, unsigned long long, unsigned long long)   \
    X(B200_TYPE_UINT16, B200_TYPE_UINT64, uint16_t, unsigned long long, unsigned long long) \
    X(B200_TYPE_UINT32, B200_TYPE_UINT64, uint32_t, unsigned long long, unsigned long long) \
    X(B200_TYPE_UINT64, B200_TYPE_UINT64, unsigned long long, unsigned long long, unsigned long long) \
    X(B200_TYPE_FLOAT32, B200_TYPE_FLOAT32, float, float, float)         \
    X(B200_TYPE_FLOAT64, B200_TYPE_FLOAT64, double, double, double)      \
    X(B200_TYPE_FLOAT16, B200_TYPE_FLOAT16, float16, float, float16)     \
    X(B200_TYPE_FLOAT16, B200_TYPE_FLOAT32, float16, float, float)       \
    X(B200_TYPE_FLOAT32, B200_TYPE_FLOAT64, float, double, double)

