// Test infrastructure (see oracle/Makefile).  Compiles the REFERENCE's native CUB
// wrapper from where it lies (/root/reference/cupy/cuda/cupy_cub.cu, included by
// path -- no reference source is copied into this repository) and exports its
// C++ entry points (cupy/cuda/cupy_cub.h:29-38) under plain C names so that the
// GPU baseline harness can call them through ctypes.
#include B200_REF_CUB_SOURCE

extern "C" {

size_t ref_cub_reduce_workspace(void* x, void* y, int n, cudaStream_t s, int op, int dtype_id) {
    return cub_device_reduce_get_workspace_size(x, y, n, s, op, dtype_id);
}
void ref_cub_reduce(void* ws, size_t ws_bytes, void* x, void* y, int n, cudaStream_t s, int op, int dtype_id) {
    cub_device_reduce(ws, ws_bytes, x, y, n, s, op, dtype_id);
}
size_t ref_cub_segmented_reduce_workspace(void* x, void* y, int n_seg, int seg_size, cudaStream_t s, int op, int dtype_id) {
    return cub_device_segmented_reduce_get_workspace_size(x, y, n_seg, seg_size, s, op, dtype_id);
}
void ref_cub_segmented_reduce(void* ws, size_t ws_bytes, void* x, void* y, int n_seg, int seg_size,
                              cudaStream_t s, int op, int dtype_id) {
    cub_device_segmented_reduce(ws, ws_bytes, x, y, n_seg, seg_size, s, op, dtype_id);
}
size_t ref_cub_scan_workspace(void* x, void* y, int n, cudaStream_t s, int op, int dtype_id) {
    return cub_device_scan_get_workspace_size(x, y, n, s, op, dtype_id);
}
void ref_cub_scan(void* ws, size_t ws_bytes, void* x, void* y, int n, cudaStream_t s, int op, int dtype_id) {
    cub_device_scan(ws, ws_bytes, x, y, n, s, op, dtype_id);
}

}  // extern "C"
