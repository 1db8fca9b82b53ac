/*
 * cupy_b200.h -- C ABI of the B200-native data-parallel engine.
 *
 * This is the drop-in boundary for ONE hot path of cupy/cupy: the
 * ufunc / ElementwiseKernel launcher, the ReductionKernel / axis-reduction
 * path and the cumsum/cumprod scan.  Every entry point below replaces a native
 * interface of the reference (paths relative to the reference tree):
 *
 *   reference interface                                   replaced by
 *   ----------------------------------------------------  ---------------------------
 *   cupy/cuda/cupy_cub.h:29      cub_device_reduce          b200_reduce_run (layout FULL)
 *   cupy/cuda/cupy_cub.h:30      cub_device_segmented_reduce b200_reduce_run (layout ROWS)
 *   cupy/cuda/cupy_cub.h:31      cub_device_scan            b200_scan_run
 *   cupy/cuda/cupy_cub.h:34-36   *_get_workspace_size       b200_reduce_workspace_bytes,
 *                                                           b200_scan_workspace_bytes
 *   cupy/cuda/cupy_cub.h:4-11    CUPY_CUB_* op codes        B200_OP_* (same values 0..7)
 *   cupy/_core/include/cupy/type_dispatcher.cuh:15-28       B200_TYPE_* (same values 0..13)
 *   cupy/_core/_kernel.pyx:360-461 _reduce_dims/_reduced_view_core
 *                                                           b200_ew_plan (collapse + classify)
 *   cupy/_core/_kernel.pyx:1024-1100 _get_ufunc_kernel + function.pyx:153-171 linear_launch
 *                                                           b200_ufunc_launch (prebuilt kernels)
 *   cupy/cuda/compiler.py:655-790 _compile_with_cache_cuda (NVRTC)
 *                                                           b200_jit_compile
 *   cupy/cuda/function.pyx:193-232 Module.load / get_function
 *                                                           b200_module_load / b200_module_get_function
 *   cupy/cuda/function.pyx:92-171 Function._launch -> driver.pyx:273-286 cuLaunchKernel
 *                                                           b200_jit_ew_launch / b200_jit_reduce_launch
 *   cupy/_core/_reduction.pyx:481-508 _launch (generic reduction geometry)
 *                                                           b200_reduce_run (layout COLS / GENERIC)
 *
 * Conventions:
 *   - plain C: pointers, sizes, integer codes.  No torch / C++ types.
 *   - every function returns an int status: 0 = ok, negative = B200_E_* (invalid
 *     argument / unsupported: the Python host raises), positive = cudaError_t /
 *     CUresult / nvrtcResult of the failing call.  b200_last_error_string() gives
 *     the text for the calling thread.
 *   - no hidden allocation and no hidden synchronisation: outputs and workspaces
 *     are caller-owned device pointers valid on `stream`; all work is enqueued on
 *     `stream` (a cudaStream_t passed as void*).
 *   - strides are in BYTES (NumPy/CuPy convention).
 */
#ifndef CUPY_B200_H_
#define CUPY_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

/* ---- dtype ids: identical to cupy/_core/include/cupy/type_dispatcher.cuh:15-28 */
#define B200_TYPE_INT8        0
#define B200_TYPE_UINT8       1
#define B200_TYPE_INT16       2
#define B200_TYPE_UINT16      3
#define B200_TYPE_INT32       4
#define B200_TYPE_UINT32      5
#define B200_TYPE_INT64       6
#define B200_TYPE_UINT64      7
#define B200_TYPE_FLOAT16     8
#define B200_TYPE_FLOAT32     9
#define B200_TYPE_FLOAT64    10
#define B200_TYPE_COMPLEX64  11   /* recognised, no prebuilt kernels (out of scope) */
#define B200_TYPE_COMPLEX128 12
#define B200_TYPE_BOOL       13
#define B200_NUM_TYPES       14

/* ---- reduction / scan op codes: 0..7 identical to cupy/cuda/cupy_cub.h:4-11 */
#define B200_OP_SUM      0
#define B200_OP_MIN      1
#define B200_OP_MAX      2
#define B200_OP_ARGMIN   3
#define B200_OP_ARGMAX   4
#define B200_OP_CUMSUM   5
#define B200_OP_CUMPROD  6
#define B200_OP_PROD     7
/* extensions (the reference builds these from ReductionKernel strings,
 * cupy/_core/_routines_statistics.pyx:611-655) */
#define B200_OP_MEAN     8
#define B200_OP_VAR      9    /* param = ddof; single pass Welford/Chan */
#define B200_OP_MOMENTS  10   /* full reductions only: y = three doubles (n, mean, M2), accumulated in float for
                                 fp16/fp32 inputs and double otherwise -- what a caller needs to merge shards
                                 (Chan) without a second pass; out_dtype = float64.  The reference has no
                                 counterpart (two passes, cupy/_core/_routines_statistics.pyx:556-600) */

/* ---- status codes */
#define B200_OK               0
#define B200_E_INVALID       -1   /* bad argument */
#define B200_E_UNSUPPORTED   -2   /* valid request without a prebuilt kernel: caller falls back to JIT */
#define B200_E_WORKSPACE     -3   /* workspace too small */
#define B200_E_NOLIB         -4   /* libcuda / libnvrtc could not be loaded */
#define B200_E_COMPILE       -5   /* NVRTC compilation failed (log available) */

#define B200_MAX_NDIM 10
#define B200_MAX_ARGS 12

/* ---- operands of an elementwise call (already broadcast to a common shape
 *      by the host: broadcast dims carry stride 0), cf. the by-value CArray
 *      struct of the reference, cupy/_core/_carray.pyx:90-128 */
#define B200_KIND_ARRAY  0
#define B200_KIND_SCALAR 1
#define B200_KIND_RAW    2   /* `raw` array: not broadcast, indexed by user code */

typedef struct b200_operand {
    void*   data;                      /* device pointer (ARRAY / RAW) */
    int64_t scalar[2];                 /* raw little-endian scalar bytes (SCALAR), already cast to `dtype` */
    int32_t kind;                      /* B200_KIND_* */
    int32_t dtype;                     /* B200_TYPE_* */
    int32_t ndim;
    int32_t is_output;
    int64_t shape[B200_MAX_NDIM];
    int64_t strides[B200_MAX_NDIM];    /* bytes */
} b200_operand_t;

/* ---- launch plan produced by the classifier */
#define B200_EW_FLAT     0   /* every array operand is dense with the same layout: 1-D, 128-bit vector access */
#define B200_EW_ROWWISE  1   /* N-D strided / broadcast: vectors along the innermost dim, one index decomposition per vector */
#define B200_EW_TILED    2   /* some input is unit-stride along another dim: 32x32 shared-memory tile transpose */
#define B200_EW_TILED_TMA 3  /* same call shape, equal item sizes, 16-byte aligned: TMA-pipelined swizzled tiles (A/B knob
                                B200_EW_TILED_MODE=tma; tensor copies are translation-bound on 4 MB-strided rows) */
#define B200_EW_TILED_REG 4  /* same conditions: register-block transpose (16-byte vectors along both dims, no shared
                                memory) -- the default for transposing calls */

typedef struct b200_ew_plan {
    int32_t  variant;                  /* B200_EW_* */
    int32_t  ndim;                     /* collapsed rank */
    int32_t  vec;                      /* elements per vector access (1,2,4,8,16) */
    int32_t  idx32;                    /* 1 if every offset fits in int32 */
    int32_t  tile_axis;                /* TILED*: collapsed dim along which staged operands are unit-stride;
                                          FLAT with periodic operands: the period in elements */
    int32_t  nargs;
    uint32_t staged_mask;              /* TILED*: bit k set = operand k is staged (unit-stride along tile_axis);
                                          FLAT: bit k set = operand k is a row vector broadcast over the rows (periodic) */
    uint32_t reserved;
    int64_t  size;                     /* number of loop elements */
    int64_t  shape[B200_MAX_NDIM];
    int64_t  strides[B200_MAX_ARGS][B200_MAX_NDIM];   /* bytes, collapsed */
} b200_ew_plan_t;

/* ---- prebuilt ufunc ids (routines of cupy/_core/_routines_math.pyx:878-1178,
 *      cupy/_math/explog.py, cupy/_core/_ufuncs.py:7-13 elementwise_copy) */
enum b200_ufunc {
    B200_UF_COPY = 0,       /* out0 = in0 (dtype cast) */
    B200_UF_ADD,
    B200_UF_SUBTRACT,
    B200_UF_MULTIPLY,
    B200_UF_TRUE_DIVIDE,
    B200_UF_NEGATIVE,
    B200_UF_ABSOLUTE,
    B200_UF_SQUARE,
    B200_UF_SQRT,
    B200_UF_EXP,
    B200_UF_LOG,
    B200_UF_MAXIMUM,
    B200_UF_MINIMUM,
    B200_UF_FMA,            /* out0 = in0 * in1 + in2 (one rounding) */
    B200_UF_COUNT
};

/* ---- library / device */
int         b200_abi_version(void);
const char* b200_last_error_string(void);
int         b200_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes);
int         b200_dtype_itemsize(int dtype);

/* Zero-fills a freshly allocated reduction / scan workspace ON `stream` (cudaMemsetAsync), i.e. ordered before
 * the kernels the caller enqueues on that stream afterwards.  Workspaces must be initialised this way once; the
 * kernels leave them zeroed (self-resetting tickets), so a workspace is reused without further memsets as
 * long as it is used on ONE stream at a time. */
int b200_workspace_init(void* workspace, size_t bytes, void* stream);

/* ---- elementwise launcher */
int b200_ew_plan(int nargs, const b200_operand_t* args, b200_ew_plan_t* plan);
/* Same, with flags.  B200_PLAN_KEEP_ORDER: the kernel observes the C-order linear index (`i`, `_ind`, raw
 * operands), so the loop dims keep their order (what b200_ew_plan does).  Without it the dims are sorted by
 * the output's strides first, so F-ordered and permuted operands collapse and vectorise like C-ordered
 * ones (the reference only collapses C-contiguous operands, cupy/_core/_kernel.pyx:360-461). */
#define B200_PLAN_KEEP_ORDER 1u
int b200_ew_plan_ex(int nargs, const b200_operand_t* args, uint32_t flags, b200_ew_plan_t* plan);
int b200_ufunc_supported(int ufunc, int nin, const int32_t* in_dtypes, int32_t out_dtype);
int b200_ufunc_launch(int ufunc, const b200_ew_plan_t* plan, int nargs,
                      const b200_operand_t* args, void* stream);

/* ---- reductions */
#define B200_RED_FULL    0   /* x[n]                       -> y[1]           */
#define B200_RED_ROWS    1   /* x[rows][n]  (n contiguous) -> y[rows]        */
#define B200_RED_COLS    2   /* x[batch][n][cols]          -> y[batch][cols] */

typedef struct b200_reduce_desc {
    int32_t op;            /* B200_OP_* */
    int32_t layout;        /* B200_RED_* */
    int32_t in_dtype;
    int32_t out_dtype;
    int64_t batch;         /* COLS: leading batch count (>=1); else 1 */
    int64_t n_reduce;      /* elements reduced per output */
    int64_t n_out;         /* ROWS: rows; COLS: cols; FULL: 1 */
    double  param;         /* VAR: ddof */
} b200_reduce_desc_t;

int b200_reduce_supported(const b200_reduce_desc_t* d);
int b200_reduce_workspace_bytes(const b200_reduce_desc_t* d, size_t* bytes);
/* workspace must be 16-byte aligned and zero-filled once when it is allocated; kernels leave it zeroed */
int b200_reduce_run(const b200_reduce_desc_t* d, const void* x, void* y,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Sharded FULL reduction with the cross-GPU combine fused into the kernel's last block: every rank
 * runs this call on its shard; the per-GPU partial accumulators are exchanged through `slots` -- one small
 * buffer per rank that every rank has mapped (symmetric memory over NVLink / NVSwitch peer access), written
 * with plain stores on the mapped peer pointers -- folded in rank order and post-processed with the TOTAL
 * element count, so that y holds the global sum / prod / min / max / mean / var, bit-identical on every rank,
 * after ONE launch.  Replaces per-GPU reduction + ncclAllReduce(count=1) (cupyx/distributed/_nccl_comm.py:
 * 312-320 -> cupy_backends/cuda/libs/nccl.pyx:470-477).  Collective: every rank must make the same call in the
 * same order on one stream.  B200_E_UNSUPPORTED for arg-reductions and non-FULL layouts. */
#define B200_MAX_PEERS 16
#define B200_EXCHANGE_BYTES 2048     /* per-rank buffer: [2 parities][16 sources][8 words] x 8 bytes; zeroed once */
typedef struct b200_peer_exchange {
    int32_t  rank, nranks;
    uint32_t tag;                    /* per-communicator call counter: same on every rank, changes every call, != 0 */
    uint32_t reserved;
    int64_t  n_total;                /* elements over all ranks */
    void*    slots[B200_MAX_PEERS];  /* slots[r] = rank r's exchange buffer as mapped into this process */
} b200_peer_exchange_t;
int b200_reduce_run_sharded(const b200_reduce_desc_t* d, const void* x, void* y, void* workspace, size_t workspace_bytes,
                            const b200_peer_exchange_t* exchange, void* stream);

/* Sharded variance (the reference has no distributed var, cupyx/distributed/array/_array.py:744-747):
 * Chan merge, in index order, of `count` (n, mean, M2) double triples as written by B200_OP_MOMENTS and
 * all-gathered over the ranks.  out[0] = M2 / (n - ddof) (NaN when n - ddof <= 0), out[1] = mean, out[2] = n. */
int b200_moments_merge(const double* triples, int count, double ddof, double* out, void* stream);

/* ---- scan (inclusive, flat, decoupled look-back single pass) */
int b200_scan_supported(int op, int in_dtype, int out_dtype);
int b200_scan_workspace_bytes(int64_t n, int out_dtype, size_t* bytes);
int b200_scan_run(int op, int in_dtype, int out_dtype, const void* x, void* y, int64_t n,
                  void* workspace, size_t workspace_bytes, void* stream);

/* cumsum / cumprod along one axis of a dense array x[outer][n][inner] (C order; scans n), dtype
 * conversion fused into the load, result y in the same layout.  Replaces _proc_as_batch +
 * _batch_scan_op (cupy/_core/_routines_math.pyx:499-699: two transposing copies + log2(n) passes)
 * with one pass.  Same dtype table as b200_scan_supported.  A workspace is needed only when
 * outer * inner columns cannot fill the GPU and n is cut into segments (segment totals);
 * it does not have to be zeroed. */
int b200_scan_axis_workspace_bytes(int64_t outer, int64_t n, int64_t inner, size_t* bytes);
int b200_scan_axis_run(int op, int in_dtype, int out_dtype, const void* x, void* y,
                       int64_t outer, int64_t n, int64_t inner,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- JIT of user code strings (NVRTC -> cubin -> module -> function) */
int  b200_jit_compile(const char* source, const char* name, int n_options,
                      const char* const* options, void** image, size_t* image_bytes);
const char* b200_jit_last_log(void);
void b200_jit_free_image(void* image);
int  b200_module_load(const void* image, void** module);
int  b200_module_unload(void* module);
int  b200_module_get_function(void* module, const char* name, void** function);

/* raw (non-broadcast) array views handed to user code as CArray objects */
typedef struct b200_raw_view {
    void*   data;
    int64_t size;
    int32_t ndim;
    int32_t reserved;
    int64_t shape[B200_MAX_NDIM];
    int64_t strides[B200_MAX_NDIM];
} b200_raw_view_t;

/* Launch an NVRTC-built elementwise kernel generated over the skeleton in
 * cupy_b200/csrc/include/b200/elementwise.cuh. */
int b200_jit_ew_launch(void* function, const b200_ew_plan_t* plan, int nargs,
                       const b200_operand_t* args, int block_size, void* stream);
/* Same; additionally hands the kernel the UN-collapsed loop shape (`ind_ndim`, `ind_shape`) for kernels
 * created with reduce_dims=False whose code reads `_ind` (the reference builds the Indexer from the original
 * shape then, cupy/_core/_kernel.pyx:931-940).  ind_ndim = 0: as b200_jit_ew_launch.  Any number of `raw`
 * operands (<= B200_MAX_ARGS) is accepted. */
int b200_jit_ew_launch_ex(void* function, const b200_ew_plan_t* plan, int nargs,
                          const b200_operand_t* args, int block_size, int ind_ndim, const int64_t* ind_shape,
                          void* stream);

/* Launch an NVRTC-built reduction kernel (skeleton: b200/reduce.cuh).
 * `params` is the packed by-value kernel parameter block built by the host. */
int b200_jit_launch(void* function, unsigned grid_x, unsigned grid_y, unsigned grid_z,
                    unsigned block_x, unsigned shared_bytes,
                    const void* params, size_t params_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CUPY_B200_H_ */
