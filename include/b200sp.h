/*
 * b200sp.h — C ABI of the B200-native sparse matrix-vector engine (libb200sp.so).
 *
 * This is the drop-in boundary for the SpMV / BLAS-1 / CG hot path of
 * bigno78/cusp-autotuned.  The reference has no C ABI: its boundary is a C++
 * overload set found by ADL on the execution policy.  Every entry point below
 * names the reference overload / function it replaces (paths relative to the
 * reference root).  The templated `cusp::` compatibility headers under
 * include/cusp/ unpack containers into the raw pointers + sizes taken here.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name says `_host`;
 *   - indices are int32 (the reference's tests and benchmarks only use
 *     IndexType=int); slab arithmetic (pitch*K) is done in 64 bit inside;
 *   - ELL / DIA slabs are column-major with an explicit pitch, exactly the
 *     cusp::array2d<column_major> layout (cusp/detail/ell_matrix.inl:35-36,
 *     cusp/detail/dia_matrix.inl:34): entry (row, k) at values[k*pitch + row];
 *   - `accumulate` selects the `initialize` functor of the reference's 7-arg
 *     multiply: 0 -> constant_functor(0)  (y  = A x,  generic/multiply.inl:158-162)
 *               1 -> identity             (y += A x,  testing/multiply.cu:514-645);
 *     `combine` is multiplies and `reduce` plus in these entry points; other functor
 *     triples go through b200sp_spmv_generalized;
 *   - every call is asynchronous on `stream` unless it returns a scalar to the
 *     host, in which case it synchronises that stream only;
 *   - scratch memory is owned by the handle (the reference allocates per call,
 *     cuda/detail/multiply/coo_flat_spmv.h:445-446);
 *   - a handle is bound to the device current at creation and is not
 *     thread-safe: use one handle per host thread.
 *   - status != B200SP_OK  <=>  nothing usable was written; the message is
 *     available from b200sp_last_error_string().  The C++ shim maps
 *     B200SP_INVALID_INPUT -> cusp::invalid_input_exception and everything else
 *     -> cusp::runtime_exception (cusp/exception.h:31-84).
 *   - There is NO CPU fallback anywhere behind this ABI.
 */
#ifndef B200SP_H
#define B200SP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SP_VERSION 100 /* 0.1.0 */

typedef struct b200sp_context *b200sp_handle;
typedef void *b200sp_stream; /* cudaStream_t */

typedef enum {
  B200SP_OK = 0,
  B200SP_INVALID_INPUT = 1, /* bad sizes / null pointers / unsupported cfg */
  B200SP_CUDA_ERROR = 2,    /* a CUDA runtime call or launch failed */
  B200SP_NOT_IMPLEMENTED = 3,
  B200SP_ALLOC_FAILED = 4,
  B200SP_COMM_ERROR = 5 /* NCCL / peer-access failure */
} b200sp_status;

/* ---------------------------------------------------------------------------
 * Tuning configuration — replaces the KTT parameter spaces of
 * cusp/system/cuda/ktt/{csr,ell,dia,coo}_multiply.h.
 * A zero field means "engine default".  b200sp_cfg_space_* enumerates the
 * valid points per format.
 * ------------------------------------------------------------------------- */
typedef struct {
  int kernel;          /* variant id, see B200SP_K_*                           */
  int block_size;      /* threads per CTA: 128, 256, 512                        */
  int threads_per_row; /* CSR row-split width 1,2,4,8,16,32 (1 = scalar)        */
  int unroll;          /* independent rows (CSR/ELL/DIA) or nnz (COO) per thread */
  int vector_width;    /* consecutive entries per lane per global load: K_COO_WARP 4 (128-bit
                          loads) or 8 (256-bit loads); other kernels: 1                  */
  int tile_rows;       /* rows per staged slab tile (bulk-async kernels)         */
  int stages;          /* smem pipeline depth (bulk-async kernels)               */
  int ctas_per_sm;     /* persistent-grid multiplier (bulk-async kernels)        */
} b200sp_cfg;

enum {
  B200SP_K_AUTO = 0,
  /* CSR  (replaces spmv_csr_vector_kernel / spmv_csr_scalar_kernel,
   *       cuda/detail/multiply/csr_vector_spmv.h:66-161, csr_scalar.h:48-73,
   *       and KTT csr_kernel_{naive,warp,block,balanced}, ktt/kernels/csr_kernel.h) */
  B200SP_K_CSR_VECTOR = 1, /* sub-warp per row, shuffle reduction              */
  B200SP_K_CSR_STREAM = 2, /* CTA streams a row block's nnz through smem        */
  B200SP_K_CSR_RING = 3,   /* persistent CTAs, producer warp + mbarrier ring of
                              bulk-async staged row blocks (default)            */
  B200SP_K_CSR_BALANCED = 4, /* nnz-balanced tiles + segmented scan, rows rebuilt per
                              tile from row_offsets: power-law rows; replaces KTT's
                              DYNAMIC=2 kernel + gpu_compute_row_starts
                              (cuda/ktt/csr_multiply.h:64-133), no atomics        */
  /* ELL  (replaces spmv_ell_kernel ell_spmv.h:47-93, ktt_ell_kernel)          */
  B200SP_K_ELL_LDG = 1,  /* thread per row, coalesced loads, deep unroll      */
  B200SP_K_ELL_BULK = 2, /* slabs staged by cp.async.bulk + mbarrier pipeline  */
  /* DIA  (replaces spmv_dia_kernel dia_spmv.h:66-126, ktt_dia_vector_kernel)  */
  B200SP_K_DIA_LDG = 1,
  B200SP_K_DIA_BULK = 2,
  /* COO  (replaces thrust reduce_by_key generic/multiply/spmv.h:182-238,
   *       spmv_coo_flat_kernel coo_flat_spmv.h:225-463, KTT coo_spmv)         */
  B200SP_K_COO_SEGSCAN = 1, /* nnz-balanced tiles, smem segmented scan, no atomics */
  B200SP_K_COO_RING = 2,    /* same tiles and summation order; persistent CTAs, entry streams
                               staged by cp.async.bulk into an mbarrier ring (default for
                               16-byte aligned arrays with enough tiles)                     */
  B200SP_K_COO_WARP = 3     /* warp-autonomous tiles: 128- / 256-bit loads of vector_width
                               consecutive entries per lane, products in registers, shuffle
                               segmented scan, no shared memory or CTA barrier (default for
                               scattered column streams: power-law graphs)                   */
};

/* ---- lifecycle ---------------------------------------------------------- */
int b200sp_version(void);
b200sp_status b200sp_create(b200sp_handle *out);
b200sp_status b200sp_destroy(b200sp_handle h);
/* last error text of this handle (never NULL). h may be NULL -> global text. */
const char *b200sp_last_error_string(b200sp_handle h);
const char *b200sp_status_string(b200sp_status s);
/* number of kernels this handle has launched since creation (bench bookkeeping) */
uint64_t b200sp_launch_count(b200sp_handle h);
/* pin `bytes` at `ptr` (typically x) in L2 with an access-policy window on
 * `stream` ("x through the read-only / L2-persisting path"); bytes==0 clears. */
b200sp_status b200sp_set_l2_persist(b200sp_handle h, b200sp_stream stream,
                                    const void *ptr, size_t bytes);

/* ---- SpMV ---------------------------------------------------------------
 * y = A x  (accumulate==0)   or   y += A x  (accumulate==1)
 * cfg may be NULL (engine default / cached tuned configuration).
 *
 * csr: cusp::system::cuda::detail::multiply(exec, A, x, y, init, combine, reduce,
 *      csr_format, array1d_format, array1d_format)   csr_vector_spmv.h:218-258
 * ell: ... ell_format ...                             ell_spmv.h:96-155
 * dia: ... dia_format ...                             dia_spmv.h:129-188
 * coo: ... coo_format ...                             coo_flat_spmv.h:486-502 /
 *                                                     generic/multiply/spmv.h:182-238
 * hyb: generic/multiply/spmv.h:272-290 (ELL pass, then COO pass with identity)
 */
#define B200SP_DECL_SPMV(T, sfx)                                                   \
  b200sp_status b200sp_spmv_csr_##sfx(                                             \
      b200sp_handle h, b200sp_stream stream, int64_t num_rows, int64_t num_cols,   \
      int64_t num_entries, const int32_t *row_offsets,                             \
      const int32_t *column_indices, const T *values, const T *x, T *y,            \
      int accumulate, const b200sp_cfg *cfg);                                      \
  b200sp_status b200sp_spmv_ell_##sfx(                                             \
      b200sp_handle h, b200sp_stream stream, int64_t num_rows, int64_t num_cols,   \
      int64_t num_cols_per_row, int64_t pitch, const int32_t *column_indices,      \
      const T *values, const T *x, T *y, int accumulate, const b200sp_cfg *cfg);   \
  b200sp_status b200sp_spmv_dia_##sfx(                                             \
      b200sp_handle h, b200sp_stream stream, int64_t num_rows, int64_t num_cols,   \
      int64_t num_diagonals, int64_t pitch, const int32_t *diagonal_offsets,       \
      const T *values, const T *x, T *y, int accumulate, const b200sp_cfg *cfg);   \
  b200sp_status b200sp_spmv_coo_##sfx(                                             \
      b200sp_handle h, b200sp_stream stream, int64_t num_rows, int64_t num_cols,   \
      int64_t num_entries, const int32_t *row_indices,                             \
      const int32_t *column_indices, const T *values, const T *x, T *y,            \
      int accumulate, const b200sp_cfg *cfg);                                      \
  b200sp_status b200sp_spmv_hyb_##sfx(                                             \
      b200sp_handle h, b200sp_stream stream, int64_t num_rows, int64_t num_cols,   \
      int64_t ell_cols_per_row, int64_t ell_pitch,                                 \
      const int32_t *ell_column_indices, const T *ell_values,                      \
      int64_t coo_num_entries, const int32_t *coo_row_indices,                     \
      const int32_t *coo_column_indices, const T *coo_values, const T *x, T *y,    \
      int accumulate, const b200sp_cfg *ell_cfg, const b200sp_cfg *coo_cfg);       \
  /* ELL-R (cusp::ktt::ellr_matrix, cusp/ktt/ellr_matrix.h:18): ELL plus         \
   * row_lengths[num_rows]; the kernel stops at row_lengths[row]. */              \
  b200sp_status b200sp_spmv_ellr_##sfx(                                            \
      b200sp_handle h, b200sp_stream stream, int64_t num_rows, int64_t num_cols,   \
      int64_t num_cols_per_row, int64_t pitch, const int32_t *column_indices,      \
      const T *values, const int32_t *row_lengths, const T *x, T *y,               \
      int accumulate, const b200sp_cfg *cfg);

B200SP_DECL_SPMV(float, f32)
B200SP_DECL_SPMV(double, f64)
#undef B200SP_DECL_SPMV

/* Preprocessing of the balanced CSR kernel (cusp/system/cuda/ktt/csr_multiply.h:38-85, cpu_compute_row_starts /
 * gpu_compute_row_starts): row_starts[w] = the row containing entry w * chunk, chunk = ceil(num_entries / workers);
 * 0 for workers that start beyond the last entry.  K_CSR_BALANCED computes this array (chunk = its tile size) before
 * every product; the entry point exposes it (bit-exact parity with the reference's host version). */
b200sp_status b200sp_csr_row_starts(b200sp_handle h, b200sp_stream stream, int64_t num_rows,
                                    int64_t num_entries, const int32_t *row_offsets, int64_t workers,
                                    int32_t *row_starts);

/* row_lengths for ELL-R: count of leading non-negative column slots per row
 * (cusp/ktt/detail/ellr_matrix.inl:16-52). */
b200sp_status b200sp_ell_row_lengths(b200sp_handle h, b200sp_stream stream,
                                     int64_t num_rows, int64_t num_cols_per_row,
                                     int64_t pitch, const int32_t *column_indices,
                                     int32_t *row_lengths);

/* ---- BLAS-1 (cusp::blas::*, cusp/detail/blas.inl:84-461,
 *      cusp/system/detail/generic/blas.h:180-340) ---------------------------
 * Reductions are deterministic (fixed two-level order).  `result_dev` may be
 * NULL; `result_host` may be NULL.  If result_host != NULL the call
 * synchronises `stream` (the reference's reductions always block).
 */
#define B200SP_DECL_BLAS(T, sfx)                                                   \
  /* y <- alpha*x + y            generic/blas.h:180-198 */                         \
  b200sp_status b200sp_axpy_##sfx(b200sp_handle h, b200sp_stream s, int64_t n,     \
                                  T alpha, const T *x, T *y);                      \
  /* z <- alpha*x + beta*y       generic/blas.h:200-220 */                         \
  b200sp_status b200sp_axpby_##sfx(b200sp_handle h, b200sp_stream s, int64_t n,    \
                                   T alpha, const T *x, T beta, const T *y, T *z); \
  /* y <- x                      generic/blas.h copy */                            \
  b200sp_status b200sp_copy_##sfx(b200sp_handle h, b200sp_stream s, int64_t n,     \
                                  const T *x, T *y);                               \
  b200sp_status b200sp_fill_##sfx(b200sp_handle h, b200sp_stream s, int64_t n,     \
                                  T alpha, T *x);                                  \
  b200sp_status b200sp_scal_##sfx(b200sp_handle h, b200sp_stream s, int64_t n,     \
                                  T alpha, T *x);                                  \
  /* out <- alpha*x + beta*y + gamma*z   generic/blas.h:99-117,222-246 */          \
  b200sp_status b200sp_axpbypcz_##sfx(b200sp_handle h, b200sp_stream s, int64_t n, \
                                      T alpha, const T *x, T beta, const T *y,     \
                                      T gamma, const T *z, T *out);                \
  /* z <- x .* y                 generic/blas.h:119-127,248-266 */                 \
  b200sp_status b200sp_xmy_##sfx(b200sp_handle h, b200sp_stream s, int64_t n,      \
                                 const T *x, const T *y, T *z);                    \
  /* sum |x_i| (asum == nrm1)    generic/blas.h:160-176 */                         \
  b200sp_status b200sp_asum_##sfx(b200sp_handle h, b200sp_stream s, int64_t n,     \
                                  const T *x, T *result_dev, T *result_host);      \
  /* max |x_i|                   generic/blas.h nrmmax */                          \
  b200sp_status b200sp_nrmmax_##sfx(b200sp_handle h, b200sp_stream s, int64_t n,   \
                                    const T *x, T *result_dev, T *result_host);    \
  /* index of the first max |x_i| generic/blas.h:138-158; synchronises */          \
  b200sp_status b200sp_amax_##sfx(b200sp_handle h, b200sp_stream s, int64_t n,     \
                                  const T *x, int *index_host);                    \
  /* sum x_i*y_i                 generic/blas.h:284-313 (dot == dotc for reals) */ \
  b200sp_status b200sp_dot_##sfx(b200sp_handle h, b200sp_stream s, int64_t n,      \
                                 const T *x, const T *y, T *result_dev,            \
                                 T *result_host);                                  \
  /* sqrt(sum |x_i|^2), unscaled generic/blas.h:328-340 */                         \
  b200sp_status b200sp_nrm2_##sfx(b200sp_handle h, b200sp_stream s, int64_t n,     \
                                  const T *x, T *result_dev, T *result_host);

B200SP_DECL_BLAS(float, f32)
B200SP_DECL_BLAS(double, f64)
#undef B200SP_DECL_BLAS

/* ---- sparse matrix descriptor (for CG / tuning / host-buffer calls) ------ */
typedef enum {
  B200SP_FMT_CSR = 0,
  B200SP_FMT_ELL = 1,
  B200SP_FMT_DIA = 2,
  B200SP_FMT_COO = 3,
  B200SP_FMT_HYB = 4,
  B200SP_FMT_ELLR = 5
} b200sp_format;

typedef enum { B200SP_F32 = 0, B200SP_F64 = 1 } b200sp_dtype;

/* Non-owning view of a device-resident matrix.  Unused fields are 0/NULL.
 *   CSR : row_offsets[num_rows+1], column_indices[nnz], values[nnz]
 *   ELL : column_indices[pitch*K], values[pitch*K], K = num_cols_per_row
 *   ELLR: ELL + row_offsets := row_lengths[num_rows]
 *   DIA : diagonal_offsets[K], values[pitch*K],     K = num_cols_per_row (#diagonals)
 *   COO : row_indices[nnz], column_indices[nnz], values[nnz] (sorted by row)
 *   HYB : ELL fields + coo_* fields                                          */
typedef struct {
  b200sp_format format;
  b200sp_dtype dtype;
  int64_t num_rows, num_cols, num_entries;
  int64_t num_cols_per_row, pitch;
  const int32_t *row_offsets;
  const int32_t *row_indices;
  const int32_t *column_indices;
  const int32_t *diagonal_offsets;
  const void *values;
  int64_t coo_num_entries;
  const int32_t *coo_row_indices;
  const int32_t *coo_column_indices;
  const void *coo_values;
} b200sp_matrix;

/* generic dispatch on the descriptor (same kernels as the typed calls) */
b200sp_status b200sp_spmv(b200sp_handle h, b200sp_stream stream,
                          const b200sp_matrix *A, const void *x, void *y,
                          int accumulate, const b200sp_cfg *cfg);

/* ---- generalized product: y[i] = reduce(initialize(y[i]), combine(a_ij, x_j) ...) ---------------------
 * cusp::multiply(A, x, y, initialize, combine, reduce) / cusp::generalized_spmv (cusp/multiply.h:163-280,
 * cusp/system/detail/generic/multiply/generalized_spmv.h:61-303; the reference's device kernels are templated on
 * the functor triple, e.g. cuda/detail/multiply/csr_vector_spmv.h:66-161).  A C ABI cannot carry C++ functor
 * objects, so the triple is named by code; include/cusp/multiply.h maps cusp:: / thrust:: / std:: functor types to
 * these codes and throws cusp::not_implemented_exception for anything else.  Every format of b200sp_matrix.
 * (multiplies, plus) with constant(0) | identity is b200sp_spmv.  ELL / DIA / short-row CSR keep the host loop's
 * order of operations (bit-identical for every pair); long-row CSR and COO regroup (exact for minimum / maximum). */
typedef enum { B200SP_INIT_CONSTANT = 0 /* cusp::constant_functor(init_value) */, B200SP_INIT_IDENTITY = 1 } b200sp_init_op;
typedef enum {
  B200SP_COMBINE_MULTIPLIES = 0, /* a * x                       */
  B200SP_COMBINE_PLUS = 1,       /* a + x        (min-plus / max-plus semirings) */
  B200SP_COMBINE_MINIMUM = 2,    /* min(a, x)                   */
  B200SP_COMBINE_MAXIMUM = 3,    /* max(a, x)                   */
  B200SP_COMBINE_PROJECT2ND = 4  /* x            (structure-only products: reachability, label propagation) */
} b200sp_combine_op;
typedef enum { B200SP_REDUCE_PLUS = 0, B200SP_REDUCE_MINIMUM = 1, B200SP_REDUCE_MAXIMUM = 2 } b200sp_reduce_op;
typedef struct {
  int initialize;    /* b200sp_init_op    */
  double init_value; /* B200SP_INIT_CONSTANT: the constant (converted to A.dtype) */
  int combine;       /* b200sp_combine_op */
  int reduce;        /* b200sp_reduce_op  */
} b200sp_functors;
b200sp_status b200sp_spmv_generalized(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A,
                                      const void *x, void *y, const b200sp_functors *functors);

/* Same product through HOST buffers: x_host -> device, SpMV, y -> y_host.
 * The matrix stays device-resident like a cusp::*_matrix<.., device_memory>;
 * this is what `cusp::array1d<T,host_memory> y_h = y_d` after cusp::multiply
 * costs end to end.  Synchronises `stream`. */
b200sp_status b200sp_spmv_host(b200sp_handle h, b200sp_stream stream,
                               const b200sp_matrix *A, const void *x_host,
                               void *y_host, int accumulate, const b200sp_cfg *cfg);

/* ---- captured products: the launch-bound regime ---------------------------------------------------------
 * A product on an L2-resident operator (BASELINE configs[0], poisson5pt 512^2: 21 MB) takes ~5 us on the GPU and
 * ~10 us of host time per call.  b200sp_spmv_graph_create captures `count` back-to-back products y = A x (fixed
 * pointers: a solver's inner loop, a power iteration) into one CUDA graph; b200sp_graph_launch replays it with one
 * host call.  Same kernels and results as `count` calls of b200sp_spmv.  The graph borrows every array of A, x
 * and y.  b200sp_cg does the same internally for small systems (check_interval iterations per replay). */
typedef struct b200sp_graph_s *b200sp_graph;
b200sp_status b200sp_spmv_graph_create(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A,
                                       const void *x, void *y, int accumulate, const b200sp_cfg *cfg,
                                       int count, b200sp_graph *out);
b200sp_status b200sp_graph_launch(b200sp_handle h, b200sp_stream stream, b200sp_graph graph);
b200sp_status b200sp_graph_destroy(b200sp_handle h, b200sp_graph graph);

/* ---- conjugate gradients -------------------------------------------------
 * cusp::krylov::cg(A, x, b, monitor) with the identity preconditioner
 * (cusp/krylov/detail/cg.inl:35-107) and cusp::monitor semantics
 * (cusp/detail/monitor.inl:107-111,178-208): stop when
 *   ||r||_2 <= absolute_tolerance + relative_tolerance*||b||_2
 * or iteration_count >= iteration_limit; the residual norm is recorded before
 * every iteration and once more at exit, exactly like monitor.residuals.
 * Same iterate sequence as the reference (same operation order per entry);
 * fused into 3 kernels / iteration with device-resident scalars, chained by programmatic
 * dependent launch; on systems of up to 2^22 rows the check_interval iterations between two
 * host polls are replayed from one CUDA graph (launch-bound regime).
 */
typedef struct {
  int64_t iteration_limit;   /* monitor default 500  */
  double relative_tolerance; /* monitor default 1e-5 */
  double absolute_tolerance; /* monitor default 0    */
  int check_interval;        /* iterations between host convergence polls; 0 -> 16 */
} b200sp_cg_params;

typedef struct {
  int64_t iteration_count;
  int converged; /* residual_norm <= tolerance */
  double residual_norm;
  double b_norm;
  int64_t num_residuals; /* entries written to residuals_host */
} b200sp_cg_result;

/* x (in: initial guess, out: solution) and b are device vectors of A.dtype.
 * residuals_host: optional host array with room for iteration_limit+1 doubles. */
b200sp_status b200sp_cg(b200sp_handle h, b200sp_stream stream,
                        const b200sp_matrix *A, void *x, const void *b,
                        const b200sp_cg_params *params, const b200sp_cfg *spmv_cfg,
                        b200sp_cg_result *result, double *residuals_host);

/* ---- CSR x dense block ("block SpMV", cusp::multiply(csr_matrix, array2d, array2d)) ----------
 * Y[num_rows x block_cols] = (accumulate ? Y : 0) + A * X[num_cols x block_cols]; X and Y are
 * row-major with leading dimensions ldx, ldy (elements).  Replaces BlockSpmvKernel /
 * __spmv_csr_block (cusp/system/cuda/detail/multiply/csr_block_spmv.h:36-222); per (row, column)
 * the entries are added in storage order like the host loop
 * (cusp/system/detail/sequential/multiply/csr_block_spmv.h:52-77), so results are bit-identical
 * to it (a single contiguous column, block_cols == ldx == ldy == 1, is handed to b200sp_spmv_csr and
 * follows its parity rule).  Any block width (the reference's kernel handles at most 32 columns). */
b200sp_status b200sp_spmm_csr_f32(b200sp_handle h, b200sp_stream stream, int64_t num_rows,
                                  int64_t num_cols, int64_t num_entries,
                                  const int32_t *row_offsets, const int32_t *column_indices,
                                  const float *values, int64_t block_cols, const float *X,
                                  int64_t ldx, float *Y, int64_t ldy, int accumulate);
b200sp_status b200sp_spmm_csr_f64(b200sp_handle h, b200sp_stream stream, int64_t num_rows,
                                  int64_t num_cols, int64_t num_entries,
                                  const int32_t *row_offsets, const int32_t *column_indices,
                                  const double *values, int64_t block_cols, const double *X,
                                  int64_t ldx, double *Y, int64_t ldy, int accumulate);

/* ---- inspector / executor COO product for gather-bound operators (power-law graphs) ----------------
 * The analogue of tuning tied to one matrix (cusp::ktt::tune(A, x, y) keeps its results per kernel and
 * matrix, cusp/ktt/detail/ktt.inl:108-142).  b200sp_coo_plan_create inspects the sparsity pattern once
 * (column histogram -> the most frequent columns that fit a `table_bytes` shared-memory table, default:
 * all of an SM's shared memory -> a second copy of column_indices in which those columns are replaced by
 * a table slot); the executor keeps x of those columns in shared memory, so their gathers never reach
 * the L1 / L2 path that bounds COO on an R-MAT.  Same tiles, summation order and carry
 * fix-up as K_COO_WARP with equal (vector_width, unroll): identical bits.  The plan borrows row_indices
 * and remembers column_indices by address (caller keeps both alive and unchanged while the plan exists);
 * values are passed per call and may change between calls.  dtype fixes the table's element size.
 *
 * b200sp_coo_plan_attach registers the plan with the handle: from then on b200sp_spmv_coo_<t> /
 * b200sp_spmv (COO, and the COO part of HYB) called with exactly the plan's (row_indices, column_indices,
 * num_entries, shape) runs the executor; any other arrays take the default kernels.  Detach (or destroy)
 * before changing the matrix's structure in place.  cusp::ktt::tune on a device coo_matrix attaches a
 * plan, cusp::ktt::reset_tuning drops it (include/cusp/ktt/ktt.h). */
typedef struct b200sp_coo_plan_s *b200sp_coo_plan;
b200sp_status b200sp_coo_plan_create(b200sp_handle h, b200sp_stream stream, int64_t num_rows,
                                     int64_t num_cols, int64_t num_entries,
                                     const int32_t *row_indices, const int32_t *column_indices,
                                     b200sp_dtype dtype, int64_t table_bytes, b200sp_coo_plan *out);
b200sp_status b200sp_coo_plan_destroy(b200sp_handle h, b200sp_coo_plan plan);
b200sp_status b200sp_coo_plan_attach(b200sp_handle h, b200sp_coo_plan plan);
b200sp_status b200sp_coo_plan_detach(b200sp_handle h, b200sp_coo_plan plan);
/* columns held in the table, stored entries they serve, table capacity in elements */
b200sp_status b200sp_coo_plan_info(b200sp_coo_plan plan, int64_t *hot_columns, int64_t *hot_entries,
                                   int64_t *capacity);
/* cfg: NULL or a K_COO_WARP shape (vector_width, unroll, stages) */
b200sp_status b200sp_spmv_coo_plan_f32(b200sp_handle h, b200sp_stream stream, b200sp_coo_plan plan,
                                       const float *values, const float *x, float *y, int accumulate,
                                       const b200sp_cfg *cfg);
b200sp_status b200sp_spmv_coo_plan_f64(b200sp_handle h, b200sp_stream stream, b200sp_coo_plan plan,
                                       const double *values, const double *x, double *y, int accumulate,
                                       const b200sp_cfg *cfg);

/* ---- multi-GPU: row-block partitioned operator ---------------------------
 * One process per GPU.  Each rank owns a contiguous block of rows of A and the
 * matching slices of x, b.  Column j of the global matrix is owned by the rank
 * whose row block contains j (square operator).  A rank's local matrix has
 * num_rows = local rows and addresses x through a window
 *        [halo_lo | local | halo_hi]
 * i.e. local column index c (0-based in the window) = global j - (row_begin - halo_lo).
 * For DIA the diagonal offsets are unchanged and the kernel is simply given
 * x_window + halo_lo.  Halo planes are exchanged with the two neighbouring
 * ranks before every SpMV (NCCL send/recv on `stream`), dot products are
 * all-reduced with NCCL.  The communicator is created from a 128-byte NCCL
 * unique id that the host distributes (any out-of-band channel: MPI, a file, a socket).
 */
#define B200SP_NCCL_UNIQUE_ID_BYTES 128
b200sp_status b200sp_comm_unique_id(void *id128);
b200sp_status b200sp_comm_init(b200sp_handle h, const void *id128, int world_size,
                               int rank);
b200sp_status b200sp_comm_destroy(b200sp_handle h);
/* 1 when the NVLink peer-memory path is active for this communicator: every rank's
 * mailbox is CUDA-IPC mapped into every peer, b200sp_cg_dist then stores halo planes and
 * the two CG scalars straight into peer memory from inside its three kernels (no NCCL
 * call per iteration).  0: NCCL send/recv + all-reduce per iteration (also forced by the
 * environment variable B200SP_DISABLE_P2P=1). */
int b200sp_comm_p2p_enabled(b200sp_handle h);
/* Number of cross-GPU waits of the peer-memory path that gave up (a peer never published
 * its data within ~4 s).  Non-zero means results of `_dist` calls since b200sp_comm_init
 * are invalid; b200sp_cg_dist checks it itself and returns B200SP_COMM_ERROR.  Synchronises
 * `stream`.  Always 0 on the NCCL path. */
int64_t b200sp_comm_timeouts(b200sp_handle h, b200sp_stream stream);

typedef struct {
  int64_t halo_lo; /* elements received from rank-1 (0 on rank 0)            */
  int64_t halo_hi; /* elements received from rank+1 (0 on the last rank)     */
} b200sp_halo;

/* Partitioned CG: A is the LOCAL block (num_rows = local rows, num_cols =
 * halo_lo + local + halo_hi window width), x/b are local slices. */
b200sp_status b200sp_cg_dist(b200sp_handle h, b200sp_stream stream,
                             const b200sp_matrix *A_local, const b200sp_halo *halo,
                             void *x_local, const void *b_local,
                             const b200sp_cg_params *params,
                             const b200sp_cfg *spmv_cfg, b200sp_cg_result *result,
                             double *residuals_host);
/* Partitioned SpMV (halo exchange + local product); x_window has room for
 * halo_lo + local + halo_hi elements with the local slice already in place. */
b200sp_status b200sp_spmv_dist(b200sp_handle h, b200sp_stream stream,
                               const b200sp_matrix *A_local, const b200sp_halo *halo,
                               void *x_window, void *y_local,
                               const b200sp_cfg *cfg);

/* b200sp_spmv_dist through HOST buffers: this rank's slice of x (num_rows elements) goes up, the halo
 * planes are exchanged between the GPUs, this rank's slice of y comes down.  For DIA blocks the call is
 * a pipeline over row chunks (edge pieces first, halo exchange and interior uploads overlapped, both
 * PCIe directions busy at once), bit-identical to upload + b200sp_spmv_dist + download.  Collective:
 * every rank of the communicator calls it.  Synchronises `stream`. */
b200sp_status b200sp_spmv_dist_host(b200sp_handle h, b200sp_stream stream,
                                    const b200sp_matrix *A_local, const b200sp_halo *halo,
                                    const void *x_host_local, void *y_host_local,
                                    const b200sp_cfg *cfg);

/* Partitioned SpMV for operators whose rows read the whole of x (graphs): the rank's
 * block of rows keeps GLOBAL column indices (num_cols = global size); x is partitioned like the
 * rows, slice r = [slice_offsets[r], slice_offsets[r+1]) (host array of world_size+1 entries,
 * slice_offsets[0] = 0, slice_offsets[world_size] = num_cols).  x_full has room for num_cols
 * elements with this rank's slice in place; the call gathers the other slices (one kernel
 * pulling them from the peers' IPC-shared staging buffers over NVLink, or grouped NCCL
 * send/recv) and then runs the local product.  Every y entry is computed by exactly one rank. */
b200sp_status b200sp_spmv_dist_gather(b200sp_handle h, b200sp_stream stream,
                                      const b200sp_matrix *A_local, const int64_t *slice_offsets,
                                      void *x_full, void *y_local, const b200sp_cfg *cfg);

/* ---- the other Krylov solvers + the diagonal (Jacobi) preconditioner, fused like CG ------------------
 * cusp::krylov::cg with cusp::precond::diagonal (cusp/krylov/detail/cg.inl:35-107, cusp/precond/diagonal.h),
 * cusp::krylov::bicgstab (cusp/krylov/detail/bicgstab.inl:35-123), cusp::krylov::cr (cusp/krylov/detail/cr.inl:39-128)
 * with cusp::monitor semantics.  Scalars and the monitor live on the device; an iteration is the products (dot
 * products fused into their epilogues) plus 2-3 fused vector kernels; the host polls every check_interval
 * iterations.  Same element expressions in the same order as the reference: same iterate sequence (the dot
 * products are summed in a different order).  residuals_host: monitor.residuals — one entry per
 * monitor.finished() call (BiCGStab calls it twice per iteration: room for 2 * iteration_limit + 2).
 *   diagonal_inverse  NULL: identity preconditioner; else the device vector cusp::precond::diagonal holds
 *                     (1 / a_ii, this rank's rows), applied as z_i = diagonal_inverse_i * r_i
 *   halo              NULL: one GPU; else the row-block partitioned form of b200sp_cg_dist (A_local in window
 *                     coordinates, halo planes exchanged before every product, sums all-reduced)
 * B200SP_SOLVER_CG with diagonal_inverse == NULL is b200sp_cg / b200sp_cg_dist. */
typedef enum { B200SP_SOLVER_CG = 0, B200SP_SOLVER_BICGSTAB = 1, B200SP_SOLVER_CR = 2 } b200sp_solver;
b200sp_status b200sp_krylov(b200sp_handle h, b200sp_stream stream, b200sp_solver solver,
                            const b200sp_matrix *A, const b200sp_halo *halo, void *x, const void *b,
                            const void *diagonal_inverse, const b200sp_cg_params *params,
                            const b200sp_cfg *spmv_cfg, b200sp_cg_result *result, double *residuals_host);

/* ---- autotuning (cusp::ktt::{multiply,tune,reset_tuning},
 *      cusp/ktt/detail/ktt.inl:83-142, cuda/ktt/multiply.h:56-153) ----------- */
typedef enum {
  B200SP_TUNE_OK = 0,
  B200SP_TUNE_LAUNCH_FAILED = 1,     /* ~ ktt::ResultStatus::ComputationFailed   */
  B200SP_TUNE_VALIDATION_FAILED = 2, /* ~ ktt::ResultStatus::ValidationFailed    */
  B200SP_TUNE_UNSUPPORTED = 3        /* ~ ktt::ResultStatus::DeviceLimitsExceeded */
} b200sp_tune_status;

typedef struct {
  b200sp_cfg cfg;
  b200sp_tune_status status;
  float milliseconds; /* mean over `repeats` launches */
  double max_rel_error; /* vs the reference output */
} b200sp_tune_result;

/* enumerate the valid configurations for (format,dtype); returns the count and
 * fills up to `capacity` entries of `out` (out may be NULL to query). */
int64_t b200sp_cfg_space(b200sp_format format, b200sp_dtype dtype, b200sp_cfg *out,
                         int64_t capacity);

/* Exhaustive offline tuning (cusp::ktt::tune): runs every configuration
 * `repeats` times, validates y against `y_reference` (device, may be NULL ->
 * the engine-default configuration's output is the reference) with
 * |y - y_ref| <= tol*|y_ref| + tol, records results, stores the winner in the
 * handle's tuning cache (key: format, dtype, log2 rows bucket, nnz/row bucket and the
 * structure class the engine's own defaults distinguish — CSR: banded / scattered /
 * skewed row lengths, COO and HYB tails: banded / one entry per row / scattered columns —
 * so a winner is only replayed on matrices of the kind it was found on)
 * and returns it in *best.  y is restored semantics-wise: on return it holds
 * A x computed by the best configuration. */
b200sp_status b200sp_tune(b200sp_handle h, b200sp_stream stream,
                          const b200sp_matrix *A, const void *x, void *y,
                          const void *y_reference, double tol, int repeats,
                          b200sp_tune_result *results, int64_t capacity,
                          int64_t *num_results, b200sp_cfg *best);
/* The same search with the two hooks cusp::ktt::tune passes down (cusp/system/cuda/ktt/multiply.h:106-153):
 *   order / order_len  the searcher: indices into b200sp_cfg_space() in the order to visit them (a permutation for
 *                      ::ktt::RandomSearcher, a prefix or subset for a budgeted search); NULL = the whole space in
 *                      its own order (::ktt::DeterministicSearcher)
 *   callback / user    the stop condition: called after every configuration with its result; a non-zero return
 *                      ends the search (Tune(kernel, stop_condition)).  The winner among the configurations visited
 *                      so far is cached and returned like b200sp_tune does. */
typedef int (*b200sp_tune_callback)(const b200sp_tune_result *result, void *user);
b200sp_status b200sp_tune_ex(b200sp_handle h, b200sp_stream stream, const b200sp_matrix *A, const void *x,
                             void *y, const void *y_reference, double tol, int repeats, const int64_t *order,
                             int64_t order_len, b200sp_tune_callback callback, void *user,
                             b200sp_tune_result *results, int64_t capacity, int64_t *num_results,
                             b200sp_cfg *best);
/* One step of dynamic tuning (cusp::ktt::multiply(A,x,y) / what plain
 * cusp::multiply does for ELL & DIA when ktt is enabled, cuda/ktt/multiply.h:56-77):
 * runs the next untried configuration (timed), or the best one when the space
 * is exhausted.  result may be NULL. */
b200sp_status b200sp_tune_step(b200sp_handle h, b200sp_stream stream,
                               const b200sp_matrix *A, const void *x, void *y,
                               b200sp_tune_result *result);
/* cusp::ktt::reset_tuning: forget cached winners / dynamic-tuning progress for
 * the class of A (A == NULL: everything). */
b200sp_status b200sp_tune_reset(b200sp_handle h, const b200sp_matrix *A);
/* look up the cached winner for A's class; returns 1 and fills *cfg if present */
int b200sp_tune_lookup(b200sp_handle h, const b200sp_matrix *A, b200sp_cfg *cfg);
/* persist / restore the tuning cache (SURVEY §5 "checkpoint": the fork never
 * saves KTT results) as a small text file */
b200sp_status b200sp_tune_save(b200sp_handle h, const char *path);
b200sp_status b200sp_tune_load(b200sp_handle h, const char *path);

/* ---- format conversions on the device (cusp::convert) ----------------
 * Layouts are the reference's, bit for bit (cusp/system/detail/generic/conversions/
 * csr_to_other.h:73-306, cusp/system/detail/generic/format_utils.inl:36-110,281-321).
 * The caller sizes the outputs from b200sp_csr_convert_query. */
typedef struct {
  int64_t max_entries_per_row; /* cusp::compute_max_entries_per_row                      */
  int64_t hyb_entries_per_row; /* cusp::compute_optimal_entries_per_row(relative_speed,
                                  breakeven_threshold): ELL width of the HYB split        */
  int64_t hyb_coo_entries;     /* entries left for the COO part at that width            */
  int64_t num_diagonals;       /* cusp::count_diagonals (0 if column_indices == NULL)    */
} b200sp_convert_info;

/* row_indices[k] = i for row_offsets[i] <= k < row_offsets[i+1]   (CSR -> COO) */
b200sp_status b200sp_offsets_to_indices(b200sp_handle h, b200sp_stream stream, int64_t num_rows,
                                        const int32_t *row_offsets, int32_t *row_indices);
/* row_offsets[i] = #(row_indices < i), i = 0..num_rows; row_indices sorted  (COO -> CSR) */
b200sp_status b200sp_indices_to_offsets(b200sp_handle h, b200sp_stream stream, int64_t num_rows,
                                        int64_t num_entries, const int32_t *row_indices,
                                        int32_t *row_offsets);
/* structure of a CSR matrix that the conversions need; the reference's defaults are
 * relative_speed = 3, breakeven_threshold = 4096 (csr_to_other.h:250-253).
 * Returns scalars to the host: synchronises `stream`. */
b200sp_status b200sp_csr_convert_query(b200sp_handle h, b200sp_stream stream, int64_t num_rows,
                                       int64_t num_cols, int64_t num_entries,
                                       const int32_t *row_offsets, const int32_t *column_indices,
                                       float relative_speed, int64_t breakeven_threshold,
                                       b200sp_convert_info *info_host);

#define B200SP_DECL_CONVERT(T, sfx)                                                       \
  /* k-th entry of row i -> slot [k*pitch + i] for k < num_cols_per_row; padding column   \
   * -1 / value 0, rows [num_rows, pitch) padded too          csr_to_other.h:155-227 */   \
  b200sp_status b200sp_csr_to_ell_##sfx(                                                  \
      b200sp_handle h, b200sp_stream s, int64_t num_rows, int64_t num_cols_per_row,       \
      int64_t pitch, const int32_t *row_offsets, const int32_t *column_indices,           \
      const T *values, int32_t *ell_column_indices, T *ell_values);                       \
  /* entries k >= num_cols_per_row of every row, in CSR order, as COO (the tail of the    \
   * HYB split)                                               csr_to_other.h:229-306 */   \
  b200sp_status b200sp_csr_to_coo_tail_##sfx(                                             \
      b200sp_handle h, b200sp_stream s, int64_t num_rows, int64_t num_cols_per_row,       \
      const int32_t *row_offsets, const int32_t *column_indices, const T *values,         \
      int32_t *coo_row_indices, int32_t *coo_column_indices, T *coo_values);              \
  /* occupied diagonals ascending in diagonal_offsets[num_diagonals]; values[d*pitch+i],  \
   * zero elsewhere                                           csr_to_other.h:73-153 */    \
  b200sp_status b200sp_csr_to_dia_##sfx(                                                  \
      b200sp_handle h, b200sp_stream s, int64_t num_rows, int64_t num_cols,               \
      int64_t num_diagonals, int64_t pitch, const int32_t *row_offsets,                   \
      const int32_t *column_indices, const T *values, int32_t *diagonal_offsets,          \
      T *dia_values);                                                                     \
  /* number of stored values equal to 0 (ELL num_entries = nnz - zeros, :205-212) */      \
  b200sp_status b200sp_count_zeros_##sfx(b200sp_handle h, b200sp_stream s, int64_t n,     \
                                         const T *values, int64_t *count_host);
B200SP_DECL_CONVERT(float, f32)
B200SP_DECL_CONVERT(double, f64)
#undef B200SP_DECL_CONVERT

/* DIA / ELL / HYB sources: to CSR on the device (and from there to every other format with the calls above), in two
 * steps because the caller owns the outputs: `_offsets` computes row_offsets[num_rows + 1] and returns the number of
 * kept entries to the host (synchronises), `_fill` writes column_indices / values of that size.
 *   DIA  row-major scan of the [rows x diagonals] slab, keeps value != 0, column = row + diagonal_offsets[d]
 *        cusp/system/detail/generic/conversions/dia_to_other.h:61-161
 *   ELL  row-major scan of the [rows x K] slabs, keeps value != 0                           ell_to_other.h:55-143
 *   HYB  per row the ELL entries with a valid column merged by column with the row's COO entries (ties: ELL first)
 *        hyb_to_other.h:45-56, cusp/detail/coo_matrix.inl:269-341
 * b200sp_dia_to_ell: the fork's direct DIA -> ELL (dia_to_other.h:163-251): K = num_diagonals, pitch = the DIA pitch,
 * non-zero values left-packed per row in diagonal order, padding column -1 / value 0 — one thread per row instead of
 * two thrust::stable_partition calls per row. */
#define B200SP_DECL_TO_CSR(T, sfx)                                                                    \
  b200sp_status b200sp_dia_to_csr_offsets_##sfx(b200sp_handle h, b200sp_stream s, int64_t num_rows,   \
                                                int64_t num_diagonals, int64_t pitch, const T *values, \
                                                int32_t *row_offsets, int64_t *num_entries_host);     \
  b200sp_status b200sp_dia_to_csr_fill_##sfx(b200sp_handle h, b200sp_stream s, int64_t num_rows,      \
                                             int64_t num_diagonals, int64_t pitch,                    \
                                             const int32_t *diagonal_offsets, const T *values,        \
                                             const int32_t *row_offsets, int32_t *column_indices,     \
                                             T *csr_values);                                          \
  b200sp_status b200sp_ell_to_csr_offsets_##sfx(b200sp_handle h, b200sp_stream s, int64_t num_rows,   \
                                                int64_t num_cols_per_row, int64_t pitch,              \
                                                const int32_t *ell_column_indices,                    \
                                                const T *ell_values, int32_t *row_offsets,            \
                                                int64_t *num_entries_host);                           \
  b200sp_status b200sp_ell_to_csr_fill_##sfx(b200sp_handle h, b200sp_stream s, int64_t num_rows,      \
                                             int64_t num_cols_per_row, int64_t pitch,                 \
                                             const int32_t *ell_column_indices, const T *ell_values,  \
                                             const int32_t *row_offsets, int32_t *column_indices,     \
                                             T *csr_values);                                          \
  b200sp_status b200sp_hyb_to_csr_offsets_##sfx(                                                      \
      b200sp_handle h, b200sp_stream s, int64_t num_rows, int64_t ell_cols_per_row,                   \
      int64_t ell_pitch, const int32_t *ell_column_indices, const T *ell_values,                      \
      int64_t coo_num_entries, const int32_t *coo_row_indices, int32_t *row_offsets,                  \
      int64_t *num_entries_host);                                                                     \
  b200sp_status b200sp_hyb_to_csr_fill_##sfx(                                                         \
      b200sp_handle h, b200sp_stream s, int64_t num_rows, int64_t ell_cols_per_row,                   \
      int64_t ell_pitch, const int32_t *ell_column_indices, const T *ell_values,                      \
      int64_t coo_num_entries, const int32_t *coo_row_indices, const int32_t *coo_column_indices,     \
      const T *coo_values, const int32_t *row_offsets, int32_t *column_indices, T *csr_values);       \
  b200sp_status b200sp_dia_to_ell_##sfx(b200sp_handle h, b200sp_stream s, int64_t num_rows,           \
                                        int64_t num_diagonals, int64_t pitch,                         \
                                        const int32_t *diagonal_offsets, const T *values,             \
                                        int32_t *ell_column_indices, T *ell_values);
B200SP_DECL_TO_CSR(float, f32)
B200SP_DECL_TO_CSR(double, f64)
#undef B200SP_DECL_TO_CSR

/* ---- device-side input builders (cusp::gallery::poisson5pt/7pt via
 *      generate_matrix_from_stencil, gallery/detail/stencil.inl:143-206, then
 *      cusp::convert; produce bit-identical arrays to that pipeline without
 *      the O(rows) launches of conversions/dia_to_other.h:227-251) -----------
 * grid = (nx, ny, nz) with nz == 1 for the 2-D 5-point stencil; row index
 * = ix + nx*(iy + ny*iz).  Rows [row_begin, row_begin+num_rows) of the global
 * operator are produced (row-block partition); column indices are global minus
 * `col_shift`.  `stencil` = 5 or 7.
 */
#define B200SP_DECL_GALLERY(T, sfx)                                                \
  /* DIA: values[K*pitch] (K = stencil), offsets[K] ascending; the stored        \
   * offset is global_offset + row_begin - col_shift (window coordinates) */       \
  b200sp_status b200sp_poisson_dia_##sfx(                                          \
      b200sp_handle h, b200sp_stream s, int stencil, int64_t nx, int64_t ny,       \
      int64_t nz, int64_t row_begin, int64_t num_rows, int64_t col_shift,          \
      int64_t pitch, int32_t *diagonal_offsets, T *values);                                       \
  /* ELL: left-packed, K = stencil columns, pad col=-1 val=0 */                    \
  b200sp_status b200sp_poisson_ell_##sfx(                                          \
      b200sp_handle h, b200sp_stream s, int stencil, int64_t nx, int64_t ny,       \
      int64_t nz, int64_t row_begin, int64_t num_rows, int64_t col_shift,          \
      int64_t pitch, int32_t *column_indices, T *values);                          \
  /* CSR: row_offsets must be computed first with b200sp_poisson_csr_offsets */    \
  b200sp_status b200sp_poisson_csr_##sfx(                                          \
      b200sp_handle h, b200sp_stream s, int stencil, int64_t nx, int64_t ny,       \
      int64_t nz, int64_t row_begin, int64_t num_rows, int64_t col_shift,          \
      const int32_t *row_offsets, int32_t *column_indices, T *values);

B200SP_DECL_GALLERY(float, f32)
B200SP_DECL_GALLERY(double, f64)
#undef B200SP_DECL_GALLERY

/* number of stored entries of the rows [row_begin,row_begin+num_rows) */
int64_t b200sp_poisson_num_entries(int stencil, int64_t nx, int64_t ny, int64_t nz,
                                   int64_t row_begin, int64_t num_rows);
b200sp_status b200sp_poisson_csr_offsets(b200sp_handle h, b200sp_stream s, int stencil,
                                         int64_t nx, int64_t ny, int64_t nz,
                                         int64_t row_begin, int64_t num_rows,
                                         int32_t *row_offsets);

#ifdef __cplusplus
}
#endif
#endif /* B200SP_H */
