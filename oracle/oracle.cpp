// oracle.cpp — CPU restatement of the reference's algorithms on the SpMV / CG
// hot path.  TEST INFRASTRUCTURE ONLY: nothing under oracle/ is linked, imported
// or executed by the product (cusp_autotuned_b200/, include/); only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use
// it, as the checker or the CPU baseline.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).  Parity pins: the restatement is checked (tests/
// test_oracle.py) against
//   * the reference's own host loops compiled unmodified into oracle/_ref/
//     (cusp/system/detail/sequential/multiply/{csr,coo,dia,ell,hyb}_spmv.h),
//   * the golden vectors of testing/{multiply,convert,poisson,format_utils,blas,
//     cg}.cu transcribed into tests/golden/.
// Built with -ffp-contract=off and no -march flags: like the reference's default
// x86-64 host build, no FMA contraction.
//
// Plain loops, int32 indices, column-major pitch layout for ELL/DIA
// (cusp/detail/ell_matrix.inl:35-36, cusp/detail/dia_matrix.inl:34).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

typedef int64_t i64;

// ---------------------------------------------------------------------------
// SpMV  (y = init(y); y += A x), init = 0 (accumulate==0) or identity
// ---------------------------------------------------------------------------
// cusp/system/detail/sequential/multiply/csr_spmv.h:35-74
template <typename T>
static void spmv_csr(i64 rows, const int *Ap, const int *Aj, const T *Ax, const T *x, T *y, int acc) {
  for (i64 i = 0; i < rows; i++) {
    T a = acc ? y[i] : T(0);
    for (int jj = Ap[i]; jj < Ap[i + 1]; jj++) a = a + Ax[jj] * x[Aj[jj]];
    y[i] = a;
  }
}

// cusp/system/detail/sequential/multiply/coo_spmv.h:35-68
template <typename T>
static void spmv_coo(i64 rows, i64 nnz, const int *Ai, const int *Aj, const T *Ax, const T *x, T *y, int acc) {
  if (!acc)
    for (i64 i = 0; i < rows; i++) y[i] = T(0);
  for (i64 n = 0; n < nnz; n++) y[Ai[n]] = y[Ai[n]] + Ax[n] * x[Aj[n]];
}

// cusp/system/detail/sequential/multiply/dia_spmv.h:36-82
template <typename T>
static void spmv_dia(i64 rows, i64 cols, i64 ndiag, i64 pitch, const int *offs, const T *vals, const T *x,
                     T *y, int acc) {
  if (!acc)
    for (i64 i = 0; i < rows; i++) y[i] = T(0);
  for (i64 d = 0; d < ndiag; d++) {
    const i64 k = offs[d];
    const i64 i_start = std::max<i64>(0, -k);
    const i64 j_start = std::max<i64>(0, k);
    if (i_start >= rows || j_start >= cols) continue;  // reference: N wraps to 0 iterations via size_t min
    const i64 N = std::min(rows - i_start, cols - j_start);
    for (i64 n = 0; n < N; n++)
      y[i_start + n] = y[i_start + n] + vals[d * pitch + i_start + n] * x[j_start + n];
  }
}

// cusp/system/detail/sequential/multiply/ell_spmv.h:34-76
template <typename T>
static void spmv_ell(i64 rows, i64 K, i64 pitch, const int *cidx, const T *vals, const T *x, T *y, int acc) {
  if (!acc)
    for (i64 i = 0; i < rows; i++) y[i] = T(0);
  for (i64 n = 0; n < K; n++)
    for (i64 i = 0; i < rows; i++) {
      const int j = cidx[n * pitch + i];
      if (j != -1) y[i] = y[i] + vals[n * pitch + i] * x[j];
    }
}

// ELL-R: ktt_ellr_kernel stops at row_lengths[row] (cusp/system/cuda/ktt/kernels/
// ell_kernel.h:181-213); row_lengths counts the leading non-negative slots
// (cusp/ktt/detail/ellr_matrix.inl:16-52)
template <typename T>
static void spmv_ellr(i64 rows, i64 K, i64 pitch, const int *cidx, const T *vals, const int *len, const T *x,
                      T *y, int acc) {
  for (i64 i = 0; i < rows; i++) {
    T a = acc ? y[i] : T(0);
    for (i64 n = 0; n < K && n < len[i]; n++) {
      const int j = cidx[n * pitch + i];
      if (j != -1) a = a + vals[n * pitch + i] * x[j];
    }
    y[i] = a;
  }
}

// libgomp is not installed in this image, so the row-parallel loops use
// std::thread over contiguous row blocks (static schedule, like the reference's
// `#pragma omp parallel for`).
template <typename F>
static void parallel_rows(i64 rows, int nthreads, F f) {
  if (nthreads <= 1 || rows < 2 * (i64)nthreads) {
    f((i64)0, rows);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) {
    const i64 r0 = rows * t / nthreads, r1 = rows * (t + 1) / nthreads;
    th.emplace_back([=]() { f(r0, r1); });
  }
  for (auto &t : th) t.join();
}

// row-parallel restatements for the multi-core CPU baseline.  CSR follows
// cusp/system/omp/detail/multiply/csr_spmv.h:65-85 (parallel for over rows); ELL/DIA use the same per-row arithmetic order as the sequential
// loops above (slots / diagonals ascending), so results are bit-identical.
template <typename T>
static void spmv_csr_mt(i64 rows, const int *Ap, const int *Aj, const T *Ax, const T *x, T *y, int acc,
                        int nthreads) {
  parallel_rows(rows, nthreads, [=](i64 r0, i64 r1) {
    for (i64 i = r0; i < r1; i++) {
      T a = acc ? y[i] : T(0);
      for (int jj = Ap[i]; jj < Ap[i + 1]; jj++) a = a + Ax[jj] * x[Aj[jj]];
      y[i] = a;
    }
  });
}
template <typename T>
static void spmv_dia_mt(i64 rows, i64 cols, i64 ndiag, i64 pitch, const int *offs, const T *vals, const T *x,
                        T *y, int acc, int nthreads) {
  parallel_rows(rows, nthreads, [=](i64 r0, i64 r1) {
    for (i64 i = r0; i < r1; i++) {
      T a = acc ? y[i] : T(0);
      for (i64 d = 0; d < ndiag; d++) {
        const i64 c = i + offs[d];
        if (c >= 0 && c < cols) a = a + vals[d * pitch + i] * x[c];
      }
      y[i] = a;
    }
  });
}
template <typename T>
static void spmv_ell_mt(i64 rows, i64 K, i64 pitch, const int *cidx, const T *vals, const T *x, T *y,
                        int acc, int nthreads) {
  parallel_rows(rows, nthreads, [=](i64 r0, i64 r1) {
    for (i64 i = r0; i < r1; i++) {
      T a = acc ? y[i] : T(0);
      for (i64 n = 0; n < K; n++) {
        const int j = cidx[n * pitch + i];
        if (j != -1) a = a + vals[n * pitch + i] * x[j];
      }
      y[i] = a;
    }
  });
}

// ---------------------------------------------------------------------------
// BLAS-1  (cusp/system/detail/generic/blas.h:64-96,180-340), sequential order
// (thrust::cpp inner_product / transform_reduce run left to right)
// ---------------------------------------------------------------------------
template <typename T>
static void axpy(i64 n, T alpha, const T *x, T *y) {
  for (i64 i = 0; i < n; i++) y[i] = alpha * x[i] + y[i];
}
template <typename T>
static void axpby(i64 n, T alpha, const T *x, T beta, const T *y, T *z) {
  for (i64 i = 0; i < n; i++) z[i] = alpha * x[i] + beta * y[i];
}
template <typename T>
static T dot(i64 n, const T *x, const T *y) {
  T s = T(0);
  for (i64 i = 0; i < n; i++) s = s + x[i] * y[i];
  return s;
}
template <typename T>
static T nrm2(i64 n, const T *x) {
  T s = T(0);
  for (i64 i = 0; i < n; i++) s = s + x[i] * x[i];
  return std::sqrt(s);
}

// ---------------------------------------------------------------------------
// CG (cusp/krylov/detail/cg.inl:35-107) with identity preconditioner and
// cusp::monitor (cusp/detail/monitor.inl:107-111,178-208).  CSR operator.
// Returns the iteration count; residuals gets one entry per finished() call.
// ---------------------------------------------------------------------------
// dot / nrm2 accumulated in long double (x87 80-bit: 64-bit mantissa) and rounded once: the rounding-free
// reference the sequential sums above and the engine's tree sums are both compared with at 10^8 terms
template <typename T>
static T dot_ld(i64 n, const T *x, const T *y) {
  long double s = 0.0L;
  for (i64 i = 0; i < n; i++) s += (long double)x[i] * (long double)y[i];
  return (T)s;
}

template <typename T, bool COMPENSATED = false>
static i64 cg_csr(i64 n, const int *Ap, const int *Aj, const T *Ax, T *x, const T *b, i64 limit, double rel,
                  double abs_tol, double *residuals, i64 *nres, int *converged) {
  auto dot_ = [](i64 m, const T *u, const T *v) { return COMPENSATED ? dot_ld<T>(m, u, v) : dot<T>(m, u, v); };
  auto nrm2_ = [](i64 m, const T *u) { return COMPENSATED ? (T)std::sqrt((T)dot_ld<T>(m, u, u)) : nrm2<T>(m, u); };
  std::vector<T> y(n), z(n), r(n), p(n);
  const T bnorm = nrm2_(n, b);
  const T tol = (T)abs_tol + (T)rel * bnorm;
  spmv_csr<T>(n, Ap, Aj, Ax, x, y.data(), 0);
  axpby<T>(n, T(1), b, T(-1), y.data(), r.data());
  z = r;
  p = z;
  T rz = dot_(n, r.data(), z.data());
  i64 it = 0, k = 0;
  *converged = 0;
  for (;;) {
    const T rn = nrm2_(n, r.data());
    residuals[k++] = (double)rn;
    if (rn <= tol) {
      *converged = 1;
      break;
    }
    if (it >= limit) break;
    spmv_csr<T>(n, Ap, Aj, Ax, p.data(), y.data(), 0);
    const T alpha = rz / dot_(n, y.data(), p.data());
    axpy<T>(n, alpha, p.data(), x);
    axpy<T>(n, -alpha, y.data(), r.data());
    z = r;
    const T rz_old = rz;
    rz = dot_(n, r.data(), z.data());
    const T beta = rz / rz_old;
    axpby<T>(n, T(1), z.data(), beta, p.data(), p.data());
    ++it;
  }
  *nres = k;
  return it;
}

// ---------------------------------------------------------------------------
// The other Krylov solvers with the identity or a diagonal (Jacobi) preconditioner (dinv == nullptr: identity;
// cusp::precond::diagonal applies z = diagonal_reciprocals .* r through blas::xmy, precond/detail/diagonal.inl:52-56),
// CSR operator, sequential sums, the monitor called exactly where the reference calls it.
// ---------------------------------------------------------------------------
template <typename T>
struct OracleMonitor {  // cusp/detail/monitor.inl:178-208
  T tol;
  i64 limit, iter = 0, k = 0;
  double *residuals;
  int converged = 0;
  bool finished(i64 n, const T *v) {
    const T rn = nrm2<T>(n, v);
    residuals[k++] = (double)rn;
    if (rn <= tol) {
      converged = 1;
      return true;
    }
    return iter >= limit;
  }
};
template <typename T>
static void apply_precond(i64 n, const T *dinv, const T *r, T *z) {
  for (i64 i = 0; i < n; i++) z[i] = dinv ? dinv[i] * r[i] : r[i];
}
template <typename T>
static void axpbypcz(i64 n, T a, const T *x, T b, const T *y, T c, const T *z, T *out) {
  for (i64 i = 0; i < n; i++) out[i] = a * x[i] + b * y[i] + c * z[i];
}

// cusp/krylov/detail/cg.inl:35-107 with a preconditioner
template <typename T>
static i64 pcg_csr(i64 n, const int *Ap, const int *Aj, const T *Ax, const T *dinv, T *x, const T *b, i64 limit, double rel,
                   double abs_tol, double *residuals, i64 *nres, int *converged) {
  std::vector<T> y(n), z(n), r(n), p(n);
  OracleMonitor<T> mon{(T)abs_tol + (T)rel * nrm2<T>(n, b), limit, 0, 0, residuals};
  spmv_csr<T>(n, Ap, Aj, Ax, x, y.data(), 0);
  axpby<T>(n, T(1), b, T(-1), y.data(), r.data());
  apply_precond<T>(n, dinv, r.data(), z.data());
  p = z;
  T rz = dot<T>(n, r.data(), z.data());
  while (!mon.finished(n, r.data())) {
    spmv_csr<T>(n, Ap, Aj, Ax, p.data(), y.data(), 0);
    const T alpha = rz / dot<T>(n, y.data(), p.data());
    axpy<T>(n, alpha, p.data(), x);
    axpy<T>(n, -alpha, y.data(), r.data());
    apply_precond<T>(n, dinv, r.data(), z.data());
    const T rz_old = rz;
    rz = dot<T>(n, r.data(), z.data());
    const T beta = rz / rz_old;
    axpby<T>(n, T(1), z.data(), beta, p.data(), p.data());
    ++mon.iter;
  }
  *nres = mon.k;
  *converged = mon.converged;
  return mon.iter;
}

// cusp/krylov/detail/bicgstab.inl:35-123
template <typename T>
static i64 bicgstab_csr(i64 n, const int *Ap, const int *Aj, const T *Ax, const T *dinv, T *x, const T *b, i64 limit,
                        double rel, double abs_tol, double *residuals, i64 *nres, int *converged) {
  std::vector<T> p(n), r(n), r_star(n), s(n), Mp(n), AMp(n), Ms(n), AMs(n);
  OracleMonitor<T> mon{(T)abs_tol + (T)rel * nrm2<T>(n, b), limit, 0, 0, residuals};
  spmv_csr<T>(n, Ap, Aj, Ax, x, r.data(), 0);
  axpby<T>(n, T(1), b, T(-1), r.data(), r.data());
  p = r;
  r_star = r;
  T rho_old = dot<T>(n, r_star.data(), r.data());
  while (!mon.finished(n, r.data())) {
    apply_precond<T>(n, dinv, p.data(), Mp.data());
    spmv_csr<T>(n, Ap, Aj, Ax, Mp.data(), AMp.data(), 0);
    const T alpha = rho_old / dot<T>(n, r_star.data(), AMp.data());
    axpby<T>(n, T(1), r.data(), T(-alpha), AMp.data(), s.data());
    if (mon.finished(n, s.data())) {
      axpby<T>(n, T(1), x, T(alpha), Mp.data(), x);
      break;
    }
    apply_precond<T>(n, dinv, s.data(), Ms.data());
    spmv_csr<T>(n, Ap, Aj, Ax, Ms.data(), AMs.data(), 0);
    const T omega = dot<T>(n, AMs.data(), s.data()) / dot<T>(n, AMs.data(), AMs.data());
    axpbypcz<T>(n, T(1), x, alpha, Mp.data(), omega, Ms.data(), x);
    axpby<T>(n, T(1), s.data(), -omega, AMs.data(), r.data());
    const T rho_new = dot<T>(n, r_star.data(), r.data());
    const T beta = (rho_new / rho_old) * (alpha / omega);
    rho_old = rho_new;
    axpbypcz<T>(n, T(1), r.data(), beta, p.data(), -beta * omega, AMp.data(), p.data());
    ++mon.iter;
  }
  *nres = mon.k;
  *converged = mon.converged;
  return mon.iter;
}

// cusp/krylov/detail/cr.inl:39-128 (r recomputed from b - A x every 8 iterations)
template <typename T>
static i64 cr_csr(i64 n, const int *Ap, const int *Aj, const T *Ax, const T *dinv, T *x, const T *b, i64 limit, double rel,
                  double abs_tol, double *residuals, i64 *nres, int *converged) {
  std::vector<T> y(n), z(n), r(n), p(n), Az(n), Axv(n);
  OracleMonitor<T> mon{(T)abs_tol + (T)rel * nrm2<T>(n, b), limit, 0, 0, residuals};
  spmv_csr<T>(n, Ap, Aj, Ax, x, Axv.data(), 0);
  axpby<T>(n, T(1), b, T(-1), Axv.data(), r.data());
  apply_precond<T>(n, dinv, r.data(), z.data());
  p = z;
  spmv_csr<T>(n, Ap, Aj, Ax, p.data(), y.data(), 0);
  spmv_csr<T>(n, Ap, Aj, Ax, z.data(), Az.data(), 0);
  T rz = dot<T>(n, r.data(), Az.data());
  while (!mon.finished(n, r.data())) {
    const T alpha = rz / dot<T>(n, y.data(), y.data());
    axpy<T>(n, alpha, p.data(), x);
    if ((mon.iter % 8) && (mon.iter > 0)) {
      axpy<T>(n, -alpha, y.data(), r.data());
    } else {
      spmv_csr<T>(n, Ap, Aj, Ax, x, Axv.data(), 0);
      axpby<T>(n, T(1), b, T(-1), Axv.data(), r.data());
    }
    apply_precond<T>(n, dinv, r.data(), z.data());
    spmv_csr<T>(n, Ap, Aj, Ax, z.data(), Az.data(), 0);
    const T rz_old = rz;
    rz = dot<T>(n, r.data(), Az.data());
    const T beta = rz / rz_old;
    axpby<T>(n, T(1), z.data(), beta, p.data(), p.data());
    axpby<T>(n, T(1), Az.data(), beta, y.data(), y.data());
    ++mon.iter;
  }
  *nres = mon.k;
  *converged = mon.converged;
  return mon.iter;
}

// ---------------------------------------------------------------------------
// gallery: generate_matrix_from_stencil -> DIA
// (cusp/gallery/detail/stencil.inl:33-63 inside_grid, :114-134 fill, :143-188)
// stencil points: npts x ndim integer offsets + value; grid: ndim extents;
// dimension 0 is the fastest varying.  pitch = num_rows (4-arg resize, :174).
// ---------------------------------------------------------------------------
template <typename T>
static i64 stencil_dia(int ndim, const i64 *grid, int npts, const int *pts, const T *pvals, int *offsets,
                       T *values /* npts*rows */) {
  i64 rows = 1;
  std::vector<i64> strides(ndim);
  for (int j = 0; j < ndim; j++) {
    strides[j] = rows;
    rows *= grid[j];
  }
  i64 nnz = 0;
  for (int p = 0; p < npts; p++) {
    i64 off = 0;
    for (int j = 0; j < ndim; j++) off += strides[j] * pts[p * ndim + j];
    offsets[p] = (int)off;
    for (i64 idx = 0; idx < rows; idx++) {
      i64 rem = idx;
      bool inside = true;
      for (int j = 0; j < ndim; j++) {
        const i64 xj = rem % grid[j] + pts[p * ndim + j];
        if (xj < 0 || xj >= grid[j]) {
          inside = false;
          break;
        }
        rem /= grid[j];
      }
      const T v = inside ? pvals[p] : T(0);
      values[(i64)p * rows + idx] = v;
      if (v != T(0)) nnz++;
    }
  }
  return nnz;  // num_entries = size - count(0), stencil.inl:188
}

// ---------------------------------------------------------------------------
// conversions
// ---------------------------------------------------------------------------
// DIA -> COO (row-major scan of the K x rows logical array, drop value == 0)
// cusp/system/detail/generic/conversions/dia_to_other.h:61-108; ->CSR :110-161
template <typename T>
static i64 dia_to_coo(i64 rows, i64 cols, i64 ndiag, i64 pitch, const int *offs, const T *vals, int *Ai,
                      int *Aj, T *Ax) {
  (void)cols;
  i64 n = 0;
  for (i64 i = 0; i < rows; i++)
    for (i64 d = 0; d < ndiag; d++) {
      const T v = vals[d * pitch + i];
      if (v != T(0)) {
        if (Ai) {
          Ai[n] = (int)i;
          Aj[n] = (int)(i + offs[d]);
          Ax[n] = v;
        }
        n++;
      }
    }
  return n;
}

// ELL -> COO / CSR (cusp/system/detail/generic/conversions/ell_to_other.h:55-143): row-major scan of the
// [rows x K] logical array (thrust::copy_if over a row_major -> column_major permutation), keep value != 0
// (the stencil is the VALUE, not the column: explicit zeros are dropped, like DIA).  Returns the count.
template <typename T>
static i64 ell_to_coo(i64 rows, i64 K, i64 pitch, const int *cidx, const T *vals, int *Ai, int *Aj, T *Ax) {
  i64 n = 0;
  for (i64 i = 0; i < rows; i++)
    for (i64 k = 0; k < K; k++) {
      const T v = vals[k * pitch + i];
      if (v != T(0)) {
        if (Ai) {
          Ai[n] = (int)i;
          Aj[n] = cidx[k * pitch + i];
          Ax[n] = v;
        }
        n++;
      }
    }
  return n;
}

// HYB -> COO (hyb_to_other.h:45-56 -> coo view of a hyb matrix, cusp/detail/coo_matrix.inl:269-341): the ELL part
// in row-major order and the COO part are merged by (row, column) — thrust::merge_by_key, ties: ELL first — and the
// slots whose COLUMN is the invalid index are removed (explicit zeros stay, unlike ELL -> COO).  Restated as a
// sequential two-pointer merge of the two row-sorted sequences.
template <typename T>
static i64 hyb_to_coo(i64 rows, i64 K, i64 pitch, const int *ecidx, const T *evals, i64 cnnz, const int *ci,
                      const int *cj, const T *cv, int *Ai, int *Aj, T *Ax) {
  i64 n = 0, b = 0;  // b: next COO entry
  auto emit = [&](int r, int c, T v) {
    if (c != -1) {
      if (Ai) {
        Ai[n] = r;
        Aj[n] = c;
        Ax[n] = v;
      }
      n++;
    }
  };
  for (i64 i = 0; i < rows; i++)
    for (i64 k = 0; k < K; k++) {
      const int c = ecidx[k * pitch + i];
      // COO entries strictly smaller than this ELL entry come first
      while (b < cnnz && (ci[b] < (int)i || (ci[b] == (int)i && cj[b] < c))) {
        emit(ci[b], cj[b], cv[b]);
        b++;
      }
      emit((int)i, c, evals[k * pitch + i]);
    }
  for (; b < cnnz; b++) emit(ci[b], cj[b], cv[b]);
  return n;
}

// cpu_compute_row_starts (cusp/system/cuda/ktt/csr_multiply.h:38-61): the preprocessing of the balanced CSR kernel.
// out[w] = the row whose entry range contains w * chunk, chunk = ceil(nnz / workers); workers beyond the matrix get 0.
static void compute_row_starts(i64 rows, i64 nnz, const int *Ap, i64 workers, int *out) {
  i64 count = 0, w = 0;
  const i64 chunk = workers > 0 ? (nnz + workers - 1) / workers : 0;
  for (i64 i = 0; i < rows; i++) {
    const i64 next = Ap[i + 1];
    while (count <= w * chunk && w * chunk < next && w < workers) {
      out[w] = (int)i;
      ++w;
    }
    count = next;
  }
  for (; w < workers; ++w) out[w] = 0;
}

// indices_to_offsets / offsets_to_indices (cusp/format_utils.h, testing/format_utils.cu:13-75)
static void indices_to_offsets(i64 nnz, const int *idx, i64 rows, int *offs) {
  // offsets[i] = number of indices < i  (lower_bound over sorted indices)
  i64 k = 0;
  for (i64 i = 0; i <= rows; i++) {
    while (k < nnz && idx[k] < i) k++;
    offs[i] = (int)k;
  }
}
static void offsets_to_indices(i64 rows, const int *offs, int *idx) {
  for (i64 i = 0; i < rows; i++)
    for (int j = offs[i]; j < offs[i + 1]; j++) idx[j] = (int)i;
}

// DIA -> ELL as the fork does it (dia_to_other.h:163-251): K = #diagonals,
// pitch = DIA pitch, col = -1 where value == 0, rows stably left-packed.
template <typename T>
static void dia_to_ell(i64 rows, i64 ndiag, i64 pitch, const int *offs, const T *vals, int *cidx, T *evals) {
  for (i64 i = 0; i < pitch; i++)
    for (i64 d = 0; d < ndiag; d++) {
      cidx[d * pitch + i] = -1;
      evals[d * pitch + i] = T(0);
    }
  for (i64 i = 0; i < rows; i++) {
    i64 k = 0;
    for (i64 d = 0; d < ndiag; d++) {
      const T v = vals[d * pitch + i];
      if (v != T(0)) {
        cidx[k * pitch + i] = (int)(i + offs[d]);
        evals[k * pitch + i] = v;
        k++;
      }
    }
  }
}

// CSR -> ELL (cusp/system/detail/generic/conversions/csr_to_other.h:155-227):
// k-th entry of row i -> slot k*pitch+i, pad col=-1/val=0, pitch=round_up(rows,align)
template <typename T>
static void csr_to_ell(i64 rows, const int *Ap, const int *Aj, const T *Ax, i64 K, i64 pitch, int *cidx,
                       T *vals) {
  for (i64 s = 0; s < K * pitch; s++) {
    cidx[s] = -1;
    vals[s] = T(0);
  }
  for (i64 i = 0; i < rows; i++)
    for (int jj = Ap[i]; jj < Ap[i + 1]; jj++) {
      const i64 k = jj - Ap[i];
      if (k < K) {
        cidx[k * pitch + i] = Aj[jj];
        vals[k * pitch + i] = Ax[jj];
      }
    }
}

static i64 max_entries_per_row(i64 rows, const int *Ap) {
  i64 m = 0;
  for (i64 i = 0; i < rows; i++) m = std::max<i64>(m, Ap[i + 1] - Ap[i]);
  return m;
}

// compute_optimal_entries_per_row (cusp/system/detail/generic/format_utils.inl:281-321)
// + speed_threshold_functor (cusp/detail/functional.inl:114-132)
static i64 optimal_entries_per_row(i64 rows, const int *Ap, float relative_speed, i64 breakeven) {
  const i64 maxc = max_entries_per_row(rows, Ap);
  std::vector<i64> cum(maxc + 1, 0);  // cum[k] = #rows with length <= k  (upper_bound)
  for (i64 i = 0; i < rows; i++) cum[Ap[i + 1] - Ap[i]]++;
  for (i64 k = 1; k <= maxc; k++) cum[k] += cum[k - 1];
  for (i64 k = 0; k < maxc; k++) {
    const size_t r = (size_t)cum[k];
    // relative_speed * (num_rows-rows) < num_rows || (num_rows-rows) < breakeven
    if (relative_speed * (float)((size_t)rows - r) < (float)(size_t)rows || ((size_t)rows - r) < (size_t)breakeven)
      return k;
  }
  return maxc;
}

// CSR -> HYB (csr_to_other.h:229-306): first K entries of each row to ELL, the
// rest to COO in CSR order.
template <typename T>
static i64 csr_to_hyb(i64 rows, const int *Ap, const int *Aj, const T *Ax, i64 K, i64 pitch, int *ecidx,
                      T *evals, int *ci, int *cj, T *cv) {
  for (i64 s = 0; s < K * pitch; s++) {
    ecidx[s] = -1;
    evals[s] = T(0);
  }
  i64 n = 0;
  for (i64 i = 0; i < rows; i++)
    for (int jj = Ap[i]; jj < Ap[i + 1]; jj++) {
      const i64 k = jj - Ap[i];
      if (k < K) {
        ecidx[k * pitch + i] = Aj[jj];
        evals[k * pitch + i] = Ax[jj];
      } else {
        if (ci) {
          ci[n] = (int)i;
          cj[n] = Aj[jj];
          cv[n] = Ax[jj];
        }
        n++;
      }
    }
  return n;
}

// CSR -> DIA (csr_to_other.h:73-153): occupied diagonals ascending, pitch given,
// zero fill, values scattered.  Returns #diagonals (offsets may be NULL to count).
template <typename T>
static i64 csr_to_dia(i64 rows, i64 cols, const int *Ap, const int *Aj, const T *Ax, i64 pitch, int *offs,
                      T *vals) {
  std::vector<int> occ(rows + cols, 0);
  for (i64 i = 0; i < rows; i++)
    for (int jj = Ap[i]; jj < Ap[i + 1]; jj++) occ[Aj[jj] - i + rows] = 1;
  std::vector<int> map(rows + cols, -1);
  i64 nd = 0;
  for (i64 k = 0; k < rows + cols; k++)
    if (occ[k]) {
      if (offs) offs[nd] = (int)(k - rows);
      map[k] = (int)nd++;
    }
  if (!vals) return nd;
  for (i64 s = 0; s < nd * pitch; s++) vals[s] = T(0);
  for (i64 i = 0; i < rows; i++)
    for (int jj = Ap[i]; jj < Ap[i + 1]; jj++) vals[(i64)map[Aj[jj] - i + rows] * pitch + i] = Ax[jj];
  return nd;
}

// cusp::gallery::random (cusp/gallery/detail/random.inl:33-63): srand(m^n^samples),
// glibc rand()%m / rand()%n, sort by (row,col), unique, values 1.
static i64 gallery_random(i64 m, i64 n, i64 samples, int *Ai, int *Aj) {
  std::vector<std::pair<int, int>> e((size_t)samples);
  srand((unsigned)(m ^ n ^ samples));
  for (i64 k = 0; k < samples; k++) {
    const int r = rand() % m;
    const int c = rand() % n;
    e[(size_t)k] = std::make_pair(r, c);
  }
  std::sort(e.begin(), e.end());
  e.erase(std::unique(e.begin(), e.end()), e.end());
  for (size_t k = 0; k < e.size(); k++) {
    Ai[k] = e[k].first;
    Aj[k] = e[k].second;
  }
  return (i64)e.size();
}

// cusp::ktt::make_diagonal_matrix (cusp/ktt/matrix_generation.h:14-61): ones on
// the given diagonals, pitch = rows
static i64 make_diagonal(i64 rows, i64 cols, i64 nd, const int *offs, float *vals) {
  i64 nnz = 0;
  for (i64 s = 0; s < nd * rows; s++) vals[s] = 0.f;
  for (i64 d = 0; d < nd; d++) {
    const i64 sr = offs[d] < 0 ? -offs[d] : 0, sc = offs[d] < 0 ? 0 : offs[d];
    if (sr >= rows || sc >= cols) return -1;
    const i64 er = sr + std::min(rows - sr, cols - sc);
    for (i64 r = sr; r < er; r++) {
      vals[d * rows + r] = 1.f;
      nnz++;
    }
  }
  return nnz;
}

// ---------------------------------------------------------------------------
extern "C" {
#define DEF(T, sfx)                                                                                          \
  void oracle_spmv_csr_##sfx(i64 rows, const int *Ap, const int *Aj, const T *Ax, const T *x, T *y, int acc) { \
    spmv_csr<T>(rows, Ap, Aj, Ax, x, y, acc);                                                                \
  }                                                                                                          \
  void oracle_spmv_csr_mt_##sfx(i64 rows, const int *Ap, const int *Aj, const T *Ax, const T *x, T *y,       \
                                int acc, int nthreads) {                                                     \
    spmv_csr_mt<T>(rows, Ap, Aj, Ax, x, y, acc, nthreads);                                                   \
  }                                                                                                          \
  void oracle_spmv_coo_##sfx(i64 rows, i64 nnz, const int *Ai, const int *Aj, const T *Ax, const T *x, T *y, \
                             int acc) {                                                                      \
    spmv_coo<T>(rows, nnz, Ai, Aj, Ax, x, y, acc);                                                           \
  }                                                                                                          \
  void oracle_spmv_dia_##sfx(i64 rows, i64 cols, i64 nd, i64 pitch, const int *offs, const T *vals,          \
                             const T *x, T *y, int acc) {                                                    \
    spmv_dia<T>(rows, cols, nd, pitch, offs, vals, x, y, acc);                                               \
  }                                                                                                          \
  void oracle_spmv_dia_mt_##sfx(i64 rows, i64 cols, i64 nd, i64 pitch, const int *offs, const T *vals,       \
                                const T *x, T *y, int acc, int nthreads) {                                   \
    spmv_dia_mt<T>(rows, cols, nd, pitch, offs, vals, x, y, acc, nthreads);                                  \
  }                                                                                                          \
  void oracle_spmv_ell_##sfx(i64 rows, i64 K, i64 pitch, const int *cidx, const T *vals, const T *x, T *y,   \
                             int acc) {                                                                      \
    spmv_ell<T>(rows, K, pitch, cidx, vals, x, y, acc);                                                      \
  }                                                                                                          \
  void oracle_spmv_ell_mt_##sfx(i64 rows, i64 K, i64 pitch, const int *cidx, const T *vals, const T *x,      \
                                T *y, int acc, int nthreads) {                                               \
    spmv_ell_mt<T>(rows, K, pitch, cidx, vals, x, y, acc, nthreads);                                         \
  }                                                                                                          \
  void oracle_spmv_ellr_##sfx(i64 rows, i64 K, i64 pitch, const int *cidx, const T *vals, const int *len,    \
                              const T *x, T *y, int acc) {                                                   \
    spmv_ellr<T>(rows, K, pitch, cidx, vals, len, x, y, acc);                                                \
  }                                                                                                          \
  /* HYB: ELL pass then COO pass with identity, sequential/multiply/hyb_spmv.h:35-57 */                      \
  void oracle_spmv_hyb_##sfx(i64 rows, i64 K, i64 pitch, const int *ecidx, const T *evals, i64 cnnz,         \
                             const int *ci, const int *cj, const T *cv, const T *x, T *y, int acc) {         \
    spmv_ell<T>(rows, K, pitch, ecidx, evals, x, y, acc);                                                    \
    spmv_coo<T>(rows, cnnz, ci, cj, cv, x, y, 1);                                                            \
  }                                                                                                          \
  void oracle_axpy_##sfx(i64 n, T a, const T *x, T *y) { axpy<T>(n, a, x, y); }                              \
  void oracle_axpby_##sfx(i64 n, T a, const T *x, T b, const T *y, T *z) { axpby<T>(n, a, x, b, y, z); }     \
  T oracle_dot_##sfx(i64 n, const T *x, const T *y) { return dot<T>(n, x, y); }                              \
  T oracle_nrm2_##sfx(i64 n, const T *x) { return nrm2<T>(n, x); }                                           \
  i64 oracle_cg_csr_##sfx(i64 n, const int *Ap, const int *Aj, const T *Ax, T *x, const T *b, i64 limit,     \
                          double rel, double abs_tol, double *residuals, i64 *nres, int *converged) {        \
    return cg_csr<T>(n, Ap, Aj, Ax, x, b, limit, rel, abs_tol, residuals, nres, converged);                  \
  }                                                                                                          \
  i64 oracle_krylov_csr_##sfx(int solver, i64 n, const int *Ap, const int *Aj, const T *Ax, const T *dinv, T *x,   \
                              const T *b, i64 limit, double rel, double abs_tol, double *residuals, i64 *nres,  \
                              int *converged) {                                                              \
    if (solver == 0) return pcg_csr<T>(n, Ap, Aj, Ax, dinv, x, b, limit, rel, abs_tol, residuals, nres, converged);      \
    if (solver == 1) return bicgstab_csr<T>(n, Ap, Aj, Ax, dinv, x, b, limit, rel, abs_tol, residuals, nres, converged); \
    return cr_csr<T>(n, Ap, Aj, Ax, dinv, x, b, limit, rel, abs_tol, residuals, nres, converged);                        \
  }                                                                                                          \
  i64 oracle_cg_csr_compensated_##sfx(i64 n, const int *Ap, const int *Aj, const T *Ax, T *x, const T *b,    \
                                      i64 limit, double rel, double abs_tol, double *residuals, i64 *nres,  \
                                      int *converged) {                                                      \
    return cg_csr<T, true>(n, Ap, Aj, Ax, x, b, limit, rel, abs_tol, residuals, nres, converged);            \
  }                                                                                                          \
  i64 oracle_stencil_dia_##sfx(int ndim, const i64 *grid, int npts, const int *pts, const T *pvals,          \
                               int *offsets, T *values) {                                                    \
    return stencil_dia<T>(ndim, grid, npts, pts, pvals, offsets, values);                                    \
  }                                                                                                          \
  i64 oracle_dia_to_coo_##sfx(i64 rows, i64 cols, i64 nd, i64 pitch, const int *offs, const T *vals,         \
                              int *Ai, int *Aj, T *Ax) {                                                     \
    return dia_to_coo<T>(rows, cols, nd, pitch, offs, vals, Ai, Aj, Ax);                                     \
  }                                                                                                          \
  i64 oracle_ell_to_coo_##sfx(i64 rows, i64 K, i64 pitch, const int *cidx, const T *vals, int *Ai, int *Aj,  \
                              T *Ax) {                                                                       \
    return ell_to_coo<T>(rows, K, pitch, cidx, vals, Ai, Aj, Ax);                                            \
  }                                                                                                          \
  i64 oracle_hyb_to_coo_##sfx(i64 rows, i64 K, i64 pitch, const int *ecidx, const T *evals, i64 cnnz,        \
                              const int *ci, const int *cj, const T *cv, int *Ai, int *Aj, T *Ax) {          \
    return hyb_to_coo<T>(rows, K, pitch, ecidx, evals, cnnz, ci, cj, cv, Ai, Aj, Ax);                        \
  }                                                                                                          \
  void oracle_dia_to_ell_##sfx(i64 rows, i64 nd, i64 pitch, const int *offs, const T *vals, int *cidx,       \
                               T *evals) {                                                                   \
    dia_to_ell<T>(rows, nd, pitch, offs, vals, cidx, evals);                                                 \
  }                                                                                                          \
  void oracle_csr_to_ell_##sfx(i64 rows, const int *Ap, const int *Aj, const T *Ax, i64 K, i64 pitch,        \
                               int *cidx, T *vals) {                                                         \
    csr_to_ell<T>(rows, Ap, Aj, Ax, K, pitch, cidx, vals);                                                   \
  }                                                                                                          \
  i64 oracle_csr_to_hyb_##sfx(i64 rows, const int *Ap, const int *Aj, const T *Ax, i64 K, i64 pitch,         \
                              int *ecidx, T *evals, int *ci, int *cj, T *cv) {                               \
    return csr_to_hyb<T>(rows, Ap, Aj, Ax, K, pitch, ecidx, evals, ci, cj, cv);                              \
  }                                                                                                          \
  i64 oracle_csr_to_dia_##sfx(i64 rows, i64 cols, const int *Ap, const int *Aj, const T *Ax, i64 pitch,      \
                              int *offs, T *vals) {                                                          \
    return csr_to_dia<T>(rows, cols, Ap, Aj, Ax, pitch, offs, vals);                                         \
  }
DEF(float, f32)
DEF(double, f64)
#undef DEF

void oracle_compute_row_starts(i64 rows, i64 nnz, const int *Ap, i64 workers, int *out) {
  compute_row_starts(rows, nnz, Ap, workers, out);
}
void oracle_indices_to_offsets(i64 nnz, const int *idx, i64 rows, int *offs) {
  indices_to_offsets(nnz, idx, rows, offs);
}
void oracle_offsets_to_indices(i64 rows, const int *offs, int *idx) { offsets_to_indices(rows, offs, idx); }
i64 oracle_max_entries_per_row(i64 rows, const int *Ap) { return max_entries_per_row(rows, Ap); }
i64 oracle_optimal_entries_per_row(i64 rows, const int *Ap, float rel, i64 breakeven) {
  return optimal_entries_per_row(rows, Ap, rel, breakeven);
}
i64 oracle_gallery_random(i64 m, i64 n, i64 samples, int *Ai, int *Aj) {
  return gallery_random(m, n, samples, Ai, Aj);
}
i64 oracle_make_diagonal(i64 rows, i64 cols, i64 nd, const int *offs, float *vals) {
  return make_diagonal(rows, cols, nd, offs, vals);
}
int oracle_num_threads(void) { return (int)std::thread::hardware_concurrency(); }
}
