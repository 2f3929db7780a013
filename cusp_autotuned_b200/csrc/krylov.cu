// krylov.cu — fused Krylov solvers beside CG: Jacobi-preconditioned CG, BiCGStab and CR.
//
// Replaces, for device operators with the identity or a diagonal (Jacobi) preconditioner,
//   cusp::krylov::cg       with M = cusp::precond::diagonal  (cusp/krylov/detail/cg.inl:35-107, precond/diagonal.h)
//   cusp::krylov::bicgstab                                    (cusp/krylov/detail/bicgstab.inl:35-123)
//   cusp::krylov::cr                                          (cusp/krylov/detail/cr.inl:39-128)
// + cusp::monitor (cusp/detail/monitor.inl:178-208).  The reference runs every BLAS-1 step as its own pass and
// synchronises with the host for every dot product and every monitor.finished() (BiCGStab: 6 per iteration).  Here the
// scalars (alpha, beta, omega, <.,.>, the monitor) live in device memory; each iteration is the SpMVs (dot products
// fused into their epilogues where the kernel has one) plus 2-3 fused vector kernels whose last CTA performs the
// scalar step; the host polls a flag every `check_interval` iterations and kernels launched after the monitor said
// "finished" return immediately, so x holds the iterate at which the reference would have stopped.  Every element
// update is the reference's expression in the reference's order (library built with -fmad=false): same iterate
// sequence, only the summation order of the dot products differs.
//
// Row-block partitioned form (halo != nullptr): operands of A live in windows [halo_lo | local | halo_hi]; halo
// planes are exchanged before each product (comm_halo_exchange_auto: NVLink peer memory or NCCL) and the local sums
// are all-reduced with NCCL, after which a one-thread kernel performs the scalar step.
#include <vector>

#include "comm.h"
#include "common.cuh"

namespace b200sp {

template <typename T>
b200sp_status spmv_any(b200sp_handle h, cudaStream_t st, const b200sp_matrix *A, const T *x, T *y, int accumulate,
                       const b200sp_cfg *cfg, const T *dotv, T *dot_result);  // cg.cu
template <typename T, int MODE>
b200sp_status reduce(b200sp_handle, cudaStream_t, i64, const T *, const T *, T *, T *);  // blas1.cu

template <typename T>
struct KrState {
  T rho;      // CG: <r,z>   BiCGStab: <r*, r>   CR: <r, A z>
  T d1, d2, d3;  // dot products written by SpMV epilogues / reductions (CG: <Ap,p>; BiCGStab: <r*,AMp>, <AMs,s>, <AMs,AMs>; CR: d3 = <y,y>)
  T alpha, omega, beta;
  T rr;       // ||r||^2 waiting for the monitor (CR)
  T acc0, acc1;  // local sums of the partitioned form, all-reduced before the scalar step
  T tol, bnorm, rnorm;
  int iter, limit, done, converged, nres, pad;
};

// monitor.finished(v): record ||v||, decide (cusp/detail/monitor.inl:178-208)
template <typename T>
__device__ __forceinline__ void kr_finished(KrState<T> *S, T sumsq, double *residuals) {
  const T rn = (T)sqrt((double)sumsq);
  S->rnorm = rn;
  residuals[S->nres++] = (double)rn;
  if (rn <= S->tol) {
    S->converged = 1;
    S->done = 1;
  } else if (S->iter >= S->limit) {
    S->done = 1;
  }
}

constexpr int KR_BLOCK = 256;
constexpr int KR_UNROLL = 4;

// two-value version of grid_reduce_finish (common.cuh): partials[2 * gridDim.x]
template <typename T, typename Fin>
__device__ __forceinline__ void kr_grid_finish(T v0, T v1, T *partials, unsigned int *ticket, T *smem, Fin fin) {
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = v0;
    partials[gridDim.x + blockIdx.x] = v1;
    __threadfence();
    is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    T a0 = 0, a1 = 0;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += KR_BLOCK) {
      a0 += *((volatile T *)(partials + i));
      a1 += *((volatile T *)(partials + gridDim.x + i));
    }
    a0 = block_sum<KR_BLOCK>(a0, smem);
    a1 = block_sum<KR_BLOCK>(a1, smem);
    if (threadIdx.x == 0) {
      fin(a0, a1);
      *ticket = 0;
      __threadfence();
    }
  }
}

// One fused vector step.  F provides: GATED (skip once the monitor has finished), NACC (0..2 sums), Regs,
// prepare(S) (scalars into registers), load(i, Regs&), apply(i, Regs&, a0, a1), finish(S, t0, t1, residuals).
template <typename T, typename F, bool DIST>
__global__ void __launch_bounds__(KR_BLOCK) kr_kernel(i64 n, F f, KrState<T> *S, T *partials, unsigned int *ticket,
                                                      double *residuals) {
  __shared__ T s_red[32];
  if (F::GATED && S->done) return;
  f.prepare(S);
  T a0 = T(0), a1 = T(0);
  const i64 stride = (i64)gridDim.x * KR_BLOCK;
  i64 i = (i64)blockIdx.x * KR_BLOCK + threadIdx.x;
  for (; i + (KR_UNROLL - 1) * stride < n; i += KR_UNROLL * stride) {
    typename F::Regs rg[KR_UNROLL];
#pragma unroll
    for (int u = 0; u < KR_UNROLL; ++u) f.load(i + u * stride, rg[u]);
#pragma unroll
    for (int u = 0; u < KR_UNROLL; ++u) f.apply(i + u * stride, rg[u], a0, a1);
  }
  for (; i < n; i += stride) {
    typename F::Regs rg;
    f.load(i, rg);
    f.apply(i, rg, a0, a1);
  }
  if (F::NACC == 0) return;
  const T b0 = block_sum<KR_BLOCK>(a0, s_red);
  const T b1 = F::NACC > 1 ? block_sum<KR_BLOCK>(a1, s_red) : T(0);
  kr_grid_finish<T>(b0, b1, partials, ticket, s_red, [&](T t0, T t1) {
    if (DIST) {
      S->acc0 = t0;
      S->acc1 = t1;
    } else {
      f.finish(S, t0, t1, residuals);
    }
  });
}
template <typename T, typename F>
__global__ void kr_scalar_kernel(F f, KrState<T> *S, double *residuals) {
  if (F::GATED && S->done) return;
  f.prepare(S);
  f.finish(S, S->acc0, S->acc1, residuals);
}

template <typename T>
__global__ void kr_setup_kernel(KrState<T> *S, const T *bnorm, double rel, double abs_tol, int limit) {
  memset(S, 0, sizeof(KrState<T>));
  S->bnorm = *bnorm;
  S->tol = (T)abs_tol + (T)rel * (*bnorm);  // monitor::tolerance() in Real
  S->limit = limit;
  S->d1 = S->d3 = S->rho = T(1);
}

// ============================== functors ===========================================
// z = M r with M = diag^-1: cusp::blas::xmy(diagonal_reciprocals, r, z) -> dinv[i] * r[i] (precond/detail/diagonal.inl:52-56)
template <typename T>
__device__ __forceinline__ T apply_m(const T *dinv, i64 i, T v) {
  return dinv ? dinv[i] * v : v;
}

// ---- preconditioned CG ----
template <typename T>
struct PcgInit {  // r = b - A x0; z = M r; p = z; rho = <r,z>; monitor(r)
  static constexpr bool GATED = false;
  static constexpr int NACC = 2;
  const T *b, *y, *dinv;
  T *r, *p;
  struct Regs { T b, y; };
  __device__ void prepare(const KrState<T> *) {}
  __device__ void load(i64 i, Regs &g) const { g.b = b[i]; g.y = y[i]; }
  __device__ void apply(i64 i, Regs &g, T &a0, T &a1) const {
    const T ri = T(1) * g.b + T(-1) * g.y;
    const T zi = apply_m(dinv, i, ri);
    r[i] = ri;
    p[i] = zi;
    a0 = a0 + ri * zi;
    a1 = a1 + ri * ri;
  }
  __device__ void finish(KrState<T> *S, T t0, T t1, double *res) const {
    S->rho = t0;
    kr_finished(S, t1, res);
  }
};
template <typename T>
struct PcgUpdate {  // alpha = rho/<y,p>; x += alpha p; r -= alpha y; z = M r; rho' = <r,z>; beta; ++monitor; monitor(r)
  static constexpr bool GATED = true;
  static constexpr int NACC = 2;
  const T *p, *y, *dinv;
  T *x, *r;
  T alpha, nalpha;
  struct Regs { T p, y, x, r; };
  __device__ void prepare(const KrState<T> *S) { alpha = S->rho / S->d1; nalpha = -alpha; }
  __device__ void load(i64 i, Regs &g) const { g.p = p[i]; g.y = y[i]; g.x = x[i]; g.r = r[i]; }
  __device__ void apply(i64 i, Regs &g, T &a0, T &a1) const {
    x[i] = alpha * g.p + g.x;
    const T rn = nalpha * g.y + g.r;
    r[i] = rn;
    a0 = a0 + rn * apply_m(dinv, i, rn);
    a1 = a1 + rn * rn;
  }
  __device__ void finish(KrState<T> *S, T t0, T t1, double *res) const {
    S->beta = t0 / S->rho;
    S->rho = t0;
    S->iter += 1;
    kr_finished(S, t1, res);
  }
};
template <typename T>
struct PcgDirection {  // p = z + beta p
  static constexpr bool GATED = true;
  static constexpr int NACC = 0;
  const T *r, *dinv;
  T *p;
  T beta;
  struct Regs { T r, p; };
  __device__ void prepare(const KrState<T> *S) { beta = S->beta; }
  __device__ void load(i64 i, Regs &g) const { g.r = r[i]; g.p = p[i]; }
  __device__ void apply(i64 i, Regs &g, T &, T &) const { p[i] = T(1) * apply_m(dinv, i, g.r) + beta * g.p; }
  __device__ void finish(KrState<T> *, T, T, double *) const {}
};

// ---- BiCGStab ----
template <typename T>
struct BiInit {  // r = b - A x0; p = r; r* = r; (Mp = M p); rho = <r*, r>; monitor(r)
  static constexpr bool GATED = false;
  static constexpr int NACC = 2;
  const T *b, *y, *dinv;
  T *r, *p, *rstar, *Mp;  // Mp == nullptr for the identity (the products read p itself)
  struct Regs { T b, y; };
  __device__ void prepare(const KrState<T> *) {}
  __device__ void load(i64 i, Regs &g) const { g.b = b[i]; g.y = y[i]; }
  __device__ void apply(i64 i, Regs &g, T &a0, T &a1) const {
    const T ri = T(1) * g.b + T(-1) * g.y;
    r[i] = ri;
    p[i] = ri;
    rstar[i] = ri;
    if (Mp) Mp[i] = dinv[i] * ri;
    a0 = a0 + ri * ri;
    a1 = a1 + ri * ri;
  }
  __device__ void finish(KrState<T> *S, T t0, T t1, double *res) const {
    S->rho = t0;
    kr_finished(S, t1, res);
  }
};
template <typename T>
struct BiS {  // alpha = rho/<r*,AMp>; s = r - alpha AMp; x += alpha Mp; (Ms = M s); monitor(s)
  static constexpr bool GATED = true;
  static constexpr int NACC = 1;
  const T *r, *AMp, *Mp, *dinv;
  T *s, *x, *Ms;
  T alpha, nalpha;
  struct Regs { T r, a, m, x; };
  __device__ void prepare(const KrState<T> *S) { alpha = S->rho / S->d1; nalpha = -alpha; }
  __device__ void load(i64 i, Regs &g) const { g.r = r[i]; g.a = AMp[i]; g.m = Mp[i]; g.x = x[i]; }
  __device__ void apply(i64 i, Regs &g, T &a0, T &) const {
    const T si = T(1) * g.r + nalpha * g.a;
    s[i] = si;
    x[i] = T(1) * g.x + alpha * g.m;  // the first half of x + alpha Mp + omega Ms (and the whole early-exit update)
    if (Ms) Ms[i] = dinv[i] * si;
    a0 = a0 + si * si;
  }
  __device__ void finish(KrState<T> *S, T t0, T, double *res) const {
    S->alpha = alpha;
    kr_finished(S, t0, res);
  }
};
template <typename T>
struct BiR {  // omega = <AMs,s>/<AMs,AMs>; x += omega Ms; r = s - omega AMs; rho' = <r*,r>; beta; ++monitor; monitor(r)
  static constexpr bool GATED = true;
  static constexpr int NACC = 2;
  const T *s, *AMs, *Ms, *rstar;
  T *x, *r;
  T omega, nomega;
  struct Regs { T s, a, m, x, q; };
  __device__ void prepare(const KrState<T> *S) { omega = S->d2 / S->d3; nomega = -omega; }
  __device__ void load(i64 i, Regs &g) const { g.s = s[i]; g.a = AMs[i]; g.m = Ms[i]; g.x = x[i]; g.q = rstar[i]; }
  __device__ void apply(i64 i, Regs &g, T &a0, T &a1) const {
    x[i] = g.x + omega * g.m;
    const T ri = T(1) * g.s + nomega * g.a;
    r[i] = ri;
    a0 = a0 + g.q * ri;
    a1 = a1 + ri * ri;
  }
  __device__ void finish(KrState<T> *S, T t0, T t1, double *res) const {
    S->omega = omega;
    S->beta = (t0 / S->rho) * (S->alpha / omega);
    S->rho = t0;
    S->iter += 1;
    kr_finished(S, t1, res);
  }
};
template <typename T>
struct BiP {  // p = r + beta p - beta omega AMp; (Mp = M p)
  static constexpr bool GATED = true;
  static constexpr int NACC = 0;
  const T *r, *AMp, *dinv;
  T *p, *Mp;
  T beta, c;
  struct Regs { T r, p, a; };
  __device__ void prepare(const KrState<T> *S) { beta = S->beta; c = -beta * S->omega; }
  __device__ void load(i64 i, Regs &g) const { g.r = r[i]; g.p = p[i]; g.a = AMp[i]; }
  __device__ void apply(i64 i, Regs &g, T &, T &) const {
    const T pi = T(1) * g.r + beta * g.p + c * g.a;
    p[i] = pi;
    if (Mp) Mp[i] = dinv[i] * pi;
  }
  __device__ void finish(KrState<T> *, T, T, double *) const {}
};

// ---- CR ----
template <typename T>
struct CrInit {  // r = b - A x0; z = M r; p = z; ||r||^2 for the first monitor call
  static constexpr bool GATED = false;
  static constexpr int NACC = 1;
  const T *b, *ax, *dinv;
  T *r, *z, *p;  // z == nullptr for the identity (z is r)
  struct Regs { T b, a; };
  __device__ void prepare(const KrState<T> *) {}
  __device__ void load(i64 i, Regs &g) const { g.b = b[i]; g.a = ax[i]; }
  __device__ void apply(i64 i, Regs &g, T &a0, T &) const {
    const T ri = T(1) * g.b + T(-1) * g.a;
    const T zi = apply_m(dinv, i, ri);
    r[i] = ri;
    if (z) z[i] = zi;
    if (p) p[i] = zi;
    a0 = a0 + ri * ri;
  }
  __device__ void finish(KrState<T> *S, T t0, T, double *) const { S->rr = t0; }
};
template <typename T>
struct CrInit2 {  // Az = y (= A p = A z); rho = <r, Az>; <y,y>; then the first monitor.finished(r)
  static constexpr bool GATED = false;
  static constexpr int NACC = 2;
  const T *r, *y;
  T *Az;
  struct Regs { T r, y; };
  __device__ void prepare(const KrState<T> *) {}
  __device__ void load(i64 i, Regs &g) const { g.r = r[i]; g.y = y[i]; }
  __device__ void apply(i64 i, Regs &g, T &a0, T &a1) const {
    Az[i] = g.y;
    a0 = a0 + g.r * g.y;
    a1 = a1 + g.y * g.y;
  }
  __device__ void finish(KrState<T> *S, T t0, T t1, double *res) const {
    S->rho = t0;
    S->d3 = t1;
    kr_finished(S, S->rr, res);
  }
};
template <typename T>
struct CrX {  // alpha = rho/<y,y>; x += alpha p; and, unless r is recomputed from b - A x, r -= alpha y; z = M r
  static constexpr bool GATED = true;
  static constexpr int NACC = 1;
  const T *p, *y, *dinv;
  T *x, *r, *z;
  int update_r;
  T alpha, nalpha;
  struct Regs { T p, x, y, r; };
  __device__ void prepare(const KrState<T> *S) { alpha = S->rho / S->d3; nalpha = -alpha; }
  __device__ void load(i64 i, Regs &g) const {
    g.p = p[i];
    g.x = x[i];
    if (update_r) { g.y = y[i]; g.r = r[i]; }
  }
  __device__ void apply(i64 i, Regs &g, T &a0, T &) const {
    x[i] = alpha * g.p + g.x;
    if (update_r) {
      const T rn = nalpha * g.y + g.r;
      r[i] = rn;
      if (z) z[i] = dinv[i] * rn;
      a0 = a0 + rn * rn;
    }
  }
  __device__ void finish(KrState<T> *S, T t0, T, double *) const {
    S->alpha = alpha;
    if (update_r) S->rr = t0;
  }
};
template <typename T>
struct CrPY {  // beta = rho'/rho; p = z + beta p; y = Az + beta y; <y,y>; ++monitor; monitor(r)
  static constexpr bool GATED = true;
  static constexpr int NACC = 1;
  const T *z, *Az;
  T *p, *y;
  T beta;
  struct Regs { T z, a, p, y; };
  __device__ void prepare(const KrState<T> *S) { beta = S->d1 / S->rho; }
  __device__ void load(i64 i, Regs &g) const { g.z = z[i]; g.a = Az[i]; g.p = p[i]; g.y = y[i]; }
  __device__ void apply(i64 i, Regs &g, T &a0, T &) const {
    p[i] = T(1) * g.z + beta * g.p;
    const T yi = T(1) * g.a + beta * g.y;
    y[i] = yi;
    a0 = a0 + yi * yi;
  }
  __device__ void finish(KrState<T> *S, T t0, T, double *res) const {
    S->beta = beta;
    S->rho = S->d1;
    S->d3 = t0;
    S->iter += 1;
    kr_finished(S, S->rr, res);
  }
};

// ============================== host driver ========================================
template <typename T>
struct KrCtx {
  b200sp_handle h;
  cudaStream_t st;
  const b200sp_matrix *A;
  const b200sp_cfg *cfg;
  i64 n, lo, hi;
  bool dist;
  KrState<T> *S;
  T *partials;
  unsigned int *ticket;
  double *res;
  i64 grid;

  template <typename F>
  b200sp_status step(F f) {  // fused vector kernel (+ all-reduce + scalar step when partitioned)
    if (dist) {
      kr_kernel<T, F, true><<<(unsigned)grid, KR_BLOCK, 0, st>>>(n, f, S, partials, ticket, res);
      B200SP_LAUNCH_CHECK(h, "kr_kernel");
      if (F::NACC > 0) {
        b200sp_status s = comm_allreduce_sum(h, st, &S->acc0, 2, sizeof(T) == 8);
        if (s != B200SP_OK) return s;
        kr_scalar_kernel<T, F><<<1, 1, 0, st>>>(f, S, res);
        B200SP_LAUNCH_CHECK(h, "kr_scalar_kernel");
      }
    } else {
      kr_kernel<T, F, false><<<(unsigned)grid, KR_BLOCK, 0, st>>>(n, f, S, partials, ticket, res);
      B200SP_LAUNCH_CHECK(h, "kr_kernel");
    }
    return B200SP_OK;
  }
  // out = A v (v: window base), optional dot <out, dotv> -> *dot (all-reduced when partitioned)
  b200sp_status product(T *vwin, T *out, const T *dotv, T *dot) {
    b200sp_status s;
    if (dist) {
      s = comm_halo_exchange_auto(h, st, vwin, n, lo, hi, sizeof(T));
      if (s != B200SP_OK) return s;
    }
    s = spmv_any<T>(h, st, A, vwin, out, 0, cfg, dotv, dot);
    if (s != B200SP_OK) return s;
    if (dist && dot) return comm_allreduce_sum(h, st, dot, 1, sizeof(T) == 8);
    return B200SP_OK;
  }
  b200sp_status dot(const T *u, const T *v, T *out) {
    b200sp_status s = reduce<T, 0>(h, st, n, u, v, out, nullptr);
    if (s != B200SP_OK) return s;
    if (dist) return comm_allreduce_sum(h, st, out, 1, sizeof(T) == 8);
    return B200SP_OK;
  }
};

template <typename T>
static b200sp_status krylov_impl(b200sp_handle h, cudaStream_t st, int solver, const b200sp_matrix *A,
                                 const b200sp_halo *halo, T *x, const T *b, const T *dinv, const b200sp_cg_params *params,
                                 const b200sp_cfg *cfg, b200sp_cg_result *result, double *residuals_host) {
  const bool dist = halo != nullptr;
  const i64 n = A->num_rows;
  const i64 lo = dist ? halo->halo_lo : 0, hi = dist ? halo->halo_hi : 0;
  B200SP_REQUIRE(h, dist || A->num_rows == A->num_cols, "krylov: matrix must be square");
  B200SP_REQUIRE(h, !dist || A->num_cols == n + lo + hi, "krylov: num_cols != halo_lo + local + halo_hi");
  B200SP_REQUIRE(h, !dist || h->nccl_comm, "krylov: partitioned call without b200sp_comm_init");
  b200sp_cg_params prm = params ? *params : b200sp_cg_params{500, 1e-5, 0.0, 0};
  if (prm.check_interval <= 0) prm.check_interval = 16;
  B200SP_REQUIRE(h, prm.iteration_limit >= 0 && prm.iteration_limit < (1ll << 30), "krylov: bad iteration limit");

  // workspace: plain vectors of n and windows of lo + n + hi (operands of A), 256-byte aligned pieces
  const i64 w = lo + n + hi;
  const size_t vec = (((size_t)n * sizeof(T)) + 255) & ~(size_t)255, win = (((size_t)w * sizeof(T)) + 255) & ~(size_t)255;
  const bool pre = dinv != nullptr;
  size_t need = 0;
  int nvec = 0, nwin = 0;
  if (solver == B200SP_SOLVER_CG) { nvec = 2; nwin = 1; }                       // y, r | p
  if (solver == B200SP_SOLVER_BICGSTAB) { nvec = pre ? 6 : 4; nwin = 2; }       // r, r*, AMp, AMs (, p, s) | Mp, Ms  (identity: p, s are the windows)
  if (solver == B200SP_SOLVER_CR) { nvec = pre ? 4 : 3; nwin = 3; }             // y, Az, Ax (, r) | z (identity: r), p, x copy
  need = (size_t)nvec * vec + (size_t)nwin * win + 256;
  if (h->cg_ws_bytes < need) {
    if (h->cg_ws) cudaFree(h->cg_ws);
    h->cg_ws = nullptr;
    h->cg_ws_bytes = 0;
    if (cudaMalloc(&h->cg_ws, need) != cudaSuccess) {
      cudaGetLastError();
      return set_error(h, B200SP_ALLOC_FAILED, "krylov: cannot allocate %zu B workspace", need);
    }
    h->cg_ws_bytes = need;
  }
  const size_t nres_cap = 2 * (size_t)prm.iteration_limit + 4;  // BiCGStab records two norms per iteration
  if (h->cg_residuals_cap < nres_cap) {
    if (h->cg_residuals) cudaFree(h->cg_residuals);
    h->cg_residuals = nullptr;
    h->cg_residuals_cap = 0;
    if (cudaMalloc(&h->cg_residuals, nres_cap * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      return set_error(h, B200SP_ALLOC_FAILED, "krylov: cannot allocate residual log");
    }
    h->cg_residuals_cap = nres_cap;
  }
  char *base = reinterpret_cast<char *>(h->cg_ws);
  auto take_vec = [&]() { T *p = reinterpret_cast<T *>(base); base += vec; return p; };
  auto take_win = [&]() { T *p = reinterpret_cast<T *>(base); base += win; return p; };

  KrCtx<T> c;
  c.h = h; c.st = st; c.A = A; c.cfg = cfg; c.n = n; c.lo = lo; c.hi = hi; c.dist = dist;
  c.S = reinterpret_cast<KrState<T> *>(h->dev_scalars);
  c.partials = reinterpret_cast<T *>(h->red_partials);
  c.ticket = h->red_counters + 10;
  c.res = h->cg_residuals;
  c.grid = ceil_div(n, (i64)KR_BLOCK * KR_UNROLL);
  if (c.grid > (i64)h->num_sms * 8) c.grid = (i64)h->num_sms * 8;
  if (c.grid < 1) c.grid = 1;
  KrState<T> *S = c.S;
  T *bn = reinterpret_cast<T *>(h->dev_scalars + 40);
  b200sp_status s;

  // ||b|| (monitor constructor)
  if (dist) {
    s = reduce<T, 0>(h, st, n, b, b, bn, nullptr);
    if (s != B200SP_OK) return s;
    s = comm_allreduce_sum(h, st, bn, 1, sizeof(T) == 8);
    if (s != B200SP_OK) return s;
    comm_sqrt_inplace(h, st, bn, sizeof(T) == 8);
  } else {
    s = reduce<T, 1>(h, st, n, b, nullptr, bn, nullptr);
    if (s != B200SP_OK) return s;
  }
  kr_setup_kernel<T><<<1, 1, 0, st>>>(S, bn, prm.relative_tolerance, prm.absolute_tolerance, (int)prm.iteration_limit);
  B200SP_LAUNCH_CHECK(h, "kr_setup_kernel");

  KrState<T> *hs = reinterpret_cast<KrState<T> *>(h->pinned_scalars);
  auto poll = [&]() -> b200sp_status {
    B200SP_CUDA(h, cudaMemcpyAsync(hs, S, sizeof(KrState<T>), cudaMemcpyDeviceToHost, st));
    B200SP_CUDA(h, cudaStreamSynchronize(st));
    return B200SP_OK;
  };
#define KR_TRY(expr)             \
  do {                           \
    s = (expr);                  \
    if (s != B200SP_OK) return s; \
  } while (0)

  if (solver == B200SP_SOLVER_CG) {
    T *y = take_vec(), *r = take_vec(), *pwin = take_win(), *p = pwin + lo;
    // y = A x0 (x0 through the p window: it needs its halo too)
    B200SP_CUDA(h, cudaMemcpyAsync(p, x, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    KR_TRY(c.product(pwin, y, nullptr, nullptr));
    KR_TRY(c.step(PcgInit<T>{b, y, dinv, r, p}));
    KR_TRY(poll());
    while (!hs->done) {
      for (int k = 0; k < prm.check_interval; ++k) {
        KR_TRY(c.product(pwin, y, p, &S->d1));
        KR_TRY(c.step(PcgUpdate<T>{p, y, dinv, x, r, T(0), T(0)}));
        KR_TRY(c.step(PcgDirection<T>{r, dinv, p, T(0)}));
      }
      KR_TRY(poll());
    }
  } else if (solver == B200SP_SOLVER_BICGSTAB) {
    T *r = take_vec(), *rstar = take_vec(), *AMp = take_vec(), *AMs = take_vec();
    T *Mpwin = take_win(), *Mswin = take_win();
    T *Mp = Mpwin + lo, *Ms = Mswin + lo;
    T *p = pre ? take_vec() : Mp, *sv = pre ? take_vec() : Ms;  // identity: M p is p, M s is s
    B200SP_CUDA(h, cudaMemcpyAsync(Mp, x, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    KR_TRY(c.product(Mpwin, AMp, nullptr, nullptr));  // r <- A x0 (in AMp)
    KR_TRY(c.step(BiInit<T>{b, AMp, dinv, r, p, rstar, pre ? Mp : nullptr}));
    KR_TRY(poll());
    while (!hs->done) {
      for (int k = 0; k < prm.check_interval; ++k) {
        KR_TRY(c.product(Mpwin, AMp, rstar, &S->d1));                       // AMp = A M p, <r*, AMp>
        KR_TRY(c.step(BiS<T>{r, AMp, Mp, dinv, sv, x, pre ? Ms : nullptr, T(0), T(0)}));
        KR_TRY(c.product(Mswin, AMs, sv, &S->d2));                          // AMs = A M s, <AMs, s>
        KR_TRY(c.dot(AMs, AMs, &S->d3));
        KR_TRY(c.step(BiR<T>{sv, AMs, Ms, rstar, x, r, T(0), T(0)}));
        KR_TRY(c.step(BiP<T>{r, AMp, dinv, p, pre ? Mp : nullptr, T(0), T(0)}));
      }
      KR_TRY(poll());
    }
  } else {  // CR
    T *y = take_vec(), *Az = take_vec(), *Ax = take_vec();
    T *zwin = take_win(), *pwin = take_win(), *xwin = take_win();
    T *z = zwin + lo, *p = pwin + lo, *xw = xwin + lo;
    T *r = pre ? take_vec() : z;  // identity: z is r, so r lives in the window the products read
    B200SP_CUDA(h, cudaMemcpyAsync(xw, x, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    KR_TRY(c.product(xwin, Ax, nullptr, nullptr));
    KR_TRY(c.step(CrInit<T>{b, Ax, dinv, r, pre ? z : nullptr, p}));
    KR_TRY(c.product(pwin, y, nullptr, nullptr));
    KR_TRY(c.step(CrInit2<T>{r, y, Az}));
    KR_TRY(poll());
    i64 it = 0;
    while (!hs->done) {
      for (int k = 0; k < prm.check_interval; ++k, ++it) {
        const bool recompute = !((it % 8) && (it > 0));  // cr.inl:94-108
        KR_TRY(c.step(CrX<T>{p, y, dinv, x, r, pre ? z : nullptr, recompute ? 0 : 1, T(0), T(0)}));
        if (recompute) {
          B200SP_CUDA(h, cudaMemcpyAsync(xw, x, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, st));
          KR_TRY(c.product(xwin, Ax, nullptr, nullptr));
          // r = b - A x; z = M r; ||r||^2 (the update kernels after the monitor has finished are gated; this one
          // is not, but x no longer changes then, so r is merely recomputed to the same values)
          KR_TRY(c.step(CrInit<T>{b, Ax, dinv, r, pre ? z : nullptr, nullptr}));
        }
        KR_TRY(c.product(zwin, Az, r, &S->d1));  // Az = A z, <r, Az>
        KR_TRY(c.step(CrPY<T>{z, Az, p, y, T(0)}));
      }
      KR_TRY(poll());
    }
  }
#undef KR_TRY
  result->iteration_count = hs->iter;
  result->converged = hs->converged;
  result->residual_norm = (double)hs->rnorm;
  result->b_norm = (double)hs->bnorm;
  result->num_residuals = hs->nres;
  if (residuals_host && hs->nres > 0) {
    B200SP_CUDA(h, cudaMemcpyAsync(residuals_host, c.res, (size_t)hs->nres * sizeof(double), cudaMemcpyDeviceToHost, st));
    B200SP_CUDA(h, cudaStreamSynchronize(st));
  }
  return B200SP_OK;
}

}  // namespace b200sp

extern "C" b200sp_status b200sp_krylov(b200sp_handle h, b200sp_stream stream, b200sp_solver solver,
                                       const b200sp_matrix *A, const b200sp_halo *halo, void *x, const void *b,
                                       const void *diagonal_inverse, const b200sp_cg_params *params,
                                       const b200sp_cfg *spmv_cfg, b200sp_cg_result *result, double *residuals_host) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, A && x && b && result, "krylov: null argument");
  B200SP_REQUIRE(h, solver == B200SP_SOLVER_CG || solver == B200SP_SOLVER_BICGSTAB || solver == B200SP_SOLVER_CR,
                 "krylov: unknown solver");
  // plain CG: the hand-fused three-kernel iteration (peer-memory exchange inside the kernels when partitioned)
  if (solver == B200SP_SOLVER_CG && !diagonal_inverse)
    return halo ? b200sp_cg_dist(h, stream, A, halo, x, b, params, spmv_cfg, result, residuals_host)
                : b200sp_cg(h, stream, A, x, b, params, spmv_cfg, result, residuals_host);
  if (A->dtype == B200SP_F32)
    return b200sp::krylov_impl<float>(h, (cudaStream_t)stream, (int)solver, A, halo, (float *)x, (const float *)b,
                                      (const float *)diagonal_inverse, params, spmv_cfg, result, residuals_host);
  if (A->dtype == B200SP_F64)
    return b200sp::krylov_impl<double>(h, (cudaStream_t)stream, (int)solver, A, halo, (double *)x, (const double *)b,
                                       (const double *)diagonal_inverse, params, spmv_cfg, result, residuals_host);
  return b200sp::set_error(h, B200SP_INVALID_INPUT, "krylov: unknown dtype %d", (int)A->dtype);
}
