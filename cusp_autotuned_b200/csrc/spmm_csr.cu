// spmm_csr.cu — CSR x dense block:  Y[rows x k] = (init Y) + A * X[cols x k], X and Y row-major
// ("block SpMV" for block Krylov methods).  Replaces, for csr_matrix x array2d on device_memory,
// BlockSpmvKernel / __spmv_csr_block (cusp/system/cuda/detail/multiply/csr_block_spmv.h:36-222).
// Semantics: host loop cusp/system/detail/sequential/multiply/csr_block_spmv.h:52-77
//     acc[j] = init(Y(i,j));  for jj ascending: acc[j] += Ax[jj] * X(Aj[jj], j);  Y(i,j) = acc[j]
// — per (row, column) the same summation order, so results are bit-identical to it (-fmad=false).
//
// Two kernels:
//   * csr_spmm_ring_kernel (default; needs 16-byte-aligned matrix arrays): persistent CTAs, the matrix streams
//     staged by a producer warp with cp.async.bulk into an mbarrier ring, a lane owning 4 fp32 / 2 fp64 adjacent
//     columns of the block (128-bit gathers of X, 128-bit stores of Y) — described at the kernel below;
//   * csr_spmm_kernel (round 1; unaligned arrays, B200SP_SPMM_LDG=1): a sub-warp of K lanes (K = 2..32 >= min(k,32))
//     owns a row, lane j < k owns column j; the K lanes load K consecutive entries of the row coalesced
//     (ld.global.cs) and hand them round with shuffles, one coalesced K*sizeof(T) gather of X (ld.global.nc) per
//     entry, no shared memory.
// Blocks wider than one pass (64 fp32 columns with 128-bit accesses, otherwise 32) are processed in column chunks.
// Algorithmic bytes: (rows+1)*4 + nnz*(4+V) + cols*k*V + rows*k*V.
#include <stdlib.h>

#include "common.cuh"

namespace b200sp {

// spmv_csr.cu
template <typename T>
b200sp_status spmv_csr(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 nnz, const int *Ap, const int *Aj,
                       const T *Ax, const T *x, T *y, int accumulate, const b200sp_cfg *cfg, const T *dotv,
                       T *dot_result);

template <typename T>
struct SpmmArgs {
  i64 rows, cols;
  const int *Ap, *Aj;
  const T *Ax;
  const T *X;
  T *Y;
  i64 ldx, ldy;
  int k;  // columns handled by this launch (<= 32), starting at X / Y
  int accumulate;
};

// R rows per sub-warp, interleaved over the CTA (row = first + i*(256/K) + sub-warp): their
// offsets, entries and gathers are issued together, so a thread has R*G independent gathers in
// flight instead of a chain of dependent loads per row (one row per sub-warp ran at 1-1.6 TB/s,
// latency-bound).
template <typename T, int K, int R>
__global__ void __launch_bounds__(256) csr_spmm_kernel(SpmmArgs<T> a) {
  constexpr int G = 4;          // gathers in flight per lane and row
  constexpr int SW = 256 / K;   // sub-warps per CTA
  const int lane = threadIdx.x & (K - 1);
  const i64 first = (i64)blockIdx.x * (SW * R) + threadIdx.x / K;
  const bool col_ok = lane < a.k;
  const int xl = col_ok ? lane : 0;
  int lo[R], hi[R];
  T acc[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const i64 row = first + (i64)i * SW;
    lo[i] = hi[i] = 0;
    if (row < a.rows) {
      lo[i] = ld_ro(a.Ap + row);
      hi[i] = ld_ro(a.Ap + row + 1);
    }
  }
  int span = 0;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const i64 row = first + (i64)i * SW;
    acc[i] = (row < a.rows && col_ok && a.accumulate) ? a.Y[row * a.ldy + lane] : T(0);
    span = max(span, hi[i] - lo[i]);
  }
  // the longest row among the warp's rows bounds the trip count (shuffles need uniform control)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) span = max(span, __shfl_xor_sync(0xffffffffu, span, o));
  for (int base = 0; base < span; base += K) {
    int c[R];
    T v[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int jj = lo[i] + base + lane;
      c[i] = 0;
      v[i] = T(0);
      if (jj < hi[i]) {
        c[i] = ld_stream(a.Aj + jj);
        v[i] = ld_stream(a.Ax + jj);
      }
    }
    const int nmax = min(K, span - base);  // warp-uniform
#pragma unroll
    for (int g0 = 0; g0 < K; g0 += G) {
      if (g0 >= nmax) break;
      constexpr int GQ = G < K ? G : K;  // sub-warps narrower than G hold fewer entries per chunk
      T vv[R][GQ], xv[R][GQ];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int q = 0; q < GQ; ++q) {
          const int cc = __shfl_sync(0xffffffffu, c[i], g0 + q, K);
          vv[i][q] = __shfl_sync(0xffffffffu, v[i], g0 + q, K);
          xv[i][q] = ld_ro(a.X + (i64)cc * a.ldx + xl);  // slots past the row's end gather row 0 of X (valid)
        }
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int n = min(K, hi[i] - lo[i] - base);  // entries of row i in this chunk (<= 0: none)
#pragma unroll
        for (int q = 0; q < GQ; ++q) {
          pin(xv[i][q]);
          const T t = acc[i] + vv[i][q] * xv[i][q];
          acc[i] = (g0 + q < n) ? t : acc[i];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const i64 row = first + (i64)i * SW;
    if (row < a.rows && col_ok) a.Y[row * a.ldy + lane] = acc[i];
  }
}

template <typename T, int K, int R>
static b200sp_status launch_spmm(b200sp_handle h, cudaStream_t st, const SpmmArgs<T> &a) {
  const i64 grid = ceil_div(a.rows, (i64)(256 / K) * R);
  csr_spmm_kernel<T, K, R><<<(unsigned)grid, 256, 0, st>>>(a);
  B200SP_LAUNCH_CHECK(h, "csr_spmm_kernel");
  return B200SP_OK;
}


// ---------------------------------------------------------------------------
// csr_spmm_ring_kernel — the block product on the persistent bulk-copy ring of K_CSR_RING.
//
// The one-shot kernel above spends its time issuing instructions, not moving bytes: ncu on poisson7pt 256^3, k = 32
// (profiles/r04_spmm.md) shows 23 warp instructions per stored entry, issue slots 64 % busy, DRAM 18 %, L1TEX 44 %.
// Two changes here:
//   * the matrix streams are decoupled from the gathers: persistent CTAs of 256 consumer threads + one producer
//     warp; tiles of R consecutive rows; per tile the producer warp writes the tile's R+1 row offsets into the stage
//     (prefetched one tile ahead in registers) and its first lane issues two cp.async.bulk copies of the tile's
//     contiguous [Ap[r0], Ap[r0+R)) range of Aj / Ax (L2 evict-first), up to `stages` chunks ahead;
//   * a lane owns V adjacent columns of the block (V = 4 fp32 / 2 fp64: one 128-bit ld.global.nc per entry and lane,
//     one 128-bit store of Y), so a sub-warp of K = k / V lanes covers a row and a warp works on 32 / K rows at
//     once: index, address and shared-memory instructions are shared by V columns.  Blocks whose width, leading
//     dimensions or base addresses do not allow 16-byte accesses run the same kernel with V = 1.
// Entry order per (row, column) is storage order: bit-identical to the host loop.  A tile with more entries than a
// stage becomes several chunks; a row that continues in the next chunk parks its partial sums in Y (the same thread
// picks them up again).
// ---------------------------------------------------------------------------
constexpr int SPMM_BLOCK = 256;
constexpr int SPMM_CAP = 2048;           // entries per stage
constexpr int SPMM_OSTR = 264;           // row offsets per stage (R + 1 <= 257)

template <typename T, int V>
struct BlockVec {
  T v[V];
};
template <typename T, int V>
__device__ __forceinline__ BlockVec<T, V> ldv_ro(const T *p) {
  BlockVec<T, V> r;
  if constexpr (V == 1) {
    r.v[0] = ld_ro(p);
  } else if constexpr (sizeof(T) == 4 && V == 2) {
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(r.v[0]), "=f"(r.v[1]) : "l"(p));
  } else if constexpr (sizeof(T) == 4) {
    static_assert(V == 4, "fp32: two or four columns per lane");
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p));
  } else {
    static_assert(V == 2, "fp64: two columns per lane");
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p));
  }
  return r;
}
template <typename T, int V>
__device__ __forceinline__ BlockVec<T, V> ldv(const T *p) {  // coherent load (Y: partial sums parked by this thread)
  BlockVec<T, V> r;
  if constexpr (V == 1) r.v[0] = *p;
  else if constexpr (sizeof(T) == 4 && V == 2) {
    const float2 t = *reinterpret_cast<const float2 *>(p);
    r.v[0] = t.x; r.v[1] = t.y;
  } else if constexpr (sizeof(T) == 4) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    const double2 t = *reinterpret_cast<const double2 *>(p);
    r.v[0] = t.x; r.v[1] = t.y;
  }
  return r;
}
template <typename T, int V>
__device__ __forceinline__ void stv(T *p, const BlockVec<T, V> &r) {
  if constexpr (V == 1) *p = r.v[0];
  else if constexpr (sizeof(T) == 4 && V == 2) *reinterpret_cast<float2 *>(p) = make_float2(r.v[0], r.v[1]);
  else if constexpr (sizeof(T) == 4) *reinterpret_cast<float4 *>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  else *reinterpret_cast<double2 *>(p) = make_double2(r.v[0], r.v[1]);
}

template <typename T, int K, int V, int MINB, int U>
__global__ void __launch_bounds__(SPMM_BLOCK + 32, MINB) csr_spmm_ring_kernel(SpmmArgs<T> a, int R, int stages,
                                                                            i64 num_tiles, i64 nnz, i64 per_cta, int cap) {
  constexpr int BLOCK = SPMM_BLOCK, OSTR = SPMM_OSTR;
  const int CAP = cap, STR = cap + 16;  // entries per stage; stride with alignment shift + read-ahead slack
  constexpr int EPV = 16 / (int)sizeof(T);
  constexpr int G = 8;           // gathers in flight per lane
  constexpr int SW = BLOCK / K;  // rows in flight per CTA
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *s_val = reinterpret_cast<T *>(smem_raw);
  int *s_col = reinterpret_cast<int *>(smem_raw + (size_t)stages * STR * sizeof(T));
  int *s_off = s_col + (size_t)stages * STR;
  uint64_t *full = reinterpret_cast<uint64_t *>(s_off + (size_t)stages * OSTR);
  uint64_t *empty = full + stages;

  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], BLOCK / 32);
    }
    mbar_fence_init();
  }
  // read-ahead slots must hold valid columns from the first tile on
  for (int i = tid; i < stages * STR; i += BLOCK + 32) s_col[i] = 0;
  __syncthreads();

  const i64 rows = a.rows;
  // tile sequence of this CTA: runs of `run` consecutive tiles, the runs dealt round-robin over the grid — the X
  // rows a tile leaves in L1 serve the next tile's gathers (banded operators) while all CTAs still sweep the matrix
  // together, so far neighbours stay in L2.  n-th tile of this CTA:
  const i64 run = per_cta > 0 ? per_cta : 1;
  auto tile_of = [&](i64 n) { return ((n / run) * (i64)gridDim.x + (i64)blockIdx.x) * run + (n % run); };
  if (tid >= BLOCK) {
    // ------------------------------ producer warp ---------------------------
    const int pl = tid - BLOCK;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint64_t pol = l2_policy_evict_first();
    const int nnz_c = (int)nnz & ~3, nnz_v = (int)nnz & ~(EPV - 1);  // last 16-byte-complete entry
    constexpr int OQ = (SPMM_BLOCK + 1 + 31) / 32;                   // offsets held per lane
    int off[OQ], noff[OQ];
    int s = 0, s0 = 0, s1 = 0;
    uint32_t ph = 0;
    i64 tn = 0;
    i64 tile = tile_of(0);
    if (tile < num_tiles) {
      const i64 r0 = tile * R;
      const int nr = (int)min((i64)R, rows - r0);
#pragma unroll
      for (int q = 0; q < OQ; ++q) off[q] = (pl + 32 * q <= nr) ? ld_ro(a.Ap + r0 + pl + 32 * q) : 0;
      s0 = ld_ro(a.Ap + r0);
      s1 = ld_ro(a.Ap + r0 + nr);
    }
    while (tile < num_tiles) {
      const int nr = (int)min((i64)R, rows - tile * R);
      const i64 next = tile_of(++tn);
      int n0 = 0, n1 = 0;
#pragma unroll
      for (int q = 0; q < OQ; ++q) noff[q] = 0;
      if (next < num_tiles) {  // the next tile's offsets: in flight while this tile is issued
        const i64 r0 = next * R;
        const int nnr = (int)min((i64)R, rows - r0);
#pragma unroll
        for (int q = 0; q < OQ; ++q) noff[q] = (pl + 32 * q <= nnr) ? ld_ro(a.Ap + r0 + pl + 32 * q) : 0;
        n0 = ld_ro(a.Ap + r0);
        n1 = ld_ro(a.Ap + r0 + nnr);
      }
      int lo = s0;
      do {  // at least one chunk per tile (its offsets travel with it), also when the tile has no entries
        const int hi = (s1 - lo > CAP) ? lo + CAP : s1;
        mbar_wait(&empty[s], ph ^ 1);
        int *so = s_off + (size_t)s * OSTR;
#pragma unroll
        for (int q = 0; q < OQ; ++q)
          if (pl + 32 * q <= nr) so[pl + 32 * q] = off[q];
        __syncwarp();
        if (pl == 0) {
          int *dc = s_col + (size_t)s * STR;
          T *dv = s_val + (size_t)s * STR;
          const int ga_c = lo & ~3, ga_v = lo & ~(EPV - 1);
          const int end_c = min((hi + 3) & ~3, nnz_c), end_v = min((hi + EPV - 1) & ~(EPV - 1), nnz_v);
          // the (at most 3) entries after the last complete 16 bytes of the arrays
          for (int j = max(end_c, ga_c); j < hi; ++j) dc[j - ga_c] = a.Aj[j];
          for (int j = max(end_v, ga_v); j < hi; ++j) dv[j - ga_v] = a.Ax[j];
          const int bc = (hi > lo) ? max(end_c - ga_c, 0) : 0, bv = (hi > lo) ? max(end_v - ga_v, 0) : 0;
          mbar_expect_tx(&full[s], (uint32_t)(bc * sizeof(int) + bv * sizeof(T)));
          if (bc > 0) bulk_g2s(dc, a.Aj + ga_c, (uint32_t)(bc * sizeof(int)), &full[s], pol);
          if (bv > 0) bulk_g2s(dv, a.Ax + ga_v, (uint32_t)(bv * sizeof(T)), &full[s], pol);
        }
        if (++s == stages) {
          s = 0;
          ph ^= 1;
        }
        lo = hi;
      } while (lo < s1);
      tile = next;
      s0 = n0;
      s1 = n1;
#pragma unroll
      for (int q = 0; q < OQ; ++q) off[q] = noff[q];
    }
  } else {
    // ------------------------------ consumers -------------------------------
    const int lane = tid & (K - 1), sub = tid / K;
    const bool col_ok = lane * V < a.k;  // vector mode: k % V == 0, so a lane's columns are all in or all out
    const T *Xl = a.X + (col_ok ? lane * V : 0);
    int s = 0;
    uint32_t ph = 0;
    for (i64 tn = 0;; ++tn) {
      const i64 tile = tile_of(tn);
      if (tile >= num_tiles) break;  // the tiles of a CTA ascend: nothing behind the first one past the end
      const i64 r0 = tile * R;
      const int nr = (int)min((i64)R, rows - r0);
      int lo = -1, s1 = 0;
      do {
        mbar_wait(&full[s], ph);
        const int *offs = s_off + (size_t)s * OSTR;
        if (lo < 0) lo = offs[0];
        s1 = offs[nr];
        const int hi = (s1 - lo > CAP) ? lo + CAP : s1;
        const bool last = hi == s1;
        const int *pc = s_col + (size_t)s * STR + (lo & 3) - lo;  // pc[j], absolute entry index j
        const T *pv = s_val + (size_t)s * STR + (lo & (EPV - 1)) - lo;
        T keep = T(0);
        for (int i0 = sub; i0 < nr; i0 += SW * U) {
          // U rows of the sub-warp are worked on together: U * G independent X-row gathers in flight per lane
          int b[U], e[U];
          bool act[U];
          BlockVec<T, V> acc[U];
          int len_max = 0;
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int i = i0 + u * SW;
            act[u] = false;
            b[u] = e[u] = lo;
#pragma unroll
            for (int v = 0; v < V; ++v) acc[u].v[v] = T(0);
            if (i < nr) {
              const int B = offs[i], E = offs[i + 1];
              const bool first = B >= lo && (B < hi || last);  // the chunk that initialises the row
              if (first || (B < hi && E > lo)) {               // else the row has nothing in this chunk
                act[u] = true;
                b[u] = max(B, lo);
                e[u] = max(min(E, hi), b[u]);
                // partial sums of a row continued from the previous chunk were parked in Y by this thread
                if (col_ok && (!first || a.accumulate)) acc[u] = ldv<T, V>(a.Y + (r0 + i) * a.ldy + lane * V);
              }
            }
            len_max = max(len_max, e[u] - b[u]);
          }
          for (int jj = 0; jj < len_max; jj += G) {
            BlockVec<T, V> xv[U][G];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int q = 0; q < G; ++q) {
                // slots past a row's end hold valid columns (next rows / zero fill); with several rows in step the
                // shorter ones stay at their own end
                const int idx = (U == 1) ? b[u] + jj + q : min(b[u] + jj + q, e[u]);
                xv[u][q] = ldv_ro<T, V>(Xl + (i64)(unsigned)pc[idx] * a.ldx);
              }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int q = 0; q < G; ++q)
#pragma unroll
                for (int v = 0; v < V; ++v) pin(xv[u][q].v[v]);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
              for (int q = 0; q < G; ++q) {
                if (b[u] + jj + q < e[u]) {
                  const T val = pv[b[u] + jj + q];
#pragma unroll
                  for (int v = 0; v < V; ++v) acc[u].v[v] = acc[u].v[v] + val * xv[u][q].v[v];
                }
              }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (act[u] && col_ok) stv<T, V>(a.Y + (r0 + i0 + u * SW) * a.ldy + lane * V, acc[u]);
            keep = keep + acc[u].v[0];
          }
        }
        consume_before_release(keep);
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        if (++s == stages) {
          s = 0;
          ph ^= 1;
        }
        lo = hi;
      } while (lo < s1);
    }
  }
}

template <typename T, int K, int V, int MINB, int U>
static b200sp_status launch_spmm_ring_m(b200sp_handle h, cudaStream_t st, const SpmmArgs<T> &a, i64 nnz, int stages,
                                        int cap) {
  const double mean = (double)nnz / (double)a.rows;
  i64 R = (i64)(0.9 * (double)cap / (mean > 1.0 ? mean : 1.0));
  constexpr int PASS = (SPMM_BLOCK / K) * U;  // rows per pass of the CTA
  R = (R / PASS) * PASS;
  if (R < PASS) R = PASS;
  if (R > SPMM_BLOCK) R = SPMM_BLOCK;
  const i64 num_tiles = ceil_div(a.rows, R);
  const size_t smem = (size_t)stages * ((size_t)(cap + 16) * (sizeof(T) + sizeof(int)) + SPMM_OSTR * sizeof(int)) +
                      2 * (size_t)stages * sizeof(uint64_t) + 16;
  auto kern = csr_spmm_ring_kernel<T, K, V, MINB, U>;
  B200SP_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int resident = 0;
  B200SP_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, SPMM_BLOCK + 32, smem));
  if (resident < 1) return set_error(h, B200SP_INVALID_INPUT, "csr spmm ring: configuration does not fit on an SM");
  // persistent, one wave; MINB CTAs per SM, the rest of the unified array stays L1 for the X rows
  i64 grid = (i64)h->num_sms * (resident < MINB ? resident : MINB);
  if (grid > num_tiles) grid = num_tiles;
  // runs of consecutive tiles per CTA, the runs dealt round-robin (kernel header); B200SP_SPMM_RUN overrides
  const char *rn = getenv("B200SP_SPMM_RUN");
  i64 per_cta = rn ? atoi(rn) : 1;
  if (per_cta < 1) per_cta = 1;
  if (grid * per_cta > num_tiles) grid = ceil_div(num_tiles, per_cta);
  kern<<<(unsigned)grid, SPMM_BLOCK + 32, smem, st>>>(a, (int)R, stages, num_tiles, nnz, per_cta, cap);
  B200SP_LAUNCH_CHECK(h, "csr_spmm_ring_kernel");
  return B200SP_OK;
}

template <typename T, int K, int V>
static b200sp_status launch_spmm_ring(b200sp_handle h, cudaStream_t st, const SpmmArgs<T> &a, i64 nnz) {
  // Defaults by measurement (poisson7pt 256^3, profiles/r04_spmm.md): three resident CTAs per SM (72 registers) with a
  // two-stage ring.  Measurement switches: B200SP_SPMM_MINB = resident CTAs per SM (2 | 3), B200SP_SPMM_STAGES = ring
  // depth (2..4), B200SP_SPMM_CAP = entries per stage (1024 | 2048).
  const char *mb = getenv("B200SP_SPMM_MINB"), *sg = getenv("B200SP_SPMM_STAGES"), *cp = getenv("B200SP_SPMM_CAP");
  const int minb = mb ? atoi(mb) : 3, stages = sg ? atoi(sg) : 2, cap = cp ? atoi(cp) : SPMM_CAP;
  if (stages < 2 || stages > 4) return set_error(h, B200SP_INVALID_INPUT, "csr spmm ring: stages %d", stages);
  if (cap != 1024 && cap != 2048) return set_error(h, B200SP_INVALID_INPUT, "csr spmm ring: stage capacity %d", cap);
  // (two rows in flight per sub-warp — U = 2 — measured: no change, 1.913 / 1.915 ms at k = 32 fp32; not instantiated)
  if (minb >= 3) return launch_spmm_ring_m<T, K, V, 3, 1>(h, st, a, nnz, stages, cap);
  return launch_spmm_ring_m<T, K, V, 2, 1>(h, st, a, nnz, stages, cap);
}

template <typename T>
b200sp_status spmm_csr(b200sp_handle h, cudaStream_t st, i64 rows, i64 cols, i64 nnz, const int *Ap, const int *Aj,
                       const T *Ax, i64 k, const T *X, i64 ldx, T *Y, i64 ldy, int accumulate) {
  B200SP_CHECK_HANDLE(h);
  B200SP_REQUIRE(h, rows >= 0 && cols >= 0 && nnz >= 0 && k >= 0, "spmm csr: negative dimension");
  B200SP_REQUIRE(h, rows < (1ll << 31) && cols < (1ll << 31) && nnz < (1ll << 31), "spmm csr: int32 index range");
  B200SP_REQUIRE(h, ldx >= k && ldy >= k, "spmm csr: leading dimension smaller than the block width");
  if (rows == 0 || k == 0) return B200SP_OK;
  B200SP_REQUIRE(h, Ap && Y, "spmm csr: null pointer");
  B200SP_REQUIRE(h, nnz == 0 || (Aj && Ax && X && cols > 0), "spmm csr: null pointer");
  B200SP_REQUIRE(h, rows * 32 / 256 < (1ll << 31), "spmm csr: too many rows for one launch");
  SpmmArgs<T> a;
  a.rows = rows; a.cols = cols; a.Ap = Ap; a.Aj = Aj; a.Ax = Ax; a.ldx = ldx; a.ldy = ldy; a.accumulate = accumulate;
  if (k == 1 && ldx == 1 && ldy == 1)  // a single contiguous column is a vector (csr_block_spmv.h:198-201)
    return spmv_csr<T>(h, st, rows, cols, nnz, Ap, Aj, Ax, X, Y, accumulate, nullptr, nullptr, nullptr);
  const T dummy = T(0);
  const char *force = getenv("B200SP_SPMM_LDG");  // measurement switch: the one-shot kernel
  const char *nv = getenv("B200SP_SPMM_NOVEC");  // measurement switch: one column per lane
  const bool novec = nv && nv[0] == '1';
  const bool ring = nnz > 0 && aligned16(Aj) && aligned16(Ax) && !(force && force[0] == '1');
  // column chunks: 16 lanes x 4 fp32 columns (64 wide) when the whole block allows 128-bit accesses — half as many
  // passes over the matrix for wide fp32 blocks — otherwise 32 columns per pass
  const bool wide = ring && sizeof(T) == 4 && !novec && k > 32 && ldx % 4 == 0 && ldy % 4 == 0 && aligned16(X) && aligned16(Y);
  i64 cw = 32;
  for (i64 c0 = 0; c0 < k; c0 += cw) {
    const i64 left = k - c0;
    // a 33..64-column chunk only when every lane gets four whole columns (16 lanes x 4); otherwise 32 at a time
    cw = (wide && left > 32 && (left >= 64 || left % 4 == 0)) ? (left < 64 ? left : 64) : (left < 32 ? left : 32);
    a.k = (int)cw;
    a.X = X ? X + c0 : &dummy;
    a.Y = Y + c0;
    b200sp_status s;
    // sub-warp width = block width rounded up to a power of two
    if (ring) {  // bulk copies need 16-byte-aligned array bases
      constexpr int VEC = 16 / (int)sizeof(T);  // columns per lane with 128-bit accesses
      const bool vec = a.k % VEC == 0 && ldx % VEC == 0 && ldy % VEC == 0 && aligned16(a.X) && aligned16(a.Y) && !novec;
      const bool vec2 = sizeof(T) == 4 && !vec && a.k % 2 == 0 && ldx % 2 == 0 && ldy % 2 == 0 &&
                        ((uintptr_t)a.X & 7) == 0 && ((uintptr_t)a.Y & 7) == 0 && !novec;  // fp32: 64-bit accesses
      const int lanes = vec ? a.k / VEC : (vec2 ? a.k / 2 : a.k);
      if (vec) {
        if (lanes <= 1) s = launch_spmm_ring<T, 1, VEC>(h, st, a, nnz);
        else if (lanes <= 2) s = launch_spmm_ring<T, 2, VEC>(h, st, a, nnz);
        else if (lanes <= 4) s = launch_spmm_ring<T, 4, VEC>(h, st, a, nnz);
        else if (lanes <= 8) s = launch_spmm_ring<T, 8, VEC>(h, st, a, nnz);
        else s = launch_spmm_ring<T, 16, VEC>(h, st, a, nnz);
      } else if (vec2) {
        if constexpr (sizeof(T) == 4) {
          if (lanes <= 1) s = launch_spmm_ring<T, 1, 2>(h, st, a, nnz);
          else if (lanes <= 2) s = launch_spmm_ring<T, 2, 2>(h, st, a, nnz);
          else if (lanes <= 4) s = launch_spmm_ring<T, 4, 2>(h, st, a, nnz);
          else if (lanes <= 8) s = launch_spmm_ring<T, 8, 2>(h, st, a, nnz);
          else s = launch_spmm_ring<T, 16, 2>(h, st, a, nnz);
        } else {
          s = B200SP_INVALID_INPUT;
        }
      } else {
        if (lanes <= 2) s = launch_spmm_ring<T, 2, 1>(h, st, a, nnz);
        else if (lanes <= 4) s = launch_spmm_ring<T, 4, 1>(h, st, a, nnz);
        else if (lanes <= 8) s = launch_spmm_ring<T, 8, 1>(h, st, a, nnz);
        else if (lanes <= 16) s = launch_spmm_ring<T, 16, 1>(h, st, a, nnz);
        else s = launch_spmm_ring<T, 32, 1>(h, st, a, nnz);
      }
      if (s != B200SP_OK) return s;
      continue;
    }
    // one-shot LDG kernel: 4 rows (2 for 32 lanes) per sub-warp
    if (a.k <= 2) s = launch_spmm<T, 2, 4>(h, st, a);
    else if (a.k <= 4) s = launch_spmm<T, 4, 4>(h, st, a);
    else if (a.k <= 8) s = launch_spmm<T, 8, 4>(h, st, a);
    else if (a.k <= 16) s = launch_spmm<T, 16, 4>(h, st, a);
    else s = launch_spmm<T, 32, 2>(h, st, a);
    if (s != B200SP_OK) return s;
  }
  return B200SP_OK;
}

}  // namespace b200sp

extern "C" {
#define DEF(T, sfx)                                                                                         \
  b200sp_status b200sp_spmm_csr_##sfx(b200sp_handle h, b200sp_stream stream, int64_t num_rows, int64_t num_cols, \
                                      int64_t num_entries, const int32_t *row_offsets,                      \
                                      const int32_t *column_indices, const T *values, int64_t block_cols,   \
                                      const T *X, int64_t ldx, T *Y, int64_t ldy, int accumulate) {         \
    return b200sp::spmm_csr<T>(h, (cudaStream_t)stream, num_rows, num_cols, num_entries, row_offsets,       \
                               column_indices, values, block_cols, X, ldx, Y, ldy, accumulate);             \
  }
DEF(float, f32)
DEF(double, f64)
#undef DEF
}
