// probe_kernels.cu — measurement-only kernels (NOT part of libb200sp.so; built into tools/_build/libprobe.so by
// tools/Makefile and driven by tools/gather_probe.py).  They answer one question of VERDICT r1 task 2: how fast can
// *any* kernel consume a scattered column stream on this GPU, with everything a sparse product does besides the
// gathers taken away?
//
// gather_sum_kernel<T, VPL, ROWS>: every lane reads VPL consecutive column indices and values (and, with ROWS, the
// row indices of a COO stream) with 128/256-bit streaming loads exactly like K_COO_WARP, gathers x[col] through
// ld.global.nc, adds the VPL products in a register and stores ONE value per lane (coalesced): no segmented scan, no
// shuffles, no carries, no row logic.  Its time on the R-MAT / random column streams is the floor for COO / CSR SpMV
// kernels that gather through L1TEX; compare with the engine's product on the same arrays.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

template <int W>
__device__ __forceinline__ void ld_words(const void *p, uint32_t *o) {
  if constexpr (W == 8) {
    asm volatile("ld.global.cs.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7])
                 : "l"(p));
  } else {
    static_assert(W == 4, "128- or 256-bit pieces");
    asm volatile("ld.global.cs.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]) : "l"(p));
  }
}
__device__ __forceinline__ float ld_x(const float *p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

template <int VPL, bool ROWS, bool VALS>
__global__ void __launch_bounds__(256) gather_sum_kernel(long long nnz, const int *__restrict__ Ai,
                                                         const int *__restrict__ Aj, const float *__restrict__ Ax,
                                                         const float *__restrict__ x, float *__restrict__ out) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long e0 = t * VPL;
  if (e0 + VPL > nnz) return;
  uint32_t c[VPL], r[VPL], v[VPL];
  ld_words<VPL>(Aj + e0, c);
  if (ROWS) ld_words<VPL>(Ai + e0, r);
  if (VALS) ld_words<VPL>(Ax + e0, v);
  float xv[VPL];
#pragma unroll
  for (int q = 0; q < VPL; ++q) xv[q] = ld_x(x + c[q]);
  float acc = 0.f;
#pragma unroll
  for (int q = 0; q < VPL; ++q) {
    float p = VALS ? __uint_as_float(v[q]) * xv[q] : xv[q];
    if (ROWS) p = ((int)r[q] >= 0) ? p : 0.f;  // consume the row word
    acc += p;
  }
  out[t] = acc;
}

// ---------------------------------------------------------------------------------------------------------------
// release_order_kernel<MODE>: the ring-stage release hazard of common.cuh (consume_before_release), reproduced in
// isolation, with the candidate cures side by side ("try fence.proxy.async.shared::cta").
//
// Persistent CTAs of 256 consumers + one producer lane; tiles of 2048 ints; src[(t % period) * 2048 + i] = t % period.
// The producer refills a stage as soon as its `empty` barrier completes (cp.async.bulk, mbarrier complete_tx).  A
// consumer loads its 8 ints of the stage into registers, RELEASES THE STAGE, and only then looks at the registers:
// every value must equal the tile's tag.  A mismatch means the load was still queued when the arrive let the producer
// overwrite the stage.
//   MODE 0  arrive right after the loads (nothing in between)
//   MODE 1  fence.proxy.async.shared::cta between the loads and the arrive (the architected generic->async ordering)
//   MODE 2  the library's cure: a compare-and-branch that consumes every loaded register before the arrive
//   MODE 3  membar.cta (fence.acq_rel.cta) between the loads and the arrive
template <int MODE>
__global__ void __launch_bounds__(288) release_order_kernel(const int *__restrict__ src, long long tiles, int period,
                                                            int stages, unsigned long long *bad) {
  constexpr int TILE = 2048, NPT = 8, GATHERS = 16;
  const unsigned span = (unsigned)period * TILE;
  int sink = 0;
  extern __shared__ __align__(128) unsigned char smem[];
  int *stage = reinterpret_cast<int *>(smem);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)stages * TILE * sizeof(int));
  uint64_t *empty = full + stages;
  const int tid = threadIdx.x;
  auto sptr = [](const void *p) { return (uint32_t)__cvta_generic_to_shared(p); };
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sptr(&empty[s])), "r"(8));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto wait = [&](uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(sptr(bar)), "r"(parity) : "memory");
  };
  if (tid >= 256) {
    if (tid == 256) {
      int s = 0;
      uint32_t ph = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        wait(&empty[s], ph ^ 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sptr(&full[s])), "r"(TILE * 4) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         sptr(stage + (size_t)s * TILE)), "l"(src + (size_t)(t % period) * TILE), "r"(TILE * 4), "r"(sptr(&full[s]))
                     : "memory");
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
    return;
  }
  int s = 0;
  uint32_t ph = 0;
  unsigned long long mism = 0;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    wait(&full[s], ph);
    const int *p = stage + (size_t)s * TILE;
    // what a sparse product has in the memory pipe in front of its shared-memory loads: scattered global gathers
    int g[GATHERS];
    unsigned hsh = (unsigned)(t * 2654435761u) + tid * 40503u;
#pragma unroll
    for (int q = 0; q < GATHERS; ++q) {
      hsh = hsh * 1664525u + 1013904223u;
      asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(g[q]) : "l"(src + (hsh % span)));
    }
    int v[NPT];
#pragma unroll
    for (int q = 0; q < NPT; ++q) asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v[q]) : "r"(sptr(p + q * 256 + tid)));
    if (MODE == 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (MODE == 3) asm volatile("fence.acq_rel.cta;" ::: "memory");
    if (MODE == 2) {
      int acc = 0;
#pragma unroll
      for (int q = 0; q < NPT; ++q) acc += v[q];
      if (acc == 0x7ff4dead) asm volatile("nanosleep.u32 0;");
    }
    __syncwarp();
    if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sptr(&empty[s])) : "memory");
    const int tag = (int)(t % period);
#pragma unroll
    for (int q = 0; q < NPT; ++q) mism += (v[q] != tag);
#pragma unroll
    for (int q = 0; q < GATHERS; ++q) sink += g[q];
    if (++s == stages) { s = 0; ph ^= 1; }
  }
  if (mism) atomicAdd(bad, mism);
  if (sink == 0x7ff4dead) atomicAdd(bad + 1, 1ull);  // keeps the gathers alive
}

// ---------------------------------------------------------------------------------------------------------------
// dma_fed_kernel: can ONE resident kernel, fed by the copy engine, keep both PCIe directions busy?  (The host
// pipeline of b200sp_spmv_host chains copy -> kernel -> copy per chunk with stream events and reaches 40 GB/s each
// way; the same copies without kernels reach 48.)  The host queues, on one copy stream, for every chunk k: the H2D
// copy of x piece k, then a 4-byte H2D copy that sets flag[k] = epoch.  This kernel is launched once on another
// stream; its CTAs walk the chunks in order, spin on flag[k + lag] and then stream chunk k from the device staging
// buffer to y in MAPPED HOST memory (y = 2 x: a stand-in for the chunk product, which adds 58 MB of HBM reads per chunk).
__global__ void __launch_bounds__(256) dma_fed_kernel(const double *__restrict__ x_dev, double *__restrict__ y_host,
                                                      long long n, int chunks, int lag, const volatile int *flag,
                                                      int epoch) {
  const long long per = (n + chunks - 1) / chunks;
  for (int k = 0; k < chunks; ++k) {
    const int need = min(k + lag, chunks - 1);
    if (threadIdx.x == 0)
      while (flag[need] != epoch) __nanosleep(200);
    __syncthreads();
    __threadfence();
    const long long b = (long long)k * per, e = min(n, b + per);
    for (long long i = b + (long long)blockIdx.x * 256 + threadIdx.x; i < e; i += (long long)gridDim.x * 256)
      y_host[i] = 2.0 * __ldcv(x_dev + i);
  }
}

}  // namespace

extern "C" {
int probe_dma_fed(const double *x_dev, double *y_host_mapped, long long n, int chunks, int lag, const int *flag, int epoch,
                  int ctas, void *stream) {
  dma_fed_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>(x_dev, y_host_mapped, n, chunks, lag, flag, epoch);
  return (int)cudaGetLastError();
}

// mode 0..3 as above; `bad` (device, zeroed by the caller) receives the number of stale values seen
int probe_release_order(int mode, long long tiles, int period, int stages, int ctas, const int *src,
                        unsigned long long *bad, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)stages * 2048 * sizeof(int) + 2 * (size_t)stages * sizeof(uint64_t);
#define RUN(M)                                                                                          \
  {                                                                                                     \
    cudaFuncSetAttribute(release_order_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    release_order_kernel<M><<<ctas, 288, smem, st>>>(src, tiles, period, stages, bad);                  \
  }
  if (mode == 0) RUN(0) else if (mode == 1) RUN(1) else if (mode == 2) RUN(2) else RUN(3)
#undef RUN
  return (int)cudaGetLastError();
}

// variant: 0 = columns only, 1 = columns + values (CSR stream), 2 = rows + columns + values (COO stream);
// vpl 4 | 8.  Returns cudaError_t of the launch.
int probe_gather_sum(int variant, int vpl, long long nnz, const int *Ai, const int *Aj, const float *Ax, const float *x,
                     float *out, void *stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long threads = nnz / vpl;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  if (grid == 0) return 0;
#define LAUNCH(V, R, X) gather_sum_kernel<V, R, X><<<grid, 256, 0, st>>>(nnz, Ai, Aj, Ax, x, out)
  if (vpl == 8) {
    if (variant == 0) LAUNCH(8, false, false);
    else if (variant == 1) LAUNCH(8, false, true);
    else LAUNCH(8, true, true);
  } else {
    if (variant == 0) LAUNCH(4, false, false);
    else if (variant == 1) LAUNCH(4, false, true);
    else LAUNCH(4, true, true);
  }
#undef LAUNCH
  return (int)cudaGetLastError();
}
}
