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

}  // namespace

extern "C" {
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
