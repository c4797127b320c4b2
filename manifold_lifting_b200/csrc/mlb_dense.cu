// Dense helpers of `mlift.linalg` (reference mlift/linalg.py:6-43), batched over chains:
//
//   lu_triangular_solve(l, u, b)               solve l u x = b, l lower-, u upper-triangular
//   jvp_cholesky_mtx_mult_by_vct(t, chol, v)   d/dt [cholesky(mtx + t) v] for a symmetric
//                                              tangent t: chol Phi(chol^-1 t chol^-T) v,
//                                              Phi = lower triangle with the diagonal halved
//
// the two routines under the Gaussian-process systems (systems.py:871-899: Jacobian of
// y = cholesky(covar(u)) n with respect to u, and the solves with the non-symmetric product
// chol_l chol_r^T of the Newton iteration).  Arrays are the reference's with a leading chain
// axis, chain-major and row-major: matrices [n][dim][dim], right-hand sides [n][dim][n_rhs].
//
// One CTA per chain (per matrix), one warp per right-hand side at a time.  A triangular
// solve is row-oriented: the lanes of the warp split the columns of row i (coalesced
// 256-byte segments of the row), a shuffle tree adds the partial sums in a fixed order
// (deterministic) and lane 0 finishes x_i; the right-hand side lives in shared memory.
// The rows of the NEXT block are on their way while the current one is reduced (the loads
// do not depend on the solution), so a row costs a reduction, not a memory round trip.
// HBM view: every right-hand side re-reads the triangle (from L2 for all but the first warp
// of the CTA); the work is dim^2 / 2 fused multiply-adds per right-hand side.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/mlb.h"
#include "mlb_launch.cuh"

namespace mlb {

constexpr int kDenseWarps = 8;

MLB_DEV double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// x <- A^-1 x for the warp's right-hand side x [dim] in shared memory; A [dim][dim] row-major
// in global memory, lower (rows top-down) or upper (rows bottom-up) triangular.
template <bool LOWER>
MLB_DEV void warp_tri_solve(const double* __restrict__ A, int dim, double* x, int lane) {
  constexpr int KP = 4;  // 128 columns of the next row prefetched per lane
  double nxt[KP], nxt_diag = 1.0;
  auto row_of = [&](int step) { return LOWER ? step : dim - 1 - step; };
  auto col0 = [&](int i) { return LOWER ? 0 : i + 1; };        // first off-diagonal column
  auto cols = [&](int i) { return LOWER ? i : dim - 1 - i; };  // number of them
  auto fetch = [&](int i) {
    const double* r = A + (int64_t)i * dim + col0(i);
    const int m = cols(i);
#pragma unroll
    for (int k = 0; k < KP; ++k) nxt[k] = lane + 32 * k < m ? r[lane + 32 * k] : 0.0;
    nxt_diag = A[(int64_t)i * dim + i];
  };
  if (dim > 0) fetch(row_of(0));
  for (int step = 0; step < dim; ++step) {
    const int i = row_of(step), m = cols(i), c0 = col0(i);
    double cur[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) cur[k] = nxt[k];
    const double diag = nxt_diag;
    if (step + 1 < dim) fetch(row_of(step + 1));
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < KP; ++k)
      if (lane + 32 * k < m) s = fma(cur[k], x[c0 + lane + 32 * k], s);
    const double* r = A + (int64_t)i * dim + c0;
    for (int j = lane + 32 * KP; j < m; j += 32) s = fma(r[j], x[c0 + j], s);
    s = warp_sum(s);
    if (lane == 0) x[i] = (x[i] - s) / diag;
    __syncwarp();
  }
}

// ---- linalg.py:6-9
__global__ void __launch_bounds__(32 * kDenseWarps)
    k_lu_triangular_solve(int dim, int n_rhs, const double* __restrict__ l,
                          const double* __restrict__ u, const double* __restrict__ b,
                          double* __restrict__ x) {
  extern __shared__ double sh[];  // [kDenseWarps][dim]
  const int lane = threadIdx.x, w = threadIdx.y;
  const int64_t chain = blockIdx.x;
  const double* L = l + chain * dim * dim;
  const double* U = u + chain * dim * dim;
  const double* B = b + chain * dim * n_rhs;
  double* X = x + chain * dim * n_rhs;
  double* xs = sh + (int64_t)w * dim;
  for (int r = w; r < n_rhs; r += kDenseWarps) {
    for (int k = lane; k < dim; k += 32) xs[k] = B[(int64_t)k * n_rhs + r];
    __syncwarp();
    warp_tri_solve<true>(L, dim, xs, lane);
    warp_tri_solve<false>(U, dim, xs, lane);
    for (int k = lane; k < dim; k += 32) X[(int64_t)k * n_rhs + r] = xs[k];
    __syncwarp();
  }
}

// ---- linalg.py:12-43.  One CTA per (chain, direction); W [dim][dim] is its scratch.
__global__ void __launch_bounds__(32 * kDenseWarps)
    k_jvp_cholesky_mult(int dim, int n_dir, const double* __restrict__ tangents,
                        const double* __restrict__ chol, const double* __restrict__ vct,
                        double* __restrict__ work, double* __restrict__ out) {
  extern __shared__ double sh[];  // [kDenseWarps][dim] right-hand sides, [dim] y
  const int lane = threadIdx.x, w = threadIdx.y;
  const int64_t m = blockIdx.x, chain = m / n_dir;
  const double* T = tangents + m * dim * dim;
  const double* L = chol + chain * dim * dim;
  const double* v = vct + chain * dim;
  double* W = work + m * dim * dim;
  double* xs = sh + (int64_t)w * dim;
  double* y = sh + (int64_t)kDenseWarps * dim;
  // W = T L^-T: row r of W solves L w^T = T[r, :]^T
  for (int r = w; r < dim; r += kDenseWarps) {
    for (int k = lane; k < dim; k += 32) xs[k] = T[(int64_t)r * dim + k];
    __syncwarp();
    warp_tri_solve<true>(L, dim, xs, lane);
    for (int k = lane; k < dim; k += 32) W[(int64_t)r * dim + k] = xs[k];
    __syncwarp();
  }
  __syncthreads();
  // W <- L^-1 W column by column (only the lower triangle of the result is used)
  for (int c = w; c < dim; c += kDenseWarps) {
    for (int k = lane; k < dim; k += 32) xs[k] = W[(int64_t)k * dim + c];
    __syncwarp();
    warp_tri_solve<true>(L, dim, xs, lane);
    for (int k = c + lane; k < dim; k += 32) W[(int64_t)k * dim + c] = xs[k];
    __syncwarp();
  }
  __syncthreads();
  // y = Phi(W) v: lower triangle, diagonal halved
  for (int i = w; i < dim; i += kDenseWarps) {
    double s = 0.0;
    for (int c = lane; c <= i; c += 32)
      s = fma(W[(int64_t)i * dim + c] * (c == i ? 0.5 : 1.0), v[c], s);
    s = warp_sum(s);
    if (lane == 0) y[i] = s;
  }
  __syncthreads();
  // out = L y
  for (int i = w; i < dim; i += kDenseWarps) {
    double s = 0.0;
    for (int c = lane; c <= i; c += 32) s = fma(L[(int64_t)i * dim + c], y[c], s);
    s = warp_sum(s);
    if (lane == 0) out[m * dim + i] = s;
  }
}

}  // namespace mlb

using namespace mlb;

namespace {
template <class K>
bool allow_shared(K kernel, size_t bytes) {
  return bytes <= 48 * 1024 ||
         cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) ==
             cudaSuccess;
}
}  // namespace

extern "C" {

int mlb_lu_triangular_solve(int64_t n, int32_t dim, int32_t n_rhs, const double* l,
                            const double* u, const double* b, double* x, void* stream) {
  MLB_ARG(n >= 0 && dim >= 1 && dim <= MLB_DENSE_MAX_DIM && n_rhs >= 1 && l && u && b && x);
  if (n == 0) return MLB_OK;
  const size_t sh = sizeof(double) * kDenseWarps * (size_t)dim;
  if (!allow_shared(k_lu_triangular_solve, sh)) {
    g_err = "lu_triangular_solve: shared-memory configuration failed";
    return MLB_ERR_CUDA;
  }
  k_lu_triangular_solve<<<(unsigned)n, dim3(32, kDenseWarps), sh, (cudaStream_t)stream>>>(
      dim, n_rhs, l, u, b, x);
  return finish_launch();
}

int mlb_jvp_cholesky_mtx_mult_by_vct(int64_t n, int32_t n_dir, int32_t dim,
                                     const double* tangents, const double* chol_mtx,
                                     const double* vct, double* out, void* stream) {
  MLB_ARG(n >= 0 && n_dir >= 1 && dim >= 1 && dim <= MLB_DENSE_MAX_DIM && tangents &&
          chol_mtx && vct && out);
  if (n == 0) return MLB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t sh = sizeof(double) * (kDenseWarps + 1) * (size_t)dim;
  if (!allow_shared(k_jvp_cholesky_mult, sh)) {
    g_err = "jvp_cholesky_mtx_mult_by_vct: shared-memory configuration failed";
    return MLB_ERR_CUDA;
  }
  double* work = nullptr;
  MLB_CUDA_OK(scratch_alloc((void**)&work,
                            sizeof(double) * (size_t)n * n_dir * (size_t)dim * dim, s));
  k_jvp_cholesky_mult<<<(unsigned)(n * n_dir), dim3(32, kDenseWarps), sh, s>>>(
      dim, n_dir, tangents, chol_mtx, vct, work, out);
  const int rc = finish_launch();
  cudaFreeAsync(work, s);
  return rc;
}

}  // extern "C"
