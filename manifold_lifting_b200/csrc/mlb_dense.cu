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


// x <- A^-T x for LOWER-triangular A read by rows (the transposed solve of a Cholesky pair,
// `cho_solve`, and of `lu_triangular_solve(l_left, l_right^T, .)`): bottom-up, x_i is
// final once the rows below it are done; its multiple of row i is then taken off
// x_0 .. x_{i-1} by the lanes (no reduction).  Next row prefetched as above.
MLB_DEV void warp_tri_solve_lt(const double* __restrict__ A, int dim, double* x, int lane) {
  constexpr int KP = 4;
  double nxt[KP], nxt_diag = 1.0;
  auto fetch = [&](int i) {
    const double* r = A + (int64_t)i * dim;
#pragma unroll
    for (int k = 0; k < KP; ++k) nxt[k] = lane + 32 * k < i ? r[lane + 32 * k] : 0.0;
    nxt_diag = r[i];
  };
  if (dim > 0) fetch(dim - 1);
  for (int i = dim - 1; i >= 0; --i) {
    double cur[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) cur[k] = nxt[k];
    const double diag = nxt_diag;
    if (i > 0) fetch(i - 1);
    const double xi = x[i] / diag;
    __syncwarp();
    if (lane == 0) x[i] = xi;
#pragma unroll
    for (int k = 0; k < KP; ++k)
      if (lane + 32 * k < i) x[lane + 32 * k] = fma(-cur[k], xi, x[lane + 32 * k]);
    const double* r = A + (int64_t)i * dim;
    for (int j = lane + 32 * KP; j < i; j += 32) x[j] = fma(-r[j], xi, x[j]);
    __syncwarp();
  }
}

// ---- batched dense Cholesky (cholesky(covar) of systems.py:869), column by column: the
// diagonal entry by warp 0, then the rows below it in parallel over the warps, each a
// dot product of two rows of L split over the lanes (Cholesky-Crout; in place in global
// memory, L2-resident for the sizes of these models)
__global__ void __launch_bounds__(32 * kDenseWarps)
    k_dense_cholesky(int dim, const double* __restrict__ a, double* __restrict__ l) {
  __shared__ double diag_sh;
  const int lane = threadIdx.x, w = threadIdx.y;
  const int64_t chain = blockIdx.x;
  const double* A = a + chain * dim * dim;
  double* L = l + chain * dim * dim;
  for (int j = 0; j < dim; ++j) {
    if (w == 0) {
      double s = 0.0;
      for (int k = lane; k < j; k += 32) s = fma(L[(int64_t)j * dim + k], L[(int64_t)j * dim + k], s);
      s = warp_sum(s);
      if (lane == 0) {
        const double d = A[(int64_t)j * dim + j] - s;
        const double r = d > 0.0 ? sqrt(d) : nan("");  // not positive definite: NaN factor
        L[(int64_t)j * dim + j] = r, diag_sh = r;
      }
    }
    __syncthreads();
    const double rd = 1.0 / diag_sh;
    for (int i = j + 1 + w; i < dim; i += kDenseWarps) {
      double s = 0.0;
      for (int k = lane; k < j; k += 32) s = fma(L[(int64_t)i * dim + k], L[(int64_t)j * dim + k], s);
      s = warp_sum(s);
      if (lane == 0) L[(int64_t)i * dim + j] = (A[(int64_t)i * dim + j] - s) * rd;
    }
    for (int i = w * 32 + lane; i < j; i += 32 * kDenseWarps) L[(int64_t)i * dim + j] = 0.0;
    __syncthreads();
  }
}

// ---- block algebra of GaussianProcessModelSystem (systems.py:878-918): Jacobian blocks
// dy_du [Y][U] (row-major) and dy_dn = chol_covar [Y][Y]; Gram = dy_du dy_du^T + L L^T.
constexpr int kGpMaxU = MLB_GP_MAX_DIM_U;

// U x U helpers for one thread, matrices in shared memory
MLB_DEV void small_cholesky(double* c, int U) {  // in place, lower; upper zeroed
  for (int j = 0; j < U; ++j) {
    double d = c[j * U + j];
    for (int k = 0; k < j; ++k) d -= c[j * U + k] * c[j * U + k];
    const double r = d > 0.0 ? sqrt(d) : nan("");
    c[j * U + j] = r;
    for (int i = j + 1; i < U; ++i) {
      double t = c[i * U + j];
      for (int k = 0; k < j; ++k) t -= c[i * U + k] * c[j * U + k];
      c[i * U + j] = t / r;
    }
    for (int i = 0; i < j; ++i) c[i * U + j] = 0.0;
  }
}
MLB_DEV void small_cho_solve(const double* c, int U, double* x) {
  for (int i = 0; i < U; ++i) {
    double t = x[i];
    for (int k = 0; k < i; ++k) t -= c[i * U + k] * x[k];
    x[i] = t / c[i * U + i];
  }
  for (int i = U - 1; i >= 0; --i) {
    double t = x[i];
    for (int k = i + 1; k < U; ++k) t -= c[k * U + i] * x[k];
    x[i] = t / c[i * U + i];
  }
}
MLB_DEV void small_lu_solve(double* m, int U, double* x) {  // partial pivoting; m destroyed
  for (int j = 0; j < U; ++j) {
    int piv = j;
    for (int i = j + 1; i < U; ++i)
      if (fabs(m[i * U + j]) > fabs(m[piv * U + j])) piv = i;
    if (piv != j) {
      for (int k = 0; k < U; ++k) {
        const double t = m[j * U + k];
        m[j * U + k] = m[piv * U + k], m[piv * U + k] = t;
      }
      const double t = x[j];
      x[j] = x[piv], x[piv] = t;
    }
    const double r = 1.0 / m[j * U + j];
    for (int i = j + 1; i < U; ++i) {
      const double f = m[i * U + j] * r;
      for (int k = j + 1; k < U; ++k) m[i * U + k] -= f * m[j * U + k];
      x[i] -= f * x[j];
    }
  }
  for (int i = U - 1; i >= 0; --i) {
    double t = x[i];
    for (int k = i + 1; k < U; ++k) t -= m[i * U + k] * x[k];
    x[i] = t / m[i * U + i];
  }
}

// out_a = sum_k B[k][a] x[k], a < U (B [Y][U] row-major); warps over a, lanes over k
MLB_DEV void block_tmatvec(const double* __restrict__ B, int Y, int U, const double* x,
                           double* out, int lane, int w) {
  for (int a = w; a < U; a += kDenseWarps) {
    double s = 0.0;
    for (int k = lane; k < Y; k += 32) s = fma(B[(int64_t)k * U + a], x[k], s);
    s = warp_sum(s);
    if (lane == 0) out[a] = s;
  }
}

// capacitance matrix I + R^T Z of the solved columns Z [U][Y] (scratch) against R [Y][U]
MLB_DEV void capacitance(const double* __restrict__ R, const double* __restrict__ Z, int Y,
                         int U, double* cap, int lane, int w) {
  for (int p = w; p < U * U; p += kDenseWarps) {
    const int a = p / U, b = p % U;
    double s = 0.0;
    for (int k = lane; k < Y; k += 32) s = fma(R[(int64_t)k * U + a], Z[(int64_t)b * Y + k], s);
    s = warp_sum(s);
    if (lane == 0) cap[a * U + b] = s + (a == b ? 1.0 : 0.0);
  }
}

// shared layout of the GP kernels: xs [kDenseWarps][Y], then small [2 U^2 + 2 U]
// MODE 0: decompose_gram (systems.py:885-888)   -> chol_cap [U][U]
// MODE 1: lmult_by_inv_gram (systems.py:890-897), needs chol_cap
// MODE 2: lmult_by_inv_jacob_product (systems.py:899-911): (du, dn) = left blocks,
//         (du_r, dn_r) = right blocks
template <int MODE>
__global__ void __launch_bounds__(32 * kDenseWarps)
    k_gp_solve(int Y, int U, const double* __restrict__ du, const double* __restrict__ dn,
               const double* __restrict__ du_r, const double* __restrict__ dn_r,
               const double* __restrict__ chol_cap_in, const double* __restrict__ vct,
               double* __restrict__ work, double* __restrict__ out) {
  extern __shared__ double sh[];
  const int lane = threadIdx.x, w = threadIdx.y;
  const int64_t chain = blockIdx.x;
  const double* A = du + chain * Y * U;    // dy_du (left)
  const double* L = dn + chain * Y * Y;    // chol_covar (left)
  const double* Ar = MODE == 2 ? du_r + chain * Y * U : A;
  const double* Lr = MODE == 2 ? dn_r + chain * Y * Y : L;
  double* Z = work + chain * U * Y;        // solved columns [U][Y]
  double* xs = sh + (int64_t)w * Y;
  double* cap = sh + (int64_t)kDenseWarps * Y;  // [U][U]
  double* t2 = cap + U * U;                       // [U]
  if (MODE != 1) {
    // Z[a] = (L Lr^T)^-1 A[:, a]  (MODE 0: only the forward half, cap = I + W^T W)
    for (int a = w; a < U; a += kDenseWarps) {
      for (int k = lane; k < Y; k += 32) xs[k] = A[(int64_t)k * U + a];
      __syncwarp();
      warp_tri_solve<true>(L, Y, xs, lane);
      if (MODE == 2) warp_tri_solve_lt(Lr, Y, xs, lane);
      for (int k = lane; k < Y; k += 32) Z[(int64_t)a * Y + k] = xs[k];
      __syncwarp();
    }
    __syncthreads();
    if (MODE == 0) {  // cap = I + W^T W with W = L^-1 A: both factors are the solved columns
      for (int p = w; p < U * U; p += kDenseWarps) {
        const int a = p / U, b = p % U;
        double s = 0.0;
        for (int k = lane; k < Y; k += 32) s = fma(Z[(int64_t)a * Y + k], Z[(int64_t)b * Y + k], s);
        s = warp_sum(s);
        if (lane == 0) cap[a * U + b] = s + (a == b ? 1.0 : 0.0);
      }
      __syncthreads();
      if (w == 0 && lane == 0) small_cholesky(cap, U);
      __syncthreads();
      for (int p = w * 32 + lane; p < U * U; p += 32 * kDenseWarps) out[chain * U * U + p] = cap[p];
      return;
    }
    capacitance(Ar, Z, Y, U, cap, lane, w);
  } else {
    for (int p = w * 32 + lane; p < U * U; p += 32 * kDenseWarps) cap[p] = chol_cap_in[chain * U * U + p];
  }
  // t1 = (L Lr^T)^-1 v in xs of warp 0
  double* x0 = sh;
  for (int k = w * 32 + lane; k < Y; k += 32 * kDenseWarps) x0[k] = vct[chain * Y + k];
  __syncthreads();
  if (w == 0) {
    warp_tri_solve<true>(L, Y, x0, lane);
    warp_tri_solve_lt(Lr, Y, x0, lane);
  }
  __syncthreads();
  block_tmatvec(Ar, Y, U, x0, t2, lane, w);
  __syncthreads();
  if (w == 0 && lane == 0) {
    if (MODE == 1) small_cho_solve(cap, U, t2); else small_lu_solve(cap, U, t2);
  }
  __syncthreads();
  for (int k = w * 32 + lane; k < Y; k += 32 * kDenseWarps) {
    double s = vct[chain * Y + k];
    for (int a = 0; a < U; ++a) s = fma(-A[(int64_t)k * U + a], t2[a], s);
    x0[k] = s;
  }
  __syncthreads();
  if (w == 0) {
    warp_tri_solve<true>(L, Y, x0, lane);
    warp_tri_solve_lt(Lr, Y, x0, lane);
  }
  __syncthreads();
  for (int k = w * 32 + lane; k < Y; k += 32 * kDenseWarps) out[chain * Y + k] = x0[k];
}

// lmult_by_jacob_constr (systems.py:878-880): J v = dy_du v_u + L v_n; rmult (882-883):
// J^T w = (w^T dy_du, w^T L); log_det_sqrt_gram from the factors (913-918)
__global__ void __launch_bounds__(32 * kDenseWarps)
    k_gp_jacob_mult(int Y, int U, int transpose, const double* __restrict__ du,
                    const double* __restrict__ dn, const double* __restrict__ vct,
                    double* __restrict__ out) {
  const int lane = threadIdx.x, w = threadIdx.y;
  const int64_t chain = blockIdx.x;
  const double* A = du + chain * Y * U;
  const double* L = dn + chain * Y * Y;
  if (!transpose) {
    const double* v = vct + chain * (U + Y);
    for (int i = w; i < Y; i += kDenseWarps) {
      double s = 0.0;
      for (int a = lane; a < U; a += 32) s = fma(A[(int64_t)i * U + a], v[a], s);
      for (int k = lane; k <= i; k += 32) s = fma(L[(int64_t)i * Y + k], v[U + k], s);
      s = warp_sum(s);
      if (lane == 0) out[chain * Y + i] = s;
    }
  } else {
    const double* v = vct + chain * Y;
    double* o = out + chain * (U + Y);
    // column sums: thread per output column, rows walked together by a warp's lanes
    for (int c = w * 32 + lane; c < U + Y; c += 32 * kDenseWarps) {
      double s = 0.0;
      if (c < U) {
        for (int k = 0; k < Y; ++k) s = fma(v[k], A[(int64_t)k * U + c], s);
      } else {
        const int j = c - U;
        for (int k = j; k < Y; ++k) s = fma(v[k], L[(int64_t)k * Y + j], s);
      }
      o[c] = s;
    }
  }
}

__global__ void __launch_bounds__(128)
    k_gp_log_det(int64_t n, int Y, int U, const double* __restrict__ dn,
                 const double* __restrict__ chol_cap, double* __restrict__ out) {
  const int64_t chain = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= n) return;
  double s = 0.0;
  for (int a = 0; a < U; ++a) s += log(chol_cap[chain * U * U + a * U + a]);
  for (int k = 0; k < Y; ++k) s += log(dn[chain * Y * Y + (int64_t)k * Y + k]);
  out[chain] = s;
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

namespace {
template <int MODE>
int gp_solve(int64_t n, int Y, int U, const double* du, const double* dn, const double* du_r,
             const double* dn_r, const double* chol_cap, const double* vct, double* out,
             cudaStream_t s) {
  const size_t sh = sizeof(double) * ((size_t)kDenseWarps * Y + (size_t)U * U + U);
  if (!allow_shared(k_gp_solve<MODE>, sh)) {
    g_err = "Gaussian-process block algebra: shared-memory configuration failed";
    return MLB_ERR_CUDA;
  }
  double* work = nullptr;
  if (MODE != 1) MLB_CUDA_OK(scratch_alloc((void**)&work, sizeof(double) * (size_t)n * U * Y, s));
  k_gp_solve<MODE><<<(unsigned)n, dim3(32, kDenseWarps), sh, s>>>(Y, U, du, dn, du_r, dn_r,
                                                                   chol_cap, vct, work, out);
  const int rc = finish_launch();
  if (work) cudaFreeAsync(work, s);
  return rc;
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

int mlb_dense_cholesky(int64_t n, int32_t dim, const double* a, double* l, void* stream) {
  MLB_ARG(n >= 0 && dim >= 1 && dim <= MLB_DENSE_MAX_DIM && a && l);
  if (n == 0) return MLB_OK;
  k_dense_cholesky<<<(unsigned)n, dim3(32, kDenseWarps), 0, (cudaStream_t)stream>>>(dim, a, l);
  return finish_launch();
}


#define MLB_GP_ARGS() \
  MLB_ARG(n >= 0 && dim_y >= 1 && dim_y <= MLB_DENSE_MAX_DIM && dim_u >= 1 && dim_u <= kGpMaxU)

int mlb_gp_decompose_gram(int64_t n, int32_t dim_y, int32_t dim_u, const double* dy_du,
                          const double* dy_dn, double* chol_cap_mtx, void* stream) {
  MLB_GP_ARGS();
  MLB_ARG(dy_du && dy_dn && chol_cap_mtx);
  if (n == 0) return MLB_OK;
  return gp_solve<0>(n, dim_y, dim_u, dy_du, dy_dn, nullptr, nullptr, nullptr, nullptr,
                     chol_cap_mtx, (cudaStream_t)stream);
}

int mlb_gp_lmult_by_inv_gram(int64_t n, int32_t dim_y, int32_t dim_u, const double* dy_du,
                             const double* dy_dn, const double* chol_cap_mtx, const double* vct,
                             double* out, void* stream) {
  MLB_GP_ARGS();
  MLB_ARG(dy_du && dy_dn && chol_cap_mtx && vct && out);
  if (n == 0) return MLB_OK;
  return gp_solve<1>(n, dim_y, dim_u, dy_du, dy_dn, nullptr, nullptr, chol_cap_mtx, vct, out,
                     (cudaStream_t)stream);
}

int mlb_gp_lmult_by_inv_jacob_product(int64_t n, int32_t dim_y, int32_t dim_u,
                                      const double* dy_du_l, const double* dy_dn_l,
                                      const double* dy_du_r, const double* dy_dn_r,
                                      const double* vct, double* out, void* stream) {
  MLB_GP_ARGS();
  MLB_ARG(dy_du_l && dy_dn_l && dy_du_r && dy_dn_r && vct && out);
  if (n == 0) return MLB_OK;
  return gp_solve<2>(n, dim_y, dim_u, dy_du_l, dy_dn_l, dy_du_r, dy_dn_r, nullptr, vct, out,
                     (cudaStream_t)stream);
}

int mlb_gp_jacob_constr_mult(int64_t n, int32_t dim_y, int32_t dim_u, int32_t transpose,
                             const double* dy_du, const double* dy_dn, const double* vct,
                             double* out, void* stream) {
  MLB_GP_ARGS();
  MLB_ARG(dy_du && dy_dn && vct && out);
  if (n == 0) return MLB_OK;
  k_gp_jacob_mult<<<(unsigned)n, dim3(32, kDenseWarps), 0, (cudaStream_t)stream>>>(
      dim_y, dim_u, transpose, dy_du, dy_dn, vct, out);
  return finish_launch();
}

int mlb_gp_log_det_sqrt_gram(int64_t n, int32_t dim_y, int32_t dim_u, const double* dy_dn,
                             const double* chol_cap_mtx, double* out, void* stream) {
  MLB_GP_ARGS();
  MLB_ARG(dy_dn && chol_cap_mtx && out);
  if (n == 0) return MLB_OK;
  k_gp_log_det<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      n, dim_y, dim_u, dy_dn, chol_cap_mtx, out);
  return finish_launch();
}

}  // extern "C"
