// Batched tridiagonal primitives of `mlift.linalg` (reference mlift/linalg.py:46-114), the
// linear algebra under the state-space systems' "tridiagonal + low rank" Gram matrices
// (mlift/systems.py:1189-1238): one chain per thread, every array component-major
// ([rows][ld], element (k, chain i) at p[k * ld + i]), so a warp's access to one row of 32
// chains is one coalesced 256-byte transaction and the kernels stream at HBM rate.  The
// arithmetic is the reference's Thomas recurrence and log-continuant recursion in the same
// order (the same log1m_exp branch; the single-right-hand-side solve multiplies by 1 / b'
// where the reference divides), so results match a NumPy restatement of those scans to a
// few ulp.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/mlb.h"
#include "mlb_launch.cuh"

namespace mlb {

// The recurrences are sequential in k, but their loads do not depend on computed values:
// each thread fetches kPre rows ahead into registers (3 kPre independent 8-byte loads in
// flight per thread) and then runs kPre dependent steps, so a DRAM round trip is paid once
// per kPre steps instead of once per step.
constexpr int kPre = 16;

// The library's own scratch pools (see mlb_launch.cuh), created on first use per device.
namespace {
constexpr int kMaxDevices = 64;
std::mutex g_pool_mutex;
cudaMemPool_t g_pools[kMaxDevices] = {};
}  // namespace

cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t s) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  cudaMemPool_t pool;
  {
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pools[dev]) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      if ((e = cudaMemPoolCreate(&g_pools[dev], &props)) != cudaSuccess) {
        g_pools[dev] = nullptr;
        return e;
      }
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &thr);
    }
    pool = g_pools[dev];
  }
  return cudaMallocFromPoolAsync(p, bytes, pool, s);
}

int scratch_trim(size_t keep_bytes) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  for (int d = 0; d < kMaxDevices; ++d)
    if (g_pools[d] && cudaMemPoolTrimTo(g_pools[d], keep_bytes) != cudaSuccess) return MLB_ERR_CUDA;
  return MLB_OK;
}

// forward elimination of mlift/linalg.py:67-75: w_i = a_{i-1} / b'_{i-1},
// b'_i = b_i - w_i c_{i-1};  fac rows: [0, dim) b', [dim, 2 dim - 1) w
__global__ void __launch_bounds__(128)
    k_tridiag_factor(int64_t n, int dim, int64_t ld, const double* a, const double* b,
                     const double* c, double* fac, int64_t ldf) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double bp = b[i];
  fac[i] = bp;
  for (int k0 = 1; k0 < dim; k0 += kPre) {
    double av[kPre], bv[kPre], cv[kPre];
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int k = k0 + j < dim ? k0 + j : dim - 1;
      av[j] = a[(int64_t)(k - 1) * ld + i], bv[j] = b[(int64_t)k * ld + i];
      cv[j] = c[(int64_t)(k - 1) * ld + i];
    }
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int k = k0 + j;
      if (k < dim) {
        const double w = av[j] / bp;
        bp = bv[j] - w * cv[j];
        fac[(int64_t)k * ldf + i] = bp;
        fac[(int64_t)(dim + k - 1) * ldf + i] = w;
      }
    }
  }
}

// d'_i = d_i - w_i d'_{i-1} (linalg.py:72-73), then the backward scan of :80-85
// x_i = (d'_i - c_i x_{i+1}) / b'_i; one right-hand side per blockIdx.y
__global__ void __launch_bounds__(128)
    k_tridiag_apply(int64_t n, int dim, int64_t ld, const double* c, const double* fac,
                    int64_t ldf, const double* d, double* x) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t off = (int64_t)blockIdx.y * dim * ld;
  d += off, x += off;
  double dp = d[i];
  x[i] = dp;
  for (int k0 = 1; k0 < dim; k0 += kPre) {
    double dv[kPre], wv[kPre];
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int k = k0 + j < dim ? k0 + j : dim - 1;
      dv[j] = d[(int64_t)k * ld + i], wv[j] = fac[(int64_t)(dim + k - 1) * ldf + i];
    }
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int k = k0 + j;
      if (k < dim) {
        dp = dv[j] - wv[j] * dp;
        x[(int64_t)k * ld + i] = dp;
      }
    }
  }
  double xn = dp / fac[(int64_t)(dim - 1) * ldf + i];
  x[(int64_t)(dim - 1) * ld + i] = xn;
  for (int k0 = dim - 2; k0 >= 0; k0 -= kPre) {
    double xv[kPre], cv[kPre], bv[kPre];
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int k = k0 - j >= 0 ? k0 - j : 0;
      xv[j] = x[(int64_t)k * ld + i], cv[j] = c[(int64_t)k * ld + i];
      bv[j] = fac[(int64_t)k * ldf + i];
    }
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int k = k0 - j;
      if (k >= 0) {
        xn = (xv[j] - cv[j] * xn) / bv[j];
        x[(int64_t)k * ld + i] = xn;
      }
    }
  }
}

// One right-hand side (the common case, systems.py:1199-1210, 1224-1231): elimination and
// both substitutions in ONE kernel.  The forward sweep reads a, b, c, d once and leaves
// 1 / b' (scratch) and d' (in x); the backward sweep reads them back with c: 10 rows of
// traffic against 12 for the two-kernel form, one reciprocal per row instead of two IEEE
// divisions on the dependent chain (x_i = (d'_i - c_i x_{i+1}) * (1 / b'_i): results agree
// with the reference's recurrence to a few ulp).
__global__ void __launch_bounds__(128)
    k_tridiag_solve_one(int64_t n, int dim, int64_t ld, const double* a, const double* b,
                        const double* c, const double* d, double* x, double* rb, int64_t ldf) {
  constexpr int P = 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double rbp = 0.0, dp = 0.0;
  for (int k0 = 0; k0 < dim; k0 += P) {
    double av[P], bv[P], cv[P], dv[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = k0 + j < dim ? k0 + j : dim - 1;
      av[j] = k > 0 ? a[(int64_t)(k - 1) * ld + i] : 0.0;
      cv[j] = k > 0 ? c[(int64_t)(k - 1) * ld + i] : 0.0;
      bv[j] = b[(int64_t)k * ld + i], dv[j] = d[(int64_t)k * ld + i];
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = k0 + j;
      if (k < dim) {
        const double w = av[j] * rbp;  // 0 in the first row
        rbp = 1.0 / (bv[j] - w * cv[j]);
        dp = dv[j] - w * dp;
        rb[(int64_t)k * ldf + i] = rbp;
        x[(int64_t)k * ld + i] = dp;
      }
    }
  }
  double xn = 0.0;
  for (int k0 = dim - 1; k0 >= 0; k0 -= P) {
    double xv[P], cv[P], rv[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = k0 - j >= 0 ? k0 - j : 0;
      xv[j] = x[(int64_t)k * ld + i], rv[j] = rb[(int64_t)k * ldf + i];
      cv[j] = k < dim - 1 ? c[(int64_t)k * ld + i] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int k = k0 - j;
      if (k >= 0) {
        xn = (xv[j] - cv[j] * xn) * rv[j];
        x[(int64_t)k * ld + i] = xn;
      }
    }
  }
}

// log(1 - exp(v)), mlift/math.py:7-14
__device__ __forceinline__ double log1m_exp(double v) {
  return v > 0.69314718055994530942 ? log(-expm1(v)) : log1p(-exp(v));
}

// log det of the symmetric positive-definite tridiagonal matrix (a, b, a) by the
// log-continuant recursion of linalg.py:104-114:
// l_{i+1} = log_diff_exp(log b_{i+1} + l_i, 2 log|a_i| + l_{i-1}), l_0 = log b_0, l_{-1} = 0
__global__ void __launch_bounds__(128)
    k_tridiag_log_det(int64_t n, int dim, int64_t ld, const double* a, const double* b,
                      double* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double l = log(b[i]), lm = 0.0;
  for (int k0 = 1; k0 < dim; k0 += kPre) {
    // the logs of the diagonals do not depend on the recursion: they are taken with the
    // prefetched values, off its critical path
    double lb[kPre], la[kPre];
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int k = k0 + j < dim ? k0 + j : dim - 1;
      lb[j] = log(b[(int64_t)k * ld + i]);
      la[j] = 2.0 * log(fabs(a[(int64_t)(k - 1) * ld + i]));
    }
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      if (k0 + j < dim) {
        const double v1 = lb[j] + l, v2 = la[j] + lm;
        lm = l;
        l = v1 + log1m_exp(v2 - v1);
      }
    }
  }
  out[i] = l;
}

}  // namespace mlb

using namespace mlb;

extern "C" {

int mlb_tridiagonal_solve(int64_t n, int32_t dim, int32_t n_rhs, int64_t ld, const double* a,
                          const double* b, const double* c, const double* d, double* x,
                          void* stream) {
  MLB_ARG(n >= 0 && dim >= 1 && n_rhs >= 1 && ld >= n && b && d && x && (dim == 1 || (a && c)));
  if (n == 0) return MLB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t ldf = (n + 31) / 32 * 32;
  double* fac = nullptr;
  const unsigned g = (unsigned)((n + 127) / 128);
  if (n_rhs == 1) {
    MLB_CUDA_OK(scratch_alloc((void**)&fac, sizeof(double) * (size_t)dim * (size_t)ldf, s));
    k_tridiag_solve_one<<<g, 128, 0, s>>>(n, dim, ld, a, b, c, d, x, fac, ldf);
    const int rc1 = finish_launch();
    cudaFreeAsync(fac, s);
    return rc1;
  }
  MLB_CUDA_OK(scratch_alloc((void**)&fac, sizeof(double) * (size_t)(2 * dim) * (size_t)ldf, s));
  k_tridiag_factor<<<g, 128, 0, s>>>(n, dim, ld, a, b, c, fac, ldf);
  int rc = finish_launch();
  if (rc == MLB_OK) {
    k_tridiag_apply<<<dim3(g, (unsigned)n_rhs), 128, 0, s>>>(n, dim, ld, c, fac, ldf, d, x);
    rc = finish_launch();
  }
  cudaFreeAsync(fac, s);
  return rc;
}

int mlb_tridiagonal_pos_def_log_det(int64_t n, int32_t dim, int64_t ld, const double* a,
                                    const double* b, double* out, void* stream) {
  MLB_ARG(n >= 0 && dim >= 1 && ld >= n && b && out && (dim == 1 || a));
  if (n == 0) return MLB_OK;
  k_tridiag_log_det<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n, dim, ld, a,
                                                                                  b, out);
  return finish_launch();
}

}  // extern "C"
