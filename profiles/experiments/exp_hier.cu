// Experiment harness for the register-resident fused trajectory kernel (eight-schools J=8,
// Newton): times variants selected by -D flags on synthetic inputs of the bench's shape.
//   nvcc -O3 -std=c++20 -gencode arch=compute_100a,code=sm_100a -lineinfo -DKBLOCK=64 ...
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <random>
#include <vector>

#include "../manifold_lifting_b200/csrc/mlb_chmc.cuh"
#include "../manifold_lifting_b200/csrc/mlb_hier_lean.cuh"

#ifndef KBLOCK
#define KBLOCK 64
#endif
#ifndef MINB
#define MINB 1
#endif
#ifndef SOLVER
#define SOLVER MLB_SOLVER_NEWTON
#endif
#ifndef NSCHOOL
#define NSCHOOL 8
#endif
using namespace mlb;
using M = HierModel<NSCHOOL, 0>;
constexpr int Q = M::Q;

#ifdef MAXNREG
__global__ void __maxnreg__(MAXNREG)
#else
__global__ void __launch_bounds__(KBLOCK, MINB)
#endif
    k_lf(const __grid_constant__ M::Params P, int64_t n, int64_t ld, double* q, double* p,
         double dt, int n_step, Opts o, const int32_t* order, int32_t* status, int32_t* n_done,
         int32_t* itf, int32_t* itr) {
  __shared__ double commit[2 * Q][KBLOCK];
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int64_t i = order ? order[t] : t;
  double qv[Q], pv[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) qv[k] = q[k * ld + i], pv[k] = p[k * ld + i];
  const TrajOut r = run_trajectory<M, SOLVER, false, false, KBLOCK>(
      P, qv, pv, &commit[0][threadIdx.x], &commit[Q][threadIdx.x], n_step, dt, o, 0.0);
#pragma unroll
  for (int k = 0; k < Q; ++k) q[k * ld + i] = qv[k], p[k * ld + i] = pv[k];
  status[i] = r.status, n_done[i] = r.done, itf[i] = r.it_fwd, itr[i] = r.it_rev;
}

#ifndef LEANREG
#define LEANREG 144
#endif
__global__ void __maxnreg__(LEANREG)
    k_lean(const __grid_constant__ M::Params P, int64_t n, int64_t ld, double* q, double* p,
           double dt, int n_step, Opts o, const int32_t* order, int32_t* status, int32_t* n_done,
           int32_t* itf, int32_t* itr) {
  extern __shared__ double lean_sh[];
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int64_t i = order ? order[t] : t;
  LeanCol<M, KBLOCK> S{lean_sh + threadIdx.x};
  const TrajOut r = run_trajectory_lean<M, SOLVER, KBLOCK>(P, S, q + i, p + i, ld, n_step, dt, o);
  status[i] = r.status, n_done[i] = r.done, itf[i] = r.it_fwd, itr[i] = r.it_rev;
}

__global__ void k_project(const __grid_constant__ M::Params P, int64_t n, int64_t ld,
                          const double* q, double* p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double qv[Q], pv[Q];
  for (int k = 0; k < Q; ++k) qv[k] = q[k * ld + i], pv[k] = p[k * ld + i];
  Eval<M> E;
  evaluate_at<M>(P, qv, E);
  M::B::project_cotangent(E.J, E.G, pv);
  for (int k = 0; k < Q; ++k) p[k * ld + i] = pv[k];
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 65536;
  const int n_step = argc > 2 ? atoi(argv[2]) : 32;
  const int order_mode = argc > 3 ? atoi(argv[3]) : 0;  // 0 none, 1 sorted by work, 2 sorted by other-momentum work
  const int reps = 10;
  std::mt19937_64 rng(202101);
  std::normal_distribution<double> N01;
  std::uniform_real_distribution<double> U01;
  std::cauchy_distribution<double> C01;
  M::Params P;
  for (int j = 0; j < NSCHOOL; ++j) P.sigma[j] = 9.0 + 9.0 * U01(rng);
  for (int j = 0; j < NSCHOOL; ++j) P.y_obs[j] = 4.4 + 3.6 * N01(rng) + P.sigma[j] * N01(rng);
  const int64_t ld = n;
  std::vector<double> q((size_t)Q * n), p((size_t)Q * n), p2((size_t)Q * n);
  for (int64_t i = 0; i < n; ++i) {
    double u0 = 5.0 * N01(rng);
    double u1 = std::min(3.0, std::max(-2.0, std::log(std::fabs(5.0 * C01(rng)))));
    q[0 * ld + i] = u0, q[1 * ld + i] = u1;
    for (int j = 0; j < NSCHOOL; ++j) {
      double v = N01(rng);
      q[(2 + j) * ld + i] = v;
      q[(2 + NSCHOOL + j) * ld + i] = (P.y_obs[j] - u0 - std::exp(u1) * v) / P.sigma[j];
    }
    for (int k = 0; k < Q; ++k) p[k * ld + i] = N01(rng), p2[k * ld + i] = N01(rng);
  }
  double *dq0, *dp0, *dp2, *dq, *dp;
  int32_t *st, *nd, *itf, *itr, *dorder;
  size_t bytes = sizeof(double) * Q * n;
  CK(cudaMalloc(&dq0, bytes)); CK(cudaMalloc(&dp0, bytes)); CK(cudaMalloc(&dp2, bytes));
  CK(cudaMalloc(&dq, bytes)); CK(cudaMalloc(&dp, bytes));
  CK(cudaMalloc(&st, 4 * n)); CK(cudaMalloc(&nd, 4 * n)); CK(cudaMalloc(&itf, 4 * n));
  CK(cudaMalloc(&itr, 4 * n)); CK(cudaMalloc(&dorder, 4 * n));
  CK(cudaMemcpy(dq0, q.data(), bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dp0, p.data(), bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dp2, p2.data(), bytes, cudaMemcpyHostToDevice));
  k_project<<<(n + 127) / 128, 128>>>(P, n, ld, dq0, dp0);
  k_project<<<(n + 127) / 128, 128>>>(P, n, ld, dq0, dp2);
  Opts o{1e-9, 1e-8, 1e10, 2e-8, 50, 10, MLB_NORM_MAX, MLB_NORM_MAX, 0};
  const double dt = 0.1;
  char* flush;
  CK(cudaMalloc(&flush, 256 << 20));
  const unsigned grid = (unsigned)((n + KBLOCK - 1) / KBLOCK);
#ifdef LEAN
  CK(cudaFuncSetAttribute(k_lean, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  CK(cudaFuncSetAttribute(k_lean, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)(sizeof(double) * LeanCol<M, KBLOCK>::ROWS * KBLOCK)));
  { int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_lean, KBLOCK, sizeof(double) * LeanCol<M, KBLOCK>::ROWS * KBLOCK); printf("lean occupancy: %d CTAs/SM\n", nb); }
#endif
  auto launch = [&](const double* psrc, const int32_t* ord) {
    CK(cudaMemcpy(dq, dq0, bytes, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(dp, psrc, bytes, cudaMemcpyDeviceToDevice));
    CK(cudaMemset(flush, 0, 256 << 20));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
#ifdef LEAN
    k_lean<<<grid, KBLOCK, sizeof(double) * LeanCol<M, KBLOCK>::ROWS * KBLOCK>>>(
        P, n, ld, dq, dp, dt, n_step, o, ord, st, nd, itf, itr);
#else
    k_lf<<<grid, KBLOCK>>>(P, n, ld, dq, dp, dt, n_step, o, ord, st, nd, itf, itr);
#endif
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
  };
  std::vector<int32_t> hf(n), hr(n), hs(n), hd(n), order(n);
  const int32_t* ord = nullptr;
  if (order_mode) {
    launch(order_mode == 2 ? dp2 : dp0, nullptr);
    CK(cudaMemcpy(hf.data(), itf, 4 * n, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hr.data(), itr, 4 * n, cudaMemcpyDeviceToHost));
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return hf[a] + hr[a] > hf[b] + hr[b]; });
    CK(cudaMemcpy(dorder, order.data(), 4 * n, cudaMemcpyHostToDevice));
    ord = dorder;
  }
  for (int w = 0; w < 3; ++w) launch(dp0, ord);
  double tot = 0, best = 1e9;
  for (int r = 0; r < reps; ++r) {
    float ms = launch(dp0, ord);
    tot += ms, best = std::min(best, (double)ms);
  }
  CK(cudaMemcpy(hf.data(), itf, 4 * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hr.data(), itr, 4 * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hs.data(), st, 4 * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hd.data(), nd, 4 * n, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(q.data(), dq, bytes, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(p.data(), dp, bytes, cudaMemcpyDeviceToHost));
  double its = 0, done = 0, cs = 0, cp = 0;
  int64_t fail = 0;
  for (int64_t i = 0; i < n; ++i) its += hf[i] + hr[i], done += hd[i], fail += hs[i] != 0;
  for (size_t k = 0; k < q.size(); ++k) cs += q[k], cp += p[k];
  cudaFuncAttributes fa;
#ifdef LEAN
  cudaFuncGetAttributes(&fa, k_lean);
#else
  cudaFuncGetAttributes(&fa, k_lf);
#endif
  printf("n=%lld n_step=%d order=%d kblock=%d regs=%d local=%zuB | mean %.4f ms best %.4f ms | "
         "%.3e steps/s | iters/step %.3f failed %lld | checksum q %.10e p %.10e\n",
         (long long)n, n_step, order_mode, KBLOCK, fa.numRegs, fa.localSizeBytes, tot / reps, best,
         done / (tot / reps * 1e-3), its / done, (long long)fail, cs, cp);
  return 0;
}
