// fp64 latency / throughput microbenchmark: ILP chains x warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, int iters, long long* cyc) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
  double b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], b, c);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int threads) {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  int iters = 2000;
  k<ILP><<<1, threads>>>(out, iters, cyc);
  k<ILP><<<1, threads>>>(out, iters, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * 8.0 * ILP);
  printf("ILP=%d threads=%4d (warps/SMSP=%.2f): %.2f cycles per DFMA per warp; SM fp64 pipe: %.1f%% of 2-cycle issue\n",
         ILP, threads, threads / 128.0, per, 100.0 * 2.0 / per / 1.0 * (threads/128.0 >= 1 ? 1 : 1));
  cudaFree(out); cudaFree(cyc);
}
// MUFU.RCP64H + 2 NR, dependent chain
__global__ void krcp(double* out, int iters, long long* cyc) {
  double x = 1.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0); y = fma(y, e, y); e = fma(-x, y, 1.0); x = fma(y, e, y) + 1.0;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void kexp(double* out, int iters, long long* cyc) {
  double x = 0.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) x = exp(x) * 0.3;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  for (int t : {32, 128, 256, 512}) { run<1>(t); run<2>(t); run<4>(t); run<8>(t); }
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8); long long h;
  krcp<<<1, 32>>>(out, 1000, cyc); krcp<<<1, 32>>>(out, 1000, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("rcp_fast+add dependent: %.1f cycles\n", h / 1000.0);
  kexp<<<1, 32>>>(out, 1000, cyc); kexp<<<1, 32>>>(out, 1000, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("exp+mul dependent: %.1f cycles\n", h / 1000.0);
  return 0;
}
