// Micro-benchmark: FP64 FMA issue rate of ONE warp per SM sub-partition as a function of the
// number of independent accumulator chains (ILP) and of resident warps per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double *out, int iters, long long *cyc)
{
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i)
    a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      a[i] = fma(a[i], b, c);
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i)
    s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0)
    *cyc = t1 - t0;
}
template <int ILP>
void run(int warps_per_smsp)
{
  double *out;
  long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 8);
  const int iters = 4096;
  k<ILP><<<148, 128 * warps_per_smsp>>>(out, iters, cyc);
  k<ILP><<<148, 128 * warps_per_smsp>>>(out, iters, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("warps/SMSP %d ILP %2d: %.2f clk per DFMA per warp, pipe util %.0f%%\n", warps_per_smsp, ILP,
         double(h) / (double(iters) * ILP), 100.0 * 2.0 * warps_per_smsp * iters * ILP / double(h));
  cudaFree(out);
  cudaFree(cyc);
}
int main()
{
  for (int w = 1; w <= 4; w *= 2)
    {
      run<1>(w); run<2>(w); run<4>(w); run<6>(w); run<8>(w); run<12>(w); run<16>(w);
    }
  return 0;
}
