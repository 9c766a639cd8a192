// Micro-benchmark behind the "no tensor cores" decision: does the FP64 tensor
// instruction (DMMA, mma.sync m8n8k4 / m16n8k16 f64) on sm_100a deliver more FMA/clk/SM than the
// FP64 vector pipe (DFMA, 64 FMA/clk/SM), and do the two issue concurrently?
//   mode 0  DFMA only                      (ILP independent chains per warp)
//   mode 1  DMMA only                      (ILP independent accumulator fragments per warp)
//   mode 2  DFMA and DMMA interleaved in every warp
//   mode 3  sibling warps: even warps DFMA, odd warps DMMA
// Prints FMA per clock per SM for each stream and their sum.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_concurrency dmma_concurrency.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 8;

__device__ __forceinline__ void dmma884(double &d0, double &d1, const double a, const double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4])
{
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
               "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// FMA counts per warp-level instruction: DFMA 32, DMMA m8n8k4 256, DMMA m16n8k16 2048
template <int MODE, bool BIG>
__global__ void k(double *out, const int iters, long long *cyc)
{
  const int  warp    = threadIdx.x >> 5;
  const bool do_fma  = MODE == 0 || MODE == 2 || (MODE == 3 && (warp & 1) == 0);
  const bool do_mma  = MODE == 1 || MODE == 2 || (MODE == 3 && (warp & 1) == 1);
  double     f[ILP], c2[ILP][2], c4[ILP][4], a8[8], b4[4];
#pragma unroll
  for (int i = 0; i < ILP; ++i)
    {
      f[i] = threadIdx.x * 1e-3 + i;
      c2[i][0] = c2[i][1] = 0.;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        c4[i][j] = 0.;
    }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    a8[j] = 1e-9 * (threadIdx.x + j);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    b4[j] = 1e-9 * (threadIdx.x - j);
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
    {
      if (do_fma)
#pragma unroll
        for (int i = 0; i < ILP; ++i)
          f[i] = fma(f[i], b, c);
      if (do_mma)
#pragma unroll
        for (int i = 0; i < ILP; ++i)
          {
            if (BIG)
              dmma16816(c4[i], a8, b4);
            else
              dmma884(c2[i][0], c2[i][1], a8[0], b4[0]);
          }
    }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i)
    s += f[i] + c2[i][0] + c2[i][1] + c4[i][0] + c4[i][1] + c4[i][2] + c4[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0)
    cyc[warp] = t1 - t0; // per warp: in mode 3 the two kinds of warps finish at different times
}

template <int MODE, bool BIG>
void run(const int warps_per_sm, const char *name)
{
  double    *out;
  long long *cyc, h[32];
  cudaMalloc(&out, 148 * 1024 * 8);
  cudaMalloc(&cyc, 8 * 32);
  const int iters = 2048;
  k<MODE, BIG><<<148, 32 * warps_per_sm>>>(out, iters, cyc);
  k<MODE, BIG><<<148, 32 * warps_per_sm>>>(out, iters, cyc);
  cudaMemcpy(h, cyc, 8 * warps_per_sm, cudaMemcpyDeviceToHost);
  const cudaError_t e = cudaGetLastError();
  const int    w_fma = MODE == 0 || MODE == 2 ? warps_per_sm : (MODE == 3 ? warps_per_sm / 2 : 0);
  const int    w_mma = MODE == 1 || MODE == 2 ? warps_per_sm : (MODE == 3 ? warps_per_sm / 2 : 0);
  // the window in which BOTH streams run is the shorter of the two kinds' durations; rates are
  // quoted over the longest warp of all (everything issued / time until the last warp is done)
  long long t_all = 0;
  for (int w = 0; w < warps_per_sm; ++w)
    t_all = h[w] > t_all ? h[w] : t_all;
  const double fma_per_clk = double(w_fma) * iters * ILP * 32. / double(t_all);
  const double mma_per_clk = double(w_mma) * iters * ILP * (BIG ? 2048. : 256.) / double(t_all);
  printf("%-34s %-9s warps/SM %2d: DFMA %6.1f + DMMA %6.1f = %6.1f FMA/clk/SM (vector peak 64)  %s\n", name,
         BIG ? "m16n8k16" : "m8n8k4", warps_per_sm, fma_per_clk, mma_per_clk, fma_per_clk + mma_per_clk,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(cyc);
}

int main()
{
  for (int w : {4, 8, 16})
    {
      run<0, false>(w, "DFMA only");
      run<1, false>(w, "DMMA only");
      run<1, true>(w, "DMMA only");
      run<2, false>(w, "both, interleaved in each warp");
      run<2, true>(w, "both, interleaved in each warp");
      run<3, false>(w, "both, sibling warps");
      run<3, true>(w, "both, sibling warps");
    }
  return 0;
}
