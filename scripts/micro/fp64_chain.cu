// FP64 pipe behaviour on one SM: dependent DMUL+DADD chains, W warps per block (1 block per SM), C chains per thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o gpurun_out/fp64_chain scripts/micro/fp64_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void chain(double *out, int iters, long long *cycles)
{
  double x[C];
#pragma unroll
  for (int c = 0; c < C; c++) x[c] = 1.0 + threadIdx.x * 1e-9 + c * 1e-7;
  const double a = 1.0000001, b = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int k = 0; k < 8; k++)
#pragma unroll
      for (int c = 0; c < C; c++) x[c] = x[c] * a + b;   // DMUL then dependent DADD (-fmad=false)
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; c++) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main()
{
  double *out; long long *cyc, h;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  printf("warps/SM chains/thread cycles/dependent-instr  warp-instr/cycle/SM\n");
  for (int C : {1, 2, 4})
    for (int W : {1, 2, 4, 8, 16, 32})
    {
      for (int rep = 0; rep < 2; rep++)
      {
        if (C == 1) chain<1><<<148, 32 * W>>>(out, iters, cyc);
        if (C == 2) chain<2><<<148, 32 * W>>>(out, iters, cyc);
        if (C == 4) chain<4><<<148, 32 * W>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      const double ninstr = (double)iters * 8 * 2;  // per chain
      printf("%6d %6d %10.2f %10.3f\n", W, C, h / ninstr, ninstr * C * W / h);
    }
  return 0;
}
