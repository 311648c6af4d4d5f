// Latency of rss_rect_dist for ONE warp (the wide kernel's situation: a query alone on its warp), as a function of how the
// 32 lanes' inputs differ: (a) every lane the same rectangle pair, (b) every lane its own pair, (c) as (b) with only the
// even / only 16 / only 8 lanes active.  Inputs: random rotations, rectangles of sides U(1,5), offsets of about one side.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I c2a_b200/csrc -o gpurun_out/rect_latency scripts/micro/rect_latency.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include <random>
#include <cuda_runtime.h>
#include "c2a_geom.cuh"
using namespace c2a;

__global__ void bench(const double *cases, int n_cases, int mode, unsigned active, double *out, long long *cycles, int *trips_out)
{
  const int lane = threadIdx.x;
  double acc = 0;
  long long t0 = 0, t1 = 0;
  __syncwarp();
  if ((active >> lane) & 1u)
  {
    t0 = clock64();
    for (int i = 0; i < n_cases; i++)
    {
      const double *c = cases + (size_t)18 * ((mode == 0) ? i : ((i * 32 + lane) % (n_cases * 32)));
      double R[9], T[3], S[3] = {0, 0, 0};
#pragma unroll
      for (int k = 0; k < 9; k++) R[k] = c[k];
      T[0] = c[9]; T[1] = c[10]; T[2] = c[11];
      acc += rss_rect_dist(R, T, c[12], c[13], c[14], c[15], S) + S[0];
    }
    t1 = clock64();
  }
  __syncwarp();
  out[lane] = acc;
  if (lane == __ffs(active) - 1) *cycles = t1 - t0;
}

int main()
{
  const int N = 2000;
  std::mt19937_64 rng(7);
  std::normal_distribution<double> g(0, 1);
  std::uniform_real_distribution<double> u(1, 5);
  std::vector<double> h((size_t)18 * N * 32);
  for (int i = 0; i < N * 32; i++)
  {
    double q[4]; double n = 0;
    for (double &x : q) { x = g(rng); n += x * x; }
    n = std::sqrt(n); for (double &x : q) x /= n;
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    double *c = &h[(size_t)18 * i];
    const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
                         2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
    for (int k = 0; k < 9; k++) c[k] = R[k];
    for (int k = 0; k < 3; k++) c[9 + k] = 3.0 * g(rng);
    for (int k = 0; k < 4; k++) c[12 + k] = u(rng);
  }
  double *d_cases, *d_out; long long *d_cyc, cyc; int *d_tr;
  cudaMalloc(&d_cases, h.size() * 8); cudaMalloc(&d_out, 32 * 8); cudaMalloc(&d_cyc, 8); cudaMalloc(&d_tr, 4);
  cudaMemcpy(d_cases, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  struct { const char *name; int mode; unsigned active; } runs[] = {
      {"all 32 lanes, the same pair", 0, 0xffffffffu}, {"all 32 lanes, different pairs", 1, 0xffffffffu},
      {"16 lanes (even), different pairs", 1, 0x55555555u}, {"8 lanes, different pairs", 1, 0x11111111u},
      {"2 lanes, different pairs", 1, 0x00010001u}, {"1 lane", 1, 0x1u}};
  for (auto &r : runs)
    for (int rep = 0; rep < 2; rep++)
    {
      bench<<<1, 32>>>(d_cases, N, r.mode, r.active, d_out, d_cyc, d_tr);
      cudaDeviceSynchronize();
      cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
      if (rep) printf("%-36s %8.0f cycles per call\n", r.name, (double)cyc / N);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
