// fp64 latency probe (development): dependent DFMA / DMUL / rsqrt / reciprocal chains, one warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x0) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.000000001, z = 1e-9;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int j = 0; j < 16; j++) x = fma(x, y, z);
  }
  long long t1 = clock64();
  double a = x;
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) a = rsqrt(a) + 1.5;
  }
  long long t2 = clock64();
  double b = a;
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) b = 1.0 / b + 0.5;
  }
  long long t3 = clock64();
  double c = b;
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int j = 0; j < 4; j++) c = sqrt(c) + 1.5;
  }
  long long t4 = clock64();
  // 8 independent chains (ILP)
  double e[8];
  for (int j = 0; j < 8; j++) e[j] = c + j;
#pragma unroll 1
  for (int i = 0; i < 64; i++) {
#pragma unroll
    for (int j = 0; j < 16; j++) e[j & 7] = fma(e[j & 7], y, z);
  }
  long long t5 = clock64();
  double s = 0; for (int j = 0; j < 8; j++) s += e[j];
  out[threadIdx.x + blockIdx.x * blockDim.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; }
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 1 << 20); cudaMalloc(&c, 64);
  for (int threads : {32, 128, 256, 512}) {
    k<<<1, threads>>>(o, c, 1.0); cudaDeviceSynchronize();
    k<<<1, threads>>>(o, c, 1.0); cudaDeviceSynchronize();
    long long h[5]; cudaMemcpy(h, c, 40, cudaMemcpyDeviceToHost);
    printf("threads %d: dep DFMA %.1f cyc, rsqrt+add %.1f, rcp+add %.1f, sqrt+add %.1f, 8-chain DFMA %.2f cyc/op\n", threads, h[0] / 1024.0, h[1] / 256.0, h[2] / 256.0,
           h[3] / 256.0, h[4] / 1024.0);
  }
  return 0;
}
