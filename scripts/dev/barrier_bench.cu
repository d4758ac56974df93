#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;
__global__ void k_cg(int n, long long* out) {
  cg::grid_group g = cg::this_grid();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { __threadfence(); g.sync(); }
  if (blockIdx.x == 0 && threadIdx.x == 0) *out = clock64() - t0;
}
__device__ __forceinline__ void my_barrier(unsigned* ctr, unsigned& epoch, unsigned nb) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += nb;
    __threadfence();
    atomicAdd(ctr, 1u);
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(ctr));
      if ((int)(v - epoch) >= 0) break;
    }
  }
  __syncthreads();
}
__global__ void k_my(int n, unsigned* ctr, long long* out) {
  unsigned epoch = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) my_barrier(ctr, epoch, gridDim.x);
  if (blockIdx.x == 0 && threadIdx.x == 0) *out = clock64() - t0;
}
int main() {
  long long* d; unsigned* c; cudaMalloc(&d, 8); cudaMalloc(&c, 4);
  int n = 2000; long long h;
  for (int grid : {148, 296}) for (int th : {256}) {
    void* a1[] = {&n, &d};
    cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(th), a1, 0, 0); cudaDeviceSynchronize();
    cudaLaunchCooperativeKernel((void*)k_cg, dim3(grid), dim3(th), a1, 0, 0); cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("cg grid.sync grid=%d: %.0f cycles/barrier (%s)\n", grid, (double)h / n, cudaGetErrorString(cudaGetLastError()));
    cudaMemset(c, 0, 4);
    void* a2[] = {&n, &c, &d};
    cudaLaunchCooperativeKernel((void*)k_my, dim3(grid), dim3(th), a2, 0, 0); cudaDeviceSynchronize();
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("atomic barrier grid=%d: %.0f cycles/barrier (%s)\n", grid, (double)h / n, cudaGetErrorString(cudaGetLastError()));
  }
}
