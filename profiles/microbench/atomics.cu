// Micro-benchmark: cost of the "every block adds its per-channel partials to one global array" epilogue used
// by the reduction passes (BN statistics, bias gradients, loss).  nblocks x 256 threads, each block issues C
// atomicAdds (double or float) to C consecutive addresses; optional stripes spread blocks over S copies.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics atomics.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

template <typename T>
__global__ void __launch_bounds__(256) atom_kernel(T* acc, int C, int S) {
  T* dst = acc + (size_t)(blockIdx.x % S) * C;
  for (int k = threadIdx.x; k < C; k += 256) atomicAdd(&dst[k], (T)1);
}
__global__ void empty_kernel() {}
__global__ void __launch_bounds__(256) touch_kernel(float* p, int n) {
  int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) p[i] += 1.f;
}

template <typename F>
static float time_it(F f, int reps = 20) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps * 1e3f;
}

int main() {
  void* acc; CK(cudaMalloc(&acc, 64 << 20)); CK(cudaMemset(acc, 0, 64 << 20));
  printf("kind,nblocks,C,stripes,us_per_launch\n");
  printf("empty,1,0,0,%.2f\n", time_it([&] { empty_kernel<<<1, 32>>>(); }));
  printf("empty,1184,0,0,%.2f\n", time_it([&] { empty_kernel<<<1184, 256>>>(); }));
  printf("touch,2048,0,0,%.2f\n", time_it([&] { touch_kernel<<<2048, 256>>>((float*)acc, 2048 * 256); }));
  for (int nb : {148, 296, 592, 1184, 2368})
    for (int C : {32, 64, 128, 512, 1024, 2048})
      for (int S : {1, 16}) {
        printf("double,%d,%d,%d,%.2f\n", nb, C, S, time_it([&] { atom_kernel<double><<<nb, 256>>>((double*)acc, C, S); }));
        printf("float,%d,%d,%d,%.2f\n", nb, C, S, time_it([&] { atom_kernel<float><<<nb, 256>>>((float*)acc, C, S); }));
      }
  return 0;
}
