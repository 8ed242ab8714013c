// Micro-benchmark: how to stream two bf16 tensors (a, dy) from HBM on B200 -- plain 16-byte loads with
// different unroll factors / occupancies versus a 1-D TMA (cp.async.bulk) shared-memory ring.
// Pattern A (reduce): sum a*dy.  Pattern B (apply): dz = f(a, dy) written back (2 reads + 1 write).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_read stream_read.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ float dot8(uint4 a, uint4 b) {
  const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 p = __bfloat1622float2(x[i]), q = __bfloat1622float2(y[i]);
    s = fmaf(p.x, q.x, s);
    s = fmaf(p.y, q.y, s);
  }
  return s;
}
__device__ __forceinline__ uint4 mix8(uint4 a, uint4 b) {
  uint4 r;
  __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(&r);
  const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = __hfma2(x[i], y[i], x[i]);
  return r;
}

template <int U, bool WRITE>
__global__ void __launch_bounds__(256) ldg_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                  uint4* __restrict__ out, float* res, size_t n) {
  const size_t stride = (size_t)gridDim.x * 256;
  size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  float s = 0.f;
  for (; i + (U - 1) * stride < n; i += U * stride) {
    uint4 va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) va[u] = __ldcs(a + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) vb[u] = __ldcs(b + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (WRITE) out[i + u * stride] = mix8(va[u], vb[u]);
      else s += dot8(va[u], vb[u]);
    }
  }
  for (; i < n; i += stride) {
    if (WRITE) out[i] = mix8(a[i], b[i]);
    else s += dot8(a[i], b[i]);
  }
  if (!WRITE) {
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(res, s);
  }
}

// ---------------------------------------------------------------- TMA bulk ring
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t"
      "@P bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

// NT consumer threads + 1 producer warp.  chunk = CH bytes per tensor per stage.
template <int NST, int CH, bool WRITE>
__global__ void __launch_bounds__(288) tma_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                  uint8_t* __restrict__ out, float* res, size_t bytes) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* bufA = smem;
  uint8_t* bufB = smem + (size_t)NST * CH;
  uint8_t* bufO = bufB + (size_t)NST * CH;   // WRITE only: 2 output staging buffers
  uint64_t* full = reinterpret_cast<uint64_t*>(bufO + (WRITE ? 2 * CH : 0));
  uint64_t* empty = full + NST;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t nchunks = bytes / CH;
  if (warp == 8) {
    if ((threadIdx.x & 31) == 0) {
      int st = 0, ph = 0;
      for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], 2 * CH);
        bulk_load(bufA + (size_t)st * CH, a + c * CH, CH, &full[st]);
        bulk_load(bufB + (size_t)st * CH, b + c * CH, CH, &full[st]);
        if (++st == NST) { st = 0; ph ^= 1; }
      }
    }
  } else {
    int st = 0, ph = 0, ob = 0;
    float s = 0.f;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
      mbar_wait(&full[st], ph);
      const uint4* pa = reinterpret_cast<const uint4*>(bufA + (size_t)st * CH);
      const uint4* pb = reinterpret_cast<const uint4*>(bufB + (size_t)st * CH);
      if (WRITE) {
        // staging buffer `ob` must have been read by the bulk store issued two chunks ago
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        uint4* po = reinterpret_cast<uint4*>(bufO + (size_t)ob * CH);
#pragma unroll 4
        for (int k = threadIdx.x; k < CH / 16; k += 256) po[k] = mix8(pa[k], pb[k]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (threadIdx.x == 0) {
          bulk_store(out + c * CH, po, CH);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ob ^= 1;
      } else {
#pragma unroll 4
        for (int k = threadIdx.x; k < CH / 16; k += 256) s += dot8(pa[k], pb[k]);
      }
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[st]);
      if (++st == NST) { st = 0; ph ^= 1; }
    }
    if (WRITE) {
      if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(res, s);
    }
  }
}

template <typename F>
static float time_it(F f, int reps = 10) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  const size_t bytes = (size_t)512 << 20;   // per tensor; 2-3 tensors >> 126 MB L2
  uint8_t *a, *b, *o; float* res;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&o, bytes)); CK(cudaMalloc(&res, 4));
  CK(cudaMemset(a, 0x3c, bytes)); CK(cudaMemset(b, 0x3c, bytes)); CK(cudaMemset(o, 0, bytes));
  const size_t n = bytes / 16;
  printf("pattern,variant,blocks_per_sm,ms,GB/s\n");
#define RUN_LDG(U, W, BPS) { float ms = time_it([&] { ldg_kernel<U, W><<<148 * BPS, 256>>>((const uint4*)a, (const uint4*)b, (uint4*)o, res, n); }); \
    printf("%s,ldg_u%d,%d,%.4f,%.0f\n", W ? "apply" : "reduce", U, BPS, ms, (W ? 3.0 : 2.0) * bytes / ms / 1e6); }
  RUN_LDG(1, false, 4) RUN_LDG(1, false, 8) RUN_LDG(2, false, 4) RUN_LDG(2, false, 8) RUN_LDG(4, false, 2) RUN_LDG(4, false, 4)
  RUN_LDG(4, false, 8) RUN_LDG(8, false, 2) RUN_LDG(8, false, 4)
  RUN_LDG(1, true, 4) RUN_LDG(1, true, 8) RUN_LDG(2, true, 4) RUN_LDG(2, true, 8) RUN_LDG(4, true, 2) RUN_LDG(4, true, 4)
  RUN_LDG(4, true, 8) RUN_LDG(8, true, 2) RUN_LDG(8, true, 4)
#define RUN_TMA(NST, CH, W, BPS) { size_t sm = (size_t)2 * NST * CH + (W ? 2 * CH : 0) + 16 * NST + 64; \
    CK(cudaFuncSetAttribute(tma_kernel<NST, CH, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    float ms = time_it([&] { tma_kernel<NST, CH, W><<<148 * BPS, 288, sm>>>(a, b, o, res, bytes); }); \
    CK(cudaGetLastError()); \
    printf("%s,tma_st%d_ch%d,%d,%.4f,%.0f\n", W ? "apply" : "reduce", NST, CH, BPS, ms, (W ? 3.0 : 2.0) * bytes / ms / 1e6); }
  RUN_TMA(4, 8192, false, 1) RUN_TMA(4, 16384, false, 1) RUN_TMA(6, 16384, false, 1) RUN_TMA(3, 32768, false, 1)
  RUN_TMA(4, 8192, false, 2) RUN_TMA(3, 16384, false, 2) RUN_TMA(4, 4096, false, 4)
  RUN_TMA(4, 8192, true, 1) RUN_TMA(4, 16384, true, 1) RUN_TMA(3, 16384, true, 2) RUN_TMA(4, 8192, true, 2) RUN_TMA(4, 4096, true, 4)
  // pure copy reference (cudaMemcpyAsync d2d)
  { float ms = time_it([&] { cudaMemcpyAsync(o, a, bytes, cudaMemcpyDeviceToDevice); }); printf("copy,memcpy_d2d,0,%.4f,%.0f\n", ms, 2.0 * bytes / ms / 1e6); }
  return 0;
}
