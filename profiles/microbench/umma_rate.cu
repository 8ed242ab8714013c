// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, bf16 in / fp32 out, cta_group::1, M = 128) with both
// operands in shared memory, as a function of N and of operand major-ness.  One CTA per SM, no loads: the
// operands are whatever the shared memory holds.  Answers: how many cycles does a 128 x N x 16 MMA cost when
// the A tile (4 KB) and B tile (N*32 B) must be re-read from shared memory for every instruction?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../cmr_landmark_detection_b200/csrc -o umma_rate umma_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "tc_prims.cuh"
using namespace rvip::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)

// MODE 0: K-major A and B, SW128 (64 bf16 per row), K advance = +32 B inside the swizzle atom
// MODE 1: MN-major A and B, SW128: 64-element M/N chunks, K advance = 16 rows of 128 B
// DISTINCT: number of distinct A tiles cycled through (1 = same operands every MMA)
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) umma_kernel(int iters, int distinct, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 40 * 1024; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 1 && elect_one()) {   // elect form: straight-line UTCHMMA issue (lane == 0 costs ~72 cycles per MMA)
    constexpr uint32_t idesc = make_idesc_bf16(128, N, MODE, MODE);
    const uint32_t a_base = smem_u32(smem);                 // A tiles: 16 KB each (128 rows x 128 B)
    const uint32_t b_base = a_base + 96 * 1024;             // B tiles: up to 32 KB (256 rows x 128 B)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t ao = (uint32_t)(it % distinct) * 16384u;
      uint64_t adesc, bdesc;
      if (MODE == 0) {
        adesc = make_smem_desc(a_base + ao, 16, 1024, kLayoutSW128);
        bdesc = make_smem_desc(b_base, 16, 1024, kLayoutSW128);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_bf16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
      } else {
        // MN-major: 64-element chunks along M/N are `LBO` apart (one 64-row x 128 B block each = 8 KB),
        // 8-row K groups are 1024 B apart
        adesc = make_smem_desc(a_base + ao, 8192, 1024, kLayoutSW128);
        bdesc = make_smem_desc(b_base, 8192, 1024, kLayoutSW128);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_bf16_ss(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, 1);
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int N, int MODE>
static void run(int distinct, long long* d_out) {
  const int iters = 2000;
  const size_t smem = 1024 + 96 * 1024 + 64 * 1024;
  CK(cudaFuncSetAttribute(umma_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_kernel<N, MODE><<<148, 128, smem>>>(iters, distinct, d_out);
  CK(cudaDeviceSynchronize());
  long long cyc;
  CK(cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost));
  const double per = (double)cyc / (iters * 4.0);
  printf("%s,%d,%d,%.1f,%.1f,%.0f\n", MODE ? "mn_major" : "k_major", N, distinct, per, 128.0 * N / 256.0,
         (128.0 + N) * 32.0 / per);
}

int main() {
  long long* d_out;
  CK(cudaMalloc(&d_out, 64));
  printf("layout,N,distinct_A_tiles,cycles_per_mma,ideal_cycles,smem_operand_bytes_per_cycle\n");
  for (int distinct : {1, 4}) {
    run<32, 0>(distinct, d_out); run<64, 0>(distinct, d_out); run<128, 0>(distinct, d_out); run<256, 0>(distinct, d_out);
    run<32, 1>(distinct, d_out); run<64, 1>(distinct, d_out); run<128, 1>(distinct, d_out); run<256, 1>(distinct, d_out);
  }
  return 0;
}
