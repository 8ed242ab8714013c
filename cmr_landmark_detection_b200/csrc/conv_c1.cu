// Cin == 1 first layer (enc0.conv_a), forward and weight gradient, on CUDA cores.
// K = 9: arithmetic intensity 8.7 FLOP/B -- an HBM streaming problem (write / read 64 B per pixel), tensor cores
// have nothing to chew on.  A thread owns 8 output channels (its 72 weights live in registers) of FOUR
// consecutive pixels of a row, so the 3 x 6 input patch is loaded once for 288 FMAs and all index math is
// 32-bit; G = Cout/8 adjacent threads share the pixel quad, so stores are 16 B (bf16) and fully coalesced.
// Replaces tf.keras Conv2D(1 -> FILTERS) + ReLU and its Conv2DBackpropFilter (src/models/KerasLayers.py:689).
#include "kernels.cuh"

namespace rvip {

constexpr int kQuad = 4;

// loads the 3 x (kQuad + 2) input patch around pixels (yy, x0 .. x0+3) with zero padding
__device__ __forceinline__ void load_patch(const float* __restrict__ img, int H, int W, int yy, int x0,
                                           float (&v)[3][kQuad + 2]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int y2 = yy + r - 1;
    const bool rowok = (unsigned)y2 < (unsigned)H;
    const float* row = img + (size_t)(rowok ? y2 : 0) * W;
#pragma unroll
    for (int k = 0; k < kQuad + 2; ++k) {
      const int x2 = x0 + k - 1;
      v[r][k] = (rowok && (unsigned)x2 < (unsigned)W) ? __ldg(row + x2) : 0.f;
    }
  }
}

template <typename Tout>
__global__ void __launch_bounds__(256, 2) conv3x3_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, Tout* __restrict__ out,
                                                             double* __restrict__ stats, int B, int H, int W, int Cout,
                                                             int want_stats, const float* __restrict__ scale,
                                                             const float* __restrict__ shift) {
  extern __shared__ float red_s[];  // [2][Cout]
  pdl_wait();
  const uint32_t G = Cout >> 3, lg = 31 - __clz(G);
  const uint32_t Wq = W / kQuad;
  const uint32_t n_items = ((uint32_t)B * H * Wq) << lg;
  for (int k = threadIdx.x; k < 2 * Cout; k += 256) red_s[k] = 0.f;
  __syncthreads();
  const uint32_t i0 = blockIdx.x * 256 + threadIdx.x;
  if ((i0 & ~31u) < n_items) {
    const int c = (int)(i0 & (G - 1)) * 8;
    float wr[9][8], br[8], s[8], q[8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) wr[t][j] = w[t * Cout + c + j];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      br[j] = bias[c + j];
      s[j] = q[j] = 0.f;
    }
    for (uint32_t i = i0; i < n_items; i += gridDim.x * 256) {
      const uint32_t quad = i >> lg;
      const uint32_t xq = quad % Wq, t2 = quad / Wq;
      const uint32_t yy = t2 % H, b = t2 / H;
      const int x0 = (int)xq * kQuad;
      float v[3][kQuad + 2];
      load_patch(x + (size_t)b * H * W, H, W, (int)yy, x0, v);
      Tout* dst = out + ((size_t)(b * H + yy) * W + x0) * Cout + c;
#pragma unroll
      for (int p = 0; p < kQuad; ++p) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = br[j];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float xin = v[t / 3][p + t % 3];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(xin, wr[t][j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[j] = fmaxf(acc[j], 0.f);
          s[j] += acc[j];
          q[j] = fmaf(acc[j], acc[j], q[j]);
        }
        if (scale) {   // inference: BatchNorm (moving statistics) folded into the epilogue
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(acc[j], __ldg(scale + c + j), __ldg(shift + c + j));
        }
        Vec8<Tout>::store(dst + (size_t)p * Cout, acc);
      }
    }
    pdl_launch_dependents();
    if (want_stats) {
      block_accumulate8(red_s, c, s, G);
      block_accumulate8(red_s + Cout, c, q, G);
    }
  }
  if (want_stats) {
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * Cout; k += 256) atomicAdd(&stats[k], (double)red_s[k]);
  }
}

// dW[tap][0][co] = sum_p x[p + off(tap)] * dz[p][co]
template <typename Tdz>
__global__ void __launch_bounds__(256, 2) wgrad3x3_c1_kernel(const float* __restrict__ x, const Tdz* __restrict__ dz,
                                                          float* __restrict__ dw, int B, int H, int W, int Cout) {
  extern __shared__ float red_s[];  // [9][Cout]
  pdl_wait();
  const uint32_t G = Cout >> 3, lg = 31 - __clz(G);
  const uint32_t Wq = W / kQuad;
  const uint32_t n_items = ((uint32_t)B * H * Wq) << lg;
  for (int k = threadIdx.x; k < 9 * Cout; k += 256) red_s[k] = 0.f;
  __syncthreads();
  const uint32_t i0 = blockIdx.x * 256 + threadIdx.x;
  if ((i0 & ~31u) < n_items) {
    const int c = (int)(i0 & (G - 1)) * 8;
    float acc[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;
    for (uint32_t i = i0; i < n_items; i += gridDim.x * 256) {
      const uint32_t quad = i >> lg;
      const uint32_t xq = quad % Wq, t2 = quad / Wq;
      const uint32_t yy = t2 % H, b = t2 / H;
      const int x0 = (int)xq * kQuad;
      float v[3][kQuad + 2];
      load_patch(x + (size_t)b * H * W, H, W, (int)yy, x0, v);
      const Tdz* src = dz + ((size_t)(b * H + yy) * W + x0) * Cout + c;
#pragma unroll
      for (int p = 0; p < kQuad; ++p) {
        float g[8];
        Vec8<Tdz>::load(src + (size_t)p * Cout, g);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float xin = v[t / 3][p + t % 3];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[t][j] = fmaf(xin, g[j], acc[t][j]);
        }
      }
    }
    pdl_launch_dependents();
#pragma unroll
    for (int t = 0; t < 9; ++t) block_accumulate8(red_s + t * Cout, c, acc[t], G);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 9 * Cout; k += 256) atomicAdd(&dw[k], red_s[k]);
}

static int c1_grid(size_t n_items) {
  size_t g = (n_items + 255) / 256;
  const size_t cap = (size_t)kNumSMs * 4;
  return (int)(g < cap ? (g ? g : 1) : cap);
}
static int c1_check(int B, int H, int W, int Cout) {
  const int G = Cout / 8;
  RVIP_REQUIRE(Cout % 8 == 0 && G <= 32 && (G & (G - 1)) == 0, "conv_c1: Cout=%d must be 8 * power of two <= 256", Cout);
  RVIP_REQUIRE(W % kQuad == 0, "conv_c1: W=%d must be a multiple of %d", W, kQuad);
  RVIP_REQUIRE((size_t)B * H * W * G < 0x7fffffffULL, "conv_c1: tensor too large for 32-bit indexing");
  return 0;
}

int conv_c1_fwd_launch(const float* x, const float* w, const float* bias, void* out, double* stats, int B, int H, int W,
                       int Cout, int want_stats, int out_is_bf16, const float* scale, const float* shift,
                       cudaStream_t st) {
  if (c1_check(B, H, W, Cout)) return 1;
  const int grid = c1_grid((size_t)B * H * (W / kQuad) * (Cout / 8));
  if (out_is_bf16)
    launch_kernel(conv3x3_c1_fwd_kernel<__nv_bfloat16>, grid, 256, 2 * Cout * sizeof(float), st, x, w, bias,
                  static_cast<__nv_bfloat16*>(out), stats, B, H, W, Cout, want_stats, scale, shift);
  else
    launch_kernel(conv3x3_c1_fwd_kernel<float>, grid, 256, 2 * Cout * sizeof(float), st, x, w, bias,
                  static_cast<float*>(out), stats, B, H, W, Cout, want_stats, scale, shift);
  RVIP_LAUNCH_CHECK();
  return 0;
}

int wgrad_c1_launch(const float* x, const void* dz, float* dw, int B, int H, int W, int Cout, int dz_is_bf16,
                    cudaStream_t st) {
  if (c1_check(B, H, W, Cout)) return 1;
  const int grid = c1_grid((size_t)B * H * (W / kQuad) * (Cout / 8));
  if (dz_is_bf16)
    launch_kernel(wgrad3x3_c1_kernel<__nv_bfloat16>, grid, 256, 9 * Cout * sizeof(float), st, x,
                  static_cast<const __nv_bfloat16*>(dz), dw, B, H, W, Cout);
  else
    launch_kernel(wgrad3x3_c1_kernel<float>, grid, 256, 9 * Cout * sizeof(float), st, x, static_cast<const float*>(dz),
                  dw, B, H, W, Cout);
  RVIP_LAUNCH_CHECK();
  return 0;
}

}  // namespace rvip
