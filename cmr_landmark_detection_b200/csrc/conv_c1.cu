// Cin == 1 first layer (enc0.conv_a), forward and weight gradient, on CUDA cores.
// K = 9: arithmetic intensity 8.7 FLOP/B -- an HBM streaming problem (write / read 64 B per pixel), tensor cores
// have nothing to chew on.  A thread owns 8 output channels (its 72 weights live in registers) of FOUR
// consecutive pixels of a row, so the 3 x 6 input patch is loaded once for 288 FMAs and all index math is
// 32-bit; G = Cout/8 adjacent threads share the pixel quad, so stores are 16 B (bf16) and fully coalesced.
// Replaces tf.keras Conv2D(1 -> FILTERS) + ReLU and its Conv2DBackpropFilter (src/models/KerasLayers.py:689).
#include "kernels.cuh"

#include <stdlib.h>

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

// ---- second mapping: 4 output channels x 8 consecutive pixels per thread.
// ncu on the 8-channel x 4-pixel kernels above: issue slots 61 % busy at 25 % occupancy (128 registers), ~2x the
// instructions the arithmetic needs -- the 72 weights + 18 patch values + accumulators do not fit 128 registers, so the
// weights are re-read from L1 every iteration.  With 4 channels a thread keeps 36 weights, a 3 x 10 patch and its
// accumulators in ~90 registers: nothing is reloaded and three blocks per SM are resident.
constexpr int kOct = 8;
template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&o)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&o)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&o)[4]) {
  uint2 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
  h[0] = __floats2bfloat162_rn(o[0], o[1]);
  h[1] = __floats2bfloat162_rn(o[2], o[3]);
  *reinterpret_cast<uint2*>(p) = raw;
}
template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&o)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&o)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&o)[4]) {
  const uint2 raw = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
  const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
// 3 x (kOct + 2) input patch around pixels (yy, x0 .. x0+7), zero padding; x0 is a multiple of 8, so the eight centre
// values of a row are two aligned 16-byte loads
__device__ __forceinline__ void load_patch8(const float* __restrict__ img, int H, int W, int yy, int x0,
                                            float (&v)[3][kOct + 2]) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int y2 = yy + r - 1;
    if ((unsigned)y2 < (unsigned)H) {
      const float* row = img + (size_t)y2 * W + x0;
      const float4 a = __ldg(reinterpret_cast<const float4*>(row));
      const float4 b = __ldg(reinterpret_cast<const float4*>(row + 4));
      v[r][0] = x0 > 0 ? __ldg(row - 1) : 0.f;
      v[r][1] = a.x; v[r][2] = a.y; v[r][3] = a.z; v[r][4] = a.w;
      v[r][5] = b.x; v[r][6] = b.y; v[r][7] = b.z; v[r][8] = b.w;
      v[r][9] = x0 + kOct < W ? __ldg(row + kOct) : 0.f;
    } else {
#pragma unroll
      for (int k = 0; k < kOct + 2; ++k) v[r][k] = 0.f;
    }
  }
}
// G4 = Cout / 4 lanes share a pixel octet; sums of a channel are folded over the lanes that hold the same channels
__device__ __forceinline__ void block_accumulate4(float* smem_acc, int c, const float (&v)[4], uint32_t G4) {
  float r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) r[j] = v[j];
  for (uint32_t o = 16; o >= G4 && o > 0; o >>= 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) r[j] += __shfl_xor_sync(0xffffffffu, r[j], o);
  }
  if ((threadIdx.x & 31) < G4 || G4 >= 32) {
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&smem_acc[c + j], r[j]);
  }
}

template <typename Tout>
__global__ void __launch_bounds__(256, 2) conv3x3_c1_fwd8_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, Tout* __restrict__ out,
                                                              double* __restrict__ stats, int B, int H, int W, int Cout,
                                                              int want_stats, const float* __restrict__ scale,
                                                              const float* __restrict__ shift) {
  extern __shared__ float red_s[];  // [2][Cout]
  pdl_wait();
  const uint32_t G = Cout >> 2, lg = 31 - __clz(G);
  const uint32_t Wo = W / kOct;
  const uint32_t n_items = ((uint32_t)B * H * Wo) << lg;
  for (int k = threadIdx.x; k < 2 * Cout; k += 256) red_s[k] = 0.f;
  __syncthreads();
  const uint32_t i0 = blockIdx.x * 256 + threadIdx.x;
  if ((i0 & ~31u) < n_items) {
    const int c = (int)(i0 & (G - 1)) * 4;
    // channel pairs ride in packed fp32 registers: 18 FFMA2 per pixel instead of 36 FFMA (the kernel is issue bound)
    f32x2_t wr2[9][2], br2[2], s2[2], q2[2];
    float sc4[4], sh4[4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int h = 0; h < 2; ++h) wr2[t][h] = f2_pack(w[t * Cout + c + 2 * h], w[t * Cout + c + 2 * h + 1]);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      br2[h] = f2_pack(bias[c + 2 * h], bias[c + 2 * h + 1]);
      s2[h] = q2[h] = f2_pack(0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sc4[j] = scale ? scale[c + j] : 1.f;
      sh4[j] = scale ? shift[c + j] : 0.f;
    }
    for (uint32_t i = i0; i < n_items; i += gridDim.x * 256) {
      const uint32_t oct = i >> lg;
      const uint32_t xo = oct % Wo, t2 = oct / Wo;
      const uint32_t yy = t2 % H, b = t2 / H;
      const int x0 = (int)xo * kOct;
      float v[3][kOct + 2];
      load_patch8(x + (size_t)b * H * W, H, W, (int)yy, x0, v);
      Tout* dst = out + ((size_t)(b * H + yy) * W + x0) * Cout + c;
#pragma unroll
      for (int p = 0; p < kOct; ++p) {
        f32x2_t acc2[2] = {br2[0], br2[1]};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float xin = v[t / 3][p + t % 3];
          const f32x2_t xb = f2_pack(xin, xin);
#pragma unroll
          for (int h = 0; h < 2; ++h) acc2[h] = f2_fma(xb, wr2[t][h], acc2[h]);
        }
        float acc[4];
        f2_unpack(acc2[0], acc[0], acc[1]);
        f2_unpack(acc2[1], acc[2], acc[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaxf(acc[j], 0.f);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const f32x2_t r2 = f2_pack(acc[2 * h], acc[2 * h + 1]);
          s2[h] = f2_add(s2[h], r2);
          q2[h] = f2_fma(r2, r2, q2[h]);
        }
        if (scale) {   // inference: BatchNorm (moving statistics) folded into the epilogue
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] = fmaf(acc[j], sc4[j], sh4[j]);
        }
        store4<Tout>(dst + (size_t)p * Cout, acc);
      }
    }
    float s[4], q[4];
    f2_unpack(s2[0], s[0], s[1]);
    f2_unpack(s2[1], s[2], s[3]);
    f2_unpack(q2[0], q[0], q[1]);
    f2_unpack(q2[1], q[2], q[3]);
    pdl_launch_dependents();
    if (want_stats) {
      block_accumulate4(red_s, c, s, G);
      block_accumulate4(red_s + Cout, c, q, G);
    }
  }
  if (want_stats) {
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * Cout; k += 256) atomicAdd(&stats[k], (double)red_s[k]);
  }
}

template <typename Tdz>
__global__ void __launch_bounds__(256, 2) wgrad3x3_c1_8_kernel(const float* __restrict__ x, const Tdz* __restrict__ dz,
                                                            float* __restrict__ dw, int B, int H, int W, int Cout) {
  extern __shared__ float red_s[];  // [9][Cout]
  pdl_wait();
  const uint32_t G = Cout >> 2, lg = 31 - __clz(G);
  const uint32_t Wo = W / kOct;
  const uint32_t n_items = ((uint32_t)B * H * Wo) << lg;
  for (int k = threadIdx.x; k < 9 * Cout; k += 256) red_s[k] = 0.f;
  __syncthreads();
  const uint32_t i0 = blockIdx.x * 256 + threadIdx.x;
  if ((i0 & ~31u) < n_items) {
    const int c = (int)(i0 & (G - 1)) * 4;
    f32x2_t acc2[9][2];      // channel pairs in packed fp32 registers: 144 FFMA2 per pixel octet instead of 288 FFMA
#pragma unroll
    for (int t = 0; t < 9; ++t) acc2[t][0] = acc2[t][1] = f2_pack(0.f, 0.f);
    for (uint32_t i = i0; i < n_items; i += gridDim.x * 256) {
      const uint32_t oct = i >> lg;
      const uint32_t xo = oct % Wo, t2 = oct / Wo;
      const uint32_t yy = t2 % H, b = t2 / H;
      const int x0 = (int)xo * kOct;
      float v[3][kOct + 2];
      load_patch8(x + (size_t)b * H * W, H, W, (int)yy, x0, v);
      const Tdz* src = dz + ((size_t)(b * H + yy) * W + x0) * Cout + c;
      float g[kOct][4];
#pragma unroll
      for (int p = 0; p < kOct; ++p) load4<Tdz>(src + (size_t)p * Cout, g[p]);
#pragma unroll
      for (int p = 0; p < kOct; ++p) {
        const f32x2_t g2[2] = {f2_pack(g[p][0], g[p][1]), f2_pack(g[p][2], g[p][3])};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float xin = v[t / 3][p + t % 3];
          const f32x2_t xb = f2_pack(xin, xin);
#pragma unroll
          for (int h = 0; h < 2; ++h) acc2[t][h] = f2_fma(xb, g2[h], acc2[t][h]);
        }
      }
    }
    float acc[9][4];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      f2_unpack(acc2[t][0], acc[t][0], acc[t][1]);
      f2_unpack(acc2[t][1], acc[t][2], acc[t][3]);
    }
    pdl_launch_dependents();
#pragma unroll
    for (int t = 0; t < 9; ++t) block_accumulate4(red_s + t * Cout, c, acc[t], G);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 9 * Cout; k += 256) atomicAdd(&dw[k], red_s[k]);
}

static bool c1_use_oct(int W, int Cout) {
  const bool off = getenv("RVIP_C1_QUAD") != nullptr;   // read per launch: the parity tests run both mappings
  const int G4 = Cout / 4;
  return !off && W % kOct == 0 && Cout % 4 == 0 && G4 <= 32 && (G4 & (G4 - 1)) == 0;
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
  if (c1_use_oct(W, Cout)) {
    const int grid8 = c1_grid((size_t)B * H * (W / kOct) * (Cout / 4));
    if (out_is_bf16)
      launch_kernel(conv3x3_c1_fwd8_kernel<__nv_bfloat16>, grid8, 256, 2 * Cout * sizeof(float), st, x, w, bias,
                    static_cast<__nv_bfloat16*>(out), stats, B, H, W, Cout, want_stats, scale, shift);
    else
      launch_kernel(conv3x3_c1_fwd8_kernel<float>, grid8, 256, 2 * Cout * sizeof(float), st, x, w, bias,
                    static_cast<float*>(out), stats, B, H, W, Cout, want_stats, scale, shift);
    RVIP_LAUNCH_CHECK();
    return 0;
  }
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
  if (c1_use_oct(W, Cout)) {
    const int grid8 = c1_grid((size_t)B * H * (W / kOct) * (Cout / 4));
    if (dz_is_bf16)
      launch_kernel(wgrad3x3_c1_8_kernel<__nv_bfloat16>, grid8, 256, 9 * Cout * sizeof(float), st, x,
                    static_cast<const __nv_bfloat16*>(dz), dw, B, H, W, Cout);
    else
      launch_kernel(wgrad3x3_c1_8_kernel<float>, grid8, 256, 9 * Cout * sizeof(float), st, x, static_cast<const float*>(dz),
                    dw, B, H, W, Cout);
    RVIP_LAUNCH_CHECK();
    return 0;
  }
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
