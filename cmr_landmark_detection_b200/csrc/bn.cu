// Memory-bound passes around the convolutions: BatchNorm finalize / apply (fused with Dropout,
// MaxPool 2x2 or nearest UpSampling x2) and the fused BatchNorm+ReLU backward (fused with the
// dropout-mask replay, max-pool gradient routing or up-sampling 2x2 gradient sum).
// Reference ops replaced (all TF kernels reached from src/models/KerasLayers.py):
//   BatchNormalization(axis=-1) :684,691   Dropout :718,772 (Unets.py:813)   MaxPooling2D :714,721
//   UpSampling2D :756-757   and their gradients.  Block order is Conv -> ReLU -> BN (BN_FIRST false).
// Every thread moves 8 channels (16 B bf16 / 32 B fp32) of one pixel: fully coalesced NHWC traffic.
// These passes were issue-bound before they were HBM-bound, so the per-element instruction count is
// kept minimal: 32-bit index math with shifts (C/8 is a power of two), a 7-op counter hash for the
// dropout mask, and the backward formula folded into two FMAs per element:
//   dz = [a>0] * ( sc*dy - k1*a + c0 ),  sc = gamma*rstd, k1 = sc*rstd*mean(dy*ahat), c0 = k1*mu - sc*mean(dy)
#include "kernels.cuh"

#include <stdlib.h>

namespace rvip {

__device__ __forceinline__ void scale_shift8(const BnArgs& a, int c, float (&sc)[8], float (&sh)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = a.gamma[c + j] * a.rstd[c + j];
    sh[j] = fmaf(-a.mean[c + j], sc[j], a.beta[c + j]);
  }
}

// ------------------------------------------------------------------------------------- inference statistics
__global__ void bn_eval_coef_kernel(const float* __restrict__ params, const float* __restrict__ state,
                                    const BnEvalEntry* __restrict__ table, float eps, float* __restrict__ scale,
                                    float* __restrict__ shift) {
  const BnEvalEntry e = table[blockIdx.y];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= e.C) return;
  const float sc = params[e.off_g + c] * (1.f / sqrtf(state[e.off_mv + c] + eps));
  scale[e.off_stat + c] = sc;
  shift[e.off_stat + c] = fmaf(-state[e.off_mm + c], sc, params[e.off_be + c]);
}
int bn_eval_coef_launch(const float* params, const float* bn_state, const BnEvalEntry* table_dev, int n_layers, int max_c,
                        float eps, float* scale, float* shift, cudaStream_t st) {
  if (n_layers == 0) return 0;
  bn_eval_coef_kernel<<<dim3((max_c + 127) / 128, n_layers), 128, 0, st>>>(params, bn_state, table_dev, eps, scale, shift);
  RVIP_LAUNCH_CHECK();
  return 0;
}
__global__ void bn_eval_prepare_kernel(const float* mm, const float* mv, float* mean, float* rstd, int n, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  mean[c] = mm[c];
  rstd[c] = 1.f / sqrtf(mv[c] + eps);
}
int bn_eval_prepare_launch(const float* mm, const float* mv, float* mean, float* rstd, int n, float eps,
                           cudaStream_t st) {
  bn_eval_prepare_kernel<<<(n + 255) / 256, 256, 0, st>>>(mm, mv, mean, rstd, n, eps);
  RVIP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------- first layer recompute
// out[j] = relu(b[c+j] + sum_t x[p + off(t)] * w[t][c+j]), 8 channels of pixel p of the 1-channel image (zero padding).
// ONE function for the statistics pass, the forward apply and both backward passes: identical arithmetic everywhere, so
// the ReLU mask and the normalised values seen by backward are exactly those of forward.
__device__ __forceinline__ void c1_recompute8(const float* __restrict__ x, int H, int W, uint32_t p, const float* w_s,
                                              const float* b_s, int C, int c, float (&out)[8]) {
  const int xx = (int)(p % (uint32_t)W), yy = (int)((p / (uint32_t)W) % (uint32_t)H);
#pragma unroll
  for (int j = 0; j < 8; ++j) out[j] = b_s[c + j];
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const bool rowok = (unsigned)(yy + dy) < (unsigned)H;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const bool ok = rowok && (unsigned)(xx + dx) < (unsigned)W;
      const float xv = ok ? __ldg(x + (size_t)p + dy * W + dx) : 0.f;
      const float4 w0 = *reinterpret_cast<const float4*>(w_s + ((dy + 1) * 3 + dx + 1) * C + c);
      const float4 w1 = *reinterpret_cast<const float4*>(w_s + ((dy + 1) * 3 + dx + 1) * C + c + 4);
      out[0] = fmaf(xv, w0.x, out[0]); out[1] = fmaf(xv, w0.y, out[1]);
      out[2] = fmaf(xv, w0.z, out[2]); out[3] = fmaf(xv, w0.w, out[3]);
      out[4] = fmaf(xv, w1.x, out[4]); out[5] = fmaf(xv, w1.y, out[5]);
      out[6] = fmaf(xv, w1.z, out[6]); out[7] = fmaf(xv, w1.w, out[7]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) out[j] = fmaxf(out[j], 0.f);
}
// stages w [9][C] and b [C] of the first layer into shared memory at dst (10 * C floats)
__device__ __forceinline__ void c1_stage_weights(const BnArgs& a, float* dst) {
  for (int k = threadIdx.x; k < 9 * a.C; k += 256) dst[k] = a.w0[k];
  for (int k = threadIdx.x; k < a.C; k += 256) dst[9 * a.C + k] = a.b0[k];
}

__global__ void __launch_bounds__(256) c1_stats_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, double* __restrict__ stats, int B,
                                                       int H, int W, int C) {
  extern __shared__ float sm_c1[];    // w [9][C], b [C], sum [C], sumsq [C]
  pdl_wait();
  float* w_s = sm_c1;
  float* b_s = w_s + 9 * C;
  float* red = b_s + C;
  for (int k = threadIdx.x; k < 9 * C; k += 256) w_s[k] = w[k];
  for (int k = threadIdx.x; k < C; k += 256) b_s[k] = bias[k];
  for (int k = threadIdx.x; k < 2 * C; k += 256) red[k] = 0.f;
  __syncthreads();
  const uint32_t G = C >> 3, lg = 31 - __clz(G);
  const uint32_t n_items = ((uint32_t)B * H * W) << lg;
  const uint32_t i0 = blockIdx.x * 256 + threadIdx.x;
  if ((i0 & ~31u) < n_items) {
    const int c = (int)(i0 & (G - 1)) * 8;
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
    for (uint32_t i = i0; i < n_items; i += gridDim.x * 256) {
      float v[8];
      c1_recompute8(x, H, W, i >> lg, w_s, b_s, C, c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += v[j];
        q[j] = fmaf(v[j], v[j], q[j]);
      }
    }
    block_accumulate8(red, c, s, G);
    block_accumulate8(red + C, c, q, G);
  }
  pdl_launch_dependents();
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * C; k += 256) atomicAdd(&stats[k], (double)red[k]);
}
int c1_stats_launch(const float* x, const float* w, const float* bias, double* stats, int B, int H, int W, int C,
                    cudaStream_t st) {
  const int G = C / 8;
  RVIP_REQUIRE(C % 8 == 0 && G <= 32 && (G & (G - 1)) == 0, "c1_stats: C=%d must be 8 * power of two <= 256", C);
  RVIP_REQUIRE((size_t)B * H * W * G < 0x7fffffffULL, "c1_stats: tensor too large for 32-bit indexing");
  launch_kernel(c1_stats_kernel, kNumSMs * 4, 256, (size_t)12 * C * sizeof(float), st, x, w, bias, stats, B, H, W, C);
  RVIP_LAUNCH_CHECK();
  return 0;
}

// item index -> pixel coordinates (32-bit; all tensors of this path have < 2^31 16-byte vectors)
struct Geo {
  uint32_t lg;       // log2(C / 8)
  uint32_t n_items;  // (#pixels or #pooling windows) * C/8
  uint32_t stride;   // gridDim.x * 256
};
__device__ __forceinline__ Geo make_geo(const BnArgs& a, bool pool) {
  Geo g;
  g.lg = 31 - __clz(a.C >> 3);
  const uint32_t P = (uint32_t)a.B * a.H * a.W;
  g.n_items = (pool ? P / 4 : P) << g.lg;
  g.stride = gridDim.x * 256;
  return g;
}
// pixel indices of the 2x2 window `win` (row-major order) of a [B,H,W] tensor
__device__ __forceinline__ void window_pixels(const BnArgs& a, uint32_t win, uint32_t (&p)[4]) {
  const uint32_t Wo = a.W >> 1, Ho = a.H >> 1;
  const uint32_t xo = win % Wo, t = win / Wo;
  const uint32_t yo = t % Ho, b = t / Ho;
  const uint32_t base = (b * a.H + 2 * yo) * a.W + 2 * xo;
  p[0] = base; p[1] = base + 1; p[2] = base + a.W; p[3] = base + a.W + 1;
}
// the four pixels of the x2 up-sampled tensor [B,2H,2W] that replicate pixel p of [B,H,W]
__device__ __forceinline__ void upsampled_pixels(const BnArgs& a, uint32_t p, uint32_t (&q)[4]) {
  const uint32_t xx = p % a.W, t = p / a.W;
  const uint32_t yy = t % a.H, b = t / a.H;
  const uint32_t base = (b * 2 * a.H + 2 * yy) * (2 * a.W) + 2 * xx;
  q[0] = base; q[1] = base + 1; q[2] = base + 2 * a.W; q[3] = base + 2 * a.W + 1;
}

// ------------------------------------------------------------------------------------- forward apply
// Training: every block first derives mean / rstd of all C channels from the conv epilogue's sum / sum^2
// (a.stats, double) into shared memory -- the former one-block bn_finalize launch, folded in; block 0 also
// publishes mean / rstd for the backward pass and updates the moving statistics (momentum 0.99, unbiased
// variance: TF fused batch norm).  Inference: mean / rstd were prepared from the moving statistics.
template <typename T, int POST, bool FROMX, bool BNF = false>
__global__ void __launch_bounds__(256) bn_apply_kernel(BnArgs a) {
  extern __shared__ float coef_s[];   // [2][C]: scale, shift  (+ FROMX: first-layer w [9][C], b [C])
  pdl_wait();
  for (int k = threadIdx.x; k < a.C; k += 256) {
    float m, r;
    if (a.stats) {
      const double mean = a.stats[k] * a.inv_count;
      double var = a.stats[a.C + k] * a.inv_count - mean * mean;   // biased batch variance
      if (var < 0) var = 0;
      m = (float)mean;
      // only the variance needs double (cancellation); 1/sqrt in fp32 is exact to ~1e-7 and 50x cheaper -- this
      // prologue runs in every block
      r = rsqrtf((float)var + a.eps);
      if (blockIdx.x == 0) {
        a.mean_out[k] = m;
        a.rstd_out[k] = r;
        const double unb = a.count > 1.0 ? var * a.count / (a.count - 1.0) : var;
        a.mov_mean[k] = a.momentum * a.mov_mean[k] + (1.f - a.momentum) * m;
        a.mov_var[k] = a.momentum * a.mov_var[k] + (1.f - a.momentum) * (float)unb;
      }
    } else if (a.identity) {
      coef_s[k] = 1.f;
      coef_s[a.C + k] = 0.f;
      continue;
    } else {
      m = a.mean[k];
      r = a.rstd[k];
    }
    const float sck = a.gamma[k] * r;
    coef_s[k] = sck;
    coef_s[a.C + k] = fmaf(-m, sck, a.beta[k]);
  }
  if (FROMX) c1_stage_weights(a, coef_s + 2 * a.C);
  __syncthreads();
  const Geo g = make_geo(a, POST == POST_POOL);
  const uint32_t i0 = blockIdx.x * 256 + threadIdx.x;
  if (i0 >= g.n_items) return;
  const int c = (int)(i0 & ((1u << g.lg) - 1)) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = coef_s[c + j];
    sh[j] = coef_s[a.C + c + j];
    // inverted dropout scales the kept values by 1 / (1 - rate): folded into the affine (ReLU commutes with a positive
    // factor, so this also holds for BN_FIRST)
    if (POST == POST_DROPOUT) {
      sc[j] *= a.keep_scale;
      sh[j] *= a.keep_scale;
    }
  }
  const T* av = static_cast<const T*>(a.a) + c;
  T* y = static_cast<T*>(a.y) + c;
  T* y2 = static_cast<T*>(a.y2) + c;
  const DropKey key = dropout_key(a.seed, a.site);
  if constexpr ((POST == POST_NONE || POST == POST_DROPOUT) && !FROMX) {
    // kU independent 16-byte loads in flight per thread before the first dependent instruction (level-1 dropout layers
    // 35 -> 27 us; profiles/r2p_sweep.jsonl).  The same pipelining of the BACKWARD passes (two items = four loads in
    // flight) left their profile-mode time unchanged and slowed the overlapped step by 1.5 % (64 registers: no room
    // beside a weight-gradient CTA of the side stream; profiles/r2r_bn_bwd_u2_sweep.jsonl) -- not kept.
    constexpr int kU = 4;
    for (uint32_t ib = i0; ib < g.n_items; ib += kU * g.stride) {
      Raw8<T> raw[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const uint32_t i = ib + u * g.stride;
        if (i < g.n_items) load_raw8(av + (size_t)(i >> g.lg) * a.C, raw[u]);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const uint32_t i = ib + u * g.stride;
        if (i >= g.n_items) break;
        float v[8];
        unpack_raw8(raw[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = fmaf(v[j], sc[j], sh[j]);
          if (BNF) v[j] = fmaxf(v[j], 0.f);     // BN_FIRST: the ReLU follows the normalisation
        }
        if (POST == POST_DROPOUT) {
          bool keep[8];
          dropout_keep8(key, i, a.thr16, keep);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = keep[j] ? v[j] : 0.f;
        }
        Vec8<T>::store(y + (size_t)(i >> g.lg) * a.C, v);
      }
    }
  } else {
  for (uint32_t i = i0; i < g.n_items; i += g.stride) {
    if (POST == POST_NONE || POST == POST_DROPOUT) {
      const size_t off = (size_t)(i >> g.lg) * a.C;
      float v[8];
      if (FROMX)
        c1_recompute8(a.x0, a.H, a.W, i >> g.lg, coef_s + 2 * a.C, coef_s + 11 * a.C, a.C, c, v);
      else
        Vec8<T>::load(av + off, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = fmaf(v[j], sc[j], sh[j]);
        if (BNF) v[j] = fmaxf(v[j], 0.f);     // BN_FIRST: the ReLU follows the normalisation
      }
      if (POST == POST_DROPOUT) {
        bool keep[8];
        dropout_keep8(key, i, a.thr16, keep);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = keep[j] ? v[j] : 0.f;
      }
      Vec8<T>::store(y + off, v);
    } else if (POST == POST_POOL) {
      const uint32_t win = i >> g.lg;
      uint32_t p[4];
      window_pixels(a, win, p);
      float mx[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float v[8];
        Vec8<T>::load(av + (size_t)p[k] * a.C, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = fmaf(v[j], sc[j], sh[j]);
          if (BNF) v[j] = fmaxf(v[j], 0.f);
          mx[j] = (k == 0) ? v[j] : fmaxf(mx[j], v[j]);
        }
        if (a.y) Vec8<T>::store(y + (size_t)p[k] * a.C, v);   // nullptr: pooling-only pass over an already normalised tensor
      }
      Vec8<T>::store(y2 + (size_t)win * a.C, mx);
    } else {  // POST_UPSAMPLE
      const uint32_t p = i >> g.lg;
      uint32_t q[4];
      upsampled_pixels(a, p, q);
      float v[8];
      Vec8<T>::load(av + (size_t)p * a.C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = fmaf(v[j], sc[j], sh[j]);
        if (BNF) v[j] = fmaxf(v[j], 0.f);
      }
      if (a.y2) {
#pragma unroll
        for (int k = 0; k < 4; ++k) Vec8<T>::store(y2 + (size_t)q[k] * a.C, v);
      }
      if (a.y) Vec8<T>::store(y + (size_t)p * a.C, v);   // low-resolution copy for the phase-decomposed up-conv
    }
  }
  }
  pdl_launch_dependents();
}

static int ew_grid(size_t n_items, int per_sm) {
  size_t g = (n_items + 255) / 256;
  const size_t cap = (size_t)kNumSMs * per_sm;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

template <typename T>
static int bn_apply_t(const BnArgs& a, cudaStream_t st) {
  const size_t P = (size_t)a.B * a.H * a.W;
  const int G = a.C / 8;
  const size_t n = (a.post == POST_POOL ? P / 4 : P) * G;
  // 56-60 registers x 256 threads (pooling variant; plain / dropout variants with four loads in flight): four blocks
  // fit an SM.  Larger grids ran their surplus blocks as a second, nearly empty wave (six per SM: 0.481 ms per step,
  // five: 0.448, four with the 4-deep loads: 0.410; profiles/r2j_sweep.jsonl, r2p_sweep.jsonl)
  static const int fwd_per_sm = getenv("RVIP_BN_FWD_BLOCKS") ? atoi(getenv("RVIP_BN_FWD_BLOCKS")) : 4;
  const int grid = ew_grid(n, fwd_per_sm);   // persistent blocks: the per-block coefficient prologue is amortised
  const size_t sm = (a.x0 ? 12 : 2) * a.C * sizeof(float);
  if (a.x0) {
    RVIP_REQUIRE(a.post == POST_NONE || a.post == POST_DROPOUT, "bn: first-layer recompute with post op %d", a.post);
    RVIP_REQUIRE(!a.bn_first, "bn: first-layer recompute is not available with BN_FIRST");
    if (a.post == POST_NONE) launch_kernel(bn_apply_kernel<T, POST_NONE, true>, grid, 256, sm, st, a);
    else launch_kernel(bn_apply_kernel<T, POST_DROPOUT, true>, grid, 256, sm, st, a);
    RVIP_LAUNCH_CHECK();
    return 0;
  }
  if (a.bn_first) {
    switch (a.post) {
      case POST_NONE: launch_kernel(bn_apply_kernel<T, POST_NONE, false, true>, grid, 256, sm, st, a); break;
      case POST_DROPOUT: launch_kernel(bn_apply_kernel<T, POST_DROPOUT, false, true>, grid, 256, sm, st, a); break;
      case POST_POOL: launch_kernel(bn_apply_kernel<T, POST_POOL, false, true>, grid, 256, sm, st, a); break;
      default: launch_kernel(bn_apply_kernel<T, POST_UPSAMPLE, false, true>, grid, 256, sm, st, a); break;
    }
    RVIP_LAUNCH_CHECK();
    return 0;
  }
  switch (a.post) {
    case POST_NONE: launch_kernel(bn_apply_kernel<T, POST_NONE, false>, grid, 256, sm, st, a); break;
    case POST_DROPOUT: launch_kernel(bn_apply_kernel<T, POST_DROPOUT, false>, grid, 256, sm, st, a); break;
    case POST_POOL: launch_kernel(bn_apply_kernel<T, POST_POOL, false>, grid, 256, sm, st, a); break;
    default: launch_kernel(bn_apply_kernel<T, POST_UPSAMPLE, false>, grid, 256, sm, st, a); break;
  }
  RVIP_LAUNCH_CHECK();
  return 0;
}
static int check_bn(const BnArgs& a) {
  const int G = a.C / 8;
  RVIP_REQUIRE(a.C % 8 == 0 && G <= 256 && (G & (G - 1)) == 0, "bn: C=%d must be 8 * power of two (<= 2048)", a.C);
  RVIP_REQUIRE(a.post != POST_POOL || (a.H % 2 == 0 && a.W % 2 == 0), "bn: max-pool needs even H, W (got %dx%d)", a.H,
               a.W);
  RVIP_REQUIRE((size_t)a.B * a.H * a.W * 4 * G < 0x7fffffffULL, "bn: tensor too large for 32-bit vector indexing");
  return 0;
}
int bn_apply_launch(const BnArgs& a, int is_bf16, cudaStream_t st) {
  if (check_bn(a)) return 1;
  return is_bf16 ? bn_apply_t<__nv_bfloat16>(a, st) : bn_apply_t<float>(a, st);
}

// ------------------------------------------------------------------------------------- backward
// Work item -> K pixels (4 for a pooling window, else 1) with dy = dL/d(BN output) gathered from the
// consumers' gradient buffers and the stored relu(conv) values.
template <typename T, int POST, bool FROMX = false, bool BNF = false>
struct Gather {
  static constexpr int K = POST == POST_POOL ? 4 : 1;
  // w1_s: first-layer weights / bias staged in shared memory (FROMX only)
  __device__ static __forceinline__ void run(const BnArgs& a, const Geo& g, const DropKey& key, uint32_t i, int c,
                                             const float (&sc)[8], const float (&sh)[8], uint32_t (&pix)[K],
                                             float (&av)[K][8], float (&dy)[K][8], const float* w1_s = nullptr) {
    const T* A = static_cast<const T*>(a.a) + c;
    const T* g0 = static_cast<const T*>(a.g0) + c;
    if constexpr (POST == POST_NONE || POST == POST_DROPOUT) {
      const uint32_t p = i >> g.lg;
      pix[0] = p;
      if (FROMX)
        c1_recompute8(a.x0, a.H, a.W, p, w1_s, w1_s + 9 * a.C, a.C, c, av[0]);
      else
        Vec8<T>::load(A + (size_t)p * a.C, av[0]);
      Vec8<T>::load(g0 + (size_t)p * a.C, dy[0]);
      if (POST == POST_DROPOUT) {
        bool keep[8];
        dropout_keep8(key, i, a.thr16, keep);
#pragma unroll
        for (int j = 0; j < 8; ++j) dy[0][j] = keep[j] ? dy[0][j] : 0.f;   // x keep_scale: folded into the callers' coefficients
      }
    } else if constexpr (POST == POST_UPSAMPLE) {
      const uint32_t p = i >> g.lg;
      pix[0] = p;
      uint32_t q[4];
      upsampled_pixels(a, p, q);
      Vec8<T>::load(A + (size_t)p * a.C, av[0]);
      float t[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k) Vec8<T>::load(g0 + (size_t)q[k] * a.C, t[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) dy[0][j] = (t[0][j] + t[1][j]) + (t[2][j] + t[3][j]);
    } else {  // POST_POOL: skip gradient + pooled gradient routed to the first maximum (row-major, strict >)
      const T* g1 = static_cast<const T*>(a.g1) + c;
      const uint32_t win = i >> g.lg;
      window_pixels(a, win, pix);
      float best[8], dp[8];
      int arg[8];
      Vec8<T>::load(g1 + (size_t)win * a.C, dp);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        Vec8<T>::load(A + (size_t)pix[k] * a.C, av[k]);
        Vec8<T>::load(g0 + (size_t)pix[k] * a.C, dy[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float yv = fmaf(av[k][j], sc[j], sh[j]);
          if (BNF) yv = fmaxf(yv, 0.f);
          if (k == 0 || yv > best[j]) {
            best[j] = yv;
            arg[j] = k;
          }
        }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (arg[j] == k) dy[k][j] += dp[j];
    }
    if (BNF) {
      // y = relu(BN(z)): the gradient passes where the normalised value was positive (same float expression as forward)
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (!(fmaf(av[k][j], sc[j], sh[j]) > 0.f)) dy[k][j] = 0.f;
    }
  }
};

// pass 1: red[stripe][c] += sum dy, red[stripe][C + c] += sum dy * a   (raw a; normalised by the finalize kernel).
// Blocks spread their double atomics over kRedStripes copies: ~1200 blocks adding to ONE copy serialise in the L2
// atomic unit for ~17 us (profiles/microbench/atomics.cu), 16 copies cost nothing measurable.
template <typename T, int POST, bool FROMX, bool BNF = false>
__global__ void __launch_bounds__(256, POST == POST_POOL ? 2 : 4) bn_bwd_reduce_kernel(BnArgs a) {
  extern __shared__ float red_s[];  // [2][C]  (+ FROMX: first-layer w [9][C], b [C])
  pdl_wait();
  constexpr int K = Gather<T, POST>::K;
  const Geo g = make_geo(a, POST == POST_POOL);
  for (int k = threadIdx.x; k < 2 * a.C; k += 256) red_s[k] = 0.f;
  if (FROMX) c1_stage_weights(a, red_s + 2 * a.C);
  __syncthreads();
  const uint32_t i0 = blockIdx.x * 256 + threadIdx.x;
  if ((i0 & ~31u) < g.n_items) {   // warp-uniform: n_items is a multiple of 32 vectors or the warp is partial
    const int c = (int)(i0 & ((1u << g.lg) - 1)) * 8;
    float sc[8], sh[8], s1[8], s2[8];
    if (POST == POST_POOL || BNF) scale_shift8(a, c, sc, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    const DropKey key = dropout_key(a.seed, a.site);
    for (uint32_t i = i0; i < g.n_items; i += g.stride) {
      uint32_t pix[K];
      float av[K][8], dy[K][8];
      Gather<T, POST, FROMX, BNF>::run(a, g, key, i, c, sc, sh, pix, av, dy, red_s + 2 * a.C);
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += dy[k][j];
          s2[j] = fmaf(dy[k][j], av[k][j], s2[j]);
        }
    }
    pdl_launch_dependents();
    if (POST == POST_DROPOUT) {   // dy of a kept element is keep_scale times the stored gradient
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] *= a.keep_scale;
        s2[j] *= a.keep_scale;
      }
    }
    block_accumulate8(red_s, c, s1, 1u << g.lg);
    block_accumulate8(red_s + a.C, c, s2, 1u << g.lg);
  }
  __syncthreads();
  double* dst = a.red + (size_t)(blockIdx.x % kRedStripes) * 2 * a.C;
  for (int k = threadIdx.x; k < 2 * a.C; k += 256) atomicAdd(&dst[k], (double)red_s[k]);
}

// pass 2: dz and the conv-bias gradient.  Prologue (every block, one thread per channel): fold the stripes and
// derive the coefficients of  dz = [a>0] * ( sc*dy - k1*a + c0 )  in double, once per channel; block 0 also
// writes dgamma / dbeta.
template <typename T, int POST, bool FROMX, bool BNF = false>
__global__ void __launch_bounds__(256, POST == POST_POOL ? 2 : 4) bn_bwd_apply_kernel(BnArgs a) {
  extern __shared__ float red_s[];  // [C] bias-gradient partials, then [4][C] sc, k1, c0, shift (+ FROMX: w [9][C], b [C])
  pdl_wait();
  float* coef_s = red_s + a.C;
  constexpr int K = Gather<T, POST>::K;
  const Geo g = make_geo(a, POST == POST_POOL);
  const double inv_P = 1.0 / ((double)a.B * a.H * a.W);
  for (int k = threadIdx.x; k < a.C; k += 256) {
    red_s[k] = 0.f;
    if (a.identity) {      // no BatchNorm in this block: dz = [a > 0] * dy
      coef_s[k] = 1.f;
      coef_s[a.C + k] = coef_s[2 * a.C + k] = coef_s[3 * a.C + k] = 0.f;
      continue;
    }
    double s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int s = 0; s < kRedStripes; ++s) {
      s1 += a.red[(size_t)s * 2 * a.C + k];
      s2 += a.red[(size_t)s * 2 * a.C + a.C + k];
    }
    const double mu = a.mean[k], r = a.rstd[k], sck = (double)a.gamma[k] * r;
    const double sda = r * (s2 - mu * s1);          // sum dy * ahat
    const double k1 = sck * r * (sda * inv_P);
    coef_s[k] = (float)sck;
    coef_s[a.C + k] = (float)k1;
    coef_s[2 * a.C + k] = (float)(k1 * mu - sck * (s1 * inv_P));
    coef_s[3 * a.C + k] = fmaf(-a.mean[k], a.gamma[k] * a.rstd[k], a.beta[k]);
    if (blockIdx.x == 0) {
      a.dbeta[k] = (float)s1;
      a.dgamma[k] = (float)sda;
    }
  }
  if (FROMX) c1_stage_weights(a, red_s + 5 * a.C);
  __syncthreads();
  const uint32_t i0 = blockIdx.x * 256 + threadIdx.x;
  if ((i0 & ~31u) < g.n_items) {
    const int c = (int)(i0 & ((1u << g.lg) - 1)) * 8;
    float sc[8], sh[8], k1[8], c0[8], db[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = coef_s[c + j];
      k1[j] = coef_s[a.C + c + j];
      c0[j] = coef_s[2 * a.C + c + j];
      if ((POST == POST_POOL || BNF) && !a.identity) {
        // the pooling argmax / BN_FIRST ReLU mask replay the forward values: same float expressions as scale_shift8
        sc[j] = a.gamma[c + j] * a.rstd[c + j];
        sh[j] = coef_s[3 * a.C + c + j];
      } else {
        sh[j] = 0.f;
      }
    }
    // coefficient of the gathered dy in dz: the dropout variant's gather leaves out the 1 / (1 - rate) factor
    float scd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      db[j] = 0.f;
      // (the pooling variant keeps ONE copy -- the float product its argmax replay uses: it runs at the 128-register cap)
      scd[j] = POST == POST_DROPOUT ? coef_s[c + j] * a.keep_scale : (POST == POST_POOL ? sc[j] : coef_s[c + j]);
    }
    T* dzp = static_cast<T*>(a.dz) + c;
    const DropKey key = dropout_key(a.seed, a.site);
    for (uint32_t i = i0; i < g.n_items; i += g.stride) {
      uint32_t pix[K];
      float av[K][8], dy[K][8];
      Gather<T, POST, FROMX, BNF>::run(a, g, key, i, c, sc, sh, pix, av, dy, red_s + 5 * a.C);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float dz[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float da = fmaf(scd[j], dy[k][j], fmaf(-k1[j], av[k][j], c0[j]));
          dz[j] = (BNF || av[k][j] > 0.f) ? da : 0.f;
          db[j] += dz[j];
        }
        Vec8<T>::store(dzp + (size_t)pix[k] * a.C, dz);
      }
    }
    pdl_launch_dependents();
    block_accumulate8(red_s, c, db, 1u << g.lg);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < a.C; k += 256) atomicAdd(&a.dbias[k], red_s[k]);
}

template <typename T, int WHICH>
static int bn_bwd_t(const BnArgs& a, cudaStream_t st) {
  const size_t P = (size_t)a.B * a.H * a.W;
  const int G = a.C / 8;
  const size_t n = (a.post == POST_POOL ? P / 4 : P) * G;
  // persistent grids: every block ends with C global atomics, so few, long-lived blocks.  Four per SM (all the
  // 64-register kernels allow): ncu shows these passes latency bound (long-scoreboard stalls 13-17 warps per issue, DRAM
  // at 51-59 %), so occupancy buys bandwidth -- 1.388 -> 1.316 ms per step, and the overlapped step gains too
  // (4.63 -> 4.54 ms) although a co-resident weight-gradient CTA of the side stream now finds fewer free registers
  static const int per_sm = getenv("RVIP_BN_BWD_BLOCKS") ? atoi(getenv("RVIP_BN_BWD_BLOCKS")) : 4;
  const int grid = ew_grid(n, a.post == POST_POOL ? 2 : per_sm);
  const size_t sm = ((WHICH == 0 ? 2 : 5) + (a.x0 ? 10 : 0)) * a.C * sizeof(float);
  if (a.x0) {
    RVIP_REQUIRE(a.post == POST_NONE || a.post == POST_DROPOUT, "bn: first-layer recompute with post op %d", a.post);
    if (a.post == POST_NONE) {
      if (WHICH == 0) launch_kernel(bn_bwd_reduce_kernel<T, POST_NONE, true>, grid, 256, sm, st, a);
      else launch_kernel(bn_bwd_apply_kernel<T, POST_NONE, true>, grid, 256, sm, st, a);
    } else {
      if (WHICH == 0) launch_kernel(bn_bwd_reduce_kernel<T, POST_DROPOUT, true>, grid, 256, sm, st, a);
      else launch_kernel(bn_bwd_apply_kernel<T, POST_DROPOUT, true>, grid, 256, sm, st, a);
    }
    RVIP_LAUNCH_CHECK();
    return 0;
  }
#define RVIP_BWD(POSTV)                                                                \
  if (a.bn_first) {                                                                    \
    if (WHICH == 0)                                                                    \
      launch_kernel(bn_bwd_reduce_kernel<T, POSTV, false, true>, grid, 256, sm, st, a); \
    else                                                                               \
      launch_kernel(bn_bwd_apply_kernel<T, POSTV, false, true>, grid, 256, sm, st, a);  \
  } else if (WHICH == 0)                                                               \
    launch_kernel(bn_bwd_reduce_kernel<T, POSTV, false>, grid, 256, sm, st, a);         \
  else                                                                                 \
    launch_kernel(bn_bwd_apply_kernel<T, POSTV, false>, grid, 256, sm, st, a);
  switch (a.post) {
    case POST_NONE: RVIP_BWD(POST_NONE) break;
    case POST_DROPOUT: RVIP_BWD(POST_DROPOUT) break;
    case POST_POOL: RVIP_BWD(POST_POOL) break;
    default: RVIP_BWD(POST_UPSAMPLE) break;
  }
#undef RVIP_BWD
  RVIP_LAUNCH_CHECK();
  return 0;
}
int bn_bwd_reduce_launch(const BnArgs& a, int is_bf16, cudaStream_t st) {
  if (check_bn(a)) return 1;
  return is_bf16 ? bn_bwd_t<__nv_bfloat16, 0>(a, st) : bn_bwd_t<float, 0>(a, st);
}
int bn_bwd_apply_launch(const BnArgs& a, int is_bf16, cudaStream_t st) {
  if (check_bn(a)) return 1;
  return is_bf16 ? bn_bwd_t<__nv_bfloat16, 1>(a, st) : bn_bwd_t<float, 1>(a, st);
}

// ------------------------------------------------------------------------------------- ReLU backward (up-conv)
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_kernel(const T* __restrict__ u, const T* __restrict__ du,
                                                       T* __restrict__ dz, float* dbias, uint32_t n_items, uint32_t lg,
                                                       int C) {
  extern __shared__ float red_s[];
  pdl_wait();
  for (int k = threadIdx.x; k < C; k += 256) red_s[k] = 0.f;
  __syncthreads();
  const uint32_t i0 = blockIdx.x * 256 + threadIdx.x;
  if ((i0 & ~31u) < n_items) {
    const int c = (int)(i0 & ((1u << lg) - 1)) * 8;
    float db[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) db[j] = 0.f;
    for (uint32_t i = i0; i < n_items; i += gridDim.x * 256) {
      const size_t off = (size_t)(i >> lg) * C + c;
      float uv[8], gv[8];
      Vec8<T>::load(u + off, uv);
      Vec8<T>::load(du + off, gv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        gv[j] = uv[j] > 0.f ? gv[j] : 0.f;
        db[j] += gv[j];
      }
      Vec8<T>::store(dz + off, gv);
    }
    pdl_launch_dependents();
    block_accumulate8(red_s, c, db, 1u << lg);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < C; k += 256) atomicAdd(&dbias[k], red_s[k]);
}
int relu_bwd_launch(const void* u, const void* du, void* dz, float* dbias, size_t pixels, int C, int is_bf16,
                    cudaStream_t st) {
  const int G = C / 8;
  RVIP_REQUIRE(C % 8 == 0 && G <= 256 && (G & (G - 1)) == 0, "relu_bwd: C=%d must be 8 * power of two", C);
  RVIP_REQUIRE(pixels * G < 0x7fffffffULL, "relu_bwd: tensor too large for 32-bit vector indexing");
  uint32_t lg = 0;
  while ((1 << lg) < G) ++lg;
  const uint32_t n = (uint32_t)(pixels * G);
  const int grid = ew_grid(n, 4);
  if (is_bf16)
    launch_kernel(relu_bwd_kernel<__nv_bfloat16>, grid, 256, C * sizeof(float), st, static_cast<const __nv_bfloat16*>(u),
                  static_cast<const __nv_bfloat16*>(du), static_cast<__nv_bfloat16*>(dz), dbias, n, lg, C);
  else
    launch_kernel(relu_bwd_kernel<float>, grid, 256, C * sizeof(float), st, static_cast<const float*>(u),
                  static_cast<const float*>(du), static_cast<float*>(dz), dbias, n, lg, C);
  RVIP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------- dropout mask export (tests)
__global__ void dropout_mask_kernel(uint64_t seed, uint32_t site, uint32_t thr16, size_t n_vec8, uint8_t* keep) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vec8) return;
  bool k[8];
  dropout_keep8(seed, site, i, thr16, k);
#pragma unroll
  for (int j = 0; j < 8; ++j) keep[i * 8 + j] = k[j] ? 1 : 0;
}
int dropout_mask_launch(uint64_t seed, uint32_t site, uint32_t thr16, size_t n_vec8, uint8_t* keep, cudaStream_t st) {
  dropout_mask_kernel<<<(unsigned)((n_vec8 + 255) / 256), 256, 0, st>>>(seed, site, thr16, n_vec8, keep);
  RVIP_LAUNCH_CHECK();
  return 0;
}

}  // namespace rvip
