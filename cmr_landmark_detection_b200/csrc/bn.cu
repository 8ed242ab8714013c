// Memory-bound passes around the convolutions: BatchNorm finalize / apply (fused with Dropout,
// MaxPool 2x2 or nearest UpSampling x2) and the fused BatchNorm+ReLU backward (fused with the
// dropout-mask replay, max-pool gradient routing or up-sampling 2x2 gradient sum).
// Reference ops replaced (all TF kernels reached from src/models/KerasLayers.py):
//   BatchNormalization(axis=-1) :684,691   Dropout :718,772 (Unets.py:813)   MaxPooling2D :714,721
//   UpSampling2D :756-757   and their gradients.  Block order is Conv -> ReLU -> BN (BN_FIRST false).
// Every thread moves 8 channels (16 B bf16 / 32 B fp32) of one pixel: fully coalesced NHWC traffic.
#include "kernels.cuh"

namespace rvip {

__device__ __forceinline__ void scale_shift8(const BnArgs& a, int c, float (&sc)[8], float (&sh)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = a.gamma[c + j] * a.rstd[c + j];
    sh[j] = fmaf(-a.mean[c + j], sc[j], a.beta[c + j]);
  }
}

// ------------------------------------------------------------------------------------- finalize
__global__ void bn_finalize_kernel(const double* stats, double count, float* mean, float* rstd, float* mm, float* mv,
                                   int C, float momentum, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = stats[c] / count;
  double var = stats[C + c] / count - m * m;   // biased batch variance
  if (var < 0) var = 0;
  mean[c] = (float)m;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  // moving statistics: momentum 0.99, unbiased variance (TF fused batch norm)
  const double unb = count > 1 ? var * count / (count - 1) : var;
  mm[c] = momentum * mm[c] + (1.f - momentum) * (float)m;
  mv[c] = momentum * mv[c] + (1.f - momentum) * (float)unb;
}
int bn_finalize_launch(const double* stats, double count, float* mean, float* rstd, float* mm, float* mv, int C,
                       float momentum, float eps, cudaStream_t st) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(stats, count, mean, rstd, mm, mv, C, momentum, eps);
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

// ------------------------------------------------------------------------------------- forward apply
template <typename T, int POST>
__global__ void __launch_bounds__(256) bn_apply_kernel(BnArgs a) {
  const int G = a.C >> 3;
  const size_t P = (size_t)a.B * a.H * a.W;
  const size_t n_items = (POST == POST_POOL ? P / 4 : P) * G;
  const size_t i0 = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i0 >= n_items) return;
  const int c = (int)(i0 % G) * 8;
  float sc[8], sh[8];
  scale_shift8(a, c, sc, sh);
  const T* av = static_cast<const T*>(a.a);
  T* y = static_cast<T*>(a.y);
  T* y2 = static_cast<T*>(a.y2);
  for (size_t i = i0; i < n_items; i += (size_t)gridDim.x * 256) {
    if (POST == POST_NONE || POST == POST_DROPOUT) {
      const size_t p = i / G;
      float v[8];
      Vec8<T>::load(av + p * a.C + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
      if (POST == POST_DROPOUT) {
        bool keep[8];
        dropout_keep8(a.seed, a.site, i, a.thr16, keep);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = keep[j] ? v[j] * a.keep_scale : 0.f;
      }
      Vec8<T>::store(y + p * a.C + c, v);
    } else if (POST == POST_POOL) {
      const size_t win = i / G;
      const int Wo = a.W >> 1, Ho = a.H >> 1;
      const int xo = (int)(win % Wo), yo = (int)((win / Wo) % Ho);
      const size_t b = win / ((size_t)Wo * Ho);
      float mx[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const size_t p = (b * a.H + 2 * yo + (k >> 1)) * a.W + 2 * xo + (k & 1);
        float v[8];
        Vec8<T>::load(av + p * a.C + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = fmaf(v[j], sc[j], sh[j]);
          mx[j] = (k == 0) ? v[j] : fmaxf(mx[j], v[j]);
        }
        Vec8<T>::store(y + p * a.C + c, v);
      }
      Vec8<T>::store(y2 + win * a.C + c, mx);
    } else {  // POST_UPSAMPLE
      const size_t p = i / G;
      const int xx = (int)(p % a.W), yy = (int)((p / a.W) % a.H);
      const size_t b = p / ((size_t)a.W * a.H);
      float v[8];
      Vec8<T>::load(av + p * a.C + c, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const size_t q = (b * 2 * a.H + 2 * yy + (k >> 1)) * (2 * a.W) + 2 * xx + (k & 1);
        Vec8<T>::store(y2 + q * a.C + c, v);
      }
    }
  }
}

static int ew_grid(size_t n_items) {
  size_t g = (n_items + 255) / 256;
  const size_t cap = (size_t)kNumSMs * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

template <typename T>
static int bn_apply_t(const BnArgs& a, cudaStream_t st) {
  const size_t P = (size_t)a.B * a.H * a.W;
  const int G = a.C / 8;
  const size_t n = (a.post == POST_POOL ? P / 4 : P) * G;
  const int grid = ew_grid(n);
  switch (a.post) {
    case POST_NONE: bn_apply_kernel<T, POST_NONE><<<grid, 256, 0, st>>>(a); break;
    case POST_DROPOUT: bn_apply_kernel<T, POST_DROPOUT><<<grid, 256, 0, st>>>(a); break;
    case POST_POOL: bn_apply_kernel<T, POST_POOL><<<grid, 256, 0, st>>>(a); break;
    default: bn_apply_kernel<T, POST_UPSAMPLE><<<grid, 256, 0, st>>>(a); break;
  }
  RVIP_LAUNCH_CHECK();
  return 0;
}
static int check_bn(const BnArgs& a) {
  const int G = a.C / 8;
  RVIP_REQUIRE(a.C % 8 == 0 && G <= 256 && (G & (G - 1)) == 0, "bn: C=%d must be 8 * power of two (<= 2048)", a.C);
  RVIP_REQUIRE(a.post != POST_POOL || (a.H % 2 == 0 && a.W % 2 == 0), "bn: max-pool needs even H, W (got %dx%d)", a.H,
               a.W);
  return 0;
}
int bn_apply_launch(const BnArgs& a, int is_bf16, cudaStream_t st) {
  if (check_bn(a)) return 1;
  return is_bf16 ? bn_apply_t<__nv_bfloat16>(a, st) : bn_apply_t<float>(a, st);
}

// ------------------------------------------------------------------------------------- backward
// Work item -> K pixels (4 for a pooling window, else 1) with dy = dL/d(BN output) gathered from the
// consumers' gradient buffers and the stored relu(conv) values.
template <typename T, int POST>
struct Gather {
  static constexpr int K = POST == POST_POOL ? 4 : 1;
  __device__ static __forceinline__ void run(const BnArgs& a, size_t i, int G, int c, const float (&sc)[8],
                                             const float (&sh)[8], size_t (&pix)[K], float (&av)[K][8],
                                             float (&dy)[K][8]) {
    const T* A = static_cast<const T*>(a.a);
    const T* g0 = static_cast<const T*>(a.g0);
    if (POST == POST_NONE || POST == POST_DROPOUT) {
      const size_t p = i / G;
      pix[0] = p;
      Vec8<T>::load(A + p * a.C + c, av[0]);
      Vec8<T>::load(g0 + p * a.C + c, dy[0]);
      if (POST == POST_DROPOUT) {
        bool keep[8];
        dropout_keep8(a.seed, a.site, i, a.thr16, keep);
#pragma unroll
        for (int j = 0; j < 8; ++j) dy[0][j] = keep[j] ? dy[0][j] * a.keep_scale : 0.f;
      }
    } else if (POST == POST_UPSAMPLE) {
      const size_t p = i / G;
      pix[0] = p;
      const int xx = (int)(p % a.W), yy = (int)((p / a.W) % a.H);
      const size_t b = p / ((size_t)a.W * a.H);
      Vec8<T>::load(A + p * a.C + c, av[0]);
#pragma unroll
      for (int j = 0; j < 8; ++j) dy[0][j] = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const size_t q = (b * 2 * a.H + 2 * yy + (k >> 1)) * (2 * a.W) + 2 * xx + (k & 1);
        float t[8];
        Vec8<T>::load(g0 + q * a.C + c, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) dy[0][j] += t[j];
      }
    } else {  // POST_POOL: skip gradient + pooled gradient routed to the first maximum (row-major, strict >)
      const T* g1 = static_cast<const T*>(a.g1);
      const size_t win = i / G;
      const int Wo = a.W >> 1, Ho = a.H >> 1;
      const int xo = (int)(win % Wo), yo = (int)((win / Wo) % Ho);
      const size_t b = win / ((size_t)Wo * Ho);
      float best[8], dp[8];
      int arg[8];
      Vec8<T>::load(g1 + win * a.C + c, dp);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const size_t p = (b * a.H + 2 * yo + (k >> 1)) * a.W + 2 * xo + (k & 1);
        pix[k] = p;
        Vec8<T>::load(A + p * a.C + c, av[k]);
        Vec8<T>::load(g0 + p * a.C + c, dy[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float yv = fmaf(av[k][j], sc[j], sh[j]);
          if (k == 0 || yv > best[j]) {
            best[j] = yv;
            arg[j] = k;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (arg[j] == k) dy[k][j] += dp[j];
    }
  }
};

template <typename T, int POST>
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(BnArgs a) {
  extern __shared__ float red_s[];  // [2][C]
  constexpr int K = Gather<T, POST>::K;
  const int G = a.C >> 3;
  const size_t P = (size_t)a.B * a.H * a.W;
  const size_t n_items = (POST == POST_POOL ? P / 4 : P) * G;
  for (int k = threadIdx.x; k < 2 * a.C; k += 256) red_s[k] = 0.f;
  __syncthreads();
  const size_t i0 = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i0 < n_items) {
    const int c = (int)(i0 % G) * 8;
    float sc[8], sh[8], s1[8], s2[8], mean[8], rstd[8];
    scale_shift8(a, c, sc, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1[j] = s2[j] = 0.f;
      mean[j] = a.mean[c + j];
      rstd[j] = a.rstd[c + j];
    }
    // U independent work items per iteration keep 2U..9U 16-byte loads in flight per thread
    constexpr int U = K == 1 ? 2 : 1;
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = i0; i < n_items; i += U * stride) {
      size_t pix[U][K];
      float av[U][K][8], dy[U][K][8];
      bool live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t iu = i + u * stride;
        live[u] = iu < n_items;
        Gather<T, POST>::run(a, live[u] ? iu : i, G, c, sc, sh, pix[u], av[u], dy[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!live[u]) continue;
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            s1[j] += dy[u][k][j];
            s2[j] = fmaf(dy[u][k][j], (av[u][k][j] - mean[j]) * rstd[j], s2[j]);
          }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red_s[c + j], s1[j]);
      atomicAdd(&red_s[a.C + c + j], s2[j]);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * a.C; k += 256) atomicAdd(&a.red[k], (double)red_s[k]);
}

template <typename T, int POST>
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(BnArgs a) {
  extern __shared__ float red_s[];  // [C] bias-gradient partials
  constexpr int K = Gather<T, POST>::K;
  const int G = a.C >> 3;
  const size_t P = (size_t)a.B * a.H * a.W;
  const size_t n_items = (POST == POST_POOL ? P / 4 : P) * G;
  for (int k = threadIdx.x; k < a.C; k += 256) red_s[k] = 0.f;
  if (blockIdx.x == 0) {
    for (int k = threadIdx.x; k < a.C; k += 256) {
      a.dbeta[k] = (float)a.red[k];
      a.dgamma[k] = (float)a.red[a.C + k];
    }
  }
  __syncthreads();
  const size_t i0 = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i0 < n_items) {
    const int c = (int)(i0 % G) * 8;
    float sc[8], sh[8], mean[8], rstd[8], m1[8], m2[8], db[8];
    scale_shift8(a, c, sc, sh);
    const double invP = 1.0 / (double)P;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean[j] = a.mean[c + j];
      rstd[j] = a.rstd[c + j];
      m1[j] = (float)(a.red[c + j] * invP);
      m2[j] = (float)(a.red[a.C + c + j] * invP);
      db[j] = 0.f;
    }
    T* dzp = static_cast<T*>(a.dz);
    constexpr int U = K == 1 ? 2 : 1;
    const size_t stride = (size_t)gridDim.x * 256;
    for (size_t i = i0; i < n_items; i += U * stride) {
      size_t pix[U][K];
      float av[U][K][8], dy[U][K][8];
      bool live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t iu = i + u * stride;
        live[u] = iu < n_items;
        Gather<T, POST>::run(a, live[u] ? iu : i, G, c, sc, sh, pix[u], av[u], dy[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!live[u]) continue;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float dz[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float ahat = (av[u][k][j] - mean[j]) * rstd[j];
            const float da = sc[j] * (dy[u][k][j] - m1[j] - ahat * m2[j]);
            dz[j] = av[u][k][j] > 0.f ? da : 0.f;
            db[j] += round_to<T>(dz[j]);
          }
          Vec8<T>::store(dzp + pix[u][k] * a.C + c, dz);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&red_s[c + j], db[j]);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < a.C; k += 256) atomicAdd(&a.dbias[k], red_s[k]);
}

template <typename T, int WHICH>
static int bn_bwd_t(const BnArgs& a, cudaStream_t st) {
  const size_t P = (size_t)a.B * a.H * a.W;
  const int G = a.C / 8;
  const size_t n = (a.post == POST_POOL ? P / 4 : P) * G;
  size_t g = (n + 255) / 256;
  const size_t cap = (size_t)kNumSMs * 8;
  const int grid = (int)(g < cap ? (g ? g : 1) : cap);
  const size_t sm = (WHICH == 0 ? 2 : 1) * a.C * sizeof(float);
#define RVIP_BWD(POSTV)                                                 \
  if (WHICH == 0)                                                       \
    bn_bwd_reduce_kernel<T, POSTV><<<grid, 256, sm, st>>>(a);           \
  else                                                                  \
    bn_bwd_apply_kernel<T, POSTV><<<grid, 256, sm, st>>>(a);
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
__global__ void __launch_bounds__(256) relu_bwd_kernel(const T* u, const T* du, T* dz, float* dbias, size_t pixels,
                                                       int C) {
  extern __shared__ float red_s[];
  const int G = C >> 3;
  const size_t n_items = pixels * G;
  for (int k = threadIdx.x; k < C; k += 256) red_s[k] = 0.f;
  __syncthreads();
  const size_t i0 = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i0 < n_items) {
    const int c = (int)(i0 % G) * 8;
    float db[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) db[j] = 0.f;
    for (size_t i = i0; i < n_items; i += (size_t)gridDim.x * 256) {
      const size_t p = i / G;
      float uv[8], g[8];
      Vec8<T>::load(u + p * C + c, uv);
      Vec8<T>::load(du + p * C + c, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        g[j] = uv[j] > 0.f ? g[j] : 0.f;
        db[j] += g[j];
      }
      Vec8<T>::store(dz + p * C + c, g);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&red_s[c + j], db[j]);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < C; k += 256) atomicAdd(&dbias[k], red_s[k]);
}
int relu_bwd_launch(const void* u, const void* du, void* dz, float* dbias, size_t pixels, int C, int is_bf16,
                    cudaStream_t st) {
  const int G = C / 8;
  RVIP_REQUIRE(C % 8 == 0 && G <= 256 && (G & (G - 1)) == 0, "relu_bwd: C=%d must be 8 * power of two", C);
  size_t g = (pixels * G + 255) / 256;
  const size_t cap = (size_t)kNumSMs * 8;
  const int grid = (int)(g < cap ? (g ? g : 1) : cap);
  if (is_bf16)
    relu_bwd_kernel<__nv_bfloat16><<<grid, 256, C * sizeof(float), st>>>(
        static_cast<const __nv_bfloat16*>(u), static_cast<const __nv_bfloat16*>(du), static_cast<__nv_bfloat16*>(dz),
        dbias, pixels, C);
  else
    relu_bwd_kernel<float><<<grid, 256, C * sizeof(float), st>>>(static_cast<const float*>(u),
                                                                  static_cast<const float*>(du),
                                                                  static_cast<float*>(dz), dbias, pixels, C);
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
