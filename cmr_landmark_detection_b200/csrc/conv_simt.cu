// CUDA-core 3x3 convolution kernels (fp32 accumulate).
// Used for (1) "fp32 mode" -- every conv runs here on fp32 NHWC tensors, which is the mode whose
// argmax indices must be bit-exact against the oracle -- and (2) the Cin=1 first layer in bf16 mode
// (K = 9: HBM-bound, tensor cores have nothing to chew on).  Same math as conv_tc.cu:
//   fwd/dgrad:  out[p, n] = epi( sum_{tap,c} in[p+off(tap), c] * w[tap][c][n] )
//   wgrad:      dw[tap][c][n] += sum_p in[p+off(tap), c] * dz[p, n]
// Reference ops replaced: tf.keras Conv2D fwd + autodiff (src/models/KerasLayers.py:683,689,758).
#include "kernels.cuh"

namespace rvip {

constexpr int TP = 8;        // 8x8 pixel patch per tile
constexpr int CK = 16;       // input channels per smem chunk
constexpr int TN = 32;       // output channels per block

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(256) conv3x3_simt_kernel(ConvSimtArgs a) {
  __shared__ float in_s[(TP + 2) * (TP + 2)][CK + 1];
  __shared__ __align__(16) float w_s[9][CK][TN];
  __shared__ float s_sum[TN], s_sq[TN];
  const Tin* in0 = static_cast<const Tin*>(a.in0);
  const Tin* in1 = static_cast<const Tin*>(a.in1);
  const int t = threadIdx.x;
  const int pix = t >> 2, tq = t & 3;
  const int py = pix >> 3, px = pix & 7;
  const int n0 = blockIdx.y * TN;
  const int tiles_x = (a.W + TP - 1) / TP, tiles_y = (a.H + TP - 1) / TP;
  const int n_tiles = tiles_x * tiles_y * a.B;
  const int C1 = a.Ctot - a.C0;
  if (t < TN) s_sum[t] = s_sq[t] = 0.f;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int bx = tile % tiles_x, by = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int x0 = bx * TP, y0 = by * TP;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int c0 = 0; c0 < a.Ctot; c0 += CK) {
      const int cw = (a.Ctot - c0) < CK ? (a.Ctot - c0) : CK;
      __syncthreads();
      for (int e = t; e < (TP + 2) * (TP + 2) * CK; e += 256) {
        const int ci = e % CK, p = e / CK;
        const int yy = y0 + p / (TP + 2) - 1, xx = x0 + p % (TP + 2) - 1;
        float v = 0.f;
        const int c = c0 + ci;
        if (ci < cw && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) {
          if (a.in_stuffed) {
            if ((yy & 1) && (xx & 1))
              v = to_f32<Tin>(in0[(((size_t)b * (a.H >> 1) + (yy >> 1)) * (a.W >> 1) + (xx >> 1)) * a.C0 + c]);
          } else {
            const size_t pidx = ((size_t)b * a.H + yy) * a.W + xx;
            v = (c < a.C0) ? to_f32<Tin>(in0[pidx * a.C0 + c]) : to_f32<Tin>(in1[pidx * C1 + (c - a.C0)]);
          }
        }
        in_s[p][ci] = v;
      }
      for (int e = t; e < 9 * CK * TN; e += 256) {
        const int n = e % TN, ci = (e / TN) % CK, tap = e / (TN * CK);
        w_s[tap][ci][n] = (ci < cw) ? a.w[((size_t)tap * a.Ctot + c0 + ci) * a.Cout + n0 + n] : 0.f;
      }
      __syncthreads();
      for (int ci = 0; ci < cw; ++ci) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const float xin = in_s[(py + tap / 3) * (TP + 2) + px + tap % 3][ci];
          const float4 w0 = *reinterpret_cast<const float4*>(&w_s[tap][ci][tq * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&w_s[tap][ci][tq * 8 + 4]);
          acc[0] = fmaf(xin, w0.x, acc[0]);
          acc[1] = fmaf(xin, w0.y, acc[1]);
          acc[2] = fmaf(xin, w0.z, acc[2]);
          acc[3] = fmaf(xin, w0.w, acc[3]);
          acc[4] = fmaf(xin, w1.x, acc[4]);
          acc[5] = fmaf(xin, w1.y, acc[5]);
          acc[6] = fmaf(xin, w1.z, acc[6]);
          acc[7] = fmaf(xin, w1.w, acc[7]);
        }
      }
    }
    const int yy = y0 + py, xx = x0 + px;
    if (yy < a.H && xx < a.W) {
      const int n = n0 + tq * 8;
      if (a.mode != EPI_LINEAR) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j] + a.bias[n + j], a.floor);
      }
      size_t pidx = ((size_t)b * a.H + yy) * a.W + xx;
      if (a.out_stuffed) {
        if (!((yy & 1) && (xx & 1))) continue;
        pidx = ((size_t)b * (a.H >> 1) + (yy >> 1)) * (a.W >> 1) + (xx >> 1);
      }
      if (a.mode == EPI_LINEAR && n >= a.out_split)
        Vec8<Tout>::store(static_cast<Tout*>(a.out1) + pidx * (a.Cout - a.out_split) + (n - a.out_split), acc);
      else
        Vec8<Tout>::store(static_cast<Tout*>(a.out0) + pidx * (a.mode == EPI_LINEAR ? a.out_split : a.Cout) + n, acc);
      if (a.mode == EPI_RELU_STATS) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = round_to<Tout>(acc[j]);
          atomicAdd(&s_sum[tq * 8 + j], v);
          atomicAdd(&s_sq[tq * 8 + j], v * v);
        }
      }
    }
  }
  if (a.mode == EPI_RELU_STATS) {
    __syncthreads();
    if (t < TN) {
      atomicAdd(&a.stats[n0 + t], (double)s_sum[t]);
      atomicAdd(&a.stats[a.Cout + n0 + t], (double)s_sq[t]);
    }
  }
}

template <typename Tin, typename Tout>
static int launch_simt(const ConvSimtArgs& a, cudaStream_t st) {
  const int tiles = ((a.W + TP - 1) / TP) * ((a.H + TP - 1) / TP) * a.B;
  const int ny = a.Cout / TN;
  int gx = tiles;
  const int cap = (kNumSMs * 8 + ny - 1) / ny;
  if (gx > cap) gx = cap;
  conv3x3_simt_kernel<Tin, Tout><<<dim3(gx, ny), 256, 0, st>>>(a);
  RVIP_LAUNCH_CHECK();
  return 0;
}

int conv_simt_launch(const ConvSimtArgs& a, int in_is_bf16, int out_is_bf16, cudaStream_t st) {
  RVIP_REQUIRE(a.Cout % TN == 0, "conv_simt: Cout=%d must be a multiple of %d", a.Cout, TN);
  RVIP_REQUIRE(a.mode != EPI_LINEAR || (a.out_split % 8 == 0), "conv_simt: bad out_split %d", a.out_split);
  if (!in_is_bf16 && !out_is_bf16) return launch_simt<float, float>(a, st);
  if (!in_is_bf16 && out_is_bf16) return launch_simt<float, __nv_bfloat16>(a, st);
  if (in_is_bf16 && out_is_bf16) return launch_simt<__nv_bfloat16, __nv_bfloat16>(a, st);
  set_error("conv_simt: unsupported dtype combination");
  return 1;
}

// ------------------------------------------------------------------------------------- wgrad
template <typename Tin, typename Tdz>
__global__ void __launch_bounds__(256) wgrad3x3_simt_kernel(WgradSimtArgs a) {
  __shared__ float in_s[(TP + 2) * (TP + 2)][CK + 1];
  __shared__ __align__(8) float dz_s[TP * TP][TN];
  const Tin* in0 = static_cast<const Tin*>(a.in0);
  const Tin* in1 = static_cast<const Tin*>(a.in1);
  const Tdz* dz = static_cast<const Tdz*>(a.dz);
  const int t = threadIdx.x;
  const int ci = t >> 4, co = (t & 15) * 2;
  const int c0 = blockIdx.x * CK, n0 = blockIdx.y * TN;
  const int cw = (a.Ctot - c0) < CK ? (a.Ctot - c0) : CK;
  const int C1 = a.Ctot - a.C0;
  const int tiles_x = (a.W + TP - 1) / TP, tiles_y = (a.H + TP - 1) / TP;
  const int n_tiles = tiles_x * tiles_y * a.B;
  float acc[9][2];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k][0] = acc[k][1] = 0.f;

  for (int tile = blockIdx.z; tile < n_tiles; tile += gridDim.z) {
    const int bx = tile % tiles_x, by = (tile / tiles_x) % tiles_y, b = tile / (tiles_x * tiles_y);
    const int x0 = bx * TP, y0 = by * TP;
    __syncthreads();
    for (int e = t; e < (TP + 2) * (TP + 2) * CK; e += 256) {
      const int cc = e % CK, p = e / CK;
      const int yy = y0 + p / (TP + 2) - 1, xx = x0 + p % (TP + 2) - 1;
      float v = 0.f;
      const int c = c0 + cc;
      if (cc < cw && yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) {
        if (a.in_stuffed) {
          if ((yy & 1) && (xx & 1))
            v = to_f32<Tin>(in0[(((size_t)b * (a.H >> 1) + (yy >> 1)) * (a.W >> 1) + (xx >> 1)) * a.C0 + c]);
        } else {
          const size_t pidx = ((size_t)b * a.H + yy) * a.W + xx;
          v = (c < a.C0) ? to_f32<Tin>(in0[pidx * a.C0 + c]) : to_f32<Tin>(in1[pidx * C1 + (c - a.C0)]);
        }
      }
      in_s[p][cc] = v;
    }
    for (int e = t; e < TP * TP * TN; e += 256) {
      const int n = e % TN, p = e / TN;
      const int yy = y0 + p / TP, xx = x0 + p % TP;
      float v = 0.f;
      if (yy < a.H && xx < a.W) v = to_f32<Tdz>(dz[(((size_t)b * a.H + yy) * a.W + xx) * a.Cout + n0 + n]);
      dz_s[p][n] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int p = 0; p < TP * TP; ++p) {
      const float2 d = *reinterpret_cast<const float2*>(&dz_s[p][co]);
      const int base = (p / TP) * (TP + 2) + (p % TP);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const float xin = in_s[base + (tap / 3) * (TP + 2) + tap % 3][ci];
        acc[tap][0] = fmaf(xin, d.x, acc[tap][0]);
        acc[tap][1] = fmaf(xin, d.y, acc[tap][1]);
      }
    }
  }
  if (ci < cw) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      if (a.transposed_out) {
        // dK[8 - tap][co][ci] of the Conv2DTranspose kernel (kh, kw, out, in)
        float* dst = a.dw + ((size_t)(8 - tap) * a.Cout + n0 + co) * a.Ctot + c0 + ci;
        atomicAdd(dst, acc[tap][0]);
        atomicAdd(dst + a.Ctot, acc[tap][1]);
        continue;
      }
      float* dst = a.dw + ((size_t)tap * a.Ctot + c0 + ci) * a.Cout + n0 + co;
      atomicAdd(dst, acc[tap][0]);
      atomicAdd(dst + 1, acc[tap][1]);
    }
  }
}

int wgrad_simt_launch(const WgradSimtArgs& a, int in_is_bf16, int dz_is_bf16, cudaStream_t st) {
  RVIP_REQUIRE(a.Cout % TN == 0, "wgrad_simt: Cout=%d must be a multiple of %d", a.Cout, TN);
  const int tiles = ((a.W + TP - 1) / TP) * ((a.H + TP - 1) / TP) * a.B;
  const int gx = (a.Ctot + CK - 1) / CK, gy = a.Cout / TN;
  int gz = (kNumSMs * 4 + gx * gy - 1) / (gx * gy);
  if (gz > tiles) gz = tiles;
  if (gz < 1) gz = 1;
  dim3 grid(gx, gy, gz);
  if (!in_is_bf16 && !dz_is_bf16)
    wgrad3x3_simt_kernel<float, float><<<grid, 256, 0, st>>>(a);
  else if (!in_is_bf16 && dz_is_bf16)
    wgrad3x3_simt_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(a);
  else if (in_is_bf16 && dz_is_bf16)
    wgrad3x3_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(a);
  else {
    set_error("wgrad_simt: unsupported dtype combination");
    return 1;
  }
  RVIP_LAUNCH_CHECK();
  return 0;
}

}  // namespace rvip
