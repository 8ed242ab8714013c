// Head (1x1 conv -> sigmoid, src/models/Unets.py:128) fused with the heat-map loss and its gradient:
//   mse ...................... tf.keras.losses.mse (src/models/Loss_and_metrics.py:6): mean over every element
//   masked / weighted ........ loss_with_zero_mask (src/models/Loss_and_metrics.py:40-89)
// One pass reads y (and the target), writes the fp32 heat map, dL/dy for the last decoder block, and
// reduces the loss, dW_head and db_head with warp shuffles -> shared atomics -> one global atomic per
// block and output.  G = Cin/CPT adjacent lanes share a pixel (CPT = 8 or 16 channels per thread), so all global traffic
// is coalesced.  The kernel is ISSUE bound (ncu: 62 % issue slots busy at 25 % occupancy, DRAM at 24 %): every lane of a
// pixel repeats the sigmoid / loss / dlogit arithmetic, so CPT = 16 (two lanes per pixel at 32 channels) halves that
// redundancy and the logit shuffles.
#include "kernels.cuh"
#include "tc_prims.cuh"

#include <stdlib.h>

namespace rvip {

constexpr int kMaxNC = 4;

// VAR: tuning variant of the software pipeline / occupancy (0 = default: depth 2 at 2 blocks/SM for CPT = 16)
template <typename T, bool TRAIN, int NCT, int CPT, bool FOLD, int VAR = 0>
__global__ void __launch_bounds__(VAR == 3 ? 384 : 256, (CPT == 16 && VAR < 2) ? 2 : 1) head_kernel(HeadArgs a) {
  constexpr int NT = VAR == 3 ? 384 : 256;   // threads per block = items per tile
  constexpr int NV = CPT / 8;              // 8-channel vectors per thread
  extern __shared__ float sm[];            // w [Cin*NC], b [NC], dw acc [Cin*NC], db acc [NC], (FOLD) scale, shift [Cin]
  pdl_wait();
  constexpr int NC = NCT;                  // the launcher instantiates the exact class count
  const int Cin = a.Cin, G = Cin / CPT;
  float* w_s = sm;
  float* b_s = w_s + Cin * NC;
  float* dw_s = b_s + NC;
  float* db_s = dw_s + Cin * NC;
  float* sc_s = db_s + NC;
  float* sh_s = sc_s + Cin;
  __shared__ double loss_s;
  for (int k = threadIdx.x; k < Cin * NC; k += NT) {
    w_s[k] = a.w[k];
    if (TRAIN) dw_s[k] = 0.f;
  }
  if (threadIdx.x < NC) {
    b_s[threadIdx.x] = a.b[threadIdx.x];
    if (TRAIN) db_s[threadIdx.x] = 0.f;
  }
  if (threadIdx.x == 0) loss_s = 0.0;
  if (FOLD) {
    // BatchNorm of the producing block, exactly as bn_apply_kernel's prologue derives it (same float expressions, so
    // the backward pass sees the mean / rstd the forward used)
    for (int k = threadIdx.x; k < Cin; k += NT) {
      const double mean = a.bn_stats[k] * a.bn_inv_count;
      double var = a.bn_stats[Cin + k] * a.bn_inv_count - mean * mean;
      if (var < 0) var = 0;
      const float m = (float)mean, r = rsqrtf((float)var + a.bn_eps);
      if (blockIdx.x == 0 && a.bn_publish) {
        a.bn_mean_out[k] = m;
        a.bn_rstd_out[k] = r;
        const double unb = a.bn_count > 1.0 ? var * a.bn_count / (a.bn_count - 1.0) : var;
        a.bn_mov_mean[k] = a.bn_momentum * a.bn_mov_mean[k] + (1.f - a.bn_momentum) * m;
        a.bn_mov_var[k] = a.bn_momentum * a.bn_mov_var[k] + (1.f - a.bn_momentum) * (float)unb;
      }
      const float sck = a.bn_gamma[k] * r;
      sc_s[k] = sck;
      sh_s[k] = fmaf(-m, sck, a.bn_beta[k]);
    }
  }
  __syncthreads();
  if (FOLD) {
    // logits = sum_c w[c][k] (sc[c] a[c] + sh[c]) + b[k]: the shift folds into the bias
    if (threadIdx.x < NC) {
      float s = b_s[threadIdx.x];
      for (int c2 = 0; c2 < Cin; ++c2) s = fmaf(sh_s[c2], w_s[c2 * NC + threadIdx.x], s);
      b_s[threadIdx.x] = s;
    }
    __syncthreads();
  }

  const uint32_t P = (uint32_t)a.B * a.H * a.W;
  const uint32_t lg = 31 - __clz(G);
  const uint32_t n_items = P << lg;
  const uint32_t i0 = blockIdx.x * NT + threadIdx.x;
  const int cg = (int)(i0 & (G - 1)), c = cg * CPT;
  const uint32_t HW = (uint32_t)a.H * a.W;
  const T* y = static_cast<const T*>(a.y);
  T* dy = static_cast<T*>(a.dy);
  float wreg[CPT][NCT];
#pragma unroll
  for (int j = 0; j < CPT; ++j)
#pragma unroll
    for (int k = 0; k < NCT; ++k) wreg[j][k] = k < NC ? w_s[(c + j) * NC + k] * (FOLD ? sc_s[c + j] : 1.f) : 0.f;
  float dw_acc[CPT][NCT], db_acc[NCT];
#pragma unroll
  for (int k = 0; k < NCT; ++k) {
    db_acc[k] = 0.f;
#pragma unroll
    for (int j = 0; j < CPT; ++j) dw_acc[j][k] = 0.f;
  }
  float loss_acc = 0.f;
  const float inv_n = 1.f / ((float)P * (float)NC);
  float dice_D = 1.f, dice_num = 0.f, dice_invD2 = 0.f;
  if (TRAIN && a.loss_kind == LOSS_BCE_DICE) {
    const double I = a.dice_sums[0], D = a.dice_sums[1] + a.dice_sums[2] + 1.0;
    dice_D = (float)D;
    dice_num = (float)(2.0 * I + 1.0);
    dice_invD2 = (float)(1.0 / (D * D));
  }
  // all lanes of a warp run the same trip count (n_items and the stride are multiples of 32)
  const uint32_t n_round = (n_items + 31) / 32 * 32;
  const uint32_t stride = gridDim.x * NT;
  // Software pipeline: the loads of y and of the target for the item kDepth grid-strides ahead are in flight
  // while the current item is processed (ncu: with the loads issued at the point of use, 55 % of the stall
  // samples of this kernel were long-scoreboard waits on exactly those two loads).
  constexpr int kDepth = VAR == 1 ? 1 : (VAR == 2 ? 4 : (VAR == 3 ? 1 : (CPT == 8 ? 4 : 2)));
  Raw8<T> ybuf[kDepth][NV];
  float tbuf[kDepth][NCT];
  auto issue = [&](int d, uint32_t i) {
    if (i < n_items) {
      const size_t p = i >> lg;
#pragma unroll
      for (int u = 0; u < NV; ++u) load_raw8(y + p * Cin + c + 8 * u, ybuf[d][u]);
      if (TRAIN) {
#pragma unroll
        for (int k = 0; k < NCT; ++k) tbuf[d][k] = k < NC ? a.target[p * NC + k] : 0.f;
      }
    } else {
#pragma unroll
      for (int u = 0; u < NV; ++u) zero_raw8(ybuf[d][u]);
#pragma unroll
      for (int k = 0; k < NCT; ++k) tbuf[d][k] = 0.f;
    }
  };
  // one item: logits (lanes of a pixel combine by shuffle), sigmoid, heat map, loss terms, dL/dy, dW / db partials
  auto process = [&](bool live, size_t p, const float (&v)[CPT], const float (&tgt)[NCT]) {
      float logit[NCT];
#pragma unroll
      for (int k = 0; k < NCT; ++k) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < CPT; ++j) s = fmaf(v[j], wreg[j][k], s);
        for (int o = 1; o < G; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        logit[k] = s + (k < NC ? b_s[k] : 0.f);
      }
      if (!live) return;
      float prob[NCT];
#pragma unroll
      for (int k = 0; k < NCT; ++k) prob[k] = 1.f / (1.f + expf(-logit[k]));
      if (cg == 0) {
#pragma unroll
        for (int k = 0; k < NCT; ++k)
          if (k < NC) a.heat[p * NC + k] = prob[k];
      }
      if (TRAIN) {
        float wpx = 1.f;
        bool any = false;
#pragma unroll
        for (int k = 0; k < NCT; ++k) any = any || (k < NC && tgt[k] > a.mask_thr);
        if (a.loss_kind == LOSS_MASKED || a.loss_kind == LOSS_WEIGHTED) wpx = any ? 1.f : 0.f;
        if (a.loss_kind == LOSS_WEIGHTED) wpx *= a.inplane[(uint32_t)p % HW];
        float dl[NCT], se = 0.f;
        if (a.loss_kind == LOSS_BCE_DICE) {
          // w_bce * mean_c BCE(t, clip(p)) - w_dice * dice;  d dice / d p_i = (2 t_i D - (2 I + 1)) / D^2
#pragma unroll
          for (int k = 0; k < NCT; ++k) {
            if (k < NC) {
              const float pc = fminf(fmaxf(prob[k], a.eps), 1.f - a.eps);
              const bool inside = prob[k] > a.eps && prob[k] < 1.f - a.eps;     // clip passes no gradient outside
              se -= tgt[k] * logf(pc + a.eps) + (1.f - tgt[k]) * logf(1.f - pc + a.eps);
              const float dbce = inside ? (-tgt[k] / (pc + a.eps) + (1.f - tgt[k]) / (1.f - pc + a.eps)) : 0.f;
              const float ddice = (2.f * tgt[k] * dice_D - dice_num) * dice_invD2;
              dl[k] = (a.w_bce * inv_n * dbce - a.w_dice * ddice) * prob[k] * (1.f - prob[k]);
            } else {
              dl[k] = 0.f;
            }
          }
          se *= a.w_bce;                    // the common accumulation below divides by NC (mean over channels)
        } else {
#pragma unroll
          for (int k = 0; k < NCT; ++k) {
            const float d2 = k < NC ? prob[k] - tgt[k] : 0.f;
            se = fmaf(d2, d2, se);
            dl[k] = 2.f * d2 * wpx * inv_n * prob[k] * (1.f - prob[k]);
          }
        }
        if (cg == 0) {
          loss_acc += se / (float)NC * wpx + (a.loss_kind == LOSS_WEIGHTED ? a.eps : 0.f);
#pragma unroll
          for (int k = 0; k < NCT; ++k) db_acc[k] += dl[k];
        }
#pragma unroll
        for (int u = 0; u < NV; ++u) {
          float g[8];
          // dy = dL/d(BN output) needs the PLAIN head weights; when folded wreg carries the BatchNorm scale, so they
          // come from shared memory (vector loads, broadcast: all lanes of a channel group read the same address)
          float wd[8][NCT];
          if (FOLD) {
            if constexpr (NC == 2) {
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                const float4 t4 = *reinterpret_cast<const float4*>(w_s + (c + 8 * u + j) * 2);
                wd[j][0] = t4.x; wd[j][1] = t4.y; wd[j + 1][0] = t4.z; wd[j + 1][1] = t4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int k = 0; k < NCT; ++k) wd[j][k] = w_s[(c + 8 * u + j) * NC + k];
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < NCT; ++k) {
              s = fmaf(FOLD ? wd[j][k] : wreg[8 * u + j][k], dl[k], s);
              dw_acc[8 * u + j][k] = fmaf(v[8 * u + j], dl[k], dw_acc[8 * u + j][k]);
            }
            g[j] = s;
          }
          Vec8<T>::store(dy + p * Cin + c + 8 * u, g);
        }
      }
  };
  if constexpr (VAR == 3) {
    using namespace tc;
    // Shared-memory staged variant: one thread keeps kSlots tiles (NT items = NT / G consecutive pixels: 12 KB of y and
    // their targets, both contiguous) in flight with 1-D bulk copies, so no register holds a prefetch: 160 registers
    // instead of 233, which buys 12 warps per SM instead of 8 (the kernel is bound by instruction latency at this
    // occupancy, not by its loads: staging alone, at 8 warps, measured no change).
    constexpr int kSlots = 6;
    constexpr uint32_t kYBytes = (uint32_t)NT * CPT * sizeof(T);
    const uint32_t tile_px = (uint32_t)NT >> lg;
    const uint32_t t_full = TRAIN ? tile_px * NC * (uint32_t)sizeof(float) : 0u;
    const uint32_t slot_bytes = kYBytes + ((t_full + 127u) & ~127u);
    __shared__ uint64_t full_bar[kSlots];
    uint8_t* ring = reinterpret_cast<uint8_t*>(sh_s + Cin);
    ring += (128u - (smem_u32(ring) & 127u)) & 127u;
    const uint32_t n_tiles = (n_items + NT - 1) / NT;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s2 = 0; s2 < kSlots; ++s2) mbar_init(&full_bar[s2], 1);
      fence_barrier_init();
    }
    __syncthreads();
    auto fill = [&](uint32_t k) {
      const uint32_t tile = blockIdx.x + k * gridDim.x;
      if (tile >= n_tiles) return;
      const uint32_t s2 = k % kSlots;
      uint8_t* dst = ring + s2 * slot_bytes;
      // the last tile may be partial: whole pixels (the launcher checks that their bytes are 16-byte multiples)
      const uint32_t px = min(tile_px, P - tile * tile_px);
      const uint32_t yb = px * Cin * (uint32_t)sizeof(T), tb = TRAIN ? px * NC * (uint32_t)sizeof(float) : 0u;
      mbar_expect_tx(&full_bar[s2], yb + tb);
      bulk_load_1d(dst, y + (size_t)tile * tile_px * Cin, yb, &full_bar[s2]);
      if (TRAIN) bulk_load_1d(dst + kYBytes, a.target + (size_t)tile * tile_px * NC, tb, &full_bar[s2]);
    };
    if (threadIdx.x == 0) {
      for (uint32_t k = 0; k < kSlots; ++k) fill(k);
    }
    for (uint32_t k = 0;; ++k) {
      const uint32_t tile = blockIdx.x + k * gridDim.x;
      if (tile >= n_tiles) break;
      const uint32_t s2 = k % kSlots;
      mbar_wait(&full_bar[s2], (k / kSlots) & 1u);
      const uint8_t* src = ring + s2 * slot_bytes;
      const uint32_t item = tile * NT + threadIdx.x;
      const bool live = item < n_items;
      const size_t p = live ? (item >> lg) : 0;
      float v[CPT], tgt[NCT];
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        Raw8<T> r8;
        if (live) load_raw8(reinterpret_cast<const T*>(src) + threadIdx.x * CPT + 8 * u, r8);
        else zero_raw8(r8);
        float t8[8];
        unpack_raw8(r8, t8);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * u + j] = t8[j];
      }
#pragma unroll
      for (int k2 = 0; k2 < NCT; ++k2)
        tgt[k2] = (TRAIN && live && k2 < NC) ? reinterpret_cast<const float*>(src + kYBytes)[(threadIdx.x >> lg) * NC + k2] : 0.f;
      __syncthreads();                       // every thread has its values: the slot may be refilled
      if (threadIdx.x == 0) fill(k + kSlots);
      process(live, p, v, tgt);
    }
  } else {
#pragma unroll
  for (int d = 0; d < kDepth; ++d) issue(d, i0 + d * stride);
  for (uint32_t ib = i0; ib < n_round; ib += kDepth * stride) {
#pragma unroll
    for (int d = 0; d < kDepth; ++d) {
      const uint32_t i = ib + d * stride;
      if (i >= n_round) break;           // warp-uniform
      const bool live = i < n_items;
      const size_t p = live ? (i >> lg) : 0;
      float v[CPT], tgt[NCT];
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        float t8[8];
        unpack_raw8(ybuf[d][u], t8);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 * u + j] = t8[j];
      }
#pragma unroll
      for (int k = 0; k < NCT; ++k) tgt[k] = tbuf[d][k];
      issue(d, i + kDepth * stride);
      process(live, p, v, tgt);
    }
  }
  }
  pdl_launch_dependents();
  if (TRAIN) {
    // lanes l, l+G, l+2G, ... hold the same channels: fold them with shuffles so that G lanes per warp (not 32)
    // touch the shared accumulators (float atomicAdd on smem is a CAS loop; 64-way contention cost ~100 us)
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < CPT; ++j)
#pragma unroll
      for (int k = 0; k < NCT; ++k) {
        float v = dw_acc[j][k];
        for (int o = 16; o >= G; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (k < NC && lane < G) atomicAdd(&dw_s[(c + j) * NC + k], v);
      }
#pragma unroll
    for (int k = 0; k < NCT; ++k) {
      const float v = warp_sum(db_acc[k]);     // only the cg == 0 lanes carry non-zero partials
      if (k < NC && lane == 0) atomicAdd(&db_s[k], v);
    }
    float l = warp_sum(loss_acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(&loss_s, (double)l);
    __syncthreads();
    for (int k = threadIdx.x; k < Cin * NC; k += NT) atomicAdd(&a.dw[k], dw_s[k]);
    if (threadIdx.x < NC) atomicAdd(&a.db[threadIdx.x], db_s[threadIdx.x]);
    if (threadIdx.x == 0) {
      double l = loss_s / (double)P;                                  // mean over B*H*W
      if (a.loss_kind == LOSS_BCE_DICE && blockIdx.x == 0)
        l -= (double)a.w_dice * (2.0 * a.dice_sums[0] + 1.0) / (a.dice_sums[1] + a.dice_sums[2] + 1.0);
      atomicAdd(a.loss_acc, l);
    }
  }
}

__global__ void __launch_bounds__(256) head_dice_sums_kernel(const float* __restrict__ heat, const float* __restrict__ tgt,
                                                             size_t n, double* __restrict__ sums) {
  float s_tp = 0.f, s_p = 0.f, s_t = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float p = heat[i], t = tgt[i];
    s_tp = fmaf(t, p, s_tp);
    s_p += p;
    s_t += t;
  }
  s_tp = warp_sum(s_tp); s_p = warp_sum(s_p); s_t = warp_sum(s_t);
  __shared__ double acc[3];
  if (threadIdx.x < 3) acc[threadIdx.x] = 0.0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&acc[0], (double)s_tp);
    atomicAdd(&acc[1], (double)s_p);
    atomicAdd(&acc[2], (double)s_t);
  }
  __syncthreads();
  if (threadIdx.x < 3) atomicAdd(&sums[threadIdx.x], acc[threadIdx.x]);
}
int head_dice_sums_launch(const float* heat, const float* target, size_t n, double* sums, cudaStream_t st) {
  head_dice_sums_kernel<<<kNumSMs * 2, 256, 0, st>>>(heat, target, n, sums);
  RVIP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------- folded-BatchNorm finalize
__global__ void head_bn_finalize_kernel(const float* __restrict__ dwa, const float* __restrict__ db,
                                        const float* __restrict__ w, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, int Cin, int NC, float* __restrict__ dw,
                                        double* __restrict__ red) {
  pdl_wait();
  for (int c = threadIdx.x; c < Cin; c += blockDim.x) {
    const float sc = gamma[c] * rstd[c];
    const float sh = fmaf(-mean[c], sc, beta[c]);
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < NC; ++k) {
      const float g = dwa[c * NC + k];
      dw[c * NC + k] = fmaf(sc, g, sh * db[k]);
      s1 += (double)w[c * NC + k] * (double)db[k];
      s2 += (double)w[c * NC + k] * (double)g;
    }
    red[c] = s1;
    red[Cin + c] = s2;
  }
  pdl_launch_dependents();
}
int head_bn_finalize_launch(const float* dwa, const float* db, const float* w, const float* gamma, const float* beta,
                            const float* mean, const float* rstd, int Cin, int NC, float* dw, double* red,
                            cudaStream_t st) {
  launch_kernel(head_bn_finalize_kernel, 1, 256, 0, st, dwa, db, w, gamma, beta, mean, rstd, Cin, NC, dw, red);
  RVIP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------- validation loss / Dice metrics
// One pass over a heat map and its target (both [P][NC] fp32): the sum of the per-pixel loss terms of the compiled loss
// (model.evaluate / validation_data of fit, inference-mode heat maps) and, per channel, {sum t*p, sum p, sum t} -- the
// three sums every dice_coef* metric of src/models/Loss_and_metrics.py:124-171 is made of.  out[0] += sum loss terms,
// out[1 + 3c ..] += channel c's sums (double).  The host divides (mean over pixels, Dice ratio): no torch arithmetic.
template <int NC>
__global__ void __launch_bounds__(256) heat_stats_kernel(const float* __restrict__ heat, const float* __restrict__ tgt,
                                                         const float* __restrict__ inplane, size_t P, uint32_t HW,
                                                         int loss_kind, float mask_thr, float eps,
                                                         double* __restrict__ out) {
  float l = 0.f, s_tp[NC], s_p[NC], s_t[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) s_tp[k] = s_p[k] = s_t[k] = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < P; i += (size_t)gridDim.x * 256) {
    float p[NC], t[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      p[k] = heat[i * NC + k];
      t[k] = tgt[i * NC + k];
      s_tp[k] = fmaf(t[k], p[k], s_tp[k]);
      s_p[k] += p[k];
      s_t[k] += t[k];
    }
    float se = 0.f;
    bool any = false;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      any = any || t[k] > mask_thr;
      if (loss_kind == LOSS_BCE_DICE) {
        const float pc = fminf(fmaxf(p[k], eps), 1.f - eps);
        se -= t[k] * logf(pc + eps) + (1.f - t[k]) * logf(1.f - pc + eps);
      } else {
        const float d = p[k] - t[k];
        se = fmaf(d, d, se);
      }
    }
    se *= 1.f / (float)NC;
    if (loss_kind == LOSS_MASKED || loss_kind == LOSS_WEIGHTED) se = any ? se : 0.f;
    if (loss_kind == LOSS_WEIGHTED) se = fmaf(se, inplane[i % HW], eps);
    l += se;
  }
  __shared__ double acc[1 + 3 * NC];
  if (threadIdx.x < 1 + 3 * NC) acc[threadIdx.x] = 0.0;
  __syncthreads();
  l = warp_sum(l);
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    s_tp[k] = warp_sum(s_tp[k]);
    s_p[k] = warp_sum(s_p[k]);
    s_t[k] = warp_sum(s_t[k]);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&acc[0], (double)l);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      atomicAdd(&acc[1 + 3 * k], (double)s_tp[k]);
      atomicAdd(&acc[2 + 3 * k], (double)s_p[k]);
      atomicAdd(&acc[3 + 3 * k], (double)s_t[k]);
    }
  }
  __syncthreads();
  if (threadIdx.x < 1 + 3 * NC) atomicAdd(&out[threadIdx.x], acc[threadIdx.x]);
}
int heat_stats_launch(const float* heat, const float* target, const float* inplane, size_t n_pixels, int HW, int NC,
                      int loss_kind, float mask_thr, double* out, cudaStream_t st) {
  RVIP_REQUIRE(NC >= 1 && NC <= kMaxNC, "heat_stats: %d channels not in [1,%d]", NC, kMaxNC);
  RVIP_REQUIRE(loss_kind != LOSS_WEIGHTED || inplane, "heat_stats: weighted loss needs in-plane weights");
  RVIP_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (1 + 3 * NC), st));
  size_t g = (n_pixels + 255) / 256;
  const int grid = (int)(g < (size_t)kNumSMs * 4 ? (g ? g : 1) : (size_t)kNumSMs * 4);
  switch (NC) {
    case 1: heat_stats_kernel<1><<<grid, 256, 0, st>>>(heat, target, inplane, n_pixels, HW, loss_kind, mask_thr, 1e-7f, out); break;
    case 2: heat_stats_kernel<2><<<grid, 256, 0, st>>>(heat, target, inplane, n_pixels, HW, loss_kind, mask_thr, 1e-7f, out); break;
    case 3: heat_stats_kernel<3><<<grid, 256, 0, st>>>(heat, target, inplane, n_pixels, HW, loss_kind, mask_thr, 1e-7f, out); break;
    default: heat_stats_kernel<4><<<grid, 256, 0, st>>>(heat, target, inplane, n_pixels, HW, loss_kind, mask_thr, 1e-7f, out); break;
  }
  RVIP_LAUNCH_CHECK();
  return 0;
}

int head_launch(const HeadArgs& a, int training, int is_bf16, cudaStream_t st) {
  RVIP_REQUIRE(a.Cin % 8 == 0 && a.Cin / 8 <= 32 && ((a.Cin / 8) & (a.Cin / 8 - 1)) == 0,
               "head: Cin=%d must be 8 * power of two <= 256", a.Cin);
  // 16 channels per thread where the class count keeps the per-thread weight / dW registers within budget
  const int cpt = (a.Cin % 16 == 0 && a.NC <= 2 && getenv("RVIP_HEAD_CPT8") == nullptr) ? 16 : 8;
  const int G = a.Cin / cpt;
  RVIP_REQUIRE((size_t)a.B * a.H * a.W * G < 0x7fffffffULL, "head: tensor too large for 32-bit indexing");
  RVIP_REQUIRE(a.NC >= 1 && a.NC <= kMaxNC, "head: MASK_CLASSES=%d not in [1,%d]", a.NC, kMaxNC);
  const size_t n = (size_t)a.B * a.H * a.W * G;
  size_t g = (n + 255) / 256;
  const size_t cap = (size_t)kNumSMs * 2;   // 2 blocks/SM resident (122 registers); every block ends with Cin*NC + NC + 1 global atomics: keep them few
  const int grid = (int)(g < cap ? (g ? g : 1) : cap);
  const size_t smem = (size_t)(2 * (a.Cin * a.NC + a.NC) + 2 * a.Cin) * sizeof(float);
  const bool fold = a.bn_stats != nullptr;
  // folded training head, 2 classes x 16 channels per thread: 128 registers spill (200 B) at 2 blocks/SM; one block per SM
  // with a 4-deep load pipeline measured 108 us against 156 (depth 2, 2 blocks) and 128 (depth 1), profiles/r2d_headvar.jsonl
  // (variant 3: shared-memory staged loads, 384 threads at 156 registers: 96 us, profiles/r2zc_head_sweep.jsonl)
  int var = getenv("RVIP_HEAD_VAR") ? atoi(getenv("RVIP_HEAD_VAR")) : 3;
  const int grid1 = (int)(g < (size_t)kNumSMs ? (g ? g : 1) : (size_t)kNumSMs);
  // variant 3 stages whole 384-item tiles in shared memory with bulk copies (16-byte granules, whole pixels)
  const size_t tile_px3 = 384 / G;
  const size_t g3 = (n + 383) / 384;
  const int grid3 = (int)(g3 < (size_t)kNumSMs ? (g3 ? g3 : 1) : (size_t)kNumSMs);
  if (var == 3 && (!is_bf16 || 384 % G != 0 || (a.NC * sizeof(float)) % 16 != 0 && ((size_t)a.B * a.H * a.W) % 2 != 0)) var = 2;
  const size_t smem3 = smem + 128 + 6 * (384 * 16 * 2 + ((tile_px3 * a.NC * sizeof(float) + 127) & ~size_t(127)));
  if (var == 3) {
    static bool attr_set = false;
    if (!attr_set) {
      RVIP_CUDA(cudaFuncSetAttribute(head_kernel<__nv_bfloat16, true, 2, 16, true, 3>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr_set = true;
    }
  }
#define RVIP_HEAD_T(TT, NCV, CPTV)                                                             \
  {                                                                                            \
    if (training && fold && NCV == 2 && CPTV == 16 && var == 1)                                \
      launch_kernel(head_kernel<TT, true, 2, 16, true, 1>, grid, 256, smem, st, a);            \
    else if (training && fold && NCV == 2 && CPTV == 16 && var == 2)                           \
      launch_kernel(head_kernel<TT, true, 2, 16, true, 2>, grid1, 256, smem, st, a);           \
    else if (training && fold && NCV == 2 && CPTV == 16 && var == 3)                           \
      launch_kernel(head_kernel<TT, true, 2, 16, true, 3>, grid3, 384, smem3, st, a);           \
    else if (training && fold) launch_kernel(head_kernel<TT, true, NCV, CPTV, true>, grid, 256, smem, st, a);         \
    else if (training) launch_kernel(head_kernel<TT, true, NCV, CPTV, false>, grid, 256, smem, st, a);           \
    else if (fold) launch_kernel(head_kernel<TT, false, NCV, CPTV, true>, grid, 256, smem, st, a);               \
    else launch_kernel(head_kernel<TT, false, NCV, CPTV, false>, grid, 256, smem, st, a);                        \
  }
#define RVIP_HEAD(NCV, CPTV)                          \
  if (a.NC == NCV && cpt == CPTV) {                   \
    if (is_bf16) RVIP_HEAD_T(__nv_bfloat16, NCV, CPTV) \
    else RVIP_HEAD_T(float, NCV, CPTV)                \
  }
  RVIP_HEAD(1, 8) RVIP_HEAD(2, 8) RVIP_HEAD(3, 8) RVIP_HEAD(4, 8) RVIP_HEAD(1, 16) RVIP_HEAD(2, 16)
#undef RVIP_HEAD
#undef RVIP_HEAD_T
  RVIP_LAUNCH_CHECK();
  return 0;
}

}  // namespace rvip
