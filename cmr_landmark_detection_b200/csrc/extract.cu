// Heat map -> landmark extraction (warp-level reductions, integer accumulators => bit-exact).
// Reference steps replaced:
//   threshold -> label map    src/models/predict_model.py:153-156
//       flat = 0; flat[p[...,0] > thr] = 1; flat[p[...,1] > thr] = 2   (later channel wins, strict >,
//       NaN compares false) => label (c+1) set = {p_c > thr and no later channel > thr}
//   per-slice centroid        src/models/evaluate_cv.py:418-442 (get_mean_rvip_2d): mean (row, col)
//       of each label's pixels, float64
//   argmax / max per channel  (no reference symbol): numpy.argmax semantics --
//       first occurrence in row-major order, NaN counts as the maximum.
// Pass 1 streams the fp32 heat map once (float2/float4 loads, one slice segment per block) and folds
// count / sum(row) / sum(col) as 64-bit integers and argmax as a packed (ordered value, ~index) key;
// pass 2 (one thread per slice-channel) turns the accumulators into coordinates.
#include "kernels.cuh"

namespace rvip {

constexpr int kMaxC = 4;

size_t extract_scratch_bytes(int Z, int C) { return (size_t)Z * C * 4 * sizeof(unsigned long long); }

__device__ __forceinline__ uint32_t order_key(float v) {
  if (v != v) return 0xFFFFFFFFu;  // NaN == maximum (numpy.argmax)
  uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int o) {
  uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
  lo = __shfl_xor_sync(0xffffffffu, lo, o);
  hi = __shfl_xor_sync(0xffffffffu, hi, o);
  return ((unsigned long long)hi << 32) | lo;
}

template <int C>
__global__ void __launch_bounds__(256) extract_accum_kernel(const float* __restrict__ heat, int H, int W, float thr,
                                                            unsigned long long* __restrict__ acc) {
  const int z = blockIdx.y;
  const int npix = H * W;
  const int seg = (((npix + gridDim.x - 1) / gridDim.x) + 1) & ~1;   // even: segments start on a pixel pair
  const int p0 = blockIdx.x * seg;
  const int p1 = min(npix, p0 + seg);
  const float* base = heat + (size_t)z * npix * C;
  unsigned long long cnt[C], sr[C], sc[C], best[C];
#pragma unroll
  for (int c = 0; c < C; ++c) cnt[c] = sr[c] = sc[c] = best[c] = 0ull;
  // one pixel: threshold predicates (later channel wins), integer row / column sums, packed argmax key.  The
  // row = p / W division only runs for labelled pixels (a few dozen per slice).
  auto visit = [&](int p, const float (&v)[C]) {
    int label = -1;
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (v[c] > thr) label = c;  // later channel overwrites
    if (label >= 0) {
      const int row = p / W, col = p - row * W;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (label == c) {
          cnt[c] += 1;
          sr[c] += row;
          sc[c] += col;
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const unsigned long long key = ((unsigned long long)order_key(v[c]) << 32) | (0xFFFFFFFFu - (uint32_t)p);
      best[c] = key > best[c] ? key : best[c];
    }
  };
  if (C == 2 && (npix & 1) == 0 && (p0 & 1) == 0) {
    // fast path: 16-byte loads (two pixels), four of them in flight per thread -- with 8-byte loads and no
    // unrolling this scan ran at 39 % of the HBM roofline, latency bound
    const float4* b4 = reinterpret_cast<const float4*>(base);
    const int q1 = p1 >> 1;                      // pixel pairs [q0, q1); an odd last pixel is handled below
    int q = (p0 >> 1) + threadIdx.x;
    for (; q + 3 * 256 < q1; q += 4 * 256) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = __ldcs(b4 + q + u * 256);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float va[2] = {t[u].x, t[u].y}, vb[2] = {t[u].z, t[u].w};
        visit(2 * (q + u * 256), reinterpret_cast<const float(&)[C]>(va));
        visit(2 * (q + u * 256) + 1, reinterpret_cast<const float(&)[C]>(vb));
      }
    }
    for (; q < q1; q += 256) {
      const float4 t = __ldcs(b4 + q);
      const float va[2] = {t.x, t.y}, vb[2] = {t.z, t.w};
      visit(2 * q, reinterpret_cast<const float(&)[C]>(va));
      visit(2 * q + 1, reinterpret_cast<const float(&)[C]>(vb));
    }
    if ((p1 & 1) && threadIdx.x == 0) {
      const float vl[2] = {base[(size_t)(p1 - 1) * 2], base[(size_t)(p1 - 1) * 2 + 1]};
      visit(p1 - 1, reinterpret_cast<const float(&)[C]>(vl));
    }
  } else {
    for (int p = p0 + threadIdx.x; p < p1; p += 256) {
      float v[C];
      if (C == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(base) + p);
        v[0] = t.x;
        v[1] = t.y;
      } else if (C == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(base) + p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int c = 0; c < C; ++c) v[c] = __ldg(base + (size_t)p * C + c);
      }
      visit(p, v);
    }
  }
  __shared__ unsigned long long s_acc[8][C][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < C; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt[c] += shfl_xor_u64(cnt[c], o);
      sr[c] += shfl_xor_u64(sr[c], o);
      sc[c] += shfl_xor_u64(sc[c], o);
      const unsigned long long other = shfl_xor_u64(best[c], o);
      best[c] = other > best[c] ? other : best[c];
    }
    if (lane == 0) {
      s_acc[warp][c][0] = cnt[c];
      s_acc[warp][c][1] = sr[c];
      s_acc[warp][c][2] = sc[c];
      s_acc[warp][c][3] = best[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    unsigned long long a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int w = 0; w < 8; ++w) {
      a0 += s_acc[w][c][0];
      a1 += s_acc[w][c][1];
      a2 += s_acc[w][c][2];
      a3 = s_acc[w][c][3] > a3 ? s_acc[w][c][3] : a3;
    }
    unsigned long long* dst = acc + ((size_t)z * C + c) * 4;
    if (a0) {
      atomicAdd(dst + 0, a0);
      atomicAdd(dst + 1, a1);
      atomicAdd(dst + 2, a2);
    }
    atomicMax(dst + 3, a3);
  }
}

__global__ void extract_finalize_kernel(const float* __restrict__ heat, const unsigned long long* __restrict__ acc,
                                        int Z, int H, int W, int C, double* yx, int* count, int* argmax, float* maxv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Z * C) return;
  const unsigned long long n = acc[i * 4 + 0];
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  count[i] = (int)n;
  yx[i * 2 + 0] = n ? (double)acc[i * 4 + 1] / (double)n : nan;
  yx[i * 2 + 1] = n ? (double)acc[i * 4 + 2] / (double)n : nan;
  const uint32_t idx = 0xFFFFFFFFu - (uint32_t)(acc[i * 4 + 3] & 0xFFFFFFFFull);
  argmax[i] = (int)idx;
  const int z = i / C, c = i % C;
  maxv[i] = heat[((size_t)z * H * W + idx) * C + c];
}

__global__ void label_map_kernel(const float* __restrict__ heat, size_t n_pix, int C, float thr,
                                 uint8_t* __restrict__ out) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix; p += (size_t)gridDim.x * blockDim.x) {
    uint8_t label = 0;
    for (int c = 0; c < C; ++c)
      if (__ldg(heat + p * C + c) > thr) label = (uint8_t)(c + 1);  // later channel wins (predict_model.py:155-156)
    out[p] = label;
  }
}
int label_map_launch(const float* heat, size_t n_pix, int C, float thr, uint8_t* out, cudaStream_t st) {
  if (n_pix == 0) return 0;
  size_t g = (n_pix + 255) / 256;
  if (g > (size_t)kNumSMs * 16) g = (size_t)kNumSMs * 16;
  label_map_kernel<<<(unsigned)g, 256, 0, st>>>(heat, n_pix, C, thr, out);
  RVIP_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------- per-volume landmark metrics
// The step after extraction (src/models/evaluate_cv.py): septum angle per slice (get_angle2x :508-536), distances to
// the ground truth with spacing / threshold (get_distances :549-561) and with the missing-prediction upper bound
// (get_distances_upper_bound :572-595), mean insertion points (calc_mean_ip :113-120) and the TP / FN / FP counters
// behind calc_tpr_thresh :267-308 / calc_ppv_thresh :311-353.  One block per call, one thread per slice, float64.
// Points are (y, x) pairs, NaN = missing (the reference's None).
__device__ __forceinline__ bool pt_ok(const double* p) { return isfinite(p[0]) && isfinite(p[1]); }

__global__ void __launch_bounds__(256) landmark_metrics_kernel(const double* __restrict__ gt, const double* __restrict__ pr,
                                                               int Z, double spacing, double thr, double dim,
                                                               double* __restrict__ angle, double* __restrict__ dist,
                                                               double* __restrict__ dist_thr, double* __restrict__ dist_ub,
                                                               double* __restrict__ summary) {
  // shared accumulators: [which (0 gt, 1 pred)][landmark][y, x, n]  then counters [landmark][tp, fn, fp]
  __shared__ double s_sum[2][2][3];
  __shared__ unsigned int s_cnt[2][3];
  if (threadIdx.x < 12) (&s_sum[0][0][0])[threadIdx.x] = 0.0;
  if (threadIdx.x < 6) (&s_cnt[0][0])[threadIdx.x] = 0u;
  __syncthreads();
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (int z = threadIdx.x; z < Z; z += blockDim.x) {
    const double* g = gt + (size_t)z * 4;   // [landmark][y, x]
    const double* p = pr + (size_t)z * 4;
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const double* a = w == 0 ? g : p;
      double ang = nan;
      if (pt_ok(a) && pt_ok(a + 2)) {
        ang = atan2(a[2] - a[0], a[3] - a[1]) * (180.0 / 3.14159265358979323846);
        if (ang < 0) ang = 360.0 + ang;
      }
      angle[(size_t)w * Z + z] = ang;
    }
#pragma unroll
    for (int l = 0; l < 2; ++l) {
      const bool hg = pt_ok(g + 2 * l), hp = pt_ok(p + 2 * l);
      double d = nan, dt = nan, du = nan;
      if (hg && hp) {
        const double dy = g[2 * l] - p[2 * l], dx = g[2 * l + 1] - p[2 * l + 1];
        d = sqrt(dy * dy + dx * dx) * spacing;
        dt = d <= thr ? d : nan;
        du = d;
        atomicAdd(&s_cnt[l][d <= thr ? 0 : 2], 1u);
      } else if (hg) {
        // farthest corner of the dim x dim image
        const double y = g[2 * l], x = g[2 * l + 1];
        const double my = fmax(fabs(y), fabs(y - dim)), mx = fmax(fabs(x), fabs(x - dim));
        du = sqrt(my * my + mx * mx) * spacing;
        atomicAdd(&s_cnt[l][1], 1u);
      } else if (hp) {
        atomicAdd(&s_cnt[l][2], 1u);
      }
      dist[(size_t)l * Z + z] = d;
      dist_thr[(size_t)l * Z + z] = dt;
      dist_ub[(size_t)l * Z + z] = du;
      if (hg) {
        atomicAdd(&s_sum[0][l][0], g[2 * l]);
        atomicAdd(&s_sum[0][l][1], g[2 * l + 1]);
        atomicAdd(&s_sum[0][l][2], 1.0);
      }
      if (hp) {
        atomicAdd(&s_sum[1][l][0], p[2 * l]);
        atomicAdd(&s_sum[1][l][1], p[2 * l + 1]);
        atomicAdd(&s_sum[1][l][2], 1.0);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // summary: mean points [which][landmark][2] (8), tpr[2], ppv[2], counters [landmark][tp, fn, fp] (6)
    for (int w = 0; w < 2; ++w) {
      const bool both = s_sum[w][0][2] > 0 && s_sum[w][1][2] > 0;
      for (int l = 0; l < 2; ++l)
        for (int k = 0; k < 2; ++k) summary[(w * 2 + l) * 2 + k] = both ? s_sum[w][l][k] / s_sum[w][l][2] : nan;
    }
    for (int l = 0; l < 2; ++l) {
      const double tp = s_cnt[l][0], fn = s_cnt[l][1], fp = s_cnt[l][2];
      summary[8 + l] = tp > 0 ? tp / (tp + fn) : 0.0;
      summary[10 + l] = tp > 0 ? tp / (tp + fp) : 0.0;
      summary[12 + l * 3 + 0] = tp;
      summary[12 + l * 3 + 1] = fn;
      summary[12 + l * 3 + 2] = fp;
    }
  }
}
int landmark_metrics_launch(const double* gt, const double* pred, int Z, double spacing, double thr, double dim,
                            double* angle, double* dist, double* dist_thr, double* dist_ub, double* summary,
                            cudaStream_t st) {
  RVIP_REQUIRE(Z >= 0, "landmark_metrics: bad slice count");
  landmark_metrics_kernel<<<1, 256, 0, st>>>(gt, pred, Z, spacing, thr, dim, angle, dist, dist_thr, dist_ub, summary);
  RVIP_LAUNCH_CHECK();
  return 0;
}

int extract_launch(const float* heat, int Z, int H, int W, int C, float thr, double* yx, int* count, int* argmax,
                   float* maxv, unsigned long long* scratch, cudaStream_t st) {
  RVIP_REQUIRE(C >= 1 && C <= kMaxC, "extract: C=%d not in [1,%d]", C, kMaxC);
  RVIP_REQUIRE(Z >= 0 && H > 0 && W > 0 && (long long)H * W < 0x7fffffffLL, "extract: bad shape");
  if (Z == 0) return 0;
  RVIP_CUDA(cudaMemsetAsync(scratch, 0, extract_scratch_bytes(Z, C), st));
  const int npix = H * W;
  // enough segments to fill the chip, at least 1024 pixels per block
  int segs = (2 * kNumSMs + Z - 1) / Z;
  const int max_segs = (npix + 1023) / 1024;
  if (segs > max_segs) segs = max_segs;
  if (segs < 1) segs = 1;
  dim3 grid(segs, Z);
  switch (C) {
    case 1: extract_accum_kernel<1><<<grid, 256, 0, st>>>(heat, H, W, thr, scratch); break;
    case 2: extract_accum_kernel<2><<<grid, 256, 0, st>>>(heat, H, W, thr, scratch); break;
    case 3: extract_accum_kernel<3><<<grid, 256, 0, st>>>(heat, H, W, thr, scratch); break;
    default: extract_accum_kernel<4><<<grid, 256, 0, st>>>(heat, H, W, thr, scratch); break;
  }
  RVIP_LAUNCH_CHECK();
  extract_finalize_kernel<<<(Z * C + 127) / 128, 128, 0, st>>>(heat, scratch, Z, H, W, C, yx, count, argmax, maxv);
  RVIP_LAUNCH_CHECK();
  return 0;
}

}  // namespace rvip
