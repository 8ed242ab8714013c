// Adam (tf.keras.optimizers.Adam defaults, src/models/ModelUtils.py:107) over the flat fp32
// parameter / gradient / moment buffers, and the re-pack of the fp32 master weights (Keras HWIO)
// into the K-major bf16 operand layouts the tensor-core kernels read.
#include "kernels.cuh"

namespace rvip {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, size_t n, float lr_t,
                                                   float b1, float b2, float eps, float gs) {
  pdl_wait();
  const size_t n4 = n / 4;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define RVIP_ADAM(x)                              \
  {                                               \
    const float gr = gg.x * gs;                   \
    mm.x = b1 * mm.x + (1.f - b1) * gr;           \
    vv.x = b2 * vv.x + (1.f - b2) * gr * gr;      \
    pp.x -= lr_t * mm.x / (sqrtf(vv.x) + eps);    \
  }
    RVIP_ADAM(x) RVIP_ADAM(y) RVIP_ADAM(z) RVIP_ADAM(w)
#undef RVIP_ADAM
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  pdl_launch_dependents();
  if (blockIdx.x == 0) {
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += 256) {
      const float gr = g[i] * gs;
      m[i] = b1 * m[i] + (1.f - b1) * gr;
      v[i] = b2 * v[i] + (1.f - b2) * gr * gr;
      p[i] -= lr_t * m[i] / (sqrtf(v[i]) + eps);
    }
  }
}
int adam_launch(float* p, const float* g, float* m, float* v, size_t n, float lr_t, float b1, float b2, float eps,
                float grad_scale, cudaStream_t st) {
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks > (size_t)kNumSMs * 8) blocks = (size_t)kNumSMs * 8;
  if (blocks < 1) blocks = 1;
  launch_kernel(adam_kernel, (unsigned)blocks, 256, 0, st, p, g, m, v, n, lr_t, b1, b2, eps, grad_scale);
  RVIP_LAUNCH_CHECK();
  return 0;
}

// tf.keras.optimizers.SGD (the optimizer the reference switches to in finetune_with_SGD, utils/KerasCallbacks.py:280-306,
// and OPTIMIZER='sgd', ModelUtils.py:109-111):  v = momentum * v - lr * g;  w += nesterov ? momentum * v - lr * g : v;
// with momentum == 0 (both call sites) no velocity buffer is needed: w -= lr * g.
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ vel,
                                                  size_t n, float lr, float momentum, int nesterov, float gs) {
  pdl_wait();
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float gr = g[i] * gs;
    if (vel != nullptr) {
      const float v = momentum * vel[i] - lr * gr;
      vel[i] = v;
      p[i] += nesterov ? momentum * v - lr * gr : v;
    } else {
      p[i] -= lr * gr;
    }
  }
  pdl_launch_dependents();
}
int sgd_launch(float* p, const float* g, float* vel, size_t n, float lr, float momentum, int nesterov, float grad_scale,
               cudaStream_t st) {
  size_t blocks = (n + 255) / 256;
  if (blocks > (size_t)kNumSMs * 8) blocks = (size_t)kNumSMs * 8;
  if (blocks < 1) blocks = 1;
  launch_kernel(sgd_kernel, (unsigned)blocks, 256, 0, st, p, g, vel, n, lr, momentum, nesterov, grad_scale);
  RVIP_LAUNCH_CHECK();
  return 0;
}

// packed forward  operand: Wf[n][tap][c]  = W[tap][c][n]           (rows = output channels, K-major)
// packed dgrad    operand: Wd[c][tap'][n] = W[8 - tap'][c][n]      (rows = input channels, taps rotated)
// fp32 dgrad copy (CUDA-core path): Wr[tap'][n][c] = W[8 - tap'][c][n]
template <typename TO, bool BF16>
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ params, TO* __restrict__ packed,
                                                           const PackEntry* __restrict__ table) {
  pdl_wait();
  const PackEntry e = table[blockIdx.y];
  const long long n = 9LL * e.Ctot * e.Cout;
  const float* W = params + e.src;
  // i enumerates DESTINATION elements so the narrow stores coalesce; the strided fp32 reads hit L2
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    if (BF16) {
      {  // Wf[co][tap][c]
        const int c = (int)(i % e.Ctot), tap = (int)((i / e.Ctot) % 9), co = (int)(i / (9LL * e.Ctot));
        packed[e.dst_f + i] = from_f32<TO>(W[((long long)tap * e.Ctot + c) * e.Cout + co]);
      }
      if (e.dst_d >= 0) {  // Wd[c][tap'][co] = W[8 - tap'][c][co]
        const int co = (int)(i % e.Cout), tp = (int)((i / e.Cout) % 9), c = (int)(i / (9LL * e.Cout));
        packed[e.dst_d + i] = from_f32<TO>(W[((long long)(8 - tp) * e.Ctot + c) * e.Cout + co]);
      }
    } else if (e.dst_d >= 0) {  // Wr[tap'][co][c] = W[8 - tap'][c][co]
      const int c = (int)(i % e.Ctot), co = (int)((i / e.Ctot) % e.Cout), tp = (int)(i / ((long long)e.Ctot * e.Cout));
      packed[e.dst_d + i] = from_f32<TO>(W[((long long)(8 - tp) * e.Ctot + c) * e.Cout + co]);
    }
  }
}
// bf16 re-pack through a shared-memory tile transpose: one coalesced read of a 32 (c) x 32 (co) tile of W[tap]
// feeds BOTH operand copies -- Wd rows (c) take 32 consecutive co straight from the tile, Wf rows (co) take 32
// consecutive c from its transpose.  (The gather version above read W twice, once with a Cout-float stride: 8x sector
// amplification in L2, 95 us per step.)  Channel counts on this path are multiples of 32.
__global__ void __launch_bounds__(256) pack_weights_tile_kernel(const float* __restrict__ params,
                                                                __nv_bfloat16* __restrict__ packed,
                                                                const PackEntry* __restrict__ table) {
  pdl_wait();
  __shared__ float tile[32][33];
  const PackEntry e = table[blockIdx.y];
  const float* W = params + e.src;
  const int ct_n = e.Ctot >> 5, cot_n = e.Cout >> 5;
  const int n_tiles = 9 * ct_n * cot_n;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 rows per pass
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const int cot = t % cot_n, ct = (t / cot_n) % ct_n, tap = t / (cot_n * ct_n);
    const int c0 = ct << 5, co0 = cot << 5;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int c = ty + 8 * r;
      tile[c][tx] = W[((long long)tap * e.Ctot + c0 + c) * e.Cout + co0 + tx];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = ty + 8 * r;
      // Wf[co][tap][c]: row = co, tx = c
      packed[e.dst_f + ((long long)(co0 + row) * 9 + tap) * e.Ctot + c0 + tx] = __float2bfloat16_rn(tile[tx][row]);
      // Wd[c][tap'][co] = W[8 - tap'][c][co]: row = c, tx = co
      if (e.dst_d >= 0)
        packed[e.dst_d + ((long long)(c0 + row) * 9 + (8 - tap)) * e.Cout + co0 + tx] = __float2bfloat16_rn(tile[row][tx]);
    }
    __syncthreads();
  }
  pdl_launch_dependents();
}

// Phase-decomposed up-convolution operands (conv_halo.cuh).  With ky, kx the 3x3 tap indices (0..2 = offsets -1..+1)
// and S(p, 0) = {0} / {0, 1}, S(p, 1) = {1, 2} / {2} for phase parity p = 0 / 1 the taps that fall on the low-resolution
// neighbour 0 / 1 of the phase:   Wp[a][b][r][s][ci][co] = sum_{ky in S(a,r)} sum_{kx in S(b,s)} W[ky][kx][ci][co]  (fp32 sums,
// one rounding to bf16).
//   forward copy  [phase][row][tap][ci]:  ns = 2: phase = 2a+b, row = co, tap = 2r+s;
//                                         ns = 3: phase = a, row = b*C+co, tap = 3r+tx with s = tx - b (zero if outside)
//   dgrad copy    [ci][phase][tap][cz]:   halo taps run in increasing offset: rr = 1-r, ss = 1-s (ns = 2);
//                                         ns = 3: cz = b*C+co, tap = 3rr+tx with s = 2 - b - tx (zero if outside)
// One block pass = one 32 (ci) x 32 (co) tile position: the nine 3x3 taps of it are read ONCE (coalesced) into registers,
// the 16 pre-summed phase weights Wp[a][b][r][s] are formed from them, and each goes through a shared-memory tile so that
// both operand copies are written coalesced -- the dgrad copy (co fastest) straight from it, the forward copy (ci
// fastest) from its transpose.  The structural zeros of the ns = 3 copies are never written: the packed buffer is
// cleared once when the handle is bound.
__global__ void __launch_bounds__(256) pack_up_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ packed,
                                                      const UpPackEntry* __restrict__ table) {
  pdl_wait();
  __shared__ float tile[2][32][33];
  const UpPackEntry e = table[blockIdx.y];
  const float* W = params + e.src;
  const int Cin = e.Cin, C = e.C;
  const int ct_n = Cin >> 5, cot_n = C >> 5;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int t = blockIdx.x; t < ct_n * cot_n; t += gridDim.x) {
    const int cot = t % cot_n, ct = t / cot_n;
    const int c0 = ct << 5, co0 = cot << 5;
    // tile rows (ty + 8q) / columns (tx): (ci, co) for a Conv2D kernel (kh, kw, Cin, C), (co, ci) for a Conv2DTranspose
    // kernel (kh, kw, C, Cin) -- the contiguous axis of the source is always the column
    const bool tr = e.transposed != 0;
    float w[9][4];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        w[tap][q] = tr ? W[((long long)tap * C + co0 + ty + 8 * q) * Cin + c0 + tx]
                       : W[((long long)tap * Cin + c0 + ty + 8 * q) * C + co0 + tx];
    int nbuf = 0;
#pragma unroll
    for (int combo = 0; combo < 16; ++combo) {
      const int a = combo >> 3, b = (combo >> 2) & 1, r = (combo >> 1) & 1, s = combo & 1;
      const int ky0 = (a == 0) ? (r == 0 ? 0 : 1) : (r == 0 ? 0 : 2), ky1 = (a == 0) ? (r == 0 ? 0 : 2) : (r == 0 ? 1 : 2);
      const int kx0 = (b == 0) ? (s == 0 ? 0 : 1) : (s == 0 ? 0 : 2), kx1 = (b == 0) ? (s == 0 ? 0 : 2) : (s == 0 ? 1 : 2);
      // Conv2DTranspose(3, strides 2, 'same'): out[2i + a] = sum_{2i' + k = 2i + a} x[i'] w[k] -> parity 0 sees tap 2 on
      // neighbour i - 1 (r = 0) and tap 0 on i (r = 1), parity 1 sees tap 1 on i (r = 0) and nothing on i + 1
      const int kyt = a == 0 ? (r == 0 ? 2 : 0) : (r == 0 ? 1 : -1), kxt = b == 0 ? (s == 0 ? 2 : 0) : (s == 0 ? 1 : -1);
      if (tr && (kyt < 0 || kxt < 0)) continue;   // structurally zero (cleared at bind time); uniform over the block
      float (*tl)[33] = tile[nbuf & 1];      // double-buffered over the combinations actually written: one barrier each
      ++nbuf;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float acc = 0.f;
        if (tr) {
          acc = w[(kyt < 0 ? 0 : kyt) * 3 + (kxt < 0 ? 0 : kxt)][q];
        } else {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              if (ky >= ky0 && ky <= ky1 && kx >= kx0 && kx <= kx1) acc += w[ky * 3 + kx][q];
        }
        tl[ty + 8 * q][tx] = acc;
      }
      __syncthreads();
      long long f_base, d_base;      // element offsets of (row = co0, ci = c0) / (ci = c0, cz = co0) in the two copies
      long long f_row, d_row;        // strides between consecutive rows
      if (e.ns == 2) {
        const int ph = 2 * a + b;
        f_row = 4LL * Cin;           // [ph][co][t][ci]
        f_base = ((long long)(ph * C + co0) * 4 + (2 * r + s)) * Cin + c0;
        d_row = 16LL * C;            // [ci][ph][t'][co]
        d_base = ((long long)c0 * 4 + ph) * 4 * C + (long long)(2 * (1 - r) + (1 - s)) * C + co0;
      } else {
        f_row = 6LL * Cin;           // [a][b*C+co][t = 3r + s + b][ci]
        f_base = ((long long)(a * 2 * C + b * C + co0) * 6 + (3 * r + s + b)) * Cin + c0;
        d_row = 12LL * 2 * C;        // [ci][a][t' = 3(1-r) + 2-b-s][b*C+co]
        d_base = ((long long)c0 * 2 + a) * 6 * 2 * C + (long long)(3 * (1 - r) + (2 - b - s)) * 2 * C + b * C + co0;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int row = ty + 8 * q;
        // forward copy: row = co, tx = ci; dgrad copy: row = ci, tx = co
        packed[e.dst_f + f_base + row * f_row + tx] = __float2bfloat16_rn(tr ? tl[row][tx] : tl[tx][row]);
        if (e.dst_d >= 0) packed[e.dst_d + d_base + row * d_row + tx] = __float2bfloat16_rn(tr ? tl[tx][row] : tl[row][tx]);
      }
    }
    __syncthreads();
  }
  pdl_launch_dependents();
}
int pack_up_launch(const float* params, void* packed, const UpPackEntry* table_dev, int n_entries, cudaStream_t st) {
  if (n_entries == 0) return 0;
  launch_kernel(pack_up_kernel, dim3(128, n_entries), 256, 0, st, params, static_cast<__nv_bfloat16*>(packed), table_dev);
  RVIP_LAUNCH_CHECK();
  return 0;
}

int pack_weights_launch(const float* params, void* packed, const PackEntry* table_dev, int n_entries, int to_bf16,
                        cudaStream_t st) {
  if (n_entries == 0) return 0;
  dim3 grid(148, n_entries);
  if (to_bf16)
    launch_kernel(pack_weights_tile_kernel, dim3(74, n_entries), 256, 0, st, params, static_cast<__nv_bfloat16*>(packed),
                  table_dev);
  else
    launch_kernel(pack_weights_kernel<float, false>, grid, 256, 0, st, params, static_cast<float*>(packed), table_dev);
  RVIP_LAUNCH_CHECK();
  return 0;
}

}  // namespace rvip
