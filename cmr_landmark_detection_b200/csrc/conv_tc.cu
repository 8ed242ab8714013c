// tcgen05 / TMEM / TMA implicit-GEMM 3x3 convolution for sm_100a.
//
// Replaces what the reference reaches through tf.keras Conv2D (src/models/KerasLayers.py:683,689,758)
// and its autodiff (Conv2DBackpropInput / Conv2DBackpropFilter):
//
//   conv3x3_tc_kernel  forward and dgrad.  GEMM  D[128 pixels, BN] = A[128, K] * B[BN, K]^T with
//     K = 9 taps x channels.  A tiles are NOT im2col'ed: for every tap the TMA engine fetches the
//     shifted NHWC box {KC ch, TW, TH, NB} (out-of-bounds rows/cols zero-filled == 'same' padding)
//     straight into the 128B/64B-swizzled K-major layout tcgen05.mma consumes.  A second tensor map
//     supplies the skip connection's channels, so Concatenate never materialises.
//     Warp roles: warp0 = TMA producer, warp1 = MMA issuer (one thread), warp2 = TMEM allocator,
//     warps4-7 = epilogue (tcgen05.ld -> bias/ReLU -> bf16 -> swizzled smem -> TMA store, plus
//     per-channel sum / sum-of-squares for the BatchNorm that follows).  Accumulators are double
//     buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1; the kernel is
//     persistent (grid = min(#tiles, 148)).
//
//   wgrad3x3_tc_kernel  dW[tap][ci][co] = sum_p x[p+off(tap), ci] * dz[p, co].  Both operands are
//     used MN-major exactly as TMA lands them ([pixel rows] x [channels contiguous]); the M
//     dimension packs (tap, ci) chunks, pixels are the GEMM K dimension, split across CTAs and
//     reduced with vector fp32 atomics straight into the HWIO gradient buffer.
#include "conv_tc.cuh"

#include "common.cuh"
#include "tc_prims.cuh"

namespace rvip {
using namespace tc;

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}
static int pow2_divisor(int v) {
  int p = 1;
  while (v % (p * 2) == 0) p *= 2;
  return p;
}

TileGeom make_tile_geom(int B, int H, int W, int P) {
  TileGeom g;
  int tw = pow2_divisor(W);
  if (tw > P) tw = P;
  if (tw < 8 && W > tw) tw = pow2_floor(W) < P ? pow2_floor(W) : P;
  g.TW = tw;
  int th = P / tw;
  if (th > pow2_floor(H)) th = pow2_floor(H);
  g.TH = th;
  g.NB = P / (tw * th);
  g.tiles_x = (W + g.TW - 1) / g.TW;
  g.tiles_y = (H + g.TH - 1) / g.TH;
  g.tiles_b = (B + g.NB - 1) / g.NB;
  g.full = (W % g.TW == 0) && (H % g.TH == 0) && (B % g.NB == 0);
  return g;
}

constexpr int kMaxDynSmem = 227 * 1024;
constexpr int kMaxStages = 8;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint8_t* align_1024(uint8_t* p) {
  uint32_t a = smem_u32(p);
  return p + ((1024u - (a & 1023u)) & 1023u);
}

struct SmemCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
  uint8_t valid[128];
};

// =====================================================================================
// forward / dgrad
// =====================================================================================
template <int KC, int BN>
struct FwdCfg {
  static constexpr int ROWB = KC * 2;
  static constexpr int A_BYTES = 128 * ROWB;
  static constexpr int B_BYTES = BN * ROWB;
  static constexpr int STAGE = A_BYTES + B_BYTES;
  static constexpr int OCH = BN >= 64 ? 64 : 32;  // channels per TMA store box
  static constexpr int OROWB = OCH * 2;
  static constexpr int OCHUNK_BYTES = 128 * OROWB;
  static constexpr int STAGING = 128 * BN * 2;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr uint64_t LAYOUT = KC == 64 ? kLayoutSW128 : kLayoutSW64;
};

size_t conv_tc_smem_bytes(int KC, int BN, int Cout) {
  // stages are added by the launcher; this is the fixed part
  return 1024 + (size_t)128 * BN * 2 + (size_t)5 * Cout * sizeof(float) + sizeof(SmemCtl) + 64;
}

template <int KC, int BN>
__global__ void __launch_bounds__(256, 1) conv3x3_tc_kernel(const __grid_constant__ ConvTcArgs a, int nst) {
  using Cfg = FwdCfg<KC, BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  uint8_t* staging = smem + (size_t)nst * Cfg::STAGE;
  float* s_sum = reinterpret_cast<float*>(staging + Cfg::STAGING);
  float* s_sq = s_sum + a.Cout;
  float* s_bias = s_sq + a.Cout;     // [3][Cout]: bias, scale, shift
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>((reinterpret_cast<uintptr_t>(s_bias + 3 * a.Cout) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TileGeom g = a.g;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&a.in0);
    prefetch_tmap(&a.in1);
    prefetch_tmap(&a.w);
    prefetch_tmap(&a.out0);
    prefetch_tmap(&a.out1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < nst; ++i) {
      mbar_init(&ctl->full[i], 1);
      mbar_init(&ctl->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->tfull[i], 1);
      mbar_init(&ctl->tempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&ctl->tmem_base, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  pdl_wait();   // everything above is independent of the previous kernel's output
  if (warp >= 4) {
    for (int c = threadIdx.x - 128; c < 2 * a.Cout; c += 128) s_sum[c] = 0.f;
    // bias in shared memory: a per-element __ldg in the epilogue exposes an L2 latency per 8 channels
    for (int c = threadIdx.x - 128; c < a.Cout; c += 128) {
      s_bias[c] = a.mode != EPI_LINEAR ? a.bias[c] : 0.f;
      s_bias[a.Cout + c] = a.mode == EPI_RELU_AFFINE ? a.scale[c] : 1.f;
      s_bias[2 * a.Cout + c] = a.mode == EPI_RELU_AFFINE ? a.shift[c] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const int iters = 9 * (a.Ctot / KC);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int nt = tile % a.n_ntiles, pt = tile / a.n_ntiles;
        const int x0 = (pt % g.tiles_x) * g.TW;
        const int y0 = ((pt / g.tiles_x) % g.tiles_y) * g.TH;
        const int b0 = (pt / (g.tiles_x * g.tiles_y)) * g.NB;
        const int n0 = nt * BN;
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          for (int c = 0; c < a.Ctot; c += KC) {
            mbar_wait(&ctl->empty[stage], phase ^ 1);
            mbar_expect_tx(&ctl->full[stage], Cfg::STAGE);
            uint8_t* A = smem + (size_t)stage * Cfg::STAGE;
            if (c < a.C0)
              tma_load_4d(A, &a.in0, &ctl->full[stage], c, x0 + dx, y0 + dy, b0);
            else
              tma_load_4d(A, &a.in1, &ctl->full[stage], c - a.C0, x0 + dx, y0 + dy, b0);
            tma_load_2d(A + Cfg::A_BYTES, &a.w, &ctl->full[stage], tap * a.Ctot + c, n0);
            if (++stage == nst) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      int stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        mbar_wait(&ctl->tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * BN;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&ctl->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * Cfg::STAGE);
          const uint64_t adesc = make_smem_desc(a_addr, 16, 8 * Cfg::ROWB, Cfg::LAYOUT);
          const uint64_t bdesc = make_smem_desc(a_addr + Cfg::A_BYTES, 16, 8 * Cfg::ROWB, Cfg::LAYOUT);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k)
            mma_bf16_ss(d, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
          mma_commit(&ctl->empty[stage]);
          if (++stage == nst) {
            stage = 0;
            phase ^= 1;
          }
        }
        mma_commit(&ctl->tfull[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;        // TMEM lane quarter == warp % 4
    const int r = ew * 32 + lane;   // accumulator row == pixel slot inside the tile
    const int et = threadIdx.x - 128;
    int acc = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int nt = tile % a.n_ntiles, pt = tile / a.n_ntiles;
      const int x0 = (pt % g.tiles_x) * g.TW;
      const int y0 = ((pt / g.tiles_x) % g.tiles_y) * g.TH;
      const int b0 = (pt / (g.tiles_x * g.tiles_y)) * g.NB;
      const int n0 = nt * BN;
      bool valid = true;
      if (!g.full) {
        const int xl = r % g.TW, yl = (r / g.TW) % g.TH, bl = r / (g.TW * g.TH);
        valid = (x0 + xl < a.W) && (y0 + yl < a.H) && (b0 + bl < a.B);
      }
      // staging buffer must be free: previous TMA store has read it, previous stats pass is done
      if (et == 0) tma_store_wait_read0();
      named_bar_sync(1, 128);
      ctl->valid[r] = valid ? 1 : 0;

      mbar_wait(&ctl->tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + acc * BN + ch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            f[j] = __uint_as_float(v[q * 8 + j]);
            if (a.mode != EPI_LINEAR) f[j] = fmaxf(f[j] + s_bias[n0 + ch * 32 + q * 8 + j], a.floor);
          }
          if (a.mode == EPI_RELU_AFFINE) {   // inference only
            const float* sa = s_bias + a.Cout + n0 + ch * 32 + q * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sa[j], sa[a.Cout + j]);
          }
          uint4 pk;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
          const int col = ch * 32 + q * 8;
          const int oc = col / Cfg::OCH, cidx = (col % Cfg::OCH) / 8;
          *reinterpret_cast<uint4*>(staging + oc * Cfg::OCHUNK_BYTES + swz_off<Cfg::OROWB>(r, cidx)) = pk;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->tempty[acc]);  // accumulator may be overwritten by tile i+2
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (et == 0) {
#pragma unroll 1
        for (int oc = 0; oc < BN / Cfg::OCH; ++oc) {
          const int n = n0 + oc * Cfg::OCH;
          if (a.mode == EPI_LINEAR && n >= a.out_split)
            tma_store_4d(&a.out1, staging + oc * Cfg::OCHUNK_BYTES, n - a.out_split, x0, y0, b0);
          else
            tma_store_4d(&a.out0, staging + oc * Cfg::OCHUNK_BYTES, n, x0, y0, b0);
        }
        tma_store_commit();
      }
      if (a.mode == EPI_RELU_STATS) {
        // per-channel sum / sum^2 over the tile's valid rows, from the bf16 values just staged
        constexpr int G8 = BN / 8;     // 16-byte column groups per row
        constexpr int RT = 128 / G8;   // threads sharing one column group
        const int cg = et % G8, rt = et / G8;
        float s[8], q2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] = q2[j] = 0.f;
        const int col = cg * 8;
        const int oc = col / Cfg::OCH, cidx = (col % Cfg::OCH) / 8;
#pragma unroll 1
        for (int k = 0; k < G8; ++k) {
          const int row = rt + k * RT;
          if (!ctl->valid[row]) continue;
          uint4 raw = *reinterpret_cast<const uint4*>(staging + oc * Cfg::OCHUNK_BYTES + swz_off<Cfg::OROWB>(row, cidx));
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f = __bfloat1622float2(h[j]);
            s[2 * j] += f.x;
            s[2 * j + 1] += f.y;
            q2[2 * j] += f.x * f.x;
            q2[2 * j + 1] += f.y * f.y;
          }
        }
        // lanes with equal (lane % G8) hold the same channels: fold them before touching smem
        if (G8 < 32) {
#pragma unroll
          for (int o = 16; o >= G8; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
              q2[j] += __shfl_xor_sync(0xffffffffu, q2[j], o);
            }
          }
        }
        if (G8 >= 32 || lane < G8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            atomicAdd(&s_sum[n0 + col + j], s[j]);
            atomicAdd(&s_sq[n0 + col + j], q2[j]);
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (et == 0) tma_store_wait_all0();
    if (a.mode == EPI_RELU_STATS) {
      named_bar_sync(1, 128);
      for (int c = et; c < a.Cout; c += 128) {
        atomicAdd(&a.stats[c], (double)s_sum[c]);
        atomicAdd(&a.stats[a.Cout + c], (double)s_sq[c]);
      }
    }
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int KC, int BN>
static int launch_fwd(const ConvTcArgs& a, cudaStream_t st) {
  using Cfg = FwdCfg<KC, BN>;
  size_t fixed = conv_tc_smem_bytes(KC, BN, a.Cout);
  int nst = (int)((kMaxDynSmem - fixed) / Cfg::STAGE);
  if (nst > kMaxStages) nst = kMaxStages;
  RVIP_REQUIRE(nst >= 2, "conv_tc: not enough shared memory for KC=%d BN=%d Cout=%d", KC, BN, a.Cout);
  size_t smem = fixed + (size_t)nst * Cfg::STAGE;
  static bool attr_set = false;
  if (!attr_set) {
    RVIP_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<KC, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kMaxDynSmem));
    attr_set = true;
  }
  int grid = a.total_tiles < kNumSMs ? a.total_tiles : kNumSMs;
  launch_kernel(conv3x3_tc_kernel<KC, BN>, dim3(grid), dim3(256), smem, st, a, nst);
  RVIP_LAUNCH_CHECK();
  return 0;
}

int conv_tc_launch(const ConvTcArgs& a, int KC, int BN, cudaStream_t st) {
  RVIP_REQUIRE(a.Ctot % KC == 0 && a.C0 % KC == 0, "conv_tc: channels (%d,%d) not divisible by KC=%d", a.C0,
               a.Ctot, KC);
  RVIP_REQUIRE(a.Cout % BN == 0, "conv_tc: Cout=%d not divisible by BN=%d", a.Cout, BN);
#define RVIP_FWD_CASE(kc, bn) \
  if (KC == kc && BN == bn) return launch_fwd<kc, bn>(a, st);
  RVIP_FWD_CASE(64, 32)
  RVIP_FWD_CASE(64, 64)
  RVIP_FWD_CASE(64, 128)
  RVIP_FWD_CASE(64, 256)
  RVIP_FWD_CASE(32, 32)
  RVIP_FWD_CASE(32, 64)
  RVIP_FWD_CASE(32, 128)
  RVIP_FWD_CASE(32, 256)
#undef RVIP_FWD_CASE
  set_error("conv_tc: unsupported tile KC=%d BN=%d", KC, BN);
  return 1;
}

// =====================================================================================
// wgrad
// =====================================================================================
constexpr int KP = 64;  // pixels per K tile

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <int CBA, int CBB>
__global__ void __launch_bounds__(256, 1) wgrad3x3_tc_kernel(const __grid_constant__ WgradTcArgs a, int nst,
                                                             int tmem_cols) {
  constexpr int A_CHUNK = KP * CBA * 2;  // bytes of one TMA box of the A operand
  constexpr int B_CHUNK = KP * CBB * 2;
  constexpr int CPT = 128 / CBA;         // A chunks per 128-row M tile
  constexpr int A_TILE = CPT * A_CHUNK;  // == KP * 128 * 2
  constexpr uint64_t LAYA = CBA == 64 ? kLayoutSW128 : kLayoutSW64;
  constexpr uint64_t LAYB = CBB == 64 ? kLayoutSW128 : kLayoutSW64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_1024(smem_raw);
  const int BN = a.BN, MT = a.MT;
  const int stage_bytes = MT * A_TILE + KP * BN * 2;
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem + (size_t)nst * stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TileGeom g = a.g;
  const int mg = blockIdx.x % a.n_mgroups, nt = blockIdx.x / a.n_mgroups;
  const int mt0 = mg * MT;
  const int mcount = (a.n_mtiles - mt0) < MT ? (a.n_mtiles - mt0) : MT;
  const int n0 = nt * BN;
  const int k_per = (a.k_tiles + a.n_split - 1) / a.n_split;
  const int kt0 = blockIdx.y * k_per;
  const int kt1 = (kt0 + k_per) < a.k_tiles ? (kt0 + k_per) : a.k_tiles;
  const int cpc = a.Ctot / CBA;   // chunks per tap
  const int nchunk = 9 * cpc;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&a.x0);
    prefetch_tmap(&a.x1);
    prefetch_tmap(&a.dz);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < nst; ++i) {
      mbar_init(&ctl->full[i], 1);
      mbar_init(&ctl->empty[i], 1);
    }
    mbar_init(&ctl->tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&ctl->tmem_base, tmem_cols);
    tmem_relinquish();
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (kt0 < kt1) {
    if (warp == 0) {
      if (elect_one()) {
        int stage = 0, phase = 0;
        const uint32_t tx_bytes = mcount * A_TILE + KP * BN * 2;
        for (int kt = kt0; kt < kt1; ++kt) {
          const int x0 = (kt % g.tiles_x) * g.TW;
          const int y0 = ((kt / g.tiles_x) % g.tiles_y) * g.TH;
          const int b0 = (kt / (g.tiles_x * g.tiles_y)) * g.NB;
          mbar_wait(&ctl->empty[stage], phase ^ 1);
          mbar_expect_tx(&ctl->full[stage], tx_bytes);
          uint8_t* Asm = smem + (size_t)stage * stage_bytes;
          uint8_t* Bsm = Asm + MT * A_TILE;
          for (int j = 0; j < BN / CBB; ++j)
            tma_load_4d(Bsm + j * B_CHUNK, &a.dz, &ctl->full[stage], n0 + j * CBB, x0, y0, b0);
          for (int i = 0; i < mcount; ++i) {
            for (int j = 0; j < CPT; ++j) {
              int q = (mt0 + i) * CPT + j;
              if (q >= nchunk) q = 0;  // padding rows of the last M tile: any valid data, results dropped
              const int tap = q / cpc, c = (q % cpc) * CBA;
              const int dy = tap / 3 - 1, dx = tap % 3 - 1;
              uint8_t* dst = Asm + (i * CPT + j) * A_CHUNK;
              if (c < a.C0)
                tma_load_4d(dst, &a.x0, &ctl->full[stage], c, x0 + dx, y0 + dy, b0);
              else
                tma_load_4d(dst, &a.x1, &ctl->full[stage], c - a.C0, x0 + dx, y0 + dy, b0);
            }
          }
          if (++stage == nst) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
        int stage = 0, phase = 0;
        for (int kt = kt0; kt < kt1; ++kt) {
          mbar_wait(&ctl->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t b_addr = a_addr + MT * A_TILE;
          for (int i = 0; i < mcount; ++i) {
#pragma unroll
            for (int kk = 0; kk < KP / 16; ++kk) {
              // MN-major operands: LBO = distance between channel chunks, SBO = between 8-pixel groups
              const uint64_t adesc = make_smem_desc(a_addr + i * A_TILE + kk * 16 * CBA * 2, A_CHUNK, 8 * CBA * 2, LAYA);
              const uint64_t bdesc = make_smem_desc(b_addr + kk * 16 * CBB * 2, B_CHUNK, 8 * CBB * 2, LAYB);
              mma_bf16_ss(tmem_base + i * BN, adesc, bdesc, idesc, (kt > kt0 || kk > 0) ? 1u : 0u);
            }
          }
          mma_commit(&ctl->empty[stage]);
          if (++stage == nst) {
            stage = 0;
            phase ^= 1;
          }
        }
        mma_commit(&ctl->tfull[0]);
      }
    } else if (warp >= 4) {
      const int ew = warp - 4;
      const int m = ew * 32 + lane;
      mbar_wait(&ctl->tfull[0], 0);
      tc_fence_after();
      for (int i = 0; i < mcount; ++i) {
        const int q = (mt0 + i) * CPT + m / CBA;
        const int tap = q / cpc, c = (q % cpc) * CBA + (m % CBA);
        float* dst = a.dw + ((size_t)(tap * a.Ctot + c) * a.Cout + n0);
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + i * BN + ch * 32, v);
          tmem_ld_wait();
          if (q < nchunk) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              red_add_v4(dst + ch * 32 + j * 4, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
        }
      }
    }
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

template <int CBA, int CBB>
static int launch_wgrad(const WgradTcArgs& a, cudaStream_t st) {
  const int stage_bytes = a.MT * KP * 128 * 2 + KP * a.BN * 2;
  size_t fixed = 1024 + sizeof(SmemCtl) + 64;
  int nst = (int)((kMaxDynSmem - fixed) / stage_bytes);
  if (nst > kMaxStages) nst = kMaxStages;
  RVIP_REQUIRE(nst >= 2, "wgrad_tc: tile too large for shared memory (MT=%d BN=%d)", a.MT, a.BN);
  int cols = 32;
  while (cols < a.MT * a.BN) cols *= 2;
  RVIP_REQUIRE(cols <= 512, "wgrad_tc: MT*BN=%d exceeds TMEM", a.MT * a.BN);
  static bool attr_set = false;
  if (!attr_set) {
    RVIP_CUDA(cudaFuncSetAttribute(wgrad3x3_tc_kernel<CBA, CBB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kMaxDynSmem));
    attr_set = true;
  }
  dim3 grid(a.n_mgroups * a.n_ntiles, a.n_split);
  launch_kernel(wgrad3x3_tc_kernel<CBA, CBB>, grid, dim3(256), fixed + (size_t)nst * stage_bytes, st, a, nst, cols);
  RVIP_LAUNCH_CHECK();
  return 0;
}

int wgrad_tc_launch(const WgradTcArgs& a, int CBA, int CBB, cudaStream_t st) {
  RVIP_REQUIRE(a.C0 % CBA == 0 && a.Ctot % CBA == 0 && a.Cout % CBB == 0 && a.BN % CBB == 0 && a.BN % 32 == 0,
               "wgrad_tc: bad channel tiling C0=%d Ctot=%d Cout=%d CBA=%d CBB=%d BN=%d", a.C0, a.Ctot, a.Cout, CBA,
               CBB, a.BN);
  if (CBA == 64 && CBB == 64) return launch_wgrad<64, 64>(a, st);
  if (CBA == 64 && CBB == 32) return launch_wgrad<64, 32>(a, st);
  if (CBA == 32 && CBB == 64) return launch_wgrad<32, 64>(a, st);
  if (CBA == 32 && CBB == 32) return launch_wgrad<32, 32>(a, st);
  set_error("wgrad_tc: unsupported chunking %d/%d", CBA, CBB);
  return 1;
}

}  // namespace rvip
