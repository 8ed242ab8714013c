// Halo-staged tcgen05 weight gradient for every level of the U-Net (any image size, any channel count
// that is a multiple of 32):
//
//   dW[tap][ci][co] = sum_p x[p + off(tap), ci] * dz[p, co]
//
// The generic kernel (conv_tc.cu: wgrad3x3_tc_kernel) fetches one shifted x box per tap, i.e. moves x nine
// times through L2 -- and L2 -> SM bandwidth (~43 B/cycle/SM) is what bounds these layers.  Here a CTA owns ONE
// input-channel chunk (CIC = 32 or 64 channels) and one BN-wide slice of the output channels, stages the x tile
// of a 128-pixel block once WITH its halo ({CIC ch, TW+2, TH+2} -> [pixel][CIC*2 B], swizzled as TMA lands
// it) and applies all nine taps from it; the nine taps' accumulators (3 or 5 M-tiles of 128 x BN fp32) stay in
// TMEM over every pixel tile of the CTA and are flushed once with vector fp32 reductions.
//
// Both operands are MN-major exactly as TMA lands them (rows = pixels = the GEMM K dimension, 16 pixels of one
// image row per MMA).  The M dimension packs taps as channel chunks that are whole pixels apart in the halo
// block (leading-dimension byte offset = pixel shift; the swizzle is a function of the absolute smem address,
// so shifted views stay consistent with what TMA wrote):
//   CIC = 32: M-tile dy = { dx=-1, 0, +1, (pad) } x 32 channels, chunks one pixel (64 B) apart, SWIZZLE_64B
//   CIC = 64: M-tile t  = taps (2t, 2t+1) x 64 channels (tap 9 = pad), chunks off(2t+1)-off(2t) pixels apart,
//             SWIZZLE_128B
// Replaces Conv2DBackpropFilter behind tf.keras Conv2D (src/models/KerasLayers.py:683,689,758).
#include "wgrad_halo.cuh"

#include "common.cuh"
#include "tc_prims.cuh"

namespace rvip {
using namespace tc;

constexpr int kWhMaxSmem = 227 * 1024;
constexpr int kWhMaxStages = 6;

struct WhCtl {
  uint64_t full[kWhMaxStages];
  uint64_t empty[kWhMaxStages];
  uint64_t done;
  uint32_t tmem_base;
};

__host__ __device__ constexpr int wh_round1k(int v) { return (v + 1023) & ~1023; }

__device__ __forceinline__ void wh_red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__device__ __forceinline__ void wh_red_add(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

template <int CIC, int BN, bool UP = false>
struct WhCfg {
  static constexpr int XPIXB = CIC * 2;                       // bytes per pixel of the x chunk
  static constexpr int CB = BN >= 64 ? 64 : 32;               // channels per dz TMA box
  static constexpr int DPIXB = CB * 2;
  static constexpr int NDZ = (UP ? 4 : 1) * (BN / CB);        // dz boxes per stage (UP: four phase views)
  static constexpr int MT = UP ? 8 : (CIC == 32 ? 3 : 5);     // M tiles (accumulators) per CTA
  static_assert(!UP || CIC == 64, "the phase-decomposed weight gradient packs column-neighbour pairs of 64 channels");
  static constexpr uint64_t LAYA = CIC == 64 ? kLayoutSW128 : kLayoutSW64;
  static constexpr uint64_t LAYB = CB == 64 ? kLayoutSW128 : kLayoutSW64;
  static_assert(MT * BN <= 512, "accumulators exceed TMEM");
};

// halo-pixel offset of tap t (0..8) inside a halo block that is `pitch` pixels wide
__host__ __device__ constexpr int wh_tap_off(int t, int pitch) { return (t / 3) * pitch + (t % 3); }

// TW (16 | 32) is a template parameter so that every operand offset of the MT * 8 MMAs of a stage is an immediate:
// with run-time tile geometry the single issuing thread spent ~120 dependent cycles per MMA on address
// arithmetic -- twice the MMA itself (ncu: tensor pipe 27..55 % active, L2 at 10 %).
// does 3x3 tap index k (0..2) fall on low-resolution neighbour r (0 / 1) of an output row / column of parity p?
__host__ __device__ constexpr bool wh_in_phase(int p, int r, int k) {
  return p == 0 ? (r == 0 ? k == 0 : k >= 1) : (r == 0 ? k <= 1 : k == 2);
}

// UP: weight gradient of the phase-decomposed up-convolution (conv_halo.cuh).  x0 = LOW-resolution input; for output
// phase (a, b) the pre-summed weight Wp[a][b][r][s] pairs low-resolution neighbour (a - 1 + r, b - 1 + s) with the
// phase view dz_ab:  M tile (a, b, r) = column neighbours s = 0, 1 as two 64-channel chunks one pixel apart, N = the
// phase's BN output channels -> 8 accumulators, 8 MMAs per 16 low-resolution pixels instead of 5 per 16 high-resolution
// ones (2.5x fewer), and the up-sampled copy of x is never read.  dWp is folded onto the 3x3 taps in the flush:
// dW[ky][kx] = sum over (a, r) with ky in S(a, r) and (b, s) with kx in S(b, s).
template <int CIC, int BN, int TW, bool UP>
__global__ void __launch_bounds__(256, 1) wgrad3x3_halo_kernel(const __grid_constant__ WgradHaloArgs a, int nst,
                                                               int tmem_cols) {
  using Cfg = WhCfg<CIC, BN, UP>;
  constexpr bool MERGE = !UP && CIC == 32 && BN <= 64;   // row-merged MMAs (see the issue loop)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int TH = 128 / TW, pitch = TW + 2;
  constexpr int x_tx = (TH + 2) * pitch * Cfg::XPIXB;         // bytes TMA delivers for the x halo block
  constexpr int x_st = wh_round1k(x_tx + 4 * Cfg::XPIXB);     // + slack: the pad chunk reads a few pixels past the end
  constexpr int dz_box = TH * TW * Cfg::DPIXB;
  constexpr int stage_bytes = x_st + Cfg::NDZ * dz_box;
  WhCtl* ctl = reinterpret_cast<WhCtl*>(smem + (size_t)nst * stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cchunk = blockIdx.x % a.n_cchunks, nt = blockIdx.x / a.n_cchunks;
  const int c0 = cchunk * CIC, n0 = nt * BN;
  const int k_per = (a.pixel_tiles + a.k_split - 1) / a.k_split;
  const int kt0 = blockIdx.y * k_per;
  const int kt1 = (kt0 + k_per) < a.pixel_tiles ? (kt0 + k_per) : a.pixel_tiles;
  const bool has_work = kt0 < kt1;

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
    mbar_init(&ctl->done, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&ctl->tmem_base, tmem_cols);
    tmem_relinquish();
  }
  pdl_wait();   // everything above is independent of the previous kernel's output
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (has_work && elect_one()) {
      int stage = 0, phase = 0;
      for (int kt = kt0; kt < kt1; ++kt) {
        const int x0 = (kt % a.tiles_x) * TW;
        const int y0 = ((kt / a.tiles_x) % a.tiles_y) * TH;
        const int b = kt / (a.tiles_x * a.tiles_y);
        mbar_wait(&ctl->empty[stage], phase ^ 1);
        mbar_expect_tx(&ctl->full[stage], x_tx + Cfg::NDZ * dz_box);
        uint8_t* X = smem + (size_t)stage * stage_bytes;
        if (c0 < a.C0)
          tma_load_4d(X, &a.x0, &ctl->full[stage], c0, x0 - 1, y0 - 1, b);
        else
          tma_load_4d(X, &a.x1, &ctl->full[stage], c0 - a.C0, x0 - 1, y0 - 1, b);
        if (UP) {
#pragma unroll
          for (int j = 0; j < Cfg::NDZ; ++j)
            tma_load_4d(X + x_st + j * dz_box, &a.dzp[j / (BN / Cfg::CB)], &ctl->full[stage],
                        n0 + (j % (BN / Cfg::CB)) * Cfg::CB, x0, y0, b);
        } else {
#pragma unroll
          for (int j = 0; j < Cfg::NDZ; ++j)
            tma_load_4d(X + x_st + j * dz_box, &a.dz, &ctl->full[stage], n0 + j * Cfg::CB, x0, y0, b);
        }
        if (++stage == nst) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (has_work && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
      constexpr int runs_x = TW >> 4;             // 16-pixel K runs per tile row
      int stage = 0, phase = 0;
      uint32_t accumulate = 0;
      for (int kt = kt0; kt < kt1; ++kt) {
        mbar_wait(&ctl->full[stage], phase);
        tc_fence_after();
        const uint32_t x_base = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t dz_base = x_base + x_st;
        if constexpr (MERGE) {
          // Row-merged form for N = BN <= 64 (an N = 32 MMA costs 48 cycles for 16 cycles of work): input row hr of the halo
          // block pairs with the output rows i = hr - ty of the dz tile for the taps ty = 0, 1, 2 -- ONE M tile (the dx taps
          // packed in M as before) times N = up to 3 * BN, the dz rows i_lo .. i_hi being consecutive N chunks TW pixels
          // apart (LBO) that land in accumulator column blocks 2 - ty (ascending with i).  (TH + 2) MMAs of N <= 3 BN per
          // 16-pixel column run instead of 3 TH of N = BN: ~1.8x (TH = 4) / 2x (TH = 8) fewer tensor cycles.
          constexpr uint32_t idesc2 = make_idesc_bf16(128, 2 * BN, 1, 1), idesc3 = make_idesc_bf16(128, 3 * BN, 1, 1);
          const uint64_t bdescM = make_smem_desc(dz_base, TW * Cfg::DPIXB, 8 * Cfg::DPIXB, Cfg::LAYB);
#pragma unroll
          for (int hr = 0; hr < TH + 2; ++hr) {
            const int i_lo = hr - 2 < 0 ? 0 : hr - 2, i_hi = hr > TH - 1 ? TH - 1 : hr;
            const int nblk = i_hi - i_lo + 1;
            const uint64_t adesc0 =
                make_smem_desc(x_base + hr * pitch * Cfg::XPIXB, Cfg::XPIXB, 8 * Cfg::XPIXB, Cfg::LAYA);
#pragma unroll
            for (int xb = 0; xb < runs_x; ++xb) {
              const uint32_t a_off = (uint32_t)((xb * 16) * Cfg::XPIXB) >> 4;
              if (accumulate == 0) {
                // first stage of this CTA: column block 2 - ty is first touched by (hr = ty, i = 0, xb = 0)
#pragma unroll
                for (int i = i_lo; i <= i_hi; ++i) {
                  const uint32_t b_off = (uint32_t)((i * TW + xb * 16) * Cfg::DPIXB) >> 4;
                  mma_bf16_ss(tmem_base + (2 - hr + i) * BN, adesc0 + a_off, bdescM + b_off, idesc,
                              (i == 0 && xb == 0) ? 0u : 1u);
                }
              } else {
                const uint32_t b_off = (uint32_t)((i_lo * TW + xb * 16) * Cfg::DPIXB) >> 4;
                mma_bf16_ss(tmem_base + (2 - hr + i_lo) * BN, adesc0 + a_off, bdescM + b_off,
                            nblk == 3 ? idesc3 : (nblk == 2 ? idesc2 : idesc), 1u);
              }
            }
          }
        }
        // B: dz, MN-major: 64-channel chunks one box apart, 8-pixel groups 8 * DPIXB apart
        const uint64_t bdesc0 = make_smem_desc(dz_base, dz_box, 8 * Cfg::DPIXB, Cfg::LAYB);
#pragma unroll
        for (int mt = 0; mt < (MERGE ? 0 : Cfg::MT); ++mt) {
          // A: x halo block, MN-major: channel chunks LBO apart (= pixel shift between packed taps)
          int start_px, lbo_px;
          if (UP) {
            const int ph = mt >> 1, r = mt & 1;   // phase (a, b) = (ph >> 1, ph & 1): halo rows a + r, columns b, b + 1
            start_px = ((ph >> 1) + r) * pitch + (ph & 1);
            lbo_px = 1;
          } else if (CIC == 32) {
            start_px = mt * pitch;                // tap (dy = mt, dx = 0); chunks dx = 0, 1, 2, (3 = pad)
            lbo_px = 1;
          } else {
            const int t0 = 2 * mt, t1 = 2 * mt + 1 > 8 ? 8 : 2 * mt + 1;
            start_px = wh_tap_off(t0, pitch);
            lbo_px = wh_tap_off(t1, pitch) - start_px;
            if (lbo_px == 0) lbo_px = 1;          // pad chunk of the last M tile
          }
          const uint64_t adesc0 =
              make_smem_desc(x_base + start_px * Cfg::XPIXB, lbo_px * Cfg::XPIXB, 8 * Cfg::XPIXB, Cfg::LAYA);
          const uint32_t d = tmem_base + mt * BN;
#pragma unroll
          for (int y = 0; y < TH; ++y) {
#pragma unroll
            for (int xb = 0; xb < runs_x; ++xb) {
              const uint32_t a_off = (uint32_t)((y * pitch + xb * 16) * Cfg::XPIXB) >> 4;
              const uint32_t b_off = (uint32_t)((y * TW + xb * 16) * Cfg::DPIXB) >> 4;
              const uint32_t b_ph = UP ? (uint32_t)((mt >> 1) * (BN / Cfg::CB) * dz_box) >> 4 : 0u;   // phase view of dz
              mma_bf16_ss(d, adesc0 + a_off, bdesc0 + b_ph + b_off, idesc, (y | xb) != 0 ? 1u : accumulate);
            }
          }
        }
        mma_commit(&ctl->empty[stage]);
        accumulate = 1;
        if (++stage == nst) {
          stage = 0;
          phase ^= 1;
        }
      }
      mma_commit(&ctl->done);
      pdl_launch_dependents();   // all MMAs issued: the next kernel may start launching behind the flush
    }
  } else if (warp >= 4 && has_work) {
    // ------------------------------------------------------------------ epilogue: TMEM -> fp32 reductions into HWIO
    const int ew = warp - 4;
    const int m = ew * 32 + lane;               // accumulator row
    mbar_wait(&ctl->done, 0);
    tc_fence_after();
    if (UP && a.transposed) {
      // Conv2DTranspose(3, strides 2, 'same'): tap k of an axis pairs output parity / neighbour (p, r) = (0, 1) for k = 0,
      // (1, 0) for k = 1, (0, 0) for k = 2 -- one accumulator per tap, nothing to fold.  dW is (kh, kw, Cout, Cin):
      // rows of the accumulator (ci) are the contiguous axis, so a column is one coalesced scalar reduction per warp.
      const int s = m >> 6, ci = m & 63;
#pragma unroll 1
      for (int ky = 0; ky < 3; ++ky) {
        const int pa = ky == 1 ? 1 : 0, r = ky == 0 ? 1 : 0;
#pragma unroll 1
        for (int opt = 0; opt < (s == 0 ? 2 : 1); ++opt) {
          const int kx = s == 0 ? (opt == 0 ? 2 : 1) : 0, pb = (s == 0 && opt == 1) ? 1 : 0;
          float* dst = a.dw + ((size_t)((ky * 3 + kx) * a.Cout + n0) * a.Ctot + c0 + ci);
#pragma unroll 1
          for (int ch = 0; ch < BN / 32; ++ch) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + ((pa * 2 + pb) * 2 + r) * BN + ch * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) wh_red_add(dst + (size_t)(ch * 32 + j) * a.Ctot, __uint_as_float(v[j]));
          }
        }
      }
    } else if (UP) {
      const int s = m >> 6, ci = m & 63;        // column neighbour of this row's chunk (warp-uniform), input channel
#pragma unroll 1
      for (int ky = 0; ky < 3; ++ky) {
#pragma unroll 1
        for (int kk = 0; kk < 2; ++kk) {
          const int kx = s + kk;                 // neighbour s = 0 carries kx in {0, 1}, s = 1 carries kx in {1, 2}
          float* dst = a.dw + ((size_t)((ky * 3 + kx) * a.Ctot + c0 + ci) * a.Cout + n0);
#pragma unroll 1
          for (int ch = 0; ch < BN / 32; ++ch) {
            float acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0.f;
#pragma unroll
            for (int ar = 0; ar < 4; ++ar) {
              if (!wh_in_phase(ar >> 1, ar & 1, ky)) continue;
#pragma unroll
              for (int pb = 0; pb < 2; ++pb) {
                if (!wh_in_phase(pb, s, kx)) continue;
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (((ar >> 1) * 2 + pb) * 2 + (ar & 1)) * BN + ch * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(v[j]);
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              wh_red_add_v4(dst + ch * 32 + j * 4, acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          }
        }
      }
    }
    if (MERGE) {
      // one M tile: rows = (dx chunk, ci); column block cb holds tap ty = 2 - cb
      const int dxc = m >> 5, ci = m & 31;
#pragma unroll 1
      for (int cb = 0; cb < 3; ++cb) {
        const int tap = (2 - cb) * 3 + (dxc < 3 ? dxc : 0);
        float* dst = a.dw + ((size_t)(tap * a.Ctot + c0 + ci) * a.Cout + n0);
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + cb * BN + ch * 32, v);
          tmem_ld_wait();
          if (dxc < 3) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              wh_red_add_v4(dst + ch * 32 + j * 4, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                            __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          }
        }
      }
    }
#pragma unroll 1
    for (int mt = 0; mt < ((UP || MERGE) ? 0 : Cfg::MT); ++mt) {
      int tap, ci;
      if (CIC == 32) {
        tap = mt * 3 + (m >> 5);                // dy = mt, dx = chunk; chunk 3 is padding
        ci = m & 31;
        if ((m >> 5) == 3) tap = -1;
      } else {
        tap = 2 * mt + (m >> 6);
        ci = m & 63;
        if (tap > 8) tap = -1;
      }
      float* dst = a.dw + ((size_t)((tap < 0 ? 0 : tap) * a.Ctot + c0 + ci) * a.Cout + n0);
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + mt * BN + ch * 32, v);
        tmem_ld_wait();
        if (tap >= 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            wh_red_add_v4(dst + ch * 32 + j * 4, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                          __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------- host
static size_t wh_stage_bytes(int CIC, int BN, int TW, int TH, bool up = false) {
  const int xpixb = CIC * 2, cb = BN >= 64 ? 64 : 32;
  const int x_st = wh_round1k((TH + 2) * (TW + 2) * xpixb + 4 * xpixb);
  return (size_t)x_st + (size_t)TH * TW * cb * 2 * (BN / cb) * (up ? 4 : 1);
}
// pixel-tile width for an image row of W pixels: the wider tile unless it wastes more zero-filled columns
static int wh_tile_w(int W) {
  if (W < 32) return 16;
  const int waste32 = (W + 31) / 32 * 32 - W, waste16 = (W + 15) / 16 * 16 - W;
  return waste32 <= waste16 ? 32 : 16;
}
static size_t wh_fixed_bytes() { return 1024 + sizeof(WhCtl) + 64; }
// Stages are small (128 pixels) and many: the ring has to cover the TMA latency (~2000 cycles) with loads in
// flight while a stage's MMAs (1500..2000 cycles) run; two 256-pixel stages measured load-latency bound.
static int wh_stages(int CIC, int BN, int TW, int TH, bool up = false) {
  int n = (int)((kWhMaxSmem - wh_fixed_bytes()) / wh_stage_bytes(CIC, BN, TW, TH, up));
  return n > kWhMaxStages ? kWhMaxStages : n;
}

bool wgrad_halo_up_plan(int B, int h, int w, int Cin, int Cout, WgradHaloPlan* p) {
  if (w < 1 || Cin % 64 != 0 || Cout % 32 != 0) return false;   // columns past the image are zero-filled by TMA
  p->CIC = 64;
  p->BN = Cout % 64 == 0 ? 64 : 32;          // 8 accumulators x BN columns <= 512 TMEM columns
  p->TW = wh_tile_w(w);
  p->TH = 128 / p->TW;
  if (wh_stages(64, p->BN, p->TW, p->TH, true) < 2) return false;
  p->tiles_x = (w + p->TW - 1) / p->TW;
  p->tiles_y = (h + p->TH - 1) / p->TH;
  p->pixel_tiles = p->tiles_x * p->tiles_y * B;
  p->n_cchunks = Cin / 64;
  p->n_ntiles = Cout / p->BN;
  const int units = p->n_cchunks * p->n_ntiles;
  int split = kNumSMs / units;
  if (split < 1) split = 1;
  if (split > (p->pixel_tiles + 3) / 4) split = (p->pixel_tiles + 3) / 4;
  if (split < 1) split = 1;
  p->k_split = split;
  return true;
}

bool wgrad_halo_plan(int B, int H, int W, int C0, int C1, int Cout, WgradHaloPlan* p) {
  if (W < 1 || C0 % 32 != 0 || C1 % 32 != 0 || Cout % 32 != 0) return false;   // columns past the image are zero-filled by TMA
  const int Ctot = C0 + C1;
  int cic, bn;
  if (Cout % 128 == 0) {
    cic = 32; bn = 128;                      // 3 M tiles x 128 columns, N = 128 runs the tensor pipe at ~97 %
  } else if (Cout % 64 == 0) {
    if (C0 % 64 == 0 && C1 % 64 == 0) { cic = 64; bn = 64; }   // 5 M tiles (taps paired: 90 % useful rows)
    else { cic = 32; bn = 64; }
  } else {
    if (C0 % 64 == 0 && C1 % 64 == 0) { cic = 64; bn = 32; }
    else { cic = 32; bn = 32; }
  }
  p->CIC = cic; p->BN = bn;
  p->TW = wh_tile_w(W);
  p->TH = 128 / p->TW;                     // rows / columns past the image are zero-filled by TMA and contribute nothing
  if (wh_stages(cic, bn, p->TW, p->TH) < 2) return false;
  p->tiles_x = (W + p->TW - 1) / p->TW;
  p->tiles_y = (H + p->TH - 1) / p->TH;
  p->pixel_tiles = p->tiles_x * p->tiles_y * B;
  p->n_cchunks = Ctot / cic;
  p->n_ntiles = Cout / bn;
  const int units = p->n_cchunks * p->n_ntiles;
  // split the pixel tiles so that about one wave of CTAs runs, each with at least 2 tiles to pipeline
  int split = kNumSMs / units;               // never more CTAs than SMs: a second, nearly empty wave doubles the time
  if (split < 1) split = 1;
  if (split > (p->pixel_tiles + 3) / 4) split = (p->pixel_tiles + 3) / 4;
  if (split < 1) split = 1;
  p->k_split = split;
  return true;
}

template <int CIC, int BN, int TW, bool UP = false>
static int launch_wh(const WgradHaloArgs& a, cudaStream_t st) {
  const int nst = wh_stages(CIC, BN, a.TW, a.TH, UP);
  RVIP_REQUIRE(nst >= 2, "wgrad_halo: tile %dx%d does not fit shared memory", a.TW, a.TH);
  const size_t smem = wh_fixed_bytes() + (size_t)nst * wh_stage_bytes(CIC, BN, a.TW, a.TH, UP);
  int cols = 32;
  while (cols < WhCfg<CIC, BN, UP>::MT * BN) cols *= 2;
  static bool attr_set = false;
  if (!attr_set) {
    RVIP_CUDA(cudaFuncSetAttribute(wgrad3x3_halo_kernel<CIC, BN, TW, UP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kWhMaxSmem));
    attr_set = true;
  }
  launch_kernel(wgrad3x3_halo_kernel<CIC, BN, TW, UP>, dim3(a.n_cchunks * a.n_ntiles, a.k_split), dim3(256), smem, st, a, nst,
                cols);
  RVIP_LAUNCH_CHECK();
  return 0;
}

int wgrad_halo_launch(const WgradHaloArgs& a, int CIC, int BN, cudaStream_t st) {
  RVIP_REQUIRE(a.C0 % CIC == 0 && a.Ctot % CIC == 0 && a.Cout % BN == 0, "wgrad_halo: bad channel tiling C0=%d Ctot=%d "
               "Cout=%d CIC=%d BN=%d", a.C0, a.Ctot, a.Cout, CIC, BN);
  RVIP_REQUIRE((a.TW == 16 || a.TW == 32) && a.TH == 128 / a.TW, "wgrad_halo: unsupported pixel tile %dx%d", a.TW, a.TH);
  if (a.up) {
    RVIP_REQUIRE(CIC == 64 && (BN == 64 || BN == 32), "wgrad_halo: up-convolution tile CIC=%d BN=%d", CIC, BN);
    if (BN == 64) return a.TW == 32 ? launch_wh<64, 64, 32, true>(a, st) : launch_wh<64, 64, 16, true>(a, st);
    return a.TW == 32 ? launch_wh<64, 32, 32, true>(a, st) : launch_wh<64, 32, 16, true>(a, st);
  }
#define RVIP_WH_CASE(cic, bn)                                                        \
  if (CIC == cic && BN == bn)                                                        \
    return a.TW == 32 ? launch_wh<cic, bn, 32>(a, st) : launch_wh<cic, bn, 16>(a, st);
  RVIP_WH_CASE(32, 128)
  RVIP_WH_CASE(32, 64)
  RVIP_WH_CASE(32, 32)
  RVIP_WH_CASE(64, 64)
  RVIP_WH_CASE(64, 32)
#undef RVIP_WH_CASE
  set_error("wgrad_halo: unsupported tile CIC=%d BN=%d", CIC, BN);
  return 1;
}

}  // namespace rvip
