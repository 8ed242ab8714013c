// Row-tiled tcgen05 3x3 convolution for the wide, few-channel layers (rows of 128-pixel tiles: the full- and
// half-resolution levels of the U-Net), forward and dgrad.
//
// The generic kernel (conv_tc.cu) fetches one shifted input box per tap, i.e. reads the activation nine
// times through L2; these layers are HBM/L2-bound (arithmetic intensity 144..384 FLOP/B), so here the
// input is staged ONCE per tile with its halo:  TMA box {32 ch, 130 px, R+2 rows} -> smem as a dense
// [pixel][32 ch] (64 B, SWIZZLE_64B) array.  Output row i of the tile is one M=128 accumulator; for tap
// (dy, dx) its A operand is simply the same smem block read from pixel (i+dy+1)*130 + (dx+1) on: 128
// consecutive pixels = 16 core-matrix groups at a uniform 512 B stride, so a K-major descriptor whose
// start address is shifted by whole pixels addresses it directly (the 64B swizzle is a function of the
// absolute smem address, identical for the TMA write and the MMA read).  Input traffic drops from 9x
// to (R+2)/R x.  Small weight sets stay resident in shared memory for the whole kernel.
#include "conv_row.cuh"

#include "common.cuh"
#include "tc_prims.cuh"

namespace rvip {
using namespace tc;

constexpr int kMaxDynSmemRow = 227 * 1024;
constexpr int kRowStagesMax = 4;
constexpr int kHaloW = 130;           // 128 output columns + 1 halo column each side
constexpr int kPixB = 64;             // bytes per pixel of a 32-channel chunk

__device__ __forceinline__ void row_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct RowCtl {
  uint64_t full[kRowStagesMax];
  uint64_t empty[kRowStagesMax];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint64_t wfull;
  uint32_t tmem_base;
};

__host__ __device__ constexpr int row_a_bytes(int R) { return (R + 2) * kHaloW * kPixB; }
// Shared-memory slot of filter tap (dy, dx) inside a chunk's weight block: grouped by dx, dy DESCENDING.  Input (halo) row
// hr feeds output rows hr-2, hr-1, hr through taps dy = 2, 1, 0: in this order their weights are consecutive B rows in
// the order of the accumulator columns (row i at columns i * BN), so ONE tcgen05.mma of N = 3 * BN applies all three.
__host__ __device__ constexpr int row_w_slot(int tap) { return (tap % 3) * 3 + (2 - tap / 3); }
__host__ __device__ constexpr int round1k(int v) { return (v + 1023) & ~1023; }

template <int BN, int R, bool ZERO_BASE>
__device__ __forceinline__ void row_mma_loop(const ConvRowArgs& a, RowCtl* ctl, uint8_t* stages, uint8_t* wres,
                                             int stage_bytes, int nst, int nchunks, uint32_t tmem_base_rt) {
  constexpr int A_ST = round1k(row_a_bytes(R));
  constexpr int W_TILE = BN * kPixB;
  constexpr int ACC_COLS = R * BN;
  const uint32_t tmem_base = ZERO_BASE ? 0u : tmem_base_rt;
  int stage = 0, phase = 0, acc = 0, acc_phase = 0;
  if (a.wres) mbar_wait(&ctl->wfull, 0);
  for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
    const int nt = tile % a.n_ntiles;
    mbar_wait(&ctl->tempty[acc], acc_phase ^ 1);
    tc_fence_after();
    const uint32_t d0 = tmem_base + acc * ACC_COLS;
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(&ctl->full[stage], phase);
      tc_fence_after();
      const uint32_t a_base = smem_u32(stages + (size_t)stage * stage_bytes);
      const uint32_t w_base = a.wres ? smem_u32(wres) + ((nt * nchunks + c) * 9) * W_TILE : a_base + A_ST;
      // one descriptor per operand block; every row / dx / k-half is a compile-time offset of its start field.
      // N = 32 MMAs run at 48 cycles against an ideal 16 (the 4 KB A tile is re-read from shared memory for every MMA):
      // instead of nine taps x R rows of N = BN, every input row is applied ONCE per dx with N = up to 3 * BN -- the
      // same A tile feeds the three output rows it contributes to (weights ordered as the accumulator columns,
      // row_w_slot).  (R + 2) * 3 MMAs of N <= 3 BN per K step instead of 9 R of N = BN: 1.8x fewer tensor cycles.
      const uint64_t adesc0 = make_smem_desc(a_base, 16, 512, kLayoutSW64);
      const uint64_t bdesc0 = make_smem_desc(w_base, 16, 512, kLayoutSW64);
#pragma unroll
      for (int hr = 0; hr < R + 2; ++hr) {
        constexpr uint32_t idesc1 = make_idesc_bf16(128, BN, 0, 0), idesc2 = make_idesc_bf16(128, 2 * BN, 0, 0),
                           idesc3 = make_idesc_bf16(128, 3 * BN, 0, 0);
        const int i_lo = hr - 2 < 0 ? 0 : hr - 2, i_hi = hr > R - 1 ? R - 1 : hr;
        const int nblk = i_hi - i_lo + 1;                  // output rows fed by this input row
        const int wrow = (2 - (hr - i_lo)) * BN;           // first B row: tap dy = hr - i_lo, slots are dy-descending
        const uint32_t d = d0 + i_lo * BN;
        const uint32_t idn = nblk == 3 ? idesc3 : (nblk == 2 ? idesc2 : idesc1);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t ad = adesc0 + (uint64_t)((((hr * kHaloW + dx) * kPixB) >> 4) + 2 * k);
            const uint64_t bd = bdesc0 + (uint64_t)((((dx * 3 * BN + wrow) * kPixB) >> 4) + 2 * k);
            if (dx == 0 && k == 0 && hr < R) {
              // output row hr is touched for the first time in chunk 0: its columns start from zero, the others accumulate
              if (c == 0) {
                if (nblk > 1) mma_bf16_ss(d, ad, bd, nblk == 3 ? idesc2 : idesc1, 1u);
                mma_bf16_ss(d0 + hr * BN, ad, bdesc0 + (uint64_t)((((2 * BN) * kPixB) >> 4)), idesc1, 0u);
              } else {
                mma_bf16_ss(d, ad, bd, idn, 1u);
              }
            } else {
              mma_bf16_ss(d, ad, bd, idn, 1u);
            }
          }
        }
      }
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
  pdl_launch_dependents();   // all MMAs issued: the next kernel may start launching behind the last epilogue
}

// MODE (ConvEpilogue) is a template parameter: ncu showed the 8 epilogue warps -- two per scheduler, ~950 instructions per
// warp and tile -- setting the tile period of the level-0 layers (2.6 us against 1.2 us of MMAs), and with a run-time
// mode every element of a dgrad tile still issued the predicated-off bias / ReLU / statistics instructions.
// NEW = epilogue warps (8 or 16).  With 8, two warps share a scheduler and each walks ~700-950 dependent instructions per
// tile: the epilogue, not the MMAs (1.2 us) or HBM, set the tile period of the level-0 layers (2.6 us).  With 16 every
// warp owns ONE 16-channel sub-chunk of the N tile (fixed for the whole kernel when the layer has a single N tile), half
// the instructions per tile, four warps per scheduler; bias and the BatchNorm partial sums of its 16 channels stay in
// registers across all tiles.
template <int BN, int R, int MODE, int OCH, int NEW = 8>
__global__ void __launch_bounds__(128 + 32 * NEW, 1) conv3x3_row_kernel(const __grid_constant__ ConvRowArgs a, int nst) {
  constexpr bool AFFINE = MODE == EPI_RELU_AFFINE;
  constexpr bool LIN = MODE == EPI_LINEAR || MODE == EPI_LINEAR_BNRED;   // no bias / activation in the epilogue
  constexpr bool RED = MODE == EPI_LINEAR_BNRED;                        // ... plus the BatchNorm-backward sums
  constexpr bool SUMS = MODE == EPI_RELU_STATS || RED;                   // per-channel partial sums of some kind
  constexpr int A_TX = row_a_bytes(R);
  constexpr int A_ST = round1k(A_TX);
  constexpr int W_TILE = BN * kPixB;           // one (chunk, tap) weight tile: BN rows x 64 B
  constexpr int OROWB = OCH * 2;
  constexpr int OCHUNK = R * 128 * OROWB;
  constexpr int STAGING = R * 128 * BN * 2;
  constexpr int ACC_COLS = R * BN;
  constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nchunks = a.Ctot / 32;
  const int wres_bytes = a.wres ? nchunks * 9 * a.Cout * kPixB : 0;
  const int stage_bytes = A_ST + (a.wres ? 0 : 9 * W_TILE);
  uint8_t* wres = smem;
  uint8_t* stages = smem + round1k(wres_bytes);
  uint8_t* staging = stages + (size_t)nst * stage_bytes;
  float* s_bias = reinterpret_cast<float*>(staging + STAGING);     // [Cout]
  float* s_slot = s_bias + a.Cout;                                  // [8 warps][2][Cout] sum / sum^2 partials
  float* s_aff = s_slot + 16 * a.Cout;                              // [2][Cout] scale, shift (EPI_RELU_AFFINE)
  RowCtl* ctl = reinterpret_cast<RowCtl*>((reinterpret_cast<uintptr_t>(s_aff + 2 * a.Cout) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
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
      mbar_init(&ctl->tempty[i], NEW);
    }
    mbar_init(&ctl->wfull, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&ctl->tmem_base, TMEM_COLS);
    tmem_relinquish();
  }
  pdl_wait();   // everything above is independent of the previous kernel's output
  if (warp >= 4) {
    for (int c = threadIdx.x - 128; c < 16 * a.Cout; c += 32 * NEW) s_slot[c] = 0.f;
    for (int c = threadIdx.x - 128; c < a.Cout; c += 32 * NEW) {
      s_bias[c] = !LIN ? a.bias[c] : 0.f;
      if (AFFINE) {
        s_aff[c] = a.scale[c];
        s_aff[a.Cout + c] = a.shift[c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      if (a.wres) {
        mbar_expect_tx(&ctl->wfull, wres_bytes);
        for (int nt = 0; nt < a.n_ntiles; ++nt)
          for (int c = 0; c < nchunks; ++c)
            for (int tap = 0; tap < 9; ++tap)
              tma_load_2d(wres + ((nt * nchunks + c) * 9 + row_w_slot(tap)) * W_TILE, &a.w, &ctl->wfull,
                          tap * a.Ctot + c * 32, nt * BN);
      }
      int stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int nt = tile % a.n_ntiles, pt = tile / a.n_ntiles;
        const int x0 = (pt % a.tiles_x) * 128;
        const int y0 = ((pt / a.tiles_x) % a.tiles_y) * R;
        const int b = pt / (a.tiles_x * a.tiles_y);
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&ctl->empty[stage], phase ^ 1);
          mbar_expect_tx(&ctl->full[stage], A_TX + (a.wres ? 0 : 9 * W_TILE));
          uint8_t* A = stages + (size_t)stage * stage_bytes;
          const int cc = c * 32;
          if (cc < a.C0)
            tma_load_4d(A, &a.in0, &ctl->full[stage], cc, x0 - 1, y0 - 1, b);
          else
            tma_load_4d(A, &a.in1, &ctl->full[stage], cc - a.C0, x0 - 1, y0 - 1, b);
          if (!a.wres)
            for (int tap = 0; tap < 9; ++tap)
              tma_load_2d(A + A_ST + row_w_slot(tap) * W_TILE, &a.w, &ctl->full[stage], tap * a.Ctot + cc, nt * BN);
          if (++stage == nst) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      // The first (and only) TMEM allocation of an SM-exclusive CTA starts at column 0; with the base a literal
      // every tcgen05.mma operand lives in uniform registers (no per-instruction R2UR/ELECT loop).
      if (tmem_base == 0)
        row_mma_loop<BN, R, true>(a, ctl, stages, wres, stage_bytes, nst, nchunks, 0u);
      else
        row_mma_loop<BN, R, false>(a, ctl, stages, wres, stage_bytes, nst, nchunks, tmem_base);
    }
  } else if (warp >= 4 && NEW == 16) {
    // ------------------------------------------------------------------ epilogue (16 warps; modes without BNRED; the
    // statistics mode only for layers with one N tile, conv_row_launch)
    constexpr int NSUB = BN / 16;              // 16-channel sub-chunks of the N tile
    constexpr int RSTEP = 4 / NSUB;            // warp groups per sub-chunk: rows of a tile are dealt round-robin to them
    constexpr int NU = R / RSTEP;              // (row, sub-chunk) units per thread and tile
    static_assert(RSTEP >= 1 && NU >= 1, "row kernel: 16-warp epilogue geometry");
    const int q = (warp - 4) & 3, eg = (warp - 4) >> 2;
    const int sub = eg % NSUB, rbase = eg / NSUB;
    const int r = q * 32 + lane;              // pixel column inside the tile == TMEM lane
    const int et = threadIdx.x - 128;         // 0..511
    int acc = 0, acc_phase = 0;
    float p1[16], p2[16], bias[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) p1[j] = p2[j] = bias[j] = 0.f;
    for (int tile_i = blockIdx.x; tile_i < a.total_tiles; tile_i += gridDim.x) {
      const int tile = tile_i;
      const int nt = tile % a.n_ntiles, pt = tile / a.n_ntiles;
      const int x0 = (pt % a.tiles_x) * 128;
      const int y0 = ((pt / a.tiles_x) % a.tiles_y) * R;
      const int b = pt / (a.tiles_x * a.tiles_y);
      const int n0 = nt * BN, c0 = n0 + sub * 16;
      const bool oobx = MODE == EPI_RELU_STATS && x0 + 128 > a.W && x0 + r >= a.W;
      if (et == 0) tma_store_wait_read0();   // previous tile's TMA store has drained the staging buffer
      row_bar_sync(1, 32 * NEW);
      if (!LIN) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = *reinterpret_cast<const float4*>(s_bias + c0 + j * 4);
          bias[4 * j] = t.x; bias[4 * j + 1] = t.y; bias[4 * j + 2] = t.z; bias[4 * j + 3] = t.w;
        }
      }
      mbar_wait(&ctl->tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < NU; ++k) {
        const int i = rbase + k * RSTEP;
        const uint32_t row = i * 128 + r;
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + i * BN + sub * 16, v);
        tmem_ld_wait();
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          f[j] = __uint_as_float(v[j]);
          if (!LIN) f[j] = fmaxf(f[j] + bias[j], a.floor);
        }
        if (MODE == EPI_RELU_STATS) {
          if (oobx) {                      // pixel past the end of the image row: clipped by the store, not counted
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            p1[j] += f[j];
            p2[j] = fmaf(f[j], f[j], p2[j]);
          }
        }
        if (AFFINE) {
          const float* sa = s_aff + c0;
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaf(f[j], sa[j], sa[a.Cout + j]);
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint4 pk;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[g * 8 + 2 * j], f[g * 8 + 2 * j + 1]);
          const int col = sub * 16 + g * 8;
          const int oc = col / OCH, cidx = (col % OCH) / 8;
          *reinterpret_cast<uint4*>(staging + oc * OCHUNK + swz_off<OROWB>(row, cidx)) = pk;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->tempty[acc]);
      fence_proxy_async_smem();
      row_bar_sync(1, 32 * NEW);
      if (et == 0) {
#pragma unroll 1
        for (int oc = 0; oc < BN / OCH; ++oc) {
          const int n = n0 + oc * OCH;
          if (MODE == EPI_LINEAR && n >= a.out_split)
            tma_store_4d(&a.out1, staging + oc * OCHUNK, n - a.out_split, x0, y0, b);
          else
            tma_store_4d(&a.out0, staging + oc * OCHUNK, n, x0, y0, b);
        }
        tma_store_commit();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (et == 0) tma_store_wait_all0();
    if (MODE == EPI_RELU_STATS) {
      // transpose-reduce of the 16 per-pixel sums: after four steps lane l holds channel (l & 15) over its half warp
#pragma unroll
      for (int S = 8; S >= 1; S >>= 1) {
        const bool up = (lane & S) != 0;
#pragma unroll
        for (int k = 0; k < S; ++k) {
          const float send1 = up ? p1[k] : p1[k + S], keep1 = up ? p1[k + S] : p1[k];
          const float send2 = up ? p2[k] : p2[k + S], keep2 = up ? p2[k + S] : p2[k];
          p1[k] = keep1 + __shfl_xor_sync(0xffffffffu, send1, S);
          p2[k] = keep2 + __shfl_xor_sync(0xffffffffu, send2, S);
        }
      }
      p1[0] += __shfl_xor_sync(0xffffffffu, p1[0], 16);
      p2[0] += __shfl_xor_sync(0xffffffffu, p2[0], 16);
      float* slot = s_slot + (size_t)(warp - 4) * 32;       // [16 warps][2][16]
      if (lane < 16) {
        slot[lane] = p1[0];
        slot[16 + lane] = p2[0];
      }
      row_bar_sync(1, 32 * NEW);
      // single N tile (Cout == BN): channel c lives in sub-chunk c / 16, held by the warps with eg % NSUB == c / 16
      for (int c = et; c < 2 * a.Cout; c += 32 * NEW) {
        const int which = c / a.Cout, ch = c % a.Cout;
        float t = 0.f;
        for (int w = 0; w < 16; ++w)
          if (((w >> 2) % NSUB) == ch / 16) t += s_slot[(size_t)w * 32 + which * 16 + (ch & 15)];
        atomicAdd(&a.stats[c], (double)t);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (8 warps)
    // Two warps per TMEM lane quarter: warp (4+q) takes rows [0, R/2), warp (8+q) rows [R/2, R) of the tile.
    // Per 32-column chunk a thread keeps its pixel's values in registers: bias + ReLU, bf16 pack into the
    // swizzled staging tile, and running per-channel sum / sum-of-squares that are folded across the warp
    // with a 31-shuffle transpose-reduce (lane j ends up owning channel j) into a per-warp smem slot.
    const int q = (warp - 4) & 3, eh = (warp - 4) >> 2;
    const int r = q * 32 + lane;            // pixel column inside the tile == TMEM lane
    const int et = threadIdx.x - 128;       // 0..255
    float* slot = s_slot + (size_t)(warp - 4) * 2 * a.Cout;
    int acc = 0, acc_phase = 0;
    // BN == 32 (one 32-channel chunk per tile): the per-thread BatchNorm partial sums live in registers across ALL tiles
    // of the CTA (same N tile every time when Cout == 32) and the 31-shuffle transpose-reduce runs ONCE at the end -- per
    // tile it was a quarter of the epilogue's instructions, and the epilogue (two warps per scheduler) sets the tile period
    constexpr bool PERSIST = SUMS && BN == 32;
    // EPI_LINEAR_BNRED: the thread's items of a tile (rows x 32-channel chunks) of the BatchNorm block's stored relu(conv)
    // tensor, fetched with plain coalesced loads BEFORE the accumulator is waited for (the latency hides behind the MMAs)
    constexpr int NIT = (R / 2) * (BN / 32);
    static_assert(!RED || NIT <= 2, "EPI_LINEAR_BNRED keeps at most two items of the aux tensor in registers");
    uint4 aux[RED ? NIT * 4 : 1];
    const DropKey dkey = {a.bnred.k0, a.bnred.k1};
    const bool persist = PERSIST && a.n_ntiles == 1;
    float p1[PERSIST ? 32 : 1], p2[PERSIST ? 32 : 1];
#pragma unroll
    for (int j = 0; j < (PERSIST ? 32 : 1); ++j) p1[j] = p2[j] = 0.f;
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int nt = tile % a.n_ntiles, pt = tile / a.n_ntiles;
      const int x0 = (pt % a.tiles_x) * 128;
      const int y0 = ((pt / a.tiles_x) % a.tiles_y) * R;
      const int b = pt / (a.tiles_x * a.tiles_y);
      const int n0 = nt * BN;
      const bool oobx = SUMS && x0 + 128 > a.W && x0 + r >= a.W;   // only a partial last tile has any
      if (RED) {
        const __nv_bfloat16* A = static_cast<const __nv_bfloat16*>(a.bnred.a);
#pragma unroll
        for (int ch = 0; ch < BN / 32; ++ch)
#pragma unroll
          for (int ii = 0; ii < R / 2; ++ii) {
            const int i = eh * (R / 2) + ii;
            const size_t pix = ((size_t)b * a.H + y0 + i) * a.W + x0 + r;
            const uint4* src = reinterpret_cast<const uint4*>(A + pix * a.Cout + n0 + ch * 32);
#pragma unroll
            for (int g = 0; g < 4; ++g) aux[(ch * (R / 2) + ii) * 4 + g] = oobx ? make_uint4(0u, 0u, 0u, 0u) : __ldg(src + g);
          }
      }
      if (et == 0) tma_store_wait_read0();   // previous tile's TMA store has drained the staging buffer
      row_bar_sync(1, 256);
      mbar_wait(&ctl->tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        float bias[32], s1[32], s2[32];
        if (!LIN) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = *reinterpret_cast<const float4*>(s_bias + n0 + ch * 32 + j * 4);
            bias[4 * j] = t.x; bias[4 * j + 1] = t.y; bias[4 * j + 2] = t.z; bias[4 * j + 3] = t.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          s1[j] = PERSIST ? p1[j] : 0.f;
          s2[j] = PERSIST ? p2[j] : 0.f;
        }
#pragma unroll
        for (int ii = 0; ii < R / 2; ++ii) {
          const int i = eh * (R / 2) + ii;
          const uint32_t row = i * 128 + r;
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * ACC_COLS + i * BN + ch * 32, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            f[j] = __uint_as_float(v[j]);
            if (!LIN) f[j] = fmaxf(f[j] + bias[j], a.floor);
          }
          if (RED) {
            // sum keep * dy and sum keep * dy * a per channel (x keep_scale at the very end); dy = this accumulator
            const uint32_t pix = ((uint32_t)b * a.H + y0 + i) * a.W + x0 + r;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint4 raw = aux[(ch * (R / 2) + ii) * 4 + g];
              const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&raw);
              bool keep[8];
              if (a.bnred.thr16) {
                dropout_keep8(dkey, (pix << a.bnred.lg) | (uint32_t)((n0 + ch * 32) / 8 + g), a.bnred.thr16, keep);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) keep[j] = true;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 av = __bfloat1622float2(hh[j]);
                const float d0 = (keep[2 * j] && !oobx) ? f[g * 8 + 2 * j] : 0.f;
                const float d1 = (keep[2 * j + 1] && !oobx) ? f[g * 8 + 2 * j + 1] : 0.f;
                s1[g * 8 + 2 * j] += d0;
                s1[g * 8 + 2 * j + 1] += d1;
                s2[g * 8 + 2 * j] = fmaf(d0, av.x, s2[g * 8 + 2 * j]);
                s2[g * 8 + 2 * j + 1] = fmaf(d1, av.y, s2[g * 8 + 2 * j + 1]);
              }
            }
          }
          if (MODE == EPI_RELU_STATS) {
            if (oobx) {                      // pixel past the end of the image row: clipped by the store, not counted
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              s1[j] += f[j];
              s2[j] = fmaf(f[j], f[j], s2[j]);
            }
          }
          if (AFFINE) {   // inference-only instantiation: the training kernels carry none of this
            const float* sa = s_aff + n0 + ch * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaf(f[j], sa[j], sa[a.Cout + j]);
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 pk;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[g * 8 + 2 * j], f[g * 8 + 2 * j + 1]);
            const int col = ch * 32 + g * 8;
            const int oc = col / OCH, cidx = (col % OCH) / 8;
            *reinterpret_cast<uint4*>(staging + oc * OCHUNK + swz_off<OROWB>(row, cidx)) = pk;
          }
        }
        if (PERSIST && persist) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            p1[j] = s1[j];
            p2[j] = s2[j];
          }
        } else if (SUMS) {
          if (PERSIST) {     // several N tiles per CTA: the running sums are per tile after all
#pragma unroll
            for (int j = 0; j < 32; ++j) p1[j] = p2[j] = 0.f;
          }
          // transpose-reduce: after the 5 steps lane j holds the sum over the warp's 32 pixels of channel j
#pragma unroll
          for (int S = 16; S >= 1; S >>= 1) {
            const bool up = (lane & S) != 0;
#pragma unroll
            for (int k = 0; k < S; ++k) {
              const float send1 = up ? s1[k] : s1[k + S], keep1 = up ? s1[k + S] : s1[k];
              const float send2 = up ? s2[k] : s2[k + S], keep2 = up ? s2[k + S] : s2[k];
              s1[k] = keep1 + __shfl_xor_sync(0xffffffffu, send1, S);
              s2[k] = keep2 + __shfl_xor_sync(0xffffffffu, send2, S);
            }
          }
          slot[n0 + ch * 32 + lane] += s1[0];
          slot[a.Cout + n0 + ch * 32 + lane] += s2[0];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->tempty[acc]);
      fence_proxy_async_smem();
      row_bar_sync(1, 256);
      if (et == 0) {
#pragma unroll 1
        for (int oc = 0; oc < BN / OCH; ++oc) {
          const int n = n0 + oc * OCH;
          if (MODE == EPI_LINEAR && n >= a.out_split)
            tma_store_4d(&a.out1, staging + oc * OCHUNK, n - a.out_split, x0, y0, b);
          else
            tma_store_4d(&a.out0, staging + oc * OCHUNK, n, x0, y0, b);
        }
        tma_store_commit();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (et == 0) tma_store_wait_all0();
    if (PERSIST && persist) {
#pragma unroll
      for (int S = 16; S >= 1; S >>= 1) {
        const bool up = (lane & S) != 0;
#pragma unroll
        for (int k = 0; k < S; ++k) {
          const float send1 = up ? p1[k] : p1[k + S], keep1 = up ? p1[k + S] : p1[k];
          const float send2 = up ? p2[k] : p2[k + S], keep2 = up ? p2[k + S] : p2[k];
          p1[k] = keep1 + __shfl_xor_sync(0xffffffffu, send1, S);
          p2[k] = keep2 + __shfl_xor_sync(0xffffffffu, send2, S);
        }
      }
      slot[lane] += p1[0];
      slot[a.Cout + lane] += p2[0];
    }
    if (SUMS) {
      row_bar_sync(1, 256);
      // EPI_LINEAR_BNRED: striped like bn_bwd_reduce_kernel's partial sums; the dropout factor 1 / (1 - rate) goes in here
      double* dst = RED ? a.bnred.red + (size_t)(blockIdx.x % kBnRedStripes) * 2 * a.Cout : a.stats;
      const float fac = RED ? a.bnred.keep_scale : 1.f;
      for (int c = et; c < 2 * a.Cout; c += 256) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_slot[(size_t)w * 2 * a.Cout + c];
        atomicAdd(&dst[c], (double)(t * fac));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------- host
static size_t row_fixed_bytes(int BN, int R, int Cout, int wres_bytes) {
  return 1024 + (size_t)round1k(wres_bytes) + (size_t)R * 128 * BN * 2 + 19 * (size_t)Cout * sizeof(float) +
         sizeof(RowCtl) + 64;
}
static size_t row_stage_bytes(int BN, int R, int wres) { return round1k(row_a_bytes(R)) + (wres ? 0 : 9 * BN * kPixB); }

bool conv_row_plan(int H, int W, int C0, int C1, int Cout, int mode, int out_split, int* BN, int* R, int* wres,
                   int* nst) {
  // rows are cut into 128-pixel tiles; a last partial tile loads zero-filled pixels, has its stores clipped by TMA and is
  // kept out of the BatchNorm sums.  Accepted when at least 3/4 of the tile columns are real pixels (224 = 128 + 96).
  if (W < 96 || 4 * W < 3 * ((W + 127) / 128) * 128 || C0 % 32 != 0 || C1 % 32 != 0 || Cout % 32 != 0) return false;
  int bn = (Cout % 64 == 0) ? 64 : 32;
  // a concat split on a 32-channel boundary keeps the N = 64 tile (54-cycle MMAs for twice the work of an N = 32
  // one); only the TMA store boxes shrink to 32 channels (ConvRowArgs::och)
  if (mode == EPI_LINEAR && out_split < Cout && out_split % 32 != 0) return false;
  const int nchunks = (C0 + C1) / 32;
  const int wbytes = nchunks * 9 * Cout * kPixB;
  const int wr = wbytes <= 40960 ? 1 : 0;
  for (int r : {4, 2}) {
    if (H % r != 0) continue;
    const size_t fixed = row_fixed_bytes(bn, r, Cout, wr ? wbytes : 0);
    if (fixed >= (size_t)kMaxDynSmemRow) continue;
    int n = (int)((kMaxDynSmemRow - fixed) / row_stage_bytes(bn, r, wr));
    if (n > kRowStagesMax) n = kRowStagesMax;
    if (n >= 2) {
      *BN = bn; *R = r; *wres = wr; *nst = n;
      return true;
    }
  }
  return false;
}

template <int BN, int R, int MODE, int OCH, int NEW = 8>
static int launch_row(const ConvRowArgs& a, int nst, cudaStream_t st) {
  const int nchunks = a.Ctot / 32;
  const int wres_bytes = a.wres ? nchunks * 9 * a.Cout * kPixB : 0;
  const size_t smem = row_fixed_bytes(BN, R, a.Cout, wres_bytes) + (size_t)nst * row_stage_bytes(BN, R, a.wres);
  RVIP_REQUIRE(smem <= (size_t)kMaxDynSmemRow, "conv_row: %zu bytes of shared memory needed", smem);
  static bool attr_set = false;
  if (!attr_set) {
    RVIP_CUDA(cudaFuncSetAttribute(conv3x3_row_kernel<BN, R, MODE, OCH, NEW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   kMaxDynSmemRow));
    attr_set = true;
  }
  const int grid = a.total_tiles < kNumSMs ? a.total_tiles : kNumSMs;
  launch_kernel(conv3x3_row_kernel<BN, R, MODE, OCH, NEW>, grid, 128 + 32 * NEW, smem, st, a, nst);
  RVIP_LAUNCH_CHECK();
  return 0;
}

int conv_row_launch(const ConvRowArgs& a, int BN, int R, int nst, cudaStream_t st) {
  RVIP_REQUIRE(a.och == 32 || (a.och == 64 && BN == 64), "conv_row: bad store box width %d for BN=%d", a.och, BN);
  RVIP_REQUIRE(a.mode >= EPI_RELU_STATS && a.mode <= EPI_LINEAR_BNRED, "conv_row: bad epilogue mode %d", a.mode);
  if (a.mode == EPI_LINEAR_BNRED) {
    RVIP_REQUIRE((R / 2) * (BN / 32) <= 2 && a.out_split == a.Cout && a.bnred.a && a.bnred.red,
                 "conv_row: EPI_LINEAR_BNRED needs a single output and a tile of at most two items per thread (BN=%d R=%d)",
                 BN, R);
    // the instantiations that exist: (32, 4), (32, 2), (64, 2)
    if (BN == 32 && R == 4 && a.och == 32) return launch_row<32, 4, EPI_LINEAR_BNRED, 32>(a, nst, st);
    if (BN == 32 && R == 2 && a.och == 32) return launch_row<32, 2, EPI_LINEAR_BNRED, 32>(a, nst, st);
    if (BN == 64 && R == 2 && a.och == 64) return launch_row<64, 2, EPI_LINEAR_BNRED, 64>(a, nst, st);
    set_error("conv_row: no EPI_LINEAR_BNRED instantiation for BN=%d R=%d och=%d", BN, R, a.och);
    return 1;
  }
  // 16 epilogue warps wherever the geometry allows (the statistics mode needs the layer's single N tile: every warp keeps
  // the sums of its 16 channels in registers for the whole kernel); RVIP_ROW_EPI8=1 = the 8-warp epilogue everywhere
  const bool wide = getenv("RVIP_ROW_EPI8") == nullptr && (a.mode != EPI_RELU_STATS || a.n_ntiles == 1);
#define RVIP_ROW_CASE(bn, r, oc)                                                               \
  if (BN == bn && R == r && a.och == oc && wide) {                                             \
    switch (a.mode) {                                                                          \
      case EPI_RELU_STATS: return launch_row<bn, r, EPI_RELU_STATS, oc, 16>(a, nst, st);       \
      case EPI_RELU: return launch_row<bn, r, EPI_RELU, oc, 16>(a, nst, st);                   \
      case EPI_LINEAR: return launch_row<bn, r, EPI_LINEAR, oc, 16>(a, nst, st);               \
      case EPI_RELU_AFFINE: return launch_row<bn, r, EPI_RELU_AFFINE, oc, 16>(a, nst, st);     \
      default: break;                                                                          \
    }                                                                                          \
  }                                                                                            \
  if (BN == bn && R == r && a.och == oc) {                                                     \
    switch (a.mode) {                                                                          \
      case EPI_RELU_STATS: return launch_row<bn, r, EPI_RELU_STATS, oc>(a, nst, st);           \
      case EPI_RELU: return launch_row<bn, r, EPI_RELU, oc>(a, nst, st);                       \
      case EPI_LINEAR: return launch_row<bn, r, EPI_LINEAR, oc>(a, nst, st);                   \
      case EPI_RELU_AFFINE: return launch_row<bn, r, EPI_RELU_AFFINE, oc>(a, nst, st);         \
      default: break;                                                                          \
    }                                                                                          \
  }
  RVIP_ROW_CASE(32, 4, 32)
  RVIP_ROW_CASE(32, 2, 32)
  RVIP_ROW_CASE(64, 4, 64)
  RVIP_ROW_CASE(64, 2, 64)
  RVIP_ROW_CASE(64, 4, 32)
  RVIP_ROW_CASE(64, 2, 32)
#undef RVIP_ROW_CASE
  set_error("conv_row: unsupported tile BN=%d R=%d", BN, R);
  return 1;
}

}  // namespace rvip
