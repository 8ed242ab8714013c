// Row-tiled tcgen05 weight gradient for the wide, few-channel layers (W % 128 == 0).
//
//   dW[tap][ci][co] = sum_p x[p + off(tap), ci] * dz[p, co]
//
// Both operands are consumed MN-major exactly as TMA lands them (rows = pixels = the GEMM K dimension).
// The input block is staged once per tile WITH its halo ({32 ch, 130 px, R+2 rows} -> [pixel][64 B]); for
// dz row i and vertical tap dy the A operand starts at pixel (i+dy)*130 of that block, and the M dimension
// packs the horizontal taps: M = 4 x 32 channels, where "chunk" j is the same block shifted by j pixels --
// i.e. a leading-dimension byte offset of one pixel (64 B).  Chunks 0..2 are dx = -1, 0, +1; chunk 3 is
// padding whose accumulator rows are dropped.  One CTA accumulates ALL of its pixel tiles in TMEM
// (3 vertical taps x Ctot/32 chunks accumulators of 128 x BN fp32) and flushes once with vector fp32 atomics,
// so x is read ~(R+2)/R times and dz once, instead of 9..12 times in the generic wgrad kernel.
#include "conv_row.cuh"

#include "common.cuh"
#include "tc_prims.cuh"

namespace rvip {
using namespace tc;

constexpr int kWgMaxSmem = 227 * 1024;
constexpr int kWgStagesMax = 4;
constexpr int kWgHaloW = 130;
constexpr int kWgPixB = 64;

struct WgCtl {
  uint64_t full[kWgStagesMax];
  uint64_t empty[kWgStagesMax];
  uint64_t dfull[2];
  uint64_t dempty[2];
  uint64_t done;
  uint32_t tmem_base;
};

__host__ __device__ constexpr int wg_round1k(int v) { return (v + 1023) & ~1023; }
__host__ __device__ constexpr int wg_x_bytes(int R) { return (R + 2) * kWgHaloW * kWgPixB; }

__device__ __forceinline__ void wg_red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <int BN, int R, bool ZERO_BASE>
__device__ __forceinline__ void wg_mma_loop(const WgradRowArgs& a, WgCtl* ctl, uint8_t* dzbuf, uint8_t* stages, int nst,
                                            int nchunks, uint32_t tmem_base_rt) {
  constexpr int X_ST = wg_round1k(wg_x_bytes(R));
  constexpr int DZ_ROWB = BN * 2;
  constexpr int DZ_BYTES = R * 128 * DZ_ROWB;
  constexpr uint64_t LAYB = BN == 64 ? kLayoutSW128 : kLayoutSW64;
  constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
  const uint32_t tmem_base = ZERO_BASE ? 0u : tmem_base_rt;
  int stage = 0, phase = 0, db = 0, dphase = 0;
  uint32_t not_first = 0;
  for (int pt = blockIdx.x; pt < a.pixel_tiles; pt += gridDim.x) {
    mbar_wait(&ctl->dfull[db], dphase);
    tc_fence_after();
    // B: dz rows, MN-major; 8-pixel groups 8*DZ_ROWB apart
    const uint64_t bdesc0 = make_smem_desc(smem_u32(dzbuf + db * DZ_BYTES), 16, 8 * DZ_ROWB, LAYB);
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(&ctl->full[stage], phase);
      tc_fence_after();
      // A: MN-major, 32-channel chunks one pixel (64 B) apart (the dx taps), 8-pixel groups 512 B apart
      const uint64_t adesc0 = make_smem_desc(smem_u32(stages + (size_t)stage * X_ST), kWgPixB, 8 * kWgPixB, kLayoutSW64);
      const uint32_t dc = tmem_base + c * 3 * BN;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
        for (int i = 0; i < R; ++i)   // 8 MMAs of 16 pixels each = one dz row, issued as one statement
          mma_bf16_ss_k8<((16 * kWgPixB) >> 4), ((16 * DZ_ROWB) >> 4)>(
              dc + dy * BN, adesc0 + (uint64_t)((((i + dy) * kWgHaloW) * kWgPixB) >> 4),
              bdesc0 + (uint64_t)(((i * 128) * DZ_ROWB) >> 4), idesc, i != 0 ? 1u : not_first);
      }
      mma_commit(&ctl->empty[stage]);
      if (++stage == nst) {
        stage = 0;
        phase ^= 1;
      }
    }
    mma_commit(&ctl->dempty[db]);
    db ^= 1;
    if (db == 0) dphase ^= 1;
    not_first = 1;
  }
  mma_commit(&ctl->done);
}

template <int BN, int R>
__global__ void __launch_bounds__(256, 1) wgrad3x3_row_kernel(const __grid_constant__ WgradRowArgs a, int nst,
                                                              int tmem_cols) {
  constexpr int X_TX = wg_x_bytes(R);
  constexpr int X_ST = wg_round1k(X_TX);
  constexpr int DZ_ROWB = BN * 2;                      // 64 B (SW64) or 128 B (SW128) per pixel
  constexpr int DZ_BYTES = R * 128 * DZ_ROWB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* dzbuf = smem;                               // 2 x DZ_BYTES
  uint8_t* stages = smem + 2 * DZ_BYTES;
  WgCtl* ctl = reinterpret_cast<WgCtl*>(stages + (size_t)nst * X_ST);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = a.Ctot / 32;
  const int nt = blockIdx.y, n0 = nt * BN;

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
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->dfull[i], 1);
      mbar_init(&ctl->dempty[i], 1);
    }
    mbar_init(&ctl->done, 1);
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
  const bool has_work = (int)blockIdx.x < a.pixel_tiles;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0, phase = 0, db = 0, dphase = 0;
      for (int pt = blockIdx.x; pt < a.pixel_tiles; pt += gridDim.x) {
        const int x0 = (pt % a.tiles_x) * 128;
        const int y0 = ((pt / a.tiles_x) % a.tiles_y) * R;
        const int b = pt / (a.tiles_x * a.tiles_y);
        mbar_wait(&ctl->dempty[db], dphase ^ 1);
        mbar_expect_tx(&ctl->dfull[db], DZ_BYTES);
        tma_load_4d(dzbuf + db * DZ_BYTES, &a.dz, &ctl->dfull[db], n0, x0, y0, b);
        db ^= 1;
        if (db == 0) dphase ^= 1;
        for (int c = 0; c < nchunks; ++c) {
          mbar_wait(&ctl->empty[stage], phase ^ 1);
          mbar_expect_tx(&ctl->full[stage], X_TX);
          const int cc = c * 32;
          if (cc < a.C0)
            tma_load_4d(stages + (size_t)stage * X_ST, &a.x0, &ctl->full[stage], cc, x0 - 1, y0 - 1, b);
          else
            tma_load_4d(stages + (size_t)stage * X_ST, &a.x1, &ctl->full[stage], cc - a.C0, x0 - 1, y0 - 1, b);
          if (++stage == nst) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (has_work && elect_one()) {
      if (tmem_base == 0)
        wg_mma_loop<BN, R, true>(a, ctl, dzbuf, stages, nst, nchunks, 0u);
      else
        wg_mma_loop<BN, R, false>(a, ctl, dzbuf, stages, nst, nchunks, tmem_base);
    }
  } else if (warp >= 4 && has_work) {
    const int ew = warp - 4;
    const int m = ew * 32 + lane;           // accumulator row = (dx chunk, ci)
    const int dxi = m >> 5, ci = m & 31;
    mbar_wait(&ctl->done, 0);
    tc_fence_after();
    for (int c = 0; c < nchunks; ++c) {
      for (int dy = 0; dy < 3; ++dy) {
        float* dst = a.dw + ((size_t)((dy * 3 + dxi) * a.Ctot + c * 32 + ci) * a.Cout + n0);
#pragma unroll 1
        for (int ch = 0; ch < BN / 32; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (c * 3 + dy) * BN + ch * 32, v);
          tmem_ld_wait();
          if (dxi < 3) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              wg_red_add_v4(dst + ch * 32 + j * 4, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
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

bool wgrad_row_plan(int H, int W, int C0, int C1, int Cout, int* BN, int* R, int* nst) {
  if (W % 128 != 0 || C0 % 32 != 0 || C1 % 32 != 0 || Cout % 32 != 0) return false;
  const int nchunks = (C0 + C1) / 32;
  int bn = (Cout % 64 == 0) ? 64 : 32;
  if (nchunks * 3 * bn > 512) bn = 32;
  if (nchunks * 3 * bn > 512) return false;
  for (int r : {4, 2}) {
    if (H % r != 0) continue;
    const size_t fixed = 1024 + 2 * (size_t)r * 128 * bn * 2 + sizeof(WgCtl) + 64;
    if (fixed >= (size_t)kWgMaxSmem) continue;
    int n = (int)((kWgMaxSmem - fixed) / wg_round1k(wg_x_bytes(r)));
    if (n > kWgStagesMax) n = kWgStagesMax;
    if (n >= 2) {
      *BN = bn; *R = r; *nst = n;
      return true;
    }
  }
  return false;
}

template <int BN, int R>
static int launch_wg_row(const WgradRowArgs& a, int nst, cudaStream_t st) {
  const size_t smem = 1024 + 2 * (size_t)R * 128 * BN * 2 + (size_t)nst * wg_round1k(wg_x_bytes(R)) + sizeof(WgCtl) + 64;
  RVIP_REQUIRE(smem <= (size_t)kWgMaxSmem, "wgrad_row: %zu bytes of shared memory needed", smem);
  int cols = 32;
  while (cols < (a.Ctot / 32) * 3 * BN) cols *= 2;
  RVIP_REQUIRE(cols <= 512, "wgrad_row: accumulators exceed TMEM");
  static bool attr_set = false;
  if (!attr_set) {
    RVIP_CUDA(cudaFuncSetAttribute(wgrad3x3_row_kernel<BN, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgMaxSmem));
    attr_set = true;
  }
  int gx = kNumSMs / a.n_ntiles;
  if (gx > a.pixel_tiles) gx = a.pixel_tiles;
  if (gx < 1) gx = 1;
  launch_kernel(wgrad3x3_row_kernel<BN, R>, dim3(gx, a.n_ntiles), dim3(256), smem, st, a, nst, cols);
  RVIP_LAUNCH_CHECK();
  return 0;
}

int wgrad_row_launch(const WgradRowArgs& a, int BN, int R, int nst, cudaStream_t st) {
  if (BN == 32 && R == 4) return launch_wg_row<32, 4>(a, nst, st);
  if (BN == 32 && R == 2) return launch_wg_row<32, 2>(a, nst, st);
  if (BN == 64 && R == 4) return launch_wg_row<64, 4>(a, nst, st);
  if (BN == 64 && R == 2) return launch_wg_row<64, 2>(a, nst, st);
  set_error("wgrad_row: unsupported tile BN=%d R=%d", BN, R);
  return 1;
}

}  // namespace rvip
