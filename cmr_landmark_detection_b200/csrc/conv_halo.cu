// Halo-staged tcgen05 3x3 convolution (forward and dgrad) for the deep levels of the U-Net
// (any image size -- blocks sticking out of the image are zero-filled / clipped by TMA; input channels in chunks of
// 64; Cout a multiple of 64).
//
// The generic kernel (conv_tc.cu) moves one shifted activation box AND one weight tile per (tap, 64-channel
// chunk) for a 128-pixel tile: (128 + BN) * 128 B per 4 MMAs, i.e. 96 B/cycle at BN = 256 -- more than twice
// what L2 delivers to one SM (~43 B/cycle), so those layers ran at ~45 % of the tensor pipe.  This kernel cuts
// the L2 traffic per MMA cycle 2.6x:
//   * a CTA computes a 16 x 16 pixel block = two 128-row accumulators (left / right 8-pixel-wide halves), so
//     every weight tile fetched is used for 8 MMAs instead of 4;
//   * the activation block is staged ONCE per 64-channel chunk with its halo (TMA box {64 ch, 18, 18} ->
//     [pixel][128 B], SWIZZLE_128B) and all nine taps read it through K-major descriptors whose start address
//     is shifted by whole pixels: rows of an accumulator are 8-pixel groups (one image row each) at a uniform
//     18-pixel stride, which is exactly a stride-dimension byte offset of 18 * 128 B.  The swizzle is a function
//     of the absolute shared-memory address, so shifted views stay consistent with what TMA wrote.
// Per 64-channel chunk: 41.5 KB of activations + 9 weight tiles for 72 MMAs (36 B/cycle at BN = 256).
// Warp roles: warp 0 TMA producer of the weight tiles, warp 3 TMA producer of the activation blocks, warp 1 MMA issuer
// (elected thread, all operand offsets immediates), warp 2 TMEM allocator, warps 4-7 / 8-11 epilogue of the left / right half (tcgen05.ld -> bias + ReLU -> bf16 ->
// swizzled staging -> TMA store per 64-channel slice, plus the BatchNorm sum / sum of squares).
// Replaces tf.keras Conv2D forward and Conv2DBackpropInput (src/models/KerasLayers.py:683,689,758).
#include "conv_halo.cuh"

#include <stdlib.h>

#include "common.cuh"
#include "tc_prims.cuh"

namespace rvip {
using namespace tc;

constexpr int kHaloMaxSmem = 227 * 1024;
constexpr int kHaloMaxB = 6;                 // weight-tile ring slots
constexpr int kHaloPitch = 18;               // 16 + 2 halo pixels per row
constexpr int kHaloATx = kHaloPitch * kHaloPitch * 128;       // 41472 B delivered per activation block
constexpr int kHaloASlot = (kHaloATx + 1023) & ~1023;         // 41984

struct HaloCtl {
  uint64_t afull[2], aempty[2];
  uint64_t bfull[kHaloMaxB], bempty[kHaloMaxB];
  uint64_t tfull[2], tempty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void halo_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int BN>
struct HaloCfg {
  static constexpr int NBUF = BN <= 128 ? 2 : 1;              // TMEM accumulator buffers (2 halves x BN columns each)
  static constexpr int SB = BN <= 128 ? 2 : 1;                // staging buffers per epilogue group
  static constexpr int B_BYTES = BN * 128;                    // one (tap, chunk) weight tile
  static constexpr int STG = 128 * 128;                       // one 64-channel slice of a 128-row half, bf16
  static constexpr int TMEM_COLS = 2 * BN * NBUF;
  static_assert(TMEM_COLS <= 512, "TMEM budget");
};

// NS = 0: plain 3x3 convolution (9 taps).  NS = 2 / 3: phase-decomposed up-convolution (conv_halo.cuh): 2 x NS taps
// per phase; the taps of a phase are the halo offsets (ty0 + rr, tx0 + ss), rr < 2, ss < NS, so every operand offset
// is still an immediate on top of one per-phase shift of the activation descriptor.
template <int BN, int NS>
__global__ void __launch_bounds__(384, 1) conv3x3_halo_kernel(const __grid_constant__ ConvHaloArgs a, int nbst) {
  using Cfg = HaloCfg<BN>;
  constexpr int NSX = NS ? NS : 3;             // taps per halo row
  constexpr int NT = NS ? 2 * NS : 9;          // taps per K group
  const bool up_k = NS && a.up_dir == 1;       // phases enumerate K groups (dgrad of the up-convolution)
  const int n_eff = a.n_ntiles * ((NS && a.up_dir == 0) ? a.up_nph : 1);   // (phase, N tile) pairs per pixel block
  const int cpp = up_k ? a.up_cz / 64 : 1;     // 64-channel chunks per phase view
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_ring = smem;                                     // 2 activation blocks
  uint8_t* b_ring = a_ring + 2 * kHaloASlot;                  // nbst weight tiles
  uint8_t* staging = b_ring + (size_t)nbst * Cfg::B_BYTES;    // 2 groups x SB slices
  // the per-channel arrays exist only where the epilogue needs them (EPI_LINEAR = dgrad: none), so wide dgrads
  // (Cout = 2048 input channels at the bottom of the 5-level net) keep their weight ring
  const int auxc = a.mode == EPI_LINEAR ? 0 : a.Cout;
  float* s_part = reinterpret_cast<float*>(staging + 2 * Cfg::SB * Cfg::STG);   // [2 groups][2][Cout] sum, sum^2
  float* s_scr = s_part + 4 * auxc;                                              // [2 groups][4 warps][8][16]
  float* s_bias = s_scr + 2 * 4 * 8 * 16;                                        // [3][Cout]: bias, scale, shift
  HaloCtl* ctl = reinterpret_cast<HaloCtl*>((reinterpret_cast<uintptr_t>(s_bias + 3 * auxc) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t_kernel = a.dbg ? clock64() : 0;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&a.in0);
    prefetch_tmap(&a.in1);
    prefetch_tmap(&a.w);
    prefetch_tmap(&a.out0);
    prefetch_tmap(&a.out1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ctl->afull[i], 1);
      mbar_init(&ctl->aempty[i], 1);
      mbar_init(&ctl->tfull[i], 1);
      mbar_init(&ctl->tempty[i], 8);
    }
    for (int i = 0; i < nbst; ++i) {
      mbar_init(&ctl->bfull[i], 1);
      mbar_init(&ctl->bempty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(&ctl->tmem_base, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  pdl_wait();   // everything above is independent of the previous kernel's output
  if (warp >= 4) {
    for (int c = threadIdx.x - 128; c < 4 * auxc; c += 256) s_part[c] = 0.f;
    // the bias lives in shared memory: per-element __ldg in the epilogue exposed one L2 latency per 8 channels
    for (int c = threadIdx.x - 128; c < auxc; c += 256) {
      s_bias[c] = (a.mode != EPI_LINEAR && a.mode != EPI_LINEAR_BNRED) ? a.bias[NS ? c % a.bias_mod : c] : 0.f;
      s_bias[a.Cout + c] = a.mode == EPI_RELU_AFFINE ? a.scale[c] : 1.f;
      s_bias[2 * a.Cout + c] = a.mode == EPI_RELU_AFFINE ? a.shift[c] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const int nchunks = up_k ? a.up_nph * cpp : a.Ctot / 64;   // K groups per tile

  if (warp == 3) {
    // ------------------------------------------------------------------ TMA producer: activation blocks
    // Its own warp, so that a block is requested the moment its ring slot is released (= when the MMAs of the chunk
    // before the previous one complete), a whole chunk ahead of its use, without ever stalling the weight stream.
    if (elect_one()) {
      int as = 0, aphase = 0;
      auto load_act = [&](int tile, int c) {
        const int pt = tile / n_eff;
        const int x0 = (pt % a.tiles_x) * 16;
        const int y0 = ((pt / a.tiles_x) % a.tiles_y) * 16;
        const int b = pt / (a.tiles_x * a.tiles_y);
        const int cc = up_k ? (c % cpp) * 64 : c * 64;
        mbar_wait(&ctl->aempty[as], aphase ^ 1);
        mbar_expect_tx(&ctl->afull[as], kHaloATx);
        if (up_k)
          tma_load_4d(a_ring + as * kHaloASlot, &a.upin[c / cpp], &ctl->afull[as], cc, x0 - 1, y0 - 1, b);
        else if (cc < a.C0)
          tma_load_4d(a_ring + as * kHaloASlot, &a.in0, &ctl->afull[as], cc, x0 - 1, y0 - 1, b);
        else
          tma_load_4d(a_ring + as * kHaloASlot, &a.in1, &ctl->afull[as], cc - a.C0, x0 - 1, y0 - 1, b);
        as ^= 1;
        if (as == 0) aphase ^= 1;
      };
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x)
        for (int c = 0; c < nchunks; ++c) load_act(tile, c);
    }
  } else if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: weight tiles
    if (elect_one()) {
      int bs = 0, bphase = 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        const int n0 = (tile % n_eff) * BN;     // weight rows: (phase, N tile) pairs are consecutive row blocks
        for (int c = 0; c < nchunks; ++c) {
          // K coordinate of tap t of this group = kb + t * ks
          const int kb = up_k ? (c / cpp) * NT * a.up_cz + (c % cpp) * 64 : c * 64;
          const int ks = up_k ? a.up_cz : a.Ctot;
          for (int tap = 0; tap < NT; ++tap) {
            mbar_wait(&ctl->bempty[bs], bphase ^ 1);
            mbar_expect_tx(&ctl->bfull[bs], Cfg::B_BYTES);
            tma_load_2d(b_ring + (size_t)bs * Cfg::B_BYTES, &a.w, &ctl->bfull[bs], tap * ks + kb, n0);
            if (++bs == nbst) {
              bs = 0;
              bphase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      int as = 0, aphase = 0, bs = 0, bphase = 0, buf = 0, tphase = 0;
      // optional wait-time accounting of the issuing thread (a.dbg != nullptr): where does the main loop stall?
      long long w_t = 0, w_a = 0, w_b = 0, t_begin = a.dbg ? clock64() : 0;
      for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
        long long t0 = a.dbg ? clock64() : 0;
        mbar_wait(&ctl->tempty[buf], tphase ^ 1);
        if (a.dbg) w_t += clock64() - t0;
        tc_fence_after();
        const uint32_t d0 = tmem_base + buf * 2 * BN;
        const int ph_t = (tile % n_eff) / a.n_ntiles;   // output phase of this tile (forward up-convolution)
        for (int c = 0; c < nchunks; ++c) {
          t0 = a.dbg ? clock64() : 0;
          mbar_wait(&ctl->afull[as], aphase);
          if (a.dbg) w_a += clock64() - t0;
          tc_fence_after();
          // rows = 8-pixel groups (one image row of the half tile each), kHaloPitch pixels apart
          uint64_t adesc0 = make_smem_desc(smem_u32(a_ring + as * kHaloASlot), 16, kHaloPitch * 128, kLayoutSW128);
          if (NS) {
            // first halo offset (ty0, tx0) of the phase: forward (a, b) -> rows {a, a+1}, columns {b, b+1};
            // dgrad -> rows {1-a, 2-a}, columns {1-b, 2-b}; NS == 3 (column phase merged into channels): all 3 columns
            const int ph = up_k ? c / cpp : ph_t;
            const int pa = NS == 2 ? ph >> 1 : ph, pb = NS == 2 ? ph & 1 : 0;
            const int ty0 = up_k ? 1 - pa : pa;
            const int tx0 = NS == 2 ? (up_k ? 1 - pb : pb) : 0;
            adesc0 += (uint64_t)(((ty0 * kHaloPitch + tx0) * 128) >> 4);
          }
#pragma unroll
          for (int tap = 0; tap < NT; ++tap) {
            t0 = a.dbg ? clock64() : 0;
            mbar_wait(&ctl->bfull[bs], bphase);
            if (a.dbg) w_b += clock64() - t0;
            tc_fence_after();
            const uint64_t bdesc0 = make_smem_desc(smem_u32(b_ring + (size_t)bs * Cfg::B_BYTES), 16, 1024, kLayoutSW128);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t a_off = (uint32_t)((((tap / NSX) * kHaloPitch + (tap % NSX) + 8 * half) * 128 + k * 32) >> 4);
                mma_bf16_ss(d0 + half * BN, adesc0 + a_off, bdesc0 + 2 * k, idesc, (tap | k) != 0 ? 1u : (uint32_t)(c != 0));
              }
            }
            mma_commit(&ctl->bempty[bs]);
            if (++bs == nbst) {
              bs = 0;
              bphase ^= 1;
            }
          }
          mma_commit(&ctl->aempty[as]);
          as ^= 1;
          if (as == 0) aphase ^= 1;
        }
        mma_commit(&ctl->tfull[buf]);
        if (Cfg::NBUF == 2) {
          buf ^= 1;
          if (buf == 0) tphase ^= 1;
        } else {
          tphase ^= 1;
        }
      }
      pdl_launch_dependents();   // all MMAs issued: the next kernel may start launching behind this CTA's epilogue
      if (a.dbg) {
        a.dbg[blockIdx.x * 8 + 0] = clock64() - t_begin;
        a.dbg[blockIdx.x * 8 + 1] = w_t;
        a.dbg[blockIdx.x * 8 + 2] = w_a;
        a.dbg[blockIdx.x * 8 + 3] = w_b;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: group 0 = left half, 1 = right half
    const int grp = (warp - 4) >> 2;
    const int ew = warp & 3;                 // TMEM lane quarter
    const int r = ew * 32 + lane;            // accumulator row: pixel (y = r / 8, x = r % 8) of the half tile
    const int gt = threadIdx.x - 128 - grp * 128;   // thread index inside the group
    uint8_t* stg = staging + (size_t)grp * Cfg::SB * Cfg::STG;
    int buf = 0, tphase = 0, sb = 0;
    long long t_epi = 0;
    // EPI_LINEAR_BNRED (conv_tc.cuh): BatchNorm-backward sums of the block this gradient feeds, in the statistics mapping
    // below (thread = 8 channels x 8 rows of a staged slice); its relu(conv) values come straight from global memory,
    // requested before the accumulator slice is read so that their latency hides behind the TMEM loads and the packing
    const bool red = a.mode == EPI_LINEAR_BNRED;
    const bool lin = a.mode == EPI_LINEAR || red;
    const bool sums = a.mode == EPI_RELU_STATS || red;
    const DropKey dkey = {a.bnred.k0, a.bnred.k1};
    for (int tile = blockIdx.x; tile < a.total_tiles; tile += gridDim.x) {
      const int nt = tile % a.n_ntiles, pt = tile / n_eff;
      const int ph_t = (tile % n_eff) / a.n_ntiles;
      const int x0 = (pt % a.tiles_x) * 16 + 8 * grp;
      const int y0 = ((pt / a.tiles_x) % a.tiles_y) * 16;
      const int b = pt / (a.tiles_x * a.tiles_y);
      const int n0 = nt * BN;
      mbar_wait(&ctl->tfull[buf], tphase);
      tc_fence_after();
      const long long t_e0 = a.dbg ? clock64() : 0;
      const uint32_t acc = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * 2 * BN + grp * BN;
      // rows of the accumulator that lie outside the image (image sizes that are not multiples of 16): their stores are
      // clipped by TMA, but they must not enter the BatchNorm sums
      const bool oob = a.mode == EPI_RELU_STATS && (x0 + (r & 7) >= a.W || y0 + (r >> 3) >= a.H);
#pragma unroll 1
      for (int sl = 0; sl < BN / 64; ++sl) {
        uint8_t* sbuf = stg + (size_t)sb * Cfg::STG;
        // the staging slice must have been drained by the TMA store issued SB slices ago
        if (gt == 0) {
          if (Cfg::SB == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else tma_store_wait_read0();
        }
        halo_bar_sync(1 + grp, 128);
        uint4 auxv[8];
        if (red) {
          const int cg = gt & 7, rt = gt >> 3;
          const __nv_bfloat16* A = static_cast<const __nv_bfloat16*>(a.bnred.a);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int row = rt + k * 16, yy = y0 + (row >> 3), xx = x0 + (row & 7);
            const bool ok = yy < a.H && xx < a.W;
            const size_t pix = ((size_t)b * a.H + yy) * a.W + xx;
            auxv[k] = ok ? __ldg(reinterpret_cast<const uint4*>(A + pix * a.Cout + n0 + sl * 64 + cg * 8))
                         : make_uint4(0u, 0u, 0u, 0u);
          }
        }
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          uint32_t v[32];
          tmem_ld_32x32(acc + sl * 64 + hc * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[q * 8 + j]);
            if (!lin) {
              const float4 b0 = *reinterpret_cast<const float4*>(s_bias + n0 + sl * 64 + hc * 32 + q * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(s_bias + n0 + sl * 64 + hc * 32 + q * 8 + 4);
              const float fl = a.floor;
              f[0] = fmaxf(f[0] + b0.x, fl); f[1] = fmaxf(f[1] + b0.y, fl);
              f[2] = fmaxf(f[2] + b0.z, fl); f[3] = fmaxf(f[3] + b0.w, fl);
              f[4] = fmaxf(f[4] + b1.x, fl); f[5] = fmaxf(f[5] + b1.y, fl);
              f[6] = fmaxf(f[6] + b1.z, fl); f[7] = fmaxf(f[7] + b1.w, fl);
            }
            if (oob) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = 0.f;
            }
            if (a.mode == EPI_RELU_AFFINE) {
              const float* sc = s_bias + a.Cout + n0 + sl * 64 + hc * 32 + q * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], sc[j], sc[a.Cout + j]);
            }
            uint4 pk;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
            *reinterpret_cast<uint4*>(sbuf + swz_off<128>(r, hc * 4 + q)) = pk;
          }
        }
        if (sl == BN / 64 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ctl->tempty[buf]);   // all TMEM reads of this warp for this tile are done
        }
        fence_proxy_async_smem();
        halo_bar_sync(1 + grp, 128);
        if (gt == 0) {
          const int n = n0 + sl * 64;
          if (NS && a.up_dir == 0)
            tma_store_4d(&a.upout[ph_t], sbuf, n, x0, y0, b);
          else if (a.mode == EPI_LINEAR && n >= a.out_split)
            tma_store_4d(&a.out1, sbuf, n - a.out_split, x0, y0, b);
          else
            tma_store_4d(&a.out0, sbuf, n, x0, y0, b);
          tma_store_commit();
        }
        if (sums) {
          // per-channel sum / sum^2 of the slice from the bf16 values just staged: 8 column groups x 16 threads
          // (EPI_LINEAR_BNRED: sum keep * dy / sum keep * dy * a instead, `a` from auxv, the dropout mask replayed)
          const int cg = gt & 7, rt = gt >> 3;
          float s[8], q2[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) s[j] = q2[j] = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int row = rt + k * 16;
            const uint4 raw = *reinterpret_cast<const uint4*>(sbuf + swz_off<128>(row, cg));
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
            if (red) {
              const int yy = y0 + (row >> 3), xx = x0 + (row & 7);
              const bool ok = yy < a.H && xx < a.W;
              const uint32_t pix = ((uint32_t)b * a.H + yy) * a.W + xx;
              bool keep[8];
              if (a.bnred.thr16) {
                dropout_keep8(dkey, (pix << a.bnred.lg) | (uint32_t)((n0 + sl * 64) / 8 + cg), a.bnred.thr16, keep);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) keep[j] = true;
              }
              const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&auxv[k]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(h[j]);
                const float2 av = __bfloat1622float2(ha[j]);
                const float d0 = (ok && keep[2 * j]) ? f.x : 0.f, d1 = (ok && keep[2 * j + 1]) ? f.y : 0.f;
                s[2 * j] += d0;
                s[2 * j + 1] += d1;
                q2[2 * j] = fmaf(d0, av.x, q2[2 * j]);
                q2[2 * j + 1] = fmaf(d1, av.y, q2[2 * j + 1]);
              }
              continue;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __bfloat1622float2(h[j]);
              s[2 * j] += f.x;
              s[2 * j + 1] += f.y;
              q2[2 * j] = fmaf(f.x, f.x, q2[2 * j]);
              q2[2 * j + 1] = fmaf(f.y, f.y, q2[2 * j + 1]);
            }
          }
          // lanes with equal (lane & 7) hold the same channels
#pragma unroll
          for (int o = 16; o >= 8; o >>= 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
              q2[j] += __shfl_xor_sync(0xffffffffu, q2[j], o);
            }
          }
          // fold the group's four warps through a small scratch tile; channel ch of the slice is then owned by
          // exactly one thread of the group (plain +=, no shared-memory float atomics = CAS loops)
          float* scr = s_scr + (size_t)grp * 4 * 8 * 16;
          if (lane < 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              scr[(ew * 8 + cg) * 16 + j] = s[j];
              scr[(ew * 8 + cg) * 16 + 8 + j] = q2[j];
            }
          }
          halo_bar_sync(1 + grp, 128);
          {
            const int ch = gt & 63, which = gt >> 6;
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) v += scr[(w * 8 + (ch >> 3)) * 16 + which * 8 + (ch & 7)];
            s_part[(size_t)(grp * 2 + which) * a.Cout + n0 + sl * 64 + ch] += v;
          }
        }
        if (Cfg::SB == 2) sb ^= 1;
      }
      if (Cfg::NBUF == 2) {
        buf ^= 1;
        if (buf == 0) tphase ^= 1;
      } else {
        tphase ^= 1;
      }
      if (a.dbg) t_epi += clock64() - t_e0;
    }
    if (a.dbg && threadIdx.x == 128) a.dbg[blockIdx.x * 8 + 5] = t_epi;
    if (gt == 0) tma_store_wait_all0();
    if (sums) {
      halo_bar_sync(3, 256);
      double* dst = red ? a.bnred.red + (size_t)(blockIdx.x % kBnRedStripes) * 2 * a.Cout : a.stats;
      const double fac = red ? (double)a.bnred.keep_scale : 1.0;
      for (int c = threadIdx.x - 128; c < 2 * a.Cout; c += 256)
        atomicAdd(&dst[c], ((double)s_part[c] + (double)s_part[2 * a.Cout + c]) * fac);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (a.dbg && threadIdx.x == 0) a.dbg[blockIdx.x * 8 + 4] = clock64() - t_kernel;
}

// ------------------------------------------------------------------------------------- host
static size_t halo_fixed_bytes(int BN, int Cout, int mode) {
  const int sb = BN <= 128 ? 2 : 1;
  const size_t auxc = mode == EPI_LINEAR ? 0 : Cout;
  return 1024 + 2 * (size_t)kHaloASlot + 2 * (size_t)sb * 128 * 128 + (7 * auxc + 2 * 4 * 8 * 16) * sizeof(float) + sizeof(HaloCtl) + 64;
}

bool conv_halo_plan(int B, int H, int W, int C0, int C1, int Cout, int mode, int out_split, int* BN, int* nbst) {
  // any image size: blocks that stick out of the image load zero-filled pixels and their stores are clipped by TMA;
  // the BatchNorm statistics skip the out-of-image rows of the accumulator (the reference trains at 224 x 224, whose
  // deeper levels are 56, 28 and 14 pixels wide)
  if (H < 1 || W < 1 || C0 % 64 != 0 || C1 % 64 != 0 || Cout % 64 != 0) return false;
  if (mode == EPI_LINEAR && out_split < Cout && out_split % 64 != 0) return false;
  int bn = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
  // a wider N tile halves the activation re-reads, a narrower one fills the 148 SMs: take the narrower tile when
  // the wide one would leave more than a third of the SMs idle in the last wave
  const long mtiles = (long)B * ((H + 15) / 16) * ((W + 15) / 16);
  auto waves_eff = [&](int n) {
    const long t = mtiles * (Cout / n);
    return (double)t / (double)(((t + kNumSMs - 1) / kNumSMs) * kNumSMs);
  };
  if (bn == 256 && waves_eff(256) < 0.67 && waves_eff(128) > waves_eff(256)) bn = 128;
  // BN = 256 fills TMEM with one tile (no double buffering): with several tiles per CTA every epilogue is exposed
  if (bn == 256 && mtiles * (Cout / 256) > kNumSMs && getenv("RVIP_HALO_BN256") == nullptr) bn = 128;
  // ... and once more: 64 units of N = 128 on 148 SMs (the 256-channel dgrad at the bottom of the depth-4 net, ncu: 36 us
  // at 26 % SM throughput) become 128 units of N = 64
  if (bn == 128 && waves_eff(128) < 0.67 && waves_eff(64) > waves_eff(128) && getenv("RVIP_HALO_NO_BN64") == nullptr) bn = 64;
  const size_t fixed = halo_fixed_bytes(bn, Cout, mode);
  if (fixed >= (size_t)kHaloMaxSmem) return false;
  int n = (int)((kHaloMaxSmem - fixed) / ((size_t)bn * 128));
  if (n > kHaloMaxB) n = kHaloMaxB;
  if (n < 3) return false;
  *BN = bn; *nbst = n;
  return true;
}

template <int BN, int NS>
static int launch_halo(const ConvHaloArgs& a, int nbst, cudaStream_t st) {
  const size_t smem = halo_fixed_bytes(BN, a.Cout, a.mode) + (size_t)nbst * BN * 128;
  RVIP_REQUIRE(smem <= (size_t)kHaloMaxSmem, "conv_halo: %zu bytes of shared memory needed", smem);
  static bool attr_set = false;
  if (!attr_set) {
    RVIP_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<BN, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloMaxSmem));
    attr_set = true;
  }
  const int grid = a.total_tiles < kNumSMs ? a.total_tiles : kNumSMs;
  launch_kernel(conv3x3_halo_kernel<BN, NS>, grid, 384, smem, st, a, nbst);
  RVIP_LAUNCH_CHECK();
  return 0;
}

int conv_halo_up_variant(int h, int w, int Cin, int Cout) {
  if (h < 1 || w < 1 || Cin % 64 != 0) return 0;
  if (Cout % 64 == 0) return 2;
  if (Cout == 32) return 3;
  return 0;
}
long long conv_halo_up_pack_elems(int Cin, int Cout) {
  return Cout % 64 == 0 ? 16LL * Cin * Cout : 2LL * 6 * 64 * Cin;
}
bool conv_halo_up_plan(int B, int h, int w, int Cin, int Cout, int dir, int* BN, int* nbst, bool bnred) {
  const int ns = conv_halo_up_variant(h, w, Cin, Cout);
  if (!ns) return false;
  const int nph = ns == 2 ? 4 : 2;
  const int n_phase = ns == 2 ? Cout : 64;              // channels of one phase view
  const int N = dir == 0 ? n_phase : Cin;               // accumulator columns needed per pixel block (and phase)
  int bn = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : 64);
  const long tiles_at = (long)B * ((h + 15) / 16) * ((w + 15) / 16) * (dir == 0 ? nph : 1);
  auto waves_eff = [&](int n) {
    const long t = tiles_at * (N / n);
    return (double)t / (double)(((t + kNumSMs - 1) / kNumSMs) * kNumSMs);
  };
  if (bn == 256 && waves_eff(256) < 0.67 && waves_eff(128) > waves_eff(256)) bn = 128;
  if (bn == 256 && tiles_at * (N / 256) > kNumSMs) bn = 128;
  // narrower N tiles need less shared memory (8 KB per weight-ring slot at BN = 64): fall back until one fits
  for (; bn >= 64; bn >>= 1) {
    if (N % bn != 0) continue;
    const size_t fixed = halo_fixed_bytes(bn, N, dir == 0 ? EPI_RELU : (bnred ? EPI_LINEAR_BNRED : EPI_LINEAR));
    if (fixed >= (size_t)kHaloMaxSmem) continue;
    int n = (int)((kHaloMaxSmem - fixed) / ((size_t)bn * 128));
    if (n > kHaloMaxB) n = kHaloMaxB;
    if (n < 3) continue;
    *BN = bn; *nbst = n;
    return true;
  }
  return false;
}

int conv_halo_launch(const ConvHaloArgs& a, int BN, int nbst, cudaStream_t st) {
  RVIP_REQUIRE(a.up_ns == 0 || a.up_ns == 2 || a.up_ns == 3, "conv_halo: bad up-convolution variant %d", a.up_ns);
  RVIP_REQUIRE(a.C0 % 64 == 0 && a.Ctot % 64 == 0 && a.Cout % BN == 0 && a.tiles_x == (a.W + 15) / 16 &&
                   a.tiles_y == (a.H + 15) / 16,
               "conv_halo: bad shape %dx%d C0=%d Ctot=%d Cout=%d BN=%d", a.H, a.W, a.C0, a.Ctot, a.Cout, BN);
  if (a.up_ns == 2) {
    if (BN == 256) return launch_halo<256, 2>(a, nbst, st);
    if (BN == 128) return launch_halo<128, 2>(a, nbst, st);
    if (BN == 64) return launch_halo<64, 2>(a, nbst, st);
  } else if (a.up_ns == 3) {
    if (BN == 64) return launch_halo<64, 3>(a, nbst, st);
  } else {
    if (BN == 256) return launch_halo<256, 0>(a, nbst, st);
    if (BN == 128) return launch_halo<128, 0>(a, nbst, st);
    if (BN == 64) return launch_halo<64, 0>(a, nbst, st);
  }
  set_error("conv_halo: unsupported N tile %d", BN);
  return 1;
}

}  // namespace rvip
