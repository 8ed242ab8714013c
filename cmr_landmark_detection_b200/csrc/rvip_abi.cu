// C ABI (include/rvip.h): network plan, workspace layout, TMA descriptors and the launch sequences of
// the forward / backward / optimizer passes.  Host code only; every FLOP and byte moves in the kernels.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/rvip.h"
#include "conv_row.cuh"
#include "wgrad_halo.cuh"
#include "conv_halo.cuh"
#include "kernels.cuh"

namespace rvip {

static thread_local char g_err[1024] = "";
bool pdl_enabled() {
  // On by default (RVIP_NO_PDL=1 disables).  With every kernel triggering at its END it measured neutral; with the
  // tcgen05 kernels triggering as soon as their last MMA is issued -- so the next kernel's launch, barrier / TMEM
  // setup and (for the streaming kernels, which fit beside a conv CTA) coefficient prologue run behind the epilogue --
  // the bench step gains ~2 % (5.69 -> 5.58 ms, alternating runs on one box).
  static const bool on = getenv("RVIP_NO_PDL") == nullptr;
  return on;
}
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------- TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
static int swizzle_for(int inner_bytes, CUtensorMapSwizzle* sw) {
  if (inner_bytes == 128) *sw = CU_TENSOR_MAP_SWIZZLE_128B;
  else if (inner_bytes == 64) *sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else {
    set_error("tensor map: inner box of %d bytes has no matching swizzle mode", inner_bytes);
    return 1;
  }
  return 0;
}
// NHWC bf16 activation viewed as (C, W, H, B); box = {boxC, TW, TH, NB}
static int make_act_map(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int boxC, const TileGeom& g) {
  EncodeTiledFn enc = get_encode();
  RVIP_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  CUtensorMapSwizzle sw;
  if (swizzle_for(boxC * 2, &sw)) return 1;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)g.TW, (cuuint32_t)g.TH, (cuuint32_t)g.NB};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RVIP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(act %dx%dx%dx%d box %d,%d,%d,%d) failed: %d", B, H, W, C,
               boxC, g.TW, g.TH, g.NB, (int)r);
  return 0;
}
// packed weights [rows][K] bf16; box = {boxK, boxRows}
static int make_w_map(CUtensorMap* m, const void* ptr, int rows, int K, int boxK, int boxRows) {
  EncodeTiledFn enc = get_encode();
  RVIP_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  CUtensorMapSwizzle sw;
  if (swizzle_for(boxK * 2, &sw)) return 1;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)boxK, (cuuint32_t)boxRows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RVIP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights %dx%d box %d,%d) failed: %d", rows, K, boxK,
               boxRows, (int)r);
  return 0;
}

// NHWC bf16 activation with an explicit box {boxC, bw, bh, bb}
static int make_act_map_box(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int boxC, int bw, int bh,
                            int bb) {
  TileGeom g;
  g.TW = bw; g.TH = bh; g.NB = bb;
  g.tiles_x = g.tiles_y = g.tiles_b = 0; g.full = 0;
  return make_act_map(m, ptr, B, H, W, C, boxC, g);
}

static int setup_conv_row(ConvRowArgs* a, int BN, int R, int wres, const void* in0, const void* in1, int C0, int C1,
                          const void* wpk, void* out0, void* out1, int out_split, int B, int H, int W, int Cout, int mode,
                          const float* bias, double* stats) {
  const int Ctot = C0 + C1;
  a->B = B; a->H = H; a->W = W; a->C0 = C0; a->Ctot = Ctot; a->Cout = Cout;
  a->n_ntiles = Cout / BN;
  a->tiles_x = (W + 127) / 128; a->tiles_y = H / R;
  a->total_tiles = a->n_ntiles * a->tiles_x * a->tiles_y * B;
  a->mode = mode;
  a->out_split = (mode == EPI_LINEAR) ? out_split : Cout;
  a->wres = wres;
  a->floor = 0.f;
  memset(&a->bnred, 0, sizeof(a->bnred));
  a->base_offset_mode = 0;
  a->bias = bias; a->stats = stats;
  a->dbg = nullptr;
  if (make_act_map_box(&a->in0, in0, B, H, W, C0, 32, 130, R + 2, 1)) return 1;
  if (C1 > 0) {
    if (make_act_map_box(&a->in1, in1, B, H, W, C1, 32, 130, R + 2, 1)) return 1;
  } else {
    a->in1 = a->in0;
  }
  if (make_w_map(&a->w, wpk, Cout, 9 * Ctot, 32, BN)) return 1;
  const bool split32 = mode == EPI_LINEAR && out_split < Cout && out_split % 64 != 0;
  const int och = (BN >= 64 && !split32) ? 64 : 32;
  a->och = och;
  const int c_out0 = (mode == EPI_LINEAR) ? a->out_split : Cout;
  if (make_act_map_box(&a->out0, out0, B, H, W, c_out0, och, 128, R, 1)) return 1;
  if (mode == EPI_LINEAR && out_split < Cout) {
    if (make_act_map_box(&a->out1, out1, B, H, W, Cout - out_split, och, 128, R, 1)) return 1;
  } else {
    a->out1 = a->out0;
  }
  return 0;
}

static int setup_wgrad_row(WgradRowArgs* a, int BN, int R, const void* x0, const void* x1, int C0, int C1,
                           const void* dz, float* dw, int B, int H, int W, int Cout) {
  a->B = B; a->H = H; a->W = W; a->C0 = C0; a->Ctot = C0 + C1; a->Cout = Cout;
  a->n_ntiles = Cout / BN;
  a->tiles_x = W / 128; a->tiles_y = H / R;
  a->pixel_tiles = a->tiles_x * a->tiles_y * B;
  a->dw = dw;
  if (make_act_map_box(&a->x0, x0, B, H, W, C0, 32, 130, R + 2, 1)) return 1;
  if (C1 > 0) {
    if (make_act_map_box(&a->x1, x1, B, H, W, C1, 32, 130, R + 2, 1)) return 1;
  } else {
    a->x1 = a->x0;
  }
  if (make_act_map_box(&a->dz, dz, B, H, W, Cout, BN, 128, R, 1)) return 1;
  return 0;
}

static int setup_conv_halo(ConvHaloArgs* a, int BN, const void* in0, const void* in1, int C0, int C1, const void* wpk,
                           void* out0, void* out1, int out_split, int B, int H, int W, int Cout, int mode,
                           const float* bias, double* stats) {
  const int Ctot = C0 + C1;
  a->B = B; a->H = H; a->W = W; a->C0 = C0; a->Ctot = Ctot; a->Cout = Cout;
  a->n_ntiles = Cout / BN;
  a->tiles_x = (W + 15) / 16; a->tiles_y = (H + 15) / 16;
  a->total_tiles = a->n_ntiles * a->tiles_x * a->tiles_y * B;
  a->mode = mode;
  a->out_split = (mode == EPI_LINEAR) ? out_split : Cout;
  a->bias = bias; a->stats = stats;
  a->dbg = nullptr;
  a->floor = 0.f;
  memset(&a->bnred, 0, sizeof(a->bnred));
  a->up_ns = 0; a->up_dir = 0; a->up_nph = 1; a->up_cz = 64; a->bias_mod = Cout;
  if (make_act_map_box(&a->in0, in0, B, H, W, C0, 64, 18, 18, 1)) return 1;
  if (C1 > 0) {
    if (make_act_map_box(&a->in1, in1, B, H, W, C1, 64, 18, 18, 1)) return 1;
  } else {
    a->in1 = a->in0;
  }
  if (make_w_map(&a->w, wpk, Cout, 9 * Ctot, 64, BN)) return 1;
  const int c_out0 = (mode == EPI_LINEAR) ? a->out_split : Cout;
  if (make_act_map_box(&a->out0, out0, B, H, W, c_out0, 64, 8, 16, 1)) return 1;
  if (mode == EPI_LINEAR && out_split < Cout) {
    if (make_act_map_box(&a->out1, out1, B, H, W, Cout - out_split, 64, 8, 16, 1)) return 1;
  } else {
    a->out1 = a->out0;
  }
  return 0;
}

// Phase view of a high-resolution NHWC bf16 tensor [B, 2h, 2w, C] as a low-resolution tensor (conv_halo.cuh):
//   ns = 2: view (a, b) = pixels (2i + a, 2j + b), C channels, pixel stride 2C, row stride 2 * (2w) * C
//   ns = 3: view (a)    = rows 2i + a with the column phase merged into 2C channels, pixel stride 2C
static int make_phase_map(CUtensorMap* m, const void* ptr, int B, int h, int w, int C, int ns, int pa, int pb, int bw,
                          int bh, int boxC = 64) {
  EncodeTiledFn enc = get_encode();
  RVIP_REQUIRE(enc, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  CUtensorMapSwizzle sw;
  if (swizzle_for(boxC * 2, &sw)) return 1;
  const int Cv = ns == 2 ? C : 2 * C;
  const size_t Wh = 2 * (size_t)w, Hh = 2 * (size_t)h;
  const uint8_t* base = static_cast<const uint8_t*>(ptr) + ((size_t)pa * Wh + (ns == 2 ? pb : 0)) * C * 2;
  cuuint64_t dims[4] = {(cuuint64_t)Cv, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)2 * C * 2, (cuuint64_t)2 * Wh * C * 2, (cuuint64_t)Hh * Wh * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)bw, (cuuint32_t)bh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<uint8_t*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RVIP_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(phase view %dx%dx%dx%d ns %d) failed: %d", B, h, w, C, ns,
               (int)r);
  return 0;
}

// dir 0: u[B,2h,2w,C] = relu(conv3x3(upsample2x(x[B,h,w,Cin])) + bias) from the packed forward copy;
// dir 1: dx[B,h,w,Cin] = the matching input gradient from dz[B,2h,2w,C] and the packed dgrad copy.
static int setup_conv_halo_up(ConvHaloArgs* a, int dir, int BN, const void* low, const void* high, const void* wpk,
                              const float* bias, int B, int h, int w, int Cin, int C) {
  const int ns = conv_halo_up_variant(h, w, Cin, C);
  RVIP_REQUIRE(ns != 0, "up-convolution %dx%d %d->%d is not eligible for the phase-decomposed kernel", h, w, Cin, C);
  const int nph = ns == 2 ? 4 : 2, NT = 2 * ns, Cv = ns == 2 ? C : 2 * C;
  memset(a, 0, sizeof(*a));
  a->B = B; a->H = h; a->W = w;
  a->up_ns = ns; a->up_dir = dir; a->up_nph = nph; a->up_cz = Cv; a->bias_mod = C;
  a->tiles_x = (w + 15) / 16; a->tiles_y = (h + 15) / 16;
  a->bias = bias; a->stats = nullptr; a->scale = a->shift = nullptr; a->dbg = nullptr;
  if (dir == 0) {
    a->C0 = a->Ctot = Cin; a->Cout = Cv;
    a->n_ntiles = Cv / BN;
    a->total_tiles = a->n_ntiles * nph * a->tiles_x * a->tiles_y * B;
    a->mode = EPI_RELU; a->out_split = Cv;
    if (make_act_map_box(&a->in0, low, B, h, w, Cin, 64, 18, 18, 1)) return 1;
    a->in1 = a->in0;
    if (make_w_map(&a->w, wpk, nph * Cv, NT * Cin, 64, BN)) return 1;
    for (int ph = 0; ph < nph; ++ph)
      if (make_phase_map(&a->upout[ph], high, B, h, w, C, ns, ns == 2 ? ph >> 1 : ph, ph & 1, 8, 16)) return 1;
    for (int ph = nph; ph < 4; ++ph) a->upout[ph] = a->upout[0];
    a->out0 = a->out1 = a->upout[0];
    for (int ph = 0; ph < 4; ++ph) a->upin[ph] = a->in0;
  } else {
    a->C0 = a->Ctot = nph * Cv; a->Cout = Cin;
    a->n_ntiles = Cin / BN;
    a->total_tiles = a->n_ntiles * a->tiles_x * a->tiles_y * B;
    a->mode = EPI_LINEAR; a->out_split = Cin;
    for (int ph = 0; ph < nph; ++ph)
      if (make_phase_map(&a->upin[ph], high, B, h, w, C, ns, ns == 2 ? ph >> 1 : ph, ph & 1, 18, 18)) return 1;
    for (int ph = nph; ph < 4; ++ph) a->upin[ph] = a->upin[0];
    a->in0 = a->in1 = a->upin[0];
    if (make_w_map(&a->w, wpk, Cin, nph * NT * Cv, 64, BN)) return 1;
    if (make_act_map_box(&a->out0, const_cast<void*>(low), B, h, w, Cin, 64, 8, 16, 1)) return 1;
    a->out1 = a->out0;
    for (int ph = 0; ph < 4; ++ph) a->upout[ph] = a->out0;
  }
  return 0;
}

static int setup_wgrad_halo(WgradHaloArgs* a, const WgradHaloPlan& p, const void* x0, const void* x1, int C0, int C1,
                            const void* dz, float* dw, int B, int H, int W, int Cout) {
  a->B = B; a->H = H; a->W = W; a->C0 = C0; a->Ctot = C0 + C1; a->Cout = Cout;
  a->TW = p.TW; a->TH = p.TH; a->tiles_x = p.tiles_x; a->tiles_y = p.tiles_y; a->pixel_tiles = p.pixel_tiles;
  a->n_cchunks = p.n_cchunks; a->n_ntiles = p.n_ntiles; a->k_split = p.k_split;
  a->dw = dw;
  a->up = 0;
  a->transposed = 0;
  if (make_act_map_box(&a->x0, x0, B, H, W, C0, p.CIC, p.TW + 2, p.TH + 2, 1)) return 1;
  if (C1 > 0) {
    if (make_act_map_box(&a->x1, x1, B, H, W, C1, p.CIC, p.TW + 2, p.TH + 2, 1)) return 1;
  } else {
    a->x1 = a->x0;
  }
  if (make_act_map_box(&a->dz, dz, B, H, W, Cout, p.BN >= 64 ? 64 : 32, p.TW, p.TH, 1)) return 1;
  return 0;
}

// weight gradient of the phase-decomposed up-convolution: x_low [B,h,w,Cin], dz [B,2h,2w,C] -> dw [3][3][Cin][C]
static int setup_wgrad_halo_up(WgradHaloArgs* a, const WgradHaloPlan& p, const void* x_low, const void* dz, float* dw,
                               int B, int h, int w, int Cin, int C, int transposed) {
  memset(a, 0, sizeof(*a));
  a->B = B; a->H = h; a->W = w; a->C0 = Cin; a->Ctot = Cin; a->Cout = C;
  a->TW = p.TW; a->TH = p.TH; a->tiles_x = p.tiles_x; a->tiles_y = p.tiles_y; a->pixel_tiles = p.pixel_tiles;
  a->n_cchunks = p.n_cchunks; a->n_ntiles = p.n_ntiles; a->k_split = p.k_split;
  a->dw = dw;
  a->up = 1;
  a->transposed = transposed;
  if (make_act_map_box(&a->x0, x_low, B, h, w, Cin, 64, p.TW + 2, p.TH + 2, 1)) return 1;
  a->x1 = a->x0;
  for (int ph = 0; ph < 4; ++ph)
    if (make_phase_map(&a->dzp[ph], dz, B, h, w, C, 2, ph >> 1, ph & 1, p.TW, p.TH, p.BN >= 64 ? 64 : 32)) return 1;
  a->dz = a->dzp[0];
  return 0;
}

static int pick_kc(int C0, int C1) { return (C0 % 64 == 0 && C1 % 64 == 0) ? 64 : 32; }

// fills the geometry / tiling fields of a forward-style launch (also used by the single-op entry point)
static int setup_conv_tc(ConvTcArgs* a, int* KC, int* BN, const void* in0, const void* in1, int C0, int C1,
                         const void* wpk, void* out0, void* out1, int out_split, int B, int H, int W, int Cout,
                         int mode, const float* bias, double* stats) {
  const int Ctot = C0 + C1;
  RVIP_REQUIRE(C0 % 32 == 0 && C1 % 32 == 0 && Cout % 32 == 0, "conv_tc: channels must be multiples of 32 (%d+%d->%d)",
               C0, C1, Cout);
  *KC = pick_kc(C0, C1);
  a->g = make_tile_geom(B, H, W, 128);
  // widest N tile that divides the channels (measured: narrower tiles lose more to activation re-reads than they
  // gain in wave efficiency on this per-tap kernel)
  int bn = 0;
  for (int cand : {256, 128, 64, 32}) {
    if (cand > Cout || Cout % cand != 0) continue;
    if (mode == EPI_LINEAR && out_split < Cout && (cand > out_split || out_split % cand != 0)) continue;
    bn = cand;
    break;
  }
  RVIP_REQUIRE(bn >= 32, "conv_tc: no valid N tile for Cout=%d split=%d", Cout, out_split);
  *BN = bn;
  a->B = B; a->H = H; a->W = W;
  a->C0 = C0; a->Ctot = Ctot; a->Cout = Cout;
  a->n_ntiles = Cout / bn;
  a->total_tiles = a->n_ntiles * a->g.tiles_x * a->g.tiles_y * a->g.tiles_b;
  a->mode = mode;
  a->out_split = (mode == EPI_LINEAR) ? out_split : Cout;
  a->floor = 0.f;
  a->bias = bias;
  a->stats = stats;
  if (make_act_map(&a->in0, in0, B, H, W, C0, *KC, a->g)) return 1;
  if (C1 > 0) {
    if (make_act_map(&a->in1, in1, B, H, W, C1, *KC, a->g)) return 1;
  } else {
    a->in1 = a->in0;
  }
  if (make_w_map(&a->w, wpk, Cout, 9 * Ctot, *KC, bn)) return 1;
  const int och = bn >= 64 ? 64 : 32;
  const int c_out0 = (mode == EPI_LINEAR) ? a->out_split : Cout;
  if (make_act_map(&a->out0, out0, B, H, W, c_out0, och, a->g)) return 1;
  if (mode == EPI_LINEAR && out_split < Cout) {
    if (make_act_map(&a->out1, out1, B, H, W, Cout - out_split, och, a->g)) return 1;
  } else {
    a->out1 = a->out0;
  }
  return 0;
}

static int setup_wgrad_tc(WgradTcArgs* a, int* CBA, int* CBB, const void* x0, const void* x1, int C0, int C1,
                          const void* dz, float* dw, int B, int H, int W, int Cout) {
  const int Ctot = C0 + C1;
  RVIP_REQUIRE(C0 % 32 == 0 && C1 % 32 == 0 && Cout % 32 == 0,
               "wgrad_tc: channels must be multiples of 32 (%d+%d->%d)", C0, C1, Cout);
  *CBA = pick_kc(C0, C1);
  *CBB = (Cout % 64 == 0) ? 64 : 32;
  a->g = make_tile_geom(B, H, W, 64);
  a->B = B; a->H = H; a->W = W;
  a->C0 = C0; a->Ctot = Ctot; a->Cout = Cout;
  // Per 64-pixel K step a CTA streams MT*16 KB of x and 128*BN bytes of dz for MT*4 MMAs of 128 x BN x 16:
  // the wider the N tile, the fewer L2->smem bytes per MMA cycle (BN=256, MT=2: 64 B/cycle/SM).
  a->BN = std::min(Cout, 256);
  const int cpt = 128 / *CBA;
  const int nchunk = 9 * Ctot / *CBA;
  a->n_mtiles = (nchunk + cpt - 1) / cpt;
  int mt = (72 * 1024 - 128 * a->BN) / 16384;
  mt = std::max(1, std::min(mt, std::min(a->n_mtiles, 512 / a->BN)));
  a->MT = mt;
  a->n_mgroups = (a->n_mtiles + mt - 1) / mt;
  a->n_ntiles = Cout / a->BN;
  a->k_tiles = a->g.tiles_x * a->g.tiles_y * a->g.tiles_b;
  const int units = a->n_mgroups * a->n_ntiles;
  int split = (2 * kNumSMs + units - 1) / units;
  split = std::max(1, std::min(split, std::max(1, a->k_tiles / 4)));
  a->n_split = split;
  a->dw = dw;
  if (make_act_map(&a->x0, x0, B, H, W, C0, *CBA, a->g)) return 1;
  if (C1 > 0) {
    if (make_act_map(&a->x1, x1, B, H, W, C1, *CBA, a->g)) return 1;
  } else {
    a->x1 = a->x0;
  }
  if (make_act_map(&a->dz, dz, B, H, W, Cout, *CBB, a->g)) return 1;
  return 0;
}

// ------------------------------------------------------------------------------------- plan
enum KClass { KC_CONV_FWD_TC = 0, KC_CONV_DGRAD_TC, KC_CONV_WGRAD_TC, KC_CONV_SIMT, KC_BN_FWD, KC_BN_BWD, KC_HEAD,
              KC_OPTIM, KC_EXTRACT, KC_MISC };
static const char* kClassNames[RVIP_NUM_KERNEL_CLASSES] = {"conv_fwd_tcgen05", "conv_dgrad_tcgen05", "conv_wgrad_tcgen05",
                                                           "conv_cuda_core",   "bn_forward",         "bn_backward",
                                                           "head_loss",        "adam_pack",          "extract",
                                                           "memset_misc"};

struct Layer {
  std::string name;
  int C0 = 0, C1 = 0, Cout = 0, H = 0, W = 0;
  bool bn = false, first = false;      // bn: a conv_layer_fn block (conv -> [BN] -> dropout / pool / up-sample pass)
  bool has_bn = false;                 // ... whose BatchNormalization layer exists (BATCH_NORMALISATION)
  int post = POST_NONE;
  float drop = 0.f;
  uint32_t site = 0;
  // producers / consumers (indices into layers, -1 = none)
  int in0_layer = -1, in0_which = 1;   // which buffer of the producer feeds in0 (1 = y, 2 = y2, 0 = a)
  int in1_layer = -1;                  // skip connection (producer's y)
  int g0_layer = -1, g0_which = 3;     // gradient source(s) for this layer's output (3 = dx0, 4 = dx1)
  int g1_layer = -1;
  // offsets (floats) in params / state
  long long off_k = 0, off_b = 0, off_g = 0, off_be = 0, off_mm = 0, off_mv = 0;
  long long off_stat = 0;              // per-channel offset into mean/rstd/stats/red arrays
  long long pk_f = -1, pk_d = -1;      // element offsets in the packed-weight buffer
  // bound buffers
  void *a = nullptr, *y = nullptr, *y2 = nullptr, *dx0 = nullptr, *dx1 = nullptr;
  void* dz = nullptr;                  // this layer's dz scratch buffer (one of two, alternating by layer index)
  void* dz_own = nullptr;              // ... or its own, when its weight gradient is deferred (defer_wgrad)
  int defer_wgrad = 0;                 // weight gradient queued late, behind the wide encoder levels' BatchNorm passes
  int flush_deferred = 0;              // the deferred weight gradients are queued right behind this layer's own
  int dz_prev_user = -1;               // layer whose weight gradient last read this layer's dz scratch buffer
  // tensor-core launch descriptors (bf16 mode, non-first layers)
  ConvTcArgs fwd, dgrad;
  WgradTcArgs wg;
  int fKC = 0, fBN = 0, dKC = 0, dBN = 0, wCBA = 0, wCBB = 0;
  // row-tiled variants for the wide levels (rows of 128-pixel tiles)
  ConvRowArgs rfwd, rdgrad;
  WgradRowArgs rwg;
  WgradHaloArgs hwg;
  WgradHaloPlan hwp;
  ConvHaloArgs hfwd, hdgrad;
  // phase-decomposed up-convolution (decoder up-conv layers, bf16): forward reads the producer's low-resolution y,
  // dgrad writes the low-resolution gradient directly (conv_halo.cuh)
  int up_ns = 0, up_dgrad = 0;         // variant (0 = off); dgrad also phase-decomposed
  int up_wgrad = 0;                    // weight gradient from the low-resolution input (no up-sampled copy needed)
  int needs_y2 = 1;                    // POST_UPSAMPLE layers: some consumer still reads the up-sampled copy
  int tr_simt = 0;                     // ... in fp32 parity mode: CUDA-core convolution over the virtually zero-stuffed low-resolution
                                       // input (conv_simt.cu), forward weights = the fp32 rotated copy, dgrad = the raw kernel
  int transposed = 0;                  // decoder up-conv is a Conv2DTranspose(3, strides 2, 'same') (USE_UPSAMPLE falsy):
                                       // kernel layout (kh, kw, out, in); runs on the same phase kernels, other packing
  int red_fused = 0;                   // this block's BatchNorm-backward sums come out of the dgrad that produces its output
                                       // gradient (EPI_LINEAR_BNRED): no bn_bwd_reduce pass
  int feeds_up = 0;                    // this (POST_UPSAMPLE) layer's y feeds a phase-decomposed up-convolution
  int g0_lowres = 0;                   // ... and its gradient arrives at its own (low) resolution
  long long pk_uf = -1, pk_ud = -1;    // packed up-convolution operands
  ConvHaloArgs ufwd, udgrad;
  int ufBN = 0, ufNb = 0, udBN = 0, udNb = 0;
  int use_hfwd = 0, use_hdgrad = 0, hfBN = 0, hfNb = 0, hdBN = 0, hdNb = 0;
  int use_rfwd = 0, use_rdgrad = 0, use_rwg = 0, use_hwg = 0;
  int rfBN = 0, rfR = 0, rfNst = 0, rdBN = 0, rdR = 0, rdNst = 0, rwBN = 0, rwR = 0, rwNst = 0;
};

struct TensorInfo {
  std::string name;
  int is_state;
  long long offset;
  int ndim;
  int dims[4];
};

struct ProfRec {
  int cls;
  cudaEvent_t e0, e1;
  std::string tag;   // "<layer>:<op>" of the launch group
};

}  // namespace rvip

struct rvip_handle {
  rvip_cfg cfg;
  std::vector<rvip::Layer> L;     // 3x3 conv layers in Keras creation order
  int head_in = -1;               // layer feeding the head
  long long head_k = 0, head_b = 0;
  std::vector<rvip::TensorInfo> tensors;
  long long n_params = 0, n_state = 0, n_stat_ch = 0, n_packed = 0;
  // binding
  int batch = 0, training = 0, bound = 0;
  float *params = nullptr, *grads = nullptr, *bn_state = nullptr;
  uint8_t* ws = nullptr;
  float *mean = nullptr, *rstd = nullptr;
  float *aff_scale = nullptr, *aff_shift = nullptr;   // inference: BatchNorm folded into the conv epilogues
  rvip::BnEvalEntry* bn_table_dev = nullptr;
  int n_bn = 0, max_bn_c = 0;
  double *stats = nullptr, *red = nullptr;
  void *packed = nullptr, *dz = nullptr, *head_dy = nullptr;
  void* dz2[2] = {nullptr, nullptr};        // alternating dz scratch buffers (layer i uses dz2[i & 1])
  cudaStream_t side = nullptr;              // low-priority stream running the weight gradients beside the main chain
  std::vector<cudaEvent_t> ev_dz, ev_wg;    // per layer: dz ready (main) / wgrad done (side)
  int overlap_wgrad = 1;
  double* dice_sums = nullptr;              // {sum t*p, sum p, sum t} of the BCE+Dice loss
  int head_fold = 0;                        // training: BatchNorm of the last block folded into the head (head_loss.cu)
  float* head_dwa = nullptr;                // [Cin][classes] sum_p a * dlogit of the folded head
  float w_bce = 1.f, w_dice = 1.f;
  rvip::PackEntry* pack_table_dev = nullptr;
  int n_pack = 0;
  rvip::UpPackEntry* up_pack_table_dev = nullptr;
  int n_up_pack = 0;
  std::vector<std::pair<long long, long long>> buckets;   // (offset, count) in grads
  std::vector<int> bucket_after_layer;                    // bucket i completes after backward of this layer index
  std::vector<cudaEvent_t> bucket_events;
  // per bucket: slices of the two pack tables (entries are in layer order, a bucket is a contiguous run of layers)
  std::vector<int> bucket_pack0, bucket_packn, bucket_up0, bucket_upn;
  // one-shot request (rvip_set_inline_adam): rvip_train_step applies Adam + the operand re-pack bucket by bucket on the
  // weight-gradient stream as soon as a bucket's gradients are complete, hidden behind the rest of backward
  struct {
    int armed = 0;
    float *m = nullptr, *v = nullptr;
    float lr_t = 0.f, b1 = 0.f, b2 = 0.f, eps = 0.f, gs = 1.f;
  } inline_adam;
  cudaStream_t opt = nullptr;               // third stream: the inline optimizer (behind neither chain's kernels)
  cudaEvent_t ev_opt = nullptr, ev_grad = nullptr;   // optimizer work complete / main-chain gradients of a bucket final
  // profiling
  int profile = 0;
  std::vector<rvip::ProfRec> prof;
  std::string cur_tag;            // label attached to the launch groups recorded next
  std::string detail;             // per-launch-group CSV of the last rvip_profile_read
  long long launches = 0;
};

namespace rvip {

// host restatement of common.cuh: mix32 / dropout_key (the dgrad epilogues of EPI_LINEAR_BNRED take the key by value)
static uint32_t host_mix32(uint32_t v) {
  v ^= v >> 16;
  v *= 0x85ebca6bu;
  v ^= v >> 13;
  v *= 0xc2b2ae35u;
  v ^= v >> 16;
  return v;
}
static void host_dropout_key(uint64_t seed, uint32_t site, uint32_t* k0, uint32_t* k1) {
  *k0 = host_mix32((uint32_t)seed ^ (site * 0x9E3779B9u) ^ 0xa511e9b3u);
  *k1 = host_mix32((uint32_t)(seed >> 32) + site * 0x85ebca6bu + 0x6a09e667u);
}

static bool is_bf16(const rvip_handle* h) { return h->cfg.precision == 1; }
static size_t esize(const rvip_handle* h) { return is_bf16(h) ? 2 : 4; }

template <typename F>
static int timed(rvip_handle* h, int cls, int n_launch, cudaStream_t st, F&& f) {
  h->launches += n_launch;
  if (!h->profile) return f();
  ProfRec r;
  r.cls = cls;
  r.tag = h->cur_tag;
  RVIP_CUDA(cudaEventCreate(&r.e0));
  RVIP_CUDA(cudaEventCreate(&r.e1));
  RVIP_CUDA(cudaEventRecord(r.e0, st));
  int rc = f();
  RVIP_CUDA(cudaEventRecord(r.e1, st));
  h->prof.push_back(r);
  return rc;
}

// Deferred weight gradients.  A weight gradient needs the tensor pipe for 30-75 us at the deep levels, where the BatchNorm
// passes it is meant to hide behind last ~30 us: the rest of it holds the SMs (whole TMEM, ~200 KB shared memory) against
// the next dgrad.  The wide encoder levels at the END of the backward pass have the opposite imbalance (BatchNorm passes of
// 70-140 us over weight gradients of 35-70 us).  So the largest deep-level weight gradients keep their dz in a buffer of
// their own and are queued on the side stream when the backward pass reaches the first wide encoder layer.
// RVIP_DEFER_WGRAD = comma-separated layer names | "none" | unset (default rule below); RVIP_DEFER_FLUSH = layer name.
static void plan_deferred_wgrads(rvip_handle* h) {
  if (h->cfg.precision != 1 || h->L.size() < 6) return;
  const char* env = getenv("RVIP_DEFER_WGRAD");
  const char* flush_env = getenv("RVIP_DEFER_FLUSH");
  if (env && strcmp(env, "none") == 0) return;
  const int nL = (int)h->L.size();
  // flush point: the first layer of the encoder half (in backward order) on a level at least as wide as H / 2
  int flush = -1;
  for (int i = nL - 1; i >= 0; --i) {
    Layer& l = h->L[i];
    if (flush_env ? l.name == flush_env : (l.name.rfind("enc", 0) == 0 && l.H * 2 >= h->cfg.H && !l.first)) {
      flush = i;
      break;
    }
  }
  if (flush < 0) return;
  int n = 0;
  for (int i = nL - 1; i > flush; --i) {
    Layer& l = h->L[i];
    if (l.first) continue;
    bool pick;
    if (env) {
      const std::string list = std::string(",") + env + ",";
      pick = list.find("," + l.name + ",") != std::string::npos;
    } else {
      pick = false;
    }
    if (pick) {
      l.defer_wgrad = 1;
      ++n;
    }
  }
  if (n) h->L[flush].flush_deferred = 1;
}

static int build_plan(rvip_handle* h) {
  const rvip_cfg& c = h->cfg;
  RVIP_REQUIRE(c.depth >= 1 && c.depth <= RVIP_MAX_DEPTH, "DEPTH=%d not in [1,%d]", c.depth, RVIP_MAX_DEPTH);
  RVIP_REQUIRE(c.H % (1 << c.depth) == 0 && c.W % (1 << c.depth) == 0, "DIM %dx%d must be divisible by 2^DEPTH=%d",
               c.H, c.W, 1 << c.depth);
  RVIP_REQUIRE(c.filters % 32 == 0 && c.filters >= 32, "FILTERS=%d must be a multiple of 32", c.filters);
  RVIP_REQUIRE(c.classes >= 1 && c.classes <= 4, "MASK_CLASSES=%d not in [1,4]", c.classes);
  RVIP_REQUIRE(c.in_ch >= 1 && c.in_ch < 32, "IMG_CHANNELS=%d not in [1,31]", c.in_ch);
  const int d = c.depth;
  auto add = [&](const std::string& name, int C0, int C1, int Cout, int H, int W, bool bn, int post, float drop) {
    Layer l;
    l.name = name; l.C0 = C0; l.C1 = C1; l.Cout = Cout; l.H = H; l.W = W; l.bn = bn;
    l.has_bn = bn && c.batch_norm != 0;
    l.post = post; l.drop = drop;
    if (post == POST_DROPOUT && !(drop > 0.f)) l.post = POST_NONE;
    l.site = (uint32_t)h->L.size();
    h->L.push_back(l);
    return (int)h->L.size() - 1;
  };
  std::vector<int> enc_b(d);
  int f = c.filters, H = c.H, W = c.W, cin = c.in_ch, prev = -1;
  for (int l = 0; l < d; ++l) {
    const int ia = add("enc" + std::to_string(l) + ".conv_a", cin, 0, f, H, W, true, POST_DROPOUT, c.dropout[l]);
    h->L[ia].first = (l == 0);
    h->L[ia].in0_layer = prev; h->L[ia].in0_which = 2;
    const int ib = add("enc" + std::to_string(l) + ".conv_b", f, 0, f, H, W, true, POST_POOL, 0.f);
    h->L[ib].in0_layer = ia; h->L[ib].in0_which = 1;
    h->L[ia].g0_layer = ib; h->L[ia].g0_which = 3;
    enc_b[l] = ib;
    prev = ib; cin = f; f *= 2; H /= 2; W /= 2;
  }
  const int ma = add("mid.conv_a", cin, 0, f, H, W, true, POST_DROPOUT, c.dropout_mid);
  h->L[ma].in0_layer = prev; h->L[ma].in0_which = 2;
  h->L[prev].g1_layer = ma;               // pooled gradient of the deepest encoder block
  const int mb = add("mid.conv_b", f, 0, f, H, W, true, POST_UPSAMPLE, 0.f);
  h->L[mb].in0_layer = ma; h->L[mb].in0_which = 1;
  h->L[ma].g0_layer = mb;
  for (int l = 1; l < d; ++l) h->L[enc_b[l - 1]].g1_layer = enc_b[l] - 1;   // pooled gradient <- next conv_a
  int low = mb, clow = f;
  for (int l = 0; l < d; ++l) {
    f /= 2; H *= 2; W *= 2;
    const std::string p = "dec" + std::to_string(l);
    const int iu = add(p + ".upconv", clow, 0, f, H, W, false, POST_NONE, 0.f);
    h->L[iu].in0_layer = low; h->L[iu].in0_which = 2;
    h->L[low].g0_layer = iu; h->L[low].g0_which = 3;
    const int skip = enc_b[d - 1 - l];
    const int ia = add(p + ".conv_a", f, f, f, H, W, true, POST_DROPOUT, c.dropout[d - 1 - l]);
    h->L[ia].in0_layer = iu; h->L[ia].in0_which = 0;
    h->L[ia].in1_layer = skip;
    h->L[iu].g0_layer = ia; h->L[iu].g0_which = 3;
    h->L[skip].g0_layer = ia; h->L[skip].g0_which = 4;
    const int ib = add(p + ".conv_b", f, 0, f, H, W, true, l < d - 1 ? POST_UPSAMPLE : POST_NONE, 0.f);
    h->L[ib].in0_layer = ia; h->L[ib].in0_which = 1;
    h->L[ia].g0_layer = ib; h->L[ia].g0_which = 3;
    low = ib; clow = f;
  }
  h->head_in = low;
  if (is_bf16(h) && getenv("RVIP_NO_PHASED") == nullptr) {
    for (Layer& l : h->L) {
      if (l.bn || l.in0_layer < 0 || h->L[l.in0_layer].post != POST_UPSAMPLE) continue;   // decoder up-conv layers
      l.up_ns = conv_halo_up_variant(l.H / 2, l.W / 2, l.C0, l.Cout);
      int bn_ = 0, nb_ = 0;   // shared-memory feasibility does not depend on the batch size
      if (l.up_ns && !conv_halo_up_plan(1, l.H / 2, l.W / 2, l.C0, l.Cout, 0, &bn_, &nb_)) l.up_ns = 0;
      if (!l.up_ns) continue;
      l.up_dgrad = getenv("RVIP_NO_PHASED_DGRAD") == nullptr &&
                   conv_halo_up_plan(1, l.H / 2, l.W / 2, l.C0, l.Cout, 1, &bn_, &nb_);
      WgradHaloPlan wp;
      l.up_wgrad = l.up_dgrad && getenv("RVIP_NO_PHASED_WGRAD") == nullptr &&
                   wgrad_halo_up_plan(1, l.H / 2, l.W / 2, l.C0, l.Cout, &wp);
      h->L[l.in0_layer].feeds_up = 1;
      h->L[l.in0_layer].g0_lowres = l.up_dgrad;
      h->L[l.in0_layer].needs_y2 = !l.up_wgrad;
    }
  }
  if (!c.use_upsample) {
    for (Layer& l : h->L) {
      if (l.bn || l.in0_layer < 0 || h->L[l.in0_layer].post != POST_UPSAMPLE) continue;
      l.transposed = 1;
      if (!is_bf16(h)) {
        // fp32 parity mode: the producer keeps its low-resolution y, its gradient arrives at low resolution
        l.tr_simt = 1;
        h->L[l.in0_layer].feeds_up = 1;
        h->L[l.in0_layer].g0_lowres = 1;
        h->L[l.in0_layer].needs_y2 = 0;
        continue;
      }
      RVIP_REQUIRE(l.up_ns && l.up_dgrad && l.up_wgrad,
                   "USE_UPSAMPLE=false: %s (%dx%d, %d -> %d channels) does not fit the phase-decomposed kernels (input "
                   "channels %% 64 == 0, FILTERS 32 or a multiple of 64)",
                   l.name.c_str(), l.H, l.W, l.C0, l.Cout);
    }
  }
  // flat offsets + tensor table in model.get_weights() order
  long long po = 0, so = 0, ch = 0;
  auto push = [&](const std::string& n, int is_state, long long off, std::initializer_list<int> dims) {
    TensorInfo t;
    t.name = n; t.is_state = is_state; t.offset = off; t.ndim = (int)dims.size();
    int i = 0;
    for (int v : dims) t.dims[i++] = v;
    for (; i < 4; ++i) t.dims[i] = 1;
    h->tensors.push_back(t);
  };
  for (Layer& l : h->L) {
    const int Ct = l.C0 + l.C1;
    l.off_k = po;
    if (l.transposed) push(l.name + "/kernel", 0, po, {3, 3, l.Cout, Ct});   // Conv2DTranspose: (kh, kw, out, in)
    else push(l.name + "/kernel", 0, po, {3, 3, Ct, l.Cout});
    po += 9LL * Ct * l.Cout;
    l.off_b = po; push(l.name + "/bias", 0, po, {l.Cout}); po += l.Cout;
    if (l.has_bn) {
      l.off_g = po; push(l.name + "/bn/gamma", 0, po, {l.Cout}); po += l.Cout;
      l.off_be = po; push(l.name + "/bn/beta", 0, po, {l.Cout}); po += l.Cout;
      l.off_mm = so; push(l.name + "/bn/moving_mean", 1, so, {l.Cout}); so += l.Cout;
      l.off_mv = so; push(l.name + "/bn/moving_variance", 1, so, {l.Cout}); so += l.Cout;
      l.off_stat = ch; ch += l.Cout;
    }
  }
  const int hc = h->L[h->head_in].Cout;
  h->head_k = po; push("head/kernel", 0, po, {1, 1, hc, c.classes}); po += (long long)hc * c.classes;
  h->head_b = po; push("head/bias", 0, po, {c.classes}); po += c.classes;
  h->n_params = po; h->n_state = so; h->n_stat_ch = ch;
  // packed operand copies
  long long pk = 0;
  for (Layer& l : h->L) {
    const long long n = 9LL * (l.C0 + l.C1) * l.Cout;
    if (is_bf16(h)) {
      if (l.first) continue;
      if (l.up_ns) {
        const long long nu = conv_halo_up_pack_elems(l.C0, l.Cout);
        l.pk_uf = pk; pk += nu;
        if (l.up_dgrad) {
          l.pk_ud = pk; pk += nu;
          continue;                       // the plain 3x3 operand copies are not needed
        }
      }
      l.pk_f = pk; pk += n;
      l.pk_d = pk; pk += n;
    } else {
      if (l.first) continue;
      l.pk_d = pk; pk += n;
    }
  }
  h->n_packed = pk;
  // gradient buckets: contiguous suffixes of the flat buffer in backward order, ~4 buckets
  const long long target = std::max<long long>(1, h->n_params / 4);
  long long end = h->n_params;
  for (int i = (int)h->L.size() - 1; i >= 0; --i) {
    const long long begin = h->L[i].off_k;
    if (end - begin >= target || i == 0) {
      h->buckets.push_back({begin, end - begin});
      h->bucket_after_layer.push_back(i);
      end = begin;
    }
  }
  h->bucket_events.assign(h->buckets.size(), nullptr);
  plan_deferred_wgrads(h);
  return 0;
}

// ------------------------------------------------------------------------------------- workspace
struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(uint8_t* b) : base(b) {}
  void* take(size_t bytes) {
    off = (off + 1023) & ~size_t(1023);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

static size_t carve(rvip_handle* h, uint8_t* base, int B, int training, bool assign) {
  Carver cv(base);
  const size_t es = esize(h);
  auto T = [&](void** dst, size_t elems) {
    void* p = cv.take(elems * es);
    if (assign) *dst = p;
  };
  void* tmp;
  void** sink = &tmp;
  {
    void* p = cv.take(sizeof(float) * h->n_stat_ch);
    if (assign) h->mean = (float*)p;
    p = cv.take(sizeof(float) * h->n_stat_ch);
    if (assign) h->rstd = (float*)p;
    p = cv.take(sizeof(double) * 4);
    if (assign) h->dice_sums = (double*)p;
    p = cv.take(sizeof(float) * 4096);
    if (assign) h->head_dwa = (float*)p;
    p = cv.take(sizeof(float) * h->n_stat_ch);
    if (assign) h->aff_scale = (float*)p;
    p = cv.take(sizeof(float) * h->n_stat_ch);
    if (assign) h->aff_shift = (float*)p;
    p = cv.take(sizeof(double) * 2 * h->n_stat_ch);
    if (assign) h->stats = (double*)p;
    p = cv.take(sizeof(double) * 2 * kRedStripes * h->n_stat_ch);
    if (assign) h->red = (double*)p;

    p = cv.take((is_bf16(h) ? 2 : 4) * (size_t)std::max<long long>(h->n_packed, 1));
    if (assign) h->packed = p;
  }
  size_t max_dz = 0;
  for (Layer& l : h->L) {
    const size_t P = (size_t)B * l.H * l.W;
    T(assign ? &l.a : sink, P * l.Cout);
    if (l.bn) {
      if (l.post == POST_UPSAMPLE) {
        // the up-sampled copy feeds the plain 3x3 up-conv kernels and (training) the up-conv's weight gradient; the
        // phase-decomposed up-conv reads the low-resolution y instead
        if (assign) l.y = l.y2 = nullptr;
        if (l.feeds_up) T(assign ? &l.y : sink, P * l.Cout);
        if ((training && l.needs_y2) || !l.feeds_up) T(assign ? &l.y2 : sink, 4 * P * l.Cout);
      } else {
        T(assign ? &l.y : sink, P * l.Cout);
        if (l.post == POST_POOL) T(assign ? &l.y2 : sink, P / 4 * l.Cout);
      }
    } else if (assign) {
      l.y = l.a;   // conv + ReLU without BN: the conv output is the block output
    }
    if (training) {
      if (l.defer_wgrad) T(assign ? &l.dz_own : sink, P * l.Cout);
      else max_dz = std::max(max_dz, P * l.Cout);
      if (!l.first) {
        T(assign ? &l.dx0 : sink, ((l.up_dgrad || l.tr_simt) ? P / 4 : P) * l.C0);
        if (l.C1) T(assign ? &l.dx1 : sink, P * l.C1);
      }
    }
  }
  if (training) {
    void* p = cv.take(max_dz * es);
    if (assign) h->dz = h->dz2[0] = p;
    p = cv.take(max_dz * es);
    if (assign) h->dz2[1] = p;
    const Layer& hl = h->L[h->head_in];
    p = cv.take((size_t)B * hl.H * hl.W * hl.Cout * es);
    if (assign) h->head_dy = p;
  }
  return cv.off + 1024;
}

static const void* buffer_of(const rvip_handle* h, int layer, int which) {
  const Layer& l = h->L[layer];
  switch (which) {
    case 0: return l.a;
    case 1: return l.y;
    case 2: return l.y2;
    case 3: return l.dx0;
    default: return l.dx1;
  }
}

// Inference in bf16 mode folds BatchNorm (moving statistics) into the conv epilogue: the conv writes the block
// output y directly (where the layer has one; up-sampling layers keep writing `a`, replicated by the pass after).
static bool fused_inference(const rvip_handle* h, const Layer& l) {
  return !h->training && is_bf16(h) && l.has_bn && !h->cfg.bn_first;   // BN_FIRST: relu(affine(z)) is not an epilogue affine
}
// activation floor of a block's conv epilogue: BN_FIRST stores the pre-activation z = conv + bias
static float conv_floor(const rvip_handle* h, const Layer& l) { return (l.has_bn && h->cfg.bn_first) ? -INFINITY : 0.f; }
static void* conv_output(const rvip_handle* h, const Layer& l) {
  return (fused_inference(h, l) && l.y) ? l.y : l.a;
}

// Can the dgrad of layer i also produce the BatchNorm-backward sums of the block it feeds (EPI_LINEAR_BNRED)?  That block
// must take its whole output gradient from this one tensor (no pooling / skip sum), through at most a dropout mask.
// OPT-IN (RVIP_BNRED_FUSION=all | halo): measured at C2 it LOSES -- the level-0/1 row-kernel epilogues (8 warps, two per
// scheduler) already set the tile period, so the extra mask replay + multiply-adds cost more there (dgrad 51 -> 138 us)
// than the streaming pass they replace (52 us); at the deep levels the halo epilogues hide it better but still add
// 8-14 us against 12-15 us saved.  bn_backward 1.31 -> 1.01 ms, dgrad 0.96 -> 1.32 ms per step
// (profiles/r2u_bnred_fusion_*).  Kept as a tested variant; DESIGN section 7.
static bool bnred_candidate(const rvip_handle* h, int i) {
  const char* mode = getenv("RVIP_BNRED_FUSION");
  if (!h->training || !is_bf16(h) || h->cfg.bn_first || !mode || !(strcmp(mode, "all") == 0 || strcmp(mode, "halo") == 0))
    return false;
  const Layer& l = h->L[i];
  if (l.first || l.in0_layer < 0 || l.C1 != 0) return false;
  const Layer& c = h->L[l.in0_layer];
  if (!c.has_bn || c.g0_layer != i || c.g0_which != 3 || c.Cout != l.C0 || c.Cout % 64 != 0 && c.Cout != 32) return false;
  if (c.first && getenv("RVIP_C1_RECOMPUTE")) return false;
  if (c.post == POST_NONE || c.post == POST_DROPOUT) return c.H == l.H && c.W == l.W;
  return c.post == POST_UPSAMPLE && c.g0_lowres && l.up_dgrad;     // low-resolution gradient straight from the up-conv dgrad
}

static int build_descriptors(rvip_handle* h) {
  const int B = h->batch;
  for (size_t i = 0; i < h->L.size(); ++i) {
    h->L[i].dz = h->training ? h->dz2[i & 1] : nullptr;
    h->L[i].red_fused = 0;
  }
  if (h->training) {
    // layers with a deferred weight gradient keep dz to themselves; the others alternate between the two shared buffers
    int last_user[2] = {-1, -1};
    int slot = 0;
    for (int i = (int)h->L.size() - 1; i >= 0; --i) {
      Layer& l = h->L[i];
      if (l.defer_wgrad) {
        l.dz = l.dz_own;
        l.dz_prev_user = -1;
        continue;
      }
      l.dz = h->dz2[slot];
      l.dz_prev_user = last_user[slot];
      last_user[slot] = i;
      slot ^= 1;
    }
  }
  for (size_t i = 0; i < h->L.size(); ++i) {
    Layer& l = h->L[i];
    if (l.first || !is_bf16(h)) continue;
    if (l.up_ns) {
      const __nv_bfloat16* pk = static_cast<const __nv_bfloat16*>(h->packed);
      const Layer& lo = h->L[l.in0_layer];
      RVIP_REQUIRE(conv_halo_up_plan(B, lo.H, lo.W, l.C0, l.Cout, 0, &l.ufBN, &l.ufNb), "%s: no up-conv plan", l.name.c_str());
      if (setup_conv_halo_up(&l.ufwd, 0, l.ufBN, lo.y, l.a, pk + l.pk_uf, h->params + l.off_b, B, lo.H, lo.W, l.C0, l.Cout))
        return 1;
      if (h->training) {
        if (l.up_dgrad) {
          // with room for the per-channel arrays of EPI_LINEAR_BNRED where this gradient feeds a BatchNorm block directly
          if (bnred_candidate(h, (int)i) && conv_halo_up_plan(B, lo.H, lo.W, l.C0, l.Cout, 1, &l.udBN, &l.udNb, true))
            h->L[l.in0_layer].red_fused = 1;
          else
          RVIP_REQUIRE(conv_halo_up_plan(B, lo.H, lo.W, l.C0, l.Cout, 1, &l.udBN, &l.udNb), "%s: no up-conv dgrad plan",
                       l.name.c_str());
          if (setup_conv_halo_up(&l.udgrad, 1, l.udBN, l.dx0, l.dz, pk + l.pk_ud, nullptr, B, lo.H, lo.W, l.C0, l.Cout))
            return 1;
        }
        if (l.up_wgrad) {
          l.use_hwg = wgrad_halo_up_plan(B, lo.H, lo.W, l.C0, l.Cout, &l.hwp);
          RVIP_REQUIRE(l.use_hwg, "%s: no up-conv weight-gradient plan", l.name.c_str());
          if (setup_wgrad_halo_up(&l.hwg, l.hwp, lo.y, l.dz, h->grads + l.off_k, B, lo.H, lo.W, l.C0, l.Cout, l.transposed))
            return 1;
        } else {
          l.use_hwg = wgrad_halo_plan(B, l.H, l.W, l.C0, 0, l.Cout, &l.hwp);
          RVIP_REQUIRE(l.use_hwg, "%s: the halo weight-gradient kernel does not fit", l.name.c_str());
          if (setup_wgrad_halo(&l.hwg, l.hwp, lo.y2, nullptr, l.C0, 0, l.dz, h->grads + l.off_k, B, l.H, l.W, l.Cout))
            return 1;
        }
      }
      if (l.up_dgrad || !h->training) continue;
    }
    const void* in0 = buffer_of(h, l.in0_layer, l.in0_which);
    const void* in1 = l.in1_layer >= 0 ? buffer_of(h, l.in1_layer, 1) : nullptr;
    const __nv_bfloat16* pk = static_cast<const __nv_bfloat16*>(h->packed);
    const int mode = (l.has_bn && h->training) ? EPI_RELU_STATS : EPI_RELU;
    void* conv_out = conv_output(h, l);
    if (setup_conv_tc(&l.fwd, &l.fKC, &l.fBN, in0, in1, l.C0, l.C1, pk + l.pk_f, conv_out, nullptr, l.Cout, B, l.H, l.W,
                      l.Cout, mode, h->params + l.off_b, h->stats + 2 * l.off_stat))
      return 1;
    const bool allow_row = getenv("RVIP_NO_ROW") == nullptr;
    int wres = 0;
    l.use_rfwd = allow_row && conv_row_plan(l.H, l.W, l.C0, l.C1, l.Cout, mode, l.Cout, &l.rfBN, &l.rfR, &wres, &l.rfNst);
    if (l.use_rfwd && setup_conv_row(&l.rfwd, l.rfBN, l.rfR, wres, in0, in1, l.C0, l.C1, pk + l.pk_f, conv_out, nullptr,
                                     l.Cout, B, l.H, l.W, l.Cout, mode, h->params + l.off_b, h->stats + 2 * l.off_stat))
      return 1;
    const bool allow_halo = getenv("RVIP_NO_HALO_CONV") == nullptr;
    // both kernels stage the input once with its halo; where both apply (128-wide levels) the halo kernel wins when
    // it can run N >= 128 (the row kernel is limited to N = 64 tiles: 54-cycle MMAs instead of 66 per 2x the work)
    l.use_hfwd = allow_halo && conv_halo_plan(B, l.H, l.W, l.C0, l.C1, l.Cout, mode, l.Cout, &l.hfBN, &l.hfNb);
    if (l.use_hfwd && l.use_rfwd) {
      if (l.hfBN >= 128 && getenv("RVIP_PREFER_ROW") == nullptr) l.use_rfwd = 0;
      else l.use_hfwd = 0;
    }
    if (l.use_hfwd && setup_conv_halo(&l.hfwd, l.hfBN, in0, in1, l.C0, l.C1, pk + l.pk_f, conv_out, nullptr, l.Cout, B, l.H,
                                      l.W, l.Cout, mode, h->params + l.off_b, h->stats + 2 * l.off_stat))
      return 1;
    if (l.has_bn) {
      l.fwd.scale = l.rfwd.scale = l.hfwd.scale = h->aff_scale + l.off_stat;
      l.fwd.shift = l.rfwd.shift = l.hfwd.shift = h->aff_shift + l.off_stat;
    }
    l.fwd.floor = l.rfwd.floor = l.hfwd.floor = conv_floor(h, l);
    if (h->training) {
      const int dsplit = l.C1 ? l.C0 : l.C0 + l.C1;
      l.use_rdgrad = allow_row && conv_row_plan(l.H, l.W, l.Cout, 0, l.C0 + l.C1, EPI_LINEAR, dsplit, &l.rdBN, &l.rdR,
                                                &wres, &l.rdNst);
      if (l.use_rdgrad && setup_conv_row(&l.rdgrad, l.rdBN, l.rdR, wres, l.dz, nullptr, l.Cout, 0, pk + l.pk_d, l.dx0,
                                         l.dx1, dsplit, B, l.H, l.W, l.C0 + l.C1, EPI_LINEAR, nullptr, nullptr))
        return 1;
      const bool cand = bnred_candidate(h, (int)i);
      l.use_hdgrad = allow_halo && conv_halo_plan(B, l.H, l.W, l.Cout, 0, l.C0 + l.C1, cand ? EPI_LINEAR_BNRED : EPI_LINEAR,
                                                  dsplit, &l.hdBN, &l.hdNb);
      if (l.use_hdgrad && l.use_rdgrad) {
        if (l.hdBN >= 128 && getenv("RVIP_PREFER_ROW") == nullptr) l.use_rdgrad = 0;
        else l.use_hdgrad = 0;
      }
      // the row kernel keeps at most two (row, 32-channel chunk) items of the aux tensor per thread
      const bool row_too = cand && strcmp(getenv("RVIP_BNRED_FUSION"), "all") == 0;
      if (cand && (l.use_hdgrad ||
                   (row_too && l.use_rdgrad && (l.rdR / 2) * (l.rdBN / 32) <= 2 && (l.rdBN == 32 || l.rdR == 2))))
        h->L[l.in0_layer].red_fused = 1;
      if (l.use_hdgrad && setup_conv_halo(&l.hdgrad, l.hdBN, l.dz, nullptr, l.Cout, 0, pk + l.pk_d, l.dx0, l.dx1, dsplit,
                                          B, l.H, l.W, l.C0 + l.C1, EPI_LINEAR, nullptr, nullptr))
        return 1;
      // weight gradient: halo-staged kernel wherever it applies (fastest at every level of the bench network),
      // then the row-tiled kernel (RVIP_NO_HALO_WGRAD=1), then the generic per-tap kernel
      l.use_hwg = getenv("RVIP_NO_HALO_WGRAD") == nullptr && wgrad_halo_plan(B, l.H, l.W, l.C0, l.C1, l.Cout, &l.hwp);
      if (l.use_hwg && setup_wgrad_halo(&l.hwg, l.hwp, in0, in1, l.C0, l.C1, l.dz, h->grads + l.off_k, B, l.H, l.W,
                                        l.Cout))
        return 1;
      l.use_rwg = !l.use_hwg && allow_row && getenv("RVIP_NO_ROW_WGRAD") == nullptr &&
                  wgrad_row_plan(l.H, l.W, l.C0, l.C1, l.Cout, &l.rwBN, &l.rwR, &l.rwNst);
      if (l.use_rwg && setup_wgrad_row(&l.rwg, l.rwBN, l.rwR, in0, in1, l.C0, l.C1, l.dz, h->grads + l.off_k, B, l.H,
                                       l.W, l.Cout))
        return 1;
      if (setup_conv_tc(&l.dgrad, &l.dKC, &l.dBN, l.dz, nullptr, l.Cout, 0, pk + l.pk_d, l.dx0, l.dx1,
                        l.C1 ? l.C0 : l.C0 + l.C1, B, l.H, l.W, l.C0 + l.C1, EPI_LINEAR, nullptr, nullptr))
        return 1;
      if (setup_wgrad_tc(&l.wg, &l.wCBA, &l.wCBB, in0, in1, l.C0, l.C1, l.dz, h->grads + l.off_k, B, l.H, l.W,
                         l.Cout))
        return 1;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------- passes
// First layer in training, opt-in (RVIP_C1_RECOMPUTE=1): `a` = relu(conv(x) + b) is recomputed from the 1-channel image
// by every pass that needs it (bn.cu: c1_recompute8) instead of being written once and read three times.  Measured
// SLOWER at C2 (forward 134 -> 199 us, BN backward 140 -> 274 us): 9 FMAs + 2.25 shared-memory weight loads per output
// element make those HBM-bound passes issue-bound.  Kept as a tested variant (it saves the 134 MB buffer).
static bool c1_recompute(const rvip_handle* h, const Layer& l) {
  return l.first && l.has_bn && !h->cfg.bn_first && l.C0 == 1 && l.Cout <= 256 && l.W % 4 == 0 &&
         getenv("RVIP_C1_RECOMPUTE") != nullptr &&
         (l.post == POST_NONE || l.post == POST_DROPOUT);
}
static void set_c1_source(const rvip_handle* h, const Layer& l, BnArgs* a, const float* x) {
  a->x0 = x;
  a->w0 = h->params + l.off_k;
  a->b0 = h->params + l.off_b;
}

static int conv_forward(rvip_handle* h, Layer& l, const float* x, bool training, cudaStream_t st) {
  const int mode = (l.has_bn && training) ? EPI_RELU_STATS : (fused_inference(h, l) ? EPI_RELU_AFFINE : EPI_RELU);
  if (is_bf16(h) && !l.first) {
    if (l.up_ns) return timed(h, KC_CONV_FWD_TC, 1, st, [&] { return conv_halo_launch(l.ufwd, l.ufBN, l.ufNb, st); });
    if (l.use_rfwd) {
      l.rfwd.mode = mode;
      return timed(h, KC_CONV_FWD_TC, 1, st, [&] { return conv_row_launch(l.rfwd, l.rfBN, l.rfR, l.rfNst, st); });
    }
    if (l.use_hfwd) {
      l.hfwd.mode = mode;
      return timed(h, KC_CONV_FWD_TC, 1, st, [&] { return conv_halo_launch(l.hfwd, l.hfBN, l.hfNb, st); });
    }
    l.fwd.mode = mode;
    return timed(h, KC_CONV_FWD_TC, 1, st, [&] { return conv_tc_launch(l.fwd, l.fKC, l.fBN, st); });
  }
  if (training && c1_recompute(h, l)) {
    return timed(h, KC_CONV_SIMT, 1, st, [&] {
      return c1_stats_launch(x, h->params + l.off_k, h->params + l.off_b, h->stats + 2 * l.off_stat, h->batch, l.H, l.W,
                             l.Cout, st);
    });
  }
  if (l.first && l.C0 == 1 && l.Cout <= 256 && l.W % 4 == 0 && conv_floor(h, l) == 0.f) {
    return timed(h, KC_CONV_SIMT, 1, st, [&] {
      const bool aff = mode == EPI_RELU_AFFINE;
      return conv_c1_fwd_launch(x, h->params + l.off_k, h->params + l.off_b, conv_output(h, l), h->stats + 2 * l.off_stat,
                                h->batch, l.H, l.W, l.Cout, mode == EPI_RELU_STATS, is_bf16(h),
                                aff ? h->aff_scale + l.off_stat : nullptr, aff ? h->aff_shift + l.off_stat : nullptr, st);
    });
  }
  ConvSimtArgs a;
  a.in0 = l.first ? (const void*)x : buffer_of(h, l.in0_layer, l.in0_which);
  a.in1 = l.in1_layer >= 0 ? buffer_of(h, l.in1_layer, 1) : nullptr;
  a.w = h->params + l.off_k;
  if (l.tr_simt) {   // Conv2DTranspose: low-resolution input, flipped + (out, in)-swapped kernel copy
    a.in0 = buffer_of(h, l.in0_layer, 1);
    a.in_stuffed = 1;
    a.w = static_cast<const float*>(h->packed) + l.pk_d;
  }
  a.bias = h->params + l.off_b;
  a.out0 = l.a; a.out1 = nullptr;
  a.stats = h->stats + 2 * l.off_stat;
  a.B = h->batch; a.H = l.H; a.W = l.W; a.C0 = l.C0; a.Ctot = l.C0 + l.C1; a.Cout = l.Cout;
  a.mode = mode; a.out_split = l.Cout;
  a.floor = conv_floor(h, l);
  const int in_bf16 = l.first ? 0 : is_bf16(h);
  return timed(h, KC_CONV_SIMT, 1, st, [&] { return conv_simt_launch(a, in_bf16, is_bf16(h), st); });
}

static void fill_bn(const rvip_handle* h, const Layer& l, BnArgs* a, bool training, uint64_t seed) {
  memset(a, 0, sizeof(*a));
  a->B = h->batch; a->H = l.H; a->W = l.W; a->C = l.Cout;
  a->post = l.post;
  if (!training && l.post == POST_DROPOUT) a->post = POST_NONE;
  a->a = l.a;
  a->gamma = h->params + l.off_g;
  a->beta = h->params + l.off_be;
  a->mean = h->mean + l.off_stat;
  a->rstd = h->rstd + l.off_stat;
  a->y = l.y; a->y2 = l.y2;
  a->seed = seed; a->site = l.site;
  a->thr16 = (uint32_t)std::lround((double)l.drop * 65536.0);
  a->keep_scale = 1.f / (1.f - l.drop);
  a->identity = l.has_bn ? 0 : 1;          // BATCH_NORMALISATION false: only the dropout / pool / up-sample part remains
  a->bn_first = (l.has_bn && h->cfg.bn_first) ? 1 : 0;
}

static int forward_body(rvip_handle* h, const float* x, bool training, uint64_t seed, cudaStream_t st,
                        bool fold_head = false) {
  if (training) {
    if (timed(h, KC_MISC, 1, st, [&] {
          RVIP_CUDA(cudaMemsetAsync(h->stats, 0, sizeof(double) * 2 * h->n_stat_ch, st));
          return 0;
        }))
      return 1;
  } else if (is_bf16(h) && !h->cfg.bn_first) {
    // one launch: scale / shift of every BatchNorm layer from the moving statistics
    if (timed(h, KC_BN_FWD, 1, st, [&] {
          return bn_eval_coef_launch(h->params, h->bn_state, h->bn_table_dev, h->n_bn, h->max_bn_c, h->cfg.bn_eps,
                                     h->aff_scale, h->aff_shift, st);
        }))
      return 1;
  } else {
    int n_bn = 0;
    for (Layer& l : h->L) n_bn += l.has_bn ? 1 : 0;
    if (timed(h, KC_BN_FWD, n_bn, st, [&] {
          // moving_mean / moving_variance are interleaved per layer in bn_state; prepare per layer
          for (Layer& l : h->L)
            if (l.has_bn && bn_eval_prepare_launch(h->bn_state + l.off_mm, h->bn_state + l.off_mv, h->mean + l.off_stat,
                                               h->rstd + l.off_stat, l.Cout, h->cfg.bn_eps, st))
              return 1;
          return 0;
        }))
      return 1;
  }
  for (Layer& l : h->L) {
    h->cur_tag = l.name + ":conv_fwd";
    if (conv_forward(h, l, x, training, st)) return 1;
    if (!l.bn) continue;
    if (fold_head && &l == &h->L[h->head_in]) continue;   // normalised on the fly by the head kernel
    h->cur_tag = l.name + ":bn_fwd";
    BnArgs a;
    fill_bn(h, l, &a, training, seed);
    if (training && c1_recompute(h, l)) set_c1_source(h, l, &a, x);
    if (fused_inference(h, l)) {
      // the conv epilogue already produced y = BN(relu(conv)); only pooling / up-sampling remain
      if (a.post != POST_POOL && a.post != POST_UPSAMPLE) continue;
      if (a.post == POST_UPSAMPLE && !l.y2) continue;   // consumed at low resolution by the phase-decomposed up-conv
      a.identity = 1;
      a.a = conv_output(h, l);
      if (a.post == POST_POOL) a.y = nullptr;
    }
    if (training && l.has_bn) {
      a.stats = h->stats + 2 * l.off_stat;
      a.count = (double)h->batch * l.H * l.W;
      a.inv_count = 1.0 / a.count;
      a.momentum = h->cfg.bn_momentum;
      a.eps = h->cfg.bn_eps;
      a.mean_out = h->mean + l.off_stat;
      a.rstd_out = h->rstd + l.off_stat;
      a.mov_mean = h->bn_state + l.off_mm;
      a.mov_var = h->bn_state + l.off_mv;
    }
    if (timed(h, KC_BN_FWD, 1, st, [&] { return bn_apply_launch(a, is_bf16(h), st); })) return 1;
  }
  return 0;
}

static void fill_head(const rvip_handle* h, HeadArgs* a, float* heat) {
  memset(a, 0, sizeof(*a));
  const Layer& l = h->L[h->head_in];
  a->B = h->batch; a->H = l.H; a->W = l.W; a->Cin = l.Cout; a->NC = h->cfg.classes;
  a->y = l.y;
  a->w = h->params + h->head_k;
  a->b = h->params + h->head_b;
  a->heat = heat;
}
// training with the last block's BatchNorm folded into the head: read `a`, derive scale / shift from the conv epilogue's sums
static void fold_head_bn(const rvip_handle* h, HeadArgs* a, int publish) {
  const Layer& l = h->L[h->head_in];
  a->y = l.a;
  a->bn_stats = h->stats + 2 * l.off_stat;
  a->bn_count = (double)h->batch * l.H * l.W;
  a->bn_inv_count = 1.0 / a->bn_count;
  a->bn_momentum = h->cfg.bn_momentum;
  a->bn_eps = h->cfg.bn_eps;
  a->bn_gamma = h->params + l.off_g;
  a->bn_beta = h->params + l.off_be;
  a->bn_mean_out = h->mean + l.off_stat;
  a->bn_rstd_out = h->rstd + l.off_stat;
  a->bn_mov_mean = h->bn_state + l.off_mm;
  a->bn_mov_var = h->bn_state + l.off_mv;
  a->bn_publish = publish;
}

// Adam on gradient bucket `b` plus the re-pack of exactly the operand copies its layers own, on `st`
static int adam_pack_bucket(rvip_handle* h, size_t b, float* m, float* v, float lr_t, float b1, float b2, float eps, float gs,
                            cudaStream_t st) {
  const long long off = h->buckets[b].first, cnt = h->buckets[b].second;
  h->cur_tag = "step:adam";
  return timed(h, KC_OPTIM, 3, st, [&] {
    if (adam_launch(h->params + off, h->grads + off, m + off, v + off, (size_t)cnt, lr_t, b1, b2, eps, gs, st)) return 1;
    if (pack_weights_launch(h->params, h->packed, h->pack_table_dev + h->bucket_pack0[b], h->bucket_packn[b], is_bf16(h), st))
      return 1;
    return pack_up_launch(h->params, h->packed, h->up_pack_table_dev + h->bucket_up0[b], h->bucket_upn[b], st);
  });
}

// Backward pass.  Main chain per layer (reverse order):  BN/ReLU backward -> dz,  dgrad -> dx.  The weight gradient
// of a layer only needs dz and the stored forward input, and nothing downstream needs it before the optimizer
// (or the layer's gradient bucket): it runs on a low-priority side stream, so wgrad CTAs fill SMs the main chain
// leaves idle (kernel tails, the <= 128-CTA deep-level kernels) and share SMs with the HBM-bound BN passes, which
// need almost no shared memory.  dz alternates between two scratch buffers; before a buffer is overwritten the
// main stream waits for the wgrad that read it.  Measured 5.95 -> 5.68 ms/step.  (Serialising each wgrad behind
// its layer's dgrad so that the two 200-KB-smem kernels never compete was slower: 5.82 ms.)
static int backward_body(rvip_handle* h, const float* x, uint64_t seed, cudaStream_t st, bool fold_head) {
  const int bf = is_bf16(h);
  const int nL = (int)h->L.size();
  const bool overlap = h->overlap_wgrad && !h->profile && h->side != nullptr;
  cudaStream_t ws = overlap ? h->side : st;      // stream of the weight-gradient kernels
  size_t next_bucket = 0;
  std::vector<int> deferred;          // layers whose weight gradient has not been queued yet (Layer::defer_wgrad)
  cudaEvent_t last_ws_ev = nullptr;   // most recent event on the side stream: covers every weight gradient queued so far
  auto launch_wgrad_tc = [&](Layer& l) -> int {
    return timed(h, KC_CONV_WGRAD_TC, 1, ws, [&] {
      if (l.use_hwg) return wgrad_halo_launch(l.hwg, l.hwp.CIC, l.hwp.BN, ws);
      if (l.use_rwg) return wgrad_row_launch(l.rwg, l.rwBN, l.rwR, l.rwNst, ws);
      return wgrad_tc_launch(l.wg, l.wCBA, l.wCBB, ws);
    });
  };
  // Hands gradient bucket `b` on: every gradient in it is final once the main chain has reached this point and the side
  // stream has finished what is queued on it.  Single replica + Adam: optimizer + operand re-pack of exactly these layers
  // on a third stream behind both.  Otherwise the caller's bucket event is recorded on the SIDE stream behind the main
  // chain's position, so the main chain itself never waits for a weight gradient here.
  auto close_bucket = [&](size_t b) -> int {
    const auto& ia = h->inline_adam;
    if (!overlap) {
      if (ia.armed && adam_pack_bucket(h, b, ia.m, ia.v, ia.lr_t, ia.b1, ia.b2, ia.eps, ia.gs, st)) return 1;
      if (h->bucket_events[b]) RVIP_CUDA(cudaEventRecord(h->bucket_events[b], st));
      return 0;
    }
    RVIP_CUDA(cudaEventRecord(h->ev_grad, st));
    if (ia.armed) {
      cudaStream_t os = h->opt;
      RVIP_CUDA(cudaStreamWaitEvent(os, h->ev_grad, 0));
      if (last_ws_ev) RVIP_CUDA(cudaStreamWaitEvent(os, last_ws_ev, 0));
      if (adam_pack_bucket(h, b, ia.m, ia.v, ia.lr_t, ia.b1, ia.b2, ia.eps, ia.gs, os)) return 1;
      RVIP_CUDA(cudaEventRecord(h->ev_opt, os));
      if (h->bucket_events[b]) RVIP_CUDA(cudaEventRecord(h->bucket_events[b], os));
    } else if (h->bucket_events[b]) {
      RVIP_CUDA(cudaStreamWaitEvent(ws, h->ev_grad, 0));
      RVIP_CUDA(cudaEventRecord(h->bucket_events[b], ws));
    }
    return 0;
  };
  for (int i = nL - 1; i >= 0; --i) {
    Layer& l = h->L[i];
    const size_t P = (size_t)h->batch * l.H * l.W;
    // the weight gradient that last read the dz buffer this layer is about to overwrite must have finished
    if (overlap && l.dz_prev_user >= 0) RVIP_CUDA(cudaStreamWaitEvent(st, h->ev_wg[l.dz_prev_user], 0));
    h->cur_tag = l.name + ":bn_bwd";
    if (l.bn) {
      BnArgs a;
      fill_bn(h, l, &a, true, seed);
      if (c1_recompute(h, l)) set_c1_source(h, l, &a, x);
      if (i == h->head_in) {
        a.g0 = h->head_dy;
      } else {
        a.g0 = buffer_of(h, l.g0_layer, l.g0_which);
        if (l.post == POST_POOL) a.g1 = buffer_of(h, l.g1_layer, 3);
        if (l.g0_lowres) a.post = POST_NONE;   // the phase-decomposed up-conv dgrad already summed the 2x2 replicas
      }
      a.red = h->red + 2 * kRedStripes * l.off_stat;
      a.dz = l.dz;
      a.dgamma = h->grads + l.off_g;
      a.dbeta = h->grads + l.off_be;
      a.dbias = h->grads + l.off_b;
      // the folded head already left sum dy / sum dy * a of its input block in `red` (head_bn_finalize)
      const bool have_sums = (fold_head && i == h->head_in) || a.identity || l.red_fused;
      if (timed(h, KC_BN_BWD, have_sums ? 1 : 2, st, [&] {
            if (!have_sums && bn_bwd_reduce_launch(a, bf, st)) return 1;
            return bn_bwd_apply_launch(a, bf, st);
          }))
        return 1;
    } else {
      const void* du = buffer_of(h, l.g0_layer, l.g0_which);
      if (timed(h, KC_BN_BWD, 1, st,
                [&] { return relu_bwd_launch(l.a, du, l.dz, h->grads + l.off_b, P, l.Cout, bf, st); }))
        return 1;
    }
    if (overlap) {
      RVIP_CUDA(cudaEventRecord(h->ev_dz[i], st));
      RVIP_CUDA(cudaStreamWaitEvent(ws, h->ev_dz[i], 0));
    }
    h->cur_tag = l.name + ":conv_bwd";
    bool queued_now = true;
    if (bf && !l.first) {
      if (overlap && l.defer_wgrad) {
        deferred.push_back(i);
        queued_now = false;
      } else if (launch_wgrad_tc(l)) {
        return 1;
      }
      if (l.in0_layer >= 0 && h->L[l.in0_layer].red_fused) {
        // this gradient is dL/dy of the BatchNorm block below: its epilogue also leaves that block's backward sums
        const Layer& c = h->L[l.in0_layer];
        BnRedArgs r;
        r.a = c.a;
        r.red = h->red + 2 * kRedStripes * c.off_stat;
        const bool drop = c.post == POST_DROPOUT;
        host_dropout_key(seed, c.site, &r.k0, &r.k1);
        r.thr16 = drop ? (uint32_t)std::lround((double)c.drop * 65536.0) : 0u;
        r.keep_scale = drop ? 1.f / (1.f - c.drop) : 1.f;
        r.lg = 0;
        while ((8 << r.lg) < c.Cout) ++r.lg;
        ConvHaloArgs* ha = l.up_dgrad ? &l.udgrad : (l.use_hdgrad ? &l.hdgrad : nullptr);
        if (ha) {
          ha->mode = EPI_LINEAR_BNRED;
          ha->bnred = r;
        } else {
          l.rdgrad.mode = EPI_LINEAR_BNRED;
          l.rdgrad.bnred = r;
        }
      }
      if (timed(h, KC_CONV_DGRAD_TC, 1, st, [&] {
            if (l.up_dgrad) return conv_halo_launch(l.udgrad, l.udBN, l.udNb, st);
            if (l.use_rdgrad) return conv_row_launch(l.rdgrad, l.rdBN, l.rdR, l.rdNst, st);
            if (l.use_hdgrad) return conv_halo_launch(l.hdgrad, l.hdBN, l.hdNb, st);
            return conv_tc_launch(l.dgrad, l.dKC, l.dBN, st);
          }))
        return 1;
    } else if (l.first && l.C0 == 1 && l.Cout <= 256 && l.W % 4 == 0) {
      if (timed(h, KC_CONV_SIMT, 1, ws, [&] {
            return wgrad_c1_launch(x, l.dz, h->grads + l.off_k, h->batch, l.H, l.W, l.Cout, bf, ws);
          }))
        return 1;
    } else {
      WgradSimtArgs w;
      w.in0 = l.first ? (const void*)x : buffer_of(h, l.in0_layer, l.in0_which);
      w.in1 = l.in1_layer >= 0 ? buffer_of(h, l.in1_layer, 1) : nullptr;
      w.dz = l.dz;
      w.dw = h->grads + l.off_k;
      w.B = h->batch; w.H = l.H; w.W = l.W; w.C0 = l.C0; w.Ctot = l.C0 + l.C1; w.Cout = l.Cout;
      if (l.tr_simt) {
        w.in0 = buffer_of(h, l.in0_layer, 1);
        w.in_stuffed = 1;
        w.transposed_out = 1;
      }
      const int in_bf16 = l.first ? 0 : bf;
      if (timed(h, KC_CONV_SIMT, 1, ws, [&] { return wgrad_simt_launch(w, in_bf16, bf, ws); })) return 1;
      if (!l.first) {
        ConvSimtArgs a;
        a.in0 = l.dz; a.in1 = nullptr;
        a.w = static_cast<const float*>(h->packed) + l.pk_d;
        if (l.tr_simt) {   // the raw (kh, kw, out, in) kernel IS the rotated dgrad operand; keep the odd outputs only
          a.w = h->params + l.off_k;
          a.out_stuffed = 1;
        }
        a.bias = nullptr;
        a.out0 = l.dx0; a.out1 = l.dx1;
        a.stats = nullptr;
        a.B = h->batch; a.H = l.H; a.W = l.W; a.C0 = l.Cout; a.Ctot = l.Cout; a.Cout = l.C0 + l.C1;
        a.mode = EPI_LINEAR;
        a.out_split = l.C1 ? l.C0 : l.C0 + l.C1;
        if (timed(h, KC_CONV_SIMT, 1, st, [&] { return conv_simt_launch(a, bf, bf, st); })) return 1;
      }
    }
    if (overlap && queued_now) {
      RVIP_CUDA(cudaEventRecord(h->ev_wg[i], ws));
      last_ws_ev = h->ev_wg[i];
    }
    if (overlap && (l.flush_deferred || i == 0) && !deferred.empty()) {
      // the wide encoder levels start here: their long BatchNorm passes hide the weight gradients held back so far
      for (int j : deferred) {
        h->cur_tag = h->L[j].name + ":conv_bwd";
        if (launch_wgrad_tc(h->L[j])) return 1;
        RVIP_CUDA(cudaEventRecord(h->ev_wg[j], ws));
        last_ws_ev = h->ev_wg[j];
      }
      deferred.clear();
    }
    // a bucket is handed on once none of its layers still holds a weight gradient back
    while (next_bucket < h->buckets.size() && h->bucket_after_layer[next_bucket] >= i) {
      const int lo = h->bucket_after_layer[next_bucket];
      bool pending = false;
      for (int j : deferred) pending = pending || j >= lo;
      if (pending) break;
      if (close_bucket(next_bucket)) return 1;
      ++next_bucket;
    }
    // join at the end of the step: the side stream and the optimizer stream
    if (overlap && i == 0) {
      if (last_ws_ev) RVIP_CUDA(cudaStreamWaitEvent(st, last_ws_ev, 0));
      if (h->inline_adam.armed) RVIP_CUDA(cudaStreamWaitEvent(st, h->ev_opt, 0));
    }
  }
  h->inline_adam.armed = 0;
  return 0;
}

static int pack_weights(rvip_handle* h, cudaStream_t st) {
  return timed(h, KC_OPTIM, h->n_up_pack ? 2 : 1, st, [&] {
    if (pack_weights_launch(h->params, h->packed, h->pack_table_dev, h->n_pack, is_bf16(h), st)) return 1;
    return pack_up_launch(h->params, h->packed, h->up_pack_table_dev, h->n_up_pack, st);
  });
}

}  // namespace rvip

// =====================================================================================
// extern "C"
// =====================================================================================
using namespace rvip;

extern "C" {

const char* rvip_last_error(void) { return g_err; }
int rvip_abi_version(void) { return 1; }

int rvip_create(const rvip_cfg* cfg, rvip_handle** out) {
  RVIP_REQUIRE(cfg && out, "rvip_create: null argument");
  rvip_handle* h = new rvip_handle();
  h->cfg = *cfg;
  if (build_plan(h)) {
    delete h;
    return 1;
  }
  *out = h;
  return 0;
}

void rvip_destroy(rvip_handle* h) {
  if (!h) return;
  if (h->pack_table_dev) cudaFree(h->pack_table_dev);
  if (h->up_pack_table_dev) cudaFree(h->up_pack_table_dev);
  if (h->bn_table_dev) cudaFree(h->bn_table_dev);
  for (cudaEvent_t e : h->ev_dz) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : h->ev_wg) if (e) cudaEventDestroy(e);
  if (h->ev_opt) cudaEventDestroy(h->ev_opt);
  if (h->ev_grad) cudaEventDestroy(h->ev_grad);
  if (h->opt) cudaStreamDestroy(h->opt);
  if (h->side) cudaStreamDestroy(h->side);
  for (auto& r : h->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  delete h;
}

long long rvip_param_count(const rvip_handle* h) { return h->n_params; }
long long rvip_state_count(const rvip_handle* h) { return h->n_state; }
int rvip_num_tensors(const rvip_handle* h) { return (int)h->tensors.size(); }
int rvip_tensor_info(const rvip_handle* h, int index, char* name, int name_cap, int* is_state, long long* offset,
                     int* ndim, int dims[4]) {
  RVIP_REQUIRE(index >= 0 && index < (int)h->tensors.size(), "rvip_tensor_info: index %d out of range", index);
  const TensorInfo& t = h->tensors[index];
  if (name && name_cap > 0) {
    strncpy(name, t.name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (is_state) *is_state = t.is_state;
  if (offset) *offset = t.offset;
  if (ndim) *ndim = t.ndim;
  if (dims) memcpy(dims, t.dims, sizeof(int) * 4);
  return 0;
}

size_t rvip_workspace_bytes(const rvip_handle* h, int batch, int training) {
  return carve(const_cast<rvip_handle*>(h), nullptr, batch, training, false);
}

int rvip_bind(rvip_handle* h, float* params, float* grads, float* bn_state, void* workspace, size_t workspace_bytes,
              int batch, int training) {
  RVIP_REQUIRE(h && params && workspace && batch > 0 && (bn_state || h->n_state == 0), "rvip_bind: null/invalid argument");
  RVIP_REQUIRE(!training || grads, "rvip_bind: training needs a gradient buffer");
  const size_t need = carve(h, nullptr, batch, training, false);
  RVIP_REQUIRE(workspace_bytes >= need, "rvip_bind: workspace too small (%zu < %zu)", workspace_bytes, need);
  int dev_count = 0;
  RVIP_CUDA(cudaGetDeviceCount(&dev_count));
  RVIP_REQUIRE(dev_count > 0, "rvip_bind: no CUDA device (this library has no CPU fallback)");
  h->params = params; h->grads = grads; h->bn_state = bn_state;
  h->ws = static_cast<uint8_t*>(workspace);
  h->batch = batch; h->training = training;
  carve(h, h->ws, batch, training, true);
  // pack table
  std::vector<PackEntry> tab;
  std::vector<UpPackEntry> utab;
  std::vector<int> tab_layer, utab_layer;      // layer index of every entry
  for (size_t li = 0; li < h->L.size(); ++li) {
    const Layer& l = h->L[li];
    if (l.first) continue;
    if (l.up_ns) {
      UpPackEntry u;
      u.src = l.off_k; u.dst_f = l.pk_uf; u.dst_d = l.pk_ud; u.Cin = l.C0; u.C = l.Cout; u.ns = l.up_ns;
      u.transposed = l.transposed;
      utab.push_back(u);
      utab_layer.push_back((int)li);
    }
    if (l.pk_d < 0) continue;
    PackEntry e;
    e.src = l.off_k; e.dst_f = l.pk_f; e.dst_d = l.pk_d; e.Ctot = l.C0 + l.C1; e.Cout = l.Cout;
    if (l.tr_simt) {   // (kh, kw, out, in) read as HWIO with I = out, O = in: its rotated copy is the forward operand
      e.Ctot = l.Cout;
      e.Cout = l.C0;
    }
    tab.push_back(e);
    tab_layer.push_back((int)li);
  }
  {
    const size_t nb = h->buckets.size();
    h->bucket_pack0.assign(nb, 0); h->bucket_packn.assign(nb, 0);
    h->bucket_up0.assign(nb, 0); h->bucket_upn.assign(nb, 0);
    int hi = (int)h->L.size();                 // bucket b covers layers [bucket_after_layer[b], hi)
    for (size_t b = 0; b < nb; ++b) {
      const int lo = h->bucket_after_layer[b];
      auto slice = [&](const std::vector<int>& layers, int* first, int* count) {
        *first = -1; *count = 0;
        for (size_t k = 0; k < layers.size(); ++k)
          if (layers[k] >= lo && layers[k] < hi) {
            if (*first < 0) *first = (int)k;
            ++*count;
          }
        if (*first < 0) *first = 0;
      };
      slice(tab_layer, &h->bucket_pack0[b], &h->bucket_packn[b]);
      slice(utab_layer, &h->bucket_up0[b], &h->bucket_upn[b]);
      hi = lo;
    }
  }
  h->n_pack = (int)tab.size();
  if (h->pack_table_dev) cudaFree(h->pack_table_dev);
  h->pack_table_dev = nullptr;
  if (!tab.empty()) {
    RVIP_CUDA(cudaMalloc(&h->pack_table_dev, sizeof(PackEntry) * tab.size()));
    RVIP_CUDA(cudaMemcpy(h->pack_table_dev, tab.data(), sizeof(PackEntry) * tab.size(), cudaMemcpyHostToDevice));
  }
  h->n_up_pack = (int)utab.size();
  if (h->up_pack_table_dev) cudaFree(h->up_pack_table_dev);
  h->up_pack_table_dev = nullptr;
  if (!utab.empty()) {
    RVIP_CUDA(cudaMalloc(&h->up_pack_table_dev, sizeof(UpPackEntry) * utab.size()));
    RVIP_CUDA(cudaMemcpy(h->up_pack_table_dev, utab.data(), sizeof(UpPackEntry) * utab.size(), cudaMemcpyHostToDevice));
  }
  if (h->bn_table_dev) cudaFree(h->bn_table_dev);
  h->bn_table_dev = nullptr;
  {
    std::vector<BnEvalEntry> bt;
    h->max_bn_c = 0;
    for (const Layer& l : h->L) {
      if (!l.has_bn) continue;
      BnEvalEntry e;
      e.off_g = l.off_g; e.off_be = l.off_be; e.off_mm = l.off_mm; e.off_mv = l.off_mv; e.off_stat = l.off_stat;
      e.C = l.Cout;
      bt.push_back(e);
      h->max_bn_c = std::max(h->max_bn_c, l.Cout);
    }
    h->n_bn = (int)bt.size();
    if (!bt.empty()) {
      RVIP_CUDA(cudaMalloc(&h->bn_table_dev, sizeof(BnEvalEntry) * bt.size()));
      RVIP_CUDA(cudaMemcpy(h->bn_table_dev, bt.data(), sizeof(BnEvalEntry) * bt.size(), cudaMemcpyHostToDevice));
    }
  }
  // the phase-merged up-convolution operands have structural zeros that the packer never writes
  if (h->n_up_pack) {
    RVIP_CUDA(cudaMemset(h->packed, 0, 2 * (size_t)std::max<long long>(h->n_packed, 1)));
    RVIP_CUDA(cudaStreamSynchronize(0));   // the packer runs on the caller's (possibly non-blocking) stream
  }
  {
    const Layer& hl = h->L[h->head_in];
    // bf16 mode only: the fp32 parity mode keeps the plain pass structure (its sums of dy * a then come from the double
    // accumulators of bn_bwd_reduce instead of fp32 atomics, which the 1e-4 run-to-run tests of that mode rely on)
    h->head_fold = training && is_bf16(h) && hl.has_bn && !h->cfg.bn_first && hl.post == POST_NONE && hl.Cout * h->cfg.classes <= 4096 &&
                   getenv("RVIP_NO_HEAD_FOLD") == nullptr;
  }
  if (build_descriptors(h)) return 1;
  if (training && !h->side) {
    h->overlap_wgrad = getenv("RVIP_NO_WGRAD_OVERLAP") == nullptr;
    int lo = 0, hi = 0;
    RVIP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = least priority
    RVIP_CUDA(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, getenv("RVIP_SIDE_PRIO_HIGH") ? hi : lo));
    RVIP_CUDA(cudaEventCreateWithFlags(&h->ev_opt, cudaEventDisableTiming));
    RVIP_CUDA(cudaEventCreateWithFlags(&h->ev_grad, cudaEventDisableTiming));
    RVIP_CUDA(cudaStreamCreateWithPriority(&h->opt, cudaStreamNonBlocking, lo));
    h->ev_dz.assign(h->L.size(), nullptr);
    h->ev_wg.assign(h->L.size(), nullptr);
    for (size_t i = 0; i < h->L.size(); ++i) {
      RVIP_CUDA(cudaEventCreateWithFlags(&h->ev_dz[i], cudaEventDisableTiming));
      RVIP_CUDA(cudaEventCreateWithFlags(&h->ev_wg[i], cudaEventDisableTiming));
    }
  }
  h->bound = 1;
  return 0;
}

int rvip_pack_weights(rvip_handle* h, void* stream) {
  RVIP_REQUIRE(h && h->bound, "rvip_pack_weights: handle not bound");
  return pack_weights(h, (cudaStream_t)stream);
}

int rvip_predict(rvip_handle* h, const float* x, float* heat, void* stream) {
  RVIP_REQUIRE(h && h->bound, "rvip_predict: handle not bound");
  cudaStream_t st = (cudaStream_t)stream;
  if (forward_body(h, x, false, 0, st)) return 1;
  HeadArgs a;
  fill_head(h, &a, heat);
  return timed(h, KC_HEAD, 1, st, [&] { return head_launch(a, 0, is_bf16(h), st); });
}

int rvip_train_step(rvip_handle* h, const float* x, const float* target, const float* inplane, int loss_kind,
                    float mask_thr, uint64_t seed, float* heat, double* loss_out, void* stream) {
  RVIP_REQUIRE(h && h->bound && h->training, "rvip_train_step: handle not bound for training");
  RVIP_REQUIRE(loss_kind != RVIP_LOSS_WEIGHTED || inplane, "rvip_train_step: weighted loss needs in-plane weights");
  cudaStream_t st = (cudaStream_t)stream;
  h->cur_tag = "step:memset";
  if (timed(h, KC_MISC, 3, st, [&] {
        RVIP_CUDA(cudaMemsetAsync(h->grads, 0, sizeof(float) * h->n_params, st));
        RVIP_CUDA(cudaMemsetAsync(h->red, 0, sizeof(double) * 2 * kRedStripes * h->n_stat_ch, st));
        RVIP_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(double), st));
        return 0;
      }))
    return 1;
  const bool fold = h->head_fold != 0;
  if (forward_body(h, x, true, seed, st, fold)) return 1;
  HeadArgs a;
  fill_head(h, &a, heat);
  if (fold) {
    fold_head_bn(h, &a, 1);
    RVIP_CUDA(cudaMemsetAsync(h->head_dwa, 0, sizeof(float) * a.Cin * a.NC, st));
  }
  a.target = target; a.inplane = inplane; a.loss_kind = loss_kind; a.mask_thr = mask_thr; a.eps = 1e-7f;
  a.dy = h->head_dy;
  a.dw = fold ? h->head_dwa : h->grads + h->head_k;
  a.db = h->grads + h->head_b;
  a.loss_acc = loss_out;
  a.dice_sums = h->dice_sums; a.w_bce = h->w_bce; a.w_dice = h->w_dice;
  h->cur_tag = "head:loss";
  if (loss_kind == RVIP_LOSS_BCE_DICE) {
    // the Dice term needs batch-global sums first: inference-mode head (writes the heat map), a small reduction,
    // then the training head with the sums in hand
    if (timed(h, KC_HEAD, 3, st, [&] {
          RVIP_CUDA(cudaMemsetAsync(h->dice_sums, 0, sizeof(double) * 4, st));
          if (head_launch(a, 0, is_bf16(h), st)) return 1;
          a.bn_publish = 0;   // the inference-mode pass above already published mean / rstd and the moving statistics
          const Layer& hl = h->L[h->head_in];
          return head_dice_sums_launch(heat, target, (size_t)h->batch * hl.H * hl.W * h->cfg.classes, h->dice_sums, st);
        }))
      return 1;
  }
  if (timed(h, KC_HEAD, fold ? 2 : 1, st, [&] {
        if (head_launch(a, 1, is_bf16(h), st)) return 1;
        if (!fold) return 0;
        const Layer& hl = h->L[h->head_in];
        return head_bn_finalize_launch(h->head_dwa, a.db, a.w, a.bn_gamma, a.bn_beta, a.bn_mean_out, a.bn_rstd_out, a.Cin,
                                       a.NC, h->grads + h->head_k, h->red + 2 * kRedStripes * hl.off_stat, st);
      }))
    return 1;
  return backward_body(h, x, seed, st, fold);
}

int rvip_set_loss_weights(rvip_handle* h, float w_bce, float w_dice) {
  RVIP_REQUIRE(h, "rvip_set_loss_weights: null handle");
  h->w_bce = w_bce;
  h->w_dice = w_dice;
  return 0;
}

int rvip_adam_step(rvip_handle* h, float* m, float* v, float lr, float beta1, float beta2, float eps, long long step,
                   float grad_scale, void* stream) {
  RVIP_REQUIRE(h && h->bound && h->training && step >= 1, "rvip_adam_step: handle not bound for training");
  cudaStream_t st = (cudaStream_t)stream;
  h->cur_tag = "step:adam";
  const double lr_t = (double)lr * std::sqrt(1.0 - std::pow((double)beta2, (double)step)) /
                      (1.0 - std::pow((double)beta1, (double)step));
  if (timed(h, KC_OPTIM, 1, st, [&] {
        return adam_launch(h->params, h->grads, m, v, (size_t)h->n_params, (float)lr_t, beta1, beta2, eps, grad_scale,
                           st);
      }))
    return 1;
  return pack_weights(h, st);
}

int rvip_set_inline_adam(rvip_handle* h, float* m, float* v, float lr, float beta1, float beta2, float eps, long long step,
                         float grad_scale) {
  RVIP_REQUIRE(h && h->bound && h->training && m && v && step >= 1, "rvip_set_inline_adam: handle not bound for training");
  h->inline_adam.m = m; h->inline_adam.v = v;
  h->inline_adam.lr_t = (float)((double)lr * std::sqrt(1.0 - std::pow((double)beta2, (double)step)) /
                                (1.0 - std::pow((double)beta1, (double)step)));
  h->inline_adam.b1 = beta1; h->inline_adam.b2 = beta2; h->inline_adam.eps = eps; h->inline_adam.gs = grad_scale;
  h->inline_adam.armed = 1;
  return 0;
}

int rvip_adam_bucket(rvip_handle* h, int bucket, float* m, float* v, float lr, float beta1, float beta2, float eps,
                     long long step, float grad_scale, void* stream) {
  RVIP_REQUIRE(h && h->bound && h->training && m && v && step >= 1, "rvip_adam_bucket: handle not bound for training");
  RVIP_REQUIRE(bucket >= 0 && bucket < (int)h->buckets.size(), "rvip_adam_bucket: bucket %d out of range", bucket);
  cudaStream_t st = (cudaStream_t)stream;
  const float lr_t = (float)((double)lr * std::sqrt(1.0 - std::pow((double)beta2, (double)step)) /
                             (1.0 - std::pow((double)beta1, (double)step)));
  return adam_pack_bucket(h, (size_t)bucket, m, v, lr_t, beta1, beta2, eps, grad_scale, st);
}

int rvip_sgd_step(rvip_handle* h, float* velocity, float lr, float momentum, int nesterov, float grad_scale, void* stream) {
  RVIP_REQUIRE(h && h->bound && h->training, "rvip_sgd_step: handle not bound for training");
  RVIP_REQUIRE(momentum == 0.f || velocity, "rvip_sgd_step: momentum needs a velocity buffer");
  cudaStream_t st = (cudaStream_t)stream;
  h->cur_tag = "step:sgd";
  if (timed(h, KC_OPTIM, 1, st, [&] {
        return sgd_launch(h->params, h->grads, momentum == 0.f ? nullptr : velocity, (size_t)h->n_params, lr, momentum,
                          nesterov, grad_scale, st);
      }))
    return 1;
  return pack_weights(h, st);
}

int rvip_num_buckets(const rvip_handle* h) { return (int)h->buckets.size(); }
int rvip_bucket(const rvip_handle* h, int index, long long* offset, long long* count) {
  RVIP_REQUIRE(index >= 0 && index < (int)h->buckets.size(), "rvip_bucket: index out of range");
  *offset = h->buckets[index].first;
  *count = h->buckets[index].second;
  return 0;
}
int rvip_set_bucket_event(rvip_handle* h, int index, void* ev) {
  RVIP_REQUIRE(index >= 0 && index < (int)h->buckets.size(), "rvip_set_bucket_event: index out of range");
  h->bucket_events[index] = (cudaEvent_t)ev;
  return 0;
}

int rvip_heat_stats(const float* heat, const float* target, const float* inplane, long long n_pixels, int hw, int classes,
                    int loss_kind, float mask_thr, double* out, void* stream) {
  RVIP_REQUIRE(heat && target && out && n_pixels >= 0 && hw > 0, "rvip_heat_stats: bad argument");
  return heat_stats_launch(heat, target, inplane, (size_t)n_pixels, hw, classes, loss_kind, mask_thr, out,
                           (cudaStream_t)stream);
}

size_t rvip_extract_scratch_bytes(int Z, int C) { return extract_scratch_bytes(Z, C); }
int rvip_extract(const float* heat, int Z, int H, int W, int C, float thr, double* yx, int* count, int* argmax,
                 float* maxv, void* scratch, void* stream) {
  RVIP_REQUIRE(heat && yx && count && argmax && maxv && scratch, "rvip_extract: null argument");
  return extract_launch(heat, Z, H, W, C, thr, yx, count, argmax, maxv, static_cast<unsigned long long*>(scratch),
                        (cudaStream_t)stream);
}

int rvip_label_map(const float* heat, long long n_pixels, int C, float thr, uint8_t* labels, void* stream) {
  RVIP_REQUIRE(heat && labels && n_pixels >= 0 && C >= 1, "rvip_label_map: bad argument");
  return label_map_launch(heat, (size_t)n_pixels, C, thr, labels, (cudaStream_t)stream);
}

size_t rvip_cc_scratch_bytes(int Z, int H, int W) { return cc_scratch_bytes(Z, H, W); }
int rvip_cc_filter(const uint8_t* labels, int Z, int H, int W, int connectivity, uint8_t* out, void* scratch, void* stream) {
  RVIP_REQUIRE(labels && out && scratch, "rvip_cc_filter: null argument");
  return cc_filter_launch(labels, Z, H, W, connectivity, out, scratch, (cudaStream_t)stream);
}

int rvip_landmark_metrics(const double* gt_yx, const double* pred_yx, int Z, double spacing, double threshold, double dim,
                          double* angle, double* dist, double* dist_thr, double* dist_ub, double* summary, void* stream) {
  RVIP_REQUIRE(gt_yx && pred_yx && angle && dist && dist_thr && dist_ub && summary, "rvip_landmark_metrics: null argument");
  return landmark_metrics_launch(gt_yx, pred_yx, Z, spacing, threshold, dim, angle, dist, dist_thr, dist_ub, summary,
                                 (cudaStream_t)stream);
}

int rvip_debug_buffer(const rvip_handle* h, const char* name, int which, void** ptr, long long* count,
                      int* elem_bytes) {
  RVIP_REQUIRE(h && h->bound, "rvip_debug_buffer: handle not bound");
  for (size_t i = 0; i < h->L.size(); ++i) {
    const Layer& l = h->L[i];
    if (l.name != name) continue;
    const long long P = (long long)h->batch * l.H * l.W;
    long long n = 0;
    switch (which) {
      case 0: n = P * l.Cout; break;
      case 1: n = l.y ? P * l.Cout : 0; break;
      case 2: n = l.post == POST_POOL ? P / 4 * l.Cout : (l.post == POST_UPSAMPLE ? 4 * P * l.Cout : 0); break;
      case 3: n = l.dx0 ? ((l.up_dgrad || l.tr_simt) ? P / 4 : P) * l.C0 : 0; break;
      case 4: n = l.dx1 ? P * l.C1 : 0; break;
      default: set_error("rvip_debug_buffer: bad selector %d", which); return 1;
    }
    *ptr = const_cast<void*>(buffer_of(h, (int)i, which));
    *count = *ptr ? n : 0;
    *elem_bytes = (int)esize(h);
    return 0;
  }
  set_error("rvip_debug_buffer: unknown layer '%s'", name);
  return 1;
}

int rvip_dropout_mask(uint64_t seed, uint32_t site, float rate, long long n_elems, uint8_t* keep, void* stream) {
  RVIP_REQUIRE(n_elems % 8 == 0, "rvip_dropout_mask: element count must be a multiple of 8");
  return dropout_mask_launch(seed, site, (uint32_t)std::lround((double)rate * 65536.0), (size_t)(n_elems / 8), keep,
                             (cudaStream_t)stream);
}
int rvip_dropout_site(const rvip_handle* h, const char* name, uint32_t* site, float* rate) {
  for (const Layer& l : h->L)
    if (l.name == name) {
      *site = l.site;
      *rate = l.post == POST_DROPOUT ? l.drop : 0.f;
      return 0;
    }
  set_error("rvip_dropout_site: unknown layer '%s'", name);
  return 1;
}

int rvip_profile(rvip_handle* h, int enable) {
  h->profile = enable;
  return 0;
}
int rvip_profile_read(rvip_handle* h, float ms[RVIP_NUM_KERNEL_CLASSES], long long launches[RVIP_NUM_KERNEL_CLASSES]) {
  for (int i = 0; i < RVIP_NUM_KERNEL_CLASSES; ++i) {
    ms[i] = 0.f;
    launches[i] = 0;
  }
  h->detail.clear();
  for (auto& r : h->prof) {
    RVIP_CUDA(cudaEventSynchronize(r.e1));
    float t = 0.f;
    RVIP_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    char line[256];
    snprintf(line, sizeof(line), "%s,%s,%.5f\n", kClassNames[r.cls], r.tag.c_str(), t);
    h->detail += line;
    ms[r.cls] += t;
    launches[r.cls] += 1;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  h->prof.clear();
  return 0;
}
const char* rvip_profile_detail(const rvip_handle* h) { return h->detail.c_str(); }
const char* rvip_kernel_class_name(int cls) {
  return (cls >= 0 && cls < RVIP_NUM_KERNEL_CLASSES) ? kClassNames[cls] : "?";
}
long long rvip_launch_count(const rvip_handle* h) { return h->launches; }

int rvip_conv3x3_tc(const void* in0, const void* in1, int C0, int C1, const void* w_packed, const float* bias,
                    void* out0, void* out1, int out_split, double* stats, int B, int H, int W, int Cout, int mode,
                    void* stream) {
  ConvTcArgs a;
  int KC, BN;
  if (setup_conv_tc(&a, &KC, &BN, in0, in1, C0, C1, w_packed, out0, out1, out_split, B, H, W, Cout, mode, bias, stats))
    return 1;
  return conv_tc_launch(a, KC, BN, (cudaStream_t)stream);
}

int rvip_wgrad3x3_tc(const void* x0, const void* x1, int C0, int C1, const void* dz, float* dw, int B, int H, int W,
                     int Cout, void* stream) {
  WgradTcArgs a;
  int CBA, CBB;
  if (setup_wgrad_tc(&a, &CBA, &CBB, x0, x1, C0, C1, dz, dw, B, H, W, Cout)) return 1;
  return wgrad_tc_launch(a, CBA, CBB, (cudaStream_t)stream);
}

/* debug: device buffer [148][4] of long long receiving the issue-loop wait accounting of rvip_conv3x3_halo */
static long long* g_halo_dbg = nullptr;
int rvip_conv3x3_halo_debug(long long* dbg) {
  g_halo_dbg = dbg;
  return 0;
}

int rvip_conv3x3_halo(const void* in0, const void* in1, int C0, int C1, const void* w_packed, const float* bias,
                      void* out0, void* out1, int out_split, double* stats, int B, int H, int W, int Cout, int mode,
                      void* stream) {
  ConvHaloArgs a;
  int BN, nb;
  RVIP_REQUIRE(conv_halo_plan(B, H, W, C0, C1, Cout, mode, out_split, &BN, &nb),
               "rvip_conv3x3_halo: shape not eligible for the halo kernel");
  if (setup_conv_halo(&a, BN, in0, in1, C0, C1, w_packed, out0, out1, out_split, B, H, W, Cout, mode, bias, stats))
    return 1;
  a.dbg = g_halo_dbg;
  return conv_halo_launch(a, BN, nb, (cudaStream_t)stream);
}

int rvip_upconv3x3_halo(int dir, const void* low, const void* high, const float* w_hwio, const float* bias,
                        void* packed_scratch, int B, int h, int w, int Cin, int C, int transposed, void* stream) {
  const int ns = conv_halo_up_variant(h, w, Cin, C);
  RVIP_REQUIRE(ns != 0, "rvip_upconv3x3_halo: shape not eligible for the phase-decomposed kernel");
  cudaStream_t st = (cudaStream_t)stream;
  UpPackEntry e;
  e.src = 0; e.dst_f = 0; e.dst_d = conv_halo_up_pack_elems(Cin, C); e.Cin = Cin; e.C = C; e.ns = ns;
  e.transposed = transposed;
  UpPackEntry* e_dev = nullptr;
  RVIP_CUDA(cudaMalloc(&e_dev, sizeof(e)));
  RVIP_CUDA(cudaMemcpy(e_dev, &e, sizeof(e), cudaMemcpyHostToDevice));
  RVIP_CUDA(cudaMemsetAsync(packed_scratch, 0, 2 * sizeof(__nv_bfloat16) * (size_t)e.dst_d, st));
  int rc = pack_up_launch(w_hwio, packed_scratch, e_dev, 1, st);
  ConvHaloArgs a;
  int BN = 0, nb = 0;
  if (!rc && !conv_halo_up_plan(B, h, w, Cin, C, dir, &BN, &nb)) {
    set_error("rvip_upconv3x3_halo: no plan");
    rc = 1;
  }
  const __nv_bfloat16* pk = static_cast<const __nv_bfloat16*>(packed_scratch) + (dir ? e.dst_d : 0);
  if (!rc) rc = setup_conv_halo_up(&a, dir, BN, low, high, pk, bias, B, h, w, Cin, C);
  if (!rc) {
    a.dbg = g_halo_dbg;
    rc = conv_halo_launch(a, BN, nb, st);
  }
  cudaStreamSynchronize(st);
  cudaFree(e_dev);
  return rc;
}

int rvip_upconv_wgrad_halo(const void* x_low, const void* dz, float* dw, int B, int h, int w, int Cin, int C,
                           int transposed, void* stream) {
  WgradHaloArgs a;
  WgradHaloPlan p;
  RVIP_REQUIRE(wgrad_halo_up_plan(B, h, w, Cin, C, &p), "rvip_upconv_wgrad_halo: shape not eligible");
  if (setup_wgrad_halo_up(&a, p, x_low, dz, dw, B, h, w, Cin, C, transposed)) return 1;
  return wgrad_halo_launch(a, p.CIC, p.BN, (cudaStream_t)stream);
}

int rvip_wgrad3x3_halo(const void* x0, const void* x1, int C0, int C1, const void* dz, float* dw, int B, int H, int W,
                       int Cout, void* stream) {
  WgradHaloArgs a;
  WgradHaloPlan p;
  RVIP_REQUIRE(wgrad_halo_plan(B, H, W, C0, C1, Cout, &p), "rvip_wgrad3x3_halo: shape not eligible for the halo kernel");
  if (setup_wgrad_halo(&a, p, x0, x1, C0, C1, dz, dw, B, H, W, Cout)) return 1;
  return wgrad_halo_launch(a, p.CIC, p.BN, (cudaStream_t)stream);
}

int rvip_conv3x3_row(const void* in0, const void* in1, int C0, int C1, const void* w_packed, const float* bias,
                     void* out0, void* out1, int out_split, double* stats, int B, int H, int W, int Cout, int mode,
                     int base_offset_mode, void* stream) {
  ConvRowArgs a;
  int BN, R, wres, nst;
  RVIP_REQUIRE(conv_row_plan(H, W, C0, C1, Cout, mode, out_split, &BN, &R, &wres, &nst),
               "rvip_conv3x3_row: shape not eligible for the row-tiled kernel");
  if (setup_conv_row(&a, BN, R, wres, in0, in1, C0, C1, w_packed, out0, out1, out_split, B, H, W, Cout, mode, bias,
                     stats))
    return 1;
  a.base_offset_mode = base_offset_mode;
  a.dbg = g_halo_dbg;
  return conv_row_launch(a, BN, R, nst, (cudaStream_t)stream);
}

int rvip_wgrad3x3_row(const void* x0, const void* x1, int C0, int C1, const void* dz, float* dw, int B, int H, int W,
                      int Cout, void* stream) {
  WgradRowArgs a;
  int BN, R, nst;
  RVIP_REQUIRE(wgrad_row_plan(H, W, C0, C1, Cout, &BN, &R, &nst),
               "rvip_wgrad3x3_row: shape not eligible for the row-tiled kernel");
  if (setup_wgrad_row(&a, BN, R, x0, x1, C0, C1, dz, dw, B, H, W, Cout)) return 1;
  return wgrad_row_launch(a, BN, R, nst, (cudaStream_t)stream);
}

}  // extern "C"
