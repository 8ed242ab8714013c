// Host-visible argument blocks + launchers of the tcgen05 convolution kernels (conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvip {

// pixel-tile geometry: a GEMM row block of P pixels = NB images x TH rows x TW columns
struct TileGeom {
  int TW, TH, NB;
  int tiles_x, tiles_y, tiles_b;
  int full;  // 1 when every tile lies completely inside the tensor (no row masking needed)
};
TileGeom make_tile_geom(int B, int H, int W, int P);

enum ConvEpilogue {
  EPI_RELU_STATS = 0,  // a = relu(acc + bias), bf16 store, per-channel sum / sum-of-squares (training, BN follows)
  EPI_RELU = 1,        // a = relu(acc + bias), bf16 store (up-conv; inference)
  EPI_LINEAR = 2,      // acc, bf16 store (dgrad); output channels >= out_split go to out1
  EPI_RELU_AFFINE = 3, // y = scale * relu(acc + bias) + shift: inference, BatchNorm (moving statistics) folded in
  EPI_LINEAR_BNRED = 4,  // dgrad whose output IS dL/dy of a BatchNorm block (row / halo kernels, single output): stores it like
                         // EPI_LINEAR and also leaves that block's BatchNorm-backward sums (BnRedArgs) -- the separate
                         // two-tensor statistics pass (bn_bwd_reduce) of that block disappears
};

// EPI_LINEAR_BNRED: the epilogue reads the block's stored relu(conv) tile `a` (plain coalesced loads issued before the
// accumulator is waited for), replays the dropout keep-mask that sits between the block and this convolution, and adds
//   red[stripe][c] += keep_scale * sum keep * dy,   red[stripe][C + c] += keep_scale * sum keep * dy * a
// (the sums bn_bwd_reduce_kernel produces; striped like there).
constexpr int kBnRedStripes = 4;
struct BnRedArgs {
  const void* a;          // [B, H, W, C] bf16: relu(conv) of the block whose output gradient this convolution produces
  double* red;            // [kBnRedStripes][2][C]
  uint32_t k0, k1;        // dropout key of the step (common.cuh: dropout_key(seed, site))
  uint32_t thr16;         // 0 = no dropout between the block and this convolution
  float keep_scale;       // 1 / (1 - rate)
  int lg;                 // log2(C / 8): vector index of the mask generator = (pixel << lg) | (channel / 8)
};

// ---- forward / dgrad: out[p, n] = epi( sum_{tap, c} in[p + off(tap), c] * Wp[n][tap][c] )
struct ConvTcArgs {
  CUtensorMap in0, in1;  // NHWC bf16 activation(s), box {KC, TW, TH, NB}; in1 = second concat source
  CUtensorMap w;         // packed weights [Cout][9 * Ctot] bf16 (K-major), box {KC, BN}
  CUtensorMap out0, out1;  // NHWC bf16 outputs, box {min(BN, 64), TW, TH, NB}
  TileGeom g;
  int B, H, W;
  int C0, Ctot, Cout;
  int n_ntiles, total_tiles;
  int mode, out_split;
  float floor;        // activation floor of the EPI_RELU* modes: 0 = ReLU, -inf = none (BN_FIRST: Conv -> BN -> ReLU)
  const float* bias;  // [Cout] (EPI_RELU*)
  double* stats;      // [2][Cout] (EPI_RELU_STATS)
  const float* scale; // [Cout] (EPI_RELU_AFFINE)
  const float* shift;
};
int conv_tc_launch(const ConvTcArgs& a, int KC, int BN, cudaStream_t st);
size_t conv_tc_smem_bytes(int KC, int BN, int Cout);

// ---- wgrad: dW[tap][c][n] += sum_p x[p + off(tap), c] * dz[p, n]   (fp32 atomics into HWIO layout)
struct WgradTcArgs {
  CUtensorMap x0, x1;  // conv input(s) NHWC bf16, box {CBA, TW, TH, NB} over 64-pixel tiles
  CUtensorMap dz;      // NHWC bf16, box {CBB, TW, TH, NB}
  TileGeom g;          // geometry of the 64-pixel K tiles
  int B, H, W;
  int C0, Ctot, Cout;
  int BN, MT;          // N tile (<= 256) and M tiles (of 128 rows) per CTA; MT * BN <= 512
  int n_mtiles, n_mgroups, n_ntiles, n_split, k_tiles;
  float* dw;           // [9][Ctot][Cout] fp32, accumulated with red.global.add
};
int wgrad_tc_launch(const WgradTcArgs& a, int CBA, int CBB, cudaStream_t st);

}  // namespace rvip
