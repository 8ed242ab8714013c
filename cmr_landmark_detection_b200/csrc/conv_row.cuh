// Row-tiled tcgen05 convolution kernels for the wide levels (conv_row.cu: rows of 128-pixel tiles, the last one may be
// partial; wgrad_row.cu: W % 128 == 0).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_tc.cuh"

namespace rvip {

struct ConvRowArgs {
  CUtensorMap in0, in1;    // NHWC bf16, box {32, 130, R+2, 1}, SWIZZLE_64B
  CUtensorMap w;           // packed weights [Cout][9*Ctot], box {32, BN}, SWIZZLE_64B
  CUtensorMap out0, out1;  // NHWC bf16, box {min(BN,64), 128, R, 1}
  int B, H, W;
  int C0, Ctot, Cout;
  int n_ntiles, tiles_x, tiles_y, total_tiles;
  int mode, out_split, wres;
  float floor;             // activation floor of the EPI_RELU* modes: 0 = ReLU, -inf = none (BN_FIRST)
  int och;                 // channels per TMA store box: 64, or 32 when a dgrad's concat split falls on a 32-channel boundary
  int base_offset_mode;    // debug knob: 1 = encode (addr>>7)&3 in the descriptor base-offset field
  const float* bias;
  double* stats;
  const float* scale;      // [Cout] (EPI_RELU_AFFINE)
  const float* shift;
  BnRedArgs bnred;         // EPI_LINEAR_BNRED
  long long* dbg;          // optional [grid][8]: issue-loop cycles, waits on TMEM / input stage, -, kernel, epilogue cycles
};
// Chooses the tile (BN, R), weight residency and stage count; false if the layer does not fit this kernel.
bool conv_row_plan(int H, int W, int C0, int C1, int Cout, int mode, int out_split, int* BN, int* R, int* wres,
                   int* nst);
int conv_row_launch(const ConvRowArgs& a, int BN, int R, int nst, cudaStream_t st);

struct WgradRowArgs {
  CUtensorMap x0, x1;      // conv input(s) NHWC bf16, box {32, 130, R+2, 1}, SWIZZLE_64B
  CUtensorMap dz;          // NHWC bf16, box {BN, 128, R, 1} (SWIZZLE_64B for BN=32, 128B for BN=64)
  int B, H, W;
  int C0, Ctot, Cout;
  int n_ntiles, tiles_x, tiles_y, pixel_tiles;
  float* dw;               // [9][Ctot][Cout] fp32 (atomically accumulated)
};
bool wgrad_row_plan(int H, int W, int C0, int C1, int Cout, int* BN, int* R, int* nst);
int wgrad_row_launch(const WgradRowArgs& a, int BN, int R, int nst, cudaStream_t st);

}  // namespace rvip
