// Halo-staged tcgen05 convolution for the deep levels (conv_halo.cu): argument block, planner, launcher.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_tc.cuh"

namespace rvip {

struct ConvHaloArgs {
  CUtensorMap in0, in1;    // NHWC bf16, box {64, 18, 18, 1} (16 x 16 pixel block + halo), SWIZZLE_128B
  CUtensorMap w;           // packed weights [Cout][9*Ctot] bf16 (K-major), box {64, BN}
  CUtensorMap out0, out1;  // NHWC bf16, box {64, 8, 16, 1}: one 64-channel slice of one half block
  int B, H, W;
  int C0, Ctot, Cout;
  int n_ntiles, tiles_x, tiles_y, total_tiles;
  int mode, out_split;     // ConvEpilogue; EPI_LINEAR: output channels >= out_split go to out1
  const float* bias;
  double* stats;
  const float* scale;      // [Cout] (EPI_RELU_AFFINE)
  const float* shift;
  long long* dbg;          // optional [grid][8]: issue-loop cycles, waits on TMEM / activation block / weight tile, kernel cycles, epilogue cycles
};
// false if the layer does not fit (H, W not multiples of 16; channels not multiples of 64)
bool conv_halo_plan(int B, int H, int W, int C0, int C1, int Cout, int mode, int out_split, int* BN, int* nbst);
int conv_halo_launch(const ConvHaloArgs& a, int BN, int nbst, cudaStream_t st);

}  // namespace rvip
