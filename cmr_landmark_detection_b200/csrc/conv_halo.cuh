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
  float floor;             // activation floor of the EPI_RELU* modes: 0 = ReLU, -inf = none (BN_FIRST)
  const float* bias;
  double* stats;
  const float* scale;      // [Cout] (EPI_RELU_AFFINE)
  const float* shift;
  BnRedArgs bnred;         // EPI_LINEAR_BNRED
  long long* dbg;          // optional [grid][8]: issue-loop cycles, waits on TMEM / activation block / weight tile, kernel cycles, epilogue cycles
  // ---- phase-decomposed up-convolution (UpSampling2D(2) -> Conv3x3 folded into 2x2 / 2x3-tap convolutions on the
  // LOW-resolution tensor, one per output phase (row parity a, column parity b); conv_halo_up_* below).  H, W above are
  // then the low-resolution geometry and Cout the channels of ONE phase view.
  int up_ns;               // 0 = plain 3x3; 2 = four phases (a, b) of 2x2 taps; 3 = two phases (a) of 2x3 taps, b merged into channels
  int up_dir;              // 0 = forward (phases enumerate output tiles, stores through upout[ph]);
                           // 1 = dgrad (phases enumerate K groups, loads through upin[ph])
  int up_nph, up_cz;       // phases; channels of one phase view on the dgrad input side
  int bias_mod;            // bias index = channel % bias_mod (merged phases replicate the bias)
  CUtensorMap upin[4];     // dgrad: strided phase views of the high-resolution dz, box {64, 18, 18, 1}
  CUtensorMap upout[4];    // forward: strided phase views of the high-resolution output, box {64, 8, 16, 1}
};
// false if the layer does not fit (channels not multiples of 64, or no shared memory for a weight ring)
bool conv_halo_plan(int B, int H, int W, int C0, int C1, int Cout, int mode, int out_split, int* BN, int* nbst);
int conv_halo_launch(const ConvHaloArgs& a, int BN, int nbst, cudaStream_t st);

// Phase-decomposed up-convolution: u = relu(conv3x3(upsample2x(x)) + b) computed from the low-resolution x.
// Output pixel (2i + a, 2j + b) only sees the 2 x 2 low-resolution neighbourhood rows {i + a - 1, i + a}, columns
// {j + b - 1, j + b}; the 3x3 taps that land on the same low-resolution pixel are pre-summed (fp32) when the weights are
// packed, so a phase costs 4 taps instead of 9 (2.25x fewer MMAs) and the 4x larger up-sampled tensor is never read.
// Cout % 64 == 0: four phases, N = Cout; Cout == 32: two row phases with the column phase merged into N = 64
// (2 x 3 taps, structurally zero weights where a column phase does not see a tap).
// Returns the variant (2 / 3) or 0 when the layer is not eligible (h, w = LOW-resolution size).
int conv_halo_up_variant(int h, int w, int Cin, int Cout);
bool conv_halo_up_plan(int B, int h, int w, int Cin, int Cout, int dir, int* BN, int* nbst, bool bnred = false);
// elements of one packed operand copy (forward and dgrad copies have the same size)
long long conv_halo_up_pack_elems(int Cin, int Cout);

}  // namespace rvip
