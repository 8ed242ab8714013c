// Halo-staged tcgen05 weight gradient (wgrad_halo.cu): host-visible argument block, planner and launcher.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvip {

struct WgradHaloPlan {
  int CIC, BN;               // input channels per CTA (32 | 64), output-channel tile (32 | 64 | 128)
  int TW, TH;                // pixel tile: TW (16 | 32) columns x TH rows, at most 256 pixels
  int tiles_x, tiles_y, pixel_tiles;
  int n_cchunks, n_ntiles, k_split;
};

struct WgradHaloArgs {
  CUtensorMap x0, x1;        // conv input(s) NHWC bf16, box {CIC, TW+2, TH+2, 1} (halo), swizzle = CIC*2 bytes
  CUtensorMap dz;            // NHWC bf16, box {min(BN,64), TW, TH, 1}
  float* dw;                 // [9][Ctot][Cout] fp32, accumulated with red.global.add
  int B, H, W;
  int C0, Ctot, Cout;
  int TW, TH, tiles_x, tiles_y, pixel_tiles;
  int n_cchunks, n_ntiles, k_split;
  // phase-decomposed up-convolution (conv_halo.cuh): x0 is the LOW-resolution input (H, W = its size), dzp[2a + b] the
  // phase view (pixels (2i + a, 2j + b)) of the high-resolution dz, box {min(BN,64), TW, TH, 1}
  int up;
  int transposed;            // up: dw is a Conv2DTranspose kernel gradient (kh, kw, Cout, Cin); each tap has ONE accumulator
  CUtensorMap dzp[4];
};
// weight gradient of the phase-decomposed up-convolution: eight accumulators (phase (a, b) x low-resolution row
// neighbour r, each a pair of column neighbours s = 0, 1 packed in M), folded onto the nine 3x3 taps in the flush
bool wgrad_halo_up_plan(int B, int h, int w, int Cin, int Cout, WgradHaloPlan* p);
// false if the layer does not fit this kernel (channels not multiples of 32)
bool wgrad_halo_plan(int B, int H, int W, int C0, int C1, int Cout, WgradHaloPlan* p);
int wgrad_halo_launch(const WgradHaloArgs& a, int CIC, int BN, cudaStream_t st);

}  // namespace rvip
