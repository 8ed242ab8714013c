// Largest-connected-component filter on label volumes (row N4 of the scope table): per slice and per non-zero label
// value keep only the connected component with the largest area.  Replaces clean_3d_prediction_2d_cc
// (src/data/Postprocess.py:108-120), switched on by CC_FILTER in predict_model.py:159-161 and always on in
// predict_4d_on_seg.py:99.  The reference's cv2 call passes 4 positionally (the `labels` slot), so OpenCV actually
// runs 8-connected; connectivity is a parameter here and the Python mirror defaults to 8 (oracle/cc_ref.py).
// Lock-free union-find in global memory (one thread per pixel, atomicMin linking towards the smaller raster index, so
// the root of a component is its first pixel in raster order), then size count, a packed (size, ~first pixel) 64-bit
// atomicMax per (slice, label) and a final select pass.  Integer work only: results are bit-exact; on equal maximal
// areas the component whose first pixel comes first in raster order wins.
#include "kernels.cuh"

namespace rvip {

constexpr int kCcMaxLabels = 4;

size_t cc_scratch_bytes(int Z, int H, int W) {
  const size_t n = (size_t)Z * H * W;
  return n * sizeof(int) + (n + (n & 1)) * sizeof(unsigned int) + (size_t)Z * kCcMaxLabels * sizeof(unsigned long long) + 16;
}

__device__ __forceinline__ int cc_find(int* parent, int i) {
  int p = __ldcg(parent + i);
  while (p != i) {
    const int gp = __ldcg(parent + p);
    if (gp != p) parent[i] = gp;   // path halving (benign race: parents only ever move towards the root)
    i = p;
    p = gp;
  }
  return i;
}
__device__ __forceinline__ void cc_union(int* parent, int a, int b) {
  while (true) {
    a = cc_find(parent, a);
    b = cc_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(parent + a, b);   // link the larger root under the smaller one
    if (old == a) return;
    a = old;                                     // somebody re-linked a first: retry from its new parent
  }
}

__global__ void __launch_bounds__(256) cc_init_kernel(const uint8_t* __restrict__ lab, int* __restrict__ parent, int n) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p < n) parent[p] = lab[p] ? p : -1;
}
__global__ void __launch_bounds__(256) cc_union_kernel(const uint8_t* __restrict__ lab, int* parent, int Z, int H, int W,
                                                       int conn8) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= Z * H * W) return;
  const uint8_t v = lab[p];
  if (!v) return;
  const int x = p % W, y = (p / W) % H;
  if (x > 0 && lab[p - 1] == v) cc_union(parent, p, p - 1);
  if (y > 0) {
    if (lab[p - W] == v) cc_union(parent, p, p - W);
    if (conn8) {
      if (x > 0 && lab[p - W - 1] == v) cc_union(parent, p, p - W - 1);
      if (x < W - 1 && lab[p - W + 1] == v) cc_union(parent, p, p - W + 1);
    }
  }
}
__global__ void __launch_bounds__(256) cc_count_kernel(int* parent, unsigned int* __restrict__ size, int n) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= n || parent[p] < 0) return;
  // read-only walk: a path-halving store of another thread's walk could land AFTER this thread's final
  // parent[p] = root and leave p pointing at an inner node (cc_write_kernel compares parent[p] with the root)
  int r = p;
  for (int q = __ldcg(parent + r); q != r; q = __ldcg(parent + r)) r = q;
  parent[p] = r;
  atomicAdd(size + r, 1u);
}
__global__ void __launch_bounds__(256) cc_select_kernel(const uint8_t* __restrict__ lab, const int* __restrict__ parent,
                                                        const unsigned int* __restrict__ size,
                                                        unsigned long long* __restrict__ best, int n, int HW) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= n || parent[p] != p) return;           // roots only
  const int z = p / HW, v = lab[p];
  if (v > kCcMaxLabels) return;
  const unsigned long long key = ((unsigned long long)size[p] << 32) | (0xFFFFFFFFu - (unsigned int)(p - z * HW));
  atomicMax(best + (size_t)z * kCcMaxLabels + (v - 1), key);
}
__global__ void __launch_bounds__(256) cc_write_kernel(const uint8_t* __restrict__ lab, const int* __restrict__ parent,
                                                       const unsigned long long* __restrict__ best,
                                                       uint8_t* __restrict__ out, int n, int HW) {
  const int p = blockIdx.x * 256 + threadIdx.x;
  if (p >= n) return;
  const uint8_t v = lab[p];
  uint8_t o = 0;
  if (v && v <= kCcMaxLabels) {
    const int z = p / HW;
    const unsigned long long key = best[(size_t)z * kCcMaxLabels + (v - 1)];
    const int root = z * HW + (int)(0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull));
    if (parent[p] == root) o = v;
  }
  out[p] = o;
}

int cc_filter_launch(const uint8_t* labels, int Z, int H, int W, int connectivity, uint8_t* out, void* scratch,
                     cudaStream_t st) {
  RVIP_REQUIRE(connectivity == 4 || connectivity == 8, "cc_filter: connectivity %d not in {4, 8}", connectivity);
  RVIP_REQUIRE(Z >= 0 && H > 0 && W > 0 && (long long)Z * H * W < 0x7fffffffLL, "cc_filter: bad shape");
  if (Z == 0) return 0;
  const int n = Z * H * W, grid = (n + 255) / 256;
  int* parent = static_cast<int*>(scratch);
  unsigned int* size = reinterpret_cast<unsigned int*>(parent + n);
  unsigned long long* best = reinterpret_cast<unsigned long long*>(size + n + (n & 1));
  RVIP_CUDA(cudaMemsetAsync(size, 0, ((size_t)n + (n & 1)) * sizeof(unsigned int) +
                                         (size_t)Z * kCcMaxLabels * sizeof(unsigned long long), st));
  cc_init_kernel<<<grid, 256, 0, st>>>(labels, parent, n);
  cc_union_kernel<<<grid, 256, 0, st>>>(labels, parent, Z, H, W, connectivity == 8);
  cc_count_kernel<<<grid, 256, 0, st>>>(parent, size, n);
  cc_select_kernel<<<grid, 256, 0, st>>>(labels, parent, size, best, n, H * W);
  cc_write_kernel<<<grid, 256, 0, st>>>(labels, parent, best, out, n, H * W);
  RVIP_LAUNCH_CHECK();
  return 0;
}

}  // namespace rvip
