// Shared device/host helpers for the RVIP hot path (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace rvip {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);
#define RVIP_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      rvip::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)
#define RVIP_LAUNCH_CHECK() RVIP_CUDA(cudaGetLastError())
#define RVIP_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      rvip::set_error(__VA_ARGS__);    \
      return 1;                        \
    }                                  \
  } while (0)

constexpr int kNumSMs = 148;

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// A training step is ~165 stream-ordered launches, many of them a few microseconds long at the deep levels.
// Every hot-path kernel can be launched with programmatic stream serialization: its CTAs may be scheduled while the
// previous kernel drains, run their prologue (barrier init, TMEM allocation, tensor-map prefetch) and then block
// in pdl_wait() until the previous grid has completed and its memory is visible.  pdl_launch_dependents() at the
// top of a kernel allows the NEXT kernel to start that early.  Both are no-ops for a normal launch.
// On by default; RVIP_NO_PDL=1 falls back to plain stream serialisation (see rvip_abi.cu:pdl_enabled).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------- storage-type traits
// Activations are NHWC in T = float ("fp32 mode") or __nv_bfloat16 ("bf16 mode").
// All element-wise kernels move 8 channels per thread (16 B for bf16, 32 B for fp32).
struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};
struct alignas(16) f32x8 {
  float4 lo, hi;
};

template <typename T>
struct Vec8;
template <>
struct Vec8<float> {
  __device__ static __forceinline__ void load(const float* p, float (&o)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
    o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
  }
  __device__ static __forceinline__ void store(float* p, const float (&o)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
};
template <>
struct Vec8<__nv_bfloat16> {
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      o[2 * i] = f.x;
      o[2 * i + 1] = f.y;
    }
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&o)[8]) {
    uint4 raw;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
  }
};

// Raw (still packed) 8-channel vectors: a load can be issued iterations ahead and unpacked only when consumed.
template <typename T>
struct Raw8;
template <>
struct Raw8<float> {
  float4 lo, hi;
};
template <>
struct Raw8<__nv_bfloat16> {
  uint4 r;
};
__device__ __forceinline__ void load_raw8(const float* p, Raw8<float>& o) {
  o.lo = *reinterpret_cast<const float4*>(p);
  o.hi = *reinterpret_cast<const float4*>(p + 4);
}
__device__ __forceinline__ void load_raw8(const __nv_bfloat16* p, Raw8<__nv_bfloat16>& o) {
  o.r = *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void unpack_raw8(const Raw8<float>& r, float (&o)[8]) {
  o[0] = r.lo.x; o[1] = r.lo.y; o[2] = r.lo.z; o[3] = r.lo.w;
  o[4] = r.hi.x; o[5] = r.hi.y; o[6] = r.hi.z; o[7] = r.hi.w;
}
__device__ __forceinline__ void unpack_raw8(const Raw8<__nv_bfloat16>& r, float (&o)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    o[2 * i] = f.x;
    o[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void zero_raw8(Raw8<float>& o) { o.lo = o.hi = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void zero_raw8(Raw8<__nv_bfloat16>& o) { o.r = make_uint4(0u, 0u, 0u, 0u); }

template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Rounds through the storage type: statistics and masks are computed on the values that
// are actually stored, so forward and backward see identical numbers.
template <typename T>
__device__ __forceinline__ float round_to(float v) { return to_f32<T>(from_f32<T>(v)); }

// ---------------------------------------------------------------- Philox4x32-10 (dropout)
struct Philox {
  __device__ static __forceinline__ uint4 gen(uint64_t seed, uint64_t ctr, uint32_t stream) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0x5eed5eedu;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};
// Counter-based dropout mask: 16 random bits per element, keep iff bits >= thr16 (thr16 = round(rate*65536)).
// The keep decision for the 8 channels of 16-byte vector `vec_idx` at dropout site `site` is a pure function
// of (seed, site, vec_idx), so backward replays the forward mask without storing it.  The generator is the
// murmur3 32-bit finaliser (full avalanche, bijective) over a keyed counter, once per 8-channel vector.
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return h;
}
struct DropKey {
  uint32_t k0, k1;
};
__device__ __forceinline__ DropKey dropout_key(uint64_t seed, uint32_t site) {
  DropKey k;
  k.k0 = mix32((uint32_t)seed ^ (site * 0x9E3779B9u) ^ 0xa511e9b3u);
  k.k1 = mix32((uint32_t)(seed >> 32) + site * 0x85ebca6bu + 0x6a09e667u);
  return k;
}
__device__ __forceinline__ void dropout_keep8(const DropKey& key, uint32_t vec_idx, uint32_t thr16, bool (&keep)[8]) {
  // ONE full-avalanche hash of the keyed vector index, then four cheap odd-constant multiplies of it (bijections of h)
  // supply the 8 x 16 random bits.  ncu (profiles/r2h_*): with a full mix32 per 32 bits the dropout variants of the
  // BatchNorm passes executed 2.0 - 2.3x the instructions of the plain ones and sat on the integer pipe
  // (math-pipe throttle 2.0 warps per issue, 58 % issue slots) instead of on HBM.
  const uint32_t h = mix32(vec_idx ^ key.k0) ^ key.k1;
  const uint32_t thr_hi = thr16 << 16;
  const uint32_t w0 = h * 0x9E3779B1u, w1 = h * 0x85EBCA6Bu, w2 = h * 0xC2B2AE35u, w3 = h * 0x27D4EB2Fu;
  keep[0] = (w0 << 16) >= thr_hi; keep[1] = w0 >= thr_hi;
  keep[2] = (w1 << 16) >= thr_hi; keep[3] = w1 >= thr_hi;
  keep[4] = (w2 << 16) >= thr_hi; keep[5] = w2 >= thr_hi;
  keep[6] = (w3 << 16) >= thr_hi; keep[7] = w3 >= thr_hi;
}
__device__ __forceinline__ void dropout_keep8(uint64_t seed, uint32_t site, uint64_t vec_idx, uint32_t thr16,
                                              bool (&keep)[8]) {
  dropout_keep8(dropout_key(seed, site), (uint32_t)vec_idx, thr16, keep);
}

// Streaming passes keep few bytes in flight per thread (registers go to per-channel constants), so one thread
// per block asks the L2 to fetch the block's data a few iterations ahead (bulk prefetch, no smem, no
// registers); the vector loads that follow then pay L2 instead of HBM latency.
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// prefetch `count` items of `item_bytes` starting at item `first` of a tensor holding `n_items`
__device__ __forceinline__ void l2_prefetch_items(const void* base, uint32_t first, uint32_t count, uint32_t n_items,
                                                  uint32_t item_bytes) {
  if (first >= n_items) return;
  if (first + count > n_items) count = n_items - first;
  l2_prefetch(static_cast<const char*>(base) + (size_t)first * item_bytes, count * item_bytes);
}
constexpr uint32_t kPrefetchAhead = 4;   // grid-stride iterations

// Per-channel block reduction used by the streaming passes: thread t owns channels [c, c+8) with
// c = (t % G) * 8, so lanes l, l+G, l+2G, ... of a warp hold the same channels.  Fold those with shuffles and
// let one lane per channel group touch shared memory (float atomicAdd on smem is a CAS loop).
__device__ __forceinline__ void block_accumulate8(float* smem_acc, int c, const float (&v)[8], uint32_t G) {
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = v[j];
  for (uint32_t o = 16; o >= G && o > 0; o >>= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] += __shfl_xor_sync(0xffffffffu, r[j], o);
  }
  if ((threadIdx.x & 31) < G || G >= 32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&smem_acc[c + j], r[j]);
  }
}

// Packed fp32 pairs (Blackwell FFMA2 / FADD2): two IEEE fp32 operations per issued instruction, each lane rounded exactly
// like the scalar instruction; a pair built from one value twice compiles to the broadcast operand form (no packing move).
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t r, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace rvip
