// Helpers shared by the bandwidth-bound (CUDA-core) kernels: 16-byte bf16 vectors, counter-based dropout RNG,
// per-channel block reductions.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lun {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
// 16-byte packed bf16 vector <-> 8 floats (kept packed while loads are in flight to save registers)
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 ldg16(const bf16* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Counter-based, stateless dropout RNG: a 32-bit integer hash (lowbias32) of (seed, element index); one hash yields
// two 16-bit uniforms. The same function regenerates the mask in the backward pass.
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t drop_key(uint64_t seed, uint64_t idx8) {
  return static_cast<uint32_t>(seed) ^ (static_cast<uint32_t>(idx8) * 4u) ^
         (static_cast<uint32_t>(idx8 >> 30) * 0x9E3779B9U) ^ static_cast<uint32_t>(seed >> 32) * 0x85EBCA6BU;
}
// keep[i] for elements 8*idx8 .. 8*idx8+7 of a logical tensor; element is dropped when its 16-bit uniform < thresh16.
__device__ __forceinline__ void drop_keep8(uint64_t seed, uint64_t idx8, uint32_t thresh16, bool (&keep)[8]) {
  const uint32_t k = drop_key(seed, idx8);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t h = hash32(k + i);
    keep[2 * i] = (h & 0xFFFFu) >= thresh16;
    keep[2 * i + 1] = (h >> 16) >= thresh16;
  }
}
__device__ __forceinline__ bool drop_keep1(uint64_t seed, uint64_t idx, uint32_t thresh16) {
  const uint32_t h = hash32(drop_key(seed, idx >> 3) + ((idx >> 1) & 3));
  return ((idx & 1) ? (h >> 16) : (h & 0xFFFFu)) >= thresh16;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace lun
