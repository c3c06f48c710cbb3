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
__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// splitmix64: counter-based, stateless. One call yields four 16-bit uniforms.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// keep[i] for elements 8*idx8 .. 8*idx8+7 of a logical tensor; element is dropped when its 16-bit uniform < thresh16.
__device__ __forceinline__ void drop_keep8(uint64_t seed, uint64_t idx8, uint32_t thresh16, bool (&keep)[8]) {
  const uint64_t a = splitmix64(seed ^ (idx8 * 2));
  const uint64_t b = splitmix64(seed ^ (idx8 * 2 + 1));
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    keep[i] = ((a >> (16 * i)) & 0xFFFF) >= thresh16;
    keep[4 + i] = ((b >> (16 * i)) & 0xFFFF) >= thresh16;
  }
}
__device__ __forceinline__ bool drop_keep1(uint64_t seed, uint64_t idx, uint32_t thresh16) {
  const uint64_t a = splitmix64(seed ^ ((idx >> 3) * 2 + ((idx >> 2) & 1)));
  return ((a >> (16 * (idx & 3))) & 0xFFFF) >= thresh16;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace lun
