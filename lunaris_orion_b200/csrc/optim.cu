// Optimizer boundary of the step (train_hybrid.py:906-922): clip_grad_norm_ + AdamW as two multi-tensor kernels.
//   kernel 1: sum of squares of every gradient tensor of one model -> norm2[]             (one pass over the grads)
//   kernel 2: coef = min(1, max_norm / (sqrt(norm2) + 1e-6)); g *= coef (the reference's in-place clip);
//             decoupled weight decay + Adam moments + parameter update (torch.optim.AdamW semantics, fp32)
// Both read the gradient as g * grad_scale: data-parallel ranks all-reduce SUMS and fold the 1/world average in here.
// The tensors are described by a device table of {param, grad, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps,
// weight_decay, 1 - beta1^t, sqrt(1 - beta2^t)}: hyper-parameters and the step count are PER TENSOR, so several param
// groups and parameters whose first gradient arrives late update exactly like torch.optim.AdamW. Work is cut into
// fixed chunks so one launch covers all ~100 tensors of a model. State layout stays torch's (checkpoint contract).
#include "../../include/lunaris_b200.h"
#include "elem_common.cuh"
#include "launch_count.cuh"

namespace lun {

struct OptTensor {
  float* p;
  float* g;
  float* m;
  float* v;
  long long n;
  float lr, beta1, beta2, eps, wd, bias_c1, bias_c2_sqrt, pad;
};
static_assert(sizeof(OptTensor) == 72, "host table rows are 72 bytes (lunaris_orion_b200/optim.py)");
constexpr int kOptChunk = 8192;   // elements per block-iteration

__global__ void __launch_bounds__(256) multi_sumsq_kernel(const OptTensor* __restrict__ tab, const int2* __restrict__ chunks,
                                                          int nchunks, float grad_scale,
                                                          float* __restrict__ norm2) {
  float acc = 0.f;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const OptTensor t = tab[chunks[c].x];
    const long long base = (long long)chunks[c].y * kOptChunk;
    const long long end = base + kOptChunk < t.n ? base + kOptChunk : t.n;
    for (long long i = base + threadIdx.x; i < end; i += 256) {
      const float g = t.g[i] * grad_scale;
      acc += g * g;
    }
  }
  acc = warp_sum(acc);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += s[i];
    norm2[blockIdx.x] = tot;     // per-block partial: the final sum is taken in a fixed order (bitwise reproducible,
  }                              // so data-parallel ranks that hold identical gradients take identical steps)
}

__global__ void __launch_bounds__(256) multi_adamw_kernel(const OptTensor* __restrict__ tab, const int2* __restrict__ chunks,
                                                          int nchunks, const float* __restrict__ norm2, int npart,
                                                          float max_norm, float grad_scale) {
  __shared__ float s_norm2;
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < npart; ++i) tot += norm2[i];
    s_norm2 = tot;
  }
  __syncthreads();
  const float norm = sqrtf(s_norm2);
  const float coef = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const OptTensor t = tab[chunks[c].x];
    const float step_size = t.lr / t.bias_c1;
    const float decay = 1.f - t.lr * t.wd;
    const float beta1 = t.beta1, beta2 = t.beta2, eps = t.eps, bias_c2_sqrt = t.bias_c2_sqrt;
    const long long base = (long long)chunks[c].y * kOptChunk;
    const long long end = base + kOptChunk < t.n ? base + kOptChunk : t.n;
    for (long long i = base + threadIdx.x; i < end; i += 256) {
      const float g = t.g[i] * grad_scale * coef;
      t.g[i] = g;
      const float m = beta1 * t.m[i] + (1.f - beta1) * g;
      const float v = beta2 * t.v[i] + (1.f - beta2) * g * g;
      t.m[i] = m;
      t.v[i] = v;
      const float denom = sqrtf(v) / bias_c2_sqrt + eps;
      t.p[i] = t.p[i] * decay - step_size * (m / denom);
    }
  }
}

// Kernel-operand shadow of an fp32 parameter after an optimizer step: src [A][B][T] fp32 (a conv weight [Cout][Cin][kh*kw]
// or a transposed-conv weight [Cin][Cout][kh*kw]) -> dst bf16 [T][A][B] (transpose == 0) or [T][B][A] (transpose == 1),
// the K-major slabs the implicit-GEMM kernels read. One launch per weight instead of torch's permute copy + cast.
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int A,
                                                          int B, int T, int transpose) {
  const long n = (long)A * B * T;
  const int X = transpose ? B : A, Y = transpose ? A : B;       // dst is [T][X][Y]
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) {
    const int y = (int)(i % Y);
    const int x = (int)((i / Y) % X);
    const int t = (int)(i / ((long)X * Y));
    const int a = transpose ? y : x, b = transpose ? x : y;
    dst[i] = __float2bfloat16_rn(__ldg(src + ((long)a * B + b) * T + t));
  }
}

}  // namespace lun

using namespace lun;

extern "C" {

int lun_multi_grad_sumsq(const void* table, const void* chunks, int nchunks, float grad_scale, float* norm2,
                         void* stream) {
  if (nchunks <= 0) return LUN_OK;
  int grid = nchunks < 1024 ? nchunks : 1024;      // == number of partial sums written to norm2[]
  multi_sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const OptTensor*)table, (const int2*)chunks, nchunks,
                                                            grad_scale, norm2);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_multi_clip_adamw(const void* table, const void* chunks, int nchunks, const float* norm2, float max_norm,
                         float grad_scale, void* stream) {
  if (nchunks <= 0) return LUN_OK;
  int grid = nchunks < 148 * 8 ? nchunks : 148 * 8;
  const int npart = nchunks < 1024 ? nchunks : 1024;
  multi_adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const OptTensor*)table, (const int2*)chunks, nchunks,
                                                            norm2, npart, max_norm, grad_scale);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_pack_weight_bf16(const float* src, void* dst, int A, int B, int T, int transpose, void* stream) {
  if (A < 1 || B < 1 || T < 1) return LUN_E_SHAPE;
  const long n = (long)A * B * T;
  long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  lun::pack_weight_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, (lun::bf16*)dst, A, B, T, transpose);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

}  // extern "C"
