// PixelArtFeatureExtractor front end (lunar_evaluator.py:71-96, 105-110): HBM-bound CUDA-core kernels.
//   fe_conv1_kernel : conv3x3 3->32 + bias + LeakyReLU(0.2) on NCHW fp32 images -> NHWC bf16, + BN batch statistics
//   fe_branch_kernel: BN(conv1) applied on load, three depthwise convs (3x3, 5x5, 3x3) each followed by a pointwise
//                     32->64 conv + bias + LeakyReLU, written into the 192-channel concat buffer (torch.cat of :110)
// Arithmetic mirrors the reference's bf16 autocast: conv operands rounded to bf16, fp32 accumulate, bf16 results.
#include "../../include/lunaris_b200.h"
#include "elem_common.cuh"
#include "launch_count.cuh"

namespace lun {

// ------------------------------------------------------------------------------------------- conv1 3->32
__global__ void __launch_bounds__(256) fe_conv1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, bf16* __restrict__ y,
                                                       float* __restrict__ stats, int B, int H, int W, float slope) {
  __shared__ float sw[27][32];
  __shared__ float sb[32];
  __shared__ float sstat[64];
  for (int i = threadIdx.x; i < 27 * 32; i += 256) {
    const int o = i % 32, k = i / 32;          // k = (c*3 + kh)*3 + kw ; reference layout [o][c][kh][kw]
    sw[k][o] = rbf(w[o * 27 + k]);
  }
  if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
  if (threadIdx.x < 64) sstat[threadIdx.x] = 0.f;
  __syncthreads();
  const long HW = (long)H * W;
  const long total = (long)B * HW;
  const long p = (long)blockIdx.x * 256 + threadIdx.x;
  const bool valid = p < total;
  float acc[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) acc[o] = 0.f;
  if (valid) {
    const int b = (int)(p / HW), hw = (int)(p % HW), h = hw / W, wq = hw % W;
    for (int c = 0; c < 3; ++c) {
      const float* xc = x + ((long)b * 3 + c) * HW;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = h + kh - 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = wq + kw - 1;
          float v = 0.f;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = rbf(__ldg(xc + (long)ih * W + iw));
          const float* wr = sw[(c * 3 + kh) * 3 + kw];
#pragma unroll
          for (int o = 0; o < 32; ++o) acc[o] += v * wr[o];
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 32; ++o) {
      const float t = rbf(acc[o] + sb[o]);    // conv output is a bf16 tensor; LeakyReLU runs on it
      acc[o] = rbf(t > 0.f ? t : t * slope);
    }
    bf16* dst = y + p * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v8[i] = acc[8 * j + i];
      store8(dst + 8 * j, v8);
    }
  } else {
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = 0.f;
  }
  // per-channel sum / sumsq: transpose-reduce over the warp, then shared and global atomics
  float s1[32], s2[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    s1[o] = acc[o];
    s2[o] = acc[o] * acc[o];
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = lane & off;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float a1 = upper ? s1[i] : s1[i + off], k1 = upper ? s1[i + off] : s1[i];
      s1[i] = k1 + __shfl_xor_sync(0xffffffffu, a1, off);
      const float a2 = upper ? s2[i] : s2[i + off], k2 = upper ? s2[i + off] : s2[i];
      s2[i] = k2 + __shfl_xor_sync(0xffffffffu, a2, off);
    }
  }
  atomicAdd(&sstat[lane], s1[0]);
  atomicAdd(&sstat[32 + lane], s2[0]);
  __syncthreads();
  if (threadIdx.x < 64) atomicAdd(stats + threadIdx.x, sstat[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------- branches
struct FeBranchW {
  const float* dw_w[3];   // [32][1][k][k]  (edge k=3, color k=5, detail k=3)
  const float* dw_b[3];   // [32]
  const float* pw_w[3];   // [64][32]
  const float* pw_b[3];   // [64]
};

constexpr int kT = 16;          // 16x16 pixel tile
constexpr int kHalo = 2;
constexpr int kTP = kT + 2 * kHalo;   // 20
constexpr int kPlane = kTP * kTP + 1; // +1 padding

__global__ void __launch_bounds__(256) fe_branch_kernel(const bf16* __restrict__ y0, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, FeBranchW wt,
                                                        bf16* __restrict__ cat, int H, int W, float slope) {
  extern __shared__ float sm[];
  float* tile = sm;                           // [32][kPlane]
  float* s_dw = tile + 32 * kPlane;           // [3][32][25]
  float* s_dwb = s_dw + 3 * 32 * 25;          // [3][32]
  float* s_pw = s_dwb + 96;                   // [3][64][32]
  float* s_pwb = s_pw + 3 * 64 * 32;          // [3][64]
  const int b = blockIdx.z, h0 = blockIdx.y * kT, w0 = blockIdx.x * kT;
  const int ksz[3] = {3, 5, 3};
  for (int i = threadIdx.x; i < 3 * 32 * 25; i += 256) {
    const int br = i / 800, c = (i / 25) % 32, t = i % 25;
    const int k = ksz[br];
    s_dw[i] = t < k * k ? rbf(wt.dw_w[br][c * k * k + t]) : 0.f;
  }
  for (int i = threadIdx.x; i < 96; i += 256) s_dwb[i] = wt.dw_b[i / 32][i % 32];
  for (int i = threadIdx.x; i < 3 * 64 * 32; i += 256) s_pw[i] = rbf(wt.pw_w[i / 2048][i % 2048]);
  for (int i = threadIdx.x; i < 192; i += 256) s_pwb[i] = wt.pw_b[i / 64][i % 64];
  // BN-applied input tile with halo (zero padding is applied AFTER BatchNorm, as in the reference)
  for (int i = threadIdx.x; i < kTP * kTP * 4; i += 256) {
    const int cg8 = i & 3, px = i >> 2;
    const int ty = px / kTP, tx = px % kTP;
    const int ih = h0 + ty - kHalo, iw = w0 + tx - kHalo;
    float v[8];
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      load8(y0 + (((long)b * H + ih) * W + iw) * 32 + cg8 * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = rbf(v[j] * scale[cg8 * 8 + j] + shift[cg8 * 8 + j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) tile[(cg8 * 8 + j) * kPlane + px] = v[j];
  }
  __syncthreads();
  const int ty = threadIdx.x / kT, tx = threadIdx.x % kT;
  const long pix = ((long)b * H + h0 + ty) * W + w0 + tx;
  for (int br = 0; br < 3; ++br) {
    const int k = ksz[br], pad = k / 2;
    float d[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      float a = 0.f;
      const float* tp = tile + c * kPlane + (ty + kHalo - pad) * kTP + (tx + kHalo - pad);
      const float* wp = s_dw + (br * 32 + c) * 25;
      for (int kh = 0; kh < k; ++kh)
        for (int kw = 0; kw < k; ++kw) a += tp[kh * kTP + kw] * wp[kh * k + kw];
      d[c] = rbf(a + s_dwb[br * 32 + c]);
    }
    bf16* dst = cat + pix * 192 + br * 64;
    for (int o8 = 0; o8 < 8; ++o8) {
      float out8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int o = o8 * 8 + j;
        const float4* wr = reinterpret_cast<const float4*>(s_pw + (br * 64 + o) * 32);
        float a = 0.f;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 wv = wr[c4];
          a += d[4 * c4] * wv.x + d[4 * c4 + 1] * wv.y + d[4 * c4 + 2] * wv.z + d[4 * c4 + 3] * wv.w;
        }
        const float t = rbf(a + s_pwb[br * 64 + o]);
        out8[j] = t > 0.f ? t : t * slope;
      }
      store8(dst + o8 * 8, out8);
    }
  }
}

}  // namespace lun

using namespace lun;

extern "C" {

int lun_fe_conv1(const float* x_nchw, const float* w, const float* bias, void* y, float* stats, int B, int H, int W,
                 float slope, void* stream) {
  const long total = (long)B * H * W;
  fe_conv1_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_nchw, w, bias, (bf16*)y, stats, B,
                                                                               H, W, slope);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_fe_branches(const void* y0, const float* scale, const float* shift, const float* const* dw_w,
                    const float* const* dw_b, const float* const* pw_w, const float* const* pw_b, void* cat, int B,
                    int H, int W, float slope, void* stream) {
  if (H % kT || W % kT) return LUN_E_SHAPE;
  FeBranchW wt;
  for (int i = 0; i < 3; ++i) {
    wt.dw_w[i] = dw_w[i]; wt.dw_b[i] = dw_b[i]; wt.pw_w[i] = pw_w[i]; wt.pw_b[i] = pw_b[i];
  }
  const int smem = (32 * kPlane + 3 * 32 * 25 + 96 + 3 * 64 * 32 + 192) * (int)sizeof(float);
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(fe_branch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return LUN_E_ATTR;
    configured = true;
  }
  dim3 grid(W / kT, H / kT, B);
  fe_branch_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)y0, scale, shift, wt, (bf16*)cat, H, W,
                                                             slope);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

}  // extern "C"
