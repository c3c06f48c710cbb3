// Bandwidth-bound kernels of the Teacher (LunarMoETeacher) path: BatchNorm finalize / apply, the ExpertBlock residual
// epilogue (forward and backward), channel statistics, the as-executed local attention rows and the proj expansion.
// All tensors are NHWC bf16 [B, HW, C]; a thread owns 8 consecutive channels (16-byte accesses), a block owns a
// set of pixel lanes so per-channel reductions finish in shared memory with one atomic per channel per block.
//
// Reference call sites: lunar_evaluator.py:71-103 (FE BatchNorm/LeakyReLU/Dropout), :189-227 (attention),
// :241-258 (ExpertBlock conv->LeakyReLU->BN->Dropout2d, layer_scale), :260-275 (residual + leaky_relu).
#include "../../include/lunaris_b200.h"
#include "elem_common.cuh"
#include "launch_count.cuh"
#include <stdlib.h>

namespace lun {

constexpr int kEThreads = 256;
constexpr int kUR = 2;  // three input streams in the backward reduce pass
constexpr int kU = 4;   // pixels per thread per iteration: all loads are issued before any is consumed

struct ChanGeom {
  int cg;      // channel groups of 8
  int lanes;   // pixel lanes per block
  __device__ __forceinline__ ChanGeom(int C) : cg(C >> 3), lanes(kEThreads / (C >> 3)) {}
};


__device__ __forceinline__ void lds8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// sums[k][c] partial per thread -> block reduce -> atomicAdd(dst_k[c]). NS = number of statistics.
template <int NS>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[NS][8], float* smem, int C, int lane_px, int cgi,
                                                     bool active, float* const (&dst)[NS]) {
  // smem layout [lanes][NS][C]
  const ChanGeom g(C);
  if (active) {
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) smem[(lane_px * NS + k) * C + cgi * 8 + j] = acc[k][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NS * C; i += kEThreads) {
    float s = 0.f;
    for (int l = 0; l < g.lanes; ++l) s += smem[l * NS * C + i];
    const int k = i / C, c = i % C;
    atomicAdd(dst[k] + c, s);
  }
}

// ------------------------------------------------------------------------------------------- channel statistics
__global__ void __launch_bounds__(kEThreads) channel_stats_kernel(const bf16* __restrict__ x, float* __restrict__ stats,
                                                                  long P, int C) {
  extern __shared__ float smem[];
  const ChanGeom g(C);
  const int cgi = threadIdx.x % g.cg, lane_px = threadIdx.x / g.cg;
  const bool active = lane_px < g.lanes;
  float acc[2][8] = {};
  if (active) {
    for (long p0 = (long)blockIdx.x * g.lanes * kU + lane_px; p0 < P; p0 += (long)gridDim.x * g.lanes * kU) {
      // raw 16-byte vectors first, unpacked only after ALL loads are issued (unpacking into floats right behind each
      // load made the compiler reuse one destination register quad: one load in flight per thread, 2.9 TB/s)
      uint4 xr[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const long p = p0 + (long)u * g.lanes;
        if (p < P) xr[u] = ldg16(x + p * C + cgi * 8);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const long p = p0 + (long)u * g.lanes;
        if (p >= P) continue;
        float v[8];
        unpack8(xr[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] += v[j];
          acc[1][j] += v[j] * v[j];
        }
      }
    }
  }
  float* const dst[2] = {stats, stats + C};
  block_channel_reduce<2>(acc, smem, C, lane_px, cgi, active, dst);
}

// ------------------------------------------------------------------------------------------- BN finalize
// stats = [sum | sumsq] over n elements per channel. Produces the normalisation affine (scale, shift), keeps
// mean / rstd for backward and applies the running-statistics update `n_updates` times (momentum, unbiased var).
__global__ void bn_finalize_kernel(const float* __restrict__ stats, double n, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean,
                                   float* __restrict__ rvar, long long* __restrict__ nbt, int n_updates,
                                   float momentum, float eps, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = stats[c] / n;
  double var = stats[C + c] / n - mean * mean;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;
  mean_out[c] = (float)mean;
  rstd_out[c] = rstd;
  if (n_updates > 0 && rmean) {
    const float var_u = (float)(n > 1 ? var * n / (n - 1) : var);
    float rm = rmean[c], rv = rvar[c];
    for (int i = 0; i < n_updates; ++i) {
      rm = (1.f - momentum) * rm + momentum * (float)mean;
      rv = (1.f - momentum) * rv + momentum * var_u;
    }
    rmean[c] = rm;
    rvar[c] = rv;
    if (c == 0 && nbt) *nbt += n_updates;
  }
}

// ------------------------------------------------------------------------------------------- affine forward
struct AffineArgs {
  const bf16* x;          // [B,HW,C]
  const float* scale;     // [C]
  const float* shift;     // [C]
  const float* m2;        // [B,C] Dropout2d keep-mask (0 or 1/(1-p), bf16-representable) or null
  const float* ls;        // [C] layer scale or null
  const bf16* idn;        // identity [B,HW,C] or null
  const float* id_scale;  // [C] affine on the identity (shortcut BatchNorm) or null
  const float* id_shift;
  bf16* y;
  float* pool;            // [B,C] sums over HW of the (unrounded) result or null
  unsigned long long seed;
  unsigned int thresh16;  // elementwise dropout threshold (0 = off)
  float drop_scale;
  int B, HW, C;
  float slope;            // leaky slope applied after the residual add when idn != null
};

__global__ void __launch_bounds__(kEThreads, 3) affine_fwd_kernel(const AffineArgs a) {
  extern __shared__ __align__(16) float smem[];
  const ChanGeom g(a.C);
  const int C = a.C;
  float* sP = smem;                       // [6][C]: scale, shift, m2, ls, id_scale, id_shift
  float* sRed = smem + 6 * C;             // [lanes][C] pooling partials
  const int cgi = threadIdx.x % g.cg, lane_px = threadIdx.x / g.cg;
  const bool active = lane_px < g.lanes;
  const int b = blockIdx.y;
  const int c0 = cgi * 8;
  for (int c = threadIdx.x; c < C; c += kEThreads) {
    sP[c] = a.scale ? a.scale[c] : 1.f;
    sP[C + c] = a.shift ? a.shift[c] : 0.f;
    sP[2 * C + c] = a.m2 ? a.m2[(size_t)b * C + c] : 1.f;
    sP[3 * C + c] = a.ls ? a.ls[c] : 1.f;
    sP[4 * C + c] = a.id_scale ? a.id_scale[c] : 1.f;
    sP[5 * C + c] = a.id_scale ? a.id_shift[c] : 0.f;
  }
  __syncthreads();
  float acc[1][8] = {};
  if (active) {
    for (int p0 = blockIdx.x * g.lanes * kU + lane_px; p0 < a.HW; p0 += gridDim.x * g.lanes * kU) {
      uint4 xr[kU], ir[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int p = p0 + u * g.lanes;
        if (p < a.HW) {
          const size_t off = ((size_t)b * a.HW + p) * C + c0;
          xr[u] = ldg16(a.x + off);
          if (a.idn) ir[u] = ldg16(a.idn + off);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int p = p0 + u * g.lanes;
        if (p >= a.HW) continue;
        const size_t off = ((size_t)b * a.HW + p) * C + c0;
        float v[8], t0[8], t1[8];
        unpack8(xr[u], v);
        lds8(sP + c0, t0);
        lds8(sP + C + c0, t1);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = rbf(v[j] * t0[j] + t1[j]);
        if (a.m2) {
          lds8(sP + 2 * C + c0, t0);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = rbf(v[j] * t0[j]);
        }
        if (a.thresh16) {
          bool keep[8];
          drop_keep8(a.seed, off >> 3, a.thresh16, keep);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = keep[j] ? rbf(v[j] * a.drop_scale) : 0.f;
        }
        if (a.ls) {
          lds8(sP + 3 * C + c0, t0);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= t0[j];
        }
        if (a.idn) {
          float iv[8];
          unpack8(ir[u], iv);
          if (a.id_scale) {
            lds8(sP + 4 * C + c0, t0);
            lds8(sP + 5 * C + c0, t1);
#pragma unroll
            for (int j = 0; j < 8; ++j) iv[j] = rbf(iv[j] * t0[j] + t1[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float sum = v[j] + iv[j];
            v[j] = sum > 0.f ? sum : sum * a.slope;
          }
        }
        store8(a.y + off, v);
        if (a.pool) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[0][j] += v[j];
        }
      }
    }
  }
  if (a.pool) {
    float* const dst[1] = {a.pool + (size_t)b * C};
    block_channel_reduce<1>(acc, sRed, C, lane_px, cgi, active, dst);
  }
}

// ExpertBlock tail, specialised (the 24 launches per step that carry layer scale + residual):
//   out = leaky( bf16(bf16(x * s + t) * m2) * ls + identity' ),  identity' = bf16(idn * ids + idt) or idn
// Same arithmetic as affine_fwd_kernel with half the instructions: the per-thread channel parameters live in registers
// (a thread keeps its channel octet for all its pixels), the Dropout2d product is ONE packed HMUL2.BF16 per channel pair
// (m2 is bf16-representable, so the packed product rounds exactly like bf16(fp32 product)), layer scale and residual
// are one FFMA, leaky_relu is max(v, slope * v). ncu on the generic kernel: 57 % issue-slot utilisation at 61 % of DRAM
// peak - it was bound by instruction issue, not by HBM.
template <bool ID_AFFINE>
__global__ void __launch_bounds__(kEThreads, 2) affine_tail_kernel(const AffineArgs a) {
  extern __shared__ __align__(16) float smem[];
  const ChanGeom g(a.C);
  const int C = a.C;
  const int cgi = threadIdx.x % g.cg, lane_px = threadIdx.x / g.cg;
  const bool active = lane_px < g.lanes;
  const int b = blockIdx.y;
  const int c0 = cgi * 8;
  float sc[8], sh[8], ls[8], isc[ID_AFFINE ? 8 : 1], ish[ID_AFFINE ? 8 : 1];
  __nv_bfloat162 m2p[4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = a.scale[c0 + j];
    sh[j] = a.shift[c0 + j];
    ls[j] = a.ls[c0 + j];
    if (ID_AFFINE) {
      isc[j] = a.id_scale[c0 + j];
      ish[j] = a.id_shift[c0 + j];
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    m2p[j] = a.m2 ? __floats2bfloat162_rn(a.m2[(size_t)b * C + c0 + 2 * j], a.m2[(size_t)b * C + c0 + 2 * j + 1])
                  : __floats2bfloat162_rn(1.f, 1.f);
  const float slope = a.slope;
  float acc[1][8] = {};
  if (active) {
    for (int p0 = blockIdx.x * g.lanes * kU + lane_px; p0 < a.HW; p0 += gridDim.x * g.lanes * kU) {
      uint4 xr[kU], ir[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int p = p0 + u * g.lanes;
        if (p < a.HW) {
          const size_t off = ((size_t)b * a.HW + p) * C + c0;
          xr[u] = ldg16(a.x + off);
          ir[u] = ldg16(a.idn + off);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int p = p0 + u * g.lanes;
        if (p >= a.HW) continue;
        const size_t off = ((size_t)b * a.HW + p) * C + c0;
        const uint32_t xw[4] = {xr[u].x, xr[u].y, xr[u].z, xr[u].w};
        const uint32_t iw[4] = {ir[u].x, ir[u].y, ir[u].z, ir[u].w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // BatchNorm affine in fp32, rounded to bf16 as a pair; Dropout2d as one packed bf16 multiply
          const float x0 = __uint_as_float(xw[j] << 16), x1 = __uint_as_float(xw[j] & 0xffff0000u);
          __nv_bfloat162 v = __floats2bfloat162_rn(fmaf(x0, sc[2 * j], sh[2 * j]), fmaf(x1, sc[2 * j + 1], sh[2 * j + 1]));
          v = __hmul2(v, m2p[j]);
          const uint32_t vw = *reinterpret_cast<const uint32_t*>(&v);
          float i0 = __uint_as_float(iw[j] << 16), i1 = __uint_as_float(iw[j] & 0xffff0000u);
          if (ID_AFFINE) {
            const __nv_bfloat162 ib = __floats2bfloat162_rn(fmaf(i0, isc[2 * j], ish[2 * j]),
                                                            fmaf(i1, isc[2 * j + 1], ish[2 * j + 1]));
            const uint32_t ibw = *reinterpret_cast<const uint32_t*>(&ib);
            i0 = __uint_as_float(ibw << 16);
            i1 = __uint_as_float(ibw & 0xffff0000u);
          }
          float o0 = fmaf(__uint_as_float(vw << 16), ls[2 * j], i0);
          float o1 = fmaf(__uint_as_float(vw & 0xffff0000u), ls[2 * j + 1], i1);
          o0 = fmaxf(o0, o0 * slope);                 // leaky_relu for 0 < slope < 1
          o1 = fmaxf(o1, o1 * slope);
          acc[0][2 * j] += o0;
          acc[0][2 * j + 1] += o1;
          const __nv_bfloat162 ob = __floats2bfloat162_rn(o0, o1);
          ow[j] = *reinterpret_cast<const uint32_t*>(&ob);
        }
        *reinterpret_cast<uint4*>(a.y + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
  }
  if (a.pool) {
    float* const dst[1] = {a.pool + (size_t)b * C};
    block_channel_reduce<1>(acc, smem, C, lane_px, cgi, active, dst);
  }
}

// ------------------------------------------------------------------------------------------- block epilogue backward
struct BlkBwdArgs {
  const bf16* dout;     // [B,HW,C] upstream gradient, or null when gpool is used
  const float* gpool;   // [B,C] gradient of the HW-sum pooled features (dout = gpool broadcast), or null
  const bf16* out;      // [B,HW,C] block output (sign gives the leaky_relu derivative), or null when slope_out == 1
  const bf16* a;        // [B,HW,C] BatchNorm input (post-LeakyReLU conv output)
  const float* mean;    // [C]
  const float* rstd;    // [C]
  const float* m2;      // [B,C] or null
  bf16* dpre;           // [B,HW,C] gradient w.r.t. the pre-activation sum (= gradient of the identity branch)
  float* t1;            // [C] sum dpre*m2
  float* t2;            // [C] sum dpre*m2*xhat
  int B, HW, C;
  float slope_out;
};

__global__ void __launch_bounds__(kEThreads, 3) blk_bwd_reduce_kernel(const BlkBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const ChanGeom g(a.C);
  const int C = a.C;
  float* sP = smem;                       // [4][C]: mean, rstd, m2, gpool
  float* sRed = smem + 4 * C;             // [lanes][2][C]
  const int cgi = threadIdx.x % g.cg, lane_px = threadIdx.x / g.cg;
  const bool active = lane_px < g.lanes;
  const int b = blockIdx.y, c0 = cgi * 8;
  for (int c = threadIdx.x; c < C; c += kEThreads) {
    sP[c] = a.mean[c];
    sP[C + c] = a.rstd[c];
    sP[2 * C + c] = a.m2 ? a.m2[(size_t)b * C + c] : 1.f;
    sP[3 * C + c] = a.gpool ? a.gpool[(size_t)b * C + c] : 0.f;
  }
  __syncthreads();
  float acc[2][8] = {};
  if (active) {
    for (int p0 = blockIdx.x * g.lanes * kUR + lane_px; p0 < a.HW; p0 += gridDim.x * g.lanes * kUR) {
      uint4 dr[kUR], ar[kUR], orr[kUR];
#pragma unroll
      for (int u = 0; u < kUR; ++u) {
        const int p = p0 + u * g.lanes;
        if (p < a.HW) {
          const size_t off = ((size_t)b * a.HW + p) * C + c0;
          if (a.dout) dr[u] = ldg16(a.dout + off);
          if (a.out) orr[u] = ldg16(a.out + off);
          ar[u] = ldg16(a.a + off);
        }
      }
#pragma unroll
      for (int u = 0; u < kUR; ++u) {
        const int p = p0 + u * g.lanes;
        if (p >= a.HW) continue;
        const size_t off = ((size_t)b * a.HW + p) * C + c0;
        float d[8], t0[8], t1[8];
        if (a.dout) unpack8(dr[u], d);
        else lds8(sP + 3 * C + c0, d);
        if (a.out) {
          unpack8(orr[u], t0);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = t0[j] > 0.f ? d[j] : d[j] * a.slope_out;
        }
        if (a.dpre) {
          store8(a.dpre + off, d);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = rbf(d[j]);
        }
        lds8(sP + 2 * C + c0, t0);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] *= t0[j];
        unpack8(ar[u], t0);
        lds8(sP + c0, t1);
#pragma unroll
        for (int j = 0; j < 8; ++j) t0[j] -= t1[j];
        lds8(sP + C + c0, t1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] += d[j];
          acc[1][j] += d[j] * t0[j] * t1[j];
        }
      }
    }
  }
  float* const dst[2] = {a.t1, a.t2};
  block_channel_reduce<2>(acc, sRed, C, lane_px, cgi, active, dst);
}

// blk_bwd_reduce_kernel specialised for its common call (stored upstream gradient, block output and dpre all present):
// mean / rstd / Dropout2d factor of the thread's channel octet in registers, packed sign select, as affine_tail_kernel.
template <bool HAS_DOUT>   // false: the upstream gradient is the broadcast pooled gradient gpool[b][c] (last block of an expert)
__global__ void __launch_bounds__(kEThreads, 2) blk_bwd_reduce_fast_kernel(const BlkBwdArgs a) {
  extern __shared__ __align__(16) float smem[];
  const ChanGeom g(a.C);
  const int C = a.C;
  const int cgi = threadIdx.x % g.cg, lane_px = threadIdx.x / g.cg;
  const bool active = lane_px < g.lanes;
  const int b = blockIdx.y, c0 = cgi * 8;
  float mean[8], rstd[8], m2[8], gp[HAS_DOUT ? 1 : 8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    mean[j] = a.mean[c0 + j];
    rstd[j] = a.rstd[c0 + j];
    m2[j] = a.m2 ? a.m2[(size_t)b * C + c0 + j] : 1.f;
    if (!HAS_DOUT) gp[j] = a.gpool[(size_t)b * C + c0 + j];
  }
  const float slope = a.slope_out;
  float acc[2][8] = {};
  if (active) {
    for (int p0 = blockIdx.x * g.lanes * kUR + lane_px; p0 < a.HW; p0 += gridDim.x * g.lanes * kUR) {
      uint4 dr[kUR], ar[kUR], orr[kUR];
#pragma unroll
      for (int u = 0; u < kUR; ++u) {
        const int p = p0 + u * g.lanes;
        if (p < a.HW) {
          const size_t off = ((size_t)b * a.HW + p) * C + c0;
          if (HAS_DOUT) dr[u] = ldg16(a.dout + off);
          orr[u] = ldg16(a.out + off);
          ar[u] = ldg16(a.a + off);
        }
      }
#pragma unroll
      for (int u = 0; u < kUR; ++u) {
        const int p = p0 + u * g.lanes;
        if (p >= a.HW) continue;
        const size_t off = ((size_t)b * a.HW + p) * C + c0;
        const uint32_t dw[4] = {HAS_DOUT ? dr[u].x : 0u, HAS_DOUT ? dr[u].y : 0u, HAS_DOUT ? dr[u].z : 0u,
                                HAS_DOUT ? dr[u].w : 0u};
        const uint32_t ow_[4] = {orr[u].x, orr[u].y, orr[u].z, orr[u].w};
        const uint32_t aw[4] = {ar[u].x, ar[u].y, ar[u].z, ar[u].w};
        uint32_t pw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float d0 = HAS_DOUT ? __uint_as_float(dw[j] << 16) : gp[HAS_DOUT ? 0 : 2 * j];
          float d1 = HAS_DOUT ? __uint_as_float(dw[j] & 0xffff0000u) : gp[HAS_DOUT ? 0 : 2 * j + 1];
          const float o0 = __uint_as_float(ow_[j] << 16), o1 = __uint_as_float(ow_[j] & 0xffff0000u);
          d0 = o0 > 0.f ? d0 : d0 * slope;
          d1 = o1 > 0.f ? d1 : d1 * slope;
          const __nv_bfloat162 db = __floats2bfloat162_rn(d0, d1);
          pw[j] = *reinterpret_cast<const uint32_t*>(&db);
          // the sums run over what was stored (bf16), like the generic kernel
          const float g0 = __uint_as_float(pw[j] << 16) * m2[2 * j], g1 = __uint_as_float(pw[j] & 0xffff0000u) * m2[2 * j + 1];
          const float a0 = __uint_as_float(aw[j] << 16), a1 = __uint_as_float(aw[j] & 0xffff0000u);
          acc[0][2 * j] += g0;
          acc[0][2 * j + 1] += g1;
          acc[1][2 * j] += g0 * (a0 - mean[2 * j]) * rstd[2 * j];
          acc[1][2 * j + 1] += g1 * (a1 - mean[2 * j + 1]) * rstd[2 * j + 1];
        }
        *reinterpret_cast<uint4*>(a.dpre + off) = make_uint4(pw[0], pw[1], pw[2], pw[3]);
      }
    }
  }
  float* const dst[2] = {a.t1, a.t2};
  block_channel_reduce<2>(acc, smem, C, lane_px, cgi, active, dst);
}

struct BlkBwdApplyArgs {
  const bf16* dpre;     // [B,HW,C] (or null with gpool/out as in the reduce pass)
  const float* gpool;
  const bf16* out;
  const bf16* a;        // BN input
  const float* mean; const float* rstd; const float* gamma;
  const float* ls;      // [C] or null (1)
  const float* m2;      // [B,C] or null
  const float* t1; const float* t2;   // from the reduce pass
  bf16* dz;             // [B,HW,C] gradient w.r.t. the conv output (before LeakyReLU)
  float* dbias;         // [C] column sums of dz (conv bias gradient) or null
  int B, HW, C;
  float slope_out;      // leaky slope of the block output (used when dpre == null)
  float slope_a;        // leaky slope between conv and BN (1 = none)
  float inv_n;          // 1 / (B*HW)
};

__global__ void __launch_bounds__(kEThreads, 3) blk_bwd_apply_kernel(const BlkBwdApplyArgs a) {
  extern __shared__ __align__(16) float smem[];
  const ChanGeom g(a.C);
  const int C = a.C;
  float* sP = smem;                       // [4][C]: P1, P2, P3 (dz = P1*d + P2*a + P3), gpool
  float* sRed = smem + 4 * C;             // [lanes][C]
  const int cgi = threadIdx.x % g.cg, lane_px = threadIdx.x / g.cg;
  const bool active = lane_px < g.lanes;
  const int b = blockIdx.y, c0 = cgi * 8;
  for (int c = threadIdx.x; c < C; c += kEThreads) {
    // dz = gamma*rstd * (g - S1/n - xhat*S2/n),  g = d*ls*m2,  xhat = (a - mean)*rstd
    const float ls = a.ls ? a.ls[c] : 1.f;
    const float rstd = a.rstd[c], mean = a.mean[c];
    const float k0 = a.gamma[c] * rstd;
    const float s1 = ls * a.t1[c] * a.inv_n, s2 = ls * a.t2[c] * a.inv_n;
    sP[c] = k0 * ls * (a.m2 ? a.m2[(size_t)b * C + c] : 1.f);
    sP[C + c] = -k0 * s2 * rstd;
    sP[2 * C + c] = k0 * (s2 * rstd * mean - s1);
    sP[3 * C + c] = a.gpool ? a.gpool[(size_t)b * C + c] : 0.f;
  }
  __syncthreads();
  float acc[1][8] = {};
  if (active) {
    for (int p0 = blockIdx.x * g.lanes * kU + lane_px; p0 < a.HW; p0 += gridDim.x * g.lanes * kU) {
      uint4 dr[kU], ar[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int p = p0 + u * g.lanes;
        if (p < a.HW) {
          const size_t off = ((size_t)b * a.HW + p) * C + c0;
          if (a.dpre) dr[u] = ldg16(a.dpre + off);
          else if (a.out) dr[u] = ldg16(a.out + off);
          ar[u] = ldg16(a.a + off);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int p = p0 + u * g.lanes;
        if (p >= a.HW) continue;
        const size_t off = ((size_t)b * a.HW + p) * C + c0;
        float d[8], av[8], t0[8], t1[8];
        if (a.dpre) unpack8(dr[u], d);
        else {
          lds8(sP + 3 * C + c0, d);
          if (a.out) {
            unpack8(dr[u], t0);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = t0[j] > 0.f ? d[j] : d[j] * a.slope_out;
          }
        }
        unpack8(ar[u], av);
        lds8(sP + c0, t0);
        lds8(sP + C + c0, t1);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = d[j] * t0[j] + av[j] * t1[j];
        lds8(sP + 2 * C + c0, t0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float z = d[j] + t0[j];
          if (av[j] <= 0.f) z *= a.slope_a;
          d[j] = z;
        }
        store8(a.dz + off, d);
        if (a.dbias) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[0][j] += rbf(d[j]);
        }
      }
    }
  }
  if (a.dbias) {
    float* const dst[1] = {a.dbias};
    block_channel_reduce<1>(acc, sRed, C, lane_px, cgi, active, dst);
  }
}

// blk_bwd_apply_kernel specialised for its common call (the upstream gradient is the stored dpre tensor): the three
// per-channel coefficients of  dz = P1 * d + P2 * a + P3  live in registers (a thread keeps its channel octet and a
// block stays inside one image, so the Dropout2d factor inside P1 is fixed too), as in affine_tail_kernel - the generic
// kernel re-read them from shared memory for every pixel and ran at 5.0 TB/s where the tail pass reaches 6.2.
__global__ void __launch_bounds__(kEThreads, 2) blk_bwd_apply_fast_kernel(const BlkBwdApplyArgs a) {
  extern __shared__ __align__(16) float smem[];
  const ChanGeom g(a.C);
  const int C = a.C;
  const int cgi = threadIdx.x % g.cg, lane_px = threadIdx.x / g.cg;
  const bool active = lane_px < g.lanes;
  const int b = blockIdx.y, c0 = cgi * 8;
  float p1[8], p2[8], p3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    const float ls = a.ls ? a.ls[c] : 1.f;
    const float rstd = a.rstd[c], mean = a.mean[c];
    const float k0 = a.gamma[c] * rstd;
    const float s1 = ls * a.t1[c] * a.inv_n, s2 = ls * a.t2[c] * a.inv_n;
    p1[j] = k0 * ls * (a.m2 ? a.m2[(size_t)b * C + c] : 1.f);
    p2[j] = -k0 * s2 * rstd;
    p3[j] = k0 * (s2 * rstd * mean - s1);
  }
  const float slope_a = a.slope_a;
  float acc[1][8] = {};
  if (active) {
    for (int p0 = blockIdx.x * g.lanes * kU + lane_px; p0 < a.HW; p0 += gridDim.x * g.lanes * kU) {
      uint4 dr[kU], ar[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int p = p0 + u * g.lanes;
        if (p < a.HW) {
          const size_t off = ((size_t)b * a.HW + p) * C + c0;
          dr[u] = ldg16(a.dpre + off);
          ar[u] = ldg16(a.a + off);
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int p = p0 + u * g.lanes;
        if (p >= a.HW) continue;
        const size_t off = ((size_t)b * a.HW + p) * C + c0;
        const uint32_t dw[4] = {dr[u].x, dr[u].y, dr[u].z, dr[u].w}, aw[4] = {ar[u].x, ar[u].y, ar[u].z, ar[u].w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float d0 = __uint_as_float(dw[j] << 16), d1 = __uint_as_float(dw[j] & 0xffff0000u);
          const float a0 = __uint_as_float(aw[j] << 16), a1 = __uint_as_float(aw[j] & 0xffff0000u);
          float z0 = fmaf(d0, p1[2 * j], fmaf(a0, p2[2 * j], p3[2 * j]));
          float z1 = fmaf(d1, p1[2 * j + 1], fmaf(a1, p2[2 * j + 1], p3[2 * j + 1]));
          z0 = a0 <= 0.f ? z0 * slope_a : z0;
          z1 = a1 <= 0.f ? z1 * slope_a : z1;
          const __nv_bfloat162 zb = __floats2bfloat162_rn(z0, z1);
          ow[j] = *reinterpret_cast<const uint32_t*>(&zb);
          acc[0][2 * j] += __uint_as_float(ow[j] << 16);            // the conv bias gradient sums what was stored
          acc[0][2 * j + 1] += __uint_as_float(ow[j] & 0xffff0000u);
        }
        *reinterpret_cast<uint4*>(a.dz + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
    }
  }
  if (a.dbias) {
    float* const dst[1] = {a.dbias};
    block_channel_reduce<1>(acc, smem, C, lane_px, cgi, active, dst);
  }
}

// ------------------------------------------------------------------------------------------- attention (as executed)
// One warp per (image b, query slot i < nq, head). Query slot i < nc-1 is token 32*i (row 0 of chunk i); slots
// nc-1 .. nc+30 are the 32 tokens of the last chunk. Output row i of att_small is what the reference leaves at
// position i of `out` before proj (lunar_evaluator.py:203-216). Lane j owns key j of the chunk.
// Loads are laid out so that HD/8 adjacent lanes read one 128-byte (or shorter) head row: every LDG.128 of the warp
// touches whole cache lines (the earlier one-row-per-lane layout was L1-tag bound: 32 lines per instruction).
//   q: [B, nq_pad, C] rows already gathered/projected (q_stride = C) or the full qkv tensor (q_stride = 3C)
//   k, v: base pointers of the K and V channel blocks, row stride kv_stride elements
template <int HD>
__global__ void __launch_bounds__(256) attn_ref_rows_kernel(const bf16* __restrict__ q, long q_img_stride,
                                                             int q_stride, int q_is_token_indexed,
                                                             const bf16* __restrict__ k, const bf16* __restrict__ v,
                                                             int kv_stride, bf16* __restrict__ att, int B, int N,
                                                             int C, int heads, int nq_pad, unsigned long long seed,
                                                             unsigned int thresh16, float drop_scale) {
  constexpr int LPR = HD / 8;        // lanes per row (16-byte chunks per head row)
  constexpr int RPI = 32 / LPR;      // rows per load instruction
  constexpr int NT = 32 / RPI;       // load instructions per 32-row chunk (== LPR)
  const int warp = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int nc = N / 32;
  const int nq = nc + 31;
  if (warp >= B * nq * heads) return;
  const int h = warp % heads;
  const int i = (warp / heads) % nq;
  const int b = warp / (heads * nq);
  const int chunk = i < nc - 1 ? i : nc - 1;
  const int qtok = i < nc - 1 ? 32 * i : 32 * (nc - 1) + (i - (nc - 1));
  const int c8 = lane % LPR, rg = lane / LPR;
  const long qrow = q_is_token_indexed ? qtok : i;
  const uint4 qv = __ldg(reinterpret_cast<const uint4*>(q + b * q_img_stride + qrow * q_stride + h * HD + c8 * 8));
  uint4 kr[NT], vr[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const size_t roff = ((size_t)b * N + 32 * chunk + t * RPI + rg) * kv_stride + h * HD + c8 * 8;
    kr[t] = __ldg(reinterpret_cast<const uint4*>(k + roff));
    vr[t] = __ldg(reinterpret_cast<const uint4*>(v + roff));
  }
  float qf[8];
  unpack8(qv, qf);
  const float scale = rbf(rsqrtf((float)HD));
  float s[NT];
  float m = -3.0e38f;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    float kf[8];
    unpack8(kr[t], kf);
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) d += qf[j] * kf[j];
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    // reference dtype flow under bf16 autocast: scores and the scale product are bf16, softmax is fp32
    s[t] = rbf(rbf(d) * scale);
    m = fmaxf(m, s[t]);
  }
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    s[t] = __expf(s[t] - m);
    sum += s[t];
  }
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  float o8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    float p = s[t] * inv;
    if (thresh16) {
      const unsigned long long idx = ((unsigned long long)warp << 5) + (t * RPI + rg);
      p = drop_keep1(seed, idx, thresh16) ? p * drop_scale : 0.f;
    }
    p = rbf(p);
    float vf[8];
    unpack8(vr[t], vf);
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] += p * vf[j];
  }
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] += __shfl_xor_sync(0xffffffffu, o8[j], o);
  }
  if (rg == 0) store8(att + ((size_t)b * nq_pad + i) * C + h * HD + c8 * 8, o8);
}

// out[b, i, :] = x[b, qtok(i), :]: the N/32+31 query tokens the as-executed attention actually uses.
__global__ void gather_query_rows_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int C, int nq,
                                         int nq_pad) {
  const int i = blockIdx.x, b = blockIdx.y;
  const int nc = N / 32;
  const int qtok = i < nc - 1 ? 32 * i : 32 * (nc - 1) + (i - (nc - 1));
  const uint4* src = reinterpret_cast<const uint4*>(x + ((size_t)b * N + qtok) * C);
  uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * nq_pad + i) * C);
  for (int c = threadIdx.x; c < C / 8; c += blockDim.x) dst[c] = __ldg(src + c);
}

// h2[b,p,:] = dropout( p < nq ? proj_small[b,p,:] : bf16(bias) )   (lunar_evaluator.py:224-225)
__global__ void __launch_bounds__(kEThreads) proj_expand_kernel(const bf16* __restrict__ small,
                                                               const float* __restrict__ bias,
                                                               bf16* __restrict__ y, int B, int HW, int C, int nq,
                                                               int nq_pad, unsigned long long seed,
                                                               unsigned int thresh16, float drop_scale) {
  const ChanGeom g(C);
  const int cgi = threadIdx.x % g.cg, lane_px = threadIdx.x / g.cg;
  if (lane_px >= g.lanes) return;
  const int b = blockIdx.y, c0 = cgi * 8;
  // Outside the first nq rows every pixel is dropout(bf16(bias)): the kept value bf16(bias * scale) is a per-channel
  // constant, so those rows (97 % of the tensor) cost one hash per two elements and a select - no multiply, no rounding
  float bv[8];
  uint32_t bkeep[4];                                  // packed bf16 pairs of the kept value
#pragma unroll
  for (int j = 0; j < 8; ++j) bv[j] = rbf(bias[c0 + j]);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(thresh16 ? rbf(bv[2 * j] * drop_scale) : bv[2 * j],
                                                   thresh16 ? rbf(bv[2 * j + 1] * drop_scale) : bv[2 * j + 1]);
    bkeep[j] = *reinterpret_cast<const uint32_t*>(&t);
  }
  for (int p = blockIdx.x * g.lanes + lane_px; p < HW; p += gridDim.x * g.lanes) {
    const size_t off = ((size_t)b * HW + p) * C + c0;
    if (p >= nq) {
      uint32_t w[4] = {bkeep[0], bkeep[1], bkeep[2], bkeep[3]};
      if (thresh16) {
        const uint32_t k = drop_key(seed, off >> 3);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t h = hash32(k + i);           // same stream as drop_keep8
          w[i] = ((h & 0xFFFFu) >= thresh16 ? w[i] & 0x0000FFFFu : 0u) | ((h >> 16) >= thresh16 ? w[i] & 0xFFFF0000u : 0u);
        }
      }
      *reinterpret_cast<uint4*>(y + off) = make_uint4(w[0], w[1], w[2], w[3]);
      continue;
    }
    float v[8];
    load8(small + ((size_t)b * nq_pad + p) * C + c0, v);
    if (thresh16) {
      bool keep[8];
      drop_keep8(seed, off >> 3, thresh16, keep);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = keep[j] ? rbf(v[j] * drop_scale) : 0.f;
    }
    store8(y + off, v);
  }
}

// Backward of proj_drop + what proj's gradients need: dpo = mask * dh2; column sums of dpo (proj bias gradient);
// the first nq rows of every image (the only rows where proj's input is non-zero) gathered into dpo_small.
__global__ void __launch_bounds__(kEThreads) proj_bwd_gather_kernel(const bf16* __restrict__ dh2,
                                                                   bf16* __restrict__ dpo_small,
                                                                   float* __restrict__ dbias, int B, int HW, int C,
                                                                   int nq, int nq_pad, unsigned long long seed,
                                                                   unsigned int thresh16, float drop_scale) {
  extern __shared__ float smem[];
  const ChanGeom g(C);
  const int cgi = threadIdx.x % g.cg, lane_px = threadIdx.x / g.cg;
  const bool active = lane_px < g.lanes;
  const int b = blockIdx.y, c0 = cgi * 8;
  float acc[1][8] = {};
  // dbias == null: the column sums came out of the producing conv's epilogue (lun_conv_taps_dropsum_bf16); only the
  // first nq rows of every image are visited (gather + mask)
  const int lim = dbias ? HW : nq;
  if (active) {
    for (int p0 = blockIdx.x * g.lanes * kU + lane_px; p0 < lim; p0 += gridDim.x * g.lanes * kU) {
      uint4 xr[kU];                                   // raw vectors: all loads in flight before the first unpack
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int p = p0 + u * g.lanes;
        if (p < lim) xr[u] = ldg16(dh2 + ((size_t)b * HW + p) * C + c0);
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int p = p0 + u * g.lanes;
        if (p >= lim) continue;
        const size_t off = ((size_t)b * HW + p) * C + c0;
        float v[8];
        unpack8(xr[u], v);
        if (thresh16) {
          bool keep[8];
          drop_keep8(seed, off >> 3, thresh16, keep);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = keep[j] ? rbf(v[j] * drop_scale) : 0.f;
        }
        if (p < nq) store8(dpo_small + ((size_t)b * nq_pad + p) * C + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[0][j] += v[j];
      }
    }
  }
  if (dbias == nullptr) return;
  float* const dst[1] = {dbias};
  block_channel_reduce<1>(acc, smem, C, lane_px, cgi, active, dst);
}

static int elem_blocks_per_image(int HW, int C, int B) {
  const int lanes = kEThreads / (C / 8) * kU;
  int per = (HW + lanes - 1) / lanes;
  // many more blocks than resident slots so the tail wave is a small fraction of the work
  int want = (148 * 32 + B - 1) / B;
  if (want < 1) want = 1;
  // ... but a block should still walk >= 8 iterations, or its prologue (channel parameters, reduction epilogue) shows:
  // at C2 (batch 16, 256 channels) the 32 x SMs rule alone gave blocks of 1.7 iterations
  const int cap = per >= 8 ? per / 8 : 1;
  if (want > cap) want = cap;
  return per < want ? per : want;
}
static bool chan_ok(int C) { return C % 8 == 0 && C / 8 <= kEThreads && C >= 8; }

}  // namespace lun

using namespace lun;

extern "C" {

int lun_channel_stats_bf16(const void* x, long P, int C, float* stats, void* stream) {
  if (!chan_ok(C)) return LUN_E_SHAPE;
  const int lanes = kEThreads / (C / 8);
  long blocks = (P + lanes * kU - 1) / (lanes * kU);
  if (blocks > 148 * 8) blocks = 148 * 8;             // 8 x SMs measured best (32 x: 182 vs 165 us at C3 shapes)
  channel_stats_kernel<<<(int)blocks, kEThreads, lanes * 2 * C * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)x, stats, P, C);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_bn_finalize(const float* stats, double n, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, long long* num_batches_tracked, int n_updates, float momentum, float eps,
                    float* scale, float* shift, float* mean, float* rstd, int C, void* stream) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, n, gamma, beta, running_mean,
                                                                       running_var, num_batches_tracked, n_updates,
                                                                       momentum, eps, scale, shift, mean, rstd, C);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_affine_fwd_bf16(const void* x, const float* scale, const float* shift, const float* mask2d, const float* ls,
                        const void* identity, const float* id_scale, const float* id_shift, void* y, float* pool,
                        unsigned long long seed, float drop_p, int B, int HW, int C, float slope, void* stream) {
  if (!chan_ok(C)) return LUN_E_SHAPE;
  AffineArgs a;
  a.x = (const bf16*)x; a.scale = scale; a.shift = shift; a.m2 = mask2d; a.ls = ls; a.idn = (const bf16*)identity;
  a.id_scale = id_scale; a.id_shift = id_shift; a.y = (bf16*)y; a.pool = pool; a.seed = seed;
  a.thresh16 = drop_p > 0.f ? (unsigned int)(drop_p * 65536.f + 0.5f) : 0u;
  a.drop_scale = drop_p > 0.f ? __bfloat162float(__float2bfloat16_rn(1.f / (1.f - drop_p))) : 1.f;
  a.B = B; a.HW = HW; a.C = C; a.slope = slope;
  const int lanes = kEThreads / (C / 8);
  dim3 grid(elem_blocks_per_image(HW, C, B), B);
  static int tail_mode = -1;
  if (tail_mode < 0) {
    const char* e = getenv("LUN_AFFINE_TAIL");
    tail_mode = e ? atoi(e) : 1;
  }
  if (tail_mode && scale && shift && ls && identity && a.thresh16 == 0 && slope > 0.f && slope < 1.f) {
    // ExpertBlock tail: specialised kernel (parameters in registers, packed Dropout2d multiply)
    const size_t sm = pool ? (size_t)lanes * C * sizeof(float) : 0;
    if (id_scale) affine_tail_kernel<true><<<grid, kEThreads, sm, (cudaStream_t)stream>>>(a);
    else affine_tail_kernel<false><<<grid, kEThreads, sm, (cudaStream_t)stream>>>(a);
    lun::note_launch(1);
    return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
  }
  affine_fwd_kernel<<<grid, kEThreads, (6 * C + (pool ? lanes * C : 0)) * sizeof(float), (cudaStream_t)stream>>>(a);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_block_bwd_reduce_bf16(const void* dout, const float* gpool, const void* out, const void* bn_in,
                              const float* mean, const float* rstd, const float* mask2d, void* dpre, float* t1,
                              float* t2, int B, int HW, int C, float slope_out, void* stream) {
  if (!chan_ok(C)) return LUN_E_SHAPE;
  BlkBwdArgs a;
  a.dout = (const bf16*)dout; a.gpool = gpool; a.out = (const bf16*)out; a.a = (const bf16*)bn_in; a.mean = mean;
  a.rstd = rstd; a.m2 = mask2d; a.dpre = (bf16*)dpre; a.t1 = t1; a.t2 = t2; a.B = B; a.HW = HW; a.C = C;
  a.slope_out = slope_out;
  const int lanes = kEThreads / (C / 8);
  dim3 grid(elem_blocks_per_image(HW, C, B), B);
  if (dout && out && dpre)
    blk_bwd_reduce_fast_kernel<true><<<grid, kEThreads, lanes * 2 * C * sizeof(float), (cudaStream_t)stream>>>(a);
  else if (!dout && gpool && out && dpre)
    blk_bwd_reduce_fast_kernel<false><<<grid, kEThreads, lanes * 2 * C * sizeof(float), (cudaStream_t)stream>>>(a);
  else
    blk_bwd_reduce_kernel<<<grid, kEThreads, (4 * C + lanes * 2 * C) * sizeof(float), (cudaStream_t)stream>>>(a);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_block_bwd_apply_bf16(const void* dpre, const float* gpool, const void* out, const void* bn_in,
                             const float* mean, const float* rstd, const float* gamma, const float* ls,
                             const float* mask2d, const float* t1, const float* t2, void* dz, float* dbias, int B,
                             int HW, int C, float slope_out, float slope_a, void* stream) {
  if (!chan_ok(C)) return LUN_E_SHAPE;
  BlkBwdApplyArgs a;
  a.dpre = (const bf16*)dpre; a.gpool = gpool; a.out = (const bf16*)out; a.a = (const bf16*)bn_in; a.mean = mean;
  a.rstd = rstd; a.gamma = gamma; a.ls = ls; a.m2 = mask2d; a.t1 = t1; a.t2 = t2; a.dz = (bf16*)dz;
  a.dbias = dbias; a.B = B; a.HW = HW; a.C = C; a.slope_out = slope_out; a.slope_a = slope_a;
  a.inv_n = 1.f / ((float)B * (float)HW);
  const int lanes = kEThreads / (C / 8);
  dim3 grid(elem_blocks_per_image(HW, C, B), B);
  if (dpre)
    blk_bwd_apply_fast_kernel<<<grid, kEThreads, (dbias ? lanes * C : 0) * sizeof(float), (cudaStream_t)stream>>>(a);
  else
    blk_bwd_apply_kernel<<<grid, kEThreads, (4 * C + (dbias ? lanes * C : 0)) * sizeof(float), (cudaStream_t)stream>>>(a);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

static int launch_attn(const bf16* q, long q_img_stride, int q_stride, int q_tok, const bf16* k, const bf16* v,
                       int kv_stride, bf16* att, int B, int N, int C, int heads, int nq_pad, unsigned long long seed,
                       float drop_p, cudaStream_t st) {
  if (N % 32 || C % heads || (C / heads) % 8) return LUN_E_SHAPE;
  const int nq = N / 32 + 31;
  if (nq_pad < nq) return LUN_E_SHAPE;
  const long warps = (long)B * nq * heads;
  const unsigned int th = drop_p > 0.f ? (unsigned int)(drop_p * 65536.f + 0.5f) : 0u;
  const float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int hd = C / heads;
  const int blocks = (int)((warps + 7) / 8);
#define LUN_ATTN(HD_)                                                                                             \
  attn_ref_rows_kernel<HD_><<<blocks, 256, 0, st>>>(q, q_img_stride, q_stride, q_tok, k, v, kv_stride, att, B, N, C, \
                                                     heads, nq_pad, seed, th, ds)
  if (hd == 64) LUN_ATTN(64);
  else if (hd == 32) LUN_ATTN(32);
  else if (hd == 16) LUN_ATTN(16);
  else if (hd == 8) LUN_ATTN(8);
  else if (hd == 128) LUN_ATTN(128);
  else return LUN_E_SHAPE;
#undef LUN_ATTN
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_attn_ref_rows_bf16(const void* qkv, void* att_small, int B, int N, int C, int heads, int nq_pad,
                           unsigned long long seed, float drop_p, void* stream) {
  const bf16* base = (const bf16*)qkv;
  return launch_attn(base, (long)N * 3 * C, 3 * C, 1, base + C, base + 2 * C, 3 * C, (bf16*)att_small, B, N, C, heads,
                     nq_pad, seed, drop_p, (cudaStream_t)stream);
}

int lun_attn_ref_rows_split_bf16(const void* q_small, const void* kv, void* att_small, int B, int N, int C, int heads,
                                 int nq_pad, unsigned long long seed, float drop_p, void* stream) {
  const bf16* kvb = (const bf16*)kv;
  return launch_attn((const bf16*)q_small, (long)nq_pad * C, C, 0, kvb, kvb + C, 2 * C, (bf16*)att_small, B, N, C,
                     heads, nq_pad, seed, drop_p, (cudaStream_t)stream);
}

int lun_gather_query_rows_bf16(const void* x, void* out, int B, int N, int C, int nq_pad, void* stream) {
  if (N % 32 || C % 8) return LUN_E_SHAPE;
  const int nq = N / 32 + 31;
  if (nq_pad < nq) return LUN_E_SHAPE;
  dim3 grid(nq, B);
  gather_query_rows_kernel<<<grid, 64, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)out, N, C, nq, nq_pad);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_proj_expand_bf16(const void* proj_small, const float* bias, void* y, int B, int HW, int C, int nq, int nq_pad,
                         unsigned long long seed, float drop_p, void* stream) {
  if (!chan_ok(C)) return LUN_E_SHAPE;
  const unsigned int th = drop_p > 0.f ? (unsigned int)(drop_p * 65536.f + 0.5f) : 0u;
  const float ds = drop_p > 0.f ? __bfloat162float(__float2bfloat16_rn(1.f / (1.f - drop_p))) : 1.f;
  dim3 grid(elem_blocks_per_image(HW, C, B), B);
  proj_expand_kernel<<<grid, kEThreads, 0, (cudaStream_t)stream>>>((const bf16*)proj_small, bias, (bf16*)y, B, HW, C,
                                                                  nq, nq_pad, seed, th, ds);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_proj_bwd_gather_bf16(const void* dh2, void* dpo_small, float* dbias, int B, int HW, int C, int nq, int nq_pad,
                             unsigned long long seed, float drop_p, void* stream) {
  if (!chan_ok(C)) return LUN_E_SHAPE;
  const unsigned int th = drop_p > 0.f ? (unsigned int)(drop_p * 65536.f + 0.5f) : 0u;
  const float ds = drop_p > 0.f ? __bfloat162float(__float2bfloat16_rn(1.f / (1.f - drop_p))) : 1.f;
  const int lanes = kEThreads / (C / 8);
  dim3 grid(elem_blocks_per_image(dbias ? HW : nq, C, B), B);
  proj_bwd_gather_kernel<<<grid, kEThreads, lanes * C * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)dh2, (bf16*)dpo_small, dbias, B, HW, C, nq, nq_pad, seed, th, ds);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

}  // extern "C"
