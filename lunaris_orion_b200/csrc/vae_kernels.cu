// CUDA-core kernels of the VAE path (LunarisCoreVAE, lunar_generate.py): GroupNorm(8)+Mish forward / backward with the
// ResBlock and skip-connection adds fused, the 3-channel first / last convolutions (too thin for the tensor cores),
// and the reparameterisation. Activations are NHWC bf16 [B, HW, C]; images are NCHW fp32 as the trainer hands them.
#include "../../include/lunaris_b200.h"
#include "elem_common.cuh"
#include "launch_count.cuh"

namespace lun {

constexpr int kVT = 256;

// mish(v) = v * tanh(softplus(v)). With e = exp(v): tanh(log(1+e)) = n / (n + 2), n = e^2 + 2e  (one MUFU.EX2, one
// MUFU.RCP). v is clamped at 20 (torch's softplus threshold): there e = 4.9e8, n = 2.4e17 and n / (n + 2) rounds to
// exactly 1.0f, so no select is needed for the linear branch.
__device__ __forceinline__ float mish_tanh_sp(float v, float* e_out) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(v, 20.f) * 1.4426950408889634f));
  *e_out = e;
  const float n = e * (e + 2.f);
  // MUFU.RCP directly: n + 2 lies in [2, 2.4e17], so the range guards of __fdividef (FSETP + two predicated FMULs, a
  // quarter of the instructions of this function) protect nothing here
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n + 2.f));
  return n * r;
}
__device__ __forceinline__ float mish_f(float v) {
  float e;
  return v * mish_tanh_sp(v, &e);
}
// tanh on the SFU (MUFU.TANH, relative error ~2^-11): the reference's tanh output is a bf16 tensor under autocast
// (2^-9), so the approximation is below the rounding the reference itself applies
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// d mish / dv = t + v * (1 - t^2) * sigmoid(v),  t = tanh(softplus(v)),  sigmoid(v) = e / (1 + e)
__device__ __forceinline__ float mish_grad_f(float v) {
  float e;
  const float t = mish_tanh_sp(v, &e);
  float r1;                                     // 1 + e lies in [1, 4.9e8]: plain MUFU.RCP, as above
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(1.f + e));
  const float sg = e * r1;                      // == 1.0f for v >= 20 (e is clamped at exp(20))
  return t + v * (1.f - t * t) * sg;
}

// GroupNorm affine + Mish of one element in 7 instructions (FFMA, MUFU.EX2, FADD, FFMA, MUFU.RCP, FFMA, FMUL). The affine
// is applied in the log2 domain, a = x * sc2 + sh2 with sc2 = sc * log2(e), sh2 = sh * log2(e), so that a feeds the
// exponential directly and u = gn(x) = a * ln 2 never has to exist on its own:
//   e = 2^a = exp(u),  s = 1 + e,  tanh(softplus(u)) = (s^2 - 1) / (s^2 + 1) = 1 - 2 / (s^2 + 1)
//   mish(u) = a * (ln 2 - 2 ln 2 / (s^2 + 1))
// No clamp: e = inf gives 1 / inf = 0 and the linear branch u; for u << 0 the factor cancels to an absolute error of
// ~1e-7, i.e. |u| * 1e-7 on a value that is rounded to bf16 next.
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
__device__ __forceinline__ float gn_mish_elem(float x, float sc2, float sh2) {
  const float a = fmaf(x, sc2, sh2);
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
  const float s = e + 1.f;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(s, s, 1.f)));
  return a * fmaf(r, -2.f * kLn2, kLn2);
}
// ------------------------------------------------------------------------------------------- per-image channel sums
// stats[b][0][c] = sum_hw x, stats[b][1][c] = sum_hw x^2   (caller zeroes)
__global__ void __launch_bounds__(kVT) image_channel_stats_kernel(const bf16* __restrict__ x,
                                                                  float* __restrict__ stats, int HW, int C) {
  extern __shared__ float smem[];
  const int cg = C >> 3, lanes = kVT / cg;
  const int cgi = threadIdx.x % cg, lane_px = threadIdx.x / cg;
  const bool active = lane_px < lanes;
  const int b = blockIdx.y;
  float a1[8] = {}, a2[8] = {};
  if (active) {
    for (int p = blockIdx.x * lanes + lane_px; p < HW; p += gridDim.x * lanes) {
      float v[8];
      load8(x + ((size_t)b * HW + p) * C + cgi * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a1[j] += v[j];
        a2[j] += v[j] * v[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      smem[(lane_px * 2 + 0) * C + cgi * 8 + j] = a1[j];
      smem[(lane_px * 2 + 1) * C + cgi * 8 + j] = a2[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += kVT) {
    float s = 0.f;
    for (int l = 0; l < lanes; ++l) s += smem[l * 2 * C + i];
    atomicAdd(stats + (size_t)b * 2 * C + i, s);
  }
}

// group mean / rstd of channel c's group from the per-image channel sums
__device__ __forceinline__ void group_stats(const float* __restrict__ st, int C, int cpg, int c, float inv_m, float eps,
                                            float* mean, float* rstd) {
  const int g0 = (c / cpg) * cpg;
  float s1 = 0.f, s2 = 0.f;
  for (int k = 0; k < cpg; ++k) {
    s1 += st[g0 + k];
    s2 += st[C + g0 + k];
  }
  const float m = s1 * inv_m;
  float var = s2 * inv_m - m * m;
  var = var < 0.f ? 0.f : var;
  *mean = m;
  *rstd = rsqrtf(var + eps);
}


// Group statistics of image b computed ONCE per block into shared memory: sg[g] = {mean, rstd}; with red != null also
// the gamma-weighted backward sums sg[groups + g] = {S1/m, S2/m}. Thread i < groups reduces group i.
__device__ __forceinline__ void block_group_stats(const float* __restrict__ st, const float* __restrict__ red,
                                                  const float* __restrict__ gamma, int C, int groups, float inv_m,
                                                  float eps, float2* sg) {
  const int cpg = C / groups;
  if ((int)threadIdx.x < groups) {
    const int g0 = threadIdx.x * cpg;
    float s1 = 0.f, s2 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < cpg; ++k) {
      s1 += st[g0 + k];
      s2 += st[C + g0 + k];
      if (red) {
        t1 += gamma[g0 + k] * red[g0 + k];
        t2 += gamma[g0 + k] * red[C + g0 + k];
      }
    }
    const float m = s1 * inv_m;
    float var = s2 * inv_m - m * m;
    var = var < 0.f ? 0.f : var;
    sg[threadIdx.x] = make_float2(m, rsqrtf(var + eps));
    if (red) sg[groups + threadIdx.x] = make_float2(t1 * inv_m, t2 * inv_m);
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------- GroupNorm + Mish forward
// y = mish(gn(x));  if (res) y = mish(y + res);  if (add) y = y + add      (lunar_generate.py:37-38,53,96-97,212-222)
__global__ void __launch_bounds__(kVT) gn_mish_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ stats,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const bf16* __restrict__ res,
                                                          const bf16* __restrict__ add, bf16* __restrict__ y, int HW,
                                                          int C, int groups, float eps) {
  __shared__ float2 sg[64];
  const int cg = C >> 3, lanes = kVT / cg;
  const int cgi = threadIdx.x % cg, lane_px = threadIdx.x / cg;
  const int b = blockIdx.y, c0 = cgi * 8, cpg = C / groups;
  const float inv_m = 1.f / ((float)cpg * (float)HW);
  block_group_stats(stats + (size_t)b * 2 * C, nullptr, nullptr, C, groups, inv_m, eps, sg);
  if (lane_px >= lanes) return;
  float sc2[8], sh2[8];                              // GroupNorm affine in the log2 domain (gn_mish_elem)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 ms = sg[(c0 + j) / cpg];
    const float sc = gamma[c0 + j] * ms.y;
    sc2[j] = sc * kLog2e;
    sh2[j] = (beta[c0 + j] - ms.x * sc) * kLog2e;
  }
  for (int p = blockIdx.x * lanes + lane_px; p < HW; p += gridDim.x * lanes) {
    const size_t off = ((size_t)b * HW + p) * C + c0;
    float v[8];
    load8(x + off, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = gn_mish_elem(v[j], sc2[j], sh2[j]);
    if (res) {
      float r[8];
      load8(res + off, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = mish_f(v[j] + r[j]);
    }
    if (add) {
      float r[8];
      load8(add + off, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    store8(y + off, v);
  }
}

// ------------------------------------------------------------------------------------------- GroupNorm + Mish backward
// Recomputes the forward from x; dyh = gradient at the GroupNorm output. pass 0 (reduce): red[b][0][c] = sum dyh,
// red[b][1][c] = sum dyh*xhat. pass 1 (apply): dx = rstd*(gamma*dyh - S1/m - xhat*S2/m) with group sums S1,S2 of
// gamma-weighted red; when res != null also writes dres = dy * mish'(mish(gn)+res).
template <int PASS>
__global__ void __launch_bounds__(kVT) gn_mish_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ dy2,
                                                          const bf16* __restrict__ x,
                                                          const float* __restrict__ stats,
                                                          const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const bf16* __restrict__ res,
                                                          float* __restrict__ red, bf16* __restrict__ dx,
                                                          bf16* __restrict__ dres, int HW, int C, int groups,
                                                          float eps) {
  extern __shared__ float smem[];
  const int cg = C >> 3, lanes = kVT / cg;
  const int cgi = threadIdx.x % cg, lane_px = threadIdx.x / cg;
  const bool active = lane_px < lanes;
  const int b = blockIdx.y, c0 = cgi * 8, cpg = C / groups;
  const float inv_m = 1.f / ((float)cpg * (float)HW);
  __shared__ float2 sg[128];
  block_group_stats(stats + (size_t)b * 2 * C, PASS == 1 ? red + (size_t)b * 2 * C : nullptr, gamma, C, groups, inv_m,
                    eps, sg);
  float mean[8], rstd[8], gm[8], bt[8], s1[8], s2[8], a1[8] = {}, a2[8] = {};
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const float2 ms = sg[c / cpg];
      mean[j] = ms.x;
      rstd[j] = ms.y;
      gm[j] = gamma[c];
      bt[j] = beta[c];
      if (PASS == 1) {
        const float2 ts = sg[groups + c / cpg];
        s1[j] = ts.x;
        s2[j] = ts.y;
      }
    }
    for (int p = blockIdx.x * lanes + lane_px; p < HW; p += gridDim.x * lanes) {
      const size_t off = ((size_t)b * HW + p) * C + c0;
      float d[8], xv[8], r[8], dr[8];
      // every stream's 16-byte vector is requested before the first is unpacked (up to four loads in flight)
      const uint4 q_d = ldg16(dy + off), q_x = ldg16(x + off);
      uint4 q_d2 = make_uint4(0u, 0u, 0u, 0u), q_r = make_uint4(0u, 0u, 0u, 0u);
      if (dy2) q_d2 = ldg16(dy2 + off);
      if (res) q_r = ldg16(res + off);
      unpack8(q_d, d);
      if (dy2) {
        float d2[8];
        unpack8(q_d2, d2);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] += d2[j];
      }
      unpack8(q_x, xv);
      if (res) unpack8(q_r, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (xv[j] - mean[j]) * rstd[j];
        const float u = xh * gm[j] + bt[j];
        float g = d[j];
        if (res) {
          g *= mish_grad_f(mish_f(u) + r[j]);
          dr[j] = g;
        }
        g *= mish_grad_f(u);
        if (PASS == 0) {
          a1[j] += g;
          a2[j] += g * xh;
        } else {
          d[j] = rstd[j] * (gm[j] * g - s1[j] - xh * s2[j]);
        }
      }
      if (PASS == 1) {
        store8(dx + off, d);
        if (res && dres) store8(dres + off, dr);
      }
    }
  }
  if (PASS == 0) {
    if (active) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        smem[(lane_px * 2 + 0) * C + c0 + j] = a1[j];
        smem[(lane_px * 2 + 1) * C + c0 + j] = a2[j];
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kVT) {
      float s = 0.f;
      for (int l = 0; l < lanes; ++l) s += smem[l * 2 * C + i];
      atomicAdd(red + (size_t)b * 2 * C + i, s);
    }
  }
}

// ------------------------------------------------------------------------------------------- 3-channel input conv
// y[b,oh,ow,:] = act(conv3x3(x_nchw[b,0:3], w) + bias), stride 1 or 2, pad 1; NHWC bf16 output with COUT channels.
template <int COUT>
__global__ void __launch_bounds__(256) conv3x3_c3_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, bf16* __restrict__ y,
                                                             int B, int H, int W, int stride) {
  __shared__ float sw[27][COUT];
  __shared__ float sb[COUT];
  for (int i = threadIdx.x; i < 27 * COUT; i += 256) sw[i / COUT][i % COUT] = rbf(w[(i % COUT) * 27 + i / COUT]);
  for (int i = threadIdx.x; i < COUT; i += 256) sb[i] = bias[i];
  __syncthreads();
  const int OH = H / stride, OW = W / stride;
  const long total = (long)B * OH * OW;
  const long p = (long)blockIdx.x * 256 + threadIdx.x;
  if (p >= total) return;
  const int b = (int)(p / ((long)OH * OW)), r = (int)(p % ((long)OH * OW)), oh = r / OW, ow = r % OW;
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
  for (int c = 0; c < 3; ++c) {
    const float* xc = x + ((long)b * 3 + c) * H * W;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int ih = oh * stride + kh - 1;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int iw = ow * stride + kw - 1;
        float v = 0.f;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = rbf(__ldg(xc + (long)ih * W + iw));
        const float* wr = sw[(c * 3 + kh) * 3 + kw];
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] += v * wr[o];
      }
    }
  }
  bf16* dst = y + p * COUT;
#pragma unroll
  for (int j = 0; j < COUT / 8; ++j) {
    float v8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v8[i] = acc[8 * j + i] + sb[8 * j + i];
    store8(dst + 8 * j, v8);
  }
}

// dW[o][c][kh][kw] += sum_pixels dy[b,oh,ow,o] * x[b,c,ih,iw];  db[o] += sum dy.  256 pixels per block iteration.
template <int COUT>
__global__ void __launch_bounds__(256) conv3x3_c3_wgrad_kernel(const bf16* __restrict__ dy,
                                                               const float* __restrict__ x, float* __restrict__ dw,
                                                               float* __restrict__ db, int B, int H, int W,
                                                               int stride) {
  __shared__ float s_dy[64][COUT + 1];
  __shared__ float s_x[64][28];
  const int OH = H / stride, OW = W / stride;
  const long total = (long)B * OH * OW;
  // thread -> (o, tap group): 256 threads = COUT x (256/COUT) groups; 28 slots (27 taps + 1 for the bias) split up
  const int o = threadIdx.x % COUT, grp = threadIdx.x / COUT, ngrp = 256 / COUT;
  const int k_per = (28 + ngrp - 1) / ngrp;
  float acc[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) acc[k] = 0.f;
  for (long base = (long)blockIdx.x * 64; base < total; base += (long)gridDim.x * 64) {
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * COUT; i += 256) {
      const long p = base + i / COUT;
      s_dy[i / COUT][i % COUT] = p < total ? __bfloat162float(dy[p * COUT + i % COUT]) : 0.f;
    }
    for (int i = threadIdx.x; i < 64 * 28; i += 256) {
      const int pi = i / 28, k = i % 28;
      const long p = base + pi;
      float v = 0.f;
      if (p < total) {
        if (k == 27) v = 1.f;
        else {
          const int b = (int)(p / ((long)OH * OW)), r = (int)(p % ((long)OH * OW)), oh = r / OW, ow = r % OW;
          const int c = k / 9, kh = (k / 3) % 3, kw = k % 3;
          const int ih = oh * stride + kh - 1, iw = ow * stride + kw - 1;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = rbf(x[(((long)b * 3 + c) * H + ih) * W + iw]);
        }
      }
      s_x[pi][k] = v;
    }
    __syncthreads();
    for (int pi = 0; pi < 64; ++pi) {
      const float d = s_dy[pi][o];
#pragma unroll
      for (int kk = 0; kk < 28; ++kk) {
        if (kk < k_per) {
          const int k = grp * k_per + kk;
          if (k < 28) acc[kk] += d * s_x[pi][k];
        }
      }
    }
  }
#pragma unroll
  for (int kk = 0; kk < 28; ++kk) {
    if (kk < k_per) {
      const int k = grp * k_per + kk;
      if (k < 27) atomicAdd(dw + o * 27 + k, acc[kk]);
      else if (k == 27) atomicAdd(db + o, acc[kk]);
    }
  }
}

// ------------------------------------------------------------------------------------------- final conv 32->3 + tanh
// recon[b,o,h,w] = tanh(conv3x3(x_nhwc[b], w[o]) + bias[o]), NCHW fp32 output (lunar_generate.py:226-228).
// N = 3 outputs is far too thin for tcgen05 tiles, so each warp runs warp-level mma.sync m16n8k16 (bf16, fp32 acc):
// M = 16 consecutive pixels of one image row (2 M-tiles per warp), N = 8 (3 used), K = 9 taps x 32 channels.
// A fragments are 4-byte loads straight from the NHWC tensor (tap shift = pixel offset, zero outside the image),
// B fragments (weights) are built once per warp.
__global__ void __launch_bounds__(256) final_conv_tanh_fwd_kernel(const bf16* __restrict__ x,
                                                                  const float* __restrict__ w,
                                                                  const float* __restrict__ bias,
                                                                  float* __restrict__ recon, int B, int H, int W) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // B fragment of (tap, k-half): b0 = W[k = 2t, 2t+1][n = g], b1 = W[k = 2t+8, 2t+9][n = g]; n >= 3 is zero padding
  uint32_t bfrag[9][2][2];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int kh = 0; kh < 2; ++kh)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float w0 = 0.f, w1 = 0.f;
        if (g < 3) {
          const int ci = kh * 16 + r * 8 + 2 * t;
          w0 = w[(g * 32 + ci) * 9 + tap];          // reference layout [o][ci][kh][kw]
          w1 = w[(g * 32 + ci + 1) * 9 + tap];
        }
        __nv_bfloat162 p = __floats2bfloat162_rn(w0, w1);
        bfrag[tap][kh][r] = *reinterpret_cast<uint32_t*>(&p);
      }
  const float b_lo = (2 * t < 3) ? bias[2 * t] : 0.f, b_hi = (2 * t + 1 < 3) ? bias[2 * t + 1] : 0.f;
  const long hw = (long)H * W;
  const long ngroups = (long)B * H * (W / 32);      // one warp per 32 consecutive pixels of a row
  const long warp0 = ((long)blockIdx.x * 256 + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * 256) >> 5;
  for (long grp = warp0; grp < ngroups; grp += nwarps) {
    const int wseg = (int)(grp % (W / 32));
    const int h = (int)((grp / (W / 32)) % H);
    const int b = (int)(grp / ((long)(W / 32) * H));
    float d[2][4] = {};
#pragma unroll
    for (int kh3 = 0; kh3 < 3; ++kh3) {
      const int ih = h + kh3 - 1;
      if (ih < 0 || ih >= H) continue;
      const bf16* rowp = x + ((long)b * H + ih) * W * 32;
#pragma unroll
      for (int kw3 = 0; kw3 < 3; ++kw3) {
        const int tap = kh3 * 3 + kw3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int w_lo = wseg * 32 + mt * 16 + g + kw3 - 1, w_hi = w_lo + 8;   // pixels of fragment rows g, g+8
          const bool ok_lo = w_lo >= 0 && w_lo < W, ok_hi = w_hi >= 0 && w_hi < W;
          const uint32_t* plo = reinterpret_cast<const uint32_t*>(rowp + (long)w_lo * 32);
          const uint32_t* phi = reinterpret_cast<const uint32_t*>(rowp + (long)w_hi * 32);
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {
            uint32_t a[4];
            a[0] = ok_lo ? __ldg(plo + kh * 8 + t) : 0u;          // row g,   k = 2t, 2t+1   (+16*kh)
            a[1] = ok_hi ? __ldg(phi + kh * 8 + t) : 0u;          // row g+8
            a[2] = ok_lo ? __ldg(plo + kh * 8 + 4 + t) : 0u;      // row g,   k = 2t+8, 2t+9
            a[3] = ok_hi ? __ldg(phi + kh * 8 + 4 + t) : 0u;
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
                "{%0, %1, %2, %3};"
                : "+f"(d[mt][0]), "+f"(d[mt][1]), "+f"(d[mt][2]), "+f"(d[mt][3])
                : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(bfrag[tap][kh][0]), "r"(bfrag[tap][kh][1]));
          }
        }
      }
    }
    // C fragment: d[.][0], d[.][1] = (pixel g, outputs 2t, 2t+1); d[.][2], d[.][3] = (pixel g+8, same outputs)
    float* obase = recon + (long)b * 3 * hw + (long)h * W + wseg * 32;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int px = mt * 16 + g;
      if (2 * t < 3) {
        obase[(2 * t) * hw + px] = tanhf(rbf(d[mt][0] + b_lo));
        obase[(2 * t) * hw + px + 8] = tanhf(rbf(d[mt][2] + b_lo));
      }
      if (2 * t + 1 < 3) {
        obase[(2 * t + 1) * hw + px] = tanhf(rbf(d[mt][1] + b_hi));
        obase[(2 * t + 1) * hw + px + 8] = tanhf(rbf(d[mt][3] + b_hi));
      }
    }
  }
}

// Decoder tail in ONE pass (lunar_generate.py:187-189 up4's GroupNorm + Mish, then :226-228 final conv + tanh):
//   h = mish(gn(t))  (bf16, optionally written for the backward)   recon = tanh(conv3x3(h, w) + bias)   NCHW fp32
// A block owns a 32 x 16 pixel tile of one image: the 34 x 18 halo of the raw transposed-conv output t is normalised and
// activated on the way into shared memory (80-byte pixel pitch: conflict-free ldmatrix), then each warp runs
// mma.sync m16n8k16 over 4 rows x one sixteen-pixel group with K = 9 taps x 32 channels, N = 8 (3 used). The
// normalised 32-channel tensor (the largest of the decoder) is neither written nor re-read on the sampling path.
constexpr int kFtW = 32, kFtH = 16, kFtPitch = 80;
__global__ void __launch_bounds__(256, 3) gn_mish_final_conv_tanh_kernel(
    const bf16* __restrict__ t, const float* __restrict__ stats, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ w, const float* __restrict__ bias,
    bf16* __restrict__ h_out, float* __restrict__ recon, int H, int W, int groups, float eps) {
  constexpr int C = 32, HX = kFtW + 2, HY = kFtH + 2;
  extern __shared__ __align__(16) unsigned char tile[];            // [HY][HX] pixels of kFtPitch bytes
  __shared__ float2 sg[64];
  __shared__ __align__(16) __nv_bfloat16 s_w[9][3][C + 2];         // [tap][output][ci] (+2: outputs 0 / 2 on different banks)
  const int b = blockIdx.z, y0 = blockIdx.y * kFtH, x0 = blockIdx.x * kFtW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tq = lane & 3;
  const int cpg = C / groups, cpg_shift = __ffs(cpg) - 1;      // 32 % groups == 0: channels per group is a power of two
  // ---- phase 1: halo tile -> GroupNorm + Mish -> bf16 smem (zeros outside the image = the conv's padding).
  // A thread keeps its channel octet; halo pixel k of the thread = (threadIdx.x >> 2) + 64 k, whose (hy, hx) advance by
  // (1, 30) with one wrap, so neither the pixel coordinates nor the addresses need a division. ALL ten 16-byte loads of
  // the thread are issued before anything else: their DRAM latency overlaps the weight staging, the group statistics
  // and the parameter loads below (with three CTAs per SM there is little else to hide it behind).
  constexpr int kPix = HX * HY, kSteps = (kPix + 63) / 64;
  // lane -> (pixel lane & 7, channel octet lane >> 3): a quarter-warp's eight 16-byte shared-memory stores then hit eight
  // different bank groups at the 80-byte pixel pitch ((5 p + c) mod 8 is a permutation in p; with four octets of two
  // pixels per quarter-warp every store was a 2-way conflict), and the warp still reads 512 contiguous bytes
  const int c0 = ((threadIdx.x >> 3) & 3) * 8;
  const int hp0 = (threadIdx.x >> 5) * 8 + (threadIdx.x & 7);
  const bf16* img = t + (size_t)b * H * W * C + c0;
  uint4 raw[kSteps];
  unsigned in_mask = 0;
  {
    int hp = hp0, hy = hp >= HX ? 1 : 0, hx = hp - hy * HX;
#pragma unroll
    for (int k = 0; k < kSteps; ++k) {
      const int y = y0 + hy - 1, x = x0 + hx - 1;
      const bool in = hp < kPix && (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
      raw[k] = in ? ldg16(img + (y * W + x) * C) : make_uint4(0u, 0u, 0u, 0u);
      in_mask |= in ? 1u << k : 0u;
      hp += 64;
      hx += 64 - HX;
      hy += 1;
      if (hx >= HX) {
        hx -= HX;
        hy += 1;
      }
    }
  }
  for (int i = threadIdx.x; i < 3 * C * 9; i += 256) {             // reference layout [o][ci][kh][kw], read coalesced
    const int tap = i % 9, oc = i / 9;
    s_w[tap][oc / C][oc % C] = __float2bfloat16_rn(w[i]);
  }
  block_group_stats(stats + (size_t)b * 2 * C, nullptr, nullptr, C, groups, 1.f / ((float)cpg * (float)H * (float)W),
                    eps, sg);
  {
    float sc2[8], sh2[8];                            // GroupNorm affine in the log2 domain (gn_mish_elem)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 ms = sg[(c0 + j) >> cpg_shift];
      const float sc = gamma[c0 + j] * ms.y;
      sc2[j] = sc * kLog2e;
      sh2[j] = (beta[c0 + j] - ms.x * sc) * kLog2e;
    }
    unsigned char* sdst = tile + hp0 * kFtPitch + c0 * 2;
    int hy = hp0 >= HX ? 1 : 0, hx = hp0 - hy * HX;
#pragma unroll
    for (int k = 0; k < kSteps; ++k) {
      uint4 packed = make_uint4(0u, 0u, 0u, 0u);
      if (in_mask >> k & 1u) {
        const uint32_t xw[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a0 = gn_mish_elem(__uint_as_float(xw[j] << 16), sc2[2 * j], sh2[2 * j]);
          const float a1 = gn_mish_elem(__uint_as_float(xw[j] & 0xffff0000u), sc2[2 * j + 1], sh2[2 * j + 1]);
          const __nv_bfloat162 pb = __floats2bfloat162_rn(a0, a1);
          ow[j] = *reinterpret_cast<const uint32_t*>(&pb);
        }
        packed = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        if (h_out != nullptr && hy >= 1 && hy <= kFtH && hx >= 1 && hx <= kFtW)    // interior pixel: kept for the backward
          *reinterpret_cast<uint4*>(h_out + (size_t)b * H * W * C + c0 + ((y0 + hy - 1) * W + x0 + hx - 1) * C) = packed;
      }
      if (k < kSteps - 1 || hp0 + 64 * (kSteps - 1) < kPix) *reinterpret_cast<uint4*>(sdst + k * 64 * kFtPitch) = packed;
      hx += 64 - HX;
      hy += 1;
      if (hx >= HX) {
        hx -= HX;
        hy += 1;
      }
    }
  }
  __syncthreads();                                   // s_w and the halo tile are complete
  // B fragment of (tap, k-half): b0 = W[k = 2t, 2t+1][n = g], b1 = W[k = 2t+8, 2t+9][n = g]; n >= 3 is zero padding
  uint32_t bfrag[9][2][2];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int kh = 0; kh < 2; ++kh)
#pragma unroll
      for (int r = 0; r < 2; ++r)
        bfrag[tap][kh][r] = g < 3 ? *reinterpret_cast<const uint32_t*>(&s_w[tap][g][kh * 16 + r * 8 + 2 * tq]) : 0u;
  const float b_lo = (2 * tq < 3) ? bias[2 * tq] : 0.f, b_hi = (2 * tq + 1 < 3) ? bias[2 * tq + 1] : 0.f;
  // ---- phase 2: warp w -> 16-pixel column group cb = w & 1, output rows 4 (w >> 1) .. 4 (w >> 1) + 3. Every A fragment
  // (halo row, dx shift, channel half) is loaded ONCE and feeds the three output rows it reaches (dy = 0, 1, 2): 36
  // ldmatrix for 72 MMAs per warp instead of 72 (the kernel is bound by the shared-memory pipe: ncu, L1/TEX 76 %), and
  // the four rows are four independent accumulator chains instead of one chain of 18 dependent MMAs.
  const uint32_t tile_base = static_cast<uint32_t>(__cvta_generic_to_shared(tile));
  const int mat = lane >> 3;                                      // ldmatrix.x4: matrix fed by this lane's address
  const int lpix = (lane & 7) + ((mat & 1) ? 8 : 0), lch = (mat >> 1) ? 8 : 0;
  const long hw = (long)H * W;
  const int cb = warp & 1, r0 = (warp >> 1) * 4;
  float d[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
#pragma unroll
  for (int hr = 0; hr < 6; ++hr)
#pragma unroll
    for (int kw3 = 0; kw3 < 3; ++kw3) {
      const uint32_t pix_addr = tile_base + ((r0 + hr) * HX + cb * 16 + kw3 + lpix) * kFtPitch + lch * 2;
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        uint32_t a[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                     : "r"(pix_addr + kh * 32));
#pragma unroll
        for (int kh3 = 0; kh3 < 3; ++kh3) {
          const int orow = hr - kh3;                 // output row (within the band) whose tap (kh3, kw3) reads halo row hr
          if (orow < 0 || orow > 3) continue;
          asm volatile(
              "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
              "{%0, %1, %2, %3};"
              : "+f"(d[orow][0]), "+f"(d[orow][1]), "+f"(d[orow][2]), "+f"(d[orow][3])
              : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(bfrag[kh3 * 3 + kw3][kh][0]),
                "r"(bfrag[kh3 * 3 + kw3][kh][1]));
        }
      }
    }
  // C fragment: d[.][0], d[.][1] = (pixel g, outputs 2t, 2t+1); d[.][2], d[.][3] = (pixel g+8, same outputs)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* obase = recon + (long)b * 3 * hw + (long)(y0 + r0 + i) * W + x0 + cb * 16;
    if (2 * tq < 3) {
      obase[(2 * tq) * hw + g] = tanh_fast(rbf(d[i][0] + b_lo));
      obase[(2 * tq) * hw + g + 8] = tanh_fast(rbf(d[i][2] + b_lo));
    }
    if (2 * tq + 1 < 3) {
      obase[(2 * tq + 1) * hw + g] = tanh_fast(rbf(d[i][1] + b_hi));
      obase[(2 * tq + 1) * hw + g + 8] = tanh_fast(rbf(d[i][3] + b_hi));
    }
  }
}

// dx[b,h,w,ci] = sum_{o,taps} dpre[b,o,h+1-kh,w+1-kw] * w[o][ci][kh][kw],  dpre = drecon * (1 - recon^2)
__global__ void __launch_bounds__(256) final_conv_dgrad_kernel(const float* __restrict__ drecon,
                                                               const float* __restrict__ recon,
                                                               const float* __restrict__ w, bf16* __restrict__ dx,
                                                               int B, int H, int W) {
  __shared__ float sw[9][3][32];
  for (int i = threadIdx.x; i < 864; i += 256) {
    const int o = i / 288, ci = (i / 9) % 32, t = i % 9;
    sw[t][o][ci] = rbf(w[i]);
  }
  __syncthreads();
  const long total = (long)B * H * W, hw = (long)H * W;
  const long p = (long)blockIdx.x * 256 + threadIdx.x;
  if (p >= total) return;
  const int b = (int)(p / hw), r = (int)(p % hw), h = r / W, wq = r % W;
  float acc[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) acc[c] = 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) {
    const int oh = h + 1 - kh;
    if (oh < 0 || oh >= H) continue;
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int ow = wq + 1 - kw;
      if (ow < 0 || ow >= W) continue;
      const long q = (long)b * 3 * hw + (long)oh * W + ow;
      float d[3];
#pragma unroll
      for (int o = 0; o < 3; ++o) {
        const float rc = recon[q + o * hw];
        d[o] = rbf(drecon[q + o * hw] * (1.f - rc * rc));
      }
      const int t = kh * 3 + kw;
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] += d[0] * sw[t][0][c] + d[1] * sw[t][1][c] + d[2] * sw[t][2][c];
    }
  }
  bf16* dst = dx + p * 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v8[i] = acc[8 * j + i];
    store8(dst + 8 * j, v8);
  }
}

// dW[o][ci][kh][kw] += sum_pixels dpre[b,o,h,w] * x[b,h+kh-1,w+kw-1,ci];  db[o] += sum dpre.
// One block per 16x16 pixel tile; thread (tap, ci) for 288 threads, 3 more for the bias.
__global__ void __launch_bounds__(320) final_conv_wgrad_kernel(const float* __restrict__ drecon,
                                                               const float* __restrict__ recon,
                                                               const bf16* __restrict__ x, float* __restrict__ dw,
                                                               float* __restrict__ db, int H, int W) {
  __shared__ float s_x[18 * 18][33];
  __shared__ float s_d[256][3];
  const int b = blockIdx.z, h0 = blockIdx.y * 16, w0 = blockIdx.x * 16;
  const long hw = (long)H * W;
  for (int i = threadIdx.x; i < 18 * 18 * 4; i += 320) {
    const int c8 = i & 3, px = i >> 2, ty = px / 18, tx = px % 18;
    const int ih = h0 + ty - 1, iw = w0 + tx - 1;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) load8(x + (((long)b * H + ih) * W + iw) * 32 + c8 * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s_x[px][c8 * 8 + j] = v[j];
  }
  for (int i = threadIdx.x; i < 768; i += 320) {
    const int o = i / 256, px = i % 256;
    const long q = ((long)b * 3 + o) * hw + (long)(h0 + px / 16) * W + w0 + px % 16;
    const float rc = recon[q];
    s_d[px][o] = rbf(drecon[q] * (1.f - rc * rc));
  }
  __syncthreads();
  if (threadIdx.x < 288) {
    const int ci = threadIdx.x % 32, t = threadIdx.x / 32, kh = t / 3, kw = t % 3;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int px = 0; px < 256; ++px) {
      const float xv = s_x[(px / 16 + kh) * 18 + px % 16 + kw][ci];
      a0 += s_d[px][0] * xv;
      a1 += s_d[px][1] * xv;
      a2 += s_d[px][2] * xv;
    }
    atomicAdd(dw + (0 * 32 + ci) * 9 + t, a0);
    atomicAdd(dw + (1 * 32 + ci) * 9 + t, a1);
    atomicAdd(dw + (2 * 32 + ci) * 9 + t, a2);
  } else if (threadIdx.x < 291) {
    const int o = threadIdx.x - 288;
    float a = 0.f;
    for (int px = 0; px < 256; ++px) a += s_d[px][o];
    atomicAdd(db + o, a);
  }
}

// ------------------------------------------------------------------------------------------- reparameterisation
// z = mu + eps * exp(0.5*logvar)  (lunar_generate.py:259-261); mu/logvar come packed as mulv[B, 2L] (fused fc).
__global__ void reparam_fwd_kernel(const float* __restrict__ mulv, const float* __restrict__ eps,
                                   bf16* __restrict__ z, int B, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L) return;
  const int b = i / L, l = i % L;
  const float mu = mulv[(size_t)b * 2 * L + l], lv = mulv[(size_t)b * 2 * L + L + l];
  z[i] = __float2bfloat16_rn(mu + eps[i] * __expf(0.5f * lv));
}
// d(mulv) = [dmu + dz | dlogvar + dz * eps * 0.5 * exp(0.5*logvar)] as bf16 for the fused fc backward
__global__ void reparam_bwd_kernel(const float* __restrict__ mulv, const float* __restrict__ eps,
                                   const bf16* __restrict__ dz, const float* __restrict__ dmu,
                                   const float* __restrict__ dlv, bf16* __restrict__ dmulv, int B, int L) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * L) return;
  const int b = i / L, l = i % L;
  const float lv = mulv[(size_t)b * 2 * L + L + l];
  const float g = __bfloat162float(dz[i]);
  const float gm = (dmu ? dmu[i] : 0.f) + g;
  const float gl = (dlv ? dlv[i] : 0.f) + g * eps[i] * 0.5f * __expf(0.5f * lv);
  dmulv[(size_t)b * 2 * L + l] = __float2bfloat16_rn(gm);
  dmulv[(size_t)b * 2 * L + L + l] = __float2bfloat16_rn(gl);
}


// ------------------------------------------------------------------------------------------- step losses
// One pass over the reconstruction and the latent statistics (train_hybrid.py:859-862):
//   sums[0] += sum (recon - x)^2        sums[1] += sum (1 + logvar - mu^2 - exp(logvar))
__global__ void __launch_bounds__(256) vae_loss_fwd_kernel(const float* __restrict__ recon,
                                                           const float* __restrict__ x,
                                                           const float* __restrict__ mulv, float* __restrict__ sums,
                                                           long n_img4, int B, int L) {
  float a = 0.f, k = 0.f;
  const long stride = (long)gridDim.x * blockDim.x;
  const float4* r4 = reinterpret_cast<const float4*>(recon);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_img4; i += stride) {
    const float4 r = __ldg(r4 + i), t = __ldg(x4 + i);
    const float d0 = r.x - t.x, d1 = r.y - t.y, d2 = r.z - t.z, d3 = r.w - t.w;
    a += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)B * L; i += stride) {
    const int b = (int)(i / L), l = (int)(i % L);
    const float mu = mulv[(size_t)b * 2 * L + l], lv = mulv[(size_t)b * 2 * L + L + l];
    k += 1.f + lv - mu * mu - __expf(lv);
  }
  a = warp_sum(a);
  k = warp_sum(k);
  __shared__ float sa[8], sk[8];
  if ((threadIdx.x & 31) == 0) {
    sa[threadIdx.x >> 5] = a;
    sk[threadIdx.x >> 5] = k;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.f, tk = 0.f;
    for (int i = 0; i < 8; ++i) {
      ta += sa[i];
      tk += sk[i];
    }
    atomicAdd(sums, ta);
    atomicAdd(sums + 1, tk);
  }
}
// d_recon = g[0] * 2 (recon - x) / n_img;  d_mu = g[1] * mu / (B L);  d_logvar = g[1] * (-0.5)(1 - exp(logvar)) / (B L)
// where g[0], g[1] (device scalars) are the upstream gradients of recon_loss and kl_loss.
__global__ void __launch_bounds__(256) vae_loss_bwd_kernel(const float* __restrict__ recon,
                                                           const float* __restrict__ x,
                                                           const float* __restrict__ mulv,
                                                           const float* __restrict__ g, float* __restrict__ drecon,
                                                           float* __restrict__ dmulv, long n_img4, int B, int L) {
  const float gr = g[0] * 2.f / (float)(n_img4 * 4), gk = g[1] / ((float)B * (float)L);
  const long stride = (long)gridDim.x * blockDim.x;
  const float4* r4 = reinterpret_cast<const float4*>(recon);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4* d4 = reinterpret_cast<float4*>(drecon);
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_img4; i += stride) {
    const float4 r = __ldg(r4 + i), t = __ldg(x4 + i);
    d4[i] = make_float4(gr * (r.x - t.x), gr * (r.y - t.y), gr * (r.z - t.z), gr * (r.w - t.w));
  }
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < (long)B * L; i += stride) {
    const int b = (int)(i / L), l = (int)(i % L);
    const float mu = mulv[(size_t)b * 2 * L + l], lv = mulv[(size_t)b * 2 * L + L + l];
    dmulv[(size_t)b * 2 * L + l] = gk * mu;
    dmulv[(size_t)b * 2 * L + L + l] = gk * (-0.5f) * (1.f - __expf(lv));
  }
}

// ------------------------------------------------------------------------------------------- sprite loader
// uint8 NHWC sprites [B,H,W,3] (the on-disk format of sprites_*.npy) -> fp32 NCHW in [-1,1]  (x/127.5 - 1,
// train_hybrid.py:181-182)
__global__ void __launch_bounds__(256) sprites_u8_to_f32_kernel(const unsigned char* __restrict__ u8,
                                                                float* __restrict__ out, long n_pix, long hw) {
  const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pix) return;
  const long b = p / hw, r = p % hw;
  const unsigned char* s = u8 + p * 3;
  float* d = out + b * 3 * hw + r;
  d[0] = (float)s[0] / 127.5f - 1.f;
  d[hw] = (float)s[1] / 127.5f - 1.f;
  d[2 * hw] = (float)s[2] / 127.5f - 1.f;
}

static int vae_blocks(int HW, int C, int B) {
  const int lanes = kVT / (C / 8);
  int per = (HW + lanes - 1) / lanes;
  int want = (148 * 4 + B - 1) / B;
  if (want < 1) want = 1;
  return per < want ? per : want;
}

}  // namespace lun

using namespace lun;

#define LUN_LAUNCH_OK() (cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH)

extern "C" {

int lun_image_channel_stats_bf16(const void* x, float* stats, int B, int HW, int C, void* stream) {
  if (C % 8 || C / 8 > kVT) return LUN_E_SHAPE;
  const int lanes = kVT / (C / 8);
  dim3 grid(vae_blocks(HW, C, B), B);
  image_channel_stats_kernel<<<grid, kVT, lanes * 2 * C * sizeof(float), (cudaStream_t)stream>>>((const bf16*)x,
                                                                                                 stats, HW, C);
  lun::note_launch(1);
  return LUN_LAUNCH_OK();
}

int lun_gn_mish_fwd_bf16(const void* x, const float* stats, const float* gamma, const float* beta, const void* res,
                         const void* add, void* y, int B, int HW, int C, int groups, float eps, void* stream) {
  if (C % 8 || C / 8 > kVT || C % groups) return LUN_E_SHAPE;
  // 40 registers: 6 blocks are resident per SM; ~12 blocks per SM in all keep them resident to the end (the 4 per SM of
  // vae_blocks left the pass at 57 % of its warp slots: ncu, profiles/r02_ncu_c5_decoder_kernels.txt)
  const int lanes = kVT / (C / 8);
  int per = (HW + lanes - 1) / lanes, want = (148 * 12 + B - 1) / B;
  dim3 grid(per < want ? per : want, B);
  gn_mish_fwd_kernel<<<grid, kVT, 0, (cudaStream_t)stream>>>((const bf16*)x, stats, gamma, beta, (const bf16*)res,
                                                             (const bf16*)add, (bf16*)y, HW, C, groups, eps);
  lun::note_launch(1);
  return LUN_LAUNCH_OK();
}

int lun_gn_mish_bwd_bf16(const void* dy, const void* dy2, const void* x, const float* stats, const float* gamma, const float* beta,
                         const void* res, float* red, void* dx, void* dres, int B, int HW, int C, int groups,
                         float eps, void* stream) {
  if (C % 8 || C / 8 > kVT || C % groups) return LUN_E_SHAPE;
  const int lanes = kVT / (C / 8);
  dim3 grid(vae_blocks(HW, C, B), B);
  gn_mish_bwd_kernel<0><<<grid, kVT, lanes * 2 * C * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)dy, (const bf16*)dy2, (const bf16*)x, stats, gamma, beta, (const bf16*)res, red, nullptr, nullptr,
      HW, C, groups, eps);
  gn_mish_bwd_kernel<1><<<grid, kVT, 0, (cudaStream_t)stream>>>((const bf16*)dy, (const bf16*)dy2, (const bf16*)x,
                                                                stats, gamma, beta, (const bf16*)res, red, (bf16*)dx,
                                                                (bf16*)dres, HW, C, groups, eps);
  lun::note_launch(2);
  return LUN_LAUNCH_OK();
}

int lun_conv3x3_c3_fwd(const float* x_nchw, const float* w, const float* bias, void* y, int B, int H, int W, int cout,
                       int stride, void* stream) {
  if (cout != 64 || (stride != 1 && stride != 2)) return LUN_E_SHAPE;
  const long total = (long)B * (H / stride) * (W / stride);
  conv3x3_c3_fwd_kernel<64><<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_nchw, w, bias, (bf16*)y, B,
                                                                                         H, W, stride);
  lun::note_launch(1);
  return LUN_LAUNCH_OK();
}

int lun_conv3x3_c3_wgrad(const void* dy, const float* x_nchw, float* dw, float* db, int B, int H, int W, int cout,
                         int stride, void* stream) {
  if (cout != 64) return LUN_E_SHAPE;
  conv3x3_c3_wgrad_kernel<64><<<148 * 2, 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, x_nchw, dw, db, B, H, W,
                                                                        stride);
  lun::note_launch(1);
  return LUN_LAUNCH_OK();
}

int lun_final_conv_tanh_fwd(const void* x, const float* w, const float* bias, float* recon, int B, int H, int W,
                            void* stream) {
  if (W % 32) return LUN_E_SHAPE;
  const long groups = (long)B * H * (W / 32);
  long blocks = (groups + 7) / 8;
  if (blocks > 148 * 16) blocks = 148 * 16;
  final_conv_tanh_fwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, w, bias, recon, B, H, W);
  lun::note_launch(1);
  return LUN_LAUNCH_OK();
}

int lun_gn_mish_final_conv_tanh_fwd(const void* t, const float* stats, const float* gamma, const float* beta,
                                    const float* w, const float* bias, void* h_out, float* recon, int B, int H, int W,
                                    int groups, float eps, void* stream) {
  if (W % kFtW || H % kFtH || 32 % groups) return LUN_E_SHAPE;
  dim3 grid(W / kFtW, H / kFtH, B);
  const int smem = (kFtW + 2) * (kFtH + 2) * kFtPitch;
  static bool configured_dev[64] = {};   // the opt-in smem size is a per-device function attribute
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(gn_mish_final_conv_tanh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) !=
        cudaSuccess)
      return LUN_E_ATTR;
    configured = true;
  }
  gn_mish_final_conv_tanh_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)t, stats, gamma, beta, w, bias,
                                                                         (bf16*)h_out, recon, H, W, groups, eps);
  lun::note_launch(1);
  return LUN_LAUNCH_OK();
}

int lun_final_conv_bwd(const float* drecon, const float* recon, const void* x, const float* w, void* dx, float* dw,
                       float* db, int B, int H, int W, void* stream) {
  if (H % 16 || W % 16) return LUN_E_SHAPE;
  const long total = (long)B * H * W;
  final_conv_dgrad_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(drecon, recon, w, (bf16*)dx, B,
                                                                                       H, W);
  dim3 grid(W / 16, H / 16, B);
  final_conv_wgrad_kernel<<<grid, 320, 0, (cudaStream_t)stream>>>(drecon, recon, (const bf16*)x, dw, db, H, W);
  lun::note_launch(2);
  return LUN_LAUNCH_OK();
}

int lun_reparam_fwd(const float* mulv, const float* eps, void* z, int B, int L, void* stream) {
  reparam_fwd_kernel<<<(B * L + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mulv, eps, (bf16*)z, B, L);
  lun::note_launch(1);
  return LUN_LAUNCH_OK();
}

int lun_reparam_bwd(const float* mulv, const float* eps, const void* dz, const float* dmu, const float* dlogvar,
                    void* dmulv, int B, int L, void* stream) {
  reparam_bwd_kernel<<<(B * L + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mulv, eps, (const bf16*)dz, dmu, dlogvar,
                                                                            (bf16*)dmulv, B, L);
  lun::note_launch(1);
  return LUN_LAUNCH_OK();
}

int lun_vae_loss_fwd(const float* recon, const float* images, const float* mulv, float* sums, long n_img, int B, int L,
                     void* stream) {
  if (n_img % 4) return LUN_E_SHAPE;
  vae_loss_fwd_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(recon, images, mulv, sums, n_img / 4, B, L);
  lun::note_launch(1);
  return LUN_LAUNCH_OK();
}

int lun_vae_loss_bwd(const float* recon, const float* images, const float* mulv, const float* g, float* drecon,
                     float* dmulv, long n_img, int B, int L, void* stream) {
  if (n_img % 4) return LUN_E_SHAPE;
  vae_loss_bwd_kernel<<<296, 256, 0, (cudaStream_t)stream>>>(recon, images, mulv, g, drecon, dmulv, n_img / 4, B, L);
  lun::note_launch(1);
  return LUN_LAUNCH_OK();
}

int lun_sprites_u8_to_f32(const void* u8_nhwc, float* out_nchw, int B, int H, int W, void* stream) {
  const long n = (long)B * H * W;
  sprites_u8_to_f32_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const unsigned char*)u8_nhwc,
                                                                                    out_nchw, n, (long)H * W);
  lun::note_launch(1);
  return LUN_LAUNCH_OK();
}

}  // extern "C"
