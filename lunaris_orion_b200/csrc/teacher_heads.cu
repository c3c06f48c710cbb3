// Teacher heads (reference lunar_evaluator.py:353-397 definitions, :417-456 use): gate, four quality heads, semantic
// head, style / prompt nets - every one AdaptiveAvgPool -> [LayerNorm] -> Linear -> LeakyReLU(0.2) -> Dropout -> Linear on
// a pooled [B, C] feature - plus the mixing arithmetic between them (softmax gate, expert-weighted quality logits and
// features, sigmoids, the cosine factor of :452-453). The reference spends ~100 ATen / cuBLAS launches per forward on
// these 25 MFLOP; here one launch does the forward (one CTA per sample, fp32), one launch the per-sample backward and
// one launch the batch reductions of the weight gradients.
//
// The pooled inputs arrive as per-image channel SUMS from the trunk's last fused pass (lun_affine_fwd_bf16 `pool`);
// the 1 / (H W) of AdaptiveAvgPool2d is applied here. Dropout uses the stateless counter RNG of elem_common.cuh,
// element index = sample * hidden + unit, one seed per head.
#include "../../include/lunaris_b200.h"
#include "elem_common.cuh"
#include "launch_count.cuh"

namespace lun {

constexpr int kHeadThreads = 256;
constexpr int kMaxExperts = 8;
constexpr float kLnEps = 1e-5f;

struct Mlp {                       // one head; ln_w == nullptr: no LayerNorm (gate)
  const float *ln_w, *ln_b, *w1, *b1, *w2, *b2;
  int in, hid, out;
  unsigned long long seed;
};
struct MlpGrad {                   // parameter gradients of one head (any pointer may be null: not wanted)
  float *ln_w, *ln_b, *w1, *b1, *w2, *b2;
};
struct HeadsDesc {
  Mlp gate, quality[kMaxExperts], semantic, style, prompt;
  int B, E, F, C, emb;
  float inv_hw, slope, drop_scale;
  unsigned int thresh16;
};
struct HeadsGrads {
  MlpGrad gate, quality[kMaxExperts], semantic, style, prompt;
};

// ---------------------------------------------------------------------------------------------- device helpers
__device__ __forceinline__ float block_sum(float v, float* red) {   // all threads get the sum; red: 8 floats of smem
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int i = 0; i < kHeadThreads / 32; ++i) t += red[i];
  return t;
}
// ys[r] = act(W[r,:] . xs + b[r]), r < R: one warp per row, lanes stride over K (K % 4 == 0, rows 16-byte aligned)
__device__ __forceinline__ void gemv_rows(const float* __restrict__ W, const float* __restrict__ b, const float* xs,
                                          float* ys, int R, int K, bool leaky, float slope) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < R; r += kHeadThreads / 32) {
    const float4* w4 = reinterpret_cast<const float4*>(W + (size_t)r * K);
    float acc = 0.f;
    for (int k = lane; k < K / 4; k += 32) {
      const float4 w = __ldg(w4 + k);
      const float4 x = *reinterpret_cast<const float4*>(xs + 4 * k);
      acc += w.x * x.x + w.y * x.y + w.z * x.z + w.w * x.w;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += b[r];
      ys[r] = leaky ? (acc > 0.f ? acc : acc * slope) : acc;
    }
  }
}
// ys[k] = sum_r W[r,k] * ds[r], k < K: one thread per column (coalesced over k)
__device__ __forceinline__ void gemv_cols(const float* __restrict__ W, const float* ds, float* ys, int R, int K) {
  for (int k = threadIdx.x; k < K; k += kHeadThreads) {
    float acc = 0.f;
    for (int r = 0; r < R; ++r) acc = fmaf(__ldg(W + (size_t)r * K + k), ds[r], acc);
    ys[k] = acc;
  }
}
__device__ __forceinline__ bool head_keep(const HeadsDesc& d, unsigned long long seed, size_t idx) {
  return d.thresh16 == 0 || drop_keep1(seed, idx, d.thresh16);
}
// LayerNorm statistics of xs[0:C] -> xhat[0:C] (smem) and rstd; `red` is 8 floats of scratch
__device__ __forceinline__ float layer_norm_hat(const float* xs, float* xhat, int C, float* red) {
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += kHeadThreads) s += xs[c];
  const float mean = block_sum(s, red) / C;
  float v = 0.f;
  for (int c = threadIdx.x; c < C; c += kHeadThreads) {
    const float t = xs[c] - mean;
    v += t * t;
  }
  const float rstd = rsqrtf(block_sum(v, red) / C + kLnEps);
  for (int c = threadIdx.x; c < C; c += kHeadThreads) xhat[c] = (xs[c] - mean) * rstd;
  __syncthreads();
  return rstd;
}
// Linear -> LeakyReLU -> Dropout -> Linear of head m on xin (smem, [m.in]); hs: smem [m.hid]; outs: smem [m.out].
// hid_save (global, [m.hid]): post-activation, pre-dropout hidden units of this sample (for the backward).
__device__ __forceinline__ void mlp_forward(const HeadsDesc& d, const Mlp& m, const float* xin, float* hs, float* outs,
                                            float* hid_save, int b, bool training) {
  gemv_rows(m.w1, m.b1, xin, hs, m.hid, m.in, true, d.slope);
  __syncthreads();
  for (int j = threadIdx.x; j < m.hid; j += kHeadThreads) {
    const float a = hs[j];
    if (hid_save) hid_save[j] = a;
    if (training) hs[j] = head_keep(d, m.seed, (size_t)b * m.hid + j) ? a * d.drop_scale : 0.f;
  }
  __syncthreads();
  gemv_rows(m.w2, m.b2, hs, outs, m.out, m.hid, false, 0.f);
  __syncthreads();
}
// xn = xhat * gamma + beta
__device__ __forceinline__ void ln_affine(const Mlp& m, const float* xhat, float* xn, int C) {
  for (int c = threadIdx.x; c < C; c += kHeadThreads) xn[c] = xhat[c] * m.ln_w[c] + m.ln_b[c];
  __syncthreads();
}

// Saved-for-backward layout per sample (floats): see heads_save_floats()
struct SaveOff {
  int hid_gate, hid_q, hid_sem, hid_style, hid_prompt, xhat_e, xhat_c, rstd, q_e, wq_sig, sem_sig, cosv, total;
};
__host__ __device__ inline SaveOff save_offsets(const HeadsDesc& d) {
  SaveOff o;
  int p = 0;
  o.hid_gate = p; p += d.gate.hid;
  o.hid_q = p; p += d.E * d.quality[0].hid;
  o.hid_sem = p; p += d.semantic.hid;
  o.hid_style = p; p += d.style.hid;
  o.hid_prompt = p; p += d.prompt.hid;
  o.xhat_e = p; p += d.E * d.C;
  o.xhat_c = p; p += d.C;
  o.rstd = p; p += d.E + 1;
  o.q_e = p; p += d.E * 4;
  o.wq_sig = p; p += 4;
  o.sem_sig = p; p += 1;
  o.cosv = p; p += 1;
  o.total = (p + 3) & ~3;
  return o;
}

// ---------------------------------------------------------------------------------------------- forward
// pooled_fe: [B, F] sums; pooled: [E][B][C] sums (one tensor). Outputs: quality [B,4] (sigmoid),
// weights [B,E] (softmax), style / prompt [B,emb], semantic [B,1]. save: [B, save_offsets().total] or null (eval).
__global__ void __launch_bounds__(kHeadThreads) heads_fwd_kernel(const HeadsDesc d, const float* __restrict__ pooled_fe,
                                                                 const float* __restrict__ pooled,
                                                                 float* __restrict__ quality, float* __restrict__ weights,
                                                                 float* __restrict__ style, float* __restrict__ prompt,
                                                                 float* __restrict__ semantic, float* __restrict__ save,
                                                                 int training) {
  extern __shared__ __align__(16) float sh[];
  const int b = blockIdx.x, C = d.C, E = d.E;
  float* sfe = sh;                       // [F]
  float* sm = sfe + d.F;                 // [E][C] pooled means
  float* xhat = sm + E * C;              // [C]
  float* xn = xhat + C;                  // [C]
  float* comb = xn + C;                  // [C]
  float* hs = comb + C;                  // [max hidden]
  const int maxh = max(max(d.gate.hid, d.quality[0].hid), max(d.semantic.hid, max(d.style.hid, d.prompt.hid)));
  float* outs = hs + maxh;               // [max(out)]
  const int maxo = max(max(E, 4), d.emb);
  float* small = outs + maxo;            // w[E] | q[E][4] | red[8] | misc[4]
  float* w = small;
  float* q = w + kMaxExperts;
  float* red = q + kMaxExperts * 4;
  const SaveOff so = save_offsets(d);
  float* sv = save ? save + (size_t)b * so.total : nullptr;

  for (int i = threadIdx.x; i < d.F; i += kHeadThreads) sfe[i] = pooled_fe[(size_t)b * d.F + i] * d.inv_hw;
  for (int e = 0; e < E; ++e)
    for (int c = threadIdx.x; c < C; c += kHeadThreads)
      sm[e * C + c] = pooled[((size_t)e * d.B + b) * C + c] * d.inv_hw;
  __syncthreads();

  // gate (no LayerNorm) -> softmax
  mlp_forward(d, d.gate, sfe, hs, outs, sv ? sv + so.hid_gate : nullptr, b, training);
  if (threadIdx.x == 0) {
    float mx = outs[0];
    for (int e = 1; e < E; ++e) mx = fmaxf(mx, outs[e]);
    float s = 0.f;
    for (int e = 0; e < E; ++e) {
      w[e] = expf(outs[e] - mx);
      s += w[e];
    }
    for (int e = 0; e < E; ++e) {
      w[e] /= s;
      weights[(size_t)b * E + e] = w[e];
    }
  }
  __syncthreads();

  // quality heads (and the semantic head on expert 0's normalised features)
  for (int e = 0; e < E; ++e) {
    const float rstd = layer_norm_hat(sm + e * C, xhat, C, red);
    if (sv) {
      for (int c = threadIdx.x; c < C; c += kHeadThreads) sv[so.xhat_e + e * C + c] = xhat[c];
      if (threadIdx.x == 0) sv[so.rstd + e] = rstd;
    }
    ln_affine(d.quality[e], xhat, xn, C);
    mlp_forward(d, d.quality[e], xn, hs, outs, sv ? sv + so.hid_q + e * d.quality[0].hid : nullptr, b, training);
    if (threadIdx.x < 4) {
      q[e * 4 + threadIdx.x] = outs[threadIdx.x];
      if (sv) sv[so.q_e + e * 4 + threadIdx.x] = outs[threadIdx.x];
    }
    __syncthreads();
    if (e == 0) {
      ln_affine(d.semantic, xhat, xn, C);
      mlp_forward(d, d.semantic, xn, hs, outs, sv ? sv + so.hid_sem : nullptr, b, training);
      if (threadIdx.x == 0) red[8] = 1.f / (1.f + expf(-outs[0]));       // sigmoid(semantic logit), finished below
      __syncthreads();
    }
  }
  if (threadIdx.x < 4) {
    float a = 0.f;
    for (int e = 0; e < E; ++e) a += q[e * 4 + threadIdx.x] * w[e];
    const float s = 1.f / (1.f + expf(-a));
    quality[(size_t)b * 4 + threadIdx.x] = s;
    if (sv) sv[so.wq_sig + threadIdx.x] = s;
  }
  // expert-weighted pooled features -> style / prompt nets
  for (int c = threadIdx.x; c < C; c += kHeadThreads) {
    float a = 0.f;
    for (int e = 0; e < E; ++e) a += sm[e * C + c] * w[e];
    comb[c] = a;
  }
  __syncthreads();
  const float rstd_c = layer_norm_hat(comb, xhat, C, red);
  if (sv) {
    for (int c = threadIdx.x; c < C; c += kHeadThreads) sv[so.xhat_c + c] = xhat[c];
    if (threadIdx.x == 0) sv[so.rstd + E] = rstd_c;
  }
  ln_affine(d.style, xhat, xn, C);
  mlp_forward(d, d.style, xn, hs, outs, sv ? sv + so.hid_style : nullptr, b, training);
  for (int i = threadIdx.x; i < d.emb; i += kHeadThreads) style[(size_t)b * d.emb + i] = outs[i];
  __syncthreads();
  ln_affine(d.prompt, xhat, xn, C);
  mlp_forward(d, d.prompt, xn, hs, outs, sv ? sv + so.hid_prompt : nullptr, b, training);
  float ss = 0.f;
  for (int i = threadIdx.x; i < d.emb; i += kHeadThreads) {
    prompt[(size_t)b * d.emb + i] = outs[i];
    ss += outs[i] * outs[i];
  }
  // semantic = sigmoid(logit) * cosine_similarity(prompt, prompt.detach()) (lunar_evaluator.py:452-453): the cosine of a
  // vector with itself, x.x / sqrt(max(|x|^2 |x|^2, eps^2)) with torch's eps 1e-8 - 1 up to rounding, zero gradient
  ss = block_sum(ss, red);
  if (threadIdx.x == 0) {
    const float cosv = ss / sqrtf(fmaxf(ss * ss, 1e-16f));
    semantic[b] = red[8] * cosv;
    if (sv) {
      sv[so.sem_sig] = red[8];
      sv[so.cosv] = cosv;
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward, per sample
// Work buffers (global, per sample): dpre [B, sum of hidden sizes] (gradient at the first Linear's output, dropout and
// LeakyReLU already applied), dout2 [B, E + 4E + 1 + 2 emb] (gradient at the second Linear's output), dxn [E + 3][B, C]
// (gradient at the LayerNorm output of quality_e, semantic, style, prompt). d_pooled: [E][B, C] (gradient of the pooled
// SUMS). Any incoming gradient pointer may be null.
struct BwdOff {
  int pre_gate, pre_q, pre_sem, pre_style, pre_prompt, pre_total;
  int o_gate, o_q, o_sem, o_style, o_prompt, o_total;
};
__host__ __device__ inline BwdOff bwd_offsets(const HeadsDesc& d) {
  BwdOff o;
  int p = 0;
  o.pre_gate = p; p += d.gate.hid;
  o.pre_q = p; p += d.E * d.quality[0].hid;
  o.pre_sem = p; p += d.semantic.hid;
  o.pre_style = p; p += d.style.hid;
  o.pre_prompt = p; p += d.prompt.hid;
  o.pre_total = p;
  p = 0;
  o.o_gate = p; p += d.E;
  o.o_q = p; p += 4 * d.E;
  o.o_sem = p; p += 1;
  o.o_style = p; p += d.emb;
  o.o_prompt = p; p += d.emb;
  o.o_total = p;
  return o;
}
// dout (smem, [m.out]) -> dpre (global [m.hid], also left in hs) -> dxin (smem, [m.in]) when want_dx
__device__ __forceinline__ void mlp_backward(const HeadsDesc& d, const Mlp& m, const float* dout, const float* hid_saved,
                                             float* hs, float* dpre_g, float* dxin, int b, bool want_dx) {
  gemv_cols(m.w2, dout, hs, m.out, m.hid);
  __syncthreads();
  for (int j = threadIdx.x; j < m.hid; j += kHeadThreads) {
    const float a = hid_saved[j];
    float g = head_keep(d, m.seed, (size_t)b * m.hid + j) ? hs[j] * d.drop_scale : 0.f;
    g *= a > 0.f ? 1.f : d.slope;
    hs[j] = g;
    dpre_g[j] = g;
  }
  __syncthreads();
  if (want_dx) {
    gemv_cols(m.w1, hs, dxin, m.hid, m.in);
    __syncthreads();
  }
}
// LayerNorm backward: dxn (smem) -> adds rstd * (dxh - mean(dxh) - xhat * mean(dxh * xhat)) * scale to dst[0:C]
__device__ __forceinline__ void ln_backward_add(const Mlp& m, const float* dxn, const float* xhat_g, float rstd, float* dst,
                                                float scale, int C, float* red) {
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < C; c += kHeadThreads) {
    const float g = dxn[c] * m.ln_w[c];
    s1 += g;
    s2 += g * xhat_g[c];
  }
  s1 = block_sum(s1, red) / C;
  s2 = block_sum(s2, red) / C;
  for (int c = threadIdx.x; c < C; c += kHeadThreads)
    dst[c] += rstd * (dxn[c] * m.ln_w[c] - s1 - xhat_g[c] * s2) * scale;
  __syncthreads();
}

__global__ void __launch_bounds__(kHeadThreads) heads_bwd_sample_kernel(
    const HeadsDesc d, const float* __restrict__ pooled, const float* __restrict__ weights,
    const float* __restrict__ save, const float* __restrict__ g_quality, const float* __restrict__ g_weights,
    const float* __restrict__ g_style, const float* __restrict__ g_prompt, const float* __restrict__ g_semantic,
    float* __restrict__ dpre, float* __restrict__ dout2, float* __restrict__ dxn_all, float* __restrict__ d_pooled) {
  extern __shared__ __align__(16) float sh[];
  const int b = blockIdx.x, C = d.C, E = d.E, B = d.B;
  float* dsm = sh;                       // [E][C] gradient of the pooled means
  float* dxn = dsm + E * C;              // [C]
  float* dcomb = dxn + C;                // [C]
  float* hs = dcomb + C;                 // [max hidden]
  const int maxh = max(max(d.gate.hid, d.quality[0].hid), max(d.semantic.hid, max(d.style.hid, d.prompt.hid)));
  float* dout = hs + maxh;               // [max out]
  const int maxo = max(max(E, 4), d.emb);
  float* small = dout + maxo;
  float* dw = small;                     // [E] gradient of the softmax gate weights
  float* dwq = dw + kMaxExperts;         // [4]
  float* red = dwq + 4;                  // [8]
  const SaveOff so = save_offsets(d);
  const BwdOff bo = bwd_offsets(d);
  const float* sv = save + (size_t)b * so.total;
  float* pre = dpre + (size_t)b * bo.pre_total;
  float* o2 = dout2 + (size_t)b * bo.o_total;
  const float* w = weights + (size_t)b * E;

  for (int i = threadIdx.x; i < E * C; i += kHeadThreads) dsm[i] = 0.f;
  for (int c = threadIdx.x; c < C; c += kHeadThreads) dcomb[c] = 0.f;
  if (threadIdx.x < E) dw[threadIdx.x] = g_weights ? g_weights[(size_t)b * E + threadIdx.x] : 0.f;
  if (threadIdx.x < 4) {
    const float s = sv[so.wq_sig + threadIdx.x];
    dwq[threadIdx.x] = g_quality ? g_quality[(size_t)b * 4 + threadIdx.x] * s * (1.f - s) : 0.f;
  }
  __syncthreads();

  // style / prompt nets -> gradient of the LayerNorm input `comb`
  const bool has_sp = g_style || g_prompt;
  for (int which = 0; which < 2; ++which) {
    const float* gin = which == 0 ? g_style : g_prompt;
    const Mlp& m = which == 0 ? d.style : d.prompt;
    float* o2g = o2 + (which == 0 ? bo.o_style : bo.o_prompt);
    float* preg = pre + (which == 0 ? bo.pre_style : bo.pre_prompt);
    float* dxg = dxn_all + ((size_t)(E + 1 + which) * B + b) * C;
    if (!gin) {
      if (has_sp || true) {            // buffers are consumed by the reduction kernel: keep them defined
        for (int i = threadIdx.x; i < d.emb; i += kHeadThreads) o2g[i] = 0.f;
        for (int j = threadIdx.x; j < m.hid; j += kHeadThreads) preg[j] = 0.f;
        for (int c = threadIdx.x; c < C; c += kHeadThreads) dxg[c] = 0.f;
      }
      continue;
    }
    for (int i = threadIdx.x; i < d.emb; i += kHeadThreads) {
      dout[i] = gin[(size_t)b * d.emb + i];
      o2g[i] = dout[i];
    }
    __syncthreads();
    mlp_backward(d, m, dout, sv + (which == 0 ? so.hid_style : so.hid_prompt), hs, preg, dxn, b, true);
    for (int c = threadIdx.x; c < C; c += kHeadThreads) dxg[c] = dxn[c];
    ln_backward_add(m, dxn, sv + so.xhat_c, sv[so.rstd + E], dcomb, 1.f, C, red);
  }
  // comb = sum_e w_e mean_e
  if (has_sp) {
    for (int e = 0; e < E; ++e) {
      float s = 0.f;
      for (int c = threadIdx.x; c < C; c += kHeadThreads) {
        const float me = pooled[((size_t)e * B + b) * C + c] * d.inv_hw;
        s += dcomb[c] * me;
        dsm[e * C + c] += w[e] * dcomb[c];
      }
      s = block_sum(s, red);
      if (threadIdx.x == 0) dw[e] += s;
    }
    __syncthreads();
  }
  // semantic head (expert 0): d logit = g * cos * s (1 - s); the cosine factor itself has zero gradient
  {
    float* preg = pre + bo.pre_sem;
    float* dxg = dxn_all + ((size_t)E * B + b) * C;
    if (g_semantic) {
      if (threadIdx.x == 0) {
        const float s = sv[so.sem_sig];
        dout[0] = g_semantic[b] * sv[so.cosv] * s * (1.f - s);
        o2[bo.o_sem] = dout[0];
      }
      __syncthreads();
      mlp_backward(d, d.semantic, dout, sv + so.hid_sem, hs, preg, dxn, b, true);
      for (int c = threadIdx.x; c < C; c += kHeadThreads) dxg[c] = dxn[c];
      ln_backward_add(d.semantic, dxn, sv + so.xhat_e, sv[so.rstd], dsm, 1.f, C, red);
    } else {
      if (threadIdx.x == 0) o2[bo.o_sem] = 0.f;
      for (int j = threadIdx.x; j < d.semantic.hid; j += kHeadThreads) preg[j] = 0.f;
      for (int c = threadIdx.x; c < C; c += kHeadThreads) dxg[c] = 0.f;
    }
  }
  // quality heads: wq = sum_e w_e q_e
  for (int e = 0; e < E; ++e) {
    if (threadIdx.x < 4) {
      dout[threadIdx.x] = w[e] * dwq[threadIdx.x];
      o2[bo.o_q + e * 4 + threadIdx.x] = dout[threadIdx.x];
    }
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int k = 0; k < 4; ++k) s += sv[so.q_e + e * 4 + k] * dwq[k];
      dw[e] += s;
    }
    __syncthreads();
    mlp_backward(d, d.quality[e], dout, sv + so.hid_q + e * d.quality[0].hid, hs, pre + bo.pre_q + e * d.quality[0].hid,
                 dxn, b, true);
    float* dxg = dxn_all + ((size_t)e * B + b) * C;
    for (int c = threadIdx.x; c < C; c += kHeadThreads) dxg[c] = dxn[c];
    ln_backward_add(d.quality[e], dxn, sv + so.xhat_e + e * C, sv[so.rstd + e], dsm + e * C, 1.f, C, red);
  }
  // softmax gate: d logit_e = w_e (dw_e - sum_j w_j dw_j); the gate's input (feature extractor) takes no gradient
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int e = 0; e < E; ++e) s += w[e] * dw[e];
    for (int e = 0; e < E; ++e) {
      dout[e] = w[e] * (dw[e] - s);
      o2[bo.o_gate + e] = dout[e];
    }
  }
  __syncthreads();
  mlp_backward(d, d.gate, dout, sv + so.hid_gate, hs, pre + bo.pre_gate, nullptr, b, false);
  // pooled SUMS: mean = sum / (H W)
  for (int e = 0; e < E; ++e)
    for (int c = threadIdx.x; c < C; c += kHeadThreads)
      d_pooled[((size_t)e * B + b) * C + c] = dsm[e * C + c] * d.inv_hw;
}

// ---------------------------------------------------------------------------------------------- backward, batch sums
// One thread per parameter element; sums over the B samples in a fixed order (deterministic, no atomics).
// blockIdx.y enumerates (head, tensor): head 0 gate, 1..E quality, E+1 semantic, E+2 style, E+3 prompt; tensor 0 ln_w,
// 1 ln_b, 2 w1, 3 b1, 4 w2, 5 b2. A null destination means the gradient is not wanted.
struct RedJob {
  float* dst;
  int kind;              // 0: w1 [hid,in]  1: b1 [hid]  2: w2 [out,hid]  3: b2 [out]  4: ln_w [C]  5: ln_b [C]
  int n;
  int head;
};
__device__ __forceinline__ const Mlp& head_of(const HeadsDesc& d, int h) {
  return h == 0 ? d.gate : h <= d.E ? d.quality[h - 1] : h == d.E + 1 ? d.semantic : h == d.E + 2 ? d.style : d.prompt;
}
__device__ __forceinline__ const MlpGrad& grad_of(const HeadsGrads& g, const HeadsDesc& d, int h) {
  return h == 0 ? g.gate : h <= d.E ? g.quality[h - 1] : h == d.E + 1 ? g.semantic : h == d.E + 2 ? g.style : g.prompt;
}
__global__ void __launch_bounds__(256) heads_bwd_reduce_kernel(const HeadsDesc d, const HeadsGrads gr,
                                                               const float* __restrict__ pooled_fe,
                                                               const float* __restrict__ save,
                                                               const float* __restrict__ dpre,
                                                               const float* __restrict__ dout2,
                                                               const float* __restrict__ dxn_all) {
  RedJob job;
  {
    const int h = blockIdx.y / 6, t = blockIdx.y % 6;
    const Mlp& m = head_of(d, h);
    const MlpGrad& mg = grad_of(gr, d, h);
    job.head = h;
    job.dst = t == 0 ? mg.ln_w : t == 1 ? mg.ln_b : t == 2 ? mg.w1 : t == 3 ? mg.b1 : t == 4 ? mg.w2 : mg.b2;
    job.kind = t == 0 ? 4 : t == 1 ? 5 : t - 2;
    job.n = t < 2 ? d.C : t == 2 ? m.hid * m.in : t == 3 ? m.hid : t == 4 ? m.out * m.hid : m.out;
    if (job.dst == nullptr || (h == 0 && t < 2)) return;
  }
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= job.n) return;
  const int B = d.B, E = d.E, C = d.C, h = job.head;
  const Mlp& m = head_of(d, h);
  const SaveOff so = save_offsets(d);
  const BwdOff bo = bwd_offsets(d);
  const int hq = d.quality[0].hid;
  const int pre_off = h == 0 ? bo.pre_gate : h <= E ? bo.pre_q + (h - 1) * hq : h == E + 1 ? bo.pre_sem
                      : h == E + 2 ? bo.pre_style : bo.pre_prompt;
  const int o_off = h == 0 ? bo.o_gate : h <= E ? bo.o_q + (h - 1) * 4 : h == E + 1 ? bo.o_sem
                    : h == E + 2 ? bo.o_style : bo.o_prompt;
  const int hid_off = h == 0 ? so.hid_gate : h <= E ? so.hid_q + (h - 1) * hq : h == E + 1 ? so.hid_sem
                      : h == E + 2 ? so.hid_style : so.hid_prompt;
  // normalised input of the head: xhat of expert e (quality e, semantic: expert 0) or of comb (style, prompt)
  const int xhat_off = h == 0 ? -1 : h <= E ? so.xhat_e + (h - 1) * C : h == E + 1 ? so.xhat_e : so.xhat_c;
  const int ln_slot = h - 1;                                  // dxn_all slot: quality e, E: semantic, E+1 style, E+2 prompt
  float acc = 0.f;
  if (job.kind == 0) {
    const int j = i / m.in, k = i % m.in;
    for (int b = 0; b < B; ++b) {
      const float x = h == 0 ? pooled_fe[(size_t)b * d.F + k] * d.inv_hw
                             : save[(size_t)b * so.total + xhat_off + k] * m.ln_w[k] + m.ln_b[k];
      acc = fmaf(dpre[(size_t)b * bo.pre_total + pre_off + j], x, acc);
    }
  } else if (job.kind == 1) {
    for (int b = 0; b < B; ++b) acc += dpre[(size_t)b * bo.pre_total + pre_off + i];
  } else if (job.kind == 2) {
    const int n = i / m.hid, j = i % m.hid;
    for (int b = 0; b < B; ++b) {
      const float a = save[(size_t)b * so.total + hid_off + j];
      const float hd = head_keep(d, m.seed, (size_t)b * m.hid + j) ? a * d.drop_scale : 0.f;
      acc = fmaf(dout2[(size_t)b * bo.o_total + o_off + n], hd, acc);
    }
  } else if (job.kind == 3) {
    for (int b = 0; b < B; ++b) acc += dout2[(size_t)b * bo.o_total + o_off + i];
  } else if (job.kind == 4) {
    for (int b = 0; b < B; ++b)
      acc = fmaf(dxn_all[((size_t)ln_slot * B + b) * C + i], save[(size_t)b * so.total + xhat_off + i], acc);
  } else {
    for (int b = 0; b < B; ++b) acc += dxn_all[((size_t)ln_slot * B + b) * C + i];
  }
  job.dst[i] = acc;
}

static int check_desc(const HeadsDesc& d) {
  if (d.E < 1 || d.E > kMaxExperts || d.B < 1 || d.C % 4 || d.F % 4 || d.emb < 1) return LUN_E_SHAPE;
  const Mlp* all[4] = {&d.gate, &d.semantic, &d.style, &d.prompt};
  for (const Mlp* m : all)
    if (m->hid % 4 || m->in % 4) return LUN_E_SHAPE;
  for (int e = 0; e < d.E; ++e)
    if (d.quality[e].hid != d.quality[0].hid || d.quality[e].hid % 4 || d.quality[e].out != 4) return LUN_E_SHAPE;
  return LUN_OK;
}
static int smem_floats(const HeadsDesc& d) {
  int maxh = d.gate.hid;
  for (int v : {d.quality[0].hid, d.semantic.hid, d.style.hid, d.prompt.hid}) maxh = v > maxh ? v : maxh;
  int maxo = d.E > 4 ? d.E : 4;
  maxo = d.emb > maxo ? d.emb : maxo;
  return d.F + d.E * d.C + 3 * d.C + maxh + maxo + kMaxExperts * 5 + 32;
}

}  // namespace lun

using namespace lun;

// The C ABI passes the head descriptions as flat arrays (plain pointers and sizes):
//   params: 6 pointers per head {ln_w, ln_b, w1, b1, w2, b2} (ln_* null for the gate), heads in the order
//           gate, quality[0..E), semantic, style, prompt  -> (E + 4) * 6 entries
//   dims:   {B, E, F, C, emb, gate_hid, quality_hid, semantic_hid, style_hid, prompt_hid}
//   seeds:  E + 4 dropout seeds in head order (ignored when drop_p == 0 or training == 0)
static int build_desc(HeadsDesc& d, const void* const* params, const int* dims, const unsigned long long* seeds,
                      float inv_hw, float slope, float drop_p) {
  d.B = dims[0]; d.E = dims[1]; d.F = dims[2]; d.C = dims[3]; d.emb = dims[4];
  if (d.E < 1 || d.E > kMaxExperts) return LUN_E_SHAPE;
  auto fill = [&](Mlp& m, int h, int in, int hid, int out) {
    const float* const* p = reinterpret_cast<const float* const*>(params) + 6 * h;
    m.ln_w = p[0]; m.ln_b = p[1]; m.w1 = p[2]; m.b1 = p[3]; m.w2 = p[4]; m.b2 = p[5];
    m.in = in; m.hid = hid; m.out = out;
    m.seed = seeds ? seeds[h] : 0ull;
  };
  fill(d.gate, 0, d.F, dims[5], d.E);
  for (int e = 0; e < d.E; ++e) fill(d.quality[e], 1 + e, d.C, dims[6], 4);
  fill(d.semantic, d.E + 1, d.C, dims[7], 1);
  fill(d.style, d.E + 2, d.C, dims[8], d.emb);
  fill(d.prompt, d.E + 3, d.C, dims[9], d.emb);
  d.inv_hw = inv_hw; d.slope = slope;
  d.thresh16 = drop_p > 0.f ? (unsigned int)(drop_p * 65536.f + 0.5f) : 0u;
  d.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  return check_desc(d);
}

extern "C" {

// dims -> {floats per sample of `save`, floats per sample of dpre, floats per sample of dout2} (forward / backward work
// buffer sizes; dxn needs (E + 3) * B * C floats)
int lun_heads_buffer_sizes(const int* dims, int* sizes) {
  HeadsDesc d{};
  d.B = dims[0]; d.E = dims[1]; d.F = dims[2]; d.C = dims[3]; d.emb = dims[4];
  d.gate.hid = dims[5];
  for (int e = 0; e < kMaxExperts; ++e) d.quality[e].hid = dims[6];
  d.semantic.hid = dims[7]; d.style.hid = dims[8]; d.prompt.hid = dims[9];
  if (d.E < 1 || d.E > kMaxExperts) return LUN_E_SHAPE;
  sizes[0] = save_offsets(d).total;
  const BwdOff bo = bwd_offsets(d);
  sizes[1] = bo.pre_total;
  sizes[2] = bo.o_total;
  return LUN_OK;
}

int lun_heads_fwd(const void* const* params, const int* dims, const unsigned long long* seeds, float inv_hw, float slope,
                  float drop_p, int training, const float* pooled_fe, const float* pooled, float* quality,
                  float* weights, float* style, float* prompt, float* semantic, float* save, void* stream) {
  HeadsDesc d{};
  const int rc = build_desc(d, params, dims, seeds, inv_hw, slope, training ? drop_p : 0.f);
  if (rc) return rc;
  const int smem = smem_floats(d) * 4;
  if (smem > 48 * 1024) return LUN_E_SHAPE;
  heads_fwd_kernel<<<d.B, kHeadThreads, smem, (cudaStream_t)stream>>>(
      d, pooled_fe, pooled, quality, weights, style, prompt, semantic, save, training);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

// grads: 6 pointers per head in the same order as params (null: gradient not wanted). Work buffers: dpre
// [B, sum of hidden sizes], dout2 [B, E + 4 E + 1 + 2 emb], dxn [(E + 3), B, C] floats; d_pooled [E][B][C].
int lun_heads_bwd(const void* const* params, const int* dims, const unsigned long long* seeds, float inv_hw, float slope,
                  float drop_p, const float* pooled_fe, const float* pooled, const float* weights, const float* save,
                  const float* g_quality, const float* g_weights, const float* g_style, const float* g_prompt,
                  const float* g_semantic, float* dpre, float* dout2, float* dxn, float* d_pooled, void* const* grads,
                  void* stream) {
  HeadsDesc d{};
  const int rc = build_desc(d, params, dims, seeds, inv_hw, slope, drop_p);
  if (rc) return rc;
  const int smem = smem_floats(d) * 4;
  if (smem > 48 * 1024) return LUN_E_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  heads_bwd_sample_kernel<<<d.B, kHeadThreads, smem, st>>>(d, pooled, weights, save, g_quality, g_weights, g_style,
                                                          g_prompt, g_semantic, dpre, dout2, dxn, d_pooled);
  if (cudaGetLastError() != cudaSuccess) return LUN_E_LAUNCH;
  HeadsGrads gr{};
  int maxn = 0;
  bool any = false;
  const int nheads = d.E + 4;
  for (int h = 0; h < nheads; ++h) {
    const Mlp& m = h == 0 ? d.gate : h <= d.E ? d.quality[h - 1] : h == d.E + 1 ? d.semantic : h == d.E + 2 ? d.style
                                                                                                        : d.prompt;
    MlpGrad& mg = h == 0 ? gr.gate : h <= d.E ? gr.quality[h - 1] : h == d.E + 1 ? gr.semantic : h == d.E + 2 ? gr.style
                                                                                                           : gr.prompt;
    float* const* g = reinterpret_cast<float* const*>(grads) + 6 * h;
    mg.ln_w = g[0]; mg.ln_b = g[1]; mg.w1 = g[2]; mg.b1 = g[3]; mg.w2 = g[4]; mg.b2 = g[5];
    for (int t = 0; t < 6; ++t) any = any || g[t];
    const int n = m.hid * m.in > d.C ? m.hid * m.in : d.C;
    maxn = n > maxn ? n : maxn;
    maxn = m.out * m.hid > maxn ? m.out * m.hid : maxn;
  }
  if (!any) { lun::note_launch(1); return LUN_OK; }
  dim3 grid((maxn + 255) / 256, nheads * 6);
  heads_bwd_reduce_kernel<<<grid, 256, 0, st>>>(d, gr, pooled_fe, save, dpre, dout2, dxn);
  lun::note_launch(2);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

}  // extern "C"
