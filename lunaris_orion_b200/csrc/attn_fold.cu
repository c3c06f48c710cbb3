// As-executed Teacher attention without ever forming K or V (lunar_evaluator.py:153-156, 203-216).
//
// After the reference's chunk-index scatter only ONE query per 32-token chunk survives (plus the 32 queries of the
// last chunk). For such a query q of head h, with k_j = Wk_h x_j + bk_h and v_j = Wv_h x_j + bv_h:
//     score_j = q.k_j / sqrt(hd) = (Wk_h^T q).x_j / sqrt(hd) + const        (const drops out of the softmax)
//     out     = sum_j p_j v_j     = Wv_h (sum_j p_j x_j) + bv_h
// so the kernel needs the attention INPUT rows x_j only: per (image, surviving row i) it computes, for all 8 heads,
//     S[h][j] = qt[h] . x_j,   P = softmax_j(S) (+ attn_drop),   xbar[h] = sum_j P[h][j] x_j
// where qt = (Wk_h^T Wq_h / sqrt(hd)) x_q + Wk_h^T bq_h / sqrt(hd) comes from one small GEMM and the head outputs
// Wv_h xbar[h] + bv_h from another. Both contractions here are tiny (M = 8 heads): warp-level mma.sync m16n8k16 on
// a shared-memory copy of the 32 x C chunk; the kernel is bound by reading the chunk once from HBM.
// The BatchNorm + Dropout2d affine of the attention input is applied while the chunk is staged, so the normalised
// tensor x1 is never written to HBM either.
#include "../../include/lunaris_b200.h"
#include "elem_common.cuh"
#include "ptx.cuh"
#include "launch_count.cuh"
#include <stdlib.h>

namespace lun {

constexpr int kAfThreads = 128;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

// Persistent blocks (4 warps) loop over items (image b, surviving row i); the raw 32 x C chunk of the NEXT item is
// fetched with cp.async into the other half of a double buffer while the current one is contracted.
// The BatchNorm + Dropout2d affine x = m*(s*y + t) never touches the 32 x C chunk:
//     scores:  qt . x_j = (qt*m*s) . y_j + const            -> the queries are scaled once (8 x C values)
//     output:  sum_j p_j x_j = m*(s * sum_j p_j y_j + t * sum_j p_j)   -> applied to the 8 x C result
//   y     [B, N, C]        attention input BEFORE BatchNorm (conv1 output)
//   qt    [B, nq_pad, 8*C] folded queries (head-major)
//   xbar  [B, nq_pad, 8*C] output
template <int C, int NBUF>
__global__ void __launch_bounds__(kAfThreads) attn_fold_kernel(const bf16* __restrict__ y, const float* __restrict__ scale,
                                                               const float* __restrict__ shift,
                                                               const float* __restrict__ m2,
                                                               const bf16* __restrict__ qt, bf16* __restrict__ xbar,
                                                               int B, int N, int nq, int nq_pad,
                                                               unsigned long long seed, unsigned int thresh16,
                                                               float drop_scale) {
  constexpr int PITCH = C + 8;                     // bf16 elements per smem row (+16 B: conflict-free ldmatrix)
  constexpr int CH8 = C / 8;
  constexpr int NL = 32 * CH8 / kAfThreads;        // 16-byte chunk copies per thread
  constexpr int NQ = (8 * CH8 + kAfThreads - 1) / kAfThreads;   // query chunks per thread
  constexpr int RSTEP = kAfThreads / CH8;
  static_assert(kAfThreads % CH8 == 0 && NL >= 1, "unsupported channel count");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* Xs0 = reinterpret_cast<bf16*>(smem_raw);            // [2][32][PITCH] raw chunk rows (tokens), double buffer
  bf16* Qs = Xs0 + NBUF * 32 * PITCH;                       // [8][PITCH] scaled queries
  float* Sp = reinterpret_cast<float*>(Qs + 8 * PITCH);     // [8 warps][8][32] partial scores
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nc = N / 32;
  const int items = B * nq;
  const int c8 = tid % CH8, r0 = tid / CH8;
  const int g = lane >> 2, t4 = lane & 3;
  constexpr int NWARPS = kAfThreads / 32;
  constexpr int NW = (C / 16 < NWARPS) ? C / 16 : NWARPS;   // warps that take part in the two contractions
  constexpr int KW = C / NW;                       // channels per warp

  auto issue_chunk = [&](int item, int buf) {
    const int b = item / nq, i = item % nq;
    const int chunk = i < nc - 1 ? i : nc - 1;
    const bf16* src = y + ((size_t)b * N + 32 * chunk + r0) * C + c8 * 8;
    const uint32_t dst = smem_u32(Xs0 + (buf * 32 + r0) * PITCH + c8 * 8);
#pragma unroll
    for (int k = 0; k < NL; ++k) cp_async16(dst + k * RSTEP * PITCH * 2, src + (size_t)k * RSTEP * C);
    cp_async_commit();
  };

  // query chunks owned by this thread: idx = tid + k*128 over [8 heads][CH8]; for C >= 128 the channel group of a
  // thread is fixed (128 % CH8 == 0). Queries and the Dropout2d mask of the NEXT item are prefetched into registers
  // while the current item is contracted.
  auto load_q = [&](int item, uint4 (&q)[NQ], float4 (&mk)[NQ][2]) {
    const int b = item / nq, i = item % nq;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const int idx = tid + k * kAfThreads;
      if (idx < 8 * CH8) {
        const int qc = (idx % CH8) * 8;
        q[k] = __ldg(reinterpret_cast<const uint4*>(qt + (((size_t)b * nq_pad + i) * 8 + idx / CH8) * C + qc));
        if (m2) {
          mk[k][0] = __ldg(reinterpret_cast<const float4*>(m2 + (size_t)b * C + qc));
          mk[k][1] = __ldg(reinterpret_cast<const float4*>(m2 + (size_t)b * C + qc + 4));
        } else {
          mk[k][0] = mk[k][1] = make_float4(1.f, 1.f, 1.f, 1.f);
        }
      }
    }
  };
  int item = blockIdx.x;
  uint4 qraw[NQ];
  float4 mraw[NQ][2];
  if (item < items) {
    issue_chunk(item, 0);
    load_q(item, qraw, mraw);
  }
  int buf = 0;
  for (; item < items; item += gridDim.x, buf ^= (NBUF - 1)) {
    const int b = item / nq, i = item % nq;
    // stage this item's queries scaled by m*s, then prefetch the next item's
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const int idx = tid + k * kAfThreads;
      if (idx < 8 * CH8) {
        const int qc = (idx % CH8) * 8;
        float v[8];
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&qraw[k]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __bfloat1622float2(h2[j]);
          v[2 * j] = t.x;
          v[2 * j + 1] = t.y;
        }
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + qc));
        const float4 s1 = __ldg(reinterpret_cast<const float4*>(scale + qc + 4));
        v[0] *= s0.x * mraw[k][0].x; v[1] *= s0.y * mraw[k][0].y; v[2] *= s0.z * mraw[k][0].z; v[3] *= s0.w * mraw[k][0].w;
        v[4] *= s1.x * mraw[k][1].x; v[5] *= s1.y * mraw[k][1].y; v[6] *= s1.z * mraw[k][1].z; v[7] *= s1.w * mraw[k][1].w;
        store8(Qs + (idx / CH8) * PITCH + qc, v);
      }
    }
    const int next = item + gridDim.x;
    if (next < items) {
      if (NBUF == 2) issue_chunk(next, buf ^ 1);
      load_q(next, qraw, mraw);
    }
    if (NBUF == 2 && next < items) cp_async_wait<1>(); else cp_async_wait<0>();   // this item's chunk has landed
    __syncthreads();

    // ---- S = Q' Y^T : each warp owns a quarter of the channel (k) range, partials are summed through smem
    const uint32_t xs_base = smem_u32(Xs0 + buf * 32 * PITCH), qs_base = smem_u32(Qs);
    float s[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[n][e] = 0.f;
#pragma unroll 2
    for (int k0 = warp * KW; warp < NW && k0 < (warp + 1) * KW; k0 += 16) {
      uint32_t a[4], a2[2];
      ldsm_x2(a2, qs_base + ((lane & 7) * PITCH + k0 + ((lane >> 3) & 1) * 8) * 2);
      a[0] = a2[0]; a[1] = 0u; a[2] = a2[1]; a[3] = 0u;          // rows 8..15 of the M=16 tile are padding
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        uint32_t bb[2];
        ldsm_x2(bb, xs_base + ((n * 8 + (lane & 7)) * PITCH + k0 + ((lane >> 3) & 1) * 8) * 2);
        mma_bf16_16816(s[n], a, bb);
      }
    }
    if (warp < NW) {
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        float* sp = Sp + (warp * 8 + g) * 32 + n * 8 + t4 * 2;
        sp[0] = s[n][0];
        sp[1] = s[n][1];
      }
    }
    __syncthreads();
    // every warp rebuilds the full score fragment (rows = heads g, cols = tokens)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      s[n][0] = 0.f;
      s[n][1] = 0.f;
    }
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const float2 t = *reinterpret_cast<const float2*>(Sp + (w * 8 + g) * 32 + n * 8 + t4 * 2);
        s[n][0] += t.x;
        s[n][1] += t.y;
      }
    // ---- softmax over the 32 tokens of head g
    float m0 = -3.0e38f;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      s[n][0] = rbf(s[n][0]);                      // the reference's scores are a bf16 tensor
      s[n][1] = rbf(s[n][1]);
      m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    float sum = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      s[n][0] = __expf(s[n][0] - m0);
      s[n][1] = __expf(s[n][1] - m0);
      sum += s[n][0] + s[n][1];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.f / sum;
    uint32_t pa[2][4];                             // P as A fragments of the two token k-steps
    float psum = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      float p0 = s[n][0] * inv, p1 = s[n][1] * inv;
      if (thresh16) {
        const unsigned long long base = ((((unsigned long long)b * nq + i) * 8 + g) << 5) + n * 8 + t4 * 2;
        p0 = drop_keep1(seed, base, thresh16) ? p0 * drop_scale : 0.f;
        p1 = drop_keep1(seed, base + 1, thresh16) ? p1 * drop_scale : 0.f;
      }
      p0 = rbf(p0);
      p1 = rbf(p1);
      psum += p0 + p1;
      pa[n >> 1][(n & 1) * 2] = pack2(p0, p1);     // rows g     (a0 / a2)
      pa[n >> 1][(n & 1) * 2 + 1] = 0u;            // rows g + 8 (a1 / a3): padding heads contribute nothing
    }
    psum += __shfl_xor_sync(0xffffffffu, psum, 1);
    psum += __shfl_xor_sync(0xffffffffu, psum, 2);
    // ---- Ybar = P Y, then xbar = m*(s*Ybar + t*psum); each warp owns a quarter of the output channels
    bf16* out = xbar + (((size_t)b * nq_pad + i) * 8 + g) * C;
#pragma unroll 4
    for (int n0 = warp * KW; warp < NW && n0 < (warp + 1) * KW; n0 += 8) {
      float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t bb[2];
        ldsm_x2_trans(bb, xs_base + ((kk * 16 + (lane & 15)) * PITCH + n0) * 2);
        mma_bf16_16816(d, pa[kk], bb);
      }
      const int c = n0 + t4 * 2;
      const float2 sc = *reinterpret_cast<const float2*>(scale + c), sh = *reinterpret_cast<const float2*>(shift + c);
      float2 mk = make_float2(1.f, 1.f);
      if (m2) mk = *reinterpret_cast<const float2*>(m2 + (size_t)b * C + c);
      *reinterpret_cast<uint32_t*>(out + c) = pack2(mk.x * (sc.x * d[0] + sh.x * psum), mk.y * (sc.y * d[1] + sh.y * psum));
    }
    __syncthreads();                               // Qs / Sp / this X buffer are rewritten by the next iterations
    if (NBUF == 1 && next < items) issue_chunk(next, 0);
  }
}

// out[b, i, :] = drop2d(bn(y[b, qtok(i), :])): the query rows of the as-executed attention with its input affine.
__global__ void gather_query_rows_affine_kernel(const bf16* __restrict__ y, const float* __restrict__ scale,
                                                const float* __restrict__ shift, const float* __restrict__ m2,
                                                bf16* __restrict__ out, int N, int C, int nq, int nq_pad) {
  const int i = blockIdx.x, b = blockIdx.y;
  const int nc = N / 32;
  const int qtok = i < nc - 1 ? 32 * i : 32 * (nc - 1) + (i - (nc - 1));
  for (int c8 = threadIdx.x; c8 < C / 8; c8 += blockDim.x) {
    float v[8];
    load8(y + ((size_t)b * N + qtok) * C + c8 * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 * 8 + j;
      float t = rbf(v[j] * scale[c] + shift[c]);
      if (m2) t = rbf(t * m2[(size_t)b * C + c]);
      v[j] = t;
    }
    store8(out + ((size_t)b * nq_pad + i) * C + c8 * 8, v);
  }
}

template <int C>
static int launch_fold(const bf16* y, const float* scale, const float* shift, const float* m2, const bf16* qt,
                       bf16* xbar, int B, int N, int nq, int nq_pad, unsigned long long seed, unsigned int th,
                       float ds, cudaStream_t st) {
  static int nbuf = 0;
  if (!nbuf) {
    const char* e = getenv("LUN_ATTN_NBUF");
    nbuf = e ? atoi(e) : 1;
    if (nbuf != 2) nbuf = 1;
  }
  const int smem = (32 * nbuf + 8) * (C + 8) * 2 + 8 * 8 * 32 * 4;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attn_fold_kernel<C, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(attn_fold_kernel<C, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess)
      return LUN_E_ATTR;
    configured = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = (224 * 1024) / (smem + 1024);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int grid = sms * per_sm;
  if (grid > B * nq) grid = B * nq;
  if (nbuf == 2)
    attn_fold_kernel<C, 2><<<grid, kAfThreads, smem, st>>>(y, scale, shift, m2, qt, xbar, B, N, nq, nq_pad, seed, th, ds);
  else
    attn_fold_kernel<C, 1><<<grid, kAfThreads, smem, st>>>(y, scale, shift, m2, qt, xbar, B, N, nq, nq_pad, seed, th, ds);
  return LUN_OK;
}

}  // namespace lun

using namespace lun;

extern "C" {

int lun_attn_fold_rows_bf16(const void* y, const float* scale, const float* shift, const float* mask2d,
                            const void* qt, void* xbar, int B, int N, int C, int heads, int nq_pad,
                            unsigned long long seed, float drop_p, void* stream) {
  if (heads != 8 || N % 32) return LUN_E_SHAPE;
  const int nq = N / 32 + 31;
  if (nq_pad < nq) return LUN_E_SHAPE;
  const unsigned int th = drop_p > 0.f ? (unsigned int)(drop_p * 65536.f + 0.5f) : 0u;
  const float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (C == 512) rc = launch_fold<512>((const bf16*)y, scale, shift, mask2d, (const bf16*)qt, (bf16*)xbar, B, N, nq, nq_pad, seed, th, ds, st);
  else if (C == 256) rc = launch_fold<256>((const bf16*)y, scale, shift, mask2d, (const bf16*)qt, (bf16*)xbar, B, N, nq, nq_pad, seed, th, ds, st);
  else if (C == 128) rc = launch_fold<128>((const bf16*)y, scale, shift, mask2d, (const bf16*)qt, (bf16*)xbar, B, N, nq, nq_pad, seed, th, ds, st);
  else if (C == 64) rc = launch_fold<64>((const bf16*)y, scale, shift, mask2d, (const bf16*)qt, (bf16*)xbar, B, N, nq, nq_pad, seed, th, ds, st);
  else return LUN_E_SHAPE;
  if (rc) return rc;
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_gather_query_rows_affine_bf16(const void* y, const float* scale, const float* shift, const float* mask2d,
                                      void* out, int B, int N, int C, int nq_pad, void* stream) {
  if (N % 32 || C % 8) return LUN_E_SHAPE;
  const int nq = N / 32 + 31;
  if (nq_pad < nq) return LUN_E_SHAPE;
  dim3 grid(nq, B);
  gather_query_rows_affine_kernel<<<grid, 64, 0, (cudaStream_t)stream>>>((const bf16*)y, scale, shift, mask2d,
                                                                        (bf16*)out, N, C, nq, nq_pad);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

}  // extern "C"
