// As-executed Teacher attention without ever forming K or V (lunar_evaluator.py:153-156, 203-216).
//
// After the reference's chunk-index scatter only ONE query per 32-token chunk survives (plus the 32 queries of the
// last chunk). For such a query q of head h, with k_j = Wk_h x_j + bk_h and v_j = Wv_h x_j + bv_h:
//     score_j = q.k_j / sqrt(hd) = (Wk_h^T q).x_j / sqrt(hd) + const        (const drops out of the softmax)
//     out     = sum_j p_j v_j     = Wv_h (sum_j p_j x_j) + bv_h
// so the kernel needs the attention INPUT rows x_j only: per (image, surviving row i) it computes, for all 8 heads,
//     S[h][j] = qt[h] . x_j,   P = softmax_j(S) (+ attn_drop),   xbar[h] = sum_j P[h][j] x_j
// where qt = (Wk_h^T Wq_h / sqrt(hd)) x_q + Wk_h^T bq_h / sqrt(hd) comes from one small GEMM and the head outputs
// Wv_h xbar[h] + bv_h from another. Both contractions here are tiny (8 heads): warp-level mma.sync m16n8k16 on a
// shared-memory copy of the 32 x C chunk; the kernel is bound by reading the chunk once from HBM.
// The BatchNorm + Dropout2d affine of the attention input is applied while the chunk is staged, so the normalised
// tensor x1 is never written to HBM either.
#include "../../include/lunaris_b200.h"
#include "elem_common.cuh"
#include "ptx.cuh"
#include "launch_count.cuh"
#include <stdlib.h>

namespace lun {

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
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

// Persistent CTAs (8 warps, two per SM) each own a contiguous range of items (image b, surviving row i).
// The BatchNorm + Dropout2d affine x = m*(s*y + t) never touches the 32 x C chunk:
//     scores:  qt . x_j = (qt*m*s) . y_j + const            -> the queries are scaled once (8 x C values)
//     output:  sum_j p_j x_j = m*(s * sum_j p_j y_j + t * sum_j p_j)   -> applied to the 8 x C result
//   y     [B, N, C]        attention input BEFORE BatchNorm (conv1 output)
//   qt    [B, nq_pad, 8*C] folded queries (head-major)
//   xbar  [B, nq_pad, 8*C] output
//   * the 32 x C chunk of the NEXT item travels by cp.async.bulk (one 2C-byte copy per token row, mbarrier
//     completion) into the other half of a two-stage ring while the current item is contracted: no thread spends
//     instructions or registers on the copy, and each of the two resident CTAs per SM always has a chunk in flight;
//   * S^T = Y Q'^T  (M = tokens, N = 8 heads) and  out^T = Y^T P^T  (M = channels, N = 8 heads): with the 8 heads on
//     the MMA's N side no half of an m16n8k16 tile is padding;
//   * a CTA owns a contiguous range of items, so the per-image affine lives in registers and is reloaded only when
//     the image changes; the 8 x C result leaves through shared memory and eight 2C-byte bulk stores.
constexpr int kAfThreads = 256;

__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Resident CTAs per SM: the kernel is bound by the latency of its four block-wide phases per item, not by issue slots or
// bandwidth, so every extra resident CTA adds throughput (C = 256, batch 16: 76.5 -> 54.3 us with 4 instead of 2);
// shared memory allows 2 at C = 512 and 4 from C = 256 down. An L2 prefetch of the chunks ahead of the two-stage ring
// (cp.async.bulk.prefetch.L2, 1-4 items ahead) was measured and changes nothing (355 -> 362 us at C = 512): the wait
// is not DRAM latency.
template <int C>
constexpr int af_ctas_per_sm() { return C <= 256 ? 4 : 2; }

template <int C>
__global__ void __launch_bounds__(kAfThreads, af_ctas_per_sm<C>())
attn_fold_kernel(const bf16* __restrict__ y, const float* __restrict__ scale, const float* __restrict__ shift,
                  const float* __restrict__ m2, const bf16* __restrict__ qt, bf16* __restrict__ xbar, int B, int N,
                  int nq, int nq_pad, unsigned long long seed, unsigned int thresh16, float drop_scale) {
  constexpr int PITCH = C + 8;                      // bf16 per smem row: +16 B keeps ldmatrix conflict-free
  constexpr int CH8 = C / 8;
  constexpr int NQ = (8 * CH8 + kAfThreads - 1) / kAfThreads;   // 16-byte query chunks per thread
  constexpr int KQ = C / 4;                         // channels per K-quarter of the score contraction
  constexpr int MT_TOTAL = C / 16;                  // 16-channel m-tiles of the output contraction
  constexpr int MTW = MT_TOTAL >= 8 ? MT_TOTAL / 8 : 1;           // m-tiles per warp
  constexpr int PV_WARPS = MT_TOTAL / MTW;
  constexpr int SPS = 36;                           // padded token stride of the partial-score buffer (floats)
  static_assert(kAfThreads % CH8 == 0 && KQ % 16 == 0, "unsupported channel count");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  bf16* Xs = reinterpret_cast<bf16*>(smem_raw);                  // [2][32][PITCH] chunk ring
  bf16* Qs = Xs + 2 * 32 * PITCH;                                // [8][PITCH] scaled queries
  bf16* Os = Qs + 8 * PITCH;                                     // [2][8][PITCH] results awaiting their bulk stores
  float* Sp = reinterpret_cast<float*>(Os + 2 * 8 * PITCH);      // [4 K-quarters][8 heads][SPS]
  uint32_t* Ps = reinterpret_cast<uint32_t*>(Sp + 4 * 8 * SPS);  // [8 heads][16] probabilities, bf16 pairs
  float* Psum = reinterpret_cast<float*>(Ps + 8 * 16);           // [8]
  uint64_t* full = reinterpret_cast<uint64_t*>(Psum + 8);        // [2]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int nc = N / 32;
  const int items = B * nq;
  const int per = (items + gridDim.x - 1) / gridDim.x;
  const int it0 = blockIdx.x * per;
  const int it1 = it0 + per < items ? it0 + per : items;
  if (it0 >= it1) return;
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    fence_barrier_init();
  }
  __syncthreads();

  // The 32 row copies of a chunk are issued by ALL eight warps, four rows each (lanes 0-3): a bulk copy is a uniform-
  // datapath instruction, so a warp issues its lanes' copies one after the other (~30 cycles each) - with one warp
  // issuing all 32 the other seven waited ~1000 cycles per item at the next block barrier (ncu source view: 29 % of all
  // stall samples sat behind that barrier). Warp 0 arms the barrier; a row may complete before the barrier is armed
  // (the transaction count goes transiently negative, the phase cannot complete before the arming arrival).
  auto issue_chunk = [&](int b, int i, int stage) {   // every warp
    const int chunk = i < nc - 1 ? i : nc - 1;
    if (tid == 0) mbar_expect_tx(&full[stage], 32 * C * 2);
    if (lane < 4) {
      const int row = warp * 4 + lane;
      bulk_load(smem_u32(Xs + (stage * 32 + row) * PITCH), y + ((size_t)b * N + 32 * chunk + row) * C, C * 2,
                &full[stage]);
    }
  };
  auto load_q = [&](int b, int i, uint4 (&q)[NQ]) {
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const int idx = tid + k * kAfThreads;
      if (idx < 8 * CH8)
        q[k] = __ldg(reinterpret_cast<const uint4*>(qt + (((size_t)b * nq_pad + i) * 8 + idx / CH8) * C + (idx % CH8) * 8));
    }
  };
  // per-image affine in registers: ms = m*s for this thread's query channels; A = m*s, Bc = m*t for its output channels
  const int qc = (tid % CH8) * 8;
  float ms[8], pa[MTW][2], pb[MTW][2];
  auto load_affine = [&](int b) {
#pragma unroll
    for (int j = 0; j < 8; ++j) ms[j] = scale[qc + j] * (m2 ? m2[(size_t)b * C + qc + j] : 1.f);
#pragma unroll
    for (int mt = 0; mt < MTW; ++mt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int c = (warp * MTW + mt) * 16 + g + hh * 8;
        const float mk = (m2 && c < C) ? m2[(size_t)b * C + c] : 1.f;
        pa[mt][hh] = c < C ? mk * scale[c] : 0.f;
        pb[mt][hh] = c < C ? mk * shift[c] : 0.f;
      }
  };
  // (image, row) of an item, advanced incrementally (no division in the loop)
  auto advance = [&](int& b, int& i) {
    if (++i == nq) {
      i = 0;
      ++b;
    }
  };

  int b = it0 / nq, i = it0 % nq;
  uint4 qraw[NQ];
  {
    int b1 = b, i1 = i;
    advance(b1, i1);
    issue_chunk(b, i, 0);
    if (it0 + 1 < it1) issue_chunk(b1, i1, 1);
  }
  load_q(b, i, qraw);
  int cur_b = -1;
  const uint32_t qs_base = smem_u32(Qs);
  const int mat = lane >> 3, lrow = lane & 7;

  for (int item = it0; item < it1; ++item) {
    const int n = item - it0, stage = n & 1;
    int b1 = b, i1 = i;
    advance(b1, i1);
    if (b != cur_b) {
      load_affine(b);
      cur_b = b;
    }
    // ---- queries scaled by m*s -> Qs; next item's raw queries -> registers
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      const int idx = tid + k * kAfThreads;
      if (idx < 8 * CH8) {
        float v[8];
        unpack8(qraw[k], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= ms[j];
        store8(Qs + (idx / CH8) * PITCH + qc, v);
      }
    }
    if (item + 1 < it1) load_q(b1, i1, qraw);
    mbar_wait(&full[stage], (n >> 1) & 1);         // this item's chunk has landed
    __syncthreads();                               // (1) Qs visible

    // ---- S^T partial: warp -> (16-token m-tile, K-quarter)
    const uint32_t xs_base = smem_u32(Xs + stage * 32 * PITCH);
    {
      const int mt = warp & 1, kq = warp >> 1;
      float d[4] = {0.f, 0.f, 0.f, 0.f}, e[4] = {0.f, 0.f, 0.f, 0.f};     // two accumulators: independent MMA chains
      // A (tokens x k): matrices {tok 0-7,k 0-7}, {tok 8-15,k 0-7}, {tok 0-7,k 8-15}, {tok 8-15,k 8-15}
      const uint32_t a_addr = xs_base + ((mt * 16 + lrow + (mat & 1) * 8) * PITCH + kq * KQ + (mat >> 1) * 8) * 2;
      // B (heads x k) for two k-steps: {k 0-7}, {k 8-15}, {k 16-23}, {k 24-31}
      const uint32_t b_addr = qs_base + (lrow * PITCH + kq * KQ + mat * 8) * 2;
#pragma unroll
      for (int k0 = 0; k0 < KQ; k0 += 32) {
        uint32_t bq[4], a0[4], a1[4];
        ldsm_x4(bq, b_addr + k0 * 2);
        ldsm_x4(a0, a_addr + k0 * 2);
        if (k0 + 16 < KQ) ldsm_x4(a1, a_addr + (k0 + 16) * 2);
        {
          const uint32_t bb[2] = {bq[0], bq[1]};
          mma_bf16_16816(d, a0, bb);
        }
        if (k0 + 16 < KQ) {
          const uint32_t bb[2] = {bq[2], bq[3]};
          mma_bf16_16816(e, a1, bb);
        }
      }
      // d0,d1: token g, heads 2t4, 2t4+1;  d2,d3: token g+8
      float* sp = Sp + (kq * 8 + 2 * t4) * SPS + mt * 16 + g;
      sp[0] = d[0] + e[0];
      sp[SPS] = d[1] + e[1];
      sp[8] = d[2] + e[2];
      sp[SPS + 8] = d[3] + e[3];
    }
    __syncthreads();                               // (2) partial scores visible

    // ---- softmax: warp = head, lane = token (one probability per thread, reductions by shuffles)
    {
      float sc = Sp[warp * SPS + lane] + Sp[(8 + warp) * SPS + lane] + Sp[(16 + warp) * SPS + lane] +
                 Sp[(24 + warp) * SPS + lane];
      sc = rbf(sc);                                // the reference's scores are a bf16 tensor
      float mx = sc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      const float ex = __expf(sc - mx);
      float p = ex / warp_sum(ex);
      if (thresh16) {
        // same stream as drop_keep1(seed, ((b*nq + i)*8 + head) << 5 | token, .)
        const unsigned long long base = (((unsigned long long)b * nq + i) * 8 + warp) << 5;
        const uint32_t h = hash32(drop_key(seed, (base >> 3) + (lane >> 3)) + ((lane >> 1) & 3));
        p = ((lane & 1) ? (h >> 16) : (h & 0xFFFFu)) >= thresh16 ? p * drop_scale : 0.f;
      }
      p = rbf(p);
      const float psum = warp_sum(p);
      const float pn = __shfl_down_sync(0xffffffffu, p, 1);
      if (!(lane & 1)) Ps[warp * 16 + (lane >> 1)] = pack2(p, pn);
      if (lane == 0) Psum[warp] = psum;
    }
    __syncthreads();                               // (3) probabilities visible

    // ---- out^T = Y^T P^T: warp -> MTW m-tiles of 16 channels; xbar = A * d + Bc * psum
    bf16* os = Os + stage * 8 * PITCH;
    if (warp < PV_WARPS) {
      // B fragments of P^T: [k-step][b0, b1] = head g, tokens 16ks + {2t4, 2t4+1}, {8+2t4, 9+2t4}
      uint32_t pfrag[4];
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) pfrag[nn] = Ps[g * 16 + nn * 4 + t4];
      const float ps0 = Psum[2 * t4], ps1 = Psum[2 * t4 + 1];
#pragma unroll
      for (int mt = 0; mt < MTW; ++mt) {
        const int c0 = (warp * MTW + mt) * 16;
        float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          // A^T from [token][channel] storage: {tok 0-7,ch 0-7}, {tok 0-7,ch 8-15}, {tok 8-15,ch 0-7}, {tok 8-15,ch 8-15}
          uint32_t a[4];
          ldsm_x4_trans(a, xs_base + ((ks * 16 + lrow + (mat >> 1) * 8) * PITCH + c0 + (mat & 1) * 8) * 2);
          const uint32_t bb[2] = {pfrag[2 * ks], pfrag[2 * ks + 1]};
          mma_bf16_16816(d, a, bb);
        }
        // d0,d1: channel c0+g, heads 2t4, 2t4+1;  d2,d3: channel c0+g+8
        os[(2 * t4) * PITCH + c0 + g] = __float2bfloat16_rn(pa[mt][0] * d[0] + pb[mt][0] * ps0);
        os[(2 * t4 + 1) * PITCH + c0 + g] = __float2bfloat16_rn(pa[mt][0] * d[1] + pb[mt][0] * ps1);
        os[(2 * t4) * PITCH + c0 + g + 8] = __float2bfloat16_rn(pa[mt][1] * d[2] + pb[mt][1] * ps0);
        os[(2 * t4 + 1) * PITCH + c0 + g + 8] = __float2bfloat16_rn(pa[mt][1] * d[3] + pb[mt][1] * ps1);
      }
      fence_proxy_async();                         // results visible to the bulk store
    }
    __syncthreads();                               // (4) chunk stage, Qs, Sp, Ps and Os[stage] are settled
    if (lane == 0) {                               // one head row per warp (bulk groups are per thread)
      bulk_store(xbar + (((size_t)b * nq_pad + i) * 8 + warp) * C, smem_u32(os + warp * PITCH), C * 2);
      // Os[stage ^ 1] is rewritten by the next item: its stores (committed one item ago) must have read it
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
    if (item + 2 < it1) {
      int b2 = b1, i2 = i1;
      advance(b2, i2);
      issue_chunk(b2, i2, stage);
    }
    b = b1;
    i = i1;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = (2 * 32 + 8 + 2 * 8) * (C + 8) * 2 + 4 * 8 * 36 * 4 + 8 * 16 * 4 + 8 * 4 + 16;
  static bool configured_dev[64] = {};   // the opt-in smem size is a per-device function attribute
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(attn_fold_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return LUN_E_ATTR;
    configured = true;
  }
  long grid = (long)af_ctas_per_sm<C>() * sms;       // all resident at once, contiguous item ranges
  if (grid > (long)B * nq) grid = (long)B * nq;
  attn_fold_kernel<C><<<(int)grid, kAfThreads, smem, st>>>(y, scale, shift, m2, qt, xbar, B, N, nq, nq_pad, seed, th, ds);
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
