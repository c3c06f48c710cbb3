// Backward of SelfAttention2d (lunar_generate.py:66-78; autograd of softmax(q k^T) v) for the query and key projections,
// flash-style on tcgen05: no N x N matrix, no atomics, deterministic.
//
//   S = Q K^T,  P = exp(S - lse),  dP = dY V^T,  dS = P o (dP - D),  D_i = sum_c dY_ic O_ic
//   dQ = gamma * dS K          dK = gamma * dS^T Q          (gamma folded here: dO = gamma * dY)
//
// One kernel serves both gradients by swapping roles. A CTA OWNS 128 rows (queries for dQ, keys for dK) and STREAMS
// the 128-row tiles of the other kind:
//   S'  = Xqk Yqk^T            (Xqk = own q|k rows,  Yqk = streamed k|q rows; 64 zero-padded dims, K-major)
//   dP' = Xc  Yc^T             (Xc = own dY|V rows, resident in smem;  Yc = streamed V|dY rows in 64-channel chunks)
//   dS' = exp(S' - lse) o (dP' - D)   with lse, D indexed by the QUERY: the own row (dQ) or the streamed column (dK)
//   out += dS' Yqk             (dS' through a 128B-swizzled smem tile as K-major A, Yqk as N-major B)
// TMEM: S' 128 columns, dP' 2 x 128 columns (the next tile's dP' runs while the softmax warps work on this one),
// out 64 columns. Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = dS math (thread == own row == TMEM lane).
// The value gradient dV = gamma * P^T dY is the forward kernel in its dv mode (flash_attn2d_sm100.cu).
#include "../../include/lunaris_b200.h"
#include "conv_gemm.cuh"
#include "launch_count.cuh"
#include "ptx.cuh"

namespace lun {

int make_tmap_2d(CUtensorMap* m, const void* base, long rows, long cols, int box_rows);
int make_tmap_nhwc(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h, int box_b,
                   int estride);

constexpr int kFbThreads = 192;
constexpr int kFbTile = 128 * 128;     // bytes of a [128 rows][64 bf16] tile
constexpr int kFbMaxYc = 4;

struct __align__(16) FbBars {
  uint64_t x_full;
  uint64_t yqk_full[2], yqk_empty[2];
  uint64_t yc_full[kFbMaxYc], yc_empty[kFbMaxYc];
  uint64_t s_full, s_empty;
  uint64_t dp_full[2], dp_empty[2];
  uint64_t p_full, p_empty;
  uint64_t o_full;
  uint32_t tmem_base;
  uint32_t pad;
};

// qk: [B*N, 128] bf16 = [q | k] (64 zero-padded dims each); tmDY / tmV: NHWC maps of the [B, N, C] tensors (own side
// and streamed side swap with the role); lse, dsum: [B*N] fp32; dqk: [B*N, 128] bf16 (dq | dk).
__global__ void __launch_bounds__(kFbThreads, 1)
flash_attn2d_bwd_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmDY,
                        const __grid_constant__ CUtensorMap tmV, const float* __restrict__ lse,
                        const float* __restrict__ dsum, const float* __restrict__ gamma,
                        __nv_bfloat16* __restrict__ dqk, int N, int C, int nbuf, int ycst) {
  // blockIdx.z selects the role, so both gradients share one launch (twice the CTAs: small N x B grids fill the GPU):
  //   z = 0: dQ - own = queries with dY resident, V streamed;   z = 1: dK - own = keys with V resident, dY streamed
  const int own_is_key = blockIdx.z;
  const CUtensorMap* tmXp = own_is_key ? &tmV : &tmDY;
  const CUtensorMap* tmYp = own_is_key ? &tmDY : &tmV;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nch = C / 64;
  uint8_t* sXqk = smem;                          // [128][64]
  uint8_t* sYqk = sXqk + kFbTile;                // nbuf x [128][64]
  uint8_t* sDS = sYqk + nbuf * kFbTile;          // 2 atoms x [128 own rows][64 streamed rows]
  uint8_t* sXc = sDS + 2 * kFbTile;              // nch x [128][64 ch]
  uint8_t* sYc = sXc + nch * kFbTile;            // ycst x [128][64 ch]
  FbBars* bars = reinterpret_cast<FbBars*>(sYc + ycst * kFbTile);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x, b = blockIdx.y;
  const int ntiles = N / 128;
  const int own_col = own_is_key ? 64 : 0, other_col = own_is_key ? 0 : 64;
  const uint32_t tmem_cols = 512;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQK);
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmV);
    mbar_init(&bars->x_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->yqk_full[i], 1);
      mbar_init(&bars->yqk_empty[i], 1);
      mbar_init(&bars->dp_full[i], 1);
      mbar_init(&bars->dp_empty[i], 4);
    }
    for (int i = 0; i < kFbMaxYc; ++i) {
      mbar_init(&bars->yc_full[i], 1);
      mbar_init(&bars->yc_empty[i], 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->s_empty, 4);
    mbar_init(&bars->p_full, 4);
    mbar_init(&bars->p_empty, 1);
    mbar_init(&bars->o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t tmem_s = tmem_base;
  const uint32_t tmem_dp = tmem_base + 128;       // two buffers of 128 columns
  const uint32_t tmem_o = tmem_base + 384;        // 64 columns

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const long row0 = (long)b * N;
      mbar_expect_tx(&bars->x_full, (1 + nch) * kFbTile);
      tma_load_2d(sXqk, &tmQK, &bars->x_full, own_col, (int)(row0 + t * 128));
      for (int ch = 0; ch < nch; ++ch) tma_load_4d(sXc + ch * kFbTile, tmXp, &bars->x_full, ch * 64, t * 128, 0, b);
      long g = 0;                                   // running chunk counter of the Yc ring
      for (int j = 0; j < ntiles; ++j) {
        // the order mirrors the MMA warp's (see there): with a single Yqk buffer the chunks of tile j must not wait
        // behind a Yqk slot that is only released by the previous tile's last MMAs
        for (int step = 0; step < 2; ++step) {
          if ((step == 0) == (nbuf == 2)) {
            const int ys = j % nbuf;
            mbar_wait(&bars->yqk_empty[ys], ((j / nbuf) & 1) ^ 1);
            mbar_expect_tx(&bars->yqk_full[ys], kFbTile);
            tma_load_2d(sYqk + ys * kFbTile, &tmQK, &bars->yqk_full[ys], other_col, (int)(row0 + j * 128));
          } else {
            for (int ch = 0; ch < nch; ++ch, ++g) {
              const int cs = (int)(g % ycst);
              mbar_wait(&bars->yc_empty[cs], ((g / ycst) & 1) ^ 1);
              mbar_expect_tx(&bars->yc_full[cs], kFbTile);
              tma_load_4d(sYc + cs * kFbTile, tmYp, &bars->yc_full[cs], ch * 64, j * 128, 0, b);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc_ss = make_idesc_bf16(128, 128, false, false);   // S', dP': both operands K-major
    const uint32_t idesc_o = make_idesc_bf16(128, 64, false, true);      // out += dS' Yqk: B N-major
    mbar_wait(&bars->x_full, 0);
    long g = 0;
    auto step_s = [&](int j) {                       // S'(j) = Xqk Yqk(j)^T
      const int ys = j % nbuf;
      mbar_wait(&bars->yqk_full[ys], (j / nbuf) & 1);
      mbar_wait(&bars->s_empty, (j & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(sXqk), 0, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sYqk + ys * kFbTile), 0, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, adesc + 2 * k, bdesc + 2 * k, idesc_ss, k != 0);
        umma_commit(&bars->s_full);
      }
      __syncwarp();
    };
    auto step_dp = [&](int j) {                      // dP'(j) = Xc Yc(j)^T over the channel chunks
      const int buf = j & 1;
      mbar_wait(&bars->dp_empty[buf], ((j >> 1) & 1) ^ 1);
      for (int ch = 0; ch < nch; ++ch, ++g) {
        const int cs = (int)(g % ycst);
        mbar_wait(&bars->yc_full[cs], (g / ycst) & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(sXc + ch * kFbTile), 0, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sYc + cs * kFbTile), 0, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_dp + buf * 128, adesc + 2 * k, bdesc + 2 * k, idesc_ss, (ch | k) != 0);
          umma_commit(&bars->yc_empty[cs]);
          if (ch == nch - 1) umma_commit(&bars->dp_full[buf]);
        }
        __syncwarp();
      }
    };
    auto step_out = [&](int j) {                     // out += dS'(j) Yqk(j)
      const int ys = j % nbuf;
      mbar_wait(&bars->p_full, j & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(sDS + (k >> 2) * kFbTile) + (k & 3) * 32, 0, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sYqk + ys * kFbTile) + k * 2048, kFbTile, 1024);
          umma_bf16(tmem_o, adesc, bdesc, idesc_o, (j | k) != 0);
        }
        umma_commit(&bars->yqk_empty[ys]);
        umma_commit(&bars->p_empty);
        if (j == ntiles - 1) umma_commit(&bars->o_full);
      }
      __syncwarp();
    };
    if (nbuf == 2) {
      step_s(0);
      step_dp(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) {
          step_s(j + 1);
          step_dp(j + 1);
        }
        step_out(j);
      }
    } else {
      step_dp(0);
      step_s(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) step_dp(j + 1);
        step_out(j);
        if (j + 1 < ntiles) step_s(j + 1);
      }
    }
  } else {
    // ------------------------------------------------------------ dS math + epilogue (warps 2..5)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const size_t own_row = (size_t)b * N + t * 128 + row;
    const float lse_r = own_is_key ? 0.f : lse[own_row];
    const float d_r = own_is_key ? 0.f : dsum[own_row];
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(&bars->s_full, j & 1);
      tc_fence_after();
      uint32_t r[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tmem_s + lane_addr + c * 32, r[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_empty);
      const float4* lse_t = reinterpret_cast<const float4*>(lse + (size_t)b * N + j * 128);
      const float4* d_t = reinterpret_cast<const float4*>(dsum + (size_t)b * N + j * 128);
      // P' = exp(S' - lse[query]) kept in the same registers
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          // exp(s - lse) = exp2(s * log2e - lse * log2e): one FFMA + one MUFU.EX2 per score
          constexpr float kLog2e = 1.4426950408889634f;
          float4 l = make_float4(lse_r, lse_r, lse_r, lse_r);
          if (own_is_key) l = __ldg(lse_t + c * 8 + i4);
          r[c][i4 * 4 + 0] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(r[c][i4 * 4 + 0]), kLog2e, -l.x * kLog2e)));
          r[c][i4 * 4 + 1] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(r[c][i4 * 4 + 1]), kLog2e, -l.y * kLog2e)));
          r[c][i4 * 4 + 2] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(r[c][i4 * 4 + 2]), kLog2e, -l.z * kLog2e)));
          r[c][i4 * 4 + 3] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(r[c][i4 * 4 + 3]), kLog2e, -l.w * kLog2e)));
        }
      const int buf = j & 1;
      mbar_wait(&bars->dp_full[buf], (j >> 1) & 1);
      tc_fence_after();
      mbar_wait(&bars->p_empty, (j & 1) ^ 1);          // previous dS' tile consumed by the out MMAs
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t d[32];
        tmem_ld32(tmem_dp + buf * 128 + lane_addr + c * 32, d);
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          float dd[8];
          if (own_is_key) {
            const float4 d0 = __ldg(d_t + c * 8 + g8 * 2), d1 = __ldg(d_t + c * 8 + g8 * 2 + 1);
            dd[0] = d0.x; dd[1] = d0.y; dd[2] = d0.z; dd[3] = d0.w;
            dd[4] = d1.x; dd[5] = d1.y; dd[6] = d1.z; dd[7] = d1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) dd[e] = d_r;
          }
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i0 = g8 * 8 + 2 * e;
            const float s0 = __uint_as_float(r[c][i0]) * (__uint_as_float(d[i0]) - dd[2 * e]);
            const float s1 = __uint_as_float(r[c][i0 + 1]) * (__uint_as_float(d[i0 + 1]) - dd[2 * e + 1]);
            pk[e] = pack_bf16x2(s0, s1);
          }
          const int col0 = c * 32 + g8 * 8;            // first streamed row of this 16-byte chunk
          const int atom = col0 >> 6, chunk = (col0 & 63) >> 3;
          const uint32_t addr = smem_u32(sDS + atom * kFbTile) + row * 128 + ((chunk ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                       "r"(pk[3])
                       : "memory");
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars->dp_empty[buf]);
        mbar_arrive(&bars->p_full);
      }
    }
    // epilogue: dq | dk rows, scaled by gamma
    mbar_wait(&bars->o_full, 0);
    tc_fence_after();
    const float gm = gamma[0];
    __nv_bfloat16* dst = dqk + own_row * 128 + own_col;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t o[32];
      tmem_ld32(tmem_o + lane_addr + c0, o);
      tmem_ld_wait();
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          w[e] = pack_bf16x2(gm * __uint_as_float(o[v4 * 8 + 2 * e]), gm * __uint_as_float(o[v4 * 8 + 2 * e + 1]));
        reinterpret_cast<uint4*>(dst + c0)[v4] = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// D[row] = sum_c dY[row,c] * O[row,c] (fp32) and dgamma += sum of all D. One warp per row, 16-byte loads.
__global__ void __launch_bounds__(256)
attn2d_bwd_prep_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ o,
                       float* __restrict__ dsum, float* __restrict__ dgamma, long rows, int C) {
  __shared__ float part[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float block_acc = 0.f;
  for (long row = (long)blockIdx.x * 8 + warp; row < rows; row += (long)gridDim.x * 8) {
    const uint4* a = reinterpret_cast<const uint4*>(dy + row * C);
    const uint4* bq = reinterpret_cast<const uint4*>(o + row * C);
    float acc = 0.f;
    for (int i = lane; i < C / 8; i += 32) {
      const uint4 u = __ldg(a + i), w = __ldg(bq + i);
      const uint32_t uu[4] = {u.x, u.y, u.z, u.w}, ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc += __uint_as_float(uu[e] << 16) * __uint_as_float(ww[e] << 16);
        acc += __uint_as_float(uu[e] & 0xffff0000u) * __uint_as_float(ww[e] & 0xffff0000u);
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) dsum[row] = acc;
    block_acc += acc;
  }
  if (lane == 0) part[warp] = block_acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += part[i];
    atomicAdd(dgamma, s);
  }
}

}  // namespace lun

using namespace lun;

extern "C" {

int lun_flash_attn2d_bwd_prep_bf16(const void* dy, const void* o, float* dsum, float* dgamma, long rows, int C,
                                   void* stream) {
  if (C % 8) return LUN_E_SHAPE;
  long blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  attn2d_bwd_prep_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)dy,
                                                                        (const __nv_bfloat16*)o, dsum, dgamma, rows, C);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_flash_attn2d_dqk_bf16(const void* qk, const void* v, const void* dy, const float* lse, const float* dsum,
                              const float* gamma, void* dqk, int B, int N, int C, void* stream) {
  if (N % 128 || C % 64 || C > 512) return LUN_E_SHAPE;
  CUtensorMap tmQK, tmV, tmDY;
  int rc = make_tmap_2d(&tmQK, qk, (long)B * N, 128, 128);
  if (rc) return rc;
  rc = make_tmap_nhwc(&tmV, v, B, 1, N, C, 128, 1, 1, 1);
  if (rc) return rc;
  rc = make_tmap_nhwc(&tmDY, dy, B, 1, N, C, 128, 1, 1, 1);
  if (rc) return rc;
  const int nch = C / 64;
  const int budget = 227 * 1024 - 1024 - (int)sizeof(FbBars);
  // Xqk + dS' (2 atoms) + resident Xc, then as many Yqk buffers (<= 2) and Yc ring slots (2..4) as fit
  int left = budget / kFbTile - (1 + 2 + nch);
  const int nbuf = left >= 4 ? 2 : 1;
  int ycst = left - nbuf;
  if (ycst > kFbMaxYc) ycst = kFbMaxYc;
  if (ycst < 2) return LUN_E_SHAPE;
  const int smem = (1 + nbuf + 2 + nch + ycst) * kFbTile + (int)sizeof(FbBars) + 1024;
  static bool configured_dev[64] = {};   // the opt-in smem size is a per-device function attribute
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(flash_attn2d_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess)
      return LUN_E_ATTR;
    configured = true;
  }
  dim3 grid(N / 128, B, 2);                  // z: dQ / dK
  flash_attn2d_bwd_kernel<<<grid, kFbThreads, smem, (cudaStream_t)stream>>>(
      tmQK, tmDY, tmV, lse, dsum, gamma, (__nv_bfloat16*)dqk, N, C, nbuf, ycst);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

}  // extern "C"
