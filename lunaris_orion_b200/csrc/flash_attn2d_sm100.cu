// Global spatial self-attention for SelfAttention2d (lunar_generate.py:56-78) as a flash-style tcgen05 kernel:
//   energy[i,j] = q_i . k_j (no scaling), attention = softmax_j, out_i = sum_j attention[i,j] v_j,  y = gamma*out + x
// with q,k of width DQK = C/8 (zero-padded to a multiple of 16) and v of width C. The N x N matrix never exists.
//
// One CTA owns 128 queries and a slice of <= 256 value channels. Because DQK is 8x narrower than the value width,
// Q K^T is cheap next to P V, so the softmax is done in TWO passes instead of rescaling the TMEM accumulator:
//   pass A: S = Q K_j^T per 128-key tile (tcgen05, TMEM) -> running row max / row sum in registers
//   pass B: S again, P = exp(S - max) / sum -> bf16 -> 128B-swizzled smem tile -> O += P V_j (tcgen05, O in TMEM)
// Warp roles: warp 0 = TMA producer (Q once, K / V tiles through mbarrier rings), warp 1 = MMA issuer,
// warps 2..5 = softmax + epilogue (each thread owns one query row: TMEM lane == row).
// K tiles are [128 keys][64 (padded) dims] K-major; V tiles are [128 keys][64-channel atoms] N-major (MN-major B
// operand straight from the NHWC tensor); P is the K-major A operand of the second GEMM.
//
// The same kernel computes the VALUE GRADIENT of the backward pass (lse_in != nullptr): the CTA owns 128 KEYS, streams
// the query tiles, rebuilds P^T[key, query] = exp(k.q - lse[query]) in one pass (the row statistics were saved by the
// forward) and accumulates dV = gamma * P^T dY with dY in the place of V. The query / key gradients are in
// flash_attn2d_bwd_sm100.cu.
#include "../../include/lunaris_b200.h"
#include "conv_gemm.cuh"
#include "launch_count.cuh"
#include "ptx.cuh"

namespace lun {

int make_tmap_2d(CUtensorMap* m, const void* base, long rows, long cols, int box_rows);
int make_tmap_nhwc(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h, int box_b,
                   int estride);

constexpr int kFaThreads = 192;
constexpr int kFaTileBytes = 128 * 128;   // 128 rows x 64 bf16

struct __align__(16) FaBars {
  uint64_t q_full;
  uint64_t k_full[2], k_empty[2];
  uint64_t v_full[2], v_empty[2];
  uint64_t s_full, s_empty;      // S accumulator (TMEM) MMA -> softmax -> MMA
  uint64_t p_full, p_empty;      // P tile (smem) softmax -> MMA
  uint64_t o_full;               // all PV MMAs done
  uint32_t tmem_base;
  uint32_t pad;
};

// qk: [B*N, 128] bf16 rows = [q (64, zero padded) | k (64, zero padded)];  v: [B, N, C] bf16;  x, y: [B, N, C] bf16
__global__ void __launch_bounds__(kFaThreads, 1)
flash_attn2d_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV,
                    const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ gamma, int N, int C,
                    int vw, const float* __restrict__ lse_in, __nv_bfloat16* __restrict__ o_out,
                    float* __restrict__ lse_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int vatoms = vw / 64;                       // 64-channel atoms of the value slice
  uint8_t* sQ = smem;                               // [128][64]
  uint8_t* sK = sQ + kFaTileBytes;                  // 2 x [128][64]
  uint8_t* sP = sK + 2 * kFaTileBytes;              // 2 atoms x [128 rows][64 keys]  (K-major A operand)
  uint8_t* sV = sP + 2 * kFaTileBytes;              // 2 x vatoms x [128 keys][64 ch] (N-major B operand)
  FaBars* bars = reinterpret_cast<FaBars*>(sV + 2 * vatoms * kFaTileBytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, b = blockIdx.y, vs = blockIdx.z;
  const int ntiles = N / 128;
  const uint32_t tmem_cols = 512;
  const bool dv_mode = lse_in != nullptr;           // backward: own tile = keys (qk columns 64..127), streamed = queries
  const int first_pass = dv_mode ? 1 : 0;
  const int own_col = dv_mode ? 64 : 0, other_col = dv_mode ? 0 : 64;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQK);
    prefetch_tmap(&tmV);
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], 1);
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
  const uint32_t tmem_s = tmem_base;                // 128 columns: S
  const uint32_t tmem_o = tmem_base + 128;          // vw (<= 256) columns: O

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const long row0 = (long)b * N;
      mbar_expect_tx(&bars->q_full, kFaTileBytes);
      tma_load_2d(sQ, &tmQK, &bars->q_full, own_col, (int)(row0 + qt * 128));
      int ks = 0, vsx = 0;
      uint32_t kph = 0, vph = 0;
      for (int pass = first_pass; pass < 2; ++pass) {
        for (int j = 0; j < ntiles; ++j) {
          mbar_wait(&bars->k_empty[ks], kph ^ 1);
          mbar_expect_tx(&bars->k_full[ks], kFaTileBytes);
          tma_load_2d(sK + ks * kFaTileBytes, &tmQK, &bars->k_full[ks], other_col, (int)(row0 + j * 128));
          if (++ks == 2) { ks = 0; kph ^= 1; }
          if (pass == 1) {
            mbar_wait(&bars->v_empty[vsx], vph ^ 1);
            mbar_expect_tx(&bars->v_full[vsx], vatoms * kFaTileBytes);
            for (int a = 0; a < vatoms; ++a)
              tma_load_4d(sV + (vsx * vatoms + a) * kFaTileBytes, &tmV, &bars->v_full[vsx], vs * vw + a * 64, j * 128,
                          0, b);
            if (++vsx == 2) { vsx = 0; vph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc_s = make_idesc_bf16(128, 128, false, false);   // S = Q K^T      (both K-major)
    const uint32_t idesc_o = make_idesc_bf16(128, vw, false, true);     // O += P V       (A K-major, B N-major)
    mbar_wait(&bars->q_full, 0);
    int ks = 0, vsx = 0;
    uint32_t kph = 0, vph = 0, sph = 0, pph = 0;
    for (int pass = first_pass; pass < 2; ++pass) {
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(&bars->k_full[ks], kph);
        mbar_wait(&bars->s_empty, sph ^ 1);          // softmax warps finished reading the previous S
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(sQ), 0, 1024);
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sK + ks * kFaTileBytes), 0, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_s, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
          umma_commit(&bars->k_empty[ks]);
          umma_commit(&bars->s_full);
        }
        __syncwarp();
        if (++ks == 2) { ks = 0; kph ^= 1; }
        sph ^= 1;
        if (pass == 1) {
          mbar_wait(&bars->v_full[vsx], vph);
          mbar_wait(&bars->p_full, pph);              // P tile written by the softmax warps
          tc_fence_after();
          if (elect_one()) {
            // A = P: two K-major atoms of 64 keys (8 KB... 16 KB apart); B = V: N-major atoms 16 KB apart
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint64_t adesc = make_smem_desc_sw128(smem_u32(sP + (k >> 2) * kFaTileBytes) + (k & 3) * 32, 0, 1024);
              const uint64_t bdesc =
                  make_smem_desc_sw128(smem_u32(sV + vsx * vatoms * kFaTileBytes) + k * 2048, kFaTileBytes, 1024);
              umma_bf16(tmem_o, adesc, bdesc, idesc_o, (j | k) != 0);
            }
            umma_commit(&bars->v_empty[vsx]);
            umma_commit(&bars->p_empty);
            if (j == ntiles - 1) umma_commit(&bars->o_full);
          }
          __syncwarp();
          if (++vsx == 2) { vsx = 0; vph ^= 1; }
          pph ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax + epilogue (warps 2..5)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float m = -3.0e38f, l = 0.f;
    uint32_t sph = 0, peph = 0;
    // pass A: row max and row sum (the backward reads the saved statistics instead)
    for (int j = 0; j < (dv_mode ? 0 : ntiles); ++j) {
      mbar_wait(&bars->s_full, sph);
      sph ^= 1;
      tc_fence_after();
      float tmax = -3.0e38f;
      uint32_t r[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tmem_s + lane_addr + c * 32, r[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_empty);
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) tmax = fmaxf(tmax, __uint_as_float(r[c][i]));
      const float mn = fmaxf(m, tmax);
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) s += __expf(__uint_as_float(r[c][i]) - mn);
      l = l * __expf(m - mn) + s;
      m = mn;
    }
    const float inv_l = dv_mode ? 1.f : 1.f / l;
    // pass B: P tiles
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(&bars->s_full, sph);
      sph ^= 1;
      tc_fence_after();
      uint32_t r[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(tmem_s + lane_addr + c * 32, r[c]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_empty);
      mbar_wait(&bars->p_empty, peph ^ 1);           // previous P consumed by the PV MMAs
      peph ^= 1;
      // row `row` of P: 128 keys -> two 64-key atoms, 8 chunks of 16 bytes each, chunk index swizzled by row & 7
      const float4* lse_t = reinterpret_cast<const float4*>(lse_in + (size_t)b * N + j * 128);   // dv_mode only
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t pk[4];
          float sub[8];
          if (dv_mode) {
            const float4 l0 = __ldg(lse_t + c * 8 + g * 2), l1 = __ldg(lse_t + c * 8 + g * 2 + 1);
            sub[0] = l0.x; sub[1] = l0.y; sub[2] = l0.z; sub[3] = l0.w;
            sub[4] = l1.x; sub[5] = l1.y; sub[6] = l1.z; sub[7] = l1.w;
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) sub[e] = m;
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 = __expf(__uint_as_float(r[c][g * 8 + 2 * e]) - sub[2 * e]) * inv_l;
            const float p1 = __expf(__uint_as_float(r[c][g * 8 + 2 * e + 1]) - sub[2 * e + 1]) * inv_l;
            pk[e] = pack_bf16x2(p0, p1);
          }
          const int key0 = c * 32 + g * 8;             // first key of this 16-byte chunk
          const int atom = key0 >> 6, chunk = (key0 & 63) >> 3;
          const uint32_t addr = smem_u32(sP + atom * kFaTileBytes) + row * 128 + ((chunk ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
                       "r"(pk[3])
                       : "memory");
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);
    }
    // epilogue: y = gamma * O + x  (backward: dV = gamma * O); training forward also keeps O and logsumexp
    mbar_wait(&bars->o_full, 0);
    tc_fence_after();
    const float gm = gamma[0];
    const size_t base = ((size_t)b * N + qt * 128 + row) * C + vs * vw;
    if (lse_out != nullptr && vs == 0) lse_out[(size_t)b * N + qt * 128 + row] = m + __logf(l);
    for (int c0 = 0; c0 < vw; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_o + lane_addr + c0, r);
      tmem_ld_wait();
      const uint4* xs = reinterpret_cast<const uint4*>(x + base + c0);
      uint4* ys = reinterpret_cast<uint4*>(y + base + c0);
#pragma unroll
      for (int v4 = 0; v4 < 4; ++v4) {
        uint4 xv = make_uint4(0u, 0u, 0u, 0u);
        if (!dv_mode) xv = __ldg(xs + v4);
        const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float x0 = __uint_as_float(xw[e] << 16), x1 = __uint_as_float(xw[e] & 0xffff0000u);
          o[e] = pack_bf16x2(gm * __uint_as_float(r[v4 * 8 + 2 * e]) + x0, gm * __uint_as_float(r[v4 * 8 + 2 * e + 1]) + x1);
        }
        ys[v4] = make_uint4(o[0], o[1], o[2], o[3]);
        if (o_out != nullptr) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = pack_bf16x2(__uint_as_float(r[v4 * 8 + 2 * e]), __uint_as_float(r[v4 * 8 + 2 * e + 1]));
          reinterpret_cast<uint4*>(o_out + base + c0)[v4] = make_uint4(o[0], o[1], o[2], o[3]);
        }
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

}  // namespace lun

using namespace lun;

static int launch_flash(const void* qk, const void* v, const void* x, void* y, const float* gamma, int B, int N, int C,
                        const float* lse_in, void* o_out, float* lse_out, void* stream) {
  if (N % 128 || C % 64 || C > 512) return LUN_E_SHAPE;
  const int vw = C > 256 ? 256 : C;
  if (C % vw) return LUN_E_SHAPE;
  CUtensorMap tmQK, tmV;
  int rc = make_tmap_2d(&tmQK, qk, (long)B * N, 128, 128);
  if (rc) return rc;
  rc = make_tmap_nhwc(&tmV, v, B, 1, N, C, 128, 1, 1, 1);     // {C, W=N, H=1, B}: boxes of 64 ch x 128 keys
  if (rc) return rc;
  const int vatoms = vw / 64;
  const int smem = (1 + 2 + 2 + 2 * vatoms) * kFaTileBytes + (int)sizeof(FaBars) + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(flash_attn2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
        cudaSuccess)
      return LUN_E_ATTR;
    configured = true;
  }
  dim3 grid(N / 128, B, C / vw);
  flash_attn2d_kernel<<<grid, kFaThreads, smem, (cudaStream_t)stream>>>(
      tmQK, tmV, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, gamma, N, C, vw, lse_in, (__nv_bfloat16*)o_out, lse_out);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

extern "C" {

int lun_flash_attn2d_bf16(const void* qk, const void* v, const void* x, void* y, const float* gamma, int B, int N,
                          int C, void* o_out, float* lse_out, void* stream) {
  return launch_flash(qk, v, x, y, gamma, B, N, C, nullptr, o_out, lse_out, stream);
}

int lun_flash_attn2d_dv_bf16(const void* qk, const void* dy, const float* lse, const float* gamma, void* dv, int B,
                             int N, int C, void* stream) {
  if (lse == nullptr) return LUN_E_SHAPE;
  return launch_flash(qk, dy, nullptr, dv, gamma, B, N, C, lse, nullptr, nullptr, stream);
}

}  // extern "C"
