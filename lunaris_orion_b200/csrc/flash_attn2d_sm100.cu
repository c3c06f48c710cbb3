// Global spatial self-attention for SelfAttention2d (lunar_generate.py:56-78) as a flash-style tcgen05 kernel:
//   energy[i,j] = q_i . k_j (no scaling), attention = softmax_j, out_i = sum_j attention[i,j] v_j,  y = gamma*out + x
// with q,k of width DQK = C/8 (zero-padded to 64) and v of width C. The N x N matrix never exists.
//
// One CTA owns 128 queries and a slice of <= 256 value channels. Q K^T is 8x cheaper than P V, so the row maximum is
// found in a first pass over the keys (S only - no exponentials) instead of rescaling the TMEM accumulator:
//   pass A: S = Q K_j^T per 128-key tile (tcgen05, TMEM)  -> running row max in registers; four S buffers (the O
//           columns are still free), buffer i read by the four softmax warps with column quarter i
//   pass B: S again, P = exp(S - max) UNNORMALISED -> bf16 -> 128B-swizzled smem tile -> O += P V_j (O in TMEM),
//           row sums of P accumulated alongside; the epilogue divides O by them.
// The SFU does one exponential per score (pass B only) and everything overlaps: in pass B S is double buffered, so the
// MMA warp computes S_{j+1} and runs P_j V_j while the 16 softmax warps (four per TMEM lane quarter, 32 score columns
// each) are still exponentiating.
// Warp roles: warp 0 = TMA producer (Q once, K / V tiles through mbarrier rings), warp 1 = MMA issuer,
// warps 2..17 = softmax + epilogue (thread = query row = TMEM lane; the four warps of a quarter split the columns).
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
#include <stdlib.h>

namespace lun {

int make_tmap_2d(CUtensorMap* m, const void* base, long rows, long cols, int box_rows);
int make_tmap_nhwc(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h, int box_b,
                   int estride);

constexpr int kFaThreads = 576;           // TMA warp, MMA warp, 16 softmax warps
constexpr int kFaTileBytes = 128 * 128;   // 128 rows x 64 bf16
// K ring depth. Pass A (row maxima: S only, 256 tensor cycles per tile) is bound by the ring's round trip
// (S done -> k_empty -> TMA issue -> L2 latency -> k_full -> S issue, ~1300 cycles): two stages gave 690 cycles per tile.
// A CTA of a pair holds half a K tile (8 KB) and has smem to spare; a single CTA with a 256-channel slice does not.
__host__ __device__ constexpr int kFaKStages(int cg) { return cg == 2 ? 4 : 2; }
// P stages. With one stage the softmax warps of tile j+1 hold their finished probabilities in registers until the PV
// MMAs of tile j have drained the tile (11 % of the kernel's stall samples). A second stage (a CTA of a pair has the
// 32 KB) was measured and is OFF: forward 124.5 -> 130.8 us, tensor pipe 60.0 -> 57.3 % of active cycles at C=512,
// N=9472 (profiles/r01d_flash_attn2d.txt); the per-atom hand-off already hides most of the drain. Cause not isolated.
constexpr int kFaPairPStages = 1;         // 2 = the measured variant
__host__ __device__ constexpr int kFaPStages(int cg) { return cg == 2 ? kFaPairPStages : 1; }

// S buffers in TMEM. Pass B double-buffers S in columns 0..255 (O lives in 256..511) and all 16 softmax warps work on
// every tile. Pass A (row maxima only, no O yet) rotates through FOUR 128-column buffers (columns 0..511) and buffer i
// belongs to the four warps with column quarter i, which read all 128 columns of their lane quarter: a warp then sees
// every fourth tile, so its serial wait -> tcgen05.ld -> arrive latency (~500 cycles, twice the 256 tensor cycles of
// an S tile) no longer paces the pass. The first PV MMA overwrites O only after every softmax warp has left pass A
// (p_full is signalled after the named barrier that follows pass A), so the aliasing is safe.

struct __align__(16) FaBars {
  uint64_t q_full;
  uint64_t k_full[4], k_empty[4];   // K ring: kFaKStages(CG) stages
  uint64_t v_full[2], v_empty[2];
  uint64_t s_full[2], s_empty[2];   // pass B: S accumulators (TMEM) MMA -> all 16 softmax warps -> MMA
  uint64_t sa_full[4], sa_empty[4]; // pass A: four S buffers, buffer i read by the four warps with column quarter i
  uint64_t p_full[4], p_empty[4];   // P tile (smem) softmax -> MMA, one pair per 64-key atom (= per softmax warp half)
                                    // and per P stage (kFaPStages): index = stage * 2 + atom
  uint64_t o_full;                  // all PV MMAs done
  uint32_t tmem_base;
  uint32_t pad;
  float xch[4][128];                // row statistics exchanged between the four warps of a lane quarter
};

// qk: [B*N, 128] bf16 rows = [q (64, zero padded) | k (64, zero padded)];  v: [B, N, C] bf16;  x, y: [B, N, C] bf16
// CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) owns 256 queries. Each CTA keeps its own 128 query rows, P tile
// and accumulators, but loads only HALF of every streamed tile (64 of the 128 keys of K, half of V's channels): the
// kernel is L2->SM bound with one CTA per 128 queries (7.8 TB/s), and pairing halves the operand bytes per FLOP.
// tmK: the qk matrix with 128/CG-row boxes (the streamed K / Q tiles).
template <int CG>
__global__ void __launch_bounds__(kFaThreads, 1)
flash_attn2d_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV,
                    const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ gamma, int N, int C,
                    int vw, const float* __restrict__ lse_in, __nv_bfloat16* __restrict__ o_out,
                    float* __restrict__ lse_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int vatoms = vw / 64 / CG;                  // 64-channel atoms of the value slice held by THIS CTA
  constexpr int kKBytes = kFaTileBytes / CG;        // this CTA's share of a streamed K tile: 128/CG keys
  uint8_t* sQ = smem;                               // [128][64]
  constexpr int KST = kFaKStages(CG);
  uint8_t* sK = sQ + kFaTileBytes;                  // KST x [128/CG][64]
  constexpr int PST = kFaPStages(CG);
  uint8_t* sP = sK + KST * kKBytes;                 // PST x 2 atoms x [128 rows][64 keys]  (K-major A operand)
  uint8_t* sV = sP + PST * 2 * kFaTileBytes;             // 2 x vatoms x [128 keys][64 ch] (N-major B operand)
  FaBars* bars = reinterpret_cast<FaBars*>(sV + 2 * vatoms * kFaTileBytes);
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, b = blockIdx.y, vs = blockIdx.z;
  const int ntiles = N / 128;
  const uint32_t tmem_cols = 512;
  const bool dv_mode = lse_in != nullptr;           // backward: own tile = keys (qk columns 64..127), streamed = queries
  const int own_col = dv_mode ? 64 : 0, other_col = dv_mode ? 0 : 64;
  const int steps = dv_mode ? ntiles : 2 * ntiles;  // S tiles computed: pass A (forward only) then pass B
  const int first_b = steps - ntiles;               // first step of pass B

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmQK);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], 1);
      mbar_init(&bars->sa_full[i], 1);
      mbar_init(&bars->sa_empty[i], 4 * CG);
      mbar_init(&bars->p_full[i], 8 * CG);
      mbar_init(&bars->p_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->s_empty[i], 16 * CG);       // softmax warps of both CTAs report to the leader
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], 1);
    }
    mbar_init(&bars->o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc_pair(&bars->tmem_base, tmem_cols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(&bars->tmem_base, tmem_cols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();                  // peer barriers initialised before any remote arrive / TMA signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const uint32_t tmem_s = tmem_base;                // 128-column S buffers: 4 in pass A (0..511), 2 in pass B (0..255)
  const uint32_t tmem_o = tmem_base + 256;          // vw (<= 256) columns: O

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const long row0 = (long)b * N;
      // pair mode: the full barriers live in the leader CTA; it announces the bytes of both CTAs
      if (cta_rank == 0) mbar_expect_tx(&bars->q_full, CG * kFaTileBytes);
      if (CG == 2) tma_load_2d_pair(sQ, &tmQK, &bars->q_full, own_col, (int)(row0 + qt * 128));
      else tma_load_2d(sQ, &tmQK, &bars->q_full, own_col, (int)(row0 + qt * 128));
      for (int t = 0; t < steps; ++t) {
        const int j = t < first_b ? t : t - first_b;
        const int ks = t % KST;
        mbar_wait(&bars->k_empty[ks], ((t / KST) & 1) ^ 1);
        if (cta_rank == 0) mbar_expect_tx(&bars->k_full[ks], kFaTileBytes);
        const int krow = (int)(row0 + j * 128 + cta_rank * (128 / CG));           // this CTA's 128/CG keys
        if (CG == 2) tma_load_2d_pair(sK + ks * kKBytes, &tmK, &bars->k_full[ks], other_col, krow);
        else tma_load_2d(sK + ks * kKBytes, &tmK, &bars->k_full[ks], other_col, krow);
        if (t >= first_b) {
          const int vsx = j & 1;
          mbar_wait(&bars->v_empty[vsx], ((j >> 1) & 1) ^ 1);
          if (cta_rank == 0) mbar_expect_tx(&bars->v_full[vsx], CG * vatoms * kFaTileBytes);
          for (int a = 0; a < vatoms; ++a) {
            const int ch = vs * vw + (cta_rank * vatoms + a) * 64;                // this CTA's channels of the slice
            if (CG == 2) tma_load_4d_pair(sV + (vsx * vatoms + a) * kFaTileBytes, &tmV, &bars->v_full[vsx], ch, j * 128, 0, b);
            else tma_load_4d(sV + (vsx * vatoms + a) * kFaTileBytes, &tmV, &bars->v_full[vsx], ch, j * 128, 0, b);
          }
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ------------------------------------------------------------ MMA issuer (the leader CTA of a pair)
    const uint32_t idesc_s = make_idesc_bf16(128 * CG, 128, false, false);   // S = Q K^T      (both K-major)
    const uint32_t idesc_o = make_idesc_bf16(128 * CG, vw, false, true);     // O += P V       (A K-major, B N-major)
    auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accumulate) {
      if (CG == 2) umma_bf16_pair(d, ad, bd, idesc, accumulate);
      else umma_bf16(d, ad, bd, idesc, accumulate);
    };
    auto commit = [&](uint64_t* bar) {
      if (CG == 2) umma_commit_pair(bar);
      else umma_commit(bar);
    };
    mbar_wait(&bars->q_full, 0);
    auto issue_s = [&](int t) {                      // S_t into its TMEM buffer
      const int ks = t % KST;
      const bool pa = t < first_b;
      const int j = t - first_b;
      const int sb = pa ? (t & 3) : (j & 1);
      uint64_t* full = pa ? &bars->sa_full[sb] : &bars->s_full[sb];
      mbar_wait(&bars->k_full[ks], (t / KST) & 1);
      // the softmax warps finished reading this buffer
      if (pa) {
        mbar_wait(&bars->sa_empty[sb], ((t >> 2) & 1) ^ 1);
      } else {
        if (j < 2 && first_b > j)                            // last pass-A use of the TMEM columns of buffer j
          mbar_wait(&bars->sa_empty[j], (((first_b + 3 - j) >> 2) - 1) & 1);
        mbar_wait(&bars->s_empty[sb], ((j >> 1) & 1) ^ 1);
      }
      tc_fence_after();
      if (elect_one()) {
        const uint64_t adesc = make_smem_desc_sw128(smem_u32(sQ), 0, 1024);
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sK + ks * kKBytes), 0, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma(tmem_s + sb * 128, adesc + 2 * k, bdesc + 2 * k, idesc_s, k != 0);
        commit(&bars->k_empty[ks]);
        commit(full);
      }
      __syncwarp();
    };
    issue_s(0);
    for (int t = 0; t < steps; ++t) {
      if (t + 1 < steps) issue_s(t + 1);             // runs under the softmax of tile t
      if (t >= first_b) {
        const int j = t - first_b, vsx = j & 1;
        mbar_wait(&bars->v_full[vsx], (j >> 1) & 1);
        // The P hand-off is per 64-key atom: while the MMAs of atom 1 run, the warps of atom 0 already store the next
        // tile's probabilities, so the tensor pipe does not idle across the store -> fence -> barrier round trip.
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int pa = (j % PST) * 2 + a;          // P stage and atom
          mbar_wait(&bars->p_full[pa], (j / PST) & 1);   // this atom of P written by its four softmax warps
          tc_fence_after();
          if (elect_one()) {
            // A = P atom a: K-major, 64 keys; B = V rows 64a..64a+63: N-major atoms 16 KB apart
#pragma unroll
            for (int k = 4 * a; k < 4 * a + 4; ++k) {
              const uint64_t adesc = make_smem_desc_sw128(smem_u32(sP + pa * kFaTileBytes) + (k & 3) * 32, 0, 1024);
              const uint64_t bdesc =
                  make_smem_desc_sw128(smem_u32(sV + vsx * vatoms * kFaTileBytes) + k * 2048, kFaTileBytes, 1024);
              mma(tmem_o, adesc, bdesc, idesc_o, (j | k) != 0);
            }
            commit(&bars->p_empty[pa]);
            if (a == 1) {
              commit(&bars->v_empty[vsx]);
              if (j == ntiles - 1) commit(&bars->o_full);
            }
          }
          __syncwarp();
        }
        __syncwarp();
      }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ softmax + epilogue (warps 2..17)
    auto arrive = [&](uint64_t* bar) {              // barriers the MMA issuer waits on live in the leader CTA
      if (CG == 2) mbar_arrive_cluster(bar, 0);
      else mbar_arrive(bar);
    };
    const int q = warp & 3;                         // TMEM lane quarter
    const int cq = (warp - 2) >> 2;                 // column quarter (32 of the 128 scores of a row) of the S tile
    const int hf = cq >> 1;                         // P atom (64 keys) this warp writes into
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float m = -3.0e38f, l = 0.f;
    // pass A: row maxima (no exponentials). This warp owns S buffer cq = every fourth key tile, all 128 columns.
    {
      float m0 = m, m1 = m, m2 = m, m3 = m;
      for (int t = cq; t < first_b; t += 4) {
        mbar_wait(&bars->sa_full[cq], (t >> 2) & 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t r0[32], r1[32];
          tmem_ld32(tmem_s + cq * 128 + lane_addr + h * 64, r0);
          tmem_ld32(tmem_s + cq * 128 + lane_addr + h * 64 + 32, r1);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            m0 = fmaxf(m0, __uint_as_float(r0[i]));
            m1 = fmaxf(m1, __uint_as_float(r0[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(r1[i]));
            m3 = fmaxf(m3, __uint_as_float(r1[i + 1]));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(&bars->sa_empty[cq]);
      }
      m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
    }
    if (!dv_mode) {                                 // combine the four column quarters of every row
      bars->xch[cq][row] = m;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      m = fmaxf(fmaxf(bars->xch[0][row], bars->xch[1][row]), fmaxf(bars->xch[2][row], bars->xch[3][row]));
      asm volatile("bar.sync 1, 512;" ::: "memory");   // xch is reused for the row sums
    }
    // pass B: P tiles
    for (int j = 0; j < ntiles; ++j) {
      const int sb = j & 1;
      mbar_wait(&bars->s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(tmem_s + sb * 128 + lane_addr + cq * 32, r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive(&bars->s_empty[sb]);
      // this thread's 32 keys = half of atom hf of the P tile: 4 chunks of 16 bytes, chunk index swizzled by row & 7
      const float4* lse_t = reinterpret_cast<const float4*>(lse_in + (size_t)b * N + j * 128 + cq * 32);   // dv_mode only
      // exp(s - m) = exp2(s * log2e - m * log2e): one FFMA + one MUFU.EX2 per score
      // (Moving 25 % of the exponentials to the FMA pipe - round-to-int by magic add, degree-3 minimax 2^f, exponent
      // splice: ~8 instructions per score - was measured and dropped: forward 348.5 -> 363.2 us at C=512, N=16384;
      // the softmax warps are issue-bound, not MUFU-bound. profiles/r01d_flash_attn2d.txt)
      constexpr float kLog2e = 1.4426950408889634f;
      const float m2 = m * kLog2e;
      uint32_t pk[16];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float sub[8];
        if (dv_mode) {
          const float4 l0 = __ldg(lse_t + g * 2), l1 = __ldg(lse_t + g * 2 + 1);
          sub[0] = l0.x * kLog2e; sub[1] = l0.y * kLog2e; sub[2] = l0.z * kLog2e; sub[3] = l0.w * kLog2e;
          sub[4] = l1.x * kLog2e; sub[5] = l1.y * kLog2e; sub[6] = l1.z * kLog2e; sub[7] = l1.w * kLog2e;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) sub[e] = m2;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(r[g * 8 + 2 * e]), kLog2e, -sub[2 * e]));
          const float p1 = fast_exp2(fmaf(__uint_as_float(r[g * 8 + 2 * e + 1]), kLog2e, -sub[2 * e + 1]));
          pk[g * 4 + e] = pack_bf16x2(p0, p1);
          l += p0 + p1;                                // row sum (the bf16 rounding of P averages out over 10^3+ keys)
        }
      }
      const int pa = (j % PST) * 2 + hf;             // P stage and atom
      mbar_wait(&bars->p_empty[pa], ((j / PST) & 1) ^ 1);   // the previous P in this stage consumed by the PV MMAs
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int chunk = (cq & 1) * 4 + g;            // 16-byte chunk (8 keys) inside the 64-key atom
        const uint32_t addr = smem_u32(sP + pa * kFaTileBytes) + row * 128 + ((chunk ^ (row & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[g * 4]), "r"(pk[g * 4 + 1]),
                     "r"(pk[g * 4 + 2]), "r"(pk[g * 4 + 3])
                     : "memory");
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) arrive(&bars->p_full[pa]);
    }
    // epilogue: O / rowsum, y = gamma * O + x  (backward: dV = gamma * O); training forward also keeps O and logsumexp
    float inv_l = 1.f;
    if (!dv_mode) {
      bars->xch[cq][row] = l;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      l = (bars->xch[0][row] + bars->xch[1][row]) + (bars->xch[2][row] + bars->xch[3][row]);
      inv_l = 1.f / l;
    }
    mbar_wait(&bars->o_full, 0);
    tc_fence_after();
    const float gm = gamma[0];
    const size_t base = ((size_t)b * N + qt * 128 + row) * C + vs * vw;
    if (lse_out != nullptr && vs == 0 && cq == 0) lse_out[(size_t)b * N + qt * 128 + row] = m + __logf(l);
    for (int c0 = cq * 32; c0 < vw; c0 += 128) {    // 32-column chunks of the O slice, dealt round-robin to the 4 warps
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
          const float o0 = __uint_as_float(r[v4 * 8 + 2 * e]) * inv_l, o1 = __uint_as_float(r[v4 * 8 + 2 * e + 1]) * inv_l;
          o[e] = pack_bf16x2(gm * o0 + x0, gm * o1 + x1);
        }
        ys[v4] = make_uint4(o[0], o[1], o[2], o[3]);
        if (o_out != nullptr) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            o[e] = pack_bf16x2(__uint_as_float(r[v4 * 8 + 2 * e]) * inv_l, __uint_as_float(r[v4 * 8 + 2 * e + 1]) * inv_l);
          reinterpret_cast<uint4*>(o_out + base + c0)[v4] = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_pair(tmem_base, tmem_cols);
    else tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace lun

using namespace lun;

static int launch_flash(const void* qk, const void* v, const void* x, void* y, const float* gamma, int B, int N, int C,
                        const float* lse_in, void* o_out, float* lse_out, void* stream) {
  if (N % 128 || C % 64 || C > 512) return LUN_E_SHAPE;
  const int vw = C > 256 ? 256 : C;
  if (C % vw) return LUN_E_SHAPE;
  static int pair_mode = -1;
  if (pair_mode < 0) {
    const char* e = getenv("LUN_FLASH_PAIR");
    pair_mode = e ? atoi(e) : 1;
  }
  // CTA pairs: 256 queries per cluster, every streamed K / V tile split between the two CTAs
  const int cg = (pair_mode && (N / 128) % 2 == 0 && vw % 128 == 0) ? 2 : 1;
  CUtensorMap tmQK, tmK, tmV;
  int rc = make_tmap_2d(&tmQK, qk, (long)B * N, 128, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmK, qk, (long)B * N, 128, 128 / cg);
  if (rc) return rc;
  rc = make_tmap_nhwc(&tmV, v, B, 1, N, C, 128, 1, 1, 1);     // {C, W=N, H=1, B}: boxes of 64 ch x 128 keys
  if (rc) return rc;
  const int vatoms = vw / 64 / cg;
  const int smem = kFaTileBytes + kFaKStages(cg) * (kFaTileBytes / cg) + kFaPStages(cg) * 2 * kFaTileBytes + 2 * vatoms * kFaTileBytes +
                   (int)sizeof(FaBars) + 1024;
  static bool configured_dev[64] = {};   // the opt-in smem size is a per-device function attribute
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(flash_attn2d_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess ||
        cudaFuncSetAttribute(flash_attn2d_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess)
      return LUN_E_ATTR;
    configured = true;
  }
  dim3 grid(N / 128, B, C / vw);
  if (cg == 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kFaThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, flash_attn2d_kernel<2>, tmQK, tmK, tmV, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, gamma,
                           N, C, vw, lse_in, (__nv_bfloat16*)o_out, lse_out) != cudaSuccess)
      return LUN_E_LAUNCH;
  } else {
    flash_attn2d_kernel<1><<<grid, kFaThreads, smem, (cudaStream_t)stream>>>(
        tmQK, tmK, tmV, (const __nv_bfloat16*)x, (__nv_bfloat16*)y, gamma, N, C, vw, lse_in, (__nv_bfloat16*)o_out,
        lse_out);
  }
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
