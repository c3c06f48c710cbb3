// ConvTranspose2d(k=4, s=2, p=1) for the thin decoder stages (lunar_generate.py:181-187: 128->64 and 64->32 channels)
// as ONE kernel over all output phases, fed from a halo tile.
//
//   out[b, 2j+ph, 2i+pw, co] = bias[co] + sum_{(dh,kh) in R(ph)} sum_{(dw,kw) in R(pw)} sum_ci
//                              x[b, j+dh, i+dw, ci] * W[ci][co][kh][kw],     R(0) = {(0,1), (-1,3)},  R(1) = {(1,0), (0,2)}
//
// The generic tap-list kernel runs each output phase as its own launch and fetches one shifted 128-pixel A tile per
// (phase, tap): 16 tiles of 16 KB per 128 input pixels, which makes these N = 32 / 64 stages L2->SM bound. Here
//   * a CTA owns a 16 x 8 block of INPUT pixels and loads its 18 x 10 halo ONCE per 64-channel chunk (one TMA box,
//     23 KB, zero fill outside the image); every (dh, dw) shift is the same shared-memory tile read through a UMMA
//     descriptor whose start address is moved by (dh+1)*10 + (dw+1) pixel rows and whose 8-row group stride (SBO) is
//     the HALO row pitch (10 * 128 B) instead of 1024 B - the 128-byte swizzle is a function of the absolute shared
//     memory address, so TMA's writes and the MMA's reads agree for any 128-byte-aligned start and any SBO;
//   * the weight slabs of the phases a CTA computes stay RESIDENT in shared memory for the CTA's whole life;
//   * the accumulator holds NPH phases side by side ([128 pixels] x [NPH * Cout] fp32 in TMEM, double buffered);
//     the epilogue writes out[b, 2j+ph, 2i+pw, :] (adjacent pw are contiguous in memory) and, optionally, the
//     per-image GroupNorm sums.
// Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue (two per TMEM lane quarter).
#include "../../include/lunaris_b200.h"
#include "conv_gemm.cuh"
#include "launch_count.cuh"
#include "ptx.cuh"

namespace lun {

int make_tmap_2d(CUtensorMap* m, const void* base, long rows, long cols, int box_rows);
int make_tmap_nhwc(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h, int box_b,
                   int estride);
int num_sms();

constexpr int kCtThreads = 320;
constexpr int kCtTW = 8, kCtTH = 16;                       // input pixels per tile (GEMM M = 128: row = j*8 + i)
constexpr int kCtHW = kCtTW + 2, kCtHH = kCtTH + 2;        // halo tile
constexpr int kCtHaloBytes = kCtHW * kCtHH * 128;          // 23040
constexpr int kCtHaloSlot = 23 * 1024;                     // 1024-byte aligned slot
constexpr int kCtMaxStages = 6;
constexpr int kCtStageBytes = 4096;                       // per epilogue warp: 32 pixels x up to 128 contiguous output bytes

struct CtGeom {
  int B, H, W, Cin, Cout;
  int nph;                        // phases computed by one CTA (4: all; 1: blockIdx.y selects the phase)
  int stages;
  int flags;                      // EPI_BIAS, EPI_STATS | EPI_STATS_IMG
};

struct __align__(16) CtBars {
  uint64_t w_full;
  uint64_t full[kCtMaxStages], empty[kCtMaxStages];
  uint64_t tfull[2], tempty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// Tap list of a CTA in closed form. Entry e = 4 * (local phase) + 2 * a + b with a, b in {0, 1} choosing one of the
// two rows / columns of the 4 x 4 filter that reach output phase (ph, pw):
//   dh = ph - a, kh = (1 - ph) + 2a;   dw = pw - b, kw = (1 - pw) + 2b      (R(0) = {(0,1), (-1,3)}, R(1) = {(1,0), (0,2)})
__device__ __forceinline__ void ct_entry(int p, int e, int* dh, int* dw, int* slab) {
  const int ph = p >> 1, pw = p & 1, a = (e >> 1) & 1, b = e & 1;
  *dh = ph - a;
  *dw = pw - b;
  *slab = ((1 - ph) + 2 * a) * 4 + (1 - pw) + 2 * b;
}

template <int NPH, int COUT>
__global__ void __launch_bounds__(kCtThreads, 1)
convt_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const CtGeom g,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, float* __restrict__ stats) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int Cout = COUT;
  const int kchunks = g.Cin >> 6;
  constexpr int w_tile = Cout * 128;                       // one slab, one 64-channel chunk: [Cout rows][64 ch]
  const int phase_sel = NPH == 1 ? blockIdx.y : 0;
  constexpr int nent = NPH * 4;
  // tap list of this CTA, fully unrolled so that it lives in registers / immediates: entry e = 4 * (local phase) + tap
  int e_slab[nent];
  uint32_t e_aoff[nent];                               // descriptor start-address offset ((bytes) >> 4) of the shift
#pragma unroll
  for (int e = 0; e < nent; ++e) {
    int dh, dw;
    ct_entry(NPH == 1 ? phase_sel : (e >> 2), e, &dh, &dw, &e_slab[e]);
    e_aoff[e] = static_cast<uint32_t>(((dh + 1) * kCtHW + dw + 1) * 128) >> 4;
  }
  uint8_t* sW = smem;                                  // [kchunks][nent][Cout][64] resident weights
  uint8_t* sA = sW + kchunks * nent * w_tile;          // ring of halo tiles
  uint8_t* sOut = sA + g.stages * kCtHaloSlot;         // 8 x 4 KB: per-warp output staging (epilogue)
  CtBars* bars = reinterpret_cast<CtBars*>(sOut + 8 * kCtStageBytes);
  float* s_bias = reinterpret_cast<float*>(bars + 1);  // [Cout]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntw = g.W / kCtTW, nth = g.H / kCtTH;
  const int total_tiles = g.B * nth * ntw;
  // a CTA walks a CONTIGUOUS range of tiles: consecutive tiles belong to the same image, so the GroupNorm sums are
  // carried in registers and reduced across lanes only when the image changes
  const int per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
  const int tile0 = blockIdx.x * per_cta;
  const int tile1 = tile0 + per_cta < total_tiles ? tile0 + per_cta : total_tiles;
  constexpr int ncols = NPH * Cout;                      // accumulator columns per stage
  const uint32_t tmem_cols = 2 * ncols <= 64 ? 64u : 2 * ncols <= 128 ? 128u : 2 * ncols <= 256 ? 256u : 512u;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmX);
    prefetch_tmap(&tmW);
    mbar_init(&bars->w_full, 1);
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tfull[i], 1);
      mbar_init(&bars->tempty[i], 8);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < Cout; i += kCtThreads) s_bias[i] = (g.flags & EPI_BIAS) ? bias[i] : 0.f;
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(&bars->w_full, kchunks * nent * w_tile);
      for (int kc = 0; kc < kchunks; ++kc)
#pragma unroll
        for (int e = 0; e < nent; ++e)
          tma_load_2d(sW + (kc * nent + e) * w_tile, &tmW, &bars->w_full, kc * 64, e_slab[e] * Cout);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = tile0; tile < tile1; ++tile) {
        const int tw = tile % ntw, th = (tile / ntw) % nth, b = tile / (ntw * nth);
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&bars->empty[s], ph ^ 1);
          mbar_expect_tx(&bars->full[s], kCtHaloBytes);
          tma_load_4d(sA + s * kCtHaloSlot, &tmX, &bars->full[s], kc * 64, tw * kCtTW - 1, th * kCtTH - 1, b);
          if (++s == g.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const uint32_t idesc = make_idesc_bf16(128, Cout, false, false);
    mbar_wait(&bars->w_full, 0);
    int s = 0, acc = 0;
    uint32_t ph = 0, pacc = 0;
    for (int tile = tile0; tile < tile1; ++tile) {
      mbar_wait(&bars->tempty[acc], pacc ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * ncols;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&bars->full[s], ph);
        tc_fence_after();
        if (elect_one()) {                               // elect.sync: the compiler keeps the operands in uniform registers
          const uint32_t sa = smem_u32(sA + s * kCtHaloSlot);
          // the issue loop is the critical path of these thin MMAs (N = 32 / 64: a few tens of tensor cycles each):
          // descriptors are one add away from two per-stage bases, the tap list is immediates
          const uint64_t a0 = make_smem_desc_sw128(smem_u32(sA + s * kCtHaloSlot), 0, kCtHW * 128);
          const uint64_t b0 = make_smem_desc_sw128(smem_u32(sW + kc * nent * w_tile), 0, 1024);
          constexpr uint32_t wstep = static_cast<uint32_t>(w_tile) >> 4;
          // issue order: tap, k-step, phase - consecutive MMAs accumulate into DIFFERENT phase columns
#pragma unroll
          for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int lp = 0; lp < NPH; ++lp) {
                const int e = lp * 4 + t;
                umma_bf16(tmem_d + lp * Cout, a0 + e_aoff[e] + 2 * k, b0 + e * wstep + 2 * k, idesc, (kc | t | k) != 0);
              }
          umma_commit(&bars->empty[s]);
          if (kc == kchunks - 1) umma_commit(&bars->tfull[acc]);
        }
        __syncwarp();
        if (++s == g.stages) { s = 0; ph ^= 1; }
      }
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int ew = warp - 2;
    const int q = warp & 3;               // TMEM lane quarter
    const int half = ew >> 2;             // which half of the accumulator chunks this warp drains
    const int row = q * 32 + lane;        // == tile pixel j*8 + i
    const int lj = row >> 3, li = row & 7;
    // chunks (32 accumulator columns) of this warp: all of them cover the SAME 32 output channels [c0, c0 + 32)
    constexpr int cpp = Cout / 32;                                    // chunks per phase
    constexpr int my_n = cpp == 2 ? NPH : (NPH == 4 ? 2 : 1);         // chunks per warp (NPH == 1, cpp == 1: half 0 only)
    const bool idle = (cpp == 1 && NPH == 1 && half == 1);
    const int c0 = cpp == 2 ? 32 * half : 0;
    const bool do_stats = g.flags & EPI_STATS;
    const int OW = 2 * g.W, OH = 2 * g.H;
    constexpr int kPB = (NPH == 4 && cpp == 1) ? 128 : 64;          // contiguous output bytes per thread per store round
    uint8_t* stg = sOut + ew * kCtStageBytes;
    typedef __nv_bfloat16 bf16_t;
    int acc = 0;
    uint32_t pacc = 0;
    float s1[32], s2[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) s1[j] = s2[j] = 0.f;
    int cur_b = -1;
    auto flush_stats = [&](int b) {
      // transpose-reduce across the 32 lanes (rows): lane j ends with column j
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = lane & off;
#pragma unroll
        for (int i = 0; i < off; ++i) {
          const float send1 = upper ? s1[i] : s1[i + off];
          const float keep1 = upper ? s1[i + off] : s1[i];
          s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, off);
          const float send2 = upper ? s2[i] : s2[i + off];
          const float keep2 = upper ? s2[i + off] : s2[i];
          s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
        }
      }
      float* st = stats + static_cast<size_t>(b) * 2 * Cout + c0 + lane;
      atomicAdd(st, s1[0]);
      atomicAdd(st + Cout, s2[0]);
#pragma unroll
      for (int j = 0; j < 32; ++j) s1[j] = s2[j] = 0.f;
    };
    for (int tile = tile0; tile < tile1; ++tile) {
      const int tw = tile % ntw, th = (tile / ntw) % nth, b = tile / (ntw * nth);
      const int gj = th * kCtTH + lj, gi = tw * kCtTW + li;
      if (do_stats && !idle && b != cur_b) {
        if (cur_b >= 0) flush_stats(cur_b);
        cur_b = b;
      }
      mbar_wait(&bars->tfull[acc], pacc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * ncols;
      if (!idle) {
#pragma unroll
        for (int n = 0; n < my_n; ++n) {
          const int ch = cpp == 2 ? n * 2 + half : half * my_n + n;   // accumulator chunk
          const int pl = ch / cpp;                                    // local phase
          const int p = NPH == 1 ? phase_sel : pl;
          uint32_t r[32];
          tmem_ld32(taddr + ch * 32, r);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * j);
            pk[2 * j] = pack_bf16x2(__uint_as_float(r[4 * j]) + bv.x, __uint_as_float(r[4 * j + 1]) + bv.y);
            pk[2 * j + 1] = pack_bf16x2(__uint_as_float(r[4 * j + 2]) + bv.z, __uint_as_float(r[4 * j + 3]) + bv.w);
          }
          // A thread holds 64 contiguous output bytes per chunk (128 when its two chunks are the two horizontal phases of
          // one output row: NPH = 4, Cout = 32), but neighbouring threads are 128 / 256 bytes apart, so direct 16-byte
          // stores touched 32 different lines per instruction (ncu: L1/TEX 70 % busy in a kernel that moves 0.4 GB). The
          // pieces go through a warp-private staging tile (XOR-swizzled: conflict-free both ways) and leave as stores in
          // which consecutive lanes write consecutive 16-byte pieces.
          if (kPB == 128) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 * n + j) ^ (lane & 7)) << 4)) =
                  make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) =
                  make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
          if (kPB == 64 || n == my_n - 1) {
            __syncwarp();
            constexpr int G = kPB / 16;                                   // lanes per thread-slot
#pragma unroll
            for (int k = 0; k < G; ++k) {
              const int t = k * (32 / G) + lane / G, piece = lane % G;    // slot t = the lane whose pixel this piece belongs to
              const int tj = th * kCtTH + q * 4 + (t >> 3), ti = tw * kCtTW + (t & 7);
              const uint4 v = *reinterpret_cast<const uint4*>(
                  stg + t * kPB + ((kPB == 128 ? (piece ^ (t & 7)) : (piece ^ ((t >> 1) & 3))) << 4));
              // kPB == 128: pieces 0-3 are phase pw = 0, pieces 4-7 phase pw = 1 of the same output row (oh from p >> 1)
              const int ow_t = 2 * ti + (kPB == 128 ? 0 : (p & 1));
              const int oh_t = 2 * tj + (p >> 1);
              bf16_t* dstp = out + ((static_cast<size_t>(b) * OH + oh_t) * OW + ow_t) * Cout + c0 + piece * 8;
              *reinterpret_cast<uint4*>(dstp) = v;
            }
            __syncwarp();
          }
          if (do_stats) {
            // statistics of what was stored (bf16-rounded), as GroupNorm sees them
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float v0 = __uint_as_float(pk[j] << 16), v1 = __uint_as_float(pk[j] & 0xffff0000u);
              s1[2 * j] += v0;
              s2[2 * j] += v0 * v0;
              s1[2 * j + 1] += v1;
              s2[2 * j + 1] += v1 * v1;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tempty[acc]);
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
    if (do_stats && !idle && cur_b >= 0) flush_stats(cur_b);
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

extern "C" {

int lun_convT4x4s2_halo_bf16(const void* x, int B, int H, int W, int Cin, const void* w_packed, int Cout,
                             const float* bias, void* out, float* img_stats, void* stream) {
  if (H % kCtTH || W % kCtTW || Cin % 64 || (Cout != 32 && Cout != 64)) return LUN_E_SHAPE;
  const int kchunks = Cin / 64;
  // all four phases per CTA when their 16 slabs fit next to a 4-stage halo ring, otherwise one phase per CTA
  const int budget = 227 * 1024 - 1024 - (int)sizeof(CtBars) - Cout * 4 - 8 * kCtStageBytes;
  int nph = 4;
  if (kchunks * 16 * Cout * 128 + 4 * kCtHaloSlot > budget) nph = 1;
  const int nent = nph == 4 ? 16 : 4;
  const int w_bytes = kchunks * nent * Cout * 128;
  if (w_bytes + 2 * kCtHaloSlot > budget) return LUN_E_SHAPE;
  int stages = (budget - w_bytes) / kCtHaloSlot;
  if (stages > kCtMaxStages) stages = kCtMaxStages;
  CUtensorMap tmX, tmW;
  int rc = make_tmap_nhwc(&tmX, x, B, H, W, Cin, kCtHW, kCtHH, 1, 1);
  if (rc) return rc;
  rc = make_tmap_2d(&tmW, w_packed, 16L * Cout, Cin, Cout);
  if (rc) return rc;
  CtGeom g{};
  g.B = B; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout;
  g.nph = nph;
  g.stages = stages;
  g.flags = (bias ? EPI_BIAS : 0) | (img_stats ? (EPI_STATS | EPI_STATS_IMG) : 0);
  const int smem = w_bytes + stages * kCtHaloSlot + 8 * kCtStageBytes + (int)sizeof(CtBars) + Cout * 4 + 1024;
  static bool configured_dev[64] = {};   // the opt-in smem size is a per-device function attribute
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(convt_halo_kernel<1, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess ||
        cudaFuncSetAttribute(convt_halo_kernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess ||
        cudaFuncSetAttribute(convt_halo_kernel<4, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess ||
        cudaFuncSetAttribute(convt_halo_kernel<4, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess)
      return LUN_E_ATTR;
    configured = true;
  }
  const int tiles = B * (H / kCtTH) * (W / kCtTW);
  int gx = num_sms() / (nph == 1 ? 4 : 1);
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  dim3 grid(gx, nph == 1 ? 4 : 1);
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* o = (__nv_bfloat16*)out;
  if (nph == 4 && Cout == 32) convt_halo_kernel<4, 32><<<grid, kCtThreads, smem, st>>>(tmX, tmW, g, bias, o, img_stats);
  else if (nph == 4) convt_halo_kernel<4, 64><<<grid, kCtThreads, smem, st>>>(tmX, tmW, g, bias, o, img_stats);
  else if (Cout == 32) convt_halo_kernel<1, 32><<<grid, kCtThreads, smem, st>>>(tmX, tmW, g, bias, o, img_stats);
  else convt_halo_kernel<1, 64><<<grid, kCtThreads, smem, st>>>(tmX, tmW, g, bias, o, img_stats);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

}  // extern "C"
