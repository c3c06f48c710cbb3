// Implicit-GEMM convolution forward / data-gradient kernel for sm_100a.
//
//   D[128 pixels, block_n channels] = sum over taps t, channel chunks kc of  A_t,kc[128, 64] * B_t,kc[block_n, 64]^T
//
// * A tiles are fetched by TMA in tiled mode from the NHWC bf16 activation tensor viewed as a 4-D tensor
//   {C, W, H, B}; the filter-tap shift is a coordinate offset and the halo is TMA out-of-bounds zero fill, so no
//   im2col matrix ever exists. Strided convolutions use the tensor map's element strides.
// * B tiles (weights packed [tap][Cout][Cin], K-major) are fetched by TMA from a 2-D map.
// * Both land in shared memory in the 128-byte swizzle; one elected thread issues tcgen05.mma (M=128, N=block_n,
//   K=16) with fp32 accumulators in TMEM, double-buffered so the epilogue of tile i overlaps the main loop of tile i+1.
// * Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue
//   (TMEM -> registers -> bias / activation / BN batch statistics -> vectorised global stores).
//
// Replaces the reference's F.conv2d / conv_transpose2d call sites (lunar_evaluator.py:242,249,133,134,255;
// lunar_generate.py:36,41,95-116,169-187) and their autograd data-gradients.
#include "conv_gemm.cuh"
#include "elem_common.cuh"
#include "ptx.cuh"
#include "launch_count.cuh"
#include <stdlib.h>
#include <type_traits>

namespace lun {

constexpr int kThreads = 320;   // TMA warp, MMA warp, 8 epilogue warps
constexpr int kABytes = 128 * 128;  // 128 rows x 64 bf16
constexpr int kAReuseBox = 130 * 128;   // 130 pixel rows (one halo pixel on each side) x 64 bf16
constexpr int kAReuseSlot = 17 * 1024;  // rounded up so the weight tiles behind it stay 1024-byte aligned
constexpr int kBiasSmem = 2048;         // bias entries staged in shared memory

struct __align__(16) PipeBars {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

// CG = 1: one CTA per 128-pixel tile. CG = 2: a CTA pair (cluster of 2) computes a 256-pixel x block_n tile with
// tcgen05.mma.cta_group::2 - each CTA loads its own 128 pixel rows of A and HALF of the weight tile, the leader CTA
// issues the MMAs, both CTAs drain their own 128 accumulator rows.
// STATS: instantiation with the statistics epilogue (its register accumulators cost the plain instantiation nothing).
template <int CG, bool STATS>
__global__ void __launch_bounds__(kThreads, 1)
conv_fprop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO, const ConvGeom g, const float* __restrict__ bias, void* __restrict__ out,
                  float* __restrict__ stats) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int block_n = g.block_n;
  const int b_bytes = block_n / CG * 128;          // this CTA's share of the weight tile
  // a_reuse: a stage holds one (TW+2)-row A box (17 KB slot) and the weight tiles of the 3 dx taps that share it
  const int a_bytes = g.a_reuse ? kAReuseSlot : kABytes;
  const int taps_per_stage = g.a_reuse ? 3 : 1;
  const int stage_bytes = a_bytes + taps_per_stage * b_bytes;
  const int stage_tx = (g.a_reuse ? kAReuseBox : kABytes) + taps_per_stage * b_bytes;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const int stages = g.stages;
  // output staging: per epilogue WARP, stg_bufs (1 or 2) sub-tiles of 2 KB (32 rows x 64 B, 64B swizzle); then, with
  // EPI_STATS, 8 KB of warp-private column-sum slots [8 warps][4 chunks][16 column pairs] x float4
  uint8_t* s_stage = smem + stages * stage_bytes;
  const int nbuf = g.stg_bufs;
  float* s_cstat = reinterpret_cast<float*>(s_stage + 8 * nbuf * 2048);
  PipeBars* bars = reinterpret_cast<PipeBars*>(s_stage + 8 * nbuf * 2048 + ((g.flags & EPI_STATS) ? 8192 : 0));
  float* s_bias = reinterpret_cast<float*>(bars + 1);   // [Cout]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int n_blocks = g.Cout / block_n;
  const int m_tiles = g.ntb * g.nth * g.ntw;
  const int nphase = g.nphase > 1 ? g.nphase : 1;
  const int tiles_pp = m_tiles / CG * n_blocks;     // work items of one phase (tile pairs when CG == 2)
  const int total_tiles = tiles_pp * nphase;
  const int first_item = blockIdx.x / CG, item_stride = gridDim.x / CG;
  const int kchunks = (g.Cin + 63) >> 6;   // a ragged last chunk is zero-filled by TMA (both operands)
  const int ksteps = g.ntaps / taps_per_stage * kchunks;   // pipeline steps per tile
  const uint32_t tmem_cols = (2 * block_n <= 32) ? 32u : (2 * block_n <= 64) ? 64u : (2 * block_n <= 128) ? 128u
                             : (2 * block_n <= 256) ? 256u : 512u;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    if (g.flags & EPI_TMA_STORE) prefetch_tmap(&tmO);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tfull[i], 1);
      mbar_init(&bars->tempty[i], 8 * CG);
    }
    fence_barrier_init();
  }
  // grouped mode: phase p is an independent GEMM group writing channels [p * Cout, (p + 1) * Cout) of the output
  // (bias vectors longer than kBiasSmem entries - the 32768-wide decoder fc - are read from global memory instead)
  const int nbias = g.grouped ? g.Cout * nphase : g.Cout;
  const bool bias_smem = nbias <= kBiasSmem;
  if (bias_smem)
    for (int i = threadIdx.x; i < nbias; i += kThreads) s_bias[i] = (g.flags & EPI_BIAS) ? bias[i] : 0.f;
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
  if (CG == 2) cluster_sync_all();      // peer barriers must be initialised before any remote arrive / TMA signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (lane 0: A tile, lane 1: B tile)
    int s = 0;
    uint32_t ph = 0;
    for (int tile = first_item; tile < total_tiles; tile += item_stride) {
      const int phase = tile / tiles_pp, tl = tile % tiles_pp;
      const int n_blk = tl % n_blocks;
      int m = (tl / n_blocks) * CG + cta_rank;
      const int tw = m % g.ntw;
      m /= g.ntw;
      const int th = m % g.nth;
      const int tb = m / g.nth;
      const int w0 = tw * g.TW * g.in_mul, h0 = th * g.TH * g.in_mul, b0 = tb * g.TB;
      for (int t = phase * g.ntaps; t < (phase + 1) * g.ntaps; t += taps_per_stage) {
        const int cw = w0 + g.dx[t], ch = h0 + g.dy[t];
        const int nsel = n_blk * block_n + cta_rank * (block_n / CG);
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&bars->empty[s], ph ^ 1);
          uint8_t* sa = smem + s * stage_bytes;
          if (lane == 0) {
            if (cta_rank == 0) mbar_expect_tx(&bars->full[s], CG * stage_tx);
            const CUtensorMap* ma = g.a_reuse ? &tmA2 : &tmA;
            if (CG == 2) tma_load_4d_pair(sa, ma, &bars->full[s], kc * 64, cw, ch, b0);
            else tma_load_4d(sa, ma, &bars->full[s], kc * 64, cw, ch, b0);
          } else if (lane <= taps_per_stage) {
            const int j = lane - 1;
            const int wrow = g.slab[t + j] * g.Cout + nsel;
            if (CG == 2) tma_load_2d_pair(sa + a_bytes + j * b_bytes, &tmB, &bars->full[s], kc * 64, wrow);
            else tma_load_2d(sa + a_bytes + j * b_bytes, &tmB, &bars->full[s], kc * 64, wrow);
          }
          if (++s == stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (cta_rank == 0) {
      const uint32_t idesc = make_idesc_bf16(128 * CG, block_n, false, false);
      int s = 0;
      uint32_t ph = 0;
      int acc = 0;
      uint32_t pacc = 0;
      for (int tile = first_item; tile < total_tiles; tile += item_stride) {
        mbar_wait(&bars->tempty[acc], pacc ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * block_n;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&bars->full[s], ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + s * stage_bytes);
            for (int j = 0; j < taps_per_stage; ++j) {
              // tap j of the filter row reads the same A box shifted by j pixel rows (128 bytes each)
              // (the 128B swizzle is a function of absolute smem address bits, so a 128-byte row offset of the
              // start address needs no descriptor base_offset as long as the box itself is 1024-byte aligned)
              const uint64_t adesc = make_smem_desc_sw128(sa + j * 128, 0, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(sa + a_bytes + j * b_bytes, 0, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // advance 16 bf16 (32 bytes) along K inside the 128-byte swizzle row: +2 in the (addr>>4) field
                if (CG == 2) umma_bf16_pair(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (ks | j | k) != 0);
                else umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (ks | j | k) != 0);
              }
            }
            if (CG == 2) {
              umma_commit_pair(&bars->empty[s]);
              if (ks == ksteps - 1) umma_commit_pair(&bars->tfull[acc]);
            } else {
              umma_commit(&bars->empty[s]);
              if (ks == ksteps - 1) umma_commit(&bars->tfull[acc]);
            }
          }
          __syncwarp();
          if (++s == stages) { s = 0; ph ^= 1; }
        }
        if (++acc == 2) { acc = 0; pacc ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    // Two warps share each TMEM lane quarter and split the accumulator columns (half 0: first chunks, half 1: the
    // rest), so a 128x256 tile is drained by 8 warps. Every warp is self-contained: it stages its own 32 rows x 32
    // channels (2 KB, 64-byte swizzle) in private shared memory and sends them off with its own TMA store, so the
    // epilogue has no CTA-level or half-level barriers at all (a chunk used to cost two 128-thread rendezvous, which
    // bounded the 1x1 and thin-K convolutions). TMEM loads are software-pipelined one chunk ahead.
    const int ew = warp - 2;
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = ew >> 2;
    const int row = q * 32 + lane;      // accumulator row == pixel index inside the tile
    const bool do_stats = STATS && (g.flags & EPI_STATS);
    const bool out_f32 = g.flags & EPI_OUT_F32, tma_out = g.flags & EPI_TMA_STORE;
    const bool stats_img = g.flags & EPI_STATS_IMG;
    const bool col_stats = (g.flags & EPI_COL_STATS) && tma_out;   // sums read back from the staged tile
    const bool drop_sum = g.flags & EPI_DROP_SUM;
    const float slope = (g.flags & EPI_LEAKY) ? g.slope : 1.f;
    const uint32_t stg_warp = smem_u32(s_stage) + ew * nbuf * 2048;
    const int nchunks = block_n >> 5;
    const int ch_lo = half == 0 ? 0 : (nchunks + 1) >> 1;
    const int ch_hi = half == 0 ? (nchunks + 1) >> 1 : nchunks;
    const int nch = ch_hi - ch_lo;      // <= 4 (block_n <= 256)
    // this warp's 32 rows inside the tile, as an offset of its TMA store box (box = {32 ch, sub_w, sub_h, sub_b})
    const int sub_w0 = (q * 32) % g.TW, sub_h0 = ((q * 32) / g.TW) % g.TH, sub_b0 = (q * 32) / (g.TW * g.TH);
    // Column statistics (col_stats): after a chunk is staged the warp re-reads its bf16 sub-tile (what was stored):
    // lane (cp = lane & 15, rh = lane >> 4) owns the column PAIR {2 cp, 2 cp + 1} of rows 2 i + rh - sixteen conflict-
    // free 32-bit loads. The two row halves are combined by one shuffle and added to a warp-private shared-memory slot
    // (plain read-modify-write: no ATOMS.CAST.SPIN loops), which lives across tiles - a CTA keeps its channel block
    // while item_stride % n_blocks == 0 - and is flushed with global reductions when the channel block (or, for
    // per-image statistics, the image) changes and at the end.
    const int cp = lane & 15, rh = lane >> 4;
    float4* cs_slot = reinterpret_cast<float4*>(s_cstat) + (ew * 4) * 16 + cp;    // + ci * 16
    if (do_stats && col_stats && lane < 16) {
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) cs_slot[ci * 16] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();
    int acc_blk = -1, acc_img = 0;      // channel block / image the slot sums belong to
    auto flush_stats = [&]() {
      if (acc_blk < 0) return;
      __syncwarp();
      if (lane < 16) {
        float* base = stats + (stats_img ? static_cast<size_t>(acc_img) * 2 * g.Cout : 0) + acc_blk * block_n + 2 * cp;
        for (int ci = 0; ci < nch; ++ci) {
          const float4 t = cs_slot[ci * 16];
          float* d = base + (ch_lo + ci) * 32;
          atomicAdd(d, t.x);
          atomicAdd(d + 1, t.y);
          if (!drop_sum) {
            atomicAdd(d + g.Cout, t.z);
            atomicAdd(d + g.Cout + 1, t.w);
          }
          cs_slot[ci * 16] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      __syncwarp();
    };
    // drop_key(seed, idx8) for idx8 < 2^30: seed_lo ^ (idx8 * 4) ^ seed_hi * 0x85EBCA6B
    const uint32_t drop_k0 = static_cast<uint32_t>(g.drop_seed) ^ static_cast<uint32_t>(g.drop_seed >> 32) * 0x85EBCA6BU;
    int acc = 0;
    uint32_t pacc = 0;
    int stores_issued = 0;              // TMA stores committed by this warp so far (staging ring position)
    for (int tile = first_item; tile < total_tiles; tile += item_stride) {
      const int phase = tile / tiles_pp, tl = tile % tiles_pp;
      const bool spatial_phase = nphase > 1 && !g.grouped;       // transposed conv: phase = output (row, col) parity
      const int o_ph = spatial_phase ? phase >> 1 : g.o_ph, o_pw = spatial_phase ? phase & 1 : g.o_pw;
      const int cbase = g.grouped ? phase * g.Cout : 0;          // grouped GEMM: phase = output channel group
      const int n_blk = tl % n_blocks;
      int m = (tl / n_blocks) * CG + cta_rank;
      const int tw = m % g.ntw;
      m /= g.ntw;
      const int th = m % g.nth;
      const int tb = m / g.nth;
      const int lw = row % g.TW, lh = (row / g.TW) % g.TH, lb = row / (g.TW * g.TH);
      const int gb = tb * g.TB + lb, gh = th * g.TH + lh, gw = tw * g.TW + lw;
      const bool valid = gb < g.GB;
      const int rows_left = (g.GB - tb * g.TB) * g.TW * g.TH;     // rows of this tile that lie inside the batch
      const int nvalid = rows_left < 128 ? rows_left : 128;
      const size_t pix = (static_cast<size_t>(gb) * g.OH + (gh * g.o_mul + o_ph)) * g.OW + (gw * g.o_mul + o_pw);
      const size_t obase = pix * g.ldo + g.o_coff + cbase + n_blk * block_n;
      if (do_stats && col_stats && (n_blk != acc_blk || (stats_img && tb != acc_img))) {
        flush_stats();
        acc_blk = n_blk;
        acc_img = tb;
      }

      mbar_wait(&bars->tfull[acc], pacc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * block_n;
      // process one 32-column chunk held in registers `r` (bias, activation, store, statistics)
      auto process = [&](uint32_t (&r)[32], int ch) {
        const int c0 = ch << 5;
        const int nb = cbase + n_blk * block_n + c0;
        float bv[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = bias_smem ? *reinterpret_cast<const float4*>(s_bias + nb + 4 * j)
                           : (g.flags & EPI_BIAS) ? __ldg(reinterpret_cast<const float4*>(bias + nb) + j)
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
          bv[4 * j] = t.x; bv[4 * j + 1] = t.y; bv[4 * j + 2] = t.z; bv[4 * j + 3] = t.w;
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(r[j]) + bv[j];
          v[j] = fmaxf(x, 0.f) + slope * fminf(x, 0.f);
        }
        if (out_f32) {
          if (valid) {
            float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + obase + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else {
          uint32_t p[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) p[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          if (tma_out) {
            const uint32_t stg_base = stg_warp + (nbuf == 2 ? (stores_issued & 1) * 2048 : 0);
            // the staging tile is reused: wait until the TMA store that last read it has finished with it
            if (stores_issued >= nbuf) {
              if (lane == 0) {
                if (nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              }
              __syncwarp();
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t chunk = static_cast<uint32_t>(j ^ ((lane >> 1) & 3));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_base + lane * 64 + (chunk << 4)),
                           "r"(p[4 * j]), "r"(p[4 * j + 1]), "r"(p[4 * j + 2]), "r"(p[4 * j + 3])
                           : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              asm volatile(
                  "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                      &tmO),
                  "r"(stg_base), "r"(g.o_coff + nb), "r"((tw * g.TW + sub_w0) * g.o_mul + o_pw),
                  "r"((th * g.TH + sub_h0) * g.o_mul + o_ph), "r"(tb * g.TB + sub_b0)
                  : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            ++stores_issued;
            if (do_stats && col_stats) {
              // column-pair sums from the staged bf16 sub-tile (valid until this warp's next wait on the buffer; the
              // TMA store only reads it). Word (row, pair cp) sits in 16-byte chunk (cp >> 2) ^ ((row >> 1) & 3)
              // = (cp >> 2) ^ (i & 3) of the 64-byte row 2 i + rh.
              const uint32_t colw = stg_base + rh * 64 + (cp & 3) * 4;
              const uint32_t cj = static_cast<uint32_t>(cp) >> 2;
              const int left = nvalid - q * 32 - rh;         // valid rows: 2 i < left
              float a10 = 0.f, a11 = 0.f, a20 = 0.f, a21 = 0.f;
              // replay of the elementwise dropout on the stored tensor (EPI_DROP_SUM): element index e = pixel * ldo +
              // channel (32-bit, tile = 128 consecutive pixels - checked by the launcher); e is even for the low column
              // of the pair, so both columns share one hash word (same stream as drop_keep1())
              const uint32_t e_row0 =
                  static_cast<uint32_t>((static_cast<size_t>(tb) * g.OH + th) * g.OW + tw * 128 + q * 32 + rh) *
                      static_cast<uint32_t>(g.ldo) + static_cast<uint32_t>(g.o_coff + nb + 2 * cp);
#pragma unroll 8
              for (int i = 0; i < 16; ++i) {
                uint32_t w;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(colw + i * 128 + ((cj ^ (i & 3)) << 4)) : "memory");
                float f0 = 2 * i < left ? __uint_as_float(w << 16) : 0.f;
                float f1 = 2 * i < left ? __uint_as_float(w & 0xffff0000u) : 0.f;
                if (drop_sum) {
                  const uint32_t e = e_row0 + static_cast<uint32_t>(2 * i) * static_cast<uint32_t>(g.ldo);
                  const uint32_t h = hash32((drop_k0 ^ ((e >> 3) << 2)) + ((e >> 1) & 3u));
                  f0 = (h & 0xFFFFu) >= g.drop_thresh16 ? rbf(f0 * g.drop_scale) : 0.f;
                  f1 = (h >> 16) >= g.drop_thresh16 ? rbf(f1 * g.drop_scale) : 0.f;
                }
                a10 += f0;
                a11 += f1;
                a20 = fmaf(f0, f0, a20);
                a21 = fmaf(f1, f1, a21);
              }
              a10 += __shfl_xor_sync(0xffffffffu, a10, 16);
              a11 += __shfl_xor_sync(0xffffffffu, a11, 16);
              a20 += __shfl_xor_sync(0xffffffffu, a20, 16);
              a21 += __shfl_xor_sync(0xffffffffu, a21, 16);
              if (rh == 0) {
                float4* slot = cs_slot + (ch - ch_lo) * 16;
                float4 t = *slot;
                t.x += a10; t.y += a11; t.z += a20; t.w += a21;
                *slot = t;
              }
            }
          } else if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + obase + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_uint4(p[4 * j], p[4 * j + 1], p[4 * j + 2], p[4 * j + 3]);
          }
          if (do_stats && !col_stats) {
            // statistics of what was stored (bf16-rounded), as the reference's batch_norm sees them
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              v[2 * j] = __uint_as_float(p[j] << 16);
              v[2 * j + 1] = __uint_as_float(p[j] & 0xffff0000u);
            }
          }
        }
        if (do_stats && !col_stats) {
          float s1[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            s1[j] = valid ? v[j] : 0.f;
            s2[j] = s1[j] * s1[j];
          }
          // transpose-reduce across the 32 lanes (rows): 31 shuffles per statistic, lane j ends with column j
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
          // one coalesced reduction per warp straight to global (per image for GroupNorm)
          float* st = stats + (stats_img ? static_cast<size_t>(tb) * 2 * g.Cout : 0) + nb + lane;
          atomicAdd(st, s1[0]);
          atomicAdd(st + g.Cout, s2[0]);
        }
      };
      uint32_t ra[32], rb[32];
      if (ch_lo < ch_hi) tmem_ld32(taddr + ch_lo * 32, ra);
      for (int ch = ch_lo; ch < ch_hi; ch += 2) {
        tmem_ld_wait();                                               // ra has landed
        if (ch + 1 < ch_hi) tmem_ld32(taddr + (ch + 1) * 32, rb);     // prefetch while ra is processed
        process(ra, ch);
        if (ch + 1 < ch_hi) {
          tmem_ld_wait();                                             // rb has landed
          if (ch + 2 < ch_hi) tmem_ld32(taddr + (ch + 2) * 32, ra);
          process(rb, ch + 1);
        }
      }
      // every TMEM read of this accumulator stage by this warp has completed (tmem_ld_wait above)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(&bars->tempty[acc], 0);   // the leader's MMA warp owns the accumulator hand-back
        else mbar_arrive(&bars->tempty[acc]);
      }
      if (++acc == 2) { acc = 0; pacc ^= 1; }
    }
    if (do_stats && col_stats) flush_stats();
    if (stores_issued && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 4-D bf16 NHWC activation map {C, W, H, B}; box = {64, bw, bh, bb} elements traversed with strides {1, es, es, 1}.
static int make_tmap_nhwc_ex(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_c, int box_w,
                            int box_h, int box_b, int estride, CUtensorMapSwizzle swz) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return 101;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_b};
  cuuint32_t es[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 102;
}
int make_tmap_nhwc(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h, int box_b,
                   int estride) {
  return make_tmap_nhwc_ex(m, base, B, H, W, C, 64, box_w, box_h, box_b, estride, CU_TENSOR_MAP_SWIZZLE_128B);
}

// 2-D bf16 row-major matrix map {cols, rows}; box = {64, box_rows}.
int make_tmap_2d(CUtensorMap* m, const void* base, long rows, long cols, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return 101;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 103;
}

int num_sms() {
  static int n[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!n[dev]) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}

static bool pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

// Split an output grid into 128-pixel tiles; returns false if the grid cannot be tiled.
bool tile_grid(int GB, int GH, int GW, int pixels, int* TB, int* TH, int* TW) {
  if (!pow2(GW) || !pow2(GH)) return false;
  *TW = GW < pixels ? GW : pixels;
  int rest = pixels / *TW;
  *TH = GH < rest ? GH : rest;
  *TB = rest / *TH;
  return true;
}

// x: NHWC bf16 [XB, XH, XW, Cin]; wpk: [nslabs][Cout][Cin] bf16; out: [*, OH, OW, ldo]
int launch_conv_fprop(const void* x, int XB, int XH, int XW, const void* wpk, int nslabs, ConvGeom g,
                      const float* bias, void* out, float* stats, cudaStream_t stream) {
  if (g.Cin % 8 != 0 || g.block_n % 32 != 0 || g.block_n > 256 || g.Cout % g.block_n != 0) return 2;
  if (g.nphase < 1) g.nphase = 1;
  if (g.ntaps < 1 || g.ntaps * g.nphase > kMaxTaps) return 3;
  if (!tile_grid(g.GB, g.GH, g.GW, 128, &g.TB, &g.TH, &g.TW)) return 4;
  g.ntw = g.GW / g.TW;
  g.nth = g.GH / g.TH;
  g.ntb = (g.GB + g.TB - 1) / g.TB;
  if (g.TW * g.in_mul > 256 || g.TH * g.in_mul > 256 || g.TB > 256) return 5;
  if (g.Cout > 2048 && (g.flags & EPI_STATS)) return 6;
  if ((g.flags & EPI_STATS_IMG) && (!(g.flags & EPI_STATS) || g.TB != 1)) return 6;
  if (g.grouped && ((g.flags & EPI_STATS) || g.nphase < 2)) return 6;
  if ((g.flags & EPI_DROP_SUM) &&
      (!(g.flags & EPI_STATS) || (g.flags & (EPI_OUT_F32 | EPI_STATS_IMG)) || g.o_mul != 1 || g.TW != 128 || g.TH != 1 ||
       g.TB != 1 || (double)g.GB * g.OH * g.OW * g.ldo >= 4294967296.0))
    return 6;                                    // the fused replay indexes 128-pixel row tiles with 32-bit elements
  if (!(g.flags & EPI_OUT_F32) && (g.ldo % 8 || g.o_coff % 8)) return 7;

  CUtensorMap tmA, tmA2, tmB, tmO;
  const int m_tiles = g.ntb * g.nth * g.ntw;
  // CTA pairs for the big tiles: full 256-wide N block, an even number of pixel tiles, enough work for every pair
  const int sms = num_sms();
  static int pair_mode = -1;
  if (pair_mode < 0) {
    const char* e = getenv("LUN_CONV_PAIR");
    pair_mode = e ? atoi(e) : 1;
  }
  // (pairs for 128-wide N blocks were measured on the 256->128 transposed conv: 93.8 vs 93.0 us, not kept)
  const int cg = (pair_mode && g.block_n == 256 && m_tiles % 2 == 0 && sms % 2 == 0 &&
                  (long)m_tiles / 2 * (g.Cout / g.block_n) * g.nphase >= sms / 2) ? 2 : 1;
  // A-tile reuse: taps sorted by (dy, dx) come in runs of 3 with equal dy and consecutive dx, one image row per tile
  {
    for (int i = 1; i < g.ntaps && g.nphase == 1; ++i)   // insertion sort of the tap list by (dy, dx)
      for (int j = i; j > 0 && (g.dy[j] < g.dy[j - 1] || (g.dy[j] == g.dy[j - 1] && g.dx[j] < g.dx[j - 1])); --j) {
        int t;
        t = g.dy[j]; g.dy[j] = g.dy[j - 1]; g.dy[j - 1] = t;
        t = g.dx[j]; g.dx[j] = g.dx[j - 1]; g.dx[j - 1] = t;
        t = g.slab[j]; g.slab[j] = g.slab[j - 1]; g.slab[j - 1] = t;
      }
    static int reuse_mode = -1;
    if (reuse_mode < 0) {
      const char* e = getenv("LUN_CONV_AREUSE");
      reuse_mode = e ? atoi(e) : 1;
    }
    bool ok = reuse_mode && g.nphase == 1 && cg == 2 && g.ntaps % 3 == 0 && g.TW == 128 && g.TH == 1 && g.TB == 1 && g.in_mul == 1 &&
              g.block_n == 256;
    for (int i = 0; ok && i < g.ntaps; i += 3)
      ok = g.dy[i + 1] == g.dy[i] && g.dy[i + 2] == g.dy[i] && g.dx[i + 1] == g.dx[i] + 1 && g.dx[i + 2] == g.dx[i] + 2;
    g.a_reuse = ok ? 1 : 0;
  }
  // coalesced asynchronous output path: bf16, dense pixel mapping, 32-channel boxes
  {
    static int col_mode = -1;
    if (col_mode < 0) {
      const char* e = getenv("LUN_CONV_COLSTATS");
      col_mode = e ? atoi(e) : 1;
    }
    if (col_mode || (g.flags & EPI_DROP_SUM)) g.flags |= EPI_COL_STATS;    // the masked sum only exists on this path
  }
  // (a transposed-conv phase writes every other pixel of every other row: the same box with element strides 2)
  if (!(g.flags & EPI_OUT_F32) && (g.o_mul == 1 || g.o_mul == 2))
    g.flags |= EPI_TMA_STORE;
  else
    g.flags &= ~EPI_TMA_STORE;
  int rc = make_tmap_nhwc(&tmA, x, XB, XH, XW, g.Cin, g.TW * g.in_mul, g.TH * g.in_mul, g.TB, g.in_mul);
  if (rc) return rc;
  if (g.a_reuse) {
    rc = make_tmap_nhwc(&tmA2, x, XB, XH, XW, g.Cin, g.TW + 2, 1, 1, 1);
    if (rc) return rc;
  } else {
    tmA2 = tmA;
  }
  rc = make_tmap_2d(&tmB, wpk, (long)nslabs * g.Cout, g.Cin, g.block_n / cg);
  if (rc) return rc;
  if (g.flags & EPI_TMA_STORE) {
    // one store per epilogue warp: its 32 accumulator rows = a {32 ch, sub_w, sub_h, sub_b} box of the output tensor
    const int sub_w = g.TW < 32 ? g.TW : 32;
    const int sub_h = g.TH < 32 / sub_w ? g.TH : 32 / sub_w;
    const int sub_b = 32 / (sub_w * sub_h);
    if (sub_b > g.TB) return 5;
    rc = make_tmap_nhwc_ex(&tmO, out, g.GB, g.OH, g.OW, g.ldo, 32, sub_w * g.o_mul, sub_h * g.o_mul, sub_b, g.o_mul,
                           CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  } else {
    tmO = tmA;
  }
  const int stage_bytes = g.a_reuse ? kAReuseSlot + 3 * (g.block_n / cg * 128) : kABytes + g.block_n / cg * 128;
  // two staging tiles per epilogue half (the next chunk is written while the previous TMA store drains) whenever the
  // operand ring keeps at least 4 stages; the 65 KB stages of the A-reuse configuration leave room for one
  static int stg_mode = -1;
  if (stg_mode < 0) {
    const char* e = getenv("LUN_CONV_STG2");
    stg_mode = e ? atoi(e) : 1;
  }
  int stages = 0, extra = 0;
  for (g.stg_bufs = (stg_mode && (g.flags & EPI_TMA_STORE)) ? 2 : 1; g.stg_bufs >= 1; --g.stg_bufs) {
    extra = 8 * g.stg_bufs * 2048 + ((g.flags & EPI_STATS) ? 8192 : 0) + (int)sizeof(PipeBars) +
            (g.Cout * (g.grouped ? g.nphase : 1) <= kBiasSmem ? g.Cout * (g.grouped ? g.nphase : 1) * 4 : 0) + 1024;
    stages = (227 * 1024 - extra) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages >= 4 || g.stg_bufs == 1 || stg_mode == 2) break;
  }
  if (stages < 1) return 5;
  g.stages = stages;
  const int smem_bytes = stages * stage_bytes + extra;
  // the opt-in shared-memory size is a per-device function attribute: set it once for every device this process uses
  {
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 8;
    if (!configured[dev]) {
      const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
      if (cudaFuncSetAttribute(conv_fprop_kernel<1, false>, attr, 227 * 1024) != cudaSuccess ||
          cudaFuncSetAttribute(conv_fprop_kernel<1, true>, attr, 227 * 1024) != cudaSuccess ||
          cudaFuncSetAttribute(conv_fprop_kernel<2, false>, attr, 227 * 1024) != cudaSuccess ||
          cudaFuncSetAttribute(conv_fprop_kernel<2, true>, attr, 227 * 1024) != cudaSuccess)
        return 8;
      configured[dev] = true;
    }
  }
  const bool with_stats = g.flags & EPI_STATS;
  const int total_items = m_tiles / cg * (g.Cout / g.block_n) * g.nphase;
  int grid = sms;
  if (grid > total_items * cg) grid = total_items * cg;
  if (cg == 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const cudaError_t e = with_stats
        ? cudaLaunchKernelEx(&cfg, conv_fprop_kernel<2, true>, tmA, tmA2, tmB, tmO, g, bias, out, stats)
        : cudaLaunchKernelEx(&cfg, conv_fprop_kernel<2, false>, tmA, tmA2, tmB, tmO, g, bias, out, stats);
    if (e != cudaSuccess) return 9;
  } else if (with_stats) {
    conv_fprop_kernel<1, true><<<grid, kThreads, smem_bytes, stream>>>(tmA, tmA2, tmB, tmO, g, bias, out, stats);
  } else {
    conv_fprop_kernel<1, false><<<grid, kThreads, smem_bytes, stream>>>(tmA, tmA2, tmB, tmO, g, bias, out, stats);
  }
  note_launch(1);
  return cudaGetLastError() == cudaSuccess ? 0 : 9;
}

}  // namespace lun
