// Weight-gradient kernel for sm_100a:  dW[slab][co][ci] += sum_pixels dY[pixel, co] * X[shift_tap(pixel), ci]
//
// The reduction (GEMM-K) dimension is the pixel index, so with NHWC activations both operands are "MN-major":
// a TMA box of {64 channels, 64 pixels} lands in shared memory as 64 rows of 128 bytes (128-byte swizzle), which is
// exactly the canonical MN-major SWIZZLE_128B UMMA layout (LBO = 8 KB between 64-channel atoms, SBO = 1 KB between
// 8-pixel groups). tcgen05.mma M=128 (Cout block), N=block_n (Cin block), K=16 pixels, fp32 accumulators in TMEM.
// Work unit = (Cout block, Cin block, tap, K-split); each finishes with vectorised fp32 red.global.add into the packed
// gradient buffer (zeroed by the caller). Two schedules:
//   lockstep (tiles <= workers): a worker owns ONE output tile and a contiguous 1/wpt of the pixel range, walked as
//     nsub consecutive sub-units (the drain of one overlaps the main loop of the next). All workers of a K slice move
//     through the same pixels at the same time, so every dY / X byte comes from DRAM once and from L2 for the other
//     tiles (round-robin units straddled slices and re-fetched each of them about twice: 4.8 GB read for 2.15 GB);
//   round-robin (more tiles than workers): units dealt to a persistent grid.
//
// Replaces the autograd weight-gradients of the reference's conv2d / conv_transpose2d call sites
// (lunar_evaluator.py:249,134,255; lunar_generate.py:36,41,95-116,169-187).
#include "conv_gemm.cuh"
#include "ptx.cuh"
#include "launch_count.cuh"
#include <stdlib.h>

namespace lun {

int make_tmap_nhwc(CUtensorMap* m, const void* base, int B, int H, int W, int C, int box_w, int box_h, int box_b,
                   int estride);
int num_sms();
bool tile_grid(int GB, int GH, int GW, int pixels, int* TB, int* TH, int* TW);

constexpr int kWgThreads = 192;
constexpr int kWgXBox = 66 * 128;     // reuse3: 66 pixel rows (one halo pixel each side) x 64 channels
constexpr int kWgXSlot = 9 * 1024;    // ... in a 1024-byte aligned slot

struct __align__(8) WgBars {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// CG = 2: a CTA pair accumulates a 256 (Cout) x block_n (Cin) tile with tcgen05.mma.cta_group::2 - each CTA loads its own
// 128 Cout columns of dY and HALF of the Cin columns of X, the leader issues the MMAs, both drain their 128 rows.
template <int CG>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ CUtensorMap tmX2, const WgradGeom g, float* __restrict__ dw) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int block_n = g.block_n;
  const bool r3 = g.reuse3 != 0;
  const int atom = g.kpx * 128;                           // bytes of one 64-channel atom: kpx pixel rows x 128 B
  const int a_bytes = 2 * atom;                           // dY operand: 128 Cout columns
  const int xslot = r3 ? kWgXSlot : atom;                 // bytes between the 64-channel atoms of the X operand
  const int nb_cta = block_n / CG;                        // Cin columns of the X operand loaded by this CTA
  const int b_bytes = nb_cta / 64 * xslot;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const int stage_bytes = a_bytes + b_bytes;
  const int stage_tx = a_bytes + nb_cta / 64 * (r3 ? kWgXBox : atom);
  const int tps = r3 ? 3 : 1;                             // taps accumulated per unit
  const int nacc = r3 ? 1 : 2;                            // accumulator sets in TMEM
  const int stages = g.stages;
  WgBars* bars = reinterpret_cast<WgBars*>(smem + stages * stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_blocks = (g.Cout + 128 * CG - 1) / (128 * CG);
  const int n_blocks = (g.Cin + block_n - 1) / block_n;
  const int worker = blockIdx.x / CG, nworkers = gridDim.x / CG;
  const int tiles = m_blocks * n_blocks * (g.ntaps / tps);
  const int units = tiles * g.splits;
  // k-th unit of this worker -> (K split, tile)
  auto unit_of = [&](int k, int& split, int& t) -> bool {
    if (g.lockstep) {
      if (k >= g.nsub) return false;
      t = worker % tiles;
      split = (worker / tiles) * g.nsub + k;
      return true;
    }
    const int u = worker + k * nworkers;
    if (u >= units) return false;
    split = u / tiles;
    t = u % tiles;
    return true;
  };
  const int chunks = g.ntb * g.nth * g.ntw;  // 64-pixel k-chunks
  const int acc_cols = r3 ? 3 * block_n : 2 * block_n;
  const uint32_t tmem_cols = acc_cols <= 64 ? 64u : acc_cols <= 128 ? 128u : acc_cols <= 256 ? 256u : 512u;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmDY);
    prefetch_tmap(&tmX);
    for (int i = 0; i < stages; ++i) {
      mbar_init(&bars->full[i], 1);
      mbar_init(&bars->empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->tfull[i], 1);
      mbar_init(&bars->tempty[i], 4 * CG);
    }
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
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  // unit -> (split, tap, n_blk, m_blk); chunk range of a split: [c_lo, c_hi)
  auto chunk_lo = [&](int split) { return static_cast<int>(static_cast<long long>(chunks) * split / g.splits); };

  if (warp == 0) {
    // TMA producer: the whole warp walks the ring; lane 0 arms the barrier, lanes 0..nbox-1 each issue one box so
    // the 2 + block_n/64 loads of a stage are issued in parallel instead of back to back by one thread.
    int s = 0;
    uint32_t ph = 0;
    int split, t;
    for (int k = 0; unit_of(k, split, t); ++k) {
      const int m_blk = t % m_blocks;
      t /= m_blocks;
      const int n_blk = t % n_blocks;
      const int tap = (t / n_blocks) * tps;
      const int c_lo = chunk_lo(split), c_hi = chunk_lo(split + 1);
      // chunk -> (tb, th, tw) once per unit, then incremental (no integer divisions inside the ring loop)
      int tw = c_lo % g.ntw, th = (c_lo / g.ntw) % g.nth, tb = c_lo / (g.ntw * g.nth);
      for (int c = c_lo; c < c_hi; ++c) {
        mbar_wait(&bars->empty[s], ph ^ 1);
        uint8_t* sa = smem + s * stage_bytes;
        if (lane == 0 && cta_rank == 0) mbar_expect_tx(&bars->full[s], CG * stage_tx);
        __syncwarp();
        if (lane < 2) {
          const int yw = tw * g.TW * g.dy_mul + g.dy_pw, yh = th * g.TH * g.dy_mul + g.dy_ph;
          const int co0 = (m_blk * CG + cta_rank) * 128 + lane * 64;
          if (CG == 2) tma_load_4d_pair(sa + lane * atom, &tmDY, &bars->full[s], co0, yw, yh, tb * g.TB);
          else tma_load_4d(sa + lane * atom, &tmDY, &bars->full[s], co0, yw, yh, tb * g.TB);
        } else if (lane < 2 + nb_cta / 64) {
          const int j = lane - 2;
          const int xw = tw * g.TW * g.in_mul + g.dx[tap], xh = th * g.TH * g.in_mul + g.dy[tap];
          const int ci0 = n_blk * block_n + cta_rank * nb_cta + j * 64;
          const CUtensorMap* mx = r3 ? &tmX2 : &tmX;
          if (CG == 2) tma_load_4d_pair(sa + a_bytes + j * xslot, mx, &bars->full[s], ci0, xw, xh, tb * g.TB);
          else tma_load_4d(sa + a_bytes + j * xslot, mx, &bars->full[s], ci0, xw, xh, tb * g.TB);
        }
        if (++s == stages) { s = 0; ph ^= 1; }
        if (++tw == g.ntw) {
          tw = 0;
          if (++th == g.nth) { th = 0; ++tb; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128 * CG, block_n, true, true);
    int s = 0;
    uint32_t ph = 0;
    int acc = 0;
    uint32_t pacc = 0;
    int split, t;
    for (int k = 0; cta_rank == 0 && unit_of(k, split, t); ++k) {
      const int nk = chunk_lo(split + 1) - chunk_lo(split);
      if (nk == 0) continue;
      mbar_wait(&bars->tempty[acc], pacc ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * block_n;
      for (int ks = 0; ks < nk; ++ks) {
        mbar_wait(&bars->full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + s * stage_bytes);
          // MN-major, 128B swizzle: LBO = distance to the next 64-channel atom, SBO = 1024 (next 8-pixel group)
          const uint64_t adesc = make_smem_desc_sw128(sa, atom, 1024);
          const int nkk = g.kpx >> 4;              // K=16 pixel MMAs per stage
          for (int j = 0; j < tps; ++j) {
            // tap j of the filter row: same X box, start shifted by j pixel rows (128 B each); the swizzle is a
            // function of absolute smem address bits, so the shifted start needs no base_offset
            const uint64_t bdesc = make_smem_desc_sw128(sa + a_bytes + j * 128, xslot, 1024);
#pragma unroll 4
            for (int k = 0; k < nkk; ++k) {
              // advance 16 pixels = 16 rows of 128 B = 2048 B -> +128 in the (addr>>4) field
              if (CG == 2) umma_bf16_pair(tmem_d + j * block_n, adesc + 128 * k, bdesc + 128 * k, idesc, (ks | k) != 0);
              else umma_bf16(tmem_d + j * block_n, adesc + 128 * k, bdesc + 128 * k, idesc, (ks | k) != 0);
            }
          }
          if (CG == 2) {
            umma_commit_pair(&bars->empty[s]);
            if (ks == nk - 1) umma_commit_pair(&bars->tfull[acc]);
          } else {
            umma_commit(&bars->empty[s]);
            if (ks == nk - 1) umma_commit(&bars->tfull[acc]);
          }
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      if (++acc == nacc) { acc = 0; pacc ^= 1; }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t pacc = 0;
    int split, t;
    for (int k = 0; unit_of(k, split, t); ++k) {
      const int nk = chunk_lo(split + 1) - chunk_lo(split);
      if (nk == 0) continue;
      const int m_blk = t % m_blocks;
      t /= m_blocks;
      const int n_blk = t % n_blocks;
      const int tap0 = (t / n_blocks) * tps;
      const int co = (m_blk * CG + cta_rank) * 128 + row;
      mbar_wait(&bars->tfull[acc], pacc);
      tc_fence_after();
      for (int j = 0; j < tps; ++j) {
        float* dst_row = dw + (static_cast<size_t>(g.slab[tap0 + j]) * g.Cout + co) * g.Cin + n_blk * block_n;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (acc + j) * block_n;
        for (int c0 = 0; c0 < block_n; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
          if (co < g.Cout) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int ci = n_blk * block_n + c0 + 4 * jj;
              if (ci < g.Cin)
                red_add_v4(dst_row + c0 + 4 * jj, __uint_as_float(r[4 * jj]), __uint_as_float(r[4 * jj + 1]),
                           __uint_as_float(r[4 * jj + 2]), __uint_as_float(r[4 * jj + 3]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(&bars->tempty[acc], 0);
        else mbar_arrive(&bars->tempty[acc]);
      }
      if (++acc == nacc) { acc = 0; pacc ^= 1; }
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

// dy: NHWC bf16 [YB, YH, YW, Cout]; x: NHWC bf16 [XB, XH, XW, Cin]; dw: fp32 [nslabs][Cout][Cin], pre-zeroed.
int launch_conv_wgrad(const void* dy, int YB, int YH, int YW, const void* x, int XB, int XH, int XW, WgradGeom g,
                      float* dw, cudaStream_t stream) {
  if (g.ntaps < 1 || g.ntaps > kMaxTaps) return 3;
  if (g.Cin % 4 != 0 || g.Cout % 8 != 0 || g.Cin % 8 != 0) return 2;
  // first pass with 64-pixel K chunks to decide the mode; CTA-pair mode re-tiles with 128-pixel chunks below
  if (!tile_grid(g.GB, g.GH, g.GW, 64, &g.TB, &g.TH, &g.TW)) return 4;
  g.kpx = 64;
  g.ntw = g.GW / g.TW;
  g.nth = g.GH / g.TH;
  g.ntb = (g.GB + g.TB - 1) / g.TB;
  int bn = g.Cin >= 256 ? 256 : g.Cin >= 128 ? 128 : 64;
  // sort taps by (dy, dx); runs of three with equal dy and consecutive dx can share one X box (reuse3)
  for (int i = 1; i < g.ntaps; ++i)
    for (int j = i; j > 0 && (g.dy[j] < g.dy[j - 1] || (g.dy[j] == g.dy[j - 1] && g.dx[j] < g.dx[j - 1])); --j) {
      int t;
      t = g.dy[j]; g.dy[j] = g.dy[j - 1]; g.dy[j - 1] = t;
      t = g.dx[j]; g.dx[j] = g.dx[j - 1]; g.dx[j - 1] = t;
      t = g.slab[j]; g.slab[j] = g.slab[j - 1]; g.slab[j - 1] = t;
    }
  static int reuse_mode = -1;
  if (reuse_mode < 0) {
    const char* e = getenv("LUN_WGRAD_REUSE");
    reuse_mode = e ? atoi(e) : 1;
  }
  bool r3 = reuse_mode && g.ntaps % 3 == 0 && g.TW == 64 && g.TH == 1 && g.TB == 1 && g.in_mul == 1 &&
            g.dy_mul == 1 && g.Cin == 128;   // wider Cin: 256-column tiles in CTA-pair mode are faster (less smem reads per MMA)
  for (int i = 0; r3 && i < g.ntaps; i += 3)
    r3 = g.dy[i + 1] == g.dy[i] && g.dy[i + 2] == g.dy[i] && g.dx[i + 1] == g.dx[i] + 1 && g.dx[i + 2] == g.dx[i] + 2;
  g.reuse3 = r3 ? 1 : 0;
  if (r3) bn = 128;                 // three 128-column accumulators (384 TMEM columns, single set)
  g.block_n = bn;
  if (g.TW * g.in_mul > 256 || g.TH * g.in_mul > 256 || g.TW * g.dy_mul > 256 || g.TH * g.dy_mul > 256) return 5;

  CUtensorMap tmDY, tmX, tmX2;
  int rc = make_tmap_nhwc(&tmDY, dy, YB, YH, YW, g.Cout, g.TW * g.dy_mul, g.TH * g.dy_mul, g.TB, g.dy_mul);
  if (rc) return rc;
  rc = make_tmap_nhwc(&tmX, x, XB, XH, XW, g.Cin, g.TW * g.in_mul, g.TH * g.in_mul, g.TB, g.in_mul);
  if (rc) return rc;
  if (r3) {
    rc = make_tmap_nhwc(&tmX2, x, XB, XH, XW, g.Cin, g.TW + 2, 1, 1, 1);
    if (rc) return rc;
  } else {
    tmX2 = tmX;
  }

  static int pair_mode = -1;
  if (pair_mode < 0) {
    const char* e = getenv("LUN_WGRAD_PAIR");
    pair_mode = e ? atoi(e) : 1;
  }
  const int cg = (pair_mode && !r3 && bn == 256 && g.Cout % 256 == 0 && g.Cin % 256 == 0 && num_sms() % 2 == 0) ? 2 : 1;
  static int kpx_mode = -1;
  if (kpx_mode < 0) {
    const char* e = getenv("LUN_WGRAD_KPX");
    kpx_mode = e ? atoi(e) : 128;
  }
  if (cg == 2 && kpx_mode == 128 && (long)g.GB * g.GH * g.GW % 128 == 0 && g.GW * g.GH >= 128) {
    int tb, th, tw;
    if (tile_grid(g.GB, g.GH, g.GW, 128, &tb, &th, &tw) && tb == 1 && tw * g.in_mul <= 256 && tw * g.dy_mul <= 256) {
      g.kpx = 128;
      g.TB = tb; g.TH = th; g.TW = tw;
      g.ntw = g.GW / g.TW;
      g.nth = g.GH / g.TH;
      g.ntb = (g.GB + g.TB - 1) / g.TB;
      rc = make_tmap_nhwc(&tmDY, dy, YB, YH, YW, g.Cout, g.TW * g.dy_mul, g.TH * g.dy_mul, g.TB, g.dy_mul);
      if (rc) return rc;
      rc = make_tmap_nhwc(&tmX, x, XB, XH, XW, g.Cin, g.TW * g.in_mul, g.TH * g.in_mul, g.TB, g.in_mul);
      if (rc) return rc;
      tmX2 = tmX;
    }
  }
  const int stage_bytes = 2 * g.kpx * 128 + bn / cg / 64 * (r3 ? kWgXSlot : g.kpx * 128);
  const int extra = (int)sizeof(WgBars) + 1024;
  int stages = (227 * 1024 - extra) / stage_bytes;
  if (stages > 8) stages = 8;
  g.stages = stages;
  const int smem_bytes = stages * stage_bytes + extra;
  static bool configured_dev[64] = {};   // the opt-in smem size is a per-device function attribute
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(conv_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess ||
        cudaFuncSetAttribute(conv_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
            cudaSuccess)
      return 8;
    configured = true;
  }
  const int m_blocks = (g.Cout + 128 * cg - 1) / (128 * cg), n_blocks = (g.Cin + bn - 1) / bn;
  const int tiles = m_blocks * n_blocks * (g.ntaps / (r3 ? 3 : 1));
  const int chunks = g.ntb * g.nth * g.ntw;
  // K-splits: make the unit count a multiple of the SM count when the reduction is long enough (perfect balance of
  // the persistent grid), otherwise ~4 units per SM; keep at least 8 k-chunks per unit
  const int workers = num_sms() / cg;               // CTAs (or CTA pairs) that walk the unit list
  int a = tiles, b = workers;
  while (b) { const int r = a % b; a = b; b = r; }
  int splits = workers / a;                         // tiles * splits == lcm(tiles, workers)
  while (splits * tiles < 4 * workers) splits *= 2;
  if (splits > chunks / 8) splits = (4 * workers + tiles - 1) / tiles;
  if (splits > chunks / 8) splits = chunks / 8;
  if (splits < 1) splits = 1;
  g.splits = splits;
  int grid = workers;
  if (grid > tiles * splits) grid = tiles * splits;
  static int lock_mode = -1;
  if (lock_mode < 0) {
    const char* e = getenv("LUN_WGRAD_LOCKSTEP");
    lock_mode = e ? atoi(e) : 1;
  }
  g.lockstep = 0;
  g.nsub = 1;
  if (lock_mode && tiles <= workers) {
    int wpt = workers / tiles;                      // workers per output tile = major K slices
    if (wpt > chunks / 8) wpt = chunks / 8;         // keep at least 8 k-chunks per unit
    if (wpt >= 1 && tiles * wpt * 10 >= workers * 9) {      // at most 10 % of the grid may stay idle
      // sub-units of <= ~192 k-chunks: the drain of one overlaps the main loop of the next, and a TMEM accumulator never
      // sums more than ~25 k pixels before it is flushed (fp32 accumulation error grows with the run length)
      int nsub = (chunks / wpt + 191) / 192;
      if (nsub < 1) nsub = 1;
      if (nsub > 64) nsub = 64;
      g.lockstep = 1;
      g.nsub = nsub;
      g.splits = wpt * nsub;
      grid = tiles * wpt;
    }
  }
  grid *= cg;
  if (cg == 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kWgThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, conv_wgrad_kernel<2>, tmDY, tmX, tmX2, g, dw) != cudaSuccess) return 9;
  } else {
    conv_wgrad_kernel<1><<<grid, kWgThreads, smem_bytes, stream>>>(tmDY, tmX, tmX2, g, dw);
  }
  note_launch(1);
  return cudaGetLastError() == cudaSuccess ? 0 : 9;
}

}  // namespace lun
