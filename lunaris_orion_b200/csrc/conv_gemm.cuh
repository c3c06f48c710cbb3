// Shared host/device declarations for the tcgen05 implicit-GEMM kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace lun {

constexpr int kMaxTaps = 16;

// Epilogue flags
enum : int {
  EPI_BIAS = 1,       // + bias[n] (fp32)
  EPI_LEAKY = 2,      // leaky_relu(slope)
  EPI_STATS = 4,      // per-channel sum / sum-of-squares of the stored (rounded) values -> stats[0:N], stats[N:2N]
  EPI_OUT_F32 = 8,    // store fp32 instead of bf16
  EPI_TMA_STORE = 32, // (set by the launcher) bf16 output leaves through a swizzled smem tile + TMA store
  EPI_DROP_SUM = 256,  // with EPI_STATS on the TMA-store path: stats[0:N] += column sums of dropout_mask * bf16(y * scale)
                       // (the mask is the counter RNG of the elementwise kernels, keyed by the output element index)
  EPI_COL_STATS = 128, // (set by the launcher) statistics are column sums read back from the staged output tile
  EPI_STATS_IMG = 64, // with EPI_STATS: statistics per IMAGE (GroupNorm): stats[b][0:N] sums, stats[b][N:2N] squares;
                      // needs one image per 128-pixel tile (TB == 1)
};

// Implicit-GEMM geometry. GEMM-M enumerates an output grid [GB, GH, GW] in tiles of [TB, TH, TW] (TB*TH*TW == 128).
// For tap t the A operand row of grid point (b, h, w) is the input pixel (b, h*in_mul + dy[t], w*in_mul + dx[t]),
// zero outside the image (TMA out-of-bounds fill); the B operand is weight slab slab[t] = [Cout][Cin] (K-major).
struct ConvGeom {
  int GB, GH, GW;
  int TB, TH, TW;
  int ntb, nth, ntw;
  int in_mul;
  int ntaps;
  int dy[kMaxTaps], dx[kMaxTaps], slab[kMaxTaps];
  int Cin, Cout, block_n;
  int stages;
  int stg_bufs;        // (set by the launcher) output staging tiles per epilogue half: 1 or 2
  // A-tile reuse across the dx taps of a filter row (3 consecutive taps with equal dy and dx, dx+1, dx+2, full-width
  // row tiles): one TW+2-pixel A box per (dy, channel chunk) feeds three MMA groups through descriptor row offsets.
  int a_reuse;
  // output tensor [*, OH, OW, ldo]; grid point (b,h,w) -> pixel (b, h*o_mul + o_ph, w*o_mul + o_pw), channel o_coff + n
  int OH, OW, o_mul, o_ph, o_pw, ldo, o_coff;
  int flags;
  float slope;
  // nphase > 1 (transposed conv, all output phases in ONE launch): the tap list holds nphase runs of ntaps taps; a work
  // item is (phase, pixel tile, channel block) and phase p writes output phase (o_ph, o_pw) = (p >> 1, p & 1)
  int nphase;
  // grouped != 0 (with nphase > 1): the phases are independent GEMM groups instead - phase p uses taps
  // [p * ntaps, (p + 1) * ntaps) and writes output channels o_coff + p * Cout + n (bias index p * Cout + n)
  int grouped;
  // EPI_DROP_SUM: the elementwise dropout this sum replays (seed, 16-bit drop threshold, bf16 1/(1-p))
  unsigned long long drop_seed;
  unsigned int drop_thresh16;
  float drop_scale;
};

// Weight-gradient geometry: dW[slab[t]][co][ci] += sum over grid points of dY[b,h,w,co] * X[b, h*in_mul+dy, w*in_mul+dx, ci]
struct WgradGeom {
  int GB, GH, GW;      // dY grid (pixels reduced over)
  int TB, TH, TW;      // k-chunk tile (TB*TH*TW == 64 pixels)
  int ntb, nth, ntw;
  int in_mul;
  int ntaps;
  int dy[kMaxTaps], dx[kMaxTaps], slab[kMaxTaps];
  int Cin, Cout;       // Cout = GEMM M (blocks of 128), Cin = GEMM N (blocks of block_n)
  int block_n;
  int stages;
  int splits;          // split-K factor: split s covers k-chunks [chunks * s / splits, chunks * (s + 1) / splits)
  int lockstep, nsub;  // lockstep schedule: worker w owns tile w % tiles and splits (w / tiles) * nsub + [0, nsub)
  int dy_mul, dy_ph, dy_pw;  // dY pixel = grid*dy_mul + phase  (transposed-conv phases)
  // reuse3: the three dx taps of a filter row share one (TW+2)-pixel X box; a unit accumulates 3 tiles (one per tap)
  int reuse3;
  int kpx;   // pixels (GEMM-K rows) per pipeline stage: 64, or 128 in CTA-pair mode (halves the per-stage issue overhead)
};

}  // namespace lun
