// extern "C" surface of liblunaris_b200.so (declared in include/lunaris_b200.h).
#include "../../include/lunaris_b200.h"
#include "conv_gemm.cuh"
#include "launch_count.cuh"
#include <atomic>
#include <cuda_bf16.h>

namespace lun {
int launch_conv_fprop(const void* x, int XB, int XH, int XW, const void* wpk, int nslabs, ConvGeom g,
                      const float* bias, void* out, float* stats, cudaStream_t stream);
int launch_conv_wgrad(const void* dy, int YB, int YH, int YW, const void* x, int XB, int XH, int XW, WgradGeom g,
                      float* dw, cudaStream_t stream);
int num_sms();
static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace lun

extern "C" {

int lun_abi_version(void) { return 1; }
int lun_num_sms(void) { return lun::num_sms(); }
long long lun_launch_count(void) { return lun::g_launches.load(); }

static int conv_taps_impl(const void* x, int XB, int XH, int XW, int Cin, const void* w_packed, int nslabs, int Cout,
                          int GB, int GH, int GW, int in_mul, int ntaps, const int* dy, const int* dx, const int* slab,
                          const float* bias, void* out, int OH, int OW, int o_mul, int o_ph, int o_pw, int ldo,
                          int o_coff, int flags, float slope, float* stats, unsigned long long drop_seed,
                          float drop_p, void* stream) {
  if (ntaps < 1 || ntaps > lun::kMaxTaps) return LUN_E_TAPS;
  lun::ConvGeom g{};
  g.GB = GB; g.GH = GH; g.GW = GW;
  g.in_mul = in_mul;
  g.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) { g.dy[t] = dy[t]; g.dx[t] = dx[t]; g.slab[t] = slab[t]; }
  g.Cin = Cin; g.Cout = Cout;
  g.block_n = Cout >= 256 && Cout % 256 == 0 ? 256 : Cout >= 128 && Cout % 128 == 0 ? 128 : Cout % 64 == 0 ? 64 : 32;
  g.OH = OH; g.OW = OW; g.o_mul = o_mul; g.o_ph = o_ph; g.o_pw = o_pw; g.ldo = ldo; g.o_coff = o_coff;
  g.flags = flags; g.slope = slope;
  g.drop_seed = drop_seed;
  g.drop_thresh16 = drop_p > 0.f ? (unsigned int)(drop_p * 65536.f + 0.5f) : 0u;
  g.drop_scale = drop_p > 0.f ? __bfloat162float(__float2bfloat16_rn(1.f / (1.f - drop_p))) : 1.f;
  return lun::launch_conv_fprop(x, XB, XH, XW, w_packed, nslabs, g, bias, out, stats,
                                static_cast<cudaStream_t>(stream));
}

int lun_conv_taps_bf16(const void* x, int XB, int XH, int XW, int Cin, const void* w_packed, int nslabs, int Cout,
                       int GB, int GH, int GW, int in_mul, int ntaps, const int* dy, const int* dx, const int* slab,
                       const float* bias, void* out, int OH, int OW, int o_mul, int o_ph, int o_pw, int ldo,
                       int o_coff, int flags, float slope, float* stats, void* stream) {
  return conv_taps_impl(x, XB, XH, XW, Cin, w_packed, nslabs, Cout, GB, GH, GW, in_mul, ntaps, dy, dx, slab, bias, out,
                        OH, OW, o_mul, o_ph, o_pw, ldo, o_coff, flags & ~lun::EPI_DROP_SUM, slope, stats, 0ull, 0.f, stream);
}

int lun_conv_taps_dropsum_bf16(const void* x, int XB, int XH, int XW, int Cin, const void* w_packed, int nslabs,
                               int Cout, int GB, int GH, int GW, int in_mul, int ntaps, const int* dy, const int* dx,
                               const int* slab, void* out, int OH, int OW, int ldo, float* colsum,
                               unsigned long long drop_seed, float drop_p, void* stream) {
  // dense bf16 output through the TMA-store path; colsum[0:Cout] += sum over pixels of mask * bf16(out * 1/(1-p))
  return conv_taps_impl(x, XB, XH, XW, Cin, w_packed, nslabs, Cout, GB, GH, GW, in_mul, ntaps, dy, dx, slab, nullptr, out,
                        OH, OW, 1, 0, 0, ldo, 0, lun::EPI_STATS | lun::EPI_DROP_SUM, 1.f, colsum, drop_seed, drop_p,
                        stream);
}

// defined in convt_halo_sm100.cu
extern "C" int lun_convT4x4s2_halo_bf16(const void* x, int B, int H, int W, int Cin, const void* w_packed, int Cout,
                                        const float* bias, void* out, float* img_stats, void* stream);

int lun_convT4x4s2_bf16(const void* x, int B, int H, int W, int Cin, const void* w_packed, int Cout, const float* bias,
                        void* out, float* img_stats, void* stream) {
  // thin stages: fused-phase halo kernel; otherwise the tap-list kernel with the four phases batched into one launch
  if ((Cout == 32 || Cout == 64) && Cin % 64 == 0 && H % 16 == 0 && W % 8 == 0)
    return lun_convT4x4s2_halo_bf16(x, B, H, W, Cin, w_packed, Cout, bias, out, img_stats, stream);
  lun::ConvGeom g{};
  g.GB = B; g.GH = H; g.GW = W;
  g.in_mul = 1;
  g.ntaps = 4;
  g.nphase = 4;
  for (int p = 0; p < 4; ++p) {
    const int ph = p >> 1, pw = p & 1;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        // output row 2j+ph gathers input rows j+dh through filter row kh: dh = ph - a, kh = (1 - ph) + 2a (same in w)
        const int t = p * 4 + a * 2 + b;
        g.dy[t] = ph - a;
        g.dx[t] = pw - b;
        g.slab[t] = ((1 - ph) + 2 * a) * 4 + (1 - pw) + 2 * b;
      }
  }
  g.Cin = Cin; g.Cout = Cout;
  g.block_n = Cout >= 256 && Cout % 256 == 0 ? 256 : Cout >= 128 && Cout % 128 == 0 ? 128 : Cout % 64 == 0 ? 64 : 32;
  g.OH = 2 * H; g.OW = 2 * W; g.o_mul = 2; g.o_ph = 0; g.o_pw = 0; g.ldo = Cout; g.o_coff = 0;
  g.flags = (bias ? LUN_EPI_BIAS : 0) | (img_stats ? (LUN_EPI_STATS | LUN_EPI_STATS_IMG) : 0);
  g.slope = 1.f;
  return lun::launch_conv_fprop(x, B, H, W, w_packed, 16, g, bias, out, img_stats, static_cast<cudaStream_t>(stream));
}

int lun_head_linear_bf16(const void* x, int rows, int heads, int cin, const void* w, int cout, const float* bias,
                         void* out, void* stream) {
  // out[r, h*cout + n] = bias[h*cout + n] + sum_k x[r, h*cin + k] * w[h][n][k]: `heads` independent GEMMs in one launch.
  // x is read as an NHWC tensor [rows, 1, heads, cin]: group h is the filter tap dx = h of a 1-pixel output grid.
  if (heads < 1 || heads > lun::kMaxTaps) return LUN_E_TAPS;
  lun::ConvGeom g{};
  g.GB = rows; g.GH = 1; g.GW = 1;
  g.in_mul = 1;
  g.ntaps = 1;
  g.nphase = heads;
  g.grouped = 1;
  for (int h = 0; h < heads; ++h) { g.dy[h] = 0; g.dx[h] = h; g.slab[h] = h; }
  g.Cin = cin; g.Cout = cout;
  g.block_n = cout >= 256 && cout % 256 == 0 ? 256 : cout >= 128 && cout % 128 == 0 ? 128 : cout % 64 == 0 ? 64 : 32;
  g.OH = 1; g.OW = 1; g.o_mul = 1; g.o_ph = 0; g.o_pw = 0; g.ldo = heads * cout; g.o_coff = 0;
  g.flags = bias ? LUN_EPI_BIAS : 0;
  g.slope = 1.f;
  return lun::launch_conv_fprop(x, rows, 1, heads, w, heads, g, bias, out, nullptr, static_cast<cudaStream_t>(stream));
}

int lun_wgrad_taps_bf16(const void* dy, int YB, int YH, int YW, int Cout, int dy_mul, int dy_ph, int dy_pw,
                        const void* x, int XB, int XH, int XW, int Cin, int in_mul, int GB, int GH, int GW,
                        int ntaps, const int* tdy, const int* tdx, const int* slab, float* dw, void* stream) {
  if (ntaps < 1 || ntaps > lun::kMaxTaps) return LUN_E_TAPS;
  lun::WgradGeom g{};
  g.GB = GB; g.GH = GH; g.GW = GW;
  g.in_mul = in_mul;
  g.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) { g.dy[t] = tdy[t]; g.dx[t] = tdx[t]; g.slab[t] = slab[t]; }
  g.Cin = Cin; g.Cout = Cout;
  g.dy_mul = dy_mul; g.dy_ph = dy_ph; g.dy_pw = dy_pw;
  return lun::launch_conv_wgrad(dy, YB, YH, YW, x, XB, XH, XW, g, dw, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
