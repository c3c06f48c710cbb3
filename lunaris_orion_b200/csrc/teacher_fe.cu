// PixelArtFeatureExtractor front end (lunar_evaluator.py:71-96, 105-110): HBM-bound CUDA-core kernels.
//   fe_conv1_kernel : conv3x3 3->32 + bias + LeakyReLU(0.2) on NCHW fp32 images -> NHWC bf16, + BN batch statistics
//   fe_branch_kernel: BN(conv1) applied on load, three depthwise convs (3x3, 5x5, 3x3) each followed by a pointwise
//                     32->64 conv + bias + LeakyReLU, written into the 192-channel concat buffer (torch.cat of :110)
// Arithmetic mirrors the reference's bf16 autocast: conv operands rounded to bf16, fp32 accumulate, bf16 results.
#include "../../include/lunaris_b200.h"
#include "elem_common.cuh"
#include "launch_count.cuh"

namespace lun {

// ------------------------------------------------------------------------------------------- conv1 3->32
__global__ void __launch_bounds__(256) fe_conv1_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ bias, bf16* __restrict__ y,
                                                       float* __restrict__ stats, int B, int H, int W, float slope) {
  __shared__ float sw[27][32];
  __shared__ float sb[32];
  __shared__ float sstat[64];
  for (int i = threadIdx.x; i < 27 * 32; i += 256) {
    const int o = i % 32, k = i / 32;          // k = (c*3 + kh)*3 + kw ; reference layout [o][c][kh][kw]
    sw[k][o] = rbf(w[o * 27 + k]);
  }
  if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
  if (threadIdx.x < 64) sstat[threadIdx.x] = 0.f;
  __syncthreads();
  const long HW = (long)H * W;
  const long total = (long)B * HW;
  const long p = (long)blockIdx.x * 256 + threadIdx.x;
  const bool valid = p < total;
  float acc[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) acc[o] = 0.f;
  if (valid) {
    const int b = (int)(p / HW), hw = (int)(p % HW), h = hw / W, wq = hw % W;
    for (int c = 0; c < 3; ++c) {
      const float* xc = x + ((long)b * 3 + c) * HW;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int ih = h + kh - 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int iw = wq + kw - 1;
          float v = 0.f;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = rbf(__ldg(xc + (long)ih * W + iw));
          const float* wr = sw[(c * 3 + kh) * 3 + kw];
#pragma unroll
          for (int o = 0; o < 32; ++o) acc[o] += v * wr[o];
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 32; ++o) {
      const float t = rbf(acc[o] + sb[o]);    // conv output is a bf16 tensor; LeakyReLU runs on it
      acc[o] = rbf(t > 0.f ? t : t * slope);
    }
    bf16* dst = y + p * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v8[i] = acc[8 * j + i];
      store8(dst + 8 * j, v8);
    }
  } else {
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = 0.f;
  }
  // per-channel sum / sumsq: transpose-reduce over the warp, then shared and global atomics
  float s1[32], s2[32];
#pragma unroll
  for (int o = 0; o < 32; ++o) {
    s1[o] = acc[o];
    s2[o] = acc[o] * acc[o];
  }
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = lane & off;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float a1 = upper ? s1[i] : s1[i + off], k1 = upper ? s1[i + off] : s1[i];
      s1[i] = k1 + __shfl_xor_sync(0xffffffffu, a1, off);
      const float a2 = upper ? s2[i] : s2[i + off], k2 = upper ? s2[i + off] : s2[i];
      s2[i] = k2 + __shfl_xor_sync(0xffffffffu, a2, off);
    }
  }
  atomicAdd(&sstat[lane], s1[0]);
  atomicAdd(&sstat[32 + lane], s2[0]);
  __syncthreads();
  if (threadIdx.x < 64) atomicAdd(stats + threadIdx.x, sstat[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------- branches
struct FeBranchW {
  const float* dw_w[3];   // [32][1][k][k]  (edge k=3, color k=5, detail k=3)
  const float* dw_b[3];   // [32]
  const float* pw_w[3];   // [64][32]
  const float* pw_b[3];   // [64]
};

constexpr int kT = 16;          // 16x16 pixel tile
constexpr int kHalo = 2;
constexpr int kTP = kT + 2 * kHalo;   // 20
constexpr int kPlane = kTP * kTP + 1; // +1 padding
constexpr int kDPitch = 80;     // bytes per pixel row of the bf16 [pixel][32 ch] staging tile (conflict-free ldmatrix)

// Depthwise K x K on the fp32 smem planes for one pixel: weights padded to rows of 4 (K = 3) or 8 (K = 5) floats so a
// filter row is one or two 16-byte broadcast loads.
template <int K>
__device__ __forceinline__ float dw_pixel(const float* __restrict__ tp, const float* __restrict__ wp) {
  constexpr int WR = K == 3 ? 4 : 8;
  float a = 0.f;
#pragma unroll
  for (int kh = 0; kh < K; ++kh) {
    float w[WR];
    *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(wp + kh * WR);
    if (K == 5) *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(wp + kh * WR + 4);
#pragma unroll
    for (int kw = 0; kw < K; ++kw) a = fmaf(tp[kh * kTP + kw], w[kw], a);
  }
  return a;
}

// One block = one 16x16 pixel tile of one image, thread = pixel.
//   phase 1: BN(conv1)-applied input tile with a 2-pixel halo -> fp32 channel planes in smem
//   per branch: depthwise conv on CUDA cores (thread-local, 32 channels) -> bf16 [pixel][32] staging rows ->
//               pointwise 32->64 on the tensor cores (mma.sync m16n8k16: each warp multiplies ITS 32 pixels, so only a
//               warp-level sync separates the two steps) -> + bias, LeakyReLU -> concat buffer.
// The pointwise conv is 82 % of the FLOPs of this kernel; on the FMA pipe it made the kernel 5x slower than its HBM time.
__global__ void __launch_bounds__(256) fe_branch_kernel(const bf16* __restrict__ y0, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, FeBranchW wt,
                                                        bf16* __restrict__ cat, int H, int W, float slope) {
  extern __shared__ __align__(16) float sm[];
  float* tile = sm;                           // [32][kPlane]
  float* s_dw = tile + 32 * kPlane;           // [3][32][40] padded depthwise filters (32 * 401 floats: 16-byte aligned)
  float* s_dwb = s_dw + 3 * 32 * 40;          // [3][32]
  float* s_pwb = s_dwb + 96;                  // [3][64]
  unsigned char* s_pw = reinterpret_cast<unsigned char*>(s_pwb + 192);      // [3][64 out][kDPitch] bf16 pointwise weights
  unsigned char* s_d = s_pw + 3 * 64 * kDPitch;                             // [256 px][kDPitch] bf16 depthwise outputs
  const int b = blockIdx.z, h0 = blockIdx.y * kT, w0 = blockIdx.x * kT;
  const int ksz[3] = {3, 5, 3};
  for (int i = threadIdx.x; i < 3 * 32 * 40; i += 256) {
    const int br = i / 1280, c = (i / 40) % 32, r = (i % 40) / 8, col = i % 8;
    const int k = ksz[br], wr = k == 3 ? 4 : 8;
    // rows are stored with pitch wr (4 for 3x3, 8 for 5x5); slot (r, col) of the padded [5][8] block
    float v = 0.f;
    const int flat = r * 8 + col;             // position inside this filter's 40-float block
    const int rr = flat / wr, cc = flat % wr;
    if (rr < k && cc < k) v = rbf(wt.dw_w[br][c * k * k + rr * k + cc]);
    s_dw[i] = v;
  }
  for (int i = threadIdx.x; i < 96; i += 256) s_dwb[i] = wt.dw_b[i / 32][i % 32];
  for (int i = threadIdx.x; i < 192; i += 256) s_pwb[i] = wt.pw_b[i / 64][i % 64];
  for (int i = threadIdx.x; i < 3 * 64 * 32; i += 256) {
    const int br = i / 2048, o = (i / 32) % 64, c = i % 32;
    *reinterpret_cast<bf16*>(s_pw + (br * 64 + o) * kDPitch + c * 2) = __float2bfloat16_rn(wt.pw_w[br][o * 32 + c]);
  }
  // BN-applied input tile with halo (zero padding is applied AFTER BatchNorm, as in the reference)
  for (int i = threadIdx.x; i < kTP * kTP * 4; i += 256) {
    const int cg8 = i & 3, px = i >> 2;
    const int ty = px / kTP, tx = px % kTP;
    const int ih = h0 + ty - kHalo, iw = w0 + tx - kHalo;
    float v[8];
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      load8(y0 + (((long)b * H + ih) * W + iw) * 32 + cg8 * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = rbf(v[j] * scale[cg8 * 8 + j] + shift[cg8 * 8 + j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) tile[(cg8 * 8 + j) * kPlane + px] = v[j];
  }
  __syncthreads();
  const int ty = threadIdx.x / kT, tx = threadIdx.x % kT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
  const int mat = lane >> 3, lrow = lane & 7;
  const uint32_t d_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_d)) + warp * 32 * kDPitch;
  unsigned char* d_row = s_d + threadIdx.x * kDPitch;
  for (int br = 0; br < 3; ++br) {
    // ---- depthwise: this pixel, 32 channels -> bf16 row of the staging tile
    const int pad = ksz[br] / 2;
    const float* tp0 = tile + (ty + kHalo - pad) * kTP + (tx + kHalo - pad);
    const float* wp0 = s_dw + br * 32 * 40;
#pragma unroll 4
    for (int c = 0; c < 32; c += 2) {
      float a0, a1;
      if (br == 1) {
        a0 = dw_pixel<5>(tp0 + c * kPlane, wp0 + c * 40);
        a1 = dw_pixel<5>(tp0 + (c + 1) * kPlane, wp0 + (c + 1) * 40);
      } else {
        a0 = dw_pixel<3>(tp0 + c * kPlane, wp0 + c * 40);
        a1 = dw_pixel<3>(tp0 + (c + 1) * kPlane, wp0 + (c + 1) * 40);
      }
      const __nv_bfloat162 pk = __floats2bfloat162_rn(a0 + s_dwb[br * 32 + c], a1 + s_dwb[br * 32 + c + 1]);
      *reinterpret_cast<__nv_bfloat162*>(d_row + c * 2) = pk;
    }
    __syncwarp();
    // ---- pointwise 32 -> 64 for the warp's 32 pixels: A = staged rows (2 m-tiles x 2 k-steps), B = weights
    uint32_t afr[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        // matrices {px 0-7,k 0-7}, {px 8-15,k 0-7}, {px 0-7,k 8-15}, {px 8-15,k 8-15}
        const uint32_t addr = d_base + (mt * 16 + lrow + (mat & 1) * 8) * kDPitch + (ks * 16 + (mat >> 1) * 8) * 2;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(afr[mt][ks][0]), "=r"(afr[mt][ks][1]), "=r"(afr[mt][ks][2]), "=r"(afr[mt][ks][3])
                     : "r"(addr));
      }
    const uint32_t w_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_pw)) + br * 64 * kDPitch;
    const long pix0 = ((long)b * H + h0 + 2 * warp) * W + w0;       // warp = tile rows 2w, 2w+1 (16 pixels each)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      // B for n-tile nt: matrices {k 0-7}, {k 8-15}, {k 16-23}, {k 24-31} of outputs 8nt..8nt+7
      uint32_t bfr[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                   : "=r"(bfr[0]), "=r"(bfr[1]), "=r"(bfr[2]), "=r"(bfr[3])
                   : "r"(w_base + (nt * 8 + lrow) * kDPitch + mat * 16));
      const float b0 = s_pwb[br * 64 + nt * 8 + 2 * t4], b1 = s_pwb[br * 64 + nt * 8 + 2 * t4 + 1];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float d[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
          asm volatile(
              "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
              "{%0, %1, %2, %3};"
              : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
              : "r"(afr[mt][ks][0]), "r"(afr[mt][ks][1]), "r"(afr[mt][ks][2]), "r"(afr[mt][ks][3]), "r"(bfr[2 * ks]),
                "r"(bfr[2 * ks + 1]));
        // d0,d1: pixel g of m-tile mt (tile row 2*warp + mt, column g), outputs 8nt+2t4, +1;  d2,d3: column g+8
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float t = rbf(d[e] + ((e & 1) ? b1 : b0));
          o[e] = t > 0.f ? t : t * slope;
        }
        bf16* dst = cat + (pix0 + (long)mt * W + g) * 192 + br * 64 + nt * 8 + 2 * t4;
        *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(o[0], o[1]);
        *reinterpret_cast<__nv_bfloat162*>(dst + 8 * 192) = __floats2bfloat162_rn(o[2], o[3]);
      }
    }
    __syncwarp();                                  // the staging rows are rewritten by the next branch
  }
}

}  // namespace lun

using namespace lun;

extern "C" {

int lun_fe_conv1(const float* x_nchw, const float* w, const float* bias, void* y, float* stats, int B, int H, int W,
                 float slope, void* stream) {
  const long total = (long)B * H * W;
  fe_conv1_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_nchw, w, bias, (bf16*)y, stats, B,
                                                                               H, W, slope);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

int lun_fe_branches(const void* y0, const float* scale, const float* shift, const float* const* dw_w,
                    const float* const* dw_b, const float* const* pw_w, const float* const* pw_b, void* cat, int B,
                    int H, int W, float slope, void* stream) {
  if (H % kT || W % kT) return LUN_E_SHAPE;
  FeBranchW wt;
  for (int i = 0; i < 3; ++i) {
    wt.dw_w[i] = dw_w[i]; wt.dw_b[i] = dw_b[i]; wt.pw_w[i] = pw_w[i]; wt.pw_b[i] = pw_b[i];
  }
  const int smem = (32 * kPlane + 3 * 32 * 40 + 96 + 192) * (int)sizeof(float) + (3 * 64 + 256) * kDPitch;
  static bool configured_dev[64] = {};   // the opt-in smem size is a per-device function attribute
  int cur_dev = 0;
  cudaGetDevice(&cur_dev);
  bool& configured = configured_dev[cur_dev & 63];
  if (!configured) {
    if (cudaFuncSetAttribute(fe_branch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return LUN_E_ATTR;
    configured = true;
  }
  dim3 grid(W / kT, H / kT, B);
  fe_branch_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)y0, scale, shift, wt, (bf16*)cat, H, W,
                                                             slope);
  lun::note_launch(1);
  return cudaGetLastError() == cudaSuccess ? LUN_OK : LUN_E_LAUNCH;
}

}  // extern "C"
