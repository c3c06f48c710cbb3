/* lunaris_b200 C ABI — the drop-in boundary of the B200-native hybrid training step.
 *
 * Every entry point takes raw device pointers, plain sizes and a cudaStream_t (passed as void*), allocates
 * nothing, never throws, and returns 0 on success or a small positive error code (see LUN_E_*).
 * Activations are NHWC bf16 unless noted; parameters are handed over already packed.
 * Each function names the reference call site (file:line in MeryylleA/Lunaris-Orion) whose math it replaces.
 */
#ifndef LUNARIS_B200_H
#define LUNARIS_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define LUN_OK 0
#define LUN_E_SHAPE 2      /* channel counts / block sizes not supported */
#define LUN_E_TAPS 3       /* tap list empty or longer than 16 */
#define LUN_E_GRID 4       /* spatial grid is not a power of two */
#define LUN_E_BOX 5        /* TMA box would exceed 256 elements */
#define LUN_E_STATS 6
#define LUN_E_ALIGN 7
#define LUN_E_ATTR 8       /* cudaFuncSetAttribute failed */
#define LUN_E_LAUNCH 9     /* kernel launch failed */
#define LUN_E_DRIVER 101   /* cuTensorMapEncodeTiled entry point unavailable */
#define LUN_E_TMAP 102     /* tensor-map encode failed */

/* epilogue flags for lun_conv_taps_bf16 */
#define LUN_EPI_BIAS 1
#define LUN_EPI_LEAKY 2
#define LUN_EPI_STATS 4
#define LUN_EPI_OUT_F32 8
#define LUN_EPI_TANH 16

/* Library / device probe: returns the SM count of the current device (148 on B200), <=0 on error. */
int lun_num_sms(void);
/* ABI version of this header. */
int lun_abi_version(void);

/* Implicit-GEMM convolution in tap-list form (tcgen05 + TMA).  out[b, h*o_mul+o_ph, w*o_mul+o_pw, o_coff+n] =
 *   epi( sum_t sum_c x[b, h*in_mul+dy[t], w*in_mul+dx[t], c] * w_packed[slab[t]][n][c] )   for (b,h,w) in [GB,GH,GW]
 * Covers F.conv2d (any k/stride/pad), each output phase of F.conv_transpose2d, nn.Linear (GH=GW=1), and all of
 * their data-gradients (with re-packed weights).  Reference: lunar_evaluator.py:242,249,133,134,255,79-100;
 * lunar_generate.py:36,41,95,102,109,116,124,125,165,169-187.
 * stats (if LUN_EPI_STATS): fp32 [2*Cout], accumulated with atomics: sum and sum of squares per output channel. */
int lun_conv_taps_bf16(const void* x, int XB, int XH, int XW, int Cin, const void* w_packed, int nslabs, int Cout,
                       int GB, int GH, int GW, int in_mul, int ntaps, const int* dy, const int* dx, const int* slab,
                       const float* bias, void* out, int OH, int OW, int o_mul, int o_ph, int o_pw, int ldo,
                       int o_coff, int flags, float slope, float* stats, void* stream);

/* Weight gradient in tap-list form (tcgen05, MN-major operands, split-K with fp32 red.add):
 *   dw[slab[t]][co][ci] += sum_{(b,h,w) in [GB,GH,GW]} dy[b, h*dy_mul+dy_ph, w*dy_mul+dy_pw, co]
 *                                                     * x[b, h*in_mul+tdy[t], w*in_mul+tdx[t], ci]
 * dw is fp32 and must be zeroed by the caller.  Reference: autograd of the call sites above. */
int lun_wgrad_taps_bf16(const void* dy, int YB, int YH, int YW, int Cout, int dy_mul, int dy_ph, int dy_pw,
                        const void* x, int XB, int XH, int XW, int Cin, int in_mul, int GB, int GH, int GW,
                        int ntaps, const int* tdy, const int* tdx, const int* slab, float* dw, void* stream);

#ifdef __cplusplus
}
#endif
#endif
