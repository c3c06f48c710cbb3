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
#define LUN_EPI_DROP_SUM 256 /* internal: set by lun_conv_taps_dropsum_bf16 */
#define LUN_EPI_STATS_IMG 64 /* with LUN_EPI_STATS: per-image sums, stats = fp32 [GB][2][Cout] (GroupNorm); the output
                              * grid of one image must hold at least 128 pixels */

/* Library / device probe: returns the SM count of the current device (148 on B200), <=0 on error. */
int lun_num_sms(void);
/* ABI version of this header. */
int lun_abi_version(void);
/* Number of kernels this library has launched in the calling process (monotonic). */
long long lun_launch_count(void);

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

/* Same implicit GEMM, dense bf16 output, with the backward of an elementwise Dropout fused into the epilogue:
 * colsum[0:Cout] += sum over output pixels of mask * bf16(out * 1/(1-p)), where mask is the counter-RNG dropout mask of
 * lun_proj_expand_bf16 for (drop_seed, element index = pixel * ldo + channel). Used for the data-gradient of conv2
 * (lunar_evaluator.py:249) whose result feeds proj_drop's backward and the proj bias gradient (:224-225): the
 * separate pass over the 1 GB gradient tensor disappears. colsum: fp32 [2*Cout], first half written. */
int lun_conv_taps_dropsum_bf16(const void* x, int XB, int XH, int XW, int Cin, const void* w_packed, int nslabs,
                               int Cout, int GB, int GH, int GW, int in_mul, int ntaps, const int* dy, const int* dx,
                               const int* slab, void* out, int OH, int OW, int ldo, float* colsum,
                               unsigned long long drop_seed, float drop_p, void* stream);

/* ConvTranspose2d(k=4, s=2, p=1) for thin stages (lunar_generate.py:181-187), all four output phases in one launch:
 * x [B,H,W,Cin] bf16 NHWC -> out [B,2H,2W,Cout] bf16 (+ bias). A CTA loads the 18x10 halo of a 16x8 input block once
 * per 64-channel chunk and reads every tap shift from it through UMMA descriptor offsets; weights stay resident.
 * w_packed: [16][Cout][Cin] bf16 (slab kh*4+kw); img_stats: optional fp32 [B][2][Cout] per-image sums / sums of
 * squares of the stored output (GroupNorm), accumulated with atomics. H % 16 == 0, W % 8 == 0, Cin % 64 == 0,
 * Cout in {32, 64}. */
int lun_convT4x4s2_halo_bf16(const void* x, int B, int H, int W, int Cin, const void* w_packed, int Cout,
                             const float* bias, void* out, float* img_stats, void* stream);

/* ConvTranspose2d(k=4, s=2, p=1) forward, any stage (lunar_generate.py:169-187), ONE launch: the halo kernel above when
 * its shape limits hold, otherwise the tap-list kernel with the four output phases batched as an extra work-item
 * dimension (phase, pixel tile, channel block). Same arguments; img_stats needs H*W >= 128. */
int lun_convT4x4s2_bf16(const void* x, int B, int H, int W, int Cin, const void* w_packed, int Cout, const float* bias,
                        void* out, float* img_stats, void* stream);

/* Weight gradient in tap-list form (tcgen05, MN-major operands, split-K with fp32 red.add):
 *   dw[slab[t]][co][ci] += sum_{(b,h,w) in [GB,GH,GW]} dy[b, h*dy_mul+dy_ph, w*dy_mul+dy_pw, co]
 *                                                     * x[b, h*in_mul+tdy[t], w*in_mul+tdx[t], ci]
 * dw is fp32 and must be zeroed by the caller.  Reference: autograd of the call sites above. */
int lun_wgrad_taps_bf16(const void* dy, int YB, int YH, int YW, int Cout, int dy_mul, int dy_ph, int dy_pw,
                        const void* x, int XB, int XH, int XW, int Cin, int in_mul, int GB, int GH, int GW,
                        int ntaps, const int* tdy, const int* tdx, const int* slab, float* dw, void* stream);

/* ---- Bandwidth-bound Teacher kernels (NHWC bf16 [B, HW, C]; C % 8 == 0) -------------------------------------- */

/* Per-channel sum and sum of squares of x[P, C] accumulated into stats[2*C] (caller zeroes).
 * Reference: the batch statistics nn.BatchNorm2d computes in training mode (lunar_evaluator.py:74,81,88,95,102). */
int lun_channel_stats_bf16(const void* x, long P, int C, float* stats, void* stream);

/* BatchNorm2d training-mode finalize: stats -> (scale, shift) with y = x*scale + shift, (mean, rstd) for backward,
 * and n_updates momentum updates of running_mean / running_var (unbiased) / num_batches_tracked.
 * n_updates reproduces the reference's 1x (no_grad pass), 1x or 2x (checkpoint recompute) updates, SURVEY.md 0.4.
 * Reference: lunar_evaluator.py:244,251,256 (nn.BatchNorm2d inside ExpertBlock), :74-102. */
int lun_bn_finalize(const float* stats, double n, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, long long* num_batches_tracked, int n_updates, float momentum, float eps,
                    float* scale, float* shift, float* mean, float* rstd, int C, void* stream);

/* y = act( ls * drop( drop2d( x*scale + shift ) ) + identity' )  with every optional stage switched by a null
 * pointer / zero probability: mask2d [B,C] is the Dropout2d keep-mask (values 0 or 1/(1-p)); drop_p>0 applies
 * elementwise dropout from the counter-based RNG (seed); ls [C] is ExpertBlock.layer_scale; identity [B,HW,C] is the
 * residual (optionally through its own BN affine id_scale/id_shift) followed by leaky_relu(slope); pool [B,C]
 * receives the per-image channel sums of y (AdaptiveAvgPool2d numerators).
 * Reference: lunar_evaluator.py:244-245,251-253 (BN + Dropout2d), :264 (layer_scale), :275 (residual + leaky_relu),
 * :97 (Dropout), :355,366,378,390,435 (pooling). */
int lun_affine_fwd_bf16(const void* x, const float* scale, const float* shift, const float* mask2d, const float* ls,
                        const void* identity, const float* id_scale, const float* id_shift, void* y, float* pool,
                        unsigned long long seed, float drop_p, int B, int HW, int C, float slope, void* stream);

/* Backward of the ExpertBlock tail, pass 1: dpre = dout * leaky'(out) (dout may be the broadcast pooled gradient
 * gpool[B,C]); accumulates t1[c] = sum dpre*mask2d, t2[c] = sum dpre*mask2d*xhat (xhat from bn_in, mean, rstd).
 * Reference: autograd of lunar_evaluator.py:263-264,275. */
int lun_block_bwd_reduce_bf16(const void* dout, const float* gpool, const void* out, const void* bn_in,
                              const float* mean, const float* rstd, const float* mask2d, void* dpre, float* t1,
                              float* t2, int B, int HW, int C, float slope_out, void* stream);

/* Pass 2: dz = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)) * leaky'(bn_in), g = dpre*ls*mask2d; dbias[c] += sum dz.
 * dz is the gradient at the conv output feeding conv dgrad / wgrad. */
int lun_block_bwd_apply_bf16(const void* dpre, const float* gpool, const void* out, const void* bn_in,
                             const float* mean, const float* rstd, const float* gamma, const float* ls,
                             const float* mask2d, const float* t1, const float* t2, void* dz, float* dbias, int B,
                             int HW, int C, float slope_out, float slope_a, void* stream);

/* Block-local attention exactly as the reference executes it (chunk-index scatter, lunar_evaluator.py:203-216):
 * att_small[b, i, :] for i < N/32 + 31 is the only non-zero part of the pre-proj tensor. qkv: [B, N, 3C] with
 * channel order [3][heads][hd]. drop_p is attn_drop (:212). */
int lun_attn_ref_rows_bf16(const void* qkv, void* att_small, int B, int N, int C, int heads, int nq_pad,
                           unsigned long long seed, float drop_p, void* stream);

/* Same attention with the dead work removed: only the N/32+31 query rows that survive the scatter are projected
 * (q_small [B, nq_pad, C], rows gathered by lun_gather_query_rows_bf16 then multiplied by Wq), K and V come from a
 * [B, N, 2C] tensor (channel blocks K | V). Results are identical to lun_attn_ref_rows_bf16 on the full qkv. */
int lun_attn_ref_rows_split_bf16(const void* q_small, const void* kv, void* att_small, int B, int N, int C, int heads,
                                 int nq_pad, unsigned long long seed, float drop_p, void* stream);
/* out[b, i, :] = x[b, qtok(i), :] for the query tokens the as-executed attention uses: qtok(i) = 32 i for
 * i < N/32 - 1, else the 32 tokens of the last chunk (lunar_evaluator.py:203-216). */
int lun_gather_query_rows_bf16(const void* x, void* out, int B, int N, int C, int nq_pad, void* stream);

/* As-executed attention WITHOUT forming K or V. Only one query per 32-token chunk survives the reference's
 * chunk-index scatter, so with qt = (Wk_h^T Wq_h) x_q / sqrt(hd) + Wk_h^T bq_h / sqrt(hd) (a small GEMM, head-major
 * [B, nq_pad, 8*C]) the scores are qt_h . x_j and the head output is Wv_h (sum_j p_j x_j) + bv_h. This kernel computes
 * xbar[b,i,h,:] = sum_j softmax_j(qt[b,i,h,:] . x_j) x_j over the 32 tokens of the row's chunk, where
 * x = drop2d(bn(y)) is formed on load from the pre-BatchNorm tensor y (scale/shift [C], mask2d [B,C] or null).
 * Same function as lun_attn_ref_rows_bf16 composed with the qkv conv (lunar_evaluator.py:153-156,203-216). */
int lun_attn_fold_rows_bf16(const void* y, const float* scale, const float* shift, const float* mask2d,
                            const void* qt, void* xbar, int B, int N, int C, int heads, int nq_pad,
                            unsigned long long seed, float drop_p, void* stream);
/* out[b,i,:] = drop2d(bn(y[b, qtok(i), :])) — the surviving query rows with the attention-input affine applied. */
int lun_gather_query_rows_affine_bf16(const void* y, const float* scale, const float* shift, const float* mask2d,
                                      void* out, int B, int N, int C, int nq_pad, void* stream);

/* y[b,p,:] = proj_drop( p < nq ? proj_small[b,p,:] : bias )  (lunar_evaluator.py:224-225 on the mostly-zero input). */
int lun_proj_expand_bf16(const void* proj_small, const float* bias, void* y, int B, int HW, int C, int nq, int nq_pad,
                         unsigned long long seed, float drop_p, void* stream);

/* Backward of proj_drop: dpo = mask*dh2; dbias[c] += column sums; rows p < nq gathered into dpo_small. */
int lun_proj_bwd_gather_bf16(const void* dh2, void* dpo_small, float* dbias, int B, int HW, int C, int nq, int nq_pad,
                             unsigned long long seed, float drop_p, void* stream);

/* PixelArtFeatureExtractor.conv1: conv3x3 3->32 + LeakyReLU on NCHW fp32 images -> NHWC bf16, + BN statistics
 * (lunar_evaluator.py:71-75). w is the reference layout [32,3,3,3] fp32. */
int lun_fe_conv1(const float* x_nchw, const float* w, const float* bias, void* y, float* stats, int B, int H, int W,
                 float slope, void* stream);

/* The three branches (depthwise 3x3 / 5x5 / 3x3 -> pointwise 32->64 -> LeakyReLU) written into the 192-channel concat
 * buffer; BN of conv1 (scale, shift) is applied on load (lunar_evaluator.py:77-96,106-110).
 * dw_w/dw_b/pw_w/pw_b: host arrays of 3 device pointers (edge, color, detail), reference layouts, fp32. */
int lun_fe_branches(const void* y0, const float* scale, const float* shift, const float* const* dw_w,
                    const float* const* dw_b, const float* const* pw_w, const float* const* pw_b, void* cat, int B,
                    int H, int W, float slope, void* stream);

/* ---- VAE kernels (LunarisCoreVAE, lunar_generate.py) --------------------------------------------------------- */

/* Per-image channel sums: stats[b][0][c] = sum_hw x, stats[b][1][c] = sum_hw x^2 (caller zeroes). Feeds GroupNorm. */
int lun_image_channel_stats_bf16(const void* x, float* stats, int B, int HW, int C, void* stream);

/* y = mish(group_norm(x));  if res: y = mish(y + res) (ResBlock tail, lunar_generate.py:53);  if add: y += add
 * (decoder skip connection, :212-222). Reference: nn.GroupNorm(8,C) + nn.Mish, lunar_generate.py:37-38,96-97,170-171. */
int lun_gn_mish_fwd_bf16(const void* x, const float* stats, const float* gamma, const float* beta, const void* res,
                         const void* add, void* y, int B, int HW, int C, int groups, float eps, void* stream);

/* Backward of the above (two launches: reduce + apply). dy2 (nullable) is added to dy (two consumers of the same
 * activation). red[b][0][c] = sum_hw dyh, red[b][1][c] = sum_hw dyh*xhat (caller zeroes) give dbeta / dgamma after a
 * sum over b. dx = gradient at the conv output; dres (when res != null) = gradient of the residual input. */
int lun_gn_mish_bwd_bf16(const void* dy, const void* dy2, const void* x, const float* stats, const float* gamma,
                         const float* beta, const void* res, float* red, void* dx, void* dres, int B, int HW, int C,
                         int groups, float eps, void* stream);

/* Encoder first layer: conv3x3 (3 -> cout=64, stride 1|2, pad 1) on NCHW fp32 images -> NHWC bf16
 * (lunar_generate.py:95), and its weight / bias gradient (dw [cout,3,3,3], db [cout], caller zeroes). */
int lun_conv3x3_c3_fwd(const float* x_nchw, const float* w, const float* bias, void* y, int B, int H, int W, int cout,
                       int stride, void* stream);
int lun_conv3x3_c3_wgrad(const void* dy, const float* x_nchw, float* dw, float* db, int B, int H, int W, int cout,
                         int stride, void* stream);

/* Decoder last layer: recon = tanh(conv3x3(x, 32 -> 3) + bias) as NCHW fp32 (lunar_generate.py:226-228) and its
 * backward: dx (NHWC bf16), dw [3,32,3,3], db [3] (caller zeroes dw, db). */
int lun_final_conv_tanh_fwd(const void* x, const float* w, const float* bias, float* recon, int B, int H, int W,
                            void* stream);

/* Decoder tail fused (lunar_generate.py:187-189 + 226-228): h = mish(groupnorm(t)) for the 32-channel up4 output,
 * recon = tanh(conv3x3(h) + bias) in one pass over t. t: [B,H,W,32] bf16 raw ConvTranspose output; stats: [B,2,32]
 * per-image channel sums (lun_image_channel_stats_bf16); h_out: optional [B,H,W,32] bf16 (kept for the backward, NULL
 * when sampling); recon: [B,3,H,W] fp32. H % 16 == 0, W % 32 == 0. */
int lun_gn_mish_final_conv_tanh_fwd(const void* t, const float* stats, const float* gamma, const float* beta,
                                    const float* w, const float* bias, void* h_out, float* recon, int B, int H, int W,
                                    int groups, float eps, void* stream);
int lun_final_conv_bwd(const float* drecon, const float* recon, const void* x, const float* w, void* dx, float* dw,
                       float* db, int B, int H, int W, void* stream);

/* z = mu + eps*exp(logvar/2) with mu|logvar packed as mulv[B,2L] (lunar_generate.py:259-261), and the backward that
 * merges dz with the direct gradients of mu / logvar (KL term) into dmulv[B,2L] bf16. */
int lun_reparam_fwd(const float* mulv, const float* eps, void* z, int B, int L, void* stream);
int lun_reparam_bwd(const float* mulv, const float* eps, const void* dz, const float* dmu, const float* dlogvar,
                    void* dmulv, int B, int L, void* stream);

/* Step losses in one pass (train_hybrid.py:859-862): sums[0] += sum (recon-x)^2 over n_img elements,
 * sums[1] += sum (1 + logvar - mu^2 - exp(logvar)) over B*L (caller zeroes sums; recon_loss = sums[0]/n_img,
 * kl_loss = -0.5*sums[1]/(B*L)). mulv = [mu | logvar] packed [B, 2L]. The backward turns the two upstream scalar
 * gradients g[0] (recon_loss), g[1] (kl_loss) (device memory) into d_recon and d_mulv. */
int lun_vae_loss_fwd(const float* recon, const float* images, const float* mulv, float* sums, long n_img, int B, int L,
                     void* stream);
int lun_vae_loss_bwd(const float* recon, const float* images, const float* mulv, const float* g, float* drecon,
                     float* dmulv, long n_img, int B, int L, void* stream);

/* Sprite loader: uint8 NHWC [B,H,W,3] (sprites_*.npy as written by the reference's generate.py:858-904) -> fp32 NCHW
 * in [-1,1], x/127.5 - 1 (PixelArtDataset.__getitem__, train_hybrid.py:181-182). */
int lun_sprites_u8_to_f32(const void* u8_nhwc, float* out_nchw, int B, int H, int W, void* stream);

/* SelfAttention2d (lunar_generate.py:56-78) as a flash-style tcgen05 kernel: y = gamma * softmax(q k^T) v + x over
 * all N = H*W positions of each image, no N x N matrix. qk: [B*N, 128] bf16 rows = [q | k], each zero-padded from
 * C/8 to 64 columns (the 1x1 query / key convs write it); v, x, y: [B, N, C] bf16; gamma: device scalar.
 * Optional (training): o_out [B, N, C] bf16 = softmax(q k^T) v and lse_out [B*N] fp32 = row logsumexp, kept for the
 * backward; pass NULL for inference. N % 128 == 0, C % 64 == 0, C <= 512. */
int lun_flash_attn2d_bf16(const void* qk, const void* v, const void* x, void* y, const float* gamma, int B, int N,
                          int C, void* o_out, float* lse_out, void* stream);

/* Backward of the same module (what autograd derives from lunar_generate.py:70-77), three steps:
 *   prep: dsum[row] = sum_c dy[row,c] * o[row,c]; dgamma[0] += sum of all dsum (dgamma zeroed by the caller).
 *   dv:   dv[B,N,C] bf16 = gamma * P^T dy              (P rebuilt from qk and lse)
 *   dqk:  dqk[B*N,128] bf16 = gamma * [dS K | dS^T Q], dS = P o (dy v^T - dsum); one launch, grid.z = {queries, keys}.
 * dy: [B, N, C] bf16 gradient of y (the residual branch dx += dy is the caller's). */
int lun_flash_attn2d_bwd_prep_bf16(const void* dy, const void* o, float* dsum, float* dgamma, long rows, int C,
                                   void* stream);
int lun_flash_attn2d_dv_bf16(const void* qk, const void* dy, const float* lse, const float* gamma, void* dv, int B,
                             int N, int C, void* stream);
int lun_flash_attn2d_dqk_bf16(const void* qk, const void* v, const void* dy, const float* lse, const float* dsum,
                              const float* gamma, void* dqk, int B, int N, int C, void* stream);

/* bf16 kernel-operand shadow of an fp32 weight (rebuilt after every optimizer step): src [A][B][T] fp32 - a conv weight
 * [Cout][Cin][kh*kw] or a ConvTranspose2d weight [Cin][Cout][kh*kw] - -> dst bf16 [T][A][B] (transpose = 0) or
 * [T][B][A] (transpose = 1), the K-major slabs lun_conv_taps_bf16 / lun_convT4x4s2_bf16 read. */
int lun_pack_weight_bf16(const float* src, void* dst, int A, int B, int T, int transpose, void* stream);

/* `heads` independent GEMMs in ONE launch of the tap-list kernel (grouped mode): out[r, h*cout + n] = bias[h*cout + n] +
 * sum_k x[r, h*cin + k] * w[h][n][k]; x [rows, heads*cin] bf16, w [heads][cout][cin] bf16, out [rows, heads*cout] bf16.
 * Replaces the value projection of the as-executed attention (lunar_evaluator.py:153-156, 213) on the surviving rows:
 * one [hd, C] GEMM per head instead of a block-diagonal [C, heads*C] one that multiplies 7/8 zeros. heads <= 16. */
int lun_head_linear_bf16(const void* x, int rows, int heads, int cin, const void* w, int cout, const float* bias,
                         void* out, void* stream);

/* Teacher heads (lunar_evaluator.py:353-397, 417-456): gate, quality heads, semantic head, style / prompt nets and the
 * mixing between them, fp32, on the pooled per-image channel SUMS of the trunk (the 1 / (H W) of AdaptiveAvgPool2d is
 * inv_hw). One launch forward, two launches backward.
 *   params / grads: 6 pointers per head {ln_w, ln_b, w1, b1, w2, b2} (torch layouts [out, in]; ln_* null for the gate),
 *     heads in the order gate, quality[0..E), semantic, style, prompt -> (E + 4) * 6 entries; a null grads entry means
 *     "not wanted";
 *   dims: {B, E, F, C, emb, gate_hidden, quality_hidden, semantic_hidden, style_hidden, prompt_hidden};
 *   seeds: E + 4 dropout seeds of the counter RNG (element index = sample * hidden + unit), used when drop_p > 0;
 *   pooled_fe [B, F], pooled [E][B][C] sums; outputs quality [B,4] (sigmoid), weights [B,E] (softmax), style / prompt
 *   [B, emb], semantic [B,1]; save [B, sizes[0]] keeps what the backward needs (null in eval);
 *   backward work buffers: dpre [B, sizes[1]], dout2 [B, sizes[2]], dxn [(E + 3), B, C] floats (lun_heads_buffer_sizes);
 *   g_*: incoming gradients of the five outputs (any may be null); d_pooled [E][B][C]: gradient of the pooled sums. */
int lun_heads_buffer_sizes(const int* dims, int* sizes);
int lun_heads_fwd(const void* const* params, const int* dims, const unsigned long long* seeds, float inv_hw, float slope,
                  float drop_p, int training, const float* pooled_fe, const float* pooled, float* quality,
                  float* weights, float* style, float* prompt, float* semantic, float* save, void* stream);
int lun_heads_bwd(const void* const* params, const int* dims, const unsigned long long* seeds, float inv_hw, float slope,
                  float drop_p, const float* pooled_fe, const float* pooled, const float* weights, const float* save,
                  const float* g_quality, const float* g_weights, const float* g_style, const float* g_prompt,
                  const float* g_semantic, float* dpre, float* dout2, float* dxn, float* d_pooled, void* const* grads,
                  void* stream);

/* Optimizer boundary (train_hybrid.py:906-922): clip_grad_norm_(max_norm) + AdamW over all tensors of one model in two
 * multi-tensor launches. table: device array of 72-byte rows {float* param, grad, exp_avg, exp_avg_sq; long long numel;
 * float lr, beta1, beta2, eps, weight_decay, bias_c1 = 1 - beta1^t, bias_c2_sqrt = sqrt(1 - beta2^t), pad} - hyper-
 * parameters and step count per tensor (param groups, late first gradients); chunks: device array of int2 {tensor index,
 * chunk index} with 8192 elements per chunk; norm2: device float[1024] of per-block partial sums of (g * grad_scale)^2,
 * summed in a fixed order by the update kernel (no host sync, bitwise reproducible). grad_scale folds the data-parallel
 * 1/world average into the pass. Semantics = torch.optim.AdamW (decoupled decay, eps outside the bias-corrected sqrt) on
 * the gradients scaled in place by grad_scale * min(1, max_norm / (norm + 1e-6)). */
int lun_multi_grad_sumsq(const void* table, const void* chunks, int nchunks, float grad_scale, float* norm2,
                         void* stream);
int lun_multi_clip_adamw(const void* table, const void* chunks, int nchunks, const float* norm2, float max_norm,
                         float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif
