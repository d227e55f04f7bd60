/*
 * floodplanet_b200.h -- C ABI of the B200 (sm_100a) UNet hot path.
 *
 * This is the drop-in boundary for what `st_water_seg`'s UNet training / inference step
 * dispatches to (reference: st_water_seg/models/unet.py, water_seg_model.py, ef_model.py).
 * The reference has no FFI of its own -- every op below is a torch ATen call there -- so each
 * entry point cites the reference line whose library dispatch it replaces.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless said
 *     otherwise; `stream` is a cudaStream_t passed as void*.
 *   - activations are NHWC bf16 "views": base pointer + pixel pitch `ld` (elements between
 *     consecutive pixels), so a view can be a channel slice of a wider (concat) buffer.
 *     Base pointers must be 16-byte aligned and `ld` a multiple of 8.
 *   - no allocation, no host synchronisation, no global state: the caller owns workspaces.
 *   - every function returns FPB200_OK (0) or a negative FPB200_ERR_* code; the Python host
 *     side turns a non-zero status into RuntimeError(kernel name + shape).  There is no CPU
 *     fallback anywhere behind this ABI.
 */
#ifndef FLOODPLANET_B200_H_
#define FLOODPLANET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPB200_OK 0
#define FPB200_ERR_SHAPE (-1)     /* unsupported / inconsistent dimensions            */
#define FPB200_ERR_ALIGN (-2)     /* pointer or pitch alignment violated              */
#define FPB200_ERR_LAUNCH (-3)    /* CUDA launch / runtime error (message on stderr)  */
#define FPB200_ERR_DRIVER (-4)    /* driver entry point unavailable (no GPU driver)   */
#define FPB200_ERR_TENSORMAP (-5) /* cuTensorMapEncodeTiled rejected the view         */
#define FPB200_ERR_NCCL (-6)      /* NCCL runtime missing, or an NCCL call failed     */

/* library identification: returns the ABI version (bumped on any signature change). */
int fpb200_abi_version(void);

/* ------------------------------------------------------------------------------------------
 * Layout / ingest
 * ---------------------------------------------------------------------------------------- */

/* NCHW fp32 image(s) -> NHWC bf16 with channels zero-padded to `c_pad` (multiple of 8).
 * Up to 8 source tensors are concatenated along C in the given order: this fuses
 * EarlyFusionModel.forward's torch.concat chain (models/ef_model.py:24-47) and the implicit
 * layout/cast step into one pass.  srcs / src_channels are HOST arrays of length n_src. */
int fpb200_ingest_nchw_f32_to_nhwc_bf16(const float* const* srcs, const int* src_channels,
                                        int n_src, void* dst, int c_pad, int N, int H, int W,
                                        void* stream);

/* Sliding-window inference ingest: `n_tiles` crops of ONE scene resident on the device
 * ([C][H][W] fp32) -> NHWC bf16 [n_tiles][th][tw][c_pad].  `tiles` is a DEVICE int32 array
 * [n_tiles][4] = (h0, w0, valid_h, valid_w) as produced by get_crop_slices(mode='exact')
 * (datasets/utils.py:86-212); pixels beyond the valid extent / scene edge are zero padded. */
int fpb200_ingest_scene_tiles(const float* scene, int C, long H, long W, const int* tiles,
                              int n_tiles, int th, int tw, void* dst, int c_pad, void* stream);

/* Conv weight repack, OIHW fp32 [Cout][Cin][3][3] ->
 *   fprop packing  bf16 [Cout][9][cin_pad]            (K-major GEMM B operand)
 *   dgrad packing  bf16 [cin_pad_out][9][Cout]  with the filter rotated by 180 degrees
 * Reference: nn.Conv2d parameters of DoubleConv (models/unet.py:14,16). */
int fpb200_repack_weights_fprop(const float* w_oihw, void* w_packed, int Cout, int Cin,
                                int cin_pad, void* stream);
int fpb200_repack_weights_dgrad(const float* w_oihw, void* w_packed, int Cout, int Cin,
                                void* stream);

/* Every 3x3 operand copy of a model in one launch (what an optimiser step invalidates).
 * table: device array of n_entries records, 40 bytes each, sorted by first_block:
 *   { const float* w_oihw; void* packed; int32 cout, cin, cin_pad, kind (0 fprop layout, 1 dgrad
 *     layout); int64 first_block }   -- entry e owns blocks [first_block[e], first_block[e+1]): one block
 *   per output channel (kind 0) or per input channel (kind 1); cin <= 1024; total_blocks = end of the
 *   last entry. */
int fpb200_repack_weights_batch(const void* table, int n_entries, long total_blocks, void* stream);

/* ------------------------------------------------------------------------------------------
 * 3x3 convolution (tcgen05 implicit GEMM)  -- models/unet.py:14,16 and their autograd
 * ---------------------------------------------------------------------------------------- */

/* rows of the BatchNorm partial-statistics workspace written by fprop: [rows][2][Cout] fp32 */
int fpb200_conv_stat_rows(void);

/* y = conv3x3(x, w) (no bias), bf16 NHWC in/out, fp32 accumulate.
 *   Cin  multiple of 16 (channel-padded input), Cout multiple of 64.
 *   scale/shift (nullable, [Cout]): epilogue y = acc*scale+shift, then ReLU if relu != 0
 *     (eval-mode BatchNorm folded with the conv bias: models/unet.py:15,17 in .eval()).
 *   stat_partials (nullable, [fpb200_conv_stat_rows()][2][Cout]): per-channel sum / sum of
 *     squares of the fp32 accumulators (training-mode BatchNorm statistics, fused). */
int fpb200_conv3x3_fprop_bf16_nhwc(const void* x, long ldx, const void* w_packed, void* y,
                                   long ldy, int N, int H, int W, int Cin, int Cout,
                                   const float* scale, const float* shift, int relu,
                                   float* stat_partials, void* stream);

/* dx = conv3x3_transpose(dy, w): Cout channels in, Cin (multiple of 64) channels out.
 * Optional fused BatchNorm-backward reduction (bn_y != NULL): when dx is the gradient w.r.t. the
 * activation a = relu(bn(y_prev)) of the preceding conv layer, pass that layer's raw output
 * bn_y (NHWC bf16 view, Cin channels) and its coefficients; the epilogue then also writes the
 * partial sums  sum g, sum g*xhat  (g = dx*[a>0]) to bn_partials
 * ([fpb200_conv_stat_rows()][2][Cin] fp32, the input of fpb200_bn_bwd_finalize), replacing
 * fpb200_bn_relu_bwd_reduce for that layer.  Requires Cin <= 512. */
int fpb200_conv3x3_dgrad_bf16_nhwc(const void* dy, long lddy, const void* w_packed_dgrad,
                                   void* dx, long lddx, int N, int H, int W, int Cout, int Cin,
                                   const void* bn_y, long ld_bn_y, const float* bn_scale,
                                   const float* bn_shift, const float* bn_mean,
                                   const float* bn_invstd, float* bn_partials, void* stream);

/* dW = sum_pixels dy (x) shifted x.  Two stages, deterministic:
 *   stage 1 (tensor cores) writes split-K partials into `workspace`
 *           (fp32, fpb200_conv3x3_wgrad_workspace_bytes(...) bytes),
 *   stage 2 reduces them into dw_oihw fp32 [Cout][Cin_real][3][3] (overwrites). */
long fpb200_conv3x3_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout);
int fpb200_conv3x3_wgrad_bf16_nhwc(const void* x, long ldx, const void* dy, long lddy,
                                   float* dw_oihw, void* workspace, int N, int H, int W,
                                   int Cin, int cin_real, int Cout, void* stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm2d + ReLU (+ MaxPool2d(2))  -- models/unet.py:15,17,29 and their autograd
 * ---------------------------------------------------------------------------------------- */

/* Reduce the fprop partials to batch statistics and fold them:
 *   mean/var(biased) over `count` = N*H*W elements; invstd = rsqrt(var+eps)
 *   scale = gamma*invstd, shift = beta - mean*scale           (apply pass coefficients)
 *   save_mean/save_invstd kept for backward
 *   running_mean = (1-m)*running_mean + m*(mean + conv_bias)  (conv bias is not in the GEMM;
 *   running_var  = (1-m)*running_var  + m*var*count/(count-1)  it only shifts the mean)
 * `num_partials` rows of [2][C] fp32. */
int fpb200_bn_stats_finalize(const float* partials, int num_partials, int C, double count,
                             const float* gamma, const float* beta, const float* conv_bias,
                             float eps, float momentum, float* running_mean, float* running_var,
                             float* scale, float* shift, float* save_mean, float* save_invstd,
                             void* stream);

/* eval-mode fold: scale = gamma/sqrt(running_var+eps), shift = beta + (bias-running_mean)*scale */
int fpb200_bn_fold_eval(const float* gamma, const float* beta, const float* conv_bias,
                        const float* running_mean, const float* running_var, float eps, int C,
                        float* scale, float* shift, void* stream);

/* Backward THROUGH an eval-mode BatchNorm (frozen-BatchNorm fine-tuning, saliency: the reference's nn.BatchNorm2d
 * supports autograd in .eval(), models/unet.py:15,17).  The running statistics are constants, so with
 * xhat = (y + conv_bias - running_mean) * invstd:
 *   fpb200_bn_eval_stats          mean_eff = running_mean - conv_bias, invstd = 1/sqrt(running_var + eps)
 *                                 (the form fpb200_bn_relu_bwd_reduce and the fused reductions take)
 *   fpb200_bn_bwd_finalize_frozen dgamma = sum g*xhat, dbeta = sum g, dbias = scale * sum g (the conv bias is no
 *                                 longer cancelled by a batch mean), coef = 0 so that fpb200_bn_relu_bwd_apply
 *                                 produces dy = scale * g.  partials as for fpb200_bn_bwd_finalize. */
int fpb200_bn_eval_stats(const float* conv_bias, const float* running_mean, const float* running_var,
                         float eps, int C, float* mean_eff, float* invstd, void* stream);
int fpb200_bn_bwd_finalize_frozen(const float* partials, int num_partials, int C, const float* scale,
                                  float* dgamma, float* dbeta, float* dbias, float* coef, void* stream);

/* a = relu(y*scale+shift) elementwise over an NHWC bf16 view. */
int fpb200_bn_apply_relu(const void* y, long ldy, void* a, long lda, const float* scale,
                         const float* shift, long num_pixels, int C, void* stream);

/* Same, fused with MaxPool2d(2) (floor mode): writes the full-resolution activation `a`
 * (nullable: pure pooling of an already activated input when scale == NULL), the pooled map
 * `pooled` [N][H/2][W/2][C] and the window argmax (0..3, row-major in the 2x2 window, first
 * maximum wins, NaN propagates -- torch's max_pool2d scan order) as one byte per element. */
int fpb200_bn_apply_relu_maxpool2(const void* y, long ldy, void* a, long lda, void* pooled,
                                  long ldp, uint8_t* pool_idx, const float* scale,
                                  const float* shift, int N, int H, int W, int C, void* stream);

/* MaxPool2d(2) backward fused with the skip-connection gradient add:
 *   dx[n,h,w,c] = dskip[n,h,w,c] (nullable) + (pool_idx selects (h,w)) ? dpooled : 0
 * Optional (bn_y != NULL): dx is the activation gradient of the conv+BN+ReLU layer whose raw
 * output is bn_y; that layer's BatchNorm-backward partial sums (as fpb200_bn_relu_bwd_reduce
 * would produce) are written to bn_partials, fp32 [2*fpb200_bn_bwd_rows()][2][C]. */
int fpb200_maxpool2_bwd(const void* dpooled, long lddp, const uint8_t* pool_idx, const void* dskip,
                        long ldds, void* dx, long lddx, int N, int H, int W, int C, const void* bn_y,
                        long ld_bn_y, const float* bn_scale, const float* bn_shift,
                        const float* bn_mean, const float* bn_invstd, float* bn_partials,
                        void* stream);

/* BatchNorm+ReLU backward, pass 1: with g = da * (y*scale+shift > 0) and
 * xhat = (y-mean)*invstd, writes per-block partial sums of g and g*xhat:
 * partials [fpb200_bn_bwd_rows()][2][C] fp32. */
int fpb200_bn_bwd_rows(void);
int fpb200_bn_relu_bwd_reduce(const void* da, long ldda, const void* y, long ldy,
                              const float* scale, const float* shift, const float* save_mean,
                              const float* save_invstd, float* partials, long num_pixels, int C,
                              void* stream);
/* pass 1b: reduce partials -> dgamma = sum(g*xhat), dbeta = sum(g) (fp32 [C], nullable) and
 * the folded per-channel coefficients of pass 2, coef[2][C] = (P, Q):
 *   dy = scale*(g - c1 - xhat*c2) = scale*g - P*y - Q,  c1 = sum(g)/count, c2 = sum(g*xhat)/count,
 *   P = scale*c2*invstd, Q = scale*c1 - P*mean. */
int fpb200_bn_bwd_finalize(const float* partials, int num_partials, int C, double count,
                           const float* scale, const float* save_mean, const float* save_invstd,
                           float* dgamma, float* dbeta, float* coef, void* stream);
/* pass 2: dy = scale*g - P*y - Q with g = da * (y*scale+shift > 0), bf16 NHWC. */
int fpb200_bn_relu_bwd_apply(const void* da, long ldda, const void* y, long ldy, void* dy,
                             long lddy, const float* scale, const float* shift, const float* coef,
                             long num_pixels, int C, void* stream);

/* ------------------------------------------------------------------------------------------
 * Up: bilinear x2 (align_corners=True) + zero pad + concat  -- models/unet.py:43-45,54-66
 * ---------------------------------------------------------------------------------------- */

/* Writes up(x) zero-padded to (Ho, Wo) (pad_top = (Ho-2h)/2, pad_left = (Wo-2w)/2) into the
 * NHWC view `out` (typically the second half of the concat buffer whose first half already
 * holds the skip tensor: torch.cat([x2, x1], dim=1)). */
int fpb200_upsample2x_pad_concat_fwd(const void* x, long ldx, void* out, long ldo, int N, int h,
                                     int w, int Ho, int Wo, int C, void* stream);
/* Gradient of the above w.r.t. x (gather form, deterministic): dx [N][h][w][C]. */
int fpb200_upsample2x_pad_concat_bwd(const void* dout, long lddo, void* dx, long lddx, int N,
                                     int h, int w, int Ho, int Wo, int C, void* stream);

/* ------------------------------------------------------------------------------------------
 * OutConv 1x1 head  -- models/unet.py:70-77
 * ---------------------------------------------------------------------------------------- */

/* logits[n,k,h,w] (fp32 NCHW) = sum_c a[n,h,w,c]*w[k,c] + b[k];  C = 64, n_classes <= 8.
 * bn_scale/bn_shift NULL: x is the activation a.  Non-NULL (training): x is the RAW output y of
 * the last conv3x3 and a = relu(y*scale+shift) (rounded to bf16) is formed on load, so the
 * separate BatchNorm-apply pass of that layer (4 B/element) never runs. */
int fpb200_head1x1_fwd(const void* x, long ldx, const float* w, const float* b, float* logits,
                       int N, int H, int W, int C, int n_classes, const float* bn_scale,
                       const float* bn_shift, void* stream);
/* dx (bf16 NHWC, gradient w.r.t. the activation a), and per-block partials of dW [n_classes][C]
 * and db [n_classes] in `partials` (fp32 [fpb200_head_bwd_rows()][n_classes*(C+1)]), reduced
 * into dw/db.  With the bn_* pointers non-NULL x is again the raw conv output: a is recomputed
 * on load and the BatchNorm-backward sums of that layer (sum g, sum g*xhat, g = dx*[a>0]) are
 * written to bn_partials (fp32 [fpb200_head_bwd_rows()][2][C]) -- the input of
 * fpb200_bn_bwd_finalize -- so fpb200_bn_relu_bwd_reduce is not needed for that layer. */
int fpb200_head_bwd_rows(void);
int fpb200_head1x1_bwd(const float* dlogits, const void* x, long ldx, const float* w, void* dx,
                       long lddx, float* dw, float* db, float* partials, int N, int H, int W,
                       int C, int n_classes, const float* bn_scale, const float* bn_shift,
                       const float* bn_mean, const float* bn_invstd, float* bn_partials,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * Masked cross-entropy + argmax (+ confusion counts)
 *   -- models/water_seg_model.py:40,103-107 (nn.CrossEntropyLoss(ignore_index), argmax(dim=1))
 * ---------------------------------------------------------------------------------------- */

/* result (fp64 [4]):  [0] sum of -log softmax(logits)[t] over non-ignored pixels
 *                     [1] number of non-ignored pixels   [2] number of invalid targets
 *                     [3] mean loss = [0]/[1]  (NaN when [1] == 0, exactly like torch)
 * pred (nullable, int64 [N][H][W]): argmax over classes, first maximum wins, NaN is maximal.
 * confusion (nullable, int64 [n_classes][n_classes], row = target, col = pred; ignored pixels
 * excluded) is ACCUMULATED into.  `partials` fp64 [fpb200_ce_rows()][4] workspace. */
int fpb200_ce_rows(void);
int fpb200_softmax_ce_argmax_fwd(const float* logits, const int64_t* target, long ignore_index,
                                 double* result, int64_t* pred, int64_t* confusion,
                                 double* partials, int N, int n_classes, long hw, void* stream);
/* dlogits = (softmax - onehot) * (t != ignore) * grad_out / count; count = result[1].
 * All-ignored batch (count == 0) gives zeros, matching nan_to_num'd reference behaviour. */
int fpb200_softmax_ce_bwd(const float* logits, const int64_t* target, long ignore_index,
                          const double* result, const float* grad_out, float* dlogits, int N,
                          int n_classes, long hw, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sliding-window inference post-processing  -- infer.py:122-184, utils/utils_image.py:410-494
 * ---------------------------------------------------------------------------------------- */

/* canvas[h0+y, w0+x, :] += softmax(logits[t, :, y, x]); weight[h0+y, w0+x] += 1 for every valid
 * pixel of every tile (logits fp32 [n_tiles][C][th][tw], canvas fp32 [H][W][C], weight fp32
 * [H][W], tiles as in fpb200_ingest_scene_tiles). */
int fpb200_softmax_stitch_add(const float* logits, float* canvas, float* weight, const int* tiles,
                              int n_tiles, int n_classes, int th, int tw, long H, long W,
                              void* stream);
/* mask = clip(argmax_c(nan_to_num(canvas / (weight + 1e-5))), 0, 1) * 255  (uint8 [H][W]). */
int fpb200_canvas_to_mask_u8(const float* canvas, const float* weight, uint8_t* mask, long npix,
                             int n_classes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Feature-level seams: encoder/decoder API and late fusion
 *   models/unet.py:113-131 (UNet.encode / UNet.decode), :134-191 (UNetEncoder / UNetDecoder),
 *   models/lf_model.py:29-92 (per-modality encoders, torch.concat of the five feature levels,
 *   concat_convs = nn.Conv2d(fs*k, fs, 1, 1), decoder) and their autograd
 * ---------------------------------------------------------------------------------------- */

/* Feature-map layout/cast at the encode()/decode() boundary (callers exchange fp32 NCHW lists):
 *   fp32 [N][C][H][W]  ->  bf16 NHWC view (pitch ld, channels [0,C)), and back.  C % 8 == 0. */
int fpb200_nchw_f32_to_nhwc_bf16(const float* src, void* dst, long ld, int N, int C, int H, int W,
                                 void* stream);
int fpb200_nhwc_bf16_to_nchw_f32(const void* src, long ld, float* dst, int N, int C, int H, int W,
                                 void* stream);

/* 1x1 conv weight repack: fp32 [Cout][Cin] -> bf16 [Cout][Cin] (transpose == 0, forward operand)
 * or bf16 [Cin][Cout] (transpose != 0, data-gradient operand).  lf_model.py:44-45. */
int fpb200_repack_weights_1x1(const float* w_oi, void* w_packed, int Cout, int Cin, int transpose,
                              void* stream);

/* y = conv1x1(x, w) on the tensor cores (same TMA/tcgen05 pipeline as the 3x3 kernel with a
 * halo-free box and one tap), bf16 NHWC views, fp32 accumulate.  Cin % 64 == 0, Cout % 64 == 0.
 *   scale/shift (both or neither, [Cout], Cout <= 512): y = acc*scale+shift (+ReLU); the
 *   late-fusion forward passes scale = 1, shift = bias (lf_model.py:88 `concat_conv(img_feat)`).
 * The data gradient is the same call with the transposed packing and Cin/Cout swapped. */
int fpb200_conv1x1_bf16_nhwc(const void* x, long ldx, const void* w_packed, void* y, long ldy, int N,
                             int H, int W, int Cin, int Cout, const float* scale, const float* shift,
                             int relu, void* stream);

/* dW[co][ci] = sum_pixels dy[p][co] * x[p][ci]  (split-K on the tensor cores + deterministic
 * reduction, as for the 3x3 weight gradient); dw_oi fp32 [Cout][Cin] is overwritten. */
long fpb200_conv1x1_wgrad_workspace_bytes(int N, int H, int W, int Cin, int Cout);
int fpb200_conv1x1_wgrad_bf16_nhwc(const void* x, long ldx, const void* dy, long lddy, float* dw_oi,
                                   void* workspace, int N, int H, int W, int Cin, int Cout,
                                   void* stream);

/* out[c] = sum_pixels x[p][c]: bias gradient of a pointwise conv.  partials: fp32
 * [fpb200_channel_sum_rows()][C] workspace.  C/8 must divide 256 (C = 64 ... 512 here). */
int fpb200_channel_sum_rows(void);
int fpb200_channel_sum_bf16_nhwc(const void* x, long ld, float* partials, float* out,
                                 long num_pixels, int C, void* stream);

/* ------------------------------------------------------------------------------------------
 * Per-sample normalise + augment on the device (datasets/base_dataset.py:77-113 `normalize`,
 * :494-555 `sample_transforms` / `apply_transforms`; call order datasets/floodplanet.py:616-640)
 * ---------------------------------------------------------------------------------------- */

/* One gather pass: out[n,:,y,x] = normalise(img[n,:,src(y,x)]) with src = hflip o vflip o
 * rotate(nearest, centre, fill 0) per sample, exactly torchvision's CPU index arithmetic.
 *   img      fp32 NCHW [N,C,H,W]
 *   out_f32  fp32 NCHW (the tensor the reference's DataLoader yields) and / or
 *   out_nhwc_bf16  NHWC bf16 [N,H,W,c_pad] (channels >= C zero: the first conv's operand); either may be NULL
 *   tgt / tgt_out  int64 [N,H,W] annotation, same geometric chain, fill 0 (both or neither)
 *   theta    fp32 [N][6]: rows of theta^T / [W/2, H/2] = (x->gx, x->gy, y->gx, y->gy, 1->gx, 1->gy)
 *   flags    int32 [N]: bit0 hflip, bit1 vflip, bit2 rotate
 *   xgrid [W], ygrid [H]  fp32 base grid = linspace(-size/2 + .5, size/2 - .5, size)
 *   mean / stdv  float64 [N][C] or both NULL: numpy `image -= mean; image /= std` on the fp32 image */
int fpb200_augment_nchw_f32(const float* img, float* out_f32, void* out_nhwc_bf16, int c_pad,
                            const int64_t* tgt, int64_t* tgt_out, const float* theta,
                            const int* flags, const float* xgrid, const float* ygrid,
                            const double* mean, const double* stdv, int N, int C, int H, int W,
                            void* stream);

/* norm_mode 'local' (base_dataset.py:98-104): mean and population std of each of the `planes`
 * (= N*C) fp32 planes of hw pixels, rounded to fp32 like numpy's, stored as float64. */
int fpb200_plane_mean_std_f32(const float* img, double* mean, double* stdv, int planes, long hw,
                              void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser (water_seg_model.py:198-205, optim.Adam defaults) and misc
 * ---------------------------------------------------------------------------------------- */

/* One Adam step over a flat fp32 parameter slab (p, g, m, v all length n):
 * torch.optim.Adam semantics (no amsgrad, no weight decay), grad pre-scaled by grad_scale. */
int fpb200_adam_step(float* p, const float* g, float* m, float* v, long n, float lr, float beta1,
                     float beta2, float eps, int step, float grad_scale, void* stream);

/* Same update with the step counter resident on the device (step_state: int32[2] = {completed
 * steps, internal ticket}, zero-initialised by the caller); the kernel derives the bias
 * corrections from it and increments it, so the launch can be replayed from a CUDA graph. */
int fpb200_adam_step_graphable(float* p, const float* g, float* m, float* v, long n, float lr,
                               float beta1, float beta2, float eps, int* step_state,
                               float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data-parallel exchange step: gradient all-reduce over an ncclComm_t.
 * NEW capability -- the reference trains on one GPU (fit.py:86-88 `devices=1`); the call these
 * replace is what torch-DDP would add around `loss.backward()` (water_seg_model.py:98-136).
 * NCCL is resolved at RUN time from the libnccl.so.2 already loaded in the process (the one
 * torch bundles) or else from the system library; the library has no link-time dependency on it
 * and loads on a box without NCCL (fpb200_nccl_version() then returns 0).
 * ---------------------------------------------------------------------------------------- */
#define FPB200_NCCL_UNIQUE_ID_BYTES 128

/* NCCL_VERSION_CODE of the runtime that was found (e.g. 22809), 0 if none. */
int fpb200_nccl_version(void);
/* Rank 0: a fresh ncclUniqueId (128 bytes) to hand to every rank out of band (torch.distributed
 * store, MPI, a file ...). */
int fpb200_nccl_unique_id(void* id128);
/* Every rank: *comm = ncclCommInitRankConfig(world, id, rank) on the CURRENT device.  max_ctas > 0
 * caps the CTAs (= SMs) one collective may occupy (ncclConfig_t.maxCTAs): the all-reduce overlaps
 * persistent one-CTA-per-SM convolution kernels, every SM it takes delays a conv wave; 0 = NCCL default. */
int fpb200_nccl_comm_create(void** comm, int world, int rank, const void* id128, int max_ctas);
int fpb200_nccl_comm_destroy(void* comm);
/* In-place all-reduce of `count` fp32 values at `buf` on `stream`: average (op_avg = 1, ncclAvg: the
 * 1/world scaling happens inside the collective) or sum (op_avg = 0).  Asynchronous w.r.t. the host. */
int fpb200_allreduce_f32(void* comm, float* buf, long count, int op_avg, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FLOODPLANET_B200_H_ */
