/*
 * libdlv3p — C-ABI of the B200-native (sm_100a) DeepLabV3+ encoder/decoder hot path.
 *
 * The reference (tonandr/deeplabv3plus_keras) has NO FFI/plugin interface: its hot path is the list of
 * tf.keras layer constructors called in bodhi/deeplabv3plus_keras/semantic_segmentation.py
 * (_make_encoder :790-876, _make_decoder :878-913, _refine_boundary :915-954, class_balanced_loss :438-447)
 * whose arithmetic runs inside TensorFlow 2.4 raw ops.  Each entry point below replaces one of those raw ops
 * (named in the comment above it) and is what a `tf.load_op_library` custom-op wrapper (tf_ops/, see
 * INTEGRATION.md) or the ctypes host layer (deeplabv3plus_keras_b200/_lib.py) binds.
 *
 * Conventions
 *   - tensors are NHWC, dense unless an explicit leading dimension (`ld*`, in ELEMENTS) is given;
 *   - `dtype` is the storage type of activation tensors: DLV3P_F32 or DLV3P_BF16; math is fp32;
 *   - parameters (kernels of depthwise convs, BN vectors, loss weights) and all gradients of parameters are fp32;
 *   - every pointer is a DEVICE pointer owned by the caller; the library never allocates, frees or retains
 *     device memory; every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 on success, a negative DLV3P_ERR_* otherwise; dlv3p_last_error() gives the message
 *     (thread-local).  There is no CPU fallback: unsupported arguments are an error.
 *   - "addend" arguments (nullable) are added to the result before it is stored; they may alias the output,
 *     which is how gradient accumulation for fan-out tensors is expressed.
 */
#ifndef DLV3P_H_
#define DLV3P_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DLV3P_F32 0
#define DLV3P_BF16 1

#define DLV3P_ACT_NONE 0
#define DLV3P_ACT_RELU 1
#define DLV3P_ACT_RELU6 2

#define DLV3P_OK 0
#define DLV3P_ERR_SHAPE (-1)
#define DLV3P_ERR_DTYPE (-2)
#define DLV3P_ERR_ALIGN (-3)
#define DLV3P_ERR_CUDA (-4)
#define DLV3P_ERR_UNSUPPORTED (-5)

const char* dlv3p_last_error(void);
int dlv3p_version(void);
/* compute capability major*10+minor of the current device, or a negative error */
int dlv3p_device_arch(void);
/* Programmatic dependent launch (the next kernel of the stream is made resident while the current one drains; every
 * kernel waits on griddepcontrol before touching memory) on (default) / off for all subsequent launches.  A launch
 * ATTRIBUTE only — results are identical; profilers switch it off so that a kernel's recorded duration does not include
 * the time it spent resident behind its predecessor.  Returns the previous setting. */
int dlv3p_set_pdl(int enabled);

/* ------------------------------------------------------------------------------------------------
 * K1 — atrous depthwise 3x3 (TF DepthwiseConv2dNative / ...BackpropInput / ...BackpropFilter; the depthwise
 * half of SeparableConv2D at ss.py:823-830 and of every keras.applications Xception/MobileNetV2 block).
 * x [N,H,W,C]; w [3,3,C] fp32 (= Keras depthwise_kernel [3,3,C,1]); y [N,Ho,Wo,C].  C % 8 == 0.
 * Zero padding pad_t/pad_l before the first row/col (TF SAME: total//2 before, rest after; VALID: 0).
 * Optional prologue on x: v = act(in_scale[c]*x + in_shift[c]) (in_scale/in_shift nullable => identity affine),
 * evaluated BEFORE zero padding — i.e. the fused BatchNormalization+Activation that feeds the conv.
 * ---------------------------------------------------------------------------------------------- */
int dlv3p_dwconv3x3_fwd(const void* x, const float* w, void* y, int N, int H, int W, int C, int stride, int dil_h,
                        int dil_w, int pad_t, int pad_l, int Ho, int Wo, const float* in_scale,
                        const float* in_shift, int in_act, int dtype, void* stream);
/* dx = mask * conv_transpose(dy) (+ addend), mask = act'(in_scale*x_pre+in_shift) when in_act != NONE.
 * The result is the gradient w.r.t. the activation INPUT z = in_scale*x_pre+in_shift (not w.r.t. x_pre). */
/* Inference form of DepthwiseConv2D -> BatchNormalization -> ReLU/ReLU6 (every MobileNetV2 block, keras.applications
 * mobilenet_v2 `*_depthwise`, `*_depthwise_BN`, `*_depthwise_relu`): y = act(out_scale[c] * conv(x)[c] + out_shift[c]) with
 * the folded BN applied in the convolution's epilogue (TF runs DepthwiseConv2dNative, FusedBatchNormV3 and Relu6 as three
 * kernels over the tensor).  No input prologue; other arguments as dlv3p_dwconv3x3_fwd. */
int dlv3p_dwconv3x3_fwd_epi(const void* x, const float* w, void* y, int N, int H, int W, int C, int stride, int dil_h,
                            int dil_w, int pad_t, int pad_l, int Ho, int Wo, const float* out_scale,
                            const float* out_shift, int out_act, int dtype, void* stream);
int dlv3p_dwconv3x3_dgrad(const void* dy, const float* w, void* dx, int N, int H, int W, int C, int stride,
                          int dil_h, int dil_w, int pad_t, int pad_l, int Ho, int Wo, const void* x_pre,
                          const float* in_scale, const float* in_shift, int in_act, const void* addend, int dtype,
                          void* stream);
/* dlv3p_dwconv3x3_fwd (stride 1, dilation 1, bf16) whose input is act(BatchNormalization(x)) of the producing layer in
 * TRAINING mode, with dlv3p_bn_finalize folded in: every CTA derives scale/shift of its channels from the batch sums
 * (bn_sums[0..C) = sum x, [C..2C) = sum x^2 over `count` samples, as written by dlv3p_gemm_bf16's col_stats); the first
 * CTA of each channel block publishes scale/shift/mean/invstd (read by the backward kernels) and applies `updates`
 * momentum updates to the moving statistics (tf.raw_ops.FusedBatchNormV3 conventions, see dlv3p_bn_finalize). */
int dlv3p_dwconv3x3_bn_fwd(const void* x, const float* w, void* y, int N, int H, int W, int C, int pad_t, int pad_l,
                           int Ho, int Wo, const float* bn_sums, const float* gamma, const float* beta,
                           float* moving_mean, float* moving_var, double count, float eps, float momentum, int updates,
                           int in_act, float* scale, float* shift, float* mean, float* invstd, int dtype, void* stream);
/* dlv3p_dwconv3x3_dgrad (stride 1, dilation 1, bf16, no addend) that ALSO produces the BatchNormalization-backward
 * reductions of the layer whose raw conv output is x_pre (tf.raw_ops.FusedBatchNormGradV3's two sums, i.e. what
 * dlv3p_bn_bwd_reduce computes with act = NONE from the dx written here):
 *   bn_red[0..C) += sum dx,  bn_red[C..2C) += sum dx * (x_pre - bn_mean) * bn_invstd.
 * The engine uses it when Conv -> BatchNormalization -> ReLU feeds a SeparableConv2D (Xception blocks,
 * keras.applications.xception; ss.py:823-830): the BN+ReLU output is never written, the consumer applies it on load
 * (in_scale/in_shift/in_act of dlv3p_dwconv3x3_fwd / _wgrad) and this call hands the masked gradient and its
 * reductions straight to dlv3p_bn_bwd_apply. */
int dlv3p_dwconv3x3_dgrad_bnred(const void* dy, const float* w, void* dx, int N, int H, int W, int C, int pad_t,
                                int pad_l, int Ho, int Wo, const void* x_pre, const float* in_scale,
                                const float* in_shift, int in_act, const float* bn_mean, const float* bn_invstd,
                                float* bn_red, int dtype, void* stream);
/* dw[3,3,C] (fp32) += sum over N,Ho,Wo of act(in_scale*x+in_shift)[tap] * dy */
int dlv3p_dwconv3x3_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int C, int stride,
                          int dil_h, int dil_w, int pad_t, int pad_l, int Ho, int Wo, const float* in_scale,
                          const float* in_shift, int in_act, int dtype, void* stream);
/* Whole backward of a stride-1, dilation-1, SAME depthwise convolution in ONE pass over dy (the depthwise half of every
 * Xception SeparableConv2D, keras.applications.xception / ss.py:823-830; TF runs DepthwiseConv2dNativeBackpropInput,
 * DepthwiseConv2dNativeBackpropFilter, ReluGrad and — for the producing layer — FusedBatchNormGradV3's reductions as
 * separate kernels): equivalent to dlv3p_dwconv3x3_dgrad (or _dgrad_bnred when bn_red != NULL) followed by
 * dlv3p_dwconv3x3_wgrad on the same operands.
 *   x: what the forward convolution read BEFORE its fused prologue, i.e. conv input = act(in_scale*x + in_shift);
 *   dx = act'(in_scale*x + in_shift) * conv_transpose(dy) (+ addend);  dw[3,3,C] (fp32) += sum act(..)[tap] * dy;
 *   bn_red (optional): BatchNormalization-backward reductions (dlv3p_bn_bwd_reduce's contract, accumulated):
 *     bn_y == NULL (needs in_scale/in_shift, no addend): of the layer whose raw output is x —
 *       [0..C) += sum g, [C..2C) += sum g*(x-bn_mean)*bn_invstd with g = the masked gradient (dx before any addend);
 *     bn_y != NULL (in_act = RELU, no in_scale): of the layer whose raw output is bn_y and whose BN output (+ residual)
 *       is x (the block-closing SeparableConv2D of an Xception block read through the next block's pre-activation) —
 *       [0..C) += sum dx, [C..2C) += sum dx*(bn_y-bn_mean)*bn_invstd with dx the FINAL gradient (addend included).
 * Written around the input pixel both gradients need the same 3x3 window of dy, so the bf16 kernel stages dy once
 * (TMA halo box) and keeps the filter-gradient partial sums in registers next to the taps. */
int dlv3p_dwconv3x3_bwd(const void* dy, const void* x, const float* w, void* dx, float* dw, int N, int H, int W, int C,
                        const float* in_scale, const float* in_shift, int in_act, const void* addend,
                        const float* bn_mean, const float* bn_invstd, float* bn_red, const void* bn_y, int dtype,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2 — pointwise / projection convolutions as bf16 tcgen05 tensor-core GEMMs (TF Conv2D 1x1, the pointwise
 * half of SeparableConv2D, and — after dlv3p_im2col3x3 — the dense 3x3 convs; ss.py:814-818,833-838,
 * 843-847,865-869,893-897,931-935).
 *   C[M,N] = epilogue( A[M,K] * B[N,K]^T ),  A,B bf16 row-major (K contiguous), fp32 accumulation in TMEM.
 *   epilogue: v = acc; if col_scale: v = v*col_scale[n] + col_shift[n]; v = act(v); if addend: v += addend[m,n];
 *   C stored as c_dtype.  If col_stats != NULL, col_stats[0..N) += sum_m acc[m,n] and col_stats[N..2N) +=
 *   sum_m acc[m,n]^2 (raw accumulator, i.e. the BatchNormalization batch statistics of the conv output).
 *   lda/ldb multiples of 8 elements; A,B 16-byte aligned.  ldc arbitrary (channel-slice writes into a concat).
 * ---------------------------------------------------------------------------------------------- */
int dlv3p_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int M, int N,
                    int K, int c_dtype, const float* col_scale, const float* col_shift, int act,
                    const void* addend, int64_t ld_addend, float* col_stats, void* stream);
/* filter gradient: dW[K,N] (fp32, row stride ldw) += X[M,K]^T * dY[M,N]   (split over M, fp32 atomics) */
int dlv3p_gemm_wgrad_bf16(const void* X, int64_t ldx, const void* dY, int64_t ldy, float* dW, int64_t ldw, int M,
                          int K, int N, void* stream);

/* Implicit-GEMM 3x3 VALID stride-1 convolution, bf16 NHWC, no im2col matrix (Xception block1_conv2 =
 * keras.applications.xception Conv2D(64, (3,3), use_bias=False), reached through ss.py:512-515; TF runs it as one
 * cuDNN Conv2D + Conv2DBackpropInput + Conv2DBackpropFilter).  Ho = H-2, Wo = W-2.
 *   fwd:   y[N,Ho,Wo,Cout] = epilogue(conv(x[N,H,W,Cin], W)); epilogue and col_stats as dlv3p_gemm_bf16.
 *          wt = bf16 [Cout, 9*Cin] K-major with row pitch ldw, wt[o, (i*3+j)*Cin + c] = W[i,j,c,o] (the im2col GEMM's B).
 *          64 <= Cout <= 256.  Cin == 32, Cout == 64 takes the halo-staged kernel (three input rows per stage, nine taps
 *          as shifted UMMA views of it); other shapes read overlapping 3*Cin-element runs through a rank-3 tensor map.
 *   dgrad: dx[N,H,W,Cin] = conv_transpose(dy[N,Ho,Wo,Cout], W); wd = bf16 [Cin, 9*Cout], wd[c,(i*3+j)*Cout+o] = W[i,j,c,o].
 *          One stage holds the three dy rows a 128-pixel tile needs (130-pixel boxes); the nine taps are nine views of
 *          it (UMMA descriptor base_offset), the filter slices stay resident in shared memory.  Cout == 64, Cin <= 32.
 *   wgrad: dw (fp32 HWIO [3,3,Cin,Cout]) += sum_pixels x-window * dy.  64 < 3*Cin <= 128.
 * Cin, Cout multiples of 8; all pointers 16-byte aligned; anything else returns DLV3P_ERR_UNSUPPORTED and the caller
 * uses dlv3p_im2col3x3 + dlv3p_gemm_bf16. */
int dlv3p_conv3x3_valid_fwd_bf16(const void* x, const void* wt, int64_t ldw, void* y, int N, int H, int W, int Cin,
                                 int Cout, const float* col_scale, const float* col_shift, int act, float* col_stats,
                                 void* stream);
int dlv3p_conv3x3_valid_dgrad_bf16(const void* dy, const void* wd, void* dx, int N, int H, int W, int Cin, int Cout,
                                   void* stream);
int dlv3p_conv3x3_valid_wgrad_bf16(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin, int Cout,
                                   void* stream);
/* 3x3 SAME stride-1 convolution (Conv2D(num_classes, kernel_size=3, padding='same', use_bias=False), the logits layer of
 * the decoder, ss.py:893-897; after _refine_boundary its input is the 304-channel concat at 256 x 256, ss.py:915-954) as
 * implicit GEMMs: every tap is a rank-4 TMA window (64 channels, 128 pixels of one image row, row, image) shifted by the
 * tap, what falls outside the image is zero-filled by TMA — that IS the SAME padding; no [pixels, 9*Cin] column matrix.
 *   fwd:   y[N,H,W,Cout] (c_dtype bf16 or fp32; Cout <= 256, any count: 21 classes run as one 32-column tile) =
 *          epilogue(conv(x, W)); wt = bf16 [Cout, 9*Cin] K-major, pitch ldw, as for the VALID convolution.
 *   dgrad: dx[N,H,W,Cin] (bf16) = sum_taps dy(shifted) * W^T.  dy = bf16 [N,H,W,ld_dy] (ld_dy >= Cout, multiple of 8);
 *          wd = bf16 [Cin, 9*kp], kp = 64*ceil(Cout/64): wd[c, tap*kp + o] = W[tap,c,o], zero for o >= Cout.
 *   wgrad: dw (fp32 HWIO [3,3,Cin,Cout]) += sum_pixels x(shifted) * dy; split over the pixel axis, partial tiles combined
 *          with TMA reduce-adds (Cout % 4 == 0) or fp32 atomics.
 * Cin a multiple of 8; other shapes / strides / dilations: dlv3p_im2col3x3 + dlv3p_gemm_bf16. */
int dlv3p_conv3x3_same_fwd_bf16(const void* x, const void* wt, int64_t ldw, void* y, int c_dtype, int N, int H, int W,
                                int Cin, int Cout, const float* col_scale, const float* col_shift, int act,
                                float* col_stats, void* stream);
int dlv3p_conv3x3_same_dgrad_bf16(const void* dy, int64_t ld_dy, const void* wd, void* dx, int N, int H, int W, int Cin,
                                  int Cout, void* stream);
int dlv3p_conv3x3_same_wgrad_bf16(const void* x, const void* dy, int64_t ld_dy, float* dw, int N, int H, int W, int Cin,
                                  int Cout, void* stream);
/* generic fp32-accumulate SIMT GEMM for the fp32 parity mode and shapes the TMA path cannot take:
 *   C[m,n] = sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ C if accumulate), same epilogue as above.
 *   ab_dtype: storage of A and B; c_dtype: storage of C. */
int dlv3p_gemm_simt(const void* A, int64_t sam, int64_t sak, const void* B, int64_t sbk, int64_t sbn, void* C,
                    int64_t ldc, int M, int N, int K, int ab_dtype, int c_dtype, const float* col_scale,
                    const float* col_shift, int act, const void* addend, int64_t ld_addend, int accumulate,
                    void* stream);

/* im2col for dense 3x3 convs (Xception block1_conv1/2, decoder logits conv ss.py:893-897):
 * col[(n,ho,wo), tap*C + c] = x[n, ho*stride - pad_t + i*dil, wo*stride - pad_l + j*dil, c] (0 outside);
 * columns [9C, ld_col) are zero-filled.  col2im is its transpose (gradient w.r.t. x). */
int dlv3p_im2col3x3(const void* x, void* col, int N, int H, int W, int C, int stride, int dil, int pad_t,
                    int pad_l, int Ho, int Wo, int64_t ld_col, int dtype, void* stream);
int dlv3p_col2im3x3(const void* col, void* dx, int N, int H, int W, int C, int stride, int dil, int pad_t,
                    int pad_l, int Ho, int Wo, int64_t ld_col, const void* addend, int dtype, void* stream);
/* strided 1x1 conv input gather (Xception residual Conv2D(1x1, strides 2)): y[n,ho,wo,:] = x[n,ho*s,wo*s,:] */
int dlv3p_subsample_fwd(const void* x, void* y, int N, int H, int W, int C, int stride, int Ho, int Wo, int dtype,
                        void* stream);
/* dx = scatter(dy) on the sampled positions, 0 elsewhere (+ addend) */
int dlv3p_subsample_bwd(const void* dy, void* dx, int N, int H, int W, int C, int stride, int Ho, int Wo,
                        const void* addend, int dtype, void* stream);
/* fp32 master weight [K,N] -> bf16 copies: wt [N,ldt] (transposed, K contiguous, zero padded to ldt) and,
 * if wn != NULL, wn [K,ldn] (same orientation, N contiguous, zero padded).  Used once per optimizer step. */
int dlv3p_weight_prep(const float* w, int K, int N, void* wt, int64_t ldt, void* wn, int64_t ldn, void* stream);
/* the same for `count` weights in one launch; `table` is a DEVICE array of 48-byte entries
 * { const float* w; bf16* wt; bf16* wn; int64 ldt; int64 ldn; int32 K; int32 N; } */
int dlv3p_weight_prep_batch(const void* table, int count, int blocks_per_entry, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K3 — memory-bound fused kernels.
 * BatchNormalization (TF FusedBatchNormV3 / FusedBatchNormGradV3; eps 1e-3 Keras default):
 * ---------------------------------------------------------------------------------------------- */
/* sums[0..C) += sum_m y[m,c]; sums[C..2C) += sum_m y[m,c]^2  (caller zeroes sums) */
int dlv3p_bn_stats(const void* y, int64_t ld, int64_t M, int C, float* sums, int dtype, void* stream);
/* training: mean = S1/M, var = S2/M - mean^2 (biased); scale = gamma*rsqrt(var+eps), shift = beta - mean*scale;
 * moving_mean = mom*moving_mean + (1-mom)*mean; moving_var likewise with the UNBIASED variance var*M/(M-1)
 * (TF fused batch norm convention).  gamma/beta nullable (scale=False / center=False). */
int dlv3p_bn_finalize(const float* sums, const float* gamma, const float* beta, float* moving_mean,
                      float* moving_var, int C, double count, float eps, float momentum, float* scale,
                      float* shift, float* mean, float* invstd, int update_moving, void* stream);
/* bn_finalize + affine_act in one launch (the training forward of BatchNormalization + Activation + Add):
 * statistics are finished per thread, scale/shift/mean/invstd are published for the backward pass and the moving
 * statistics receive `updates` momentum updates (0 = frozen; 2 = layer applied at two call sites).  C % 8 == 0. */
int dlv3p_bn_train_apply(const void* y, int64_t ld_y, const float* sums, const float* gamma, const float* beta,
                         float* moving_mean, float* moving_var, double count, float eps, float momentum, int updates,
                         int act, const void* addend, int64_t ld_addend, void* out, int64_t ld_out, int64_t M, int C,
                         float* scale, float* shift, float* mean, float* invstd, int dtype, void* stream);
/* inference: scale = gamma*rsqrt(moving_var+eps), shift = beta - moving_mean*scale */
int dlv3p_bn_fold(const float* gamma, const float* beta, const float* moving_mean, const float* moving_var, int C,
                  float eps, float* scale, float* shift, void* stream);
/* out[m,c] = act(scale[c]*y[m,c] + shift[c]) (+ addend[m,c]); scale/shift nullable; BN-apply + Activation + Add */
int dlv3p_affine_act(const void* y, int64_t ld_y, const float* scale, const float* shift, int act,
                     const void* addend, int64_t ld_addend, void* out, int64_t ld_out, int64_t M, int C,
                     int dtype, void* stream);
/* backward of out = act(scale*y+shift): with g = dz * act'(scale*y+shift), xhat = (y-mean)*invstd:
 *   red[0..C) += sum_m g ( = dbeta ),  red[C..2C) += sum_m g*xhat ( = dgamma )  (caller zeroes red) */
int dlv3p_bn_bwd_reduce(const void* dz, int64_t ld_dz, const void* y, int64_t ld_y, const float* scale,
                        const float* shift, const float* mean, const float* invstd, int act, int64_t M, int C,
                        float* red, int dtype, void* stream);
/*   dy = scale * (g - red[c]/M - xhat*red[C+c]/M)   (training-mode BN input gradient).
 *   With mean == NULL (inference-mode / frozen statistics): dy = scale * g. */
int dlv3p_bn_bwd_apply(const void* dz, int64_t ld_dz, const void* y, int64_t ld_y, const float* scale,
                       const float* shift, const float* mean, const float* invstd, int act, const float* red,
                       int64_t M, int C, void* dy, int64_t ld_dy, int dtype, void* stream);

/* Activation (TF Relu/Relu6 + grads) for the places it cannot be fused */
int dlv3p_act_bwd(const void* dy, const void* x, void* dx, int act, const void* addend, int64_t n, int dtype,
                  void* stream);
/* out = a + b (TF AddV2) */
int dlv3p_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);
/* y[m, 0..C) (row stride ld_y) = x[m, 0..C) (row stride ld_x): Concatenate slices / slice gradients */
int dlv3p_copy2d(const void* x, int64_t ld_x, void* y, int64_t ld_y, int64_t M, int C, const void* addend,
                 int64_t ld_addend, int dtype, void* stream);

/* MaxPooling2D(3, strides=2, padding='same') (+ residual Add): y = maxpool(x) (+ addend); argmax[n,ho,wo,c] is the
 * winning tap index 0..8 (first maximum in row-major window order; padding never wins). */
int dlv3p_maxpool3x3s2_fwd(const void* x, void* y, uint8_t* argmax, int N, int H, int W, int C, int pad_t,
                           int pad_l, int Ho, int Wo, const void* addend, int dtype, void* stream);
int dlv3p_maxpool3x3s2_bwd(const void* dy, const uint8_t* argmax, void* dx, int N, int H, int W, int C, int pad_t,
                           int pad_l, int Ho, int Wo, const void* addend, int dtype, void* stream);
/* MaxPooling2D(3, strides 2, SAME) fused with the training-mode BatchNormalization that feeds it directly (Xception
 * block2/3/4: sepconv2_bn -> MaxPooling2D -> Add, keras.applications.xception): the BN output is never written.
 *   fwd: y = maxpool(scale*x + shift) (+ addend), x = raw conv output; argmax as dlv3p_maxpool3x3s2_fwd; ymax = the raw
 *        x value of every winner (the BN-backward reductions then run on the POOLED tensors:
 *        dlv3p_bn_bwd_reduce(dz = dy_pool, y = ymax, act NONE, M = N*Ho*Wo)).
 *   bwd: dx = scale*(g - red[c]/count - xhat*red[C+c]/count) with g = the pooled gradient dy routed through argmax and
 *        xhat = (x - mean)*invstd — max-pool backward and dlv3p_bn_bwd_apply in one pass, g is never written.
 * bf16 only; anything else returns an error and the caller uses the separate entry points. */
int dlv3p_maxpool3x3s2_bn_fwd(const void* x, const float* scale, const float* shift, void* y, void* ymax,
                              uint8_t* argmax, int N, int H, int W, int C, int pad_t, int pad_l, int Ho, int Wo,
                              const void* addend, int dtype, void* stream);
int dlv3p_maxpool3x3s2_bn_bwd(const void* dy, const uint8_t* argmax, const void* x, const float* scale,
                              const float* mean, const float* invstd, const float* red, double count, void* dx, int N,
                              int H, int W, int C, int pad_t, int pad_l, int Ho, int Wo, int dtype, void* stream);
/* AveragePooling2D(pool_size=k, padding='valid') (stride = k), ss.py:842 */
int dlv3p_avgpool_fwd(const void* x, void* y, int N, int H, int W, int C, int k, int Ho, int Wo, int dtype,
                      void* stream);
int dlv3p_avgpool_bwd(const void* dy, void* dx, int N, int H, int W, int C, int k, int Ho, int Wo,
                      const void* addend, int dtype, void* stream);
/* K.resize_images(..., 'bilinear') = TF ResizeBilinear(half_pixel_centers=True), integer factors fh,fw
 * (ss.py:852-856,904-908,941-950).  y [N,H*fh,W*fw,C] with row stride ld_y (concat slice writes). */
int dlv3p_bilinear_fwd(const void* x, int64_t ld_x, void* y, int64_t ld_y, int N, int H, int W, int C, int fh,
                       int fw, int in_dtype, int out_dtype, void* stream);
/* dx[N,H,W,C] = resize^T(dy) (+ addend) */
int dlv3p_bilinear_bwd(const void* dy, int64_t ld_dy, void* dx, int64_t ld_dx, int N, int H, int W, int C, int fh,
                       int fw, const void* addend, int dy_dtype, int dx_dtype, void* stream);
/* Inference tail (segment(), ss.py:1207-1227: K.resize_images of the logits, softmax, argmax): bilinear x(fh, fw)
 * up-sampling of the LOW-RESOLUTION fp32 logits z[N,H,W,C] fused with the channel argmax (softmax is monotone) — the
 * [N,H*fh,W*fw,C] logits / probabilities are never written.  labels[N,H*fh,W*fw]: int32 (label_bytes 4) or uint8
 * (label_bytes 1, C <= 256); first maximum wins; bit-identical to dlv3p_bilinear_fwd + dlv3p_softmax_argmax. */
int dlv3p_upsample_argmax(const float* z, void* labels, int label_bytes, int N, int H, int W, int C, int fh, int fw,
                          void* stream);

/* Activation('softmax') + ClassBalancedLoss (ss.py:909, 438-447) on an integer label map:
 *   p = softmax(z[pix,:]);  L_pix = -sum_i [ pw_i*y_i*log(p_i+eps) + nw_i*(1-y_i)*log(1-p_i+eps) ], y = onehot(label)
 *   loss_sum[0] += sum_pix L_pix  (caller zeroes; the Keras mean is loss_sum / P).
 *   z fp32 [P,C]; labels int32 [P]; probs (nullable) fp32 [P,C]. C <= 32. */
int dlv3p_softmax_cbloss_fwd(const float* z, const int32_t* labels, const float* pw, const float* nw, float eps,
                             int64_t P, int C, float* loss_sum, float* probs, void* stream);
/* dz[pix,:] = grad_scale * dL_pix/dz  (grad_scale = 1/P for the Keras mean) */
int dlv3p_softmax_cbloss_bwd(const float* z, const int32_t* labels, const float* pw, const float* nw, float eps,
                             int64_t P, int C, float grad_scale, float* dz, void* stream);
/* Fused decoder tail: bilinear x f upsample of low-res logits -> softmax -> class-balanced loss, without
 * materialising the [N,H*f,W*f,C] tensors (replaces ResizeBilinear+Softmax+~10*C elementwise ops).
 * zl fp32 [N,H,W,C]; labels int32 [N,H*f,W*f]. */
int dlv3p_upsample_softmax_cbloss_fwd(const float* zl, const int32_t* labels, const float* pw, const float* nw,
                                      float eps, int N, int H, int W, int C, int f, float* loss_sum,
                                      void* stream);
/* dzl[N,H,W,C] (fp32) += grad_scale * resize^T(softmax-loss gradient)  (caller zeroes dzl) */
int dlv3p_upsample_softmax_cbloss_bwd(const float* zl, const int32_t* labels, const float* pw, const float* nw,
                                      float eps, int N, int H, int W, int C, int f, float grad_scale, float* dzl,
                                      void* stream);
/* both of the above in ONE pass over the label map (the training step): loss_sum[0] += sum_pix L_pix and
 * dzl += grad_scale * gradient; caller zeroes loss_sum and dzl. */
int dlv3p_upsample_softmax_cbloss_fwd_bwd(const float* zl, const int32_t* labels, const float* pw, const float* nw,
                                          float eps, int N, int H, int W, int C, int f, float grad_scale,
                                          float* loss_sum, float* dzl, void* stream);
/* inference tail: probs = softmax(z) and/or label map = argmax(z) (first max wins; MeanIoUExt ss.py:310-311) */
int dlv3p_softmax_argmax(const float* z, int64_t P, int C, float* probs, int32_t* labels, void* stream);
/* the Keras-signature loss on dense tensors (one-hot or soft y_true, probabilities y_pred), fwd and grad wrt y_pred */
int dlv3p_cbloss_dense_fwd(const float* y_true, const float* y_pred, const float* pw, const float* nw, float eps,
                           int64_t P, int C, float* loss_sum, void* stream);
int dlv3p_cbloss_dense_bwd(const float* y_true, const float* y_pred, const float* pw, const float* nw, float eps,
                           int64_t P, int C, float grad_scale, float* dy_pred, void* stream);
/* softmax backward for the unfused Keras path: dz = p * (dp - sum_i dp_i p_i) */
int dlv3p_softmax_bwd(const float* p, const float* dp, int64_t P, int C, float* dz, void* stream);
/* confusion matrix of MeanIoUExt (ss.py:326-334): cm[t*C+p] += 1 (float64 accumulate) */
int dlv3p_confusion_matrix(const int32_t* y_true, const int32_t* y_pred, int64_t P, int C, double* cm,
                           void* stream);

/* Dropout (ss.py:864): y = x * mask / (1-rate), mask ~ Bernoulli(1-rate) from a counter-based hash of
 * (seed + *seed_offset, element index); bwd applies the same mask.  seed_offset (nullable) is a DEVICE counter so a
 * captured CUDA graph draws a fresh mask on every replay. */
int dlv3p_dropout(const void* x, void* y, int64_t n, float rate, uint64_t seed, const uint64_t* seed_offset,
                  const void* addend, int dtype, void* stream);

/* Adam (ss.py:477-480; Keras: lr_t = lr*sqrt(1-b2^t)/(1-b1^t), w -= lr_t*m/(sqrt(v)+eps)); g may carry an L2
 * term: g_eff = g*grad_scale + l2*2*w  (Keras regularizers.l2(l) = l*sum(w^2)). */
int dlv3p_adam(float* w, const float* g, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2,
               float eps, float grad_scale, float l2, void* stream);
/* sum of squares (for the L2 regulariser term of the reported loss): out[0] += sum w^2 */
int dlv3p_sumsq(const float* w, int64_t n, float* out, void* stream);
/* dtype conversion fp32 <-> bf16 (flat, and row-strided [M,C] with independent leading dimensions) */
int dlv3p_cast(const void* x, int x_dtype, void* y, int y_dtype, int64_t n, void* stream);
int dlv3p_cast2d(const void* x, int64_t ld_x, int x_dtype, void* y, int64_t ld_y, int y_dtype, int64_t M, int C,
                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * Input pipeline (next row): what the reference's keras Sequence does per sample on one CPU thread
 * (ss.py:1528-1560 with the helpers resize() ss.py:130-195 and resize_image_to_target_symmeric_size() ss.py:198-280).
 * `table` is a DEVICE array of `count` 48-byte entries
 *   { const uint8_t* src; double inv_fy, inv_fx; int32 h, w, hp, wp, off_y, off_x; }
 * describing decoded uint8 samples (HWC, 3 channels for images, 1 for labels): source extent h x w, resized extent
 * hp x wp, top/left zero padding, inv_f = 1/(hp/h), 1/(wp/w) evaluated in fp64 by the caller.
 *   image: out[i] [S,S,3] (out_dtype) = pad(affine_transform(2*(src/255-0.5), order=1, mode='nearest'))
 *   label: out[i] [S,S] int32 = class ids > num_classes-1 cleared, resized likewise with scipy's integer rounding,
 *          cleared again (the index map of the reference's one-hot tensor)
 * ---------------------------------------------------------------------------------------------- */
int dlv3p_preprocess_image_batch(const void* table, int count, void* out, int S, int out_dtype, void* stream);
int dlv3p_preprocess_label_batch(const void* table, int count, int32_t* out, int S, int num_classes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DLV3P_H_ */
