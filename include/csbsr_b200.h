/*
 * csbsr_b200 -- C-ABI of the B200-native CSBSR hot path (degrade -> blind SR -> segment -> AIU/AHD95).
 *
 * The reference (Yuki-11/CSBSR) has no FFI of its own: its boundary is Python.
 * Every entry point below names the reference call site it replaces (file:line, relative to the
 * reference root).  Conventions: extern "C"; caller-owned DEVICE pointers unless the name ends in
 * `_host`; explicit sizes; the CUDA stream is passed last as an opaque `void*` (cudaStream_t);
 * return 0 on success, negative on error -- the message is available from csbsr_last_error().
 * No entry point allocates device memory; scratch is passed in and sized by *_workspace_bytes().
 */
#ifndef CSBSR_B200_H
#define CSBSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSBSR_MAX_TAPS 64
#define CSBSR_MAX_PHASES 16

/* activation codes of the fused conv epilogue */
enum { CSBSR_ACT_NONE = 0, CSBSR_ACT_RELU = 1, CSBSR_ACT_LEAKY = 2, CSBSR_ACT_SIGMOID = 3 };
/* output modes of the fused conv epilogue */
enum { CSBSR_OUT_BF16_NHWC = 0, CSBSR_OUT_F32_NCHW = 1, CSBSR_OUT_F32_NHWC = 2 };

const char* csbsr_last_error(void);
int csbsr_version(void);
/* 1 when a CUDA device of compute capability 10.x is visible */
int csbsr_device_ok(void);

/* ---------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores (bf16 in, fp32 accumulate in TMEM, TMA-fed).
 * Replaces every nn.Conv2d / nn.ConvTranspose2d on the path: model/modeling/kbpn.py:266-277
 * (ConvBlock/DeconvBlock), :450-489 (UpBlock/DownBlock 8x8 stride-4 conv + deconv), :493-518 (SFT),
 * :521-602 (kernel predictor), model/modeling/pspnet_pytorch/extractors.py:37-70,112-161 (dilated
 * ResNet-34), pspnet.py:23-57 (PSP module / upsample convs).
 *
 * The convolution is described as a list of taps per output "phase":
 *   out[n, oh*os+ooh[p], ow*os+oow[p], co] = epi( sum_t sum_ci x[n, oh*stride+dh[p][t], ow*stride+dw[p][t], ci]
 *                                                           * w[widx[p][t]][co][ci] )
 * A stride-s/dilation-d/pad-q conv has one phase with dh = r*d-q; the 8x8 stride-4 transposed conv is
 * 16 phases of a 2x2 conv (os = 4).  Out-of-image taps read zeros (TMA out-of-bounds fill).
 * Epilogue: v = acc + bias[n*bias_sn + cls*bias_sc + co]; v += r0; v = act(v); v *= rm; v += r1_sign*r1.
 * `cls` is the border class of the output pixel for spatially-constant conditioning folded into a
 * per-sample bias (cls_bw = 0: none, 1: 3x3 classes, 2: 5x5 classes).
 * ------------------------------------------------------------------------------------------- */
typedef struct csbsr_conv_desc {
    /* input activation: NHWC bf16, `x_pitch` channels per pixel, window [x_coff, x_coff+cin) */
    const void* x;
    int32_t n, h, w, x_pitch, x_coff, cin;          /* cin % 64 == 0 */
    /* packed weights: [w_taps][cout_pad][cin] bf16 (K-major) */
    const void* wgt;
    int32_t w_taps, cout_pad;                       /* cout_pad % 16 == 0 */
    /* taps */
    int32_t nphases, ntaps, stride;
    int8_t dh[CSBSR_MAX_TAPS], dw[CSBSR_MAX_TAPS];  /* [phase*ntaps + t] */
    int16_t widx[CSBSR_MAX_TAPS];
    /* output tile space and mapping into the stored output image */
    int32_t oh, ow;                                 /* per-phase output rows / cols */
    int32_t os;                                     /* output pixel stride (1, or 4 for the deconv) */
    int8_t ooh[CSBSR_MAX_PHASES], oow[CSBSR_MAX_PHASES];
    int32_t yh, yw;                                 /* stored output image size */
    int32_t out_mode;                               /* CSBSR_OUT_* */
    void* y;
    int32_t y_pitch, y_coff, cout_store;            /* bf16: %8==0 channels written; f32 planes: <= 16 */
    /* epilogue */
    const float* bias;                              /* may be NULL */
    int32_t bias_sn, bias_sc, cls_bw;
    int32_t act;
    float slope;
    const void* r0; int32_t r0_pitch, r0_coff;      /* bf16 NHWC at output pixels, pre-activation add */
    const void* rm; int32_t rm_pitch, rm_coff;      /* post-activation multiply */
    const void* r1; int32_t r1_pitch, r1_coff;      /* post-activation add (r1_sign = +1) / subtract (-1) */
    float r1_sign;
    const float* r32;                               /* f32 planar residual added after the activation (out_mode 1) */
    int32_t block_n;                                /* 0 = auto */
    /* out_mode 1: y / r32 are channel windows [coff, coff+cout_store) of planar buffers with `pitch` channels
     * (y_pitch / y_coff above and r32_pitch / r32_coff; pitch 0 = exactly cout_store channels) */
    int32_t r32_pitch, r32_coff;
    /* nsub > 1: every phase p is a TAP CLASS shared by `nsub` output sub-phases (the 8x8 stride-4 transposed conv has 4
     * classes of 4 sub-phases: the output phases whose rows / columns are both in {0,1} or {2,3} read the same 2x2 input
     * taps).  The sub-phases of a class are evaluated by ONE GEMM with N = 2 * cout_pad (two sub-phases per tile share the
     * A operand): widx is indexed [(p * ntaps + t) * nsub + s], ooh / oow [p * nsub + s]; cout_pad must be 128, the output
     * bf16 NHWC with whole tiles per image.  0 / 1: plain phases (one weight slice per tap). */
    int32_t nsub;
} csbsr_conv_desc;

/* Zero-initialise the descriptor (memset) before filling it: launch plans (tensor maps, tiling, launch geometry) are cached by
 * the descriptor's bytes, so a step that repeats its descriptors pays the planning once. */
int csbsr_conv_igemm(const csbsr_conv_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weight gradient of a convolution (csrc/conv_wgrad.cu): the cuDNN wgrad behind loss.backward()
 * (model/engine/trainer.py:57-72 through autograd of nn.Conv2d / nn.ConvTranspose2d, kbpn.py:266-277).
 *   wg[m][t][c] = sum_{n,y,x} g[n, y, x, m] * s[n, y*stride + dh[t], x*stride + dw[t], c]     (fp32)
 * nn.Conv2d: g = dL/dy, s = layer input (m = cout, c = cin, dh = r*dilation - pad).
 * nn.ConvTranspose2d: g = layer input, s = dL/dy (m = cin, c = cout, dh = r - pad, stride = the layer's stride).
 * Both tensors are NHWC bf16 channel windows; cg and cs are multiples of 64 (zero-padded channels).
 * `wg` holds round_up(cg, 128) * ntaps * cs floats and is overwritten.  Out-of-image taps read zeros.
 * ------------------------------------------------------------------------------------------- */
typedef struct csbsr_wgrad_desc {
    const void* g;
    int32_t n, gh, gw, g_pitch, g_coff, cg;
    const void* s;
    int32_t sh, sw, s_pitch, s_coff, cs;
    int32_t ntaps, stride;
    int8_t dh[CSBSR_MAX_TAPS], dw[CSBSR_MAX_TAPS];
    float* wg;                 /* [round_up(cg,128)][ntaps][cs] fp32 result; may be NULL when `grad` is given */
    /* Optional workspace of csbsr_conv_wgrad_workspace_bytes(d) bytes (16-byte aligned): the pixel splits then write their
     * partial sums with plain coalesced stores and a second kernel reduces them in a fixed order -- no atomics, no zero-fill,
     * bit-reproducible.  Without it the splits add into the zeroed `wg` with red.global. */
    void* ws;
    size_t ws_bytes;
    /* Optional (needs ws): fold the result straight into a parameter gradient in the parameter's own layout,
     * grad[a][grad_b0 + b][tap] += wg[a][tap][b] for a < grad_a, b < grad_b (second axis of the parameter: grad_btot entries);
     * grad_cp > 0: tap-expanded 3x3 accumulator, grad[m][grad_b0 + c][t] += wg[t * grad_cp + m][0][c] (m < grad_a, t < 9). */
    float* grad;
    int32_t grad_a, grad_b, grad_btot, grad_b0, grad_cp;
} csbsr_wgrad_desc;

size_t csbsr_conv_wgrad_workspace_bytes(const csbsr_wgrad_desc* d);
int csbsr_conv_wgrad(const csbsr_wgrad_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training-step support (csrc/train.cu).
 * csbsr_prelu_fwd / _bwd: single-parameter PReLU on `n` bf16 values (n % 8 == 0); `slope` is a device float;
 *   _bwd writes dx and overwrites *dslope with sum_{x<0} dy*x accumulated in fp32 (nn.PReLU of ConvBlock /
 *   DeconvBlock, model/modeling/kbpn.py:190-248).
 * csbsr_adam_step: torch.optim.Adam(lr, betas, eps) without weight decay / amsgrad on flat fp32 buffers of n
 *   elements (n % 4 == 0): g is first multiplied by grad_scale (1/world_size after a SUM all-reduce), `step` counts
 *   from 1 (bias corrections), zero_grad != 0 clears g for the next step (train.py:91; trainer.py:61,70).
 * ------------------------------------------------------------------------------------------- */
int csbsr_prelu_fwd(const void* x, void* y, const float* slope, long long n, void* stream);
size_t csbsr_prelu_bwd_workspace_bytes(void);   /* per-block partials of the slope gradient, summed in a fixed order */
int csbsr_prelu_bwd(const void* x, const void* dy, void* dx, const float* slope, float* dslope, long long n, float* workspace,
                    void* stream);
int csbsr_adam_step(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                    int step, float grad_scale, int zero_grad, void* stream);

/* Backward of csbsr_blur_per_sample (stride s, zero padding (k-1)/2; KBlock pseudo-LR kbpn.py:395-402, Get_pseudo_lr
 * sr_loss_functions.py:73-102) and of csbsr_resize_bicubic_aa (FactorResize, transforms.py:516-531); fp32 planar tensors.
 *   csbsr_blur_ps_bwd_input : dx[b,c,h,w]  from dy[b,c,ceil(h/s),ceil(w/s)] and the per-sample kernels kvec[b,k*k]
 *   csbsr_blur_ps_bwd_kernel: dk[b,k*k] (overwritten) = sum_{c,Y,X} dy[b,c,Y,X] * x[b,c,Y*s+i-pad,X*s+j-pad]
 *   csbsr_resize_bicubic_aa_bwd: dx[nc,h,w] from dy[nc,oh,ow] (transpose of the normalised antialiased taps) */
int csbsr_blur_ps_bwd_input(const float* dy, const float* kvec, float* dx, int b, int c, int h, int w, int ksize, int stride,
                            void* stream);
size_t csbsr_blur_ps_bwd_kernel_workspace_bytes(int b, int c, int h, int w, int ksize, int stride);
int csbsr_blur_ps_bwd_kernel(const float* x, const float* dy, float* dk, int b, int c, int h, int w, int ksize, int stride,
                             float* workspace, void* stream);
int csbsr_resize_bicubic_aa_bwd(const float* dy, float* dx, int nc, int h, int w, int oh, int ow, void* stream);

/* fp32 parameter [A][B][R][S] -> packed bf16 conv operand [R*S][rows_pad][cols_pad] (zero padding) in one launch.
 * mode 0: rows = A, cols = B (nn.Conv2d weight for the forward conv; ConvTranspose2d weight for its dgrad);
 * mode 1: rows = B, cols = A with the taps flipped (dgrad of a stride-1 conv);
 * mode 2: rows = B, cols = A (ConvTranspose2d forward phases; dgrad of the 8x8/s4 conv). */
int csbsr_pack_weights(const float* w, void* out, int a, int b, int r, int s, int rows_pad, int cols_pad, int mode, void* stream);
/* the same for the input-channel window [b0, b0 + b) of a parameter whose second axis has b_total entries (the feature /
 * conditioning halves of SFTlayer's conv0, kbpn.py:513-516, are packed straight from the one parameter) */
int csbsr_pack_weights_window(const float* w, void* out, int a, int b, int b_total, int b0, int r, int s, int rows_pad,
                              int cols_pad, int mode, void* stream);
/* all (parameter, layout) pairs of a model in ONE launch per optimisation step: `jobs_device` is an array of njobs opaque
 * job records (csbsr_pack_job_bytes() each, filled on the host by csbsr_pack_job_fill and copied to the device by the caller),
 * `start` = exclusive prefix of the jobs' tile counts csbsr_pack_job_tiles(...), `total_tiles` = their sum, max_taps = the
 * largest r*s among the jobs (<= 64).  Only the valid region of every packed buffer is written: zero-initialise them once. */
long long csbsr_pack_job_tiles(int a, int b, int r, int s, int rows_pad, int cols_pad, int mode);
size_t csbsr_pack_job_bytes(void);
int csbsr_pack_job_fill(void* job_host, const float* w, void* out, int a, int b, int b_total, int b0, int r, int s, int rows_pad,
                        int cols_pad, int mode, unsigned long long start);
int csbsr_pack_weights_multi(const void* jobs_device, int njobs, unsigned long long total_tiles, int max_taps, void* stream);
/* grad[a][b0 + b][tap] += wg[a][tap][b]: folds the accumulator of csbsr_conv_wgrad ([rows][taps][cs] fp32) into the parameter
 * gradient in the parameter's own [A][b_total][R][S] layout (autograd's permute + contiguous + accumulate, trainer.py:69) */
int csbsr_wgrad_unpack_add(const float* wg, float* grad, int a, int b, int b_total, int b0, int taps, int cs, void* stream);
/* the same for the tap-expanded accumulator: grad[m][b0 + c][t] += wg[t * cp + m][c] */
int csbsr_wgrad_unpack_add_tapexp(const float* wg, float* grad, int a, int b, int b_total, int b0, int cp, int cs, void* stream);

/* BatchNorm2d of the training graph on NHWC bf16 maps [m][pitch] whose first c channels are real (nn.BatchNorm2d in
 * pspnet_pytorch/extractors.py:52-70, pspnet.py:44-57, hrnet_backbone.py; reference runs them through cuDNN / aten).
 *   csbsr_bn_stats   : batch mean / rstd (biased variance, eps) into mean[c] / rstd[c]; updates running_mean / running_var with
 *                      `momentum` and the unbiased variance when they are given.
 *   workspace (both reductions): csbsr_bn_workspace_bytes(c) bytes of per-block partial sums, added in block order --
 *                      statistics and dgamma / dbeta are bit-reproducible (no atomics).
 *   csbsr_bn_apply   : y = relu?((x - mean) * rstd * gamma + beta + res?)   (res may be NULL; padding channels written as 0)
 *   csbsr_bn_backward: dy -> dx (+ dres = dy masked by y_relu > 0 when y_relu is given), dgamma[c], dbeta[c] (overwritten);
 *                      training != 0 subtracts the batch-statistics terms, 0 treats mean / rstd as constants (eval mode). */
size_t csbsr_bn_workspace_bytes(int c);
int csbsr_bn_stats(const void* x, int pitch, int c, long long m, float eps, float momentum, float* mean, float* rstd,
                   float* running_mean, float* running_var, float* workspace, void* stream);
int csbsr_bn_apply(const void* x, const void* res, void* y, const float* mean, const float* rstd, const float* gamma,
                   const float* beta, int pitch, int c, long long m, int relu, void* stream);
int csbsr_bn_backward(const void* dy, const void* x, const void* y_relu, const float* mean, const float* rstd, const float* gamma,
                      int pitch, int c, long long m, int training, void* dx, void* dres, float* dgamma, float* dbeta,
                      float* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * HBM-bound support kernels (csrc/support.cu).  NHWC tensors are bf16 with `*_pitch` channels per
 * pixel and a channel window starting at `*_coff`; planar tensors are fp32 NCHW.
 * ------------------------------------------------------------------------------------------- */
/* boundary layout change: fp32 NCHW (reference tensors) -> bf16 NHWC window; channels [c, cwrite) are zeroed */
int csbsr_nchw_f32_to_nhwc_bf16(const float* x, void* y, int n, int c, int h, int w, int y_pitch, int y_coff,
                                int cwrite, void* stream);
/* im2col of a few-channel fp32 image: y[n,oh,ow,(r*S+s)*c+ci] = f(x[n,ci,oh*stride+r-pad,ow*stride+s-pad]);
 * f optionally clamps to [0,1] and instance-normalises (MetaSRModel.clip_sr / norm_sr,
 * model/modeling/build_model.py:135-146).  Feeds the first conv of VGG feat (kbpn.py:42-44), fe_SR.0
 * (kbpn.py:528), KBlock.up_conv1 (kbpn.py:375) and ResNet conv1 (extractors.py:115). */
int csbsr_patchify(const float* x, void* y, int n, int c, int h, int w, int oh, int ow, int r, int s, int stride,
                   int pad, int y_pitch, int cwrite, const float* mean, const float* rstd, int clamp01, void* stream);
/* Fused kernel-predictor chains of KBlock (KernelPredictorLikeIKC.forward, model/modeling/kbpn.py:562-578, layers :528-541):
 * every intermediate of the chain stays in shared memory / TMEM (csrc/kpred_chain.cu).
 *   sr chain : img fp32 [b,3,h,w] -> fe_SR.0 (3x3, ReLU) -> fe_SR.1 (1x1) -> fe_SR.2 -> fe_SR.3 -> fe_SR.4 (3x3, LeakyReLU `slope`)
 *              -> out bf16 NHWC [b,h,w,64] (49 channels used);
 *   cat chain: in bf16 NHWC [b,h,w,64] -> fe_cat.0 (1x1 over the image branch + cls_bias fp32 [b,5,5,64]: the kernel branch
 *              fe_kernel(...) folded into a per-sample border-class bias) -> fe_cat.1 -> fe_cat.2 -> nn.AdaptiveAvgPool2d(1)
 *              -> gap_out fp32 [b, gap_c] (mean over h*w).
 * wpack: the chain's weights in the kernel's shared-memory layout, per layer [tap][cin/8][cout_pad][8] bf16, layers in order
 * (csbsr_kpred_wpack_bytes(which) bytes; which = 0 sr chain, 1 cat chain).  ws: csbsr_kpred_workspace_bytes(b,h,w) bytes. */
size_t csbsr_kpred_wpack_bytes(int which);
size_t csbsr_kpred_workspace_bytes(int b, int h, int w);
int csbsr_kpred_sr_chain(const float* img, const void* wpack, void* out, int b, int h, int w, float slope, void* stream);
int csbsr_kpred_cat_chain(const void* in, const void* wpack, const float* cls_bias, float* gap_out, int gap_c, void* ws,
                          size_t ws_bytes, int b, int h, int w, float slope, void* stream);
/* nn.AdaptiveAvgPool2d(1) (kbpn.py:324,391,565,572): out[n][c] = mean over h*w, fp32 */
int csbsr_gap_nhwc(const void* x, float* out, int n, int hw, int pitch, int coff, int c, void* stream);
/* out[b] = norm((pre ? pre[b] : 0) + bicubic_{ke x ke -> ko x ko}(v[b])), norm = divide by the sum when
 * `normalize` (predictor_withGAP.upscale_and_reshape kbpn.py:335-341; KernelPredictorLikeIKC.forward :574-578
 * followed by KBlock's renormalisation :391-392) */
int csbsr_kernel_update(const float* v, const float* pre, float* out, int b, int ke, int ko, int normalize,
                        void* stream);
/* out[b] = v[b] / sum(v[b]) (JointModel.forward, build_model.py:491-494) */
int csbsr_vec_normalize(const float* v, float* out, int b, int len, void* stream);
/* spatially constant conditioning: y[n, :, :, coff + i] = v[n][i] (i < len), zero up to cwrite */
int csbsr_broadcast_vec(const float* v, void* y, int n, int hw, int len, int y_pitch, int y_coff, int cwrite,
                        void* stream);
/* per-sample depthwise blur of a planar fp32 image with its own ksize x ksize kernel, zero padding
 * (ksize-1)/2, given stride; err = blur - lr when lr != NULL (KBlock.forward kbpn.py:395-405; stride 1:
 * Get_pseudo_lr model/utils/sr_loss_functions.py:90-94 and conv_kernel2d model/data/blur/blur.py:182-200) */
int csbsr_blur_per_sample(const float* x, const float* kvec, const float* lr, float* err, int n, int c, int h, int w,
                          int ksize, int stride, void* stream);
/* nn.Upsample(scale_factor=f, mode='bicubic') on planar fp32 (kbpn.py:70,113) */
int csbsr_bicubic_upsample(const float* x, float* y, int nc, int h, int w, int factor, void* stream);
/* clip to [0,1] in place (optional) + InstanceNorm2d statistics: mean[nc], rstd[nc] = 1/sqrt(biased var + eps)
 * (build_model.py:135-146) */
size_t csbsr_instnorm_workspace_bytes(int nc);
int csbsr_clip_instnorm_stats(float* x, float* mean, float* rstd, void* workspace, int nc, int hw, int do_clip,
                              float eps, void* stream);
/* nn.MaxPool2d(3, 2, 1) (extractors.py:119) */
int csbsr_maxpool3s2_nhwc(const void* x, void* y, int n, int h, int w, int c, int x_pitch, int x_coff, int y_pitch,
                          int y_coff, void* stream);
/* nn.AdaptiveAvgPool2d((s, s)) (pspnet.py:32) */
int csbsr_adaptive_avgpool_nhwc(const void* x, void* y, int n, int h, int w, int s, int c, int x_pitch, int x_coff,
                                int y_pitch, int y_coff, void* stream);
/* F.interpolate(mode='bilinear', align_corners=...) (pspnet.py:39,56) */
int csbsr_bilinear_nhwc(const void* x, void* y, int n, int h, int w, int oh, int ow, int c, int x_pitch, int x_coff,
                        int y_pitch, int y_coff, int align_corners, void* stream);
int csbsr_bilinear_f32(const float* x, float* y, int nc, int h, int w, int oh, int ow, int align_corners, void* stream);
/* sigmoid(F.interpolate(logits, bilinear, align_corners)) (hrnet_ocr/nets/hrnet.py:156-157) */
int csbsr_bilinear_f32_sigmoid(const float* x, float* y, int nc, int h, int w, int oh, int ow, int align_corners,
                               void* stream);
/* y = [relu](base + F.interpolate(x, bilinear)): multi-resolution fusion of HighResolutionModule.forward
 * (hrnet_ocr/backbones/hrnet/hrnet_backbone.py:274-288) */
int csbsr_bilinear_add_nhwc(const void* x, const void* base, void* y, int n, int h, int w, int oh, int ow, int c,
                            int x_pitch, int x_coff, int b_pitch, int b_coff, int y_pitch, int y_coff, int align_corners,
                            int relu, void* stream);
/* SpatialGather_Module.forward for one class (hrnet_ocr/modules/spatial_ocr_block.py:59-66): ctx[n][c] =
 * sum_hw softmax_hw(logits[n])[hw] * feats[n][hw][c]; logits fp32 [n,hw], feats bf16 NHWC, ctx fp32 [n,c] */
int csbsr_softmax_gather(const float* logits, const void* feats, float* ctx, int n, int hw, int c, int f_pitch,
                         int f_coff, void* stream);
/* 3x3 conv with <= 12 output channels as tap expansion: `z` is the bf16 NHWC result of the 1x1 conv whose output
 * channel t*cp + c (cp = co rounded up to 4) holds tap t (row-major 3x3) of output channel c; this gathers
 * out[n][out_coff+c][y][x] = (r32 ? r32[n][r_coff+c][y][x] : 0) + sum_t z[n][y+t/3-1][x+t%3-1][t*cp+c]  (fp32 planar,
 * zero outside the image).  sr_reconst / output_conv of KBPN (model/modeling/kbpn.py:361, :68, :113). */
int csbsr_tap_gather3x3(const void* z, int z_pitch, int z_coff, float* out, int out_pitch, int out_coff, const float* r32,
                        int r_pitch, int r_coff, int n, int h, int w, int co, void* stream);


/* ---------------------------------------------------------------------------------------------
 * Synthetic degradation (csrc/degrade.cu).
 * ------------------------------------------------------------------------------------------- */
/* GaussianBlur.make (model/data/blur/blur.py:128-168): params[b] = (theta [rad], sigma_x, sigma_y) fp64, the draws
 * of :129 and get_deterioration :170-179 made by the caller; kernels[b] = ksize x ksize fp32, normalised in fp64 */
int csbsr_blur_kernel_synth(const double* params, float* kernels, int b, int ksize, void* stream);
/* torchvision Resize(BICUBIC) on tensors = antialiased bicubic, a = -0.5 (FactorResize.__call__,
 * model/data/transforms/transforms.py:516-531); clamp01 as make_test_blur.py:62 */
int csbsr_resize_bicubic_aa(const float* x, float* y, int nc, int h, int w, int oh, int ow, int clamp01, void* stream);
/* CrackDataSet.__getitem__ degradation (model/data/crack_dataset.py:51-62): kernel synth -> conv_kernel2d
 * (blur.py:182-200) -> FactorResize.  hr, blurred: fp32 [b,c,h,w]; lr: fp32 [b,c,h/factor,w/factor] */
int csbsr_degrade(const float* hr, const double* params, float* kernels, float* blurred, float* lr, int b, int c,
                  int h, int w, int ksize, int factor, int clamp01, void* stream);
/* The same degradation in ONE pass over hr (crack_dataset.py:51-62 = blur.py:128-200 + transforms.py:516-531): blur and
 * antialiased bicubic are composed into 36x36 stride-4 kernels (25 per sample: 5 row x 5 column classes of bicubic weights,
 * the first / last two outputs of each axis use aten's shorter renormalised taps), so there is no `blurred` tensor and
 * 5.4x fewer FLOPs.  ksize = 21 and factor = 4 only; workspace = csbsr_degrade_workspace_bytes(b) bytes, 16-byte aligned. */
size_t csbsr_degrade_workspace_bytes(int b);
int csbsr_degrade_fused(const float* hr, const double* params, float* kernels, float* lr, void* workspace,
                        size_t workspace_bytes, int b, int c, int h, int w, int ksize, int factor, int clamp01, void* stream);
/* Throughput mode of the per-sample draws of GaussianBlur.make (blur.py:129: theta = U(0,180) deg; get_deterioration
 * :170-179: sigma = U(0.2, 4)): params[i] = (theta, sigma_x, sigma_y) fp64 from Philox4x32-10 with key = seed and counter =
 * offset + i (u = word * 2^-32).  Parity mode keeps the draws on the host (torch.rand / np.random.rand replayed). */
int csbsr_degrade_params_philox(double* params, int b, unsigned long long seed, unsigned long long offset, double theta_lo,
                                double theta_hi, double sigma_lo, double sigma_hi, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training-step glue (csrc/glue.cu): what the reference's train step runs between its convolutions, and what autograd runs
 * behind them (trainer.py:57-72 loss.backward()).  NHWC bf16 maps, channel windows (pitch, coff) in multiples of 8.
 * ------------------------------------------------------------------------------------------- */
/* bias gradient of nn.Conv2d / nn.ConvTranspose2d (kbpn.py:266-277): out[c] = sum over rows of dy[row, coff + c], fp32, fixed order */
size_t csbsr_colsum_workspace_bytes(int c);
int csbsr_bias_grad(const void* dy, int pitch, int coff, int c, long long rows, float* out, void* workspace,
                    size_t workspace_bytes, void* stream);
/* backward of ReLU (slope 0) / LeakyReLU on the saved output: dx = dy * (y > 0 ? 1 : slope) (kbpn.py:196-214, 513-516) */
int csbsr_act_bwd(const void* dy, const void* y, void* dx, long long n, float slope, void* stream);
/* out = [relu](alpha * a + beta * b), b may be NULL: residual adds / subs of the projection units (kbpn.py:464-469, 484-489) */
int csbsr_axpby(const void* a, const void* b, void* out, long long n, float alpha, float beta, int relu, void* stream);
/* SFTlayer.forward (kbpn.py:516-518): out = f * sigmoid(s) + t; backward df = dy*sig(s), ds = dy*f*sig*(1-sig) (dt = dy) */
int csbsr_sft_combine(const void* f, const void* s, const void* t, void* out, long long n, void* stream);
int csbsr_sft_combine_bwd(const void* dy, const void* f, const void* s, void* df, void* ds, long long n, void* stream);
/* torch.cat along channels and its backward slices (kbpn.py:173-186, pspnet.py:40): copy a c-channel window, optionally
 * zeroing `zero_tail` channels after it in the destination */
int csbsr_window_copy(const void* src, int src_pitch, int src_coff, void* dst, int dst_pitch, int dst_coff, int c, int zero_tail,
                      long long rows, void* stream);
/* backward of csbsr_bilinear_nhwc (F.interpolate bilinear, pspnet.py:39,56): gather form, exact transpose of the forward */
int csbsr_bilinear_nhwc_bwd(const void* dy, void* dx, int n, int h, int w, int oh, int ow, int c, int dy_pitch, int dy_coff,
                            int dx_pitch, int dx_coff, int align_corners, void* stream);
/* backward of csbsr_adaptive_avgpool_nhwc (nn.AdaptiveAvgPool2d, pspnet.py:32) */
int csbsr_adaptive_avgpool_nhwc_bwd(const void* dy, void* dx, int n, int h, int w, int s, int c, int dy_pitch, int dy_coff,
                                    int dx_pitch, int dx_coff, void* stream);
/* backward of csbsr_maxpool3s2_nhwc (extractors.py:119): the gradient goes to the first maximum of each window */
int csbsr_maxpool3s2_nhwc_bwd(const void* x, const void* dy, void* dx, int n, int h, int w, int c, int x_pitch, int x_coff,
                              int dy_pitch, int dy_coff, int dx_pitch, int dx_coff, void* stream);
/* nn.Dropout2d (pspnet.py:67,73,83): per-(sample, channel) keep mask scaled by 1/(1-p) from Philox4x32-10 keyed by `seed`,
 * counter (index, salt, *counter) -- the step counter is read on the device so CUDA-graph replays draw fresh masks;
 * csbsr_channel_scale applies it (forward and backward) */
int csbsr_dropout2d_mask(float* scale, int n, int c, int c_pad, float p, unsigned long long seed,
                         const unsigned long long* counter, unsigned int salt, void* stream);
int csbsr_counter_inc(unsigned long long* counter, void* stream);
int csbsr_channel_scale(const void* x, const float* scale, void* y, int n, long long hw, int c_pad, void* stream);
/* spatially constant conditioning (kbpn.py:404, 513-516): [n, 2bw+1, 2bw+1, c] border-class responses -> [n, h, w, c], and the
 * per-class sums of the gradient (two fixed-order stages) */
int csbsr_expand_classes(const void* small, void* out, int n, int h, int w, int bw, int c_pad, void* stream);
size_t csbsr_expand_classes_workspace_bytes(int n, int h, int bw, int c_pad);
int csbsr_expand_classes_bwd(const void* dy, void* dsmall, int n, int h, int w, int bw, int c_pad, void* workspace,
                             size_t workspace_bytes, void* stream);
/* norm_sr 'instance' in training (build_model.py:135-137, no clip): y = (x - mean) * rstd with the statistics of
 * csbsr_clip_instnorm_stats(do_clip = 0), and its backward dx = rstd * (dy - mean(dy) - xhat * mean(dy * xhat)) */
int csbsr_instnorm_apply(const float* x, const float* mean, const float* rstd, float* y, int nc, long long hw, void* stream);
size_t csbsr_instnorm_bwd_workspace_bytes(int nc);
int csbsr_instnorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, float* dx, int nc, long long hw,
                       void* workspace, size_t workspace_bytes, void* stream);
/* 3x3 / padding-1 convs with <= 4 outputs (KBlock.sr_reconst kbpn.py:361, output_conv :68) in the training step: one 1x1 GEMM
 * with 9 * cp outputs (pack modes 3 / 4 of csbsr_pack_weights_window, cp in mode >> 3), then y = gather of the nine shifted
 * taps; its backward scatters dy back to the tap-expanded layout, so dgrad and wgrad are 1x1 GEMMs too (9x fewer MMAs) */
int csbsr_tapexp_gather_nhwc(const void* z, int z_pitch, void* y, int y_pitch, int n, int h, int w, int cp, int co, void* stream);
int csbsr_tapexp_scatter_nhwc(const void* dy, int dy_pitch, void* dz, int z_pitch, int n, int h, int w, int cp, int co, void* stream);
/* Input pipeline.  SplitPatch / JointPatch (model/data/samplers/patch_sampler.py:15-50): fp32 images [b, c, h, w] <-> their
 * non-overlapping ph x pw patches [b * (h/ph) * (w/pw), c, ph, pw] (unfold with stride = size: remainders dropped on split) */
int csbsr_patch_split(const float* img, float* patches, int b, int c, int h, int w, int ph, int pw, void* stream);
int csbsr_patch_join(const float* patches, float* img, int b, int c, int h, int w, int ph, int pw, void* stream);
/* CrackDataSet.__getitem__ augmentation (crack_dataset.py:42-48, data_preprocess.py:13-46; RandomMirror / RandomVerticalFlip /
 * RandomCrop / ToTensor / 255) on decoded uint8 HWC images resident on the device: imgs[b] -> image b, dims[b] = (H, W, C),
 * params[b] = (y0, x0, hflip, vflip) drawn by the caller; out fp32 [b, c_out, th, tw] = pixel / divisor (255 for ToTensor; c >= C repeats the last) */
int csbsr_crop_flip_u8(const unsigned char* const* imgs, const int* dims, const int* params, float* out, int b, int c_out, int th,
                       int tw, float divisor, void* stream);
/* NHWC bf16 window -> fp32 NCHW (the first c channels): images / logits leaving the networks */
int csbsr_nhwc_bf16_to_nchw_f32(const void* x, float* y, int n, long long hw, int c, int x_pitch, int x_coff, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Joint-training losses (csrc/losses.cu): forward values and the gradient w.r.t. the segmentation predictions.
 * ------------------------------------------------------------------------------------------- */
/* compute_sdf1_1 (model/utils/boundary_loss.py:40-67): normalised signed distance map of mask.astype(uint8), exact EDT,
 * fp64 normalisation, 0 on the inner boundary; mask, sdf: fp32 [b,1,h,w] */
size_t csbsr_sdf_workspace_bytes(int b, int h, int w);
int csbsr_sdf(const float* mask, float* sdf, int b, int h, int w, void* workspace, size_t workspace_bytes, void* stream);
/* BoundaryComboLoss with out_map=False on the main and auxiliary heads (model/utils/loss_functions.py:49-74, 196-210,
 * 284-345; combination model/modeling/build_model.py:258-278): loss[b] = main_w*l(p_main) + aux_w*l(p_aux),
 * l = alpha*(WBCE + Dice)/2 + (1-alpha)*mean(p*sdf).  grad_* (optional): d(sum_b upstream[b]*loss[b])/dp */
size_t csbsr_seg_loss_workspace_bytes(int b);
int csbsr_seg_loss(const float* p_main, const float* p_aux, const float* target, const float* sdf, int b, int hw,
                   float alpha, float main_w, float aux_w, float* loss, float* grad_main, float* grad_aux,
                   const float* upstream, void* workspace, size_t workspace_bytes, void* stream);
/* the scalar `.mean()` of the (B,B,H,W) tensor the reference builds when the failure-oriented weight w^F =
 * exp(amp*|p_main.detach() - g|) is on (oriented_weight.py:73-83, build_model.py:422-438);
 * `out`: device double */
size_t csbsr_seg_loss_wf_workspace_bytes(int b, int hw);
int csbsr_seg_loss_wf_mean(const float* p_main, const float* p_aux, const float* target, const float* sdf, int b, int hw,
                           float alpha, float main_w, float aux_w, float wf_amp, double* out, void* workspace,
                           size_t workspace_bytes, void* stream);
/* Backward of csbsr_seg_loss_wf_mean: grad_main / grad_aux [B,HW] = (*upstream) * d mean / d prediction (w^F itself is
 * detached, oriented_weight.py:81; zero where the prediction is below the 1e-8 clamp).  `upstream` is a DEVICE float or
 * NULL (= 1).  Same workspace as the forward. */
int csbsr_seg_loss_wf_grad(const float* p_main, const float* p_aux, const float* target, const float* sdf, int b, int hw,
                           float alpha, float main_w, float aux_w, float wf_amp, const float* upstream, float* grad_main,
                           float* grad_aux, void* workspace, size_t workspace_bytes, void* stream);
/* KBPNLoss.forward (model/utils/sr_loss_functions.py:39-56) given the pseudo-LR image of Get_pseudo_lr (:84-102, built
 * with csbsr_blur_per_sample stride 1 + csbsr_resize_bicubic_aa): loss[b] = w_hr*mean|sr-hr| + w_lr*mean|plr-lr| +
 * w_k*mean((k_pred-k_gt)^2); n_* = elements per sample */
size_t csbsr_sr_loss_workspace_bytes(int b);
int csbsr_sr_loss(const float* sr, const float* hr, const float* pseudo_lr, const float* lr, const float* k_pred,
                  const float* k_gt, int b, int n_hr, int n_lr, int n_k, float w_hr, float w_lr, float w_k, float* loss,
                  void* workspace, size_t workspace_bytes, void* stream);
/* Backward of csbsr_sr_loss (autograd of KBPNLoss.forward behind loss.backward(), trainer.py:67): d_sr = upstream[b] * w_hr *
 * sign(sr - hr) / n_hr, d_plr likewise with w_lr / n_lr, d_k = upstream[b] * 2 w_k (k_pred - k_gt) / n_k (d_k may be NULL when
 * w_k = 0); `upstream`: DEVICE float[b] or NULL (= 1).  The pseudo-LR gradient continues through csbsr_resize_bicubic_aa_bwd and
 * csbsr_blur_ps_bwd_input / _bwd_kernel. */
int csbsr_sr_loss_bwd(const float* sr, const float* hr, const float* pseudo_lr, const float* lr, const float* k_pred,
                      const float* k_gt, const float* upstream, int b, int n_hr, int n_lr, int n_k, float w_hr, float w_lr,
                      float w_k, float* d_sr, float* d_plr, float* d_k, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Segmentation metrics (csrc/metrics.cu): AIU threshold sweep and the Hausdorff / mean-surface-distance
 * sweep, bit-exact with the reference.
 *   prob, mask: fp32 [b,1,h,w]; thresholds: fp32 [99] = float32(i*0.01), i = 1..99
 *   inter, uni: int64 [b,99]  -- IoU.__call__ counts (model/utils/estimate_metrics.py:72-84) for
 *               pred_i = (prob - t_i > 0) (model/engine/inference.py:49-53,111) and target = mask > 0.5
 *   hd, msd:    fp64 [b,99] or both NULL -- calc_distance_metrics (model/engine/inference.py:293-336) with
 *               gt = mask.astype(bool), `percent` the robust-Hausdorff percentile (the reference ships 50,
 *               inference.py:302); compute_surface_distances / compute_robust_hausdorff /
 *               compute_average_surface_distance of surface_distance.py:136-359.
 * ------------------------------------------------------------------------------------------- */
size_t csbsr_metrics_workspace_bytes(int b, int h, int w, int with_hd);
int csbsr_seg_metrics(const float* prob, const float* mask, const float* thresholds, int b, int h, int w,
                      long long* inter, long long* uni, double* hd, double* msd, double percent, void* workspace,
                      size_t workspace_bytes, void* stream);

/* PSNR = 10*log10(1/mse) and SSIM (11x11 Gaussian window, sigma 1.5, zero padding, C1 = 1e-4, C2 = 9e-4) per image of two
 * fp32 [b,c,h,w] tensors in [0,1]: PSNR.__call__ / SSIM.forward of model/utils/estimate_metrics.py:89-100,134-191 as used by
 * inference_for_ss (model/engine/inference.py:94-100).  psnr / ssim: DEVICE double[b]. */
size_t csbsr_psnr_ssim_workspace_bytes(int b);
int csbsr_psnr_ssim(const float* pred, const float* target, int b, int c, int h, int w, double* psnr, double* ssim,
                    void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSBSR_B200_H */
