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
enum { CSBSR_OUT_BF16_NHWC = 0, CSBSR_OUT_F32_NCHW = 1 };

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
    int32_t y_pitch, y_coff, cout_store;            /* bf16: %8==0 channels written; f32 planes: <= 8 */
    /* epilogue */
    const float* bias;                              /* may be NULL */
    int32_t bias_sn, bias_sc, cls_bw;
    int32_t act;
    float slope;
    const void* r0; int32_t r0_pitch, r0_coff;      /* bf16 NHWC at output pixels, pre-activation add */
    const void* rm; int32_t rm_pitch, rm_coff;      /* post-activation multiply */
    const void* r1; int32_t r1_pitch, r1_coff;      /* post-activation add (r1_sign = +1) / subtract (-1) */
    float r1_sign;
    const float* r32;                               /* f32 planar residual [n][cout_store][yh][yw] (out_mode 1) */
    int32_t block_n;                                /* 0 = auto */
} csbsr_conv_desc;

int csbsr_conv_igemm(const csbsr_conv_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CSBSR_B200_H */
