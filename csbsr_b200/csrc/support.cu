// HBM-bound support kernels around the tensor-core convs: layout conversion, 3-channel patch gather
// (im2col for the 3-channel image inputs), pooling, resampling, per-sample kernel-vector updates,
// per-sample depthwise blur (KBlock), clip + instance-norm statistics.  All are coalesced along the
// NHWC channel axis (128-bit vectors where the channel count allows) or along W for planar fp32.
#include <cooperative_groups.h>
#include "common.cuh"
#include "../../include/csbsr_b200.h"

namespace csbsr {

static inline int grid_for(size_t work, int threads) {
    size_t b = (work + threads - 1) / threads;
    size_t cap = static_cast<size_t>(num_sms()) * 32;
    return static_cast<int>(b < cap ? (b ? b : 1) : cap);
}

// ------------------------------------------------------------------ NCHW f32 -> NHWC bf16 (channel window)
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int C, int H,
                                    int W, int pitch, int coff, int cwrite) {
    const size_t total = static_cast<size_t>(N) * H * W * cwrite;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % cwrite);
        const size_t pix = i / cwrite;
        const int w = static_cast<int>(pix % W);
        const int h = static_cast<int>((pix / W) % H);
        const int n = static_cast<int>(pix / (static_cast<size_t>(W) * H));
        float v = c < C ? x[((static_cast<size_t>(n) * C + c) * H + h) * W + w] : 0.f;
        y[pix * pitch + coff + c] = __float2bfloat16(v);
    }
}

// 128-bit form: one thread per (pixel, 8-channel group) -> consecutive threads write consecutive 16-byte chunks of a pixel row
__global__ void nchw_to_nhwc_vec_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int C, int H, int W,
                                        int pitch, int coff, int cwrite) {
    const int G = cwrite / 8;
    const size_t HW = static_cast<size_t>(H) * W;
    const size_t total = static_cast<size_t>(N) * HW * G;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const size_t pix = i / G;
        const size_t n = pix / HW, p = pix - n * HW;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (g * 8 < C) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = g * 8 + j;
                v[j] = c < C ? x[(n * C + c) * HW + p] : 0.f;
            }
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        }
        *reinterpret_cast<uint4*>(y + pix * pitch + coff + g * 8) = o;
    }
}

// ------------------------------------------------------------------ 3-channel patch gather (im2col)
// y[n,oh,ow,(r*S+s)*C+c] = f(x[n,c,oh*st+r*dil-pad, ow*st+s*dil-pad]) (0 outside), channels >= R*S*C zero.
// f = identity, or clamp to [0,1] followed by (v-mean[n,c])*rstd[n,c] (clip_sr + norm_sr, build_model.py:135-146).
// One block = an 8 x 32 tile of output pixels: the (clamped / normalised, zero-padded) input patch is staged once in
// shared memory with coalesced reads, then each (pixel, 8-channel group) item is assembled from it and written as
// one 128-bit store (consecutive items -> consecutive 16-byte chunks of a pixel's channel vector).
__global__ void patchify_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int C, int H, int W,
                                int OH, int OW, int R, int S, int stride, int pad, int pitch, int cwrite,
                                const float* __restrict__ mean, const float* __restrict__ rstd, int clamp01) {
    constexpr int TY = 8, TX = 32;
    extern __shared__ float patch[];                 // [C][PH][PW]
    __shared__ int lut[256];                         // gathered channel k -> offset inside the patch
    const int PH = (TY - 1) * stride + R, PW = (TX - 1) * stride + S;
    const int K = R * S * C;
    const int n = blockIdx.z;
    const int oh0 = blockIdx.y * TY, ow0 = blockIdx.x * TX;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int c = k % C, rs = k / C;
        lut[k] = (c * PH + rs / S) * PW + rs % S;
    }
    const int ih0 = oh0 * stride - pad, iw0 = ow0 * stride - pad;
    const size_t plane = static_cast<size_t>(H) * W;
    const float* xn = x + static_cast<size_t>(n) * C * plane;
    for (int i = threadIdx.x; i < C * PH * PW; i += blockDim.x) {
        const int c = i / (PH * PW), r = (i / PW) % PH, col = i % PW;
        const int ih = ih0 + r, iw = iw0 + col;
        float v = 0.f;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
            v = xn[c * plane + static_cast<size_t>(ih) * W + iw];
            if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
            if (mean) v = (v - mean[n * C + c]) * rstd[n * C + c];
        }
        patch[i] = v;
    }
    __syncthreads();
    const int G = cwrite / 8;
    for (int item = threadIdx.x; item < TY * TX * G; item += blockDim.x) {
        const int g = item % G, pixel = item / G;
        const int ty = pixel / TX, tx = pixel % TX;
        const int oh = oh0 + ty, ow = ow0 + tx;
        if (oh >= OH || ow >= OW) continue;
        uint4 out = make_uint4(0, 0, 0, 0);
        if (g * 8 < K) {
            const int base = (ty * PW + tx) * stride;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = g * 8 + j;
                v[j] = k < K ? patch[lut[k] + base] : 0.f;
            }
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&out);
#pragma unroll
            for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        }
        const size_t pix = (static_cast<size_t>(n) * OH + oh) * OW + ow;
        *reinterpret_cast<uint4*>(y + pix * pitch + g * 8) = out;
    }
}

// ------------------------------------------------------------------ global average pool NHWC bf16 -> f32 [N][C]
// One block per image, 8-channel (128-bit) loads; the block's 256 threads split into (pixel lane, channel vector); per-thread
// sums run over pixels in a fixed stride order and the pixel lanes are combined in index order: bit-reproducible (the first
// version used atomicAdd across slices).  Only the initial kernel prediction still pools through here (12544 pixels x 64
// channels per image); the per-stage pools of the kernel predictor live in the fused chain (kpred_chain.cu).
// One thread-block CLUSTER of kGapCluster CTAs per sample: every CTA sums a contiguous slice of the pixels, the lanes are
// combined in index order inside the CTA, and rank 0 adds the kGapCluster partial vectors in rank order through distributed
// shared memory -- no workspace, no atomics, bit-reproducible, and 8x the memory-level parallelism of one CTA per sample.
constexpr int kGapCluster = 8;
__global__ void __cluster_dims__(kGapCluster, 1, 1)
gap_nhwc_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int HW, int pitch, int coff, int C) {
    extern __shared__ float gp[];                       // [lanes][nvec * 8] + [nvec * 8] (this CTA's partial vector)
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(cluster.block_rank());
    const int n = blockIdx.x / kGapCluster;
    const int nvec = (C + 7) / 8;
    const int lanes = blockDim.x / nvec;
    const int cv = threadIdx.x % nvec, pl = threadIdx.x / nvec;
    float* part = gp + lanes * nvec * 8;
    const int per = (HW + kGapCluster - 1) / kGapCluster;
    const int p0 = rank * per, p1 = min(HW, p0 + per);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (pl < lanes) {
        const __nv_bfloat16* base = x + static_cast<size_t>(n) * HW * pitch + coff + cv * 8;
        for (int p = p0 + pl; p < p1; p += lanes) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(p) * pitch));
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[2 * i] += __uint_as_float(w[i] << 16);
                acc[2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) gp[(pl * nvec + cv) * 8 + i] = acc[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nvec * 8; c += blockDim.x) {
        float t = 0.f;
        for (int l = 0; l < lanes; ++l) t += gp[l * nvec * 8 + c];
        part[c] = t;
    }
    cluster.sync();
    if (rank == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float t = 0.f;
            for (int r = 0; r < kGapCluster; ++r) t += cluster.map_shared_rank(part, r)[c];
            out[n * C + c] = t / static_cast<float>(HW);
        }
    }
    cluster.sync();                                     // the partial vectors stay alive until rank 0 has read them
}

// ------------------------------------------------------------------ kernel-vector update
// cubic convolution coefficients as torch's upsample_bicubic2d (A = -0.75, align_corners=False)
__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
    const float A = -0.75f;
    c[0] = cubic2(t + 1.f, A);
    c[1] = cubic1(t, A);
    c[2] = cubic1(1.f - t, A);
    c[3] = cubic2(2.f - t, A);
}
__device__ __forceinline__ float src_index(float scale, int dst) { return scale * (dst + 0.5f) - 0.5f; }

// out[b] = normalize( (pre ? pre[b] : 0) + bicubic_{ke->ko}(v[b]) ), one block per sample.
// predictor_withGAP.upscale_and_reshape (kbpn.py:335-341), KernelPredictorLikeIKC (:574-578) + KBlock (:391-392).
__global__ void kernel_update_kernel(const float* __restrict__ v, const float* __restrict__ pre,
                                     float* __restrict__ out, int ke, int ko, int normalize) {
    extern __shared__ float sh[];
    float* sv = sh;                    // ke*ke
    float* so = sh + ke * ke;          // ko*ko
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < ke * ke; i += blockDim.x) sv[i] = v[b * ke * ke + i];
    __syncthreads();
    const float scale = static_cast<float>(ke) / static_cast<float>(ko);
    float local = 0.f;
    for (int i = threadIdx.x; i < ko * ko; i += blockDim.x) {
        const int oy = i / ko, ox = i % ko;
        const float ry = src_index(scale, oy), rx = src_index(scale, ox);
        const int iy = static_cast<int>(floorf(ry)), ix = static_cast<int>(floorf(rx));
        float cy[4], cx[4];
        cubic_coeffs(ry - iy, cy);
        cubic_coeffs(rx - ix, cx);
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int yy = min(max(iy - 1 + a, 0), ke - 1);
            float row = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int xx = min(max(ix - 1 + c, 0), ke - 1);
                row += sv[yy * ke + xx] * cx[c];
            }
            acc += row * cy[a];
        }
        if (pre) acc += pre[b * ko * ko + i];
        so[i] = acc;
        local += acc;
    }
    __shared__ float red[32];
    const float total = block_sum_det(local, red);             // fixed summation order: bit-reproducible
    const float inv = normalize ? 1.f / total : 1.f;
    for (int i = threadIdx.x; i < ko * ko; i += blockDim.x) out[b * ko * ko + i] = so[i] * inv;
}

// out[b] = v[b] / sum(v[b])   (build_model.py:491-494)
__global__ void vec_normalize_kernel(const float* __restrict__ v, float* __restrict__ out, int L) {
    const int b = blockIdx.x;
    float local = 0.f;
    for (int i = threadIdx.x; i < L; i += blockDim.x) local += v[b * L + i];
    __shared__ float red[32];
    const float total = block_sum_det(local, red);
    for (int i = threadIdx.x; i < L; i += blockDim.x) out[b * L + i] = v[b * L + i] / total;
}

// broadcast a per-sample vector over a small NHWC bf16 image (spatially constant conditioning)
__global__ void broadcast_vec_kernel(const float* __restrict__ v, __nv_bfloat16* __restrict__ y, int N, int HW, int L,
                                     int pitch, int coff, int cwrite) {
    const size_t total = static_cast<size_t>(N) * HW * cwrite;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % cwrite);
        const size_t pix = i / cwrite;
        const int n = static_cast<int>(pix / HW);
        y[pix * pitch + coff + c] = __float2bfloat16(c < L ? v[n * L + c] : 0.f);
    }
}

// ------------------------------------------------------------------ per-sample depthwise blur, stride s, minus LR
// err[n,c,oh,ow] = sum_{r,s} x[n,c,oh*st+r-pad, ow*st+s-pad] * k[n,r,s] - lr[n,c,oh,ow]   (kbpn.py:395-405)
// block = one (n, c, tile of 8x32 outputs); the input patch and the kernel are staged in shared memory.
template <int KS, int ST>
__global__ void __launch_bounds__(256)
blur_ps_kernel(const float* __restrict__ x, const float* __restrict__ kvec, const float* __restrict__ lr,
               float* __restrict__ err, int C, int H, int W, int OH, int OW) {
    // register-tiled: every thread produces OPT horizontally adjacent outputs; per kernel row it loads the (OPT-1)*ST + KS
    // input values it needs with 128-bit shared-memory loads and reuses each kernel tap for OPT FMAs (the one-output-per-
    // thread form issued two shared loads per FMA and ran at 10 % of the fp32 peak)
    constexpr int OPT = ST == 1 ? 8 : 4;                     // outputs per thread
    constexpr int TXN = 16, TYN = 16;                        // thread grid -> tile of TYN x (TXN*OPT) outputs
    constexpr int TOH = TYN, TOW = TXN * OPT;
    constexpr int NIN = ((OPT - 1) * ST + KS + 3) / 4 * 4;   // input values per thread and row, padded to float4
    constexpr int PH = (TOH - 1) * ST + KS;
    constexpr int PW = ((TOW - 1) * ST + KS + 3 + 4) / 4 * 4;   // row pitch: multiple of 4 floats, covers the padded reads
    extern __shared__ __align__(16) float blur_sm[];
    float* sk = blur_sm;                                     // [KS*KS] (+ pad)
    float* sp = blur_sm + (KS * KS + 3) / 4 * 4;             // [PH][PW]
    const int nc = blockIdx.z;
    const int n = nc / C;
    const int oh0 = blockIdx.y * TOH, ow0 = blockIdx.x * TOW;
    const int pad = (KS - 1) / 2;
    const float* xp = x + static_cast<size_t>(nc) * H * W;
    for (int i = threadIdx.x; i < KS * KS; i += blockDim.x) sk[i] = kvec[n * KS * KS + i];
    const int ih0 = oh0 * ST - pad, iw0 = ow0 * ST - pad;
    for (int i = threadIdx.x; i < PH * PW; i += blockDim.x) {
        const int r = i / PW, c = i % PW;
        const int ih = ih0 + r, iw = iw0 + c;
        sp[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? xp[static_cast<size_t>(ih) * W + iw] : 0.f;
    }
    __syncthreads();
    const int tx = threadIdx.x % TXN, ty = threadIdx.x / TXN;
    float acc[OPT];
#pragma unroll
    for (int o = 0; o < OPT; ++o) acc[o] = 0.f;
    const float* prow = sp + (ty * ST) * PW + tx * OPT * ST;          // 16-byte aligned: OPT*ST is a multiple of 4
#pragma unroll 1
    for (int r = 0; r < KS; ++r) {
        float in[NIN];
#pragma unroll
        for (int v = 0; v < NIN / 4; ++v) {
            const float4 q = *reinterpret_cast<const float4*>(prow + r * PW + 4 * v);
            in[4 * v] = q.x; in[4 * v + 1] = q.y; in[4 * v + 2] = q.z; in[4 * v + 3] = q.w;
        }
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const float kv = sk[r * KS + s];
#pragma unroll
            for (int o = 0; o < OPT; ++o) acc[o] = fmaf(in[o * ST + s], kv, acc[o]);
        }
    }
    const int oh = oh0 + ty;
    if (oh < OH) {
#pragma unroll
        for (int o = 0; o < OPT; ++o) {
            const int ow = ow0 + tx * OPT + o;
            if (ow < OW) {
                const size_t oidx = (static_cast<size_t>(nc) * OH + oh) * OW + ow;
                err[oidx] = lr ? acc[o] - lr[oidx] : acc[o];
            }
        }
    }
}

// ------------------------------------------------------------------ bicubic x4 upsample, f32 planar (kbpn.py:70,113)
__global__ void bicubic_up_kernel(const float* __restrict__ x, float* __restrict__ y, int NC, int H, int W, int f) {
    const int OH = H * f, OW = W * f;
    const size_t total = static_cast<size_t>(NC) * OH * OW;
    const float scale = 1.f / f;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(i % OW);
        const int oy = static_cast<int>((i / OW) % OH);
        const int nc = static_cast<int>(i / (static_cast<size_t>(OW) * OH));
        const float ry = src_index(scale, oy), rx = src_index(scale, ox);
        const int iy = static_cast<int>(floorf(ry)), ix = static_cast<int>(floorf(rx));
        float cy[4], cx[4];
        cubic_coeffs(ry - iy, cy);
        cubic_coeffs(rx - ix, cx);
        const float* xp = x + static_cast<size_t>(nc) * H * W;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int yy = min(max(iy - 1 + a, 0), H - 1);
            float row = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int xx = min(max(ix - 1 + c, 0), W - 1);
                row += xp[static_cast<size_t>(yy) * W + xx] * cx[c];
            }
            acc += row * cy[a];
        }
        y[i] = acc;
    }
}

// ------------------------------------------------------------------ clip to [0,1] in place + per-(n,c) statistics
__global__ void clip_stats_kernel(float* __restrict__ x, double* __restrict__ part, int HW, int do_clip) {
    // grid: (slices, N*C). part[(nc*slices + slice)*2 + {0,1}] = this block's sum / sum of squares (fp64), reduced in a
    // fixed order: bit-reproducible statistics
    const int nc = blockIdx.y;
    float* xp = x + static_cast<size_t>(nc) * HW;
    double s = 0.0, ss = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        float v = xp[i];
        if (do_clip) {
            v = fminf(fmaxf(v, 0.f), 1.f);
            xp[i] = v;
        }
        s += v;
        ss += static_cast<double>(v) * v;
    }
    __shared__ double red[32];
    s = block_sum_det(s, red);
    ss = block_sum_det(ss, red);
    if (threadIdx.x == 0) {
        part[(static_cast<size_t>(nc) * gridDim.x + blockIdx.x) * 2] = s;
        part[(static_cast<size_t>(nc) * gridDim.x + blockIdx.x) * 2 + 1] = ss;
    }
}
__global__ void finish_stats_kernel(const double* __restrict__ part, float* __restrict__ mean,
                                    float* __restrict__ rstd, int NC, int HW, float eps, int slices) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NC) return;
    double sum = 0.0, sq = 0.0;
    for (int k = 0; k < slices; ++k) {                  // fixed order
        sum += part[(static_cast<size_t>(i) * slices + k) * 2];
        sq += part[(static_cast<size_t>(i) * slices + k) * 2 + 1];
    }
    const double m = sum / HW;
    double var = sq / HW - m * m;                       // biased variance (InstanceNorm2d)
    if (var < 0) var = 0;
    mean[i] = static_cast<float>(m);
    rstd[i] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
}

// ------------------------------------------------------------------ NHWC bf16 pooling / resampling (8-channel vectors)
__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 o;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return o;
}

// 3x3 stride-2 pad-1 max pool (extractors.py:119)
__global__ void maxpool3s2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                                  int W, int OH, int OW, int C, int xp, int xo, int yp, int yo) {
    const int G = C / 8;
    const size_t total = static_cast<size_t>(N) * OH * OW * G;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const size_t pix = i / G;
        const int ow = static_cast<int>(pix % OW);
        const int oh = static_cast<int>((pix / OW) % OH);
        const int n = static_cast<int>(pix / (static_cast<size_t>(OW) * OH));
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
        for (int r = 0; r < 3; ++r) {
            const int ih = oh * 2 + r - 1;
            if (ih < 0 || ih >= H) continue;
            for (int s = 0; s < 3; ++s) {
                const int iw = ow * 2 + s - 1;
                if (iw < 0 || iw >= W) continue;
                float f[8];
                unpack8(*reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(n) * H + ih) * W + iw) * xp + xo + g * 8), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
            }
        }
        *reinterpret_cast<uint4*>(y + pix * yp + yo + g * 8) = pack8(m);
    }
}

// adaptive average pool to SxS (pspnet.py:32): bin i covers [floor(i*H/S), ceil((i+1)*H/S))
__global__ void adaptive_avgpool_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N,
                                        int H, int W, int S, int C, int xp, int xo, int yp, int yo) {
    // grid: (S*S bins, N, channel slabs of 32 groups); block 256 = 32 channel groups (8 ch each) x 8 pixel lanes
    __shared__ float red[8][32][8];
    const int bin = blockIdx.x, n = blockIdx.y;
    const int g = blockIdx.z * 32 + (threadIdx.x & 31);
    const int pl = threadIdx.x >> 5;
    const int oh = bin / S, ow = bin % S;
    const int h0 = (oh * H) / S, h1 = ((oh + 1) * H + S - 1) / S;
    const int w0 = (ow * W) / S, w1 = ((ow + 1) * W + S - 1) / S;
    const int bw = w1 - w0, npix = (h1 - h0) * bw;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (g * 8 < C) {
        for (int pidx = pl; pidx < npix; pidx += 8) {
            const int ih = h0 + pidx / bw, iw = w0 + pidx % bw;
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(n) * H + ih) * W + iw) * xp + xo + g * 8), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[pl][threadIdx.x & 31][j] = acc[j];
    __syncthreads();
    if (pl == 0 && g * 8 < C) {
        const float inv = 1.f / static_cast<float>(npix);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float t = 0.f;
#pragma unroll
            for (int l = 0; l < 8; ++l) t += red[l][threadIdx.x & 31][j];
            acc[j] = t * inv;
        }
        *reinterpret_cast<uint4*>(y + ((static_cast<size_t>(n) * S + oh) * S + ow) * yp + yo + g * 8) = pack8(acc);
    }
}

__device__ __forceinline__ void bilinear_src(int dst, int in, int out, int align, int& i0, int& i1, float& l1) {
    float src;
    if (align) {
        src = out > 1 ? dst * (static_cast<float>(in - 1) / static_cast<float>(out - 1)) : 0.f;
    } else {
        src = (dst + 0.5f) * (static_cast<float>(in) / static_cast<float>(out)) - 0.5f;
        if (src < 0.f) src = 0.f;
    }
    i0 = static_cast<int>(src);
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 < in - 1 ? i0 + 1 : i0;
    l1 = src - i0;
}

// bilinear resize NHWC bf16 (F.interpolate mode='bilinear'; pspnet.py:39,56)
__global__ void bilinear_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int N, int H,
                                     int W, int OH, int OW, int C, int xp, int xo, int yp, int yo, int align) {
    const int G = C / 8;
    const size_t total = static_cast<size_t>(N) * OH * OW * G;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const size_t pix = i / G;
        const int ow = static_cast<int>(pix % OW);
        const int oh = static_cast<int>((pix / OW) % OH);
        const int n = static_cast<int>(pix / (static_cast<size_t>(OW) * OH));
        int y0, y1, x0, x1;
        float ly, lx;
        bilinear_src(oh, H, OH, align, y0, y1, ly);
        bilinear_src(ow, W, OW, align, x0, x1, lx);
        const __nv_bfloat16* base = x + static_cast<size_t>(n) * H * W * xp + xo + g * 8;
        float a[8], b[8], c[8], d[8], o[8];
        unpack8(*reinterpret_cast<const uint4*>(base + (static_cast<size_t>(y0) * W + x0) * xp), a);
        unpack8(*reinterpret_cast<const uint4*>(base + (static_cast<size_t>(y0) * W + x1) * xp), b);
        unpack8(*reinterpret_cast<const uint4*>(base + (static_cast<size_t>(y1) * W + x0) * xp), c);
        unpack8(*reinterpret_cast<const uint4*>(base + (static_cast<size_t>(y1) * W + x1) * xp), d);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            o[j] = (1.f - ly) * ((1.f - lx) * a[j] + lx * b[j]) + ly * ((1.f - lx) * c[j] + lx * d[j]);
        *reinterpret_cast<uint4*>(y + pix * yp + yo + g * 8) = pack8(o);
    }
}

// Exact x2 case of bilinear_nhwc_kernel (align_corners = false, OH = 2H, OW = 2W; PSPUpsample, pspnet.py:52-57): one thread
// owns an input pixel (n, i, j) and 8 channels, reads its 3x3 neighbourhood once (9 x 16 B instead of 16 x 16 B for the four
// outputs) and writes the 2x2 output block.  Source rows of output 2i are (i-1, i) with weight 0.75 on row i (at i = 0 the
// clamped source gives weight 0 on the second row), of output 2i+1 (i, min(i+1, H-1)) with weight 0.25: the same indices,
// weights and expression as bilinear_src / bilinear_nhwc_kernel, so both kernels produce identical bits.
__global__ void __launch_bounds__(256) bilinear_x2_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                               int H, int W, int C, int xp, int xo, int yp, int yo) {
    const int G = C >> 3;
    const int row = blockIdx.x;                       // n * H + i
    const int n = row / H, i = row - n * H;
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    if (e >= W * G) return;
    const int j = e / G, g = e - j * G;
    const int ra = i > 0 ? i - 1 : 0, rc = i < H - 1 ? i + 1 : i;
    const int ca = j > 0 ? j - 1 : 0, cc = j < W - 1 ? j + 1 : j;
    const float ly0 = i > 0 ? 0.75f : 0.f, lx0 = j > 0 ? 0.75f : 0.f;
    const __nv_bfloat16* base = x + static_cast<size_t>(n) * H * W * xp + xo + g * 8;
    const int rows[3] = {ra, i, rc}, cols[3] = {ca, j, cc};
    float v[3][3][8];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            unpack8(*reinterpret_cast<const uint4*>(base + (static_cast<size_t>(rows[r]) * W + cols[c]) * xp), v[r][c]);
    __nv_bfloat16* out = y + (static_cast<size_t>(n) * 2 * H + 2 * i) * (2 * static_cast<size_t>(W)) * yp + static_cast<size_t>(2 * j) * yp + yo + g * 8;
#pragma unroll
    for (int py = 0; py < 2; ++py) {
        const float ly = py ? 0.25f : ly0;
#pragma unroll
        for (int px = 0; px < 2; ++px) {
            const float lx = px ? 0.25f : lx0;
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
                o[k] = (1.f - ly) * ((1.f - lx) * v[py][px][k] + lx * v[py][px + 1][k]) +
                       ly * ((1.f - lx) * v[py + 1][px][k] + lx * v[py + 1][px + 1][k]);
            *reinterpret_cast<uint4*>(out + (static_cast<size_t>(py) * 2 * W + px) * yp) = pack8(o);
        }
    }
}

// y = [relu](base + bilinear(x)): branch fusion of HighResolutionModule.forward (hrnet_backbone.py:274-288)
__global__ void bilinear_add_nhwc_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ base,
                                         __nv_bfloat16* __restrict__ y, int N, int H, int W, int OH, int OW, int C, int xp,
                                         int xo, int bp, int bo, int yp, int yo, int align, int relu) {
    const int G = C / 8;
    const size_t total = static_cast<size_t>(N) * OH * OW * G;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const size_t pix = i / G;
        const int ow = static_cast<int>(pix % OW);
        const int oh = static_cast<int>((pix / OW) % OH);
        const int n = static_cast<int>(pix / (static_cast<size_t>(OW) * OH));
        int y0, y1, x0, x1;
        float ly, lx;
        bilinear_src(oh, H, OH, align, y0, y1, ly);
        bilinear_src(ow, W, OW, align, x0, x1, lx);
        const __nv_bfloat16* src = x + static_cast<size_t>(n) * H * W * xp + xo + g * 8;
        float a[8], b[8], c[8], d[8], o[8];
        unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(y0) * W + x0) * xp), a);
        unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(y0) * W + x1) * xp), b);
        unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(y1) * W + x0) * xp), c);
        unpack8(*reinterpret_cast<const uint4*>(src + (static_cast<size_t>(y1) * W + x1) * xp), d);
        unpack8(*reinterpret_cast<const uint4*>(base + pix * bp + bo + g * 8), o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j] += (1.f - ly) * ((1.f - lx) * a[j] + lx * b[j]) + ly * ((1.f - lx) * c[j] + lx * d[j]);
            if (relu) o[j] = fmaxf(o[j], 0.f);
        }
        *reinterpret_cast<uint4*>(y + pix * yp + yo + g * 8) = pack8(o);
    }
}

// SpatialGather_Module.forward for one object class (spatial_ocr_block.py:59-66): ctx[n][c] = sum_hw softmax_hw(logit)[hw] *
// feats[n][hw][c].  One block per (sample, 64-channel slab); the softmax statistics are recomputed per block (HW is small).
__global__ void softmax_gather_kernel(const float* __restrict__ logits, const __nv_bfloat16* __restrict__ feats,
                                      float* __restrict__ ctx, int HW, int C, int fp, int fo) {
    __shared__ float red[32];
    __shared__ float acc_s[4][64];
    const int n = blockIdx.y, c0 = blockIdx.x * 64;
    const float* lg = logits + static_cast<size_t>(n) * HW;
    float m = -INFINITY;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) m = fmaxf(m, lg[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int i = 1; i < (blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
    __syncthreads();
    float sum = 0.f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) sum += expf(lg[i] - m);
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) sum += red[i];
    // 256 threads = 4 pixel lanes x 64 channels
    const int c = threadIdx.x & 63, pl = threadIdx.x >> 6;
    float acc = 0.f;
    if (c0 + c < C) {
        const __nv_bfloat16* f = feats + static_cast<size_t>(n) * HW * fp + fo + c0 + c;
        for (int i = pl; i < HW; i += 4) acc += expf(lg[i] - m) * __bfloat162float(f[static_cast<size_t>(i) * fp]);
    }
    acc_s[pl][c] = acc;
    __syncthreads();
    if (pl == 0 && c0 + c < C) ctx[n * C + c0 + c] = (acc_s[0][c] + acc_s[1][c] + acc_s[2][c] + acc_s[3][c]) / sum;
}

// bilinear resize of planar f32 maps (aux head, align_corners=True; pspnet.py:122)
__global__ void bilinear_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int NC, int H, int W, int OH,
                                    int OW, int align, int sigmoid) {
    const size_t total = static_cast<size_t>(NC) * OH * OW;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ow = static_cast<int>(i % OW);
        const int oh = static_cast<int>((i / OW) % OH);
        const int nc = static_cast<int>(i / (static_cast<size_t>(OW) * OH));
        int y0, y1, x0, x1;
        float ly, lx;
        bilinear_src(oh, H, OH, align, y0, y1, ly);
        bilinear_src(ow, W, OW, align, x0, x1, lx);
        const float* b = x + static_cast<size_t>(nc) * H * W;
        float v = (1.f - ly) * ((1.f - lx) * b[y0 * W + x0] + lx * b[y0 * W + x1]) +
                  ly * ((1.f - lx) * b[y1 * W + x0] + lx * b[y1 * W + x1]);
        if (sigmoid) v = 1.f / (1.f + expf(-v));
        y[i] = v;
    }
}

}  // namespace csbsr

using namespace csbsr;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int csbsr_nchw_f32_to_nhwc_bf16(const float* x, void* y, int n, int c, int h, int w, int y_pitch,
                                           int y_coff, int cwrite, void* stream) {
    CSBSR_REQUIRE(x && y && n > 0 && c > 0 && h > 0 && w > 0 && cwrite >= c, "nchw_to_nhwc: bad arguments");
    const size_t total = static_cast<size_t>(n) * h * w * cwrite;
    if (cwrite % 8 == 0 && y_pitch % 8 == 0 && y_coff % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0)
        nchw_to_nhwc_vec_kernel<<<grid_for(total / 8, 256), 256, 0, STREAM(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n, c, h, w,
                                                                                     y_pitch, y_coff, cwrite);
    else
        nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n, c,
                                                                             h, w, y_pitch, y_coff, cwrite);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_patchify(const float* x, void* y, int n, int c, int h, int w, int oh, int ow, int r, int s,
                              int stride, int pad, int y_pitch, int cwrite, const float* mean, const float* rstd,
                              int clamp01, void* stream) {
    CSBSR_REQUIRE(x && y && n > 0 && cwrite % 8 == 0 && y_pitch % 8 == 0 && cwrite >= r * s * c && cwrite <= y_pitch,
                  "patchify: bad arguments (cwrite=%d, need >= %d)", cwrite, r * s * c);
    CSBSR_REQUIRE((mean == nullptr) == (rstd == nullptr), "patchify: mean/rstd must come together");
    CSBSR_REQUIRE(r * s * c <= 256 && r <= 255 && s <= 255 && c <= 255, "patchify: at most 256 gathered channels");
    const int ph = 7 * stride + r, pw = 31 * stride + s;
    const size_t smem = sizeof(float) * static_cast<size_t>(c) * ph * pw;
    CSBSR_REQUIRE(smem <= 48 * 1024, "patchify: patch of %zu bytes does not fit in shared memory", smem);
    dim3 grid((ow + 31) / 32, (oh + 7) / 8, n);
    patchify_kernel<<<grid, 256, smem, STREAM(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(y), n, c, h, w, oh, ow, r, s,
                                                        stride, pad, y_pitch, cwrite, mean, rstd, clamp01);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_gap_nhwc(const void* x, float* out, int n, int hw, int pitch, int coff, int c, void* stream) {
    CSBSR_REQUIRE(x && out && n > 0 && hw > 0 && c > 0 && c <= 1024, "gap_nhwc: bad arguments");
    // whole 8-channel vectors are read (channels up to the next multiple of 8 must exist inside the pixel's pitch)
    CSBSR_REQUIRE(pitch % 8 == 0 && coff % 8 == 0 && coff + (c + 7) / 8 * 8 <= pitch,
                  "gap_nhwc: pitch / offset must be multiples of 8 and cover the rounded-up channel window");
    const int nvec = (c + 7) / 8;
    const int threads = nvec <= 256 ? 256 : 1024;
    CSBSR_REQUIRE(nvec <= threads, "gap_nhwc: too many channels");
    const int lanes = threads / nvec;
    gap_nhwc_kernel<<<n * kGapCluster, threads, sizeof(float) * (lanes * nvec * 8 + nvec * 8), STREAM(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), out, hw, pitch, coff, c);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_kernel_update(const float* v, const float* pre, float* out, int b, int ke, int ko, int normalize,
                                   void* stream) {
    CSBSR_REQUIRE(v && out && b > 0 && ke > 0 && ko > 0, "kernel_update: bad arguments");
    kernel_update_kernel<<<b, 128, sizeof(float) * (ke * ke + ko * ko), STREAM(stream)>>>(v, pre, out, ke, ko, normalize);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_vec_normalize(const float* v, float* out, int b, int len, void* stream) {
    CSBSR_REQUIRE(v && out && b > 0 && len > 0, "vec_normalize: bad arguments");
    vec_normalize_kernel<<<b, 128, 0, STREAM(stream)>>>(v, out, len);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_broadcast_vec(const float* v, void* y, int n, int hw, int len, int y_pitch, int y_coff,
                                   int cwrite, void* stream) {
    CSBSR_REQUIRE(v && y && n > 0 && hw > 0 && cwrite >= len, "broadcast_vec: bad arguments");
    const size_t total = static_cast<size_t>(n) * hw * cwrite;
    broadcast_vec_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(v, reinterpret_cast<__nv_bfloat16*>(y), n, hw,
                                                                          len, y_pitch, y_coff, cwrite);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_blur_per_sample(const float* x, const float* kvec, const float* lr, float* err, int n, int c,
                                     int h, int w, int ksize, int stride, void* stream) {
    CSBSR_REQUIRE(x && kvec && err && n > 0 && c > 0, "blur_per_sample: bad arguments");
    CSBSR_REQUIRE(ksize == 21 && (stride == 4 || stride == 1), "blur_per_sample: only ksize=21, stride in {1,4}");
    const int pad = (ksize - 1) / 2;
    const int oh = (h + 2 * pad - ksize) / stride + 1, ow = (w + 2 * pad - ksize) / stride + 1;
    // tile = 16 x (16 * outputs-per-thread) outputs; dynamic shared memory = kernel + input patch (see blur_ps_kernel)
    auto smem_of = [](int st, int opt) {
        const int tow = 16 * opt, ph = 15 * st + 21, pw = ((tow - 1) * st + 21 + 3 + 4) / 4 * 4;
        return static_cast<int>(sizeof(float)) * ((21 * 21 + 3) / 4 * 4 + ph * pw);
    };
    if (stride == 4) {
        const int smem = smem_of(4, 4);
        static bool attr4 = false;
        if (!attr4) {
            CSBSR_CHECK_CUDA(cudaFuncSetAttribute(blur_ps_kernel<21, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr4 = true;
        }
        dim3 grid((ow + 63) / 64, (oh + 15) / 16, n * c);
        blur_ps_kernel<21, 4><<<grid, 256, smem, STREAM(stream)>>>(x, kvec, lr, err, c, h, w, oh, ow);
    } else {
        const int smem = smem_of(1, 8);
        dim3 grid((ow + 127) / 128, (oh + 15) / 16, n * c);
        blur_ps_kernel<21, 1><<<grid, 256, smem, STREAM(stream)>>>(x, kvec, lr, err, c, h, w, oh, ow);
    }
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_bicubic_upsample(const float* x, float* y, int nc, int h, int w, int factor, void* stream) {
    CSBSR_REQUIRE(x && y && nc > 0 && factor >= 1, "bicubic_upsample: bad arguments");
    const size_t total = static_cast<size_t>(nc) * h * w * factor * factor;
    bicubic_up_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(x, y, nc, h, w, factor);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static constexpr int kInstnormMaxSlices = 64;
extern "C" size_t csbsr_instnorm_workspace_bytes(int nc) {
    return sizeof(double) * 2 * kInstnormMaxSlices * static_cast<size_t>(nc);
}

extern "C" int csbsr_clip_instnorm_stats(float* x, float* mean, float* rstd, void* workspace, int nc, int hw,
                                         int do_clip, float eps, void* stream) {
    CSBSR_REQUIRE(x && mean && rstd && workspace && nc > 0 && hw > 0, "clip_instnorm_stats: bad arguments");
    double* part = reinterpret_cast<double*>(workspace);
    int slices = (hw + 256 * 16 - 1) / (256 * 16);
    if (slices < 1) slices = 1;
    if (slices > kInstnormMaxSlices) slices = kInstnormMaxSlices;
    dim3 grid(slices, nc);
    clip_stats_kernel<<<grid, 256, 0, STREAM(stream)>>>(x, part, hw, do_clip);
    finish_stats_kernel<<<(nc + 127) / 128, 128, 0, STREAM(stream)>>>(part, mean, rstd, nc, hw, eps, slices);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_maxpool3s2_nhwc(const void* x, void* y, int n, int h, int w, int c, int x_pitch, int x_coff,
                                     int y_pitch, int y_coff, void* stream) {
    CSBSR_REQUIRE(x && y && c % 8 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0 && x_coff % 8 == 0 && y_coff % 8 == 0,
                  "maxpool: channel counts/offsets must be multiples of 8");
    const int oh = (h + 2 - 3) / 2 + 1, ow = (w + 2 - 3) / 2 + 1;
    const size_t total = static_cast<size_t>(n) * oh * ow * (c / 8);
    maxpool3s2_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                       reinterpret_cast<__nv_bfloat16*>(y), n, h, w, oh,
                                                                       ow, c, x_pitch, x_coff, y_pitch, y_coff);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_adaptive_avgpool_nhwc(const void* x, void* y, int n, int h, int w, int s, int c, int x_pitch,
                                           int x_coff, int y_pitch, int y_coff, void* stream) {
    CSBSR_REQUIRE(x && y && s > 0 && c % 8 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0 && x_coff % 8 == 0 &&
                      y_coff % 8 == 0,
                  "adaptive_avgpool: bad arguments");
    dim3 grid(s * s, n, (c / 8 + 31) / 32);
    adaptive_avgpool_kernel<<<grid, 256, 0, STREAM(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), n, h, w, s, c, x_pitch, x_coff,
        y_pitch, y_coff);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_bilinear_nhwc(const void* x, void* y, int n, int h, int w, int oh, int ow, int c, int x_pitch,
                                   int x_coff, int y_pitch, int y_coff, int align_corners, void* stream) {
    CSBSR_REQUIRE(x && y && c % 8 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0 && x_coff % 8 == 0 && y_coff % 8 == 0,
                  "bilinear_nhwc: channel counts/offsets must be multiples of 8");
    const size_t total = static_cast<size_t>(n) * oh * ow * (c / 8);
    if (!align_corners && oh == 2 * h && ow == 2 * w && h > 1 && w > 1) {
        const dim3 grid(static_cast<unsigned>(n) * h, (static_cast<unsigned>(w) * (c / 8) + 255) / 256);
        bilinear_x2_nhwc_kernel<<<grid, 256, 0, STREAM(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                                 reinterpret_cast<__nv_bfloat16*>(y), h, w, c, x_pitch, x_coff,
                                                                 y_pitch, y_coff);
        CSBSR_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    bilinear_nhwc_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y), n, h, w, oh, ow, c, x_pitch,
        x_coff, y_pitch, y_coff, align_corners);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_bilinear_f32(const float* x, float* y, int nc, int h, int w, int oh, int ow, int align_corners,
                                  void* stream) {
    CSBSR_REQUIRE(x && y && nc > 0, "bilinear_f32: bad arguments");
    const size_t total = static_cast<size_t>(nc) * oh * ow;
    bilinear_f32_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(x, y, nc, h, w, oh, ow, align_corners, 0);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_bilinear_f32_sigmoid(const float* x, float* y, int nc, int h, int w, int oh, int ow,
                                          int align_corners, void* stream) {
    CSBSR_REQUIRE(x && y && nc > 0, "bilinear_f32_sigmoid: bad arguments");
    const size_t total = static_cast<size_t>(nc) * oh * ow;
    bilinear_f32_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(x, y, nc, h, w, oh, ow, align_corners, 1);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_bilinear_add_nhwc(const void* x, const void* base, void* y, int n, int h, int w, int oh, int ow, int c,
                                       int x_pitch, int x_coff, int b_pitch, int b_coff, int y_pitch, int y_coff,
                                       int align_corners, int relu, void* stream) {
    CSBSR_REQUIRE(x && base && y && c % 8 == 0 && x_pitch % 8 == 0 && y_pitch % 8 == 0 && b_pitch % 8 == 0 &&
                      x_coff % 8 == 0 && y_coff % 8 == 0 && b_coff % 8 == 0,
                  "bilinear_add_nhwc: channel counts/offsets must be multiples of 8");
    const size_t total = static_cast<size_t>(n) * oh * ow * (c / 8);
    bilinear_add_nhwc_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(base),
        reinterpret_cast<__nv_bfloat16*>(y), n, h, w, oh, ow, c, x_pitch, x_coff, b_pitch, b_coff, y_pitch, y_coff,
        align_corners, relu);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_softmax_gather(const float* logits, const void* feats, float* ctx, int n, int hw, int c, int f_pitch,
                                    int f_coff, void* stream) {
    CSBSR_REQUIRE(logits && feats && ctx && n > 0 && hw > 0 && c > 0, "softmax_gather: bad arguments");
    softmax_gather_kernel<<<dim3((c + 63) / 64, n), 256, 0, STREAM(stream)>>>(
        logits, reinterpret_cast<const __nv_bfloat16*>(feats), ctx, hw, c, f_pitch, f_coff);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// 3x3 conv with very few output channels (sr_reconst / output_conv, 128 -> 3: kbpn.py:361, :68) as "tap expansion":
// a 1x1 GEMM z[q][t*cp + c] = sum_ci x[q][ci] * w[c][ci][t] (cp = co rounded up to 4) on the tensor cores (N = 9*co instead of nine N = co GEMMs
// that each pay the full 128-row A-operand read), then this gather: out[p][c] = sum_t z[p + d_t][t*cp + c] (zero outside
// the image = the conv zero padding), accumulated into fp32 planar windows.
namespace csbsr {
template <int CO>
__global__ void tap_gather_kernel(const __nv_bfloat16* __restrict__ z, int z_pitch, float* out, int out_pitch,
                                  int out_coff, const float* r32, int r_pitch, int r_coff, int n, int h, int w) {
    constexpr int CP = (CO + 3) / 4 * 4;            // channels of one tap are padded to 4 -> 8-byte aligned uint2 loads
    const size_t total = static_cast<size_t>(n) * h * w;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % w);
        const int y = static_cast<int>((i / w) % h);
        const int img = static_cast<int>(i / (static_cast<size_t>(w) * h));
        float acc[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[c] = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
            if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
            const uint2* zp = reinterpret_cast<const uint2*>(z + ((static_cast<size_t>(img) * h + yy) * w + xx) * z_pitch + t * CP);
#pragma unroll
            for (int v = 0; v < CP / 4; ++v) {
                const uint2 q = __ldg(zp + v);
                acc[4 * v + 0] += __uint_as_float(q.x << 16);
                acc[4 * v + 1] += __uint_as_float(q.x & 0xFFFF0000u);
                acc[4 * v + 2] += __uint_as_float(q.y << 16);
                acc[4 * v + 3] += __uint_as_float(q.y & 0xFFFF0000u);
            }
        }
        const size_t plane = static_cast<size_t>(h) * w, pix = static_cast<size_t>(y) * w + x;
        float rv[CO];
#pragma unroll
        for (int c = 0; c < CO; ++c)
            rv[c] = r32 ? __ldg(r32 + (static_cast<size_t>(img) * r_pitch + r_coff + c) * plane + pix) : 0.f;
#pragma unroll
        for (int c = 0; c < CO; ++c) out[(static_cast<size_t>(img) * out_pitch + out_coff + c) * plane + pix] = acc[c] + rv[c];
    }
}
}  // namespace csbsr

extern "C" int csbsr_tap_gather3x3(const void* z, int z_pitch, int z_coff, float* out, int out_pitch, int out_coff,
                                   const float* r32, int r_pitch, int r_coff, int n, int h, int w, int co, void* stream) {
    CSBSR_REQUIRE(z && out && n > 0 && h > 0 && w > 0, "tap_gather3x3: bad arguments");
    CSBSR_REQUIRE(co == 3 || co == 6 || co == 9 || co == 12, "tap_gather3x3: co=%d must be 3, 6, 9 or 12", co);
    CSBSR_REQUIRE(z_pitch % 4 == 0 && z_coff % 4 == 0, "tap_gather3x3: z pitch / offset must be multiples of 4 channels");
    const __nv_bfloat16* zp = reinterpret_cast<const __nv_bfloat16*>(z) + z_coff;
    const size_t total = static_cast<size_t>(n) * h * w;
    const int grid = grid_for(total, 256);
#define CSBSR_TG(CO)                                                                                              \
    tap_gather_kernel<CO><<<grid, 256, 0, STREAM(stream)>>>(zp, z_pitch, out, out_pitch, out_coff, r32, r_pitch, \
                                                             r_coff, n, h, w)
    if (co == 3) CSBSR_TG(3);
    else if (co == 6) CSBSR_TG(6);
    else if (co == 9) CSBSR_TG(9);
    else CSBSR_TG(12);
#undef CSBSR_TG
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
