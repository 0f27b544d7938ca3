// On-the-fly synthetic degradation: anisotropic Gaussian blur-kernel synthesis, per-sample depthwise blur,
// x4 antialiased-bicubic downsample.
//   D1  GaussianBlur.make / get_deterioration   model/data/blur/blur.py:128-179  (fp64 -> fp32)
//   D2  conv_kernel2d                           model/data/blur/blur.py:182-200  (csbsr_blur_per_sample, stride 1)
//   D3  FactorResize -> torchvision Resize(BICUBIC) = antialiased bicubic (a = -0.5), model/data/transforms/transforms.py:516-531
#include "common.cuh"
#include "../../include/csbsr_b200.h"

namespace csbsr {

// one block per sample; params[b] = (theta [rad], sigma_x, sigma_y) in fp64
__global__ void kernel_synth_kernel(const double* __restrict__ params, float* __restrict__ out, int ks) {
    extern __shared__ double sv[];
    const int b = blockIdx.x;
    const double theta = params[b * 3], sx = params[b * 3 + 1], sy = params[b * 3 + 2];
    const double ct = cos(theta), st = sin(theta);
    const double ct2 = ct * ct, st2 = st * st;
    const double sx2 = 2.0 * (sx * sx), sy2 = 2.0 * (sy * sy);
    const double a = ct2 / sx2 + st2 / sy2;
    const double bb = st * ct * (1.0 / sy2 - 1.0 / sx2);
    const double c = st2 / sx2 + ct2 / sy2;
    const int r = ks / 2;
    double local = 0.0;
    for (int i = threadIdx.x; i < ks * ks; i += blockDim.x) {
        const double y = static_cast<double>(i / ks - r), x = static_cast<double>(i % ks - r);
        const double v = exp(-(a * (x * x) + 2.0 * bb * x * y + c * (y * y)));
        sv[i] = v;
        local += v;
    }
    __shared__ double red[32];
    const double total = block_sum_det(local, red);            // fixed summation order: bit-reproducible kernels
    for (int i = threadIdx.x; i < ks * ks; i += blockDim.x) out[b * ks * ks + i] = static_cast<float>(sv[i] / total);
}

// antialiased bicubic weights as aten's upsample_bicubic2d_aa (a = -0.5, support = 2*scale for scale >= 1)
__device__ __forceinline__ float aa_cubic(float x) {
    const float a = -0.5f;
    x = fabsf(x);
    if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
    if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * a;
    return 0.f;
}
__device__ __forceinline__ void aa_span(int o, int in, float scale, int& xmin, int& xsize) {
    const float support = 2.f * scale;
    const float center = scale * (o + 0.5f);
    xmin = max(static_cast<int>(center - support + 0.5f), 0);
    xsize = min(static_cast<int>(center + support + 0.5f), in) - xmin;
}

// one thread per output pixel; the <= (4f+1)^2 taps are separable (wy x wx), each axis normalised to sum 1
__global__ void resize_aa_kernel(const float* __restrict__ x, float* __restrict__ y, int NC, int H, int W, int OH,
                                 int OW, int clamp01) {
    const size_t total = static_cast<size_t>(NC) * OH * OW;
    const float sh = static_cast<float>(H) / OH, sw = static_cast<float>(W) / OW;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(i % OW);
        const int oy = static_cast<int>((i / OW) % OH);
        const int nc = static_cast<int>(i / (static_cast<size_t>(OW) * OH));
        int y0, ny, x0, nx;
        aa_span(oy, H, sh, y0, ny);
        aa_span(ox, W, sw, x0, nx);
        const float cy = sh * (oy + 0.5f), cx = sw * (ox + 0.5f);
        const float ish = 1.f / sh, isw = 1.f / sw;
        float wxs = 0.f, wys = 0.f;
        for (int k = 0; k < nx; ++k) wxs += aa_cubic((k + x0 - cx + 0.5f) * isw);
        for (int k = 0; k < ny; ++k) wys += aa_cubic((k + y0 - cy + 0.5f) * ish);
        const float* xp = x + static_cast<size_t>(nc) * H * W;
        float acc = 0.f;
        for (int r = 0; r < ny; ++r) {
            const float wy = aa_cubic((r + y0 - cy + 0.5f) * ish) / wys;
            float row = 0.f;
            const float* xr = xp + static_cast<size_t>(y0 + r) * W + x0;
            for (int k = 0; k < nx; ++k) row += xr[k] * (aa_cubic((k + x0 - cx + 0.5f) * isw) / wxs);
            acc += wy * row;
        }
        if (clamp01) acc = fminf(fmaxf(acc, 0.f), 1.f);
        y[i] = acc;
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Fused degradation (factor 4): blur and antialiased bicubic are both linear, so lr = K (*) hr with the composed kernel
//   K[u][v] = sum_{a,b} k[a][b] * wy[u-a] * wx[v-b]      (36 x 36 taps, stride 4, window origin (4i-16, 4j-16))
// -- 1296 MACs per LR pixel instead of 441 per HR pixel (5.4x fewer FLOPs) and no `blurred` [B,3,H,W] round trip through HBM.
// The bicubic weight vectors of the first / last two output rows and columns are shorter and renormalised (aten's
// upsample_bicubic2d_aa: 10 / 14 taps), so there are 5 row classes x 5 column classes = 25 composed kernels per sample; the
// zero padding of the blur (blur.py:190) is the zero fill of the staged window, exactly as in the two-step form.
constexpr int kKS = 21, kKC = 36, kKC2 = kKC * kKC, kNCls = 5;

__device__ __forceinline__ int aa_class(int o, int n) { return o < 2 ? o : (o >= n - 2 ? 5 - (n - o) : 2); }

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1)
__device__ __forceinline__ void philox4x32_10(unsigned int c[4], unsigned int k0, unsigned int k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned int n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// throughput mode of the degradation: the per-sample draws of blur.py:129 / :170-179 made on the device.
// params[i] = (theta ~ U(theta_lo, theta_hi) [rad], sigma_x, sigma_y ~ U(sig_lo, sig_hi)); sample index = offset + i
__global__ void philox_params_kernel(double* __restrict__ params, int n, unsigned long long seed, unsigned long long offset,
                                     double theta_lo, double theta_hi, double sig_lo, double sig_hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long idx = offset + i;
    unsigned int c[4] = {static_cast<unsigned int>(idx), static_cast<unsigned int>(idx >> 32), 0u, 0u};
    philox4x32_10(c, static_cast<unsigned int>(seed), static_cast<unsigned int>(seed >> 32));
    const double u0 = c[0] * (1.0 / 4294967296.0), u1 = c[1] * (1.0 / 4294967296.0), u2 = c[2] * (1.0 / 4294967296.0);
    // separate multiply and add (no FMA contraction): bit-equal to the numpy restatement of the oracle
    params[i * 3] = __dadd_rn(theta_lo, __dmul_rn(theta_hi - theta_lo, u0));
    params[i * 3 + 1] = __dadd_rn(sig_lo, __dmul_rn(sig_hi - sig_lo, u1));
    params[i * 3 + 2] = __dadd_rn(sig_lo, __dmul_rn(sig_hi - sig_lo, u2));
}

// 16-slot antialiased-bicubic weight vector of output index o (slot s <-> input index 4*o - 6 + s), scale 4
__device__ void aa_weights16(int o, int in, float* w16) {
    int x0, nx;
    aa_span(o, in, 4.f, x0, nx);
    const float c = 4.f * (o + 0.5f);
    float ws = 0.f;
    for (int k = 0; k < nx; ++k) ws += aa_cubic((k + x0 - c + 0.5f) * 0.25f);
    for (int s = 0; s < 16; ++s) w16[s] = 0.f;
    for (int k = 0; k < nx; ++k) w16[x0 + k - (4 * o - 6)] = aa_cubic((k + x0 - c + 0.5f) * 0.25f) / ws;
}

// one block per (sample, row class, column class): blur kernel (as kernel_synth_kernel) and one composed 36x36 kernel
__global__ void __launch_bounds__(256)
degrade_prep_kernel(const double* __restrict__ params, float* __restrict__ kernels, float* __restrict__ k25, int H, int W,
                    int OH, int OW) {
    __shared__ double sv[kKS * kKS];
    __shared__ float kf[kKS * kKS];
    __shared__ float wy[16], wx[16];
    __shared__ double T[kKC][kKS];                 // k (*)_y wy
    __shared__ double red[32];
    const int b = blockIdx.x, set = blockIdx.y, cy = set / kNCls, cx = set % kNCls;
    const double theta = params[b * 3], sx = params[b * 3 + 1], sy = params[b * 3 + 2];
    const double ct = cos(theta), st = sin(theta);
    const double ct2 = ct * ct, st2 = st * st;
    const double sx2 = 2.0 * (sx * sx), sy2 = 2.0 * (sy * sy);
    const double a = ct2 / sx2 + st2 / sy2;
    const double bb = st * ct * (1.0 / sy2 - 1.0 / sx2);
    const double c = st2 / sx2 + ct2 / sy2;
    double local = 0.0;
    for (int i = threadIdx.x; i < kKS * kKS; i += blockDim.x) {
        const double y = static_cast<double>(i / kKS - kKS / 2), x = static_cast<double>(i % kKS - kKS / 2);
        const double v = exp(-(a * (x * x) + 2.0 * bb * x * y + c * (y * y)));
        sv[i] = v;
        local += v;
    }
    const double total = block_sum_det(local, red);
    for (int i = threadIdx.x; i < kKS * kKS; i += blockDim.x) {
        const float v = static_cast<float>(sv[i] / total);
        kf[i] = v;
        if (set == 0) kernels[b * kKS * kKS + i] = v;
    }
    if (threadIdx.x < 2) {
        const int isx = threadIdx.x, cls = isx ? cx : cy;
        const int n = isx ? OW : OH, in = isx ? W : H;
        const int o = cls < 2 ? cls : (cls == 2 ? 2 : n - (5 - cls));
        aa_weights16(o, in, isx ? wx : wy);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kKC * kKS; i += blockDim.x) {
        const int u = i / kKS, bcol = i % kKS;
        double acc = 0.0;
        for (int aa = max(0, u - 15); aa <= min(kKS - 1, u); ++aa)
            acc += static_cast<double>(kf[aa * kKS + bcol]) * static_cast<double>(wy[u - aa]);
        T[u][bcol] = acc;
    }
    __syncthreads();
    float* out = k25 + (static_cast<size_t>(b) * kNCls * kNCls + set) * kKC2;
    for (int i = threadIdx.x; i < kKC2; i += blockDim.x) {
        const int u = i / kKC, v = i % kKC;
        double acc = 0.0;
        for (int bcol = max(0, v - 15); bcol <= min(kKS - 1, v); ++bcol) acc += T[u][bcol] * static_cast<double>(wx[v - bcol]);
        out[i] = static_cast<float>(acc);
    }
}

// interior outputs (row / column class 2).  Block = 128 threads = 8 x 16 (14 used) threads, each 2 x 4 outputs -> tile of
// 16 x 56 LR pixels; the 96 x 256 HR window is staged in shared memory with an XOR swizzle on the float4 column index
// (thread tx reads float4s 4*tx .. 4*tx+11: without it 8 neighbouring lanes hit the same two bank groups), and every input
// row feeds both output rows of the thread (kernel rows rho and rho-4), every loaded value up to 4 x 2 FMAs.
constexpr int kDT_OH = 16, kDT_OW = 56, kDT_PH = (kDT_OH - 1) * 4 + kKC, kDT_PW = 256;
__device__ __forceinline__ int dswz(int q) { return q ^ ((q >> 3) & 3); }

__global__ void __launch_bounds__(128, 2)
degrade_main_kernel(const float* __restrict__ hr, const float* __restrict__ k25, float* __restrict__ lr, int C, int H, int W,
                    int OH, int OW, int clamp01) {
    extern __shared__ __align__(16) float dsm[];
    float* sk = dsm;                       // [36*36]
    float* sp = dsm + kKC2;                // [96][256], float4 columns swizzled
    const int nc = blockIdx.z, n = nc / C;
    const int oh0 = blockIdx.y * kDT_OH, ow0 = blockIdx.x * kDT_OW;
    const int ih0 = 4 * oh0 - 16, iw0 = 4 * ow0 - 16;
    const float* xp = hr + static_cast<size_t>(nc) * H * W;
    const float* kk = k25 + (static_cast<size_t>(n) * kNCls * kNCls + 12) * kKC2;      // class (2, 2)
    for (int i = threadIdx.x; i < kKC2 / 4; i += blockDim.x)
        reinterpret_cast<float4*>(sk)[i] = reinterpret_cast<const float4*>(kk)[i];
    // 16-byte cp.async with zero fill outside the image (src-size 0): all 48 copies of a thread in flight at once
    const unsigned sp_s = static_cast<unsigned>(__cvta_generic_to_shared(sp));
#pragma unroll 8
    for (int i = threadIdx.x; i < kDT_PH * (kDT_PW / 4); i += 128) {
        const int r = i / (kDT_PW / 4), q = i % (kDT_PW / 4);
        const int ih = ih0 + r, iw = iw0 + 4 * q;
        const bool in = ih >= 0 && ih < H && iw >= 0 && iw + 3 < W;
        const float* src = in ? xp + static_cast<size_t>(ih) * W + iw : xp;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sp_s + 4u * (r * kDT_PW + 4 * dswz(q))), "l"(src),
                     "r"(in ? 16 : 0));
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    if (tx >= 14) return;
    int off[12];
#pragma unroll
    for (int v = 0; v < 12; ++v) off[v] = 4 * dswz(4 * tx + v);
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
    const float* prow = sp + (ty * 8) * kDT_PW;
#pragma unroll 1
    for (int rho = 0; rho < kKC + 4; ++rho) {
        float in[48];
#pragma unroll
        for (int v = 0; v < 12; ++v) {
            const float4 q = *reinterpret_cast<const float4*>(prow + rho * kDT_PW + off[v]);
            in[4 * v] = q.x; in[4 * v + 1] = q.y; in[4 * v + 2] = q.z; in[4 * v + 3] = q.w;
        }
        if (rho < kKC) {
            const float4* wr = reinterpret_cast<const float4*>(sk + rho * kKC);
#pragma unroll
            for (int s4 = 0; s4 < 9; ++s4) {
                const float4 w = wr[s4];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    acc0[o] = fmaf(in[4 * o + 4 * s4], w.x, acc0[o]);
                    acc0[o] = fmaf(in[4 * o + 4 * s4 + 1], w.y, acc0[o]);
                    acc0[o] = fmaf(in[4 * o + 4 * s4 + 2], w.z, acc0[o]);
                    acc0[o] = fmaf(in[4 * o + 4 * s4 + 3], w.w, acc0[o]);
                }
            }
        }
        if (rho >= 4) {
            const float4* wr = reinterpret_cast<const float4*>(sk + (rho - 4) * kKC);
#pragma unroll
            for (int s4 = 0; s4 < 9; ++s4) {
                const float4 w = wr[s4];
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    acc1[o] = fmaf(in[4 * o + 4 * s4], w.x, acc1[o]);
                    acc1[o] = fmaf(in[4 * o + 4 * s4 + 1], w.y, acc1[o]);
                    acc1[o] = fmaf(in[4 * o + 4 * s4 + 2], w.z, acc1[o]);
                    acc1[o] = fmaf(in[4 * o + 4 * s4 + 3], w.w, acc1[o]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int oh = oh0 + 2 * ty + r;
        if (oh < 2 || oh >= OH - 2) continue;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int ow = ow0 + 4 * tx + o;
            if (ow < 2 || ow >= OW - 2) continue;
            float v = r ? acc1[o] : acc0[o];
            if (clamp01) v = fminf(fmaxf(v, 0.f), 1.f);
            lr[(static_cast<size_t>(nc) * OH + oh) * OW + ow] = v;
        }
    }
}

// the border ring (first / last two output rows and columns): one block per (plane, side line), the five composed kernels the
// line can need staged in shared memory, one warp per output (the 1296 taps spread over the lanes), inputs straight from global memory (L1/L2)
__global__ void __launch_bounds__(256)
degrade_border_kernel(const float* __restrict__ hr, const float* __restrict__ k25, float* __restrict__ lr, int C, int H, int W,
                      int OH, int OW, int clamp01) {
    __shared__ __align__(16) float sw[kNCls][kKC2];
    const int nc = blockIdx.y, n = nc / C;
    const int side = blockIdx.x;                       // 0..3: rows 0, 1, OH-2, OH-1 ; 4..7: columns 0, 1, OW-2, OW-1
    const bool is_row = side < 4;
    const int line = (side & 3) < 2 ? (side & 3) : (is_row ? OH : OW) - (4 - (side & 3));
    const int lcls = aa_class(line, is_row ? OH : OW);
    const float* kb = k25 + static_cast<size_t>(n) * kNCls * kNCls * kKC2;
    for (int i = threadIdx.x; i < kNCls * kKC2; i += blockDim.x) {
        const int s = i / kKC2, e = i % kKC2;
        sw[s][e] = kb[(is_row ? lcls * kNCls + s : s * kNCls + lcls) * kKC2 + e];
    }
    __syncthreads();
    const float* xp = hr + static_cast<size_t>(nc) * H * W;
    const int len = is_row ? OW : OH;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int t = wid; t < len; t += nw) {                            // one warp per output: 1296 taps over 32 lanes
        const int oh = is_row ? line : t, ow = is_row ? t : line;
        if (!is_row && (oh < 2 || oh >= OH - 2)) continue;          // corners belong to the row lines
        const float* w = sw[aa_class(is_row ? ow : oh, is_row ? OW : OH)];
        const int ih0 = 4 * oh - 16, iw0 = 4 * ow - 16;
        float acc = 0.f;
#pragma unroll 4
        for (int e = lane; e < kKC2; e += 32) {
            const int u = e / kKC, v = e - u * kKC;
            const int ih = ih0 + u, iw = iw0 + v;
            if (ih >= 0 && ih < H && iw >= 0 && iw < W) acc = fmaf(xp[static_cast<size_t>(ih) * W + iw], w[e], acc);
        }
        acc = warp_sum(acc);
        if (clamp01) acc = fminf(fmaxf(acc, 0.f), 1.f);
        if (lane == 0) lr[(static_cast<size_t>(nc) * OH + oh) * OW + ow] = acc;
    }
}

}  // namespace csbsr

using namespace csbsr;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int csbsr_blur_kernel_synth(const double* params, float* kernels, int b, int ksize, void* stream) {
    CSBSR_REQUIRE(params && kernels && b > 0 && ksize > 0 && (ksize & 1) && ksize <= 63, "blur_kernel_synth: bad arguments");
    kernel_synth_kernel<<<b, 128, sizeof(double) * ksize * ksize, STREAM(stream)>>>(params, kernels, ksize);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_resize_bicubic_aa(const float* x, float* y, int nc, int h, int w, int oh, int ow, int clamp01,
                                       void* stream) {
    CSBSR_REQUIRE(x && y && nc > 0 && oh > 0 && ow > 0 && oh <= h && ow <= w, "resize_bicubic_aa: downscale only");
    const size_t total = static_cast<size_t>(nc) * oh * ow;
    size_t blocks = (total + 127) / 128;
    if (blocks > static_cast<size_t>(num_sms()) * 16) blocks = static_cast<size_t>(num_sms()) * 16;
    resize_aa_kernel<<<static_cast<int>(blocks), 128, 0, STREAM(stream)>>>(x, y, nc, h, w, oh, ow, clamp01);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// degrade = synth + blur + resize; `blurred` is caller-provided scratch [b,c,h,w] (also an output for parity checks)
extern "C" int csbsr_degrade(const float* hr, const double* params, float* kernels, float* blurred, float* lr, int b,
                             int c, int h, int w, int ksize, int factor, int clamp01, void* stream) {
    CSBSR_REQUIRE(hr && params && kernels && blurred && lr, "degrade: null pointer");
    CSBSR_REQUIRE(h % factor == 0 && w % factor == 0, "degrade: size must be a multiple of the factor");
    int rc = csbsr_blur_kernel_synth(params, kernels, b, ksize, stream);
    if (rc) return rc;
    rc = csbsr_blur_per_sample(hr, kernels, nullptr, blurred, b, c, h, w, ksize, 1, stream);
    if (rc) return rc;
    return csbsr_resize_bicubic_aa(blurred, lr, b * c, h, w, h / factor, w / factor, clamp01, stream);
}

extern "C" size_t csbsr_degrade_workspace_bytes(int b) {
    return sizeof(float) * static_cast<size_t>(b) * kNCls * kNCls * kKC2;
}

// fused form of csbsr_degrade (no `blurred` tensor): kernel synthesis + 25 composed kernels, interior tiles, border ring
extern "C" int csbsr_degrade_fused(const float* hr, const double* params, float* kernels, float* lr, void* workspace,
                                   size_t workspace_bytes, int b, int c, int h, int w, int ksize, int factor, int clamp01,
                                   void* stream) {
    CSBSR_REQUIRE(hr && params && kernels && lr && workspace && b > 0 && c > 0, "degrade_fused: bad arguments");
    CSBSR_REQUIRE(ksize == kKS && factor == 4, "degrade_fused: only ksize = 21, factor = 4 (the CSBSR configuration)");
    CSBSR_REQUIRE(h % 4 == 0 && w % 4 == 0 && h >= 16 && w >= 16, "degrade_fused: size must be a multiple of 4, >= 16");
    CSBSR_REQUIRE(workspace_bytes >= csbsr_degrade_workspace_bytes(b), "degrade_fused: workspace too small");
    CSBSR_REQUIRE((reinterpret_cast<uintptr_t>(hr) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0,
                  "degrade_fused: hr / workspace must be 16-byte aligned");
    const int oh = h / 4, ow = w / 4;
    float* k25 = static_cast<float*>(workspace);
    degrade_prep_kernel<<<dim3(b, kNCls * kNCls), 256, 0, STREAM(stream)>>>(params, kernels, k25, h, w, oh, ow);
    const int smem = static_cast<int>(sizeof(float)) * (kKC2 + kDT_PH * kDT_PW);
    static bool attr = false;
    if (!attr) {
        CSBSR_CHECK_CUDA(cudaFuncSetAttribute(degrade_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    dim3 grid((ow + kDT_OW - 1) / kDT_OW, (oh + kDT_OH - 1) / kDT_OH, b * c);
    degrade_main_kernel<<<grid, 128, smem, STREAM(stream)>>>(hr, k25, lr, c, h, w, oh, ow, clamp01);
    degrade_border_kernel<<<dim3(8, b * c), 256, 0, STREAM(stream)>>>(hr, k25, lr, c, h, w, oh, ow, clamp01);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// throughput mode: per-sample (theta, sigma_x, sigma_y) drawn on the device with Philox4x32-10 (the reference draws them on the
// host per sample, blur.py:129,170-179; parity mode passes host draws to csbsr_degrade / csbsr_degrade_fused instead)
extern "C" int csbsr_degrade_params_philox(double* params, int b, unsigned long long seed, unsigned long long offset,
                                           double theta_lo, double theta_hi, double sigma_lo, double sigma_hi, void* stream) {
    CSBSR_REQUIRE(params && b > 0, "degrade_params_philox: bad arguments");
    philox_params_kernel<<<(b + 127) / 128, 128, 0, STREAM(stream)>>>(params, b, seed, offset, theta_lo, theta_hi, sigma_lo,
                                                                     sigma_hi);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
