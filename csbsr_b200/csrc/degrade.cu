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
