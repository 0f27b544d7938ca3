// Training-step support kernels: HBM-bound, 128-bit vectorised.
//   csbsr_prelu_fwd / _bwd : single-slope PReLU on bf16 NHWC maps; the slope gradient is reduced in fp32
//                            (reference ConvBlock / DeconvBlock activations, model/modeling/kbpn.py:190-248)
//   csbsr_adam_step        : Adam on flat fp32 parameter / gradient / moment buffers, with the gradient zeroing of the
//                            next step fused in (train.py:91 torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8);
//                            trainer.py:61,70-71 zero_grad / step)
#include "common.cuh"
#include "../../include/csbsr_b200.h"

namespace csbsr {

__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    return r;
}

__global__ void prelu_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, const float* __restrict__ slope,
                                 size_t n8) {
    const float a = *slope;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float f[8];
        unpack8(x[i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = f[j] >= 0.f ? f[j] : a * f[j];
        y[i] = pack8(f);
    }
}

__global__ void prelu_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, uint4* __restrict__ dx,
                                 const float* __restrict__ slope, float* __restrict__ dslope, size_t n8) {
    const float a = *slope;
    float acc = 0.f;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float fx[8], fg[8];
        unpack8(x[i], fx);
        unpack8(dy[i], fg);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (fx[j] < 0.f) {
                acc = fmaf(fg[j], fx[j], acc);
                fg[j] *= a;
            }
        }
        dx[i] = pack8(fg);
    }
    acc = warp_sum(acc);
    __shared__ float part[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (warp == 0) {
        float v = lane < (blockDim.x >> 5) ? part[lane] : 0.f;
        v = warp_sum(v);
        if (lane == 0) atomicAdd(dslope, v);
    }
}

__global__ void adam_kernel(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                            size_t n4, float lr_over_bc1, float beta1, float beta2, float eps, float inv_sqrt_bc2,
                            float grad_scale, int zero_grad) {
    const float w1 = 1.f - beta1, w2 = 1.f - beta2;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
        float* pf = reinterpret_cast<float*>(&pp);
        float* gf = reinterpret_cast<float*>(&gg);
        float* mf = reinterpret_cast<float*>(&mm);
        float* vf = reinterpret_cast<float*>(&vv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = gf[j] * grad_scale;
            mf[j] = mf[j] + w1 * (gr - mf[j]);                       // exp_avg.lerp_(grad, 1 - beta1)
            vf[j] = vf[j] * beta2 + w2 * gr * gr;                    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
            const float denom = sqrtf(vf[j]) * inv_sqrt_bc2 + eps;   // sqrt(v) / sqrt(bias_correction2) + eps
            pf[j] = pf[j] - lr_over_bc1 * (mf[j] / denom);           // param.addcdiv_(exp_avg, denom, value=-lr/bias_correction1)
        }
        p[i] = pp; m[i] = mm; v[i] = vv;
        if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

static int grid_cap(size_t work, int block) {
    size_t b = (work + block - 1) / block;
    const size_t cap = static_cast<size_t>(num_sms()) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

}  // namespace csbsr

using namespace csbsr;

extern "C" int csbsr_prelu_fwd(const void* x, void* y, const float* slope, long long n, void* stream) {
    CSBSR_REQUIRE(x && y && slope && n > 0 && n % 8 == 0, "prelu_fwd: n must be a positive multiple of 8");
    const size_t n8 = static_cast<size_t>(n) / 8;
    prelu_fwd_kernel<<<grid_cap(n8, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), slope, n8);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_prelu_bwd(const void* x, const void* dy, void* dx, const float* slope, float* dslope, long long n,
                               void* stream) {
    CSBSR_REQUIRE(x && dy && dx && slope && dslope && n > 0 && n % 8 == 0, "prelu_bwd: n must be a positive multiple of 8");
    const size_t n8 = static_cast<size_t>(n) / 8;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CSBSR_CHECK_CUDA(cudaMemsetAsync(dslope, 0, sizeof(float), st));
    prelu_bwd_kernel<<<grid_cap(n8, 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(dy),
                                                        reinterpret_cast<uint4*>(dx), slope, dslope, n8);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_adam_step(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                               float eps, int step, float grad_scale, int zero_grad, void* stream) {
    CSBSR_REQUIRE(p && g && m && v && n > 0 && n % 4 == 0, "adam_step: n must be a positive multiple of 4");
    CSBSR_REQUIRE(step >= 1, "adam_step: step counts from 1");
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
    const size_t n4 = static_cast<size_t>(n) / 4;
    adam_kernel<<<grid_cap(n4, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(g), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v),
        n4, static_cast<float>(lr / bc1), beta1, beta2, eps, static_cast<float>(1.0 / sqrt(bc2)), grad_scale, zero_grad);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
