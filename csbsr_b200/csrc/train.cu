// Training-step support kernels: HBM-bound, 128-bit vectorised.
//   csbsr_prelu_fwd / _bwd : single-slope PReLU on bf16 NHWC maps; the slope gradient is reduced in fp32
//                            (reference ConvBlock / DeconvBlock activations, model/modeling/kbpn.py:190-248)
//   csbsr_adam_step        : Adam on flat fp32 parameter / gradient / moment buffers, with the gradient zeroing of the
//                            next step fused in (train.py:91 torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8);
//                            trainer.py:61,70-71 zero_grad / step)
#include "common.cuh"
#include "../../include/csbsr_b200.h"

namespace csbsr {

static int grid_cap(size_t work, int block);

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

__device__ __forceinline__ float aa_cubic_t(float x) {       // Keys cubic, a = -0.5 (antialiased bicubic of torchvision Resize)
    const float a = -0.5f;
    x = fabsf(x);
    if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
    if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * a;
    return 0.f;
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
        if (lane == 0) dslope[blockIdx.x] = v;               // per-block partial; prelu_finalize_kernel adds them in a fixed order
    }
}
__global__ void prelu_finalize_kernel(const float* __restrict__ part, int n, float* __restrict__ dslope) {
    float v = 0.f;
    for (int i = threadIdx.x; i < n; i += 32) v += part[i];
    v = warp_sum(v);
    if (threadIdx.x == 0) *dslope = v;
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

extern "C" size_t csbsr_prelu_bwd_workspace_bytes(void) { return sizeof(float) * static_cast<size_t>(num_sms()) * 16; }

extern "C" int csbsr_prelu_bwd(const void* x, const void* dy, void* dx, const float* slope, float* dslope, long long n,
                               float* workspace, void* stream) {
    CSBSR_REQUIRE(x && dy && dx && slope && dslope && workspace && n > 0 && n % 8 == 0, "prelu_bwd: n must be a positive multiple of 8");
    const size_t n8 = static_cast<size_t>(n) / 8;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int pblocks = grid_cap(n8, 256);                 // <= num_sms() * 16 partial sums (csbsr_prelu_bwd_workspace_bytes)
    prelu_bwd_kernel<<<pblocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(dy),
                                              reinterpret_cast<uint4*>(dx), slope, workspace, n8);
    prelu_finalize_kernel<<<1, 32, 0, st>>>(workspace, pblocks, dslope);
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

// ---------------------------------------------------------------------------------------------------------------
// Backward of the per-sample depthwise blur (csbsr_blur_per_sample; KBlock pseudo-LR, kbpn.py:395-402, and Get_pseudo_lr of
// KBPNLoss, sr_loss_functions.py:73-102) and of the antialiased bicubic downscale (FactorResize, transforms.py:516-531).
//   forward   y[b,c,Y,X] = sum_ij k[b,i,j] * x[b,c,Y*s+i-pad, X*s+j-pad]
//   d input   dx[b,c,y,x] = sum over (Y,X,i,j) with Y*s+i-pad = y, X*s+j-pad = x of k[b,i,j] * dy[b,c,Y,X]
//   d kernel  dk[b,i,j]   = sum_{c,Y,X} dy[b,c,Y,X] * x[b,c,Y*s+i-pad, X*s+j-pad]
namespace csbsr {

__global__ void blur_bwd_input_kernel(const float* __restrict__ dy, const float* __restrict__ kvec, float* __restrict__ dx,
                                      int C, int H, int W, int OH, int OW, int ks, int stride) {
    extern __shared__ float sk[];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < ks * ks; i += blockDim.x) sk[i] = kvec[static_cast<size_t>(b) * ks * ks + i];
    __syncthreads();
    const int pad = (ks - 1) / 2;
    const size_t per = static_cast<size_t>(C) * H * W;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < per;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % W);
        const int y = static_cast<int>((i / W) % H);
        const int c = static_cast<int>(i / (static_cast<size_t>(W) * H));
        // Y*s + ti - pad = y  ->  ti = y + pad - Y*s in [0, ks)
        const int Y0 = max(0, (y + pad - (ks - 1) + stride - 1) / stride), Y1 = min(OH - 1, (y + pad) / stride);
        const int X0 = max(0, (x + pad - (ks - 1) + stride - 1) / stride), X1 = min(OW - 1, (x + pad) / stride);
        const float* dp = dy + (static_cast<size_t>(b) * C + c) * OH * OW;
        float acc = 0.f;
        for (int Y = Y0; Y <= Y1; ++Y) {
            const int ti = y + pad - Y * stride;
            for (int X = X0; X <= X1; ++X) acc = fmaf(sk[ti * ks + (x + pad - X * stride)], dp[static_cast<size_t>(Y) * OW + X], acc);
        }
        dx[static_cast<size_t>(b) * per + i] = acc;
    }
}

// one block per (pixel tile of dy, channel, sample); thread t < ks*ks owns tap t and walks the dy tile staged in shared memory
// next to the x halo tile; partial sums are added atomically (fp32) into dk
template <int TILE>
__global__ void blur_bwd_kernel_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dk, int C,
                                       int H, int W, int OH, int OW, int ks, int stride) {
    extern __shared__ float sm[];
    const int XT = (TILE - 1) * stride + ks;                 // x halo tile edge
    float* sdy = sm;                                         // [TILE][TILE]
    float* sx = sm + TILE * TILE;                            // [XT][XT]
    const int tiles_w = (OW + TILE - 1) / TILE;
    const int ty0 = (blockIdx.x / tiles_w) * TILE, tx0 = (blockIdx.x % tiles_w) * TILE;
    const int c = blockIdx.y, b = blockIdx.z;
    const int pad = (ks - 1) / 2;
    const float* xp = x + (static_cast<size_t>(b) * C + c) * H * W;
    const float* dp = dy + (static_cast<size_t>(b) * C + c) * OH * OW;
    for (int i = threadIdx.x; i < TILE * TILE; i += blockDim.x) {
        const int Y = ty0 + i / TILE, X = tx0 + i % TILE;
        sdy[i] = (Y < OH && X < OW) ? dp[static_cast<size_t>(Y) * OW + X] : 0.f;
    }
    const int y0 = ty0 * stride - pad, x0 = tx0 * stride - pad;
    for (int i = threadIdx.x; i < XT * XT; i += blockDim.x) {
        const int yy = y0 + i / XT, xx = x0 + i % XT;
        sx[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? xp[static_cast<size_t>(yy) * W + xx] : 0.f;
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t < ks * ks) {
        const int ti = t / ks, tj = t % ks;
        float acc = 0.f;
        for (int Y = 0; Y < TILE; ++Y) {
            const float* xr = sx + (Y * stride + ti) * XT + tj;
            const float* dr = sdy + Y * TILE;
#pragma unroll 8
            for (int X = 0; X < TILE; ++X) acc = fmaf(dr[X], xr[X * stride], acc);
        }
        // partial of this (tile, channel): part[b][tile * C + c][t]; blur_bwd_kernel_finalize_kernel adds them in order
        dk[((static_cast<size_t>(b) * gridDim.x * C) + static_cast<size_t>(blockIdx.x) * C + c) * ks * ks + t] = acc;
    }
}
__global__ void blur_bwd_kernel_finalize_kernel(const float* __restrict__ part, int nparts, int kk, float* __restrict__ dk) {
    const int b = blockIdx.x;
    for (int t = threadIdx.x; t < kk; t += blockDim.x) {
        float v = 0.f;
        for (int i = 0; i < nparts; ++i) v += part[(static_cast<size_t>(b) * nparts + i) * kk + t];
        dk[static_cast<size_t>(b) * kk + t] = v;
    }
}

__global__ void resize_aa_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int NC, int H, int W, int OH, int OW) {
    const size_t total = static_cast<size_t>(NC) * H * W;
    const float sh = static_cast<float>(H) / OH, sw = static_cast<float>(W) / OW;
    const float ish = 1.f / sh, isw = 1.f / sw;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % W);
        const int y = static_cast<int>((i / W) % H);
        const int nc = static_cast<int>(i / (static_cast<size_t>(W) * H));
        // candidate outputs whose (antialiased, support 2*scale) window may contain this input pixel
        float wy[8], wx[8];
        int oy0 = static_cast<int>(floorf((y - 2.f * sh) * ish - 0.5f)), ox0 = static_cast<int>(floorf((x - 2.f * sw) * isw - 0.5f));
        for (int k = 0; k < 8; ++k) {
            wy[k] = wx[k] = 0.f;
            const int oy = oy0 + k, ox = ox0 + k;
            if (oy >= 0 && oy < OH) {
                const float center = sh * (oy + 0.5f), support = 2.f * sh;
                const int lo = max(static_cast<int>(center - support + 0.5f), 0), hi = min(static_cast<int>(center + support + 0.5f), H);
                if (y >= lo && y < hi) {
                    float s = 0.f;
                    for (int r = lo; r < hi; ++r) s += aa_cubic_t((r - center + 0.5f) * ish);
                    wy[k] = aa_cubic_t((y - center + 0.5f) * ish) / s;
                }
            }
            if (ox >= 0 && ox < OW) {
                const float center = sw * (ox + 0.5f), support = 2.f * sw;
                const int lo = max(static_cast<int>(center - support + 0.5f), 0), hi = min(static_cast<int>(center + support + 0.5f), W);
                if (x >= lo && x < hi) {
                    float s = 0.f;
                    for (int r = lo; r < hi; ++r) s += aa_cubic_t((r - center + 0.5f) * isw);
                    wx[k] = aa_cubic_t((x - center + 0.5f) * isw) / s;
                }
            }
        }
        const float* dp = dy + static_cast<size_t>(nc) * OH * OW;
        float acc = 0.f;
        for (int a = 0; a < 8; ++a) {
            if (wy[a] == 0.f) continue;
            float row = 0.f;
            for (int k = 0; k < 8; ++k)
                if (wx[k] != 0.f) row = fmaf(wx[k], dp[static_cast<size_t>(oy0 + a) * OW + ox0 + k], row);
            acc = fmaf(wy[a], row, acc);
        }
        dx[i] = acc;
    }
}

}  // namespace csbsr

extern "C" int csbsr_blur_ps_bwd_input(const float* dy, const float* kvec, float* dx, int b, int c, int h, int w, int ksize,
                                       int stride, void* stream) {
    CSBSR_REQUIRE(dy && kvec && dx && b > 0 && c > 0 && ksize > 0 && (ksize & 1) && stride >= 1, "blur_ps_bwd_input: bad arguments");
    const int oh = (h - 1) / stride + 1, ow = (w - 1) / stride + 1;
    const size_t per = static_cast<size_t>(c) * h * w;
    dim3 grid(grid_cap(per, 256), b);
    blur_bwd_input_kernel<<<grid, 256, sizeof(float) * ksize * ksize, reinterpret_cast<cudaStream_t>(stream)>>>(
        dy, kvec, dx, c, h, w, oh, ow, ksize, stride);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int blur_bwd_tiles(int h, int w, int stride) {
    const int oh = (h - 1) / stride + 1, ow = (w - 1) / stride + 1;
    const int T = stride == 1 ? 32 : 8;
    return ((oh + T - 1) / T) * ((ow + T - 1) / T);
}

extern "C" size_t csbsr_blur_ps_bwd_kernel_workspace_bytes(int b, int c, int h, int w, int ksize, int stride) {
    return sizeof(float) * static_cast<size_t>(b) * blur_bwd_tiles(h, w, stride) * c * ksize * ksize;
}

extern "C" int csbsr_blur_ps_bwd_kernel(const float* x, const float* dy, float* dk, int b, int c, int h, int w, int ksize,
                                        int stride, float* workspace, void* stream) {
    CSBSR_REQUIRE(x && dy && dk && workspace && b > 0 && c > 0 && ksize > 0 && (ksize & 1) && ksize * ksize <= 512 && stride >= 1,
                  "blur_ps_bwd_kernel: bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int oh = (h - 1) / stride + 1, ow = (w - 1) / stride + 1;
    const int tiles = blur_bwd_tiles(h, w, stride);
    if (stride == 1) {
        constexpr int T = 32;
        const int xt = (T - 1) * stride + ksize;
        dim3 grid(tiles, c, b);
        blur_bwd_kernel_kernel<T><<<grid, 512, sizeof(float) * (T * T + xt * xt), st>>>(x, dy, workspace, c, h, w, oh, ow, ksize, stride);
    } else {
        constexpr int T = 8;
        const int xt = (T - 1) * stride + ksize;
        CSBSR_REQUIRE(sizeof(float) * (T * T + xt * xt) <= 48 * 1024, "blur_ps_bwd_kernel: stride %d too large", stride);
        dim3 grid(tiles, c, b);
        blur_bwd_kernel_kernel<T><<<grid, 512, sizeof(float) * (T * T + xt * xt), st>>>(x, dy, workspace, c, h, w, oh, ow, ksize, stride);
    }
    blur_bwd_kernel_finalize_kernel<<<b, 256, 0, st>>>(workspace, tiles * c, ksize * ksize, dk);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_resize_bicubic_aa_bwd(const float* dy, float* dx, int nc, int h, int w, int oh, int ow, void* stream) {
    CSBSR_REQUIRE(dy && dx && nc > 0 && oh > 0 && ow > 0 && oh <= h && ow <= w && h <= 5 * oh && w <= 5 * ow,
                  "resize_bicubic_aa_bwd: downscale factors up to 5 only");
    const size_t total = static_cast<size_t>(nc) * h * w;
    resize_aa_bwd_kernel<<<grid_cap(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dy, dx, nc, h, w, oh, ow);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// fp32 parameter [A][B][R][S] -> packed bf16 GEMM operand [R*S][rows_pad][cols_pad] in one launch (the training step re-packs
// every weight twice per iteration: forward and dgrad layouts).  mode 0: rows = A, cols = B (nn.Conv2d forward; deconv dgrad);
// mode 1: rows = B, cols = A, taps flipped (stride-1 conv dgrad); mode 2: rows = B, cols = A (ConvTranspose2d forward phases;
// dgrad of the 8x8/s4 conv).  Padding rows / columns are written as zeros.
namespace csbsr {
// one packing job: source parameter [A][Btot][R][S], columns [b0, b0 + B) of its second axis are packed
struct PackJob {
    const float* w;
    __nv_bfloat16* out;
    int A, B, Btot, b0, R, S, rows_pad, cols_pad, mode;
    int pad_;
    unsigned long long start;      // first element of this job in the concatenated index space of a multi-job launch
};

__device__ __forceinline__ void pack_one(const PackJob& j, size_t i) {
    const int col = static_cast<int>(i % j.cols_pad);
    const int row = static_cast<int>((i / j.cols_pad) % j.rows_pad);
    if (j.mode >= 3) {
        // tap expansion of a 3x3 conv with few outputs (A <= 4): one 1x1 GEMM with 9 * cp outputs, output t * cp + m = tap t
        // of output m (cp = j.pad_).  mode 3: rows = (t, m), cols = input channel; mode 4 (its dgrad): the transpose
        const int e = j.mode == 3 ? row : col, c = j.mode == 3 ? col : row;
        const int t = e / j.pad_, m = e - t * j.pad_;
        float v = 0.f;
        if (t < 9 && m < j.A && c < j.B) v = j.w[(static_cast<size_t>(m) * j.Btot + j.b0 + c) * 9 + t];
        j.out[i] = __float2bfloat16(v);
        return;
    }
    const int t = static_cast<int>(i / (static_cast<size_t>(j.cols_pad) * j.rows_pad));
    int r = t / j.S, s = t % j.S;
    const int a = j.mode == 0 ? row : col, b = j.mode == 0 ? col : row;
    if (j.mode == 1) { r = j.R - 1 - r; s = j.S - 1 - s; }
    float v = 0.f;
    if (a < j.A && b < j.B) v = j.w[((static_cast<size_t>(a) * j.Btot + j.b0 + b) * j.R + r) * j.S + s];
    j.out[i] = __float2bfloat16(v);
}

__global__ void pack_weights_kernel(const PackJob job) {
    const size_t total = static_cast<size_t>(job.R) * job.S * job.rows_pad * job.cols_pad;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        pack_one(job, i);
}

// every registered (parameter, layout) pair of the model in ONE launch: jobs[] lives in device memory, `start` is the exclusive
// prefix of the jobs' TILE counts.  A tile is 4 x 32 (mode 0: a x b) or 32 x 4 (modes 1, 2) source rows/columns with all R*S taps:
// the source is read in runs of 32*T (4*T) contiguous floats, transposed through shared memory and written as 64-byte runs
// along the packed operand's contiguous axis -- the element-wise form reads with a stride of T floats (12.5 % sector efficiency
// for the 8x8 weights).  Only the valid region is written: the cached packed buffers are zero-initialised once.
__host__ __device__ __forceinline__ int pack_tile_factor(int T) { return T > 32 ? 1 : (T > 16 ? 2 : (T > 8 ? 4 : 8)); }
__host__ __device__ __forceinline__ long long pack_job_tiles(int A, int B, int T, int rows_pad, int cols_pad, int mode) {
    if (mode >= 3) return (static_cast<long long>(rows_pad) * cols_pad + 1023) / 1024;
    const int f = pack_tile_factor(T);
    return mode == 0 ? static_cast<long long>((A + 4 * f - 1) / (4 * f)) * ((B + 31) / 32)
                     : static_cast<long long>((A + 31) / 32) * ((B + 4 * f - 1) / (4 * f));
}
// one tile of a mode 0 / 1 / 2 job: every thread owns one (a, b) pair and walks its R*S taps -- nested loops instead of index
// divisions, no shared memory.  Lanes run along the packed operand's contiguous axis (b for mode 0, a for modes 1 / 2), so the
// bf16 stores are 64-byte runs; the fp32 source is read in per-warp regions of 32*T (mode 0) or 4f*T per lane (modes 1 / 2)
// contiguous floats that stay in L1 across the tap loop.
template <bool M0>
__device__ __forceinline__ void pack_tile(const PackJob& cur, int lt) {
    const int T = cur.R * cur.S;
    const int f = pack_tile_factor(T);
    const int na = M0 ? 4 * f : 32, nb = M0 ? 32 : 4 * f;
    const int tiles_b = (cur.B + nb - 1) / nb;
    const int a0 = (lt / tiles_b) * na, bb0 = (lt % tiles_b) * nb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t plane = static_cast<size_t>(cur.rows_pad) * cur.cols_pad;
    for (int k = warp; k < (M0 ? na : nb); k += 8) {
        const int a = a0 + (M0 ? k : lane), b = bb0 + (M0 ? lane : k);
        if (a >= cur.A || b >= cur.B) continue;
        const float* src = cur.w + (static_cast<size_t>(a) * cur.Btot + cur.b0 + b) * T;
        __nv_bfloat16* dst = cur.out + (M0 ? static_cast<size_t>(a) * cur.cols_pad + b : static_cast<size_t>(b) * cur.cols_pad + a);
        if (cur.mode == 1) {
            for (int t = 0; t < T; ++t) dst[static_cast<size_t>(T - 1 - t) * plane] = __float2bfloat16(src[t]);   // taps flipped
        } else {
            for (int t = 0; t < T; ++t) dst[static_cast<size_t>(t) * plane] = __float2bfloat16(src[t]);
        }
    }
}

__global__ void __launch_bounds__(256)
pack_weights_multi_kernel(const PackJob* __restrict__ jobs, int njobs, unsigned long long total_tiles) {
    __shared__ PackJob cur;
    // every block owns a contiguous range of tiles: the job is searched once and then advanced sequentially
    const unsigned long long per = (total_tiles + gridDim.x - 1) / gridDim.x;
    const unsigned long long lo = per * blockIdx.x, hi = min(total_tiles, lo + per);
    if (lo >= hi) return;
    int j;
    {
        int a = 0, b = njobs - 1;
        while (a < b) {
            const int m = (a + b + 1) >> 1;
            if (jobs[m].start <= lo) a = m; else b = m - 1;
        }
        j = a;
    }
    unsigned long long jend = 0;
    bool have = false;
    for (unsigned long long tile = lo; tile < hi; ++tile) {
        if (!have || tile >= jend) {
            if (have) ++j;
            __syncthreads();
            if (threadIdx.x == 0) cur = jobs[j];
            __syncthreads();
            jend = cur.start + static_cast<unsigned long long>(pack_job_tiles(cur.A, cur.B, cur.R * cur.S, cur.rows_pad, cur.cols_pad, cur.mode));
            have = true;
            if (tile >= jend) { --tile; continue; }         // (empty job: cannot happen, sizes are positive)
        }
        const int lt = static_cast<int>(tile - cur.start);
        const int T = cur.R * cur.S;
        if (cur.mode >= 3) {
            const size_t n = static_cast<size_t>(cur.rows_pad) * cur.cols_pad;
            for (size_t i = static_cast<size_t>(lt) * 1024 + threadIdx.x; i < min(n, static_cast<size_t>(lt + 1) * 1024); i += 256) pack_one(cur, i);
            continue;
        }
        if (cur.mode == 0) pack_tile<true>(cur, lt); else pack_tile<false>(cur, lt);
    }
}

// grad[a][b0 + b][r][s] += wg[a][t = r*S+s][b]  (a < A, b < B): the fp32 accumulator of csbsr_conv_wgrad ([rows][taps][cs])
// folded into the parameter's gradient in the parameter's own layout ([A][Btot][R][S]) -- replaces permute + contiguous + add
__global__ void wgrad_unpack_add_kernel(const float* __restrict__ wg, float* __restrict__ grad, int A, int B, int Btot, int b0,
                                        int T, int cs) {
    const size_t total = static_cast<size_t>(A) * B * T;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(i % T);
        const int b = static_cast<int>((i / T) % B);
        const int a = static_cast<int>(i / (static_cast<size_t>(T) * B));
        grad[(static_cast<size_t>(a) * Btot + b0 + b) * T + t] += wg[(static_cast<size_t>(a) * T + t) * cs + b];
    }
}
// tap-expanded form: grad[m][b0 + c][t] += wg[t * cp + m][c]   (wg: [rows][cs], a single "tap")
__global__ void wgrad_unpack_add_tapexp_kernel(const float* __restrict__ wg, float* __restrict__ grad, int A, int B, int Btot, int b0,
                                               int cp, int cs) {
    const size_t total = static_cast<size_t>(A) * B * 9;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(i % 9);
        const int c = static_cast<int>((i / 9) % B);
        const int m = static_cast<int>(i / (static_cast<size_t>(9) * B));
        grad[(static_cast<size_t>(m) * Btot + b0 + c) * 9 + t] += wg[static_cast<size_t>(t * cp + m) * cs + c];
    }
}
}  // namespace csbsr

extern "C" int csbsr_pack_weights(const float* w, void* out, int a, int b, int r, int s, int rows_pad, int cols_pad, int mode,
                                  void* stream) {
    return csbsr_pack_weights_window(w, out, a, b, b, 0, r, s, rows_pad, cols_pad, mode, stream);
}

static int pack_job_check(int a, int b, int b_total, int b0, int r, int s, int rows_pad, int cols_pad, int mode) {
    CSBSR_REQUIRE(a > 0 && b > 0 && r > 0 && s > 0 && mode >= 0 && mode <= 4 + 8 * 16 && b0 >= 0 && b0 + b <= b_total,
                  "pack_weights: bad arguments");
    const int m = mode & 7, cp = mode >> 3;
    if (m >= 3) {
        CSBSR_REQUIRE(m <= 4 && r == 1 && s == 1 && cp >= a && 9 * cp <= (m == 3 ? rows_pad : cols_pad) && b <= (m == 3 ? cols_pad : rows_pad),
                      "pack_weights: tap-expanded modes need r = s = 1 (the 3x3 source is implied), cp >= a, 9 * cp within the padded extent");
    } else {
        CSBSR_REQUIRE(cp == 0 && rows_pad >= (m == 0 ? a : b) && cols_pad >= (m == 0 ? b : a), "pack_weights: padded extents too small");
    }
    return 0;
}

extern "C" int csbsr_pack_weights_window(const float* w, void* out, int a, int b, int b_total, int b0, int r, int s, int rows_pad,
                                         int cols_pad, int mode, void* stream) {
    CSBSR_REQUIRE(w && out, "pack_weights: null pointer");
    if (int rc = pack_job_check(a, b, b_total, b0, r, s, rows_pad, cols_pad, mode)) return rc;
    const size_t total = static_cast<size_t>(r) * s * rows_pad * cols_pad;
    csbsr::PackJob job{w, reinterpret_cast<__nv_bfloat16*>(out), a, b, b_total, b0, r, s, rows_pad, cols_pad, mode & 7, mode >> 3, 0ull};
    csbsr::pack_weights_kernel<<<grid_cap(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(job);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" size_t csbsr_pack_job_bytes(void) { return sizeof(csbsr::PackJob); }

extern "C" int csbsr_pack_job_fill(void* job_host, const float* w, void* out, int a, int b, int b_total, int b0, int r, int s,
                                   int rows_pad, int cols_pad, int mode, unsigned long long start) {
    CSBSR_REQUIRE(job_host && w && out, "pack_job_fill: null pointer");
    if (int rc = pack_job_check(a, b, b_total, b0, r, s, rows_pad, cols_pad, mode)) return rc;
    csbsr::PackJob job{w, reinterpret_cast<__nv_bfloat16*>(out), a, b, b_total, b0, r, s, rows_pad, cols_pad, mode & 7, mode >> 3, start};
    memcpy(job_host, &job, sizeof(job));
    return 0;
}

extern "C" long long csbsr_pack_job_tiles(int a, int b, int r, int s, int rows_pad, int cols_pad, int mode) {
    return csbsr::pack_job_tiles(a, b, r * s, rows_pad, cols_pad, mode & 7);
}

extern "C" int csbsr_pack_weights_multi(const void* jobs_device, int njobs, unsigned long long total_tiles, int max_taps, void* stream) {
    CSBSR_REQUIRE(jobs_device && njobs > 0 && total_tiles > 0 && max_taps >= 1 && max_taps <= 64, "pack_weights_multi: bad arguments");
    unsigned long long blocks = total_tiles;
    const unsigned long long cap = static_cast<unsigned long long>(csbsr::num_sms()) * 8;
    if (blocks > cap) blocks = cap;
    (void)max_taps;
    const int smem = 0;
    csbsr::pack_weights_multi_kernel<<<static_cast<int>(blocks), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        static_cast<const csbsr::PackJob*>(jobs_device), njobs, total_tiles);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_wgrad_unpack_add_tapexp(const float* wg, float* grad, int a, int b, int b_total, int b0, int cp, int cs,
                                             void* stream) {
    CSBSR_REQUIRE(wg && grad && a > 0 && a <= cp && b > 0 && b <= cs && b0 >= 0 && b0 + b <= b_total, "wgrad_unpack_add_tapexp: bad arguments");
    const size_t total = static_cast<size_t>(a) * b * 9;
    csbsr::wgrad_unpack_add_tapexp_kernel<<<grid_cap(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(wg, grad, a, b,
                                                                                                                 b_total, b0, cp, cs);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_wgrad_unpack_add(const float* wg, float* grad, int a, int b, int b_total, int b0, int taps, int cs,
                                      void* stream) {
    CSBSR_REQUIRE(wg && grad && a > 0 && b > 0 && b <= cs && b0 >= 0 && b0 + b <= b_total && taps > 0, "wgrad_unpack_add: bad arguments");
    const size_t total = static_cast<size_t>(a) * b * taps;
    csbsr::wgrad_unpack_add_kernel<<<grid_cap(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(wg, grad, a, b, b_total,
                                                                                                          b0, taps, cs);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// BatchNorm2d for the training graph on NHWC bf16 maps [m pixels][pitch channels] (first c channels real, the rest zero
// padding): batch statistics or running statistics, optional fused residual add and ReLU (BasicBlock / Bottleneck tails,
// pspnet_pytorch/extractors.py:52-70; hrnet_backbone.py), with the matching backward.  Replaces at::native batch_norm_*.
//   y = relu?( (x - mean) * rstd * gamma + beta + res? )
//   dyz = dy * (y > 0 if relu);  dbeta = sum dyz;  dgamma = sum dyz * xhat;  dres = dyz
//   dx = gamma * rstd * (dyz - [training] (dbeta + xhat * dgamma) / m)
namespace csbsr {

// thread -> (pixel lane, 8-channel group); per-channel partial sums are combined through shared memory, then atomics
template <int NS>   // NS accumulators per channel
struct ChanAcc {
    float v[NS][8];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int s = 0; s < NS; ++s)
#pragma unroll
            for (int j = 0; j < 8; ++j) v[s][j] = 0.f;
    }
};

template <int NS>
__device__ __forceinline__ void chan_block_reduce(ChanAcc<NS>& a, int grp, int lane_pix, int lanes, int groups, float* smem,
                                                  float* const* out) {
    // smem: [lanes][groups*8] per accumulator, processed one accumulator at a time
    for (int s = 0; s < NS; ++s) {
        __syncthreads();
        if (grp < groups) {
#pragma unroll
            for (int j = 0; j < 8; ++j) smem[(lane_pix * groups + grp) * 8 + j] = a.v[s][j];
        }
        __syncthreads();
        for (int i = threadIdx.x; i < groups * 8; i += blockDim.x) {
            float t = 0.f;
            for (int l = 0; l < lanes; ++l) t += smem[l * groups * 8 + i];
            out[s][static_cast<size_t>(blockIdx.x) * groups * 8 + i] = t;      // per-block partial, reduced in block order later
        }
    }
}

__global__ void bn_stats_kernel(const uint4* __restrict__ x, int pitch8, int groups, long long m, float* sum, float* sumsq) {
    extern __shared__ float bn_sm[];
    const int lanes = blockDim.x / groups;
    const int grp = threadIdx.x % groups, lane_pix = threadIdx.x / groups;
    ChanAcc<2> a;
    a.zero();
    if (lane_pix < lanes) {
        for (long long p = static_cast<long long>(blockIdx.x) * lanes + lane_pix; p < m; p += static_cast<long long>(gridDim.x) * lanes) {
            float f[8];
            unpack8(x[p * pitch8 + grp], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { a.v[0][j] += f[j]; a.v[1][j] = fmaf(f[j], f[j], a.v[1][j]); }
        }
    }
    float* outs[2] = {sum, sumsq};
    chan_block_reduce<2>(a, lane_pix < lanes ? grp : groups, lane_pix < lanes ? lane_pix : 0, lanes, groups, bn_sm, outs);
}

__global__ void bn_finalize_kernel(const float* __restrict__ sum_part, const float* __restrict__ sumsq_part, int nblocks, long long m,
                                   int c, float eps, float momentum, float* mean, float* rstd, float* running_mean,
                                   float* running_var) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    double s = 0.0, ss = 0.0;
    for (int b = 0; b < nblocks; ++b) {                        // fixed order: bit-reproducible statistics
        s += sum_part[static_cast<size_t>(b) * c + i];
        ss += sumsq_part[static_cast<size_t>(b) * c + i];
    }
    const double mu = s / m;
    double var = ss / m - mu * mu;
    if (var < 0) var = 0;
    mean[i] = static_cast<float>(mu);
    rstd[i] = static_cast<float>(1.0 / sqrt(var + eps));
    if (running_mean) {                                        // nn.BatchNorm2d: unbiased variance in the running estimate
        const double unb = m > 1 ? var * m / (m - 1) : var;
        running_mean[i] = static_cast<float>((1.0 - momentum) * running_mean[i] + momentum * mu);
        running_var[i] = static_cast<float>((1.0 - momentum) * running_var[i] + momentum * unb);
    }
}

__global__ void bn_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ res, uint4* __restrict__ y,
                                const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int groups, int pitch8, long long m, int relu) {
    const long long total = m * pitch8;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % pitch8);
        float f[8];
        if (g < groups) {
            unpack8(x[i], f);
            float r[8];
            if (res) unpack8(res[i], r);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int ch = g * 8 + j;
                float v = (f[j] - mean[ch]) * rstd[ch] * gamma[ch] + beta[ch];
                if (res) v += r[j];
                f[j] = relu ? fmaxf(v, 0.f) : v;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0.f;
        }
        y[i] = pack8(f);
    }
}

__global__ void bn_bwd_reduce_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ y,
                                     const float* __restrict__ mean, const float* __restrict__ rstd, int pitch8, int groups,
                                     long long m, float* s1, float* s2) {
    extern __shared__ float bn_sm[];
    const int lanes = blockDim.x / groups;
    const int grp = threadIdx.x % groups, lane_pix = threadIdx.x / groups;
    ChanAcc<2> a;
    a.zero();
    if (lane_pix < lanes) {
        float mu[8], rs[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { mu[j] = mean[grp * 8 + j]; rs[j] = rstd[grp * 8 + j]; }
        for (long long p = static_cast<long long>(blockIdx.x) * lanes + lane_pix; p < m; p += static_cast<long long>(gridDim.x) * lanes) {
            float g[8], f[8], o[8];
            unpack8(dy[p * pitch8 + grp], g);
            unpack8(x[p * pitch8 + grp], f);
            if (y) unpack8(y[p * pitch8 + grp], o);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = (y && !(o[j] > 0.f)) ? 0.f : g[j];
                a.v[0][j] += d;
                a.v[1][j] = fmaf(d, (f[j] - mu[j]) * rs[j], a.v[1][j]);
            }
        }
    }
    float* outs[2] = {s1, s2};
    chan_block_reduce<2>(a, lane_pix < lanes ? grp : groups, lane_pix < lanes ? lane_pix : 0, lanes, groups, bn_sm, outs);
}

__global__ void bn_bwd_apply_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x, const uint4* __restrict__ y,
                                    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                    const float* __restrict__ s1, const float* __restrict__ s2, int groups, int pitch8, long long m,
                                    int training, uint4* __restrict__ dx, uint4* __restrict__ dres) {
    const long long total = m * pitch8;
    const float inv_m = 1.f / static_cast<float>(m);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % pitch8);
        float d[8], o8[8];
        if (g < groups) {
            float f[8], o[8];
            unpack8(dy[i], d);
            unpack8(x[i], f);
            if (y) unpack8(y[i], o);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int ch = g * 8 + j;
                const float dz = (y && !(o[j] > 0.f)) ? 0.f : d[j];
                const float xh = (f[j] - mean[ch]) * rstd[ch];
                float v = dz;
                if (training) v -= (s1[ch] + xh * s2[ch]) * inv_m;
                o8[j] = gamma[ch] * rstd[ch] * v;
                d[j] = dz;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { d[j] = 0.f; o8[j] = 0.f; }
        }
        dx[i] = pack8(o8);
        if (dres) dres[i] = pack8(d);
    }
}

// dbeta / dgamma = ordered sums of the per-block partials of bn_bwd_reduce_kernel
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ p1, const float* __restrict__ p2, int nblocks, int c,
                                       float* __restrict__ dbeta, float* __restrict__ dgamma) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    float a = 0.f, b2 = 0.f;
    for (int b = 0; b < nblocks; ++b) {
        a += p1[static_cast<size_t>(b) * c + i];
        b2 += p2[static_cast<size_t>(b) * c + i];
    }
    dbeta[i] = a;
    dgamma[i] = b2;
}

static constexpr int kBnMaxBlocks = 128;
static int bn_block_cfg(int groups, int& lanes) {      // threads per block: whole number of pixel lanes, <= 256 when possible
    lanes = 256 / groups;
    if (lanes < 1) lanes = 1;
    return lanes * groups;
}

}  // namespace csbsr

extern "C" size_t csbsr_bn_workspace_bytes(int c) { return sizeof(float) * 2 * kBnMaxBlocks * static_cast<size_t>(c); }

extern "C" int csbsr_bn_stats(const void* x, int pitch, int c, long long m, float eps, float momentum, float* mean, float* rstd,
                              float* running_mean, float* running_var, float* workspace, void* stream) {
    CSBSR_REQUIRE(x && mean && rstd && workspace && c > 0 && c % 8 == 0 && pitch % 8 == 0 && c <= pitch && c <= 8192 && m > 0,
                  "bn_stats: bad arguments (c=%d pitch=%d)", c, pitch);
    CSBSR_REQUIRE(!running_mean == !running_var, "bn_stats: running_mean and running_var go together");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    float* sum = workspace;                                     // [blocks][c] partial sums, then [blocks][c] partial sums of squares
    float* sumsq = workspace + static_cast<size_t>(kBnMaxBlocks) * c;
    const int groups = c / 8;
    int lanes;
    const int threads = bn_block_cfg(groups, lanes);
    CSBSR_REQUIRE(threads <= 1024, "bn_stats: too many channels");
    long long blocks = (m + lanes * 64 - 1) / (static_cast<long long>(lanes) * 64);
    if (blocks > kBnMaxBlocks) blocks = kBnMaxBlocks;
    if (blocks < 1) blocks = 1;
    bn_stats_kernel<<<static_cast<int>(blocks), threads, sizeof(float) * lanes * groups * 8, st>>>(
        reinterpret_cast<const uint4*>(x), pitch / 8, groups, m, sum, sumsq);
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(sum, sumsq, static_cast<int>(blocks), m, c, eps, momentum, mean, rstd,
                                                       running_mean, running_var);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_bn_apply(const void* x, const void* res, void* y, const float* mean, const float* rstd, const float* gamma,
                              const float* beta, int pitch, int c, long long m, int relu, void* stream) {
    CSBSR_REQUIRE(x && y && mean && rstd && gamma && beta && c > 0 && c % 8 == 0 && pitch % 8 == 0 && c <= pitch && m > 0,
                  "bn_apply: bad arguments");
    const long long total = m * (pitch / 8);
    bn_apply_kernel<<<grid_cap(static_cast<size_t>(total), 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(res), reinterpret_cast<uint4*>(y), mean, rstd, gamma,
        beta, c / 8, pitch / 8, m, relu);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_bn_backward(const void* dy, const void* x, const void* y_relu, const float* mean, const float* rstd,
                                 const float* gamma, int pitch, int c, long long m, int training, void* dx, void* dres,
                                 float* dgamma, float* dbeta, float* workspace, void* stream) {
    CSBSR_REQUIRE(dy && x && mean && rstd && gamma && dx && dgamma && dbeta && workspace && c > 0 && c % 8 == 0 && pitch % 8 == 0 &&
                      c <= pitch && c <= 8192 && m > 0, "bn_backward: bad arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    float* p1 = workspace;
    float* p2 = workspace + static_cast<size_t>(kBnMaxBlocks) * c;
    const int groups = c / 8;
    int lanes;
    const int threads = bn_block_cfg(groups, lanes);
    CSBSR_REQUIRE(threads <= 1024, "bn_backward: too many channels");
    long long blocks = (m + lanes * 64 - 1) / (static_cast<long long>(lanes) * 64);
    if (blocks > kBnMaxBlocks) blocks = kBnMaxBlocks;
    if (blocks < 1) blocks = 1;
    bn_bwd_reduce_kernel<<<static_cast<int>(blocks), threads, sizeof(float) * lanes * groups * 8, st>>>(
        reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(y_relu), mean, rstd,
        pitch / 8, groups, m, p1, p2);
    bn_bwd_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(p1, p2, static_cast<int>(blocks), c, dbeta, dgamma);
    const long long total = m * (pitch / 8);
    bn_bwd_apply_kernel<<<grid_cap(static_cast<size_t>(total), 256), 256, 0, st>>>(
        reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(y_relu), mean, rstd,
        gamma, dbeta, dgamma, groups, pitch / 8, m, training, reinterpret_cast<uint4*>(dx), reinterpret_cast<uint4*>(dres));
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
