// Training-step glue between the tcgen05 conv kernels: everything the reference's train step
// does between its convolutions -- and autograd does behind them -- as coalesced 128-bit-vectorised kernels on NHWC bf16 maps
// (8 channels per thread access), gather-form backward passes without atomics, fixed-order (bit-reproducible) reductions.
//   bias gradient                      autograd of nn.Conv2d(bias=True)           model/modeling/kbpn.py:266-277
//   ReLU / LeakyReLU backward          kbpn.py:196-214, :513-516 (F.leaky_relu 0.1), extractors.py:62-70
//   residual add / sub, SFT combine    kbpn.py:464-469, :484-489, :516-518 (x * sigmoid(scale) + shift)
//   channel concat / slice             kbpn.py:173-186, pspnet_pytorch/pspnet.py:40
//   bilinear resize backward           pspnet.py:39,56,122 (F.interpolate / F.upsample bilinear)
//   adaptive average pool backward     pspnet.py:32
//   3x3/s2 max pool backward           extractors.py:119
//   Dropout2d                          pspnet.py:67,73,83
//   instance norm forward / backward   model/modeling/build_model.py:135-137
//   border-class expansion             the spatially constant conditioning maps of kbpn.py:404,513-516 (see train_graph.py)
#include <math_constants.h>
#include "common.cuh"
#include "../../include/csbsr_b200.h"

namespace csbsr {

__device__ __forceinline__ void g_unpack8(const uint4& raw, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 g_pack8(const float (&f)[8]) {
    uint4 o;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return o;
}
static inline int glue_grid(size_t total, int threads) {
    size_t b = (total + threads - 1) / threads;
    const size_t cap = static_cast<size_t>(num_sms()) * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

// ------------------------------------------------------------------ column sums (bias gradients), two fixed-order stages
// stage 1: grid (slices, ceil(C/64)); block 256 = 8 channel groups x 32 row lanes; part[slice][c] (fp32)
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, int pitch, int coff, int C, long long rows, float* __restrict__ part,
                      int cpitch) {
    __shared__ float red[32][8][8];
    const int g = threadIdx.x & 7, rl = threadIdx.x >> 3;
    const int c0 = blockIdx.y * 64 + g * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (c0 < C) {
        for (long long r = static_cast<long long>(blockIdx.x) * 32 + rl; r < rows; r += static_cast<long long>(gridDim.x) * 32) {
            float f[8];
            g_unpack8(*reinterpret_cast<const uint4*>(x + r * pitch + coff + c0), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rl][g][j] = acc[j];
    __syncthreads();
    if (threadIdx.x < 64) {
        const int gg = threadIdx.x >> 3, j = threadIdx.x & 7;
        float t = 0.f;
        for (int l = 0; l < 32; ++l) t += red[l][gg][j];
        const int c = blockIdx.y * 64 + threadIdx.x;
        if (c < cpitch) part[static_cast<size_t>(blockIdx.x) * cpitch + c] = t;
    }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int slices, int cpitch, int C, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float t = 0.f;
    for (int s = 0; s < slices; ++s) t += part[static_cast<size_t>(s) * cpitch + c];
    out[c] = t;
}

// ------------------------------------------------------------------ elementwise (contiguous bf16, 8 per thread)
// dx = dy * (y > 0 ? 1 : slope): backward of ReLU (slope 0) / LeakyReLU evaluated on the saved OUTPUT (sign(y) = sign(x))
__global__ void act_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, uint4* __restrict__ dx, size_t n8,
                               float slope) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float a[8], b[8];
        g_unpack8(dy[i], a);
        g_unpack8(y[i], b);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = b[j] > 0.f ? a[j] : a[j] * slope;
        dx[i] = g_pack8(a);
    }
}
// out = alpha * a + beta * b (b may be null); optional ReLU
__global__ void axpby_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, size_t n8,
                             float alpha, float beta, int relu) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float x[8], y[8];
        g_unpack8(a[i], x);
        if (b) {
            g_unpack8(b[i], y);
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = alpha * x[j] + beta * y[j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = alpha * x[j];
        }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fmaxf(x[j], 0.f);
        }
        out[i] = g_pack8(x);
    }
}
// SFT combine: out = f * sigmoid(s) + t
__global__ void sft_fwd_kernel(const uint4* __restrict__ f, const uint4* __restrict__ s, const uint4* __restrict__ t,
                               uint4* __restrict__ out, size_t n8) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float a[8], b[8], c[8];
        g_unpack8(f[i], a);
        g_unpack8(s[i], b);
        g_unpack8(t[i], c);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = a[j] * (1.f / (1.f + __expf(-b[j]))) + c[j];
        out[i] = g_pack8(a);
    }
}
// df = dy * sig(s); ds = dy * f * sig * (1 - sig)   (dt = dy)
__global__ void sft_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ f, const uint4* __restrict__ s,
                               uint4* __restrict__ df, uint4* __restrict__ ds, size_t n8) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float g[8], a[8], b[8], o1[8], o2[8];
        g_unpack8(dy[i], g);
        g_unpack8(f[i], a);
        g_unpack8(s[i], b);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float sg = 1.f / (1.f + __expf(-b[j]));
            o1[j] = g[j] * sg;
            o2[j] = g[j] * a[j] * sg * (1.f - sg);
        }
        df[i] = g_pack8(o1);
        ds[i] = g_pack8(o2);
    }
}
// channel-window copy: dst[row, dcoff + c] = src[row, scoff + c] for c < C (C, pitches, offsets multiples of 8); zero_tail
// additionally clears dst channels [dcoff + C, dcoff + C + zero_tail)
__global__ void window_copy_kernel(const __nv_bfloat16* __restrict__ src, int sp, int so, __nv_bfloat16* __restrict__ dst, int dp,
                                   int dof, int C, int zero_tail, long long rows) {
    const int G = (C + zero_tail) / 8, GC = C / 8;
    const size_t total = static_cast<size_t>(rows) * G;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const size_t r = i / G;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (g < GC) v = *reinterpret_cast<const uint4*>(src + r * sp + so + g * 8);
        *reinterpret_cast<uint4*>(dst + r * dp + dof + g * 8) = v;
    }
}

// ------------------------------------------------------------------ bilinear resize backward (gather form)
__device__ __forceinline__ void gl_bilinear_src(int dst, int in, int out, int align, int& i0, int& i1, float& l1) {
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
// conservative range of outputs whose two source taps can include input index i (tested exactly by the caller)
__device__ __forceinline__ void gl_out_range(int i, int in, int out, int align, int& lo, int& hi) {
    float a, b;
    if (align) {
        const float inv = in > 1 ? static_cast<float>(out - 1) / static_cast<float>(in - 1) : static_cast<float>(out);
        a = (i - 1) * inv;
        b = (i + 1) * inv;
    } else {
        const float inv = static_cast<float>(out) / static_cast<float>(in);
        a = (i - 1 + 0.5f) * inv - 0.5f;
        b = (i + 1 + 0.5f) * inv - 0.5f;
    }
    lo = max(0, static_cast<int>(floorf(a)) - 1);
    hi = min(out - 1, static_cast<int>(ceilf(b)) + 1);
    if (in == 1) { lo = 0; hi = out - 1; }
}
__device__ __forceinline__ float gl_weight(int o, int i, int in, int out, int align) {
    int i0, i1;
    float l1;
    gl_bilinear_src(o, in, out, align, i0, i1, l1);
    return (i0 == i ? 1.f - l1 : 0.f) + (i1 == i ? l1 : 0.f);
}
// dx[n, iy, ix, :] = sum_{oy, ox} wy(oy -> iy) * wx(ox -> ix) * dy[n, oy, ox, :]; the forward's own tap / weight function decides
// membership, so forward and backward are exact transposes.  dy: [N, OH, OW] with (gp, go); dx: [N, H, W] with (xp, xo)
__global__ void bilinear_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N, int H, int W,
                                    int OH, int OW, int C, int gp, int go, int xp, int xo, int align) {
    const int G = C / 8;
    const size_t total = static_cast<size_t>(N) * H * W * G;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const size_t pix = i / G;
        const int ix = static_cast<int>(pix % W);
        const int iy = static_cast<int>((pix / W) % H);
        const int n = static_cast<int>(pix / (static_cast<size_t>(W) * H));
        int ylo, yhi, xlo, xhi;
        gl_out_range(iy, H, OH, align, ylo, yhi);
        gl_out_range(ix, W, OW, align, xlo, xhi);
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const __nv_bfloat16* base = dy + static_cast<size_t>(n) * OH * OW * gp + go + g * 8;
        for (int oy = ylo; oy <= yhi; ++oy) {
            const float wy = gl_weight(oy, iy, H, OH, align);
            if (wy == 0.f) continue;
            for (int ox = xlo; ox <= xhi; ++ox) {
                const float w = wy * gl_weight(ox, ix, W, OW, align);
                if (w == 0.f) continue;
                float f[8];
                g_unpack8(*reinterpret_cast<const uint4*>(base + (static_cast<size_t>(oy) * OW + ox) * gp), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, f[j], acc[j]);
            }
        }
        *reinterpret_cast<uint4*>(dx + pix * xp + xo + g * 8) = g_pack8(acc);
    }
}

// Exact x2 case (align_corners = false, OH = 2H, OW = 2W: the PSPUpsample blocks, pspnet.py:52-57): along each axis input i
// receives from outputs 2i-1 (0.25), 2i (0.75; 1 at i = 0, where the clamped source puts all weight on row 0), 2i+1 (0.75; 1 at
// i = H-1, both taps on the last row) and 2i+2 (0.25) -- the same weights, products and summation order as the generic kernel,
// without its per-candidate weight evaluation and 64-bit index arithmetic.  Block = one input row.
__global__ void __launch_bounds__(256) bilinear_x2_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx,
                                                              int H, int W, int C, int gp, int go, int xp, int xo) {
    const int G = C >> 3;
    const int row = blockIdx.x;                       // n * H + iy
    const int n = row / H, iy = row - n * H;
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    if (e >= W * G) return;
    const int ix = e / G, g = e - ix * G;
    const int OW = 2 * W;
    const __nv_bfloat16* base = dy + static_cast<size_t>(n) * (2 * H) * OW * gp + go + g * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int oy = 2 * iy - 1 + a;
        if (oy < 0 || oy >= 2 * H) continue;
        const float wy = (a == 0 || a == 3) ? 0.25f : ((a == 1 && iy == 0) || (a == 2 && iy == H - 1)) ? 1.f : 0.75f;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int ox = 2 * ix - 1 + b;
            if (ox < 0 || ox >= OW) continue;
            const float wx = (b == 0 || b == 3) ? 0.25f : ((b == 1 && ix == 0) || (b == 2 && ix == W - 1)) ? 1.f : 0.75f;
            const float w = wy * wx;
            float f[8];
            g_unpack8(*reinterpret_cast<const uint4*>(base + (static_cast<size_t>(oy) * OW + ox) * gp), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, f[j], acc[j]);
        }
    }
    *reinterpret_cast<uint4*>(dx + (static_cast<size_t>(row) * W + ix) * xp + xo + g * 8) = g_pack8(acc);
}

// ------------------------------------------------------------------ adaptive average pool backward
// bin o covers [floor(o*H/S), ceil((o+1)*H/S)); dx[ih, iw] = sum over covering bins of dy[bin] / npix(bin)
__global__ void adaptive_avgpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int N, int H,
                                            int W, int S, int C, int gp, int go, int xp, int xo) {
    const int G = C / 8;
    const size_t total = static_cast<size_t>(N) * H * W * G;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const size_t pix = i / G;
        const int iw = static_cast<int>(pix % W);
        const int ih = static_cast<int>((pix / W) % H);
        const int n = static_cast<int>(pix / (static_cast<size_t>(W) * H));
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int oh_c = (ih * S) / H, ow_c = (iw * S) / W;
        for (int oh = max(0, oh_c - 1); oh <= min(S - 1, oh_c + 1); ++oh) {
            const int h0 = (oh * H) / S, h1 = ((oh + 1) * H + S - 1) / S;
            if (ih < h0 || ih >= h1) continue;
            for (int ow = max(0, ow_c - 1); ow <= min(S - 1, ow_c + 1); ++ow) {
                const int w0 = (ow * W) / S, w1 = ((ow + 1) * W + S - 1) / S;
                if (iw < w0 || iw >= w1) continue;
                const float inv = 1.f / static_cast<float>((h1 - h0) * (w1 - w0));
                float f[8];
                g_unpack8(*reinterpret_cast<const uint4*>(dy + ((static_cast<size_t>(n) * S + oh) * S + ow) * gp + go + g * 8), f);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(inv, f[j], acc[j]);
            }
        }
        *reinterpret_cast<uint4*>(dx + pix * xp + xo + g * 8) = g_pack8(acc);
    }
}

// ------------------------------------------------------------------ 3x3 / stride 2 / pad 1 max pool backward
// the gradient of a window goes to its FIRST maximum in scan order (aten's max_pool2d keeps `val > max`): every input
// pixel re-evaluates the (at most four) windows that contain it
__global__ void maxpool3s2_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                      __nv_bfloat16* __restrict__ dx, int N, int H, int W, int OH, int OW, int C, int xp,
                                      int xo, int gp, int go, int dp, int dof) {
    const int G = C / 8;
    const size_t total = static_cast<size_t>(N) * H * W * G;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const size_t pix = i / G;
        const int iw = static_cast<int>(pix % W);
        const int ih = static_cast<int>((pix / W) % H);
        const int n = static_cast<int>(pix / (static_cast<size_t>(W) * H));
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const __nv_bfloat16* xb = x + static_cast<size_t>(n) * H * W * xp + xo + g * 8;
        for (int oh = max(0, (ih - 1 + 1) / 2); oh <= min(OH - 1, (ih + 1) / 2); ++oh) {
            for (int ow = max(0, iw / 2); ow <= min(OW - 1, (iw + 1) / 2); ++ow) {
                // window rows 2*oh-1 .. 2*oh+1, cols 2*ow-1 .. 2*ow+1
                float best[8];
                int arg[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { best[j] = -CUDART_INF_F; arg[j] = -1; }
                for (int r = 0; r < 3; ++r) {
                    const int yy = 2 * oh + r - 1;
                    if (yy < 0 || yy >= H) continue;
                    for (int s = 0; s < 3; ++s) {
                        const int xx = 2 * ow + s - 1;
                        if (xx < 0 || xx >= W) continue;
                        float f[8];
                        g_unpack8(*reinterpret_cast<const uint4*>(xb + (static_cast<size_t>(yy) * W + xx) * xp), f);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (f[j] > best[j] || arg[j] < 0) { best[j] = f[j]; arg[j] = yy * W + xx; }
                    }
                }
                float d[8];
                g_unpack8(*reinterpret_cast<const uint4*>(dy + ((static_cast<size_t>(n) * OH + oh) * OW + ow) * gp + go + g * 8), d);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (arg[j] == ih * W + iw) acc[j] += d[j];
            }
        }
        *reinterpret_cast<uint4*>(dx + pix * dp + dof + g * 8) = g_pack8(acc);
    }
}

// ------------------------------------------------------------------ Dropout2d
__device__ __forceinline__ void gl_philox(unsigned int c[4], unsigned int k0, unsigned int k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned int hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned int hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned int n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// scale[n*Cp + c] = keep ? 1/(1-p) : 0 for c < C, 0 for the padding channels; the step counter lives on the DEVICE so that a
// CUDA-graph replay of the training step draws fresh masks (counter[0] is advanced by counter_inc_kernel once per step)
__global__ void dropout2d_mask_kernel(float* __restrict__ scale, int N, int C, int Cp, float p, unsigned long long seed,
                                      const unsigned long long* __restrict__ counter, unsigned int salt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * Cp) return;
    const int c = i % Cp;
    const unsigned long long step = counter ? counter[0] : 0ull;
    unsigned int ctr[4] = {static_cast<unsigned int>(i), salt, static_cast<unsigned int>(step), static_cast<unsigned int>(step >> 32)};
    gl_philox(ctr, static_cast<unsigned int>(seed), static_cast<unsigned int>(seed >> 32));
    const float u = ctr[0] * (1.0f / 4294967296.0f);
    scale[i] = (c < C && u >= p) ? 1.f / (1.f - p) : 0.f;
}
__global__ void counter_inc_kernel(unsigned long long* counter) { counter[0] += 1ull; }
// y[n, pix, c] = x[n, pix, c] * scale[n, c]  (forward and backward)
__global__ void channel_scale_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ scale,
                                     __nv_bfloat16* __restrict__ y, int N, long long HW, int Cp) {
    const int G = Cp / 8;
    const size_t total = static_cast<size_t>(N) * HW * G;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const size_t pix = i / G;
        const int n = static_cast<int>(pix / HW);
        float f[8];
        g_unpack8(*reinterpret_cast<const uint4*>(x + pix * Cp + g * 8), f);
        const float4 s0 = *reinterpret_cast<const float4*>(scale + static_cast<size_t>(n) * Cp + g * 8);
        const float4 s1 = *reinterpret_cast<const float4*>(scale + static_cast<size_t>(n) * Cp + g * 8 + 4);
        f[0] *= s0.x; f[1] *= s0.y; f[2] *= s0.z; f[3] *= s0.w;
        f[4] *= s1.x; f[5] *= s1.y; f[6] *= s1.z; f[7] *= s1.w;
        *reinterpret_cast<uint4*>(y + pix * Cp + g * 8) = g_pack8(f);
    }
}

// ------------------------------------------------------------------ border-class expansion (see train_graph._expand_classes)
__device__ __forceinline__ int gl_class(int i, int n, int bw) { return i < bw ? i : (i >= n - bw ? 2 * bw - (n - 1 - i) : bw); }
// out[n, y, x, :] = small[n, cls(y), cls(x), :]     (small: [N, K, K, Cp], K = 2*bw + 1)
__global__ void expand_classes_kernel(const __nv_bfloat16* __restrict__ small, __nv_bfloat16* __restrict__ out, int N, int H, int W,
                                      int bw, int Cp) {
    const int G = Cp / 8, K = 2 * bw + 1;
    const size_t total = static_cast<size_t>(N) * H * W * G;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const size_t pix = i / G;
        const int x = static_cast<int>(pix % W);
        const int y = static_cast<int>((pix / W) % H);
        const int n = static_cast<int>(pix / (static_cast<size_t>(W) * H));
        *reinterpret_cast<uint4*>(out + pix * Cp + g * 8) =
            *reinterpret_cast<const uint4*>(small + ((static_cast<size_t>(n) * K + gl_class(y, H, bw)) * K + gl_class(x, W, bw)) * Cp + g * 8);
    }
}
// backward stage 1: per (n, image row y, 64-channel slab): sums of dy over the pixels of each COLUMN class -> part[n][y][K][Cp]
__global__ void __launch_bounds__(256)
expand_classes_bwd_rows_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ part, int H, int W, int bw, int Cp) {
    __shared__ float red[32][8][8];
    const int K = 2 * bw + 1;
    const int y = blockIdx.x, n = blockIdx.y;
    const int g = threadIdx.x & 7, xl = threadIdx.x >> 3;
    const int c0 = blockIdx.z * 64 + g * 8;
    const __nv_bfloat16* row = dy + (static_cast<size_t>(n) * H + y) * W * Cp + c0;
    float* prow = part + ((static_cast<size_t>(n) * H + y) * K) * Cp + c0;
    // border columns: one pixel each
    if (xl < 2 * bw && xl < W) {
        const int x = xl < bw ? xl : W - (2 * bw - xl);
        if (!(xl >= bw && x < bw)) {                       // tiny widths: do not count a pixel twice
            float f[8];
            g_unpack8(*reinterpret_cast<const uint4*>(row + static_cast<size_t>(x) * Cp), f);
            const int cls = gl_class(x, W, bw);
#pragma unroll
            for (int j = 0; j < 8; ++j) prow[static_cast<size_t>(cls) * Cp + j] = f[j];
        }
    }
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int x = bw + xl; x < W - bw; x += 32) {
        float f[8];
        g_unpack8(*reinterpret_cast<const uint4*>(row + static_cast<size_t>(x) * Cp), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) red[xl][g][j] = acc[j];
    __syncthreads();
    if (threadIdx.x < 64) {
        const int gg = threadIdx.x >> 3, j = threadIdx.x & 7;
        float t = 0.f;
        for (int l = 0; l < 32; ++l) t += red[l][gg][j];
        part[((static_cast<size_t>(n) * H + y) * K + bw) * Cp + blockIdx.z * 64 + gg * 8 + j] = t;
    }
}
// backward stage 2: dsmall[n][cy][cx][c] = sum over rows y of class cy of part[n][y][cx][c]   (fixed order)
__global__ void expand_classes_bwd_final_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ dsmall, int N, int H,
                                                int bw, int Cp) {
    const int K = 2 * bw + 1;
    const size_t total = static_cast<size_t>(N) * K * K * Cp;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(i % Cp);
        const int cx = static_cast<int>((i / Cp) % K);
        const int cy = static_cast<int>((i / (static_cast<size_t>(Cp) * K)) % K);
        const int n = static_cast<int>(i / (static_cast<size_t>(Cp) * K * K));
        float t = 0.f;
        int y0, y1;
        if (cy < bw) { y0 = cy; y1 = cy + 1; }
        else if (cy == bw) { y0 = bw; y1 = H - bw; }
        else { y0 = H - 1 - (2 * bw - cy); y1 = y0 + 1; }
        for (int y = y0; y < y1; ++y) t += part[((static_cast<size_t>(n) * H + y) * K + cx) * Cp + c];
        dsmall[i] = __float2bfloat16(t);
    }
}

// ------------------------------------------------------------------ instance norm (fp32 NCHW planes)
__global__ void instnorm_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                      float* __restrict__ y, long long HW) {
    const int nc = blockIdx.y;
    const float m = mean[nc], r = rstd[nc];
    const float* xp = x + static_cast<size_t>(nc) * HW;
    float* yp = y + static_cast<size_t>(nc) * HW;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < HW; i += static_cast<long long>(gridDim.x) * blockDim.x)
        yp[i] = (xp[i] - m) * r;
}
// part[(nc*slices + slice)*2 + {0,1}] = sum dy, sum dy * xhat  (fp64, fixed order)
__global__ void instnorm_bwd_partial_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                            const float* __restrict__ rstd, double* __restrict__ part, long long HW) {
    const int nc = blockIdx.y;
    const float m = mean[nc], r = rstd[nc];
    double s = 0.0, sx = 0.0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < HW; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float g = dy[static_cast<size_t>(nc) * HW + i];
        s += g;
        sx += static_cast<double>(g) * ((x[static_cast<size_t>(nc) * HW + i] - m) * r);
    }
    __shared__ double red[32];
    s = block_sum_det(s, red);
    sx = block_sum_det(sx, red);
    if (threadIdx.x == 0) {
        part[(static_cast<size_t>(nc) * gridDim.x + blockIdx.x) * 2] = s;
        part[(static_cast<size_t>(nc) * gridDim.x + blockIdx.x) * 2 + 1] = sx;
    }
}
// dx = rstd * (dy - mean(dy) - xhat * mean(dy * xhat))
__global__ void instnorm_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                          const float* __restrict__ rstd, const double* __restrict__ part, int slices,
                                          float* __restrict__ dx, long long HW) {
    const int nc = blockIdx.y;
    __shared__ float sm[2];
    if (threadIdx.x == 0) {
        double s = 0.0, sx = 0.0;
        for (int k = 0; k < slices; ++k) {
            s += part[(static_cast<size_t>(nc) * slices + k) * 2];
            sx += part[(static_cast<size_t>(nc) * slices + k) * 2 + 1];
        }
        sm[0] = static_cast<float>(s / HW);
        sm[1] = static_cast<float>(sx / HW);
    }
    __syncthreads();
    const float m = mean[nc], r = rstd[nc], mg = sm[0], mgx = sm[1];
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < HW; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const size_t o = static_cast<size_t>(nc) * HW + i;
        const float xh = (x[o] - m) * r;
        dx[o] = r * (dy[o] - mg - xh * mgx);
    }
}

// ------------------------------------------------------------------ layout conversion NHWC bf16 -> NCHW fp32 (first C channels)
__global__ void nhwc_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int N, long long HW, int C, int xp,
                                        int xo) {
    const size_t total = static_cast<size_t>(N) * C * HW;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const long long p = static_cast<long long>(i % HW);
        const int c = static_cast<int>((i / HW) % C);
        const int n = static_cast<int>(i / (static_cast<size_t>(HW) * C));
        y[i] = __bfloat162float(x[(static_cast<size_t>(n) * HW + p) * xp + xo + c]);
    }
}

// ------------------------------------------------------------------ tap-expanded 3x3 conv with <= 4 outputs (kbpn.py:361, :68)
// z[pix, t*cp + m] = tap t of output m evaluated AT pix (a 1x1 GEMM over the input channels); the conv result gathers the nine
// shifted taps: y[n, h, w, m] = sum_t z[n, h + t/3 - 1, w + t%3 - 1, t*cp + m] (zero outside); channels [co, yp) are zeroed
__global__ void tapexp_gather_kernel(const __nv_bfloat16* __restrict__ z, int zp, __nv_bfloat16* __restrict__ y, int yp, int N,
                                     int H, int W, int cp, int co) {
    const size_t total = static_cast<size_t>(N) * H * W;
    for (size_t pix = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; pix < total; pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int w = static_cast<int>(pix % W);
        const int h = static_cast<int>((pix / W) % H);
        const size_t nb = pix - static_cast<size_t>(h) * W - w;          // first pixel of the image
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
            if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
            const __nv_bfloat16* src = z + (nb + static_cast<size_t>(hh) * W + ww) * zp + t * cp;
            for (int m = 0; m < co; ++m) acc[m] += __bfloat162float(src[m]);
        }
        for (int m = co; m < 8; ++m) acc[m] = 0.f;
        uint4* dst = reinterpret_cast<uint4*>(y + pix * yp);
        dst[0] = g_pack8(acc);
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
        for (int g = 1; g < yp / 8; ++g) dst[g] = zero;
    }
}
// its transpose (the gradient w.r.t. z): dz[n, h, w, t*cp + m] = dy[n, h - (t/3 - 1), w - (t%3 - 1), m]; channels >= 9*cp zero
__global__ void tapexp_scatter_kernel(const __nv_bfloat16* __restrict__ dy, int dyp, __nv_bfloat16* __restrict__ dz, int zp, int N,
                                      int H, int W, int cp, int co) {
    const int G = zp / 8;
    const size_t total = static_cast<size_t>(N) * H * W * G;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % G);
        const size_t pix = i / G;
        const int w = static_cast<int>(pix % W);
        const int h = static_cast<int>((pix / W) % H);
        const size_t nb = pix - static_cast<size_t>(h) * W - w;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = g * 8 + j, t = e / cp, m = e - t * cp;
            float val = 0.f;
            if (t < 9 && m < co) {
                const int hh = h - (t / 3 - 1), ww = w - (t % 3 - 1);
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = __bfloat162float(dy[(nb + static_cast<size_t>(hh) * W + ww) * dyp + m]);
            }
            v[j] = val;
        }
        *reinterpret_cast<uint4*>(dz + pix * zp + g * 8) = g_pack8(v);
    }
}

// ------------------------------------------------------------------ input pipeline
// SplitPatch / JointPatch (model/data/samplers/patch_sampler.py:15-50): an image [C, H, W] <-> its non-overlapping ph x pw
// patches [(iy * nx + ix), C, ph, pw] (unfold with stride = size; remainders are dropped exactly like Tensor.unfold).
// join == 0: img -> patches;  join == 1: patches -> img  (B images: patch index b * ny * nx + iy * nx + ix)
__global__ void patch_split_join_kernel(float* __restrict__ img, float* __restrict__ patches, int B, int C, int H, int W, int ph,
                                        int pw, int ny, int nx, int join) {
    const size_t total = static_cast<size_t>(B) * ny * nx * C * ph * pw;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % pw);
        const int y = static_cast<int>((i / pw) % ph);
        const int c = static_cast<int>((i / (static_cast<size_t>(pw) * ph)) % C);
        const size_t pidx = i / (static_cast<size_t>(pw) * ph * C);
        const int ix = static_cast<int>(pidx % nx), iy = static_cast<int>((pidx / nx) % ny);
        const size_t b = pidx / (static_cast<size_t>(nx) * ny);
        const size_t io = ((b * C + c) * H + iy * ph + y) * W + ix * pw + x;
        if (join) img[io] = patches[i]; else patches[i] = img[io];
    }
}
// Training augmentation of CrackDataSet.__getitem__ (crack_dataset.py:42-48 with data_preprocess.py:13-46): ConvertFromInts,
// RandomMirror, RandomVerticalFlip, RandomCrop, ToTensor, /255 in one pass over a decoded uint8 HWC image.
// prm[b] = (y0, x0, hflip, vflip) in the coordinates of the FLIPPED image; out fp32 [B, C, th, tw]
__global__ void crop_flip_u8_kernel(const unsigned char* const* __restrict__ imgs, const int* __restrict__ dims /* [B][3]: H, W, C */,
                                    const int* __restrict__ prm, float* __restrict__ out, int B, int Cout, int th, int tw, float divisor) {
    const size_t total = static_cast<size_t>(B) * Cout * th * tw;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % tw);
        const int y = static_cast<int>((i / tw) % th);
        const int c = static_cast<int>((i / (static_cast<size_t>(tw) * th)) % Cout);
        const int b = static_cast<int>(i / (static_cast<size_t>(tw) * th * Cout));
        const int H = dims[b * 3], W = dims[b * 3 + 1], C = dims[b * 3 + 2];
        int sy = prm[b * 4] + y, sx = prm[b * 4 + 1] + x;
        if (prm[b * 4 + 2]) sx = W - 1 - sx;               // RandomMirror: img[:, ::-1]
        if (prm[b * 4 + 3]) sy = H - 1 - sy;               // RandomVerticalFlip: img[::-1]
        out[i] = static_cast<float>(imgs[b][(static_cast<size_t>(sy) * W + sx) * C + min(c, C - 1)]) / divisor;   // true division: bit-equal to `x / 255`
    }
}

}  // namespace csbsr

using namespace csbsr;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define ALIGNED16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

static constexpr int kColsumMaxSlices = 128;
extern "C" size_t csbsr_colsum_workspace_bytes(int c) { return sizeof(float) * kColsumMaxSlices * static_cast<size_t>((c + 63) / 64 * 64); }

extern "C" int csbsr_bias_grad(const void* dy, int pitch, int coff, int c, long long rows, float* out, void* workspace,
                               size_t workspace_bytes, void* stream) {
    CSBSR_REQUIRE(dy && out && workspace && c > 0 && rows > 0 && pitch % 8 == 0 && coff % 8 == 0, "bias_grad: bad arguments");
    const int cp = (c + 63) / 64 * 64;
    CSBSR_REQUIRE(coff + (c + 7) / 8 * 8 <= pitch, "bias_grad: channel window exceeds the pitch");
    CSBSR_REQUIRE(workspace_bytes >= csbsr_colsum_workspace_bytes(c), "bias_grad: workspace too small");
    int slices = static_cast<int>((rows + 32 * 8 - 1) / (32 * 8));
    if (slices > kColsumMaxSlices) slices = kColsumMaxSlices;
    if (slices < 1) slices = 1;
    const int cvec = (c + 7) / 8 * 8;                                  // channels read in whole 8-vectors
    colsum_partial_kernel<<<dim3(slices, cp / 64), 256, 0, STREAM(stream)>>>(CBF(dy), pitch, coff, cvec, rows,
                                                                            static_cast<float*>(workspace), cp);
    colsum_final_kernel<<<(c + 127) / 128, 128, 0, STREAM(stream)>>>(static_cast<const float*>(workspace), slices, cp, c, out);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_act_bwd(const void* dy, const void* y, void* dx, long long n, float slope, void* stream) {
    CSBSR_REQUIRE(dy && y && dx && n > 0 && n % 8 == 0 && ALIGNED16(dy) && ALIGNED16(y) && ALIGNED16(dx), "act_bwd: bad arguments");
    act_bwd_kernel<<<glue_grid(n / 8, 256), 256, 0, STREAM(stream)>>>(static_cast<const uint4*>(dy), static_cast<const uint4*>(y),
                                                                     static_cast<uint4*>(dx), n / 8, slope);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_axpby(const void* a, const void* b, void* out, long long n, float alpha, float beta, int relu, void* stream) {
    CSBSR_REQUIRE(a && out && n > 0 && n % 8 == 0 && ALIGNED16(a) && ALIGNED16(out) && (!b || ALIGNED16(b)), "axpby: bad arguments");
    axpby_kernel<<<glue_grid(n / 8, 256), 256, 0, STREAM(stream)>>>(static_cast<const uint4*>(a), static_cast<const uint4*>(b),
                                                                   static_cast<uint4*>(out), n / 8, alpha, beta, relu);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_sft_combine(const void* f, const void* s, const void* t, void* out, long long n, void* stream) {
    CSBSR_REQUIRE(f && s && t && out && n > 0 && n % 8 == 0, "sft_combine: bad arguments");
    sft_fwd_kernel<<<glue_grid(n / 8, 256), 256, 0, STREAM(stream)>>>(static_cast<const uint4*>(f), static_cast<const uint4*>(s),
                                                                     static_cast<const uint4*>(t), static_cast<uint4*>(out), n / 8);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_sft_combine_bwd(const void* dy, const void* f, const void* s, void* df, void* ds, long long n, void* stream) {
    CSBSR_REQUIRE(dy && f && s && df && ds && n > 0 && n % 8 == 0, "sft_combine_bwd: bad arguments");
    sft_bwd_kernel<<<glue_grid(n / 8, 256), 256, 0, STREAM(stream)>>>(static_cast<const uint4*>(dy), static_cast<const uint4*>(f),
                                                                     static_cast<const uint4*>(s), static_cast<uint4*>(df),
                                                                     static_cast<uint4*>(ds), n / 8);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_window_copy(const void* src, int src_pitch, int src_coff, void* dst, int dst_pitch, int dst_coff, int c,
                                 int zero_tail, long long rows, void* stream) {
    CSBSR_REQUIRE(src && dst && rows > 0 && c > 0 && c % 8 == 0 && zero_tail % 8 == 0 && src_pitch % 8 == 0 && dst_pitch % 8 == 0 &&
                      src_coff % 8 == 0 && dst_coff % 8 == 0 && src_coff + c <= src_pitch && dst_coff + c + zero_tail <= dst_pitch,
                  "window_copy: channel windows must be multiples of 8 inside their pitches");
    window_copy_kernel<<<glue_grid(static_cast<size_t>(rows) * ((c + zero_tail) / 8), 256), 256, 0, STREAM(stream)>>>(
        CBF(src), src_pitch, src_coff, BF(dst), dst_pitch, dst_coff, c, zero_tail, rows);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_bilinear_nhwc_bwd(const void* dy, void* dx, int n, int h, int w, int oh, int ow, int c, int dy_pitch,
                                       int dy_coff, int dx_pitch, int dx_coff, int align_corners, void* stream) {
    CSBSR_REQUIRE(dy && dx && c % 8 == 0 && dy_pitch % 8 == 0 && dx_pitch % 8 == 0 && dy_coff % 8 == 0 && dx_coff % 8 == 0,
                  "bilinear_nhwc_bwd: channel counts/offsets must be multiples of 8");
    const size_t total = static_cast<size_t>(n) * h * w * (c / 8);
    if (!align_corners && oh == 2 * h && ow == 2 * w && h > 1 && w > 1) {
        const dim3 grid(static_cast<unsigned>(n) * h, (static_cast<unsigned>(w) * (c / 8) + 255) / 256);
        bilinear_x2_bwd_kernel<<<grid, 256, 0, STREAM(stream)>>>(CBF(dy), BF(dx), h, w, c, dy_pitch, dy_coff, dx_pitch, dx_coff);
        CSBSR_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    bilinear_bwd_kernel<<<glue_grid(total, 256), 256, 0, STREAM(stream)>>>(CBF(dy), BF(dx), n, h, w, oh, ow, c, dy_pitch, dy_coff,
                                                                          dx_pitch, dx_coff, align_corners);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_adaptive_avgpool_nhwc_bwd(const void* dy, void* dx, int n, int h, int w, int s, int c, int dy_pitch,
                                               int dy_coff, int dx_pitch, int dx_coff, void* stream) {
    CSBSR_REQUIRE(dy && dx && s > 0 && c % 8 == 0 && dy_pitch % 8 == 0 && dx_pitch % 8 == 0 && dy_coff % 8 == 0 && dx_coff % 8 == 0,
                  "adaptive_avgpool_bwd: bad arguments");
    const size_t total = static_cast<size_t>(n) * h * w * (c / 8);
    adaptive_avgpool_bwd_kernel<<<glue_grid(total, 256), 256, 0, STREAM(stream)>>>(CBF(dy), BF(dx), n, h, w, s, c, dy_pitch,
                                                                                  dy_coff, dx_pitch, dx_coff);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_maxpool3s2_nhwc_bwd(const void* x, const void* dy, void* dx, int n, int h, int w, int c, int x_pitch,
                                         int x_coff, int dy_pitch, int dy_coff, int dx_pitch, int dx_coff, void* stream) {
    CSBSR_REQUIRE(x && dy && dx && c % 8 == 0 && x_pitch % 8 == 0 && dy_pitch % 8 == 0 && dx_pitch % 8 == 0 && x_coff % 8 == 0 &&
                      dy_coff % 8 == 0 && dx_coff % 8 == 0, "maxpool_bwd: channel counts/offsets must be multiples of 8");
    const int oh = (h + 2 - 3) / 2 + 1, ow = (w + 2 - 3) / 2 + 1;
    const size_t total = static_cast<size_t>(n) * h * w * (c / 8);
    maxpool3s2_bwd_kernel<<<glue_grid(total, 256), 256, 0, STREAM(stream)>>>(CBF(x), CBF(dy), BF(dx), n, h, w, oh, ow, c, x_pitch,
                                                                            x_coff, dy_pitch, dy_coff, dx_pitch, dx_coff);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_dropout2d_mask(float* scale, int n, int c, int c_pad, float p, unsigned long long seed,
                                    const unsigned long long* counter, unsigned int salt, void* stream) {
    CSBSR_REQUIRE(scale && n > 0 && c > 0 && c_pad >= c && c_pad % 8 == 0 && p >= 0.f && p < 1.f, "dropout2d_mask: bad arguments");
    dropout2d_mask_kernel<<<(n * c_pad + 127) / 128, 128, 0, STREAM(stream)>>>(scale, n, c, c_pad, p, seed, counter, salt);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_counter_inc(unsigned long long* counter, void* stream) {
    CSBSR_REQUIRE(counter, "counter_inc: null pointer");
    counter_inc_kernel<<<1, 1, 0, STREAM(stream)>>>(counter);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_channel_scale(const void* x, const float* scale, void* y, int n, long long hw, int c_pad, void* stream) {
    CSBSR_REQUIRE(x && scale && y && n > 0 && hw > 0 && c_pad % 8 == 0, "channel_scale: bad arguments");
    channel_scale_kernel<<<glue_grid(static_cast<size_t>(n) * hw * (c_pad / 8), 256), 256, 0, STREAM(stream)>>>(CBF(x), scale, BF(y), n,
                                                                                                            hw, c_pad);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_expand_classes(const void* small, void* out, int n, int h, int w, int bw, int c_pad, void* stream) {
    CSBSR_REQUIRE(small && out && n > 0 && bw >= 1 && h > 2 * bw && w > 2 * bw && c_pad % 8 == 0, "expand_classes: bad arguments");
    expand_classes_kernel<<<glue_grid(static_cast<size_t>(n) * h * w * (c_pad / 8), 256), 256, 0, STREAM(stream)>>>(CBF(small), BF(out),
                                                                                                                n, h, w, bw, c_pad);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" size_t csbsr_expand_classes_workspace_bytes(int n, int h, int bw, int c_pad) {
    return sizeof(float) * static_cast<size_t>(n) * h * (2 * bw + 1) * c_pad;
}

extern "C" int csbsr_expand_classes_bwd(const void* dy, void* dsmall, int n, int h, int w, int bw, int c_pad, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    CSBSR_REQUIRE(dy && dsmall && workspace && n > 0 && bw >= 1 && bw <= 16 && h > 2 * bw && w > 2 * bw && c_pad % 64 == 0,
                  "expand_classes_bwd: bad arguments");
    CSBSR_REQUIRE(workspace_bytes >= csbsr_expand_classes_workspace_bytes(n, h, bw, c_pad), "expand_classes_bwd: workspace too small");
    float* part = static_cast<float*>(workspace);
    expand_classes_bwd_rows_kernel<<<dim3(h, n, c_pad / 64), 256, 0, STREAM(stream)>>>(CBF(dy), part, h, w, bw, c_pad);
    const size_t total = static_cast<size_t>(n) * (2 * bw + 1) * (2 * bw + 1) * c_pad;
    expand_classes_bwd_final_kernel<<<glue_grid(total, 256), 256, 0, STREAM(stream)>>>(part, BF(dsmall), n, h, bw, c_pad);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static constexpr int kInstSlices = 64;
extern "C" size_t csbsr_instnorm_bwd_workspace_bytes(int nc) { return sizeof(double) * 2 * kInstSlices * static_cast<size_t>(nc); }

extern "C" int csbsr_instnorm_apply(const float* x, const float* mean, const float* rstd, float* y, int nc, long long hw, void* stream) {
    CSBSR_REQUIRE(x && mean && rstd && y && nc > 0 && hw > 0, "instnorm_apply: bad arguments");
    int slices = static_cast<int>((hw + 256 * 8 - 1) / (256 * 8));
    if (slices > 256) slices = 256;
    instnorm_apply_kernel<<<dim3(slices, nc), 256, 0, STREAM(stream)>>>(x, mean, rstd, y, hw);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_instnorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, float* dx, int nc,
                                  long long hw, void* workspace, size_t workspace_bytes, void* stream) {
    CSBSR_REQUIRE(dy && x && mean && rstd && dx && workspace && nc > 0 && hw > 0, "instnorm_bwd: bad arguments");
    CSBSR_REQUIRE(workspace_bytes >= csbsr_instnorm_bwd_workspace_bytes(nc), "instnorm_bwd: workspace too small");
    int slices = static_cast<int>((hw + 256 * 16 - 1) / (256 * 16));
    if (slices < 1) slices = 1;
    if (slices > kInstSlices) slices = kInstSlices;
    double* part = static_cast<double*>(workspace);
    instnorm_bwd_partial_kernel<<<dim3(slices, nc), 256, 0, STREAM(stream)>>>(dy, x, mean, rstd, part, hw);
    int aslices = static_cast<int>((hw + 256 * 8 - 1) / (256 * 8));
    if (aslices > 256) aslices = 256;
    instnorm_bwd_apply_kernel<<<dim3(aslices, nc), 256, 0, STREAM(stream)>>>(dy, x, mean, rstd, part, slices, dx, hw);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_nhwc_bf16_to_nchw_f32(const void* x, float* y, int n, long long hw, int c, int x_pitch, int x_coff, void* stream) {
    CSBSR_REQUIRE(x && y && n > 0 && hw > 0 && c > 0 && x_coff + c <= x_pitch, "nhwc_bf16_to_nchw_f32: bad arguments");
    nhwc_to_nchw_f32_kernel<<<glue_grid(static_cast<size_t>(n) * c * hw, 256), 256, 0, STREAM(stream)>>>(CBF(x), y, n, hw, c, x_pitch, x_coff);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_tapexp_gather_nhwc(const void* z, int z_pitch, void* y, int y_pitch, int n, int h, int w, int cp, int co, void* stream) {
    CSBSR_REQUIRE(z && y && n > 0 && co >= 1 && co <= cp && cp <= 8 && 9 * cp <= z_pitch && z_pitch % 8 == 0 && y_pitch % 8 == 0 && co <= 8,
                  "tapexp_gather: bad arguments");
    tapexp_gather_kernel<<<glue_grid(static_cast<size_t>(n) * h * w, 256), 256, 0, STREAM(stream)>>>(CBF(z), z_pitch, BF(y), y_pitch, n, h, w, cp, co);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_tapexp_scatter_nhwc(const void* dy, int dy_pitch, void* dz, int z_pitch, int n, int h, int w, int cp, int co, void* stream) {
    CSBSR_REQUIRE(dy && dz && n > 0 && co >= 1 && co <= cp && 9 * cp <= z_pitch && z_pitch % 8 == 0 && co <= dy_pitch, "tapexp_scatter: bad arguments");
    tapexp_scatter_kernel<<<glue_grid(static_cast<size_t>(n) * h * w * (z_pitch / 8), 256), 256, 0, STREAM(stream)>>>(CBF(dy), dy_pitch, BF(dz),
                                                                                                                  z_pitch, n, h, w, cp, co);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_patch_split(const float* img, float* patches, int b, int c, int h, int w, int ph, int pw, void* stream) {
    CSBSR_REQUIRE(img && patches && b > 0 && c > 0 && ph > 0 && pw > 0 && ph <= h && pw <= w, "patch_split: bad arguments");
    const int ny = h / ph, nx = w / pw;
    const size_t total = static_cast<size_t>(b) * ny * nx * c * ph * pw;
    patch_split_join_kernel<<<glue_grid(total, 256), 256, 0, STREAM(stream)>>>(const_cast<float*>(img), patches, b, c, h, w, ph, pw, ny, nx, 0);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_patch_join(const float* patches, float* img, int b, int c, int h, int w, int ph, int pw, void* stream) {
    CSBSR_REQUIRE(img && patches && b > 0 && c > 0 && ph > 0 && pw > 0 && h % ph == 0 && w % pw == 0, "patch_join: the image must be a whole number of patches");
    const int ny = h / ph, nx = w / pw;
    const size_t total = static_cast<size_t>(b) * ny * nx * c * ph * pw;
    patch_split_join_kernel<<<glue_grid(total, 256), 256, 0, STREAM(stream)>>>(img, const_cast<float*>(patches), b, c, h, w, ph, pw, ny, nx, 1);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int csbsr_crop_flip_u8(const unsigned char* const* imgs, const int* dims, const int* params, float* out, int b, int c_out,
                                  int th, int tw, float divisor, void* stream) {
    CSBSR_REQUIRE(imgs && dims && params && out && b > 0 && c_out > 0 && th > 0 && tw > 0, "crop_flip_u8: bad arguments");
    CSBSR_REQUIRE(divisor != 0.f, "crop_flip_u8: zero divisor");
    const size_t total = static_cast<size_t>(b) * c_out * th * tw;
    crop_flip_u8_kernel<<<glue_grid(total, 256), 256, 0, STREAM(stream)>>>(imgs, dims, params, out, b, c_out, th, tw, divisor);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
