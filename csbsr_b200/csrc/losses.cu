// Joint-training losses of the reference on the device (forward values; gradients w.r.t. the predictions where noted):
//   L3  compute_sdf1_1 / BoundaryLoss          model/utils/boundary_loss.py:26-67   (exact EDT, fp64 normalisation)
//   L2  BoundaryComboLoss = alpha*(WBCE+Dice)/2 + (1-alpha)*Boundary   model/utils/loss_functions.py:49-74,196-210,284-345
//   L4  SegmentFailerOrientedExpWeight w^F     model/utils/oriented_weight.py:73-83, applied in build_model.py:422-438
//   L1  KBPNLoss (HR L1 + pseudo-LR L1 + 0 * kernel MSE)               model/utils/sr_loss_functions.py:39-56,84-102
// The reference runs the SDF on the host every step (D2H, two scipy EDTs, skimage, H2D); here it stays on the device.
#include <math_constants.h>
#include "common.cuh"
#include "../../include/csbsr_b200.h"

namespace csbsr {

static constexpr unsigned short kLInf16 = 0xFFFFu;
static constexpr unsigned int kLInfD2 = 0xFFFFFFFFu;

// ------------------------------------------------------------------ SDF
// fg[b][y][x] = (uint8)(mask) != 0   (boundary_loss.py:51,57: astype(np.uint8) truncation, then astype(bool))
__global__ void sdf_binarise_kernel(const float* __restrict__ mask, unsigned char* __restrict__ fg, int* __restrict__ any_fg,
                                    int HW) {
    const int b = blockIdx.y;
    int local = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const float m = mask[static_cast<size_t>(b) * HW + i];
        // float -> uint8 truncates toward zero; values in (-1, 1) become 0 (values outside [0, 256) are not produced by masks)
        const int v = static_cast<int>(m);
        const unsigned char f = (static_cast<unsigned char>(v) != 0) ? 1 : 0;
        fg[static_cast<size_t>(b) * HW + i] = f;
        local |= f;
    }
    if (__any_sync(0xffffffffu, local) && (threadIdx.x & 31) == 0) atomicOr(&any_fg[b], 1);
}

// vertical distance to the nearest pixel with fg == want (column pass of the exact EDT); grid.y = b*2 + which
__global__ void sdf_column_kernel(const unsigned char* __restrict__ fg, unsigned short* __restrict__ g, int H, int W) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const int b = blockIdx.y >> 1, want = blockIdx.y & 1;          // want = 0: distance to background (posdis), 1: to foreground
    const unsigned char* f = fg + static_cast<size_t>(b) * H * W;
    unsigned short* gp = g + static_cast<size_t>(blockIdx.y) * H * W;
    int last = -1;
    for (int y = 0; y < H; ++y) {
        if (f[y * W + x] == want) last = y;
        gp[y * W + x] = last < 0 ? kLInf16 : static_cast<unsigned short>(y - last);
    }
    last = -1;
    for (int y = H - 1; y >= 0; --y) {
        const unsigned short cur = gp[y * W + x];
        if (cur == 0) last = y;
        else if (last >= 0) {
            const int dd = last - y;
            if (cur == kLInf16 || dd < cur) gp[y * W + x] = static_cast<unsigned short>(dd);
        }
    }
}

__device__ __forceinline__ unsigned int sdf_row_search(const unsigned short* __restrict__ grow, int x, int W) {
    unsigned int best = kLInfD2;
    const int maxd = max(x, W - 1 - x);
    for (int d = 0; d <= maxd; ++d) {
        const unsigned int dd = static_cast<unsigned int>(d) * d;
        if (dd >= best) break;
        if (x - d >= 0) {
            const unsigned int gv = grow[x - d];
            if (gv != kLInf16) best = min(best, dd + gv * gv);
        }
        if (d > 0 && x + d < W) {
            const unsigned int gv = grow[x + d];
            if (gv != kLInf16) best = min(best, dd + gv * gv);
        }
    }
    return best;
}

// exact squared distances + per-(sample, which) min / max
__global__ void sdf_row_kernel(const unsigned short* __restrict__ g, unsigned int* __restrict__ d2, unsigned int* __restrict__ mm,
                               int H, int W) {
    const int bw = blockIdx.y;
    const unsigned short* gp = g + static_cast<size_t>(bw) * H * W;
    unsigned int* op = d2 + static_cast<size_t>(bw) * H * W;
    unsigned int lo = 0xFFFFFFFFu, hi = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i % W;
        const unsigned int v = sdf_row_search(gp + static_cast<size_t>(y) * W, x, W);
        op[i] = v;
        lo = min(lo, v);
        hi = max(hi, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[bw * 2], lo);
        atomicMax(&mm[bw * 2 + 1], hi);
    }
}

// sdf = (negdis - min)/(max - min) - (posdis - min)/(max - min) in fp64, 0 on the inner boundary
// (4-connected foreground pixels touching background, reflect border), zeros when the mask is empty
__global__ void sdf_finish_kernel(const unsigned char* __restrict__ fg, const unsigned int* __restrict__ d2,
                                  const unsigned int* __restrict__ mm, const int* __restrict__ any_fg,
                                  float* __restrict__ sdf, int H, int W) {
    const int b = blockIdx.y;
    const size_t HW = static_cast<size_t>(H) * W;
    const unsigned char* f = fg + b * HW;
    const unsigned int* dpos = d2 + (static_cast<size_t>(b) * 2) * HW;        // distance to background
    const unsigned int* dneg = d2 + (static_cast<size_t>(b) * 2 + 1) * HW;    // distance to foreground
    const bool any = any_fg[b] != 0;
    auto dist = [](unsigned int v) { return v == kLInfD2 ? CUDART_INF : sqrt(static_cast<double>(v)); };
    const double pmin = dist(mm[(b * 2) * 2]), pmax = dist(mm[(b * 2) * 2 + 1]);
    const double nmin = dist(mm[(b * 2 + 1) * 2]), nmax = dist(mm[(b * 2 + 1) * 2 + 1]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H * W; i += gridDim.x * blockDim.x) {
        float out = 0.f;
        if (any) {
            const int y = i / W, x = i % W;
            bool boundary = false;
            if (f[i]) {
                boundary = (y > 0 && !f[i - W]) || (y < H - 1 && !f[i + W]) || (x > 0 && !f[i - 1]) || (x < W - 1 && !f[i + 1]);
            }
            if (!boundary) {
                const double v = (dist(dneg[i]) - nmin) / (nmax - nmin) - (dist(dpos[i]) - pmin) / (pmax - pmin);
                out = static_cast<float>(v);
            }
        }
        sdf[b * HW + i] = out;
    }
}

// ------------------------------------------------------------------ segmentation loss
struct SegSums {            // per sample, per head (0 = main, 1 = aux): fp64 accumulators
    double bce, pg, pp, gg, psdf, wbce, wbd, w;
};

// pass 1: reductions.  wf_amp == 0: plain per-sample terms; wf_amp != 0 additionally the w^F-weighted sums
__global__ void seg_loss_reduce_kernel(const float* __restrict__ pm, const float* __restrict__ pa,
                                       const float* __restrict__ g, const float* __restrict__ sdf, double* __restrict__ sums,
                                       float* __restrict__ wmap_sum, int HW, float wf_amp, float smooth) {
    const int b = blockIdx.y;
    double acc[2][7] = {{0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0}};
    double wsum = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const size_t o = static_cast<size_t>(b) * HW + i;
        const float t = g[o], sd = sdf[o];
        const float p_main_raw = pm[o];
        const float w = wf_amp != 0.f ? expf(wf_amp * fabsf(p_main_raw - t)) : 1.f;     // oriented_weight.py:80-83 (pred detached)
        wsum += w;
        if (wmap_sum) atomicAdd(&wmap_sum[i], w);
#pragma unroll
        for (int hd = 0; hd < 2; ++hd) {
            const float* pp = hd == 0 ? pm : pa;
            if (!pp) continue;
            const float p = fmaxf(pp[o], smooth);                                        // predict.clamp(min=smooth), :51
            const float bce = -(t * logf(p + smooth) + (1.f - t) * logf(1.f - p + smooth)) / 2.f;   // :201, pos_weight [1,1]
            const float bd = p * sd;
            acc[hd][0] += bce;
            acc[hd][1] += static_cast<double>(p) * t;
            acc[hd][2] += static_cast<double>(p) * p;
            acc[hd][3] += static_cast<double>(t) * t;
            acc[hd][4] += bd;
            acc[hd][5] += static_cast<double>(w) * bce;
            acc[hd][6] += static_cast<double>(w) * bd;
        }
    }
#pragma unroll
    for (int hd = 0; hd < 2; ++hd)
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const double v = warp_sum(acc[hd][k]);
            if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(&sums[(b * 2 + hd) * 8 + k], v);
        }
    wsum = warp_sum(wsum);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sums[(b * 2) * 8 + 7], wsum);
}

// pass 2 (out_map = False): per-sample loss [B] and gradients w.r.t. the raw predictions
//   loss_b = main_w * l(main) + aux_w * l(aux),  l = alpha*(bce_mean + dice)/2 + (1-alpha)*mean(p*sdf)
__global__ void seg_loss_finish_kernel(const float* __restrict__ pm, const float* __restrict__ pa, const float* __restrict__ g,
                                       const float* __restrict__ sdf, const double* __restrict__ sums,
                                       float* __restrict__ loss, float* __restrict__ grad_m, float* __restrict__ grad_a,
                                       const float* __restrict__ upstream, int HW, float alpha, float main_w, float aux_w,
                                       float smooth, float dice_smooth) {
    const int b = blockIdx.y;
    double l[2] = {0, 0}, N[2], D[2];
#pragma unroll
    for (int hd = 0; hd < 2; ++hd) {
        const double* s = sums + (b * 2 + hd) * 8;
        N[hd] = 2.0 * s[1] + dice_smooth;
        D[hd] = s[2] + s[3] + dice_smooth;
        const double bce = s[0] / HW, dice = 1.0 - N[hd] / D[hd], bd = s[4] / HW;
        l[hd] = alpha * (bce + dice) / 2.0 + (1.0 - alpha) * bd;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) loss[b] = static_cast<float>(main_w * l[0] + (pa ? aux_w * l[1] : 0.0));
    if (!grad_m && !grad_a) return;
    const float up = upstream ? upstream[b] : 1.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const size_t o = static_cast<size_t>(b) * HW + i;
        const float t = g[o], sd = sdf[o];
#pragma unroll
        for (int hd = 0; hd < 2; ++hd) {
            const float* pp = hd == 0 ? pm : pa;
            float* gp = hd == 0 ? grad_m : grad_a;
            if (!pp || !gp) continue;
            const float raw = pp[o];
            float gr = 0.f;
            if (raw >= smooth) {                                   // clamp(min) passes the gradient where x >= min
                const double p = raw;
                const double dbce = -(t / (p + smooth) - (1.0 - t) / (1.0 - p + smooth)) / 2.0 / HW;
                const double ddice = -(2.0 * t * D[hd] - N[hd] * 2.0 * p) / (D[hd] * D[hd]);
                const double dbd = static_cast<double>(sd) / HW;
                gr = static_cast<float>((hd == 0 ? main_w : aux_w) * up * (alpha * (dbce + ddice) / 2.0 + (1.0 - alpha) * dbd));
            }
            gp[o] = gr;
        }
    }
}

// cross term of the w^F mean: sum_hw W[hw] * sum_j (1/numel - (2 p_j g_j + eps) / Dall), Dall = sum over the whole batch
__global__ void wf_cross_kernel(const float* __restrict__ pm, const float* __restrict__ pa, const float* __restrict__ g,
                                const float* __restrict__ wmap, const double* __restrict__ sums, double* __restrict__ cross,
                                int B, int HW, float smooth, float dice_smooth) {
    double Dall[2] = {dice_smooth, dice_smooth};
    for (int j = 0; j < B; ++j) {
        Dall[0] += sums[(j * 2) * 8 + 2] + sums[(j * 2) * 8 + 3];
        Dall[1] += sums[(j * 2 + 1) * 8 + 2] + sums[(j * 2 + 1) * 8 + 3];
    }
    const double numel = static_cast<double>(B) * HW;
    double acc[2] = {0, 0}, accq[2] = {0, 0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const double w = wmap[i];
#pragma unroll
        for (int hd = 0; hd < 2; ++hd) {
            const float* pp = hd == 0 ? pm : pa;
            if (!pp) continue;
            double dj = 0.0, nj = 0.0;
            for (int j = 0; j < B; ++j) {
                const size_t o = static_cast<size_t>(j) * HW + i;
                const double p = fmaxf(pp[o], smooth);
                dj += 1.0 / numel - (2.0 * p * g[o] + dice_smooth) / Dall[hd];
                nj += 2.0 * p * g[o] + dice_smooth;
            }
            acc[hd] += w * dj;
            accq[hd] += w * nj;                       // Q = sum_hw W[hw] * sum_j (2 p_j g_j + eps): backward of the Dall term
        }
    }
#pragma unroll
    for (int hd = 0; hd < 2; ++hd) {
        const double v = warp_sum(acc[hd]), q = warp_sum(accq[hd]);
        if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(&cross[hd], v);
        if ((threadIdx.x & 31) == 0 && q != 0.0) atomicAdd(&cross[2 + hd], q);
    }
}

// gradient of the w^F-weighted (B,B,H,W) mean w.r.t. the raw predictions (w^F itself is detached, oriented_weight.py:81):
//   d/dp_k[hw] = w_hd * up * [ W_k (alpha/2 dbce + (1-alpha) sdf) / (B HW)
//                              + (alpha/2) / (B^2 HW) * ( -2 g_k SW[hw] / Dall + 2 p_k Q / Dall^2 ) ],  zero where p < smooth
__global__ void wf_grad_kernel(const float* __restrict__ pm, const float* __restrict__ pa, const float* __restrict__ g,
                               const float* __restrict__ sdf, const float* __restrict__ wmap, const double* __restrict__ sums,
                               const double* __restrict__ cross, float* __restrict__ grad_m, float* __restrict__ grad_a,
                               const float* __restrict__ upstream, int B, int HW, float alpha, float main_w, float aux_w,
                               float wf_amp, float smooth, float dice_smooth) {
    const int b = blockIdx.y;
    double Dall[2] = {dice_smooth, dice_smooth};
    for (int j = 0; j < B; ++j) {
        Dall[0] += sums[(j * 2) * 8 + 2] + sums[(j * 2) * 8 + 3];
        Dall[1] += sums[(j * 2 + 1) * 8 + 2] + sums[(j * 2 + 1) * 8 + 3];
    }
    const double up = upstream ? static_cast<double>(*upstream) : 1.0;
    const double c1 = 1.0 / (static_cast<double>(B) * HW), c2 = (alpha / 2.0) / (static_cast<double>(B) * B * HW);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const size_t o = static_cast<size_t>(b) * HW + i;
        const float t = g[o], sd = sdf[o];
        const double W = expf(wf_amp * fabsf(pm[o] - t));
        const double SW = wmap[i];
#pragma unroll
        for (int hd = 0; hd < 2; ++hd) {
            const float* pp = hd == 0 ? pm : pa;
            float* gp = hd == 0 ? grad_m : grad_a;
            if (!pp || !gp) continue;
            const float raw = pp[o];
            float gr = 0.f;
            if (raw >= smooth) {
                const double p = raw;
                const double dbce = -(t / (p + smooth) - (1.0 - t) / (1.0 - p + smooth)) / 2.0;
                const double t1 = c1 * W * (alpha / 2.0 * dbce + (1.0 - alpha) * sd);
                const double t2 = c2 * (-2.0 * t * SW / Dall[hd] + 2.0 * p * cross[2 + hd] / (Dall[hd] * Dall[hd]));
                gr = static_cast<float>((hd == 0 ? main_w : aux_w) * up * (t1 + t2));
            }
            gp[o] = gr;
        }
    }
}
__global__ void wf_final_kernel(const double* __restrict__ sums, const double* __restrict__ cross, double* __restrict__ out, int B,
                                int HW, float alpha, float main_w, float aux_w, int has_aux) {
    double total = 0.0;
    for (int hd = 0; hd < (has_aux ? 2 : 1); ++hd) {
        double wb = 0.0;
        for (int i = 0; i < B; ++i) {
            const double* s = sums + (i * 2 + hd) * 8;
            wb += alpha * s[5] / 2.0 + (1.0 - alpha) * s[6];
        }
        const double term1 = wb / (static_cast<double>(B) * HW);
        const double term2 = (alpha / 2.0) * cross[hd] / (static_cast<double>(B) * B * HW);
        total += (hd == 0 ? main_w : aux_w) * (term1 + term2);
    }
    *out = total;
}

// per-sample mean |a - b| (nn.L1Loss(reduction='none').mean((1,2,3)), sr_loss_functions.py:41-45,53)
__global__ void l1_mean_kernel(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ acc, int n) {
    const int s = blockIdx.y;
    double local = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const size_t o = static_cast<size_t>(s) * n + i;
        local += fabsf(a[o] - b[o]);
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc[s], local);
}
__global__ void mse_mean_kernel(const float* __restrict__ a, const float* __restrict__ b, double* __restrict__ acc, int n) {
    const int s = blockIdx.y;
    double local = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const size_t o = static_cast<size_t>(s) * n + i;
        const float d = a[o] - b[o];
        local += static_cast<double>(d) * d;
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) atomicAdd(&acc[s], local);
}
__global__ void sr_loss_finish_kernel(const double* __restrict__ acc, float* __restrict__ loss, int B, int n_hr, int n_lr, int n_k,
                                      float w_hr, float w_lr, float w_k) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float hr = static_cast<float>(acc[b] / n_hr), lr = static_cast<float>(acc[B + b] / n_lr),
                kk = static_cast<float>(acc[2 * B + b] / n_k);
    loss[b] = w_hr * hr + w_lr * lr + w_k * kk;
}

static inline size_t al(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace csbsr

using namespace csbsr;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" size_t csbsr_sdf_workspace_bytes(int b, int h, int w) {
    const size_t HW = static_cast<size_t>(h) * w;
    return al(b * HW) + al(sizeof(int) * b) + al(sizeof(short) * 2 * b * HW) + al(sizeof(int) * 2 * b * HW) + al(sizeof(int) * 4 * b);
}

extern "C" int csbsr_sdf(const float* mask, float* sdf, int b, int h, int w, void* workspace, size_t workspace_bytes,
                         void* stream_) {
    cudaStream_t stream = STREAM(stream_);
    CSBSR_REQUIRE(mask && sdf && workspace && b > 0 && h > 0 && w > 0 && h <= 32768 && w <= 32768, "sdf: bad arguments");
    CSBSR_REQUIRE(workspace_bytes >= csbsr_sdf_workspace_bytes(b, h, w), "sdf: workspace too small");
    const size_t HW = static_cast<size_t>(h) * w;
    char* base = static_cast<char*>(workspace);
    unsigned char* fg = reinterpret_cast<unsigned char*>(base); base += al(b * HW);
    int* any_fg = reinterpret_cast<int*>(base); base += al(sizeof(int) * b);
    unsigned short* g = reinterpret_cast<unsigned short*>(base); base += al(sizeof(short) * 2 * b * HW);
    unsigned int* d2 = reinterpret_cast<unsigned int*>(base); base += al(sizeof(int) * 2 * b * HW);
    unsigned int* mm = reinterpret_cast<unsigned int*>(base);
    CSBSR_CHECK_CUDA(cudaMemsetAsync(any_fg, 0, sizeof(int) * b, stream));
    // min slots start at 0xFFFFFFFF, max slots at 0: byte patterns differ, so initialise with two strided memsets
    CSBSR_CHECK_CUDA(cudaMemset2DAsync(mm, 8, 0xFF, 4, static_cast<size_t>(2) * b, stream));
    CSBSR_CHECK_CUDA(cudaMemset2DAsync(mm + 1, 8, 0x00, 4, static_cast<size_t>(2) * b, stream));
    const int slices = static_cast<int>((HW + 255) / 256);
    sdf_binarise_kernel<<<dim3(slices < 64 ? slices : 64, b), 256, 0, stream>>>(mask, fg, any_fg, static_cast<int>(HW));
    sdf_column_kernel<<<dim3((w + 127) / 128, 2 * b), 128, 0, stream>>>(fg, g, h, w);
    sdf_row_kernel<<<dim3(slices, 2 * b), 256, 0, stream>>>(g, d2, mm, h, w);
    sdf_finish_kernel<<<dim3(slices, b), 256, 0, stream>>>(fg, d2, mm, any_fg, sdf, h, w);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" size_t csbsr_seg_loss_workspace_bytes(int b) { return al(sizeof(double) * 16 * static_cast<size_t>(b)); }
extern "C" size_t csbsr_seg_loss_wf_workspace_bytes(int b, int hw) {
    return al(sizeof(double) * 16 * static_cast<size_t>(b)) + al(sizeof(double) * 4) + al(sizeof(float) * static_cast<size_t>(hw));
}

extern "C" int csbsr_seg_loss(const float* p_main, const float* p_aux, const float* target, const float* sdf, int b, int hw,
                              float alpha, float main_w, float aux_w, float* loss, float* grad_main, float* grad_aux,
                              const float* upstream, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = STREAM(stream_);
    CSBSR_REQUIRE(p_main && target && sdf && loss && workspace && b > 0 && hw > 0, "seg_loss: bad arguments");
    CSBSR_REQUIRE(workspace_bytes >= csbsr_seg_loss_workspace_bytes(b), "seg_loss: workspace too small");
    CSBSR_REQUIRE(!grad_aux || p_aux, "seg_loss: grad_aux without p_aux");
    double* sums = static_cast<double*>(workspace);
    CSBSR_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 16 * b, stream));
    int slices = (hw + 256 * 8 - 1) / (256 * 8);
    seg_loss_reduce_kernel<<<dim3(slices, b), 256, 0, stream>>>(p_main, p_aux, target, sdf, sums, nullptr, hw, 0.f, 1e-8f);
    seg_loss_finish_kernel<<<dim3(slices, b), 256, 0, stream>>>(p_main, p_aux, target, sdf, sums, loss, grad_main, grad_aux,
                                                                upstream, hw, alpha, main_w, aux_w, 1e-8f, 1e-6f);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// scalar the trainer takes from the (B,B,H,W) w^F loss tensor: mean over (i, j, h, w) of
//   w_i * ( alpha * (bce_i + dice_j) / 2 + (1 - alpha) * bd_i ),  main + aux_w * aux, both weighted by the MAIN prediction's w^F
// (loss_functions.py:292-296,334 ; build_model.py:413-414,433-434 ; trainer.py:407).  `out` is a DEVICE double.
extern "C" int csbsr_seg_loss_wf_mean(const float* p_main, const float* p_aux, const float* target, const float* sdf, int b,
                                      int hw, float alpha, float main_w, float aux_w, float wf_amp, double* out,
                                      void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = STREAM(stream_);
    CSBSR_REQUIRE(p_main && target && sdf && out && workspace && b > 0 && hw > 0, "seg_loss_wf_mean: bad arguments");
    const size_t need = csbsr_seg_loss_wf_workspace_bytes(b, hw);
    CSBSR_REQUIRE(workspace_bytes >= need, "seg_loss_wf_mean: workspace too small (%zu < %zu)", workspace_bytes, need);
    double* sums = static_cast<double*>(workspace);
    double* cross = reinterpret_cast<double*>(static_cast<char*>(workspace) + al(sizeof(double) * 16 * b));
    float* wmap = reinterpret_cast<float*>(static_cast<char*>(workspace) + al(sizeof(double) * 16 * b) + al(sizeof(double) * 4));
    CSBSR_CHECK_CUDA(cudaMemsetAsync(workspace, 0, need, stream));
    int slices = (hw + 256 * 8 - 1) / (256 * 8);
    seg_loss_reduce_kernel<<<dim3(slices, b), 256, 0, stream>>>(p_main, p_aux, target, sdf, sums, wmap, hw, wf_amp, 1e-8f);
    wf_cross_kernel<<<slices, 256, 0, stream>>>(p_main, p_aux, target, wmap, sums, cross, b, hw, 1e-8f, 1e-6f);
    wf_final_kernel<<<1, 1, 0, stream>>>(sums, cross, out, b, hw, alpha, main_w, p_aux ? aux_w : 0.f, p_aux != nullptr);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Backward of csbsr_seg_loss_wf_mean: grad_main / grad_aux [B,HW] = upstream * d mean / d p (upstream: DEVICE float or NULL = 1).
extern "C" int csbsr_seg_loss_wf_grad(const float* p_main, const float* p_aux, const float* target, const float* sdf, int b,
                                      int hw, float alpha, float main_w, float aux_w, float wf_amp, const float* upstream,
                                      float* grad_main, float* grad_aux, void* workspace, size_t workspace_bytes,
                                      void* stream_) {
    cudaStream_t stream = STREAM(stream_);
    CSBSR_REQUIRE(p_main && target && sdf && grad_main && workspace && b > 0 && hw > 0, "seg_loss_wf_grad: bad arguments");
    CSBSR_REQUIRE(!p_aux == !grad_aux, "seg_loss_wf_grad: p_aux and grad_aux go together");
    const size_t need = csbsr_seg_loss_wf_workspace_bytes(b, hw);
    CSBSR_REQUIRE(workspace_bytes >= need, "seg_loss_wf_grad: workspace too small (%zu < %zu)", workspace_bytes, need);
    double* sums = static_cast<double*>(workspace);
    double* cross = reinterpret_cast<double*>(static_cast<char*>(workspace) + al(sizeof(double) * 16 * b));
    float* wmap = reinterpret_cast<float*>(static_cast<char*>(workspace) + al(sizeof(double) * 16 * b) + al(sizeof(double) * 4));
    CSBSR_CHECK_CUDA(cudaMemsetAsync(workspace, 0, need, stream));
    int slices = (hw + 256 * 8 - 1) / (256 * 8);
    seg_loss_reduce_kernel<<<dim3(slices, b), 256, 0, stream>>>(p_main, p_aux, target, sdf, sums, wmap, hw, wf_amp, 1e-8f);
    wf_cross_kernel<<<slices, 256, 0, stream>>>(p_main, p_aux, target, wmap, sums, cross, b, hw, 1e-8f, 1e-6f);
    wf_grad_kernel<<<dim3(slices, b), 256, 0, stream>>>(p_main, p_aux, target, sdf, wmap, sums, cross, grad_main, grad_aux,
                                                       upstream, b, hw, alpha, main_w, aux_w, wf_amp, 1e-8f, 1e-6f);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// backward of the per-sample means: d mean|a - b| / da = sign(a - b) / n (0 where a == b, as torch's abs), d mean((a - b)^2) / da
// = 2 (a - b) / n, each scaled by the sample's upstream gradient
__global__ void l1_mean_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ up,
                                   float* __restrict__ d, int n, float w_over_n) {
    const int s = blockIdx.y;
    const float g = w_over_n * (up ? up[s] : 1.f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const size_t o = static_cast<size_t>(s) * n + i;
        const float df = a[o] - b[o];
        d[o] = df > 0.f ? g : (df < 0.f ? -g : 0.f);
    }
}
__global__ void mse_mean_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ up,
                                    float* __restrict__ d, int n, float w_over_n) {
    const int s = blockIdx.y;
    const float g = 2.f * w_over_n * (up ? up[s] : 1.f);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const size_t o = static_cast<size_t>(s) * n + i;
        d[o] = g * (a[o] - b[o]);
    }
}

extern "C" size_t csbsr_sr_loss_workspace_bytes(int b) { return al(sizeof(double) * 3 * static_cast<size_t>(b)); }

// KBPNLoss.forward given the pseudo-LR image (csbsr_blur_per_sample stride 1 + csbsr_resize_bicubic_aa of sr with the
// normalised predicted kernel): loss[b] = w_hr*mean|sr-hr| + w_lr*mean|pseudo_lr-lr| + w_k*mean((k_pred-k_gt)^2)
extern "C" int csbsr_sr_loss(const float* sr, const float* hr, const float* pseudo_lr, const float* lr, const float* k_pred,
                             const float* k_gt, int b, int n_hr, int n_lr, int n_k, float w_hr, float w_lr, float w_k,
                             float* loss, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = STREAM(stream_);
    CSBSR_REQUIRE(sr && hr && pseudo_lr && lr && k_pred && k_gt && loss && workspace && b > 0, "sr_loss: bad arguments");
    CSBSR_REQUIRE(workspace_bytes >= csbsr_sr_loss_workspace_bytes(b), "sr_loss: workspace too small");
    double* acc = static_cast<double*>(workspace);
    CSBSR_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 3 * b, stream));
    l1_mean_kernel<<<dim3((n_hr + 2047) / 2048, b), 256, 0, stream>>>(sr, hr, acc, n_hr);
    l1_mean_kernel<<<dim3((n_lr + 2047) / 2048, b), 256, 0, stream>>>(pseudo_lr, lr, acc + b, n_lr);
    mse_mean_kernel<<<dim3((n_k + 2047) / 2048, b), 256, 0, stream>>>(k_pred, k_gt, acc + 2 * b, n_k);
    sr_loss_finish_kernel<<<(b + 127) / 128, 128, 0, stream>>>(acc, loss, b, n_hr, n_lr, n_k, w_hr, w_lr, w_k);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// Backward of csbsr_sr_loss w.r.t. sr, pseudo_lr and k_pred: d(sum_b upstream[b] * loss[b]) (upstream NULL = ones); d_k may be NULL.
extern "C" int csbsr_sr_loss_bwd(const float* sr, const float* hr, const float* pseudo_lr, const float* lr, const float* k_pred,
                                 const float* k_gt, const float* upstream, int b, int n_hr, int n_lr, int n_k, float w_hr, float w_lr,
                                 float w_k, float* d_sr, float* d_plr, float* d_k, void* stream_) {
    cudaStream_t stream = STREAM(stream_);
    CSBSR_REQUIRE(sr && hr && pseudo_lr && lr && d_sr && d_plr && b > 0 && n_hr > 0 && n_lr > 0, "sr_loss_bwd: bad arguments");
    CSBSR_REQUIRE(!d_k || (k_pred && k_gt && n_k > 0), "sr_loss_bwd: d_k needs k_pred and k_gt");
    l1_mean_bwd_kernel<<<dim3((n_hr + 2047) / 2048, b), 256, 0, stream>>>(sr, hr, upstream, d_sr, n_hr, w_hr / static_cast<float>(n_hr));
    l1_mean_bwd_kernel<<<dim3((n_lr + 2047) / 2048, b), 256, 0, stream>>>(pseudo_lr, lr, upstream, d_plr, n_lr, w_lr / static_cast<float>(n_lr));
    if (d_k) mse_mean_bwd_kernel<<<dim3((n_k + 2047) / 2048, b), 256, 0, stream>>>(k_pred, k_gt, upstream, d_k, n_k, w_k / static_cast<float>(n_k));
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
