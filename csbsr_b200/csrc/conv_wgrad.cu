// Weight-gradient GEMM of a convolution for sm_100a (training path):
//
//   Wg[m][tap][c] = sum over pixels (n, y, x) of  G[n, y, x, m] * S[n, y*stride + dh[tap], x*stride + dw[tap], c]
//
// For nn.Conv2d: G = dL/dy (NHWC bf16), S = the layer input, m = cout, c = cin.  For nn.ConvTranspose2d(8, 4, 2):
// G = the layer input, S = dL/dy, m = cin, c = cout (the same expression with the roles swapped).  Replaces the cuDNN
// wgrad behind loss.backward() (reference model/engine/trainer.py:57-72 -> autograd of kbpn.py:266-277).
//
// GEMM view: M = 128 channels of G, N = up to 128 channels of S per (tap, channel chunk) "unit", K = pixels.  Both
// operands are pixel-major in memory, so they are fed to tcgen05.mma as MN-major 128B-swizzled tiles: a TMA box of
// 64 pixels x 64 channels is exactly one [K = 64][MN = 64] slab.  Up to four units share one G tile per pipeline stage
// and accumulate side by side in the 512 TMEM columns; the pixel range is split across CTAs and every CTA adds its
// partial sums into the fp32 result with vector reductions (red.global.add.v4.f32).
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/csbsr_b200.h"

namespace csbsr {

static constexpr int kWgK = 64;                      // pixels per pipeline stage
static constexpr int kWgBox = kWgK * 128;            // one TMA box: 64 pixels x 64 bf16 channels = 8 KB
static constexpr int kWgThreads = 256;               // warp0 TMA, warp1 MMA (+TMEM alloc), warps 4-7 epilogue
static constexpr int kWgMaxUnits = 4;                // 4 x 128 fp32 columns = all of TMEM
static constexpr int kWgSmem = 200 * 1024;

struct WgradParams {
    int N, OH, OW, TH, TW, tiles_h, tiles_w, ptiles;
    int stride, ntaps, ncs, cs_pad, units_total, upp, passes_per_m, npass, nsplit, tiles_per_split, nitems;
    int stages, stage_bytes;
    // tap groups: G taps that share dw and whose dh are dh0 + j*rs*stride are served by ONE S box of TH + (G-1)*rs rows per
    // 64-channel chunk (the descriptor start moves by whole swizzle atoms); a unit = (group, 128-channel chunk of S)
    int G, ngroups, rs, b_box_bytes;
    // accumulator columns: a unit owns unit_cols TMEM columns, tap j of the unit starts at j * col_stride.  merge = 1 (single
    // 64-channel chunk of S and G > 1 vertical taps): ONE MMA of N = 64 * G covers all taps of the unit -- the N dimension's
    // second-level stride (LBO) of the MN-major B descriptor is the vertical tap shift inside the shared S box
    int merge, unit_cols, col_stride;
    // ws != nullptr: every (pixel split, output element) is written once into ws[split][tap * cs_pad + c][row] with plain
    // coalesced stores (32 lanes = 32 consecutive rows) and reduced over the splits by wgrad_reduce_kernel -- no atomics, no
    // zero-fill of the result, bit-reproducible.  ws == nullptr: red.global.add into the zeroed wg (round-1 path)
    float* ws;
    int rows_pad, cg;
    float* wg;
    int* err_flag;
    int8_t dh[CSBSR_MAX_TAPS], dw[CSBSR_MAX_TAPS];      // per group: offsets of its first tap
    int8_t tap[CSBSR_MAX_TAPS];                          // [group*G + j] -> tap index in the caller's order
};

// MN-major, 128B-swizzled operand: 8 pixel rows of 128 B form one swizzle atom (SBO = 1024 B between K groups of 8 pixels),
// the next 64 channels live one TMA box further (LBO = 8 KB).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes = kWgBox) {
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmS, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
    uint64_t* full_bar = bars;                 // [stages]
    uint64_t* empty_bar = bars + 8;            // [stages]
    uint64_t* tmem_full = bars + 16;
    uint64_t* tmem_empty = bars + 17;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4);
        fence_barrier_init();
        tma_prefetch_desc(&tmG);
        tma_prefetch_desc(&tmS);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_img = p.tiles_h * p.tiles_w;
    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        uint32_t it = 0;
        // one elected thread runs the whole loop (no per-stage elect / reconvergence in the issue path)
        if (elect_one_sync())
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            const int split = item / p.npass, pass = item - split * p.npass;
            const int mt = pass / p.passes_per_m, pl = pass - mt * p.passes_per_m;
            const int u0 = pl * p.upp, u1 = min(u0 + p.upp, p.units_total);
            const int t0 = split * p.tiles_per_split, t1 = min(t0 + p.tiles_per_split, p.ptiles);
            for (int t = t0; t < t1; ++t, ++it) {
                const int stage = it % p.stages;
                const uint32_t par = (it / p.stages) & 1u;
                mbar_wait(&empty_bar[stage], par ^ 1u, p.err_flag, 21);
                {
                    const int img = t / tiles_per_img, tr = t - img * tiles_per_img;
                    const int oy0 = (tr / p.tiles_w) * p.TH, ox0 = (tr % p.tiles_w) * p.TW;
                    uint32_t bytes = 2 * kWgBox;
                    for (int u = u0; u < u1; ++u) {
                        const int csc = u % p.ncs;
                        bytes += (min(128, p.cs_pad - csc * 128) / 64) * p.b_box_bytes;
                    }
                    mbar_arrive_expect_tx(&full_bar[stage], bytes);
                    uint32_t dst = smem_base + stage * p.stage_bytes;
                    tma_load_4d(dst, &tmG, &full_bar[stage], mt * 128, ox0, oy0, img);
                    tma_load_4d(dst + kWgBox, &tmG, &full_bar[stage], mt * 128 + 64, ox0, oy0, img);
                    dst += 2 * kWgBox;
                    for (int u = u0; u < u1; ++u) {
                        const int grp = u / p.ncs, csc = u - grp * p.ncs;
                        const int nb = min(128, p.cs_pad - csc * 128) / 64;
                        const int sx = ox0 * p.stride + p.dw[grp], sy = oy0 * p.stride + p.dh[grp];
                        for (int h = 0; h < nb; ++h)
                            tma_load_4d(dst + h * p.b_box_bytes, &tmS, &full_bar[stage], csc * 128 + h * 64, sx, sy, img);
                        dst += 2 * p.b_box_bytes;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        uint32_t it = 0, nitem = 0;
        if (elect_one_sync())
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            const int split = item / p.npass, pass = item - split * p.npass;
            const int pl = pass % p.passes_per_m;
            const int u0 = pl * p.upp, u1 = min(u0 + p.upp, p.units_total);
            const int t0 = split * p.tiles_per_split, t1 = min(t0 + p.tiles_per_split, p.ptiles);
            if (t1 <= t0) continue;
            mbar_wait(tmem_empty, (nitem & 1u) ^ 1u, p.err_flag, 22);
            tcgen05_fence_after();
            for (int t = t0; t < t1; ++t, ++it) {
                const int stage = it % p.stages;
                const uint32_t par = (it / p.stages) & 1u;
                mbar_wait(&full_bar[stage], par, p.err_flag, 23);
                tcgen05_fence_after();
                {
                    const uint32_t a_addr = smem_base + stage * p.stage_bytes;
                    for (int u = u0; u < u1; ++u) {
                        const int csc = u % p.ncs;
                        const int nw = min(128, p.cs_pad - csc * 128);
                        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                               (static_cast<uint32_t>(nw >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
                        const uint32_t b_unit = a_addr + 2 * kWgBox + 2 * p.b_box_bytes * (u - u0);
                        const uint32_t d_unit = tmem_base + static_cast<uint32_t>((u - u0) * p.unit_cols);
                        if (p.merge) {
                            const uint32_t idesc_m = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                                                     (static_cast<uint32_t>((64 * p.G) >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
                            const uint32_t tap_shift = static_cast<uint32_t>(p.rs * p.TW * 128);
#pragma unroll
                            for (int k = 0; k < kWgK / 16; ++k)
                                umma_bf16(d_unit, make_desc_mn(a_addr + k * 2048), make_desc_mn(b_unit + k * 2048, tap_shift), idesc_m,
                                          (t > t0 || k > 0) ? 1u : 0u);
                            continue;
                        }
                        for (int j = 0; j < p.G; ++j) {
                            const uint32_t b_addr = b_unit + static_cast<uint32_t>(j * p.rs * p.TW * 128);   // vertical tap shift
                            const uint32_t d_addr = d_unit + static_cast<uint32_t>(j * p.col_stride);
#pragma unroll
                            for (int k = 0; k < kWgK / 16; ++k)
                                umma_bf16(d_addr, make_desc_mn(a_addr + k * 2048),
                                          make_desc_mn(b_addr + k * 2048, static_cast<uint32_t>(p.b_box_bytes)), idesc,
                                          (t > t0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (t == t1 - 1) umma_commit(tmem_full);
                }
            }
            ++nitem;
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue: TMEM -> red.add into the fp32 gradient
        const int q = warp - 4;                           // TMEM lane quarter
        uint32_t nitem = 0;
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            const int split = item / p.npass, pass = item - split * p.npass;
            const int mt = pass / p.passes_per_m, pl = pass - mt * p.passes_per_m;
            const int u0 = pl * p.upp, u1 = min(u0 + p.upp, p.units_total);
            const int t0 = split * p.tiles_per_split, t1 = min(t0 + p.tiles_per_split, p.ptiles);
            if (t1 <= t0) continue;
            mbar_wait(tmem_full, nitem & 1u, p.err_flag, 24);
            tcgen05_fence_after();
            const int row = mt * 128 + q * 32 + lane;
            float* wrow = p.wg + static_cast<size_t>(row) * p.ntaps * p.cs_pad;
            float* wsp = p.ws ? p.ws + static_cast<size_t>(split) * p.ntaps * p.cs_pad * p.rows_pad + row : nullptr;
            const bool live = mt * 128 + q * 32 < p.cg;          // rows >= cg are zero padding of the last m tile: never read
            for (int u = u0; u < u1; ++u) {
                const int grp = u / p.ncs, csc = u - grp * p.ncs;
                const int nw = min(128, p.cs_pad - csc * 128);
                for (int jt = 0; jt < p.G; ++jt) {
                    const int gcol0 = p.tap[grp * p.G + jt] * p.cs_pad + csc * 128;
                    float* dst = wrow + gcol0;
                    const uint32_t col0 = static_cast<uint32_t>((u - u0) * p.unit_cols + jt * p.col_stride);
                    for (int c = 0; c < nw; c += 16) {
                        uint32_t v[16];
                        tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + col0 + static_cast<uint32_t>(c), v);
                        tmem_ld_wait();
                        if (wsp) {
                            if (live) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    wsp[static_cast<size_t>(gcol0 + c + j) * p.rows_pad] = __uint_as_float(v[j]);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; j += 4)
                                red_add_v4(dst + c + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                           __uint_as_float(v[j + 3]));
                        }
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
            ++nitem;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// Sum of the per-split partials P[s][gcol][row] (gcol = tap * cs + c) with a shared-memory transpose so that both the reads
// (along rows) and the writes are coalesced.  Tile columns enumerate j = c * T + tap (the order of a parameter gradient).
//   mode 0: wg[row][tap][c] = sum            (rows >= cg are written as zero)
//   mode 1: grad[row][b0 + c][tap] += sum    (row < A, c < B: the parameter's own [A][Btot][R][S] layout)
__global__ void __launch_bounds__(1024)
wgrad_reduce_kernel(const float* __restrict__ ws, int nsplit, int rows_pad, int cg, int T, int cs, float* __restrict__ out, int mode,
                    int A, int B, int Btot, int b0) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 32: one element per thread, splits unrolled by 8
    const int ncols = T * cs;
    const int j0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const size_t sstride = static_cast<size_t>(ncols) * rows_pad;
    {
        const int j = j0 + ty, r = r0 + tx;
        float v = 0.f;
        if (j < ncols && r < cg) {
            const int c = j / T, t = j - c * T;
            const float* src = ws + static_cast<size_t>(t * cs + c) * rows_pad + r;
            int sidx = 0;
            for (; sidx + 8 <= nsplit; sidx += 8) {
                float a[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) a[q] = src[(sidx + q) * sstride];
#pragma unroll
                for (int q = 0; q < 8; ++q) v += a[q];
            }
            for (; sidx < nsplit; ++sidx) v += src[sidx * sstride];
        }
        tile[ty][tx] = v;
    }
    __syncthreads();
    {
        const int r = r0 + ty, j = j0 + tx;
        if (j >= ncols || r >= rows_pad) return;
        const float v = tile[tx][ty];
        const int c = j / T, t = j - c * T;
        if (mode == 0) {
            out[(static_cast<size_t>(r) * T + t) * cs + c] = v;
        } else if (r < A && c < B) {
            out[(static_cast<size_t>(r) * Btot + b0 + c) * T + t] += v;
        }
    }
}
// tap-expanded accumulator (one "tap", rows = t * cp + m): grad[m][b0 + c][t] += sum_s P[s][c][t * cp + m]
__global__ void wgrad_reduce_tapexp_kernel(const float* __restrict__ ws, int nsplit, int rows_pad, int cs, float* __restrict__ grad,
                                           int A, int B, int Btot, int b0, int cp) {
    const size_t total = static_cast<size_t>(A) * B * 9;
    const size_t sstride = static_cast<size_t>(cs) * rows_pad;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(i % 9);
        const int c = static_cast<int>((i / 9) % B);
        const int m = static_cast<int>(i / (static_cast<size_t>(9) * B));
        const float* src = ws + static_cast<size_t>(c) * rows_pad + t * cp + m;
        float v = 0.f;
        for (int sidx = 0; sidx < nsplit; ++sidx) v += src[sidx * sstride];
        grad[(static_cast<size_t>(m) * Btot + b0 + c) * 9 + t] += v;
    }
}

typedef CUresult (*PFN_encodeTiledW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiledW wg_encode_fn() {
    static PFN_encodeTiledW fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiledW>(ptr);
    }
    return fn;
}
static int* g_wg_err = nullptr;

}  // namespace csbsr

using namespace csbsr;

static int wg_check(const csbsr_wgrad_desc* d) {
    CSBSR_REQUIRE(d && d->g && d->s, "conv_wgrad: null pointer");
    CSBSR_REQUIRE(d->cg > 0 && d->cg % 64 == 0 && d->cs > 0 && d->cs % 64 == 0,
                  "conv_wgrad: channel counts (%d, %d) must be positive multiples of 64", d->cg, d->cs);
    CSBSR_REQUIRE(d->g_pitch % 8 == 0 && d->g_coff % 8 == 0 && d->s_pitch % 8 == 0 && d->s_coff % 8 == 0,
                  "conv_wgrad: pitches / offsets must be multiples of 8");
    CSBSR_REQUIRE(d->ntaps >= 1 && d->ntaps <= CSBSR_MAX_TAPS, "conv_wgrad: bad tap count %d", d->ntaps);
    CSBSR_REQUIRE(d->stride >= 1 && d->stride <= 8, "conv_wgrad: bad stride %d", d->stride);
    CSBSR_REQUIRE(d->n >= 1 && d->gh >= 1 && d->gw >= 1 && d->sh >= 1 && d->sw >= 1, "conv_wgrad: empty tensor");
    return 0;
}

// work decomposition of one wgrad launch (tap groups, passes, pixel splits, pipeline stages)
static int wg_plan(const csbsr_wgrad_desc* d, WgradParams& p) {
    int G = 1;
    memset(&p, 0, sizeof(p));
    p.N = d->n; p.OH = d->gh; p.OW = d->gw;
    p.TW = d->gw > 8 ? 16 : 8;
    p.TH = kWgK / p.TW;
    p.tiles_h = (p.OH + p.TH - 1) / p.TH;
    p.tiles_w = (p.OW + p.TW - 1) / p.TW;
    p.ptiles = p.N * p.tiles_h * p.tiles_w;
    p.stride = d->stride; p.ntaps = d->ntaps;
    p.cs_pad = d->cs;
    p.ncs = (d->cs + 127) / 128;
    // ---- tap groups (same rule as the forward kernel): taps with equal dw and equal dh modulo the stride, dh in uniform steps
    int ngroups = d->ntaps, rs = 0;
    int8_t g_dh[CSBSR_MAX_TAPS], g_dw[CSBSR_MAX_TAPS], g_tap[CSBSR_MAX_TAPS];
    for (int t = 0; t < d->ntaps; ++t) { g_dh[t] = d->dh[t]; g_dw[t] = d->dw[t]; g_tap[t] = static_cast<int8_t>(t); }
    if (d->ntaps >= 2 && !getenv("CSBSR_NO_GROUPING")) {
        const int st = d->stride;
        auto key_of = [&](int t) { return d->dw[t] * 64 + (((d->dh[t] % st) + st) % st); };
        int keys[CSBSR_MAX_TAPS], nk = 0;
        for (int t = 0; t < d->ntaps; ++t) {
            bool seen = false;
            for (int i = 0; i < nk; ++i) seen |= (keys[i] == key_of(t));
            if (!seen) keys[nk++] = key_of(t);
        }
        bool ok = d->ntaps % nk == 0 && d->ntaps / nk >= 2 && d->ntaps / nk <= 4;
        const int gsz = ok ? d->ntaps / nk : 1;
        int step = 0;
        int8_t n_dh[CSBSR_MAX_TAPS], n_dw[CSBSR_MAX_TAPS], n_tap[CSBSR_MAX_TAPS];
        for (int gi = 0; gi < nk && ok; ++gi) {
            int idx[CSBSR_MAX_TAPS], cnt = 0;
            for (int t = 0; t < d->ntaps; ++t)
                if (key_of(t) == keys[gi]) idx[cnt++] = t;
            if (cnt != gsz) { ok = false; break; }
            for (int a = 0; a < cnt; ++a)
                for (int b2 = a + 1; b2 < cnt; ++b2)
                    if (d->dh[idx[b2]] < d->dh[idx[a]]) { const int tmp = idx[a]; idx[a] = idx[b2]; idx[b2] = tmp; }
            for (int j = 0; j < cnt; ++j) {
                if (j > 0) {
                    const int sd = d->dh[idx[j]] - d->dh[idx[j - 1]];
                    if (sd <= 0 || sd % st != 0 || (step != 0 && sd != step)) ok = false;
                    step = sd;
                }
                n_tap[gi * gsz + j] = static_cast<int8_t>(idx[j]);
            }
            n_dh[gi] = d->dh[idx[0]];
            n_dw[gi] = d->dw[idx[0]];
        }
        const int rows = p.TH + (gsz - 1) * (step / (st > 0 ? st : 1));
        if (ok && rows * st <= 256 && 2 * (2 * kWgBox + 2 * rows * p.TW * 128) <= kWgSmem - 2048) {
            G = gsz; ngroups = nk; rs = step / st;
            memcpy(g_dh, n_dh, sizeof(g_dh)); memcpy(g_dw, n_dw, sizeof(g_dw)); memcpy(g_tap, n_tap, sizeof(g_tap));
        }
    }
    p.G = G; p.ngroups = ngroups; p.rs = rs;
    p.b_box_bytes = (p.TH + (G - 1) * rs) * p.TW * 128;
    p.units_total = ngroups * p.ncs;
    const int m_tiles = (d->cg + 127) / 128;
    // units per pass: every tap of a unit owns 128 TMEM columns; spread the units evenly over the passes
    p.merge = (G > 1 && d->cs == 64 && !getenv("CSBSR_WGRAD_NO_MERGE")) ? 1 : 0;
    p.col_stride = p.merge ? 64 : 128;
    p.unit_cols = G * p.col_stride;
    const int max_units = (kWgMaxUnits * 128) / p.unit_cols > 0 ? (kWgMaxUnits * 128) / p.unit_cols : 1;
    const int min_passes = (p.units_total + max_units - 1) / max_units;
    p.upp = (p.units_total + min_passes - 1) / min_passes;
    p.passes_per_m = (p.units_total + p.upp - 1) / p.upp;
    p.npass = m_tiles * p.passes_per_m;
    const int sms = num_sms();
    // Pixel splits: the work items (pass, split) are dealt round-robin to one CTA per SM and all cost the same, so the launch
    // takes rounds = ceil(items / SMs) items of tiles_per_split pipeline stages plus one accumulator drain each (the
    // accumulators are single-buffered).  Choose the split count that minimises rounds * (stages + drain): rounding the
    // items per SM UP (round 1: 160 items on 148 SMs for the 8x8/s4 layers) made 12 SMs work twice while the rest idled --
    // 54 % efficiency on those layers.  Every split also costs one more pass over the output in the reduction.
    // CSBSR_WGRAD_ITEMS_PER_SM forces the old rule (items = that many per SM, rounded up).
    static const int items_per_sm = getenv("CSBSR_WGRAD_ITEMS_PER_SM") ? atoi(getenv("CSBSR_WGRAD_ITEMS_PER_SM")) : 0;
    int nsplit = 1;
    if (items_per_sm > 0) {
        nsplit = (items_per_sm * sms + p.npass - 1) / p.npass;
    } else {
        const int drain = 2 + (p.upp * p.unit_cols) / 64;          // accumulator drain of one item, in pipeline stages (measured order)
        long long best = -1;
        const int ns_max = 4 * sms / p.npass + 2;
        for (int ns = 1; ns <= ns_max && ns <= p.ptiles; ++ns) {
            const int tps = (p.ptiles + ns - 1) / ns;
            const int ns_eff = (p.ptiles + tps - 1) / tps;
            const int rounds = (p.npass * ns_eff + sms - 1) / sms;
            const long long cost = static_cast<long long>(rounds) * (tps + drain) * 16 + ns_eff * 4;   // + reduction passes
            if (best < 0 || cost < best) { best = cost; nsplit = ns_eff; }
        }
    }
    if (nsplit > p.ptiles) nsplit = p.ptiles;
    if (nsplit < 1) nsplit = 1;
    p.tiles_per_split = (p.ptiles + nsplit - 1) / nsplit;
    p.nsplit = (p.ptiles + p.tiles_per_split - 1) / p.tiles_per_split;
    p.nitems = p.npass * p.nsplit;
    p.stage_bytes = 2 * kWgBox + 2 * p.b_box_bytes * p.upp;
    p.stages = (kWgSmem - 2048) / p.stage_bytes;
    if (p.stages > 8) p.stages = 8;
    CSBSR_REQUIRE(p.stages >= 2, "conv_wgrad: not enough shared memory for 2 stages");
    p.rows_pad = m_tiles * 128;
    p.cg = d->cg;
    memcpy(p.dh, g_dh, sizeof(p.dh));
    memcpy(p.dw, g_dw, sizeof(p.dw));
    memcpy(p.tap, g_tap, sizeof(p.tap));
    return 0;
}

extern "C" size_t csbsr_conv_wgrad_workspace_bytes(const csbsr_wgrad_desc* d) {
    WgradParams p;
    if (wg_check(d) || wg_plan(d, p)) return 0;
    return sizeof(float) * static_cast<size_t>(p.nsplit) * p.ntaps * p.cs_pad * p.rows_pad;
}

extern "C" int csbsr_conv_wgrad(const csbsr_wgrad_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (int rc = wg_check(d)) return rc;
    CSBSR_REQUIRE(d->wg || (d->ws && d->grad), "conv_wgrad: no output (wg, or ws + grad)");
    PFN_encodeTiledW encode = wg_encode_fn();
    CSBSR_REQUIRE(encode, "conv_wgrad: cuTensorMapEncodeTiled entry point unavailable");
    WgradParams p;
    if (int rc = wg_plan(d, p)) return rc;
    const int G = p.G, rs = p.rs, m_tiles = p.rows_pad / 128;
    const int sms = num_sms();
    const size_t ws_need = sizeof(float) * static_cast<size_t>(p.nsplit) * p.ntaps * p.cs_pad * p.rows_pad;
    if (d->ws) {
        CSBSR_REQUIRE(d->ws_bytes >= ws_need, "conv_wgrad: workspace too small (%zu < %zu)", (size_t)d->ws_bytes, ws_need);
        CSBSR_REQUIRE((reinterpret_cast<uintptr_t>(d->ws) & 15) == 0, "conv_wgrad: workspace must be 16-byte aligned");
        p.ws = static_cast<float*>(d->ws);
    }
    if (d->grad) {
        CSBSR_REQUIRE(d->ws, "conv_wgrad: accumulating into a parameter gradient needs the workspace path");
        if (d->grad_cp > 0)
            CSBSR_REQUIRE(d->ntaps == 1 && d->grad_a <= d->grad_cp && 9 * d->grad_cp <= d->cg && d->grad_b <= d->cs,
                          "conv_wgrad: bad tap-expanded gradient target");
        else
            CSBSR_REQUIRE(d->grad_a <= d->cg && d->grad_b <= d->cs, "conv_wgrad: gradient target larger than the operands");
        CSBSR_REQUIRE(d->grad_a > 0 && d->grad_b > 0 && d->grad_b0 >= 0 && d->grad_b0 + d->grad_b <= d->grad_btot,
                      "conv_wgrad: bad gradient window");
    }
    p.wg = d->wg;
    if (!g_wg_err) {
        CSBSR_CHECK_CUDA(cudaMalloc(&g_wg_err, sizeof(int)));
        CSBSR_CHECK_CUDA(cudaMemset(g_wg_err, 0, sizeof(int)));
    }
    p.err_flag = g_wg_err;
    if (!p.ws)
        CSBSR_CHECK_CUDA(cudaMemsetAsync(d->wg, 0, sizeof(float) * static_cast<size_t>(m_tiles) * 128 * p.ntaps * p.cs_pad, stream));

    CUtensorMap tmG, tmS;
    {
        const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(d->g) + d->g_coff;
        cuuint64_t dims[4] = {(cuuint64_t)d->cg, (cuuint64_t)d->gw, (cuuint64_t)d->gh, (cuuint64_t)d->n};
        const cuuint64_t pb = (cuuint64_t)d->g_pitch * 2;
        cuuint64_t strides[3] = {pb, pb * d->gw, pb * d->gw * d->gh};
        cuuint32_t box[4] = {64, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmG, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CSBSR_REQUIRE(r == CUDA_SUCCESS, "conv_wgrad: cuTensorMapEncodeTiled(G) failed with %d", (int)r);
    }
    {
        const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(d->s) + d->s_coff;
        cuuint64_t dims[4] = {(cuuint64_t)d->cs, (cuuint64_t)d->sw, (cuuint64_t)d->sh, (cuuint64_t)d->n};
        const cuuint64_t pb = (cuuint64_t)d->s_pitch * 2;
        cuuint64_t strides[3] = {pb, pb * d->sw, pb * d->sw * d->sh};
        cuuint32_t box[4] = {64, (cuuint32_t)(p.TW * d->stride), (cuuint32_t)((p.TH + (G - 1) * rs) * d->stride), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
        CUresult r = encode(&tmS, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CSBSR_REQUIRE(r == CUDA_SUCCESS, "conv_wgrad: cuTensorMapEncodeTiled(S) failed with %d", (int)r);
    }
    const int smem_bytes = p.stages * p.stage_bytes + 1024 + 512;
    static bool attr_set = false;
    if (!attr_set) {
        CSBSR_CHECK_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmem));
        attr_set = true;
    }
    const int grid = p.nitems < sms ? p.nitems : sms;
    conv_wgrad_kernel<<<grid, kWgThreads, smem_bytes, stream>>>(tmG, tmS, p);
    if (p.ws) {
        const int ncols = p.ntaps * p.cs_pad;
        if (d->wg)
            wgrad_reduce_kernel<<<dim3((ncols + 31) / 32, p.rows_pad / 32), 1024, 0, stream>>>(p.ws, p.nsplit, p.rows_pad, p.cg, p.ntaps,
                                                                                           p.cs_pad, d->wg, 0, 0, 0, 0, 0);
        if (d->grad && d->grad_cp > 0) {
            const size_t total = static_cast<size_t>(d->grad_a) * d->grad_b * 9;
            int blocks = static_cast<int>((total + 255) / 256);
            if (blocks > sms * 8) blocks = sms * 8;
            wgrad_reduce_tapexp_kernel<<<blocks, 256, 0, stream>>>(p.ws, p.nsplit, p.rows_pad, p.cs_pad, d->grad, d->grad_a, d->grad_b,
                                                                  d->grad_btot, d->grad_b0, d->grad_cp);
        } else if (d->grad) {
            wgrad_reduce_kernel<<<dim3((ncols + 31) / 32, (d->grad_a + 31) / 32), 1024, 0, stream>>>(
                p.ws, p.nsplit, p.rows_pad, p.cg, p.ntaps, p.cs_pad, d->grad, 1, d->grad_a, d->grad_b, d->grad_btot, d->grad_b0);
        }
    }
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
