// Implicit-GEMM convolution for sm_100a: TMA-fed, tcgen05.mma with fp32 accumulators in TMEM,
// persistent warp-specialised CTAs (1 per SM), double-buffered accumulators, fused epilogue.
//
// GEMM view: M = 128 output pixels (a TH x TW patch of one image), N = block_n output channels,
// K = taps x cin.  For each (tap, 64-channel chunk) the A operand is ONE 4-D TMA box of the NHWC
// activation tensor shifted by the tap offset (out-of-image elements are zero-filled by TMA, which
// is the conv zero padding); strided convs use the tensor map's element strides.  The B operand is a
// [block_n x 64] K-major slab of the packed weights.  Both land in 128B-swizzled shared memory and
// are consumed by tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32).
//
// Replaces the cuDNN convs behind nn.Conv2d / nn.ConvTranspose2d of the reference
// (model/modeling/kbpn.py:266-277, 450-518; pspnet_pytorch/extractors.py:37-70; pspnet.py:23-57).
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/csbsr_b200.h"

namespace csbsr {

// CTA-0 timeline instrumentation (scripts/trace_conv.py) is compiled in only with -DCSBSR_CONV_TRACE_BUILD: the runtime checks
// alone cost ~20 instructions per k-block in the single-thread producer / MMA loops, which bound the small layers
#ifdef CSBSR_CONV_TRACE_BUILD
#define CSBSR_TRACE(...) __VA_ARGS__
#else
#define CSBSR_TRACE(...)
#endif

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;                  // bf16 elements = 128 bytes = one swizzle row
static constexpr int kATileBytes = kBlockM * kBlockK * 2;   // 16 KB
static constexpr int kMaxStages = 8;
static constexpr int kEpiWarps = 16;                // 4 warps per TMEM lane quarter, each a share of the 16-column units
static constexpr int kThreads = 128 + 32 * kEpiWarps;   // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps4.. epilogue
static constexpr int kTmemCols = 512;
static constexpr int kSmemBudget = 227 * 1024;

// division by a runtime constant without the ~25-instruction integer divide: q = umulhi(a, mul) >> shr (exact for a < 2^31)
struct FastDiv {
    uint32_t mul, shr, div;
};
static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.div = d;
    if (d == 1) { f.mul = 0; f.shr = 0; return f; }
    uint32_t l = 0;
    while ((1u << l) < d) ++l;                      // ceil(log2 d)
    const uint64_t m = ((static_cast<uint64_t>(1) << (32 + l - 1)) + d - 1) / d;   // ceil(2^(31+l) / d) < 2^32 for a < 2^31
    f.mul = static_cast<uint32_t>(m);
    f.shr = l - 1;
    return f;
}
__device__ __forceinline__ void fast_divmod(uint32_t a, const FastDiv& f, uint32_t& q, uint32_t& r) {
    q = f.div == 1 ? a : (__umulhi(a, f.mul) >> f.shr);
    r = a - q * f.div;
}

struct ConvKParams {
    FastDiv fd_ntiles, fd_nphases, fd_tiles_per_img, fd_tiles_w;
    // tiles
    int tiles_h, tiles_w, m_tiles, n_tiles, nphases, total_tiles;
    int TH, TW, block_n, kchunks, ntaps, stride, stages;
    int kb;              // K elements per pipeline chunk: 64 (128B swizzle rows), 32 (64B) or 16 (32B)
    int cluster;         // 1, or 2: CTA pairs work on neighbouring pixel tiles of the same (n tile, phase) and each loads
                         // half of every weight tile, multicast to both (halves the L2 -> SM weight traffic)
    int OH, OW, os, YH, YW, N;
    int out_mode, y_pitch, y_coff, cout_store;
    int bias_sn, bias_sc, cls_bw, act;
    float slope, r1_sign;
    int r0_pitch, r0_coff, rm_pitch, rm_coff, r1_pitch, r1_coff, r32_pitch, r32_coff;
    void* y;
    const float* bias;
    const __nv_bfloat16* r0;
    const __nv_bfloat16* rm;
    const __nv_bfloat16* r1;
    const float* r32;
    int* err_flag;
    long long* trace;
    int G, ngroups, dstep, a_stage_bytes;   // tap groups: G taps sharing dw, dh = dh0 + j*dstep, one A box per group
    int cg2;             // cluster == 2 only: tcgen05.mma.cta_group::2 -- the pair's leader issues M = 256 MMAs over both CTAs'
                         // A tiles, every CTA holds half of the weight rows (no multicast), TMA bytes land on the leader's barriers
    int nsub, spt, sub_n;   // merged sub-phases (csbsr_conv_desc.nsub): nsub per tap class, spt = block_n / sub_n of them per tile,
                            // sub_n = cout_pad columns each; the two epilogue teams then split every tile by sub-phase
    int res_sets;        // staged epilogue with a residual: staging sets per team (2 = the residual tile of the team's next tile is
                         // fetched while the current one is processed; 1 when shared memory does not allow it)
    int staged;          // 1: epilogue through swizzled smem panels, residual via TMA load, output via TMA store
    int res_mode;        // staged only: 0 none, 1 pre-activation add (r0), 2 post-activation add/sub (r1)
    int8_t dh[CSBSR_MAX_TAPS], dw[CSBSR_MAX_TAPS];
    int16_t widx[CSBSR_MAX_TAPS];
    int8_t ooh[CSBSR_MAX_PHASES], oow[CSBSR_MAX_PHASES];
};

// tile index -> (n tile, phase, image, first output row / column); tiles are ordered n fastest, then phase, then m
// In cluster mode `tile` counts PAIRS of pixel tiles and the CTA of rank `crank` takes pixel tile 2*pair + crank; a pair's
// second tile may not exist: it is mapped to image N (all its loads and TMA stores fall outside the tensors).
__device__ __forceinline__ void decode_tile(const ConvKParams& p, int tile, uint32_t crank, int& nt, int& ph, int& img, int& oh0,
                                            int& ow0) {
    uint32_t rest, mt, tr, a, b, c, d2;
    fast_divmod(static_cast<uint32_t>(tile), p.fd_ntiles, rest, a);
    fast_divmod(rest, p.fd_nphases, mt, b);
    if (p.cluster == 2) mt = 2 * mt + crank;
    if (mt >= static_cast<uint32_t>(p.m_tiles)) {
        nt = static_cast<int>(a); ph = static_cast<int>(b); img = p.N; oh0 = 0; ow0 = 0;
        return;
    }
    fast_divmod(mt, p.fd_tiles_per_img, c, tr);
    uint32_t trh;
    fast_divmod(tr, p.fd_tiles_w, trh, d2);
    nt = static_cast<int>(a); ph = static_cast<int>(b); img = static_cast<int>(c);
    oh0 = static_cast<int>(trh) * p.TH;
    ow0 = static_cast<int>(d2) * p.TW;
}

__device__ __forceinline__ void team_bar_sync(int team) {          // named barriers 1 / 2: one per epilogue team
    asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(32 * kEpiWarps / 2) : "memory");
}

__device__ __forceinline__ int border_class(int i, int n, int bw) {
    if (i < bw) return i;
    int fromend = n - 1 - i;
    if (fromend < bw) return 2 * bw - fromend;
    return bw;
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == CSBSR_ACT_RELU) return fmaxf(v, 0.f);
    if (act == CSBSR_ACT_LEAKY) return v > 0.f ? v : v * slope;
    if (act == CSBSR_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    return v;
}

__device__ __forceinline__ void unpack8_bf16(const uint4& raw, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

template <int ACT>
__device__ __forceinline__ float act_fn(float v, float slope) {
    if (ACT == CSBSR_ACT_RELU) return fmaxf(v, 0.f);
    if (ACT == CSBSR_ACT_LEAKY) return v > 0.f ? v : v * slope;
    if (ACT == CSBSR_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    return v;
}

struct EpiRow {
    bool valid;
    size_t pix;
    int img, oy, ox;
    const float* bias_row;
};

__device__ __forceinline__ void bf16x8_to_f32(const uint4& raw, float (&f)[8]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}

// bf16 NHWC epilogue: one accumulator row (= one output pixel) per thread; this warp handles the 16-column units
// u0, u0+2, ...; the residual operands of a unit are requested before its TMEM load is awaited.
template <int ACT>
__device__ __forceinline__ void epilogue_bf16(const ConvKParams& p, const EpiRow& er, uint32_t taddr0, int c_base,
                                              int u0, int units) {
    const float slope = p.slope;
    const __nv_bfloat16* r0p = p.r0 ? p.r0 + er.pix * p.r0_pitch + p.r0_coff : nullptr;
    const __nv_bfloat16* rmp = p.rm ? p.rm + er.pix * p.rm_pitch + p.rm_coff : nullptr;
    const __nv_bfloat16* r1p = p.r1 ? p.r1 + er.pix * p.r1_pitch + p.r1_coff : nullptr;
    __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + er.pix * p.y_pitch + p.y_coff;
    const float r1s = p.r1_sign;
    for (int u = u0; u < units; u += kEpiWarps / 4) {
        const int c0 = c_base + u * 16;
        const bool on = er.valid && c0 < p.cout_store;
        uint4 q0[2], qm[2], q1[2];                          // always initialised: conditionally written arrays end up in local memory
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int c = c0 + g * 8;
            const bool okc = on && c < p.cout_store;
            q0[g] = (r0p && okc) ? *reinterpret_cast<const uint4*>(r0p + c) : make_uint4(0, 0, 0, 0);
            qm[g] = (rmp && okc) ? *reinterpret_cast<const uint4*>(rmp + c) : make_uint4(0, 0, 0, 0);
            q1[g] = (r1p && okc) ? *reinterpret_cast<const uint4*>(r1p + c) : make_uint4(0, 0, 0, 0);
        }
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(taddr0 + static_cast<uint32_t>(u * 16), v);
        tmem_ld_wait();
        if (on) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const int c = c0 + g * 8;
                if (c >= p.cout_store) break;
                float fg[8], r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) fg[i] = __uint_as_float(v[g * 8 + i]);
                if (er.bias_row) {
                    const float4 b0 = *reinterpret_cast<const float4*>(er.bias_row + c);
                    const float4 b1 = *reinterpret_cast<const float4*>(er.bias_row + c + 4);
                    fg[0] += b0.x; fg[1] += b0.y; fg[2] += b0.z; fg[3] += b0.w;
                    fg[4] += b1.x; fg[5] += b1.y; fg[6] += b1.z; fg[7] += b1.w;
                }
                if (r0p) {
                    bf16x8_to_f32(q0[g], r);
#pragma unroll
                    for (int i = 0; i < 8; ++i) fg[i] += r[i];
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) fg[i] = act_fn<ACT>(fg[i], slope);
                if (rmp) {
                    bf16x8_to_f32(qm[g], r);
#pragma unroll
                    for (int i = 0; i < 8; ++i) fg[i] *= r[i];
                }
                if (r1p) {
                    bf16x8_to_f32(q1[g], r);
#pragma unroll
                    for (int i = 0; i < 8; ++i) fg[i] = fmaf(r1s, r[i], fg[i]);
                }
                uint4 o;
                __nv_bfloat162* oh2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int i = 0; i < 4; ++i) oh2[i] = __floats2bfloat162_rn(fg[2 * i], fg[2 * i + 1]);
                *reinterpret_cast<uint4*>(yp + c) = o;
            }
        }
    }
}

// 16-byte shared-memory accesses by 32-bit shared-window address (generic pointers cost 64-bit address arithmetic per access)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// Staged bf16 epilogue of ONE tile: the (optional) residual tile was TMA-loaded into the swizzled staging panels
// (64 channels x 128 pixels x bf16 = 16 KB each, 128B swizzle like the A operand); every thread owns one pixel row,
// combines accumulator + bias + residual + activation in place and the panels leave through one TMA store each.
// `srow` = shared address of this thread's 128-byte row in panel 0, `sw` = (row & 7) << 4 (the swizzle XOR of the row),
// `bias_u` = this pixel's bias row at the tile's first channel (or null).  RES: 0 none, 1 add before the activation,
// 2 fused multiply-add (r1_sign) after it -- compile-time so that the residual registers never live in local memory.
template <int ACT, int RES>
__device__ __forceinline__ void epilogue_staged(uint32_t srow, uint32_t sw, uint32_t taddr0, int u_begin, int u_end,
                                                const float* bias_u, float slope, float r1s) {
    for (int u = u_begin; u < u_end; ++u) {
        const uint32_t a0 = srow + static_cast<uint32_t>(u >> 2) * kATileBytes + ((static_cast<uint32_t>(u & 3) << 5) ^ sw);
        const uint32_t a1 = a0 ^ 16u;                      // the unit's second 16-byte chunk
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(taddr0 + static_cast<uint32_t>(u * 16), v);
        // residual and bias operands are requested while the TMEM load is in flight
        uint4 q[2];
        if (RES) {
            q[0] = lds128(a0);
            q[1] = lds128(a1);
        }
        float4 b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias_u) {
            const float4* bp = reinterpret_cast<const float4*>(bias_u + u * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) b[i] = __ldg(bp + i);
        }
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            float fg[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) fg[i] = __uint_as_float(v[g * 8 + i]);
            fg[0] += b[2 * g].x; fg[1] += b[2 * g].y; fg[2] += b[2 * g].z; fg[3] += b[2 * g].w;
            fg[4] += b[2 * g + 1].x; fg[5] += b[2 * g + 1].y; fg[6] += b[2 * g + 1].z; fg[7] += b[2 * g + 1].w;
            if (RES == 1) {
                float r[8];
                bf16x8_to_f32(q[g], r);
#pragma unroll
                for (int i = 0; i < 8; ++i) fg[i] += r[i];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) fg[i] = act_fn<ACT>(fg[i], slope);
            if (RES == 2) {
                float r[8];
                bf16x8_to_f32(q[g], r);
#pragma unroll
                for (int i = 0; i < 8; ++i) fg[i] = fmaf(r1s, r[i], fg[i]);
            }
            uint4 o;
            __nv_bfloat162* oh2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int i = 0; i < 4; ++i) oh2[i] = __floats2bfloat162_rn(fg[2 * i], fg[2 * i + 1]);
            sts128(g ? a1 : a0, o);
        }
    }
}

template <int ACT>
__device__ __forceinline__ void epilogue_staged_res(int res_mode, uint32_t srow, uint32_t sw, uint32_t taddr0, int u_begin,
                                                    int u_end, const float* bias_u, float slope, float r1s) {
    if (res_mode == 0) epilogue_staged<ACT, 0>(srow, sw, taddr0, u_begin, u_end, bias_u, slope, r1s);
    else if (res_mode == 1) epilogue_staged<ACT, 1>(srow, sw, taddr0, u_begin, u_end, bias_u, slope, r1s);
    else epilogue_staged<ACT, 2>(srow, sw, taddr0, u_begin, u_end, bias_u, slope, r1s);
}

// fp32 outputs (NHWC for the class biases, planar for images / probabilities): few columns, simple loop
template <int ACT>
__device__ __forceinline__ void epilogue_units(const ConvKParams& p, const EpiRow& er, uint32_t taddr0, int c_base,
                                               int u0, int units) {
    const float slope = p.slope;
    for (int u = u0; u < units; u += kEpiWarps / 4) {
        const int c0 = c_base + u * 16;
        uint32_t v[16];
        __syncwarp();
        tmem_ld16(taddr0 + static_cast<uint32_t>(u * 16), v);
        tmem_ld_wait();
        if (!(er.valid && c0 < p.cout_store)) continue;
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
        if (er.bias_row) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 b = *reinterpret_cast<const float4*>(er.bias_row + c0 + i);
                f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
            }
        }
        if (p.out_mode == CSBSR_OUT_F32_NHWC) {
            float* y32 = reinterpret_cast<float*>(p.y) + er.pix * p.y_pitch + p.y_coff + c0;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                if (c0 + i < p.cout_store) {
                    float4 o;
                    o.x = act_fn<ACT>(f[i], slope);
                    o.y = act_fn<ACT>(f[i + 1], slope);
                    o.z = act_fn<ACT>(f[i + 2], slope);
                    o.w = act_fn<ACT>(f[i + 3], slope);
                    *reinterpret_cast<float4*>(y32 + i) = o;
                }
            }
        } else {                                        // fp32 planar [n][cout_store][YH][YW]
            float* y32 = reinterpret_cast<float*>(p.y);
            const size_t plane = static_cast<size_t>(p.YH) * p.YW;
            // planar tensors may be channel windows of wider buffers: [n][pitch][YH][YW], window start coff
            const size_t pixo = static_cast<size_t>(er.oy) * p.YW + er.ox;
            const size_t base = (static_cast<size_t>(er.img) * p.y_pitch + p.y_coff) * plane + pixo;
            const size_t rbase = (static_cast<size_t>(er.img) * p.r32_pitch + p.r32_coff) * plane + pixo;
            // y and r32 may be the same buffer (in-place accumulation), so the compiler cannot hoist the residual loads
            // above the stores: read all of them first (16 loads in flight instead of one load-store chain per channel)
            float rv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) rv[i] = (p.r32 && c0 + i < p.cout_store) ? __ldg(p.r32 + rbase + (c0 + i) * plane) : 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int c = c0 + i;
                if (c < p.cout_store) y32[base + c * plane] = act_fn<ACT>(f[i], slope) + rv[i];
            }
        }
    }
}

// ------------------------------------------------------------------ kernel
// kCg2: the cta_group::2 variant is a separate instantiation -- a kernel containing cta_group::2 instructions can only be
// launched as whole CTA pairs (a plain launch fails with "cluster misconfiguration"), so the default path must not contain them
template <bool kCg2>
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                  const __grid_constant__ ConvKParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [stages x A tile][stages x B tile][barriers]
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // cta_group::2: every CTA of the pair holds only ITS half of the weight rows (the MMA reads both halves), so a pipeline
    // stage is smaller and more stages fit
    const int b_tile_bytes = (kCg2 ? p.block_n / 2 : p.block_n) * p.kb * 2;
    const int b_stage_bytes = p.G * b_tile_bytes;
    const int b_total_bytes = (p.stages * b_stage_bytes + 1023) & ~1023;      // keeps the staging panels 1024-B aligned
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + p.stages * p.a_stage_bytes;
    const int n_panels = (p.nsub > 1 ? p.sub_n : p.block_n) >> 6;   // staged epilogue: 64-channel panels per team's staging set
    uint8_t* smem_stage = smem_b + b_total_bytes;             // 2 sets x n_panels x 16 KB (staged epilogue only)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_stage + (p.staged ? 2 * p.res_sets * n_panels * kATileBytes : 0));
    uint64_t* empty_bar = full_bar + kMaxStages;
    uint64_t* tmem_full = empty_bar + kMaxStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* res_full = tmem_empty + 2;                     // [team * 2 + staging set]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(res_full + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    uint32_t crank = 0;
    if (p.cluster == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int wid = blockIdx.x / p.cluster;                    // work stream of this CTA (cluster index)
    const int wstep = gridDim.x / p.cluster;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.staged) {
            tma_prefetch_desc(&tmY);
            if (p.res_mode) tma_prefetch_desc(&tmR);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 2);                        // one arrive.expect_tx from each of the two producer warps
            mbar_init(&empty_bar[s], kCg2 ? 1 : p.cluster);   // multicast-B mode: released by the MMAs of both CTAs
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            // one arrive per draining warp; cta_group::2: the leader's MMA thread waits for the warps of both CTAs
            mbar_init(&tmem_empty[a], ((p.staged && p.nsub <= 1) ? kEpiWarps / 2 : kEpiWarps) * (kCg2 ? 2 : 1));
            mbar_init(&res_full[2 * a], 1);
            mbar_init(&res_full[2 * a + 1], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (kCg2) tmem_alloc_cg2(tmem_ptr_smem, kTmemCols);
        else tmem_alloc(tmem_ptr_smem, kTmemCols);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (p.cluster == 2) {                                      // the peer's barriers must exist before anything remote lands
        asm volatile("barrier.cluster.arrive.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    }
    tcgen05_fence_after();
    // The CTA allocates all 512 TMEM columns, so the allocation starts at column 0 / lane 0.  Using the literal 0
    // keeps every TMEM address warp-uniform for the compiler: with a value loaded from shared memory each
    // tcgen05.mma was wrapped in an ELECT / R2UR / branch "waterfall" costing ~75 cycles per instruction.
    if (*tmem_ptr_smem != 0u) {
        if (p.err_flag) atomicExch(p.err_flag, 7);
        asm volatile("trap;");
    }
    constexpr uint32_t tmem_base = 0u;

    const int kblocks = p.ngroups * p.kchunks;

    if (warp == 0 || warp == 3) {
        // ===================== TMA producers (whole warp loops, one elected lane issues) =====================
        // Two warps share the job -- warp 0 loads the activation boxes, warp 3 the weight tiles -- because the scalar
        // per-k-block bookkeeping of a single issuing thread (~100 dependent instructions) bounds the small layers.
        {
            const bool load_a = (warp == 0);
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_a = p.a_stage_bytes, tx_b = b_stage_bytes;
            const int half_rows = p.block_n >> 1;
            const uint32_t half_off = crank * static_cast<uint32_t>(half_rows * p.kb * 2);
            const int G = p.G, kchunks = p.kchunks, ngroups = p.ngroups, kbk = p.kb, nstages = p.stages;
            // ONE elected thread runs the whole loop: no per-k-block elect / reconvergence / warp sync in the issue path
            if (elect_one_sync())
            for (int tile = wid; tile < p.total_tiles; tile += wstep) {
                int nt, ph, img, oh0, ow0;
                decode_tile(p, tile, crank, nt, ph, img, oh0, ow0);
                CSBSR_TRACE(if (p.trace && blockIdx.x == 0 && load_a && tile / wstep < 256) p.trace[0 * 256 + tile / wstep] = clock64();)
                const int n0 = nt * p.block_n + (p.cluster == 2 ? static_cast<int>(crank) * half_rows : 0);
                for (int g = 0; g < ngroups; ++g) {
                    const int gi = (ph * ngroups + g) * G;              // first tap of the group
                    const int ih0 = oh0 * p.stride + p.dh[gi];
                    const int iw0 = ow0 * p.stride + p.dw[gi];
                    for (int kc = 0; kc < kchunks; ++kc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1u, p.err_flag, 1);
                        {
                            if constexpr (kCg2) {
                                // the leader's barrier collects the bytes of both CTAs: own A tile + peer A tile, two half B tiles
                                if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], load_a ? 2 * tx_a : 2 * tx_b);
                                if (load_a) {
                                    tma_load_4d_cg2(smem_u32(smem_a + stage * p.a_stage_bytes), &tmA, &full_bar[stage], kc * kbk, iw0,
                                                    ih0, img);
                                } else {
                                    const uint32_t dst0 = smem_u32(smem_b + stage * b_stage_bytes);
#pragma unroll 1
                                    for (int j = 0; j < G; ++j)
                                        tma_load_3d_cg2(dst0 + j * b_tile_bytes, &tmB, &full_bar[stage], kc * kbk, n0, p.widx[gi + j]);
                                }
                            } else if (load_a) {
                                mbar_arrive_expect_tx(&full_bar[stage], tx_a);
                                // one A box covers the G vertically shifted taps of the group (rows TH + (G-1)*dstep)
                                tma_load_4d(smem_u32(smem_a + stage * p.a_stage_bytes), &tmA, &full_bar[stage], kc * kbk, iw0, ih0,
                                            img);
                            } else {
                                mbar_arrive_expect_tx(&full_bar[stage], tx_b);
                                const uint32_t dst0 = smem_u32(smem_b + stage * b_stage_bytes);
                                if (p.nsub > 1) {
                                    // merged sub-phases: the N = block_n tile is spt weight slices of sub_n rows, one per sub-phase
                                    const uint32_t sub_bytes = static_cast<uint32_t>(p.sub_n * kbk * 2);
#pragma unroll 1
                                    for (int j = 0; j < G; ++j)
#pragma unroll 1
                                        for (int s2 = 0; s2 < p.spt; ++s2)
                                            tma_load_3d(dst0 + j * b_tile_bytes + s2 * sub_bytes, &tmB, &full_bar[stage], kc * kbk, 0,
                                                        p.widx[(gi + j) * p.nsub + nt * p.spt + s2]);
                                } else if (p.cluster == 2) {
                                    // this CTA fetches its half of every weight tile and multicasts it to both CTAs of the pair
#pragma unroll 1
                                    for (int j = 0; j < G; ++j)
                                        tma_load_3d_mc(dst0 + j * b_tile_bytes + half_off, &tmB, &full_bar[stage], kc * kbk, n0,
                                                       p.widx[gi + j], 0x3);
                                } else {
#pragma unroll 1
                                    for (int j = 0; j < G; ++j)
                                        tma_load_3d(dst0 + j * b_tile_bytes, &tmB, &full_bar[stage], kc * kbk, n0, p.widx[gi + j]);
                                }
                            }
                        }
                        if (++stage == nstages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp loops, one elected lane issues) =====================
        {
            const bool cg2 = kCg2;
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(p.block_n >> 3) << 17) |
                                   (static_cast<uint32_t>((cg2 ? 2 * kBlockM : kBlockM) >> 4) << 24);
            // descriptor words: lo = (smem address >> 4) | LBO(1) << 16 ; hi = SBO (8 rows of kb*2 bytes) | version 1 |
            // swizzle mode of the row width (128B: 2, 64B: 4, 32B: 6)
            const uint32_t row_bytes = static_cast<uint32_t>(p.kb) * 2u;
            const uint32_t desc_hi = ((8u * row_bytes) >> 4) | (1u << 14) | ((p.kb == 64 ? 2u : (p.kb == 32 ? 4u : 6u)) << 29);
            const int ksteps = p.kb >> 4;
            const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t a_stage16 = static_cast<uint32_t>(p.a_stage_bytes) >> 4;
            const uint32_t b_stage16 = static_cast<uint32_t>(b_stage_bytes) >> 4;
            const uint32_t a_shift16 = static_cast<uint32_t>(p.dstep * p.TW) * row_bytes >> 4;   // vertical tap step inside the A box
            const uint32_t b_tile16 = static_cast<uint32_t>(b_tile_bytes) >> 4;
            const int G = p.G;
            int stage = 0;
            uint32_t phase = 0, a_lo = a_lo0, b_lo = b_lo0;
            int local = 0;
            const int nstages = p.stages, total_tiles = p.total_tiles;
            const bool mc = p.cluster == 2;
            // ONE elected thread runs the whole issue loop (see the producers)
            if (elect_one_sync() && !(cg2 && crank != 0))       // cta_group::2: only the leader issues
            for (int tile = wid; tile < total_tiles; tile += wstep, ++local) {
                const int as = local & 1;
                const uint32_t aphase = (local >> 1) & 1;
                mbar_wait(&tmem_empty[as], aphase ^ 1u, p.err_flag, 2);
                tcgen05_fence_after();
                CSBSR_TRACE(if (p.trace && blockIdx.x == 0 && local < 256) p.trace[1 * 256 + local] = clock64();)
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * 256);
                uint32_t acc = 0;
                CSBSR_TRACE(long long wait_cyc = 0;)
                for (int kb = 0; kb < kblocks; ++kb) {
                    CSBSR_TRACE(const long long tw0 = p.trace ? clock64() : 0;)
                    mbar_wait(&full_bar[stage], phase, p.err_flag, 3);
                    tcgen05_fence_after();
                    CSBSR_TRACE(if (p.trace) wait_cyc += clock64() - tw0;)
                    {
                        uint32_t ja = a_lo, jb = b_lo;
#pragma unroll 1
                        for (int j = 0; j < G; ++j) {
                            // +32 bytes per K=16 step inside the 128B swizzle row -> +2 in the >>4 address field
                            if constexpr (kCg2) {
                                for (int s2 = 0; s2 < ksteps; ++s2)
                                    umma_bf16_lohi_cg2(tmem_d, ja + 2 * s2, jb + 2 * s2, desc_hi, idesc, s2 ? 1u : acc);
                            } else {
                            umma_bf16_lohi(tmem_d, ja, jb, desc_hi, idesc, acc);
                            if (ksteps == 4) {
                                umma_bf16_lohi(tmem_d, ja + 2, jb + 2, desc_hi, idesc, 1u);
                                umma_bf16_lohi(tmem_d, ja + 4, jb + 4, desc_hi, idesc, 1u);
                                umma_bf16_lohi(tmem_d, ja + 6, jb + 6, desc_hi, idesc, 1u);
                            } else if (ksteps == 2) {
                                umma_bf16_lohi(tmem_d, ja + 2, jb + 2, desc_hi, idesc, 1u);
                            }
                            }
                            acc = 1u;
                            ja += a_shift16;
                            jb += b_tile16;
                        }
                        if constexpr (kCg2) {
                            umma_commit_cg2_mc(&empty_bar[stage], 0x3);        // slot free / accumulator ready in both CTAs
                            if (kb == kblocks - 1) umma_commit_cg2_mc(&tmem_full[as], 0x3);
                        } else {
                            if (mc) umma_commit_mc(&empty_bar[stage], 0x3);   // frees the slot in both CTAs of the pair
                            else umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                            if (kb == kblocks - 1) umma_commit(&tmem_full[as]);
                        }
                    }
                    acc = 1u;
                    CSBSR_TRACE(if (kb == kblocks - 1 && p.trace && blockIdx.x == 0 && local < 256) {
                        p.trace[2 * 256 + local] = wait_cyc;
                        p.trace[3 * 256 + local] = clock64();
                    })
                    a_lo += a_stage16;
                    b_lo += b_stage16;
                    if (++stage == nstages) {
                        stage = 0;
                        phase ^= 1u;
                        a_lo = a_lo0;
                        b_lo = b_lo0;
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===================== epilogue: TMEM -> registers -> global =====================
        const int q = warp & 3;                       // TMEM lane quarter owned by this warp
        const int row = q * 32 + lane;                // accumulator row == pixel inside the tile
        const int th = row / p.TW, tw = row % p.TW;
        if (p.staged && p.nsub > 1) {
            // ---------------- staged epilogue, merged sub-phases: BOTH teams work on every tile, team t on the sub_n accumulator
            // columns (= output sub-phase) t of it; each team keeps one staging set and its own residual barrier ----------------
            const int team = (warp - 4) >> 3;
            const int wteam = (warp - 4) & 7;
            const bool leader = (wteam == 0 && lane == 0);
            const int units = p.sub_n >> 4;
            const int share = wteam >> 2;
            const int u_begin = (units * share) / 2, u_end = (units * (share + 1)) / 2;
            const uint32_t res_bytes = static_cast<uint32_t>(n_panels) * kATileBytes;
            uint8_t* stage_set = smem_stage + team * n_panels * kATileBytes;
            const uint32_t srow = smem_u32(stage_set) + static_cast<uint32_t>(row) * 128u;
            const uint32_t sw = static_cast<uint32_t>(row & 7) << 4;
            const int res_mode = p.res_mode;
            const float slope = p.slope, r1s = p.r1_sign;
            auto load_residual = [&](int tile) {
                int nt, ph, img, oh0, ow0;
                decode_tile(p, tile, crank, nt, ph, img, oh0, ow0);
                const int sp = ph * p.nsub + nt * p.spt + team;
                mbar_arrive_expect_tx(&res_full[2 * team], res_bytes);
                for (int pn = 0; pn < n_panels; ++pn) {
                    const uint32_t dst = smem_u32(stage_set + pn * kATileBytes);
                    if (p.os == 1) tma_load_4d(dst, &tmR, &res_full[2 * team], pn * 64, ow0, oh0, img);
                    else tma_load_5d(dst, &tmR, &res_full[2 * team], pn * 64, p.oow[sp], ow0, p.ooh[sp], img * p.OH + oh0);
                }
            };
            if (leader && res_mode && wid < p.total_tiles) load_residual(wid);
            int local = 0;
            for (int tile = wid; tile < p.total_tiles; tile += wstep, ++local) {
                const int as = local & 1;
                const uint32_t aphase = (local >> 1) & 1;
                int nt, ph, img, oh0, ow0;
                decode_tile(p, tile, crank, nt, ph, img, oh0, ow0);
                const int sp = ph * p.nsub + nt * p.spt + team;
                if (res_mode) {
                    mbar_wait(&res_full[2 * team], local & 1, p.err_flag, 5);
                } else {
                    if (leader) tma_store_wait_read<0>();       // the previous store of this team is done reading the set
                    team_bar_sync(team);
                }
                const int oy = (oh0 + th) * p.os + p.ooh[sp];
                const int ox = (ow0 + tw) * p.os + p.oow[sp];
                int cls = 0;
                if (p.cls_bw > 0)
                    cls = border_class(min(oy, p.YH - 1), p.YH, p.cls_bw) * (2 * p.cls_bw + 1) +
                          border_class(min(ox, p.YW - 1), p.YW, p.cls_bw);
                const float* bias_u = p.bias ? p.bias + static_cast<size_t>(min(img, p.N - 1)) * p.bias_sn +
                                                   static_cast<size_t>(cls) * p.bias_sc
                                             : nullptr;
                mbar_wait(&tmem_full[as], aphase, p.err_flag, 4);
                tcgen05_fence_after();
                const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * 256 + team * p.sub_n);
                switch (p.act) {
                    case CSBSR_ACT_RELU:    epilogue_staged_res<CSBSR_ACT_RELU>(res_mode, srow, sw, taddr0, u_begin, u_end, bias_u, slope, r1s); break;
                    case CSBSR_ACT_LEAKY:   epilogue_staged_res<CSBSR_ACT_LEAKY>(res_mode, srow, sw, taddr0, u_begin, u_end, bias_u, slope, r1s); break;
                    case CSBSR_ACT_SIGMOID: epilogue_staged_res<CSBSR_ACT_SIGMOID>(res_mode, srow, sw, taddr0, u_begin, u_end, bias_u, slope, r1s); break;
                    default:                epilogue_staged_res<CSBSR_ACT_NONE>(res_mode, srow, sw, taddr0, u_begin, u_end, bias_u, slope, r1s); break;
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[as]);
                fence_proxy_async_smem();                       // make the generic-proxy writes visible to the TMA store
                team_bar_sync(team);
                if (leader) {
                    for (int pn = 0; pn < n_panels; ++pn) {
                        const uint32_t src = smem_u32(stage_set + pn * kATileBytes);
                        if (p.os == 1) tma_store_4d(&tmY, src, pn * 64, ow0, oh0, img);
                        else tma_store_5d(&tmY, src, pn * 64, p.oow[sp], ow0, p.ooh[sp], img * p.OH + oh0);
                    }
                    tma_store_commit();
                    if (res_mode && tile + wstep < p.total_tiles) {
                        tma_store_wait_read<0>();
                        load_residual(tile + wstep);
                    }
                }
            }
            if (leader) tma_store_wait_read<0>();
        } else if (p.staged) {
            // ---------------- staged epilogue (see epilogue_staged) ----------------
            // Two teams of 8 warps; team t owns accumulator buffer t and staging set t and handles every other tile
            // of this CTA, so two tile epilogues are in flight and their latencies overlap.
            const int team = (warp - 4) >> 3;
            const int wteam = (warp - 4) & 7;
            const bool leader = (wteam == 0 && lane == 0);
            const int units = p.block_n >> 4;
            const int share = wteam >> 2;                   // the 2 warps of a lane quarter split the 16-column units
            const int u_begin = (units * share) / 2, u_end = (units * (share + 1)) / 2;
            const uint32_t res_bytes = static_cast<uint32_t>(n_panels) * kATileBytes;
            // staging sets of this team: nset = 2 (residual tiles only, when shared memory allows) alternates them per tile, so the
            // residual of the team's NEXT tile is in flight while the current one is processed
            const int nset = p.res_sets;
            uint8_t* team_sets = smem_stage + team * nset * n_panels * kATileBytes;
            const uint32_t sw = static_cast<uint32_t>(row & 7) << 4;
            const int res_mode = p.res_mode;
            const float slope = p.slope, r1s = p.r1_sign;
            const int as = team;
            auto load_residual = [&](int tile, int k) {
                int nt, ph, img, oh0, ow0;
                decode_tile(p, tile, crank, nt, ph, img, oh0, ow0);
                uint64_t* bar = &res_full[2 * team + k];
                mbar_arrive_expect_tx(bar, res_bytes);
                for (int pn = 0; pn < n_panels; ++pn) {
                    const uint32_t dst = smem_u32(team_sets + (k * n_panels + pn) * kATileBytes);
                    const int c = nt * p.block_n + pn * 64;
                    if (p.os == 1) tma_load_4d(dst, &tmR, bar, c, ow0, oh0, img);
                    else tma_load_5d(dst, &tmR, bar, c, p.oow[ph], ow0, p.ooh[ph], img * p.OH + oh0);
                }
            };
            const int stride_tiles = 2 * wstep;
            const int first_tile = wid + team * wstep;
            if (leader && p.res_mode && first_tile < p.total_tiles) {
                load_residual(first_tile, 0);
                if (nset == 2 && first_tile + stride_tiles < p.total_tiles) load_residual(first_tile + stride_tiles, 1);
            }
            int n_use = 0;                                  // how many tiles this team has processed
            for (int tile = first_tile; tile < p.total_tiles; tile += stride_tiles, ++n_use) {
                const uint32_t aphase = n_use & 1;
                const int kset = nset == 2 ? (n_use & 1) : 0;
                uint8_t* stage_set = team_sets + kset * n_panels * kATileBytes;
                const uint32_t srow = smem_u32(stage_set) + static_cast<uint32_t>(row) * 128u;   // this thread's row in panel 0
                const int local = 2 * n_use + team;
                (void)local;                                    // only the trace build reads it
                int nt, ph, img, oh0, ow0;
                CSBSR_TRACE(if (p.trace && blockIdx.x == 0 && leader && local < 256) p.trace[6 * 256 + local] = clock64();)
                decode_tile(p, tile, crank, nt, ph, img, oh0, ow0);
                if (p.res_mode) {
                    mbar_wait(&res_full[2 * team + kset], nset == 2 ? ((n_use >> 1) & 1) : aphase, p.err_flag, 5);
                } else {
                    if (leader) tma_store_wait_read<0>();       // the previous store of this team is done reading the set
                    team_bar_sync(team);
                }
                CSBSR_TRACE(if (p.trace && blockIdx.x == 0 && leader && local < 256) p.trace[7 * 256 + local] = clock64();)
                const int oy = (oh0 + th) * p.os + p.ooh[ph];
                const int ox = (ow0 + tw) * p.os + p.oow[ph];
                int cls = 0;
                if (p.cls_bw > 0)
                    cls = border_class(min(oy, p.YH - 1), p.YH, p.cls_bw) * (2 * p.cls_bw + 1) +
                          border_class(min(ox, p.YW - 1), p.YW, p.cls_bw);
                const float* bias_row = p.bias ? p.bias + static_cast<size_t>(min(img, p.N - 1)) * p.bias_sn +
                                                     static_cast<size_t>(cls) * p.bias_sc
                                               : nullptr;
                mbar_wait(&tmem_full[as], aphase, p.err_flag, 4);
                tcgen05_fence_after();
                CSBSR_TRACE(if (p.trace && blockIdx.x == 0 && leader && local < 256) p.trace[4 * 256 + local] = clock64();)
                const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * 256);
                const int c_base = nt * p.block_n;
                const float* bias_u = bias_row ? bias_row + c_base : nullptr;
                switch (p.act) {
                    case CSBSR_ACT_RELU:    epilogue_staged_res<CSBSR_ACT_RELU>(res_mode, srow, sw, taddr0, u_begin, u_end, bias_u, slope, r1s); break;
                    case CSBSR_ACT_LEAKY:   epilogue_staged_res<CSBSR_ACT_LEAKY>(res_mode, srow, sw, taddr0, u_begin, u_end, bias_u, slope, r1s); break;
                    case CSBSR_ACT_SIGMOID: epilogue_staged_res<CSBSR_ACT_SIGMOID>(res_mode, srow, sw, taddr0, u_begin, u_end, bias_u, slope, r1s); break;
                    default:                epilogue_staged_res<CSBSR_ACT_NONE>(res_mode, srow, sw, taddr0, u_begin, u_end, bias_u, slope, r1s); break;
                }
                CSBSR_TRACE(if (p.trace && blockIdx.x == 0 && leader && local < 256) p.trace[8 * 256 + local] = clock64();)
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (kCg2) mbar_arrive_cluster(&tmem_empty[as], 0);   // the leader's MMA thread owns both accumulators
                    else mbar_arrive(&tmem_empty[as]);
                }
                fence_proxy_async_smem();                       // make the generic-proxy writes visible to the TMA store
                CSBSR_TRACE(if (p.trace && blockIdx.x == 0 && leader && local < 256) p.trace[9 * 256 + local] = clock64();)
                team_bar_sync(team);
                if (leader) {
                    for (int pn = 0; pn < n_panels; ++pn) {
                        const uint32_t src = smem_u32(stage_set + pn * kATileBytes);
                        const int c = c_base + pn * 64;
                        if (p.os == 1) tma_store_4d(&tmY, src, c, ow0, oh0, img);
                        else tma_store_5d(&tmY, src, c, p.oow[ph], ow0, p.ooh[ph], img * p.OH + oh0);
                    }
                    tma_store_commit();
                    CSBSR_TRACE(if (p.trace && blockIdx.x == 0 && local < 256) p.trace[5 * 256 + local] = clock64();)
                    // the set is single-buffered per team: its next residual tile can only land once the store has
                    // finished reading; the other team's tile hides this latency
                    if (p.res_mode && tile + nset * stride_tiles < p.total_tiles) {
                        tma_store_wait_read<0>();               // the set can only be refilled once the store has read it
                        load_residual(tile + nset * stride_tiles, kset);
                    }
                }
            }
            if (leader) tma_store_wait_read<0>();
        } else {
        int local = 0;
        for (int tile = wid; tile < p.total_tiles; tile += wstep, ++local) {
            const int as = local & 1;
            const uint32_t aphase = (local >> 1) & 1;
            int nt, ph, img, oh0d, ow0d;
            decode_tile(p, tile, crank, nt, ph, img, oh0d, ow0d);
            const int oh = oh0d + th;
            const int ow = ow0d + tw;
            const bool valid = (oh < p.OH) && (ow < p.OW) && (img < p.N);
            const int oy = oh * p.os + p.ooh[ph];
            const int ox = ow * p.os + p.oow[ph];
            const size_t pix = (static_cast<size_t>(img) * p.YH + oy) * p.YW + ox;
            int cls = 0;
            if (p.cls_bw > 0)
                cls = border_class(oy, p.YH, p.cls_bw) * (2 * p.cls_bw + 1) + border_class(ox, p.YW, p.cls_bw);
            const float* bias_row = p.bias ? p.bias + static_cast<size_t>(min(img, p.N - 1)) * p.bias_sn +
                                                 static_cast<size_t>(cls) * p.bias_sc
                                           : nullptr;

            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * 256);
            const int c_base = nt * p.block_n;
            const int units = p.block_n >> 4;               // 16-column units, interleaved between the 2 warps
            const int u0 = (warp - 4) >> 2;
            EpiRow er;
            er.valid = valid; er.pix = pix; er.img = img; er.oy = oy; er.ox = ox; er.bias_row = bias_row;
            mbar_wait(&tmem_full[as], aphase, p.err_flag, 4);
            tcgen05_fence_after();
            if (p.out_mode == CSBSR_OUT_BF16_NHWC) {
                switch (p.act) {
                    case CSBSR_ACT_RELU:    epilogue_bf16<CSBSR_ACT_RELU>(p, er, taddr0, c_base, u0, units); break;
                    case CSBSR_ACT_LEAKY:   epilogue_bf16<CSBSR_ACT_LEAKY>(p, er, taddr0, c_base, u0, units); break;
                    case CSBSR_ACT_SIGMOID: epilogue_bf16<CSBSR_ACT_SIGMOID>(p, er, taddr0, c_base, u0, units); break;
                    default:                epilogue_bf16<CSBSR_ACT_NONE>(p, er, taddr0, c_base, u0, units); break;
                }
            } else {
                switch (p.act) {
                    case CSBSR_ACT_RELU:    epilogue_units<CSBSR_ACT_RELU>(p, er, taddr0, c_base, u0, units); break;
                    case CSBSR_ACT_LEAKY:   epilogue_units<CSBSR_ACT_LEAKY>(p, er, taddr0, c_base, u0, units); break;
                    case CSBSR_ACT_SIGMOID: epilogue_units<CSBSR_ACT_SIGMOID>(p, er, taddr0, c_base, u0, units); break;
                    default:                epilogue_units<CSBSR_ACT_NONE>(p, er, taddr0, c_base, u0, units); break;
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (kCg2) mbar_arrive_cluster(&tmem_empty[as], 0);
                else mbar_arrive(&tmem_empty[as]);
            }
        }
        }  // direct epilogue
    }

    tcgen05_fence_before();
    __syncthreads();
    if (p.cluster == 2) {                                      // nobody leaves while the peer may still signal its barriers
        asm volatile("barrier.cluster.arrive.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    }
    if (warp == 2) {
        tcgen05_fence_after();
        if constexpr (kCg2) tmem_dealloc_cg2(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// Launch plans (encoded tensor maps + kernel parameters + launch geometry) are cached per descriptor: a training or eval step
// repeats the same few hundred descriptors every iteration, and encoding 2-4 CUtensorMaps plus the tap-grouping search per
// launch cost more host time than the launch itself (round-1 VERDICT: eager training 17.5 vs 21.0 steps/s graphed).  The key is
// the descriptor's bytes plus the environment switches that influence the plan; the pointers inside make stale hits impossible
// (a tensor map only encodes address, shape and strides).
struct ConvPlan {
    csbsr_conv_desc d;
    uint32_t envsig;
    CUtensorMap tmA, tmB, tmY, tmR;
    ConvKParams p;
    int smem_bytes, grid, cluster, cg2;
};
static std::unordered_map<uint64_t, std::vector<ConvPlan>> g_plans;
static size_t g_plan_count = 0;
static std::mutex g_plan_mutex;

static uint64_t plan_hash(const csbsr_conv_desc* d, uint32_t envsig) {
    uint64_t h = 1469598103934665603ull ^ envsig;
    const unsigned char* b = reinterpret_cast<const unsigned char*>(d);
    for (size_t i = 0; i < sizeof(*d); ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
static uint32_t plan_envsig() {
    uint32_t e = 0;
    const char* v;
    if ((v = getenv("CSBSR_CLUSTER"))) e |= (1u + (static_cast<uint32_t>(atoi(v)) & 3u));
    if ((v = getenv("CSBSR_CTA_GROUP"))) e |= (1u + (static_cast<uint32_t>(atoi(v)) & 3u)) << 4;
    if (getenv("CSBSR_NO_STAGED")) e |= 1u << 8;
    if (getenv("CSBSR_NO_GROUPING")) e |= 1u << 9;
    if (getenv("CSBSR_NO_RES_PREFETCH")) e |= 1u << 10;
    return e;
}
static int launch_plan(const ConvPlan& pl, cudaStream_t stream) {
    if (pl.cluster == 1) {
        conv_igemm_kernel<false><<<pl.grid, kThreads, pl.smem_bytes, stream>>>(pl.tmA, pl.tmB, pl.tmY, pl.tmR, pl.p);
    } else {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(pl.grid, 1, 1);
        cfg.blockDim = dim3(kThreads, 1, 1);
        cfg.dynamicSmemBytes = pl.smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (pl.cg2) CSBSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<true>, pl.tmA, pl.tmB, pl.tmY, pl.tmR, pl.p));
        else CSBSR_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_igemm_kernel<false>, pl.tmA, pl.tmB, pl.tmY, pl.tmR, pl.p));
    }
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static int* g_err_flag = nullptr;
static long long* g_trace = nullptr;   // debug timeline of CTA 0 (CSBSR_CONV_TRACE=1): [10][256] clock64 stamps   // device int, lazily allocated (one per process; diagnostic only)

}  // namespace csbsr

using namespace csbsr;

extern "C" int csbsr_conv_igemm(const csbsr_conv_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CSBSR_REQUIRE(d && d->x && d->wgt && d->y, "conv_igemm: null pointer");
    const bool cacheable = !getenv("CSBSR_CONV_TRACE") && !getenv("CSBSR_NO_PLAN_CACHE");
    const uint32_t envsig = plan_envsig();
    const uint64_t phash = plan_hash(d, envsig);
    if (cacheable) {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        auto it = g_plans.find(phash);
        if (it != g_plans.end())
            for (const ConvPlan& pl : it->second)
                if (pl.envsig == envsig && memcmp(&pl.d, d, sizeof(*d)) == 0) return launch_plan(pl, stream);
    }
    CSBSR_REQUIRE(d->cin > 0 && (d->cin % kBlockK == 0 || d->cin == 32 || d->cin == 16),
                  "conv_igemm: cin=%d must be a positive multiple of 64, or 32, or 16", d->cin);
    const int kb = d->cin % kBlockK == 0 ? kBlockK : d->cin;         // K elements per chunk
    const int a_tile_bytes = kBlockM * kb * 2;
    CSBSR_REQUIRE(d->cout_pad > 0 && d->cout_pad % 16 == 0, "conv_igemm: cout_pad=%d must be a multiple of 16",
                  d->cout_pad);
    CSBSR_REQUIRE(d->x_pitch % 8 == 0 && d->x_coff % 8 == 0, "conv_igemm: x pitch/offset must be multiples of 8");
    CSBSR_REQUIRE(d->nphases >= 1 && d->nphases <= CSBSR_MAX_PHASES && d->ntaps >= 1 &&
                      d->nphases * d->ntaps <= CSBSR_MAX_TAPS,
                  "conv_igemm: bad phase/tap counts (%d x %d)", d->nphases, d->ntaps);
    CSBSR_REQUIRE(d->stride >= 1 && d->stride <= 8, "conv_igemm: bad stride %d", d->stride);
    CSBSR_REQUIRE(d->n >= 1 && d->oh >= 1 && d->ow >= 1, "conv_igemm: empty output");
    if (d->out_mode == CSBSR_OUT_BF16_NHWC) {
        CSBSR_REQUIRE(d->cout_store % 8 == 0 && d->y_pitch % 8 == 0 && d->y_coff % 8 == 0,
                      "conv_igemm: bf16 output needs cout_store/pitch/offset multiples of 8");
        CSBSR_REQUIRE(!d->r32, "conv_igemm: r32 only with f32 planar output");
    } else if (d->out_mode == CSBSR_OUT_F32_NHWC) {
        CSBSR_REQUIRE(d->cout_store % 4 == 0 && d->y_pitch % 4 == 0 && d->y_coff % 4 == 0,
                      "conv_igemm: f32 NHWC output needs cout_store/pitch/offset multiples of 4");
        CSBSR_REQUIRE(!d->r0 && !d->rm && !d->r1 && !d->r32, "conv_igemm: no residuals with f32 NHWC output");
    } else {
        CSBSR_REQUIRE(d->cout_store >= 1 && d->cout_store <= 16, "conv_igemm: f32 planar output needs cout_store<=16");
        CSBSR_REQUIRE(!d->r0 && !d->rm && !d->r1, "conv_igemm: bf16 residuals only with bf16 output");
    }
    CSBSR_REQUIRE(d->cout_store <= d->cout_pad, "conv_igemm: cout_store > cout_pad");
    for (int i = 0; i < d->nphases * d->ntaps * (d->nsub > 1 ? d->nsub : 1); ++i)
        CSBSR_REQUIRE(d->widx[i] >= 0 && d->widx[i] < d->w_taps, "conv_igemm: widx[%d]=%d out of range", i, d->widx[i]);

    PFN_encodeTiled encode = get_encode_fn();
    CSBSR_REQUIRE(encode, "conv_igemm: cuTensorMapEncodeTiled entry point unavailable");

    ConvKParams p;
    memset(&p, 0, sizeof(p));
    int block_n = d->block_n;
    const int nsub = d->nsub > 1 ? d->nsub : 1;
    if (nsub > 1) {
        CSBSR_REQUIRE(d->cout_pad == 128 && nsub % 2 == 0 && d->nphases * nsub <= CSBSR_MAX_PHASES &&
                          d->nphases * d->ntaps * nsub <= CSBSR_MAX_TAPS && kb == kBlockK,
                      "conv_igemm: merged sub-phases need cout_pad = 128, an even nsub and cin %% 64 == 0");
        block_n = 256;
    }
    const int TWh = d->ow > 8 ? 16 : 8, THh = kBlockM / TWh;
    const int os_h = d->os > 0 ? d->os : 1;
    // staged epilogue (smem panels + TMA residual load + TMA store): bf16 output, at most one residual operand,
    // 64/128-column tiles; strided (deconv) outputs additionally need whole tiles per image for the merged 5-D map
    bool can_stage = d->out_mode == CSBSR_OUT_BF16_NHWC && !d->rm && !(d->r0 && d->r1) && d->cout_pad % 64 == 0 &&
                     (os_h == 1 || (d->yh == d->oh * os_h && d->yw == d->ow * os_h && d->oh % THh == 0));
    if (getenv("CSBSR_NO_STAGED")) can_stage = false;
    const bool prefer_stage = can_stage && (d->r0 || d->r1 || d->ntaps * d->cin <= 1024 || d->cout_pad <= 128);
    if (block_n <= 0) {
        if (prefer_stage) {
            block_n = d->cout_pad % 128 == 0 ? 128 : 64;
        } else {
            // widest tile (<= 256 columns, multiple of 16) that divides cout_pad: fewer re-reads of the A operand
            for (block_n = 256; block_n > 16; block_n -= 16)
                if (d->cout_pad % block_n == 0) break;
        }
    }
    const bool staged = can_stage && (block_n == 64 || block_n == 128 || nsub > 1);
    CSBSR_REQUIRE(nsub == 1 || can_stage, "conv_igemm: merged sub-phases need the staged epilogue (bf16 NHWC output, whole tiles)");
    CSBSR_REQUIRE(block_n % 16 == 0 && block_n >= 16 && block_n <= 256 && (d->cout_pad * nsub) % block_n == 0,
                  "conv_igemm: block_n=%d incompatible with cout_pad=%d", block_n, d->cout_pad);
    p.block_n = block_n;
    p.TW = d->ow > 8 ? 16 : 8;
    p.TH = kBlockM / p.TW;
    p.tiles_h = (d->oh + p.TH - 1) / p.TH;
    p.tiles_w = (d->ow + p.TW - 1) / p.TW;
    p.m_tiles = d->n * p.tiles_h * p.tiles_w;
    p.n_tiles = d->cout_pad * nsub / block_n;
    p.nsub = nsub; p.sub_n = d->cout_pad; p.spt = nsub > 1 ? block_n / d->cout_pad : 1;
    p.nphases = d->nphases;
    // CTA pairs with multicast weight tiles: worth it when the weight tile of a stage is at least as large as the activation
    // tile (the weight stream dominates the L2 -> SM traffic) and there are enough pixel tiles to keep every pair busy
    const char* cl_env = getenv("CSBSR_CLUSTER");
    int cluster = 1;
    if (cl_env) cluster = (atoi(cl_env) == 2 && block_n % 16 == 0 && p.m_tiles >= 2 && nsub == 1) ? 2 : 1;
    // cta_group::2 (CTA pairs, M = 256 per instruction, each CTA holds half of the weight rows -> smaller pipeline stages,
    // more of them in flight): measured -9..-23 % on the layers with K = taps x cin >= 1152 and at least a wave of pixel
    // tiles (8x8/s4 convs, SFT / PSP / ResNet 3x3s), +10..+40 % on short-K or tiny layers (profiles/r02_cg2_layers.md), so
    // it is switched on by that rule -- stated per IMAGE (>= 16 pixel tiles = 2048 output pixels), not per launch: M = 256
    // instructions round differently from M = 128 ones, and the result of an image must not depend on how many other images
    // share its launch (rank shards of a batch reproduce the single-process result bit for bit, tests/test_multigpu_gpu.py).
    // CSBSR_CTA_GROUP=1 / 2 forces it off / on.
    const char* cg_env = getenv("CSBSR_CTA_GROUP");
    int cg2 = 0;
    const bool cg2_ok = block_n % 32 == 0 && p.m_tiles >= 2 && nsub == 1;
    if (cg_env) cg2 = (atoi(cg_env) == 2 && cg2_ok) ? 1 : 0;
    else cg2 = (cg2_ok && d->ntaps * d->cin >= 1024 && p.tiles_h * p.tiles_w >= 16) ? 1 : 0;
    if (cg2) cluster = 2;
    p.cluster = cluster;
    p.cg2 = cg2;
    p.total_tiles = ((p.m_tiles + cluster - 1) / cluster) * p.n_tiles * p.nphases;
    p.fd_ntiles = make_fastdiv(p.n_tiles); p.fd_nphases = make_fastdiv(p.nphases);
    p.fd_tiles_per_img = make_fastdiv(p.tiles_h * p.tiles_w); p.fd_tiles_w = make_fastdiv(p.tiles_w);
    p.kchunks = d->cin / kb;
    p.kb = kb;
    p.ntaps = d->ntaps;
    p.stride = d->stride;
    // ---- tap grouping: taps of one phase that share dw and whose dh are dh0 + j*step (step a multiple of the conv
    // stride) are served by ONE A box of TH + (G-1)*step/stride rows: cuts the L2 -> smem traffic of the A operand.
    int G = 1, ngroups = d->ntaps, dstep = 0;
    int8_t g_dh[CSBSR_MAX_TAPS], g_dw[CSBSR_MAX_TAPS];
    int16_t g_widx[CSBSR_MAX_TAPS];
    memcpy(g_dh, d->dh, sizeof(g_dh)); memcpy(g_dw, d->dw, sizeof(g_dw)); memcpy(g_widx, d->widx, sizeof(g_widx));
    const int staging_bytes = staged ? 2 * ((nsub > 1 ? d->cout_pad : block_n) / 64) * kATileBytes : 0;
    const int smem_avail = kSmemBudget - 3072 - staging_bytes;
    const int b_tile = (cg2 ? block_n / 2 : block_n) * kb * 2;      // per-CTA weight tile of one tap and K chunk
    int cand = 1, cand_groups = d->ntaps, cand_step = 0;
    int8_t n_dh[CSBSR_MAX_TAPS], n_dw[CSBSR_MAX_TAPS];
    int16_t n_widx[CSBSR_MAX_TAPS];
    if (d->ntaps >= 2 && nsub == 1 && !getenv("CSBSR_NO_GROUPING")) {
        const int st = d->stride;
        auto key_of = [&](int t) { return d->dw[t] * 64 + (((d->dh[t] % st) + st) % st); };
        bool ok = true;
        int ng0 = -1, gsz = -1, step = 0;
        for (int ph = 0; ph < d->nphases && ok; ++ph) {
            int keys[CSBSR_MAX_TAPS], nk = 0;
            for (int t = 0; t < d->ntaps; ++t) {
                const int k = key_of(ph * d->ntaps + t);
                bool seen = false;
                for (int i = 0; i < nk; ++i) seen |= (keys[i] == k);
                if (!seen) keys[nk++] = k;
            }
            if (ng0 < 0) ng0 = nk;
            if (nk != ng0 || d->ntaps % nk != 0) { ok = false; break; }
            if (gsz < 0) gsz = d->ntaps / nk;
            for (int gi = 0; gi < nk && ok; ++gi) {
                int idx[CSBSR_MAX_TAPS], cnt = 0;
                for (int t = 0; t < d->ntaps; ++t)
                    if (key_of(ph * d->ntaps + t) == keys[gi]) idx[cnt++] = ph * d->ntaps + t;
                if (cnt != gsz) { ok = false; break; }
                for (int a = 0; a < cnt; ++a)
                    for (int b2 = a + 1; b2 < cnt; ++b2)
                        if (d->dh[idx[b2]] < d->dh[idx[a]]) { int tmp = idx[a]; idx[a] = idx[b2]; idx[b2] = tmp; }
                for (int j = 0; j < cnt; ++j) {
                    if (j > 0) {
                        const int sd = d->dh[idx[j]] - d->dh[idx[j - 1]];
                        if (sd <= 0 || sd % st != 0 || (step != 0 && sd != step)) ok = false;
                        step = sd;
                    }
                    const int o = (ph * nk + gi) * gsz + j;
                    n_dh[o] = d->dh[idx[j]]; n_dw[o] = d->dw[idx[j]]; n_widx[o] = d->widx[idx[j]];
                }
            }
        }
        if (ok && gsz >= 2 && (THh + (gsz - 1) * (step / st)) * st <= 256) {
            cand = gsz; cand_groups = ng0; cand_step = step / st;
        }
    }
    int a_stage_bytes = a_tile_bytes, stages = 0;
    {
        // grouped taps when they leave >= 3 pipeline stages, else one tap per stage
        const int a_b = (THh + (cand - 1) * cand_step) * TWh * kb * 2;
        const int st_g = cand > 1 ? smem_avail / (a_b + cand * b_tile) : 0;
        if (cand > 1 && st_g >= 3) {
            G = cand; ngroups = cand_groups; dstep = cand_step; a_stage_bytes = a_b; stages = st_g;
            memcpy(g_dh, n_dh, sizeof(g_dh)); memcpy(g_dw, n_dw, sizeof(g_dw)); memcpy(g_widx, n_widx, sizeof(g_widx));
        } else {
            stages = smem_avail / (a_tile_bytes + b_tile);
        }
    }
    const int stage_bytes = a_stage_bytes + G * b_tile;
    if (stages > kMaxStages) stages = kMaxStages;
    // a second staging set per team for residual tiles, when it still leaves 3 pipeline stages (short-K layers: the 3 -> 128
    // transposed conv added in place to the 448^2 feature slice is a pure read-modify-write stream of residual tiles)
    int res_sets = 1;
    if (staged && nsub == 1 && (d->r0 || d->r1) && !getenv("CSBSR_NO_RES_PREFETCH")) {
        const int st2 = (smem_avail - staging_bytes) / stage_bytes;
        if (st2 >= 3) {
            res_sets = 2;
            if (stages > st2) stages = st2;
        }
    }
    p.res_sets = res_sets;
    CSBSR_REQUIRE(stages >= 2, "conv_igemm: not enough shared memory for 2 stages");
    p.stages = stages;
    p.OH = d->oh; p.OW = d->ow; p.os = d->os > 0 ? d->os : 1; p.YH = d->yh; p.YW = d->yw; p.N = d->n;
    p.out_mode = d->out_mode; p.y_pitch = d->y_pitch; p.y_coff = d->y_coff; p.cout_store = d->cout_store;
    p.bias_sn = d->bias_sn; p.bias_sc = d->bias_sc; p.cls_bw = d->cls_bw; p.act = d->act;
    p.slope = d->slope; p.r1_sign = d->r1_sign;
    p.r0_pitch = d->r0_pitch; p.r0_coff = d->r0_coff; p.rm_pitch = d->rm_pitch; p.rm_coff = d->rm_coff;
    p.r1_pitch = d->r1_pitch; p.r1_coff = d->r1_coff;
    p.y = d->y; p.bias = d->bias;
    p.r0 = reinterpret_cast<const __nv_bfloat16*>(d->r0);
    p.rm = reinterpret_cast<const __nv_bfloat16*>(d->rm);
    p.r1 = reinterpret_cast<const __nv_bfloat16*>(d->r1);
    p.r32 = d->r32;
    if (d->out_mode == CSBSR_OUT_F32_NCHW) {
        // planar: pitch = channels of the destination / residual buffers (0 -> exactly cout_store), coff = first channel
        p.y_pitch = d->y_pitch > 0 ? d->y_pitch : d->cout_store;
        p.r32_pitch = d->r32_pitch > 0 ? d->r32_pitch : d->cout_store;
        p.r32_coff = d->r32_coff;
    }
    p.staged = staged ? 1 : 0;
    p.res_mode = staged ? (d->r0 ? 1 : (d->r1 ? 2 : 0)) : 0;
    memcpy(p.dh, g_dh, sizeof(p.dh)); memcpy(p.dw, g_dw, sizeof(p.dw)); memcpy(p.widx, g_widx, sizeof(p.widx));
    p.G = G; p.ngroups = ngroups; p.dstep = dstep; p.a_stage_bytes = a_stage_bytes;
    memcpy(p.ooh, d->ooh, sizeof(p.ooh)); memcpy(p.oow, d->oow, sizeof(p.oow));
    CSBSR_REQUIRE(!d->bias || (d->bias_sn % 4 == 0 && d->bias_sc % 4 == 0), "conv_igemm: bias strides must be multiples of 4");
    for (int ph = 0; ph < p.nphases * nsub; ++ph)
        CSBSR_REQUIRE((p.OH - 1) * p.os + p.ooh[ph] < p.YH && (p.OW - 1) * p.os + p.oow[ph] < p.YW && p.ooh[ph] >= 0 &&
                          p.oow[ph] >= 0,
                      "conv_igemm: phase %d writes outside the %dx%d output", ph, p.YH, p.YW);

    if (!g_err_flag) {
        CSBSR_CHECK_CUDA(cudaMalloc(&g_err_flag, sizeof(int)));
        CSBSR_CHECK_CUDA(cudaMemset(g_err_flag, 0, sizeof(int)));
    }
    p.err_flag = g_err_flag;
    p.trace = nullptr;
    if (getenv("CSBSR_CONV_TRACE")) {
        if (!g_trace) CSBSR_CHECK_CUDA(cudaMalloc(&g_trace, sizeof(long long) * 10 * 256));
        CSBSR_CHECK_CUDA(cudaMemsetAsync(g_trace, 0, sizeof(long long) * 10 * 256, stream));
        p.trace = g_trace;
    }

    // ---- tensor maps
    const CUtensorMapSwizzle swz_in = kb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                               : (kb == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    CUtensorMap tmA, tmB;
    {
        const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(d->x) + d->x_coff;
        cuuint64_t dims[4] = {(cuuint64_t)d->cin, (cuuint64_t)d->w, (cuuint64_t)d->h, (cuuint64_t)d->n};
        cuuint64_t strides[3] = {(cuuint64_t)d->x_pitch * 2, (cuuint64_t)d->x_pitch * 2 * d->w,
                                 (cuuint64_t)d->x_pitch * 2 * d->w * d->h};
        cuuint32_t box[4] = {(cuuint32_t)kb, (cuuint32_t)(p.TW * d->stride),
                             (cuuint32_t)((p.TH + (G - 1) * dstep) * d->stride), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)d->stride, (cuuint32_t)d->stride, 1};
        CSBSR_REQUIRE(box[1] <= 256 && box[2] <= 256, "conv_igemm: TMA box too large");
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz_in, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CSBSR_REQUIRE(r == CUDA_SUCCESS, "conv_igemm: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)d->cin, (cuuint64_t)d->cout_pad, (cuuint64_t)d->w_taps};
        cuuint64_t strides[2] = {(cuuint64_t)d->cin * 2, (cuuint64_t)d->cin * 2 * d->cout_pad};
        cuuint32_t box[3] = {(cuuint32_t)kb, (cuuint32_t)(nsub > 1 ? d->cout_pad : block_n / cluster), 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)d->wgt, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz_in, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CSBSR_REQUIRE(r == CUDA_SUCCESS, "conv_igemm: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
    }

    CUtensorMap tmY, tmR;
    memset(&tmY, 0, sizeof(tmY));
    memset(&tmR, 0, sizeof(tmR));
    if (staged) {
        auto encode_out = [&](CUtensorMap* tm, const void* base_ptr, int pitch, int coff) -> int {
            const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(base_ptr) + coff;
            const cuuint64_t pb = (cuuint64_t)pitch * 2;
            CUresult r;
            if (p.os == 1) {
                cuuint64_t dims[4] = {(cuuint64_t)d->cout_store, (cuuint64_t)p.YW, (cuuint64_t)p.YH, (cuuint64_t)d->n};
                cuuint64_t strides[3] = {pb, pb * p.YW, pb * p.YW * p.YH};
                cuuint32_t box[4] = {64, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
                cuuint32_t estr[4] = {1, 1, 1, 1};
                r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            } else {
                // pixel (n, oh*os+ooh, ow*os+oow) -> 5-D view (c, oow, ow, ooh, n*OH+oh)
                cuuint64_t dims[5] = {(cuuint64_t)d->cout_store, (cuuint64_t)p.os, (cuuint64_t)p.OW, (cuuint64_t)p.os,
                                      (cuuint64_t)d->n * p.OH};
                cuuint64_t strides[4] = {pb, pb * p.os, pb * p.YW, pb * p.YW * p.os};
                cuuint32_t box[5] = {64, 1, (cuuint32_t)p.TW, 1, (cuuint32_t)p.TH};
                cuuint32_t estr[5] = {1, 1, 1, 1, 1};
                r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            }
            return (int)r;
        };
        int r = encode_out(&tmY, d->y, d->y_pitch, d->y_coff);
        CSBSR_REQUIRE(r == 0, "conv_igemm: cuTensorMapEncodeTiled(Y) failed with %d", r);
        if (p.res_mode == 1) r = encode_out(&tmR, d->r0, d->r0_pitch, d->r0_coff);
        if (p.res_mode == 2) r = encode_out(&tmR, d->r1, d->r1_pitch, d->r1_coff);
        CSBSR_REQUIRE(r == 0, "conv_igemm: cuTensorMapEncodeTiled(R) failed with %d", r);
    }

    const int smem_bytes = stages * stage_bytes + staging_bytes * res_sets + 2048 /*align slack (base + staging)*/ + 512 /*barriers*/;
    static int smem_attr_set = 0;
    if (smem_attr_set < smem_bytes) {
        CSBSR_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              kSmemBudget));
        CSBSR_CHECK_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              kSmemBudget));
        smem_attr_set = kSmemBudget;
    }
    ConvPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.d = *d; pl.envsig = envsig;
    pl.tmA = tmA; pl.tmB = tmB; pl.tmY = tmY; pl.tmR = tmR;
    pl.p = p;
    pl.smem_bytes = smem_bytes; pl.cluster = cluster; pl.cg2 = cg2;
    pl.grid = p.total_tiles * cluster < num_sms() ? p.total_tiles * cluster : (num_sms() / cluster) * cluster;
    if (cacheable) {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        if (g_plan_count >= 16384) { g_plans.clear(); g_plan_count = 0; }       // bounded: start over (descriptors of a step: a few hundred)
        g_plans[phash].push_back(pl);
        ++g_plan_count;
    }
    return launch_plan(pl, stream);
}

// debug only (not part of the public header): copies the CTA-0 timeline recorded under CSBSR_CONV_TRACE=1
extern "C" int csbsr_conv_trace_read(long long* host_out) {
    if (!csbsr::g_trace) return -1;
    return cudaMemcpy(host_out, csbsr::g_trace, sizeof(long long) * 10 * 256, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
