// Fused kernel-predictor chains of KBPN's KBlock (reference model/modeling/kbpn.py:521-578, KernelPredictorLikeIKC):
//
//   csbsr_kpred_sr_chain :  sr_t (fp32, 3 planes) -> fe_SR.0 3x3 3->49 ReLU -> fe_SR.1 1x1 49->32 lrelu -> fe_SR.2 3x3 32->32
//                           lrelu -> fe_SR.3 3x3 32->32 lrelu -> fe_SR.4 3x3 32->49 lrelu -> bf16 NHWC [B,H,W,64]
//   csbsr_kpred_cat_chain:  that map -> fe_cat.0 1x1 (49 image channels; the 49 kernel-branch channels enter as a
//                           per-sample 5x5 border-class bias) lrelu -> fe_cat.1 3x3 32->32 lrelu -> fe_cat.2 3x3 32->49
//                           -> global average pool -> fp32 [B,64]
//
// At 448^2 these 32..64-channel layers are HBM / L2 sized when run one conv launch at a time (20 + 12 launches per 8
// images, 2.2 GB of activation traffic per stage).  Here every intermediate stays in shared memory:
//
//   * a CTA owns a 128-pixel wide column strip of one image (112..122 valid output columns + the halo the 3x3 layers
//     consume) and streams down its rows.  One image row of the strip is one MMA tile: M = 128 pixels.
//   * activations live in shared memory as 8-channel "planes" [plane][pixel][8 x bf16] (the un-swizzled K-major UMMA
//     layout with the 8-row group pitch equal to 8 x 16 B): a pixel shift is +16 B on the descriptor start address, so
//     the 9 taps of a 3x3 layer are 9 descriptors into a 3-row ring buffer -- no im2col, no halo re-load.
//   * one elected thread issues tcgen05.mma (M128, N = 32 / 64, K = 16) for all layers of a step, deepest layer first;
//     the layers are software-pipelined over rows (layer k works on row t - lag_k), so every unit of a step only
//     depends on epilogues of the previous step and the tensor pipe never waits for the epilogue warps.
//   * accumulators sit in TMEM (224 columns); 8 epilogue warps (two groups) read them with tcgen05.ld, apply the
//     activation, zero what lies outside the image (the next layer's zero padding), round to bf16 and write the next
//     layer's planes (conflict-free: consecutive lanes = consecutive pixels = consecutive 16-byte words).
//   * weights (80 KB / 58 KB, pre-packed by the host into the same plane layout) are copied to shared memory once.
//
// In-order completion of tcgen05.mma makes the ring buffers minimal: an epilogue only starts when its own unit's MMAs
// (issued after every MMA of the previous step) have retired, so a 3x3-consumed ring needs 3 rows and a pointwise one 1.
#include <cuda.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/csbsr_b200.h"

namespace csbsr {
namespace kp {

constexpr int kGapRows = 32;              // image rows per pooled partial sum (fixed: the reduction order is launch-independent)
constexpr int kPX = 128;                  // pixels per row unit = MMA M
constexpr int kPlanePx = 136;             // pad | 128 pixels | pad ... : the +-1 column taps read the pads at the strip ends; 136 makes a
                                          // plane 17 x 128 B, the alignment TMA needs for its shared-memory destination
constexpr int kPlaneB = kPlanePx * 16;    // bytes of one 8-channel plane of one row slot
constexpr int kThreads = 512;             // warp 0: MMA issue, warp 1: TMEM alloc + TMA, warps 4-7: im2col builder, warps 8-15: epilogue
constexpr int kG0 = 3;                    // first global step index (multiple of 3: ring slots are step % 3)

enum { IN_IM2COL = 0, IN_TMA = 1 };
enum { OUT_BUF = 0, OUT_GLOBAL = 1, OUT_GAP = 2 };

// ------------------------------------------------------------------ the two layer programs
struct ProgSR {      // fe_SR.0 .. fe_SR.4
    static constexpr int NU = 5, INPUT = IN_IM2COL, IN_SLOTS = 1, IN_PLANES = 4;
    __host__ __device__ static constexpr int taps(int u) { return u < 2 ? 1 : 9; }
    __host__ __device__ static constexpr int kpl(int u) { return u == 1 ? 8 : 4; }            // input channels / 8
    __host__ __device__ static constexpr int nout(int u) { return (u == 0 || u == 4) ? 64 : 32; }
    __host__ __device__ static constexpr int act(int u) { return u == 0 ? CSBSR_ACT_RELU : CSBSR_ACT_LEAKY; }
    __host__ __device__ static constexpr int lag(int u) { return u == 0 ? 0 : 2 * u - 1; }    // 0 1 3 5 7
    __host__ __device__ static constexpr int halo(int u) { return u < 2 ? 3 : 4 - u; }        // 3 3 2 1 0
    __host__ __device__ static constexpr int out_kind(int u) { return u == 4 ? OUT_GLOBAL : OUT_BUF; }
    __host__ __device__ static constexpr int out_slots(int u) { return u == 0 ? 1 : 3; }
    __host__ __device__ static constexpr int tmem_col(int u) { return u == 0 ? 0 : (u == 4 ? 64 : 128 + 32 * (u - 1)); }
    __host__ __device__ static constexpr int epi_group(int u) { return (u == 2 || u == 0) ? 1 : 0; }
    __host__ __device__ static constexpr bool cls_bias(int) { return false; }
};
struct ProgCAT {     // fe_cat.0 .. fe_cat.2 + GAP
    static constexpr int NU = 3, INPUT = IN_TMA, IN_SLOTS = 3, IN_PLANES = 8;
    __host__ __device__ static constexpr int taps(int u) { return u == 0 ? 1 : 9; }
    __host__ __device__ static constexpr int kpl(int u) { return u == 0 ? 8 : 4; }
    __host__ __device__ static constexpr int nout(int u) { return u == 2 ? 64 : 32; }
    __host__ __device__ static constexpr int act(int u) { return u == 2 ? CSBSR_ACT_NONE : CSBSR_ACT_LEAKY; }
    __host__ __device__ static constexpr int lag(int u) { return 2 * u; }                     // 0 2 4
    __host__ __device__ static constexpr int halo(int u) { return 2 - u; }                    // 2 1 0
    __host__ __device__ static constexpr int out_kind(int u) { return u == 2 ? OUT_GAP : OUT_BUF; }
    __host__ __device__ static constexpr int out_slots(int) { return 3; }
    __host__ __device__ static constexpr int tmem_col(int u) { return u == 2 ? 0 : 64 + 32 * u; }
    __host__ __device__ static constexpr int epi_group(int u) { return u == 1 ? 1 : 0; }
    __host__ __device__ static constexpr bool cls_bias(int u) { return u == 0; }
};

// shared-memory map (bytes from the 1024-aligned base): [input ring][output rings of the OUT_BUF units][weights][barriers]
template <class P>
__host__ __device__ constexpr int off_in() { return 0; }
template <class P>
__host__ __device__ constexpr int off_out(int u) {
    int o = P::IN_SLOTS * P::IN_PLANES * kPlaneB;
    for (int i = 0; i < u; ++i)
        if (P::out_kind(i) == OUT_BUF) o += P::out_slots(i) * (P::nout(i) / 8) * kPlaneB;
    return o;
}
template <class P>
__host__ __device__ constexpr int w_bytes(int u) { return P::taps(u) * P::kpl(u) * P::nout(u) * 16; }
template <class P>
__host__ __device__ constexpr int off_w(int u) {
    int o = (off_out<P>(P::NU) + 127) / 128 * 128;
    for (int i = 0; i < u; ++i) o += w_bytes<P>(i);
    return o;
}
template <class P>
__host__ __device__ constexpr int off_bar() { return (off_w<P>(P::NU) + 127) / 128 * 128; }
template <class P>
__host__ __device__ constexpr int smem_total() { return off_bar<P>() + 2048 + 1024; }   // barriers + GAP scratch, alignment slack

struct Params {
    const float* img;            // IN_IM2COL: fp32 planar [B,3,H,W]
    const void* wpack;           // packed weights, off_w(NU) - off_w(0) bytes
    void* out;                   // OUT_GLOBAL: bf16 NHWC [B,H,W,64]
    const float* cls_bias;       // fp32 [B,5,5,64] (cls_bias units)
    float* partial;              // OUT_GAP: fp32 [B][strips][gap_blocks][64] partial sums of kGapRows-row blocks
    int gap_blocks;              // ceil(H / kGapRows)
    int B, H, W;
    int nstrips, nseg, vw;       // column strips per image, row segments per strip, valid output columns per strip
    int seg_rows;                // rows per segment (last one may be shorter)
    int nitems;
    float slope;
    int* err_flag;
};

struct Item {
    int b, y0, y1, x0, vw, strip;
};
template <class P>
__device__ __forceinline__ Item decode_item(const Params& p, int item) {
    Item it;
    const int per_img = p.nstrips * p.nseg;
    it.b = item / per_img;
    const int r = item - it.b * per_img;
    const int strip = r / p.nseg, seg = r - strip * p.nseg;
    it.y0 = seg * p.seg_rows;
    it.y1 = min(p.H, it.y0 + p.seg_rows);
    const int xs = strip * p.vw;
    it.strip = strip;
    it.vw = min(p.vw, p.W - xs);
    it.x0 = xs - P::halo(0);
    return it;
}

__device__ __forceinline__ uint32_t mod3_sub(uint32_t m3, int j) {      // (g - j) % 3 given m3 = g % 3, 0 <= j <= 3
    int v = static_cast<int>(m3) - (j % 3);
    return static_cast<uint32_t>(v < 0 ? v + 3 : v);
}

__device__ __forceinline__ float act_apply(float v, int act, float slope) {
    if (act == CSBSR_ACT_RELU) return fmaxf(v, 0.f);
    if (act == CSBSR_ACT_LEAKY) return v > 0.f ? v : v * slope;
    return v;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// ------------------------------------------------------------------ MMA issue of one unit (one elected thread)
template <class P, int U>
__device__ __forceinline__ void issue_unit(uint32_t sbase, uint32_t m3) {
    constexpr int KP = P::kpl(U), N = P::nout(U), T = P::taps(U);
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (8u << 24);
    constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);                  // SBO = 8 rows x 16 B, sm_100 descriptor, no swizzle
    constexpr uint32_t lboA = kPlaneB >> 4, lboB = (N * 16) >> 4;           // distance between the two 8-channel planes of a K=16 step
    constexpr int in_slots = U == 0 ? P::IN_SLOTS : P::out_slots(U - 1);
    constexpr int in_off = U == 0 ? off_in<P>() : off_out<P>(U - 1);
    constexpr int slot_b = KP * kPlaneB;
    constexpr int D = U == 0 ? 0 : P::lag(U) - P::lag(U - 1);               // producer step of row r+dh: g - D + dh
    uint32_t a_row[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int dh = j - 1;
        uint32_t slot = 0;
        if (in_slots == 3 && (T == 9 || dh == 0)) slot = mod3_sub(m3, D - dh);
        a_row[j] = ((sbase + in_off + slot * slot_b) >> 4) + 1u;             // +1: pixel 0 sits after the left pad
    }
    const uint32_t w16 = (sbase + off_w<P>(U)) >> 4;
    const uint32_t tmem_d = static_cast<uint32_t>(P::tmem_col(U));
    uint32_t acc = 0;
#pragma unroll
    for (int t = 0; t < T; ++t) {
        const int dh = T == 9 ? t / 3 - 1 : 0, dw = T == 9 ? t % 3 - 1 : 0;
#pragma unroll
        for (int ks = 0; ks < KP / 2; ++ks) {
            const uint32_t a_lo = ((a_row[dh + 1] + static_cast<uint32_t>(dw + ks * 2 * static_cast<int>(lboA))) & 0x3FFFu) | (lboA << 16);
            const uint32_t b_lo = ((w16 + static_cast<uint32_t>((t * KP + ks * 2) * static_cast<int>(lboB))) & 0x3FFFu) | (lboB << 16);
            umma_bf16_lohi(tmem_d, a_lo, b_lo, desc_hi, idesc, acc);
            acc = 1u;
        }
    }
}

// unit U is active at item-local step t when its row t - lag lies in the rows its consumers (or the output) need
template <class P, int U>
__device__ __forceinline__ bool unit_active(const Item& it, int t) {
    const int row = t - P::lag(U);
    return row >= it.y0 - P::halo(U) && row < it.y1 + P::halo(U);
}

struct Bars {
    uint64_t* acc_full;      // [NU]   MMA -> epilogue: accumulator of unit u complete
    uint64_t* out_ready;     // [NU]   epilogue -> MMA: unit u drained (and its output row written)
    uint64_t* in_ready;      // [IN_SLOTS]
    uint64_t* in_free;       // [IN_SLOTS]
};

template <class P, int U>
__device__ __forceinline__ void mma_step_units(const Params& p, const Bars& bar, const Item& it, int t, uint32_t sbase, uint32_t m3,
                                               bool first, uint32_t& ph_out, uint32_t in_slot, uint32_t& ph_in) {
    // wait for what unit U reads: the epilogue of unit U-1 of the previous step (or the input row of this step); the unit's
    // own accumulator was drained before that (the unit above waited on out_ready[U] / the step started with out_ready[NU-1])
    if (U == P::NU - 1 && !first) {
        mbar_wait(&bar.out_ready[P::NU - 1], (ph_out >> (P::NU - 1)) & 1u, p.err_flag, 20 + U);
        ph_out ^= 1u << (P::NU - 1);
    }
    if constexpr (U > 0) {
        if (!first) {
            mbar_wait(&bar.out_ready[U - 1], (ph_out >> (U - 1)) & 1u, p.err_flag, 30 + U);
            ph_out ^= 1u << (U - 1);
        }
    } else {
        mbar_wait(&bar.in_ready[in_slot], (ph_in >> in_slot) & 1u, p.err_flag, 40);
        ph_in ^= 1u << in_slot;
    }
    tcgen05_fence_after();
    if (unit_active<P, U>(it, t)) issue_unit<P, U>(sbase, m3);
    umma_commit(&bar.acc_full[U]);
    if (U == 0) umma_commit(&bar.in_free[in_slot]);
    if constexpr (U > 0) mma_step_units<P, U - 1>(p, bar, it, t, sbase, m3, first, ph_out, in_slot, ph_in);
}

// ------------------------------------------------------------------ epilogue of one unit (one thread = one pixel)
template <class P, int U>
__device__ __forceinline__ void epilogue_unit(const Params& p, const Item& it, int t, uint32_t sbase, uint32_t m3, int q, int lane,
                                              float (&gap)[64]) {
    constexpr int N = P::nout(U);
    const int px = q * 32 + lane;
    const int row = t - P::lag(U);
    const int x = it.x0 + px;
    const bool in_img = row >= 0 && row < p.H && x >= 0 && x < p.W;
    const uint32_t taddr = (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(P::tmem_col(U));
    const float* bias = nullptr;
    if (P::cls_bias(U)) {
        const int cy = row < 2 ? row : (p.H - 1 - row < 2 ? 4 - (p.H - 1 - row) : 2);
        const int cx = x < 2 ? x : (p.W - 1 - x < 2 ? 4 - (p.W - 1 - x) : 2);
        const int cyc = min(max(cy, 0), 4), cxc = min(max(cx, 0), 4);
        bias = p.cls_bias + (static_cast<size_t>(it.b) * 25 + cyc * 5 + cxc) * 64;
    }
    uint32_t out_addr = 0;
    if (P::out_kind(U) == OUT_BUF) {
        uint32_t slot = 0;
        if (P::out_slots(U) == 3) slot = m3;
        out_addr = sbase + off_out<P>(U) + slot * ((N / 8) * kPlaneB) + static_cast<uint32_t>(1 + px) * 16u;
    }
    const bool store_px = px >= P::halo(0) && px < P::halo(0) + it.vw;
#pragma unroll
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        float b[32];
        if (P::cls_bias(U)) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + i));
                b[i] = bb.x; b[i + 1] = bb.y; b[i + 2] = bb.z; b[i + 3] = bb.w;
            }
        }
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            float a = __uint_as_float(v[i]);
            if (P::cls_bias(U)) a += b[i];
            f[i] = act_apply(a, P::act(U), p.slope);
        }
        if (P::out_kind(U) == OUT_BUF) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t w0 = pack2(f[8 * j], f[8 * j + 1]), w1 = pack2(f[8 * j + 2], f[8 * j + 3]);
                uint32_t w2 = pack2(f[8 * j + 4], f[8 * j + 5]), w3 = pack2(f[8 * j + 6], f[8 * j + 7]);
                if (!in_img) w0 = w1 = w2 = w3 = 0u;                       // the next layer's zero padding
                sts128(out_addr + static_cast<uint32_t>((c0 / 8 + j) * kPlaneB), w0, w1, w2, w3);
            }
        } else if (P::out_kind(U) == OUT_GLOBAL) {
            if (store_px && in_img) {
                uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) +
                                                      ((static_cast<size_t>(it.b) * p.H + row) * p.W + x) * 64 + c0);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    dst[j] = make_uint4(pack2(f[8 * j], f[8 * j + 1]), pack2(f[8 * j + 2], f[8 * j + 3]),
                                        pack2(f[8 * j + 4], f[8 * j + 5]), pack2(f[8 * j + 6], f[8 * j + 7]));
            }
        } else {
            if (store_px && in_img) {
#pragma unroll
                for (int i = 0; i < 32; ++i) gap[c0 + i] += f[i];
            }
        }
    }
}

template <class P, int U>
__device__ __forceinline__ void epilogue_step_units(const Params& p, const Bars& bar, const Item& it, int t, uint32_t sbase,
                                                    uint32_t m3, int group, int q, int lane, uint32_t& ph_acc, float (&gap)[64]) {
    if (P::epi_group(U) == group) {
        mbar_wait(&bar.acc_full[U], (ph_acc >> U) & 1u, p.err_flag, 50 + U);
        ph_acc ^= 1u << U;
        tcgen05_fence_after();
        if (unit_active<P, U>(it, t)) epilogue_unit<P, U>(p, it, t, sbase, m3, q, lane, gap);
        tcgen05_fence_before();
        if (P::out_kind(U) == OUT_BUF) fence_proxy_async_smem();           // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar.out_ready[U]);
    }
    if constexpr (U > 0) epilogue_step_units<P, U - 1>(p, bar, it, t, sbase, m3, group, q, lane, ph_acc, gap);
}

// ------------------------------------------------------------------ kernel
template <class P>
__global__ void __launch_bounds__(kThreads, 1) chain_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + off_bar<P>());
    Bars bar;
    bar.acc_full = bars;
    bar.out_ready = bars + P::NU;
    bar.in_ready = bars + 2 * P::NU;
    bar.in_free = bars + 2 * P::NU + P::IN_SLOTS;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * P::NU + 2 * P::IN_SLOTS);
    float* gap_smem = reinterpret_cast<float*>(bars + 2 * P::NU + 2 * P::IN_SLOTS + 1);   // [4][64] (OUT_GAP only; 1 KB slack area)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int u = 0; u < P::NU; ++u) {
            mbar_init(&bar.acc_full[u], 1);
            mbar_init(&bar.out_ready[u], 4);                    // one arrive per warp of the draining epilogue group
        }
        for (int s = 0; s < P::IN_SLOTS; ++s) {
            mbar_init(&bar.in_ready[s], P::INPUT == IN_IM2COL ? 4 : 1);
            mbar_init(&bar.in_free[s], 1);
        }
        fence_barrier_init();
        if (P::INPUT == IN_TMA) tma_prefetch_desc(&tmIn);
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, 256);
    // weights: one flat copy of the host-packed image; activation rings start zeroed (pads and never-written rows are
    // only ever read into halo pixels, but keep them finite)
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.wpack);
        uint4* dst = reinterpret_cast<uint4*>(smem + off_w<P>(0));
        constexpr int n16 = (off_w<P>(P::NU) - off_w<P>(0)) / 16;
        for (int i = threadIdx.x; i < n16; i += kThreads) dst[i] = __ldg(src + i);
        uint4* z = reinterpret_cast<uint4*>(smem);
        constexpr int z16 = off_out<P>(P::NU) / 16;
        for (int i = threadIdx.x; i < z16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (*tmem_ptr_smem != 0u) {       // 256 columns of an otherwise empty SM start at column 0 (keeps TMEM addresses compile-time)
        if (p.err_flag) atomicExch(p.err_flag, 8);
        asm volatile("trap;");
    }

    if (warp == 0) {
        // ===================== MMA issuer =====================
        if (elect_one_sync()) {
            uint32_t g = kG0, m3 = 0, ph_out = 0, ph_in = 0, in_slot = 0;
            for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
                const Item it = decode_item<P>(p, item);
                const int tb = it.y0 - P::halo(0) + P::lag(0), te = it.y1 - 1 + P::lag(P::NU - 1);
                for (int t = tb; t <= te; ++t) {
                    mma_step_units<P, P::NU - 1>(p, bar, it, t, sbase, m3, g == kG0, ph_out, in_slot, ph_in);
                    ++g;
                    m3 = m3 == 2 ? 0 : m3 + 1;
                    if (P::IN_SLOTS == 3) in_slot = m3;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== TMA producer (IN_TMA): one 64-channel row = 8 plane boxes of [136 px][8 ch] =====================
        if (P::INPUT == IN_TMA && elect_one_sync()) {
            uint32_t slot = 0, ph_free = 0, n = 0;
            for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
                const Item it = decode_item<P>(p, item);
                const int tb = it.y0 - P::halo(0) + P::lag(0), te = it.y1 - 1 + P::lag(P::NU - 1);
                for (int t = tb; t <= te; ++t) {
                    if (n >= static_cast<uint32_t>(P::IN_SLOTS)) {
                        mbar_wait(&bar.in_free[slot], (ph_free >> slot) & 1u, p.err_flag, 60);
                        ph_free ^= 1u << slot;
                    }
                    if (unit_active<P, 0>(it, t)) {
                        mbar_arrive_expect_tx(&bar.in_ready[slot], P::IN_PLANES * kPlaneB);
                        const uint32_t dst = sbase + off_in<P>() + slot * (P::IN_PLANES * kPlaneB);
                        const int row = t - P::lag(0);
#pragma unroll
                        for (int pl = 0; pl < P::IN_PLANES; ++pl)
                            tma_load_4d(dst + pl * kPlaneB, &tmIn, &bar.in_ready[slot], pl * 8, it.x0 - 1, row, it.b);
                    } else {
                        mbar_arrive(&bar.in_ready[slot]);
                    }
                    ++n;
                    slot = slot == P::IN_SLOTS - 1 ? 0 : slot + 1;
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4 && warp < 8) {
        // ===================== im2col builder (IN_IM2COL): 3x3 x 3 channels -> K = 27 (+5 zeros), one pixel per thread =====================
        if (P::INPUT == IN_IM2COL) {
            const int px = (warp - 4) * 32 + lane;
            uint32_t ph_free = 0;
            bool first = true;
            for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
                const Item it = decode_item<P>(p, item);
                const int tb = it.y0 - P::halo(0) + P::lag(0), te = it.y1 - 1 + P::lag(P::NU - 1);
                const int x = it.x0 + px;
                const float* base = p.img + static_cast<size_t>(it.b) * 3 * p.H * p.W;
                for (int t = tb; t <= te; ++t) {
                    if (!first) {
                        mbar_wait(&bar.in_free[0], ph_free, p.err_flag, 61);
                        ph_free ^= 1u;
                    }
                    first = false;
                    if (unit_active<P, 0>(it, t)) {
                        const int row = t - P::lag(0);
                        float v[32];
#pragma unroll
                        for (int i = 27; i < 32; ++i) v[i] = 0.f;
#pragma unroll
                        for (int a = 0; a < 3; ++a) {
                            const int yy = row + a - 1;
                            const bool yok = yy >= 0 && yy < p.H;
#pragma unroll
                            for (int bb = 0; bb < 3; ++bb) {
                                const int xx = x + bb - 1;
                                const bool ok = yok && xx >= 0 && xx < p.W;
#pragma unroll
                                for (int c = 0; c < 3; ++c)
                                    v[(a * 3 + bb) * 3 + c] = ok ? __ldg(base + (static_cast<size_t>(c) * p.H + yy) * p.W + xx) : 0.f;
                            }
                        }
                        const uint32_t dst = sbase + off_in<P>() + static_cast<uint32_t>(1 + px) * 16u;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            sts128(dst + j * kPlaneB, pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]),
                                   pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
                        fence_proxy_async_smem();
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar.in_ready[0]);
                }
            }
        }
    } else if (warp >= 8) {
        // ===================== epilogue: two groups of 4 warps (one per TMEM lane quarter) =====================
        const int group = (warp - 8) >> 2, q = warp & 3;
        uint32_t m3 = 0, ph_acc = 0;
        float gap[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) gap[i] = 0.f;
        for (int item = blockIdx.x; item < p.nitems; item += gridDim.x) {
            const Item it = decode_item<P>(p, item);
            const int tb = it.y0 - P::halo(0) + P::lag(0), te = it.y1 - 1 + P::lag(P::NU - 1);
            for (int t = tb; t <= te; ++t) {
                epilogue_step_units<P, P::NU - 1>(p, bar, it, t, sbase, m3, group, q, lane, ph_acc, gap);
                m3 = m3 == 2 ? 0 : m3 + 1;
                if (P::out_kind(P::NU - 1) == OUT_GAP && P::epi_group(P::NU - 1) == group) {
                    // The pooled sums leave in FIXED blocks of kGapRows image rows (segments start on block boundaries), so the
                    // order of the additions -- rows of a block per thread, lanes (shuffle tree), the 4 quarter warps, then the
                    // blocks of the image in gap_finalize_kernel -- does not depend on how the launch split the image into
                    // segments, i.e. not on how many images share the launch.
                    const int row = t - P::lag(P::NU - 1);             // the row the pooling unit handled in this step
                    if (row >= it.y0 && row < it.y1 && (((row + 1) % kGapRows) == 0 || row == it.y1 - 1)) {
#pragma unroll
                        for (int i = 0; i < 64; ++i) {
                            float s = gap[i];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                            if (lane == 0) gap_smem[q * 64 + i] = s;
                            gap[i] = 0.f;
                        }
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        if (q == 0) {
                            const size_t slot = (static_cast<size_t>(it.b) * p.nstrips + it.strip) * p.gap_blocks + row / kGapRows;
                            for (int i = lane; i < 64; i += 32)
                                p.partial[slot * 64 + i] = ((gap_smem[i] + gap_smem[64 + i]) + gap_smem[128 + i]) + gap_smem[192 + i];
                        }
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                    }
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(0u, 256);
    }
}

// GAP finalisation: out[b][c] = (sum of the image's item partials, fixed order) / (H*W)
__global__ void gap_finalize_kernel(const float* __restrict__ partial, float* __restrict__ out, int items_per_img, float inv_hw,
                                    int gap_c) {
    const int b = blockIdx.x, c = threadIdx.x;
    if (c >= gap_c) return;
    float s = 0.f;
    for (int i = 0; i < items_per_img; ++i) s += partial[(static_cast<size_t>(b) * items_per_img + i) * 64 + c];
    out[b * gap_c + c] = s * inv_hw;
}

static int* g_err_flag = nullptr;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_fn() {
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

template <class P>
static void plan(Params& p, int B, int H, int W) {
    p.B = B; p.H = H; p.W = W;
    const int maxvw = kPX - 2 * P::halo(0);
    p.nstrips = (W + maxvw - 1) / maxvw;
    p.vw = (W + p.nstrips - 1) / p.nstrips;
    // row segments: fill the SMs once; a segment re-computes 2*halo rows and pays lag(NU-1) pipeline-fill steps, so keep
    // them at least 32 rows tall
    int nseg = num_sms() / (B * p.nstrips);
    if (nseg < 1) nseg = 1;
    while (nseg > 1 && (H + nseg - 1) / nseg < 32) --nseg;
    p.seg_rows = ((H + nseg - 1) / nseg + kGapRows - 1) / kGapRows * kGapRows;    // segments start on pooling-block boundaries
    p.nseg = (H + p.seg_rows - 1) / p.seg_rows;
    p.gap_blocks = (H + kGapRows - 1) / kGapRows;
    p.nitems = B * p.nstrips * p.nseg;
}

template <class P>
static int launch(Params& p, const CUtensorMap& tm, cudaStream_t stream) {
    if (!g_err_flag) {
        CSBSR_CHECK_CUDA(cudaMalloc(&g_err_flag, sizeof(int)));
        CSBSR_CHECK_CUDA(cudaMemset(g_err_flag, 0, sizeof(int)));
    }
    p.err_flag = g_err_flag;
    static bool attr_set = false;
    if (!attr_set) {
        CSBSR_CHECK_CUDA(cudaFuncSetAttribute(chain_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total<P>()));
        attr_set = true;
    }
    const int grid = p.nitems < num_sms() ? p.nitems : num_sms();
    chain_kernel<P><<<grid, kThreads, smem_total<P>(), stream>>>(tm, p);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace kp
}  // namespace csbsr

using namespace csbsr;

extern "C" size_t csbsr_kpred_wpack_bytes(int which) {
    return which == 0 ? static_cast<size_t>(kp::off_w<kp::ProgSR>(kp::ProgSR::NU) - kp::off_w<kp::ProgSR>(0))
                      : static_cast<size_t>(kp::off_w<kp::ProgCAT>(kp::ProgCAT::NU) - kp::off_w<kp::ProgCAT>(0));
}

extern "C" size_t csbsr_kpred_workspace_bytes(int b, int h, int w) {
    kp::Params p;
    memset(&p, 0, sizeof(p));
    kp::plan<kp::ProgCAT>(p, b, h, w);
    return static_cast<size_t>(b) * p.nstrips * p.gap_blocks * 64 * sizeof(float);
}

extern "C" int csbsr_kpred_sr_chain(const float* img, const void* wpack, void* out, int b, int h, int w, float slope,
                                    void* stream) {
    CSBSR_REQUIRE(img && wpack && out && b > 0 && h > 0 && w > 0, "kpred_sr_chain: bad arguments");
    CSBSR_REQUIRE((reinterpret_cast<uintptr_t>(wpack) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                  "kpred_sr_chain: wpack / out must be 16-byte aligned");
    kp::Params p;
    memset(&p, 0, sizeof(p));
    kp::plan<kp::ProgSR>(p, b, h, w);
    p.img = img; p.wpack = wpack; p.out = out; p.slope = slope;
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    return kp::launch<kp::ProgSR>(p, tm, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int csbsr_kpred_cat_chain(const void* in, const void* wpack, const float* cls_bias, float* gap_out, int gap_c,
                                     void* ws, size_t ws_bytes, int b, int h, int w, float slope, void* stream) {
    CSBSR_REQUIRE(in && wpack && cls_bias && gap_out && ws && b > 0 && h > 0 && w > 0 && gap_c >= 1 && gap_c <= 64,
                  "kpred_cat_chain: bad arguments");
    CSBSR_REQUIRE((reinterpret_cast<uintptr_t>(wpack) & 15) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(cls_bias) & 15) == 0,
                  "kpred_cat_chain: in / wpack / cls_bias must be 16-byte aligned");
    cudaStream_t stream_ = reinterpret_cast<cudaStream_t>(stream);
    kp::Params p;
    memset(&p, 0, sizeof(p));
    kp::plan<kp::ProgCAT>(p, b, h, w);
    CSBSR_REQUIRE(ws_bytes >= static_cast<size_t>(b) * p.nstrips * p.gap_blocks * 64 * sizeof(float), "kpred_cat_chain: workspace too small");
    p.wpack = wpack; p.cls_bias = cls_bias; p.partial = reinterpret_cast<float*>(ws); p.slope = slope;
    kp::PFN_encodeTiled encode = kp::encode_fn();
    CSBSR_REQUIRE(encode, "kpred_cat_chain: cuTensorMapEncodeTiled entry point unavailable");
    CUtensorMap tm;
    {
        // [B,H,W,64] bf16; one box = 8 channels x 136 pixels of one row (one plane of a row slot); the pixels outside the
        // image (and whole rows outside it) are zero-filled by TMA = the zero padding of the 3x3 layers
        cuuint64_t dims[4] = {64, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)b};
        cuuint64_t strides[3] = {128, (cuuint64_t)w * 128, (cuuint64_t)w * 128 * h};
        cuuint32_t box[4] = {8, (cuuint32_t)kp::kPlanePx, 1, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        CSBSR_REQUIRE(r == CUDA_SUCCESS, "kpred_cat_chain: cuTensorMapEncodeTiled failed with %d", (int)r);
    }
    int rc = kp::launch<kp::ProgCAT>(p, tm, stream_);
    if (rc) return rc;
    kp::gap_finalize_kernel<<<b, 64, 0, stream_>>>(p.partial, gap_out, p.nstrips * p.gap_blocks, 1.0f / (static_cast<float>(h) * w), gap_c);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
