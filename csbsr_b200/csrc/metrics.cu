// AIU (IoU swept over 99 thresholds) and the Hausdorff / mean-surface-distance sweep, bit-exact with the
// reference's numpy/scipy path:
//   thresholds + binarisation   model/engine/inference.py:49-53,111
//   IoU counts                  model/utils/estimate_metrics.py:72-84
//   calc_distance_metrics       model/engine/inference.py:293-336
//   compute_surface_distances   model/utils/metrics/surface_distance/metrics/surface_distance.py:136-288
//   robust Hausdorff / ASD      same file :291-359 ; contour-length table lookup_tables.py:330-400
//
// Everything up to the final quantile is integer work on the (H+1)x(W+1) grid of pixel corners:
//   q[y,x]   = #{i : p[y,x] > t_i}  (uint8): the prediction at threshold i is (q >= i), so a corner is a
//              prediction-border corner for exactly the thresholds in (qlo, qhi] of its 2x2 pixel block;
//   exact squared Euclidean distance to the nearest border corner = column scan + exact row search;
//   keys (d2 << 2 | length-code) are sorted per (image, threshold, direction) with a shared-memory
//   bitonic sort, then ONE thread replays numpy's floating-point order of operations (sequential
//   cumsum, pairwise sum with 8 accumulators / blocks of 128, fp64 sqrt) so the percentile index and
//   the value are identical to the reference's, bit for bit.
#include <math_constants.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/csbsr_b200.h"

namespace csbsr {

static constexpr int kNumThr = 99;
static constexpr unsigned short kInf16 = 0xFFFFu;
static constexpr unsigned int kInfD2 = 0x3FFFFFFFu;       // "no border anywhere" marker (fits the key's 30 bits)

// ------------------------------------------------------------------ quantise + AIU histogram
// thr[i] = float32(i * 0.01) for i = 1..99, exactly what torch.Tensor([i*0.01 ...]) holds.
__global__ void quantize_hist_kernel(const float* __restrict__ prob, const float* __restrict__ mask,
                                     const float* __restrict__ thr, unsigned char* __restrict__ q,
                                     unsigned char* __restrict__ gt, int* __restrict__ hist, int HW) {
    // grid: (slices, B).  hist[b][2][100]: [0] = target false, [1] = target true (mask > 0.5)
    __shared__ int sh[2][100];
    __shared__ float sthr[kNumThr];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < 200; i += blockDim.x) (&sh[0][0])[i] = 0;
    for (int i = threadIdx.x; i < kNumThr; i += blockDim.x) sthr[i] = thr[i];
    __syncthreads();
    const float* pp = prob + static_cast<size_t>(b) * HW;
    const float* mp = mask + static_cast<size_t>(b) * HW;
    auto one = [&](float p, float m, unsigned int& qv, unsigned int& gv) {
        // count thresholds strictly below p: thresholds are increasing -> binary search for the first t >= p
        int lo = 0, hi = kNumThr;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (p > sthr[mid]) lo = mid + 1; else hi = mid;
        }
        const int t_iou = m > 0.5f ? 1 : 0;              // IoU binarisation (estimate_metrics.py:79-80)
        const int t_hd = m != 0.f ? 1 : 0;               // astype(bool) (inference.py:305)
        qv = static_cast<unsigned int>(lo);
        gv = static_cast<unsigned int>(t_iou | (t_hd << 1));
        atomicAdd(&sh[t_iou][lo], 1);
    };
    if ((HW & 3) == 0 && ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(mp)) & 15) == 0) {
        // 128-bit loads of four probabilities / mask values, one 32-bit store of the four quantised bytes
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW / 4; i += gridDim.x * blockDim.x) {
            const float4 p4 = reinterpret_cast<const float4*>(pp)[i];
            const float4 m4 = reinterpret_cast<const float4*>(mp)[i];
            unsigned int qv[4], gv[4];
            one(p4.x, m4.x, qv[0], gv[0]); one(p4.y, m4.y, qv[1], gv[1]); one(p4.z, m4.z, qv[2], gv[2]); one(p4.w, m4.w, qv[3], gv[3]);
            reinterpret_cast<unsigned int*>(q + static_cast<size_t>(b) * HW)[i] = qv[0] | (qv[1] << 8) | (qv[2] << 16) | (qv[3] << 24);
            reinterpret_cast<unsigned int*>(gt + static_cast<size_t>(b) * HW)[i] = gv[0] | (gv[1] << 8) | (gv[2] << 16) | (gv[3] << 24);
        }
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
            unsigned int qv, gv;
            one(pp[i], mp[i], qv, gv);
            q[static_cast<size_t>(b) * HW + i] = static_cast<unsigned char>(qv);
            gt[static_cast<size_t>(b) * HW + i] = static_cast<unsigned char>(gv);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 200; i += blockDim.x) {
        const int v = (&sh[0][0])[i];
        if (v) atomicAdd(&hist[b * 200 + i], v);
    }
}

// I[b][i-1] = #(pred_i & tgt), U[b][i-1] = #(pred_i | tgt), pred_i = (q >= i)
__global__ void aiu_counts_kernel(const int* __restrict__ hist, long long* __restrict__ inter,
                                  long long* __restrict__ uni) {
    const int b = blockIdx.x;
    const int i = threadIdx.x + 1;
    if (i > kNumThr) return;
    const int* h0 = hist + b * 200;
    const int* h1 = h0 + 100;
    long long I = 0, P = 0, T = 0;
    for (int k = 0; k < 100; ++k) {
        T += h1[k];
        if (k >= i) {
            I += h1[k];
            P += h0[k] + h1[k];
        }
    }
    inter[b * kNumThr + i - 1] = I;
    uni[b * kNumThr + i - 1] = P + T - I;
}

// ------------------------------------------------------------------ corner grid
__device__ __forceinline__ int pix(const unsigned char* a, int y, int x, int H, int W) {
    return (y >= 0 && y < H && x >= 0 && x < W) ? a[y * W + x] : 0;
}

// per corner: qlo/qhi of the 2x2 block (out-of-image pixels count as q = 0) and the gt neighbour code
// code = 8*in[y-1,x-1] + 4*in[y-1,x] + 2*in[y,x-1] + in[y,x]   (scipy correlate with [[8,4],[2,1]], zero padded)
__global__ void corner_kernel(const unsigned char* __restrict__ q, const unsigned char* __restrict__ gt,
                              unsigned char* __restrict__ qlo, unsigned char* __restrict__ qhi,
                              unsigned char* __restrict__ cg, int H, int W) {
    const int Hc = H + 1, Wc = W + 1;
    const int b = blockIdx.y;
    const unsigned char* qb = q + static_cast<size_t>(b) * H * W;
    const unsigned char* gb = gt + static_cast<size_t>(b) * H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Hc * Wc; i += gridDim.x * blockDim.x) {
        const int y = i / Wc, x = i % Wc;
        const int a = pix(qb, y - 1, x - 1, H, W), bb = pix(qb, y - 1, x, H, W), c = pix(qb, y, x - 1, H, W),
                  d = pix(qb, y, x, H, W);
        const size_t o = static_cast<size_t>(b) * Hc * Wc + i;
        if (qlo) {                                   // only the unfused (large-image) path keeps the per-corner range
            qlo[o] = static_cast<unsigned char>(min(min(a, bb), min(c, d)));
            qhi[o] = static_cast<unsigned char>(max(max(a, bb), max(c, d)));
        }
        const int ga = (pix(gb, y - 1, x - 1, H, W) >> 1) & 1, gbb = (pix(gb, y - 1, x, H, W) >> 1) & 1,
                  gc = (pix(gb, y, x - 1, H, W) >> 1) & 1, gd = (pix(gb, y, x, H, W) >> 1) & 1;
        cg[o] = static_cast<unsigned char>(8 * ga + 4 * gbb + 2 * gc + gd);
    }
}

// column pass of the exact EDT: g[y][x] = vertical distance to the nearest border corner in column x.
// MODE 0: gt borders (code not in {0,15}); MODE 1: prediction borders at threshold t = blockIdx.y % 99 + 1.
template <int MODE>
__global__ void column_scan_kernel(const unsigned char* __restrict__ cg, const unsigned char* __restrict__ qlo,
                                   const unsigned char* __restrict__ qhi, unsigned short* __restrict__ g, int Hc,
                                   int Wc) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= Wc) return;
    int b, t = 0;
    if (MODE == 0) {
        b = blockIdx.y;
    } else {
        b = blockIdx.y / kNumThr;
        t = blockIdx.y % kNumThr + 1;
    }
    const size_t cbase = static_cast<size_t>(b) * Hc * Wc;
    unsigned short* gp = g + static_cast<size_t>(blockIdx.y) * Hc * Wc;
    int last = -1;
    for (int y = 0; y < Hc; ++y) {
        const size_t o = cbase + static_cast<size_t>(y) * Wc + x;
        bool border;
        if (MODE == 0) {
            const int c = cg[o];
            border = (c != 0 && c != 15);
        } else {
            border = (qlo[o] < t) && (t <= qhi[o]);
        }
        if (border) last = y;
        gp[static_cast<size_t>(y) * Wc + x] = last < 0 ? kInf16 : static_cast<unsigned short>(y - last);
    }
    last = -1;
    for (int y = Hc - 1; y >= 0; --y) {
        const size_t go = static_cast<size_t>(y) * Wc + x;
        const unsigned short cur = gp[go];
        if (cur == 0) last = y;
        else if (last >= 0) {
            const int dd = last - y;
            if (cur == kInf16 || dd < cur) gp[go] = static_cast<unsigned short>(dd);
        }
    }
}

// exact row search: min over x' of (x-x')^2 + g[y][x']^2, expanding from x with early exit
__device__ __forceinline__ unsigned int row_search(const unsigned short* __restrict__ grow, int x, int Wc) {
    unsigned int best = kInfD2;
    const int maxd = max(x, Wc - 1 - x);
    for (int d = 0; d <= maxd; ++d) {
        const unsigned int dd = static_cast<unsigned int>(d) * d;
        if (dd >= best) break;
        if (x - d >= 0) {
            const unsigned int gv = grow[x - d];
            if (gv != kInf16) best = min(best, dd + gv * gv);
        }
        if (d > 0 && x + d < Wc) {
            const unsigned int gv = grow[x + d];
            if (gv != kInf16) best = min(best, dd + gv * gv);
        }
    }
    return best;
}

// exact row search with block skipping: gmin[k] = min of grow[16k .. 16k+15] (kInf16 when the block has no finite entry); a block
// whose lower bound dxmin^2 + gmin^2 cannot beat `best` is skipped as a whole -- long searches over uniform far-away borders (the
// image frame when the prediction is "everything") cost Wc/16 bound checks instead of Wc steps.  Same minimum as row_search.
__device__ __forceinline__ unsigned int row_search_blocked(const unsigned short* __restrict__ grow, const unsigned short* __restrict__ gmin,
                                                           int x, int Wc) {
    const int nb = (Wc + 15) >> 4, b0 = x >> 4;
    unsigned int best = kInfD2;
    auto scan = [&](int blk) {
        const int lo = blk << 4, hi = min(Wc, lo + 16);
        for (int xx = lo; xx < hi; ++xx) {
            const unsigned int gv = grow[xx];
            if (gv != kInf16) {
                const unsigned int dx = static_cast<unsigned int>(abs(xx - x));
                best = min(best, dx * dx + gv * gv);
            }
        }
    };
    scan(b0);
    for (int k = 1; k < nb; ++k) {
        const int bl = b0 - k, br = b0 + k;
        if (bl < 0 && br >= nb) break;
        bool any = false;
        if (bl >= 0) {
            const unsigned int dxm = static_cast<unsigned int>(x - ((bl << 4) + 15));      // distance to the block's closest column
            if (dxm * dxm < best) {
                any = true;
                const unsigned int gm = gmin[bl];
                if (gm != kInf16 && dxm * dxm + gm * gm < best) scan(bl);
            }
        }
        if (br < nb) {
            const unsigned int dxm = static_cast<unsigned int>((br << 4) - x);
            if (dxm * dxm < best) {
                any = true;
                const unsigned int gm = gmin[br];
                if (gm != kInf16 && dxm * dxm + gm * gm < best) scan(br);
            }
        }
        if (!any && (bl < 0 || br >= nb || true)) {
            // both sides are already too far in x alone (or off the row): farther blocks are farther still
            const bool left_done = bl < 0 || static_cast<unsigned int>(x - ((bl << 4) + 15)) * static_cast<unsigned int>(x - ((bl << 4) + 15)) >= best;
            const bool right_done = br >= nb || static_cast<unsigned int>((br << 4) - x) * static_cast<unsigned int>((br << 4) - x) >= best;
            if (left_done && right_done) break;
        }
    }
    return best;
}

// full squared-distance map to the gt borders (needed at every prediction-border corner): one block per (image, corner row),
// the row of column distances and its 16-column block minima staged in shared memory, block-skipping exact row search
__global__ void __launch_bounds__(256) gt_edt_kernel(const unsigned short* __restrict__ gcol, unsigned int* __restrict__ d2g, int Hc,
                                                     int Wc) {
    extern __shared__ unsigned short ge_sm[];                 // [Wc] distances + [ceil(Wc/16)] block minima
    unsigned short* grow = ge_sm;
    unsigned short* gmin = ge_sm + ((Wc + 7) & ~7);
    const int y = blockIdx.x, b = blockIdx.y;
    const unsigned short* src = gcol + (static_cast<size_t>(b) * Hc + y) * Wc;
    for (int x = threadIdx.x; x < Wc; x += blockDim.x) grow[x] = src[x];
    __syncthreads();
    const int nblk = (Wc + 15) >> 4;
    for (int k = threadIdx.x; k < nblk; k += blockDim.x) {
        unsigned int mn = kInf16;
        for (int j = 0; j < min(16, Wc - (k << 4)); ++j) mn = min(mn, static_cast<unsigned int>(grow[(k << 4) + j]));
        gmin[k] = static_cast<unsigned short>(mn);
    }
    __syncthreads();
    unsigned int* ob = d2g + (static_cast<size_t>(b) * Hc + y) * Wc;
    for (int x = threadIdx.x; x < Wc; x += blockDim.x) ob[x] = row_search_blocked(grow, gmin, x, Wc);
}

// neighbour code -> contour-length class: 0 = none, 1 = diag (0.5*sqrt2), 2 = 1.0, 3 = 2*diag (lookup_tables.py:351-398)
__device__ __forceinline__ int len_class(int code) {
    //            0  1  2  3  4  5  6  7  8  9 10 11 12 13 14 15
    const unsigned int tbl = (0u) | (1u << 2) | (1u << 4) | (2u << 6) | (1u << 8) | (2u << 10) | (3u << 12) | (1u << 14) |
                             (1u << 16) | (3u << 18) | (2u << 20) | (1u << 22) | (2u << 24) | (1u << 26) | (1u << 28) |
                             (0u << 30);
    return (tbl >> (2 * code)) & 3;
}

// compact list of gt border corners per image (order is irrelevant: keys are sorted later)
__global__ void gt_list_kernel(const unsigned char* __restrict__ cg, int* __restrict__ list, int* __restrict__ count,
                               int NC) {
    const int b = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NC; i += gridDim.x * blockDim.x) {
        const int c = cg[static_cast<size_t>(b) * NC + i];
        if (c != 0 && c != 15) {
            const int slot = atomicAdd(&count[b], 1);
            list[static_cast<size_t>(b) * NC + slot] = i;
        }
    }
}

// keys of distances_gt_to_pred: for every gt border corner and threshold, d2 to the nearest prediction border
__global__ void g2p_keys_kernel(const int* __restrict__ list, const int* __restrict__ count,
                                const unsigned char* __restrict__ cg, const unsigned short* __restrict__ gcol_t,
                                unsigned int* __restrict__ keys, int Hc, int Wc, int cap) {
    const int bt = blockIdx.y;                 // b * 99 + (t-1)
    const int b = bt / kNumThr;
    const int NC = Hc * Wc;
    const int n = count[b];
    const unsigned short* g = gcol_t + static_cast<size_t>(bt) * NC;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int i = list[static_cast<size_t>(b) * NC + k];
        const int y = i / Wc, x = i % Wc;
        const unsigned int d2 = row_search(g + static_cast<size_t>(y) * Wc, x, Wc);
        keys[static_cast<size_t>(bt) * cap + k] = (d2 << 2) | len_class(cg[static_cast<size_t>(b) * NC + i]);
    }
}

// keys of distances_pred_to_gt: every corner contributes to the thresholds in (qlo, qhi]
__global__ void p2g_keys_kernel(const unsigned char* __restrict__ q, const unsigned char* __restrict__ qlo,
                                const unsigned char* __restrict__ qhi, const unsigned int* __restrict__ d2g,
                                unsigned int* __restrict__ keys, int* __restrict__ count, int H, int W, int cap) {
    const int Hc = H + 1, Wc = W + 1, NC = Hc * Wc;
    const int b = blockIdx.y;
    const unsigned char* qb = q + static_cast<size_t>(b) * H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NC; i += gridDim.x * blockDim.x) {
        const size_t o = static_cast<size_t>(b) * NC + i;
        const int lo = qlo[o], hi = qhi[o];
        if (lo == hi) continue;
        const int y = i / Wc, x = i % Wc;
        const int a = pix(qb, y - 1, x - 1, H, W), bb = pix(qb, y - 1, x, H, W), c = pix(qb, y, x - 1, H, W),
                  d = pix(qb, y, x, H, W);
        const unsigned int dd = d2g[o];
        for (int t = lo + 1; t <= hi; ++t) {
            const int code = ((a >= t) << 3) | ((bb >= t) << 2) | ((c >= t) << 1) | (d >= t);
            const int bt = b * kNumThr + t - 1;
            const int slot = atomicAdd(&count[bt], 1);
            keys[static_cast<size_t>(bt) * cap + slot] = (dd << 2) | len_class(code);
        }
    }
}

// ------------------------------------------------------------------ numpy-order floating point replay
__device__ __forceinline__ double len_value(int cls) {
    // 0.5 * math.sqrt(1**2 + 1**2), 1, 2 * diag  (lookup_tables.py:351-398 with spacing (1, 1))
    const double diag = 0.5 * 1.4142135623730951;
    return cls == 1 ? diag : (cls == 2 ? 1.0 : (cls == 3 ? 2.0 * diag : 0.0));
}
__device__ __forceinline__ double key_dist(unsigned int key) {
    const unsigned int d2 = key >> 2;
    return d2 == kInfD2 ? CUDART_INF : sqrt(static_cast<double>(d2));    // IEEE sqrt == np.sqrt on exact integers
}
template <bool PROD>
__device__ __forceinline__ double elem(const unsigned int* keys, int i) {
    const unsigned int k = keys[i];
    const double l = len_value(k & 3);
    return PROD ? __dmul_rn(key_dist(k), l) : l;     // separate multiply: no FMA contraction with the following add
}
// numpy pairwise_sum for n <= 128 (numpy/_core/src/umath/loops_utils.h.src): 8 interleaved accumulators
template <bool PROD>
__device__ double pairwise_leaf(const unsigned int* keys, int start, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res += elem<PROD>(keys, start + i);
        return res;
    }
    double r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = elem<PROD>(keys, start + k);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] += elem<PROD>(keys, start + i + k);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += elem<PROD>(keys, start + i);
    return res;
}
// recursion of pairwise_sum for n > 128: n2 = n/2 rounded down to a multiple of 8
template <bool PROD>
__device__ double pairwise_rec(const unsigned int* keys, int start, int n) {
    // explicit stack (depth <= 20) to avoid device recursion
    int s_start[24], s_n[24], s_state[24];
    double s_left[24];
    int sp = 0;
    s_start[0] = start; s_n[0] = n; s_state[0] = 0;
    double ret = 0.0;
    while (sp >= 0) {
        const int cs = s_start[sp], cn = s_n[sp];
        if (cn <= 128) {
            ret = pairwise_leaf<PROD>(keys, cs, cn);
            --sp;
            continue;
        }
        int n2 = cn / 2;
        n2 -= n2 % 8;
        if (s_state[sp] == 0) {
            s_state[sp] = 1;
            ++sp;
            s_start[sp] = cs; s_n[sp] = n2; s_state[sp] = 0;
        } else if (s_state[sp] == 1) {
            s_left[sp] = ret;
            s_state[sp] = 2;
            ++sp;
            s_start[sp] = cs + n2; s_n[sp] = cn - n2; s_state[sp] = 0;
        } else {
            ret = s_left[sp] + ret;
            --sp;
        }
    }
    return ret;
}

// After sorting: out[0] = percentile distance (compute_robust_hausdorff :344-347), out[1] = average distance (:315-319)
__device__ void replay_list(const unsigned int* keys, int n, double pct, double* out) {
    const double total = pairwise_rec<false>(keys, 0, n);
    const double wsum = pairwise_rec<true>(keys, 0, n);
    double c = 0.0;
    int idx = n;
    for (int i = 0; i < n; ++i) {                      // np.cumsum: sequential; np.searchsorted(side='left')
        c += len_value(keys[i] & 3);
        if (c / total >= pct) { idx = i; break; }
    }
    if (idx > n - 1) idx = n - 1;
    out[0] = key_dist(keys[idx]);
    out[1] = wsum / total;
}

// Block-parallel replay with the same results as replay_list:
//  * the two pairwise sums are evaluated leaf by leaf (<= 128 elements each, one thread per leaf, numpy's own
//    8-accumulator order inside a leaf) and combined by one thread in numpy's recursion order -> bit-identical;
//  * the percentile index is located with exact integer prefix counts.  The sequential fp64 cumsum c_i differs from
//    the exactly rounded value e_i = n1*diag + n2 + n3*2diag by at most n*2^-53 relative (< 3e-11 for n <= 2^18), so
//    whenever e_i/total is further than 1e-9 from the percentile on both sides of the crossing the float predicate
//    `c_i/total >= pct` is decided; otherwise (a near tie, e.g. all-equal surfels and an even count) one thread falls
//    back to the literal sequential replay.
struct ReplayAux {
    int* leaf_start;     // [max_leaves]
    int* leaf_n;         // [max_leaves]
    double* leaf_l;      // [max_leaves] pairwise sums of lengths
    double* leaf_w;      // [max_leaves] pairwise sums of dist*length
    int* cnt;            // [blockDim][3] class counts per thread chunk
    int* misc;           // [4]: n_leaves, found index, flags
    double* dres;        // [2]: total, wsum
};

__device__ void replay_parallel(const unsigned int* keys, int n, double pct, double* out, ReplayAux ax) {
    const int tid = threadIdx.x, nt = blockDim.x;
    // ---- leaves of numpy's pairwise recursion, in left-to-right order (thread 0)
    if (tid == 0) {
        int st_s[24], st_n[24], sp = 0, nl = 0;
        st_s[0] = 0; st_n[0] = n;
        while (sp >= 0) {
            const int cs = st_s[sp], cn = st_n[sp];
            --sp;
            if (cn <= 128) {
                ax.leaf_start[nl] = cs; ax.leaf_n[nl] = cn; ++nl;
            } else {
                int n2 = cn / 2;
                n2 -= n2 % 8;
                ++sp; st_s[sp] = cs + n2; st_n[sp] = cn - n2;      // right pushed first so the left is popped first
                ++sp; st_s[sp] = cs; st_n[sp] = n2;
            }
        }
        ax.misc[0] = nl;
        ax.misc[1] = 0x7FFFFFFF;
    }
    __syncthreads();
    const int nl = ax.misc[0];
    // every leaf sum (two per leaf: lengths, distance * length) is evaluated by 8 lanes, one per numpy accumulator, and
    // combined in numpy's order ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) followed by the sequential tail
    {
        const int lane = tid & 31, sub = lane & 7, grp = lane >> 3;          // 4 groups of 8 lanes per warp
        const int warp = tid >> 5, nwarps = nt >> 5;
        for (int base = warp * 2; base < nl; base += nwarps * 2) {
            const int li = base + (grp >> 1);
            const bool prod = grp & 1;
            double r = 0.0;
            int ls = 0, ln = 0;
            if (li < nl) {
                ls = ax.leaf_start[li]; ln = ax.leaf_n[li];
                if (ln >= 8) {
                    r = prod ? elem<true>(keys, ls + sub) : elem<false>(keys, ls + sub);
                    for (int i = 8 + sub; i < ln - (ln % 8); i += 8) r += prod ? elem<true>(keys, ls + i) : elem<false>(keys, ls + i);
                }
            }
            const unsigned int full = 0xFFFFFFFFu;
            const double r1 = r + __shfl_down_sync(full, r, 1);                 // lanes 0,2,4,6 of the group: r0+r1, r2+r3, ...
            const double r2 = r1 + __shfl_down_sync(full, r1, 2);               // lanes 0,4: (r0+r1)+(r2+r3), (r4+r5)+(r6+r7)
            double res = r2 + __shfl_down_sync(full, r2, 4);                    // lane 0 of the group
            if (li < nl && sub == 0) {
                if (ln < 8) {
                    res = 0.0;
                    for (int i = 0; i < ln; ++i) res += prod ? elem<true>(keys, ls + i) : elem<false>(keys, ls + i);
                } else {
                    for (int i = ln - (ln % 8); i < ln; ++i) res += prod ? elem<true>(keys, ls + i) : elem<false>(keys, ls + i);
                }
                (prod ? ax.leaf_w : ax.leaf_l)[li] = res;
            }
        }
    }
    // ---- class counts per contiguous chunk (odd chunk length: consecutive threads then start in different shared-memory banks)
    const int L = ((n + nt - 1) / nt) | 1;
    const int c0 = min(n, tid * L), c1 = min(n, c0 + L);
    int k1 = 0, k2 = 0, k3 = 0;
    for (int i = c0; i < c1; ++i) {
        const int c = keys[i] & 3;
        k1 += (c == 1); k2 += (c == 2); k3 += (c == 3);
    }
    // exclusive prefix of the chunk counts over the block: warp scans + a scan of the warp totals (ax.cnt[0 .. 3*32))
    {
        const int lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
        int i1 = k1, i2 = k2, i3 = k3;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a1 = __shfl_up_sync(0xFFFFFFFFu, i1, o), a2 = __shfl_up_sync(0xFFFFFFFFu, i2, o), a3 = __shfl_up_sync(0xFFFFFFFFu, i3, o);
            if (lane >= o) { i1 += a1; i2 += a2; i3 += a3; }
        }
        __syncthreads();                                   // leaf tables complete; ax.cnt free
        if (lane == 31) { ax.cnt[warp * 3] = i1; ax.cnt[warp * 3 + 1] = i2; ax.cnt[warp * 3 + 2] = i3; }
        __syncthreads();
        if (tid == 0) {
            int a1 = 0, a2 = 0, a3 = 0;
            for (int w = 0; w < nwarps; ++w) {
                const int b1 = ax.cnt[w * 3], b2 = ax.cnt[w * 3 + 1], b3 = ax.cnt[w * 3 + 2];
                ax.cnt[w * 3] = a1; ax.cnt[w * 3 + 1] = a2; ax.cnt[w * 3 + 2] = a3;
                a1 += b1; a2 += b2; a3 += b3;
            }
        }
        __syncthreads();
        const int e1 = ax.cnt[warp * 3] + i1 - k1, e2 = ax.cnt[warp * 3 + 1] + i2 - k2, e3 = ax.cnt[warp * 3 + 2] + i3 - k3;
        __syncthreads();
        ax.cnt[96 + tid * 3] = e1; ax.cnt[96 + tid * 3 + 1] = e2; ax.cnt[96 + tid * 3 + 2] = e3;     // exclusive prefix of chunk tid
    }
    if (tid == 0) {
        // combine the leaf sums in recursion order: post-order over the same tree
        for (int pass = 0; pass < 2; ++pass) {
            const double* leaf = pass == 0 ? ax.leaf_l : ax.leaf_w;
            int st_n[24], st_state[24], sp = 0, li = 0;
            double st_left[24], ret = 0.0;
            st_n[0] = n; st_state[0] = 0;
            while (sp >= 0) {
                const int cn = st_n[sp];
                if (cn <= 128) { ret = leaf[li++]; --sp; continue; }
                int n2 = cn / 2;
                n2 -= n2 % 8;
                if (st_state[sp] == 0) { st_state[sp] = 1; ++sp; st_n[sp] = n2; st_state[sp] = 0; }
                else if (st_state[sp] == 1) { st_left[sp] = ret; st_state[sp] = 2; ++sp; st_n[sp] = cn - n2; st_state[sp] = 0; }
                else { ret = st_left[sp] + ret; --sp; }
            }
            ax.dres[pass] = ret;
        }
    }
    __syncthreads();
    const double total = ax.dres[0];
    const double diag = 0.5 * 1.4142135623730951;
    const double margin = 1e-9;
    k1 = ax.cnt[96 + tid * 3]; k2 = ax.cnt[96 + tid * 3 + 1]; k3 = ax.cnt[96 + tid * 3 + 2];
    for (int i = c0; i < c1; ++i) {
        const int c = keys[i] & 3;
        k1 += (c == 1); k2 += (c == 2); k3 += (c == 3);
        const double e = (static_cast<double>(k1) + 2.0 * static_cast<double>(k3)) * diag + static_cast<double>(k2);
        if (e / total >= pct - margin) {
            atomicMin(&ax.misc[1], i);
            break;
        }
    }
    __syncthreads();
    if (tid == 0) {
        int idx = ax.misc[1];
        bool decided = false;
        if (idx < n) {
            // recompute e at idx from the owning chunk
            const int t = idx / L;
            int q1 = ax.cnt[96 + t * 3], q2 = ax.cnt[96 + t * 3 + 1], q3 = ax.cnt[96 + t * 3 + 2];
            for (int i = t * L; i <= idx; ++i) {
                const int c = keys[i] & 3;
                q1 += (c == 1); q2 += (c == 2); q3 += (c == 3);
            }
            const double e = (static_cast<double>(q1) + 2.0 * static_cast<double>(q3)) * diag + static_cast<double>(q2);
            decided = (e / total >= pct + margin);
        }
        if (decided) {
            out[0] = key_dist(keys[idx]);
        } else {                                       // near tie (or nothing found): literal sequential replay
            double c = 0.0;
            idx = n;
            for (int i = 0; i < n; ++i) {
                c += len_value(keys[i] & 3);
                if (c / total >= pct) { idx = i; break; }
            }
            if (idx > n - 1) idx = n - 1;
            out[0] = key_dist(keys[idx]);
        }
        out[1] = ax.dres[1] / total;
    }
}

// bitonic sort of one list in shared memory + replay. CAP = power of two capacity; lists longer than CAP or
// not longer than MINN are handled by another instantiation (the block exits at once).
template <int CAP, int MINN>
__global__ void sort_replay_kernel(const unsigned int* __restrict__ keys_g2p, const unsigned int* __restrict__ keys_p2g,
                                   const int* __restrict__ count_gt, const int* __restrict__ count_pred, int cap,
                                   double pct, double* __restrict__ res, int force_sequential) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    unsigned int* sk = reinterpret_cast<unsigned int*>(smem_dyn);
    constexpr int kMaxLeaves = CAP / 64 + 2;
    ReplayAux ax;
    ax.leaf_l = reinterpret_cast<double*>(smem_dyn + static_cast<size_t>(CAP) * 4);
    ax.leaf_w = ax.leaf_l + kMaxLeaves;
    ax.dres = ax.leaf_w + kMaxLeaves;
    ax.leaf_start = reinterpret_cast<int*>(ax.dres + 2);
    ax.leaf_n = ax.leaf_start + kMaxLeaves;
    ax.cnt = ax.leaf_n + kMaxLeaves;
    ax.misc = ax.cnt + 96 + 3 * 1024;
    const int bt = blockIdx.x >> 1;
    const int dir = blockIdx.x & 1;                    // 0: gt->pred, 1: pred->gt
    const int b = bt / kNumThr;
    const int ng = count_gt[b], np_ = count_pred[bt];
    if (ng == 0 || np_ == 0) return;                   // empty cases are resolved by finalize_kernel
    const int n = dir == 0 ? ng : np_;
    if (n > CAP || n <= MINN) return;
    const unsigned int* src = (dir == 0 ? keys_g2p : keys_p2g) + static_cast<size_t>(bt) * cap;
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = threadIdx.x; i < m; i += blockDim.x) sk[i] = i < n ? src[i] : 0xFFFFFFFFu;
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned int a = sk[i], c = sk[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > c) == up) { sk[i] = c; sk[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    double* out = res + (static_cast<size_t>(bt) * 2 + dir) * 2;
    if (force_sequential) {
        if (threadIdx.x == 0) replay_list(sk, n, pct, out);
    } else {
        replay_parallel(sk, n, pct, out, ax);
    }
}

template <int CAP>
constexpr size_t sort_smem_bytes() {
    return static_cast<size_t>(CAP) * 4 + (CAP / 64 + 2) * (8 + 8 + 4 + 4) + 16 + (96 + 3 * 1024) * 4 + 16;
}

// lists that do not fit in shared memory: bitonic sort in place in global memory by one block (rare, slow path)
__global__ void sort_replay_global_kernel(unsigned int* __restrict__ keys_g2p, unsigned int* __restrict__ keys_p2g,
                                          const int* __restrict__ count_gt, const int* __restrict__ count_pred,
                                          int cap, int cap_pow2, int minn, double pct, double* __restrict__ res) {
    const int bt = blockIdx.x >> 1;
    const int dir = blockIdx.x & 1;
    const int b = bt / kNumThr;
    const int ng = count_gt[b], np_ = count_pred[bt];
    if (ng == 0 || np_ == 0) return;
    const int n = dir == 0 ? ng : np_;
    if (n <= minn) return;
    unsigned int* sk = (dir == 0 ? keys_g2p : keys_p2g) + static_cast<size_t>(bt) * cap;
    int m = 1;
    while (m < n) m <<= 1;
    if (m > cap_pow2) m = cap_pow2;                    // cap is allocated as a power of two >= NC
    for (int i = n + threadIdx.x; i < m; i += blockDim.x) sk[i] = 0xFFFFFFFFu;
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned int a = sk[i], c = sk[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > c) == up) { sk[i] = c; sk[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) replay_list(sk, n, pct, res + (static_cast<size_t>(bt) * 2 + dir) * 2);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int pow2_at_least(int v) { int m = 1; while (m < v) m <<= 1; return m; }

// per corner the range (qlo, qhi] of thresholds at which it is a prediction border, stored TRANSPOSED ([x][y], y padded to a
// multiple of 32 with the never-a-border range (255, 0)): a warp of the fused kernel reads 32 consecutive y of one column
// with one coalesced 64-byte load and its ballot is the column-packed mask word directly
__global__ void corner_range_t_kernel(const unsigned char* __restrict__ q, uchar2* __restrict__ qr, int H, int W, int Hp) {
    __shared__ uchar2 tile[32][33];
    const int Hc = H + 1, Wc = W + 1;
    const int b = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const unsigned char* qb = q + static_cast<size_t>(b) * H * W;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {              // coalesced along x
        const int y = y0 + r, x = x0 + threadIdx.x;
        uchar2 v = make_uchar2(255, 0);
        if (y < Hc && x < Wc) {
            const int a = pix(qb, y - 1, x - 1, H, W), bb = pix(qb, y - 1, x, H, W), c = pix(qb, y, x - 1, H, W), d = pix(qb, y, x, H, W);
            v = make_uchar2(static_cast<unsigned char>(min(min(a, bb), min(c, d))), static_cast<unsigned char>(max(max(a, bb), max(c, d))));
        }
        tile[r][threadIdx.x] = v;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {              // coalesced along y
        const int x = x0 + r, y = y0 + threadIdx.x;
        if (x < Wc && y < Hp) qr[(static_cast<size_t>(b) * Wc + x) * Hp + y] = tile[threadIdx.x][r];
    }
}

// gt border corners of one image as a ROW-SORTED list: list[row_start[y] .. row_start[y+1]) are the corners of corner-row y
// (one block per image: per-row counts with ballots, exclusive scan, ordered fill); count[b] = row_start[Hc]
__global__ void __launch_bounds__(1024) gt_rowlist_kernel(const unsigned char* __restrict__ cg, int* __restrict__ list,
                                                          int* __restrict__ row_start, int* __restrict__ count, int Hc, int Wc) {
    extern __shared__ int rl_sm[];                       // [Hc + 1]
    const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const unsigned char* cb = cg + static_cast<size_t>(b) * Hc * Wc;
    for (int y = warp; y < Hc; y += nwarps) {
        int n = 0;
        for (int x0 = 0; x0 < Wc; x0 += 32) {
            const int x = x0 + lane;
            const int c = x < Wc ? cb[y * Wc + x] : 0;
            n += __popc(__ballot_sync(0xFFFFFFFFu, c != 0 && c != 15));
        }
        if (lane == 0) rl_sm[y] = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int y = 0; y < Hc; ++y) { const int n = rl_sm[y]; rl_sm[y] = acc; acc += n; }
        rl_sm[Hc] = acc;
        count[b] = acc;
    }
    __syncthreads();
    for (int y = threadIdx.x; y <= Hc; y += blockDim.x) row_start[static_cast<size_t>(b) * (Hc + 1) + y] = rl_sm[y];
    int* lb = list + static_cast<size_t>(b) * Hc * Wc;
    for (int y = warp; y < Hc; y += nwarps) {
        int base = rl_sm[y];
        for (int x0 = 0; x0 < Wc; x0 += 32) {
            const int x = x0 + lane;
            const int c = x < Wc ? cb[y * Wc + x] : 0;
            const bool on = c != 0 && c != 15;
            const unsigned int m = __ballot_sync(0xFFFFFFFFu, on);
            if (on) lb[base + __popc(m & ((1u << lane) - 1u))] = y * Wc + x;
            base += __popc(m);
        }
    }
}

// ------------------------------------------------------------------ fused per-(image, threshold) HD / MSD
// One persistent CTA per SM takes (image, threshold) items from a global counter and keeps everything of the item on chip:
//   1. the prediction-border corners of the threshold as a column-packed BIT MASK in shared memory ((W+1) columns x
//      ceil((H+1)/32) words, 27 KB at 448^2), built with warp ballots (32x32 bit-tile transposes) straight from the map q;
//   2. distances_gt_to_pred: the exact EDT as column pass + row search, in bands of 32 corner rows: every column thread turns
//      its mask words into the vertical distances of the band's rows (clz / ffs for the nearest bits above / below, one sweep
//      inside the band's own word), the band's [32][W+1] uint16 distances stay in shared memory, and the gt border corners of
//      those rows (row-sorted list) run the expanding row search with early exit -- the per-threshold column-distance map
//      (99 x 449^2 uint16 per image) never exists;
//   3. distances_pred_to_gt: every set bit looks its squared distance up in the per-image gt EDT map d2g;
//   4. both key lists are sorted in shared memory (bitonic, <= 32768 keys) and replayed in numpy's floating-point order;
//      longer lists spill to a per-CTA global scratch area (one region per resident CTA, not per image).
// This replaces column_scan<1> / g2p_keys / p2g_keys / sort_replay<*> / finalize and their 260 MB-per-image workspace
// (99 x 449^2 uint16 column distances + 2 x 99 key lists per image).
constexpr int kFusedThreads = 1024;
constexpr int kFusedCap = 32768;                            // keys held in shared memory
constexpr int kFusedLeaves = kFusedCap / 64 + 2;
__host__ __device__ constexpr size_t fused_aux_bytes() {   // leaf tables + per-thread class counts of replay_parallel, 16-byte multiple
    return (static_cast<size_t>(kFusedLeaves) * (8 + 8 + 4 + 4) + 16 + (96 + 3 * kFusedThreads) * 4 + 64 + 15) / 16 * 16;
}

struct FusedArgs {
    const unsigned char* q;
    const uchar2* qr;                                       // [B][Wc][Hp] transposed threshold ranges (corner_range_t_kernel)
    const unsigned char* cg;
    const unsigned int* d2g;
    const int* gt_list;                                     // row-sorted (gt_rowlist_kernel)
    const int* row_start;                                   // [B][Hc + 1]
    const int* count_gt;
    int H, W, total_items;
    double pct, max_img_len;
    double* hd;
    double* msd;
    unsigned int* scratch;                                  // [gridDim.x][2][scratch_cap]: spilled key list + its sort target
    int scratch_cap;                                        // power of two >= (H+1)*(W+1)
    int* counter;
    int force_seq;
};

// Counting sort of keys[0..n) in shared memory when every key is below `kmax` <= cap - n: the histogram lives in
// keys[n .. n + kmax) of the same buffer.  O(n + kmax) instead of the O(n log^2 n) bitonic network: distances of a few dozen
// pixels (keys = d^2 * 4 + class < 16384) are the common case.  Returns false (buffer untouched) when the range is too large.
__device__ bool block_counting_sort(unsigned int* keys, int n, int cap, int* s_tmp /* [34] shared */) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    unsigned int mx = 0;
    for (int i = tid; i < n; i += nt) mx = max(mx, keys[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    if (lane == 0) s_tmp[warp] = static_cast<int>(mx);
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int w = 0; w < nwarps; ++w) m = max(m, s_tmp[w]);
        s_tmp[32] = m;
    }
    __syncthreads();
    const long long kmax = static_cast<long long>(static_cast<unsigned int>(s_tmp[32])) + 1;
    if (kmax > cap - n || kmax > 4LL * n + 1024) { __syncthreads(); return false; }
    const int K = static_cast<int>(kmax);
    unsigned int* hist = keys + n;
    for (int i = tid; i < K; i += nt) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) atomicAdd(&hist[keys[i]], 1u);
    __syncthreads();
    // exclusive scan of hist over the block: contiguous chunks per thread
    const int L = (K + nt - 1) / nt;
    const int c0 = min(K, tid * L), c1 = min(K, c0 + L);
    int sum = 0;
    for (int i = c0; i < c1; ++i) sum += static_cast<int>(hist[i]);
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_tmp[warp] = inc;
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int w = 0; w < nwarps; ++w) { const int v = s_tmp[w]; s_tmp[w] = acc; acc += v; }
    }
    __syncthreads();
    int pos = s_tmp[warp] + inc - sum;                        // first output index of this thread's key values
    __syncthreads();
    // every key value writes its run; reading hist and writing keys[0..n) never overlap (hist sits behind the keys)
    for (int i = c0; i < c1; ++i) {
        const int cnt = static_cast<int>(hist[i]);
        for (int j = 0; j < cnt; ++j) keys[pos + j] = static_cast<unsigned int>(i);
        pos += cnt;
    }
    __syncthreads();
    return true;
}

// Sort of a LONG key list (n > kFusedCap, living in the CTA's global scratch `A`) into the second scratch region `B`:
// most-significant-digit bucketing on the top 12 bits of the key range (histogram + cursors in shared memory, one scatter
// A -> B), then every bucket -- a contiguous segment of B whose keys differ only in their low <= 9 bits -- is counting-sorted
// by one warp with a private shared-memory histogram that regenerates the bucket's runs in place.  O(n) work, two passes over
// the keys, instead of the ~150 global-memory passes of the bitonic network (10.8 of the 16.5 ms of the App. E fixture).
// `sm` / `sm_words` = free shared memory for the histograms (the 32768-word key area while the list is in global memory).
// Returns B, or nullptr (A untouched) when the key range needs more than 21 bits or the histograms do not fit.
__device__ unsigned int* block_msd_sort_global(const unsigned int* A, unsigned int* B, int n, unsigned int* sm, int sm_words,
                                               int* s_tmp) {
    constexpr int kHiBits = 12, kBuckets = 1 << kHiBits, kLoMax = 9;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    unsigned int mx = 0;
    for (int i = tid; i < n; i += nt) mx = max(mx, A[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    if (lane == 0) s_tmp[warp] = static_cast<int>(mx);
    __syncthreads();
    if (tid == 0) {
        unsigned int m = 0;
        for (int w = 0; w < nwarps; ++w) m = max(m, static_cast<unsigned int>(s_tmp[w]));
        s_tmp[32] = static_cast<int>(m);
    }
    __syncthreads();
    const unsigned int kmax = static_cast<unsigned int>(s_tmp[32]);
    int shift = 0;
    while ((kmax >> shift) >= static_cast<unsigned int>(kBuckets)) ++shift;
    __syncthreads();
    if (shift > kLoMax || nwarps * (1 << shift) + 2 * kBuckets + 2 > sm_words) return nullptr;
    unsigned int* off = sm;                              // [kBuckets + 1]
    unsigned int* cur = sm + kBuckets + 1;               // [kBuckets]
    unsigned int* whist = sm + 2 * kBuckets + 1;         // [nwarps][1 << shift]
    for (int i = tid; i <= kBuckets; i += nt) off[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) atomicAdd(&off[A[i] >> shift], 1u);
    __syncthreads();
    {   // exclusive scan of the bucket counts: contiguous chunks per thread
        const int L = (kBuckets + nt - 1) / nt;
        const int c0 = min(kBuckets, tid * L), c1 = min(kBuckets, c0 + L);
        int sum = 0;
        for (int i = c0; i < c1; ++i) sum += static_cast<int>(off[i]);
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_tmp[warp] = inc;
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int w = 0; w < nwarps; ++w) { const int v = s_tmp[w]; s_tmp[w] = acc; acc += v; }
        }
        __syncthreads();
        int pos = s_tmp[warp] + inc - sum;
        for (int i = c0; i < c1; ++i) {
            const int cnt = static_cast<int>(off[i]);
            off[i] = static_cast<unsigned int>(pos);
            cur[i] = static_cast<unsigned int>(pos);
            pos += cnt;
        }
        if (tid == 0) off[kBuckets] = static_cast<unsigned int>(n);
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        const unsigned int k = A[i];
        B[atomicAdd(&cur[k >> shift], 1u)] = k;
    }
    __syncthreads();                                     // (block-scope: the scattered keys are visible to every warp)
    const int nbins = 1 << shift, per_lane = (nbins + 31) >> 5;
    const unsigned int lomask = static_cast<unsigned int>(nbins - 1);
    unsigned int* wh = whist + warp * nbins;
    if (shift > 0) {
        for (int b = warp; b < kBuckets; b += nwarps) {
            const int lo = static_cast<int>(off[b]), hi = static_cast<int>(off[b + 1]);
            if (hi - lo < 2) continue;                   // warp-uniform
            for (int v = lane; v < nbins; v += 32) wh[v] = 0;
            __syncwarp();
            for (int i = lo + lane; i < hi; i += 32) atomicAdd(&wh[B[i] & lomask], 1u);
            __syncwarp();
            // lane owns the bins [lane * per_lane, ...): exclusive prefix across the warp, then it writes its runs
            const int v0 = min(nbins, lane * per_lane), v1 = min(nbins, v0 + per_lane);
            int sum = 0;
            for (int v = v0; v < v1; ++v) sum += static_cast<int>(wh[v]);
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t2 = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += t2;
            }
            int pos = lo + inc - sum;
            const unsigned int base = static_cast<unsigned int>(b) << shift;
            for (int v = v0; v < v1; ++v) {
                const int cnt = static_cast<int>(wh[v]);
                for (int j = 0; j < cnt; ++j) B[pos + j] = base | static_cast<unsigned int>(v);
                pos += cnt;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    return B;
}

// block-wide bitonic sort of keys[0..n) (keys[n..m) padded with 0xFFFFFFFF, m = next power of two <= capacity)
__device__ void block_bitonic(unsigned int* keys, int n) {
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = n + threadIdx.x; i < m; i += blockDim.x) keys[i] = 0xFFFFFFFFu;
    __syncthreads();
    for (int k = 2; k <= m; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            // one compare-exchange per pair index p: i = the p-th index with bit j clear, partner i | j
            for (int p = threadIdx.x; p < (m >> 1); p += blockDim.x) {
                const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                const unsigned int a = keys[i], c = keys[i | j];
                const bool up = (i & k) == 0;
                if ((a > c) == up) { keys[i] = c; keys[i | j] = a; }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(kFusedThreads, 1) hd_fused_kernel(const FusedArgs a) {
    extern __shared__ __align__(16) unsigned char fsm[];
    const int Hc = a.H + 1, Wc = a.W + 1, WH = (Hc + 31) / 32, NC = Hc * Wc;
    unsigned int* bits = reinterpret_cast<unsigned int*>(fsm);                       // [Wc][WH]: bit r of word (x, k) = corner (32k + r, x)
    unsigned int* skeys = bits + ((Wc * WH + 3) & ~3);
    unsigned char* aux_base = reinterpret_cast<unsigned char*>(skeys + kFusedCap);
    unsigned short* gbuf = reinterpret_cast<unsigned short*>(aux_base + fused_aux_bytes());   // [32][Wc] vertical distances of a band
    unsigned short* gminb = gbuf + 32 * Wc;                                                   // [32][ceil(Wc/16)] block minima
    __shared__ int s_item, s_np, s_cnt;
    __shared__ int s_tmp[34];
    __shared__ double s_res[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = kFusedThreads / 32;
    unsigned int* gkeys = a.scratch + static_cast<size_t>(blockIdx.x) * 2 * a.scratch_cap;      // [2][cap]: list, sort target

    auto make_aux = [&](unsigned char* base, int max_leaves) {
        ReplayAux ax;
        ax.leaf_l = reinterpret_cast<double*>(base);
        ax.leaf_w = ax.leaf_l + max_leaves;
        ax.dres = ax.leaf_w + max_leaves;
        ax.leaf_start = reinterpret_cast<int*>(ax.dres + 2);
        ax.leaf_n = ax.leaf_start + max_leaves;
        ax.cnt = ax.leaf_n + max_leaves;
        ax.misc = ax.cnt + 96 + 3 * kFusedThreads;
        return ax;
    };
    // sort + replay of the list that was just written to `keys` (shared or this CTA's global scratch)
    auto sort_replay = [&](unsigned int* keys, int n, bool in_smem, double* out) {
        if (a.force_seq & 4) { if (tid == 0) { out[0] = 0; out[1] = 0; } __syncthreads(); return; }
        bool sorted = false;
        if (!(a.force_seq & 16)) {
            if (in_smem) {
                // (short lists with a wide key range stay on the bitonic network: routing them through the MSD sort below was
                // measured slower, 3.27 vs 2.92 ms per 16 images -- 4096 mostly empty buckets cost more than ~70 smem passes)
                sorted = block_counting_sort(keys, n, kFusedCap, s_tmp);
            } else {
                unsigned int* r = block_msd_sort_global(keys, keys + a.scratch_cap, n, skeys, kFusedCap, s_tmp);
                if (r) { keys = r; sorted = true; }
            }
        }
        if (!sorted) block_bitonic(keys, n);
        if (a.force_seq & 8) { if (tid == 0) { out[0] = 0; out[1] = 0; } __syncthreads(); return; }
        if (a.force_seq & 1) {
            if (tid == 0) replay_list(keys, n, a.pct, out);
        } else if (in_smem) {
            replay_parallel(keys, n, a.pct, out, make_aux(aux_base, kFusedLeaves));
        } else {                                              // long list: the key area of shared memory is free for the leaves
            replay_parallel(keys, n, a.pct, out, make_aux(reinterpret_cast<unsigned char*>(skeys), a.scratch_cap / 64 + 2));
        }
        __syncthreads();
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) { s_item = atomicAdd(a.counter, 1); s_np = 0; s_cnt = 0; }
        __syncthreads();
        const int item = s_item;
        if (item >= a.total_items) break;
        const int b = item / kNumThr, t = item % kNumThr + 1;
        const int ng = a.count_gt[b];
        const unsigned char* qb = a.q + static_cast<size_t>(b) * a.H * a.W;
        // ---- 1. border mask of threshold t: corner (y, x) is a border iff qlo < t <= qhi of its 2x2 pixel block; one coalesced
        // load of 32 consecutive y of a column + one ballot per mask word
        int local_np = 0;
        {
            const uchar2* qrb = a.qr + static_cast<size_t>(b) * Wc * (WH * 32);
            // four independent loads in flight per warp: one dependent L2 load per mask word made this loop latency-bound
            // (~210 round trips per warp and item, the largest single stall of the kernel in ncu's source view)
            const int nwords = Wc * WH;
            for (int w0 = warp; w0 < nwords; w0 += 4 * nwarps) {
                uchar2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int widx = w0 + u * nwarps;                                   // widx = x * WH + k  ->  (x, y = 32k + lane)
                    v[u] = widx < nwords ? qrb[static_cast<size_t>(widx) * 32 + lane] : make_uchar2(255, 0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int widx = w0 + u * nwarps;
                    const unsigned int word = __ballot_sync(0xFFFFFFFFu, v[u].x < t && t <= v[u].y);
                    if (lane == 0 && widx < nwords) { bits[widx] = word; local_np += __popc(word); }
                }
            }
        }
        local_np = warp_sum(local_np);
        if (lane == 0 && local_np) atomicAdd(&s_np, local_np);
        __syncthreads();
        const int np_ = s_np;
        if (ng == 0 || np_ == 0) {                            // empty-mask branches (inference.py:315-321)
            if (tid == 0) {
                const double v = (ng == 0 && np_ == 0) ? 0.0 : a.max_img_len;
                a.hd[item] = v;
                a.msd[item] = v;
            }
            continue;
        }
        // ---- 2. distances_gt_to_pred: column pass per band of 32 corner rows, then the row search of the band's gt corners
        {
            const bool in_smem = ng <= kFusedCap;
            unsigned int* keys = in_smem ? skeys : gkeys;
            const int* list = a.gt_list + static_cast<size_t>(b) * NC;
            const int* rs = a.row_start + static_cast<size_t>(b) * (Hc + 1);
            const unsigned char* cgb = a.cg + static_cast<size_t>(b) * NC;
            for (int ky = 0; ky < WH; ++ky) {
                const int y0 = ky * 32, y1 = min(Hc, y0 + 32);
                const int k0 = rs[y0], k1 = rs[y1];
                if (k0 == k1) continue;                                    // no gt border corner in these rows (uniform branch)
                for (int x = tid; x < Wc; x += kFusedThreads) {
                    const unsigned int* col = bits + x * WH;
                    int last = -0x10000, next = 0x20000;                   // nearest set bit above / below the band
                    for (int k = ky - 1; k >= 0; --k) { const unsigned int v = col[k]; if (v) { last = k * 32 + 31 - __clz(v); break; } }
                    for (int k = ky + 1; k < WH; ++k) { const unsigned int v = col[k]; if (v) { next = k * 32 + __ffs(v) - 1; break; } }
                    const unsigned int w = col[ky];
                    for (int r = 0; r < y1 - y0; ++r) {
                        if ((w >> r) & 1u) last = y0 + r;
                        gbuf[r * Wc + x] = static_cast<unsigned short>(min(y0 + r - last, 0xFFFF));
                    }
                    for (int r = y1 - y0 - 1; r >= 0; --r) {
                        if ((w >> r) & 1u) next = y0 + r;
                        const int dn = next - (y0 + r);
                        if (dn < gbuf[r * Wc + x]) gbuf[r * Wc + x] = static_cast<unsigned short>(dn);
                    }
                }
                __syncthreads();
                const int nblk = (Wc + 15) >> 4;
                for (int e = tid; e < (y1 - y0) * nblk; e += kFusedThreads) {        // block minima of the band's distance rows
                    const int r = e / nblk, blk = e - r * nblk;
                    const unsigned short* gr = gbuf + r * Wc + (blk << 4);
                    unsigned int mn = kInf16;
                    for (int j = 0; j < min(16, Wc - (blk << 4)); ++j) mn = min(mn, static_cast<unsigned int>(gr[j]));
                    gminb[r * nblk + blk] = static_cast<unsigned short>(mn);
                }
                __syncthreads();
                for (int k = k0 + tid; k < k1; k += kFusedThreads) {
                    const int i = list[k];
                    const int y = i / Wc, x = i - y * Wc;
                    const unsigned int d2 = (a.force_seq & 2) ? 1u : row_search_blocked(gbuf + (y - y0) * Wc, gminb + (y - y0) * nblk, x, Wc);
                    keys[k] = (d2 << 2) | len_class(cgb[i]);
                }
                __syncthreads();
            }
            sort_replay(keys, ng, in_smem, &s_res[0]);
        }
        // ---- 3. distances_pred_to_gt
        {
            const bool in_smem = np_ <= kFusedCap;
            unsigned int* keys = in_smem ? skeys : gkeys;
            const unsigned int* d2b = a.d2g + static_cast<size_t>(b) * NC;
            for (int widx = tid; widx < Wc * WH; widx += kFusedThreads) {          // one thread per mask word: most words are empty
                unsigned int word = bits[widx];
                if (word == 0) continue;
                int slot = atomicAdd(&s_cnt, __popc(word));
                const int x = widx / WH, ybase = (widx - x * WH) * 32;
                while (word) {
                    const int y = ybase + __ffs(word) - 1;
                    word &= word - 1;
                    const int p00 = pix(qb, y - 1, x - 1, a.H, a.W), p01 = pix(qb, y - 1, x, a.H, a.W), p10 = pix(qb, y, x - 1, a.H, a.W),
                              p11 = pix(qb, y, x, a.H, a.W);
                    const int code = ((p00 >= t) << 3) | ((p01 >= t) << 2) | ((p10 >= t) << 1) | (p11 >= t);
                    keys[slot++] = (d2b[y * Wc + x] << 2) | len_class(code);
                }
            }
            __syncthreads();
            sort_replay(keys, np_, in_smem, &s_res[2]);
        }
        if (tid == 0) {
            a.hd[item] = fmax(s_res[0], s_res[2]);             // compute_robust_hausdorff: max of the two directions
            a.msd[item] = (s_res[1] + s_res[3]) / 2;            // inference.py:327-334
        }
    }
}

static size_t fused_smem_bytes(int h, int w) {
    const int Wc = w + 1, WH = (h + 1 + 31) / 32;
    const size_t bits = static_cast<size_t>((Wc * WH + 3) & ~3) * 4;
    return bits + static_cast<size_t>(kFusedCap) * 4 + fused_aux_bytes() + static_cast<size_t>(32) * Wc * 2 + 32 * ((Wc + 15) / 16) * 2 + 16;
}
static bool fused_ok(int h, int w) {
    // the long-list replay borrows the key area for its leaf tables: (cap / 64 + 2) * 24 B + counters must fit into it
    const size_t cap = static_cast<size_t>(pow2_at_least((h + 1) * (w + 1)));
    const size_t long_aux = (cap / 64 + 2) * 24 + 16 + (96 + 3 * kFusedThreads) * 4 + 64;
    return !getenv("CSBSR_METRICS_UNFUSED") && fused_smem_bytes(h, w) <= 220 * 1024 && h + 1 < 0xFFFF && long_aux <= static_cast<size_t>(kFusedCap) * 4;
}

// combine the two directions and resolve the empty-mask branches (inference.py:315-334)
__global__ void finalize_kernel(const int* __restrict__ count_gt, const int* __restrict__ count_pred,
                                const double* __restrict__ res, double* __restrict__ hd, double* __restrict__ msd,
                                int total, double max_img_len) {
    const int bt = blockIdx.x * blockDim.x + threadIdx.x;
    if (bt >= total) return;
    const int b = bt / kNumThr;
    const int ng = count_gt[b], np_ = count_pred[bt];
    if (ng == 0 && np_ == 0) {
        hd[bt] = 0.0;
        msd[bt] = 0.0;
    } else if (ng == 0 || np_ == 0) {
        hd[bt] = max_img_len;
        msd[bt] = max_img_len;
    } else {
        const double* r = res + static_cast<size_t>(bt) * 4;
        hd[bt] = fmax(r[0], r[2]);
        msd[bt] = (r[1] + r[3]) / 2;
    }
}


struct MetricsWs {
    unsigned char *q, *gt, *qlo, *qhi, *cg;
    int *hist, *gt_list, *count_gt, *count_pred, *row_start;
    unsigned short *gcol_g, *gcol_t;
    unsigned int *d2g, *keys_g2p, *keys_p2g;
    double* res;
    uchar2* qr;                  // fused path: transposed per-corner threshold ranges
    unsigned int* scratch;       // fused path: [ctas][cap] spill area of the long key lists
    int* counter;                // fused path: work-item counter
    int cap, ctas;
    bool fused;
    size_t total;
};

static MetricsWs carve(void* base, int b, int h, int w, bool with_hd) {
    MetricsWs m;
    const size_t HW = static_cast<size_t>(h) * w, NC = static_cast<size_t>(h + 1) * (w + 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { void* p = base ? static_cast<char*>(base) + off : nullptr; off += align_up(bytes, 256); return p; };
    m.q = static_cast<unsigned char*>(take(b * HW));
    m.gt = static_cast<unsigned char*>(take(b * HW));
    m.hist = static_cast<int*>(take(sizeof(int) * b * 200));
    m.cap = pow2_at_least(static_cast<int>(NC));
    m.fused = with_hd && fused_ok(h, w);
    m.ctas = num_sms();
    if (m.fused) {
        // per image: q, gt, cg (1 B / pixel or corner), gt border list, gt column distances + EDT map; per resident CTA: one
        // spill region (list + sort target) for key lists longer than 32768 -- ~2.9 MB per 448^2 image + 310 MB per device instead of 260 MB per image
        m.qlo = m.qhi = nullptr;
        m.qr = static_cast<uchar2*>(take(sizeof(uchar2) * static_cast<size_t>(b) * (w + 1) * ((h + 1 + 31) / 32 * 32)));
        m.cg = static_cast<unsigned char*>(take(b * NC));
        m.gt_list = static_cast<int*>(take(sizeof(int) * b * NC));
        m.count_gt = static_cast<int*>(take(sizeof(int) * b));
        m.row_start = static_cast<int*>(take(sizeof(int) * static_cast<size_t>(b) * (h + 2)));
        m.count_pred = nullptr;
        m.gcol_g = static_cast<unsigned short*>(take(sizeof(short) * b * NC));
        m.d2g = static_cast<unsigned int*>(take(sizeof(int) * b * NC));
        m.gcol_t = nullptr;
        m.keys_g2p = m.keys_p2g = nullptr;
        m.res = nullptr;
        m.scratch = static_cast<unsigned int*>(take(sizeof(int) * static_cast<size_t>(m.ctas) * 2 * m.cap));
        m.counter = static_cast<int*>(take(sizeof(int)));
    } else if (with_hd) {
        m.qlo = static_cast<unsigned char*>(take(b * NC));
        m.qhi = static_cast<unsigned char*>(take(b * NC));
        m.cg = static_cast<unsigned char*>(take(b * NC));
        m.gt_list = static_cast<int*>(take(sizeof(int) * b * NC));
        m.count_gt = static_cast<int*>(take(sizeof(int) * b));
        m.count_pred = static_cast<int*>(take(sizeof(int) * b * kNumThr));
        m.gcol_g = static_cast<unsigned short*>(take(sizeof(short) * b * NC));
        m.d2g = static_cast<unsigned int*>(take(sizeof(int) * b * NC));
        m.gcol_t = static_cast<unsigned short*>(take(sizeof(short) * b * kNumThr * NC));
        m.keys_g2p = static_cast<unsigned int*>(take(sizeof(int) * static_cast<size_t>(b) * kNumThr * m.cap));
        m.keys_p2g = static_cast<unsigned int*>(take(sizeof(int) * static_cast<size_t>(b) * kNumThr * m.cap));
        m.res = static_cast<double*>(take(sizeof(double) * b * kNumThr * 4));
    }
    m.total = off;
    return m;
}

}  // namespace csbsr

using namespace csbsr;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" size_t csbsr_metrics_workspace_bytes(int b, int h, int w, int with_hd) {
    return carve(nullptr, b, h, w, with_hd != 0).total;
}

extern "C" int csbsr_seg_metrics(const float* prob, const float* mask, const float* thresholds, int b, int h, int w,
                                 long long* inter, long long* uni, double* hd, double* msd, double percent,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = STREAM(stream_);
    CSBSR_REQUIRE(prob && mask && thresholds && inter && uni && workspace && b > 0 && h > 0 && w > 0,
                  "seg_metrics: bad arguments");
    CSBSR_REQUIRE((hd == nullptr) == (msd == nullptr), "seg_metrics: hd and msd must come together");
    CSBSR_REQUIRE(h <= 16384 && w <= 16384, "seg_metrics: image too large (squared distances are packed in 30 bits)");
    const bool with_hd = hd != nullptr;
    MetricsWs m = carve(workspace, b, h, w, with_hd);
    CSBSR_REQUIRE(workspace_bytes >= m.total, "seg_metrics: workspace too small (%zu < %zu)", workspace_bytes, m.total);
    const int HW = h * w, Hc = h + 1, Wc = w + 1, NC = Hc * Wc;
    CSBSR_CHECK_CUDA(cudaMemsetAsync(m.hist, 0, sizeof(int) * b * 200, stream));
    {
        int slices = (HW + 256 * 8 - 1) / (256 * 8);
        dim3 grid(slices, b);
        quantize_hist_kernel<<<grid, 256, 0, stream>>>(prob, mask, thresholds, m.q, m.gt, m.hist, HW);
        aiu_counts_kernel<<<b, 128, 0, stream>>>(m.hist, inter, uni);
    }
    if (with_hd && m.fused) {
        CSBSR_CHECK_CUDA(cudaMemsetAsync(m.counter, 0, sizeof(int), stream));
        const int cslices = (NC + 255) / 256;
        corner_kernel<<<dim3(cslices, b), 256, 0, stream>>>(m.q, m.gt, nullptr, nullptr, m.cg, h, w);
        const int Hp = (Hc + 31) / 32 * 32;
        corner_range_t_kernel<<<dim3((Wc + 31) / 32, Hp / 32, b), dim3(32, 8), 0, stream>>>(m.q, m.qr, h, w, Hp);
        column_scan_kernel<0><<<dim3((Wc + 127) / 128, b), 128, 0, stream>>>(m.cg, nullptr, nullptr, m.gcol_g, Hc, Wc);
        gt_edt_kernel<<<dim3(Hc, b), 256, sizeof(short) * (((Wc + 7) & ~7) + ((Wc + 15) / 16) + 8), stream>>>(m.gcol_g, m.d2g, Hc, Wc);
        gt_rowlist_kernel<<<b, 1024, sizeof(int) * (Hc + 1), stream>>>(m.cg, m.gt_list, m.row_start, m.count_gt, Hc, Wc);
        FusedArgs fa;
        fa.q = m.q; fa.qr = m.qr; fa.cg = m.cg; fa.d2g = m.d2g; fa.gt_list = m.gt_list; fa.row_start = m.row_start; fa.count_gt = m.count_gt;
        fa.H = h; fa.W = w; fa.total_items = b * kNumThr;
        fa.pct = percent / 100.0; fa.max_img_len = static_cast<double>(w);
        fa.hd = hd; fa.msd = msd;
        fa.scratch = m.scratch; fa.scratch_cap = m.cap; fa.counter = m.counter;
        fa.force_seq = (getenv("CSBSR_METRICS_SEQUENTIAL") ? 1 : 0) | (getenv("CSBSR_HD_DEBUG") ? atoi(getenv("CSBSR_HD_DEBUG")) : 0);
        const int smem = static_cast<int>(fused_smem_bytes(h, w));
        static bool fattr = false;
        if (!fattr) {
            CSBSR_CHECK_CUDA(cudaFuncSetAttribute(hd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            fattr = true;
        }
        const int grid = fa.total_items < m.ctas ? fa.total_items : m.ctas;
        hd_fused_kernel<<<grid, kFusedThreads, smem, stream>>>(fa);
    } else if (with_hd) {
        CSBSR_CHECK_CUDA(cudaMemsetAsync(m.count_gt, 0, sizeof(int) * b, stream));
        CSBSR_CHECK_CUDA(cudaMemsetAsync(m.count_pred, 0, sizeof(int) * b * kNumThr, stream));
        const int cslices = (NC + 255) / 256;
        corner_kernel<<<dim3(cslices, b), 256, 0, stream>>>(m.q, m.gt, m.qlo, m.qhi, m.cg, h, w);
        column_scan_kernel<0><<<dim3((Wc + 127) / 128, b), 128, 0, stream>>>(m.cg, m.qlo, m.qhi, m.gcol_g, Hc, Wc);
        gt_edt_kernel<<<dim3(Hc, b), 256, sizeof(short) * (((Wc + 7) & ~7) + ((Wc + 15) / 16) + 8), stream>>>(m.gcol_g, m.d2g, Hc, Wc);
        gt_list_kernel<<<dim3(cslices, b), 256, 0, stream>>>(m.cg, m.gt_list, m.count_gt, NC);
        column_scan_kernel<1><<<dim3((Wc + 127) / 128, b * kNumThr), 128, 0, stream>>>(m.cg, m.qlo, m.qhi, m.gcol_t,
                                                                                      Hc, Wc);
        g2p_keys_kernel<<<dim3(64, b * kNumThr), 256, 0, stream>>>(m.gt_list, m.count_gt, m.cg, m.gcol_t, m.keys_g2p,
                                                                  Hc, Wc, m.cap);
        p2g_keys_kernel<<<dim3(cslices, b), 256, 0, stream>>>(m.q, m.qlo, m.qhi, m.d2g, m.keys_p2g, m.count_pred, h, w,
                                                             m.cap);
        const double pct = percent / 100.0;
        const int lists = b * kNumThr * 2;
        static bool attr_set = false;
        if (!attr_set) {
            CSBSR_CHECK_CUDA(cudaFuncSetAttribute(sort_replay_kernel<32768, 4096>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  static_cast<int>(sort_smem_bytes<32768>())));
            attr_set = true;
        }
        const int force_seq = getenv("CSBSR_METRICS_SEQUENTIAL") ? 1 : 0;
        sort_replay_kernel<4096, 0><<<lists, 256, sort_smem_bytes<4096>(), stream>>>(
            m.keys_g2p, m.keys_p2g, m.count_gt, m.count_pred, m.cap, pct, m.res, force_seq);
        sort_replay_kernel<32768, 4096><<<lists, 1024, sort_smem_bytes<32768>(), stream>>>(
            m.keys_g2p, m.keys_p2g, m.count_gt, m.count_pred, m.cap, pct, m.res, force_seq);
        sort_replay_global_kernel<<<lists, 1024, 0, stream>>>(m.keys_g2p, m.keys_p2g, m.count_gt, m.count_pred, m.cap,
                                                              m.cap, 32768, pct, m.res);
        finalize_kernel<<<(b * kNumThr + 127) / 128, 128, 0, stream>>>(m.count_gt, m.count_pred, m.res, hd, msd,
                                                                      b * kNumThr, static_cast<double>(w));
    }
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// PSNR and SSIM of the SR image (model/utils/estimate_metrics.py:89-100 PSNR, :134-191 SSIM: 11x11 Gaussian window,
// sigma 1.5, zero padding 5, C1 = 0.01^2, C2 = 0.03^2, per-image mean over (C,H,W)); evaluated by inference_for_ss
// (model/engine/inference.py:94-100).  One kernel: each block produces a 16x16 tile of the SSIM map of one (image, channel)
// plane from a 26x26 input tile of both images (separable window in shared memory) and accumulates the per-image sums.
namespace csbsr {

struct Gauss11 {
    float v[11];
};

__global__ void psnr_ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int C, int H, int W,
                                 double* __restrict__ acc /* [B][2]: sum sq diff, sum ssim */, const Gauss11 g11) {
    constexpr int T = 16, R = 5, P = T + 2 * R;
    __shared__ float sa[P][P + 1], sb[P][P + 1];
    __shared__ float hq[5][P][T + 1];
    __shared__ float sg[11];
    const int plane = blockIdx.z;                 // image * C + channel
    const int img = plane / C;
    const int y0 = blockIdx.y * T, x0 = blockIdx.x * T;
    const float* ap = a + static_cast<size_t>(plane) * H * W;
    const float* bp = b + static_cast<size_t>(plane) * H * W;
    if (threadIdx.x < 11) sg[threadIdx.x] = g11.v[threadIdx.x];
    for (int i = threadIdx.x; i < P * P; i += blockDim.x) {
        const int r = i / P, c = i % P;
        const int yy = y0 + r - R, xx = x0 + c - R;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
        sa[r][c] = in ? ap[static_cast<size_t>(yy) * W + xx] : 0.f;
        sb[r][c] = in ? bp[static_cast<size_t>(yy) * W + xx] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * T; i += blockDim.x) {          // horizontal pass
        const int r = i / T, c = i % T;
        float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float u = sa[r][c + k], v = sb[r][c + k], w = sg[k];
            m1 = fmaf(w, u, m1); m2 = fmaf(w, v, m2);
            s11 = fmaf(w, u * u, s11); s22 = fmaf(w, v * v, s22); s12 = fmaf(w, u * v, s12);
        }
        hq[0][r][c] = m1; hq[1][r][c] = m2; hq[2][r][c] = s11; hq[3][r][c] = s22; hq[4][r][c] = s12;
    }
    __syncthreads();
    const int ty = threadIdx.x / T, tx = threadIdx.x % T;          // 256 threads = 16 x 16 outputs
    double ssim = 0.0, sq = 0.0;
    if (y0 + ty < H && x0 + tx < W) {
        float q[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = sg[k];
#pragma unroll
            for (int j = 0; j < 5; ++j) q[j] = fmaf(w, hq[j][ty + k][tx], q[j]);
        }
        const float mu1 = q[0], mu2 = q[1];
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
        const float s1 = q[2] - mu1_sq, s2 = q[3] - mu2_sq, s12 = q[4] - mu12;
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        ssim = ((2.f * mu12 + C1) * (2.f * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2));
        const float d = sa[ty + R][tx + R] - sb[ty + R][tx + R];
        sq = static_cast<double>(d) * d;
    }
    ssim = warp_sum(ssim);
    sq = warp_sum(sq);
    __shared__ double red[2][8];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    if (lane == 0) { red[0][wp] = sq; red[1][wp] = ssim; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
        atomicAdd(&acc[img * 2 + threadIdx.x], t);
    }
}

__global__ void psnr_ssim_finish_kernel(const double* __restrict__ acc, double* __restrict__ psnr, double* __restrict__ ssim, int B,
                                        double n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const float mse = static_cast<float>(acc[i * 2] / n);          // the reference reduces in fp32 (torch.mean)
    psnr[i] = static_cast<double>(10.f * log10f(1.f / mse));
    ssim[i] = static_cast<double>(static_cast<float>(acc[i * 2 + 1] / n));
}

}  // namespace csbsr

extern "C" size_t csbsr_psnr_ssim_workspace_bytes(int b) { return sizeof(double) * 2 * static_cast<size_t>(b); }

extern "C" int csbsr_psnr_ssim(const float* pred, const float* target, int b, int c, int h, int w, double* psnr, double* ssim,
                               void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    CSBSR_REQUIRE(pred && target && psnr && ssim && workspace && b > 0 && c > 0 && h > 0 && w > 0, "psnr_ssim: bad arguments");
    CSBSR_REQUIRE(workspace_bytes >= csbsr_psnr_ssim_workspace_bytes(b), "psnr_ssim: workspace too small");
    double* acc = static_cast<double*>(workspace);
    CSBSR_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * b, stream));
    // gaussian(11, 1.5) of estimate_metrics.py:140-142: exp(-(x-5)^2 / (2 sigma^2)) in double, normalised as fp32
    Gauss11 g11;
    float* hg = g11.v;
    double gs[11];
    for (int i = 0; i < 11; ++i) gs[i] = exp(-static_cast<double>((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5));
    float fsum = 0.f;
    for (int i = 0; i < 11; ++i) { hg[i] = static_cast<float>(gs[i]); fsum += hg[i]; }
    for (int i = 0; i < 11; ++i) hg[i] = hg[i] / fsum;
    dim3 grid((w + 15) / 16, (h + 15) / 16, b * c);
    psnr_ssim_kernel<<<grid, 256, 0, stream>>>(pred, target, c, h, w, acc, g11);
    psnr_ssim_finish_kernel<<<(b + 127) / 128, 128, 0, stream>>>(acc, psnr, ssim, b, static_cast<double>(c) * h * w);
    CSBSR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
