// Microbenchmark (round 2): how fast does one CTA retire tcgen05.mma.cta_group::1.kind::f16 instructions of M=128, K=16
// as a function of N, of the number of independent TMEM accumulators the issue order rotates over, and of the shared
// memory layout (128B-swizzled K-major rows as in conv_igemm.cu, or the un-swizzled "plane" layout [k/8][row][8] in which
// a row shift is a plain +16 B on the descriptor start address).  Also checks the numerics of the plane layout with
// shifted start addresses, which is what a shared-memory-resident 3x3 conv chain needs.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/ubench_mma scripts/ubench_mma.cu
//   ./gpurun_out/ubench_mma
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "../csbsr_b200/csrc/tc_ptx.cuh"

using namespace csbsr;

// The issue loop is fully unrolled over 8 MMAs with compile-time accumulator / shift selection: a first version with
// run-time modulo arithmetic in the loop measured its own scalar overhead (153 cycles per MMA whatever N).
template <int N, int NACC, int LAYOUT, int DISTINCT>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 2; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16(0.01f);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_ptr, 512);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (8u << 24);
    constexpr uint32_t lbo = 256u * 16u;
    constexpr uint32_t desc_hi = LAYOUT == 0 ? ((1024u >> 4) | (1u << 14) | (2u << 29)) : ((128u >> 4) | (1u << 14));
    constexpr uint32_t kstep = LAYOUT == 0 ? 2u : 2u * (lbo >> 4);
    constexpr uint32_t lbo_field = LAYOUT == 0 ? 1u : (lbo >> 4);
    const uint32_t a_lo = ((smem_u32(smem) & 0x3FFFFu) >> 4) | (lbo_field << 16);
    const uint32_t b_lo = ((smem_u32(smem + 64 * 1024) & 0x3FFFFu) >> 4) | (lbo_field << 16);
    if (warp == 1) {
        if (elect_one_sync()) {
            uint32_t phase = 0;
            for (int rep = 0; rep < 4; ++rep) {
                const long long t0 = clock64();
#pragma unroll 1
                for (int it = 0; it < 64; ++it) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        constexpr int dummy = 0; (void)dummy;
                        const uint32_t acc = static_cast<uint32_t>((j % NACC) * N);
                        const uint32_t sh = DISTINCT ? static_cast<uint32_t>(j * (LAYOUT ? 1 : 64)) : 0u;
                        umma_bf16_lohi(acc, a_lo + sh + (j & 1) * kstep, b_lo + (j & 1) * kstep, desc_hi, idesc, 1u);
                    }
                }
                const long long t1 = clock64();
                umma_commit(&bar);
                while (!mbar_try_wait(&bar, phase)) {}
                phase ^= 1u;
                const long long t2 = clock64();
                if (rep == 3) {
                    out[blockIdx.x * 2 + 0] = t1 - t0;
                    out[blockIdx.x * 2 + 1] = t2 - t0;
                }
            }
        }
        __syncwarp();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc(0, 512);
    }
}

template <int N, int NACC, int LAYOUT, int DISTINCT>
static void run_rate(long long* d_out) {
    cudaFuncSetAttribute(mma_rate_kernel<N, NACC, LAYOUT, DISTINCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int grid : {1, 148}) {
        mma_rate_kernel<N, NACC, LAYOUT, DISTINCT><<<grid, 128, 160 * 1024>>>(d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
        long long h[2];
        cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%d %3d %d %d %3d  %7.1f  %7.1f\n", LAYOUT, N, NACC, DISTINCT, grid, h[0] / 512.0, h[1] / 512.0);
    }
}
template <int N, int LAYOUT>
static void run_n(long long* d_out) {
    run_rate<N, 1, LAYOUT, 0>(d_out);
    run_rate<N, 1, LAYOUT, 1>(d_out);
    run_rate<N, 2, LAYOUT, 1>(d_out);
    if (N * 4 <= 512) run_rate<N, (N * 4 <= 512 ? 4 : 1), LAYOUT, 1>(d_out);
    if (N * 8 <= 512) run_rate<N, (N * 8 <= 512 ? 8 : 1), LAYOUT, 1>(d_out);
}

// ---------------------------------------------------------------- numerics of the plane layout with shifted A start
// A "image" of P pixels x 32 channels lives as 4 planes [P][8]; weights for T taps as [T][4 planes][N rows][8].
// out[m][n] = sum_t sum_c A[m + shift_t][c] * W[t][n][c]  for m in [0,128)
__global__ void __launch_bounds__(128, 1) plane_conv_kernel(const __nv_bfloat16* a_g, const __nv_bfloat16* w_g, float* out, int P,
                                                             int N, int T, const int* shifts) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __nv_bfloat16* sa = reinterpret_cast<__nv_bfloat16*>(smem);                 // [4][P][8]
    __nv_bfloat16* sw = reinterpret_cast<__nv_bfloat16*>(smem + 4 * P * 16);    // [T][4][N][8]
    for (int i = threadIdx.x; i < P * 32; i += blockDim.x) {
        const int p = i / 32, ch = i % 32;
        sa[((ch >> 3) * P + p) * 8 + (ch & 7)] = a_g[i];
    }
    for (int i = threadIdx.x; i < T * N * 32; i += blockDim.x) {
        const int t = i / (N * 32), n = (i / 32) % N, ch = i % 32;
        sw[((t * 4 + (ch >> 3)) * N + n) * 8 + (ch & 7)] = w_g[i];
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(&tmem_ptr, 512);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (threadIdx.x == 32) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (8u << 24);
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);
        const uint32_t lbo_a = static_cast<uint32_t>(P) * 16u, lbo_w = static_cast<uint32_t>(N) * 16u;
        const uint32_t a0 = (smem_u32(sa) & 0x3FFFFu) >> 4, w0 = (smem_u32(sw) & 0x3FFFFu) >> 4;
        uint32_t acc = 0;
        for (int t = 0; t < T; ++t)
            for (int ks = 0; ks < 2; ++ks) {          // K = 32 channels = 2 MMAs of K=16 (2 planes each)
                const uint32_t alo = (a0 + static_cast<uint32_t>(shifts[t]) + ks * 2 * (lbo_a >> 4)) | ((lbo_a >> 4) << 16);
                const uint32_t wlo = (w0 + static_cast<uint32_t>(t * 4 + ks * 2) * (lbo_w >> 4)) | ((lbo_w >> 4) << 16);
                umma_bf16_lohi(0u, alo, wlo, desc_hi, idesc, acc);
                acc = 1u;
            }
        umma_commit(&bar);
    }
    while (!mbar_try_wait(&bar, 0)) {}
    tcgen05_fence_after();
    for (int u = 0; u < N / 16; ++u) {
        uint32_t v[16];
        tmem_ld16((static_cast<uint32_t>(warp * 32) << 16) + u * 16, v);
        tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * N + u * 16 + i] = __uint_as_float(v[i]);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
        tcgen05_fence_after();
        tmem_dealloc(0, 512);
    }
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
    cudaFuncSetAttribute(plane_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long* d_out;
    cudaMalloc(&d_out, 148 * 2 * sizeof(long long));
    printf("layout n nacc distinct grid  issue_cyc/mma  retire_cyc/mma\n");
    run_n<32, 0>(d_out); run_n<64, 0>(d_out); run_n<128, 0>(d_out); run_n<256, 0>(d_out);
    run_n<32, 1>(d_out); run_n<64, 1>(d_out); run_n<128, 1>(d_out); run_n<256, 1>(d_out);
    // numerics
    const int P = 512, T = 9, W = 20;
    for (int N : {32, 64}) {
        std::vector<__nv_bfloat16> a(P * 32), w(T * N * 32);
        std::vector<float> af(P * 32), wf(T * N * 32);
        srand(1);
        for (size_t i = 0; i < a.size(); ++i) { af[i] = bf((rand() % 2001 - 1000) / 1000.f); a[i] = __float2bfloat16(af[i]); }
        for (size_t i = 0; i < w.size(); ++i) { wf[i] = bf((rand() % 2001 - 1000) / 4000.f); w[i] = __float2bfloat16(wf[i]); }
        int shifts[9];
        for (int t = 0; t < 9; ++t) shifts[t] = (W + 1) + (t / 3 - 1) * W + (t % 3 - 1);   // centre pixel offset W+1
        __nv_bfloat16 *da, *dw;
        float* dout;
        int* dsh;
        cudaMalloc(&da, a.size() * 2); cudaMalloc(&dw, w.size() * 2); cudaMalloc(&dout, 128 * N * 4); cudaMalloc(&dsh, sizeof(shifts));
        cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(dsh, shifts, sizeof(shifts), cudaMemcpyHostToDevice);
        plane_conv_kernel<<<1, 128, 100 * 1024>>>(da, dw, dout, P, N, T, dsh);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("plane_conv error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> o(128 * N);
        cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double r = 0;
                for (int t = 0; t < T; ++t)
                    for (int ch = 0; ch < 32; ++ch) r += (double)af[(m + shifts[t]) * 32 + ch] * wf[(t * N + n) * 32 + ch];
                maxerr = fmax(maxerr, fabs(r - o[m * N + n]));
            }
        printf("plane-layout shifted conv N=%d: max abs err %.3e (%s)\n", N, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
    }
    return 0;
}
