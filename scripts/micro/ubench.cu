// Microbenchmarks behind the attention-kernel design decisions (run on the B200 box, results under profiles/):
//   mufu : per-SM throughput of ex2.approx in f32, f16x2 and bf16x2 form, and tanh.approx f32 / f16x2
//   mma  : issue-to-retire cost of back-to-back tcgen05.mma (M = 128, K = 16, bf16) for N = 16 .. 256, with the A operand in
//          shared memory (SS) or tensor memory (TS) and B K-major or MN-major -- operands are resident, so this is the
//          tensor pipe alone, with none of the L2 -> SM traffic a GEMM-level sweep also measures
//   mixed: does kind::f16 accept A = f16 (tensor memory) with B = bf16 (shared memory)?  known-answer check
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rdeic_b200/csrc -o scripts/micro/ubench scripts/micro/ubench.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include "sm100.cuh"

namespace rdeic {   // the two symbols sm100.cuh / common.cuh expect from api.cu
char* err_buf() { static char b[512]; return b; }
int set_error(const char* fmt, ...) { va_list a; va_start(a, fmt); vfprintf(stderr, fmt, a); va_end(a); fputc('\n', stderr); return 1; }
bool pdl_enabled() { return false; }
}
using namespace rdeic;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

// ---------------------------------------------------------------- MUFU
template <int kMode>
__global__ void __launch_bounds__(1024) mufu_kernel(float* out, int iters, float seed) {
    float acc = 0.f;
    if (kMode == 0) {                       // ex2.approx.ftz.f32
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = seed - 0.01f * i - 1e-3f * threadIdx.x;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += x[i];
    } else if (kMode == 1 || kMode == 2 || kMode == 4) {   // f16x2 / bf16x2 ex2, f16x2 tanh
        uint32_t x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0xb800b800u + i + threadIdx.x;       // ~ -0.5 (f16) / tiny negative (bf16)
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (kMode == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[i]));
                else if (kMode == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(x[i]));
                else asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(x[i]));
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += __uint_as_float(x[i]);
    } else if (kMode >= 5) {                // 5: cvt.rn.bf16x2.f32 alone; 6: 2 ex2 + 1 cvt (the softmax inner loop); 7: 2 ex2 + prmt pack
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = seed - 0.01f * i - 1e-3f * threadIdx.x;
        uint32_t sink = 0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                if (kMode != 5) {
                    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
                    asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i + 1]));
                }
                uint32_t pk;
                if (kMode == 7) asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(pk) : "r"(__float_as_uint(x[i])), "r"(__float_as_uint(x[i + 1])));
                else asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(x[i + 1]), "f"(x[i]));
                if (kMode == 5) { x[i] = __uint_as_float(pk); asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(x[i + 1]), "f"(x[i])); x[i + 1] = __uint_as_float(pk & 0xffff0000u); }
                else sink ^= pk;
            }
        }
        acc = __uint_as_float(sink);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += x[i];
    } else {                                // tanh.approx.f32
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = seed - 0.01f * i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += x[i];
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int kMode>
static int run_mufu(const char* name, int results_per_op) {
    float* d;
    CK(cudaMalloc(&d, 4));
    const int iters = 4096;
    mufu_kernel<kMode><<<148, 1024>>>(d, 16, -0.5f);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    mufu_kernel<kMode><<<148, 1024>>>(d, iters, -0.5f);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 148.0 * 1024 * iters * 8;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double per_sm_per_ns = ops / 148.0 / (ms * 1e6);
    printf("mufu %-22s: %.3f ms  %.2f instr-lanes/ns/SM = %.2f results/ns/SM  (at 1.9 GHz: %.1f results/clk/SM)\n", name, ms,
           per_sm_per_ns, per_sm_per_ns * results_per_op, per_sm_per_ns * results_per_op / 1.9);
    cudaFree(d);
    return 0;
}

// ---------------------------------------------------------------- MMA issue rate
// MN-major 128B-swizzled B (as attention_tc.cu's V tile: rows = K, 64 contiguous N elements per row)
__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)(16384 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// mode: 0 = SS, B K-major ; 1 = TS, B K-major ; 2 = TS, B MN-major (N <= 64)
__global__ void __launch_bounds__(128) mma_rate_kernel(int n, int mode, int count, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* s_a = smem;                 // 16 KB
    uint8_t* s_b = smem + 16384;         // 32 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const bool mn = mode == 2;
        const uint32_t idesc = make_idesc_mn(128, n, mn);
        const uint64_t da = make_smem_desc(smem_u32(s_a));
        const uint64_t db = mn ? desc_mn(smem_u32(s_b)) : make_smem_desc(smem_u32(s_b));
        const long long t0 = clock64();
        for (int i = 0; i < count; ++i) {
            const uint32_t d = tm + ((n <= 128) ? (i & 1) * 128 : 0);   // A (TS) sits at columns 480..487
            const int k = i & 3;
            if (mode == 0) umma_bf16(d, da + 2 * k, db + 2 * k, idesc, 1);
            else umma_bf16_ts(d, tm + 480 + 0 * k, mn ? db + (uint64_t)k * (2048 >> 4) : db + 2 * k, idesc, 1);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

// several issuing warps, each with its own accumulator: is the ~100-clock floor per issuing thread or per SM?
// kind: 0 = every issuer reads A from tensor memory (TS), 1 = from shared memory (SS); indep: 4 accumulators round-robin
__global__ void __launch_bounds__(256) mma_multi_kernel(int n, int issuers, int ss, int nacc, int count, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* s_a = smem;
    uint8_t* s_b = smem + 16384;
    __shared__ uint64_t bar[8];
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5;
    const long long t0 = clock64();
    if ((threadIdx.x & 31) == 0 && warp < issuers) {
        const uint32_t idesc = make_idesc_mn(128, n, false);
        const uint64_t da = make_smem_desc(smem_u32(s_a));
        const uint64_t db = make_smem_desc(smem_u32(s_b));
        for (int i = 0; i < count; ++i) {
            const uint32_t d = tm + (uint32_t)((warp * nacc + (i % nacc)) * n);
            const int k = i & 3;
            if (ss) umma_bf16(d, da + 2 * k, db + 2 * k, idesc, 1);
            else umma_bf16_ts(d, tm + 496, db + 2 * k, idesc, 1);
        }
        umma_commit(&bar[warp]);
        mbar_wait(&bar[warp], 0);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

static int run_mma_multi() {
    long long* d;
    CK(cudaMalloc(&d, 8));
    CK(cudaFuncSetAttribute(mma_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024));
    const int count = 4096;
    for (int ss = 0; ss < 2; ++ss)
        for (int n = 16; n <= 128; n *= 2)
            for (int nacc = 1; nacc <= 4; nacc *= 2)
                for (int issuers = 1; issuers <= 4; issuers *= 2) {
                    if (issuers * nacc * n > 480) continue;
                    mma_multi_kernel<<<148, 256, 16384 + 32768 + 1024>>>(n, issuers, ss, nacc, count, d);
                    CK(cudaDeviceSynchronize());
                    long long c;
                    CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
                    printf("mma-multi %s N=%3d accumulators/issuer=%d issuers=%d: %6.1f clk per MMA per SM (%6.1f per issuer)\n", ss ? "SS" : "TS", n, nacc,
                           issuers, (double)c / count / issuers, (double)c / count);
                }
    cudaFree(d);
    return 0;
}

// same loop as mma_rate_kernel (SS), but the issuing thread is chosen by elect.sync inside a warp-uniform branch (CUTLASS's
// elect_one_sync) instead of `threadIdx.x == 0`: does ptxas then drop the ELECT / BRA.U.ANY wrapper around every UTCHMMA?
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__global__ void __launch_bounds__(128) mma_elect_kernel(int n, int count, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* s_a = smem;
    uint8_t* s_b = smem + 16384;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x < 32) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_mn(128, n, false);
            const uint64_t da = make_smem_desc(smem_u32(s_a));
            const uint64_t db = make_smem_desc(smem_u32(s_b));
            const long long t0 = clock64();
            for (int i = 0; i < count; ++i) {
                const uint32_t d = tm + ((n <= 128) ? (i & 1) * 128 : 0);
                const int k = i & 3;
                umma_bf16(d, da + 2 * k, db + 2 * k, idesc, 1);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t1 = clock64();
            if (blockIdx.x == 0) cycles[0] = t1 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

static int run_mma_elect() {
    long long* d;
    CK(cudaMalloc(&d, 8));
    CK(cudaFuncSetAttribute(mma_elect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024));
    const int count = 4096;
    const int ns[6] = {16, 64, 128, 160, 192, 256};
    for (int i = 0; i < 6; ++i) {
        mma_elect_kernel<<<148, 128, 16384 + 32768 + 1024>>>(ns[i], count, d);
        CK(cudaDeviceSynchronize());
        long long c;
        CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
        printf("mma-elect SS N=%3d: %6.1f clk per MMA (ideal N/2 = %d)\n", ns[i], (double)c / count, ns[i] / 2);
    }
    cudaFree(d);
    return 0;
}

static int run_mma() {
    long long* d;
    CK(cudaMalloc(&d, 8));
    CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024));
    const int count = 4096;
    const char* names[3] = {"SS  B K-major ", "TS  B K-major ", "TS  B MN-major"};
    for (int mode = 0; mode < 3; ++mode) {
        const int ns[6] = {16, 32, 64, 128, 160, 256};
        for (int i = 0; i < 6; ++i) {
            const int n = ns[i];
            if (mode == 2 && n > 64) continue;
            for (int grid = 1; grid <= 148; grid += 147) {
                mma_rate_kernel<<<grid, 128, 16384 + 32768 + 1024>>>(n, mode, count, d);
                CK(cudaDeviceSynchronize());
                long long c;
                CK(cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost));
                printf("mma %s N=%3d grid=%3d: %6.1f clk per MMA (ideal N/2 = %d)\n", names[mode], n, grid, (double)c / count, n / 2);
            }
        }
    }
    cudaFree(d);
    return 0;
}

// ---------------------------------------------------------------- mixed f16 (TMEM A) x bf16 (smem B)
__global__ void __launch_bounds__(128) mixed_kernel(float* out, int a_is_f16) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x40004000u;   // bf16 2.0
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc<128>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    uint32_t a[16];
    for (int i = 0; i < 16; ++i) a[i] = a_is_f16 ? 0x3c003c00u /* f16 1.0 pairs */ : 0x3f803f80u /* bf16 1.0 pairs */;
    tmem_st16(tm + lane_addr + 64, a);       // 16 columns = 32 K elements of A
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        // a_format bits [7,10): 0 = f16, 1 = bf16 ; b_format bits [10,13): 1 = bf16 ; D fp32 ; M = 128, N = 64
        uint32_t idesc = (1u << 4) | ((a_is_f16 ? 0u : 1u) << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        umma_bf16_ts(tm, tm + 64, make_smem_desc(smem_u32(smem)), idesc, 0);
        umma_commit(&bar);
        mbar_wait(&bar, 0);
    }
    __syncthreads();
    tc_fence_after();
    uint32_t r[16];
    tmem_ld16(tm + lane_addr, r);
    tmem_ld_wait();
    if ((threadIdx.x & 31) == 0) out[warp] = __uint_as_float(r[3]);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<128>(tm); }
}

static int run_mixed() {
    float* d;
    CK(cudaMalloc(&d, 16));
    CK(cudaFuncSetAttribute(mixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024));
    for (int f16 = 0; f16 < 2; ++f16) {
        mixed_kernel<<<1, 128, 16384 + 1024>>>(d, f16);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mixed a_is_f16=%d: CUDA error %s\n", f16, cudaGetErrorString(e)); return 1; }
        float h[4];
        CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
        printf("mixed A=%s(TMEM, 1.0) x B=bf16(smem, 2.0), K=16: D = %g %g %g %g (expected 32)\n", f16 ? "f16" : "bf16", h[0], h[1], h[2], h[3]);
    }
    cudaFree(d);
    return 0;
}

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "all";
    if (!strcmp(what, "mufu") || !strcmp(what, "all")) {
        if (run_mufu<0>("ex2.approx.ftz.f32", 1)) return 1;
        if (run_mufu<1>("ex2.approx.f16x2", 2)) return 1;
        if (run_mufu<2>("ex2.approx.ftz.bf16x2", 2)) return 1;
        if (run_mufu<3>("tanh.approx.f32", 1)) return 1;
        if (run_mufu<4>("tanh.approx.f16x2", 2)) return 1;
        if (run_mufu<5>("cvt.rn.bf16x2.f32 (x8/it)", 1)) return 1;
        if (run_mufu<6>("2 ex2 + cvt.bf16x2", 1)) return 1;
        if (run_mufu<7>("2 ex2 + prmt", 1)) return 1;
    }
    if (!strcmp(what, "mma") || !strcmp(what, "all")) if (run_mma()) return 1;
    if (!strcmp(what, "mmaw") || !strcmp(what, "all")) if (run_mma_multi()) return 1;
    if (!strcmp(what, "elect") || !strcmp(what, "all")) if (run_mma_elect()) return 1;
    if (!strcmp(what, "mixed")) if (run_mixed()) return 1;
    return 0;
}
