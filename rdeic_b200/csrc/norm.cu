// GroupNorm(+SiLU) and LayerNorm over pixel-major (NHWC) bf16 activations, fp32 statistics.
// Both are HBM-bound: every element is read twice (stats pass hits L2 on the second read
// for UNet-sized tensors) and written once, in 16-byte vectors.
//
// Reference semantics: ldm/modules/diffusionmodules/util.py:224 GroupNorm32 (eps 1e-5, fp32
// compute), ldm/modules/diffusionmodules/model.py:48 Normalize (eps 1e-6),
// ldm/modules/attention.py:96 Normalize (eps 1e-6), model/rdeic.py:483 GroupNorm_leq32,
// ldm/modules/attention.py:273-275 nn.LayerNorm (eps 1e-5).
#include "common.cuh"
#include "../../include/rdeic_b200.h"

namespace rdeic {

constexpr int kGnThreads = 256;
constexpr int kGnMaxChunks = 256;
constexpr int kGnMaxGroups = 32;
constexpr int kGnMaxC = 2560;  // widest GN input on the path: concat 1280 + 1280

static inline int gn_num_chunks(int B, int64_t HW) {
    int64_t want = (8 * kNumSMs + B - 1) / B;   // ~8 CTAs per SM over the whole batch
    int64_t max_by_rows = HW / 32 > 0 ? HW / 32 : 1;
    if (want > max_by_rows) want = max_by_rows;
    if (want > kGnMaxChunks) want = kGnMaxChunks;
    if (want < 1) want = 1;
    return (int)want;
}

// 8 consecutive channels of one pixel as fp32, from a bf16 (16 B) or fp32 (32 B) tensor.
// `vec_index` counts 8-channel vectors.
template <bool kF32>
__device__ __forceinline__ void load8(const void* base, int64_t vec_index, float* f) {
    if (kF32) {
        const uint4 a = ld_stream_u4(reinterpret_cast<const uint4*>(base) + 2 * vec_index);
        const uint4 b = ld_stream_u4(reinterpret_cast<const uint4*>(base) + 2 * vec_index + 1);
        f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y);
        f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
        f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y);
        f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
    } else {
        const uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(base) + vec_index);
        unpack_bf16x2(v.x, f[0], f[1]); unpack_bf16x2(v.y, f[2], f[3]);
        unpack_bf16x2(v.z, f[4], f[5]); unpack_bf16x2(v.w, f[6], f[7]);
    }
}

// partial[(b*nchunk + chunk)*G + g] = (sum, sumsq) over the chunk's pixels and the group's
// channels, reduced in a fixed order (deterministic run to run).
template <bool kF32>
__global__ void __launch_bounds__(kGnThreads)
gn_stats_kernel(const void* __restrict__ x1, int C1, const void* __restrict__ x2, int C2,
                int64_t HW, int G, float2* __restrict__ partial) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_part[kGnThreads * 16];   // [row][lane][8 sum | 8 sq]
    __shared__ float s_csum[kGnMaxC];
    __shared__ float s_csq[kGnMaxC];
    const int C = C1 + C2;
    const int VL = C >> 3, VL1 = C1 >> 3, VL2 = C2 >> 3;
    const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
    const int64_t per = (HW + nchunk - 1) / nchunk;
    const int64_t p0 = chunk * per;
    const int64_t p1 = (p0 + per < HW) ? p0 + per : HW;
    const int lanes_per_pass = VL < kGnThreads ? VL : kGnThreads;
    const int rows_per_iter = kGnThreads / lanes_per_pass;
    const int tid = threadIdx.x;
    const int row = tid / lanes_per_pass, lane_in = tid - row * lanes_per_pass;

    for (int lane_base = 0; lane_base < VL; lane_base += lanes_per_pass) {
        const int l = lane_base + lane_in;
        float s[8], q[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = q[k] = 0.f;
        if (row < rows_per_iter && l < VL) {
            const bool first = l < VL1;
            const void* src = first ? x1 : x2;
            const int64_t stride = first ? VL1 : VL2;
            const int64_t off = (int64_t)b * HW * stride + (first ? l : l - VL1);
            int64_t p = p0 + row;
            // four independent 16/32-byte loads in flight per thread (memory-level parallelism)
            for (; p + 3 * (int64_t)rows_per_iter < p1; p += 4 * (int64_t)rows_per_iter) {
                float f[4][8];
#pragma unroll
                for (int u = 0; u < 4; ++u) load8<kF32>(src, off + (p + u * (int64_t)rows_per_iter) * stride, f[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        s[k] += f[u][k];
                        q[k] = fmaf(f[u][k], f[u][k], q[k]);
                    }
                }
            }
            for (; p < p1; p += rows_per_iter) {
                float f[8];
                load8<kF32>(src, off + p * stride, f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    s[k] += f[k];
                    q[k] = fmaf(f[k], f[k], q[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s_part[tid * 16 + k] = s[k];
            s_part[tid * 16 + 8 + k] = q[k];
        }
        __syncthreads();
        // reduce over rows: one thread per (lane, channel-in-vector)
        for (int i = tid; i < lanes_per_pass * 8; i += kGnThreads) {
            const int li = i >> 3, k = i & 7;
            const int lg = lane_base + li;
            if (lg < VL) {
                float a = 0.f, c = 0.f;
                for (int r = 0; r < rows_per_iter; ++r) {
                    a += s_part[(r * lanes_per_pass + li) * 16 + k];
                    c += s_part[(r * lanes_per_pass + li) * 16 + 8 + k];
                }
                s_csum[lg * 8 + k] = a;
                s_csq[lg * 8 + k] = c;
            }
        }
        __syncthreads();
    }
    const int cg = C / G;
    if (tid < G) {
        float a = 0.f, c = 0.f;
        for (int k = 0; k < cg; ++k) {
            a += s_csum[tid * cg + k];
            c += s_csq[tid * cg + k];
        }
        partial[((int64_t)b * nchunk + chunk) * G + tid] = make_float2(a, c);
    }
}

// Statistics handed over by the producing GEMMs: st[slab*Csrc + c] = (sum, sumsq) of channel c over
// 32 consecutive pixels.  Same grid and output layout as gn_stats_kernel (one CTA per pixel chunk
// and sample -> partial[(b*nchunk + chunk)*G + g]), but it reads M/32 x C pairs instead of the
// tensor: rows of the slab table are contiguous over channels, so the loads are fully coalesced.
__global__ void __launch_bounds__(kGnThreads)
gn_fold_stats_kernel(const float2* __restrict__ st1, int C1, const float2* __restrict__ st2, int C2,
                     int64_t slabs_per_sample, int G, float2* __restrict__ partial) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_csum[kGnMaxC];
    __shared__ float s_csq[kGnMaxC];
    const int C = C1 + C2;
    const int b = blockIdx.y, chunk = blockIdx.x, nchunk = gridDim.x;
    const int64_t per = (slabs_per_sample + nchunk - 1) / nchunk;
    const int64_t s0 = chunk * per;
    const int64_t s1 = (s0 + per < slabs_per_sample) ? s0 + per : slabs_per_sample;
    for (int c = threadIdx.x; c < C; c += kGnThreads) {
        const bool first = c < C1;
        const float2* src = first ? st1 + c : st2 + (c - C1);
        const int64_t stride = first ? C1 : C2;
        const int64_t base = (int64_t)b * slabs_per_sample;
        float a0 = 0.f, q0 = 0.f, a1 = 0.f, q1 = 0.f;
        int64_t sl = s0;
        for (; sl + 1 < s1; sl += 2) {               // two independent loads in flight
            const float2 u = src[(base + sl) * stride], v = src[(base + sl + 1) * stride];
            a0 += u.x; q0 += u.y; a1 += v.x; q1 += v.y;
        }
        if (sl < s1) { const float2 u = src[(base + sl) * stride]; a0 += u.x; q0 += u.y; }
        s_csum[c] = a0 + a1;
        s_csq[c] = q0 + q1;
    }
    __syncthreads();
    const int cg = C / G;
    if (threadIdx.x < G) {
        float a = 0.f, c = 0.f;
        for (int k = 0; k < cg; ++k) {
            a += s_csum[threadIdx.x * cg + k];
            c += s_csq[threadIdx.x * cg + k];
        }
        partial[((int64_t)b * nchunk + chunk) * G + threadIdx.x] = make_float2(a, c);
    }
}

template <bool kF32, bool kOutF32 = false>
__global__ void __launch_bounds__(kGnThreads)
gn_apply_kernel(const void* __restrict__ x1, int C1, const void* __restrict__ x2, int C2,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                uint4* __restrict__ out, int64_t HW, int G, float eps, int silu,
                const float2* __restrict__ partial, int nchunk, FastDiv div_vl) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_scale[kGnMaxC];
    __shared__ float s_shift[kGnMaxC];
    __shared__ float s_mean[kGnMaxGroups];
    __shared__ float s_rstd[kGnMaxGroups];
    const int C = C1 + C2;
    const int VL = C >> 3, VL1 = C1 >> 3, VL2 = C2 >> 3;
    const int b = blockIdx.y;
    const int cg = C / G;
    {   // finalise the statistics: 8 threads per group split the chunk partials (fixed order ->
        // deterministic), fp64 combine; a single thread per group would serialise nchunk L2 round trips
        const int g = threadIdx.x >> 3, part = threadIdx.x & 7;
        double a = 0.0, c = 0.0;
        if (g < G) {
            for (int k = part; k < nchunk; k += 8) {
                const float2 v = partial[((int64_t)b * nchunk + k) * G + g];
                a += (double)v.x;
                c += (double)v.y;
            }
        }
#pragma unroll
        for (int o = 4; o; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (g < G && part == 0) {
            const double n = (double)HW * cg;
            const double mean = a / n;
            double var = c / n - mean * mean;   // biased variance, as nn.GroupNorm
            if (var < 0.0) var = 0.0;
            s_mean[g] = (float)mean;
            s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kGnThreads) {
        const int g = c / cg;
        const float sc = gamma[c] * s_rstd[g];
        s_scale[c] = sc;
        s_shift[c] = beta[c] - s_mean[g] * sc;
    }
    __syncthreads();
    // Each thread keeps ONE 8-channel lane for the whole pass, so its scale/shift live in registers
    // (reading them from shared memory per element made the kernel smem-bandwidth bound: 4 LDS.128
    // per 16 bytes of payload) and the loop has no index division.
    const int64_t o1 = (int64_t)b * HW * VL1, o2 = (int64_t)b * HW * VL2;
    uint4* bo = out + (int64_t)b * HW * VL * (kOutF32 ? 2 : 1);
    const int nchunk_a = gridDim.x;
    const int64_t per = (HW + nchunk_a - 1) / nchunk_a;
    const int64_t p0 = blockIdx.x * per;
    const int64_t p1 = (p0 + per < HW) ? p0 + per : HW;
    const int lanes_per_pass = VL < kGnThreads ? VL : kGnThreads;
    const int rows_per_iter = kGnThreads / lanes_per_pass;
    const int row = threadIdx.x / lanes_per_pass, lane_in = threadIdx.x - row * lanes_per_pass;
    for (int lane_base = 0; lane_base < VL; lane_base += lanes_per_pass) {
        const int l = lane_base + lane_in;
        if (row >= rows_per_iter || l >= VL) continue;
        float sc[8], sh[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { sc[k] = s_scale[l * 8 + k]; sh[k] = s_shift[l * 8 + k]; }
        const bool first = l < VL1;
        const void* src = first ? x1 : x2;
        const int64_t stride = first ? VL1 : VL2;
        const int64_t off = (first ? o1 + l : o2 + (l - VL1));
        auto finish = [&](float* f, int64_t p) {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sh[k]);
            if (kOutF32) {               // fp32 kernel mode: exact exp / division, fp32 result
                if (silu) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] = f[k] / (1.0f + expf(-f[k]));
                }
                uint4 o0, o1;
                o0.x = __float_as_uint(f[0]); o0.y = __float_as_uint(f[1]); o0.z = __float_as_uint(f[2]); o0.w = __float_as_uint(f[3]);
                o1.x = __float_as_uint(f[4]); o1.y = __float_as_uint(f[5]); o1.z = __float_as_uint(f[6]); o1.w = __float_as_uint(f[7]);
                st_stream_u4(bo + (p * VL + l) * 2, o0);
                st_stream_u4(bo + (p * VL + l) * 2 + 1, o1);
                return;
            }
            if (silu) {
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = silu_f(f[k]);
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
            st_stream_u4(bo + p * VL + l, o);
        };
        int64_t p = p0 + row;
        for (; p + 3 * (int64_t)rows_per_iter < p1; p += 4 * (int64_t)rows_per_iter) {
            float f[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) load8<kF32>(src, off + (p + u * (int64_t)rows_per_iter) * stride, f[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) finish(f[u], p + u * (int64_t)rows_per_iter);
        }
        for (; p < p1; p += rows_per_iter) {
            float f[8];
            load8<kF32>(src, off + p * stride, f);
            finish(f, p);
        }
    }
}

// LayerNorm: one warp per row, row held in registers (C <= 8*32*kLnMaxVec); kLnMaxVec is a template
// parameter so narrow rows keep few registers live and the SM holds enough warps (= bytes in
// flight) to cover HBM latency.
constexpr int kLnMaxVecLimit = 5;
constexpr int kLnWarps = 8;

template <bool kF32, int kLnMaxVec, bool kOutF32 = false>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ beta, uint4* __restrict__ out, int64_t rows, int C,
                 float eps) {
    pdl_trigger();
    pdl_wait();
    const int VL = C >> 3;
    const int lane = threadIdx.x & 31;
    // grid-stride over rows: at most 8 resident CTAs per SM walk the tensor (32 768 short-lived CTAs
    // spent more time being scheduled than loading)
    for (int64_t row = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5); row < rows;
         row += (int64_t)gridDim.x * kLnWarps) {
    float f[kLnMaxVec][8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxVec; ++j) {
        const int l = lane + 32 * j;
        if (l < VL) {
            load8<kF32>(x, row * VL + l, f[j]);
#pragma unroll
            for (int k = 0; k < 8; ++k) sum += f[j][k];
        }
    }
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxVec; ++j) {
        const int l = lane + 32 * j;
        if (l < VL) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float d = f[j][k] - mean;
                sq = fmaf(d, d, sq);
            }
        }
    }
    for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)C + eps);
#pragma unroll
    for (int j = 0; j < kLnMaxVec; ++j) {
        const int l = lane + 32 * j;
        if (l < VL) {
            const float4 g0 = *reinterpret_cast<const float4*>(gamma + l * 8);
            const float4 g1 = *reinterpret_cast<const float4*>(gamma + l * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(beta + l * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(beta + l * 8 + 4);
            float y[8];
            y[0] = (f[j][0] - mean) * rstd * g0.x + b0.x;
            y[1] = (f[j][1] - mean) * rstd * g0.y + b0.y;
            y[2] = (f[j][2] - mean) * rstd * g0.z + b0.z;
            y[3] = (f[j][3] - mean) * rstd * g0.w + b0.w;
            y[4] = (f[j][4] - mean) * rstd * g1.x + b1.x;
            y[5] = (f[j][5] - mean) * rstd * g1.y + b1.y;
            y[6] = (f[j][6] - mean) * rstd * g1.z + b1.z;
            y[7] = (f[j][7] - mean) * rstd * g1.w + b1.w;
            if (kOutF32) {
                uint4 o0, o1;
                o0.x = __float_as_uint(y[0]); o0.y = __float_as_uint(y[1]); o0.z = __float_as_uint(y[2]); o0.w = __float_as_uint(y[3]);
                o1.x = __float_as_uint(y[4]); o1.y = __float_as_uint(y[5]); o1.z = __float_as_uint(y[6]); o1.w = __float_as_uint(y[7]);
                st_stream_u4(out + (row * VL + l) * 2, o0);
                st_stream_u4(out + (row * VL + l) * 2 + 1, o1);
                continue;
            }
            uint4 o;
            o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
            o.z = pack_bf16x2(y[4], y[5]); o.w = pack_bf16x2(y[6], y[7]);
            st_stream_u4(out + row * VL + l, o);
        }
    }
    }
}

}  // namespace rdeic

using namespace rdeic;

extern "C" {

int64_t rdeic_groupnorm_workspace_bytes(int B, int64_t HW, int C) {
    (void)HW; (void)C;
    return (int64_t)B * kGnMaxChunks * kGnMaxGroups * (int64_t)sizeof(float2);
}

int rdeic_groupnorm_nhwc(const void* x1, int C1, const void* x2, int C2, int in_is_f32,
                         const float* gamma, const float* beta, void* out, int B, int64_t HW,
                         int groups, float eps, int silu, void* workspace,
                         rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(x1 && gamma && beta && out && workspace, "rdeic_groupnorm_nhwc: null pointer");
    RDEIC_CHECK_ARG(C2 == 0 || x2, "rdeic_groupnorm_nhwc: C2 > 0 needs x2");
    if (C2 == 0) x2 = nullptr;
    const int C = C1 + C2;
    RDEIC_CHECK_ARG(B > 0 && HW > 0, "rdeic_groupnorm_nhwc: empty tensor");
    RDEIC_CHECK_ARG(B <= 65535, "rdeic_groupnorm_nhwc: B too large");
    RDEIC_CHECK_ARG(C1 > 0 && C1 % 8 == 0 && C2 >= 0 && C2 % 8 == 0,
                    "rdeic_groupnorm_nhwc: channel counts (%d, %d) must be multiples of 8", C1, C2);
    RDEIC_CHECK_ARG(C <= kGnMaxC, "rdeic_groupnorm_nhwc: C=%d exceeds %d", C, kGnMaxC);
    RDEIC_CHECK_ARG(groups > 0 && groups <= kGnMaxGroups && C % groups == 0,
                    "rdeic_groupnorm_nhwc: groups=%d invalid for C=%d", groups, C);
    RDEIC_CHECK_ARG(((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)out) % 16 == 0,
                    "rdeic_groupnorm_nhwc: tensors must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    const int nchunk = gn_num_chunks(B, HW);
    if (in_is_f32)
        launch_k(gn_stats_kernel<true>, dim3(nchunk, B), kGnThreads, 0, s, x1, C1, x2, C2, HW, groups, (float2*)workspace);
    else
        launch_k(gn_stats_kernel<false>, dim3(nchunk, B), kGnThreads, 0, s, x1, C1, x2, C2, HW, groups, (float2*)workspace);
    RDEIC_LAUNCH_CHECK();
    const FastDiv div_vl((uint32_t)(C / 8));
    int64_t blocks = nchunk;          // same pixel chunking as the statistics pass
    if (in_is_f32)
        launch_k(gn_apply_kernel<true>, dim3((unsigned)blocks, B), kGnThreads, 0, s, 
            x1, C1, x2, C2, gamma, beta, (uint4*)out, HW, groups, eps, silu, (const float2*)workspace, nchunk, div_vl);
    else
        launch_k(gn_apply_kernel<false>, dim3((unsigned)blocks, B), kGnThreads, 0, s, 
            x1, C1, x2, C2, gamma, beta, (uint4*)out, HW, groups, eps, silu, (const float2*)workspace, nchunk, div_vl);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_groupnorm_from_stats(const void* x1, int C1, const float* stats1, const void* x2, int C2,
                               const float* stats2, int in_is_f32, const float* gamma,
                               const float* beta, void* out, int B, int64_t HW, int groups,
                               float eps, int silu, void* workspace, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(x1 && stats1 && gamma && beta && out && workspace, "rdeic_groupnorm_from_stats: null pointer");
    RDEIC_CHECK_ARG(C2 == 0 || (x2 && stats2), "rdeic_groupnorm_from_stats: C2 > 0 needs x2 and stats2");
    if (C2 == 0) { x2 = nullptr; stats2 = nullptr; }
    const int C = C1 + C2;
    RDEIC_CHECK_ARG(B > 0 && B <= 65535 && HW > 0 && HW % 32 == 0,
                    "rdeic_groupnorm_from_stats: HW must be a positive multiple of 32");
    RDEIC_CHECK_ARG(C1 > 0 && C1 % 8 == 0 && C2 >= 0 && C2 % 8 == 0 && C <= kGnMaxC,
                    "rdeic_groupnorm_from_stats: bad channel counts (%d, %d)", C1, C2);
    RDEIC_CHECK_ARG(groups > 0 && groups <= kGnMaxGroups && C % groups == 0,
                    "rdeic_groupnorm_from_stats: groups=%d invalid for C=%d", groups, C);
    RDEIC_CHECK_ARG(((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)out) % 16 == 0,
                    "rdeic_groupnorm_from_stats: tensors must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    int nfold = gn_num_chunks(B, HW);
    if (nfold > HW / 32) nfold = (int)(HW / 32);
    launch_k(gn_fold_stats_kernel, dim3(nfold, B), kGnThreads, 0, s, (const float2*)stats1, C1,
             (const float2*)stats2, C2, HW / 32, groups, (float2*)workspace);
    RDEIC_LAUNCH_CHECK();
    const FastDiv div_vl((uint32_t)(C / 8));
    const int nchunk = gn_num_chunks(B, HW);
    if (in_is_f32)
        launch_k(gn_apply_kernel<true>, dim3((unsigned)nchunk, B), kGnThreads, 0, s,
            x1, C1, x2, C2, gamma, beta, (uint4*)out, HW, groups, eps, silu, (const float2*)workspace, nfold, div_vl);
    else
        launch_k(gn_apply_kernel<false>, dim3((unsigned)nchunk, B), kGnThreads, 0, s,
            x1, C1, x2, C2, gamma, beta, (uint4*)out, HW, groups, eps, silu, (const float2*)workspace, nfold, div_vl);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_groupnorm_nhwc_f32(const float* x1, int C1, const float* x2, int C2, const float* gamma,
                             const float* beta, float* out, int B, int64_t HW, int groups, float eps,
                             int silu, void* workspace, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(x1 && gamma && beta && out && workspace, "rdeic_groupnorm_nhwc_f32: null pointer");
    RDEIC_CHECK_ARG(C2 == 0 || x2, "rdeic_groupnorm_nhwc_f32: C2 > 0 needs x2");
    if (C2 == 0) x2 = nullptr;
    const int C = C1 + C2;
    RDEIC_CHECK_ARG(B > 0 && B <= 65535 && HW > 0, "rdeic_groupnorm_nhwc_f32: bad tensor dims");
    RDEIC_CHECK_ARG(C1 > 0 && C1 % 8 == 0 && C2 >= 0 && C2 % 8 == 0 && C <= kGnMaxC,
                    "rdeic_groupnorm_nhwc_f32: bad channel counts (%d, %d)", C1, C2);
    RDEIC_CHECK_ARG(groups > 0 && groups <= kGnMaxGroups && C % groups == 0,
                    "rdeic_groupnorm_nhwc_f32: groups=%d invalid for C=%d", groups, C);
    RDEIC_CHECK_ARG(((uintptr_t)x1 | (uintptr_t)x2 | (uintptr_t)out) % 16 == 0,
                    "rdeic_groupnorm_nhwc_f32: tensors must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    const int nchunk = gn_num_chunks(B, HW);
    launch_k(gn_stats_kernel<true>, dim3(nchunk, B), kGnThreads, 0, s, (const void*)x1, C1, (const void*)x2, C2, HW, groups,
             (float2*)workspace);
    RDEIC_LAUNCH_CHECK();
    const FastDiv div_vl((uint32_t)(C / 8));
    launch_k(gn_apply_kernel<true, true>, dim3((unsigned)nchunk, B), kGnThreads, 0, s, (const void*)x1, C1, (const void*)x2,
             C2, gamma, beta, (uint4*)out, HW, groups, eps, silu, (const float2*)workspace, nchunk, div_vl);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_layernorm_f32(const float* x, const float* gamma, const float* beta, float* out, int64_t rows, int C,
                        float eps, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(x && gamma && beta && out && rows >= 0, "rdeic_layernorm_f32: bad args");
    RDEIC_CHECK_ARG(C > 0 && C % 8 == 0 && C <= 8 * 32 * kLnMaxVecLimit,
                    "rdeic_layernorm_f32: C=%d must be a multiple of 8 and <= %d", C, 8 * 32 * kLnMaxVecLimit);
    RDEIC_CHECK_ARG(((uintptr_t)x | (uintptr_t)out | (uintptr_t)gamma | (uintptr_t)beta) % 16 == 0,
                    "rdeic_layernorm_f32: pointers must be 16-byte aligned");
    if (rows == 0) return 0;
    int64_t blocks = ceil_div64(rows, kLnWarps);
    if (blocks > 8ll * kNumSMs) blocks = 8ll * kNumSMs;
    launch_k(layernorm_kernel<true, 5, true>, (unsigned)blocks, kLnWarps * 32, 0, as_stream(stream), (const void*)x, gamma,
             beta, (uint4*)out, rows, C, eps);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_layernorm(const void* x, int in_is_f32, const float* gamma, const float* beta,
                    void* out, int64_t rows, int C, float eps, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(x && gamma && beta && out, "rdeic_layernorm: null pointer");
    RDEIC_CHECK_ARG(rows >= 0, "rdeic_layernorm: negative rows");
    RDEIC_CHECK_ARG(C > 0 && C % 8 == 0 && C <= 8 * 32 * kLnMaxVecLimit,
                    "rdeic_layernorm: C=%d must be a multiple of 8 and <= %d", C, 8 * 32 * kLnMaxVecLimit);
    RDEIC_CHECK_ARG(((uintptr_t)x | (uintptr_t)out | (uintptr_t)gamma | (uintptr_t)beta) % 16 == 0,
                    "rdeic_layernorm: pointers must be 16-byte aligned");
    if (rows == 0) return 0;
    int64_t blocks = ceil_div64(rows, kLnWarps);
    if (blocks > 8ll * kNumSMs) blocks = 8ll * kNumSMs;
    const int nv = (C / 8 + 31) / 32;   // 16-byte vectors per lane
    cudaStream_t s = as_stream(stream);
#define RDEIC_LN(F32, NV) launch_k(layernorm_kernel<F32, NV>, (unsigned)blocks, kLnWarps * 32, 0, s, x, gamma, beta, (uint4*)out, rows, C, eps)
    if (in_is_f32) {
        if (nv <= 1) RDEIC_LN(true, 1); else if (nv == 2) RDEIC_LN(true, 2); else if (nv == 3) RDEIC_LN(true, 3); else RDEIC_LN(true, 5);
    } else {
        if (nv <= 1) RDEIC_LN(false, 1); else if (nv == 2) RDEIC_LN(false, 2); else if (nv == 3) RDEIC_LN(false, 3); else RDEIC_LN(false, 5);
    }
#undef RDEIC_LN
    RDEIC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
