// GroupNorm(+SiLU) and LayerNorm over pixel-major (NHWC) bf16 activations, fp32 statistics.
// Both are HBM-bound: every element is read twice (stats pass hits L2 on the second read
// for UNet-sized tensors) and written once, in 16-byte vectors.
//
// Reference semantics: ldm/modules/diffusionmodules/util.py:224 GroupNorm32 (eps 1e-5, fp32
// compute), ldm/modules/diffusionmodules/model.py:48 Normalize (eps 1e-6),
// ldm/modules/attention.py:96 Normalize (eps 1e-6), model/rdeic.py:483 GroupNorm_leq32,
// ldm/modules/attention.py:273-275 nn.LayerNorm (eps 1e-5).
#include "common.cuh"
#include <stdlib.h>
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

// Fold width for statistics that arrive as a slab table: kFoldSlabs slabs per fold CTA (the apply kernel's
// prologue sums the nfold partials of its sample, 8 threads per group, four loads in flight).
// Pixel chunks of the apply pass: ONE resident wave of CTAs (the kernel's prologue and its load latency chain are paid once
// per CTA; with 8 * 148 / B chunks and 3-4 resident CTAs per SM the UNet level-0 pass ran 2.3 waves of 32-pixel CTAs).
// Tensors far larger than the L2 (the VAE's 512^2 / 256^2 levels) keep the finer 8 * 148 / B chunking: there the prologue is
// noise and the shorter tail of many small CTAs wins (A/B on one box: 217.6 vs 224.1 us on [8,512,512,128], 38.8 vs 34.8 us
// on the UNet's [8,64,64,320] fp32 stream, fold + apply).
template <int kTag, typename K>       // kTag: the instantiations share one function-pointer type, so the type alone would share the static
static inline int gn_apply_chunks(K kernel, int B, int64_t HW, int C) {
    if ((int64_t)B * HW * C > (int64_t)1 << 25) return gn_num_chunks(B, HW);
    static int per_sm = 0;              // one static per kernel instantiation
    if (per_sm == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kGnThreads, 0) != cudaSuccess || n < 1) n = 2;
        per_sm = n;
    }
    int64_t want = ((int64_t)per_sm * kNumSMs) / B;
    const int64_t max_by_rows = HW / 32 > 0 ? HW / 32 : 1;
    if (want > max_by_rows) want = max_by_rows;
    if (want > 65535) want = 65535;
    if (want < 1) want = 1;
    return (int)want;
}

static inline int gn_num_fold(int B, int64_t HW) {
    static const int per = getenv("RDEIC_GN_FOLD_SLABS") ? atoi(getenv("RDEIC_GN_FOLD_SLABS")) : 1;
    const int64_t slabs = HW / 32;
    int64_t want = (slabs + per - 1) / per;
    const int cap = gn_num_chunks(B, HW);
    if (want > cap) want = cap;
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

// the same in two steps: the raw 16 / 32 bytes now, the fp32 values later (keeps a prefetched bf16 row in 4 registers)
template <bool kF32> struct Raw8 { uint4 v[kF32 ? 2 : 1]; };
template <bool kF32>
__device__ __forceinline__ void load8_raw(const void* base, int64_t vec_index, Raw8<kF32>& r) {
    if (kF32) {
        r.v[0] = ld_stream_u4(reinterpret_cast<const uint4*>(base) + 2 * vec_index);
        r.v[kF32 ? 1 : 0] = ld_stream_u4(reinterpret_cast<const uint4*>(base) + 2 * vec_index + 1);
    } else {
        r.v[0] = ld_stream_u4(reinterpret_cast<const uint4*>(base) + vec_index);
    }
}
template <bool kF32>
__device__ __forceinline__ void unpack8(const Raw8<kF32>& r, float* f) {
    if (kF32) {
        const uint4 a = r.v[0], b = r.v[kF32 ? 1 : 0];
        f[0] = __uint_as_float(a.x); f[1] = __uint_as_float(a.y); f[2] = __uint_as_float(a.z); f[3] = __uint_as_float(a.w);
        f[4] = __uint_as_float(b.x); f[5] = __uint_as_float(b.y); f[6] = __uint_as_float(b.z); f[7] = __uint_as_float(b.w);
    } else {
        const uint4 v = r.v[0];
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
    pdl_trigger_short();
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
    pdl_trigger_short();
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

// Small tensors (UNet levels 2-3, the control adapter: <= 4 M elements): GroupNorm(+SiLU) in ONE kernel, one CTA per
// (sample, group).  At that size the statistics + apply pair (or fold + apply) is two launch latencies around a few
// microseconds of work; here the group's HW x C/G elements are read twice by the same CTA (second time from L1/L2),
// statistics are reduced in a fixed order (fp32 per thread, fp64 across the CTA: deterministic) and nothing goes
// through a workspace.  Two channels per access (C/G is even for every GroupNorm on the path).
constexpr int kGnSmallMaxElems = 1 << 22;
constexpr int kGnSmallGroupElems = 12288;        // elements of one (sample, group): 24 channel pairs per thread
constexpr int kGnSmallGroupElemsWide = 20480;    // second instantiation, 40 pairs per thread (UNet level 1, [8,32,32,640]: 1024 x 20):
                                                 // measured 23 us per launch against 18 us for fold + apply, so rdeic_groupnorm_is_small
                                                 // does not select it; kept for callers without fused statistics

template <bool kF32, int kGroupElems>
__global__ void __launch_bounds__(kGnThreads)
gn_small_kernel(const void* __restrict__ x1, int C1, const void* __restrict__ x2, int C2,
                const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ out,
                int HW, int G, float eps, int silu, FastDiv div_pairs) {
    pdl_trigger_short();
    pdl_wait();
    __shared__ double s_red[2][kGnThreads / 32];
    __shared__ float s_stat[2];
    const int C = C1 + C2, cg = C / G, pairs = cg >> 1;
    const int b = blockIdx.y, c0 = blockIdx.x * cg;
    const int total = HW * pairs;
    auto load2 = [&](int p, int c) -> float2 {
        const bool first = c < C1;
        const int64_t off = first ? ((int64_t)b * HW + p) * C1 + c : ((int64_t)b * HW + p) * C2 + (c - C1);
        const void* src = first ? x1 : x2;
        if (kF32) return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(src) + off);
        float2 v;
        unpack_bf16x2(*reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(src) + off), v.x, v.y);
        return v;
    };
    // the group's elements stay in registers between the statistics and the apply pass (<= 24 channel pairs per thread:
    // rdeic_groupnorm_is_small caps a group at 12 288 elements): every load is requested up front, one memory round trip
    // instead of a dozen dependent ones in a kernel that is nothing but latency (9.7 us per launch before)
    constexpr int kMaxPairs = kGroupElems / 2 / kGnThreads;
    float2 v[kMaxPairs];
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPairs; ++k) {
        const int i = threadIdx.x + k * kGnThreads;
        v[k] = make_float2(0.f, 0.f);
        if (i < total) {
            uint32_t p, j;
            div_pairs.divmod((uint32_t)i, p, j);
            v[k] = load2((int)p, c0 + 2 * (int)j);
        }
    }
#pragma unroll
    for (int k = 0; k < kMaxPairs; ++k) {
        s += v[k].x + v[k].y;
        q = fmaf(v[k].x, v[k].x, fmaf(v[k].y, v[k].y, q));
    }
    double ds = (double)s, dq = (double)q;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        ds += __shfl_xor_sync(0xffffffffu, ds, o);
        dq += __shfl_xor_sync(0xffffffffu, dq, o);
    }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = ds; s_red[1][threadIdx.x >> 5] = dq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < kGnThreads / 32; ++w) { a += s_red[0][w]; c += s_red[1][w]; }
        const double n = (double)HW * cg;
        const double mean = a / n;
        double var = c / n - mean * mean;        // biased variance, as nn.GroupNorm
        if (var < 0.0) var = 0.0;
        s_stat[0] = (float)mean;
        s_stat[1] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const float mean = s_stat[0], rstd = s_stat[1];
#pragma unroll
    for (int k = 0; k < kMaxPairs; ++k) {
        const int i = threadIdx.x + k * kGnThreads;
        if (i < total) {
            uint32_t p, j;
            div_pairs.divmod((uint32_t)i, p, j);
            const int c = c0 + 2 * (int)j;
            const float2 ga = *reinterpret_cast<const float2*>(gamma + c), be = *reinterpret_cast<const float2*>(beta + c);
            float y0 = fmaf((v[k].x - mean) * rstd, ga.x, be.x), y1 = fmaf((v[k].y - mean) * rstd, ga.y, be.y);
            if (silu) { y0 = silu_f(y0); y1 = silu_f(y1); }
            *reinterpret_cast<uint32_t*>(out + ((int64_t)b * HW + p) * C + c) = pack_bf16x2(y0, y1);
        }
    }
}


// (bf16 input: 4 resident CTAs per SM, i.e. at most 64 registers -- the VAE's 537 MB tensors want the bytes in flight)
template <bool kF32, bool kOutF32 = false>
__global__ void __launch_bounds__(kGnThreads, kF32 ? 3 : 4)
gn_apply_kernel(const void* __restrict__ x1, int C1, const void* __restrict__ x2, int C2,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                uint4* __restrict__ out, int64_t HW, int G, float eps, int silu,
                const float2* __restrict__ partial, int nchunk, FastDiv div_vl) {
    pdl_trigger_short();
    pdl_wait();
    __shared__ float s_scale[kGnMaxC];
    __shared__ float s_shift[kGnMaxC];
    __shared__ float s_mean[kGnMaxGroups];
    __shared__ float s_rstd[kGnMaxGroups];
    const int C = C1 + C2;
    const int VL = C >> 3, VL1 = C1 >> 3, VL2 = C2 >> 3;
    const int b = blockIdx.y;
    const int cg = C / G;
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
    // The first four rows of this thread's lane are requested BEFORE the statistics are finalised: the loads do not depend
    // on them, and the prologue below (a chain of L2 round trips, an fp64 reduction and two CTA barriers) is then hidden
    // behind the HBM latency instead of preceding it (UNet level 0: a CTA's whole share is 12 rows per thread).
    Raw8<kF32> f4[4];
    const bool pre_ok = row < rows_per_iter && lane_in < VL && p0 + row + 3 * (int64_t)rows_per_iter < p1;
    if (pre_ok) {
        const bool first = lane_in < VL1;
        const void* src = first ? x1 : x2;
        const int64_t stride = first ? VL1 : VL2;
        const int64_t off = (first ? o1 + lane_in : o2 + (lane_in - VL1));
#pragma unroll
        for (int u = 0; u < 4; ++u) load8_raw<kF32>(src, off + (p0 + row + u * (int64_t)rows_per_iter) * stride, f4[u]);
    }
    {   // finalise the statistics: 8 threads per group split the chunk partials (fixed order ->
        // deterministic), fp64 combine; a single thread per group would serialise nchunk L2 round trips
        const int g = threadIdx.x >> 3, part = threadIdx.x & 7;
        double a = 0.0, c = 0.0;
        if (g < G) {
            const float2* pp = partial + (int64_t)b * nchunk * G + g;
            int k = part;
            for (; k + 24 < nchunk; k += 32) {           // four independent L2 loads in flight, fixed order
                const float2 v0 = pp[(int64_t)k * G], v1 = pp[(int64_t)(k + 8) * G];
                const float2 v2 = pp[(int64_t)(k + 16) * G], v3 = pp[(int64_t)(k + 24) * G];
                a += (double)v0.x; c += (double)v0.y;
                a += (double)v1.x; c += (double)v1.y;
                a += (double)v2.x; c += (double)v2.y;
                a += (double)v3.x; c += (double)v3.y;
            }
            for (; k < nchunk; k += 8) {
                const float2 v = pp[(int64_t)k * G];
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
        auto finish = [&](const Raw8<kF32>& raw, int64_t p) {
            float f[8];
            unpack8<kF32>(raw, f);
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
        if (lane_base == 0 && pre_ok) {
#pragma unroll
            for (int u = 0; u < 4; ++u) finish(f4[u], p + u * (int64_t)rows_per_iter);
            p += 4 * (int64_t)rows_per_iter;
        }
        for (; p + 3 * (int64_t)rows_per_iter < p1; p += 4 * (int64_t)rows_per_iter) {
#pragma unroll
            for (int u = 0; u < 4; ++u) load8_raw<kF32>(src, off + (p + u * (int64_t)rows_per_iter) * stride, f4[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) finish(f4[u], p + u * (int64_t)rows_per_iter);
        }
        // tail: up to three rows, requested together
#pragma unroll
        for (int u = 0; u < 3; ++u)
            if (p + u * (int64_t)rows_per_iter < p1) load8_raw<kF32>(src, off + (p + u * (int64_t)rows_per_iter) * stride, f4[u]);
#pragma unroll
        for (int u = 0; u < 3; ++u)
            if (p + u * (int64_t)rows_per_iter < p1) finish(f4[u], p + u * (int64_t)rows_per_iter);
    }
}

// ------------------------------------------------------------------------------------------------
// VAE tail: norm_out -> swish -> conv_out (3x3, C -> 3) -> optional uint8 post-process, ONE kernel
// (model.py:683-686 `h = self.norm_out(h); h = nonlinearity(h); h = self.conv_out(h)`, inference.py:85-87).
// The tcgen05 implicit GEMM is the wrong tool for 3 output channels: an N = 32 tile spends its time
// reading the M = 128 A operand from shared memory (39-48 clocks per instruction for N <= 64 against N / 2 = 16
// of tensor work), and re-reads the activation 9 times from L2 (0.5-0.8 ms at 512^2, batch 8).  Here a CTA stages an (8+2) x (32+2) pixel halo tile ONCE, applying GroupNorm scale/shift and
// SiLU on the way into shared memory (zero outside the image = the conv's padding), and 8 warps run
// mma.sync m16n8k16 (16 pixels x 8 padded output channels) over the 9 taps from shared memory.
// The normalised activation never goes back to HBM (saves the gn_apply pass: 2 x 537 MB at 512^2).
// ------------------------------------------------------------------------------------------------
constexpr int kTailTH = 8, kTailTW = 32;
constexpr int kTailC = 128;
constexpr int kTailThreads = 512;
constexpr int kTailPix = (kTailTH + 2) * (kTailTW + 2);              // 340 halo pixels
constexpr int kTailPitch = kTailC * 2 + 16;                           // bytes per pixel (padded: conflict-free ldmatrix)
constexpr int kTailTileBytes = kTailPix * kTailPitch;                 // 92 480
constexpr int kTailWLane = 80;                                        // bytes per (tap, cout, k-quad) lane row: 64 + 16 pad
constexpr int kTailWBytes = 9 * 8 * 4 * kTailWLane;                   // B fragments in per-lane order: 4 LDS.128 per tap
constexpr int kTailSmem = 2 * kTailTileBytes + kTailWBytes;           // 208 000 B: one persistent CTA per SM
constexpr int kTailIters = (kTailPix + kTailThreads / 16 - 1) / (kTailThreads / 16);   // halo pixels per thread (11)

__device__ __forceinline__ float tail_u8(float x) {                  // inference.py:85-87, as image_to_u8_kernel
    float v = __fdiv_rn(__fadd_rn(x, 1.0f), 2.0f);
    v = fminf(fmaxf(v, 0.f), 1.f);
    v = __fmul_rn(v, 255.0f);
    return fminf(fmaxf(v, 0.f), 255.f);
}

// Persistent CTA: tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... over (sample, tile row, tile column).
// Two raw halo buffers: cp.async fills tile i+1 (zero-fill outside the image) while tile i is normalised in
// place (GroupNorm scale/shift + SiLU, bf16) and contracted, so the HBM latency is off the critical path.
__global__ void __launch_bounds__(kTailThreads, 1)
gn_silu_conv3x3_tail_kernel(const uint4* __restrict__ x, int B, int H, int W, const float2* __restrict__ partial,
                            int nfold, int G, float eps, const float* __restrict__ gamma,
                            const float* __restrict__ beta, const uint4* __restrict__ wpk,
                            const float* __restrict__ bias, int n_out, float* __restrict__ out_f32, int ldo,
                            uint8_t* __restrict__ out_u8, int tiles_w, int tiles_h) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) uint8_t tail_smem[];
    uint8_t* s_w = tail_smem + 2 * kTailTileBytes;
    __shared__ float s_scale[kTailC], s_shift[kTailC];
    __shared__ float s_mean[kGnMaxGroups], s_rstd[kGnMaxGroups];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cg = kTailC / G;
    const int tiles_per_sample = tiles_w * tiles_h;
    const int total = B * tiles_per_sample;
    const int l = tid & 15;                      // this thread's 8-channel lane in the staging passes

    // this thread's halo pixels are the same in every tile: p = tid/16 + 32k -> (hy, hx) packed once
    int hyx[kTailIters];
#pragma unroll
    for (int k = 0; k < kTailIters; ++k) {
        const int p = (tid >> 4) + k * (kTailThreads / 16);
        const int hy = p / (kTailTW + 2), hx = p - hy * (kTailTW + 2);
        hyx[k] = p < kTailPix ? ((hy << 8) | hx) : -1;
    }
    auto issue_load = [&](int t, int buf) {
        const int b = t / tiles_per_sample, r = t - b * tiles_per_sample;
        const int th0 = (r / tiles_w) * kTailTH - 1, tw0 = (r % tiles_w) * kTailTW - 1;
        const uint4* xb = x + (int64_t)b * H * W * (kTailC / 8) + l;
        const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(tail_smem + buf * kTailTileBytes) + l * 16 +
                              (tid >> 4) * kTailPitch;
#pragma unroll
        for (int k = 0; k < kTailIters; ++k) {
            if (hyx[k] < 0) break;
            const int gh = th0 + (hyx[k] >> 8), gw = tw0 + (hyx[k] & 255);
            const bool ok = (unsigned)gh < (unsigned)H && (unsigned)gw < (unsigned)W;
            const uint4* src = ok ? xb + ((int64_t)gh * W + gw) * (kTailC / 8) : x;
            const int nbytes = ok ? 16 : 0;      // 0 source bytes -> 16 bytes of zeros = the conv's padding
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + k * (kTailThreads / 16) * kTailPitch),
                         "l"(src), "r"(nbytes)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (blockIdx.x < total) issue_load(blockIdx.x, 0);
    // weights [tap][cout 8][C] -> per-lane B fragment order: lane (cout g, k-quad q) owns, for channel block kb,
    // words (kb*8 + q) and (kb*8 + q + 4) of its row; stored contiguously so a tap is 4 LDS.128 per lane
    {
        const uint32_t* wsrc = reinterpret_cast<const uint32_t*>(wpk);
        for (int i = tid; i < 9 * 8 * (kTailC / 2); i += kTailThreads) {
            const int row = i / (kTailC / 2), w = i % (kTailC / 2);
            const int kb = w >> 3, qq = w & 3, j = (w >> 2) & 1;
            *reinterpret_cast<uint32_t*>(s_w + (row * 4 + qq) * kTailWLane + (kb * 2 + j) * 4) = wsrc[i];
        }
    }
    const int a_row = (lane & 7) + 8 * ((lane >> 3) & 1), a_kc = 8 * (lane >> 4);
    const int orow = warp >> 1, mt = warp & 1;   // 16 warps: output row of the tile, 16-pixel half
    const int q = lane & 3, g4 = lane >> 2;
    const float bias0 = (2 * q < n_out) ? bias[2 * q] : 0.f, bias1 = (2 * q + 1 < n_out) ? bias[2 * q + 1] : 0.f;
    float sc[8], sh[8];
    int cur_b = -1, it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int buf = it & 1;
        const int b = t / tiles_per_sample, r = t - b * tiles_per_sample;
        const int th0 = (r / tiles_w) * kTailTH, tw0 = (r % tiles_w) * kTailTW;
        if (t + (int)gridDim.x < total) issue_load(t + gridDim.x, buf ^ 1);
        if (b != cur_b) {                        // new sample: statistics -> per-channel scale / shift
            cur_b = b;
            __syncthreads();
            const int g = tid >> 3, part = tid & 7;              // 8 threads per group (G <= 32 -> first 256 threads)
            double a = 0.0, c = 0.0;
            if (g < G)
                for (int k = part; k < nfold; k += 8) {
                    const float2 v = partial[((int64_t)b * nfold + k) * G + g];
                    a += (double)v.x; c += (double)v.y;
                }
#pragma unroll
            for (int o = 4; o; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                c += __shfl_xor_sync(0xffffffffu, c, o);
            }
            if (g < G && part == 0) {
                const double n = (double)H * W * cg;
                const double mean = a / n;
                double var = c / n - mean * mean;
                if (var < 0.0) var = 0.0;
                s_mean[g] = (float)mean;
                s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps));
            }
            __syncthreads();
            if (tid < kTailC) {
                const float k = gamma[tid] * s_rstd[tid / cg];
                s_scale[tid] = k;
                s_shift[tid] = beta[tid] - s_mean[tid / cg] * k;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 8; ++k) { sc[k] = s_scale[l * 8 + k]; sh[k] = s_shift[l * 8 + k]; }
        }
        if (t + (int)gridDim.x < total) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        uint8_t* s_x = tail_smem + buf * kTailTileBytes;
        // normalise + SiLU in place (pixels outside the image stay zero)
        {
            uint8_t* base = s_x + l * 16 + (tid >> 4) * kTailPitch;
            const int th1 = th0 - 1, tw1 = tw0 - 1;
#pragma unroll
            for (int k = 0; k < kTailIters; ++k) {
                if (hyx[k] < 0) break;
                const int gh = th1 + (hyx[k] >> 8), gw = tw1 + (hyx[k] & 255);
                if ((unsigned)gh >= (unsigned)H || (unsigned)gw >= (unsigned)W) continue;
                uint4* ptr = reinterpret_cast<uint4*>(base + k * (kTailThreads / 16) * kTailPitch);
                const uint4 v = *ptr;
                float f[8];
                unpack_bf16x2(v.x, f[0], f[1]); unpack_bf16x2(v.y, f[2], f[3]);
                unpack_bf16x2(v.z, f[4], f[5]); unpack_bf16x2(v.w, f[6], f[7]);
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    // y = x*scale + shift ; silu(y) = y / (1 + 2^(-y log2 e)), packed fp32x2 where it exists
                    const float2 y = __ffma2_rn(make_float2(f[j], f[j + 1]), make_float2(sc[j], sc[j + 1]), make_float2(sh[j], sh[j + 1]));
                    const float2 a = __fmul2_rn(y, make_float2(-1.4426950408889634f, -1.4426950408889634f));
                    float2 e, rr;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a.x));
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a.y));
                    const float2 d = __fadd2_rn(e, make_float2(1.0f, 1.0f));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rr.x) : "f"(d.x));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rr.y) : "f"(d.y));
                    const float2 o = __fmul2_rn(y, rr);
                    f[j] = o.x; f[j + 1] = o.y;
                }
                *ptr = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            }
        }
        __syncthreads();
        // warp -> (output row, 16-pixel half); accumulators: pixel rows g4 / g4+8, channels 2q, 2q+1.
        // Per tap all 8 channel blocks' fragments are loaded first (loads in flight together), then 8 MMAs
        // alternate between two accumulators so consecutive MMAs do not wait on each other.
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t sx = (uint32_t)__cvta_generic_to_shared(s_x);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int dy = tap / 3, dx = tap - dy * 3;               // halo coordinates: (row + dy, col + dx)
            const uint32_t a_base = sx + (uint32_t)(((orow + dy) * (kTailTW + 2) + mt * 16 + dx + a_row) * kTailPitch + a_kc * 2);
            const uint4* b_ptr = reinterpret_cast<const uint4*>(s_w + ((tap * 8 + g4) * 4 + q) * kTailWLane);
            uint32_t bf[kTailC / 16][2], af[kTailC / 16][4];
#pragma unroll
            for (int kq = 0; kq < kTailC / 64; ++kq) {
                const uint4 w4 = b_ptr[2 * kq], w5 = b_ptr[2 * kq + 1];
                bf[4 * kq][0] = w4.x; bf[4 * kq][1] = w4.y; bf[4 * kq + 1][0] = w4.z; bf[4 * kq + 1][1] = w4.w;
                bf[4 * kq + 2][0] = w5.x; bf[4 * kq + 2][1] = w5.y; bf[4 * kq + 3][0] = w5.z; bf[4 * kq + 3][1] = w5.w;
            }
#pragma unroll
            for (int kb = 0; kb < kTailC / 16; ++kb) {
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(af[kb][0]), "=r"(af[kb][1]), "=r"(af[kb][2]), "=r"(af[kb][3])
                             : "r"(a_base + (uint32_t)(kb * 32)));
            }
#pragma unroll
            for (int kb = 0; kb < kTailC / 16; kb += 2) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, "
                             "{%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3])
                             : "r"(af[kb][0]), "r"(af[kb][1]), "r"(af[kb][2]), "r"(af[kb][3]), "r"(bf[kb][0]), "r"(bf[kb][1]));
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, "
                             "{%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                             : "+f"(acc2[0]), "+f"(acc2[1]), "+f"(acc2[2]), "+f"(acc2[3])
                             : "r"(af[kb + 1][0]), "r"(af[kb + 1][1]), "r"(af[kb + 1][2]), "r"(af[kb + 1][3]),
                               "r"(bf[kb + 1][0]), "r"(bf[kb + 1][1]));
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] += acc2[k];
        // epilogue: q == 0 holds channels 0,1 ; q == 1 holds channels 2,3 of pixel rows g4 and g4+8
        const int gh = th0 + orow;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const int gw = tw0 + mt * 16 + g4 + 8 * hf;
            const float v0 = acc[2 * hf] + bias0, v1 = acc[2 * hf + 1] + bias1;
            const float v2 = __shfl_down_sync(0xffffffffu, v0, 1);        // channel 2 for the q == 0 lane
            if (gh >= H || gw >= W) continue;
            const int64_t pix = ((int64_t)b * H + gh) * W + gw;
            if (out_f32 && q < 2) {
                if (2 * q < ldo) out_f32[pix * ldo + 2 * q] = (2 * q < n_out) ? v0 : 0.f;
                if (2 * q + 1 < ldo) out_f32[pix * ldo + 2 * q + 1] = (2 * q + 1 < n_out) ? v1 : 0.f;
            }
            if (out_u8 && q == 0) {
                uint8_t* d = out_u8 + pix * 3;
                d[0] = (uint8_t)__float2int_rz(tail_u8(v0));
                d[1] = (uint8_t)__float2int_rz(tail_u8(v1));
                d[2] = (uint8_t)__float2int_rz(tail_u8(v2));
            }
        }
        __syncthreads();                         // everyone is done with this buffer before it is refilled
    }
}

// LayerNorm: one warp per row, row held in registers (C <= 8*32*kLnMaxVec); kLnMaxVec is a template
// parameter so narrow rows keep few registers live and the SM holds enough warps (= bytes in
// flight) to cover HBM latency.
constexpr int kLnMaxVecLimit = 5;
constexpr int kLnWarps = 8;

// kLanes = lanes per row: 32 (a warp per row) or 16 (two rows per warp): C = 320 is 40 vectors, which
// a full warp covers as 32 + 8 (37 % of the issued lanes idle, and the kernel is issue-bound) but a
// half warp covers as 16 + 16 + 8.
template <bool kF32, int kLnMaxVec, bool kOutF32 = false, int kLanes = 32>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_kernel(const void* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ beta, uint4* __restrict__ out, int64_t rows, int C,
                 float eps) {
    pdl_trigger_short();
    pdl_wait();
    const int VL = C >> 3;
    const int lane = threadIdx.x & (kLanes - 1);
    constexpr int kRowsPerWarp = 32 / kLanes;
    // grid-stride over rows: at most 8 resident CTAs per SM walk the tensor (32 768 short-lived CTAs
    // spent more time being scheduled than loading)
    // (a half warp whose row is past the end still runs the loop body with loads masked off, so the
    // full-mask shuffles below stay convergent)
    for (int64_t row0 = ((int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5)) * kRowsPerWarp; row0 < rows;
         row0 += (int64_t)gridDim.x * kLnWarps * kRowsPerWarp) {
    const int64_t row = row0 + ((threadIdx.x & 31) / kLanes);
    const bool live = row < rows;
    float f[kLnMaxVec][8];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxVec; ++j) {
        const int l = lane + kLanes * j;
        if (l < VL && live) {
            load8<kF32>(x, row * VL + l, f[j]);
#pragma unroll
            for (int k = 0; k < 8; ++k) sum += f[j][k];
        }
    }
    for (int o = kLanes / 2; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxVec; ++j) {
        const int l = lane + kLanes * j;
        if (l < VL && live) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float d = f[j][k] - mean;
                sq = fmaf(d, d, sq);
            }
        }
    }
    for (int o = kLanes / 2; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)C + eps);
#pragma unroll
    for (int j = 0; j < kLnMaxVec; ++j) {
        const int l = lane + kLanes * j;
        if (l < VL && live) {
            // (keeping gamma / beta in registers across rows was tried: 40 -> 77+ registers, a third of the
            // resident warps, 0.64 -> 0.60 of HBM; they are L1 hits here)
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

int rdeic_groupnorm_is_small(int B, int64_t HW, int C1, int C2, int groups) {
    const int C = C1 + C2;
    if (B <= 0 || HW <= 0 || groups <= 0 || C % groups) return 0;
    const int cg = C / groups;
    // pairs of channels must not straddle the two sources, and a group's pairs are 8-byte aligned fp32 / 4-byte bf16.
    // One CTA walks a whole group: worth it only while that is a couple of dozen iterations per thread and a pixel's
    // slice of the group fills at least one 32-byte sector (the control adapter's 64- and 128-channel levels at 64^2 /
    // 32^2 have 2 or 4 channels per group and thousands of pixels: measured 84 us here against ~15 us for fold + apply).
    return (int64_t)B * HW * C <= kGnSmallMaxElems && HW * (int64_t)cg <= kGnSmallGroupElems && cg >= 8 && cg % 2 == 0 && C1 % 2 == 0;
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
    if (rdeic_groupnorm_is_small(B, HW, C1, C2, groups)) {
        const FastDiv div_pairs((uint32_t)((C / groups) >> 1));
        const bool wide = HW * (int64_t)(C / groups) > kGnSmallGroupElems;
        auto go = [&](auto kernel) {
            launch_k(kernel, dim3((unsigned)groups, B), kGnThreads, 0, s, x1, C1, x2, C2, gamma, beta, (__nv_bfloat16*)out, (int)HW,
                     groups, eps, silu, div_pairs);
        };
        if (in_is_f32) { if (wide) go(gn_small_kernel<true, kGnSmallGroupElemsWide>); else go(gn_small_kernel<true, kGnSmallGroupElems>); }
        else { if (wide) go(gn_small_kernel<false, kGnSmallGroupElemsWide>); else go(gn_small_kernel<false, kGnSmallGroupElems>); }
        RDEIC_LAUNCH_CHECK();
        return 0;
    }
    const int nchunk = gn_num_chunks(B, HW);
    if (in_is_f32)
        launch_k(gn_stats_kernel<true>, dim3(nchunk, B), kGnThreads, 0, s, x1, C1, x2, C2, HW, groups, (float2*)workspace);
    else
        launch_k(gn_stats_kernel<false>, dim3(nchunk, B), kGnThreads, 0, s, x1, C1, x2, C2, HW, groups, (float2*)workspace);
    RDEIC_LAUNCH_CHECK();
    const FastDiv div_vl((uint32_t)(C / 8));
    if (in_is_f32)
        launch_k(gn_apply_kernel<true>, dim3((unsigned)gn_apply_chunks<0>(gn_apply_kernel<true>, B, HW, C), B), kGnThreads, 0, s, 
            x1, C1, x2, C2, gamma, beta, (uint4*)out, HW, groups, eps, silu, (const float2*)workspace, nchunk, div_vl);
    else
        launch_k(gn_apply_kernel<false>, dim3((unsigned)gn_apply_chunks<1>(gn_apply_kernel<false>, B, HW, C), B), kGnThreads, 0, s, 
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
    const int nfold = gn_num_fold(B, HW);
    launch_k(gn_fold_stats_kernel, dim3(nfold, B), kGnThreads, 0, s, (const float2*)stats1, C1,
             (const float2*)stats2, C2, HW / 32, groups, (float2*)workspace);
    RDEIC_LAUNCH_CHECK();
    const FastDiv div_vl((uint32_t)(C / 8));
    if (in_is_f32)
        launch_k(gn_apply_kernel<true>, dim3((unsigned)gn_apply_chunks<0>(gn_apply_kernel<true>, B, HW, C), B), kGnThreads, 0, s,
            x1, C1, x2, C2, gamma, beta, (uint4*)out, HW, groups, eps, silu, (const float2*)workspace, nfold, div_vl);
    else
        launch_k(gn_apply_kernel<false>, dim3((unsigned)gn_apply_chunks<1>(gn_apply_kernel<false>, B, HW, C), B), kGnThreads, 0, s,
            x1, C1, x2, C2, gamma, beta, (uint4*)out, HW, groups, eps, silu, (const float2*)workspace, nfold, div_vl);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_gn_silu_conv3x3_tail(const void* x, const float* stats, const float* gamma, const float* beta,
                               const void* w_packed, const float* bias, int n_out, float* out_f32, int ldo,
                               uint8_t* out_u8, int B, int H, int W, int C, int groups, float eps,
                               void* workspace, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(x && stats && gamma && beta && w_packed && bias && workspace, "rdeic_gn_silu_conv3x3_tail: null pointer");
    RDEIC_CHECK_ARG(out_f32 || out_u8, "rdeic_gn_silu_conv3x3_tail: no output pointer");
    RDEIC_CHECK_ARG(C == kTailC, "rdeic_gn_silu_conv3x3_tail: C must be %d (got %d)", kTailC, C);
    RDEIC_CHECK_ARG(n_out >= 1 && n_out <= 4 && (!out_u8 || n_out == 3) && (!out_f32 || ldo >= n_out),
                    "rdeic_gn_silu_conv3x3_tail: n_out must be 1..4 (3 for uint8 output), ldo >= n_out");
    RDEIC_CHECK_ARG(B > 0 && B <= 65535 && H > 0 && W > 0 && ((int64_t)H * W) % 32 == 0,
                    "rdeic_gn_silu_conv3x3_tail: H*W must be a positive multiple of 32");
    RDEIC_CHECK_ARG(groups > 0 && groups <= kGnMaxGroups && C % groups == 0, "rdeic_gn_silu_conv3x3_tail: bad groups");
    RDEIC_CHECK_ARG(((uintptr_t)x | (uintptr_t)w_packed | (uintptr_t)stats) % 16 == 0,
                    "rdeic_gn_silu_conv3x3_tail: tensors must be 16-byte aligned");
    cudaStream_t s = as_stream(stream);
    const int64_t HW = (int64_t)H * W;
    const int nfold = gn_num_fold(B, HW);
    launch_k(gn_fold_stats_kernel, dim3(nfold, B), kGnThreads, 0, s, (const float2*)stats, C, (const float2*)nullptr, 0,
             HW / 32, groups, (float2*)workspace);
    RDEIC_LAUNCH_CHECK();
    static bool attr_set = false;
    if (!attr_set) {
        RDEIC_CUDA(cudaFuncSetAttribute(gn_silu_conv3x3_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailSmem));
        attr_set = true;
    }
    const int tiles_w = (W + kTailTW - 1) / kTailTW, tiles_h = (H + kTailTH - 1) / kTailTH;
    const int64_t total = (int64_t)B * tiles_w * tiles_h;
    launch_k(gn_silu_conv3x3_tail_kernel, dim3((unsigned)(total < kNumSMs ? total : kNumSMs)), kTailThreads, kTailSmem, s,
             (const uint4*)x, B, H, W, (const float2*)workspace, nfold, groups, eps, gamma, beta, (const uint4*)w_packed,
             bias, n_out, out_f32, ldo, out_u8, tiles_w, tiles_h);
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
    launch_k(gn_apply_kernel<true, true>, dim3((unsigned)gn_apply_chunks<2>(gn_apply_kernel<true, true>, B, HW, C), B), kGnThreads, 0, s, (const void*)x1, C1, (const void*)x2,
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
    if (C / 8 > 32 && C / 8 <= 48) {    // 33..48 vectors: half a warp per row, three vectors per lane
        blocks = ceil_div64(rows, 2 * kLnWarps);
        if (blocks > 8ll * kNumSMs) blocks = 8ll * kNumSMs;
        if (in_is_f32)
            launch_k(layernorm_kernel<true, 3, false, 16>, (unsigned)blocks, kLnWarps * 32, 0, s, x, gamma, beta, (uint4*)out, rows, C, eps);
        else
            launch_k(layernorm_kernel<false, 3, false, 16>, (unsigned)blocks, kLnWarps * 32, 0, s, x, gamma, beta, (uint4*)out, rows, C, eps);
        RDEIC_LAUNCH_CHECK();
        return 0;
    }
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
