// Implicit-GEMM convolution / Linear on 5th-gen tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor) stages A and B tiles in 128B-swizzled shared memory,
//   one elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16, bf16 x bf16 -> fp32),
//   the accumulator lives in TMEM and is drained by 4 epilogue warps with tcgen05.ld.
//
// conv3x3 (pad 1, stride 1) never materialises im2col: the M tile is a TN x TH x TW box of
// output pixels and, for filter tap (dy,dx) and 64-channel block cb, the A tile is the same
// box of the NHWC input shifted by (dy,dx) — one 4-D TMA load whose out-of-bounds rows and
// columns are zero-filled by the hardware, which *is* the zero padding.  conv1x1 / Linear is
// the 1-tap case.  The channel concat feeding the UNet decoder blocks (openaimodel.py:804)
// is read as two K segments from two tensor maps.
//
// Reference call sites replaced: torch conv2d / F.linear in
//   ldm/modules/diffusionmodules/openaimodel.py:203,229,240,106,566,750,
//   model/rdeic.py:167-172,346,528,554, ldm/modules/attention.py:52,72,162-169,314,328,
//   ldm/modules/diffusionmodules/model.py:57-61,102-125,160-179,605,647.
#include "common.cuh"
#include "sm100.cuh"
#include <stdlib.h>
#include "../../include/rdeic_b200.h"

namespace rdeic {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                 // bf16 elements = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kATileBytes = kBlockM * kBlockK * 2;   // 16 KB
// warp0 TMA, warp1 MMA(+TMEM alloc), then kEW epilogue warps (8 or 12: two or three per TMEM lane
// quadrant; 12 gives the latency-bound epilogue of small-K problems one more warp per scheduler at the
// price of a 128-register cap and one pipeline stage)

struct ConvDev {
    int a_n, a_h, a_w;
    int tw_log2, th_log2;          // tile extents (powers of two), TN = 128 >> (tw+th)
    int tiles_w, tiles_h, tiles_n;
    int cblk1, cblk2, taps;
    int n_out;
    int w_batched;
    const float* bias;
    const float* row_bias; int row_bias_ld;
    const void* resid; int resid_is_f32; int ld_resid;
    float alpha;
    int act;
    float act_param;               // LeakyReLU negative slope (act == 3)
    __nv_bfloat16* out_bf16;
    float* out_f32;
    int ldo;
    float* stats_out;              // optional [M/32][n_out] (sum, sumsq) per 32-row slab and column
    float* partial;                // split-K: raw fp32 accumulators [split][M][n_out]
    int kb_per_split;              // k-blocks per k-split slice (= all of them without split-K)
    int splits, n_tiles;
    int64_t m_total;
    // ABI 5
    int up2;                       // nearest x2 upsample folded in: 4 parity classes x 2x2 taps (z = parity), output grid 2H x 2W
    int a2_center;                 // second K segment contributes the centre tap only (fused zero-conv injection / 1x1 skip)
    int in_stride;                 // 1, or 2: stride-2 conv, A is the [a_n, 2 a_h, 2 a_w] input read through a tensor map with element strides 2
    int tap_lo;                    // first tap offset of a 3x3 filter: -1 (pad 1) or 0 (pad bottom/right only)
    // work-item order inside a k-split / parity class: 0 = all M tiles of N tile 0, then N tile 1, ... (CTAs running at the
    // same time read the same weight tile); 1 = the N tiles of one M tile are consecutive items, so neighbouring CTAs fetch
    // the SAME activation boxes at the same time and the L2 serves them once (the activation side is what saturates the
    // L2 -> SM path of the N = 160 tiles: 16 KB per 320 tensor clocks per SM)
    int m_major;
};

__device__ __forceinline__ int conv_total_kb(const ConvDev& p) {
    return p.a2_center ? p.taps * p.cblk1 + p.cblk2 : p.taps * (p.cblk1 + p.cblk2);
}

// Epilogue math shared by the GEMM kernel (phase B) and the split-K reduce kernel: 4 consecutive
// output columns of one row.  v already holds act(acc + bias + row_bias).
struct EpiOut {
    const void* resid; int resid_is_f32; int ld_resid; float alpha;
    __nv_bfloat16* out_bf16; float* out_f32; int ldo; int n_cols;   // n_cols = valid output columns
};

__device__ __forceinline__ void epi_store4(const EpiOut& e, int64_t m, int n, float4 v) {
    const bool full = (n + 4 <= e.n_cols);
    if (full && (e.ldo & 3) == 0 && (!e.resid || (e.ld_resid & 3) == 0)) {
        if (e.resid) {
            float4 r;
            if (e.resid_is_f32) {
                r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(e.resid) + m * e.ld_resid + n);
            } else {
                const uint2 rv = *reinterpret_cast<const uint2*>(
                    reinterpret_cast<const __nv_bfloat16*>(e.resid) + m * e.ld_resid + n);
                unpack_bf16x2(rv.x, r.x, r.y);
                unpack_bf16x2(rv.y, r.z, r.w);
            }
            v.x = fmaf(e.alpha, v.x, r.x); v.y = fmaf(e.alpha, v.y, r.y);
            v.z = fmaf(e.alpha, v.z, r.z); v.w = fmaf(e.alpha, v.w, r.w);
        } else if (e.alpha != 1.0f) {
            v.x *= e.alpha; v.y *= e.alpha; v.z *= e.alpha; v.w *= e.alpha;
        }
        if (e.out_f32) *reinterpret_cast<float4*>(e.out_f32 + m * e.ldo + n) = v;
        if (e.out_bf16)
            *reinterpret_cast<uint2*>(e.out_bf16 + m * e.ldo + n) =
                make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    } else {
        const float vv[4] = {v.x, v.y, v.z, v.w};
        for (int j = 0; j < 4 && n + j < e.n_cols; ++j) {
            float x = vv[j];
            if (e.resid) {
                const float rv = e.resid_is_f32
                    ? reinterpret_cast<const float*>(e.resid)[m * e.ld_resid + n + j]
                    : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(e.resid)[m * e.ld_resid + n + j]);
                x = fmaf(e.alpha, x, rv);
            } else {
                x *= e.alpha;
            }
            if (e.out_f32) e.out_f32[m * e.ldo + n + j] = x;
            if (e.out_bf16) e.out_bf16[m * e.ldo + n + j] = __float2bfloat16_rn(x);
        }
    }
}

// Exact-erf GELU (torch F.gelu default) with a branch-free erf: Abramowitz & Stegun 7.1.26,
// |error| <= 1.5e-7, two MUFU ops + a short Horner chain, so 16 instances interleave freely in
// the epilogue (libdevice erff is a branchy call that serialises them).
__device__ __forceinline__ float gelu_erf(float x) {
#ifdef RDEIC_GELU_ERFF
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
#endif
    const float z = fabsf(x) * 0.70710678118654752f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z * 1.4426950408889634f));
    const float erf_abs = fmaf(-p * t, e, 1.0f);                 // erf(|x|/sqrt2)
    return 0.5f * x + 0.5f * fabsf(x) * erf_abs;                 // 0.5 x (1 + sign(x) erf|.|)
}

// Phase B of the epilogue for the common case (32 consecutive, in-bounds output rows; 32 full columns):
// the warp re-reads its 32x32 staged sub-tile so that 8 lanes cover 128 contiguous bytes of a row, adds
// the residual and writes fp32 and/or bf16.  Compile-time variants (residual type, outputs, statistics)
// selected by one warp-uniform switch per chunk: the generic form predicates every alternative, which
// doubled the issued instructions of the epilogue-bound small-K linears.
//   kRes: 0 none, 1 fp32, 2 bf16.   Pointers arrive offset to (first row of the lane, first column).
// The residual of a slab (8 row-steps x 4 columns per lane), requested BEFORE the accumulator chunk is read from tensor
// memory: the loads do not depend on it, and issued at the top of phase B their L2 / HBM latency (~1 us) was the largest
// single stall of the epilogue-bound small-K linears (ncu source page: 18 % of all samples on the first FFMA after them).
template <int kRes>
__device__ __forceinline__ void load_resid8(const void* __restrict__ resid, int64_t ld_resid, float4* rv) {
    if (kRes == 1) {
        const float* rp = reinterpret_cast<const float*>(resid);
#pragma unroll
        for (int i = 0; i < 8; ++i) rv[i] = *reinterpret_cast<const float4*>(rp + (int64_t)(4 * i) * ld_resid);
    } else if (kRes == 2) {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(resid);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint2 u = *reinterpret_cast<const uint2*>(rp + (int64_t)(4 * i) * ld_resid);
            unpack_bf16x2(u.x, rv[i].x, rv[i].y);
            unpack_bf16x2(u.y, rv[i].z, rv[i].w);
        }
    }
}

template <int kRes, bool kF32, bool kB16, bool kStats, bool kPre = false>
__device__ __forceinline__ void store_slab(const float4* __restrict__ stg, int lane, float alpha,
                                           const void* __restrict__ resid, int64_t ld_resid,
                                           float* __restrict__ of, __nv_bfloat16* __restrict__ ob, int64_t ldo,
                                           float* __restrict__ stats_dst, const float4* rv_pre = nullptr) {
    const int q = lane & 7, r0 = lane >> 3;
    float4 rv[8];
    if (kPre && kRes != 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) rv[i] = rv_pre[i];
    } else if (kRes == 1) {
        const float* rp = reinterpret_cast<const float*>(resid);
#pragma unroll
        for (int i = 0; i < 8; ++i) rv[i] = *reinterpret_cast<const float4*>(rp + (int64_t)(4 * i) * ld_resid);
    } else if (kRes == 2) {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(resid);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint2 u = *reinterpret_cast<const uint2*>(rp + (int64_t)(4 * i) * ld_resid);
            unpack_bf16x2(u.x, rv[i].x, rv[i].y);
            unpack_bf16x2(u.y, rv[i].z, rv[i].w);
        }
    }
    float4 st_s = make_float4(0.f, 0.f, 0.f, 0.f), st_q = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = 4 * i + r0;
        float4 val = stg[row * 8 + (q ^ (row & 7))];
        if (kRes != 0) {
            val.x = fmaf(alpha, val.x, rv[i].x); val.y = fmaf(alpha, val.y, rv[i].y);
            val.z = fmaf(alpha, val.z, rv[i].z); val.w = fmaf(alpha, val.w, rv[i].w);
        }
        if (kF32) *reinterpret_cast<float4*>(of + (int64_t)(4 * i) * ldo) = val;
        if (kB16)
            *reinterpret_cast<uint2*>(ob + (int64_t)(4 * i) * ldo) =
                make_uint2(pack_bf16x2(val.x, val.y), pack_bf16x2(val.z, val.w));
        if (kStats) {
            st_s.x += val.x; st_s.y += val.y; st_s.z += val.z; st_s.w += val.w;
            st_q.x = fmaf(val.x, val.x, st_q.x); st_q.y = fmaf(val.y, val.y, st_q.y);
            st_q.z = fmaf(val.z, val.z, st_q.z); st_q.w = fmaf(val.w, val.w, st_q.w);
        }
    }
    if (kStats) {
        // lanes l, l+8, l+16, l+24 hold the same 4 columns for different rows
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            st_s.x += __shfl_xor_sync(0xffffffffu, st_s.x, o); st_s.y += __shfl_xor_sync(0xffffffffu, st_s.y, o);
            st_s.z += __shfl_xor_sync(0xffffffffu, st_s.z, o); st_s.w += __shfl_xor_sync(0xffffffffu, st_s.w, o);
            st_q.x += __shfl_xor_sync(0xffffffffu, st_q.x, o); st_q.y += __shfl_xor_sync(0xffffffffu, st_q.y, o);
            st_q.z += __shfl_xor_sync(0xffffffffu, st_q.z, o); st_q.w += __shfl_xor_sync(0xffffffffu, st_q.w, o);
        }
        if (lane < 8) {
            float4* dst = reinterpret_cast<float4*>(stats_dst);
            dst[0] = make_float4(st_s.x, st_q.x, st_s.y, st_q.y);
            dst[1] = make_float4(st_s.z, st_q.z, st_s.w, st_q.w);
        }
    }
}

template <int kRes, bool kStats, bool kPre = false>
__device__ __forceinline__ void store_slab_out(const float4* stg, int lane, float alpha, const void* resid,
                                               int64_t ld_resid, float* of, __nv_bfloat16* ob, int64_t ldo,
                                               float* stats_dst, const float4* rv_pre = nullptr) {
    if (of && ob) store_slab<kRes, true, true, kStats, kPre>(stg, lane, alpha, resid, ld_resid, of, ob, ldo, stats_dst, rv_pre);
    else if (of) store_slab<kRes, true, false, kStats, kPre>(stg, lane, alpha, resid, ld_resid, of, ob, ldo, stats_dst, rv_pre);
    else store_slab<kRes, false, true, kStats, kPre>(stg, lane, alpha, resid, ld_resid, of, ob, ldo, stats_dst, rv_pre);
}

// kMT = M tiles per work item: with kMT == 2 one CTA walks two adjacent 128-row M tiles against the
// SAME B (weight) stage, so the weight tile crosses L2 -> shared memory once per 256 output rows.
// For N <= 128 the weight tile is as large as the activation tile and both are re-fetched per tile
// (conv 128->128 at 512^2: 9.7 GB of L2->SM traffic, ~13 TB/s, which is what bounds it).
// Two GELUs per call on the packed fp32x2 pipe (FFMA2 / FMUL2 / FADD2, sm_100): half the FMA-class
// issue slots of the scalar form; same formula, same rounding per lane.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
#ifdef RDEIC_GELU_ERFF
    return make_float2(gelu_erf(x.x), gelu_erf(x.y));
#endif
    const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    const float2 ahx = make_float2(fabsf(hx.x), fabsf(hx.y));
    const float2 z = __fmul2_rn(ahx, make_float2(1.41421356237309505f, 1.41421356237309505f));      // |x| / sqrt(2)
    const float2 d = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
    float2 t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(d.y));
    float2 p = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
    p = __ffma2_rn(p, t, make_float2(1.421413741f, 1.421413741f));
    p = __ffma2_rn(p, t, make_float2(-0.284496736f, -0.284496736f));
    p = __ffma2_rn(p, t, make_float2(0.254829592f, 0.254829592f));
    const float2 a = __fmul2_rn(__fmul2_rn(z, z), make_float2(-1.4426950408889634f, -1.4426950408889634f));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a.y));
    const float2 pt = __fmul2_rn(p, t);
    const float2 erf_abs = __ffma2_rn(make_float2(-pt.x, -pt.y), e, make_float2(1.0f, 1.0f));
    return __ffma2_rn(ahx, erf_abs, hx);                         // 0.5 x + 0.5 |x| erf(|x|/sqrt2)
}

// GELU in its tanh form for the GEGLU gate (attention.py:54-56) of the bf16 throughput mode:
//   0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))),  one MUFU (tanh.approx) + 5 packed fp32x2 operations per PAIR
// against two MUFU (rcp, ex2) and ~12 packed operations for the erf form above.  The GEGLU GEMMs are bound by this
// epilogue arithmetic (K = 320..1280 against N = 2560..10240 gate/value columns), not by the tensor pipe.  The tanh form
// differs from F.gelu's exact erf by at most 4.8e-4 absolute (1.8e-4 rms for gates ~ N(0, 1.5)), a ninth of the bf16
// rounding (1.7e-3 rms) the product takes on its way to the next GEMM.  Exact-GELU users (act = 4: the compressor's
// entropy-parameter nets) keep the erf form.
__device__ __forceinline__ float2 gelu_tanh2(float2 x) {
    const float2 x2 = __fmul2_rn(x, x);
    const float2 t0 = __ffma2_rn(x2, make_float2(0.0356774081f, 0.0356774081f), make_float2(0.7978845608f, 0.7978845608f));
    const float2 u = __fmul2_rn(x, t0);
    float2 th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(u.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(u.y));
    const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    return __ffma2_rn(hx, th, hx);
}

template <int BN, int kEW, int kMT = 1, int kIss = 1, int kNT = 1>
struct TileCfg {
    static_assert(kMT * kNT * BN <= 512, "the accumulators of one work item must fit the 512 TMEM columns");
    static_assert(kNT == 1 || (kNT == 2 && kMT == 1 && kIss == 1), "two N tiles per item: one M tile, one issuing warp");
    static_assert(kIss == 1 || (kIss == 2 && kMT == 2), "two issuing warps: one per M tile of the item");
    // warp 0 TMA, warp 1 (and 2 with kIss == 2) MMA, warp 3 idle: one warp group of control warps, so that setmaxnreg can
    // hand their registers to the epilogue warp groups (warps 4 ..)
    static constexpr int kThreads = 128 + 32 * kEW;
    static constexpr int kFirstEpiWarp = 4;
    static constexpr int kRegsCtl = 56;
    static constexpr int kRegsEpi = kEW == 8 ? 224 : 152;               // 128 * 56 + 256 * 224 = 384 * 168 ; 128 * 56 + 384 * 152 = 512 * 128
    static_assert(kEW == 8 || kEW == 12, "register split is written for 8 or 12 epilogue warps");
    static constexpr int kBTileBytes = BN * kBlockK * 2;
    static constexpr int kBBytes = kNT * kBTileBytes;
    static constexpr int kABytes = kMT * kATileBytes;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStagingBytes = kEW * 4096;                   // 32 rows x 128 B per epilogue warp
    static constexpr int kMaxSmem = 227 * 1024;
    static constexpr int kStagesRaw = (kMaxSmem - kStagingBytes - 1024 - 256) / kStageBytes;
    static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
    static constexpr int kAccCols = kMT * kNT * BN;
    // two accumulator buffers (the epilogue of item i overlaps the main loop of item i+1) when they fit; one otherwise
    // (2 x 160 columns: the epilogue is exposed, 2-3 thousand clocks against a main loop of tens of thousands)
    static constexpr int kAccBufs = kAccCols <= 256 ? 2 : 1;
    static constexpr int kAccStride = kAccCols <= 32 ? 32 : kAccCols <= 64 ? 64 : kAccCols <= 128 ? 128 : kAccCols <= 256 ? 256 : 512;   // TMEM columns per buffer
    static constexpr int kTmemCols = kAccBufs * kAccStride;
    static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// Persistent, warp-specialised kernel: one CTA per SM loops over (m-tile, n-tile, k-split) work
// items.  warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-3 idle (control warp group), warps 4.. = epilogue.
// Three pipelines: smem stages (TMA <-> MMA), two TMEM accumulator buffers (MMA <-> epilogue, so
// the epilogue of tile i overlaps the main loop of tile i+1), and the static tile schedule.
// kSwap (with kMT == 2, BN == 128): operand roles exchanged.  With the WEIGHT tile as the M operand (128 output
// channels) and the item's two pixel tiles as one N = 256 operand, the same FLOPs take half the tcgen05.mma instructions
// and the weight tile is read from shared memory once per 256 pixels.  (Built in round 1, when every instruction cost
// ~100 clocks behind a `lane == 0` branch -- see elect_one_sync() -- and kept: 793 -> 558 us then, still the faster form.)  The accumulator is
// then transposed (TMEM lane = output channel, column = pixel); phase A parks it in the staging buffer
// already transposed back, so phase B (residual, stores, statistics) is unchanged.
// kUp: compile-time copy of p.up2 for the epilogue (strided output rows, parity-ordered statistics slabs): the
// plain instantiations carry none of that state (the 12-warp ones sit at their 128-register cap).
// kIss: warps issuing tcgen05.mma.  Behind a `lane == 0` branch ONE thread sustained one instruction per ~100 clocks
// whatever the shape (scripts/micro/ubench.cu: N = 16 .. 128 all 100 clocks, N = 160 113, N = 256 136-161; two issuing warps
// 52 per SM, four 26).  With kIss == 2 (and kMT == 2) each M tile of the item has its own issuing warp and accumulator: both
// streams are in order, so the result is bit-identical to the one-issuer kernel.  The cause turned out to be the branch,
// not the hardware (elect_one_sync(): N / 2 clocks per instruction from one thread), so this variant is off by default.
// kNT = 2: two N tiles per work item against ONE activation stage (accumulators side by side in tensor memory).  What bounds
// an N = 160 tile with a long reduction is the L2 -> SM path, not the tensor pipe: 16 KB of activations per 320 tensor clocks
// is 51 B/clk per SM against ~43 B/clk the L2 delivers chip-wide (the weight tile is the same for every CTA at a given
// moment and is largely served once).  With both N tiles of a 320-channel layer in one item the activation box crosses
// L2 -> SM once per 640 tensor clocks.  2 x 160 columns leave no room for a second accumulator buffer: the epilogue of an
// item is exposed (a few thousand clocks against a main loop of 45+ k-blocks x 640).
template <int BN, int kResidMode, int kEW, bool kStats, int kMT = 1, bool kSwap = false, bool kUp = false, int kIss = 1, int kNT = 1>
__global__ void __launch_bounds__(128 + 32 * kEW, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_a2,
                 const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_b2, const ConvDev p) {
    pdl_trigger();
    using Cfg = TileCfg<BN, kEW, kMT, kIss, kNT>;
    constexpr int kStages = Cfg::kStages;
    constexpr int kAccBufs = Cfg::kAccBufs;
    constexpr int kCStride = kEW / 4;            // epilogue warps per TMEM lane quadrant = chunk stride
    // residual handling in the epilogue: 0 = loaded at the top of phase B, 1 = bf16 residual
    // prefetched two chunks deep, 2 = fp32 residual prefetched one phase ahead (single buffer)
    constexpr bool kPrefetchResid = (kResidMode == 1);
    constexpr bool kAheadF32 = (kResidMode == 2);
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array (not an integer round trip), so
    // the compiler keeps the address space and the staging traffic is LDS/STS, not generic LD/ST
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * Cfg::kABytes;
    uint8_t* smem_stg = smem + kStages * Cfg::kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_stg + Cfg::kStagingBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* acc_full = empty_bar + kStages;        // [2]
    uint64_t* acc_empty = acc_full + 2;              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tw = 1 << p.tw_log2, th = 1 << p.th_log2;
    const int cbt = p.cblk1 + p.cblk2;
    const int total_kb = conv_total_kb(p);
    const int m_tiles = (p.tiles_w * p.tiles_h * p.tiles_n + kMT - 1) / kMT;   // work items along M
    const int mn_tiles = m_tiles * p.n_tiles;
    const int num_items = mn_tiles * p.splits;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_b);
        if (p.cblk2) tma_prefetch_desc(&tm_a2);
        if (p.a2_center) tma_prefetch_desc(&tm_b2);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kIss);          // every issuing warp releases the stage
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], kIss);
            mbar_init(&acc_empty[b], kEW);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barrier init, descriptor prefetch, TMEM allocation) overlapped the previous
    // kernel's tail; from here on we touch memory it may still be writing
    pdl_wait();

    if (warp == 0) {
        setmaxnreg_dec<Cfg::kRegsCtl>();
        if (elect_one_sync()) {
            // ===== TMA producer =====
            // One thread feeds the whole pipeline, and with N <= 160 tiles a k-block is only ~450 tensor clocks:
            // the per-k-block instruction count of this loop is on the critical path.  So no division or modulo
            // per k-block: (segment, tap, channel block) and (stage, phase) advance as counters.
            uint32_t s = 0, ph = 0;
            const int kw = p.up2 ? 2 : (p.taps == 9 ? 3 : (p.taps == 25 ? 5 : 1));     // filter width
            const int cs = p.in_stride;               // input coordinates = cs * output coordinates + tap offset
            const int kb_seg1 = p.taps * p.cblk1;
            for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
                const int z = item / mn_tiles;
                int rem = item - z * mn_tiles;
                const int nt = p.m_major ? rem % p.n_tiles : rem / m_tiles;
                const int mi = p.m_major ? rem / p.n_tiles : rem - nt * m_tiles;
                int w0[kMT], h0[kMT], n0[kMT];
#pragma unroll
                for (int u = 0; u < kMT; ++u) {
                    int mt = mi * kMT + u;             // a tile index past the end lands at n0 >= a_n: TMA zero-fills
                    const int tiw = mt % p.tiles_w; mt /= p.tiles_w;
                    const int tih = mt % p.tiles_h; mt /= p.tiles_h;
                    w0[u] = tiw * tw; h0[u] = tih * th;
                    n0[u] = mt * (kBlockM >> (p.tw_log2 + p.th_log2));
                }
                const int col0 = nt * (kNT * BN);
                // up2: z is the output parity class (py, px), not a k slice; it also selects the weight matrix
                const int par = p.up2 ? z : 0;
                const int wz = p.up2 ? par : (p.w_batched ? n0[0] : 0);
                const int kb_begin = p.up2 ? 0 : z * p.kb_per_split;
                const int kb_end = min(total_kb, kb_begin + p.kb_per_split);
                // offset of tap (0, 0): -1 for a padded 3x3, -2 for 5x5, (py - 1, px - 1) for the folded upsample
                const int oy = p.up2 ? (par >> 1) - 1 : (kw == 3 ? p.tap_lo : (kw == 5 ? -2 : 0));
                const int ox = p.up2 ? (par & 1) - 1 : oy;
                // position of the first k-block of this item (a division only for split-K slices past the first)
                int tap = 0, cb = 0, seg = 0;
                if (kb_begin != 0) {
                    if (p.a2_center) {
                        seg = kb_begin >= kb_seg1;
                        if (seg) { tap = p.taps; cb = kb_begin - kb_seg1; }
                        else { tap = kb_begin / p.cblk1; cb = kb_begin - tap * p.cblk1; }
                    } else {
                        tap = kb_begin / cbt; cb = kb_begin - tap * cbt;
                        seg = cb >= p.cblk1;
                        if (seg) cb -= p.cblk1;
                    }
                }
                int ty = tap / kw, tx = tap - ty * kw;
                for (int kb = kb_begin; kb < kb_end; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
                    // the injected tensor (centre-tap segment) lives on the output grid, whatever the conv's stride
                    const bool ctr = seg && p.a2_center;
                    const int dy = ctr ? 0 : oy + ty, dx = ctr ? 0 : ox + tx;
                    // ... and with the upsample folded in (up2) that grid is 2H x 2W: this parity class's pixels are every
                    // other one of it (the a2 map has element strides 2), starting at (py, px)
                    const int mul = ctr ? (p.up2 ? 2 : 1) : cs;
                    const int ey = (ctr && p.up2) ? (par >> 1) : dy, ex = (ctr && p.up2) ? (par & 1) : dx;
                    const CUtensorMap* ma = seg ? &tm_a2 : &tm_a;
#pragma unroll
                    for (int u = 0; u < kMT; ++u)
                        tma_load_4d(ma, smem_a + (s * kMT + u) * kATileBytes, &full_bar[s], cb * kBlockK, w0[u] * mul + ex,
                                    h0[u] * mul + ey, n0[u]);
#pragma unroll
                    for (int v = 0; v < kNT; ++v) {      // weight rows past n_out are zero-filled by the hardware
                        uint8_t* dstb = smem_b + s * Cfg::kBBytes + v * Cfg::kBTileBytes;
                        if (ctr) tma_load_3d(&tm_b2, dstb, &full_bar[s], cb * kBlockK, col0 + v * BN, 0);
                        else tma_load_3d(&tm_b, dstb, &full_bar[s], kb * kBlockK, col0 + v * BN, wz);
                    }
                    // advance (segment, tap, channel block)
                    if (++cb == (seg ? p.cblk2 : p.cblk1)) {
                        cb = 0;
                        bool next_tap = true;
                        if (p.a2_center) { if (++tap == p.taps) { seg = 1; next_tap = false; } }
                        else if (!seg && p.cblk2) { seg = 1; next_tap = false; }
                        else seg = 0;
                        if (next_tap && ++tx == kw) { tx = 0; ++ty; }
                    }
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1 || (kIss == 2 && warp == 2)) {
        setmaxnreg_dec<Cfg::kRegsCtl>();
        if (elect_one_sync()) {
            // ===== MMA issuer (kIss == 2: warp 1 owns M tile 0 of every item, warp 2 M tile 1) =====
            constexpr uint32_t idesc = make_idesc(BN);
            const int u_lo = kIss == 2 ? warp - 1 : 0, u_hi = kIss == 2 ? warp : kMT;
            uint32_t s = 0, ph = 0, t = 0;
            const uint64_t da0 = make_smem_desc(smem_u32(smem_a)), db0 = make_smem_desc(smem_u32(smem_b));
            for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++t) {
                const int z = item / mn_tiles;
                const int kb_begin = p.up2 ? 0 : z * p.kb_per_split;
                const int num_kb = min(total_kb, kb_begin + p.kb_per_split) - kb_begin;
                const uint32_t buf = t % kAccBufs, aph = (t / kAccBufs) & 1;
                mbar_wait(&acc_empty[buf], aph ^ 1);          // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * Cfg::kAccStride;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    // descriptors advance by whole stages: (bytes >> 4) added to the 14-bit start-address field
                    const uint64_t db = db0 + (uint64_t)(s * (Cfg::kBBytes >> 4));
                    const uint64_t dst = da0 + (uint64_t)(s * (Cfg::kABytes >> 4));
                    if constexpr (kSwap) {
                        // weights (BN = 128 rows) as the M operand, both pixel tiles (256 rows, contiguous) as N
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k)
                            umma_bf16(tmem_d, db + 2 * k, dst + 2 * k, make_idesc(kMT * kBlockM), (kb | k) != 0);
                    } else
#pragma unroll
                    for (int u = 0; u < kMT; ++u) {
                        if (u < u_lo || u >= u_hi) continue;
                        const uint64_t da = dst + (uint64_t)(u * (kATileBytes >> 4));
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            // advance 16 elements = 32 bytes along K inside the swizzle row: +2 (>>4)
#pragma unroll
                            for (int v = 0; v < kNT; ++v)
                                umma_bf16(tmem_d + (u * kNT + v) * BN, da + 2 * k, db + (uint64_t)(v * (Cfg::kBTileBytes >> 4)) + 2 * k, idesc,
                                          (kb | k) != 0);
                        }
                    }
                    umma_commit(&empty_bar[s]);   // frees the stage when these MMAs retire
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
                umma_commit(&acc_full[buf]);      // accumulator complete
            }
        }
    } else if (warp < Cfg::kFirstEpiWarp) {
        setmaxnreg_dec<Cfg::kRegsCtl>();             // idle warps of the control warp group
    } else if constexpr (kSwap) {
        setmaxnreg_inc<Cfg::kRegsEpi>();
        // ===== epilogue, exchanged operands: TMEM lane = output channel, column = pixel =====
        // Host guarantees: every real tile is full and affine, n_out % 32 == 0, no per-sample bias, no GEGLU,
        // one k-split, vector-aligned rows, alpha == 1 unless there is a residual.
        const int ew = warp - Cfg::kFirstEpiWarp;
        const int quad = warp & 3;                    // channels col0 + 32 * quad + lane
        const int half = ew >> 2;
        float4* stg = reinterpret_cast<float4*>(smem_stg) + ew * 256;
        float* stg_f = reinterpret_cast<float*>(stg);
        const int total_tiles = p.tiles_w * p.tiles_h * p.tiles_n;
        const int tn_ = kBlockM >> (p.tw_log2 + p.th_log2);
        constexpr int kChunks = kMT * kBlockM / 32;   // 32-pixel chunks of the item
        uint32_t t = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++t) {
            const int nt = item / m_tiles;
            const int mi = item - nt * m_tiles;
            const int col0 = nt * BN;
            const int cbase = col0 + 32 * quad;
            const bool ch_ok = cbase < p.n_out;
            const float bias_c = (ch_ok && p.bias) ? p.bias[cbase + lane] : 0.f;
            const uint32_t buf = t % kAccBufs, aph = (t / kAccBufs) & 1;
            mbar_wait(&acc_full[buf], aph);
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + buf * Cfg::kAccStride + ((uint32_t)(quad * 32) << 16);
            int last_c = -1;
            if (ch_ok)
                for (int ci = half; ci < kChunks; ci += kCStride) last_c = ci;
            if (last_c < 0) {
                tc_fence_before();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
#pragma unroll 1
            for (int ci = half; ci < kChunks && ch_ok; ci += kCStride) {
                // coordinates of the chunk's first pixel (lane 0's row)
                int mt = mi * kMT + (ci >> 2);
                const bool tile_ok = mt < total_tiles;
                const int tiw = mt % p.tiles_w; mt /= p.tiles_w;
                const int tih = mt % p.tiles_h; mt /= p.tiles_h;
                const int r0 = (ci & 3) * 32;
                const int gw = tiw * tw + (r0 & (tw - 1)), gh = tih * th + ((r0 >> p.tw_log2) & (th - 1));
                const int gn = mt * tn_ + (r0 >> (p.tw_log2 + p.th_log2));
                const int64_t m_slab = ((int64_t)gn * p.a_h + gh) * p.a_w + gw;
                // (the residual stays at the top of phase B here: requested ahead of the accumulator read it cost 22 % on the
                // VAE's 128-channel convs with a bf16 residual, 593 -> 725 us)
                uint32_t acc[32];
                tmem_ld16(tmem_acc + (uint32_t)(ci * 32), acc);
                tmem_ld16(tmem_acc + (uint32_t)(ci * 32) + 16, acc + 16);
                tmem_ld_wait();
                if (ci == last_c) {
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                if (!tile_ok) continue;                              // the partner of an odd last tile
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) + bias_c;
                if (p.act == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
                } else if (p.act == 3) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.act_param;
                } else if (p.act == 4) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                }
                // transposed park: element (pixel j, channel lane) -> 16-byte chunk (lane/4) ^ (j & 7) of row j
#pragma unroll
                for (int j = 0; j < 32; ++j) stg_f[(j * 8 + ((lane >> 2) ^ (j & 7))) * 4 + (lane & 3)] = v[j];
                __syncwarp();
                {
                    const int q = lane & 7;
                    const int64_t m0 = m_slab + (lane >> 3);
                    const int64_t o_off = m0 * p.ldo + cbase + 4 * q;
                    float* of = p.out_f32 ? p.out_f32 + o_off : nullptr;
                    __nv_bfloat16* ob = p.out_bf16 ? p.out_bf16 + o_off : nullptr;
                    float* sd = kStats ? p.stats_out + ((m_slab >> 5) * p.n_out + cbase + 4 * q) * 2 : nullptr;
                    const int64_t r_off = m0 * p.ld_resid + cbase + 4 * q;
                    if (!p.resid)
                        store_slab_out<0, kStats>(stg, lane, p.alpha, nullptr, 0, of, ob, p.ldo, sd);
                    else if (p.resid_is_f32)
                        store_slab_out<1, kStats>(stg, lane, p.alpha, reinterpret_cast<const float*>(p.resid) + r_off,
                                                  p.ld_resid, of, ob, p.ldo, sd);
                    else
                        store_slab_out<2, kStats>(stg, lane, p.alpha, reinterpret_cast<const __nv_bfloat16*>(p.resid) + r_off,
                                                  p.ld_resid, of, ob, p.ldo, sd);
                }
                __syncwarp();
            }
        }
    } else {
        setmaxnreg_inc<Cfg::kRegsEpi>();
        // ===== epilogue: TMEM -> registers -> (swizzled smem transpose) -> coalesced global =====
        // Phase A: each thread owns one accumulator row (TMEM lane) and 32 columns per chunk; it
        // adds bias / per-sample bias, applies the activation and parks the row in shared memory
        // (16-byte chunks XOR-swizzled by row: conflict-free STS.128).  Phase B: the warp re-reads
        // its 32x32 sub-tile so that 8 lanes cover 128 contiguous bytes of one output row, adds
        // the residual (prefetched with the same coalesced mapping) and writes fp32 and/or bf16.
        // Two warps share each TMEM lane quadrant and take alternate 32-column chunks.
        const int ew = warp - Cfg::kFirstEpiWarp;     // 0..kEW-1
        const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
        const int half = ew >> 2;                     // which chunks of the tile: half, half + kCStride, ...
        const int r = quad * 32 + lane;               // row inside the M tile
        const int rw = r & (tw - 1);
        const int rh = (r >> p.tw_log2) & (th - 1);
        const int rn = r >> (p.tw_log2 + p.th_log2);
        const bool geglu = (p.act == 2);
        const bool partial = (p.partial != nullptr);
        float4* stg = reinterpret_cast<float4*>(smem_stg) + ew * 256;   // 32 rows x 8 chunks
        uint32_t t = 0;
        for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++t) {
            const int z = item / mn_tiles;
            int rem = item - z * mn_tiles;
            const int nt = p.m_major ? rem % p.n_tiles : rem / m_tiles;
            const int mi = p.m_major ? rem / p.n_tiles : rem - nt * m_tiles;
#pragma unroll 1
            for (int sub = 0; sub < kMT * kNT; ++sub) {   // the item's M (or N) tiles share one accumulator buffer
            const int u = kNT == 2 ? 0 : sub;
            int mt = mi * kMT + u;
            const int tiw = mt % p.tiles_w; mt /= p.tiles_w;
            const int tih = mt % p.tiles_h; mt /= p.tiles_h;
            const int gw = tiw * tw + rw, gh = tih * th + rh;
            const int gn = mt * (kBlockM >> (p.tw_log2 + p.th_log2)) + rn;
            const int col0 = (nt * kNT + (kNT == 2 ? sub : 0)) * BN;
            const int row_ok = (gw < p.a_w && gh < p.a_h && gn < p.a_n) ? 1 : 0;
            const int m_in = (int)(((int64_t)gn * p.a_h + gh) * p.a_w + gw);
            // row of the output tensor this accumulator row is written to: the same pixel, or with the
            // upsample folded in (up2) pixel (2 gh + py, 2 gw + px) of the 2H x 2W grid, z = 2 py + px
            const int m_own = kUp ? (int)(((int64_t)gn * 2 * p.a_h + 2 * gh + (z >> 1)) * (2 * p.a_w) + 2 * gw + (z & 1)) : m_in;
            const int rstep = kUp ? 2 : 1;          // output rows between two consecutive accumulator rows of a slab
            // Common case: the whole tile is in bounds and the 32 rows of this warp's slab are
            // consecutive output rows -> phase B needs no shuffles, predicates or divergence.
            const int tn_ = kBlockM >> (p.tw_log2 + p.th_log2);
            const bool tile_full = (tiw * tw + tw <= p.a_w) && (tih * th + th <= p.a_h) && (mt * tn_ + tn_ <= p.a_n);
            const bool affine = tile_full && (tw >= 32 || (!kUp && tw == p.a_w && (tw * th >= 32 || th == p.a_h)));
            const int m_slab = __shfl_sync(0xffffffffu, m_own, 0);

            EpiOut eo;
            eo.resid = partial ? nullptr : p.resid; eo.resid_is_f32 = p.resid_is_f32; eo.ld_resid = p.ld_resid;
            eo.alpha = partial ? 1.0f : p.alpha;
            eo.out_bf16 = partial ? nullptr : p.out_bf16;
            eo.out_f32 = partial ? p.partial + (int64_t)z * p.m_total * p.n_out : p.out_f32;
            eo.ldo = partial ? p.n_out : p.ldo;
            eo.n_cols = geglu ? (p.n_out >> 1) : p.n_out;
            const bool vec_io = (eo.ldo & 3) == 0 && (!eo.resid || (eo.ld_resid & 3) == 0);
            const uint32_t buf = t % kAccBufs, aph = (t / kAccBufs) & 1;
            constexpr int kChunks = BN / 32;
            // chunks this warp owns: half, half+2, ...
            int last_c = -1;
            for (int ci = half; ci < kChunks; ci += kCStride)
                if (col0 + ci * 32 < p.n_out) last_c = ci;
            // phase-B row mapping: row = 4*i + lane/8; its global row index / validity come from the
            // lane that owns that accumulator row (a shuffle is cheaper than 16 live registers here:
            // with 10 warps per CTA one SM sub-partition hosts 3 warps, capping threads at 168 regs)
            // bf16 residuals (VAE stream) are prefetched two chunks deep: the loads of chunk c+1 are in
            // flight while chunk c is drained, and the first chunk's are issued before we even wait
            // for the accumulator (they do not depend on it).  fp32 residuals (UNet stream) are loaded
            // at the top of phase B instead: 64 more live registers cost more than the latency they
            // hide (measured: +0.9 ms per UNet step), with the 168-register cap of a 10-warp CTA.
            auto fast_chunk = [&](int ci) -> bool {
                return !geglu && vec_io && ci <= last_c && col0 + ci * 32 + 32 <= eo.n_cols;
            };
            // (kPrefetchResid is a template flag so the fp32 instantiation carries no prefetch state)
            const bool pre_bf16 = kPrefetchResid && eo.resid && !eo.resid_is_f32;
            auto prefetch = [&](int ci, uint2* dst) {
                if (!pre_bf16 || !fast_chunk(ci)) return;
                const int n = col0 + ci * 32 + 4 * (lane & 7);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = 4 * i + (lane >> 3);
                    const int mr = __shfl_sync(0xffffffffu, m_own, row);
                    const int ok = __shfl_sync(0xffffffffu, row_ok, row);
                    dst[i] = make_uint2(0u, 0u);
                    if (ok)
                        dst[i] = *reinterpret_cast<const uint2*>(
                            reinterpret_cast<const __nv_bfloat16*>(eo.resid) + (int64_t)mr * eo.ld_resid + n);
                }
            };
            uint2 rcur[kPrefetchResid ? 8 : 1], rnext[kPrefetchResid ? 8 : 1];
            if (kPrefetchResid) prefetch(half, rcur);
            float4 rf[kAheadF32 ? 8 : 1];
            const bool ahead_f32 = kAheadF32 && eo.resid && eo.resid_is_f32;
            auto prefetch_f32 = [&](int ci) {
                if (!ahead_f32 || !fast_chunk(ci)) return;
                const int n = col0 + ci * 32 + 4 * (lane & 7);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = 4 * i + (lane >> 3);
                    const int mr = __shfl_sync(0xffffffffu, m_own, row);
                    const int ok = __shfl_sync(0xffffffffu, row_ok, row);
                    rf[kAheadF32 ? i : 0] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok)
                        rf[kAheadF32 ? i : 0] = *reinterpret_cast<const float4*>(
                            reinterpret_cast<const float*>(eo.resid) + (int64_t)mr * eo.ld_resid + n);
                }
            };
            if (kAheadF32) prefetch_f32(half);
            if (sub == 0) {
                mbar_wait(&acc_full[buf], aph);
                tc_fence_after();
            }
            const uint32_t tmem_acc = tmem_base + buf * Cfg::kAccStride + sub * BN + ((uint32_t)(quad * 32) << 16);
            if (last_c < 0 && sub == kMT * kNT - 1) {       // nothing to read: release the buffer right away
                tc_fence_before();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
#pragma unroll 1
            for (int ci = half; ci < kChunks; ci += kCStride) {
                const int c = ci * 32;
                const int nbase = col0 + c;
                if (nbase >= p.n_out) break;                                      // warp-uniform
                if (kPrefetchResid) prefetch(ci + kCStride, rnext);
                // slab fast path (below): its residual is requested now, ahead of the accumulator chunk it will be added to
                const bool use_slab = !geglu && affine && fast_chunk(ci) && !(kPrefetchResid && pre_bf16) &&
                                      !(kAheadF32 && ahead_f32) && (eo.resid || eo.alpha == 1.0f);
                // (fp32 residuals only -- the UNet's residual stream: 28.8 -> 26.9 us on the K = 320 linears; with bf16 residuals --
                // the VAE -- the early request measured 7 % slower, 502 -> 536 us, and they stay at the top of phase B)
                float4 rvp[8];
                if (use_slab && eo.resid && eo.resid_is_f32) {
                    const int64_t r_off = ((int64_t)m_slab + (lane >> 3) * rstep) * eo.ld_resid + nbase + 4 * (lane & 7);
                    load_resid8<1>(reinterpret_cast<const float*>(eo.resid) + r_off, (int64_t)eo.ld_resid * rstep, rvp);
                }
                uint32_t acc[32];
                tmem_ld16(tmem_acc + (uint32_t)c, acc);
                tmem_ld16(tmem_acc + (uint32_t)c + 16, acc + 16);
                const bool full32 = nbase + 32 <= p.n_out;
                tmem_ld_wait();
                if (ci == last_c && sub == kMT * kNT - 1) {  // all TMEM reads of this item by this warp are done
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
                if (!partial) {
                    if (full32) {
                        if (p.bias) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nbase) + j);
                                const float2 lo = __fadd2_rn(make_float2(v[4 * j], v[4 * j + 1]), make_float2(b4.x, b4.y));
                                const float2 hi = __fadd2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(b4.z, b4.w));
                                v[4 * j] = lo.x; v[4 * j + 1] = lo.y; v[4 * j + 2] = hi.x; v[4 * j + 3] = hi.y;
                            }
                        }
                        if (p.row_bias && row_ok) {
                            const float* rb = p.row_bias + (int64_t)gn * p.row_bias_ld + nbase;
#ifdef RDEIC_ROWBIAS_SCALAR
                            if (false) {
#else
                            if (((p.row_bias_ld | nbase) & 3) == 0 && (reinterpret_cast<uintptr_t>(p.row_bias) & 15) == 0) {
#endif
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const float4 r4 = __ldg(reinterpret_cast<const float4*>(rb) + j);
                                    const float2 lo = __fadd2_rn(make_float2(v[4 * j], v[4 * j + 1]), make_float2(r4.x, r4.y));
                                    const float2 hi = __fadd2_rn(make_float2(v[4 * j + 2], v[4 * j + 3]), make_float2(r4.z, r4.w));
                                    v[4 * j] = lo.x; v[4 * j + 1] = lo.y; v[4 * j + 2] = hi.x; v[4 * j + 3] = hi.y;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] += __ldg(rb + j);
                            }
                        }
                    } else {
                        for (int j = 0; j < 32; ++j) {
                            if (nbase + j < p.n_out) {
                                if (p.bias) v[j] += p.bias[nbase + j];
                                if (p.row_bias && row_ok) v[j] += p.row_bias[(int64_t)gn * p.row_bias_ld + nbase + j];
                            }
                        }
                    }
                    if (p.act == 1) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
                    } else if (p.act == 3) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.act_param;
                    } else if (p.act == 4) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float2 g2 = gelu_erf2(make_float2(v[j], v[j + 1]));
                            v[j] = g2.x; v[j + 1] = g2.y;
                        }
                    } else if (geglu) {
                        // columns [0,16) of the chunk are values, [16,32) their gates (weights are
                        // interleaved that way at load): attention.py:54-56  x * gelu(gate)
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
#ifdef RDEIC_GEGLU_ERF
                            const float2 o2 = __fmul2_rn(make_float2(v[j], v[j + 1]), gelu_erf2(make_float2(v[16 + j], v[17 + j])));
#else
                            const float2 o2 = __fmul2_rn(make_float2(v[j], v[j + 1]), gelu_tanh2(make_float2(v[16 + j], v[17 + j])));
#endif
                            v[j] = o2.x; v[j + 1] = o2.y;
                        }
                    }
                }
                const int nchunks = geglu ? 4 : 8;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q < nchunks)
                        stg[lane * 8 + (q ^ (lane & 7))] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                __syncwarp();
                if (!geglu) {
                    if (use_slab) {
                        // GroupNorm statistics of the tensor being written are fused here (kStats) so the
                        // consumer's statistics pass (a full re-read of the tensor) disappears
                        const int q = lane & 7;
                        const int64_t m0 = (int64_t)m_slab + (lane >> 3) * rstep;
                        const int64_t o_off = m0 * eo.ldo + nbase + 4 * q;
                        float* of = eo.out_f32 ? eo.out_f32 + o_off : nullptr;
                        __nv_bfloat16* ob = eo.out_bf16 ? eo.out_bf16 + o_off : nullptr;
                        float* sd = nullptr;
                        if (kStats) {
                            // statistics slab of this warp's 32 rows; up2 orders the slabs [sample][parity][h][w] so
                            // that a sample's slabs stay contiguous for the fold
                            int64_t st_slab = __shfl_sync(0xffffffffu, m_in, 0) >> 5;
                            if (kUp) st_slab += ((int64_t)__shfl_sync(0xffffffffu, gn, 0) * 3 + z) * (((int64_t)p.a_h * p.a_w) >> 5);
                            sd = p.stats_out + (st_slab * p.n_out + nbase + 4 * q) * 2;
                        }
                        const int64_t r_off = m0 * eo.ld_resid + nbase + 4 * q;
                        const int64_t ldo_s = (int64_t)eo.ldo * rstep, ldr_s = (int64_t)eo.ld_resid * rstep;
                        if (!eo.resid)
                            store_slab_out<0, kStats>(stg, lane, eo.alpha, nullptr, 0, of, ob, ldo_s, sd);
                        else if (eo.resid_is_f32)
                            store_slab_out<1, kStats, true>(stg, lane, eo.alpha, reinterpret_cast<const float*>(eo.resid) + r_off,
                                                            ldr_s, of, ob, ldo_s, sd, rvp);
                        else
                            store_slab_out<2, kStats>(stg, lane, eo.alpha,
                                                      reinterpret_cast<const __nv_bfloat16*>(eo.resid) + r_off,
                                                      ldr_s, of, ob, ldo_s, sd);
                    } else if (fast_chunk(ci)) {
                        int mr[8], okr[8];
                        float4 rv[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = 4 * i + (lane >> 3);
                            mr[i] = __shfl_sync(0xffffffffu, m_own, row);
                            okr[i] = __shfl_sync(0xffffffffu, row_ok, row);
                            rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (kAheadF32 && ahead_f32) {
                                rv[i] = rf[kAheadF32 ? i : 0];
                            } else if (eo.resid && okr[i] && !(kPrefetchResid && pre_bf16)) {
                                const int64_t off = (int64_t)mr[i] * eo.ld_resid + nbase + 4 * (lane & 7);
                                if (eo.resid_is_f32) {
                                    rv[i] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(eo.resid) + off);
                                } else {
                                    const uint2 u = *reinterpret_cast<const uint2*>(
                                        reinterpret_cast<const __nv_bfloat16*>(eo.resid) + off);
                                    unpack_bf16x2(u.x, rv[i].x, rv[i].y);
                                    unpack_bf16x2(u.y, rv[i].z, rv[i].w);
                                }
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int row = 4 * i + (lane >> 3), q = lane & 7;
                            float4 val = stg[row * 8 + (q ^ (row & 7))];
                            if (!okr[i]) continue;
                            if (kPrefetchResid && pre_bf16) {
                                unpack_bf16x2(rcur[kPrefetchResid ? i : 0].x, rv[i].x, rv[i].y);
                                unpack_bf16x2(rcur[kPrefetchResid ? i : 0].y, rv[i].z, rv[i].w);
                            }
                            if (eo.resid) {
                                val.x = fmaf(eo.alpha, val.x, rv[i].x); val.y = fmaf(eo.alpha, val.y, rv[i].y);
                                val.z = fmaf(eo.alpha, val.z, rv[i].z); val.w = fmaf(eo.alpha, val.w, rv[i].w);
                            } else if (eo.alpha != 1.0f) {
                                val.x *= eo.alpha; val.y *= eo.alpha; val.z *= eo.alpha; val.w *= eo.alpha;
                            }
                            const int64_t off = (int64_t)mr[i] * eo.ldo + nbase + 4 * q;
                            if (eo.out_f32) *reinterpret_cast<float4*>(eo.out_f32 + off) = val;
                            if (eo.out_bf16)
                                *reinterpret_cast<uint2*>(eo.out_bf16 + off) =
                                    make_uint2(pack_bf16x2(val.x, val.y), pack_bf16x2(val.z, val.w));
                        }
                    } else {
#pragma unroll 1
                        for (int i = 0; i < 8; ++i) {
                            const int row = 4 * i + (lane >> 3), q = lane & 7;
                            const float4 val = stg[row * 8 + (q ^ (row & 7))];
                            const int m_row = __shfl_sync(0xffffffffu, m_own, row);
                            const int ok = __shfl_sync(0xffffffffu, row_ok, row);
                            const int n = nbase + 4 * q;
                            if (ok && n < eo.n_cols) epi_store4(eo, (int64_t)m_row, n, val);
                        }
                    }
                } else if (affine && !kUp && vec_io && eo.out_bf16 && !eo.out_f32 && (nbase >> 1) + 16 <= eo.n_cols) {
                    // GEGLU fast path: full in-bounds tile, consecutive rows, bf16 output only
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = 8 * i + (lane & 7), q = lane >> 3;
                        const float4 val = stg[row * 8 + (q ^ (row & 7))];
                        *reinterpret_cast<uint2*>(eo.out_bf16 + ((int64_t)m_slab + row) * eo.ldo + (nbase >> 1) + 4 * q) =
                            make_uint2(pack_bf16x2(val.x, val.y), pack_bf16x2(val.z, val.w));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = 8 * i + (lane & 7), q = lane >> 3;
                        const float4 val = stg[row * 8 + (q ^ (row & 7))];
                        const int m_row = __shfl_sync(0xffffffffu, m_own, row);
                        const int ok = __shfl_sync(0xffffffffu, row_ok, row);
                        const int n = (nbase >> 1) + 4 * q;
                        if (ok && n < eo.n_cols) epi_store4(eo, (int64_t)m_row, n, val);
                    }
                }
                if (kPrefetchResid) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) rcur[kPrefetchResid ? i : 0] = rnext[kPrefetchResid ? i : 0];
                }
                if (kAheadF32) prefetch_f32(ci + kCStride);     // in flight during the next chunk's TMEM load + phase A
                __syncwarp();
            }
            }   // sub
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// split-K second pass: sum the fp32 partials in a fixed order (deterministic), then the same
// epilogue as the fused path.  One thread per 4 output columns.
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, int splits, int64_t m_total, int n_out,
                     int rows_per_sample, const float* __restrict__ bias,
                     const float* __restrict__ row_bias, int row_bias_ld, int act, float act_param, EpiOut eo) {
    pdl_trigger_short();
    pdl_wait();
    const int n4 = (n_out + 3) >> 2;
    const int64_t total = m_total * n4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t m = i / n4;
        const int n = (int)(i - m * n4) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        for (int z = 0; z < splits; ++z) {
            const float* src = partial + ((int64_t)z * m_total + m) * n_out + n;
            if (n + 4 <= n_out && (n_out & 3) == 0) {
                const float4 t = *reinterpret_cast<const float4*>(src);
                v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
            } else {
                for (int j = 0; j < 4 && n + j < n_out; ++j) v[j] += src[j];
            }
        }
        const int64_t gn = m / rows_per_sample;
        for (int j = 0; j < 4 && n + j < n_out; ++j) {
            if (bias) v[j] += bias[n + j];
            if (row_bias) v[j] += row_bias[gn * row_bias_ld + n + j];
            if (act == 1) v[j] = silu_f(v[j]);
            else if (act == 3) v[j] = v[j] > 0.f ? v[j] : v[j] * act_param;
            else if (act == 4) v[j] = gelu_erf(v[j]);
        }
        epi_store4(eo, m, n, make_float4(v[0], v[1], v[2], v[3]));
    }
}

// ----------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [n_out][taps][cp1 + cp2] bf16 (zero padded)
// ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int n_out,
                   int c1, int c2, int taps, int cp1, int cp2) {
    pdl_trigger();
    pdl_wait();
    const int kp = taps * (cp1 + cp2);
    const int64_t total = (int64_t)n_out * kp;
    const int cin = c1 + c2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / kp);
        int k = (int)(i - (int64_t)n * kp);
        const int tap = k / (cp1 + cp2);
        k -= tap * (cp1 + cp2);
        int c = -1;
        if (k < cp1) { if (k < c1) c = k; }
        else if (k - cp1 < c2) c = c1 + (k - cp1);
        float v = 0.f;
        if (c >= 0) v = w[((int64_t)n * cin + c) * taps + tap];
        dst[i] = __float2bfloat16_rn(v);
    }
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
// Pick the TN x TH x TW box (powers of two, product 128) that covers the [N,H,W] pixel grid
// with the fewest tiles; ties go to the widest TW (longest contiguous TMA rows).
static void pick_m_tile(int N, int H, int W, bool force_tn1, int* tw_o, int* th_o, int* tn_o) {
    int64_t best = -1;
    int btw = 128, bth = 1, btn = 1;
    for (int tw = 128; tw >= 1; tw >>= 1) {
        for (int th = kBlockM / tw; th >= 1; th >>= 1) {
            const int tn = kBlockM / (tw * th);
            if (force_tn1 && tn != 1) continue;
            const int64_t tiles = (int64_t)((W + tw - 1) / tw) * ((H + th - 1) / th) * ((N + tn - 1) / tn);
            if (best < 0 || tiles < best) { best = tiles; btw = tw; bth = th; btn = tn; }
        }
    }
    *tw_o = btw; *th_o = bth; *tn_o = btn;
}
static int ilog2(int x) { int l = 0; while ((1 << l) < x) ++l; return l; }

// Fused GroupNorm statistics need every M tile to be full and "affine" (each epilogue warp's 32
// accumulator rows are 32 consecutive output rows): the same predicate the kernel evaluates per tile.
static bool stats_tiling_ok(int N, int H, int W, bool force_tn1) {
    int tw, th, tn;
    pick_m_tile(N, H, W, force_tn1, &tw, &th, &tn);
    if (W % tw || H % th || N % tn) return false;
    return tw >= 32 || (tw == W && (tw * th >= 32 || th == H));
}

static inline int host_total_kb(const ConvDev& d) {
    return d.a2_center ? d.taps * d.cblk1 + d.cblk2 : d.taps * (d.cblk1 + d.cblk2);
}

template <int BN, int kPre, int kEW, bool kStats = false, int kMT = 1, bool kSwap = false, bool kUp = false, int kIss = 1, int kNT = 1>
static int launch_conv3(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const CUtensorMap& tb2,
                        const ConvDev& d, int m_tiles, int splits, cudaStream_t s) {
    using Cfg = TileCfg<BN, kEW, kMT, kIss, kNT>;
    static_assert(Cfg::kStages >= 2, "pipeline needs at least two stages");
    static bool attr_set = false;
    if (!attr_set) {
        RDEIC_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, kPre, kEW, kStats, kMT, kSwap, kUp, kIss, kNT>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        attr_set = true;
    }
    const int64_t items = (int64_t)((m_tiles + kMT - 1) / kMT) * d.n_tiles * splits;
    dim3 grid((unsigned)(items < kNumSMs ? items : kNumSMs));
    RDEIC_CUDA(launch_k(conv_gemm_kernel<BN, kPre, kEW, kStats, kMT, kSwap, kUp, kIss, kNT>, grid, Cfg::kThreads, Cfg::kSmemBytes, s, ta, ta2, tb, tb2, d));
    RDEIC_LAUNCH_CHECK();
    return 0;
}

template <int BN>
static int launch_conv(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const CUtensorMap& tb2,
                       const ConvDev& d, int m_tiles, int splits, cudaStream_t s) {
    if (d.up2) {              // nearest-upsample folded in: its own instantiations (8 epilogue warps)
        if (d.stats_out) return launch_conv3<BN, 0, 8, true, 1, false, true>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
        return launch_conv3<BN, 0, 8, false, 1, false, true>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
    }
    // Residual prefetch-ahead instantiations (1: bf16 two chunks deep, 2: fp32 one phase ahead) are
    // kept for experiments only: with the affine fast path both measured slower than loading the
    // residual at the top of phase B (VAE 128-ch conv + bf16 residual: 1335 us vs 681 us).
    // N <= 128 with a long reduction and many tiles: two M tiles per item share each weight stage
    if constexpr (BN == 128) {
        static const bool dual_m = !(getenv("RDEIC_DUAL_M") && atoi(getenv("RDEIC_DUAL_M")) == 0);
        if (dual_m && splits == 1 && !d.w_batched && !d.a2_center && d.in_stride == 1 && host_total_kb(d) >= 9 &&
            (int64_t)m_tiles * d.n_tiles >= 8 * kNumSMs) {
            // exchanged operand roles when the epilogue's fast path covers the whole problem
            static const bool swap_ok = !(getenv("RDEIC_SWAP") && atoi(getenv("RDEIC_SWAP")) == 0);
            const bool vec_io = (d.ldo & 3) == 0 && (!d.resid || (d.ld_resid & 3) == 0);
            if (swap_ok && d.n_tiles == 1 && !d.row_bias && d.act != 2 && d.n_out % 32 == 0 && vec_io &&
                (d.resid || d.alpha == 1.0f) && stats_tiling_ok(d.a_n, d.a_h, d.a_w, false)) {
                if (d.stats_out) return launch_conv3<BN, 0, 8, true, 2, true>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
                return launch_conv3<BN, 0, 8, false, 2, true>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
            }
            if (d.stats_out) return launch_conv3<BN, 0, 8, true, 2>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
            return launch_conv3<BN, 0, 8, false, 2>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
        }
    }
    // N = 160 tiles (UNet levels 0 / 1: 320 and 640 output channels) with a long reduction: two M tiles per item, one
    // issuing warp each.  Per k-block the pipe then needs 8 x 80 = 640 clocks for two tiles where one issuer took 452 per tile.
    // Items are twice as large, so it pays when the one-tile schedule needs at least two rounds of CTAs:
    //   rounds(units / 2) * 640 < rounds(units) * 452
    // N = 160 tiles with a long reduction and at least two rounds of tiles: both N tiles of a pair in one item (kNT = 2)
    if constexpr (BN == 160) {
        static const int pair160 = getenv("RDEIC_PAIR160") ? atoi(getenv("RDEIC_PAIR160")) : 0;   // off: measured 63.6 -> 74.9 us at UNet level 0 (single accumulator buffer, three stages)
        const int64_t units = (int64_t)m_tiles * d.n_tiles;
        const int64_t rounds1 = (units + kNumSMs - 1) / kNumSMs;
        const int64_t rounds2 = ((int64_t)m_tiles * ((d.n_tiles + 1) / 2) + kNumSMs - 1) / kNumSMs;
        if (pair160 && splits == 1 && !d.w_batched && d.n_tiles >= 2 && host_total_kb(d) >= 18 && rounds2 < rounds1) {
            ConvDev d2 = d;
            d2.n_tiles = (d.n_tiles + 1) / 2;
            d2.m_major = 0;
            if (d.stats_out) return launch_conv3<BN, 0, 8, true, 1, false, false, 1, 2>(ta, ta2, tb, tb2, d2, m_tiles, splits, s);
            return launch_conv3<BN, 0, 8, false, 1, false, false, 1, 2>(ta, ta2, tb, tb2, d2, m_tiles, splits, s);
        }
    }
    if constexpr (BN == 160) {
        // (off by default: built when one thread was believed to issue at most one tcgen05.mma per ~100 clocks; that was the
        // ELECT / BRA.U.ANY wrapper of a `lane == 0` branch -- see elect_one_sync() -- and with it gone one issuer keeps the
        // pipe full, while this variant pays a single-buffered accumulator and three smem stages: 64 -> 76 us at level 0)
        static const int dual160 = getenv("RDEIC_DUAL160") ? atoi(getenv("RDEIC_DUAL160")) : 0;
        const int64_t units = (int64_t)m_tiles * d.n_tiles;
        const int64_t r1 = (units + kNumSMs - 1) / kNumSMs, r2 = ((int64_t)((m_tiles + 1) / 2) * d.n_tiles + kNumSMs - 1) / kNumSMs;
        if (dual160 && splits == 1 && !d.w_batched && !d.a2_center && d.in_stride == 1 && host_total_kb(d) >= 18 &&
            m_tiles >= 2 && r2 * 640 < r1 * 452) {
            if (d.stats_out) return launch_conv3<BN, 0, 8, true, 2, false, false, 2>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
            return launch_conv3<BN, 0, 8, false, 2, false, false, 2>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
        }
    }
    if (d.stats_out) {
        static const int epi12_kb_s = getenv("RDEIC_EPI12_KB") ? atoi(getenv("RDEIC_EPI12_KB")) : 10;
        if (host_total_kb(d) <= epi12_kb_s) return launch_conv3<BN, 0, 12, true>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
        return launch_conv3<BN, 0, 8, true>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
    }
    static const bool pre16 = getenv("RDEIC_RESID_PREFETCH") != nullptr;
    if (pre16 && d.resid && !d.resid_is_f32 && !d.partial) return launch_conv3<BN, 1, 8>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
    static const bool ahead = getenv("RDEIC_RESID_AHEAD") != nullptr;
    if (ahead && d.resid && d.resid_is_f32 && !d.partial) return launch_conv3<BN, 2, 8>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
    // few k-blocks per tile -> the epilogue, not the tensor pipe, sets the pace: use 12 epilogue warps
    static const int epi12_kb = getenv("RDEIC_EPI12_KB") ? atoi(getenv("RDEIC_EPI12_KB")) : 10;
    const int total_kb = host_total_kb(d);
    if (total_kb <= epi12_kb) return launch_conv3<BN, 0, 12>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
    return launch_conv3<BN, 0, 8>(ta, ta2, tb, tb2, d, m_tiles, splits, s);
}

// Tile width: cost ~ tiles per SM * per-tile time, per-tile time ~ bn (MMA N) + a fixed overhead (pipeline
// fill, epilogue tail: ~96 columns' worth, fitted on the K <= 1280 linears of UNet levels 1-2).  A model that
// charged the main loop the same for every width was A/B-tested on one box and lost: UNet step 12.94 -> 13.60 ms,
// VAE 22.0 -> 23.8 ms.  The constants were swept again after the issue-rate fix (scripts/r02_s28.sh): unchanged.
static int pick_block_n(int n_out, int m_tiles, int hint, int total_kb) {
    (void)total_kb;
    if (hint == 32 || hint == 64 || hint == 128 || hint == 160 || hint == 256) return hint;
    if (n_out <= 32) return 32;
    if (n_out <= 64) return 64;
    static const int overhead = getenv("RDEIC_TILE_OVERHEAD") ? atoi(getenv("RDEIC_TILE_OVERHEAD")) : 96;
    const int cands[3] = {256, 160, 128};
    int best = 128;
    double best_cost = 1e30;
    for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        const int nt = (n_out + bn - 1) / bn;
        const int64_t tiles = (int64_t)m_tiles * nt;
        const int64_t per_sm = (tiles + kNumSMs - 1) / kNumSMs;      // persistent: tiles each CTA walks
        const double cost = (double)per_sm * (bn + overhead);
        if (cost < best_cost) { best_cost = cost; best = bn; }
    }
    return best;
}

}  // namespace rdeic

using namespace rdeic;

extern "C" {

int rdeic_conv_stats_supported(int a_n, int a_h, int a_w, int n_out, int k_blocks) {
    if (a_n <= 0 || a_h <= 0 || a_w <= 0 || n_out <= 0 || n_out % 32) return 0;
    if (!stats_tiling_ok(a_n, a_h, a_w, false)) return 0;
    // layers that rdeic_conv_gemm would run split-K (few tiles, long reduction) keep split-K: their
    // epilogue lives in the reduce kernel and the tensors are small enough for a statistics pass
    int tw, th, tn;
    pick_m_tile(a_n, a_h, a_w, false, &tw, &th, &tn);
    const int64_t m_tiles = (int64_t)((a_w + tw - 1) / tw) * ((a_h + th - 1) / th) * ((a_n + tn - 1) / tn);
    const int bn = pick_block_n(n_out, (int)m_tiles, 0, k_blocks);
    const int64_t tiles = m_tiles * ((n_out + bn - 1) / bn);
    if (tiles <= kNumSMs / 2 && k_blocks >= 8) return 0;
    return 1;
}

int rdeic_pack_conv_weight(const float* w_oihw, void* dst, int n_out, int c1, int c2, int kh,
                           int kw, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(w_oihw && dst, "rdeic_pack_conv_weight: null pointer");
    RDEIC_CHECK_ARG(n_out > 0 && c1 > 0 && c2 >= 0, "rdeic_pack_conv_weight: bad channel counts");
    RDEIC_CHECK_ARG((kh == 1 && kw == 1) || (kh == 2 && kw == 2) || (kh == 3 && kw == 3) || (kh == 5 && kw == 5),
                    "rdeic_pack_conv_weight: only 1x1, 2x2 (folded upsample), 3x3 and 5x5 kernels are supported (got %dx%d)", kh, kw);
    const int taps = kh * kw;
    const int cp1 = (c1 + 63) / 64 * 64, cp2 = (c2 + 63) / 64 * 64;
    const int64_t total = (int64_t)n_out * taps * (cp1 + cp2);
    launch_k(pack_weight_kernel, grid_for(total, 256), 256, 0, as_stream(stream), 
        w_oihw, (__nv_bfloat16*)dst, n_out, c1, c2, taps, cp1, cp2);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_conv_gemm(const rdeic_conv_params* p, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(p, "rdeic_conv_gemm: null params");
    RDEIC_CHECK_ARG(p->a && p->w, "rdeic_conv_gemm: null operand");
    RDEIC_CHECK_ARG(p->out_bf16 || p->out_f32, "rdeic_conv_gemm: no output pointer");
    RDEIC_CHECK_ARG(p->a_n > 0 && p->a_h > 0 && p->a_w > 0 && p->a_c > 0, "rdeic_conv_gemm: empty A");
    RDEIC_CHECK_ARG(p->a_c % 8 == 0 && p->a2_c % 8 == 0,
                    "rdeic_conv_gemm: channel counts (%d, %d) must be multiples of 8 (TMA 16-byte strides)",
                    p->a_c, p->a2_c);
    RDEIC_CHECK_ARG(p->a2_c == 0 || p->a2, "rdeic_conv_gemm: a2_c > 0 needs a2");
    RDEIC_CHECK_ARG(p->taps == 1 || p->taps == 9 || p->taps == 25 || (p->taps == 4 && p->up2),
                    "rdeic_conv_gemm: taps must be 1, 9 or 25 (4 with up2)");
    RDEIC_CHECK_ARG(p->up2 == 0 || (p->up2 == 1 && p->taps == 4 && p->w_batch_stride > 0 && (p->a2_c == 0 || p->a2_center) &&
                                    p->act != 2 && p->w_k == 0 && p->in_stride2 == 0),
                    "rdeic_conv_gemm: up2 needs taps = 4, the four parity weight matrices w_batch_stride apart, no concat source, no GEGLU");
    RDEIC_CHECK_ARG(p->a2_center == 0 || (p->a2_c > 0 && p->w2 && (uintptr_t)p->w2 % 16 == 0 && p->w_k == 0 &&
                                          (p->w_batch_stride == 0 || p->up2)),
                    "rdeic_conv_gemm: a2_center needs a second source, its packed weights w2 and a packed w");
    RDEIC_CHECK_ARG(p->in_stride2 == 0 || ((p->taps == 1 || p->taps == 9) && p->w_batch_stride == 0 && (p->pad_lo == 0 || p->pad_lo == 1)),
                    "rdeic_conv_gemm: in_stride2 needs a 1x1 or 3x3 filter and pad_lo in {0, 1}");
    const int a_ld = p->a_ld ? p->a_ld : p->a_c, a2_ld = p->a2_ld ? p->a2_ld : p->a2_c;
    RDEIC_CHECK_ARG(a_ld >= p->a_c && a2_ld >= p->a2_c && a_ld % 8 == 0 && a2_ld % 8 == 0,
                    "rdeic_conv_gemm: a_ld/a2_ld (%d, %d) must be multiples of 8 and >= the channel counts", a_ld, a2_ld);
    RDEIC_CHECK_ARG(p->n_out > 0 && p->ldo > 0, "rdeic_conv_gemm: bad n_out/ldo");
    RDEIC_CHECK_ARG(p->act >= 0 && p->act <= 4,
                    "rdeic_conv_gemm: act must be 0 (none), 1 (SiLU), 2 (GEGLU), 3 (LeakyReLU) or 4 (GELU)");
    RDEIC_CHECK_ARG(((uintptr_t)p->a | (uintptr_t)p->a2 | (uintptr_t)p->w) % 16 == 0,
                    "rdeic_conv_gemm: operands must be 16-byte aligned");
    RDEIC_CHECK_ARG(((uintptr_t)p->out_bf16 | (uintptr_t)p->out_f32 | (uintptr_t)p->resid |
                     (uintptr_t)p->bias) % 16 == 0,
                    "rdeic_conv_gemm: epilogue pointers must be 16-byte aligned");
    RDEIC_CHECK_ARG(!p->resid || p->ld_resid >= p->n_out, "rdeic_conv_gemm: bad ld_resid");
    RDEIC_CHECK_ARG(!p->row_bias || p->row_bias_ld >= p->n_out, "rdeic_conv_gemm: bad row_bias_ld");
    RDEIC_CHECK_ARG(p->w_k >= 0 && p->w_ld >= 0 && p->w_k % 8 == 0 && p->w_ld % 8 == 0 &&
                        (p->w_k == 0 || (p->taps == 1 && p->a2_c == 0 && p->w_k == p->a_c)),
                    "rdeic_conv_gemm: w_k/w_ld describe an unpacked B operand: taps=1, one source, w_k == a_c, multiples of 8");
    RDEIC_CHECK_ARG(p->w_batch_stride % 8 == 0, "rdeic_conv_gemm: w_batch_stride must be a multiple of 8");

    ConvDev d;
    d.a_n = p->a_n; d.a_h = p->a_h; d.a_w = p->a_w;
    int tw_eff, th_eff, tn;
    const bool w_batched = p->w_batch_stride != 0 && !p->up2;
    pick_m_tile(p->a_n, p->a_h, p->a_w, w_batched, &tw_eff, &th_eff, &tn);
    d.tw_log2 = ilog2(tw_eff); d.th_log2 = ilog2(th_eff);
    d.tiles_w = (p->a_w + tw_eff - 1) / tw_eff;
    d.tiles_h = (p->a_h + th_eff - 1) / th_eff;
    d.tiles_n = (p->a_n + tn - 1) / tn;
    const int64_t m_tiles64 = (int64_t)d.tiles_w * d.tiles_h * d.tiles_n;
    RDEIC_CHECK_ARG(m_tiles64 < (1ll << 31), "rdeic_conv_gemm: too many M tiles");
    const int m_tiles = (int)m_tiles64;
    d.cblk1 = (p->a_c + 63) / 64;
    d.cblk2 = (p->a2_c + 63) / 64;
    d.taps = p->taps;
    d.n_out = p->n_out;
    d.w_batched = w_batched;
    d.up2 = p->up2; d.a2_center = p->a2_center; d.in_stride = p->in_stride2 ? 2 : 1;
    d.tap_lo = (p->in_stride2 && p->pad_lo == 0) ? 0 : -1;
    d.bias = p->bias; d.row_bias = p->row_bias; d.row_bias_ld = p->row_bias_ld;
    d.resid = p->resid; d.resid_is_f32 = p->resid_is_f32; d.ld_resid = p->ld_resid;
    d.alpha = p->alpha; d.act = p->act; d.act_param = p->act_param;
    d.out_bf16 = (__nv_bfloat16*)p->out_bf16; d.out_f32 = p->out_f32; d.ldo = p->ldo;
    d.partial = nullptr; d.kb_per_split = 0; d.splits = 1; d.n_tiles = 0; d.m_major = 0;
    d.stats_out = p->stats_out;
    if (p->stats_out) {
        RDEIC_CHECK_ARG(p->act != 2 && (p->w_batch_stride == 0 || p->up2) && p->n_out % 32 == 0 && p->ldo % 4 == 0 &&
                            (!p->resid || p->ld_resid % 4 == 0) && (uintptr_t)p->stats_out % 16 == 0,
                        "rdeic_conv_gemm: stats_out needs a plain conv/linear with n_out %% 32 == 0 and 16-byte rows");
        RDEIC_CHECK_ARG(stats_tiling_ok(p->a_n, p->a_h, p->a_w, false),
                        "rdeic_conv_gemm: stats_out is not supported for a %dx%dx%d pixel grid "
                        "(ask rdeic_conv_stats_supported first)", p->a_n, p->a_h, p->a_w);
        RDEIC_CHECK_ARG(!p->up2 || tw_eff >= 32, "rdeic_conv_gemm: stats_out with up2 needs an input width that is a multiple of 32");
    }
    d.m_total = (int64_t)p->a_n * p->a_h * p->a_w;
    if (p->act == 2) {
        RDEIC_CHECK_ARG(p->n_out % 32 == 0 && !p->resid && p->ldo >= p->n_out / 2,
                        "rdeic_conv_gemm: GEGLU epilogue needs n_out %% 32 == 0, no residual, ldo >= n_out/2");
    } else {
        RDEIC_CHECK_ARG(p->ldo >= p->n_out, "rdeic_conv_gemm: bad n_out/ldo");
    }

    const int bn = pick_block_n(p->n_out, p->up2 ? 4 * m_tiles : m_tiles, p->tile_n_hint, host_total_kb(d));

    CUtensorMap ta, ta2, tb, tb2;
    {
        // stride-2 conv: the map describes the [a_n, 2 a_h, 2 a_w] input and is walked with element strides 2 along
        // W and H (a box of 2 TW x 2 TH elements delivers TW x TH pixels); the tap offset moves the box origin
        const uint64_t cs = (uint64_t)d.in_stride;
        const uint64_t iw = (uint64_t)p->a_w * cs, ih = (uint64_t)p->a_h * cs;
        uint64_t dims[4] = {(uint64_t)p->a_c, iw, ih, (uint64_t)p->a_n};
        uint64_t str[3] = {(uint64_t)a_ld * 2, (uint64_t)a_ld * 2 * iw, (uint64_t)a_ld * 2 * iw * ih};
        uint32_t box[4] = {(uint32_t)kBlockK, (uint32_t)(tw_eff * cs), (uint32_t)(th_eff * cs), (uint32_t)tn};
        uint32_t estr[4] = {1, (uint32_t)cs, (uint32_t)cs, 1};
        if (int e = encode_map(&ta, p->a, 4, dims, str, box, "A", estr)) return e;
        ta2 = ta;
        if (p->a2_c) {
            uint64_t dims2[4] = {(uint64_t)p->a2_c, iw, ih, (uint64_t)p->a_n};
            uint64_t str2[3] = {(uint64_t)a2_ld * 2, (uint64_t)a2_ld * 2 * iw, (uint64_t)a2_ld * 2 * iw * ih};
            if (d.a2_center) {           // always on the output grid: unit element strides, or (up2) every other pixel of 2H x 2W
                const uint64_t us = d.up2 ? 2 : 1;
                const uint64_t ow = (uint64_t)p->a_w * us, oh = (uint64_t)p->a_h * us;
                uint64_t dims2c[4] = {(uint64_t)p->a2_c, ow, oh, (uint64_t)p->a_n};
                uint64_t str2c[3] = {(uint64_t)a2_ld * 2, (uint64_t)a2_ld * 2 * ow, (uint64_t)a2_ld * 2 * ow * oh};
                uint32_t box2c[4] = {(uint32_t)kBlockK, (uint32_t)(tw_eff * us), (uint32_t)(th_eff * us), (uint32_t)tn};
                uint32_t estr2c[4] = {1, (uint32_t)us, (uint32_t)us, 1};
                if (int e = encode_map(&ta2, p->a2, 4, dims2c, str2c, box2c, "A2", estr2c)) return e;
            } else if (int e = encode_map(&ta2, p->a2, 4, dims2, str2, box, "A2", estr)) return e;
        }
        const uint64_t kp = (uint64_t)(d.a2_center ? d.taps * d.cblk1 : host_total_kb(d)) * kBlockK;
        const uint64_t nb = d.up2 ? 4 : (d.w_batched ? (uint64_t)p->a_n : 1);
        const uint64_t wk = p->w_k > 0 ? (uint64_t)p->w_k : kp;      // true K extent (OOB -> 0)
        const uint64_t wld = p->w_ld > 0 ? (uint64_t)p->w_ld : kp;
        uint64_t dimsb[3] = {wk, (uint64_t)p->n_out, nb};
        uint64_t strb[2] = {wld * 2, nb > 1 ? (uint64_t)p->w_batch_stride * 2 : wld * 2 * (uint64_t)p->n_out};
        uint32_t boxb[3] = {(uint32_t)kBlockK, (uint32_t)bn, 1};
        if (int e = encode_map(&tb, p->w, 3, dimsb, strb, boxb, "W")) return e;
        tb2 = tb;
        if (d.a2_center) {
            const uint64_t kp2 = (uint64_t)d.cblk2 * kBlockK;
            uint64_t dimsb2[3] = {kp2, (uint64_t)p->n_out, 1};
            uint64_t strb2[2] = {kp2 * 2, kp2 * 2 * (uint64_t)p->n_out};
            if (int e = encode_map(&tb2, p->w2, 3, dimsb2, strb2, boxb, "W2")) return e;
        }
    }
    cudaStream_t s = as_stream(stream);
    // split-K: layers with few output tiles but a long reduction (UNet levels 2/3: M = 512..2048,
    // K up to 23040) would leave most of the 148 SMs idle; slice K across blockIdx.z, write fp32
    // partials to the caller's workspace and finish with a deterministic reduce + epilogue pass.
    int splits = 1;
    const int total_kb = host_total_kb(d);
    const int n_tiles = (p->n_out + bn - 1) / bn;
    d.n_tiles = n_tiles;
    d.kb_per_split = total_kb;
    const int64_t tiles = (int64_t)m_tiles * n_tiles;
    if (d.up2) {
        splits = 4;                    // the z index of a work item is the output parity class, not a k slice
        d.splits = 4;
    } else if (p->act != 2 && !p->stats_out && p->workspace && tiles <= kNumSMs / 2 && total_kb >= 8) {
        int want = (int)(kNumSMs / tiles);
        if (want > total_kb / 4) want = total_kb / 4;
        if (want > 16) want = 16;
        const int64_t per_split = d.m_total * p->n_out * (int64_t)sizeof(float);
        while (want > 1 && want * per_split > p->workspace_bytes) --want;
        if (want >= 2) {
            d.kb_per_split = (total_kb + want - 1) / want;
            splits = (total_kb + d.kb_per_split - 1) / d.kb_per_split;
            d.partial = reinterpret_cast<float*>(p->workspace);
            d.splits = splits;
        }
    }
    {
        static const int m_major_env = getenv("RDEIC_M_MAJOR") ? atoi(getenv("RDEIC_M_MAJOR")) : 0;
        // 1: whenever there are several N tiles and several waves of items; 2: only when the N tiles divide the grid evenly
        if (m_major_env && n_tiles > 1 && tiles > kNumSMs && (m_major_env == 1 || kNumSMs % n_tiles == 0)) d.m_major = 1;
    }
    int rc;
    switch (bn) {
        case 32: rc = launch_conv<32>(ta, ta2, tb, tb2, d, m_tiles, splits, s); break;
        case 64: rc = launch_conv<64>(ta, ta2, tb, tb2, d, m_tiles, splits, s); break;
        case 128: rc = launch_conv<128>(ta, ta2, tb, tb2, d, m_tiles, splits, s); break;
        case 160: rc = launch_conv<160>(ta, ta2, tb, tb2, d, m_tiles, splits, s); break;
        case 256: rc = launch_conv<256>(ta, ta2, tb, tb2, d, m_tiles, splits, s); break;
        default: return set_error("rdeic_conv_gemm: unsupported BLOCK_N %d", bn);
    }
    if (rc) return rc;
    if (splits > 1 && !d.up2) {
        EpiOut eo;
        eo.resid = p->resid; eo.resid_is_f32 = p->resid_is_f32; eo.ld_resid = p->ld_resid; eo.alpha = p->alpha;
        eo.out_bf16 = (__nv_bfloat16*)p->out_bf16; eo.out_f32 = p->out_f32; eo.ldo = p->ldo; eo.n_cols = p->n_out;
        const int64_t work = d.m_total * ((p->n_out + 3) / 4);
        launch_k(splitk_reduce_kernel, grid_for(work, 256), 256, 0, s, 
            d.partial, splits, d.m_total, p->n_out, p->a_h * p->a_w, p->bias, p->row_bias, p->row_bias_ld,
            p->act, p->act_param, eo);
        RDEIC_LAUNCH_CHECK();
    }
    return 0;
}

}  // extern "C"
