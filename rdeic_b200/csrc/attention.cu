// Fused flash-style attention forward: softmax(Q K^T * scale) V without materialising the
// [heads, Nq, Nk] logit matrix the reference builds (ldm/modules/attention.py:171-203).
//
// v1 data path: K/V tiles staged in shared memory with cp.async (double-buffered), Q held in
// registers, S = Q K^T and O += P V on warp-level mma.sync m16n8k16 bf16 tensor-core
// instructions with fp32 accumulation and an online softmax in fp32 (exp2 domain).
// Two head widths are instantiated: 64 (SD-2.1 UNet) and 16 (control adapter, rdeic.yaml:45).
// Cross-attention (Nk = 77 text tokens) runs through the same kernel with a masked tail.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/rdeic_b200.h"

namespace rdeic {

constexpr int kAttnWarps = 4;
constexpr int kQTile = kAttnWarps * 16;   // 64 query rows per CTA
constexpr int kKTile = 64;                // keys per pipeline stage

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void mma_bf16_16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                 : "=r"(r0), "=r"(r1)
                 : "r"(s));
}

__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int D>
__global__ void __launch_bounds__(kAttnWarps * 32)
attention_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                 const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out, int Nq,
                 int Nk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t q_bs,
                 int64_t k_bs, int64_t v_bs, int64_t o_bs, float scale_log2) {
    pdl_trigger();
    pdl_wait();
    constexpr int kPitch = D + 8;                 // bf16 elements per smem row (bank spread)
    constexpr int kVecPerRow = D / 8;             // 16-byte vectors per K/V row
    constexpr int kKSteps = D / 16;               // k-steps of the QK^T mma
    constexpr int kDTiles = D / 8;                // n8 tiles of the output
    __shared__ __align__(16) __nv_bfloat16 s_k[2][kKTile][kPitch];
    __shared__ __align__(16) __nv_bfloat16 s_v[2][kKTile][kPitch];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int head = blockIdx.y, b = blockIdx.z;
    const int q0 = blockIdx.x * kQTile + warp * 16;

    const __nv_bfloat16* qb = q + (int64_t)b * q_bs + (int64_t)head * D;
    const __nv_bfloat16* kb = k + (int64_t)b * k_bs + (int64_t)head * D;
    const __nv_bfloat16* vb = v + (int64_t)b * v_bs + (int64_t)head * D;
    __nv_bfloat16* ob = out + (int64_t)b * o_bs + (int64_t)head * D;

    // Q fragments (A operand, row-major 16 x D per warp), rows clamped for the ragged tail
    uint32_t qa[kKSteps][4];
    {
        const int r0 = min(q0 + g, Nq - 1), r1 = min(q0 + g + 8, Nq - 1);
        const __nv_bfloat16* p0 = qb + (int64_t)r0 * ldq;
        const __nv_bfloat16* p1 = qb + (int64_t)r1 * ldq;
#pragma unroll
        for (int kk = 0; kk < kKSteps; ++kk) {
            qa[kk][0] = *reinterpret_cast<const uint32_t*>(p0 + 16 * kk + 2 * t);
            qa[kk][1] = *reinterpret_cast<const uint32_t*>(p1 + 16 * kk + 2 * t);
            qa[kk][2] = *reinterpret_cast<const uint32_t*>(p0 + 16 * kk + 8 + 2 * t);
            qa[kk][3] = *reinterpret_cast<const uint32_t*>(p1 + 16 * kk + 8 + 2 * t);
        }
    }

    auto load_tile = [&](int tile, int buf) {
        const int key0 = tile * kKTile;
        for (int i = threadIdx.x; i < kKTile * kVecPerRow; i += kAttnWarps * 32) {
            const int r = i / kVecPerRow, c = i - r * kVecPerRow;
            const int key = key0 + r;
            const int ok = key < Nk ? 16 : 0;           // zero-fill rows past the end
            const int keyc = key < Nk ? key : Nk - 1;
            cp_async16(&s_k[buf][r][c * 8], kb + (int64_t)keyc * ldk + c * 8, ok);
            cp_async16(&s_v[buf][r][c * 8], vb + (int64_t)keyc * ldv + c * 8, ok);
        }
    };

    float o[kDTiles][4];
#pragma unroll
    for (int j = 0; j < kDTiles; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY};
    float l_run[2] = {0.f, 0.f};

    const int num_tiles = (Nk + kKTile - 1) / kKTile;
    load_tile(0, 0);
    cp_async_commit();
    for (int tile = 0; tile < num_tiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < num_tiles) {
            load_tile(tile + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        // S = Q K^T  (16 x 64 per warp)
        float s[kKTile / 8][4];
#pragma unroll
        for (int j = 0; j < kKTile / 8; ++j) {
            s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
            for (int kk = 0; kk < kKSteps; ++kk) {
                const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&s_k[buf][8 * j + g][16 * kk + 2 * t]);
                const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&s_k[buf][8 * j + g][16 * kk + 8 + 2 * t]);
                mma_bf16_16816(s[j], qa[kk], b0, b1);
            }
        }
        // mask the ragged key tail; row maxima on the raw logits (scale_log2 > 0, so the maximum commutes
        // with the scaling and the scaling itself folds into the FFMA that feeds the exponential)
        const int key_base = tile * kKTile + 2 * t;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < kKTile / 8; ++j) {
            const int key = key_base + 8 * j;
            const bool ok0 = key < Nk, ok1 = key + 1 < Nk;
            s[j][0] = ok0 ? s[j][0] : -INFINITY;
            s[j][1] = ok1 ? s[j][1] : -INFINITY;
            s[j][2] = ok0 ? s[j][2] : -INFINITY;
            s[j][3] = ok1 ? s[j][3] : -INFINITY;
            mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
            mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        }
        float corr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float m_new = fmaxf(m_run[r], mx[r] * scale_log2);   // finite: every tile has >= 1 valid key
            corr[r] = ex2_ftz(m_run[r] - m_new);
            m_run[r] = m_new;
            l_run[r] *= corr[r];
        }
#pragma unroll
        for (int j = 0; j < kDTiles; ++j) {
            o[j][0] *= corr[0]; o[j][1] *= corr[0];
            o[j][2] *= corr[1]; o[j][3] *= corr[1];
        }
        // P = exp2(S * scale - m), packed straight into A fragments for P V.  One MUFU per element
        // (exp2f adds a range test and two multiplies for denormal results, which a probability does
        // not need): the loop is bound by the MUFU pipe instead of by instruction issue.
        uint32_t pa[kKTile / 16][4];
#pragma unroll
        for (int j = 0; j < kKTile / 8; ++j) {
            const float p0 = ex2_ftz(fmaf(s[j][0], scale_log2, -m_run[0]));
            const float p1 = ex2_ftz(fmaf(s[j][1], scale_log2, -m_run[0]));
            const float p2 = ex2_ftz(fmaf(s[j][2], scale_log2, -m_run[1]));
            const float p3 = ex2_ftz(fmaf(s[j][3], scale_log2, -m_run[1]));
            l_run[0] += p0 + p1;
            l_run[1] += p2 + p3;
            pa[j >> 1][(j & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pa[j >> 1][(j & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
        // O += P V
#pragma unroll
        for (int kk = 0; kk < kKTile / 16; ++kk) {
#pragma unroll
            for (int jd = 0; jd < kDTiles; ++jd) {
                uint32_t b0, b1;
                ldmatrix_x2_trans(b0, b1, &s_v[buf][16 * kk + (lane & 15)][8 * jd]);
                mma_bf16_16816(o[jd], pa[kk], b0, b1);
            }
        }
        __syncthreads();   // everyone done with `buf` before it is refilled two tiles later
    }

#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
    const int row0 = q0 + g, row1 = q0 + g + 8;
#pragma unroll
    for (int jd = 0; jd < kDTiles; ++jd) {
        if (row0 < Nq)
            *reinterpret_cast<uint32_t*>(ob + (int64_t)row0 * ldo + 8 * jd + 2 * t) =
                pack_bf16x2(o[jd][0] * inv0, o[jd][1] * inv0);
        if (row1 < Nq)
            *reinterpret_cast<uint32_t*>(ob + (int64_t)row1 * ldo + 8 * jd + 2 * t) =
                pack_bf16x2(o[jd][2] * inv1, o[jd][3] * inv1);
    }
}

// tcgen05 / TMEM path (attention_tc.cu)
bool attention_tc_supported(int d, int Nq, int Nk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                            int64_t q_bs, int64_t k_bs, int64_t v_bs, const void* q, const void* k, const void* v,
                            const void* out);
int launch_attention_tc(const void* q, const void* k, const void* v, void* out, int B, int heads, int Nq, int Nk,
                        int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t q_bs, int64_t k_bs,
                        int64_t v_bs, int64_t o_bs, float scale, cudaStream_t stream);

bool attention_x_supported(int d, int Nq, int Nk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                           int64_t q_bs, int64_t k_bs, int64_t v_bs, int64_t o_bs, const void* q, const void* k, const void* v,
                           const void* out);
int launch_attention_x(const void* q, const void* k, const void* v, void* out, int B, int heads, int Nq, int Nk,
                       int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t q_bs, int64_t k_bs,
                       int64_t v_bs, int64_t o_bs, float scale, cudaStream_t stream);

bool attention_wide_supported(int d, int Nq, int Nk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t q_bs,
                              int64_t k_bs, int64_t v_bs, int64_t o_bs, const void* q, const void* k, const void* v,
                              const void* out);
int launch_attention_wide(const void* q, const void* k, const void* v, void* out, int B, int heads, int Nq, int Nk, int d,
                          int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t q_bs, int64_t k_bs, int64_t v_bs,
                          int64_t o_bs, float scale, cudaStream_t stream);

}  // namespace rdeic

using namespace rdeic;

extern "C" int rdeic_attention(const void* q, const void* k, const void* v, void* out, int B,
                               int heads, int Nq, int Nk, int d, int64_t ldq, int64_t ldk,
                               int64_t ldv, int64_t ldo, int64_t q_bs, int64_t k_bs,
                               int64_t v_bs, int64_t o_bs, float scale, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(q && k && v && out, "rdeic_attention: null pointer");
    RDEIC_CHECK_ARG(B > 0 && heads > 0 && Nq > 0 && Nk > 0, "rdeic_attention: empty problem");
    RDEIC_CHECK_ARG(d == 16 || d == 64 || d == 256 || d == 512, "rdeic_attention: head dim %d not instantiated (16, 64, 256, 512)", d);
    RDEIC_CHECK_ARG(scale > 0.f, "rdeic_attention: scale must be positive (got %g)", (double)scale);
    RDEIC_CHECK_ARG(B <= 65535 && heads <= 65535, "rdeic_attention: grid too large");
    RDEIC_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0 &&
                        q_bs % 8 == 0 && k_bs % 8 == 0 && v_bs % 8 == 0 && o_bs % 2 == 0,
                    "rdeic_attention: strides must keep 16-byte row alignment");
    RDEIC_CHECK_ARG(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v) % 16 == 0 && (uintptr_t)out % 4 == 0,
                    "rdeic_attention: pointers must be 16-byte aligned");
    // wide heads (the VAE mid-block attention, d = C = 512): flash kernel with the value dimension split over two CTAs
    if (d >= 256) {
        RDEIC_CHECK_ARG((int64_t)B * heads <= 65535, "rdeic_attention: grid too large");
        RDEIC_CHECK_ARG(attention_wide_supported(d, Nq, Nk, ldq, ldk, ldv, ldo, q_bs, k_bs, v_bs, o_bs, q, k, v, out),
                        "rdeic_attention: d = %d needs Nq and Nk to be multiples of 128 (got %d, %d) and 16-byte aligned rows", d, Nq, Nk);
        return launch_attention_wide(q, k, v, out, B, heads, Nq, Nk, d, ldq, ldk, ldv, ldo, q_bs, k_bs, v_bs, o_bs, scale,
                                     as_stream(stream));
    }
    if (attention_tc_supported(d, Nq, Nk, ldq, ldk, ldv, ldo, q_bs, k_bs, v_bs, q, k, v, out) &&
        !getenv("RDEIC_ATTN_MMA_SYNC"))
        return launch_attention_tc(q, k, v, out, B, heads, Nq, Nk, ldq, ldk, ldv, ldo, q_bs, k_bs, v_bs, o_bs,
                                   scale, as_stream(stream));
    // one KV tile (the 77 text keys of cross-attention): tcgen05 kernel without a key loop
    if (attention_x_supported(d, Nq, Nk, ldq, ldk, ldv, ldo, q_bs, k_bs, v_bs, o_bs, q, k, v, out) &&
        !getenv("RDEIC_ATTN_MMA_SYNC"))
        return launch_attention_x(q, k, v, out, B, heads, Nq, Nk, ldq, ldk, ldv, ldo, q_bs, k_bs, v_bs, o_bs, scale,
                                  as_stream(stream));
    const float scale_log2 = scale * 1.4426950408889634f;
    dim3 grid((Nq + kQTile - 1) / kQTile, heads, B);
    if (d == 64)
        launch_k(attention_kernel<64>, grid, kAttnWarps * 32, 0, as_stream(stream), 
            (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
            (__nv_bfloat16*)out, Nq, Nk, ldq, ldk, ldv, ldo, q_bs, k_bs, v_bs, o_bs, scale_log2);
    else
        launch_k(attention_kernel<16>, grid, kAttnWarps * 32, 0, as_stream(stream), 
            (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
            (__nv_bfloat16*)out, Nq, Nk, ldq, ldk, ldv, ldo, q_bs, k_bs, v_bs, o_bs, scale_log2);
    RDEIC_LAUNCH_CHECK();
    return 0;
}
