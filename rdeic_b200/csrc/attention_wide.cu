// Flash attention for WIDE heads (d = 256 .. 512, a multiple of 128) on tcgen05: the VAE mid-block attention
// (ldm/modules/diffusionmodules/model.py:181-205: one head, d = C = 512, N = H*W tokens).  The reference materialises the
// [B, N, N] score tensor (torch.bmm -> softmax -> bmm); at BASELINE config 3 (768x512, batch 64, N = 6144) that is 9.7 GB of
// fp32.  Here it never exists: S tiles live in tensor memory.
//
// An O row of d fp32 values is d TMEM columns, and 512 columns is all the tensor memory an SM has, so the value
// dimension is split in two: CTA (qt, half, b) owns 128 query rows and d/2 output columns, and both halves compute the
// same S = Q K^T tiles (the price of the split: QK^T runs twice, 1.5x the FLOPs of the unsplit product, in exchange for
// never writing N x N anything).
//   TMEM   : S0 | S1 (2 x 128 columns, double-buffered so S(j+1) is computed while the softmax warps work on S(j)),
//            O (d/2 <= 256 columns).  The probabilities overwrite the first 64 columns of the S buffer they came from
//            (bf16 pairs; each thread has read its own row before it writes) and feed the PV MMAs as a TMEM A operand.
//   smem   : Q tile resident (128 rows x d, d/64 panels of 16 KB), a ring of 16 KB stages carrying K panels
//            (128 keys x 64 d, K-major) and V sub-tiles (128 keys x 64 d_v, consumed MN-major: no transpose).
//   warps  : 0 = TMA producer, 1 = issuer of the S = Q K^T MMAs (+ TMEM allocation), 2 = issuer of the PV MMAs,
//            4..7 = softmax, one query row per thread.  Two issuing warps, one per instruction stream (S and PV have
//            their own operand rings and hand-offs); a step is 32 + 32 instructions, ~3100 tensor clocks.
//   O stays in tensor memory for the whole key loop and is rescaled only when a row's maximum has grown by more than
//   2^8 over the reference maximum its exponents use (exact: O and l carry the same factor) -- as attention_tc2_kernel.
#include "common.cuh"
#include "sm100.cuh"
#include "../../include/rdeic_b200.h"

namespace rdeic {

constexpr int kWQ = 128;                       // query rows per CTA
constexpr int kWK = 128;                       // keys per tile
constexpr int kWPanel = kWK * 64 * 2;          // 16 KB: 128 rows x 64 bf16
constexpr int kWThreads = 256;
constexpr uint32_t kWColS = 0, kWColO = 256;

struct AttWDev {
    int tiles_k;
    int d;                                     // head dim (256, 384, 512)
    float scale_log2;
    __nv_bfloat16* out;
    int64_t ldo, o_bs;
    int heads;
};

__host__ __device__ constexpr int wide_stages(int d) { return (227 * 1024 - 1024 - 256 - (d / 64) * kWPanel) / kWPanel > 8 ? 8 : (227 * 1024 - 1024 - 256 - (d / 64) * kWPanel) / kWPanel; }
__host__ __device__ constexpr int wide_smem(int d) { return (d / 64) * kWPanel + wide_stages(d) * kWPanel + 1024 + 256; }

__device__ __forceinline__ uint64_t make_smem_desc_mn_w(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)(kWPanel >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ float ex2_approx_w(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int kD>
__global__ void __launch_bounds__(kWThreads, 1)
attention_wide_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                      const __grid_constant__ CUtensorMap tm_v, const AttWDev p) {
    pdl_trigger();
    constexpr int kPanels = kD / 64;           // K panels per S tile (and Q panels resident)
    constexpr int kHalf = kD / 2;              // output columns of this CTA
    constexpr int kSubs = kHalf / 64;          // V sub-tiles per key tile
    constexpr int kStagesAll = wide_stages(kD);
    constexpr int kVStages = kStagesAll / 3;   // V ring (PV issuer); the K ring (S issuer) takes the rest: a step consumes
    constexpr int kKStages = kStagesAll - kVStages;   // twice as many K panels as V sub-tiles
    static_assert(kHalf <= 256 && kVStages >= 2, "O must fit 256 TMEM columns and each ring needs two stages");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* s_q = smem;
    uint8_t* s_kring = s_q + kPanels * kWPanel;
    uint8_t* s_vring = s_kring + kKStages * kWPanel;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_vring + kVStages * kWPanel);
    uint64_t* q_full = bars;                   // [1]
    uint64_t* k_full = bars + 1;               // [kKStages]
    uint64_t* k_empty = k_full + kKStages;     // [kKStages]
    uint64_t* v_full = k_empty + kKStages;     // [kVStages]
    uint64_t* v_empty = v_full + kVStages;     // [kVStages]
    uint64_t* s_full = v_empty + kVStages;     // [2] S(j) complete
    uint64_t* p_full = s_full + 2;             // [2] P(j) written (4 softmax warps)
    uint64_t* o_full = p_full + 2;             // [2] PV(j) retired: O may be rescaled, S buffer j & 1 may be overwritten
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, half = blockIdx.y;
    const int b = blockIdx.z / p.heads, head = blockIdx.z - b * p.heads;
    const int T = p.tiles_k;
    const int c_head = head * kD;              // first column of this head in Q / K / V / out

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < kKStages; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
        for (int s = 0; s < kVStages; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&o_full[i], 1); }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        if (elect_one_sync()) {
            // ===== TMA producer: Q once, then stages in exactly the order the MMA thread consumes them =====
            mbar_expect_tx(q_full, kPanels * kWPanel);
            for (int pn = 0; pn < kPanels; ++pn) tma_load_3d(&tm_q, s_q + pn * kWPanel, q_full, c_head + pn * 64, qt * kWQ, b);
            uint32_t ks = 0, kph = 0, vs = 0, vph = 0;
            auto push_k = [&](int col, int row) {
                mbar_wait(&k_empty[ks], kph ^ 1);
                mbar_expect_tx(&k_full[ks], kWPanel);
                tma_load_3d(&tm_k, s_kring + ks * kWPanel, &k_full[ks], col, row, b);
                if (++ks == kKStages) { ks = 0; kph ^= 1; }
            };
            auto push_v = [&](int col, int row) {
                mbar_wait(&v_empty[vs], vph ^ 1);
                mbar_expect_tx(&v_full[vs], kWPanel);
                tma_load_3d(&tm_v, s_vring + vs * kWPanel, &v_full[vs], col, row, b);
                if (++vs == kVStages) { vs = 0; vph ^= 1; }
            };
            // V(j-1) is always requested before K(j+1): S(j+1) waits for PV(j-1), so the K ring can never block the V ring
            for (int pn = 0; pn < kPanels; ++pn) push_k(c_head + pn * 64, 0);
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T)
                    for (int pn = 0; pn < kPanels; ++pn) push_k(c_head + pn * 64, (j + 1) * kWK);
                for (int sb = 0; sb < kSubs; ++sb) push_v(c_head + half * kHalf + sb * 64, j * kWK);
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            // ===== issuer of S(j) = Q K(j)^T into S buffer j & 1 =====
            constexpr uint32_t idesc_s = make_idesc_mn(kWQ, kWK, false);
            uint32_t s = 0, ph = 0;
            mbar_wait(q_full, 0);
            for (int j = 0; j < T; ++j) {
                // buffer j & 1 held S(j-2), then P(j-2): PV(j-2) must have retired
                if (j >= 2) {
                    mbar_wait(&o_full[j & 1], ((j - 2) >> 1) & 1);
                    tc_fence_after();
                }
                const uint32_t ts = tmem_base + kWColS + (j & 1) * kWK;
                for (int pn = 0; pn < kPanels; ++pn) {
                    mbar_wait(&k_full[s], ph);
                    tc_fence_after();
                    const uint64_t dq = make_smem_desc(smem_u32(s_q + pn * kWPanel));
                    const uint64_t dk = make_smem_desc(smem_u32(s_kring + s * kWPanel));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(ts, dq + 2 * k, dk + 2 * k, idesc_s, (pn | k) != 0);
                    umma_commit(&k_empty[s]);
                    if (++s == kKStages) { s = 0; ph ^= 1; }
                }
                umma_commit(&s_full[j & 1]);
            }
        }
    } else if (warp == 2) {
        if (elect_one_sync()) {
            // ===== issuer of O += P(j) V(j) =====
            constexpr uint32_t idesc_o = make_idesc_mn(kWQ, 64, true);
            uint32_t s = 0, ph = 0;
            for (int j = 0; j < T; ++j) {
                mbar_wait(&p_full[j & 1], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t tp = tmem_base + kWColS + (j & 1) * kWK;      // P sits on the first 64 columns of S(j)
                for (int sb = 0; sb < kSubs; ++sb) {
                    mbar_wait(&v_full[s], ph);
                    tc_fence_after();
                    const uint64_t dv = make_smem_desc_mn_w(smem_u32(s_vring + s * kWPanel));
#pragma unroll
                    for (int ks = 0; ks < kWK / 16; ++ks)
                        umma_bf16_ts(tmem_base + kWColO + sb * 64, tp + ks * 8, dv + (uint64_t)ks * (2048 >> 4), idesc_o, (j | ks) != 0);
                    umma_commit(&v_empty[s]);
                    if (++s == kVStages) { s = 0; ph ^= 1; }
                }
                umma_commit(&o_full[j & 1]);
            }
        }
    } else if (warp >= 4) {
        // ===== softmax: one query row per thread =====
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const uint32_t to = tmem_base + lane_addr + kWColO;
        const float sc = p.scale_log2;
        float m_ref = -INFINITY, l_run = 0.f;
        for (int j = 0; j < T; ++j) {
            const uint32_t ts = tmem_base + lane_addr + kWColS + (j & 1) * kWK;
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            uint32_t sv[kWK];
#pragma unroll
            for (int c = 0; c < kWK; c += 16) tmem_ld16(ts + c, sv + c);
            tmem_ld_wait();
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int c = 0; c < kWK; c += 4) {
                mx0 = fmaxf(mx0, __uint_as_float(sv[c]));     mx1 = fmaxf(mx1, __uint_as_float(sv[c + 1]));
                mx2 = fmaxf(mx2, __uint_as_float(sv[c + 2])); mx3 = fmaxf(mx3, __uint_as_float(sv[c + 3]));
            }
            const float m_new = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
            const bool grow = m_new > m_ref + 8.0f;                  // first tile: m_ref = -inf -> true
            if (__any_sync(0xffffffffu, grow)) {
                const float m_to = grow ? m_new : m_ref;
                const float corr = ex2_approx_w(m_ref - m_to);       // 1 for the rows that keep their reference
                if (j > 0) {
                    mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);   // PV(j-1) has retired: O may be rewritten
                    tc_fence_after();
#pragma unroll 1
                    for (int c2 = 0; c2 < kHalf; c2 += 16) {
                        uint32_t r[16];
                        tmem_ld16(to + c2, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int c = 0; c < 16; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) * corr);
                        tmem_st16(to + c2, r);
                    }
                    l_run *= corr;
                }
                m_ref = m_to;
            }
            float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
            for (int c32 = 0; c32 < 4; ++c32) {
                uint32_t pk[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float p0 = ex2_approx_w(fmaf(__uint_as_float(sv[c32 * 32 + 2 * c]), sc, -m_ref));
                    const float p1 = ex2_approx_w(fmaf(__uint_as_float(sv[c32 * 32 + 2 * c + 1]), sc, -m_ref));
                    rs0 += p0; rs1 += p1;
                    pk[c] = pack_bf16x2(p0, p1);
                }
                tmem_st16(ts + c32 * 16, pk);
            }
            l_run += rs0 + rs1;
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[j & 1]);
        }
        mbar_wait(&o_full[(T - 1) & 1], ((T - 1) >> 1) & 1);
        tc_fence_after();
        const float inv = 1.0f / l_run;
        __nv_bfloat16* dst = p.out + (int64_t)b * p.o_bs + (int64_t)(qt * kWQ + row) * p.ldo + c_head + half * kHalf;
#pragma unroll 1
        for (int c2 = 0; c2 < kHalf; c2 += 32) {
            uint32_t r[32];
            tmem_ld16(to + c2, r);
            tmem_ld16(to + c2 + 16, r + 16);
            tmem_ld_wait();
#pragma unroll
            for (int v8 = 0; v8 < 4; ++v8) {
                uint4 w;
                w.x = pack_bf16x2(__uint_as_float(r[8 * v8]) * inv, __uint_as_float(r[8 * v8 + 1]) * inv);
                w.y = pack_bf16x2(__uint_as_float(r[8 * v8 + 2]) * inv, __uint_as_float(r[8 * v8 + 3]) * inv);
                w.z = pack_bf16x2(__uint_as_float(r[8 * v8 + 4]) * inv, __uint_as_float(r[8 * v8 + 5]) * inv);
                w.w = pack_bf16x2(__uint_as_float(r[8 * v8 + 6]) * inv, __uint_as_float(r[8 * v8 + 7]) * inv);
                *reinterpret_cast<uint4*>(dst + c2 + 8 * v8) = w;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

bool attention_wide_supported(int d, int Nq, int Nk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t q_bs,
                              int64_t k_bs, int64_t v_bs, int64_t o_bs, const void* q, const void* k, const void* v,
                              const void* out) {
    return (d == 256 || d == 512) && Nq % kWQ == 0 && Nk % kWK == 0 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 &&
           ldo % 8 == 0 && q_bs % 8 == 0 && k_bs % 8 == 0 && v_bs % 8 == 0 && o_bs % 8 == 0 &&
           (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) == 0;
}

template <int kD>
static int launch_wide(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttWDev& d, dim3 grid,
                       cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        RDEIC_CUDA(cudaFuncSetAttribute(attention_wide_kernel<kD>, cudaFuncAttributeMaxDynamicSharedMemorySize, wide_smem(kD)));
        attr_set = true;
    }
    RDEIC_CUDA(launch_k(attention_wide_kernel<kD>, grid, kWThreads, wide_smem(kD), stream, tq, tk, tv, d));
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int launch_attention_wide(const void* q, const void* k, const void* v, void* out, int B, int heads, int Nq, int Nk, int d,
                          int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t q_bs, int64_t k_bs, int64_t v_bs,
                          int64_t o_bs, float scale, cudaStream_t stream) {
    CUtensorMap tq, tk, tv;
    const uint32_t box[3] = {64, (uint32_t)kWQ, 1};
    auto mk = [&](CUtensorMap* m, const void* ptr, int N, int64_t ld, int64_t bs, const char* what) -> int {
        const uint64_t bstride = (B > 1 ? (uint64_t)bs : (uint64_t)ld * N) * 2;
        uint64_t dims[3] = {(uint64_t)heads * d, (uint64_t)N, (uint64_t)B};
        uint64_t str[2] = {(uint64_t)ld * 2, bstride};
        return encode_map(m, ptr, 3, dims, str, box, what);
    };
    if (int e = mk(&tq, q, Nq, ldq, q_bs, "wide attn Q")) return e;
    if (int e = mk(&tk, k, Nk, ldk, k_bs, "wide attn K")) return e;
    if (int e = mk(&tv, v, Nk, ldv, v_bs, "wide attn V")) return e;
    AttWDev dv;
    dv.tiles_k = Nk / kWK;
    dv.d = d;
    dv.scale_log2 = scale * 1.4426950408889634f;
    dv.out = (__nv_bfloat16*)out;
    dv.ldo = ldo; dv.o_bs = o_bs;
    dv.heads = heads;
    dim3 grid(Nq / kWQ, 2, B * heads);
    if (d == 512) return launch_wide<512>(tq, tk, tv, dv, grid, stream);
    return launch_wide<256>(tq, tk, tv, dv, grid, stream);
}

}  // namespace rdeic
