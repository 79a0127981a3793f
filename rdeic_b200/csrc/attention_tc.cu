// Flash-style attention forward on 5th-gen tensor cores (sm_100a) for head dim 64:
//   S = Q K^T and O_tile = P V are tcgen05.mma instructions issued by one thread, S and O tiles live
//   in TMEM (two buffers each), Q/K/V tiles arrive by TMA into 128B-swizzled shared memory, and
//   four softmax warps (one query row per thread) turn S into P with an online softmax in fp32
//   (exp2 domain).  P goes back to shared memory in the K-major swizzled layout the MMA reads;
//   V is consumed as an MN-major B operand straight from its row-major TMA tile (no transpose).
//   O_tile is never accumulated in TMEM: each thread adds it to its 64 fp32 registers and applies
//   the running-max correction there, so there is no TMEM read-modify-write on the critical path.
//
// Replaces `CrossAttention.forward` (ldm/modules/attention.py:171-203) for the SD-2.1 self-attention
// shapes (d = 64, token counts that are multiples of 128).  Ragged shapes (text keys, Nk = 77) and
// the control adapter's d = 16 heads stay on the mma.sync kernel in attention.cu.
#include "common.cuh"
#include "sm100.cuh"
#include <stdlib.h>
#include "../../include/rdeic_b200.h"

namespace rdeic {

constexpr int kAD = 64;                       // head dim
constexpr int kAQ = 128;                      // query rows per CTA (= TMEM lanes)
constexpr int kAK = 128;                      // keys per tile
constexpr int kAStages = 3;                   // K/V ring
constexpr int kATile = kAK * kAD * 2;         // 16 KB: one Q, K or V tile
constexpr int kAPBytes = kAQ * kAK * 2;       // 32 KB: one P buffer (two 64-key panels)
constexpr int kASoftWarps = 8;                 // two threads per query row (64 keys / 32 O columns each)
constexpr int kAThreads = 64 + 32 * kASoftWarps;   // warp0 TMA, warp1 MMA, warps 2-9 softmax
constexpr int kAXBytes = 2 * 2 * kAQ * 4 + 2 * kAQ * 4;   // row-max exchange [parity][half][row] + row-sum exchange
constexpr int kASmem = kATile * (1 + 2 * kAStages) + 2 * kAPBytes + kAXBytes + 1024 + 256;
constexpr uint32_t kColS = 0, kColO = 256;    // TMEM columns: S0,S1 at 0/128 ; O0,O1 at 256/320

struct AttDev {
    int tiles_k;
    float scale_log2;
    __nv_bfloat16* out;
    int64_t ldo, o_bs;
};

// MN-major, 128-byte swizzled B operand (V tile: rows = keys (K), 64 contiguous d (N) per row):
// SBO = 1024 B between groups of 8 K-rows, LBO unused for a single 64-wide MN atom.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffffu) >> 4);
    d |= (uint64_t)(kATile >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(kAThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                    const __grid_constant__ CUtensorMap tm_v, const AttDev p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint8_t* s_q = smem;
    uint8_t* s_k = s_q + kATile;
    uint8_t* s_v = s_k + kAStages * kATile;
    uint8_t* s_p = s_v + kAStages * kATile;
    float* s_xmax = reinterpret_cast<float*>(s_p + 2 * kAPBytes);          // [2][2][128]
    float* s_xsum = s_xmax + 2 * 2 * kAQ;                                  // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + 2 * kAPBytes + kAXBytes);
    uint64_t* q_full = bars;                  // [1]
    uint64_t* kv_full = bars + 1;             // [kAStages]
    uint64_t* kv_empty = kv_full + kAStages;  // [kAStages]
    uint64_t* s_full = kv_empty + kAStages;   // [2]
    uint64_t* p_full = s_full + 2;            // [2]
    uint64_t* o_full = p_full + 2;            // [2]
    uint64_t* o_empty = o_full + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int T = p.tiles_k;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < kAStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], kASoftWarps);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], kASoftWarps);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // prologue above overlapped the previous kernel's tail

    if (warp == 0) {
        if (elect_one_sync()) {
            // ===== TMA producer =====
            mbar_expect_tx(q_full, kATile);
            tma_load_3d(&tm_q, s_q, q_full, head * kAD, qt * kAQ, b);
            for (int j = 0; j < T; ++j) {
                const int s = j % kAStages;
                mbar_wait(&kv_empty[s], ((j / kAStages) & 1) ^ 1);
                mbar_expect_tx(&kv_full[s], 2 * kATile);
                tma_load_3d(&tm_k, s_k + s * kATile, &kv_full[s], head * kAD, j * kAK, b);
                tma_load_3d(&tm_v, s_v + s * kATile, &kv_full[s], head * kAD, j * kAK, b);
            }
        }
    } else if (warp == 1) {
        if (elect_one_sync()) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc_s = make_idesc_mn(kAQ, kAK, false);   // S  = Q K^T : N = 128 keys
            constexpr uint32_t idesc_o = make_idesc_mn(kAQ, kAD, true);    // O  = P V   : N = 64, V MN-major
            const uint64_t dq = make_smem_desc(smem_u32(s_q));
            auto issue_pv = [&](int i) {
                const uint32_t bf = i & 1, ph = (i >> 1) & 1;
                mbar_wait(&p_full[bf], ph);
                mbar_wait(&o_empty[bf], ph ^ 1);
                tc_fence_after();
                const uint32_t pb = smem_u32(s_p + bf * kAPBytes);
                const uint64_t dv = make_smem_desc_mn(smem_u32(s_v + (i % kAStages) * kATile));
#pragma unroll
                for (int ks = 0; ks < kAK / 16; ++ks) {
                    const uint64_t dp = make_smem_desc(pb + (ks >> 2) * (kAQ * 128) + (ks & 3) * 32);
                    umma_bf16(tmem_base + kColO + bf * kAD, dp, dv + (uint64_t)ks * (2048 >> 4), idesc_o, ks != 0);
                }
                umma_commit(&o_full[bf]);
                umma_commit(&kv_empty[i % kAStages]);
            };
            mbar_wait(q_full, 0);
            for (int j = 0; j < T; ++j) {
                const int s = j % kAStages;
                mbar_wait(&kv_full[s], (j / kAStages) & 1);
                tc_fence_after();
                const uint64_t dk = make_smem_desc(smem_u32(s_k + s * kATile));
#pragma unroll
                for (int k = 0; k < kAD / 16; ++k)
                    umma_bf16(tmem_base + kColS + (j & 1) * kAK, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[j & 1]);
                if (j > 0) issue_pv(j - 1);
            }
            issue_pv(T - 1);
        }
    } else {
        // ===== softmax / correction / epilogue: two threads per query row =====
        // Warps w and w+4 (same TMEM lane quadrant) share 32 rows: the first takes keys 0-63 of each
        // tile and O columns 0-31, the second keys 64-127 and O columns 32-63.  They exchange the row
        // maximum through shared memory once per tile (64-thread named barrier) and the row sum once
        // at the end; with 2 softmax warps per scheduler the exp / convert chains overlap.
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const float sc = p.scale_log2;
        constexpr int kOC = kAD / 2;             // O columns per thread
        float o[kOC];
#pragma unroll
        for (int i = 0; i < kOC; ++i) o[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;
        auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory"); };

        auto add_o = [&](int i) {
            const uint32_t bf = i & 1;
            mbar_wait(&o_full[bf], (i >> 1) & 1);
            tc_fence_after();
            uint32_t r[32];
            const uint32_t ta = tmem_base + lane_addr + kColO + bf * kAD + half * kOC;
            tmem_ld16(ta, r);
            tmem_ld16(ta + 16, r + 16);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] += __uint_as_float(r[c]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_empty[bf]);
        };

        for (int j = 0; j < T; ++j) {
            const uint32_t bf = j & 1;
            mbar_wait(&s_full[bf], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t ts = tmem_base + lane_addr + kColS + bf * kAK + half * 64;
            // pass 1: maximum over this thread's 64 keys, then across the pair
            float mx = -INFINITY;
#pragma unroll
            for (int c4 = 0; c4 < 2; ++c4) {
                uint32_t r[32];
                tmem_ld16(ts + c4 * 32, r);
                tmem_ld16(ts + c4 * 32 + 16, r + 16);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(r[c]));
            }
            s_xmax[(bf * 2 + half) * kAQ + row] = mx;
            pair_sync();
            mx = fmaxf(mx, s_xmax[(bf * 2 + (half ^ 1)) * kAQ + row]);
            const float m_new = fmaxf(m_run, mx * sc);
            const float corr = ex2_approx(m_run - m_new);
            // pass 2: P = exp2(S*scale - m) -> bf16 -> swizzled K-major smem (one 64-key panel row)
            float rowsum = 0.f;
            uint8_t* prow = s_p + bf * kAPBytes + half * (kAQ * 128) + row * 128;
#pragma unroll
            for (int c4 = 0; c4 < 2; ++c4) {
                uint32_t r[32];
                tmem_ld16(ts + c4 * 32, r);
                tmem_ld16(ts + c4 * 32 + 16, r + 16);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * c]), sc, -m_new));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * c + 1]), sc, -m_new));
                    rowsum += p0 + p1;
                    pk[c] = pack_bf16x2(p0, p1);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = c4 * 4 + q;                 // 16-byte chunk = 8 keys, within the panel
                    *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
                        make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
            }
            l_run = fmaf(l_run, corr, rowsum);
            m_run = m_new;
            tc_fence_before();        // S reads done before the MMA warp may overwrite this S buffer
            fence_proxy_async();      // P stores visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[bf]);
            // fold in the previous tile's O while the tensor core works on this one
            if (j > 0) add_o(j - 1);
#pragma unroll
            for (int i = 0; i < kOC; ++i) o[i] *= corr;
        }
        add_o(T - 1);
        s_xsum[half * kAQ + row] = l_run;
        pair_sync();
        const float inv = 1.0f / (l_run + s_xsum[(half ^ 1) * kAQ + row]);
        __nv_bfloat16* dst = p.out + (int64_t)b * p.o_bs + (int64_t)(qt * kAQ + row) * p.ldo + head * kAD + half * kOC;
#pragma unroll
        for (int v8 = 0; v8 < kOC / 8; ++v8) {
            uint4 w;
            w.x = pack_bf16x2(o[8 * v8] * inv, o[8 * v8 + 1] * inv);
            w.y = pack_bf16x2(o[8 * v8 + 2] * inv, o[8 * v8 + 3] * inv);
            w.z = pack_bf16x2(o[8 * v8 + 4] * inv, o[8 * v8 + 5] * inv);
            w.w = pack_bf16x2(o[8 * v8 + 6] * inv, o[8 * v8 + 7] * inv);
            *reinterpret_cast<uint4*>(dst + 8 * v8) = w;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
// Two query tiles per CTA ("ping-pong"): 256 query rows, two softmax warp groups (A: rows 0-127,
// B: rows 128-255) of four warps, ONE thread per query row — no pair exchange, no named barrier.
// The exp pass is MUFU bound; everything around it (max pass, O fold-in, rescale, barrier waits) is
// not.  With one group per CTA those phases leave the MUFU pipe idle; with two groups that drift half
// a period apart one group's MUFU phase covers the other's bookkeeping, and the single MMA thread
// always has the other group's S / PV to issue.  S, P and O are single-buffered per group: program
// order inside a group already guarantees the hand-offs (P(j+1) and PV(j+1) come after add_o(j)).
// ---------------------------------------------------------------------------------------------
constexpr int kA2Q = 256;
// warp 0 = TMA producer, warps 1 / 2 = MMA issuers of group A / B, warp 3 idle, warps 4-7 / 8-11 = softmax group A / B.
// One issuing thread sustains one tcgen05.mma per ~100 clocks whatever the shape (scripts/micro/ubench.cu: N = 16 .. 160, A in
// shared or tensor memory, all 100 clocks; two issuing warps: 52 per SM, four: 26), and a 128-key step of the two groups is
// 8 S + 16 PV instructions: 2400 clocks from one thread against 2048 clocks of MUFU work.  With one issuer per group the
// instruction stream is 1200 clocks per step and the exponentials are the only bound left.
// Warps 0-3 form one warp group so that setmaxnreg can hand their registers to the softmax groups.
constexpr int kA2Threads = 128 + 32 * 8;
constexpr int kA2RegsCtl = 56, kA2RegsSoft = 216;         // 128 * 56 + 256 * 216 = 62464 <= 384 * 168
constexpr int kA2Smem = 2 * kATile /*Q*/ + 2 * kAStages * kATile /*K,V*/ + 2 * kAPBytes /*P_A,P_B*/ + 1024 + 256;
constexpr uint32_t kA2ColS = 0, kA2ColO = 256;    // S_A 0, S_B 128 ; O_A 256, O_B 320
// kLazy keeps the probabilities in tensor memory (P_A 384, P_B 448: 128 keys = 64 packed columns each) and feeds
// them to the PV MMAs as a TMEM A operand: no P round trip through shared memory, and the shared memory that
// held P buys two more K/V stages
constexpr uint32_t kA2ColP = 384;
constexpr int kA2LazyStages = 5;
constexpr int kA2LazySmem = 2 * kATile /*Q*/ + 2 * kA2LazyStages * kATile /*K,V*/ + 1024 + 256;

// kLazy: O stays in TMEM for the whole key loop (the PV MMAs accumulate in place) and is rescaled only
// when a row's running maximum has grown by more than 2^8 since the reference maximum its exponents use
// (exact: l and O carry the same factor, which cancels in O / l; P <= 256 is harmless in bf16).  The
// per-step fold-in of PV (TMEM load + 64 adds + 64 multiplies per thread) disappears, and with the 64
// accumulator registers gone the S tile is read from TMEM once (128 registers) instead of twice.
template <bool kLazy>
__global__ void __launch_bounds__(kA2Threads, 1)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const AttDev p) {
    pdl_trigger();
    constexpr int kAStages = kLazy ? kA2LazyStages : rdeic::kAStages;      // shadows the 3-stage ring of the other kernels
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint8_t* s_q = smem;                                   // 2 tiles: rows 0-127 | 128-255
    uint8_t* s_k = s_q + 2 * kATile;
    uint8_t* s_v = s_k + kAStages * kATile;
    uint8_t* s_p = s_v + kAStages * kATile;                // P_A | P_B (not kLazy: there P lives in tensor memory)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + (kLazy ? 0 : 2 * kAPBytes));
    uint64_t* q_full = bars;                  // [1]
    uint64_t* kv_full = bars + 1;             // [kAStages]
    uint64_t* kv_empty = kv_full + kAStages;  // [kAStages]
    uint64_t* s_full = kv_empty + kAStages;   // [2] per group: S tile ready
    uint64_t* s_free = s_full + 2;            // [2] per group: S tile consumed
    uint64_t* p_full = s_free + 2;            // [2] per group: P tile written
    uint64_t* o_full = p_full + 2;            // [2] per group: PV tile ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int T = p.tiles_k;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < kAStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 2); }   // both issuers release a stage
        for (int g = 0; g < 2; ++g) {
            mbar_init(&s_full[g], 1);
            mbar_init(&s_free[g], 4);
            mbar_init(&p_full[g], 4);
            mbar_init(&o_full[g], 1);
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    // (setmaxnreg sits inside each role's branch: ptxas bounds the registers of the code that FOLLOWS one, path-insensitively)
    if (warp == 0) {
        setmaxnreg_dec<kA2RegsCtl>();
        if (elect_one_sync()) {
            mbar_expect_tx(q_full, 2 * kATile);
            tma_load_3d(&tm_q, s_q, q_full, head * kAD, qt * kA2Q, b);
            tma_load_3d(&tm_q, s_q + kATile, q_full, head * kAD, qt * kA2Q + kAQ, b);
            for (int j = 0; j < T; ++j) {
                const int s = j % kAStages;
                mbar_wait(&kv_empty[s], ((j / kAStages) & 1) ^ 1);
                mbar_expect_tx(&kv_full[s], 2 * kATile);
                tma_load_3d(&tm_k, s_k + s * kATile, &kv_full[s], head * kAD, j * kAK, b);
                tma_load_3d(&tm_v, s_v + s * kATile, &kv_full[s], head * kAD, j * kAK, b);
            }
        }
    } else if (warp == 1 || warp == 2) {
        setmaxnreg_dec<kA2RegsCtl>();
        if (elect_one_sync()) {
            const int grp = warp - 1;                                    // this thread issues group grp's S and PV MMAs
            constexpr uint32_t idesc_s = make_idesc_mn(kAQ, kAK, false);
            constexpr uint32_t idesc_o = make_idesc_mn(kAQ, kAD, true);
            auto issue_s = [&](int g, int j) {
                if (j > 0) mbar_wait(&s_free[g], (j - 1) & 1);      // softmax_g(j-1) has read S_g
                tc_fence_after();
                const uint64_t dq = make_smem_desc(smem_u32(s_q + g * kATile));
                const uint64_t dk = make_smem_desc(smem_u32(s_k + (j % kAStages) * kATile));
#pragma unroll
                for (int k = 0; k < kAD / 16; ++k)
                    umma_bf16(tmem_base + kA2ColS + g * kAK, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[g]);
            };
            auto issue_pv = [&](int g, int i) {
                mbar_wait(&p_full[g], i & 1);
                tc_fence_after();
                const uint32_t pb = smem_u32(s_p + g * kAPBytes);
                const uint64_t dv = make_smem_desc_mn(smem_u32(s_v + (i % kAStages) * kATile));
#pragma unroll
                for (int ks = 0; ks < kAK / 16; ++ks) {
                    if constexpr (kLazy) {
                        umma_bf16_ts(tmem_base + kA2ColO + g * kAD, tmem_base + kA2ColP + g * (kAK / 2) + ks * 8,
                                     dv + (uint64_t)ks * (2048 >> 4), idesc_o, (i | ks) != 0);
                    } else {
                        const uint64_t dp = make_smem_desc(pb + (ks >> 2) * (kAQ * 128) + (ks & 3) * 32);
                        umma_bf16(tmem_base + kA2ColO + g * kAD, dp, dv + (uint64_t)ks * (2048 >> 4), idesc_o, ks != 0);
                    }
                }
                umma_commit(&o_full[g]);
            };
            mbar_wait(q_full, 0);
            for (int j = 0; j < T; ++j) {
                mbar_wait(&kv_full[j % kAStages], (j / kAStages) & 1);
                issue_s(grp, j);
                if (j > 0) { issue_pv(grp, j - 1); umma_commit(&kv_empty[(j - 1) % kAStages]); }
            }
            issue_pv(grp, T - 1);
            umma_commit(&kv_empty[(T - 1) % kAStages]);
        }
    } else if (warp == 3) {
        setmaxnreg_dec<kA2RegsCtl>();
    } else {
        setmaxnreg_inc<kA2RegsSoft>();
        const int g = (warp - 4) >> 2;              // softmax group: 0 = A, 1 = B
        const int quad = warp & 3;                  // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;           // row inside the group's 128-row tile
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        const float sc = p.scale_log2;
        if constexpr (kLazy) {
            float m_ref = -INFINITY, l_run = 0.f;
            const uint32_t ts = tmem_base + lane_addr + kA2ColS + g * kAK;
            const uint32_t to = tmem_base + lane_addr + kA2ColO + g * kAD;
            const uint32_t tp = tmem_base + lane_addr + kA2ColP + g * (kAK / 2);
            for (int j = 0; j < T; ++j) {
                mbar_wait(&s_full[g], j & 1);
                tc_fence_after();
                uint32_t sv[kAK];
#pragma unroll
                for (int c = 0; c < kAK; c += 16) tmem_ld16(ts + c, sv + c);
                tmem_ld_wait();
                tc_fence_before();                  // the only read of S_g: the MMA thread may overwrite it
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_free[g]);
                // four independent chains: one serial chain of 128 dependent FMNMX is ~500 clocks of pure latency
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
                for (int c = 0; c < kAK; c += 4) {
                    mx0 = fmaxf(mx0, __uint_as_float(sv[c]));     mx1 = fmaxf(mx1, __uint_as_float(sv[c + 1]));
                    mx2 = fmaxf(mx2, __uint_as_float(sv[c + 2])); mx3 = fmaxf(mx3, __uint_as_float(sv[c + 3]));
                }
                const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                const float m_new = mx * sc;
                // PV(j-1) must have retired before P_g is overwritten (and before O_g is rescaled)
                if (j > 0) {
                    mbar_wait(&o_full[g], (j - 1) & 1);
                    tc_fence_after();
                }
                const bool grow = m_new > m_ref + 8.0f;              // first tile: m_ref = -inf -> true
                if (__any_sync(0xffffffffu, grow)) {
                    const float m_to = grow ? m_new : m_ref;
                    const float corr = ex2_approx(m_ref - m_to);     // 1 for rows that keep their reference
                    if (j > 0) {
#pragma unroll
                        for (int c2 = 0; c2 < kAD; c2 += 16) {       // 16 columns at a time: S occupies 128 registers
                            uint32_t r[16];
                            tmem_ld16(to + c2, r);
                            tmem_ld_wait();
#pragma unroll
                            for (int c = 0; c < 16; ++c) r[c] = __float_as_uint(__uint_as_float(r[c]) * corr);
                            tmem_st16(to + c2, r);
                        }
                        tmem_st_wait();
                        l_run *= corr;
                    }
                    m_ref = m_to;
                }
                // Packed fp32x2 arithmetic (FFMA2 / FADD2): the exponent argument s * scale - m and the row sum cost one
                // instruction per PAIR of keys.  The softmax warps are bound by the MUFU pipe plus whatever of their
                // other instructions does not overlap it (measured: routing a quarter of the exponentials through an
                // FMA-pipe polynomial made the kernel 5 % SLOWER, half of them 17 %), so fewer instructions is the lever.
                const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m_ref, -m_ref);
                float2 rs_a = make_float2(0.f, 0.f), rs_b = make_float2(0.f, 0.f);
#pragma unroll
                for (int c32 = 0; c32 < 4; ++c32) {                  // 32 keys -> 16 packed columns of P (S dies as we go)
                    uint32_t pk[16];
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        const float2 a0 = __ffma2_rn(make_float2(__uint_as_float(sv[c32 * 32 + 2 * c]), __uint_as_float(sv[c32 * 32 + 2 * c + 1])), sc2, nm2);
                        const float2 a1 = __ffma2_rn(make_float2(__uint_as_float(sv[c32 * 32 + 2 * c + 2]), __uint_as_float(sv[c32 * 32 + 2 * c + 3])), sc2, nm2);
                        const float2 e0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y));
                        const float2 e1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
                        rs_a = __fadd2_rn(rs_a, e0);
                        rs_b = __fadd2_rn(rs_b, e1);
                        pk[c] = pack_bf16x2(e0.x, e0.y);
                        pk[c + 1] = pack_bf16x2(e1.x, e1.y);
                    }
                    tmem_st16(tp + c32 * 16, pk);
                }
                l_run += (rs_a.x + rs_a.y) + (rs_b.x + rs_b.y);
                tmem_st_wait();                     // P (and the rescaled O) are in tensor memory ...
                tc_fence_before();                  // ... before the PV the arrive releases
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[g]);
            }
            mbar_wait(&o_full[g], (T - 1) & 1);
            tc_fence_after();
            const float inv = 1.0f / l_run;
            __nv_bfloat16* dst = p.out + (int64_t)b * p.o_bs + (int64_t)(qt * kA2Q + g * kAQ + row) * p.ldo + head * kAD;
#pragma unroll
            for (int c2 = 0; c2 < kAD; c2 += 32) {
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
        } else {
        float o[kAD];
#pragma unroll
        for (int i = 0; i < kAD; ++i) o[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;
        const uint32_t ts = tmem_base + lane_addr + kA2ColS + g * kAK;
        const uint32_t to = tmem_base + lane_addr + kA2ColO + g * kAD;
        uint8_t* prow0 = s_p + g * kAPBytes + row * 128;

        auto add_o = [&](int i) {
            mbar_wait(&o_full[g], i & 1);
            tc_fence_after();
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
                uint32_t r[32];
                tmem_ld16(to + c2 * 32, r);
                tmem_ld16(to + c2 * 32 + 16, r + 16);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) o[c2 * 32 + c] += __uint_as_float(r[c]);
            }
            tc_fence_before();
        };

        for (int j = 0; j < T; ++j) {
            mbar_wait(&s_full[g], j & 1);
            tc_fence_after();
            float mx = -INFINITY;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t r[32];
                tmem_ld16(ts + c4 * 32, r);
                tmem_ld16(ts + c4 * 32 + 16, r + 16);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(r[c]));
            }
            const float m_new = fmaxf(m_run, mx * sc);
            const float corr = ex2_approx(m_run - m_new);
            // fold in the previous tile's PV (also proves P_g and O_g are free again), rescale
            if (j > 0) add_o(j - 1);
#pragma unroll
            for (int i = 0; i < kAD; ++i) o[i] *= corr;
            float rowsum = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t r[32];
                tmem_ld16(ts + c4 * 32, r);
                tmem_ld16(ts + c4 * 32 + 16, r + 16);
                tmem_ld_wait();
                if (c4 == 3) {                     // last read of S_g: the MMA thread may overwrite it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_free[g]);
                }
                uint32_t pk[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * c]), sc, -m_new));
                    const float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * c + 1]), sc, -m_new));
                    rowsum += p0 + p1;
                    pk[c] = pack_bf16x2(p0, p1);
                }
                uint8_t* prow = prow0 + (c4 >> 1) * (kAQ * 128);      // 64-key panel
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = (c4 & 1) * 4 + q;
                    *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
                        make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
            }
            l_run = fmaf(l_run, corr, rowsum);
            m_run = m_new;
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[g]);
        }
        add_o(T - 1);
        const float inv = 1.0f / l_run;
        __nv_bfloat16* dst = p.out + (int64_t)b * p.o_bs + (int64_t)(qt * kA2Q + g * kAQ + row) * p.ldo + head * kAD;
#pragma unroll
        for (int v8 = 0; v8 < kAD / 8; ++v8) {
            uint4 w;
            w.x = pack_bf16x2(o[8 * v8] * inv, o[8 * v8 + 1] * inv);
            w.y = pack_bf16x2(o[8 * v8 + 2] * inv, o[8 * v8 + 3] * inv);
            w.z = pack_bf16x2(o[8 * v8 + 4] * inv, o[8 * v8 + 5] * inv);
            w.w = pack_bf16x2(o[8 * v8 + 6] * inv, o[8 * v8 + 7] * inv);
            *reinterpret_cast<uint4*>(dst + 8 * v8) = w;
        }
        }   // !kLazy
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// host: true when the tcgen05 kernel can take the problem
bool attention_tc_supported(int d, int Nq, int Nk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                            int64_t q_bs, int64_t k_bs, int64_t v_bs, const void* q, const void* k, const void* v,
                            const void* out) {
    return d == kAD && Nq % kAQ == 0 && Nk % kAK == 0 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 &&
           ldo % 8 == 0 && q_bs % 8 == 0 && k_bs % 8 == 0 && v_bs % 8 == 0 &&
           (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) == 0;
}

int launch_attention_tc(const void* q, const void* k, const void* v, void* out, int B, int heads, int Nq, int Nk,
                        int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t q_bs, int64_t k_bs,
                        int64_t v_bs, int64_t o_bs, float scale, cudaStream_t stream) {
    CUtensorMap tq, tk, tv;
    const uint32_t box[3] = {(uint32_t)kAD, (uint32_t)kAQ, 1};
    auto mk = [&](CUtensorMap* m, const void* ptr, int N, int64_t ld, int64_t bs, const char* what) -> int {
        // with B == 1 the batch stride is never used; TMA still wants a 16-byte multiple >= row extent
        const uint64_t bstride = (B > 1 ? (uint64_t)bs : (uint64_t)ld * N) * 2;
        uint64_t dims[3] = {(uint64_t)heads * kAD, (uint64_t)N, (uint64_t)B};
        uint64_t str[2] = {(uint64_t)ld * 2, bstride};
        return encode_map(m, ptr, 3, dims, str, box, what);
    };
    if (int e = mk(&tq, q, Nq, ldq, q_bs, "attn Q")) return e;
    if (int e = mk(&tk, k, Nk, ldk, k_bs, "attn K")) return e;
    if (int e = mk(&tv, v, Nk, ldv, v_bs, "attn V")) return e;
    static bool attr_set = false;
    if (!attr_set) {
        RDEIC_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kASmem));
        attr_set = true;
    }
    AttDev d;
    d.tiles_k = Nk / kAK;
    d.scale_log2 = scale * 1.4426950408889634f;
    d.out = (__nv_bfloat16*)out;
    d.ldo = ldo;
    d.o_bs = o_bs;
    // two query tiles per CTA when that still leaves several waves of CTAs
    static const int pp_min = getenv("RDEIC_ATTN_PINGPONG_MIN") ? atoi(getenv("RDEIC_ATTN_PINGPONG_MIN")) : 2 * kNumSMs;
    if (Nq % kA2Q == 0 && (int64_t)(Nq / kA2Q) * heads * B >= pp_min) {
        static bool attr2_set = false;
        if (!attr2_set) {
            RDEIC_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kA2Smem));
            RDEIC_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kA2LazySmem));
            attr2_set = true;
        }
        dim3 grid2(Nq / kA2Q, heads, B);
        static const bool lazy = !(getenv("RDEIC_ATTN_LAZY") && atoi(getenv("RDEIC_ATTN_LAZY")) == 0);
        if (lazy) launch_k(attention_tc2_kernel<true>, grid2, kA2Threads, kA2LazySmem, stream, tq, tk, tv, d);
        else launch_k(attention_tc2_kernel<false>, grid2, kA2Threads, kA2Smem, stream, tq, tk, tv, d);
        RDEIC_LAUNCH_CHECK();
        return 0;
    }
    dim3 grid(Nq / kAQ, heads, B);
    launch_k(attention_tc_kernel, grid, kAThreads, kASmem, stream, tq, tk, tv, d);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// One KV tile (Nk <= 128: the 77 text keys of every cross-attention, attention.py:171-203 with context = the
// [B,77,1024] prompt embedding): no key loop, no rescaling.  Small CTAs so that two fit an SM (48 KB of tiles,
// 256 TMEM columns, 160 threads) and hide each other's TMA -> MMA -> softmax -> MMA -> store latency chain:
//   warp 4: TMA (Q, K, V tiles; rows past Nq / Nk are zero-filled by the hardware), then the S and PV MMAs
//   warps 0-3: one query row per thread; S row from TMEM, masked past Nk, softmax, P back into the SAME TMEM
//              columns as bf16 pairs (each thread only ever touches its own lane), O row out, scaled, stored.
// The PV MMAs read P from tensor memory (A operand) and V as an MN-major operand from its row-major tile.
// ---------------------------------------------------------------------------------------------
constexpr int kAXThreads = 160;
constexpr int kAXSmem = 3 * kATile + 1024 + 128;
constexpr uint32_t kAXColS = 0, kAXColO = 128;       // S / P at 0, O at 128 (256 columns allocated)

struct AttXDev {
    int Nq, Nk;
    float scale_log2;
    __nv_bfloat16* out;
    int64_t ldo, o_bs;
};

__global__ void __launch_bounds__(kAXThreads)
attention_x_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const AttXDev p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* s_q = smem;
    uint8_t* s_k = s_q + kATile;
    uint8_t* s_v = s_k + kATile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_v + kATile);
    uint64_t* ld_full = bars;          // Q, K, V landed
    uint64_t* s_full = bars + 1;       // S ready
    uint64_t* p_full = bars + 2;       // P written (4 warps)
    uint64_t* o_full = bars + 3;       // O ready
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int nk16 = (p.Nk + 15) & ~15;             // keys the MMAs see (zero rows past Nk)

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_k);
        tma_prefetch_desc(&tm_v);
        mbar_init(ld_full, 1);
        mbar_init(s_full, 1);
        mbar_init(p_full, 4);
        mbar_init(o_full, 1);
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 4) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 4) {
        if (elect_one_sync()) {
            mbar_expect_tx(ld_full, 3 * kATile);
            tma_load_3d(&tm_q, s_q, ld_full, head * kAD, qt * kAQ, b);
            tma_load_3d(&tm_k, s_k, ld_full, head * kAD, 0, b);
            tma_load_3d(&tm_v, s_v, ld_full, head * kAD, 0, b);
            mbar_wait(ld_full, 0);
            tc_fence_after();
            const uint32_t idesc_s = make_idesc_mn(kAQ, nk16, false);
            const uint64_t dq = make_smem_desc(smem_u32(s_q)), dk = make_smem_desc(smem_u32(s_k));
#pragma unroll
            for (int k = 0; k < kAD / 16; ++k) umma_bf16(tmem_base + kAXColS, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
            umma_commit(s_full);
            mbar_wait(p_full, 0);
            tc_fence_after();
            constexpr uint32_t idesc_o = make_idesc_mn(kAQ, kAD, true);
            const uint64_t dv = make_smem_desc_mn(smem_u32(s_v));
            for (int ks = 0; ks < nk16 / 16; ++ks)
                umma_bf16_ts(tmem_base + kAXColO, tmem_base + kAXColS + ks * 8, dv + (uint64_t)ks * (2048 >> 4), idesc_o, ks != 0);
            umma_commit(o_full);
        }
    } else {
        const int row = warp * 32 + lane;                       // warp w owns TMEM lanes 32w .. 32w+31
        const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
        const uint32_t ts = tmem_base + lane_addr + kAXColS, to = tmem_base + lane_addr + kAXColO;
        const float sc = p.scale_log2;
        mbar_wait(s_full, 0);
        tc_fence_after();
        uint32_t sv[kAK];
#pragma unroll
        for (int c = 0; c < kAK; c += 16)
            if (c < nk16) tmem_ld16(ts + c, sv + c);
        tmem_ld_wait();
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int c = 0; c < kAK; c += 2) {
            if (c < p.Nk) mx0 = fmaxf(mx0, __uint_as_float(sv[c]));
            if (c + 1 < p.Nk) mx1 = fmaxf(mx1, __uint_as_float(sv[c + 1]));
        }
        const float m = fmaxf(mx0, mx1) * sc;
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int c16 = 0; c16 < kAK; c16 += 32) {               // 32 keys -> 16 packed columns
            if (c16 < nk16) {
                uint32_t pk[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const int k0 = c16 + 2 * c;
                    const float p0 = k0 < p.Nk ? ex2_approx(fmaf(__uint_as_float(sv[k0]), sc, -m)) : 0.f;
                    const float p1 = k0 + 1 < p.Nk ? ex2_approx(fmaf(__uint_as_float(sv[k0 + 1]), sc, -m)) : 0.f;
                    l0 += p0; l1 += p1;
                    pk[c] = pack_bf16x2(p0, p1);
                }
                tmem_st16(ts + (c16 >> 1), pk);                 // P overwrites the S columns this thread has already read
            }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        const float inv = 1.0f / (l0 + l1);
        mbar_wait(o_full, 0);
        tc_fence_after();
        const int64_t qrow = (int64_t)qt * kAQ + row;
        __nv_bfloat16* dst = p.out + (int64_t)b * p.o_bs + qrow * p.ldo + head * kAD;
#pragma unroll
        for (int c2 = 0; c2 < kAD; c2 += 32) {
            uint32_t r[32];
            tmem_ld16(to + c2, r);
            tmem_ld16(to + c2 + 16, r + 16);
            tmem_ld_wait();
            if (qrow < p.Nq) {
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
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

// host: single-KV-tile tcgen05 kernel (cross-attention): d = 64, Nk <= 128, any Nq
bool attention_x_supported(int d, int Nq, int Nk, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                           int64_t q_bs, int64_t k_bs, int64_t v_bs, int64_t o_bs, const void* q, const void* k, const void* v,
                           const void* out) {
    return d == kAD && Nk <= kAK && Nk >= 1 && Nq >= 1 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 &&
           q_bs % 8 == 0 && k_bs % 8 == 0 && v_bs % 8 == 0 && o_bs % 8 == 0 &&
           (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) == 0;
}

int launch_attention_x(const void* q, const void* k, const void* v, void* out, int B, int heads, int Nq, int Nk,
                       int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t q_bs, int64_t k_bs,
                       int64_t v_bs, int64_t o_bs, float scale, cudaStream_t stream) {
    CUtensorMap tq, tk, tv;
    const uint32_t box[3] = {(uint32_t)kAD, (uint32_t)kAQ, 1};
    auto mk = [&](CUtensorMap* m, const void* ptr, int N, int64_t ld, int64_t bs, const char* what) -> int {
        const uint64_t bstride = (B > 1 ? (uint64_t)bs : (uint64_t)ld * N) * 2;
        uint64_t dims[3] = {(uint64_t)heads * kAD, (uint64_t)N, (uint64_t)B};
        uint64_t str[2] = {(uint64_t)ld * 2, bstride};
        return encode_map(m, ptr, 3, dims, str, box, what);
    };
    if (int e = mk(&tq, q, Nq, ldq, q_bs, "xattn Q")) return e;
    if (int e = mk(&tk, k, Nk, ldk, k_bs, "xattn K")) return e;
    if (int e = mk(&tv, v, Nk, ldv, v_bs, "xattn V")) return e;
    static bool attr_set = false;
    if (!attr_set) {
        RDEIC_CUDA(cudaFuncSetAttribute(attention_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAXSmem));
        attr_set = true;
    }
    AttXDev d;
    d.Nq = Nq; d.Nk = Nk;
    d.scale_log2 = scale * 1.4426950408889634f;
    d.out = (__nv_bfloat16*)out;
    d.ldo = ldo; d.o_bs = o_bs;
    dim3 grid((Nq + kAQ - 1) / kAQ, heads, B);
    launch_k(attention_x_kernel, grid, kAXThreads, kAXSmem, stream, tq, tk, tv, d);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

}  // namespace rdeic
