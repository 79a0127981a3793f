// fp32 kernel mode of the UNet + control step (BASELINE north_star: per-step rel-L2 <= 1e-5 against the
// reference's fp32 output).  Plain fp32 NHWC tensors, fp32 FMA on the CUDA cores, no tensor cores: this
// is the verification mode of the path, not the throughput mode (that is conv_gemm.cu / attention_tc.cu).
//
// Accuracy choices: convolution / linear partial sums run in fp32 over 64-element k chunks and are
// folded into fp64 accumulators (a straight fp32 sum over K = 23 040 loses ~1e-5 by itself);
// softmax uses expf / full-precision division; GroupNorm / LayerNorm reuse the bf16 path's fp32-input
// kernels with an fp32 output (norm.cu).
//
// Reference semantics: torch conv2d / F.linear (openaimodel.py:203,229,240,106,150,566,750,
// attention.py:52,72,162-169,314,328), CrossAttention.forward (attention.py:171-203), GEGLU
// (attention.py:54-56), timestep_embedding (util.py:161-181).
#include "common.cuh"
#include "../../include/rdeic_b200.h"

namespace rdeic {

constexpr int kFM = 64, kFN = 64, kFK = 16;      // CTA tile; 256 threads, 4x4 outputs per thread

struct ConvF32Dev {
    const float* a; const float* a2;
    int n, h, w, c1, c2;            // input grid and channel counts of the two concat sources
    int a_ld, a2_ld;                // pixel strides (a / a2 may be channel windows of wider NHWC buffers)
    int ksize, stride, up;          // 1|3|5 ; 1|2 ; nearest x2 upsample before the conv (0|1)
    int oh, ow;
    const float* wt;                // [n_out][taps][c1 + c2]
    int n_out;
    const float* bias; const float* row_bias; int row_bias_ld;
    const float* resid; int ld_resid; float alpha; int act; float act_param;
    float* out; int ldo;
    double* partial;                // split-K: fp64 partial sums [split][M][n_out]; null = single pass
    int k_per_split;                // k extent of one split (a multiple of 64), = K without split-K
};

// bias, per-sample bias, activation, residual for one output element whose reduction is complete
__device__ __forceinline__ void conv_f32_finish(const ConvF32Dev& p, int64_t m, int n, double v) {
    if (p.bias) v += (double)p.bias[n];
    if (p.row_bias) v += (double)p.row_bias[(m / ((int64_t)p.oh * p.ow)) * p.row_bias_ld + n];
    float f = (float)v;
    if (p.act == 1) f = f / (1.0f + expf(-f));
    else if (p.act == 3) f = f > 0.f ? f : f * p.act_param;                        // LeakyReLU
    else if (p.act == 4) f = 0.5f * f * (1.0f + erff(f * 0.70710678118654752f));  // exact GELU (F.gelu default)
    if (p.resid) f = fmaf(p.alpha, f, p.resid[m * p.ld_resid + n]);
    else f *= p.alpha;
    p.out[m * p.ldo + n] = f;
}

__global__ void __launch_bounds__(256)
conv_f32_kernel(const ConvF32Dev p) {
    __shared__ float As[kFK][kFM + 4];
    __shared__ float Bs[kFK][kFN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int C = p.c1 + p.c2;
    const int taps = p.ksize * p.ksize;
    const int K = taps * C;
    const int64_t M = (int64_t)p.n * p.oh * p.ow;
    const int64_t m0 = (int64_t)blockIdx.x * kFM;
    const int n0 = blockIdx.y * kFN;
    const int pad = p.ksize >> 1;
    const int hu = p.h << p.up, wu = p.w << p.up;       // grid the filter slides over

    // this thread's A-load row (fixed for the whole k loop) and its 4 consecutive k
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    const int64_t lm = m0 + lr;
    int lb = 0, loy = 0, lox = 0;
    const bool lrow_ok = lm < M;
    if (lrow_ok) {
        lox = (int)(lm % p.ow);
        const int64_t t = lm / p.ow;
        loy = (int)(t % p.oh);
        lb = (int)(t / p.oh);
    }
    const int ln = n0 + lr;                              // B-load column (same thread mapping)

    double dacc[4][4];
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { dacc[i][j] = 0.0; acc[i][j] = 0.f; }

    // A tile: implicit im2col gather (zero padding, stride, nearest upsample); B tile: packed weights.
    // The next k chunk is fetched into registers while the current one is multiplied (the layers of the
    // learned compressor have M = a few thousand pixels: a handful of CTAs whose k loop is latency bound).
    auto fetch = [&](int k0, float4& av, float4& bv) {
        av = make_float4(0.f, 0.f, 0.f, 0.f);
        bv = make_float4(0.f, 0.f, 0.f, 0.f);
        const int k = k0 + lk;
        if (k >= K) return;
        if (lrow_ok) {
            const int tap = k / C, c = k - tap * C;
            const int iy = loy * p.stride + tap / p.ksize - pad, ix = lox * p.stride + tap % p.ksize - pad;
            if (iy >= 0 && iy < hu && ix >= 0 && ix < wu) {
                const int64_t pix = ((int64_t)lb * p.h + (iy >> p.up)) * p.w + (ix >> p.up);
                av = (c < p.c1) ? *reinterpret_cast<const float4*>(p.a + pix * p.a_ld + c)
                                : *reinterpret_cast<const float4*>(p.a2 + pix * p.a2_ld + (c - p.c1));
            }
        }
        if (ln < p.n_out) bv = *reinterpret_cast<const float4*>(p.wt + (int64_t)ln * K + k);
    };
    int since_flush = 0;
    float4 av, bv;
    const int k_begin = blockIdx.z * p.k_per_split;
    const int k_end = min(K, k_begin + p.k_per_split);
    fetch(k_begin, av, bv);
    for (int k0 = k_begin; k0 < k_end; k0 += kFK) {
        __syncthreads();
        As[lk][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
        Bs[lk][lr] = bv.x; Bs[lk + 1][lr] = bv.y; Bs[lk + 2][lr] = bv.z; Bs[lk + 3][lr] = bv.w;
        __syncthreads();
        if (k0 + kFK < k_end) fetch(k0 + kFK, av, bv);
#pragma unroll
        for (int kk = 0; kk < kFK; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a_[4] = {a4.x, a4.y, a4.z, a4.w}, b_[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a_[i], b_[j], acc[i][j]);
        }
        if (++since_flush == 4) {                        // 64 k per fp32 partial sum
            since_flush = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { dacc[i][j] += (double)acc[i][j]; acc[i][j] = 0.f; }
        }
    }
    // ---- epilogue: bias, per-sample bias, activation, residual (or the fp64 partial of this k split) ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.n_out) continue;
            const double v = dacc[i][j] + (double)acc[i][j];
            if (p.partial) p.partial[((int64_t)blockIdx.z * M + m) * p.n_out + n] = v;
            else conv_f32_finish(p, m, n, v);
        }
    }
}

// split-K second pass: the k slices are summed in a fixed order (deterministic: the CDF indexes built while
// compressing and while decompressing must be identical), then the same epilogue
__global__ void __launch_bounds__(256)
conv_f32_reduce_kernel(const ConvF32Dev p, int splits) {
    const int64_t M = (int64_t)p.n * p.oh * p.ow;
    const int64_t total = M * p.n_out;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        double v = 0.0;
        for (int z = 0; z < splits; ++z) v += p.partial[(int64_t)z * total + i];
        conv_f32_finish(p, i / p.n_out, (int)(i % p.n_out), v);
    }
}

// One warp per (query row, head, batch): online softmax in fp32, lane l owns head dims l, l+32, ...
// (kDL per lane: 2 for the UNet's d <= 64 heads, 16 for the VAE's single 512-wide head); logits are
// reduced across the warp with shuffles.
template <int kDL>
__global__ void __launch_bounds__(128)
attention_f32_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                     float* __restrict__ out, int heads, int Nq, int Nk, int d, int64_t ldq, int64_t ldk,
                     int64_t ldv, int64_t ldo, int64_t q_bs, int64_t k_bs, int64_t v_bs, int64_t o_bs, float scale) {
    const int lane = threadIdx.x & 31;
    const int64_t qi = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (qi >= Nq) return;
    const int head = blockIdx.y, b = blockIdx.z;
    const float* qp = q + b * q_bs + qi * ldq + head * d;
    const float* kp = k + b * k_bs + head * d;
    const float* vp = v + b * v_bs + head * d;
    float qv[kDL], o[kDL];
#pragma unroll
    for (int i = 0; i < kDL; ++i) {
        const int dd = lane + 32 * i;
        qv[i] = dd < d ? qp[dd] * scale : 0.f;
        o[i] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < Nk; ++j) {
        const float* kr = kp + (int64_t)j * ldk;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < kDL; ++i) {
            const int dd = lane + 32 * i;
            if (dd < d) s = fmaf(qv[i], kr[dd], s);
        }
#pragma unroll
        for (int o_ = 16; o_; o_ >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o_);
        const float m_new = fmaxf(m, s);
        const float corr = expf(m - m_new), pj = expf(s - m_new);
        const float* vr = vp + (int64_t)j * ldv;
        l = l * corr + pj;
#pragma unroll
        for (int i = 0; i < kDL; ++i) {
            const int dd = lane + 32 * i;
            o[i] = o[i] * corr + (dd < d ? pj * vr[dd] : 0.f);
        }
        m = m_new;
    }
    float* op = out + b * o_bs + qi * ldo + head * d;
#pragma unroll
    for (int i = 0; i < kDL; ++i) {
        const int dd = lane + 32 * i;
        if (dd < d) op[dd] = o[i] / l;
    }
}

// attention.py:54-56: out = value * gelu(gate), in [rows, 2F] (value | gate), exact erf
__global__ void __launch_bounds__(256)
geglu_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int F) {
    const int64_t total = rows * F;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / F;
        const int c = (int)(i - r * F);
        const float a = in[r * 2 * F + c], g = in[r * 2 * F + F + c];
        out[i] = a * (0.5f * g * (1.0f + erff(g * 0.70710678118654752f)));
    }
}

__global__ void timestep_embedding_f32_kernel(const long long* __restrict__ t, float* __restrict__ out, int B,
                                              int dim, float max_period) {
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * dim) return;
    const int b = i / dim, j = i - b * dim;
    float v = 0.f;
    if (j < 2 * half) {
        const int kk = j < half ? j : j - half;
        const float f = expf(-logf(max_period) * (float)kk / (float)half);
        const float arg = (float)t[b] * f;
        v = j < half ? cosf(arg) : sinf(arg);
    }
    out[i] = v;
}

}  // namespace rdeic

using namespace rdeic;

extern "C" {

int rdeic_conv_f32(const rdeic_conv_f32_params* p, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(p && p->a && p->w && p->out, "rdeic_conv_f32: null pointer");
    RDEIC_CHECK_ARG(p->a_n > 0 && p->a_h > 0 && p->a_w > 0 && p->c1 > 0 && p->c2 >= 0 && p->n_out > 0,
                    "rdeic_conv_f32: bad dims");
    RDEIC_CHECK_ARG(p->c1 % 4 == 0 && p->c2 % 4 == 0, "rdeic_conv_f32: channel counts (%d, %d) must be multiples of 4",
                    p->c1, p->c2);
    RDEIC_CHECK_ARG(p->c2 == 0 || p->a2, "rdeic_conv_f32: c2 > 0 needs a2");
    RDEIC_CHECK_ARG((p->ksize == 1 || p->ksize == 3 || p->ksize == 5) && (p->stride == 1 || p->stride == 2) &&
                        (p->up == 0 || p->up == 1),
                    "rdeic_conv_f32: ksize in {1,3,5}, stride in {1,2}, up in {0,1}");
    RDEIC_CHECK_ARG(p->act == 0 || p->act == 1 || p->act == 3 || p->act == 4,
                    "rdeic_conv_f32: act must be 0 (none), 1 (SiLU), 3 (LeakyReLU) or 4 (exact GELU)");
    const int a_ld = p->a_ld ? p->a_ld : p->c1, a2_ld = p->a2_ld ? p->a2_ld : p->c2;
    RDEIC_CHECK_ARG(a_ld >= p->c1 && a2_ld >= p->c2 && a_ld % 4 == 0 && a2_ld % 4 == 0,
                    "rdeic_conv_f32: a_ld/a2_ld (%d, %d) must be multiples of 4 and >= the channel counts", a_ld, a2_ld);
    RDEIC_CHECK_ARG(((uintptr_t)p->a | (uintptr_t)p->a2 | (uintptr_t)p->w | (uintptr_t)p->workspace) % 16 == 0,
                    "rdeic_conv_f32: operands must be 16-byte aligned");
    ConvF32Dev d;
    d.a = p->a; d.a2 = p->a2; d.n = p->a_n; d.h = p->a_h; d.w = p->a_w; d.c1 = p->c1; d.c2 = p->c2;
    d.a_ld = a_ld; d.a2_ld = a2_ld;
    d.ksize = p->ksize; d.stride = p->stride; d.up = p->up;
    const int hu = p->a_h << p->up, wu = p->a_w << p->up;
    RDEIC_CHECK_ARG(hu % p->stride == 0 && wu % p->stride == 0, "rdeic_conv_f32: grid not divisible by the stride");
    d.oh = hu / p->stride; d.ow = wu / p->stride;
    d.wt = p->w; d.n_out = p->n_out; d.bias = p->bias; d.row_bias = p->row_bias; d.row_bias_ld = p->row_bias_ld;
    d.resid = p->resid; d.ld_resid = p->ld_resid; d.alpha = p->alpha; d.act = p->act; d.act_param = p->act_param; d.out = p->out; d.ldo = p->ldo;
    RDEIC_CHECK_ARG(p->ldo >= p->n_out && (!p->resid || p->ld_resid >= p->n_out), "rdeic_conv_f32: bad ldo / ld_resid");
    const int64_t M = (int64_t)d.n * d.oh * d.ow;
    const int K = p->ksize * p->ksize * (p->c1 + p->c2);
    dim3 grid((unsigned)ceil_div64(M, kFM), (unsigned)((p->n_out + kFN - 1) / kFN));
    // few output tiles and a long reduction (the compressor's 5x5 context nets: M = a few thousand pixels, K up to
    // 6000): slice K across blockIdx.z so the chip is not left to a handful of CTAs walking hundreds of k chunks
    int splits = 1;
    d.partial = nullptr;
    d.k_per_split = (K + 63) / 64 * 64;
    const int64_t tiles = (int64_t)grid.x * grid.y;
    if (p->workspace && tiles * 2 <= kNumSMs && K >= 1024) {
        int want = (int)(2 * kNumSMs / tiles);
        if (want > K / 256) want = K / 256;
        if (want > 16) want = 16;
        while (want > 1 && (int64_t)want * M * p->n_out * (int64_t)sizeof(double) > p->workspace_bytes) --want;
        if (want >= 2) {
            d.k_per_split = ((K + want - 1) / want + 63) / 64 * 64;
            splits = (K + d.k_per_split - 1) / d.k_per_split;
            d.partial = reinterpret_cast<double*>(p->workspace);
        }
    }
    grid.z = (unsigned)splits;
    conv_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(d);
    RDEIC_LAUNCH_CHECK();
    if (splits > 1) {
        conv_f32_reduce_kernel<<<grid_for(M * p->n_out, 256), 256, 0, as_stream(stream)>>>(d, splits);
        RDEIC_LAUNCH_CHECK();
    }
    return 0;
}

int rdeic_attention_f32(const float* q, const float* k, const float* v, float* out, int B, int heads, int Nq,
                        int Nk, int d, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t q_bs,
                        int64_t k_bs, int64_t v_bs, int64_t o_bs, float scale, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(q && k && v && out, "rdeic_attention_f32: null pointer");
    RDEIC_CHECK_ARG(B > 0 && heads > 0 && Nq > 0 && Nk > 0 && d > 0 && d <= 512, "rdeic_attention_f32: bad dims (d <= 512)");
    dim3 grid((unsigned)((Nq + 3) / 4), (unsigned)heads, (unsigned)B);
    if (d <= 64)
        attention_f32_kernel<2><<<grid, 128, 0, as_stream(stream)>>>(q, k, v, out, heads, Nq, Nk, d, ldq, ldk, ldv, ldo,
                                                                      q_bs, k_bs, v_bs, o_bs, scale);
    else
        attention_f32_kernel<16><<<grid, 128, 0, as_stream(stream)>>>(q, k, v, out, heads, Nq, Nk, d, ldq, ldk, ldv, ldo,
                                                                       q_bs, k_bs, v_bs, o_bs, scale);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_geglu_f32(const float* in, float* out, int64_t rows, int F, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(in && out && rows >= 0 && F > 0, "rdeic_geglu_f32: bad args");
    if (rows == 0) return 0;
    geglu_f32_kernel<<<grid_for(rows * F, 256), 256, 0, as_stream(stream)>>>(in, out, rows, F);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_timestep_embedding_f32(const int64_t* t, float* out, int B, int dim, float max_period,
                                 rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(t && out && B > 0 && dim > 0, "rdeic_timestep_embedding_f32: bad args");
    const int n = B * dim;
    timestep_embedding_f32_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>((const long long*)t, out, B, dim, max_period);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
