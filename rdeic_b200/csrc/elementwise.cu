// Bandwidth-bound glue kernels of the relay decode path: sampler updates, layout changes,
// GEGLU, nearest upsample, stride-2 im2col, row softmax, transpose, uint8 post-process.
#include "common.cuh"
#include "../../include/rdeic_b200.h"

namespace rdeic {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------
// sampler updates.  Every product/sum is rounded separately (no FMA contraction) so the
// result is bit-identical to the reference's chain of ATen elementwise ops.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float guided_eps(float e, float eu, float scale, bool has_uncond) {
    // spaced_sampler_relay.py:283  model_uncond + scale * (model_t - model_uncond)
    return has_uncond ? __fadd_rn(eu, __fmul_rn(scale, __fsub_rn(e, eu))) : e;
}

__global__ void __launch_bounds__(kThreads)
q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                float* __restrict__ out, int64_t n, float a, float b) {
    pdl_trigger_short();
    pdl_wait();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        out[i] = __fadd_rn(__fmul_rn(a, x0[i]), __fmul_rn(b, noise[i]));
}

__global__ void __launch_bounds__(kThreads)
relay_update_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                    const float* __restrict__ eps_u, float gscale,
                    const float* __restrict__ noise, float* __restrict__ out, int64_t n,
                    float r, float rm1, float c1, float c2, float sigma) {
    pdl_trigger_short();
    pdl_wait();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float xv = x[i];
        const float e = guided_eps(eps[i], eps_u ? eps_u[i] : 0.f, gscale, eps_u != nullptr);
        // :270-275 pred_x0 = sqrt_recip*x - sqrt_recipm1*eps
        const float pred = __fsub_rn(__fmul_rn(r, xv), __fmul_rn(rm1, e));
        // :154-159 mean = coef1*pred + coef2*x
        const float mean = __fadd_rn(__fmul_rn(c1, pred), __fmul_rn(c2, xv));
        // :383 x_prev = mean + (mask*sqrt(var)) * noise   (sigma already = mask*sqrt(var))
        out[i] = __fadd_rn(mean, __fmul_rn(sigma, noise[i]));
    }
}

__global__ void __launch_bounds__(kThreads)
ddim_update_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                   const float* __restrict__ eps_u, float gscale,
                   const float* __restrict__ noise, float* __restrict__ out,
                   float* __restrict__ pred_out, int64_t n, float s1m, float sqrt_at,
                   float sqrt_aprev, float dir, float sigma) {
    pdl_trigger_short();
    pdl_wait();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float e = guided_eps(eps[i], eps_u ? eps_u[i] : 0.f, gscale, eps_u != nullptr);
        // ddim_sampler_relay.py:214 pred_x0 = (x - sqrt_one_minus_at * e_t) / a_t.sqrt()
        const float pred = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(s1m, e)), sqrt_at);
        // :225-230 x_prev = a_prev.sqrt()*pred_x0 + dir_xt + noise
        const float t0 = __fadd_rn(__fmul_rn(sqrt_aprev, pred), __fmul_rn(dir, e));
        out[i] = __fadd_rn(t0, __fmul_rn(sigma, noise[i]));
        if (pred_out) pred_out[i] = pred;
    }
}

// ------------------------------------------------------------------------------------------
// layout conversion
// ------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC bf16 (window).  Tile transpose through shared memory: 32 pixels x 32 ch.
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C,
                    int64_t HW, int ldc, int c_off) {
    pdl_trigger_short();
    pdl_wait();
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows of 32
    for (int cc = ty; cc < 32; cc += 8) {
        const int c = c0 + cc;
        const int64_t p = p0 + tx;
        tile[cc][tx] = (c < C && p < HW) ? src[((int64_t)b * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int pp = ty; pp < 32; pp += 8) {
        const int64_t p = p0 + pp;
        const int c = c0 + tx;
        if (p < HW && c < C)
            dst[((int64_t)b * HW + p) * ldc + c_off + c] = __float2bfloat16_rn(tile[tx][pp]);
    }
}

template <bool kF32>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const void* __restrict__ src, float* __restrict__ dst, int C, int64_t HW,
                    int ldc) {
    pdl_trigger_short();
    pdl_wait();
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int64_t p0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int pp = ty; pp < 32; pp += 8) {
        const int64_t p = p0 + pp;
        const int c = c0 + tx;
        float v = 0.f;
        if (p < HW && c < C) {
            const int64_t idx = ((int64_t)b * HW + p) * ldc + c;
            v = kF32 ? reinterpret_cast<const float*>(src)[idx]
                     : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[idx]);
        }
        tile[pp][tx] = v;
    }
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
        const int c = c0 + cc;
        const int64_t p = p0 + tx;
        if (c < C && p < HW) dst[((int64_t)b * C + c) * HW + p] = tile[tx][cc];
    }
}

__global__ void __launch_bounds__(kThreads)
f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    pdl_trigger_short();
    pdl_wait();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16_rn(src[i]);
}

// util.py:161-181: freqs = exp(-ln(max_period) * k / half); emb = [cos(t*f) | sin(t*f)]
__global__ void timestep_embedding_kernel(const long long* __restrict__ t,
                                          __nv_bfloat16* __restrict__ out, int B, int dim,
                                          float max_period) {
    pdl_trigger_short();
    pdl_wait();
    const int half = dim / 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * dim) return;
    const int b = i / dim, j = i - b * dim;
    float v = 0.f;
    if (j < 2 * half) {
        const int k = j < half ? j : j - half;
        const float f = expf(-logf(max_period) * (float)k / (float)half);
        const float arg = (float)t[b] * f;
        v = j < half ? cosf(arg) : sinf(arg);
    }
    out[i] = __float2bfloat16_rn(v);
}

template <bool kF32>
__global__ void __launch_bounds__(kThreads)
silu_kernel(const void* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n) {
    pdl_trigger_short();
    pdl_wait();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float v = kF32 ? reinterpret_cast<const float*>(x)[i]
                             : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[i]);
        out[i] = __float2bfloat16_rn(silu_f(v));
    }
}

// attention.py:54-56: x, gate = proj(x).chunk(2); x * gelu(gate)   (exact erf GELU)
__global__ void __launch_bounds__(kThreads)
geglu_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int64_t rows, int F) {
    pdl_trigger_short();
    pdl_wait();
    const int fv = F >> 3;  // vectors of 8 bf16
    const int64_t total = rows * fv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / fv;
        const int j = (int)(i - r * fv);
        const uint4 xv = ld_stream_u4(in + r * (2 * fv) + j);
        const uint4 gv = ld_stream_u4(in + r * (2 * fv) + fv + j);
        const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
        const uint32_t gs[4] = {gv.x, gv.y, gv.z, gv.w};
        uint32_t os[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float x0, x1, g0, g1;
            unpack_bf16x2(xs[k], x0, x1);
            unpack_bf16x2(gs[k], g0, g1);
            const float a0 = 0.5f * g0 * (1.0f + erff(g0 * 0.70710678118654752f));
            const float a1 = 0.5f * g1 * (1.0f + erff(g1 * 0.70710678118654752f));
            os[k] = pack_bf16x2(x0 * a0, x1 * a1);
        }
        st_stream_u4(out + i, make_uint4(os[0], os[1], os[2], os[3]));
    }
}

// nearest x2 upsample, NHWC bf16: one thread per INPUT 16-byte vector, four streaming stores (the 2x2
// output pixels) — a quarter of the loads and index divisions of an output-indexed copy
__global__ void __launch_bounds__(kThreads)
upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W,
                  int cv) {
    pdl_trigger_short();
    pdl_wait();
    const int64_t total = (int64_t)B * H * W * cv;
    const int64_t orow = (int64_t)2 * W * cv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv);
        int64_t p = i / cv;
        const int iw = (int)(p % W); p /= W;            // p = b*H + ih
        const uint4 v = ld_stream_u4(in + i);
        uint4* o = out + (2 * p) * orow + (int64_t)(2 * iw) * cv + c;
        st_stream_u4(o, v);
        st_stream_u4(o + cv, v);
        st_stream_u4(o + orow, v);
        st_stream_u4(o + orow + cv, v);
    }
}

// PixelShuffle(2) on NHWC bf16 whose 4C input channels are ordered (i, j, c) (the sub-pixel conv's
// weight rows are permuted that way at load): out[b, 2h+i, 2w+j, c] = in[b, h, w, (2i+j)C + c]
__global__ void __launch_bounds__(kThreads)
pixel_shuffle2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W,
                      int cv) {
    pdl_trigger_short();
    pdl_wait();
    const int64_t total = (int64_t)B * (2 * H) * (2 * W) * cv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv);
        int64_t p = i / cv;
        const int ow = (int)(p % (2 * W)); p /= (2 * W);
        const int oh = (int)(p % (2 * H));
        const int b = (int)(p / (2 * H));
        const int sub = ((oh & 1) << 1) | (ow & 1);
        out[i] = in[((((int64_t)b * H + (oh >> 1)) * W + (ow >> 1)) * 4 + sub) * cv + c];
    }
}

// stride-2 pad-1 3x3 im2col: out row m = (b, oh, ow), col = tap*Cp + c
__global__ void __launch_bounds__(kThreads)
im2col_s2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B, int H, int W,
                 int C, int Cp, int pad_lo) {
    pdl_trigger_short();
    pdl_wait();
    const int Ho = H / 2, Wo = W / 2;
    const int cpv = Cp >> 3, cv = C >> 3;
    const int64_t total = (int64_t)B * Ho * Wo * 9 * cpv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % cpv);
        int64_t q = i / cpv;
        const int tap = (int)(q % 9); q /= 9;
        const int ow = (int)(q % Wo); q /= Wo;
        const int oh = (int)(q % Ho);
        const int b = (int)(q / Ho);
        const int ih = 2 * oh + tap / 3 - pad_lo, iw = 2 * ow + tap % 3 - pad_lo;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (c < cv && ih >= 0 && ih < H && iw >= 0 && iw < W)
            v = in[(((int64_t)b * H + ih) * W + iw) * cv + c];
        out[i] = v;
    }
}

// row softmax: one block per row, row cached in registers (n <= 64K) via strided loop
template <bool kF32>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const void* __restrict__ in, __nv_bfloat16* __restrict__ out, int n,
                    float scale) {
    pdl_trigger_short();
    pdl_wait();
    const int64_t row = blockIdx.x;
    const float* inf = reinterpret_cast<const float*>(in) + row * n;
    const __nv_bfloat16* inh = reinterpret_cast<const __nv_bfloat16*>(in) + row * n;
    __shared__ float red[32];
    float m = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = (kF32 ? inf[i] : __bfloat162float(inh[i])) * scale;
        m = fmaxf(m, v);
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float s = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = (kF32 ? inf[i] : __bfloat162float(inh[i])) * scale;
        s += __expf(v - m);
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    const float inv = 1.0f / s;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = (kF32 ? inf[i] : __bfloat162float(inh[i])) * scale;
        out[row * n + i] = __float2bfloat16_rn(__expf(v - m) * inv);
    }
}

// fp32 rows of n <= 8192 (n % 4 == 0): the row lives in registers (up to 8 float4 per thread), one
// 16-byte read and one 8-byte write per 4 elements — the VAE mid-attention logits [B*N, N] are read
// once instead of three times.
template <int kVec>
__global__ void __launch_bounds__(256)
softmax_rows_f32_reg_kernel(const float4* __restrict__ in, uint2* __restrict__ out, int n4, float scale) {
    pdl_trigger_short();
    pdl_wait();
    const int64_t row = blockIdx.x;
    const float4* src = in + row * n4;
    __shared__ float red[8];
    float4 v[kVec];
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
        const int i = threadIdx.x + 256 * j;
        v[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (i < n4) {
            const uint4 u = ld_stream_u4(src + i);
            const float4 t = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
            v[j] = make_float4(t.x * scale, t.y * scale, t.z * scale, t.w * scale);
        }
        m = fmaxf(m, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
        v[j].x = __expf(v[j].x - m); v[j].y = __expf(v[j].y - m);
        v[j].z = __expf(v[j].z - m); v[j].w = __expf(v[j].w - m);
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w];
    const float inv = 1.0f / s;
#pragma unroll
    for (int j = 0; j < kVec; ++j) {
        const int i = threadIdx.x + 256 * j;
        if (i < n4)
            out[row * n4 + i] = make_uint2(pack_bf16x2(v[j].x * inv, v[j].y * inv), pack_bf16x2(v[j].z * inv, v[j].w * inv));
    }
}

__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                      int R, int C) {
    pdl_trigger_short();
    pdl_wait();
    __shared__ __nv_bfloat16 tile[32][34];
    const int64_t base = (int64_t)blockIdx.z * R * C;
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int rr = ty; rr < 32; rr += 8) {
        const int r = r0 + rr, c = c0 + tx;
        if (r < R && c < C) tile[rr][tx] = in[base + (int64_t)r * C + c];
    }
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
        const int c = c0 + cc, r = r0 + tx;
        if (r < R && c < C) out[base + (int64_t)c * R + r] = tile[tx][cc];
    }
}

// inference.py:85-87: ((x+1)/2).clamp(0,1) * 255 -> clip(0,255) -> uint8 (truncation)
__global__ void __launch_bounds__(kThreads)
image_to_u8_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int64_t pixels,
                   int ldc) {
    pdl_trigger_short();
    pdl_wait();
    const int64_t total = pixels * 3;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / 3;
        const int c = (int)(i - p * 3);
        float v = __fdiv_rn(__fadd_rn(in[p * ldc + c], 1.0f), 2.0f);
        v = fminf(fmaxf(v, 0.f), 1.f);
        v = __fmul_rn(v, 255.0f);
        v = fminf(fmaxf(v, 0.f), 255.f);
        out[i] = (uint8_t)__float2int_rz(v);
    }
}

// x (fp32) -> hi = bf16(x), lo = bf16(x - hi), written as two channel windows of one bf16 row: a GEMM over [hi | lo]
// with the weight columns duplicated sees x to ~16 mantissa bits.  Used for the tensors that ENTER the network
// (latent, hint, text context, timestep embedding): their bf16 rounding error would otherwise ride along every
// skip connection to the output.
__global__ void __launch_bounds__(kThreads)
split_bf16_hilo_kernel(const float* __restrict__ src, int64_t rows, int C, int64_t ld_src,
                       __nv_bfloat16* __restrict__ dst, int64_t ld_dst, int off_hi, int off_lo) {
    pdl_trigger_short();
    pdl_wait();
    const int64_t total = rows * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / C;
        const int c = (int)(i - r * C);
        const float x = src[r * ld_src + c];
        const __nv_bfloat16 hi = __float2bfloat16_rn(x);
        dst[r * ld_dst + off_hi + c] = hi;
        dst[r * ld_dst + off_lo + c] = __float2bfloat16_rn(x - __bfloat162float(hi));
    }
}

// Blend overlapping decoded tiles of one large image (BASELINE config 4; rdeic_b200/parallel.py).  One thread per
// output pixel: every tile that covers it contributes with a separable linear ramp over the overlap band
// (weight (i+1)/(ov+1) for the first ov pixels of a tile axis, mirrored at the far end, 1 in between), the
// weighted mean is rounded to uint8.  Tiles are uint8 HWC as the fused VAE tail wrote them.
__device__ __forceinline__ float ramp_w(int i, int n, int ov) {
    if (i < ov) return (float)(i + 1) / (float)(ov + 1);
    if (i >= n - ov) return (float)(n - i) / (float)(ov + 1);
    return 1.0f;
}

__global__ void __launch_bounds__(kThreads)
blend_tiles_u8_kernel(const uint8_t* __restrict__ tiles, const int* __restrict__ origin, int T, int th, int tw, int ov,
                      uint8_t* __restrict__ out, int H, int W) {
    pdl_trigger_short();
    pdl_wait();
    const int64_t total = (int64_t)H * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (int64_t)y * W);
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, ws = 0.f;
        for (int t = 0; t < T; ++t) {
            const int ly = y - origin[2 * t], lx = x - origin[2 * t + 1];
            if (ly < 0 || ly >= th || lx < 0 || lx >= tw) continue;
            const float w = ramp_w(ly, th, ov) * ramp_w(lx, tw, ov);
            const uint8_t* px = tiles + (((int64_t)t * th + ly) * tw + lx) * 3;
            acc0 = fmaf(w, (float)px[0], acc0); acc1 = fmaf(w, (float)px[1], acc1); acc2 = fmaf(w, (float)px[2], acc2);
            ws += w;
        }
        const float inv = ws > 0.f ? 1.0f / ws : 0.f;
        uint8_t* o = out + i * 3;
        o[0] = (uint8_t)fminf(255.f, rintf(acc0 * inv)); o[1] = (uint8_t)fminf(255.f, rintf(acc1 * inv));
        o[2] = (uint8_t)fminf(255.f, rintf(acc2 * inv));
    }
}

}  // namespace rdeic

using namespace rdeic;

extern "C" {

int rdeic_q_sample(const float* x0, const float* noise, float* out, int64_t numel, float a,
                   float b, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(x0 && noise && out, "rdeic_q_sample: null pointer");
    RDEIC_CHECK_ARG(numel >= 0, "rdeic_q_sample: negative numel");
    if (numel == 0) return 0;
    launch_k(q_sample_kernel, grid_for(numel, kThreads), kThreads, 0, as_stream(stream), x0, noise, out, numel, a, b);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_relay_update(const float* x, const float* eps, const float* eps_uncond,
                       float guidance_scale, const float* noise, float* out, int64_t numel,
                       float r, float rm1, float c1, float c2, float sigma,
                       rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(x && eps && noise && out, "rdeic_relay_update: null pointer");
    RDEIC_CHECK_ARG(numel >= 0, "rdeic_relay_update: negative numel");
    if (numel == 0) return 0;
    launch_k(relay_update_kernel, grid_for(numel, kThreads), kThreads, 0, as_stream(stream), 
        x, eps, eps_uncond, guidance_scale, noise, out, numel, r, rm1, c1, c2, sigma);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_ddim_update(const float* x, const float* eps, const float* eps_uncond,
                      float guidance_scale, const float* noise, float* out, float* pred_x0_out,
                      int64_t numel, float sqrt_one_minus_at, float sqrt_at, float sqrt_aprev,
                      float dir_coef, float sigma, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(x && eps && noise && out, "rdeic_ddim_update: null pointer");
    RDEIC_CHECK_ARG(numel >= 0, "rdeic_ddim_update: negative numel");
    if (numel == 0) return 0;
    launch_k(ddim_update_kernel, grid_for(numel, kThreads), kThreads, 0, as_stream(stream), 
        x, eps, eps_uncond, guidance_scale, noise, out, pred_x0_out, numel, sqrt_one_minus_at,
        sqrt_at, sqrt_aprev, dir_coef, sigma);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_nchw_to_nhwc_bf16(const float* src, void* dst, int B, int C, int H, int W, int ldc,
                            int c_off, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(src && dst, "rdeic_nchw_to_nhwc_bf16: null pointer");
    RDEIC_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && c_off >= 0 && c_off + C <= ldc,
                    "rdeic_nchw_to_nhwc_bf16: bad dims");
    const int64_t HW = (int64_t)H * W;
    dim3 grid((unsigned)ceil_div64(HW, 32), (unsigned)((C + 31) / 32), (unsigned)B);
    launch_k(nchw_to_nhwc_kernel, grid, 256, 0, as_stream(stream), src, (__nv_bfloat16*)dst, C, HW, ldc, c_off);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_nhwc_to_nchw_f32(const void* src, int src_is_f32, float* dst, int B, int C, int H,
                           int W, int ldc, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(src && dst, "rdeic_nhwc_to_nchw_f32: null pointer");
    RDEIC_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0 && C <= ldc, "rdeic_nhwc_to_nchw_f32: bad dims");
    const int64_t HW = (int64_t)H * W;
    dim3 grid((unsigned)ceil_div64(HW, 32), (unsigned)((C + 31) / 32), (unsigned)B);
    if (src_is_f32) launch_k(nhwc_to_nchw_kernel<true>, grid, 256, 0, as_stream(stream), src, dst, C, HW, ldc);
    else launch_k(nhwc_to_nchw_kernel<false>, grid, 256, 0, as_stream(stream), src, dst, C, HW, ldc);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_f32_to_bf16(const float* src, void* dst, int64_t numel, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(src && dst && numel >= 0, "rdeic_f32_to_bf16: bad args");
    if (numel == 0) return 0;
    launch_k(f32_to_bf16_kernel, grid_for(numel, kThreads), kThreads, 0, as_stream(stream), src, (__nv_bfloat16*)dst, numel);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_timestep_embedding(const int64_t* t, void* out_bf16, int B, int dim, float max_period,
                             rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(t && out_bf16 && B > 0 && dim > 0, "rdeic_timestep_embedding: bad args");
    const int n = B * dim;
    launch_k(timestep_embedding_kernel, (n + 255) / 256, 256, 0, as_stream(stream), 
        (const long long*)t, (__nv_bfloat16*)out_bf16, B, dim, max_period);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_silu_bf16(const void* x, int x_is_f32, void* out, int64_t numel,
                    rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(x && out && numel >= 0, "rdeic_silu_bf16: bad args");
    if (numel == 0) return 0;
    if (x_is_f32) launch_k(silu_kernel<true>, grid_for(numel, kThreads), kThreads, 0, as_stream(stream), x, (__nv_bfloat16*)out, numel);
    else launch_k(silu_kernel<false>, grid_for(numel, kThreads), kThreads, 0, as_stream(stream), x, (__nv_bfloat16*)out, numel);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_geglu(const void* in_bf16, void* out_bf16, int64_t rows, int F,
                rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(in_bf16 && out_bf16 && rows >= 0, "rdeic_geglu: bad args");
    RDEIC_CHECK_ARG(F > 0 && F % 8 == 0, "rdeic_geglu: F=%d must be a multiple of 8", F);
    RDEIC_CHECK_ARG(((uintptr_t)in_bf16 | (uintptr_t)out_bf16) % 16 == 0, "rdeic_geglu: unaligned");
    if (rows == 0) return 0;
    launch_k(geglu_kernel, grid_for(rows * (F / 8), kThreads), kThreads, 0, as_stream(stream), 
        (const uint4*)in_bf16, (uint4*)out_bf16, rows, F);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_upsample2x_nhwc(const void* in, void* out, int B, int H, int W, int C,
                          rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(in && out && B > 0 && H > 0 && W > 0, "rdeic_upsample2x_nhwc: bad args");
    RDEIC_CHECK_ARG(C > 0 && C % 8 == 0, "rdeic_upsample2x_nhwc: C=%d must be a multiple of 8", C);
    const int64_t total = (int64_t)B * H * W * (C / 8);
    launch_k(upsample2x_kernel, grid_for(total, kThreads), kThreads, 0, as_stream(stream),
        (const uint4*)in, (uint4*)out, B, H, W, C / 8);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_pixel_shuffle2_nhwc(const void* in, void* out, int B, int H, int W, int C,
                              rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(in && out && B > 0 && H > 0 && W > 0, "rdeic_pixel_shuffle2_nhwc: bad args");
    RDEIC_CHECK_ARG(C > 0 && C % 8 == 0, "rdeic_pixel_shuffle2_nhwc: C=%d must be a multiple of 8", C);
    const int64_t total = (int64_t)B * 4 * H * W * (C / 8);
    launch_k(pixel_shuffle2_kernel, grid_for(total, kThreads), kThreads, 0, as_stream(stream),
        (const uint4*)in, (uint4*)out, B, H, W, C / 8);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_im2col_3x3_s2(const void* in, void* out, int B, int H, int W, int C, int pad_lo,
                        rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(in && out && B > 0 && H > 0 && W > 0, "rdeic_im2col_3x3_s2: bad args");
    RDEIC_CHECK_ARG(C > 0 && C % 8 == 0, "rdeic_im2col_3x3_s2: C=%d must be a multiple of 8", C);
    RDEIC_CHECK_ARG(H % 2 == 0 && W % 2 == 0, "rdeic_im2col_3x3_s2: H, W must be even");
    RDEIC_CHECK_ARG(pad_lo == 0 || pad_lo == 1, "rdeic_im2col_3x3_s2: pad_lo must be 0 or 1");
    const int Cp = (C + 63) / 64 * 64;
    const int64_t total = (int64_t)B * (H / 2) * (W / 2) * 9 * (Cp / 8);
    launch_k(im2col_s2_kernel, grid_for(total, kThreads), kThreads, 0, as_stream(stream), 
        (const uint4*)in, (uint4*)out, B, H, W, C, Cp, pad_lo);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_softmax_rows(const void* in, int in_is_f32, void* out_bf16, int64_t rows, int n,
                       float scale, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(in && out_bf16 && rows >= 0 && n > 0, "rdeic_softmax_rows: bad args");
    RDEIC_CHECK_ARG(rows < (1ll << 31), "rdeic_softmax_rows: too many rows");
    if (rows == 0) return 0;
    const bool vec = in_is_f32 && n % 4 == 0 && n <= 8192 && ((uintptr_t)in | (uintptr_t)out_bf16) % 16 == 0;
    if (vec && n <= 1024) launch_k(softmax_rows_f32_reg_kernel<1>, (unsigned)rows, 256, 0, as_stream(stream), (const float4*)in, (uint2*)out_bf16, n / 4, scale);
    else if (vec && n <= 2048) launch_k(softmax_rows_f32_reg_kernel<2>, (unsigned)rows, 256, 0, as_stream(stream), (const float4*)in, (uint2*)out_bf16, n / 4, scale);
    else if (vec && n <= 4096) launch_k(softmax_rows_f32_reg_kernel<4>, (unsigned)rows, 256, 0, as_stream(stream), (const float4*)in, (uint2*)out_bf16, n / 4, scale);
    else if (vec) launch_k(softmax_rows_f32_reg_kernel<8>, (unsigned)rows, 256, 0, as_stream(stream), (const float4*)in, (uint2*)out_bf16, n / 4, scale);
    else if (in_is_f32) launch_k(softmax_rows_kernel<true>, (unsigned)rows, 256, 0, as_stream(stream), in, (__nv_bfloat16*)out_bf16, n, scale);
    else launch_k(softmax_rows_kernel<false>, (unsigned)rows, 256, 0, as_stream(stream), in, (__nv_bfloat16*)out_bf16, n, scale);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_transpose_bf16(const void* in, void* out, int batch, int R, int C,
                         rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(in && out && batch > 0 && R > 0 && C > 0, "rdeic_transpose_bf16: bad args");
    dim3 grid((C + 31) / 32, (R + 31) / 32, batch);
    launch_k(transpose_bf16_kernel, grid, 256, 0, as_stream(stream), (const __nv_bfloat16*)in, (__nv_bfloat16*)out, R, C);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_image_to_u8(const float* in, uint8_t* out, int64_t pixels, int ldc,
                      rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(in && out && pixels >= 0 && ldc >= 3, "rdeic_image_to_u8: bad args");
    if (pixels == 0) return 0;
    launch_k(image_to_u8_kernel, grid_for(pixels * 3, kThreads), kThreads, 0, as_stream(stream), in, out, pixels, ldc);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_blend_tiles_u8(const uint8_t* tiles, const int32_t* origin_yx, int T, int th, int tw, int overlap,
                         uint8_t* out, int H, int W, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(tiles && origin_yx && out, "rdeic_blend_tiles_u8: null pointer");
    RDEIC_CHECK_ARG(T > 0 && th > 0 && tw > 0 && H > 0 && W > 0 && overlap >= 0 && 2 * overlap <= th && 2 * overlap <= tw,
                    "rdeic_blend_tiles_u8: bad dims (the overlap band must not exceed half a tile)");
    launch_k(blend_tiles_u8_kernel, grid_for((int64_t)H * W, kThreads), kThreads, 0, as_stream(stream), tiles,
             (const int*)origin_yx, T, th, tw, overlap, out, H, W);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_split_bf16_hilo(const float* src, int64_t rows, int C, int64_t ld_src, void* dst, int64_t ld_dst,
                          int off_hi, int off_lo, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(src && dst && rows >= 0 && C > 0, "rdeic_split_bf16_hilo: bad args");
    RDEIC_CHECK_ARG(ld_src >= C && off_hi >= 0 && off_lo >= 0 && off_hi + C <= ld_dst && off_lo + C <= ld_dst &&
                        (off_hi + C <= off_lo || off_lo + C <= off_hi),
                    "rdeic_split_bf16_hilo: the two windows must be disjoint and inside a row");
    if (rows == 0) return 0;
    launch_k(split_bf16_hilo_kernel, grid_for(rows * C, kThreads), kThreads, 0, as_stream(stream), src, rows, C, ld_src,
             (__nv_bfloat16*)dst, ld_dst, off_hi, off_lo);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
