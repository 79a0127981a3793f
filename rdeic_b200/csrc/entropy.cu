// Entropy-model front end: checkerboard split/merge/squeeze, latent quantise/dequantise,
// CDF-index build and VQ lookup.  All integer / bit-exact, HBM-bound, NCHW fp32 like the
// reference (utils/ckbd.py, model/compression_modules.py, compressai 1.2.4 entropy_models).
//
// Layout facts used everywhere below: a tensor [B,C,H,W] is treated as R = B*C*H rows of W
// floats; the row's parity h = r % H decides which columns belong to the anchor set:
//   anchor     : (h + w) odd      (utils/ckbd.py:35-39)
//   non-anchor : (h + w) even     (utils/ckbd.py:41-45)
// Data is moved as raw 32-bit words so NaN payloads and signed zeros survive untouched.
#include "common.cuh"
#include "../../include/rdeic_b200.h"

namespace rdeic {

constexpr int kThreads = 256;

// Flat index -> (row, column-in-row, h = row % H) without integer divides on the fast path.
struct RowMap {
    FastDiv per_row, per_h;
    int fast;          // 1 when every flat index fits the 31-bit fast divider
    int cols, H;
    RowMap(int64_t total, int cols_, int H_) : per_row((uint32_t)(cols_ > 0 ? cols_ : 1)), per_h((uint32_t)(H_ > 0 ? H_ : 1)),
                                              fast(total < (1ll << 31)), cols(cols_), H(H_) {}
    __device__ __forceinline__ void map(int64_t i, int64_t& r, int& col, int& h) const {
        if (fast) {
            uint32_t rq, cq, hq, dummy;
            per_row.divmod((uint32_t)i, rq, cq);
            per_h.divmod(rq, dummy, hq);
            r = rq; col = (int)cq; h = (int)hq;
        } else {
            r = i / cols;
            col = (int)(i - r * cols);
            h = (int)(r % H);
        }
    }
};

// ------------------------------------------------------------------------------------------
// ckbd mask / split
// ------------------------------------------------------------------------------------------
template <bool kSplit>
__global__ void __launch_bounds__(kThreads)
ckbd_mask_kernel(const uint32_t* __restrict__ y, uint32_t* __restrict__ out_a,
                 uint32_t* __restrict__ out_n, int64_t rows, int H, int W, int which, RowMap rm) {
    pdl_trigger();
    pdl_wait();
    // vector path: W % 4 == 0, one uint4 per thread-iteration
    const int wv = W >> 2;
    const int64_t total = rows * wv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r; int col, h;
        rm.map(i, r, col, h);
        uint4 v = ld_stream_u4(reinterpret_cast<const uint4*>(y) + i);
        // column index of v.x is a multiple of 4 -> parity of (h + w) for lanes x,y,z,w is
        // h, h+1, h, h+1.
        const bool odd_row = h & 1;
        // anchor keeps (h+w) odd: even row -> lanes y,w ; odd row -> lanes x,z
        uint4 a, n;
        if (odd_row) {
            a = make_uint4(v.x, 0u, v.z, 0u);
            n = make_uint4(0u, v.y, 0u, v.w);
        } else {
            a = make_uint4(0u, v.y, 0u, v.w);
            n = make_uint4(v.x, 0u, v.z, 0u);
        }
        if (kSplit) {
            st_stream_u4(reinterpret_cast<uint4*>(out_a) + i, a);
            st_stream_u4(reinterpret_cast<uint4*>(out_n) + i, n);
        } else {
            st_stream_u4(reinterpret_cast<uint4*>(out_a) + i, which == 0 ? a : n);
        }
    }
}

template <bool kSplit>
__global__ void __launch_bounds__(kThreads)
ckbd_mask_scalar_kernel(const uint32_t* __restrict__ y, uint32_t* __restrict__ out_a,
                        uint32_t* __restrict__ out_n, int64_t rows, int H, int W, int which) {
    pdl_trigger();
    pdl_wait();
    const int64_t total = rows * W;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / W;
        const int w = (int)(i - r * W);
        const int h = (int)(r % H);
        const bool is_anchor = ((h + w) & 1) != 0;
        const uint32_t v = y[i];
        if (kSplit) {
            out_a[i] = is_anchor ? v : 0u;
            out_n[i] = is_anchor ? 0u : v;
        } else {
            out_a[i] = (is_anchor == (which == 0)) ? v : 0u;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
ckbd_merge_kernel(const float* __restrict__ a, const float* __restrict__ n,
                  float* __restrict__ out, int64_t numel) {
    pdl_trigger();
    pdl_wait();
    const int64_t nv = numel >> 2;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < nv; i += stride) {
        uint4 ua = ld_stream_u4(reinterpret_cast<const uint4*>(a) + i);
        uint4 un = ld_stream_u4(reinterpret_cast<const uint4*>(n) + i);
        uint4 o;
        o.x = __float_as_uint(__fadd_rn(__uint_as_float(ua.x), __uint_as_float(un.x)));
        o.y = __float_as_uint(__fadd_rn(__uint_as_float(ua.y), __uint_as_float(un.y)));
        o.z = __float_as_uint(__fadd_rn(__uint_as_float(ua.z), __uint_as_float(un.z)));
        o.w = __float_as_uint(__fadd_rn(__uint_as_float(ua.w), __uint_as_float(un.w)));
        st_stream_u4(reinterpret_cast<uint4*>(out) + i, o);
    }
    for (int64_t i = (nv << 2) + tid; i < numel; i += stride) out[i] = __fadd_rn(a[i], n[i]);
}

// ------------------------------------------------------------------------------------------
// squeeze / unsqueeze, optionally fused with index build / quantise / dequantise
// ------------------------------------------------------------------------------------------
// column offset of the kept element inside each pair (2j, 2j+1):
//   anchor: even row -> 1, odd row -> 0 ; non-anchor: even row -> 0, odd row -> 1
__device__ __forceinline__ int pair_offset(int h, int which) { return (h & 1) ^ (which == 0); }

// Index search in a sorted table, seeded so that the common case is one or two table reads:
//  mode 2 (positive lower bound, table spans <= 256 buckets): a bucket table keyed by the top bits of the
//         float (exponent + 4 mantissa bits = 16 buckets per binade) holds, per bucket, the answer for the
//         bucket's lowest value; any s in the bucket has answer >= that, so a short climb finishes it.
//         No MUFU, no float->int conversion: the kernel stays HBM-bound instead of issue-bound.
//  mode 1 (sorted, other cases): log-domain guess (the reference's table is exp(linspace(ln .11, ln 256, 64)),
//         utils/func.py:10-13), tab[L + 1] / tab[L + 2] hold ln(tab[0]) and L / (ln tab[L] - ln tab[0]).
//  mode 0: unsorted table, full count.
// Every mode compares against the table itself, so the result is exactly  first k with s <= tab[k]
// (L if none) for ANY table and any s, NaN included.
constexpr int kMaxLevels = 256;
constexpr int kLutOff = kMaxLevels + 4;          // s_tab: [0, L] table, [L+1, L+2] log seeds, then (base, nb), lut[256]
constexpr int kLutSize = 256;
constexpr int kTabFloats = kLutOff + kLutSize;

__device__ __forceinline__ int scale_index(float scale, const float* __restrict__ tab, int L,
                                           float lower_bound, int mode) {
    // compressai LowerBound = torch.max(x, bound): NaN propagates.
    const float s = (scale != scale) ? scale : fmaxf(scale, lower_bound);
    if (mode == 2) {
        const int* meta = reinterpret_cast<const int*>(tab + kMaxLevels + 2);
        int b = (int)(__float_as_uint(s) >> 19) - meta[0];       // s >= lower_bound > 0 (or NaN: huge key)
        b = min(max(b, 0), meta[1]);
        int k = reinterpret_cast<const int*>(tab + kLutOff)[b];
        while (k < L && !(s <= tab[k])) ++k;
        return k;
    }
    if (mode == 1) {
        // idx = L - #{k<L : s <= tab[k]} = first k with s <= tab[k] (L if none)
        float g = ceilf((__logf(s) - tab[L + 1]) * tab[L + 2]);
        int k = (g >= 0.f) ? ((g <= (float)L) ? (int)g : L) : 0;       // NaN -> 0
        while (k > 0 && s <= tab[k - 1]) --k;
        while (k < L && !(s <= tab[k])) ++k;
        return k;
    }
    int idx = L;
    for (int k = 0; k < L; ++k) idx -= (s <= tab[k]) ? 1 : 0;
    return idx;
}

__device__ __forceinline__ int load_table(const float* __restrict__ table, int levels, float lower_bound,
                                          float* s_tab, int* s_flag) {
    if (threadIdx.x == 0) *s_flag = 1;
    __syncthreads();
    for (int k = threadIdx.x; k < levels; k += blockDim.x) s_tab[k] = table[k];
    __syncthreads();
    for (int k = threadIdx.x; k + 1 < levels - 1; k += blockDim.x)
        if (!(s_tab[k] <= s_tab[k + 1])) *s_flag = 0;  // benign race: all writers store 0
    const int L = levels - 1;
    if (threadIdx.x == 0) {                            // seeds of the log-domain guess (see scale_index)
        const float l0 = logf(s_tab[0]), l1 = logf(s_tab[L]);
        s_tab[L + 1] = l0;
        s_tab[L + 2] = (l1 > l0) ? (float)L / (l1 - l0) : 0.f;
    }
    __syncthreads();
    if (!(*s_flag != 0 && s_tab[0] > 0.f)) return 0;
    // bucket table: keys of lower_bound .. tab[L-1]; one more bucket catches everything above (and NaN)
    const int base = (int)(__float_as_uint(lower_bound) >> 19);
    const int top = (int)(__float_as_uint(L > 0 ? s_tab[L - 1] : s_tab[0]) >> 19);
    const int nb = top - base + 1;                     // last valid bucket index
    if (!(lower_bound > 0.f) || nb < 0 || nb >= kLutSize) return 1;
    int* meta = reinterpret_cast<int*>(s_tab + kMaxLevels + 2);
    int* lut = reinterpret_cast<int*>(s_tab + kLutOff);
    if (threadIdx.x == 0) { meta[0] = base; meta[1] = nb; }
    for (int b = threadIdx.x; b <= nb; b += blockDim.x) {
        const float lo = __uint_as_float((uint32_t)(base + b) << 19);     // smallest value of the bucket
        int k = 0;
        while (k < L && !(lo <= s_tab[k])) ++k;
        lut[b] = k;
    }
    __syncthreads();
    return 2;
}


// mode 0: squeeze only (out_f)                      [ckbd.py:47-59]
// mode 1: squeeze scales+means -> means_sq, indexes [ckbd.py:99-103,108-112]
__global__ void __launch_bounds__(kThreads)
ckbd_squeeze_kernel(const uint32_t* __restrict__ y, const float* __restrict__ scales,
                    const float* __restrict__ table, int levels, float lower_bound,
                    uint32_t* __restrict__ out_f, int32_t* __restrict__ out_idx, int64_t rows,
                    int H, int W, int which, int mode) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_tab[kTabFloats];
    __shared__ int s_flag;
    int sorted = 0;
    if (mode == 1) sorted = load_table(table, levels, lower_bound, s_tab, &s_flag);
    const int Wh = W >> 1;
    const int64_t total = rows * Wh;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / Wh;
        const int j = (int)(i - r * Wh);
        const int h = (int)(r % H);
        const int64_t src = r * W + 2 * j + pair_offset(h, which);
        out_f[i] = y[src];
        if (mode == 1)
            out_idx[i] = scale_index(scales[src], s_tab, levels - 1, lower_bound, sorted);
    }
}

// vector variant: Wh % 4 == 0 -> each thread produces 4 outputs from 8 inputs
__global__ void __launch_bounds__(kThreads)
ckbd_squeeze_vec_kernel(const uint32_t* __restrict__ y, const float* __restrict__ scales,
                        const float* __restrict__ table, int levels, float lower_bound,
                        uint32_t* __restrict__ out_f, int32_t* __restrict__ out_idx,
                        int64_t rows, int H, int W, int which, int mode, RowMap rm) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_tab[kTabFloats];
    __shared__ int s_flag;
    int sorted = 0;
    if (mode == 1) sorted = load_table(table, levels, lower_bound, s_tab, &s_flag);
    const int Wq = W >> 3;  // groups of 8 inputs
    const int64_t total = rows * Wq;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r; int col, h;
        rm.map(i, r, col, h);
        const int off = pair_offset(h, which);
        uint4 lo, hi;
        ld_stream_u8(reinterpret_cast<const uint4*>(y) + 2 * i, lo, hi);
        uint4 o = off ? make_uint4(lo.y, lo.w, hi.y, hi.w) : make_uint4(lo.x, lo.z, hi.x, hi.z);
        st_stream_u4(reinterpret_cast<uint4*>(out_f) + i, o);
        if (mode == 1) {
            uint4 slo, shi;
            ld_stream_u8(reinterpret_cast<const uint4*>(scales) + 2 * i, slo, shi);
            uint4 s = off ? make_uint4(slo.y, slo.w, shi.y, shi.w)
                          : make_uint4(slo.x, slo.z, shi.x, shi.z);
            int4 id;
            const int L = levels - 1;
            id.x = scale_index(__uint_as_float(s.x), s_tab, L, lower_bound, sorted);
            id.y = scale_index(__uint_as_float(s.y), s_tab, L, lower_bound, sorted);
            id.z = scale_index(__uint_as_float(s.z), s_tab, L, lower_bound, sorted);
            id.w = scale_index(__uint_as_float(s.w), s_tab, L, lower_bound, sorted);
            st_stream_u4(reinterpret_cast<uint4*>(out_idx) + i, *reinterpret_cast<uint4*>(&id));
        }
    }
}

// mode 0: unsqueeze raw words            [ckbd.py:61-73]
// mode 1: unsqueeze(float(sym) + means)  [ckbd.py:104-105,113-114]
__global__ void __launch_bounds__(kThreads)
ckbd_unsqueeze_kernel(const uint32_t* __restrict__ sq, const int32_t* __restrict__ sym,
                      const float* __restrict__ means_sq, uint32_t* __restrict__ out,
                      int64_t rows, int H, int Wh, int which, int mode) {
    pdl_trigger();
    pdl_wait();
    const int64_t total = rows * Wh;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / Wh;
        const int h = (int)(r % H);
        const int off = pair_offset(h, which);
        uint32_t v;
        if (mode == 0) v = sq[i];
        else v = __float_as_uint(__fadd_rn(__int2float_rn(sym[i]), means_sq[i]));
        uint2 o = off ? make_uint2(0u, v) : make_uint2(v, 0u);
        reinterpret_cast<uint2*>(out)[i] = o;
    }
}

// vector variant of the raw unsqueeze: 4 inputs -> 8 outputs per thread (Wh % 4 == 0), or with kWide
// 8 inputs -> 16 outputs (Wh % 8 == 0, sq 32-byte aligned): one 256-bit load and two 256-bit stores in
// flight per lane, twice the bytes in flight of the narrow form at the same occupancy
template <bool kWide>
__global__ void __launch_bounds__(kThreads)
ckbd_unsqueeze_vec_kernel(const uint4* __restrict__ sq, uint4* __restrict__ out, int64_t rows, int H,
                          int Wh, int which, RowMap rm) {
    pdl_trigger();
    pdl_wait();
    const int wq = Wh >> (kWide ? 3 : 2);
    const int64_t total = rows * wq;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r; int col, h;
        rm.map(i, r, col, h);
        const int off = pair_offset(h, which);
        if constexpr (kWide) {
            uint4 v, w;
            ld_stream_u8(sq + 2 * i, v, w);
            const uint4 a = off ? make_uint4(0u, v.x, 0u, v.y) : make_uint4(v.x, 0u, v.y, 0u);
            const uint4 b = off ? make_uint4(0u, v.z, 0u, v.w) : make_uint4(v.z, 0u, v.w, 0u);
            const uint4 c = off ? make_uint4(0u, w.x, 0u, w.y) : make_uint4(w.x, 0u, w.y, 0u);
            const uint4 d = off ? make_uint4(0u, w.z, 0u, w.w) : make_uint4(w.z, 0u, w.w, 0u);
            st_stream_u8(out + 4 * i, a, b);
            st_stream_u8(out + 4 * i + 2, c, d);
        } else {
            const uint4 v = ld_stream_u4(sq + i);
            const uint4 lo = off ? make_uint4(0u, v.x, 0u, v.y) : make_uint4(v.x, 0u, v.y, 0u);
            const uint4 hi = off ? make_uint4(0u, v.z, 0u, v.w) : make_uint4(v.z, 0u, v.w, 0u);
            st_stream_u8(out + 2 * i, lo, hi);
        }
    }
}

// encode-side fused phase [ckbd.py:76-97]: one thread per kept element
__global__ void __launch_bounds__(kThreads)
ckbd_encode_phase_kernel(const float* __restrict__ y, const float* __restrict__ scales,
                         const float* __restrict__ means, const float* __restrict__ table,
                         int levels, float lower_bound, int32_t* __restrict__ symbols,
                         int32_t* __restrict__ indexes, float* __restrict__ y_hat,
                         int64_t rows, int H, int W, int which, RowMap rm) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_tab[kTabFloats];
    __shared__ int s_flag;
    const int sorted = load_table(table, levels, lower_bound, s_tab, &s_flag);
    const int Wh = W >> 1;
    const int64_t total = rows * Wh;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r; int j, h;
        rm.map(i, r, j, h);
        const int off = pair_offset(h, which);
        const int64_t src = r * W + 2 * j + off;
        const float m = means[src];
        const float q = rintf(__fsub_rn(y[src], m));
        const int s = __float2int_rz(q);
        symbols[i] = s;
        indexes[i] = scale_index(scales[src], s_tab, levels - 1, lower_bound, sorted);
        const float rec = __fadd_rn(__int2float_rn(s), m);
        float2 o = off ? make_float2(0.f, rec) : make_float2(rec, 0.f);
        reinterpret_cast<float2*>(y_hat)[i] = o;
    }
}

// ------------------------------------------------------------------------------------------
// quantise / dequantise / indexes on flat tensors
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
quantize_kernel(const float* __restrict__ x, const float* __restrict__ means,
                int32_t* __restrict__ sym, int64_t numel) {
    pdl_trigger();
    pdl_wait();
    const int64_t nv = numel >> 2;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < nv; i += stride) {
        uint4 ux = ld_stream_u4(reinterpret_cast<const uint4*>(x) + i);
        uint4 um = make_uint4(0, 0, 0, 0);
        if (means) um = ld_stream_u4(reinterpret_cast<const uint4*>(means) + i);
        int4 o;
        if (means) {
            o.x = __float2int_rz(rintf(__fsub_rn(__uint_as_float(ux.x), __uint_as_float(um.x))));
            o.y = __float2int_rz(rintf(__fsub_rn(__uint_as_float(ux.y), __uint_as_float(um.y))));
            o.z = __float2int_rz(rintf(__fsub_rn(__uint_as_float(ux.z), __uint_as_float(um.z))));
            o.w = __float2int_rz(rintf(__fsub_rn(__uint_as_float(ux.w), __uint_as_float(um.w))));
        } else {
            o.x = __float2int_rz(rintf(__uint_as_float(ux.x)));
            o.y = __float2int_rz(rintf(__uint_as_float(ux.y)));
            o.z = __float2int_rz(rintf(__uint_as_float(ux.z)));
            o.w = __float2int_rz(rintf(__uint_as_float(ux.w)));
        }
        st_stream_u4(reinterpret_cast<uint4*>(sym) + i, *reinterpret_cast<uint4*>(&o));
    }
    for (int64_t i = (nv << 2) + tid; i < numel; i += stride) {
        const float v = means ? __fsub_rn(x[i], means[i]) : x[i];
        sym[i] = __float2int_rz(rintf(v));
    }
}

__global__ void __launch_bounds__(kThreads)
dequantize_kernel(const int32_t* __restrict__ sym, const float* __restrict__ means,
                  float* __restrict__ out, int64_t numel) {
    pdl_trigger();
    pdl_wait();
    const int64_t nv = numel >> 2;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < nv; i += stride) {
        uint4 us = ld_stream_u4(reinterpret_cast<const uint4*>(sym) + i);
        uint4 um = ld_stream_u4(reinterpret_cast<const uint4*>(means) + i);
        uint4 o;
        o.x = __float_as_uint(__fadd_rn(__int2float_rn((int)us.x), __uint_as_float(um.x)));
        o.y = __float_as_uint(__fadd_rn(__int2float_rn((int)us.y), __uint_as_float(um.y)));
        o.z = __float_as_uint(__fadd_rn(__int2float_rn((int)us.z), __uint_as_float(um.z)));
        o.w = __float_as_uint(__fadd_rn(__int2float_rn((int)us.w), __uint_as_float(um.w)));
        st_stream_u4(reinterpret_cast<uint4*>(out) + i, o);
    }
    for (int64_t i = (nv << 2) + tid; i < numel; i += stride)
        out[i] = __fadd_rn(__int2float_rn(sym[i]), means[i]);
}

__global__ void __launch_bounds__(kThreads)
build_indexes_kernel(const float* __restrict__ scales, const float* __restrict__ table,
                     int levels, float lower_bound, int32_t* __restrict__ idx, int64_t numel) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_tab[kTabFloats];
    __shared__ int s_flag;
    const int sorted = load_table(table, levels, lower_bound, s_tab, &s_flag);
    const int L = levels - 1;
    const int64_t nv = numel >> 2;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t first = tid;
    if (((reinterpret_cast<uintptr_t>(scales) | reinterpret_cast<uintptr_t>(idx)) & 31) == 0) {
        // 32-byte aligned: one 256-bit load per lane and iteration.  A full grid of 128-bit loads keeps
        // 148 x 2048 x 16 B = 4.8 MB in flight, just under bandwidth x latency; this doubles it.
        const int64_t nw = numel >> 3;
        for (int64_t i = tid; i < nw; i += stride) {
            uint4 a, b;
            ld_stream_u8(reinterpret_cast<const uint4*>(scales) + 2 * i, a, b);
            int4 oa, ob;
            oa.x = scale_index(__uint_as_float(a.x), s_tab, L, lower_bound, sorted);
            oa.y = scale_index(__uint_as_float(a.y), s_tab, L, lower_bound, sorted);
            oa.z = scale_index(__uint_as_float(a.z), s_tab, L, lower_bound, sorted);
            oa.w = scale_index(__uint_as_float(a.w), s_tab, L, lower_bound, sorted);
            ob.x = scale_index(__uint_as_float(b.x), s_tab, L, lower_bound, sorted);
            ob.y = scale_index(__uint_as_float(b.y), s_tab, L, lower_bound, sorted);
            ob.z = scale_index(__uint_as_float(b.z), s_tab, L, lower_bound, sorted);
            ob.w = scale_index(__uint_as_float(b.w), s_tab, L, lower_bound, sorted);
            st_stream_u8(reinterpret_cast<uint4*>(idx) + 2 * i, *reinterpret_cast<uint4*>(&oa),
                         *reinterpret_cast<uint4*>(&ob));
        }
        first = 2 * nw + tid;                         // at most one 128-bit vector left
    }
    for (int64_t i = first; i < nv; i += stride) {
        uint4 us = ld_stream_u4(reinterpret_cast<const uint4*>(scales) + i);
        int4 o;
        o.x = scale_index(__uint_as_float(us.x), s_tab, L, lower_bound, sorted);
        o.y = scale_index(__uint_as_float(us.y), s_tab, L, lower_bound, sorted);
        o.z = scale_index(__uint_as_float(us.z), s_tab, L, lower_bound, sorted);
        o.w = scale_index(__uint_as_float(us.w), s_tab, L, lower_bound, sorted);
        st_stream_u4(reinterpret_cast<uint4*>(idx) + i, *reinterpret_cast<uint4*>(&o));
    }
    for (int64_t i = (nv << 2) + tid; i < numel; i += stride)
        idx[i] = scale_index(scales[i], s_tab, L, lower_bound, sorted);
}

// ------------------------------------------------------------------------------------------
// VQ: nearest codebook entry (first minimum) and lookup
// ------------------------------------------------------------------------------------------
// key = (ordered float bits of distance) << 32 | code index ; atomicMin keeps the smallest
// distance and, among equal distances, the smallest index = torch.argmin's first minimum.
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void vq_init_kernel(unsigned long long* keys, int n) {
    pdl_trigger();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ~0ull;
}

constexpr int kVqVecs = 4;      // z vectors per block (codebook row reuse)
constexpr int kVqWarps = 8;
constexpr int kVqMaxD = 512;

// grid (ceil(nvec / kVqVecs), ksplit).  One warp per codebook row: lanes stride the row in
// float4 (coalesced), z vectors live in shared memory.
__global__ void __launch_bounds__(kVqWarps * 32)
vq_search_kernel(const float* __restrict__ z, const float* __restrict__ codebook,
                 unsigned long long* __restrict__ keys, int nvec, int D, int HW, int K) {
    pdl_trigger();
    pdl_wait();
    __shared__ float s_z[kVqVecs][kVqMaxD];
    __shared__ float s_zz[kVqVecs];
    const int v0 = blockIdx.x * kVqVecs;
    // z is NCHW: element d of vector v=(b,p) lives at (b*D + d)*HW + p
    for (int i = threadIdx.x; i < kVqVecs * D; i += blockDim.x) {
        const int vv = i / D, d = i - vv * D;
        const int v = v0 + vv;
        float val = 0.f;
        if (v < nvec) {
            const int b = v / HW, p = v - b * HW;
            val = z[((int64_t)b * D + d) * HW + p];
        }
        s_z[vv][d] = val;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < kVqVecs) {  // |z|^2, sequential-in-lane then tree; only used as a common offset
        float acc = 0.f;
        for (int d = lane; d < D; d += 32) acc = fmaf(s_z[warp][d], s_z[warp][d], acc);
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_zz[warp] = acc;
    }
    __syncthreads();
    const int per = (K + gridDim.y - 1) / gridDim.y;
    const int k0 = blockIdx.y * per;
    const int k1 = min(K, k0 + per);
    unsigned long long best[kVqVecs];
#pragma unroll
    for (int vv = 0; vv < kVqVecs; ++vv) best[vv] = ~0ull;
    for (int k = k0 + warp; k < k1; k += kVqWarps) {
        const float* row = codebook + (int64_t)k * D;
        float dot[kVqVecs];
        float ee = 0.f;
#pragma unroll
        for (int vv = 0; vv < kVqVecs; ++vv) dot[vv] = 0.f;
        for (int d = lane * 4; d < D; d += 128) {
            const float4 e = *reinterpret_cast<const float4*>(row + d);
            ee = fmaf(e.x, e.x, ee); ee = fmaf(e.y, e.y, ee);
            ee = fmaf(e.z, e.z, ee); ee = fmaf(e.w, e.w, ee);
#pragma unroll
            for (int vv = 0; vv < kVqVecs; ++vv) {
                const float4 zz = *reinterpret_cast<const float4*>(&s_z[vv][d]);
                dot[vv] = fmaf(e.x, zz.x, dot[vv]); dot[vv] = fmaf(e.y, zz.y, dot[vv]);
                dot[vv] = fmaf(e.z, zz.z, dot[vv]); dot[vv] = fmaf(e.w, zz.w, dot[vv]);
            }
        }
        for (int o = 16; o; o >>= 1) {
            ee += __shfl_xor_sync(0xffffffffu, ee, o);
#pragma unroll
            for (int vv = 0; vv < kVqVecs; ++vv)
                dot[vv] += __shfl_xor_sync(0xffffffffu, dot[vv], o);
        }
#pragma unroll
        for (int vv = 0; vv < kVqVecs; ++vv) {
            // compression_modules.py:317-319: |z|^2 + |e|^2 - 2 z.e
            const float dist = __fsub_rn(__fadd_rn(s_zz[vv], ee), __fmul_rn(2.0f, dot[vv]));
            const unsigned long long key =
                ((unsigned long long)float_to_ordered(dist) << 32) | (unsigned)k;
            best[vv] = key < best[vv] ? key : best[vv];
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int vv = 0; vv < kVqVecs; ++vv)
            if (v0 + vv < nvec && best[vv] != ~0ull) atomicMin(&keys[v0 + vv], best[vv]);
    }
}

// keys -> int64 indices (in place) and zq gather (NCHW); one block per vector
__global__ void __launch_bounds__(128)
vq_finalize_kernel(unsigned long long* __restrict__ keys, const float* __restrict__ codebook,
                   float* __restrict__ zq, int nvec, int D, int HW) {
    pdl_trigger();
    pdl_wait();
    const int v = blockIdx.x;
    __shared__ long long s_idx;
    if (threadIdx.x == 0) s_idx = (long long)(keys[v] & 0xffffffffull);
    __syncthreads();
    const long long idx = s_idx;
    if (zq) {
        const int b = v / HW, p = v - b * HW;
        for (int d = threadIdx.x; d < D; d += blockDim.x)
            zq[((int64_t)b * D + d) * HW + p] = codebook[idx * D + d];
    }
    __syncthreads();
    if (threadIdx.x == 0) reinterpret_cast<long long*>(keys)[v] = idx;
}

__global__ void __launch_bounds__(128)
vq_lookup_kernel(const long long* __restrict__ indices, const float* __restrict__ codebook,
                 float* __restrict__ out, int nvec, int D, int HW, int K, int* __restrict__ bad) {
    pdl_trigger();
    pdl_wait();
    const int v = blockIdx.x;
    long long idx = indices[v];
    if (idx < 0 || idx >= K) {  // nn.Embedding raises; we flag and clamp
        if (threadIdx.x == 0 && bad) atomicExch(bad, 1);
        idx = idx < 0 ? 0 : K - 1;
    }
    const int b = v / HW, p = v - b * HW;
    for (int d = threadIdx.x; d < D; d += blockDim.x)
        out[((int64_t)b * D + d) * HW + p] = codebook[idx * D + d];
}

}  // namespace rdeic

using namespace rdeic;

extern "C" {

int rdeic_ckbd_mask(const float* y, float* out, int B, int C, int H, int W, int which,
                    rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(y && out, "rdeic_ckbd_mask: null pointer");
    RDEIC_CHECK_ARG(B >= 0 && C >= 0 && H >= 0 && W >= 0, "rdeic_ckbd_mask: negative dim");
    RDEIC_CHECK_ARG(which == 0 || which == 1, "rdeic_ckbd_mask: which must be 0 or 1");
    const int64_t rows = (int64_t)B * C * H;
    if (rows == 0 || W == 0) return 0;
    const bool vec = (W % 4 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)out % 16 == 0);
    if (vec)
        launch_k(ckbd_mask_kernel<false>, grid_for(rows * (W / 4), kThreads), kThreads, 0, as_stream(stream), 
            (const uint32_t*)y, (uint32_t*)out, nullptr, rows, H, W, which, RowMap(rows * (W / 4), W / 4, H));
    else
        launch_k(ckbd_mask_scalar_kernel<false>, grid_for(rows * W, kThreads), kThreads, 0, as_stream(stream), 
            (const uint32_t*)y, (uint32_t*)out, nullptr, rows, H, W, which);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_ckbd_split(const float* y, float* anchor, float* nonanchor, int B, int C, int H,
                     int W, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(y && anchor && nonanchor, "rdeic_ckbd_split: null pointer");
    RDEIC_CHECK_ARG(B >= 0 && C >= 0 && H >= 0 && W >= 0, "rdeic_ckbd_split: negative dim");
    const int64_t rows = (int64_t)B * C * H;
    if (rows == 0 || W == 0) return 0;
    const bool vec = (W % 4 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)anchor % 16 == 0) &&
                     ((uintptr_t)nonanchor % 16 == 0);
    if (vec)
        launch_k(ckbd_mask_kernel<true>, grid_for(rows * (W / 4), kThreads), kThreads, 0, as_stream(stream), 
            (const uint32_t*)y, (uint32_t*)anchor, (uint32_t*)nonanchor, rows, H, W, 0, RowMap(rows * (W / 4), W / 4, H));
    else
        launch_k(ckbd_mask_scalar_kernel<true>, grid_for(rows * W, kThreads), kThreads, 0, as_stream(stream), 
            (const uint32_t*)y, (uint32_t*)anchor, (uint32_t*)nonanchor, rows, H, W, 0);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_ckbd_merge(const float* anchor, const float* nonanchor, float* out, int64_t numel,
                     rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(anchor && nonanchor && out, "rdeic_ckbd_merge: null pointer");
    RDEIC_CHECK_ARG(numel >= 0, "rdeic_ckbd_merge: negative numel");
    if (numel == 0) return 0;
    RDEIC_CHECK_ARG(((uintptr_t)anchor | (uintptr_t)nonanchor | (uintptr_t)out) % 16 == 0,
                    "rdeic_ckbd_merge: pointers must be 16-byte aligned");
    launch_k(ckbd_merge_kernel, grid_for((numel + 3) / 4, kThreads), kThreads, 0, as_stream(stream), 
        anchor, nonanchor, out, numel);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

static int launch_squeeze(const float* y, const float* scales, const float* table, int levels,
                          float lower_bound, float* out_f, int32_t* out_idx, int B, int C, int H,
                          int W, int which, int mode, rdeic_stream_t stream, const char* who) {
    RDEIC_CHECK_ARG(B >= 0 && C >= 0 && H >= 0 && W >= 0, "%s: negative dim", who);
    RDEIC_CHECK_ARG(which == 0 || which == 1, "%s: which must be 0 or 1", who);
    // utils/ckbd.py:50-51 raises a shape-mismatch error for odd W; so do we.
    RDEIC_CHECK_ARG(W % 2 == 0, "%s: W=%d must be even (reference slice assignment fails)", who, W);
    if (mode == 1) RDEIC_CHECK_ARG(levels >= 2 && levels <= kMaxLevels, "%s: levels=%d out of range", who, levels);
    const int64_t rows = (int64_t)B * C * H;
    if (rows == 0 || W == 0) return 0;
    bool vec = (W % 8 == 0) && ((uintptr_t)y % 32 == 0) && ((uintptr_t)out_f % 16 == 0);     // 256-bit loads
    if (mode == 1) vec = vec && ((uintptr_t)scales % 32 == 0) && ((uintptr_t)out_idx % 16 == 0);
    if (vec)
        launch_k(ckbd_squeeze_vec_kernel, grid_for(rows * (W / 8), kThreads), kThreads, 0, as_stream(stream), 
            (const uint32_t*)y, scales, table, levels, lower_bound, (uint32_t*)out_f, out_idx,
            rows, H, W, which, mode, RowMap(rows * (W / 8), W / 8, H));
    else
        launch_k(ckbd_squeeze_kernel, grid_for(rows * (W / 2), kThreads), kThreads, 0, as_stream(stream), 
            (const uint32_t*)y, scales, table, levels, lower_bound, (uint32_t*)out_f, out_idx,
            rows, H, W, which, mode);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_ckbd_squeeze(const float* y, float* out, int B, int C, int H, int W, int which,
                       rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(y && out, "rdeic_ckbd_squeeze: null pointer");
    return launch_squeeze(y, nullptr, nullptr, 0, 0.f, out, nullptr, B, C, H, W, which, 0, stream,
                          "rdeic_ckbd_squeeze");
}

int rdeic_ckbd_squeeze_indexes(const float* scales, const float* means, const float* table,
                               int levels, float lower_bound, float* means_sq,
                               int32_t* indexes, int B, int C, int H, int W, int which,
                               rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(scales && means && table && means_sq && indexes,
                    "rdeic_ckbd_squeeze_indexes: null pointer");
    return launch_squeeze(means, scales, table, levels, lower_bound, means_sq, indexes, B, C, H, W,
                          which, 1, stream, "rdeic_ckbd_squeeze_indexes");
}

int rdeic_ckbd_unsqueeze(const float* sq, float* out, int B, int C, int H, int Wh, int which,
                         rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(sq && out, "rdeic_ckbd_unsqueeze: null pointer");
    RDEIC_CHECK_ARG(B >= 0 && C >= 0 && H >= 0 && Wh >= 0, "rdeic_ckbd_unsqueeze: negative dim");
    RDEIC_CHECK_ARG(which == 0 || which == 1, "rdeic_ckbd_unsqueeze: which must be 0 or 1");
    const int64_t rows = (int64_t)B * C * H;
    if (rows == 0 || Wh == 0) return 0;
    RDEIC_CHECK_ARG((uintptr_t)out % 8 == 0, "rdeic_ckbd_unsqueeze: out must be 8-byte aligned");
    if (Wh % 8 == 0 && (uintptr_t)sq % 32 == 0 && (uintptr_t)out % 32 == 0)     // 256-bit loads and stores
        launch_k(ckbd_unsqueeze_vec_kernel<true>, grid_for(rows * (Wh / 8), kThreads), kThreads, 0, as_stream(stream), 
            (const uint4*)sq, (uint4*)out, rows, H, Wh, which, RowMap(rows * (Wh / 8), Wh / 8, H));
    else if (Wh % 4 == 0 && (uintptr_t)sq % 16 == 0 && (uintptr_t)out % 32 == 0)     // 256-bit stores
        launch_k(ckbd_unsqueeze_vec_kernel<false>, grid_for(rows * (Wh / 4), kThreads), kThreads, 0, as_stream(stream), 
            (const uint4*)sq, (uint4*)out, rows, H, Wh, which, RowMap(rows * (Wh / 4), Wh / 4, H));
    else
        launch_k(ckbd_unsqueeze_kernel, grid_for(rows * Wh, kThreads), kThreads, 0, as_stream(stream), 
            (const uint32_t*)sq, nullptr, nullptr, (uint32_t*)out, rows, H, Wh, which, 0);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_ckbd_decode_phase(const int32_t* symbols, const float* means_sq, float* y_hat, int B,
                            int C, int H, int Wh, int which, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(symbols && means_sq && y_hat, "rdeic_ckbd_decode_phase: null pointer");
    RDEIC_CHECK_ARG(B >= 0 && C >= 0 && H >= 0 && Wh >= 0, "rdeic_ckbd_decode_phase: negative dim");
    RDEIC_CHECK_ARG(which == 0 || which == 1, "rdeic_ckbd_decode_phase: which must be 0 or 1");
    const int64_t rows = (int64_t)B * C * H;
    if (rows == 0 || Wh == 0) return 0;
    RDEIC_CHECK_ARG((uintptr_t)y_hat % 8 == 0, "rdeic_ckbd_decode_phase: y_hat must be 8-byte aligned");
    launch_k(ckbd_unsqueeze_kernel, grid_for(rows * Wh, kThreads), kThreads, 0, as_stream(stream), 
        nullptr, symbols, means_sq, (uint32_t*)y_hat, rows, H, Wh, which, 1);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_ckbd_encode_phase(const float* y, const float* scales, const float* means,
                            const float* table, int levels, float lower_bound,
                            int32_t* symbols, int32_t* indexes, float* y_hat, int B, int C,
                            int H, int W, int which, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(y && scales && means && table && symbols && indexes && y_hat,
                    "rdeic_ckbd_encode_phase: null pointer");
    RDEIC_CHECK_ARG(B >= 0 && C >= 0 && H >= 0 && W >= 0, "rdeic_ckbd_encode_phase: negative dim");
    RDEIC_CHECK_ARG(which == 0 || which == 1, "rdeic_ckbd_encode_phase: which must be 0 or 1");
    RDEIC_CHECK_ARG(W % 2 == 0, "rdeic_ckbd_encode_phase: W=%d must be even", W);
    RDEIC_CHECK_ARG(levels >= 2 && levels <= kMaxLevels, "rdeic_ckbd_encode_phase: levels=%d out of range", levels);
    const int64_t rows = (int64_t)B * C * H;
    if (rows == 0 || W == 0) return 0;
    RDEIC_CHECK_ARG((uintptr_t)y_hat % 8 == 0, "rdeic_ckbd_encode_phase: y_hat must be 8-byte aligned");
    launch_k(ckbd_encode_phase_kernel, grid_for(rows * (W / 2), kThreads), kThreads, 0, as_stream(stream), 
        y, scales, means, table, levels, lower_bound, symbols, indexes, y_hat, rows, H, W, which,
        RowMap(rows * (W / 2), W / 2, H));
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_quantize_symbols(const float* x, const float* means, int32_t* symbols, int64_t numel,
                           rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(x && symbols, "rdeic_quantize_symbols: null pointer");
    RDEIC_CHECK_ARG(numel >= 0, "rdeic_quantize_symbols: negative numel");
    if (numel == 0) return 0;
    RDEIC_CHECK_ARG(((uintptr_t)x | (uintptr_t)means | (uintptr_t)symbols) % 16 == 0,
                    "rdeic_quantize_symbols: pointers must be 16-byte aligned");
    launch_k(quantize_kernel, grid_for((numel + 3) / 4, kThreads), kThreads, 0, as_stream(stream), 
        x, means, symbols, numel);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_dequantize(const int32_t* symbols, const float* means, float* out, int64_t numel,
                     rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(symbols && means && out, "rdeic_dequantize: null pointer");
    RDEIC_CHECK_ARG(numel >= 0, "rdeic_dequantize: negative numel");
    if (numel == 0) return 0;
    RDEIC_CHECK_ARG(((uintptr_t)symbols | (uintptr_t)means | (uintptr_t)out) % 16 == 0,
                    "rdeic_dequantize: pointers must be 16-byte aligned");
    launch_k(dequantize_kernel, grid_for((numel + 3) / 4, kThreads), kThreads, 0, as_stream(stream), 
        symbols, means, out, numel);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_build_indexes(const float* scales, const float* table, int levels, float lower_bound,
                        int32_t* indexes, int64_t numel, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(scales && table && indexes, "rdeic_build_indexes: null pointer");
    RDEIC_CHECK_ARG(numel >= 0, "rdeic_build_indexes: negative numel");
    RDEIC_CHECK_ARG(levels >= 2 && levels <= kMaxLevels, "rdeic_build_indexes: levels=%d out of range", levels);
    if (numel == 0) return 0;
    RDEIC_CHECK_ARG(((uintptr_t)scales | (uintptr_t)indexes) % 16 == 0,
                    "rdeic_build_indexes: pointers must be 16-byte aligned");
    launch_k(build_indexes_kernel, grid_for((numel + 3) / 4, kThreads), kThreads, 0, as_stream(stream), 
        scales, table, levels, lower_bound, indexes, numel);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_vq_quant(const float* z, const float* codebook, int64_t* indices, float* zq, int B,
                   int D, int HW, int K, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(z && codebook && indices, "rdeic_vq_quant: null pointer");
    RDEIC_CHECK_ARG(B >= 0 && HW >= 0, "rdeic_vq_quant: negative dim");
    RDEIC_CHECK_ARG(D > 0 && D <= kVqMaxD && D % 4 == 0, "rdeic_vq_quant: D=%d must be a multiple of 4 in (0,%d]", D, kVqMaxD);
    RDEIC_CHECK_ARG(K > 0, "rdeic_vq_quant: empty codebook");
    RDEIC_CHECK_ARG((uintptr_t)codebook % 16 == 0, "rdeic_vq_quant: codebook must be 16-byte aligned");
    const int nvec = B * HW;
    if (nvec == 0) return 0;
    cudaStream_t s = as_stream(stream);
    launch_k(vq_init_kernel, (nvec + 255) / 256, 256, 0, s, (unsigned long long*)indices, nvec);
    RDEIC_LAUNCH_CHECK();
    const int vtiles = (nvec + kVqVecs - 1) / kVqVecs;
    int ksplit = (2 * kNumSMs + vtiles - 1) / vtiles;
    ksplit = ksplit < 1 ? 1 : (ksplit > 64 ? 64 : ksplit);
    if (ksplit > K) ksplit = K;
    launch_k(vq_search_kernel, dim3(vtiles, ksplit), kVqWarps * 32, 0, s, 
        z, codebook, (unsigned long long*)indices, nvec, D, HW, K);
    RDEIC_LAUNCH_CHECK();
    launch_k(vq_finalize_kernel, nvec, 128, 0, s, (unsigned long long*)indices, codebook, zq, nvec, D, HW);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

int rdeic_vq_lookup(const int64_t* indices, const float* codebook, float* out, int B, int D,
                    int HW, int K, rdeic_stream_t stream) {
    RDEIC_CHECK_ARG(indices && codebook && out, "rdeic_vq_lookup: null pointer");
    RDEIC_CHECK_ARG(B >= 0 && HW >= 0 && D > 0 && K > 0, "rdeic_vq_lookup: bad dims");
    const int nvec = B * HW;
    if (nvec == 0) return 0;
    launch_k(vq_lookup_kernel, nvec, 128, 0, as_stream(stream), (const long long*)indices, codebook, out,
                                                         nvec, D, HW, K, nullptr);
    RDEIC_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
