// Shared helpers for librdeic_b200: error reporting, launch checks, small device utilities.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

namespace rdeic {

// per-thread error string behind rdeic_last_error()
char* err_buf();
int set_error(const char* fmt, ...);

#define RDEIC_CHECK_ARG(cond, ...)                          \
    do {                                                    \
        if (!(cond)) return ::rdeic::set_error(__VA_ARGS__); \
    } while (0)

#define RDEIC_CUDA(call)                                                                   \
    do {                                                                                   \
        cudaError_t _e = (call);                                                           \
        if (_e != cudaSuccess)                                                             \
            return ::rdeic::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__,       \
                                      cudaGetErrorName(_e), cudaGetErrorString(_e));       \
    } while (0)

#define RDEIC_LAUNCH_CHECK() RDEIC_CUDA(cudaPeekAtLastError())

// Programmatic dependent launch (PDL): every kernel of the library is launched with
// programmaticStreamSerialization and calls pdl_wait() before its first global-memory access (blocks until the
// preceding grid has completed and flushed); the next kernel in the stream is launched as soon as our last CTA has
// exited instead of after the grid-completion round trip.  ~590 dependent launches per UNet step make launch gaps a
// first-order cost.  RDEIC_PDL=0 (or RDEIC_NO_PDL=1) falls back to plain stream order.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid for a grid-stride elementwise kernel: enough CTAs to fill the chip a few times over,
// in multiples of the SM count.
static inline int grid_for(int64_t work_items, int threads, int max_waves = 8) {
    int64_t blocks = ceil_div64(work_items, threads);
    int64_t cap = (int64_t)kNumSMs * max_waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// Division by a launch-invariant 32-bit divisor without the ~50-instruction integer divide
// (Granlund-Montgomery): q = (umulhi(n, mul) + n) >> shr, exact for n < 2^31.
struct FastDiv {
    uint32_t d, mul, shr;
    FastDiv() : d(1), mul(0), shr(0) {}
    explicit FastDiv(uint32_t div) : d(div) {
        shr = 0;
        while ((1ull << shr) < div) ++shr;
        mul = (uint32_t)(((1ull << 32) * ((1ull << shr) - div)) / div + 1);
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return (__umulhi(n, mul) + n) >> shr; }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
        q = div(n);
        r = n - q * d;
    }
};

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// pdl_trigger() is a no-op: the dependents of a grid launch when its last CTA EXITS (the implicit trigger).  With
// griddepcontrol.launch_dependents at kernel entry the next grid's CTAs take an SM each as soon as one frees up and sit
// there until this grid has finished -- SMs that the kernels of the step's other stream (the control adapter) would have used:
// UNet+control step 10.61 -> 11.12 ms.  With the implicit trigger programmatic launch still removes the
// grid-complete -> grid-launch latency of ~590 kernel boundaries: 10.61 -> 10.40 ms (A/B/C on one box, two repetitions).
// -DRDEIC_PDL_EARLY_TRIGGER restores the entry trigger for A/B builds.
// Short, many-CTA kernels (LayerNorm, GroupNorm, split-K reduce, elementwise) DO trigger on entry when built with
// -DRDEIC_PDL_SHORT_EARLY: their dependents are mostly GEMMs, whose prologue (barriers, TMEM allocation, tensor-map
// prefetch) then overlaps the few microseconds in which the short kernel's last wave drains.
__device__ __forceinline__ void pdl_trigger_short() {
#if defined(RDEIC_PDL_EARLY_TRIGGER) || defined(RDEIC_PDL_SHORT_EARLY)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_trigger() {
#ifdef RDEIC_PDL_EARLY_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// x * sigmoid(x); __fdividef = rcp.approx + mul (an IEEE divide costs ~10x more instructions, which made
// the GroupNorm+SiLU apply pass issue-bound instead of HBM-bound)
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) {
    return __uint_as_float(((uint32_t)b) << 16);
}
__device__ __forceinline__ void unpack_bf16x2(uint32_t v, float& lo, float& hi) {
    lo = __uint_as_float(v << 16);
    hi = __uint_as_float(v & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// streaming 128-bit accesses (read-once / write-once data: keep it out of L1)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// streaming 256-bit accesses (sm_100: LDG/STG.256): one full 32-byte sector per lane and instruction, so a
// warp moves 1 KB contiguously instead of two half-sector passes
__device__ __forceinline__ void st_stream_u8(void* p, const uint4& lo, const uint4& hi) {
    asm volatile("st.global.L1::no_allocate.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(lo.x), "r"(lo.y),
                 "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                 : "memory");
}
__device__ __forceinline__ void ld_stream_u8(const void* p, uint4& lo, uint4& hi) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
                 : "l"(p));
}

}  // namespace rdeic
