// Shared device/host helpers for libdlv3p (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <initializer_list>
#include <type_traits>
#include <utility>

#include "../../include/dlv3p.h"

namespace dlv3p {

// ---- error plumbing (thread-local message, C-ABI returns a negative enum) -------------------
void set_error(const char* fmt, ...);
int  check_launch(const char* what);      // cudaGetLastError -> DLV3P_ERR_CUDA

#define DLV3P_REQUIRE(cond, code, ...)                                   \
    do {                                                                 \
        if (!(cond)) {                                                   \
            ::dlv3p::set_error(__VA_ARGS__);                             \
            return (code);                                               \
        }                                                                \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline int  cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------
// The hot kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization (also inside a captured CUDA
// graph, where it becomes a programmatic edge): a kernel releases its dependents as soon as all of its CTAs are
// resident (pdl_launch_dependents at the top), so the next kernel's CTAs are scheduled onto SMs as they drain and run
// their prologue (barrier init, TMEM allocation, tensor-map prefetch) under the previous kernel's tail; pdl_wait()
// then blocks until the previous kernel has completed and its writes are visible.  Every kernel launched this way
// calls pdl_wait() before its first global-memory access.  DLV3P_PDL=0 in the environment disables the attribute.
bool pdl_enabled();
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);      // errors surface through check_launch()
}

// ---- element access: fp32 math, storage in T ------------------------------------------------
template <typename T> struct Vec8;     // 8 elements of T (16 B for bf16, 32 B for fp32)

template <> struct Vec8<__nv_bfloat16> {
    uint4 raw;
    __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = __ldg(reinterpret_cast<const uint4*>(p)); }
    __device__ __forceinline__ void load_stream(const __nv_bfloat16* p) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(p));
    }
    // Coherent load (L2, no non-coherent / L1 copy) for operands that may ALIAS the kernel's output: the engine
    // accumulates gradients in place (addend == out).  A non-coherent load (__ldg / ld.global.nc) of memory the same
    // kernel writes is undefined by PTX and leaves stale lines in the SM's read-only cache that a dependent kernel's
    // non-coherent loads can still hit (seen as a few stale 8-element vectors in a BN backward, 1 run in 6).
    __device__ __forceinline__ void load_rw(const __nv_bfloat16* p) {
        asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w) : "l"(p) : "memory");
    }
    __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
    __device__ __forceinline__ void zero() { raw = make_uint4(0, 0, 0, 0); }
    __device__ __forceinline__ void to_float(float (&f)[8]) const {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 v = __bfloat1622float2(h[i]);
            f[2 * i] = v.x; f[2 * i + 1] = v.y;
        }
    }
    __device__ __forceinline__ void from_float(const float (&f)[8]) {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    }
};

template <> struct Vec8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = __ldg(reinterpret_cast<const float4*>(p));
        b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    }
    __device__ __forceinline__ void load_stream(const float* p) { load(p); }
    __device__ __forceinline__ void load_rw(const float* p) {
        asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p) : "memory");
        asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4) : "memory");
    }
    __device__ __forceinline__ void store(float* p) const {
        reinterpret_cast<float4*>(p)[0] = a;
        reinterpret_cast<float4*>(p)[1] = b;
    }
    __device__ __forceinline__ void zero() { a = make_float4(0, 0, 0, 0); b = a; }
    __device__ __forceinline__ void to_float(float (&f)[8]) const {
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    __device__ __forceinline__ void from_float(const float (&f)[8]) {
        a = make_float4(f[0], f[1], f[2], f[3]); b = make_float4(f[4], f[5], f[6], f[7]);
    }
};

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
// scalar coherent load (see Vec8::load_rw)
template <typename T> __device__ __forceinline__ float ld_rw_f(const T* p) { return to_f<T>(__ldcg(p)); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// activation codes shared by every kernel (DLV3P_ACT_*)
__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == DLV3P_ACT_RELU) return fmaxf(v, 0.f);
    if (act == DLV3P_ACT_RELU6) return fminf(fmaxf(v, 0.f), 6.f);
    return v;
}
// derivative mask of the activation evaluated at pre-activation value v
__device__ __forceinline__ float act_mask(float v, int act) {
    if (act == DLV3P_ACT_RELU) return v > 0.f ? 1.f : 0.f;
    if (act == DLV3P_ACT_RELU6) return (v > 0.f && v < 6.f) ? 1.f : 0.f;
    return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// dispatch helper on the storage dtype enum
#define DLV3P_DISPATCH_DTYPE(dtype, T, ...)                                          \
    do {                                                                             \
        if ((dtype) == DLV3P_F32) { using T = float; __VA_ARGS__; }                  \
        else if ((dtype) == DLV3P_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }    \
        else { ::dlv3p::set_error("unsupported dtype %d", (int)(dtype)); return DLV3P_ERR_DTYPE; } \
    } while (0)

}  // namespace dlv3p

namespace dlv3p {
// Generic V-wide (8 or 1) element pack so every elementwise kernel has a vector and a scalar-tail flavour.
template <typename T, int V> struct Pack;
template <typename T> struct Pack<T, 8> {
    Vec8<T> v;
    __device__ __forceinline__ void load(const T* p) { v.load(p); }
    __device__ __forceinline__ void load_rw(const T* p) { v.load_rw(p); }
    __device__ __forceinline__ void store(T* p) const { v.store(p); }
    __device__ __forceinline__ void to_float(float (&f)[8]) const { v.to_float(f); }
    __device__ __forceinline__ void from_float(const float (&f)[8]) { v.from_float(f); }
};
template <typename T> struct Pack<T, 1> {
    T v;
    __device__ __forceinline__ void load(const T* p) { v = *p; }
    __device__ __forceinline__ void load_rw(const T* p) { v = __ldcg(p); }
    __device__ __forceinline__ void store(T* p) const { *p = v; }
    __device__ __forceinline__ void to_float(float (&f)[1]) const { f[0] = to_f<T>(v); }
    __device__ __forceinline__ void from_float(const float (&f)[1]) { v = from_f<T>(f[0]); }
};
}  // namespace dlv3p

namespace dlv3p {
// choose the channel-pack width of the column-reduction kernels
template <typename F>
static inline void pick_cvb(int CV, F&& f) {
    if (CV >= 32) f(std::integral_constant<int, 32>{});
    else if (CV >= 16) f(std::integral_constant<int, 16>{});
    else if (CV >= 8) f(std::integral_constant<int, 8>{});
    else f(std::integral_constant<int, 4>{});
}

// BatchNormalization finalize operands folded into a reading depthwise kernel (dlv3p_dwconv3x3_bn_fwd)
struct DwBnFold {
    const float* sums; const float* gamma; const float* beta; float* moving_mean; float* moving_var;
    float* scale; float* shift; float* mean; float* invstd;
    double count; float eps, momentum; int updates;
};
}  // namespace dlv3p
