// K2 — pointwise / projection convolutions as bf16 tcgen05 tensor-core GEMMs (sm_100a).
//
// Replaces TF Conv2D(1x1) (cuDNN/cuBLAS in the reference's TF 2.4 runtime) for the pointwise half of every
// SeparableConv2D, the ASPP / decoder projections (ss.py:814-818,833-838,843-847,865-869,931-935) and, after
// im2col, the dense 3x3 convs (ss.py:893-897).
//
// Structure (one 128 x BLOCK_N output tile per CTA, 192 threads):
//   warp 0, one lane : TMA producer  — cp.async.bulk.tensor 2D tiles (128B swizzle) into a 4-stage smem ring,
//                                      completion on mbarriers (expect_tx)
//   warp 1, one lane : MMA issuer    — tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N, K=16 per instruction,
//                                      fp32 accumulator in TMEM (BLOCK_N columns); tcgen05.commit releases smem slots
//   warps 2..5       : epilogue      — tcgen05.ld 32x32b (one accumulator row per thread), fused per-column
//                                      scale/shift (folded BatchNorm) + ReLU/ReLU6 + residual addend, optional
//                                      per-column sum / sum-of-squares (training-mode BatchNorm statistics), store
// Forward / input-gradient use K-major operands (A[M,K], B[N,K], K contiguous).  The filter gradient
// dW = X^T dY contracts over the pixel axis, which is NOT contiguous in NHWC: both operands are fed MN-major
// (64-element x 64-row TMA boxes, same 128B swizzle) so no transposed copy of the activations is ever written.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>

#include "common.cuh"

namespace dlv3p {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // 64 bf16 = 128 B = one swizzle atom row
constexpr int kMaxStages = 4;
constexpr int kThreads = 192;

struct GemmParams {
    int M, N, K;                       // logical GEMM extents (for WGRAD: rows=K(cin), cols=N(cout), reduction=M)
    void* C; long long ldc; int c_dtype;
    const float* col_scale; const float* col_shift; int act;
    const void* addend; long long ld_add;
    float* col_stats;
    int kb_per_split;                  // WGRAD: reduction blocks (of 64 rows) handled by one CTA
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor (sm_100): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48)
// | layout [61,64) (2 = SWIZZLE_128B).  Field layout as in CUTLASS cute/arch/mma_sm100_desc.hpp.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// column sums over the 32 lanes (= 32 accumulator rows) of a warp for 32 columns held as v[0..31] per lane:
// recursive halving, 31 shuffles; on return lane L holds the sum of column L in v[0].
__device__ __forceinline__ float warp_col_sums(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float send = up ? v[i] : v[i + o];
            const float keep = up ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

// kStages is matched to the depth of the K loop: the HBM-bound GEMMs of the entry flow (K = 64..256, i.e. 1-4
// k-blocks) take 1-2 stages so that 2-3 CTAs share an SM and one CTA's epilogue overlaps another's loads.
template <int BLOCK_N, bool WGRAD, int kStages>
__global__ void __launch_bounds__(kThreads, (kStages <= 2 && BLOCK_N <= 128) ? 3 : (kStages <= 2 ? 2 : 1))
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    constexpr int A_BYTES = kBlockM * kBlockK * 2;
    constexpr int B_BYTES = BLOCK_N * kBlockK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * STAGE_BYTES);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_bar = smem_u32(bars);                    // kStages barriers
    const uint32_t empty_bar = smem_u32(bars + kStages);         // kStages barriers
    const uint32_t tmem_full_bar = smem_u32(bars + 2 * kStages);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)BLOCK_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

    // tile coordinates.  Forward: rows = pixels (blockIdx.y), cols = out channels (blockIdx.x), reduce over K.
    // WGRAD: rows = in channels (blockIdx.y), cols = out channels (blockIdx.x), reduce over pixels (split blockIdx.z).
    const int row0 = blockIdx.y * kBlockM;
    const int col0 = blockIdx.x * BLOCK_N;
    int kb_begin, kb_end;
    if (WGRAD) {
        const int total_kb = (p.M + kBlockK - 1) / kBlockK;
        kb_begin = blockIdx.z * p.kb_per_split;
        kb_end = min(kb_begin + p.kb_per_split, total_kb);
    } else {
        kb_begin = 0;
        kb_end = (p.K + kBlockK - 1) / kBlockK;
    }
    const int num_kb = kb_end - kb_begin;     // host guarantees >= 1

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % kStages;
                const uint32_t ph = (uint32_t)(i / kStages) & 1u;
                mbar_wait(empty_bar + 8 * s, ph ^ 1u);
                const uint32_t a_dst = smem_base + s * STAGE_BYTES;
                const uint32_t b_dst = a_dst + A_BYTES;
                const uint32_t fb = full_bar + 8 * s;
                mbar_expect_tx(fb, STAGE_BYTES);
                const int kk = (kb_begin + i) * kBlockK;
                if (WGRAD) {
                    // MN-major operands: boxes of 64 channels (inner, 128 B) x 64 pixels
#pragma unroll
                    for (int h = 0; h < kBlockM / 64; ++h) tma_load_2d(a_dst + h * 8192, &tmA, fb, row0 + 64 * h, kk);
#pragma unroll
                    for (int h = 0; h < BLOCK_N / 64; ++h) tma_load_2d(b_dst + h * 8192, &tmB, fb, col0 + 64 * h, kk);
                } else {
                    tma_load_2d(a_dst, &tmA, fb, kk, row0);
                    tma_load_2d(b_dst, &tmB, fb, kk, col0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ |
                                       ((WGRAD ? 1u : 0u) << 15) | ((WGRAD ? 1u : 0u) << 16) |
                                       ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
            for (int i = 0; i < num_kb; ++i) {
                const int s = i % kStages;
                const uint32_t ph = (uint32_t)(i / kStages) & 1u;
                mbar_wait(full_bar + 8 * s, ph);
                tc_fence_after();
                const uint32_t a_src = smem_base + s * STAGE_BYTES;
                const uint32_t b_src = a_src + A_BYTES;
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                    uint64_t ad, bd;
                    if (WGRAD) {
                        // MN-major SW128: LBO = next 64-element MN chunk (one 8 KB box), SBO = next 8 K-rows (1 KB);
                        // a K=16 step is 16 rows of 128 B
                        ad = umma_desc(a_src + k * 2048, 8192, 1024);
                        bd = umma_desc(b_src + k * 2048, 8192, 1024);
                    } else {
                        // K-major SW128: SBO = next 8 rows (1 KB); a K=16 step is 32 B inside the swizzle atom
                        ad = umma_desc(a_src + k * 32, 16, 1024);
                        bd = umma_desc(b_src + k * 32, 16, 1024);
                    }
                    tc_mma_bf16(tmem_base, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
                }
                tc_commit(empty_bar + 8 * s);          // smem slot reusable once these MMAs retire
            }
            tc_commit(tmem_full_bar);                  // accumulator complete
        }
    } else {
        // ===== epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
        const int q = warp & 3;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int r = row0 + q * 32 + lane;            // output row owned by this thread
        const int row_limit = WGRAD ? p.K : p.M;
        const bool row_ok = r < row_limit;
#pragma unroll 1
        for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
            const int n_base = col0 + ch * 32;
            if (n_base >= p.N) break;                  // warp-uniform
            uint32_t raw[32];
            tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), raw);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);

            if (WGRAD) {
                if (row_ok) {
                    float* dst = reinterpret_cast<float*>(p.C) + (long long)r * p.ldc + n_base;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n_base + j < p.N) atomicAdd(dst + j, v[j]);
                }
                continue;
            }

            if (p.col_stats != nullptr) {
                // rows beyond M were zero-filled by TMA, so they add nothing
                float sbuf[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) sbuf[j] = v[j];
                const float c1 = warp_col_sums(sbuf, lane);
#pragma unroll
                for (int j = 0; j < 32; ++j) sbuf[j] = v[j] * v[j];
                const float c2 = warp_col_sums(sbuf, lane);
                if (n_base + lane < p.N) {
                    atomicAdd(p.col_stats + n_base + lane, c1);
                    atomicAdd(p.col_stats + p.N + n_base + lane, c2);
                }
            }
            if (!row_ok) continue;
            if (p.col_scale != nullptr) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = min(n_base + j, p.N - 1);
                    v[j] = fmaf(v[j], __ldg(p.col_scale + n), __ldg(p.col_shift + n));
                }
            }
            if (p.act != DLV3P_ACT_NONE) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
            }
            const bool full = (n_base + 32 <= p.N);
            if (p.c_dtype == DLV3P_BF16) {
                __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)r * p.ldc + n_base;
                const __nv_bfloat16* add =
                    p.addend ? reinterpret_cast<const __nv_bfloat16*>(p.addend) + (long long)r * p.ld_add + n_base : nullptr;
                const bool vec = full && ((p.ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                                 (add == nullptr || (((p.ld_add & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.addend) & 15) == 0)));
                if (vec) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float f[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) f[j] = v[g * 8 + j];
                        if (add != nullptr) {
                            Vec8<__nv_bfloat16> a; a.load(add + g * 8);
                            float af[8]; a.to_float(af);
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] += af[j];
                        }
                        Vec8<__nv_bfloat16> o; o.from_float(f);
                        o.store(dst + g * 8);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (n_base + j < p.N) {
                            float f = v[j];
                            if (add != nullptr) f += __bfloat162float(add[j]);
                            dst[j] = __float2bfloat16_rn(f);
                        }
                    }
                }
            } else {
                float* dst = reinterpret_cast<float*>(p.C) + (long long)r * p.ldc + n_base;
                const float* add = p.addend ? reinterpret_cast<const float*>(p.addend) + (long long)r * p.ld_add + n_base : nullptr;
                const bool vec = full && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (add != nullptr && n_base + j < p.N) v[j] += add[j];
                if (vec) {
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        reinterpret_cast<float4*>(dst)[g] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n_base + j < p.N) dst[j] = v[j];
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BLOCK_N)
                     : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    });
    return fn;
}

// 2D bf16 tensor map: inner extent d0 (contiguous), outer extent d1 with row pitch ld elements
static int make_tmap(CUtensorMap* map, const void* base, long long d0, long long d1, long long ld, int box0, int box1) {
    auto fn = get_encode_fn();
    DLV3P_REQUIRE(fn != nullptr, DLV3P_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DLV3P_REQUIRE(rc == CUDA_SUCCESS, DLV3P_ERR_CUDA,
                  "cuTensorMapEncodeTiled failed (%d) dims=(%lld,%lld) ld=%lld box=(%d,%d)", (int)rc, d0, d1, ld, box0,
                  box1);
    return 0;
}

template <int BLOCK_N, bool WGRAD, int kStages>
static int launch_gemm_s(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, dim3 grid, cudaStream_t st) {
    constexpr int smem = kStages * (kBlockM * kBlockK * 2 + BLOCK_N * kBlockK * 2) + 1024 /*align*/ + 256 /*barriers*/;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N, WGRAD, kStages>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
        configured = true;
    }
    gemm_tc_kernel<BLOCK_N, WGRAD, kStages><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
    return check_launch(WGRAD ? "gemm_wgrad_bf16" : "gemm_bf16");
}

template <int BLOCK_N, bool WGRAD>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, dim3 grid, cudaStream_t st,
                       int k_blocks) {
    if (k_blocks <= 1) return launch_gemm_s<BLOCK_N, WGRAD, 1>(tmA, tmB, p, grid, st);
    if (k_blocks <= 3) return launch_gemm_s<BLOCK_N, WGRAD, 2>(tmA, tmB, p, grid, st);
    return launch_gemm_s<BLOCK_N, WGRAD, kMaxStages>(tmA, tmB, p, grid, st);
}

}  // namespace dlv3p

using namespace dlv3p;

extern "C" int dlv3p_gemm_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int M,
                               int N, int K, int c_dtype, const float* col_scale, const float* col_shift, int act,
                               const void* addend, int64_t ld_addend, float* col_stats, void* stream) {
    DLV3P_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, DLV3P_ERR_SHAPE, "gemm_bf16: bad arguments M=%d N=%d K=%d", M, N, K);
    DLV3P_REQUIRE(lda >= K && ldb >= K && ldc >= N, DLV3P_ERR_SHAPE, "gemm_bf16: leading dimension smaller than extent");
    DLV3P_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0 && aligned16(A) && aligned16(B), DLV3P_ERR_ALIGN,
                  "gemm_bf16: A/B need 16-byte alignment and lda/ldb multiples of 8 (lda=%lld ldb=%lld)",
                  (long long)lda, (long long)ldb);
    DLV3P_REQUIRE((col_scale == nullptr) == (col_shift == nullptr), DLV3P_ERR_SHAPE, "gemm_bf16: scale/shift mismatch");
    DLV3P_REQUIRE(c_dtype == DLV3P_BF16 || c_dtype == DLV3P_F32, DLV3P_ERR_DTYPE, "gemm_bf16: bad c_dtype %d", c_dtype);
    cudaStream_t st = (cudaStream_t)stream;
    const int bn = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
    CUtensorMap tmA, tmB;
    int rc = make_tmap(&tmA, A, K, M, lda, kBlockK, kBlockM);
    if (rc) return rc;
    rc = make_tmap(&tmB, B, K, N, ldb, kBlockK, bn);
    if (rc) return rc;
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.c_dtype = c_dtype;
    p.col_scale = col_scale; p.col_shift = col_shift; p.act = act;
    p.addend = addend; p.ld_add = ld_addend; p.col_stats = col_stats; p.kb_per_split = 0;
    dim3 grid(cdiv(N, bn), cdiv(M, kBlockM), 1);
    const int kbs = cdiv(K, kBlockK);
    switch (bn) {
        case 32: return launch_gemm<32, false>(tmA, tmB, p, grid, st, kbs);
        case 64: return launch_gemm<64, false>(tmA, tmB, p, grid, st, kbs);
        case 128: return launch_gemm<128, false>(tmA, tmB, p, grid, st, kbs);
        default: return launch_gemm<256, false>(tmA, tmB, p, grid, st, kbs);
    }
}

extern "C" int dlv3p_gemm_wgrad_bf16(const void* X, int64_t ldx, const void* dY, int64_t ldy, float* dW, int64_t ldw,
                                     int M, int K, int N, void* stream) {
    DLV3P_REQUIRE(X && dY && dW && M > 0 && N > 0 && K > 0, DLV3P_ERR_SHAPE, "gemm_wgrad_bf16: bad arguments");
    DLV3P_REQUIRE(ldx >= K && ldy >= N && ldw >= N, DLV3P_ERR_SHAPE, "gemm_wgrad_bf16: leading dimension smaller than extent");
    DLV3P_REQUIRE((ldx % 8) == 0 && (ldy % 8) == 0 && aligned16(X) && aligned16(dY), DLV3P_ERR_ALIGN,
                  "gemm_wgrad_bf16: X/dY need 16-byte alignment and ldx/ldy multiples of 8 (ldx=%lld ldy=%lld)",
                  (long long)ldx, (long long)ldy);
    cudaStream_t st = (cudaStream_t)stream;
    const int bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    CUtensorMap tmA, tmB;
    // MN-major: inner (contiguous) axis = channels, outer axis = pixels (the contraction)
    int rc = make_tmap(&tmA, X, K, M, ldx, 64, kBlockK);
    if (rc) return rc;
    rc = make_tmap(&tmB, dY, N, M, ldy, 64, kBlockK);
    if (rc) return rc;
    const int tiles = cdiv(N, bn) * cdiv(K, kBlockM);
    const int total_kb = cdiv(M, kBlockK);
    int splits = cdiv(2 * kNumSMs, tiles);
    if (splits > total_kb) splits = total_kb;
    if (splits < 1) splits = 1;
    const int kb_per_split = cdiv(total_kb, splits);
    splits = cdiv(total_kb, kb_per_split);
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.C = dW; p.ldc = ldw; p.c_dtype = DLV3P_F32;
    p.col_scale = nullptr; p.col_shift = nullptr; p.act = 0; p.addend = nullptr; p.ld_add = 0; p.col_stats = nullptr;
    p.kb_per_split = kb_per_split;
    dim3 grid(cdiv(N, bn), cdiv(K, kBlockM), splits);
    switch (bn) {
        case 64: return launch_gemm<64, true>(tmA, tmB, p, grid, st, kb_per_split);
        case 128: return launch_gemm<128, true>(tmA, tmB, p, grid, st, kb_per_split);
        default: return launch_gemm<256, true>(tmA, tmB, p, grid, st, kb_per_split);
    }
}
