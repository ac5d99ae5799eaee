// TMA / mbarrier plumbing shared by the tensor-core GEMM (gemm_tcgen05.cu) and the halo-staged depthwise
// convolution (dwconv_tma.cu): PTX wrappers and host-side tensor-map encoding through the driver entry point
// (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>

#include "common.cuh"

namespace dlv3p {

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// smem -> global tensor stores (and fp32 reduce-adds) through TMA: coalesced full-line writes with OOB clipping
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_group_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- host side --------------------------------------------------------------------------------------------------
static inline PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
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

// 2D tensor map (bf16, or fp32 with f32 = true), 128B swizzle: inner extent d0 (contiguous), outer extent d1 with row
// pitch ld elements
static inline int make_tmap(CUtensorMap* map, const void* base, long long d0, long long d1, long long ld, int box0,
                            int box1, bool f32 = false) {
    auto fn = get_encode_fn();
    DLV3P_REQUIRE(fn != nullptr, DLV3P_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)d0, (cuuint64_t)d1};
    cuuint64_t strides[1] = {(cuuint64_t)ld * (f32 ? 4 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DLV3P_REQUIRE(rc == CUDA_SUCCESS, DLV3P_ERR_CUDA,
                  "cuTensorMapEncodeTiled failed (%d) dims=(%lld,%lld) ld=%lld box=(%d,%d)", (int)rc, d0, d1, ld, box0,
                  box1);
    return 0;
}


// General rank-3/4 tensor map with 128B swizzle (bf16, or fp32 with f32 = true): dims[0] is the contiguous axis,
// strides_bytes[i] is the pitch of dims[i+1].  The implicit-GEMM convolutions (gemm_tcgen05.cu) use it with
// OVERLAPPING rows: dims[0] = 3*Cin elements (the three horizontally adjacent pixels under one filter row) while the
// pitch of dims[1] (the output column) is only Cin elements.
static inline int make_tmap_nd(CUtensorMap* map, const void* base, int rank, const long long* dims,
                               const long long* strides_bytes, const int* box, bool f32 = false, bool swizzle64 = false) {
    auto fn = get_encode_fn();
    DLV3P_REQUIRE(fn != nullptr, DLV3P_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t d[5], st[4];
    cuuint32_t b[5], estr[5] = {1, 1, 1, 1, 1};
    for (int i = 0; i < rank; ++i) { d[i] = (cuuint64_t)dims[i]; b[i] = (cuuint32_t)box[i]; }
    for (int i = 0; i + 1 < rank; ++i) st[i] = (cuuint64_t)strides_bytes[i];
    CUresult rc = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                     const_cast<void*>(base), d, st, b, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DLV3P_REQUIRE(rc == CUDA_SUCCESS, DLV3P_ERR_CUDA,
                  "cuTensorMapEncodeTiled(rank %d) failed (%d) dims=(%lld,%lld,%lld) box=(%d,%d,%d)", rank, (int)rc,
                  dims[0], dims[1], rank > 2 ? dims[2] : 0, box[0], box[1], rank > 2 ? box[2] : 0);
    return 0;
}

// 4D bf16 NHWC tensor map (dims C, W, H, N), no swizzle, zero fill outside the tensor: the halo of a tile that
// crosses the image border comes back as the convolution's zero padding.
// nan_fill: elements outside the tensor come back as NaN instead (fused BN+ReLU on load, dwconv_tma.cu).
static inline int make_tmap_nhwc(CUtensorMap* map, const void* base, int N, int H, int W, int C, int box_c, int box_w,
                                 int box_h, bool nan_fill = false) {
    auto fn = get_encode_fn();
    DLV3P_REQUIRE(fn != nullptr, DLV3P_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult rc = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DLV3P_REQUIRE(rc == CUDA_SUCCESS, DLV3P_ERR_CUDA,
                  "cuTensorMapEncodeTiled(4D) failed (%d) N=%d H=%d W=%d C=%d box=(%d,%d,%d)", (int)rc, N, H, W, C,
                  box_c, box_w, box_h);
    return 0;
}

}  // namespace dlv3p
