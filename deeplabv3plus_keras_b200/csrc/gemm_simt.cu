// Generic strided SIMT GEMM with fp32 accumulation (FFMA).  It is NOT the product path for the pointwise
// convolutions (that is gemm_tcgen05.cu); it serves (a) the fp32 parity mode (BASELINE cfg-1 runs in float32,
// where bf16/tf32 tensor-core products cannot meet rtol 1e-3), (b) shapes the TMA path cannot describe
// (leading dimensions that are not multiples of 8 elements), and (c) an on-GPU cross-check of the tcgen05 kernel.
#include "common.cuh"

namespace dlv3p {

template <typename TA, typename TC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, long long sam, long long sak, const TA* __restrict__ B, long long sbk,
                 long long sbn, TC* __restrict__ C, long long ldc, int M, int N, int K,
                 const float* __restrict__ col_scale, const float* __restrict__ col_shift, int act,
                 const TC* __restrict__ addend, long long ld_add, int accumulate) {
    constexpr int BM = 64, BN = 64, BK = 16;
    __shared__ float As[BK][BM + 1];
    __shared__ float Bs[BK][BN + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const long long m0 = (long long)blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += BK) {
        // A tile: 64 x 16; pick the thread order that is contiguous in memory
        for (int e = threadIdx.x; e < BM * BK; e += 256) {
            int mm, kk;
            if (sak == 1) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
            const long long m = m0 + mm; const int k = k0 + kk;
            As[kk][mm] = (m < M && k < K) ? to_f<TA>(A[m * sam + k * sak]) : 0.f;
        }
        for (int e = threadIdx.x; e < BN * BK; e += 256) {
            int nn, kk;
            if (sbk == 1) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
            const int n = n0 + nn; const int k = k0 + kk;
            Bs[kk][nn] = (n < N && k < K) ? to_f<TA>(B[(long long)k * sbk + (long long)n * sbn]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (col_scale != nullptr) v = fmaf(v, col_scale[n], col_shift[n]);
            v = apply_act(v, act);
            if (addend != nullptr) v += to_f<TC>(addend[m * ld_add + n]);
            if (accumulate) v += to_f<TC>(C[m * ldc + n]);
            C[m * ldc + n] = from_f<TC>(v);
        }
    }
}

}  // namespace dlv3p

using namespace dlv3p;

extern "C" int dlv3p_gemm_simt(const void* A, int64_t sam, int64_t sak, const void* B, int64_t sbk, int64_t sbn,
                               void* C, int64_t ldc, int M, int N, int K, int ab_dtype, int c_dtype,
                               const float* col_scale, const float* col_shift, int act, const void* addend,
                               int64_t ld_addend, int accumulate, void* stream) {
    DLV3P_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0, DLV3P_ERR_SHAPE, "gemm_simt: bad arguments");
    DLV3P_REQUIRE((col_scale == nullptr) == (col_shift == nullptr), DLV3P_ERR_SHAPE, "gemm_simt: scale/shift mismatch");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(cdiv(N, 64), cdiv(M, 64));
#define DLV3P_SIMT(TA, TC)                                                                                     \
    gemm_simt_kernel<TA, TC><<<grid, 256, 0, st>>>((const TA*)A, sam, sak, (const TA*)B, sbk, sbn, (TC*)C, ldc, M, \
                                                   N, K, col_scale, col_shift, act, (const TC*)addend, ld_addend, \
                                                   accumulate)
    if (ab_dtype == DLV3P_F32 && c_dtype == DLV3P_F32) DLV3P_SIMT(float, float);
    else if (ab_dtype == DLV3P_BF16 && c_dtype == DLV3P_BF16) DLV3P_SIMT(__nv_bfloat16, __nv_bfloat16);
    else if (ab_dtype == DLV3P_BF16 && c_dtype == DLV3P_F32) DLV3P_SIMT(__nv_bfloat16, float);
    else if (ab_dtype == DLV3P_F32 && c_dtype == DLV3P_BF16) DLV3P_SIMT(float, __nv_bfloat16);
    else { set_error("gemm_simt: unsupported dtypes %d,%d", ab_dtype, c_dtype); return DLV3P_ERR_DTYPE; }
#undef DLV3P_SIMT
    return check_launch("gemm_simt");
}
