// Generic strided SIMT GEMM with fp32 accumulation (FFMA).  It is NOT the product path for the pointwise
// convolutions (that is gemm_tcgen05.cu); it serves (a) the fp32 parity mode (BASELINE cfg-1 runs in float32,
// where bf16/tf32 tensor-core products cannot meet rtol 1e-3), (b) shapes the TMA path cannot describe
// (leading dimensions that are not multiples of 8 elements), and (c) an on-GPU cross-check of the tcgen05 kernel.
#include "common.cuh"

namespace dlv3p {

// Tile BM x BN (64 x 64 or 32 x 32), 256 threads, each thread a (BM/16) x (BN/16) register block; `splits` > 1: the K
// range is cut over blockIdx.z and the partial products are accumulated into C with fp32 atomics (C pre-zeroed or
// accumulating; no epilogue in that mode).  The small-tile / split-K forms exist for the batch-1 inference shapes of
// BASELINE cfg-1 (M = 33 x 33 = 1089 pixels: a 64 x 64 grid is 18..72 CTAs on 148 SMs, and the 21-class logits
// convolution after im2col is M x N = 1089 x 21 with K = 2304: ONE column of 18 CTAs each looping 144 k-steps).
template <typename TA, typename TC, int BM, int BN>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const TA* __restrict__ A, long long sam, long long sak, const TA* __restrict__ B, long long sbk,
                 long long sbn, TC* __restrict__ C, long long ldc, int M, int N, int K,
                 const float* __restrict__ col_scale, const float* __restrict__ col_shift, int act,
                 const TC* __restrict__ addend, long long ld_add, int accumulate, int k_per_split) {
    constexpr int BK = 16, RM = BM / 16, RN = BN / 16;
    __shared__ float As[BK][BM + 1];
    __shared__ float Bs[BK][BN + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const long long m0 = (long long)blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;
    const int kb = blockIdx.z * k_per_split;
    const int ke = min(K, kb + k_per_split);
    float acc[RM][RN];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;

    for (int k0 = kb; k0 < ke; k0 += BK) {
        // A tile: BM x 16; pick the thread order that is contiguous in memory
        for (int e = threadIdx.x; e < BM * BK; e += 256) {
            int mm, kk;
            if (sak == 1) { kk = e % BK; mm = e / BK; } else { mm = e % BM; kk = e / BM; }
            const long long m = m0 + mm; const int k = k0 + kk;
            As[kk][mm] = (m < M && k < ke) ? to_f<TA>(A[m * sam + k * sak]) : 0.f;
        }
        for (int e = threadIdx.x; e < BN * BK; e += 256) {
            int nn, kk;
            if (sbk == 1) { kk = e % BK; nn = e / BK; } else { nn = e % BN; kk = e / BN; }
            const int n = n0 + nn; const int k = k0 + kk;
            Bs[kk][nn] = (n < N && k < ke) ? to_f<TA>(B[(long long)k * sbk + (long long)n * sbn]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[RM], b[RN];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = As[kk][ty * RM + i];
#pragma unroll
            for (int j = 0; j < RN; ++j) b[j] = Bs[kk][tx * RN + j];
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const long long m = m0 + ty * RM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < RN; ++j) {
            const int n = n0 + tx * RN + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (split) {                                   // fp32 C only (checked by the launcher)
                atomicAdd(reinterpret_cast<float*>(C) + m * ldc + n, v);
                continue;
            }
            if (col_scale != nullptr) v = fmaf(v, col_scale[n], col_shift[n]);
            v = apply_act(v, act);
            if (addend != nullptr) v += ld_rw_f<TC>(addend + m * ld_add + n);    // may alias C
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
    // tile / split choice: fill the 148 SMs.  64 x 64 tiles when they give at least two waves; else 32 x 32; if that is
    // still under one wave and the call has no epilogue and an fp32 C, cut K over blockIdx.z (atomic accumulation)
    const long long ctas64 = (long long)cdiv(N, 64) * cdiv(M, 64);
    const bool small = ctas64 < 2 * kNumSMs;
    const long long ctas = small ? (long long)cdiv(N, 32) * cdiv(M, 32) : ctas64;
    int splits = 1;
    const bool can_split = c_dtype == DLV3P_F32 && col_scale == nullptr && act == DLV3P_ACT_NONE && addend == nullptr;
    if (can_split && ctas < kNumSMs && K >= 256) {
        splits = (int)((2 * kNumSMs + ctas - 1) / ctas);
        if (splits > K / 64) splits = K / 64;
        if (splits < 1) splits = 1;
    }
    int kps = cdiv(cdiv(K, splits), 16) * 16;
    splits = cdiv(K, kps);
    if (splits > 1 && !accumulate) {
        // partial products are added into C: start from zero (row pitch ldc may exceed N: clear row by row)
        cudaError_t e = cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, st);
        DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "gemm_simt: memset failed: %s", cudaGetErrorString(e));
    }
#define DLV3P_SIMT(TA, TC)                                                                                          \
    do {                                                                                                            \
        if (small)                                                                                                  \
            gemm_simt_kernel<TA, TC, 32, 32><<<dim3(cdiv(N, 32), cdiv(M, 32), splits), 256, 0, st>>>(               \
                (const TA*)A, sam, sak, (const TA*)B, sbk, sbn, (TC*)C, ldc, M, N, K, col_scale, col_shift, act,    \
                (const TC*)addend, ld_addend, accumulate, kps);                                                     \
        else                                                                                                        \
            gemm_simt_kernel<TA, TC, 64, 64><<<dim3(cdiv(N, 64), cdiv(M, 64), 1), 256, 0, st>>>(                    \
                (const TA*)A, sam, sak, (const TA*)B, sbk, sbn, (TC*)C, ldc, M, N, K, col_scale, col_shift, act,    \
                (const TC*)addend, ld_addend, accumulate, K);                                                       \
    } while (0)
    if (ab_dtype == DLV3P_F32 && c_dtype == DLV3P_F32) DLV3P_SIMT(float, float);
    else if (ab_dtype == DLV3P_BF16 && c_dtype == DLV3P_BF16) DLV3P_SIMT(__nv_bfloat16, __nv_bfloat16);
    else if (ab_dtype == DLV3P_BF16 && c_dtype == DLV3P_F32) DLV3P_SIMT(__nv_bfloat16, float);
    else if (ab_dtype == DLV3P_F32 && c_dtype == DLV3P_BF16) DLV3P_SIMT(float, __nv_bfloat16);
    else { set_error("gemm_simt: unsupported dtypes %d,%d", ab_dtype, c_dtype); return DLV3P_ERR_DTYPE; }
#undef DLV3P_SIMT
    return check_launch("gemm_simt");
}
