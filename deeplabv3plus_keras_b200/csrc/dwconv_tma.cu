// K1 (dense taps) — depthwise 3x3, stride 1, dilation 1, bf16 NHWC: TMA halo staging + register sliding window.
//
// The entry-flow depthwise convolutions ([16,254,254,128] etc.) are the HBM-bound launches that decide K1's share
// of the step.  Design:
//   * persistent CTAs walk output tiles of TH x TW pixels x 64 channels; one elected thread issues ONE 4D TMA box
//     {64 ch, TW+2, TH+2, 1} per tile into a double-buffered shared-memory stage (mbarrier expect_tx) — rows/cols
//     outside the image are zero-filled by TMA, which IS the convolution's zero padding (TF SAME or VALID);
//   * 256 threads = 8 channel-octets x 32 columns; each thread slides a 3-row register window down the TH rows of
//     its column: 3 LDS.128 + 1 STG.128 per output (the direct kernel issues 4.5-9 global loads per output), the
//     bf16->fp32 conversion and the fused pre-activation are done once per loaded element, and the 72 MACs per
//     output run as 36 packed FFMA2 (fma.rn.f32x2, sm_100);
//   * the next tile's TMA is in flight while the current one is computed.
// The same kernel serves the input gradient (flipped taps, complementary padding, activation-derivative mask and
// gradient-accumulation addend in the epilogue).
#include "tma.cuh"

namespace dlv3p {

constexpr int kDwTH = 8, kDwTW = 32, kDwCB = 64;
constexpr int kDwStageBytes = (kDwTH + 2) * (kDwTW + 2) * kDwCB * 2;      // 43,520 B
constexpr int kDwThreads = 256;

struct DwTmaParams {
    int N, Hin, Win, C, Hout, Wout, pad_t, pad_l, flip, in_act;
    const float* w;                         // [3,3,C] fp32
    __nv_bfloat16* out;
    const __nv_bfloat16* mask_src; const float* m_scale; const float* m_shift; int m_act;
    const __nv_bfloat16* addend;
    int tiles_h, tiles_w, tiles_c;
    long long num_tiles;
};

__device__ __forceinline__ void bf16x8_to_f32x2(const uint4& raw, float2 (&f)[4]) {
    const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[i].x = __uint_as_float(u[i] << 16);
        f[i].y = __uint_as_float(u[i] & 0xffff0000u);
    }
}

__global__ void __launch_bounds__(kDwThreads, 1)
dw_conv_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const DwTmaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kDwStageBytes);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t stage0 = smem_u32(smem);

    const int tid = threadIdx.x;
    const int cv = tid & 7;                 // channel octet inside the 64-channel block
    const int col = tid >> 3;               // output column inside the tile, 0..31

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_in)) : "memory");
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](long long tile, int s) {
        // tile -> (n, th, tw, cb), channel block fastest so that neighbouring CTAs stream the same pixels
        const int cb = (int)(tile % p.tiles_c); long long t = tile / p.tiles_c;
        const int tw = (int)(t % p.tiles_w); t /= p.tiles_w;
        const int th = (int)(t % p.tiles_h);
        const int n = (int)(t / p.tiles_h);
        mbar_expect_tx(bar0 + 8 * s, kDwStageBytes);
        tma_load_4d(stage0 + s * kDwStageBytes, &tm_in, bar0 + 8 * s, cb * kDwCB, tw * kDwTW - p.pad_l,
                    th * kDwTH - p.pad_t, n);
    };

    long long tile = blockIdx.x;
    if (tile < p.num_tiles && tid == 0) issue(tile, 0);

    uint32_t it = 0;
    for (; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const long long next = tile + gridDim.x;
        if (next < p.num_tiles && tid == 0) issue(next, s ^ 1);      // stage s^1 was released by the barrier below

        const int cb = (int)(tile % p.tiles_c); long long t = tile / p.tiles_c;
        const int tw = (int)(t % p.tiles_w); t /= p.tiles_w;
        const int th = (int)(t % p.tiles_h);
        const int n = (int)(t / p.tiles_h);
        const int c0 = cb * kDwCB + cv * 8;
        const int wo = tw * kDwTW + col;
        const bool lane_ok = (c0 < p.C) && (wo < p.Wout);

        // filter taps of this thread's 8 channels as packed pairs (flipped for the input gradient)
        float2 wgt[9][4];
        if (lane_ok) {
#pragma unroll
            for (int a = 0; a < 9; ++a) {
                const int tap = p.flip ? (8 - a) : a;
                const float4 lo = __ldg(reinterpret_cast<const float4*>(p.w + tap * p.C + c0));
                const float4 hi = __ldg(reinterpret_cast<const float4*>(p.w + tap * p.C + c0) + 1);
                wgt[a][0] = make_float2(lo.x, lo.y); wgt[a][1] = make_float2(lo.z, lo.w);
                wgt[a][2] = make_float2(hi.x, hi.y); wgt[a][3] = make_float2(hi.z, hi.w);
            }
        }

        mbar_wait(bar0 + 8 * s, (it >> 1) & 1u);

        if (lane_ok) {
            const uint8_t* tile_smem = smem + s * kDwStageBytes;
            // smem tile layout: [row 0..TH+1][col 0..TW+1][64 ch] bf16
            auto load_row = [&](int row, float2 (&dst)[3][4]) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const uint4 raw = *reinterpret_cast<const uint4*>(
                        tile_smem + ((row * (kDwTW + 2) + col + j) * kDwCB + cv * 8) * 2);
                    bf16x8_to_f32x2(raw, dst[j]);
                    if (p.in_act != DLV3P_ACT_NONE) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            dst[j][k].x = apply_act(dst[j][k].x, p.in_act);
                            dst[j][k].y = apply_act(dst[j][k].y, p.in_act);
                        }
                    }
                }
            };
            float2 r0[3][4], r1[3][4], r2[3][4];
            load_row(0, r0);
            load_row(1, r1);
#pragma unroll
            for (int r = 0; r < kDwTH; ++r) {
                load_row(r + 2, r2);
                float2 acc[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        acc[k] = __ffma2_rn(r0[j][k], wgt[0 * 3 + j][k], acc[k]);
                        acc[k] = __ffma2_rn(r1[j][k], wgt[1 * 3 + j][k], acc[k]);
                        acc[k] = __ffma2_rn(r2[j][k], wgt[2 * 3 + j][k], acc[k]);
                    }
                }
                const int ho = th * kDwTH + r;
                if (ho < p.Hout) {
                    const long long off = (((long long)n * p.Hout + ho) * p.Wout + wo) * p.C + c0;
                    float f[8];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { f[2 * k] = acc[k].x; f[2 * k + 1] = acc[k].y; }
                    if (p.mask_src != nullptr && p.m_act != DLV3P_ACT_NONE) {
                        Vec8<__nv_bfloat16> mv; mv.load(p.mask_src + off);
                        float mf[8]; mv.to_float(mf);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            float u = mf[k];
                            if (p.m_scale != nullptr) u = fmaf(u, __ldg(p.m_scale + c0 + k), __ldg(p.m_shift + c0 + k));
                            f[k] *= act_mask(u, p.m_act);
                        }
                    }
                    if (p.addend != nullptr) {
                        Vec8<__nv_bfloat16> av; av.load(p.addend + off);
                        float af[8]; av.to_float(af);
#pragma unroll
                        for (int k = 0; k < 8; ++k) f[k] += af[k];
                    }
                    Vec8<__nv_bfloat16> o; o.from_float(f);
                    o.store(p.out + off);
                }
#pragma unroll
                for (int j = 0; j < 3; ++j)
#pragma unroll
                    for (int k = 0; k < 4; ++k) { r0[j][k] = r1[j][k]; r1[j][k] = r2[j][k]; }
            }
        }
        __syncthreads();                     // everyone is done reading stage s: it may be refilled next iteration
    }
}

// Returns 1 if the TMA kernel took the launch, 0 if the caller must use the direct kernel, < 0 on error.
int launch_dw_conv_tma(const __nv_bfloat16* in, const float* w, __nv_bfloat16* out, int N, int Hin, int Win, int C,
                       int Hout, int Wout, int pad_t, int pad_l, int flip, int in_act, const __nv_bfloat16* mask_src,
                       const float* m_scale, const float* m_shift, int m_act, const __nv_bfloat16* addend,
                       cudaStream_t st) {
    if (get_encode_fn() == nullptr) return 0;
    CUtensorMap tm;
    int rc = make_tmap_nhwc(&tm, in, N, Hin, Win, C, kDwCB, kDwTW + 2, kDwTH + 2);
    if (rc) return rc;
    DwTmaParams p;
    p.N = N; p.Hin = Hin; p.Win = Win; p.C = C; p.Hout = Hout; p.Wout = Wout; p.pad_t = pad_t; p.pad_l = pad_l;
    p.flip = flip; p.in_act = in_act; p.w = w; p.out = out; p.mask_src = mask_src; p.m_scale = m_scale;
    p.m_shift = m_shift; p.m_act = m_act; p.addend = addend;
    p.tiles_h = cdiv(Hout, kDwTH); p.tiles_w = cdiv(Wout, kDwTW); p.tiles_c = cdiv(C, kDwCB);
    p.num_tiles = (long long)N * p.tiles_h * p.tiles_w * p.tiles_c;
    constexpr int smem = 2 * kDwStageBytes + 128 + 64;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dw_conv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "cudaFuncSetAttribute(dw tma smem=%d): %s", smem, cudaGetErrorString(e));
        configured = true;
    }
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, dw_conv_tma_kernel, kDwThreads, smem);
        if (ctas_per_sm < 1) ctas_per_sm = 1;
    }
    const long long resident = (long long)kNumSMs * ctas_per_sm;      // persistent: exactly one wave
    long long grid = p.num_tiles < resident ? p.num_tiles : resident;
    dw_conv_tma_kernel<<<(int)grid, kDwThreads, smem, st>>>(tm, p);
    rc = check_launch("dwconv3x3 (tma)");
    return rc ? rc : 1;
}

}  // namespace dlv3p
