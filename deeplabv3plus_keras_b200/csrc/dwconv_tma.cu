// K1 (dense taps) — depthwise 3x3, stride 1, dilation 1, bf16 NHWC: TMA halo staging + register sliding window.
//
// The entry-flow depthwise convolutions ([16,254,254,128] etc.) are the HBM-bound launches that decide K1's share
// of the step.  Design:
//   * persistent CTAs walk output tiles of TH x TW pixels x 64 channels; one elected thread issues ONE 4D TMA box
//     {64 ch, TW+2, TH+2, 1} per tile into a 4-deep shared-memory ring (mbarrier expect_tx) — rows/cols outside the
//     image are zero-filled by TMA, which IS the convolution's zero padding (TF SAME or VALID);
//   * 512 threads = 16 channel-quads x 32 columns; each thread slides a 3-row register window down the TH rows of
//     its column: 3 LDS.64 + 1 STG.64 per output (the direct kernel issues 4.5-9 global loads per output); the
//     fused pre-activation runs on packed bf16x2 (HMNMX2) BEFORE the bf16->fp32 widening, and the 36 MACs per
//     4-channel output run as 18 packed FFMA2 (fma.rn.f32x2, sm_100).  4 channels per thread keeps the kernel at
//     <= 128 registers so that 16 warps are resident: the kernel is issue-bound, not latency-bound (measured);
//   * the TMA boxes of the next three tiles are in flight while the current one is computed.
// The same kernel serves the input gradient (flipped taps, complementary padding, activation-derivative mask and
// gradient-accumulation addend in the epilogue, whose operands are prefetched into L2 one tile ahead).
#include "tma.cuh"

namespace dlv3p {

constexpr int kDwTH = 8, kDwTW = 32, kDwCB = 64;
constexpr int kDwStageBytes = (kDwTH + 2) * (kDwTW + 2) * kDwCB * 2;      // 43,520 B
constexpr int kDwThreads = 512;
constexpr int kDwStages = 4;          // 4 x 43.5 KB boxes in flight per SM

struct DwTmaParams {
    int N, Hin, Win, C, Hout, Wout, pad_t, pad_l, flip, in_act;
    const float* w;                         // [3,3,C] fp32
    __nv_bfloat16* out;
    const __nv_bfloat16* mask_src; const float* m_scale; const float* m_shift; int m_act;
    const __nv_bfloat16* addend;
    int tiles_h, tiles_w, tiles_c;
    int num_tiles;
};

// 4 bf16 (uint2) -> optional packed ReLU/ReLU6 -> two float2
__device__ __forceinline__ void widen4(uint2 raw, int act, float2 (&f)[2]) {
    if (act != DLV3P_ACT_NONE) {
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
        const __nv_bfloat162 zero = __float2bfloat162_rn(0.f);
        a = __hmax2(a, zero); b = __hmax2(b, zero);
        if (act == DLV3P_ACT_RELU6) {
            const __nv_bfloat162 six = __float2bfloat162_rn(6.f);
            a = __hmin2(a, six); b = __hmin2(b, six);
        }
        raw.x = *reinterpret_cast<uint32_t*>(&a);
        raw.y = *reinterpret_cast<uint32_t*>(&b);
    }
    f[0].x = __uint_as_float(raw.x << 16); f[0].y = __uint_as_float(raw.x & 0xffff0000u);
    f[1].x = __uint_as_float(raw.y << 16); f[1].y = __uint_as_float(raw.y & 0xffff0000u);
}

__device__ __forceinline__ void decode_tile(const DwTmaParams& p, int tile, int& n, int& th, int& tw, int& cb) {
    // channel block slowest: a CTA's consecutive tiles (stride gridDim.x) keep their filter taps in registers;
    // different channel blocks touch disjoint bytes, so this costs no L2 locality
    tw = tile % p.tiles_w; int t = tile / p.tiles_w;
    th = t % p.tiles_h; t /= p.tiles_h;
    n = t % p.N;
    cb = t / p.N;
}

__global__ void __launch_bounds__(kDwThreads, 1)
dw_conv_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const DwTmaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDwStages * kDwStageBytes);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t stage0 = smem_u32(smem);

    const int tid = threadIdx.x;
    const int cq = tid & 15;                // channel quad inside the 64-channel block
    const int col = tid >> 4;               // output column inside the tile, 0..31

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_in)) : "memory");
        for (int s = 0; s < kDwStages; ++s) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int tile, int s) {
        int n, th, tw, cb;
        decode_tile(p, tile, n, th, tw, cb);
        mbar_expect_tx(bar0 + 8 * s, kDwStageBytes);
        tma_load_4d(stage0 + s * kDwStageBytes, &tm_in, bar0 + 8 * s, cb * kDwCB, tw * kDwTW - p.pad_l,
                    th * kDwTH - p.pad_t, n);
    };

    int tile = blockIdx.x;
    const int gstride = gridDim.x;
    if (tid == 0) {
        for (int a = 0; a < kDwStages - 1; ++a)
            if (tile + a * gstride < p.num_tiles) issue(tile + a * gstride, a);
    }

    float2 wgt[9][2];
    int wgt_c0 = -1;
    uint32_t it = 0;
    for (; tile < p.num_tiles; tile += gstride, ++it) {
        const int s = it % kDwStages;
        // refill the stage released by the barrier at the end of the previous iteration, kDwStages-1 tiles ahead
        const int ahead = tile + (kDwStages - 1) * gstride;
        if (ahead < p.num_tiles && tid == 0) issue(ahead, (it + kDwStages - 1) % kDwStages);

        int n, th, tw, cb;
        decode_tile(p, tile, n, th, tw, cb);
        const int c0 = cb * kDwCB + cq * 4;
        const int wo = tw * kDwTW + col;
        const bool lane_ok = (c0 < p.C) && (wo < p.Wout);

        // filter taps of this thread's 4 channels as packed pairs (flipped for the input gradient)
        if (c0 < p.C && c0 != wgt_c0) {
            wgt_c0 = c0;
#pragma unroll
            for (int a = 0; a < 9; ++a) {
                const int tap = p.flip ? (8 - a) : a;
                const float4 v = __ldg(reinterpret_cast<const float4*>(p.w + tap * p.C + c0));
                wgt[a][0] = make_float2(v.x, v.y); wgt[a][1] = make_float2(v.z, v.w);
            }
        }

        // the epilogue operands (activation-mask source, gradient addend) are read straight from global memory:
        // pull the NEXT tile's lines into L2 now so those loads do not pay HBM latency inside the row loop
        if (cq == 0 && (p.mask_src != nullptr || p.addend != nullptr)) {
            const int nt = tile + gstride;
            if (nt < p.num_tiles) {
                int n2, th2, tw2, cb2;
                decode_tile(p, nt, n2, th2, tw2, cb2);
                const int wo2 = tw2 * kDwTW + col;
                if (wo2 < p.Wout) {
#pragma unroll
                    for (int r = 0; r < kDwTH; ++r) {
                        const int ho2 = th2 * kDwTH + r;
                        if (ho2 < p.Hout) {
                            const long long o2 = (((long long)n2 * p.Hout + ho2) * p.Wout + wo2) * p.C + cb2 * kDwCB;
                            if (p.mask_src != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.mask_src + o2));
                            if (p.addend != nullptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.addend + o2));
                        }
                    }
                }
            }
        }

        mbar_wait(bar0 + 8 * s, (it / kDwStages) & 1u);

        if (lane_ok) {
            // smem tile layout: [row 0..TH+1][col 0..TW+1][64 ch] bf16; this thread's 3 input columns start here
            const uint32_t base = stage0 + s * kDwStageBytes + (col * kDwCB + cq * 4) * 2;
            const int in_act = p.in_act;
            auto load_row = [&](int row, float2 (&dst)[3][2]) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    uint2 raw;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                                 : "=r"(raw.x), "=r"(raw.y)
                                 : "r"(base + (uint32_t)((row * (kDwTW + 2) + j) * (kDwCB * 2))));
                    widen4(raw, in_act, dst[j]);
                }
            };
            const long long row_stride = (long long)p.Wout * p.C;
            long long off = (((long long)n * p.Hout + th * kDwTH) * p.Wout + wo) * p.C + c0;
            const int rows_valid = min(kDwTH, p.Hout - th * kDwTH);
            const bool has_mask = (p.mask_src != nullptr && p.m_act != DLV3P_ACT_NONE);
            float msc[4] = {1.f, 1.f, 1.f, 1.f}, msh[4] = {0.f, 0.f, 0.f, 0.f};
            if (has_mask && p.m_scale != nullptr) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { msc[k] = __ldg(p.m_scale + c0 + k); msh[k] = __ldg(p.m_shift + c0 + k); }
            }
            // one output row from the three window rows (ra above, rb centre, rc below)
            auto emit = [&](int r, const float2 (&ra)[3][2], const float2 (&rb)[3][2], const float2 (&rc)[3][2]) {
                float2 acc[2];
                acc[0] = make_float2(0.f, 0.f); acc[1] = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        acc[k] = __ffma2_rn(ra[j][k], wgt[0 * 3 + j][k], acc[k]);
                        acc[k] = __ffma2_rn(rb[j][k], wgt[1 * 3 + j][k], acc[k]);
                        acc[k] = __ffma2_rn(rc[j][k], wgt[2 * 3 + j][k], acc[k]);
                    }
                }
                if (r < rows_valid) {
                    float f[4] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y};
                    if (has_mask) {
                        const uint2 mraw = __ldg(reinterpret_cast<const uint2*>(p.mask_src + off));
                        float2 mf[2];
                        widen4(mraw, DLV3P_ACT_NONE, mf);
                        const float u[4] = {mf[0].x, mf[0].y, mf[1].x, mf[1].y};
#pragma unroll
                        for (int k = 0; k < 4; ++k) f[k] *= act_mask(fmaf(u[k], msc[k], msh[k]), p.m_act);
                    }
                    if (p.addend != nullptr) {
                        const uint2 araw = __ldg(reinterpret_cast<const uint2*>(p.addend + off));
                        float2 af[2];
                        widen4(araw, DLV3P_ACT_NONE, af);
                        f[0] += af[0].x; f[1] += af[0].y; f[2] += af[1].x; f[3] += af[1].y;
                    }
                    uint2 o;
                    __nv_bfloat162 lo = __floats2bfloat162_rn(f[0], f[1]), hi = __floats2bfloat162_rn(f[2], f[3]);
                    o.x = *reinterpret_cast<uint32_t*>(&lo);
                    o.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(p.out + off) = o;
                }
                off += row_stride;
            };
            // the three window rows rotate roles instead of being copied (no register moves)
            float2 r0[3][2], r1[3][2], r2[3][2];
            load_row(0, r0);
            load_row(1, r1);
            static_assert(kDwTH == 8, "row loop below is unrolled for TH = 8");
            load_row(2, r2); emit(0, r0, r1, r2);
            load_row(3, r0); emit(1, r1, r2, r0);
            load_row(4, r1); emit(2, r2, r0, r1);
            load_row(5, r2); emit(3, r0, r1, r2);
            load_row(6, r0); emit(4, r1, r2, r0);
            load_row(7, r1); emit(5, r2, r0, r1);
            load_row(8, r2); emit(6, r0, r1, r2);
            load_row(9, r0); emit(7, r1, r2, r0);
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
    const long long nt = (long long)N * p.tiles_h * p.tiles_w * p.tiles_c;
    if (nt > 0x7fffffffLL) return 0;
    p.num_tiles = (int)nt;
    constexpr int smem = kDwStages * kDwStageBytes + 128 + 64;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dw_conv_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "cudaFuncSetAttribute(dw tma smem=%d): %s", smem, cudaGetErrorString(e));
        configured = true;
    }
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;   // persistent: one CTA per SM
    dw_conv_tma_kernel<<<grid, kDwThreads, smem, st>>>(tm, p);
    rc = check_launch("dwconv3x3 (tma)");
    return rc ? rc : 1;
}

}  // namespace dlv3p
