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

struct DwTmaParams {
    int N, Hin, Win, C, Hout, Wout, pad_t, pad_l, flip, in_act;
    const float* w;                         // [3,3,C] fp32
    __nv_bfloat16* out;
    const __nv_bfloat16* mask_src; const float* m_scale; const float* m_shift; int m_act;
    const __nv_bfloat16* addend;
    const float* in_scale; const float* in_shift;             // IN_AFFINE: x := act(in_scale*x + in_shift) on load
    const float* bn_mean; const float* bn_invstd; float* bn_red;   // STATS: BN-backward reductions of the masked result
    // IN_BN: the BatchNormalization of the producing layer is FINISHED here (dlv3p_bn_finalize folded into the reader):
    // in_scale/in_shift are computed per CTA from the batch sums; the first CTA of every channel block publishes
    // scale/shift/mean/invstd for the backward pass and updates the moving statistics
    const float* f_sums; const float* f_gamma; const float* f_beta; float* f_mm; float* f_mv;
    float* f_scale; float* f_shift; float* f_mean; float* f_invstd;
    double f_count, f_inv_count; float f_eps, f_momentum; int f_updates;
    const float* out_scale; const float* out_shift;           // OUT_EPI: y := act(out_scale*conv + out_shift), out_act
    int out_act;
    int tiles_h, tiles_w, tiles_c;
    int spatial_tiles, ctas_per_cb;
};

// fused BatchNormalization + ReLU/ReLU6 on load (IN_AFFINE): the tile arrives through a tensor map whose out-of-bounds
// fill is NaN, NaN survives the affine map and max(NaN, 0) = 0 (FMNMX returns the non-NaN operand), so the
// convolution's zero padding comes out exact without a single border predicate in the row loop.
template <int ACT>
__device__ __forceinline__ void affine_act4(float2 (&f)[2], const float2 (&sc)[2], const float2 (&sh)[2]) {
    static_assert(ACT != DLV3P_ACT_NONE, "the NaN padding needs a clamping activation");
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        f[k] = __ffma2_rn(f[k], sc[k], sh[k]);
        f[k].x = fmaxf(f[k].x, 0.f); f[k].y = fmaxf(f[k].y, 0.f);
        if (ACT == DLV3P_ACT_RELU6) { f[k].x = fminf(f[k].x, 6.f); f[k].y = fminf(f[k].y, 6.f); }
    }
}

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

// Compile-time specialisation: the row loop is pure straight-line code (the first version branched on the activation /
// mask / addend codes at run time and spent 14 % of its issue slots on ISETP+BRA and 2x the HMNMX2 it needed — ncu
// source page, profiles/r1_dwfwd728_*).  IN_ACT: activation fused on load; M_ACT: activation whose derivative masks
// the result (input gradient), 0 = none; M_AFFINE: mask argument is m_scale*x+m_shift; HAS_ADD: gradient addend.
template <int ACT>
__device__ __forceinline__ void widen4_t(uint2 raw, float2 (&f)[2]) {
    if (ACT != DLV3P_ACT_NONE) {
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&raw.x);
        __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&raw.y);
        const __nv_bfloat162 zero = __float2bfloat162_rn(0.f);
        a = __hmax2(a, zero); b = __hmax2(b, zero);
        if (ACT == DLV3P_ACT_RELU6) {
            const __nv_bfloat162 six = __float2bfloat162_rn(6.f);
            a = __hmin2(a, six); b = __hmin2(b, six);
        }
        raw.x = *reinterpret_cast<uint32_t*>(&a);
        raw.y = *reinterpret_cast<uint32_t*>(&b);
    }
    f[0].x = __uint_as_float(raw.x << 16); f[0].y = __uint_as_float(raw.x & 0xffff0000u);
    f[1].x = __uint_as_float(raw.y << 16); f[1].y = __uint_as_float(raw.y & 0xffff0000u);
}

// Epilogue operands (mask source, addend) arrive through their own TMA boxes {64 ch, TW, TH} in the same stage, so
// the row loop never waits on a global load; the ring gets shallower as the stage grows (4 / 3 / 2 stages).
template <int M_ACT, bool HAS_ADD>
struct DwStageCfg {
    static constexpr int kBoxes = 1 + (M_ACT != DLV3P_ACT_NONE ? 1 : 0) + (HAS_ADD ? 1 : 0);
    static constexpr int kEpiBytes = kDwTH * kDwTW * kDwCB * 2;                    // 32,768 B
    static constexpr int kStageBytes = kDwStageBytes + (kBoxes - 1) * kEpiBytes;
    static constexpr int kStages = kBoxes == 1 ? 4 : (kBoxes == 2 ? 3 : 2);
    static constexpr int kMaskOff = kDwStageBytes;
    static constexpr int kAddOff = kDwStageBytes + (M_ACT != DLV3P_ACT_NONE ? kEpiBytes : 0);
};

template <int IN_ACT, int M_ACT, bool M_AFFINE, bool HAS_ADD, bool IN_AFFINE, bool STATS, bool IN_BN = false,
          bool OUT_EPI = false>
__global__ void __launch_bounds__(kDwThreads, 1)
dw_conv_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_mask,
                   const __grid_constant__ CUtensorMap tm_add, const DwTmaParams p) {
    using Cfg = DwStageCfg<M_ACT, HAS_ADD>;
    constexpr int kDwStages = Cfg::kStages;
    constexpr int kStageBytes = Cfg::kStageBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDwStages * kStageBytes);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t stage0 = smem_u32(smem);

    const int tid = threadIdx.x;
    const int cq = tid & 15;                // channel quad inside the 64-channel block
    const int col = tid >> 4;               // output column inside the tile, 0..31
    // every CTA owns ONE 64-channel block: its filter taps are loaded once and the tile decode needs no channel term
    const int cb = blockIdx.x / p.ctas_per_cb;
    const int gstride = p.ctas_per_cb;
    const int c0 = cb * kDwCB + cq * 4;
    pdl_launch_dependents();

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_in)) : "memory");
        if (M_ACT != DLV3P_ACT_NONE) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_mask)) : "memory");
        if (HAS_ADD) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_add)) : "memory");
        for (int s = 0; s < kDwStages; ++s) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // filter taps of this thread's 4 channels as packed pairs (flipped for the input gradient).  Loaded BEFORE the
    // programmatic-dependency wait: the fp32 master weights only change in the optimizer step, whose kernels do not
    // trigger dependent launches, so this L2 round trip overlaps the previous kernel's tail
    float2 wgt[9][2];
    const bool ch_ok = c0 < p.C;
    if (ch_ok) {
#pragma unroll
        for (int a = 0; a < 9; ++a) {
            const int tap = p.flip ? (8 - a) : a;
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.w + tap * p.C + c0));
            wgt[a][0] = make_float2(v.x, v.y); wgt[a][1] = make_float2(v.z, v.w);
        }
    }
    pdl_wait();

    auto decode = [&](int tile, int& n, int& th, int& tw) {
        tw = tile % p.tiles_w; const int t = tile / p.tiles_w;
        th = t % p.tiles_h; n = t / p.tiles_h;
    };
    auto issue = [&](int tile, int s) {
        int n, th, tw;
        decode(tile, n, th, tw);
        mbar_expect_tx(bar0 + 8 * s, kStageBytes);
        const uint32_t dst = stage0 + s * kStageBytes;
        tma_load_4d(dst, &tm_in, bar0 + 8 * s, cb * kDwCB, tw * kDwTW - p.pad_l, th * kDwTH - p.pad_t, n);
        if (M_ACT != DLV3P_ACT_NONE)
            tma_load_4d(dst + Cfg::kMaskOff, &tm_mask, bar0 + 8 * s, cb * kDwCB, tw * kDwTW, th * kDwTH, n);
        if (HAS_ADD)
            tma_load_4d(dst + Cfg::kAddOff, &tm_add, bar0 + 8 * s, cb * kDwCB, tw * kDwTW, th * kDwTH, n);
    };

    int tile = blockIdx.x % p.ctas_per_cb;
    if (tid == 0) {
        for (int a = 0; a < kDwStages - 1; ++a)
            if (tile + a * gstride < p.spatial_tiles) issue(tile + a * gstride, a);
    }

    float msc[4] = {1.f, 1.f, 1.f, 1.f}, msh[4] = {0.f, 0.f, 0.f, 0.f};
    float osc[4] = {1.f, 1.f, 1.f, 1.f}, osh[4] = {0.f, 0.f, 0.f, 0.f};
    float2 isc[2], ish[2];
    float bs1[4] = {0.f, 0.f, 0.f, 0.f}, bs2[4] = {0.f, 0.f, 0.f, 0.f};
    if (ch_ok) {
        if (IN_AFFINE && IN_BN) {
            // same arithmetic as bn_finalize_kernel (eltwise.cu): fp64 for E[x^2] - E[x]^2, TF fused-BN conventions
            float sc4[4], sh4[4];
            const bool publish = (blockIdx.x % p.ctas_per_cb == 0) && (col == 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = c0 + k;
                // fp64 only where it matters (E[x^2] - E[x]^2 cancels); the division and the square root of
                // bn_finalize_kernel are a multiplication by 1/count and an fp32 rsqrt here: every thread of the CTA runs
                // this prologue on the critical path right after the dependency wait
                const double inv_count = p.f_inv_count;
                const double m = (double)__ldcg(p.f_sums + c) * inv_count;
                double var = (double)__ldcg(p.f_sums + p.C + c) * inv_count - m * m;
                if (var < 0.0) var = 0.0;
                const float is = rsqrtf((float)var + p.f_eps);
                const float g = p.f_gamma ? __ldg(p.f_gamma + c) : 1.f;
                const float b = p.f_beta ? __ldg(p.f_beta + c) : 0.f;
                sc4[k] = g * is;
                sh4[k] = b - (float)m * sc4[k];
                if (publish) {
                    p.f_scale[c] = sc4[k]; p.f_shift[c] = sh4[k]; p.f_mean[c] = (float)m; p.f_invstd[c] = is;
                    const double unbiased = p.f_count > 1.0 ? var * p.f_count / (p.f_count - 1.0) : var;
                    float mm = p.f_mm[c], mv = p.f_mv[c];
                    for (int u = 0; u < p.f_updates; ++u) {
                        mm = p.f_momentum * mm + (1.f - p.f_momentum) * (float)m;
                        mv = p.f_momentum * mv + (1.f - p.f_momentum) * (float)unbiased;
                    }
                    p.f_mm[c] = mm; p.f_mv[c] = mv;
                }
            }
            isc[0] = make_float2(sc4[0], sc4[1]); isc[1] = make_float2(sc4[2], sc4[3]);
            ish[0] = make_float2(sh4[0], sh4[1]); ish[1] = make_float2(sh4[2], sh4[3]);
        } else if (IN_AFFINE) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p.in_scale + c0));
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.in_shift + c0));
            isc[0] = make_float2(a.x, a.y); isc[1] = make_float2(a.z, a.w);
            ish[0] = make_float2(b.x, b.y); ish[1] = make_float2(b.z, b.w);
        }
        if (M_ACT != DLV3P_ACT_NONE && M_AFFINE) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { msc[k] = __ldg(p.m_scale + c0 + k); msh[k] = __ldg(p.m_shift + c0 + k); }
        }
        if (OUT_EPI) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { osc[k] = __ldg(p.out_scale + c0 + k); osh[k] = __ldg(p.out_shift + c0 + k); }
        }
    }
    const long long row_stride = (long long)p.Wout * p.C;

    uint32_t it = 0;
    for (; tile < p.spatial_tiles; tile += gstride, ++it) {
        const int s = it % kDwStages;
        // refill the stage released by the barrier at the end of the previous iteration, kDwStages-1 tiles ahead
        const int ahead = tile + (kDwStages - 1) * gstride;
        if (ahead < p.spatial_tiles && tid == 0) issue(ahead, (it + kDwStages - 1) % kDwStages);

        int n, th, tw;
        decode(tile, n, th, tw);
        const int wo = tw * kDwTW + col;
        const bool lane_ok = ch_ok && (wo < p.Wout);

        mbar_wait(bar0 + 8 * s, (it / kDwStages) & 1u);

        if (lane_ok) {
            // smem tile layout: [row 0..TH+1][col 0..TW+1][64 ch] bf16; this thread's 3 input columns start here
            const uint32_t base = stage0 + s * kStageBytes + (col * kDwCB + cq * 4) * 2;
            auto load_row = [&](int row, float2 (&dst)[3][2]) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    uint2 raw;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                                 : "=r"(raw.x), "=r"(raw.y)
                                 : "r"(base + (uint32_t)((row * (kDwTW + 2) + j) * (kDwCB * 2))));
                    if (IN_AFFINE) {
                        widen4_t<DLV3P_ACT_NONE>(raw, dst[j]);
                        affine_act4<IN_AFFINE ? (IN_ACT == DLV3P_ACT_NONE ? DLV3P_ACT_RELU : IN_ACT) : DLV3P_ACT_RELU>(dst[j], isc, ish);
                    } else {
                        widen4_t<IN_ACT>(raw, dst[j]);
                    }
                }
            };
            auto lds_epi = [&](int byte_off, int r) {
                uint2 raw;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                             : "=r"(raw.x), "=r"(raw.y)
                             : "r"(base + (uint32_t)(byte_off + r * kDwTW * kDwCB * 2)));
                return raw;
            };
            long long off = (((long long)n * p.Hout + th * kDwTH) * p.Wout + wo) * p.C + c0;
            const int rows_valid = p.Hout - th * kDwTH;
            // one output row from the three window rows (ra above, rb centre, rc below); three independent
            // accumulation chains per channel pair (one per window row) keep the FMA pipe fed at 4 warps/scheduler
            auto emit = [&](int r, const float2 (&ra)[3][2], const float2 (&rb)[3][2], const float2 (&rc)[3][2]) {
                // two accumulation chains per channel pair (taps 0..4 / 5..8) and one packed add: the kernel is bound by the
                // FP32 pipe (a packed FMA on three distinct register pairs holds it for 3 cycles, scripts/ubench/fma_rate.cu),
                // a three-chain sum spends two packed adds per pair where one is enough to cover the FMA latency
                float2 acc[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    float2 a0 = __fmul2_rn(ra[0][k], wgt[0][k]);
                    float2 a1 = __fmul2_rn(rb[2][k], wgt[5][k]);
                    a0 = __ffma2_rn(ra[1][k], wgt[1][k], a0);
                    a1 = __ffma2_rn(rc[0][k], wgt[6][k], a1);
                    a0 = __ffma2_rn(ra[2][k], wgt[2][k], a0);
                    a1 = __ffma2_rn(rc[1][k], wgt[7][k], a1);
                    a0 = __ffma2_rn(rb[0][k], wgt[3][k], a0);
                    a1 = __ffma2_rn(rc[2][k], wgt[8][k], a1);
                    a0 = __ffma2_rn(rb[1][k], wgt[4][k], a0);
                    acc[k] = __fadd2_rn(a0, a1);
                }
                if (r < rows_valid) {
                    float f[4] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y};
                    if (M_ACT != DLV3P_ACT_NONE) {
                        const uint2 mraw = lds_epi(Cfg::kMaskOff, r);
                        float2 mf[2];
                        widen4_t<DLV3P_ACT_NONE>(mraw, mf);
                        const float u[4] = {mf[0].x, mf[0].y, mf[1].x, mf[1].y};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float v = M_AFFINE ? fmaf(u[k], msc[k], msh[k]) : u[k];
                            const bool on = (M_ACT == DLV3P_ACT_RELU) ? (v > 0.f) : (v > 0.f && v < 6.f);
                            f[k] = on ? f[k] : 0.f;
                            if (STATS) { bs1[k] += f[k]; bs2[k] = fmaf(f[k], u[k], bs2[k]); }
                        }
                    }
                    if (HAS_ADD) {
                        const uint2 araw = lds_epi(Cfg::kAddOff, r);
                        float2 af[2];
                        widen4_t<DLV3P_ACT_NONE>(araw, af);
                        f[0] += af[0].x; f[1] += af[0].y; f[2] += af[1].x; f[3] += af[1].y;
                    }
                    if (OUT_EPI) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) f[k] = apply_act(fmaf(f[k], osc[k], osh[k]), p.out_act);
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
    if (STATS) {
        // BatchNormalization-backward reductions of the layer that produced the mask source y (its raw conv output):
        // red[0..C) += sum g, red[C..2C) += sum g*xhat with g = the masked gradient written above and
        // sum g*xhat = invstd * (sum g*y - mean * sum g) — the contract of dlv3p_bn_bwd_reduce.  Every TMA box this CTA
        // issued has been consumed: stage 0 is reused as the [2][32 cols][64 ch] reduction buffer.
        float* red = reinterpret_cast<float*>(smem);
        *reinterpret_cast<float4*>(red + col * kDwCB + cq * 4) = make_float4(bs1[0], bs1[1], bs1[2], bs1[3]);
        *reinterpret_cast<float4*>(red + (kDwTW + col) * kDwCB + cq * 4) = make_float4(bs2[0], bs2[1], bs2[2], bs2[3]);
        __syncthreads();
        if (tid < kDwCB) {
            const int ch = cb * kDwCB + tid;
            if (ch < p.C) {
                float a1 = 0.f, a2 = 0.f;
#pragma unroll 8
                for (int q = 0; q < kDwTW; ++q) { a1 += red[q * kDwCB + tid]; a2 += red[(kDwTW + q) * kDwCB + tid]; }
                a2 = (a2 - __ldg(p.bn_mean + ch) * a1) * __ldg(p.bn_invstd + ch);
                atomicAdd(p.bn_red + ch, a1);
                atomicAdd(p.bn_red + p.C + ch, a2);
            }
        }
    }
}

template <int IN_ACT, int M_ACT, bool M_AFFINE, bool HAS_ADD, bool IN_AFFINE = false, bool STATS = false, bool IN_BN = false,
          bool OUT_EPI = false>
static int launch_dw_tma_inst(const CUtensorMap& tm, const CUtensorMap& tmm, const CUtensorMap& tma,
                              const DwTmaParams& p, int grid, cudaStream_t st) {
    using Cfg = DwStageCfg<M_ACT, HAS_ADD>;
    constexpr int smem = Cfg::kStages * Cfg::kStageBytes + 128 + 64;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dw_conv_tma_kernel<IN_ACT, M_ACT, M_AFFINE, HAS_ADD, IN_AFFINE, STATS, IN_BN, OUT_EPI>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "cudaFuncSetAttribute(dw tma smem=%d): %s", smem, cudaGetErrorString(e));
        configured = true;
    }
    launch_pdl(dw_conv_tma_kernel<IN_ACT, M_ACT, M_AFFINE, HAS_ADD, IN_AFFINE, STATS, IN_BN, OUT_EPI>, dim3(grid), dim3(kDwThreads), smem, st, tm, tmm, tma, p);
    return check_launch("dwconv3x3 (tma)");
}

// Returns 1 if the TMA kernel took the launch, 0 if the caller must use the direct kernel, < 0 on error.
int launch_dw_conv_tma(const __nv_bfloat16* in, const float* w, __nv_bfloat16* out, int N, int Hin, int Win, int C,
                       int Hout, int Wout, int pad_t, int pad_l, int flip, int in_act, const __nv_bfloat16* mask_src,
                       const float* m_scale, const float* m_shift, int m_act, const __nv_bfloat16* addend,
                       cudaStream_t st, const float* in_scale, const float* in_shift, const float* bn_mean,
                       const float* bn_invstd, float* bn_red, const DwBnFold* fold, const float* out_scale,
                       const float* out_shift, int out_act) {
    if (get_encode_fn() == nullptr) return 0;
    if (mask_src == nullptr) m_act = DLV3P_ACT_NONE;
    const bool out_epi = (out_scale != nullptr);
    if (out_epi && (in_act != DLV3P_ACT_NONE || m_act != DLV3P_ACT_NONE || addend != nullptr || in_scale != nullptr ||
                    fold != nullptr || bn_red != nullptr || (C & 3)))
        return 0;                       // the output epilogue exists for the plain forward only
    if (in_act != DLV3P_ACT_NONE && (m_act != DLV3P_ACT_NONE || addend != nullptr)) return 0;   // not a used combination
    const bool in_aff = (in_scale != nullptr) || (fold != nullptr);
    if (in_aff && (in_act == DLV3P_ACT_NONE || (C & 3))) return 0;          // NaN padding needs a clamping activation
    const bool stats = (bn_red != nullptr);
    if (stats && (m_act == DLV3P_ACT_NONE || m_scale == nullptr || addend != nullptr || in_aff)) return 0;
    CUtensorMap tm, tmm, tma;
    int rc = make_tmap_nhwc(&tm, in, N, Hin, Win, C, kDwCB, kDwTW + 2, kDwTH + 2, /*nan_fill=*/in_aff);
    if (rc) return rc;
    tmm = tm; tma = tm;                          // placeholders when the operand is absent (never dereferenced)
    if (m_act != DLV3P_ACT_NONE) {
        rc = make_tmap_nhwc(&tmm, mask_src, N, Hout, Wout, C, kDwCB, kDwTW, kDwTH);
        if (rc) return rc;
    }
    if (addend != nullptr) {
        rc = make_tmap_nhwc(&tma, addend, N, Hout, Wout, C, kDwCB, kDwTW, kDwTH);
        if (rc) return rc;
    }
    DwTmaParams p;
    p.N = N; p.Hin = Hin; p.Win = Win; p.C = C; p.Hout = Hout; p.Wout = Wout; p.pad_t = pad_t; p.pad_l = pad_l;
    p.flip = flip; p.in_act = in_act; p.w = w; p.out = out; p.mask_src = mask_src; p.m_scale = m_scale;
    p.m_shift = m_shift; p.m_act = m_act; p.addend = addend;
    p.in_scale = in_scale; p.in_shift = in_shift; p.bn_mean = bn_mean; p.bn_invstd = bn_invstd; p.bn_red = bn_red;
    p.f_sums = nullptr;
    p.out_scale = out_scale; p.out_shift = out_shift; p.out_act = out_act;
    if (fold != nullptr) {
        p.f_sums = fold->sums; p.f_gamma = fold->gamma; p.f_beta = fold->beta; p.f_mm = fold->moving_mean;
        p.f_mv = fold->moving_var; p.f_scale = fold->scale; p.f_shift = fold->shift; p.f_mean = fold->mean;
        p.f_invstd = fold->invstd; p.f_count = fold->count; p.f_inv_count = 1.0 / fold->count; p.f_eps = fold->eps; p.f_momentum = fold->momentum;
        p.f_updates = fold->updates;
    }
    p.tiles_h = cdiv(Hout, kDwTH); p.tiles_w = cdiv(Wout, kDwTW); p.tiles_c = cdiv(C, kDwCB);
    const long long nt = (long long)N * p.tiles_h * p.tiles_w;
    if (nt > 0x7fffffffLL) return 0;
    p.spatial_tiles = (int)nt;
    int per = kNumSMs / p.tiles_c; if (per < 1) per = 1;
    if (per > p.spatial_tiles) per = p.spatial_tiles;
    p.ctas_per_cb = per;
    const int grid = p.tiles_c * per;
    const bool aff = (m_scale != nullptr);
    const bool add = (addend != nullptr);
#define DLV3P_DW(IA, MA, AF, AD) rc = launch_dw_tma_inst<IA, MA, AF, AD>(tm, tmm, tma, p, grid, st)
    if (out_epi) {
        rc = launch_dw_tma_inst<0, 0, false, false, false, false, false, true>(tm, tmm, tma, p, grid, st);
    } else if (in_aff && fold != nullptr) {
        if (in_act == DLV3P_ACT_RELU) rc = launch_dw_tma_inst<1, 0, false, false, true, false, true>(tm, tmm, tma, p, grid, st);
        else rc = launch_dw_tma_inst<2, 0, false, false, true, false, true>(tm, tmm, tma, p, grid, st);
    } else if (in_aff) {
        if (in_act == DLV3P_ACT_RELU) rc = launch_dw_tma_inst<1, 0, false, false, true, false>(tm, tmm, tma, p, grid, st);
        else rc = launch_dw_tma_inst<2, 0, false, false, true, false>(tm, tmm, tma, p, grid, st);
    } else if (stats) {
        if (m_act == DLV3P_ACT_RELU) rc = launch_dw_tma_inst<0, 1, true, false, false, true>(tm, tmm, tma, p, grid, st);
        else rc = launch_dw_tma_inst<0, 2, true, false, false, true>(tm, tmm, tma, p, grid, st);
    } else if (m_act == DLV3P_ACT_NONE && !add) {
        if (in_act == DLV3P_ACT_NONE) DLV3P_DW(0, 0, false, false);
        else if (in_act == DLV3P_ACT_RELU) DLV3P_DW(1, 0, false, false);
        else DLV3P_DW(2, 0, false, false);
    } else if (m_act == DLV3P_ACT_NONE) {
        DLV3P_DW(0, 0, false, true);
    } else if (m_act == DLV3P_ACT_RELU) {
        if (aff) { if (add) DLV3P_DW(0, 1, true, true); else DLV3P_DW(0, 1, true, false); }
        else { if (add) DLV3P_DW(0, 1, false, true); else DLV3P_DW(0, 1, false, false); }
    } else {
        if (aff) { if (add) DLV3P_DW(0, 2, true, true); else DLV3P_DW(0, 2, true, false); }
        else { if (add) DLV3P_DW(0, 2, false, true); else DLV3P_DW(0, 2, false, false); }
    }
#undef DLV3P_DW
    return rc ? rc : 1;
}

// ---- filter gradient (dense taps, bf16): dw[i][j][c] += sum_{n,ho,wo} act(x[n,ho-pt+i,wo-pl+j,c]) * dy[n,ho,wo,c] ----
// Same tiling as the forward kernel, two TMA boxes per tile (x with halo, dy without); OOB zero fill makes ragged
// tile edges, the convolution padding and the channel tail contribute exact zeros, so no masking is needed.
// Every CTA owns ONE 64-channel block (grid = channel blocks x CTAs per block) and keeps its 9 x 4 partial sums per
// thread in registers across all its spatial tiles; one shared-memory reduction over the 32 columns and one fp32 RED
// per (tap, channel) per CTA at the end.  The direct kernels (dwconv.cu) keep 16-22 independent 16-byte global loads
// in flight per thread at <= 16 warps/SM, i.e. ~6 KB/SM — an order of magnitude short of what HBM latency needs;
// here 2 x 76 KB of TMA boxes are in flight per SM while the third is reduced.
constexpr int kWgDyBytes = kDwTH * kDwTW * kDwCB * 2;                     // 32,768 B
constexpr int kWgStageBytes = kDwStageBytes + kWgDyBytes;                 // 76,288 B
constexpr int kWgStages = 3;

struct DwWgradParams {
    const float* in_scale; const float* in_shift;             // IN_AFFINE: x := act(in_scale*x + in_shift) on load
    int N, C, Ho, Wo, pad_t, pad_l, in_act;
    float* dw;                              // [3,3,C] fp32, accumulated
    int tiles_h, tiles_w, ctas_per_cb, spatial_tiles;
};

template <int IN_ACT, bool IN_AFFINE>
__global__ void __launch_bounds__(kDwThreads, 1)
dw_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy,
                    const DwWgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t stage0 = smem_u32(smem);
    const int tid = threadIdx.x;
    const int cq = tid & 15, col = tid >> 4;
    const int cb = blockIdx.x / p.ctas_per_cb;
    const int first = blockIdx.x % p.ctas_per_cb;
    const int gstride = p.ctas_per_cb;
    pdl_launch_dependents();

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_x)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_dy)) : "memory");
        for (int s = 0; s < kWgStages; ++s) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();

    auto issue = [&](int tile, int s) {
        const int tw = tile % p.tiles_w; int t = tile / p.tiles_w;
        const int th = t % p.tiles_h; const int n = t / p.tiles_h;
        const uint32_t dst = stage0 + s * kWgStageBytes;
        mbar_expect_tx(bar0 + 8 * s, kWgStageBytes);
        tma_load_4d(dst, &tm_x, bar0 + 8 * s, cb * kDwCB, tw * kDwTW - p.pad_l, th * kDwTH - p.pad_t, n);
        tma_load_4d(dst + kDwStageBytes, &tm_dy, bar0 + 8 * s, cb * kDwCB, tw * kDwTW, th * kDwTH, n);
    };
    int tile = first;
    if (tid == 0) {
        for (int a = 0; a < kWgStages - 1; ++a)
            if (tile + a * gstride < p.spatial_tiles) issue(tile + a * gstride, a);
    }

    float2 acc[9][2];
#pragma unroll
    for (int a = 0; a < 9; ++a) { acc[a][0] = make_float2(0.f, 0.f); acc[a][1] = make_float2(0.f, 0.f); }
    float2 isc[2], ish[2];
    if (IN_AFFINE) {
        // channels beyond C compute garbage from NaN-filled lanes; they are never written (ch < p.C below)
        const int c0 = min(cb * kDwCB + cq * 4, p.C - 4);
        const float4 a = __ldg(reinterpret_cast<const float4*>(p.in_scale + c0));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.in_shift + c0));
        isc[0] = make_float2(a.x, a.y); isc[1] = make_float2(a.z, a.w);
        ish[0] = make_float2(b.x, b.y); ish[1] = make_float2(b.z, b.w);
    }
    uint32_t it = 0;
    for (; tile < p.spatial_tiles; tile += gstride, ++it) {
        const int s = it % kWgStages;
        const int ahead = tile + (kWgStages - 1) * gstride;
        if (ahead < p.spatial_tiles && tid == 0) issue(ahead, (it + kWgStages - 1) % kWgStages);
        mbar_wait(bar0 + 8 * s, (it / kWgStages) & 1u);

        const uint32_t xbase = stage0 + s * kWgStageBytes + (col * kDwCB + cq * 4) * 2;
        const uint32_t gbase = stage0 + s * kWgStageBytes + kDwStageBytes + (col * kDwCB + cq * 4) * 2;
        auto load_row = [&](int row, float2 (&dst)[3][2]) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                uint2 raw;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                             : "=r"(raw.x), "=r"(raw.y)
                             : "r"(xbase + (uint32_t)((row * (kDwTW + 2) + j) * (kDwCB * 2))));
                if (IN_AFFINE) {
                    widen4_t<DLV3P_ACT_NONE>(raw, dst[j]);
                    affine_act4<IN_ACT == DLV3P_ACT_NONE ? DLV3P_ACT_RELU : IN_ACT>(dst[j], isc, ish);
                } else {
                    widen4_t<IN_ACT>(raw, dst[j]);
                }
            }
        };
        auto accum = [&](int r, const float2 (&ra)[3][2], const float2 (&rb)[3][2], const float2 (&rc)[3][2]) {
            uint2 raw;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                         : "=r"(raw.x), "=r"(raw.y) : "r"(gbase + (uint32_t)(r * kDwTW * kDwCB * 2)));
            float2 g[2];
            widen4_t<DLV3P_ACT_NONE>(raw, g);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    acc[0 * 3 + j][k] = __ffma2_rn(ra[j][k], g[k], acc[0 * 3 + j][k]);
                    acc[1 * 3 + j][k] = __ffma2_rn(rb[j][k], g[k], acc[1 * 3 + j][k]);
                    acc[2 * 3 + j][k] = __ffma2_rn(rc[j][k], g[k], acc[2 * 3 + j][k]);
                }
            }
        };
        float2 r0[3][2], r1[3][2], r2[3][2];
        load_row(0, r0);
        load_row(1, r1);
        static_assert(kDwTH == 8, "row loop below is unrolled for TH = 8");
        load_row(2, r2); accum(0, r0, r1, r2);
        load_row(3, r0); accum(1, r1, r2, r0);
        load_row(4, r1); accum(2, r2, r0, r1);
        load_row(5, r2); accum(3, r0, r1, r2);
        load_row(6, r0); accum(4, r1, r2, r0);
        load_row(7, r1); accum(5, r2, r0, r1);
        load_row(8, r2); accum(6, r0, r1, r2);
        load_row(9, r0); accum(7, r1, r2, r0);
        __syncthreads();
    }
    // every TMA box this CTA issued has been consumed: reuse stage 0 as the reduction buffer [9][32 cols][64 ch]
    float* red = reinterpret_cast<float*>(smem);
#pragma unroll
    for (int a = 0; a < 9; ++a) {
        float4 v = make_float4(acc[a][0].x, acc[a][0].y, acc[a][1].x, acc[a][1].y);
        *reinterpret_cast<float4*>(red + (a * kDwTW + col) * kDwCB + cq * 4) = v;
    }
    __syncthreads();
    for (int o = tid; o < 9 * kDwCB; o += kDwThreads) {
        const int a = o / kDwCB, c = o % kDwCB;
        const int ch = cb * kDwCB + c;
        if (ch < p.C) {
            float sum = 0.f;
#pragma unroll 8
            for (int q = 0; q < kDwTW; ++q) sum += red[(a * kDwTW + q) * kDwCB + c];
            atomicAdd(p.dw + a * p.C + ch, sum);
        }
    }
}

// Returns 1 if the TMA kernel took the launch, 0 if the caller must use the direct kernel, < 0 on error.
int launch_dw_wgrad_tma(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, int N, int H, int W, int C, int Ho,
                        int Wo, int pad_t, int pad_l, int in_act, cudaStream_t st, const float* in_scale,
                        const float* in_shift) {
    if (get_encode_fn() == nullptr) return 0;
    const bool in_aff = (in_scale != nullptr);
    if (in_aff && (in_act == DLV3P_ACT_NONE || (C & 3))) return 0;          // NaN padding needs a clamping activation
    CUtensorMap tmx, tmg;
    int rc = make_tmap_nhwc(&tmx, x, N, H, W, C, kDwCB, kDwTW + 2, kDwTH + 2, /*nan_fill=*/in_aff);
    if (rc) return rc;
    rc = make_tmap_nhwc(&tmg, dy, N, Ho, Wo, C, kDwCB, kDwTW, kDwTH);
    if (rc) return rc;
    DwWgradParams p;
    p.N = N; p.C = C; p.Ho = Ho; p.Wo = Wo; p.pad_t = pad_t; p.pad_l = pad_l; p.in_act = in_act; p.dw = dw;
    p.in_scale = in_scale; p.in_shift = in_shift;
    p.tiles_h = cdiv(Ho, kDwTH); p.tiles_w = cdiv(Wo, kDwTW);
    const int tiles_c = cdiv(C, kDwCB);
    const long long spatial = (long long)N * p.tiles_h * p.tiles_w;
    if (spatial > 0x7fffffffLL) return 0;
    p.spatial_tiles = (int)spatial;
    int per = kNumSMs / tiles_c; if (per < 1) per = 1;
    if (per > p.spatial_tiles) per = p.spatial_tiles;
    p.ctas_per_cb = per;
    constexpr int smem = kWgStages * kWgStageBytes + 128 + 64;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dw_wgrad_tma_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dw_wgrad_tma_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dw_wgrad_tma_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dw_wgrad_tma_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(dw_wgrad_tma_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "cudaFuncSetAttribute(dw wgrad tma smem=%d): %s", smem, cudaGetErrorString(e));
        configured = true;
    }
    if (in_aff && in_act == DLV3P_ACT_RELU) launch_pdl(dw_wgrad_tma_kernel<1, true>, dim3(tiles_c * per), dim3(kDwThreads), smem, st, tmx, tmg, p);
    else if (in_aff) launch_pdl(dw_wgrad_tma_kernel<2, true>, dim3(tiles_c * per), dim3(kDwThreads), smem, st, tmx, tmg, p);
    else if (in_act == DLV3P_ACT_NONE) launch_pdl(dw_wgrad_tma_kernel<0, false>, dim3(tiles_c * per), dim3(kDwThreads), smem, st, tmx, tmg, p);
    else if (in_act == DLV3P_ACT_RELU) launch_pdl(dw_wgrad_tma_kernel<1, false>, dim3(tiles_c * per), dim3(kDwThreads), smem, st, tmx, tmg, p);
    else launch_pdl(dw_wgrad_tma_kernel<2, false>, dim3(tiles_c * per), dim3(kDwThreads), smem, st, tmx, tmg, p);
    rc = check_launch("dwconv3x3_wgrad (tma)");
    return rc ? rc : 1;
}


// =====================================================================================================================
// Fused depthwise backward (stride 1, dilation 1, SAME, bf16): input gradient + filter gradient (+ the BN-backward
// reductions of the producing layer) in ONE pass over dy.
//
// dlv3p_dwconv3x3_dgrad and dlv3p_dwconv3x3_wgrad both stream the gradient dy of the depthwise output and the tensor the
// convolution read (as the activation mask in one, as the convolved operand in the other).  Written around the INPUT
// pixel (h, w) both need the same 3x3 window of dy:
//     dx[h,w]    = mask(x) * sum_{i,j} w[i][j] * dy[h+1-i, w+1-j]        (conv-transpose = flipped taps)
//     dwg[i][j] += x_act[h,w] * dy[h+1-i, w+1-j]                          (filter gradient, re-indexed from output to
//                                                                          input pixels; dy outside the image is zero)
// so one kernel slides the dy window (TMA halo box, zero fill) once, reads x from a halo-free box of the same stage,
// and keeps the 9 x 4 filter-gradient partial sums in registers next to the 9 x 4 taps.  That is 36 registers more than
// the input-gradient kernel: at 4 channels x 1 column per thread the kernel needs ~150 registers, so a CTA is
// 12 channel quads x 32 columns = 384 threads on 48-channel blocks (<= 168 registers, 12 resident warps) instead of
// 16 x 32 = 512 threads on 64-channel blocks.  Per middle-flow layer ([16,32,32,728]) this replaces a 17.5 us and a
// 15.3 us launch, each mostly launch ramp and tail on a 24 MB L2-resident tensor.
// =====================================================================================================================
template <int CQ, int TW, bool HAS_ADD, bool YBOX>
struct DwBwdCfg {
    static constexpr int kCB = CQ * 4;                                        // channels per block
    static constexpr int kThreads = CQ * TW;
    static constexpr int kHaloBytes = (kDwTH + 2) * (TW + 2) * kCB * 2;
    static constexpr int kCtrBytes = kDwTH * TW * kCB * 2;
    static constexpr int kStageBytes = kHaloBytes + (1 + (HAS_ADD ? 1 : 0) + (YBOX ? 1 : 0)) * kCtrBytes;
    static constexpr int kBudget = 226 * 1024;
    static constexpr int kStages = kBudget / kStageBytes >= 4 ? 4 : (kBudget / kStageBytes >= 3 ? 3 : 2);
    static constexpr int kXOff = kHaloBytes;
    static constexpr int kAddOff = kHaloBytes + kCtrBytes;
    static constexpr int kYOff = kHaloBytes + (HAS_ADD ? 2 : 1) * kCtrBytes;
    static constexpr int kRedWg = 9 * TW * kCB * 4;                        // filter-gradient reduction buffer
    static constexpr int kRedBn = 2 * TW * kCB * 4;
    static constexpr int kSmem = kStages * kStageBytes + 128 + 64;
    static_assert(kRedWg + kRedBn <= kStages * kStageBytes, "reduction buffers reuse the stage ring");
    static_assert(kHaloBytes % 128 == 0 && kCtrBytes % 128 == 0, "TMA destinations stay 128-byte aligned");
};

struct DwBwdParams {
    int N, H, W, C;
    const float* w;                         // [3,3,C] fp32
    float* dwg;                             // [3,3,C] fp32, accumulated
    __nv_bfloat16* dx;
    const float* x_scale; const float* x_shift;                    // X_AFFINE: x := act(x_scale*x_src + x_shift)
    const float* bn_mean; const float* bn_invstd; float* bn_red;   // STATS
    int tiles_h, tiles_w, spatial_tiles, ctas_per_cb;
};

// X_ACT: activation between x_src and the convolution (its derivative masks dx), X_AFFINE: per-channel affine map in
// front of it (the producing layer's training-mode BatchNormalization), HAS_ADD: gradient addend, STATS: BN reductions
// of the masked gradient against x_src (x_src = raw output of the BN's layer), YSTATS: BN reductions of the FINAL
// gradient (addend included) against a separate tensor bn_y — the layer whose BatchNormalization output (+ residual) was
// materialised and is read here through a pre-activation (the block-closing sepconv of an Xception block).
template <int X_ACT, bool X_AFFINE, bool HAS_ADD, bool STATS, bool YSTATS, int CQ, int TW>
__global__ void __launch_bounds__(CQ * TW, 1)
dw_bwd_tma_kernel(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x,
                  const __grid_constant__ CUtensorMap tm_add, const __grid_constant__ CUtensorMap tm_y,
                  const DwBwdParams p) {
    static_assert(!(STATS && YSTATS), "one set of reductions per launch");
    using Cfg = DwBwdCfg<CQ, TW, HAS_ADD, YSTATS>;
    constexpr int kCB = Cfg::kCB, kStages = Cfg::kStages, kStageBytes = Cfg::kStageBytes;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t stage0 = smem_u32(smem);

    const int tid = threadIdx.x;
    const int cq = tid % CQ;                // channel quad inside the block
    const int col = tid / CQ;               // column inside the tile, 0..TW-1
    const int cb = blockIdx.x / p.ctas_per_cb;
    const int gstride = p.ctas_per_cb;
    const int c0 = cb * kCB + cq * 4;
    pdl_launch_dependents();

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_dy)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_x)) : "memory");
        if (HAS_ADD) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_add)) : "memory");
        if (YSTATS) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_y)) : "memory");
        for (int s = 0; s < kStages; ++s) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // taps in window order: window position a = (i', j') meets tap 8 - a.  Master weights / BN vectors only change in
    // kernels that do not trigger dependent launches, so these loads run ahead of the dependency wait
    float2 wgt[9][2];
    const bool ch_ok = c0 < p.C;
    float msc[4] = {1.f, 1.f, 1.f, 1.f}, msh[4] = {0.f, 0.f, 0.f, 0.f}, mu[4] = {0.f, 0.f, 0.f, 0.f};
    if (ch_ok) {
#pragma unroll
        for (int a = 0; a < 9; ++a) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.w + (8 - a) * p.C + c0));
            wgt[a][0] = make_float2(v.x, v.y); wgt[a][1] = make_float2(v.z, v.w);
        }
    }
    pdl_wait();
    if (ch_ok) {
        if (X_AFFINE) {
            const float4 a = __ldcg(reinterpret_cast<const float4*>(p.x_scale + c0));
            const float4 b = __ldcg(reinterpret_cast<const float4*>(p.x_shift + c0));
            msc[0] = a.x; msc[1] = a.y; msc[2] = a.z; msc[3] = a.w;
            msh[0] = b.x; msh[1] = b.y; msh[2] = b.z; msh[3] = b.w;
        }
        if (STATS || YSTATS) {
            const float4 m = __ldcg(reinterpret_cast<const float4*>(p.bn_mean + c0));
            mu[0] = m.x; mu[1] = m.y; mu[2] = m.z; mu[3] = m.w;
        }
    }

    auto decode = [&](int tile, int& n, int& th, int& tw) {
        tw = tile % p.tiles_w; const int t = tile / p.tiles_w;
        th = t % p.tiles_h; n = t / p.tiles_h;
    };
    auto issue = [&](int tile, int s) {
        int n, th, tw;
        decode(tile, n, th, tw);
        mbar_expect_tx(bar0 + 8 * s, kStageBytes);
        const uint32_t dst = stage0 + s * kStageBytes;
        tma_load_4d(dst, &tm_dy, bar0 + 8 * s, cb * kCB, tw * TW - 1, th * kDwTH - 1, n);
        tma_load_4d(dst + Cfg::kXOff, &tm_x, bar0 + 8 * s, cb * kCB, tw * TW, th * kDwTH, n);
        if (HAS_ADD) tma_load_4d(dst + Cfg::kAddOff, &tm_add, bar0 + 8 * s, cb * kCB, tw * TW, th * kDwTH, n);
        if (YSTATS) tma_load_4d(dst + Cfg::kYOff, &tm_y, bar0 + 8 * s, cb * kCB, tw * TW, th * kDwTH, n);
    };

    int tile = blockIdx.x % p.ctas_per_cb;
    if (tid == 0) {
        for (int a = 0; a < kStages - 1; ++a)
            if (tile + a * gstride < p.spatial_tiles) issue(tile + a * gstride, a);
    }

    float2 acc9[9][2];
#pragma unroll
    for (int a = 0; a < 9; ++a) { acc9[a][0] = make_float2(0.f, 0.f); acc9[a][1] = make_float2(0.f, 0.f); }
    float bs1[4] = {0.f, 0.f, 0.f, 0.f}, bs2[4] = {0.f, 0.f, 0.f, 0.f};
    const long long row_stride = (long long)p.W * p.C;

    uint32_t it = 0;
    for (; tile < p.spatial_tiles; tile += gstride, ++it) {
        const int s = it % kStages;
        const int ahead = tile + (kStages - 1) * gstride;
        if (ahead < p.spatial_tiles && tid == 0) issue(ahead, (it + kStages - 1) % kStages);

        int n, th, tw;
        decode(tile, n, th, tw);
        const int wo = tw * TW + col;
        const bool lane_ok = ch_ok && (wo < p.W);

        mbar_wait(bar0 + 8 * s, (it / kStages) & 1u);

        if (lane_ok) {
            // stage layout: dy [TH+2][TW+2][CB], then x_src [TH][TW][CB] (, addend [TH][TW][CB]) (, bn_y [TH][TW][CB])
            const uint32_t base = stage0 + s * kStageBytes + (col * kCB + cq * 4) * 2;
            auto load_row = [&](int row, float2 (&dst)[3][2]) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    uint2 raw;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                                 : "=r"(raw.x), "=r"(raw.y)
                                 : "r"(base + (uint32_t)((row * (TW + 2) + j) * (kCB * 2))));
                    widen4_t<DLV3P_ACT_NONE>(raw, dst[j]);
                }
            };
            auto lds_ctr = [&](int byte_off, int r) {
                uint2 raw;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
                             : "=r"(raw.x), "=r"(raw.y)
                             : "r"(base + (uint32_t)(byte_off + r * TW * kCB * 2)));
                return raw;
            };
            long long off = (((long long)n * p.H + th * kDwTH) * p.W + wo) * p.C + c0;
            const int rows_valid = p.H - th * kDwTH;
            auto emit = [&](int r, const float2 (&ra)[3][2], const float2 (&rb)[3][2], const float2 (&rc)[3][2]) {
                if (r < rows_valid) {                    // uniform across the CTA
                    // one accumulation chain per channel pair: this kernel is bound by the FP32 pipe (a packed FMA with three
                    // distinct register pairs occupies it for 3 cycles, scripts/ubench/fma_rate.cu), so the two extra
                    // packed adds of a three-chain sum cost more than the chain latency the 18 independent filter-gradient
                    // FMAs below hide anyway
                    float2 acc[2];
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        float2 a0 = __fmul2_rn(ra[0][k], wgt[0][k]);
                        a0 = __ffma2_rn(ra[1][k], wgt[1][k], a0);
                        a0 = __ffma2_rn(ra[2][k], wgt[2][k], a0);
                        a0 = __ffma2_rn(rb[0][k], wgt[3][k], a0);
                        a0 = __ffma2_rn(rb[1][k], wgt[4][k], a0);
                        a0 = __ffma2_rn(rb[2][k], wgt[5][k], a0);
                        a0 = __ffma2_rn(rc[0][k], wgt[6][k], a0);
                        a0 = __ffma2_rn(rc[1][k], wgt[7][k], a0);
                        acc[k] = __ffma2_rn(rc[2][k], wgt[8][k], a0);
                    }
                    float f[4] = {acc[0].x, acc[0].y, acc[1].x, acc[1].y};
                    const uint2 xraw = lds_ctr(Cfg::kXOff, r);
                    float2 xf[2];
                    widen4_t<DLV3P_ACT_NONE>(xraw, xf);
                    const float u[4] = {xf[0].x, xf[0].y, xf[1].x, xf[1].y};
                    float xv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float v = X_AFFINE ? fmaf(u[k], msc[k], msh[k]) : u[k];
                        if (X_ACT == DLV3P_ACT_RELU) {
                            f[k] = v > 0.f ? f[k] : 0.f;
                            xv[k] = fmaxf(v, 0.f);
                        } else if (X_ACT == DLV3P_ACT_RELU6) {
                            f[k] = (v > 0.f && v < 6.f) ? f[k] : 0.f;
                            xv[k] = fminf(fmaxf(v, 0.f), 6.f);
                        } else {
                            xv[k] = v;
                        }
                        // centred second sum: sum g*(y - mean) (the form dlv3p_bn_bwd_reduce uses; sum g*y - mean*sum g
                        // cancels to 1/(|mean|/std) of its terms)
                        if (STATS) { bs1[k] += f[k]; bs2[k] = fmaf(f[k], u[k] - mu[k], bs2[k]); }
                    }
                    const float2 x2[2] = {make_float2(xv[0], xv[1]), make_float2(xv[2], xv[3])};
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            acc9[0 * 3 + j][k] = __ffma2_rn(ra[j][k], x2[k], acc9[0 * 3 + j][k]);
                            acc9[1 * 3 + j][k] = __ffma2_rn(rb[j][k], x2[k], acc9[1 * 3 + j][k]);
                            acc9[2 * 3 + j][k] = __ffma2_rn(rc[j][k], x2[k], acc9[2 * 3 + j][k]);
                        }
                    }
                    if (HAS_ADD) {
                        const uint2 araw = lds_ctr(Cfg::kAddOff, r);
                        float2 af[2];
                        widen4_t<DLV3P_ACT_NONE>(araw, af);
                        f[0] += af[0].x; f[1] += af[0].y; f[2] += af[1].x; f[3] += af[1].y;
                    }
                    if (YSTATS) {
                        const uint2 yraw = lds_ctr(Cfg::kYOff, r);
                        float2 yf[2];
                        widen4_t<DLV3P_ACT_NONE>(yraw, yf);
                        const float yv[4] = {yf[0].x, yf[0].y, yf[1].x, yf[1].y};
#pragma unroll
                        for (int k = 0; k < 4; ++k) { bs1[k] += f[k]; bs2[k] = fmaf(f[k], yv[k] - mu[k], bs2[k]); }
                    }
                    uint2 o;
                    __nv_bfloat162 lo = __floats2bfloat162_rn(f[0], f[1]), hi = __floats2bfloat162_rn(f[2], f[3]);
                    o.x = *reinterpret_cast<uint32_t*>(&lo);
                    o.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(p.dx + off) = o;
                }
                off += row_stride;
            };
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
        __syncthreads();
    }
    // every TMA box this CTA issued has been consumed: the ring becomes [9][TW cols][CB] (+ [2][TW][CB]) fp32
    float* red = reinterpret_cast<float*>(smem);
    float* redbn = reinterpret_cast<float*>(smem + Cfg::kRedWg);
#pragma unroll
    for (int a = 0; a < 9; ++a)
        *reinterpret_cast<float4*>(red + (a * TW + col) * kCB + cq * 4) =
            make_float4(acc9[a][0].x, acc9[a][0].y, acc9[a][1].x, acc9[a][1].y);
    if (STATS || YSTATS) {
        *reinterpret_cast<float4*>(redbn + col * kCB + cq * 4) = make_float4(bs1[0], bs1[1], bs1[2], bs1[3]);
        *reinterpret_cast<float4*>(redbn + (TW + col) * kCB + cq * 4) = make_float4(bs2[0], bs2[1], bs2[2], bs2[3]);
    }
    __syncthreads();
    for (int o = tid; o < 9 * kCB; o += Cfg::kThreads) {
        const int a = o / kCB, c = o % kCB;
        const int ch = cb * kCB + c;
        if (ch < p.C) {
            float sum = 0.f;
#pragma unroll 8
            for (int q = 0; q < TW; ++q) sum += red[(a * TW + q) * kCB + c];
            atomicAdd(p.dwg + (8 - a) * p.C + ch, sum);
        }
    }
    if (STATS || YSTATS) {
        // red[0..C) += sum g, red[C..2C) += sum g*xhat, g = the gradient written above (dlv3p_bn_bwd_reduce)
        for (int o = tid; o < 2 * kCB; o += Cfg::kThreads) {
            const int q2 = o / kCB, c = o % kCB;
            const int ch = cb * kCB + c;
            if (ch < p.C) {
                float sum = 0.f;
#pragma unroll 8
                for (int q = 0; q < TW; ++q) sum += redbn[(q2 * TW + q) * kCB + c];
                if (q2 == 1) sum *= __ldg(p.bn_invstd + ch);
                atomicAdd(p.bn_red + q2 * p.C + ch, sum);
            }
        }
    }
}

template <int X_ACT, bool X_AFFINE, bool HAS_ADD, bool STATS, bool YSTATS, int CQ, int TW>
static int launch_dw_bwd_inst(const CUtensorMap& tmd, const CUtensorMap& tmx, const CUtensorMap& tma, const CUtensorMap& tmy,
                              const DwBwdParams& p, int grid, cudaStream_t st) {
    using Cfg = DwBwdCfg<CQ, TW, HAS_ADD, YSTATS>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dw_bwd_tma_kernel<X_ACT, X_AFFINE, HAS_ADD, STATS, YSTATS, CQ, TW>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
        DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "cudaFuncSetAttribute(dw bwd smem=%d): %s", Cfg::kSmem, cudaGetErrorString(e));
        configured = true;
    }
    launch_pdl(dw_bwd_tma_kernel<X_ACT, X_AFFINE, HAS_ADD, STATS, YSTATS, CQ, TW>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmem, st,
               tmd, tmx, tma, tmy, p);
    return check_launch("dwconv3x3_bwd (tma)");
}

// Two CTA geometries, both 384 threads: 12 channel quads x 32 columns (48-channel blocks) and 16 x 24 (64-channel blocks).
// The launcher takes the one that wastes fewer lanes on the ragged channel / width tails: 728 channels on 32-wide maps
// fill 48 x 32 tiles to 95 %, 64 / 128 channels on 254-wide maps fill 64 x 24 tiles to 96 % (48 x 32: 67 % / 89 %).
template <int CQ, int TW>
static int launch_dw_bwd_geom(const __nv_bfloat16* dy, const __nv_bfloat16* x_src, const float* w, __nv_bfloat16* dx,
                              float* dwg, int N, int H, int W, int C, int x_act, const float* x_scale,
                              const float* x_shift, const __nv_bfloat16* addend, const float* bn_mean,
                              const float* bn_invstd, float* bn_red, const __nv_bfloat16* bn_y, cudaStream_t st) {
    const bool aff = (x_scale != nullptr), add = (addend != nullptr), ystats = (bn_y != nullptr);
    const bool stats = (bn_red != nullptr) && !ystats;
    constexpr int kCB = CQ * 4;
    CUtensorMap tmd, tmx, tma, tmy;
    int rc = make_tmap_nhwc(&tmd, dy, N, H, W, C, kCB, TW + 2, kDwTH + 2);
    if (rc) return rc;
    rc = make_tmap_nhwc(&tmx, x_src, N, H, W, C, kCB, TW, kDwTH);
    if (rc) return rc;
    tma = tmx; tmy = tmx;
    if (add) {
        rc = make_tmap_nhwc(&tma, addend, N, H, W, C, kCB, TW, kDwTH);
        if (rc) return rc;
    }
    if (ystats) {
        rc = make_tmap_nhwc(&tmy, bn_y, N, H, W, C, kCB, TW, kDwTH);
        if (rc) return rc;
    }
    DwBwdParams p;
    p.N = N; p.H = H; p.W = W; p.C = C; p.w = w; p.dwg = dwg; p.dx = dx; p.x_scale = x_scale; p.x_shift = x_shift;
    p.bn_mean = bn_mean; p.bn_invstd = bn_invstd; p.bn_red = bn_red;
    p.tiles_h = cdiv(H, kDwTH); p.tiles_w = cdiv(W, TW);
    const int tiles_c = cdiv(C, kCB);
    const long long nt = (long long)N * p.tiles_h * p.tiles_w;
    if (nt > 0x7fffffffLL) return 0;
    p.spatial_tiles = (int)nt;
    int per = kNumSMs / tiles_c; if (per < 1) per = 1;
    if (per > p.spatial_tiles) per = p.spatial_tiles;
    p.ctas_per_cb = per;
    const int grid = tiles_c * per;
#define DLV3P_DWB(XA, AF, AD, ST) rc = launch_dw_bwd_inst<XA, AF, AD, ST, false, CQ, TW>(tmd, tmx, tma, tmy, p, grid, st)
    if (ystats) {
        if (add) rc = launch_dw_bwd_inst<1, false, true, false, true, CQ, TW>(tmd, tmx, tma, tmy, p, grid, st);
        else rc = launch_dw_bwd_inst<1, false, false, false, true, CQ, TW>(tmd, tmx, tma, tmy, p, grid, st);
    } else
    if (x_act == DLV3P_ACT_NONE) { if (add) DLV3P_DWB(0, false, true, false); else DLV3P_DWB(0, false, false, false); }
    else if (x_act == DLV3P_ACT_RELU) {
        if (stats) DLV3P_DWB(1, true, false, true);
        else if (aff) { if (add) DLV3P_DWB(1, true, true, false); else DLV3P_DWB(1, true, false, false); }
        else { if (add) DLV3P_DWB(1, false, true, false); else DLV3P_DWB(1, false, false, false); }
    } else {
        if (stats) DLV3P_DWB(2, true, false, true);
        else if (aff) { if (add) DLV3P_DWB(2, true, true, false); else DLV3P_DWB(2, true, false, false); }
        else { if (add) DLV3P_DWB(2, false, true, false); else DLV3P_DWB(2, false, false, false); }
    }
#undef DLV3P_DWB
    return rc ? rc : 1;
}

// Returns 1 if the fused kernel took the launch, 0 if the combination is not served (caller reports), < 0 on error.
int launch_dw_bwd_tma(const __nv_bfloat16* dy, const __nv_bfloat16* x_src, const float* w, __nv_bfloat16* dx, float* dwg,
                      int N, int H, int W, int C, int x_act, const float* x_scale, const float* x_shift,
                      const __nv_bfloat16* addend, const float* bn_mean, const float* bn_invstd, float* bn_red,
                      const __nv_bfloat16* bn_y, cudaStream_t st) {
    if (get_encode_fn() == nullptr || (C & 3)) return 0;
    const bool aff = (x_scale != nullptr), add = (addend != nullptr), stats = (bn_red != nullptr);
    if (x_act == DLV3P_ACT_NONE && (aff || stats)) return 0;
    if (bn_y != nullptr && (!stats || aff || x_act != DLV3P_ACT_RELU)) return 0;      // pre-activation ReLU readers only
    if (stats && bn_y == nullptr && (!aff || add)) return 0;
    auto fill = [&](int cb, int tw) {
        return ((double)C / (cdiv(C, cb) * cb)) * ((double)W / (cdiv(W, tw) * tw));
    };
    if (fill(64, 24) > fill(48, 32) + 1e-9)
        return launch_dw_bwd_geom<16, 24>(dy, x_src, w, dx, dwg, N, H, W, C, x_act, x_scale, x_shift, addend, bn_mean,
                                          bn_invstd, bn_red, bn_y, st);
    return launch_dw_bwd_geom<12, 32>(dy, x_src, w, dx, dwg, N, H, W, C, x_act, x_scale, x_shift, addend, bn_mean,
                                      bn_invstd, bn_red, bn_y, st);
}

// =====================================================================================================================
// Atrous depthwise 3x3 on small feature maps (the ASPP branches: rates 6 / 12 / 18 on [N,32,32,256], ss.py:823-830).
// With dilation d a tile of outputs needs inputs from a (TH+2d) x (TW+2d) window — for d >= 6 on a 32 x 32 map that
// is the whole image — so ONE TMA box {64 ch, W, H} stages the complete image of one (n, channel block) in shared memory
// (128 KB at 32 x 32) and the nine dilated taps become nine predicated LDS.64 per output (rows / columns outside the
// image are the convolution's zero padding).  Each (n, channel block) unit is split over `splits` CTAs by output rows so
// that the 64 units of a 256-channel tensor still fill the machine.  MODE 0: forward / input gradient (flipped taps,
// optional addend); MODE 1: filter gradient (dy read straight from global memory, one 8-byte load per output).
// The direct kernel it replaces issued 9 global loads per output and sat at 0.4-1.2 TB/s on these L2-resident tensors.
// =====================================================================================================================
struct DwImgParams {
    int N, H, W, C, dh, dw, flip, splits, tiles_c;
    const float* w;                         // [3,3,C] fp32
    __nv_bfloat16* out;                     // MODE 0
    const __nv_bfloat16* addend;            // MODE 0, optional
    const __nv_bfloat16* dy;                // MODE 1
    float* dwg;                             // MODE 1: [3,3,C] fp32, accumulated
};

template <int MODE, int IN_ACT, bool HAS_ADD>
__global__ void __launch_bounds__(kDwThreads, 1)
dw_image_kernel(const __grid_constant__ CUtensorMap tm_in, const DwImgParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    const int img_bytes = p.H * p.W * kDwCB * 2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ((img_bytes + 127) & ~127));
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t img0 = smem_u32(smem);
    const int tid = threadIdx.x;
    const int cq = tid & 15, col = tid >> 4;
    const int unit = blockIdx.x / p.splits, part = blockIdx.x % p.splits;
    const int cb = unit % p.tiles_c, n = unit / p.tiles_c;
    const int c0 = cb * kDwCB + cq * 4;
    pdl_launch_dependents();
    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_in)) : "memory");
        mbar_init(bar0, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool lane_ok = (c0 < p.C) && (col < p.W);
    float2 wgt[9][2];
    if (MODE == 0 && lane_ok) {
#pragma unroll
        for (int a = 0; a < 9; ++a) {
            const int tap = p.flip ? (8 - a) : a;
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.w + tap * p.C + c0));
            wgt[a][0] = make_float2(v.x, v.y); wgt[a][1] = make_float2(v.z, v.w);
        }
    }
    pdl_wait();
    if (tid == 0) {
        mbar_expect_tx(bar0, img_bytes);
        tma_load_4d(img0, &tm_in, bar0, cb * kDwCB, 0, 0, n);
    }
    const int rows_per = (p.H + p.splits - 1) / p.splits;
    const int r_begin = part * rows_per, r_end = min(p.H, r_begin + rows_per);
    float2 acc9[9][2];
    if (MODE == 1) {
#pragma unroll
        for (int a = 0; a < 9; ++a) { acc9[a][0] = make_float2(0.f, 0.f); acc9[a][1] = make_float2(0.f, 0.f); }
    }
    mbar_wait(bar0, 0);
    if (lane_ok) {
        // column offsets of the three tap columns inside an image row (bytes), -1 = outside the image
        int coff[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int cc = col + (j - 1) * p.dw;
            coff[j] = (cc >= 0 && cc < p.W) ? (cc * kDwCB + cq * 4) * 2 : -1;
        }
        const int row_bytes = p.W * kDwCB * 2;
        for (int r = r_begin; r < r_end; ++r) {
            float2 g[2];
            if (MODE == 1) {
                const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p.dy + (((long long)n * p.H + r) * p.W + col) * p.C + c0));
                widen4_t<DLV3P_ACT_NONE>(raw, g);
            }
            float2 o[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int rr = r + (i - 1) * p.dh;
                if (rr < 0 || rr >= p.H) continue;                      // uniform across the CTA
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    if (coff[j] < 0) continue;
                    uint2 raw;
                    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(raw.x), "=r"(raw.y)
                                 : "r"(img0 + (uint32_t)(rr * row_bytes + coff[j])));
                    float2 x[2];
                    widen4_t<IN_ACT>(raw, x);
                    if (MODE == 0) {
                        o[0] = __ffma2_rn(x[0], wgt[i * 3 + j][0], o[0]);
                        o[1] = __ffma2_rn(x[1], wgt[i * 3 + j][1], o[1]);
                    } else {
                        acc9[i * 3 + j][0] = __ffma2_rn(x[0], g[0], acc9[i * 3 + j][0]);
                        acc9[i * 3 + j][1] = __ffma2_rn(x[1], g[1], acc9[i * 3 + j][1]);
                    }
                }
            }
            if (MODE == 0) {
                const long long off = (((long long)n * p.H + r) * p.W + col) * p.C + c0;
                float f[4] = {o[0].x, o[0].y, o[1].x, o[1].y};
                if (HAS_ADD) {
                    const uint2 araw = __ldg(reinterpret_cast<const uint2*>(p.addend + off));
                    float2 af[2];
                    widen4_t<DLV3P_ACT_NONE>(araw, af);
                    f[0] += af[0].x; f[1] += af[0].y; f[2] += af[1].x; f[3] += af[1].y;
                }
                uint2 ov;
                __nv_bfloat162 lo = __floats2bfloat162_rn(f[0], f[1]), hi = __floats2bfloat162_rn(f[2], f[3]);
                ov.x = *reinterpret_cast<uint32_t*>(&lo);
                ov.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(p.out + off) = ov;
            }
        }
    }
    if (MODE == 1) {
        __syncthreads();                    // everyone is done with the staged image: reuse it as [9][32 cols][64 ch]
        float* red = reinterpret_cast<float*>(smem);
#pragma unroll
        for (int a = 0; a < 9; ++a)
            *reinterpret_cast<float4*>(red + (a * kDwTW + col) * kDwCB + cq * 4) =
                lane_ok ? make_float4(acc9[a][0].x, acc9[a][0].y, acc9[a][1].x, acc9[a][1].y) : make_float4(0.f, 0.f, 0.f, 0.f);
        __syncthreads();
        for (int o = tid; o < 9 * kDwCB; o += kDwThreads) {
            const int a = o / kDwCB, c = o % kDwCB;
            const int ch = cb * kDwCB + c;
            if (ch < p.C) {
                float sum = 0.f;
#pragma unroll 8
                for (int q = 0; q < kDwTW; ++q) sum += red[(a * kDwTW + q) * kDwCB + c];
                atomicAdd(p.dwg + a * p.C + ch, sum);
            }
        }
    }
}

// Returns 1 if it took the launch, 0 if the caller must use another kernel, < 0 on error.
// mode 0: out = conv(in) (flip = input gradient) (+ addend); mode 1: dwg += in-windows * dy.
int launch_dw_image(int mode, const __nv_bfloat16* in, const float* w, __nv_bfloat16* out, const __nv_bfloat16* dy,
                    float* dwg, int N, int H, int W, int C, int dil_h, int dil_w, int flip, int in_act,
                    const __nv_bfloat16* addend, cudaStream_t st) {
    if (get_encode_fn() == nullptr) return 0;
    const long long img_bytes = (long long)H * W * kDwCB * 2;
    const long long red_bytes = 9LL * kDwTW * kDwCB * 4;
    if (W > kDwTW || (C & 3) || img_bytes > 200 * 1024 || (mode == 1 && img_bytes < red_bytes)) return 0;
    if (mode == 0 && in_act != DLV3P_ACT_NONE && addend != nullptr) return 0;
    CUtensorMap tm;
    int rc = make_tmap_nhwc(&tm, in, N, H, W, C, kDwCB, W, H);
    if (rc) return rc;
    DwImgParams p;
    p.N = N; p.H = H; p.W = W; p.C = C; p.dh = dil_h; p.dw = dil_w; p.flip = flip; p.w = w; p.out = out;
    p.addend = addend; p.dy = dy; p.dwg = dwg;
    p.tiles_c = cdiv(C, kDwCB);
    const int units = N * p.tiles_c;
    int splits = kNumSMs / units; if (splits < 1) splits = 1; if (splits > 4) splits = 4; if (splits > H) splits = H;
    p.splits = splits;
    const int smem = (int)((img_bytes + 127) & ~127LL) + 128 + 64;
    const dim3 grid(units * splits), block(kDwThreads);
#define DLV3P_DWI(MODE, IA, AD)                                                                                      \
    do {                                                                                                             \
        static int configured = 0;                                                                                   \
        if (smem > configured) {                                                                                     \
            cudaError_t e = cudaFuncSetAttribute(dw_image_kernel<MODE, IA, AD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
            DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "cudaFuncSetAttribute(dw image smem=%d): %s", smem, cudaGetErrorString(e)); \
            configured = smem;                                                                                       \
        }                                                                                                            \
        launch_pdl(dw_image_kernel<MODE, IA, AD>, grid, block, smem, st, tm, p);                                     \
    } while (0)
    if (mode == 1) {
        if (in_act == DLV3P_ACT_NONE) DLV3P_DWI(1, 0, false);
        else if (in_act == DLV3P_ACT_RELU) DLV3P_DWI(1, 1, false);
        else DLV3P_DWI(1, 2, false);
    } else if (addend != nullptr) {
        DLV3P_DWI(0, 0, true);
    } else {
        if (in_act == DLV3P_ACT_NONE) DLV3P_DWI(0, 0, false);
        else if (in_act == DLV3P_ACT_RELU) DLV3P_DWI(0, 1, false);
        else DLV3P_DWI(0, 2, false);
    }
#undef DLV3P_DWI
    rc = check_launch("dwconv3x3 (whole-image atrous)");
    return rc ? rc : 1;
}

}  // namespace dlv3p
