// K1 — atrous depthwise 3x3 convolution, forward / input-gradient / filter-gradient, NHWC, sm_100a.
//
// Replaces TF DepthwiseConv2dNative(+BackpropInput/+BackpropFilter) as reached from the depthwise half of
// tf.keras SeparableConv2D (reference call site ss.py:823-830; keras.applications Xception blockN_sepconvM and
// MobileNetV2 block_k_depthwise).  For dilation > 1 TF 2.4 wraps the op in SpaceToBatchND/BatchToSpaceND — two
// extra full-tensor copies; here dilation is just an address stride.
//
// HBM-bound: algorithmic bytes per launch = (N*H*W*C + N*Ho*Wo*C)*esz + 9*C*4.  Design: one thread owns one
// 8-channel vector (16 B of bf16, coalesced over the contiguous NHWC channel axis) and a strip of TW output
// columns; for dense taps (stride 1, dil_w 1) the strip is a sliding window so each output costs (TW+2)*3/TW
// vector loads from L1 instead of 9, which keeps the L1 wavefront rate under the HBM rate.
// The optional prologue (BN scale/shift + ReLU/ReLU6 applied to the loaded value, zero padding afterwards)
// and epilogue (activation-derivative mask + addend) fuse the neighbouring elementwise layers.
#include "common.cuh"

namespace dlv3p {

// dwconv_tma.cu: TMA halo-staged kernel for the dense-tap bf16 case; returns 1 if it took the launch
int launch_dw_conv_tma(const __nv_bfloat16* in, const float* w, __nv_bfloat16* out, int N, int Hin, int Win, int C,
                       int Hout, int Wout, int pad_t, int pad_l, int flip, int in_act, const __nv_bfloat16* mask_src,
                       const float* m_scale, const float* m_shift, int m_act, const __nv_bfloat16* addend,
                       cudaStream_t st, const float* in_scale = nullptr, const float* in_shift = nullptr,
                       const float* bn_mean = nullptr, const float* bn_invstd = nullptr, float* bn_red = nullptr,
                       const DwBnFold* fold = nullptr, const float* out_scale = nullptr,
                       const float* out_shift = nullptr, int out_act = 0);

int launch_dw_wgrad_tma(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, int N, int H, int W, int C, int Ho,
                        int Wo, int pad_t, int pad_l, int in_act, cudaStream_t st, const float* in_scale = nullptr,
                        const float* in_shift = nullptr);

// dwconv_tma.cu: whole-image smem staging for atrous taps on small feature maps; returns 1 if it took the launch
int launch_dw_bwd_tma(const __nv_bfloat16* dy, const __nv_bfloat16* x_src, const float* w, __nv_bfloat16* dx, float* dwg,
                      int N, int H, int W, int C, int x_act, const float* x_scale, const float* x_shift,
                      const __nv_bfloat16* addend, const float* bn_mean, const float* bn_invstd, float* bn_red,
                      const __nv_bfloat16* bn_y, cudaStream_t st);
int launch_dw_image(int mode, const __nv_bfloat16* in, const float* w, __nv_bfloat16* out, const __nv_bfloat16* dy,
                    float* dwg, int N, int H, int W, int C, int dil_h, int dil_w, int flip, int in_act,
                    const __nv_bfloat16* addend, cudaStream_t st);

template <typename T, int TW, bool DENSE_W, bool HAS_AFFINE, bool HAS_EPI>
__global__ void __launch_bounds__(256)
dw_conv_kernel(const T* __restrict__ in, const float* __restrict__ w, T* __restrict__ out, int N, int Hin, int Win,
               int C, int Hout, int Wout, int stride, int dil_h, int dil_w, int pad_t, int pad_l, int flip,
               const float* __restrict__ in_scale, const float* __restrict__ in_shift, int in_act,
               const T* __restrict__ mask_src, const float* __restrict__ m_scale, const float* __restrict__ m_shift,
               int m_act, const T* __restrict__ addend, long long total) {
    const int CV = C >> 3;
    const int WS = (Wout + TW - 1) / TW;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long t = idx / CV;
    const int ws = (int)(t % WS); t /= WS;
    const int ho = (int)(t % Hout);
    const int n = (int)(t / Hout);
    const int c0 = cv << 3;
    const int wo0 = ws * TW;

    float sc[8], sh[8];
    if (HAS_AFFINE) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { sc[k] = __ldg(in_scale + c0 + k); sh[k] = __ldg(in_shift + c0 + k); }
    }

    float acc[TW][8];
#pragma unroll
    for (int o = 0; o < TW; ++o)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[o][k] = 0.f;

    const T* in_n = in + (long long)n * Hin * Win * C + c0;
    constexpr int NCOL = DENSE_W ? (TW + 2) : (3 * TW);

#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int hi = ho * stride - pad_t + i * dil_h;
        if (hi < 0 || hi >= Hin) continue;
        const T* in_row = in_n + (long long)hi * Win * C;
        // weights of this filter row (flipped for the stride-1 input gradient)
        float wr[3][8];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int tap = flip ? ((2 - i) * 3 + (2 - j)) : (i * 3 + j);
            const float4 a = __ldg(reinterpret_cast<const float4*>(w + tap * C + c0));
            const float4 b = __ldg(reinterpret_cast<const float4*>(w + tap * C + c0) + 1);
            wr[j][0] = a.x; wr[j][1] = a.y; wr[j][2] = a.z; wr[j][3] = a.w;
            wr[j][4] = b.x; wr[j][5] = b.y; wr[j][6] = b.z; wr[j][7] = b.w;
        }
        Vec8<T> v[NCOL];
        bool ok[NCOL];
#pragma unroll
        for (int q = 0; q < NCOL; ++q) {
            int wi;
            if (DENSE_W) wi = wo0 - pad_l + q;
            else wi = (wo0 + q / 3) * stride - pad_l + (q % 3) * dil_w;
            ok[q] = (wi >= 0 && wi < Win);
            if (ok[q]) v[q].load(in_row + (long long)wi * C); else v[q].zero();
        }
#pragma unroll
        for (int q = 0; q < NCOL; ++q) {
            float f[8];
            v[q].to_float(f);
            if (HAS_AFFINE || in_act != DLV3P_ACT_NONE) {
                if (ok[q]) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float u = HAS_AFFINE ? fmaf(f[k], sc[k], sh[k]) : f[k];
                        f[k] = apply_act(u, in_act);
                    }
                }
            }
            if (DENSE_W) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int o = q - j;
                    if (o >= 0 && o < TW) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc[o][k] = fmaf(f[k], wr[j][k], acc[o][k]);
                    }
                }
            } else {
                const int o = q / 3, j = q % 3;
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[o][k] = fmaf(f[k], wr[j][k], acc[o][k]);
            }
        }
    }

    const long long obase = (((long long)n * Hout + ho) * Wout) * C + c0;
#pragma unroll
    for (int o = 0; o < TW; ++o) {
        const int wo = wo0 + o;
        if (wo >= Wout) break;
        const long long off = obase + (long long)wo * C;
        if (HAS_EPI) {
            if (mask_src == nullptr && m_shift != nullptr) {
                // output epilogue (inference DepthwiseConv2D -> BatchNormalization -> ReLU/ReLU6): act(scale*acc + shift)
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    acc[o][k] = apply_act(fmaf(acc[o][k], __ldg(m_scale + c0 + k), __ldg(m_shift + c0 + k)), m_act);
            }
            if (mask_src != nullptr && m_act != DLV3P_ACT_NONE) {
                Vec8<T> mv; mv.load(mask_src + off);
                float mf[8]; mv.to_float(mf);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float u = mf[k];
                    if (m_scale != nullptr) u = fmaf(u, __ldg(m_scale + c0 + k), __ldg(m_shift + c0 + k));
                    acc[o][k] *= act_mask(u, m_act);
                }
            }
            if (addend != nullptr) {
                Vec8<T> av; av.load_rw(addend + off);
                float af[8]; av.to_float(af);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[o][k] += af[k];
            }
        }
        Vec8<T> r; r.from_float(acc[o]);
        r.store(out + off);
    }
}

// input gradient for stride > 1 (MobileNetV2 block_1/3/6 depthwise): gather form, one output pixel per thread
template <typename T>
__global__ void __launch_bounds__(256)
dw_dgrad_strided_kernel(const T* __restrict__ dy, const float* __restrict__ w, T* __restrict__ dx, int N, int H,
                        int W, int C, int Ho, int Wo, int stride, int dil_h, int dil_w, int pad_t, int pad_l,
                        const T* __restrict__ mask_src, const float* __restrict__ m_scale,
                        const float* __restrict__ m_shift, int m_act, const T* __restrict__ addend,
                        long long total) {
    const int CV = C >> 3;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long t = idx / CV;
    const int wi = (int)(t % W); t /= W;
    const int hi = (int)(t % H);
    const int n = (int)(t / H);
    const int c0 = cv << 3;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int hn = hi + pad_t - i * dil_h;
        if (hn < 0 || (hn % stride) != 0) continue;
        const int ho = hn / stride;
        if (ho >= Ho) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int wn = wi + pad_l - j * dil_w;
            if (wn < 0 || (wn % stride) != 0) continue;
            const int wo = wn / stride;
            if (wo >= Wo) continue;
            Vec8<T> v; v.load(dy + (((long long)n * Ho + ho) * Wo + wo) * C + c0);
            float f[8]; v.to_float(f);
            const float* wp = w + (i * 3 + j) * C + c0;
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(f[k], __ldg(wp + k), acc[k]);
        }
    }
    const long long off = (((long long)n * H + hi) * W + wi) * C + c0;
    if (mask_src != nullptr && m_act != DLV3P_ACT_NONE) {
        Vec8<T> mv; mv.load(mask_src + off);
        float mf[8]; mv.to_float(mf);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float u = mf[k];
            if (m_scale != nullptr) u = fmaf(u, __ldg(m_scale + c0 + k), __ldg(m_shift + c0 + k));
            acc[k] *= act_mask(u, m_act);
        }
    }
    if (addend != nullptr) {
        Vec8<T> av; av.load_rw(addend + off);
        float af[8]; av.to_float(af);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += af[k];
    }
    Vec8<T> r; r.from_float(acc);
    r.store(dx + off);
}

// filter gradient: per-channel 9-tap reduction over all N*Ho*Wo output pixels.
// block = CVB channel-vectors x (256/CVB) pixel lanes; each thread keeps 9x8 fp32 partials in registers,
// the block reduces them through shared memory and issues one fp32 atomicAdd per (tap, channel).
template <typename T, int CVB>
__global__ void __launch_bounds__(256)
dw_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, int N, int H, int W,
                int C, int Ho, int Wo, int stride, int dil_h, int dil_w, int pad_t, int pad_l,
                const float* __restrict__ in_scale, const float* __restrict__ in_shift, int in_act,
                long long npix, long long pix_per_block) {
    constexpr int PL = 256 / CVB;
    const int CV = C >> 3;
    const int tx = threadIdx.x % CVB;
    const int ty = threadIdx.x / CVB;
    const int cv = blockIdx.x * CVB + tx;
    const bool active = cv < CV;
    const int c0 = cv << 3;

    float acc[9][8];
#pragma unroll
    for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[a][k] = 0.f;

    float sc[8], sh[8];
    const bool affine = (in_scale != nullptr);
    if (active && affine) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { sc[k] = __ldg(in_scale + c0 + k); sh[k] = __ldg(in_shift + c0 + k); }
    }

    const long long p_begin = (long long)blockIdx.y * pix_per_block;
    long long p_end = p_begin + pix_per_block;
    if (p_end > npix) p_end = npix;
    if (active) {
        for (long long p = p_begin + ty; p < p_end; p += PL) {
            const int wo = (int)(p % Wo);
            long long t = p / Wo;
            const int ho = (int)(t % Ho);
            const int n = (int)(t / Ho);
            Vec8<T> g; g.load_stream(dy + p * C + c0);
            float gf[8]; g.to_float(gf);
            const T* xn = x + (long long)n * H * W * C + c0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int hi = ho * stride - pad_t + i * dil_h;
                if (hi < 0 || hi >= H) continue;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int wi = wo * stride - pad_l + j * dil_w;
                    if (wi < 0 || wi >= W) continue;
                    Vec8<T> v; v.load(xn + ((long long)hi * W + wi) * C);
                    float f[8]; v.to_float(f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        float u = affine ? fmaf(f[k], sc[k], sh[k]) : f[k];
                        u = apply_act(u, in_act);
                        acc[i * 3 + j][k] = fmaf(u, gf[k], acc[i * 3 + j][k]);
                    }
                }
            }
        }
    }
    __shared__ float red[PL][CVB][8];
#pragma unroll
    for (int a = 0; a < 9; ++a) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) red[ty][tx][k] = acc[a][k];
        __syncthreads();
        // CVB*8 channel sums, 256 threads: thread t sums column (t % (CVB*8)) when t < CVB*8
        for (int col = threadIdx.x; col < CVB * 8; col += 256) {
            const int cx = col >> 3, k = col & 7;
            const int cvv = blockIdx.x * CVB + cx;
            if (cvv < CV) {
                float s = 0.f;
#pragma unroll
                for (int r = 0; r < PL; ++r) s += red[r][cx][k];
                atomicAdd(dw + a * C + (cvv << 3) + k, s);
            }
        }
    }
}

// Dense-tap (stride 1, dil_w 1) filter gradient: each thread walks strips of TW=4 output columns, so every input
// vector is loaded / converted / activated once and reused by the three taps of its row (22 vector loads per 4
// pixels instead of 40), which is what the issue-bound per-pixel kernel above spends most of its slots on.
template <typename T, int CVB>
__global__ void __launch_bounds__(128)
dw_wgrad_strip_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, int N, int H, int W,
                      int C, int Ho, int Wo, int dil_h, int pad_t, int pad_l, const float* __restrict__ in_scale,
                      const float* __restrict__ in_shift, int in_act, long long nstrips, long long strips_per_block) {
    constexpr int TW = 4;
    constexpr int PL = 128 / CVB;
    const int CV = C >> 3;
    const int WS = (Wo + TW - 1) / TW;
    const int tx = threadIdx.x % CVB;
    const int ty = threadIdx.x / CVB;
    const int cv = blockIdx.x * CVB + tx;
    const bool active = cv < CV;
    const int c0 = cv << 3;
    float acc[9][8];
#pragma unroll
    for (int a = 0; a < 9; ++a)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[a][k] = 0.f;
    float sc[8], sh[8];
    const bool affine = (in_scale != nullptr);
    if (active && affine) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { sc[k] = __ldg(in_scale + c0 + k); sh[k] = __ldg(in_shift + c0 + k); }
    }
    const long long q_begin = (long long)blockIdx.y * strips_per_block;
    long long q_end = q_begin + strips_per_block;
    if (q_end > nstrips) q_end = nstrips;
    if (active) {
        for (long long q = q_begin + ty; q < q_end; q += PL) {
            const int ws = (int)(q % WS);
            long long t = q / WS;
            const int ho = (int)(t % Ho);
            const int n = (int)(t / Ho);
            const int wo0 = ws * TW;
            float g[TW][8];
            const T* dyp = dy + (((long long)n * Ho + ho) * Wo + wo0) * C + c0;
#pragma unroll
            for (int o = 0; o < TW; ++o) {
                if (wo0 + o < Wo) {
                    Vec8<T> v; v.load_stream(dyp + (long long)o * C);
                    v.to_float(g[o]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) g[o][k] = 0.f;
                }
            }
            const T* xn = x + (long long)n * H * W * C + c0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int hi = ho - pad_t + i * dil_h;
                if (hi < 0 || hi >= H) continue;
                const T* xr = xn + (long long)hi * W * C;
                Vec8<T> raw[TW + 2];
                bool ok[TW + 2];
#pragma unroll
                for (int c = 0; c < TW + 2; ++c) {
                    const int wi = wo0 - pad_l + c;
                    ok[c] = (wi >= 0 && wi < W);
                    if (ok[c]) raw[c].load(xr + (long long)wi * C); else raw[c].zero();
                }
#pragma unroll
                for (int c = 0; c < TW + 2; ++c) {
                    float f[8];
                    raw[c].to_float(f);
                    if (ok[c] && (affine || in_act != DLV3P_ACT_NONE)) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) f[k] = apply_act(affine ? fmaf(f[k], sc[k], sh[k]) : f[k], in_act);
                    }
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int o = c - j;            // output column of the strip that sees this input under tap j
                        if (o >= 0 && o < TW) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) acc[i * 3 + j][k] = fmaf(f[k], g[o][k], acc[i * 3 + j][k]);
                        }
                    }
                }
            }
        }
    }
    __shared__ float red[PL][CVB][8];
#pragma unroll
    for (int a = 0; a < 9; ++a) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) red[ty][tx][k] = acc[a][k];
        __syncthreads();
        for (int col = threadIdx.x; col < CVB * 8; col += 128) {
            const int cx = col >> 3, k = col & 7;
            const int cvv = blockIdx.x * CVB + cx;
            if (cvv < CV) {
                float s = 0.f;
#pragma unroll
                for (int r = 0; r < PL; ++r) s += red[r][cx][k];
                atomicAdd(dw + a * C + (cvv << 3) + k, s);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
template <typename T, bool HAS_AFFINE, bool HAS_EPI>
static int launch_dw_conv(const T* in, const float* w, T* out, int N, int Hin, int Win, int C, int Hout, int Wout,
                          int stride, int dil_h, int dil_w, int pad_t, int pad_l, int flip, const float* in_scale,
                          const float* in_shift, int in_act, const T* mask_src, const float* m_scale,
                          const float* m_shift, int m_act, const T* addend, cudaStream_t st) {
    const int CV = C / 8;
    const bool dense = (stride == 1 && dil_w == 1);
    if (dense) {
        constexpr int TW = 4;
        const long long total = (long long)N * Hout * ((Wout + TW - 1) / TW) * CV;
        dw_conv_kernel<T, TW, true, HAS_AFFINE, HAS_EPI><<<cdiv(total, 256), 256, 0, st>>>(
            in, w, out, N, Hin, Win, C, Hout, Wout, stride, dil_h, dil_w, pad_t, pad_l, flip, in_scale, in_shift,
            in_act, mask_src, m_scale, m_shift, m_act, addend, total);
    } else {
        constexpr int TW = 2;
        const long long total = (long long)N * Hout * ((Wout + TW - 1) / TW) * CV;
        dw_conv_kernel<T, TW, false, HAS_AFFINE, HAS_EPI><<<cdiv(total, 256), 256, 0, st>>>(
            in, w, out, N, Hin, Win, C, Hout, Wout, stride, dil_h, dil_w, pad_t, pad_l, flip, in_scale, in_shift,
            in_act, mask_src, m_scale, m_shift, m_act, addend, total);
    }
    return check_launch("dwconv3x3");
}

static int check_dw_args(const void* a, const void* b, const void* c, int N, int H, int W, int C, int stride,
                         int dil_h, int dil_w, int Ho, int Wo) {
    DLV3P_REQUIRE(a && b && c, DLV3P_ERR_SHAPE, "dwconv3x3: null pointer");
    DLV3P_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0, DLV3P_ERR_SHAPE,
                  "dwconv3x3: non-positive extent N=%d H=%d W=%d C=%d Ho=%d Wo=%d", N, H, W, C, Ho, Wo);
    DLV3P_REQUIRE(C % 8 == 0, DLV3P_ERR_SHAPE, "dwconv3x3: C=%d must be a multiple of 8", C);
    DLV3P_REQUIRE(stride >= 1 && stride <= 2 && dil_h >= 1 && dil_w >= 1, DLV3P_ERR_SHAPE,
                  "dwconv3x3: stride=%d dil=(%d,%d) unsupported", stride, dil_h, dil_w);
    DLV3P_REQUIRE(aligned16(a) && aligned16(c), DLV3P_ERR_ALIGN, "dwconv3x3: tensors must be 16-byte aligned");
    DLV3P_REQUIRE(aligned16(b), DLV3P_ERR_ALIGN, "dwconv3x3: filter must be 16-byte aligned");
    return 0;
}

}  // namespace dlv3p

using namespace dlv3p;

extern "C" int dlv3p_dwconv3x3_fwd(const void* x, const float* w, void* y, int N, int H, int W, int C, int stride,
                                   int dil_h, int dil_w, int pad_t, int pad_l, int Ho, int Wo,
                                   const float* in_scale, const float* in_shift, int in_act, int dtype,
                                   void* stream) {
    int rc = check_dw_args(x, w, y, N, H, W, C, stride, dil_h, dil_w, Ho, Wo);
    if (rc) return rc;
    DLV3P_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), DLV3P_ERR_SHAPE,
                  "dwconv3x3_fwd: in_scale and in_shift must both be given or both be NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DLV3P_BF16 && stride == 1 && dil_h == 1 && dil_w == 1) {
        // (with in_scale: BatchNormalization + ReLU/ReLU6 of the producing layer fused on load, NaN-filled halo)
        rc = launch_dw_conv_tma((const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, N, H, W, C, Ho, Wo, pad_t, pad_l, 0,
                                in_act, nullptr, nullptr, nullptr, 0, nullptr, st, in_scale, in_shift);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    if (dtype == DLV3P_BF16 && stride == 1 && (dil_h > 1 || dil_w > 1) && in_scale == nullptr && pad_t == dil_h &&
        pad_l == dil_w && Ho == H && Wo == W) {
        // atrous taps (ASPP): the whole image of one (n, channel block) staged in shared memory by one TMA box
        rc = launch_dw_image(0, (const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, nullptr, nullptr, N, H, W, C, dil_h,
                             dil_w, 0, in_act, nullptr, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        if (in_scale)
            return launch_dw_conv<T, true, false>((const T*)x, w, (T*)y, N, H, W, C, Ho, Wo, stride, dil_h, dil_w,
                                                  pad_t, pad_l, 0, in_scale, in_shift, in_act, nullptr, nullptr,
                                                  nullptr, 0, nullptr, st);
        return launch_dw_conv<T, false, false>((const T*)x, w, (T*)y, N, H, W, C, Ho, Wo, stride, dil_h, dil_w,
                                               pad_t, pad_l, 0, nullptr, nullptr, in_act, nullptr, nullptr,
                                               nullptr, 0, nullptr, st);
    });
    return 0;
}

extern "C" int dlv3p_dwconv3x3_fwd_epi(const void* x, const float* w, void* y, int N, int H, int W, int C, int stride,
                                       int dil_h, int dil_w, int pad_t, int pad_l, int Ho, int Wo,
                                       const float* out_scale, const float* out_shift, int out_act, int dtype,
                                       void* stream) {
    int rc = check_dw_args(x, w, y, N, H, W, C, stride, dil_h, dil_w, Ho, Wo);
    if (rc) return rc;
    DLV3P_REQUIRE(out_scale != nullptr && out_shift != nullptr, DLV3P_ERR_SHAPE,
                  "dwconv3x3_fwd_epi: out_scale and out_shift are required (use dwconv3x3_fwd otherwise)");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DLV3P_BF16 && stride == 1 && dil_h == 1 && dil_w == 1) {
        rc = launch_dw_conv_tma((const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, N, H, W, C, Ho, Wo, pad_t, pad_l, 0,
                                DLV3P_ACT_NONE, nullptr, nullptr, nullptr, 0, nullptr, st, nullptr, nullptr, nullptr,
                                nullptr, nullptr, nullptr, out_scale, out_shift, out_act);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        return launch_dw_conv<T, false, true>((const T*)x, w, (T*)y, N, H, W, C, Ho, Wo, stride, dil_h, dil_w, pad_t,
                                              pad_l, 0, nullptr, nullptr, DLV3P_ACT_NONE, nullptr, out_scale, out_shift,
                                              out_act, nullptr, st);
    });
    return 0;
}

extern "C" int dlv3p_dwconv3x3_dgrad(const void* dy, const float* w, void* dx, int N, int H, int W, int C,
                                     int stride, int dil_h, int dil_w, int pad_t, int pad_l, int Ho, int Wo,
                                     const void* x_pre, const float* in_scale, const float* in_shift, int in_act,
                                     const void* addend, int dtype, void* stream) {
    int rc = check_dw_args(dy, w, dx, N, H, W, C, stride, dil_h, dil_w, Ho, Wo);
    if (rc) return rc;
    DLV3P_REQUIRE(in_act == DLV3P_ACT_NONE || x_pre != nullptr, DLV3P_ERR_SHAPE,
                  "dwconv3x3_dgrad: x_pre required when in_act != NONE");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DLV3P_BF16 && stride == 1 && dil_h == 1 && dil_w == 1) {
        rc = launch_dw_conv_tma((const __nv_bfloat16*)dy, w, (__nv_bfloat16*)dx, N, Ho, Wo, C, H, W, 2 - pad_t, 2 - pad_l,
                                1, DLV3P_ACT_NONE, (const __nv_bfloat16*)x_pre, in_scale, in_shift, in_act,
                                (const __nv_bfloat16*)addend, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    if (dtype == DLV3P_BF16 && stride == 1 && (dil_h > 1 || dil_w > 1) && in_act == DLV3P_ACT_NONE && pad_t == dil_h &&
        pad_l == dil_w && Ho == H && Wo == W) {
        rc = launch_dw_image(0, (const __nv_bfloat16*)dy, w, (__nv_bfloat16*)dx, nullptr, nullptr, N, H, W, C, dil_h,
                             dil_w, 1, DLV3P_ACT_NONE, (const __nv_bfloat16*)addend, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        if (stride == 1) {
            // conv-transpose of a stride-1 conv = conv with flipped taps and complementary padding
            return launch_dw_conv<T, false, true>((const T*)dy, w, (T*)dx, N, Ho, Wo, C, H, W, 1, dil_h, dil_w,
                                                  2 * dil_h - pad_t, 2 * dil_w - pad_l, 1, nullptr, nullptr,
                                                  DLV3P_ACT_NONE, (const T*)x_pre, in_scale, in_shift, in_act,
                                                  (const T*)addend, st);
        }
        const long long total = (long long)N * H * W * (C / 8);
        dw_dgrad_strided_kernel<T><<<cdiv(total, 256), 256, 0, st>>>(
            (const T*)dy, w, (T*)dx, N, H, W, C, Ho, Wo, stride, dil_h, dil_w, pad_t, pad_l, (const T*)x_pre,
            in_scale, in_shift, in_act, (const T*)addend, total);
        return check_launch("dwconv3x3_dgrad");
    });
    return 0;
}

extern "C" int dlv3p_dwconv3x3_bn_fwd(const void* x, const float* w, void* y, int N, int H, int W, int C, int pad_t,
                                      int pad_l, int Ho, int Wo, const float* bn_sums, const float* gamma,
                                      const float* beta, float* moving_mean, float* moving_var, double count,
                                      float eps, float momentum, int updates, int in_act, float* scale, float* shift,
                                      float* mean, float* invstd, int dtype, void* stream) {
    int rc = check_dw_args(x, w, y, N, H, W, C, 1, 1, 1, Ho, Wo);
    if (rc) return rc;
    DLV3P_REQUIRE(bn_sums && scale && shift && mean && invstd && count > 0 && in_act != DLV3P_ACT_NONE, DLV3P_ERR_SHAPE,
                  "dwconv3x3_bn_fwd: BN sums, output arrays and a clamping activation are required");
    DLV3P_REQUIRE(updates == 0 || (moving_mean && moving_var), DLV3P_ERR_SHAPE, "dwconv3x3_bn_fwd: moving statistics required");
    DLV3P_REQUIRE(dtype == DLV3P_BF16, DLV3P_ERR_DTYPE, "dwconv3x3_bn_fwd: bf16 only (use bn_finalize + dwconv3x3_fwd)");
    DwBnFold f;
    f.sums = bn_sums; f.gamma = gamma; f.beta = beta; f.moving_mean = moving_mean; f.moving_var = moving_var;
    f.scale = scale; f.shift = shift; f.mean = mean; f.invstd = invstd; f.count = count; f.eps = eps;
    f.momentum = momentum; f.updates = updates;
    rc = launch_dw_conv_tma((const __nv_bfloat16*)x, w, (__nv_bfloat16*)y, N, H, W, C, Ho, Wo, pad_t, pad_l, 0, in_act,
                            nullptr, nullptr, nullptr, 0, nullptr, (cudaStream_t)stream, nullptr, nullptr, nullptr,
                            nullptr, nullptr, &f);
    if (rc < 0) return rc;
    DLV3P_REQUIRE(rc == 1, DLV3P_ERR_CUDA, "dwconv3x3_bn_fwd: the TMA path is unavailable");
    return 0;
}

extern "C" int dlv3p_dwconv3x3_dgrad_bnred(const void* dy, const float* w, void* dx, int N, int H, int W, int C,
                                           int pad_t, int pad_l, int Ho, int Wo, const void* x_pre,
                                           const float* in_scale, const float* in_shift, int in_act,
                                           const float* bn_mean, const float* bn_invstd, float* bn_red, int dtype,
                                           void* stream) {
    int rc = check_dw_args(dy, w, dx, N, H, W, C, 1, 1, 1, Ho, Wo);
    if (rc) return rc;
    DLV3P_REQUIRE(x_pre && in_scale && in_shift && in_act != DLV3P_ACT_NONE && bn_mean && bn_invstd && bn_red,
                  DLV3P_ERR_SHAPE, "dwconv3x3_dgrad_bnred: x_pre, in_scale/in_shift, an activation and the BN operands are required");
    DLV3P_REQUIRE(dtype == DLV3P_BF16, DLV3P_ERR_DTYPE, "dwconv3x3_dgrad_bnred: bf16 only (use dgrad + bn_bwd_reduce)");
    rc = launch_dw_conv_tma((const __nv_bfloat16*)dy, w, (__nv_bfloat16*)dx, N, Ho, Wo, C, H, W, 2 - pad_t, 2 - pad_l, 1,
                            DLV3P_ACT_NONE, (const __nv_bfloat16*)x_pre, in_scale, in_shift, in_act, nullptr,
                            (cudaStream_t)stream, nullptr, nullptr, bn_mean, bn_invstd, bn_red);
    if (rc < 0) return rc;
    DLV3P_REQUIRE(rc == 1, DLV3P_ERR_CUDA, "dwconv3x3_dgrad_bnred: the TMA path is unavailable");
    return 0;
}

extern "C" int dlv3p_dwconv3x3_wgrad(const void* x, const void* dy, float* dw, int N, int H, int W, int C,
                                     int stride, int dil_h, int dil_w, int pad_t, int pad_l, int Ho, int Wo,
                                     const float* in_scale, const float* in_shift, int in_act, int dtype,
                                     void* stream) {
    int rc = check_dw_args(x, dw, dy, N, H, W, C, stride, dil_h, dil_w, Ho, Wo);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int CV = C / 8;
    const long long npix = (long long)N * Ho * Wo;
    if (dtype == DLV3P_BF16 && stride == 1 && dil_h == 1 && dil_w == 1) {
        rc = launch_dw_wgrad_tma((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw, N, H, W, C, Ho, Wo, pad_t, pad_l,
                                 in_act, st, in_scale, in_shift);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    if (dtype == DLV3P_BF16 && stride == 1 && (dil_h > 1 || dil_w > 1) && in_scale == nullptr && pad_t == dil_h &&
        pad_l == dil_w && Ho == H && Wo == W) {
        rc = launch_dw_image(1, (const __nv_bfloat16*)x, nullptr, nullptr, (const __nv_bfloat16*)dy, dw, N, H, W, C, dil_h,
                             dil_w, 0, in_act, nullptr, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        if (stride == 1 && dil_w == 1) {
            const long long nstrips = (long long)N * Ho * ((Wo + 3) / 4);
            pick_cvb(CV, [&](auto cvb) {
                constexpr int CVB = decltype(cvb)::value;
                constexpr int PL = 128 / CVB;
                const int gx = cdiv(CV, CVB);
                // few fat blocks: every block ends with 72 x CVB fp32 atomics on the same 9*C addresses
                long long want = (long long)kNumSMs * 4 / gx; if (want < 1) want = 1;
                long long spb = (nstrips + want - 1) / want; spb = ((spb + PL - 1) / PL) * PL; if (spb < PL) spb = PL;
                const int gy = cdiv(nstrips, spb);
                dw_wgrad_strip_kernel<T, CVB><<<dim3(gx, gy), 128, 0, st>>>((const T*)x, (const T*)dy, dw, N, H, W, C,
                                                                            Ho, Wo, dil_h, pad_t, pad_l, in_scale,
                                                                            in_shift, in_act, nstrips, spb);
            });
            return check_launch("dwconv3x3_wgrad");
        }
        pick_cvb(CV, [&](auto cvb) {
            constexpr int CVB = decltype(cvb)::value;
            constexpr int PL = 256 / CVB;
            const int gx = cdiv(CV, CVB);
            long long want = (long long)kNumSMs * 6 / gx; if (want < 1) want = 1;
            long long ppb = (npix + want - 1) / want; ppb = ((ppb + PL - 1) / PL) * PL; if (ppb < PL) ppb = PL;
            const int gy = cdiv(npix, ppb);
            dw_wgrad_kernel<T, CVB><<<dim3(gx, gy), 256, 0, st>>>((const T*)x, (const T*)dy, dw, N, H, W, C, Ho, Wo,
                                                                  stride, dil_h, dil_w, pad_t, pad_l, in_scale,
                                                                  in_shift, in_act, npix, ppb);
        });
        return check_launch("dwconv3x3_wgrad");
    });
    return 0;
}

extern "C" int dlv3p_dwconv3x3_bwd(const void* dy, const void* x, const float* w, void* dx, float* dw, int N, int H,
                                   int W, int C, const float* in_scale, const float* in_shift, int in_act,
                                   const void* addend, const float* bn_mean, const float* bn_invstd, float* bn_red,
                                   const void* bn_y, int dtype, void* stream) {
    int rc = check_dw_args(dy, w, dx, N, H, W, C, 1, 1, 1, H, W);
    if (rc) return rc;
    DLV3P_REQUIRE(x != nullptr && dw != nullptr, DLV3P_ERR_SHAPE, "dwconv3x3_bwd: x and dw are required");
    DLV3P_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), DLV3P_ERR_SHAPE,
                  "dwconv3x3_bwd: in_scale and in_shift must both be given or both be NULL");
    DLV3P_REQUIRE(in_scale == nullptr || in_act != DLV3P_ACT_NONE, DLV3P_ERR_SHAPE,
                  "dwconv3x3_bwd: an affine input map needs an activation");
    DLV3P_REQUIRE(bn_red == nullptr || (bn_mean && bn_invstd), DLV3P_ERR_SHAPE,
                  "dwconv3x3_bwd: the BN reductions need bn_mean and bn_invstd");
    DLV3P_REQUIRE(bn_red == nullptr || bn_y != nullptr || (in_scale && addend == nullptr), DLV3P_ERR_SHAPE,
                  "dwconv3x3_bwd: reductions against x need in_scale/in_shift and no addend (pass bn_y otherwise)");
    DLV3P_REQUIRE(bn_y == nullptr || (bn_red != nullptr && aligned16(bn_y)), DLV3P_ERR_SHAPE,
                  "dwconv3x3_bwd: bn_y needs bn_red (and 16-byte alignment)");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DLV3P_BF16) {
        rc = launch_dw_bwd_tma((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, w, (__nv_bfloat16*)dx, dw, N, H, W, C,
                               in_act, in_scale, in_shift, (const __nv_bfloat16*)addend, bn_mean, bn_invstd, bn_red,
                               (const __nv_bfloat16*)bn_y, st);
        if (rc != 0) return rc < 0 ? rc : 0;
    }
    // fp32 (parity mode), no TMA, or an operand combination the fused kernel does not serve: separate entry points
    if (bn_red != nullptr && bn_y == nullptr) {
        DLV3P_REQUIRE(dtype == DLV3P_BF16, DLV3P_ERR_DTYPE, "dwconv3x3_bwd: reductions against x are bf16 only");
        rc = dlv3p_dwconv3x3_dgrad_bnred(dy, w, dx, N, H, W, C, 1, 1, H, W, x, in_scale, in_shift, in_act, bn_mean,
                                         bn_invstd, bn_red, dtype, stream);
    } else {
        rc = dlv3p_dwconv3x3_dgrad(dy, w, dx, N, H, W, C, 1, 1, 1, 1, 1, H, W, in_act != DLV3P_ACT_NONE ? x : nullptr,
                                   in_scale, in_shift, in_act, addend, dtype, stream);
    }
    if (rc) return rc;
    if (bn_y != nullptr) {
        rc = dlv3p_bn_bwd_reduce(dx, C, bn_y, C, nullptr, nullptr, bn_mean, bn_invstd, DLV3P_ACT_NONE, (int64_t)N * H * W, C,
                                 bn_red, dtype, stream);
        if (rc) return rc;
    }
    return dlv3p_dwconv3x3_wgrad(x, dy, dw, N, H, W, C, 1, 1, 1, 1, 1, H, W, in_scale, in_shift, in_act, dtype, stream);
}
