// K3 (part 1) — memory-bound kernels around the convolutions: BatchNormalization statistics / apply /
// backward, activation, add, concat slices, pooling, bilinear resize, im2col, dropout, Adam.
// Every kernel moves each byte once; algorithmic bytes are listed per entry point in DESIGN.md.
#include "common.cuh"

namespace dlv3p {

// ------------------------------------------------------------------------------------------------
// column reductions over a row-major [M,C] matrix (BN statistics and BN backward reductions)
// block = CVB channel-packs x PL row lanes; fp32 partials; one atomicAdd per (quantity, channel) per block.
// ------------------------------------------------------------------------------------------------
template <int CVB>
__device__ __forceinline__ void block_col_reduce_2x8(float (&q0)[8], float (&q1)[8], float* __restrict__ out0,
                                                     float* __restrict__ out1, int cv_base, int CV) {
    constexpr int PL = 256 / CVB;
    __shared__ float red[2][PL][CVB][8];
    const int tx = threadIdx.x % CVB, ty = threadIdx.x / CVB;
#pragma unroll
    for (int k = 0; k < 8; ++k) { red[0][ty][tx][k] = q0[k]; red[1][ty][tx][k] = q1[k]; }
    __syncthreads();
    for (int col = threadIdx.x; col < 2 * CVB * 8; col += 256) {
        const int which = col / (CVB * 8);
        const int r = col % (CVB * 8);
        const int cx = r >> 3, k = r & 7;
        const int cv = cv_base + cx;
        if (cv < CV) {
            float s = 0.f;
#pragma unroll
            for (int p = 0; p < PL; ++p) s += red[which][p][cx][k];
            atomicAdd((which ? out1 : out0) + (cv << 3) + k, s);
        }
    }
}

template <typename T, int CVB>
__global__ void __launch_bounds__(256)
bn_stats_kernel(const T* __restrict__ y, long long ld, long long M, int C, float* __restrict__ sums,
                long long rows_per_block) {
    constexpr int PL = 256 / CVB;
    const int CV = C >> 3;
    const int tx = threadIdx.x % CVB, ty = threadIdx.x / CVB;
    const int cv = blockIdx.x * CVB + tx;
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block; if (r1 > M) r1 = M;
    if (cv < CV) {
        const T* p = y + (cv << 3);
        long long r = r0 + ty;
        // two rows in flight per iteration
        for (; r + PL < r1; r += 2 * PL) {
            Vec8<T> a, b; a.load_stream(p + r * ld); b.load_stream(p + (r + PL) * ld);
            float fa[8], fb[8]; a.to_float(fa); b.to_float(fb);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                s1[k] += fa[k] + fb[k];
                s2[k] = fmaf(fa[k], fa[k], s2[k]); s2[k] = fmaf(fb[k], fb[k], s2[k]);
            }
        }
        for (; r < r1; r += PL) {
            Vec8<T> a; a.load_stream(p + r * ld);
            float fa[8]; a.to_float(fa);
#pragma unroll
            for (int k = 0; k < 8; ++k) { s1[k] += fa[k]; s2[k] = fmaf(fa[k], fa[k], s2[k]); }
        }
    }
    block_col_reduce_2x8<CVB>(s1, s2, sums, sums + C, blockIdx.x * CVB, CV);
}

template <int ACT>
__device__ __forceinline__ float act_mask_t(float v) {
    if (ACT == DLV3P_ACT_RELU) return v > 0.f ? 1.f : 0.f;
    if (ACT == DLV3P_ACT_RELU6) return (v > 0.f && v < 6.f) ? 1.f : 0.f;
    return 1.f;
}

// ACT is a template parameter and four rows (8 x 16-byte loads) are in flight per thread: the run-time-activation,
// one-row-at-a-time version sat at ~2.9 TB/s on L2-resident [16384,728] tensors with 24 warps/SM x 32 B in flight.
template <typename T, int CVB, int ACT>
__global__ void __launch_bounds__(256, ACT == DLV3P_ACT_NONE ? 3 : 2)
bn_bwd_reduce_kernel(const T* __restrict__ dz, long long ld_dz, const T* __restrict__ y, long long ld_y,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ mean, const float* __restrict__ invstd, long long M, int C,
                     float* __restrict__ red, long long rows_per_block) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int PL = 256 / CVB;
    constexpr int U = 4;
    const int CV = C >> 3;
    const int tx = threadIdx.x % CVB, ty = threadIdx.x / CVB;
    const int cv = blockIdx.x * CVB + tx;
    float s1[8], s2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { s1[k] = 0.f; s2[k] = 0.f; }
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block; if (r1 > M) r1 = M;
    if (cv < CV) {
        const int c0 = cv << 3;
        float sc[8], sh[8], mu[8];
        if (ACT != DLV3P_ACT_NONE) {
#pragma unroll
            for (int k = 0; k < 8; ++k) { sc[k] = __ldg(scale + c0 + k); sh[k] = __ldg(shift + c0 + k); }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) mu[k] = __ldg(mean + c0 + k);
        // sum(g) and sum(g*(y-mean)) — the centred form: sum(g*y) - mean*sum(g) cancels to 1/(|mean|/std) of its terms.
        // The fp32 (parity) instantiation compensates the serial per-thread sums (Kahan): a column of 10^5 pixels of
        // random-sign gradients otherwise loses 3 digits to the sqrt(N) cancellation of the sum itself.
#ifdef DLV3P_NO_KAHAN
        constexpr bool kComp = false;
#else
        constexpr bool kComp = sizeof(T) == 4;
#endif
        float c1[8], c2[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { c1[k] = 0.f; c2[k] = 0.f; }
        for (long long r = r0 + ty; r < r1; r += U * PL) {
            Vec8<T> a[U], b[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long rr = r + u * PL;
                if (rr < r1) { a[u].load_stream(dz + rr * ld_dz + c0); b[u].load_stream(y + rr * ld_y + c0); }
                else { a[u].zero(); b[u].zero(); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float g[8], v[8]; a[u].to_float(g); b[u].to_float(v);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float gg = (ACT == DLV3P_ACT_NONE) ? g[k] : g[k] * act_mask_t<ACT>(fmaf(v[k], sc[k], sh[k]));
                    if (kComp) {
                        const float y1 = gg - c1[k], t1 = s1[k] + y1;
                        c1[k] = (t1 - s1[k]) - y1; s1[k] = t1;
                        const float y2 = gg * (v[k] - mu[k]) - c2[k], t2 = s2[k] + y2;
                        c2[k] = (t2 - s2[k]) - y2; s2[k] = t2;
                    } else {
                        s1[k] += gg;
                        s2[k] = fmaf(gg, v[k] - mu[k], s2[k]);
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) s2[k] *= __ldg(invstd + c0 + k);
    }
    block_col_reduce_2x8<CVB>(s1, s2, red, red + C, blockIdx.x * CVB, CV);
}

__global__ void bn_finalize_kernel(const float* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ moving_mean,
                                   float* __restrict__ moving_var, int C, double count, float eps, float momentum,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean,
                                   float* __restrict__ invstd, int update_moving) {
    pdl_launch_dependents();
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    // the two sums are fp32; finish in fp64 so that E[x^2]-E[x]^2 does not lose more than the sums already did
    const double m = (double)sums[c] / count;
    double var = (double)sums[C + c] / count - m * m;
    if (var < 0.0) var = 0.0;
    const float is = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.f;
    const float b = beta ? beta[c] : 0.f;
    const float s = g * is;
    scale[c] = s;
    shift[c] = b - (float)m * s;
    mean[c] = (float)m;
    invstd[c] = is;
    if (update_moving) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * (float)m;
        moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * (float)unbiased;
    }
}

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mm, const float* __restrict__ mv, int C, float eps,
                               float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float s = (gamma ? gamma[c] : 1.f) / sqrtf(mv[c] + eps);
    scale[c] = s;
    shift[c] = (beta ? beta[c] : 0.f) - mm[c] * s;
}

// out = act(scale*y+shift) (+addend)
template <typename T, int V>
__global__ void __launch_bounds__(256)
affine_act_kernel(const T* __restrict__ y, long long ld_y, const float* __restrict__ scale,
                  const float* __restrict__ shift, int act, const T* __restrict__ addend, long long ld_a,
                  T* __restrict__ out, long long ld_o, long long M, int C) {
    const int CV = C / V;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * CV) return;
    const int c0 = (int)(idx % CV) * V;
    const long long r = idx / CV;
    Pack<T, V> a; a.load(y + r * ld_y + c0);
    float f[V]; a.to_float(f);
#pragma unroll
    for (int k = 0; k < V; ++k) {
        float u = f[k];
        if (scale != nullptr) u = fmaf(u, __ldg(scale + c0 + k), __ldg(shift + c0 + k));
        f[k] = apply_act(u, act);
    }
    if (addend != nullptr) {
        Pack<T, V> b; b.load_rw(addend + r * ld_a + c0);
        float g[V]; b.to_float(g);
#pragma unroll
        for (int k = 0; k < V; ++k) f[k] += g[k];
    }
    Pack<T, V> o; o.from_float(f);
    o.store(out + r * ld_o + c0);
}

// Row-loop variants: block = CVB channel-packs x (256/CVB) row lanes; every thread keeps the per-channel
// parameters of its 8 channels in registers and walks down the rows, so the kernel issues only the streaming
// 16-byte loads/stores (the flat variant re-fetched 40 scalar parameters per 16 bytes of payload and was LSU-bound).
template <typename T, int CVB, int ACT>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_kernel(const T* __restrict__ dz, long long ld_dz, const T* __restrict__ y, long long ld_y,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ red, long long M, int C, T* __restrict__ dy, long long ld_dy,
                    long long rows_per_block) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int PL = 256 / CVB;
    constexpr int U = 4;
    const int tx = threadIdx.x % CVB, ty = threadIdx.x / CVB;
    const int cv = blockIdx.x * CVB + tx;
    if (cv >= (C >> 3)) return;
    const int c0 = cv << 3;
    // dy = sc*g + y*A + B with A = -sc*invstd*red1/M, B = -sc*red0/M + mean*sc*invstd*red1/M
    float sc[8], sh[8], A[8], B[8];
    const float invM = 1.f / (float)M;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sc[k] = __ldg(scale + c0 + k);
        sh[k] = (ACT != DLV3P_ACT_NONE) ? __ldg(shift + c0 + k) : 0.f;
        if (mean != nullptr) {
            const float t = sc[k] * __ldg(invstd + c0 + k) * __ldg(red + C + c0 + k) * invM;
            A[k] = -t;
            B[k] = -sc[k] * __ldg(red + c0 + k) * invM + __ldg(mean + c0 + k) * t;
        } else {
            A[k] = 0.f; B[k] = 0.f;
        }
    }
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block; if (r1 > M) r1 = M;
    for (long long r = r0 + ty; r < r1; r += U * PL) {
        Vec8<T> a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + u * PL;
            if (rr < r1) { a[u].load_stream(dz + rr * ld_dz + c0); b[u].load_stream(y + rr * ld_y + c0); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + u * PL;
            if (rr < r1) {
                float g[8], v[8]; a[u].to_float(g); b[u].to_float(v);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float gg = (ACT == DLV3P_ACT_NONE) ? g[k] : g[k] * act_mask_t<ACT>(fmaf(v[k], sc[k], sh[k]));
                    g[k] = fmaf(sc[k], gg, fmaf(v[k], A[k], B[k]));
                }
                Vec8<T> o; o.from_float(g);
                o.store(dy + rr * ld_dy + c0);
            }
        }
    }
}

template <typename T, int CVB, int ACT, bool HAS_ADD>
__global__ void __launch_bounds__(256, HAS_ADD ? 2 : 3)
affine_act_rows_kernel(const T* __restrict__ y, long long ld_y, const float* __restrict__ scale,
                       const float* __restrict__ shift, const T* __restrict__ addend, long long ld_a,
                       T* __restrict__ out, long long ld_o, long long M, int C, long long rows_per_block) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int PL = 256 / CVB;
    constexpr int U = 4;
    const int tx = threadIdx.x % CVB, ty = threadIdx.x / CVB;
    const int cv = blockIdx.x * CVB + tx;
    if (cv >= (C >> 3)) return;
    const int c0 = cv << 3;
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        sc[k] = scale ? __ldg(scale + c0 + k) : 1.f;
        sh[k] = scale ? __ldg(shift + c0 + k) : 0.f;
    }
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block; if (r1 > M) r1 = M;
    for (long long r = r0 + ty; r < r1; r += U * PL) {
        Vec8<T> a[U], d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + u * PL;
            if (rr < r1) {
                a[u].load_stream(y + rr * ld_y + c0);
                if (HAS_ADD) d[u].load_stream(addend + rr * ld_a + c0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + u * PL;
            if (rr < r1) {
                float f[8]; a[u].to_float(f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    f[k] = fmaf(f[k], sc[k], sh[k]);
                    if (ACT == DLV3P_ACT_RELU) f[k] = fmaxf(f[k], 0.f);
                    if (ACT == DLV3P_ACT_RELU6) f[k] = fminf(fmaxf(f[k], 0.f), 6.f);
                }
                if (HAS_ADD) {
                    float e[8]; d[u].to_float(e);
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] += e[k];
                }
                Vec8<T> o; o.from_float(f);
                o.store(out + rr * ld_o + c0);
            }
        }
    }
}

// Training-mode BatchNormalization forward in ONE launch: every thread finishes the statistics of its own 8 channels
// (what bn_finalize_kernel does, fp64 for E[x^2]-E[x]^2) and then streams rows like affine_act_rows_kernel; the
// blockIdx.y == 0 row of blocks also publishes scale / shift / mean / invstd for the backward pass and applies the
// moving-statistics update (`updates` times: a layer shared by two call sites updates twice, ss.py:802 + :930).
template <typename T, int CVB, int ACT, bool HAS_ADD>
__global__ void __launch_bounds__(256, HAS_ADD ? 2 : 3)
bn_train_apply_kernel(const T* __restrict__ y, long long ld_y, const float* __restrict__ sums,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ moving_mean,
                      float* __restrict__ moving_var, double count, double inv_count, float eps, float momentum,
                      int updates, const T* __restrict__ addend, long long ld_a, T* __restrict__ out, long long ld_o,
                      long long M, int C, float* __restrict__ scale, float* __restrict__ shift,
                      float* __restrict__ mean, float* __restrict__ invstd, long long rows_per_block) {
    pdl_launch_dependents();
    pdl_wait();
    constexpr int PL = 256 / CVB;
    constexpr int U = 4;
    const int tx = threadIdx.x % CVB, ty = threadIdx.x / CVB;
    const int cv = blockIdx.x * CVB + tx;
    if (cv >= (C >> 3)) return;
    const int c0 = cv << 3;
    float sc[8], sh[8];
    const bool publish = (blockIdx.y == 0 && ty == 0);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = c0 + k;
        // fp64 only for the cancelling E[x^2] - E[x]^2 (two DFMA); 1/count comes from the host and the square root is an
        // fp32 rsqrt: this prologue runs in every thread on the critical path behind the dependency wait
        const double m = (double)__ldg(sums + c) * inv_count;
        double var = (double)__ldg(sums + C + c) * inv_count - m * m;
        if (var < 0.0) var = 0.0;
        const float is = rsqrtf((float)var + eps);
        const float g = gamma ? __ldg(gamma + c) : 1.f;
        const float b = beta ? __ldg(beta + c) : 0.f;
        sc[k] = g * is;
        sh[k] = b - (float)m * sc[k];
        if (publish) {
            scale[c] = sc[k]; shift[c] = sh[k]; mean[c] = (float)m; invstd[c] = is;
            if (updates > 0) {
                const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
                float mm = moving_mean[c], mv = moving_var[c];
                for (int u = 0; u < updates; ++u) {
                    mm = momentum * mm + (1.f - momentum) * (float)m;
                    mv = momentum * mv + (1.f - momentum) * (float)unbiased;
                }
                moving_mean[c] = mm; moving_var[c] = mv;
            }
        }
    }
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block; if (r1 > M) r1 = M;
    for (long long r = r0 + ty; r < r1; r += U * PL) {
        Vec8<T> a[U], d[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + u * PL;
            if (rr < r1) {
                a[u].load_stream(y + rr * ld_y + c0);
                if (HAS_ADD) d[u].load_stream(addend + rr * ld_a + c0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long rr = r + u * PL;
            if (rr < r1) {
                float f[8]; a[u].to_float(f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    f[k] = fmaf(f[k], sc[k], sh[k]);
                    if (ACT == DLV3P_ACT_RELU) f[k] = fmaxf(f[k], 0.f);
                    if (ACT == DLV3P_ACT_RELU6) f[k] = fminf(fmaxf(f[k], 0.f), 6.f);
                }
                if (HAS_ADD) {
                    float e[8]; d[u].to_float(e);
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] += e[k];
                }
                Vec8<T> o; o.from_float(f);
                o.store(out + rr * ld_o + c0);
            }
        }
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, T* __restrict__ dx, int act,
               const T* __restrict__ addend, long long n) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (i >= n) return;
    Pack<T, V> a, b; a.load(dy + i); b.load(x + i);
    float g[V], v[V]; a.to_float(g); b.to_float(v);
#pragma unroll
    for (int k = 0; k < V; ++k) g[k] *= act_mask(v[k], act);
    if (addend != nullptr) {
        Pack<T, V> c; c.load_rw(addend + i);
        float h[V]; c.to_float(h);
#pragma unroll
        for (int k = 0; k < V; ++k) g[k] += h[k];
    }
    Pack<T, V> o; o.from_float(g);
    o.store(dx + i);
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n) {
    const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (i >= n) return;
    Pack<T, V> pa, pb; pa.load_rw(a + i); pb.load_rw(b + i);      // out may alias a or b
    float fa[V], fb[V]; pa.to_float(fa); pb.to_float(fb);
#pragma unroll
    for (int k = 0; k < V; ++k) fa[k] += fb[k];
    Pack<T, V> o; o.from_float(fa);
    o.store(out + i);
}

// ---- pooling -----------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
maxpool3x3s2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ argmax, int N, int H,
                        int W, int C, int pad_t, int pad_l, int Ho, int Wo, const T* __restrict__ addend,
                        long long total) {
    const int CV = C >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c0 = (int)(idx % CV) << 3;
    long long t = idx / CV;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float best[8]; int arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; arg[k] = 0; }
    bool first = true;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int hi = ho * 2 - pad_t + i;
        if (hi < 0 || hi >= H) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int wi = wo * 2 - pad_l + j;
            if (wi < 0 || wi >= W) continue;
            Vec8<T> v; v.load(x + (((long long)n * H + hi) * W + wi) * C + c0);
            float f[8]; v.to_float(f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (first || f[k] > best[k]) { best[k] = f[k]; arg[k] = i * 3 + j; }
            }
            first = false;
        }
    }
    const long long off = (((long long)n * Ho + ho) * Wo + wo) * C + c0;
    if (argmax != nullptr) {
        uint2 pk;
        pk.x = (uint32_t)arg[0] | ((uint32_t)arg[1] << 8) | ((uint32_t)arg[2] << 16) | ((uint32_t)arg[3] << 24);
        pk.y = (uint32_t)arg[4] | ((uint32_t)arg[5] << 8) | ((uint32_t)arg[6] << 16) | ((uint32_t)arg[7] << 24);
        *reinterpret_cast<uint2*>(argmax + off) = pk;
    }
    if (addend != nullptr) {
        Vec8<T> a; a.load_rw(addend + off);
        float g[8]; a.to_float(g);
#pragma unroll
        for (int k = 0; k < 8; ++k) best[k] += g[k];
    }
    Vec8<T> o; o.from_float(best);
    o.store(y + off);
}

template <typename T>
__global__ void __launch_bounds__(256)
maxpool3x3s2_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ argmax, T* __restrict__ dx, int N,
                        int H, int W, int C, int pad_t, int pad_l, int Ho, int Wo, const T* __restrict__ addend,
                        long long total) {
    const int CV = C >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c0 = (int)(idx % CV) << 3;
    long long t = idx / CV;
    const int wi = (int)(t % W); t /= W;
    const int hi = (int)(t % H);
    const int n = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int hn = hi + pad_t - i;
        if (hn < 0 || (hn & 1)) continue;
        const int ho = hn >> 1;
        if (ho >= Ho) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int wn = wi + pad_l - j;
            if (wn < 0 || (wn & 1)) continue;
            const int wo = wn >> 1;
            if (wo >= Wo) continue;
            const long long off = (((long long)n * Ho + ho) * Wo + wo) * C + c0;
            const uint2 pk = __ldg(reinterpret_cast<const uint2*>(argmax + off));
            Vec8<T> g; g.load(dy + off);
            float gf[8]; g.to_float(gf);
            const uint32_t tap = (uint32_t)(i * 3 + j);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t a = ((k < 4 ? pk.x : pk.y) >> (8 * (k & 3))) & 0xffu;
                if (a == tap) acc[k] += gf[k];
            }
        }
    }
    const long long off = (((long long)n * H + hi) * W + wi) * C + c0;
    if (addend != nullptr) {
        Vec8<T> a; a.load_rw(addend + off);
        float g[8]; a.to_float(g);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += g[k];
    }
    Vec8<T> o; o.from_float(acc);
    o.store(dx + off);
}

template <typename T>
__global__ void __launch_bounds__(256)
avgpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int k, int Ho, int Wo,
                   long long total) {
    const int CV = C >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c0 = (int)(idx % CV) << 3;
    long long t = idx / CV;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) {
            Vec8<T> v; v.load(x + (((long long)n * H + ho * k + i) * W + wo * k + j) * C + c0);
            float f[8]; v.to_float(f);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] += f[q];
        }
    const float inv = 1.f / (float)(k * k);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] *= inv;
    Vec8<T> o; o.from_float(acc);
    o.store(y + (((long long)n * Ho + ho) * Wo + wo) * C + c0);
}

template <typename T>
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C, int k, int Ho, int Wo,
                   const T* __restrict__ addend, long long total) {
    const int CV = C >> 3;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c0 = (int)(idx % CV) << 3;
    long long t = idx / CV;
    const int wi = (int)(t % W); t /= W;
    const int hi = (int)(t % H);
    const int n = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    const int ho = hi / k, wo = wi / k;
    if (ho < Ho && wo < Wo) {
        Vec8<T> v; v.load(dy + (((long long)n * Ho + ho) * Wo + wo) * C + c0);
        v.to_float(acc);
        const float inv = 1.f / (float)(k * k);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] *= inv;
    }
    const long long off = (((long long)n * H + hi) * W + wi) * C + c0;
    if (addend != nullptr) {
        Vec8<T> a; a.load_rw(addend + off);
        float g[8]; a.to_float(g);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += g[q];
    }
    Vec8<T> o; o.from_float(acc);
    o.store(dx + off);
}

// ---- bilinear resize, TF ResizeBilinear(half_pixel_centers=True), integer factors --------------------------
__device__ __forceinline__ void hp_src(int o, float inv_f, int in_size, int& lo, int& hi, float& lerp) {
    const float in = ((float)o + 0.5f) * inv_f - 0.5f;
    const float fl = floorf(in);
    lo = max((int)fl, 0);
    hi = min((int)ceilf(in), in_size - 1);
    lerp = in - fl;
}

template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256)
bilinear_fwd_kernel(const TI* __restrict__ x, long long ld_x, TO* __restrict__ y, long long ld_y, int N, int H,
                    int W, int C, int fh, int fw, long long total) {
    const int CV = C / V;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c0 = (int)(idx % CV) * V;
    long long t = idx / CV;
    const int Wo = W * fw, Ho = H * fh;
    const int xo = (int)(t % Wo); t /= Wo;
    const int yo = (int)(t % Ho);
    const int n = (int)(t / Ho);
    int y0, y1, x0, x1; float ly, lx;
    hp_src(yo, 1.f / (float)fh, H, y0, y1, ly);
    hp_src(xo, 1.f / (float)fw, W, x0, x1, lx);
    const long long b = (long long)n * H * W;
    Pack<TI, V> ptl, ptr, pbl, pbr;
    ptl.load(x + (b + (long long)y0 * W + x0) * ld_x + c0);
    ptr.load(x + (b + (long long)y0 * W + x1) * ld_x + c0);
    pbl.load(x + (b + (long long)y1 * W + x0) * ld_x + c0);
    pbr.load(x + (b + (long long)y1 * W + x1) * ld_x + c0);
    float tl[V], tr[V], bl[V], br[V];
    ptl.to_float(tl); ptr.to_float(tr); pbl.to_float(bl); pbr.to_float(br);
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const float top = tl[k] + (tr[k] - tl[k]) * lx;
        const float bot = bl[k] + (br[k] - bl[k]) * lx;
        tl[k] = top + (bot - top) * ly;
    }
    Pack<TO, V> o; o.from_float(tl);
    o.store(y + (((long long)n * Ho + yo) * Wo + xo) * ld_y + c0);
}

// gather form of the transpose: each input pixel sums the output pixels whose 2x2 footprint contains it
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(256)
bilinear_bwd_kernel(const TI* __restrict__ dy, long long ld_dy, TO* __restrict__ dx, long long ld_dx, int N, int H,
                    int W, int C, int fh, int fw, const TO* __restrict__ addend, long long total) {
    const int CV = C / V;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c0 = (int)(idx % CV) * V;
    long long t = idx / CV;
    const int wi = (int)(t % W); t /= W;
    const int hi = (int)(t % H);
    const int n = (int)(t / H);
    const int Wo = W * fw, Ho = H * fh;
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
    const int ya = max(0, fh * hi - fh / 2 - 1), yb = min(Ho, fh * hi + (3 * fh) / 2 + 2);
    const int xa = max(0, fw * wi - fw / 2 - 1), xb = min(Wo, fw * wi + (3 * fw) / 2 + 2);
    for (int yo = ya; yo < yb; ++yo) {
        int y0, y1; float ly;
        hp_src(yo, 1.f / (float)fh, H, y0, y1, ly);
        float wy = 0.f;
        if (y0 == hi) wy += 1.f - ly;
        if (y1 == hi) wy += ly;
        if (wy == 0.f) continue;
        for (int xo = xa; xo < xb; ++xo) {
            int x0, x1; float lx;
            hp_src(xo, 1.f / (float)fw, W, x0, x1, lx);
            float wx = 0.f;
            if (x0 == wi) wx += 1.f - lx;
            if (x1 == wi) wx += lx;
            if (wx == 0.f) continue;
            Pack<TI, V> g; g.load(dy + (((long long)n * Ho + yo) * Wo + xo) * ld_dy + c0);
            float gf[V]; g.to_float(gf);
            const float wgt = wy * wx;
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] = fmaf(gf[k], wgt, acc[k]);
        }
    }
    const long long off = (((long long)n * H + hi) * W + wi) * ld_dx + c0;
    if (addend != nullptr) {
        Pack<TO, V> a; a.load_rw(addend + off);
        float g[V]; a.to_float(g);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += g[k];
    }
    Pack<TO, V> o; o.from_float(acc);
    o.store(dx + off);
}

// ---- im2col / col2im / subsample ----------------------------------------------------------------------------
template <typename T, int V, typename I>
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const T* __restrict__ x, T* __restrict__ col, int N, int H, int W, int C, int stride, int dil,
                 int pad_t, int pad_l, int Ho, int Wo, long long ld_col, I total) {
    const int JV = (int)(ld_col / V);
    const I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int j0 = (int)(idx % JV) * V;
    I p = idx / JV;
    T* dst = col + p * ld_col + j0;
    float f[V];
#pragma unroll
    for (int k = 0; k < V; ++k) f[k] = 0.f;
    Pack<T, V> o; o.from_float(f);
    if (j0 < 9 * C) {
        const int tap = j0 / C, c = j0 % C;
        const int wo = (int)(p % Wo); I t = p / Wo;
        const int ho = (int)(t % Ho);
        const int n = (int)(t / Ho);
        const int hi = ho * stride - pad_t + (tap / 3) * dil;
        const int wi = wo * stride - pad_l + (tap % 3) * dil;
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) o.load(x + (((I)n * H + hi) * W + wi) * C + c);
    }
    o.store(dst);
}

template <typename T, int V, typename I>
__global__ void __launch_bounds__(256)
col2im3x3_kernel(const T* __restrict__ col, T* __restrict__ dx, int N, int H, int W, int C, int stride, int dil,
                 int pad_t, int pad_l, int Ho, int Wo, long long ld_col, const T* __restrict__ addend,
                 I total) {
    const int CV = C / V;
    const I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c0 = (int)(idx % CV) * V;
    I t = idx / CV;
    const int wi = (int)(t % W); t /= W;
    const int hi = (int)(t % H);
    const int n = (int)(t / H);
    float acc[V];
#pragma unroll
    for (int k = 0; k < V; ++k) acc[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int hn = hi + pad_t - i * dil;
        if (hn < 0 || (hn % stride) != 0) continue;
        const int ho = hn / stride;
        if (ho >= Ho) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int wn = wi + pad_l - j * dil;
            if (wn < 0 || (wn % stride) != 0) continue;
            const int wo = wn / stride;
            if (wo >= Wo) continue;
            Pack<T, V> g; g.load(col + (((I)n * Ho + ho) * Wo + wo) * ld_col + (i * 3 + j) * C + c0);
            float gf[V]; g.to_float(gf);
#pragma unroll
            for (int k = 0; k < V; ++k) acc[k] += gf[k];
        }
    }
    const I off = (((I)n * H + hi) * W + wi) * C + c0;
    if (addend != nullptr) {
        Pack<T, V> a; a.load_rw(addend + off);
        float g[V]; a.to_float(g);
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] += g[k];
    }
    Pack<T, V> o; o.from_float(acc);
    o.store(dx + off);
}

template <typename T, typename I>
__global__ void __launch_bounds__(256)
subsample_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C, int stride, int Ho,
                     int Wo, I total) {
    const int CV = C >> 3;
    const I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c0 = (int)(idx % CV) << 3;
    I t = idx / CV;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    Vec8<T> v; v.load(x + (((I)n * H + ho * stride) * W + wo * stride) * C + c0);
    v.store(y + (((I)n * Ho + ho) * Wo + wo) * C + c0);
}

template <typename T, typename I>
__global__ void __launch_bounds__(256)
subsample_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W, int C, int stride, int Ho,
                     int Wo, const T* __restrict__ addend, I total) {
    const int CV = C >> 3;
    const I idx = (I)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c0 = (int)(idx % CV) << 3;
    I t = idx / CV;
    const int wi = (int)(t % W); t /= W;
    const int hi = (int)(t % H);
    const int n = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if ((hi % stride) == 0 && (wi % stride) == 0 && hi / stride < Ho && wi / stride < Wo) {
        Vec8<T> v; v.load(dy + (((I)n * Ho + hi / stride) * Wo + wi / stride) * C + c0);
        v.to_float(acc);
    }
    const I off = (((I)n * H + hi) * W + wi) * C + c0;
    if (addend != nullptr) {
        Vec8<T> a; a.load_rw(addend + off);
        float g[8]; a.to_float(g);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += g[k];
    }
    Vec8<T> o; o.from_float(acc);
    o.store(dx + off);
}

__global__ void __launch_bounds__(256)
weight_prep_kernel(const float* __restrict__ w, int K, int N, __nv_bfloat16* __restrict__ wt, long long ldt,
                   __nv_bfloat16* __restrict__ wn, long long ldn) {
    // 32x32 smem tile transpose so both the read of w and the write of wt are coalesced
    __shared__ float tile[32][33];
    pdl_launch_dependents();
    pdl_wait();
    const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int k = k0 + r, n = n0 + tx;
        float v = 0.f;
        if (k < K && n < N) v = w[(long long)k * N + n];
        tile[r][tx] = v;
        if (wn != nullptr && k < K && n < N) wn[(long long)k * ldn + n] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r, k = k0 + tx;
        if (n < N && k < K) wt[(long long)n * ldt + k] = __float2bfloat16_rn(tile[tx][r]);
    }
}

// All GEMM weights of the model in ONE launch (48 launches of a ~3 us kernel per optimizer step otherwise):
// blockIdx.y = table entry, blockIdx.x strides over that entry's 32x32 tiles.
struct WeightPrepEntry {
    const float* w; __nv_bfloat16* wt; __nv_bfloat16* wn; long long ldt, ldn; int K, N;
};
static_assert(sizeof(WeightPrepEntry) == 48, "table layout is part of the C-ABI (dlv3p_weight_prep_batch)");

__global__ void __launch_bounds__(256)
weight_prep_batch_kernel(const WeightPrepEntry* __restrict__ table) {
    // 64 x 64 tiles, two adjacent elements per thread: 8-byte fp32 loads and 4-byte packed bf16 stores in BOTH
    // orientations (the 32 x 32 / 2-byte-store version ran at 1.6 TB/s: 64-byte store segments)
    __shared__ float tile[64][65];
    pdl_launch_dependents();
    pdl_wait();
    const WeightPrepEntry e = table[blockIdx.y];
    const int tiles_n = (e.N + 63) / 64, tiles_k = (e.K + 63) / 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const bool n_even = (e.N % 2 == 0) && (e.wn == nullptr || e.ldn % 2 == 0);
    const bool k_even = (e.ldt % 2 == 0);
    for (int t = blockIdx.x; t < tiles_n * tiles_k; t += gridDim.x) {
        const int k0 = (t / tiles_n) * 64, n0 = (t % tiles_n) * 64;
        for (int r = ty; r < 64; r += 8) {
            const int k = k0 + r, n = n0 + 2 * tx;
            float v0 = 0.f, v1 = 0.f;
            if (k < e.K) {
                const float* src = e.w + (long long)k * e.N + n;
                if (n_even && n + 1 < e.N) { const float2 v = *reinterpret_cast<const float2*>(src); v0 = v.x; v1 = v.y; }
                else { if (n < e.N) v0 = src[0]; if (n + 1 < e.N) v1 = src[1]; }
                if (e.wn != nullptr) {
                    __nv_bfloat16* dst = e.wn + (long long)k * e.ldn + n;
                    if (n_even && n + 1 < e.N) *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(v0, v1);
                    else { if (n < e.N) dst[0] = __float2bfloat16_rn(v0); if (n + 1 < e.N) dst[1] = __float2bfloat16_rn(v1); }
                }
            }
            tile[r][2 * tx] = v0;
            tile[r][2 * tx + 1] = v1;
        }
        __syncthreads();
        for (int r = ty; r < 64; r += 8) {
            const int n = n0 + r, k = k0 + 2 * tx;
            if (n < e.N) {
                __nv_bfloat16* dst = e.wt + (long long)n * e.ldt + k;
                const float a = tile[2 * tx][r], b = tile[2 * tx + 1][r];
                if (k_even && k + 1 < e.K) *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(a, b);
                else { if (k < e.K) dst[0] = __float2bfloat16_rn(a); if (k + 1 < e.K) dst[1] = __float2bfloat16_rn(b); }
            }
        }
        __syncthreads();
    }
}

// ---- dropout / adam / misc ----------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, float rate, uint64_t seed,
               const uint64_t* __restrict__ seed_offset, const T* __restrict__ addend) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (seed_offset != nullptr) seed += *seed_offset;
    const uint64_t h = splitmix64(seed * 0xD1342543DE82EF95ull + (uint64_t)i);
    const float u = (float)(h >> 40) * (1.0f / 16777216.0f);   // [0,1)
    float v = (u >= rate) ? to_f<T>(x[i]) / (1.f - rate) : 0.f;
    if (addend != nullptr) v += ld_rw_f<T>(addend + i);
    y[i] = from_f<T>(v);
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr_t, float b1, float b2, float eps, float gscale, float l2) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float wi = w[i];
    const float gi = fmaf(g[i], gscale, 2.f * l2 * wi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    w[i] = wi - lr_t * mi / (sqrtf(vi) + eps);
}

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ w, long long n, float* __restrict__ out) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s = fmaf(w[i], w[i], s);
    s = warp_sum(s);
    __shared__ float part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tsum = 0.f;
        for (int k = 0; k < 8; ++k) tsum += part[k];
        atomicAdd(out, tsum);
    }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
cast2d_kernel(const TI* __restrict__ x, long long ld_x, TO* __restrict__ y, long long ld_y, long long M, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * C) return;
    const long long r = i / C; const int c = (int)(i % C);
    y[r * ld_y + c] = from_f<TO>(to_f<TI>(x[r * ld_x + c]));
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
cast_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = from_f<TO>(to_f<TI>(x[i]));
}

// One wave of long-lived blocks: gx column groups x gy row ranges with gx*gy <= 148 * blocks-per-SM, so that no
// second, mostly empty wave runs (ncu: 384 blocks at 2 blocks/SM = 1.3 waves cost two full rounds) and every block
// ends with a single set of atomics.  rows-per-block is a multiple of the row lanes PL (ragged tails are predicated).
static void col_reduce_grid(int CV, int CVB, long long M, int& gx, int& gy, long long& rpb, int blocks_per_sm = 2,
                            int unroll = 1) {
    (void)unroll;
    gx = cdiv(CV, CVB);
    const int PL = 256 / CVB;
    long long want = (long long)kNumSMs * blocks_per_sm / gx; if (want < 1) want = 1;
    rpb = (M + want - 1) / want;
    rpb = ((rpb + PL - 1) / PL) * PL;
    if (rpb < PL) rpb = PL;
    gy = cdiv(M, rpb);
}

static bool vec_ok(int C, std::initializer_list<long long> lds, std::initializer_list<const void*> ptrs) {
    if (C % 8) return false;
    for (long long l : lds) if (l % 8) return false;
    for (const void* p : ptrs) if (p && !aligned16(p)) return false;
    return true;
}

}  // namespace dlv3p

using namespace dlv3p;

extern "C" int dlv3p_bn_stats(const void* y, int64_t ld, int64_t M, int C, float* sums, int dtype, void* stream) {
    DLV3P_REQUIRE(y && sums && M > 0 && C > 0, DLV3P_ERR_SHAPE, "bn_stats: bad arguments");
    DLV3P_REQUIRE(C % 8 == 0 && ld % 8 == 0 && aligned16(y), DLV3P_ERR_ALIGN,
                  "bn_stats: C=%d, ld=%lld must be multiples of 8 and y 16-byte aligned", C, (long long)ld);
    cudaStream_t st = (cudaStream_t)stream;
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        pick_cvb(C / 8, [&](auto cvb) {
            constexpr int CVB = decltype(cvb)::value;
            int gx, gy; long long rpb;
            col_reduce_grid(C / 8, CVB, M, gx, gy, rpb, 4);
            bn_stats_kernel<T, CVB><<<dim3(gx, gy), 256, 0, st>>>((const T*)y, ld, M, C, sums, rpb);
        });
        return check_launch("bn_stats");
    });
    return 0;
}

extern "C" int dlv3p_bn_finalize(const float* sums, const float* gamma, const float* beta, float* moving_mean,
                                 float* moving_var, int C, double count, float eps, float momentum, float* scale,
                                 float* shift, float* mean, float* invstd, int update_moving, void* stream) {
    DLV3P_REQUIRE(sums && scale && shift && mean && invstd && C > 0 && count > 0, DLV3P_ERR_SHAPE,
                  "bn_finalize: bad arguments");
    DLV3P_REQUIRE(!update_moving || (moving_mean && moving_var), DLV3P_ERR_SHAPE,
                  "bn_finalize: moving statistics required when update_moving");
    launch_pdl(bn_finalize_kernel, dim3(cdiv(C, 128)), dim3(128), 0, (cudaStream_t)stream, sums, gamma, beta, moving_mean,
               moving_var, C, count, eps, momentum, scale, shift, mean, invstd, update_moving);
    return check_launch("bn_finalize");
}

extern "C" int dlv3p_bn_fold(const float* gamma, const float* beta, const float* moving_mean,
                             const float* moving_var, int C, float eps, float* scale, float* shift, void* stream) {
    DLV3P_REQUIRE(moving_mean && moving_var && scale && shift && C > 0, DLV3P_ERR_SHAPE, "bn_fold: bad arguments");
    bn_fold_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, moving_mean, moving_var, C, eps,
                                                                   scale, shift);
    return check_launch("bn_fold");
}

#define DLV3P_AFF(A, AD) launch_pdl(affine_act_rows_kernel<T, CVB, A, AD>, dim3(gx, gy), dim3(256), 0, st, (const T*)y, (long long)ld_y, \
                                    scale, shift, (const T*)addend, (long long)ld_addend, (T*)out, (long long)ld_out, (long long)M, C, rpb)

extern "C" int dlv3p_affine_act(const void* y, int64_t ld_y, const float* scale, const float* shift, int act,
                                const void* addend, int64_t ld_addend, void* out, int64_t ld_out, int64_t M, int C,
                                int dtype, void* stream) {
    DLV3P_REQUIRE(y && out && M > 0 && C > 0, DLV3P_ERR_SHAPE, "affine_act: bad arguments");
    DLV3P_REQUIRE((scale == nullptr) == (shift == nullptr), DLV3P_ERR_SHAPE, "affine_act: scale/shift mismatch");
    cudaStream_t st = (cudaStream_t)stream;
    const bool v8 = vec_ok(C, {ld_y, ld_out, addend ? ld_addend : 0}, {y, out, addend});
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        if (v8) {
            pick_cvb(C / 8, [&](auto cvb) {
                constexpr int CVB = decltype(cvb)::value;
                int gx, gy; long long rpb;
                col_reduce_grid(C / 8, CVB, M, gx, gy, rpb, addend ? 2 : 3);
                if (addend) {
                    if (act == DLV3P_ACT_NONE) DLV3P_AFF(0, true); else if (act == DLV3P_ACT_RELU) DLV3P_AFF(1, true); else DLV3P_AFF(2, true);
                } else {
                    if (act == DLV3P_ACT_NONE) DLV3P_AFF(0, false); else if (act == DLV3P_ACT_RELU) DLV3P_AFF(1, false); else DLV3P_AFF(2, false);
                }
            });
        } else {
            const long long total = M * C;
            affine_act_kernel<T, 1><<<cdiv(total, 256), 256, 0, st>>>((const T*)y, ld_y, scale, shift, act,
                                                                      (const T*)addend, ld_addend, (T*)out, ld_out,
                                                                      M, C);
        }
        return check_launch("affine_act");
    });
    return 0;
}

#define DLV3P_BNT(A, AD) launch_pdl(bn_train_apply_kernel<T, CVB, A, AD>, dim3(gx, gy), dim3(256), 0, st, (const T*)y, (long long)ld_y, \
                                    sums, gamma, beta, moving_mean, moving_var, count, 1.0 / count, eps, momentum, updates, (const T*)addend, \
                                    (long long)ld_addend, (T*)out, (long long)ld_out, (long long)M, C, scale, shift, mean, invstd, rpb)

extern "C" int dlv3p_bn_train_apply(const void* y, int64_t ld_y, const float* sums, const float* gamma, const float* beta,
                                    float* moving_mean, float* moving_var, double count, float eps, float momentum,
                                    int updates, int act, const void* addend, int64_t ld_addend, void* out,
                                    int64_t ld_out, int64_t M, int C, float* scale, float* shift, float* mean,
                                    float* invstd, int dtype, void* stream) {
    DLV3P_REQUIRE(y && out && sums && scale && shift && mean && invstd && M > 0 && C > 0 && count > 0, DLV3P_ERR_SHAPE,
                  "bn_train_apply: bad arguments");
    DLV3P_REQUIRE(updates == 0 || (moving_mean && moving_var), DLV3P_ERR_SHAPE,
                  "bn_train_apply: moving statistics required when updates > 0");
    DLV3P_REQUIRE(vec_ok(C, {ld_y, ld_out, addend ? ld_addend : 0}, {y, out, addend}), DLV3P_ERR_ALIGN,
                  "bn_train_apply: C/ld must be multiples of 8 (use bn_finalize + affine_act otherwise)");
    cudaStream_t st = (cudaStream_t)stream;
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        pick_cvb(C / 8, [&](auto cvb) {
            constexpr int CVB = decltype(cvb)::value;
            int gx, gy; long long rpb;
            col_reduce_grid(C / 8, CVB, M, gx, gy, rpb, addend ? 2 : 3);
            if (addend) {
                if (act == DLV3P_ACT_NONE) DLV3P_BNT(0, true); else if (act == DLV3P_ACT_RELU) DLV3P_BNT(1, true); else DLV3P_BNT(2, true);
            } else {
                if (act == DLV3P_ACT_NONE) DLV3P_BNT(0, false); else if (act == DLV3P_ACT_RELU) DLV3P_BNT(1, false); else DLV3P_BNT(2, false);
            }
        });
        return check_launch("bn_train_apply");
    });
    return 0;
}

#define DLV3P_BNR(A) launch_pdl(bn_bwd_reduce_kernel<T, CVB, A>, dim3(gx, gy), dim3(256), 0, st, (const T*)dz, (long long)ld_dz, \
                                (const T*)y, (long long)ld_y, scale, shift, mean, invstd, (long long)M, C, red, rpb)
#define DLV3P_BNA(A) launch_pdl(bn_bwd_apply_kernel<T, CVB, A>, dim3(gx, gy), dim3(256), 0, st, (const T*)dz, (long long)ld_dz, \
                                (const T*)y, (long long)ld_y, scale, shift, mean, invstd, red, (long long)M, C, (T*)dy, (long long)ld_dy, rpb)

extern "C" int dlv3p_bn_bwd_reduce(const void* dz, int64_t ld_dz, const void* y, int64_t ld_y, const float* scale,
                                   const float* shift, const float* mean, const float* invstd, int act, int64_t M,
                                   int C, float* red, int dtype, void* stream) {
    DLV3P_REQUIRE(dz && y && scale && shift && mean && invstd && red && M > 0 && C > 0, DLV3P_ERR_SHAPE,
                  "bn_bwd_reduce: bad arguments");
    DLV3P_REQUIRE(vec_ok(C, {ld_dz, ld_y}, {dz, y}), DLV3P_ERR_ALIGN, "bn_bwd_reduce: C/ld must be multiples of 8");
    cudaStream_t st = (cudaStream_t)stream;
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        pick_cvb(C / 8, [&](auto cvb) {
            constexpr int CVB = decltype(cvb)::value;
            int gx, gy; long long rpb;
            col_reduce_grid(C / 8, CVB, M, gx, gy, rpb, act == DLV3P_ACT_NONE ? 3 : 2);
            if (act == DLV3P_ACT_NONE) DLV3P_BNR(0); else if (act == DLV3P_ACT_RELU) DLV3P_BNR(1); else DLV3P_BNR(2);
        });
        return check_launch("bn_bwd_reduce");
    });
    return 0;
}

extern "C" int dlv3p_bn_bwd_apply(const void* dz, int64_t ld_dz, const void* y, int64_t ld_y, const float* scale,
                                  const float* shift, const float* mean, const float* invstd, int act,
                                  const float* red, int64_t M, int C, void* dy, int64_t ld_dy, int dtype,
                                  void* stream) {
    DLV3P_REQUIRE(dz && y && scale && shift && dy && M > 0 && C > 0, DLV3P_ERR_SHAPE, "bn_bwd_apply: bad arguments");
    DLV3P_REQUIRE(mean == nullptr || (invstd && red), DLV3P_ERR_SHAPE, "bn_bwd_apply: invstd/red required");
    DLV3P_REQUIRE(vec_ok(C, {ld_dz, ld_y, ld_dy}, {dz, y, dy}), DLV3P_ERR_ALIGN,
                  "bn_bwd_apply: C/ld must be multiples of 8");
    cudaStream_t st = (cudaStream_t)stream;
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        pick_cvb(C / 8, [&](auto cvb) {
            constexpr int CVB = decltype(cvb)::value;
            int gx, gy; long long rpb;
            col_reduce_grid(C / 8, CVB, M, gx, gy, rpb, 2);
            if (act == DLV3P_ACT_NONE) DLV3P_BNA(0); else if (act == DLV3P_ACT_RELU) DLV3P_BNA(1); else DLV3P_BNA(2);
        });
        return check_launch("bn_bwd_apply");
    });
    return 0;
}

extern "C" int dlv3p_act_bwd(const void* dy, const void* x, void* dx, int act, const void* addend, int64_t n,
                             int dtype, void* stream) {
    DLV3P_REQUIRE(dy && x && dx && n > 0, DLV3P_ERR_SHAPE, "act_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const bool v8 = (n % 8 == 0) && aligned16(dy) && aligned16(x) && aligned16(dx) && (!addend || aligned16(addend));
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        if (v8) act_bwd_kernel<T, 8><<<cdiv(n / 8, 256), 256, 0, st>>>((const T*)dy, (const T*)x, (T*)dx, act,
                                                                       (const T*)addend, n);
        else act_bwd_kernel<T, 1><<<cdiv(n, 256), 256, 0, st>>>((const T*)dy, (const T*)x, (T*)dx, act,
                                                               (const T*)addend, n);
        return check_launch("act_bwd");
    });
    return 0;
}

extern "C" int dlv3p_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream) {
    DLV3P_REQUIRE(a && b && out && n > 0, DLV3P_ERR_SHAPE, "add: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const bool v8 = (n % 8 == 0) && aligned16(a) && aligned16(b) && aligned16(out);
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        if (v8) add_kernel<T, 8><<<cdiv(n / 8, 256), 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, n);
        else add_kernel<T, 1><<<cdiv(n, 256), 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, n);
        return check_launch("add");
    });
    return 0;
}

extern "C" int dlv3p_copy2d(const void* x, int64_t ld_x, void* y, int64_t ld_y, int64_t M, int C,
                            const void* addend, int64_t ld_addend, int dtype, void* stream) {
    // a copy is the identity affine without activation
    return dlv3p_affine_act(x, ld_x, nullptr, nullptr, DLV3P_ACT_NONE, addend, ld_addend, y, ld_y, M, C, dtype,
                            stream);
}

namespace dlv3p {
// bf16 forward max-pool (+ residual add): packed bf16x2 compares (__hgt2_mask) with bitwise selects for the running
// maximum and the winning tap (16-bit lanes), 32-bit indexing, 2D blocks (16 channel-packs x 16 output columns) that
// walk kMpRowsF output rows.  First maximum wins (strict >), padding never wins (it is never visited).
constexpr int kMpRowsF = 4;

// AFFINE: x is the RAW conv output of a training-mode BatchNormalization layer that feeds the pool directly
// (Xception block2/3/4 sepconv2_bn -> MaxPooling2D): the pool runs on scale*x+shift without that tensor ever being
// written.  max(scale*x+shift) = scale*max(x)+shift for scale >= 0 and scale*min(x)+shift otherwise, so the packed bf16
// compares run on x with the sign bit of the negative-scale channels flipped (x ^ 0x8000 negates a bf16 exactly); the
// winner's raw value is also stored (ymax): the BatchNormalization backward reductions then only need the POOLED tensors.
template <bool AFFINE>
__global__ void __launch_bounds__(256)
maxpool3x3s2_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                             uint8_t* __restrict__ argmax, int N, int H, int W, int C, int pad_t, int pad_l, int Ho,
                             int Wo, const __nv_bfloat16* __restrict__ addend, int ncvb,
                             const float* __restrict__ scale, const float* __restrict__ shift,
                             __nv_bfloat16* __restrict__ ymax) {
    const int cvb = blockIdx.x % ncvb, wb = blockIdx.x / ncvb;
    const int cv = cvb * 16 + (int)threadIdx.x;
    const int wo = wb * 16 + (int)threadIdx.y;
    if (cv >= (C >> 3) || wo >= Wo) return;
    const int c0 = cv << 3;
    const int hblocks = (Ho + kMpRowsF - 1) / kMpRowsF;
    const int n = blockIdx.y / hblocks;
    const int h0 = (blockIdx.y % hblocks) * kMpRowsF;
    const int wi0 = wo * 2 - pad_l;
    uint32_t sgn[4] = {0u, 0u, 0u, 0u};
    float sc[8], sh[8];
    if (AFFINE) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            sc[k] = __ldg(scale + c0 + k); sh[k] = __ldg(shift + c0 + k);
            if (sc[k] < 0.f) sgn[k >> 1] |= (k & 1) ? 0x80000000u : 0x00008000u;
        }
    }
    for (int ho = h0; ho < min(h0 + kMpRowsF, Ho); ++ho) {
        uint32_t best[4], arg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { best[k] = 0xff80ff80u; arg[k] = 0u; }        // -inf, tap 0
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int hi = ho * 2 - pad_t + i;
            if (hi < 0 || hi >= H) continue;
            const int rowoff = ((n * H + hi) * W) * C + c0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int wi = wi0 + j;
                if (wi < 0 || wi >= W) continue;
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + rowoff + wi * C));
                const uint32_t vv[4] = {v.x ^ sgn[0], v.y ^ sgn[1], v.z ^ sgn[2], v.w ^ sgn[3]};
                const uint32_t tap2 = (uint32_t)(i * 3 + j) * 0x00010001u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&vv[k]),
                                                   *reinterpret_cast<const __nv_bfloat162*>(&best[k]));
                    best[k] = (vv[k] & m) | (best[k] & ~m);
                    arg[k] = (tap2 & m) | (arg[k] & ~m);
                }
            }
        }
        const int off = ((n * Ho + ho) * Wo + wo) * C + c0;
        if (argmax != nullptr) {
            uint2 pk;
            pk.x = __byte_perm(arg[0], arg[1], 0x6420);
            pk.y = __byte_perm(arg[2], arg[3], 0x6420);
            *reinterpret_cast<uint2*>(argmax + off) = pk;
        }
        if (AFFINE) {
#pragma unroll
            for (int k = 0; k < 4; ++k) best[k] ^= sgn[k];                 // back to the raw winner
            *reinterpret_cast<uint4*>(ymax + off) = make_uint4(best[0], best[1], best[2], best[3]);
            uint32_t av[4] = {0u, 0u, 0u, 0u};
            if (addend != nullptr) {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(addend + off));
                av[0] = a.x; av[1] = a.y; av[2] = a.z; av[3] = a.w;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&best[k]));
                const float2 q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&av[k]));
                const __nv_bfloat162 r = __floats2bfloat162_rn(fmaf(p.x, sc[2 * k], sh[2 * k]) + q.x,
                                                               fmaf(p.y, sc[2 * k + 1], sh[2 * k + 1]) + q.y);
                best[k] = *reinterpret_cast<const uint32_t*>(&r);
            }
        } else if (addend != nullptr) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(addend + off));
            const uint32_t av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&best[k]));
                const float2 q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&av[k]));
                const __nv_bfloat162 r = __floats2bfloat162_rn(p.x + q.x, p.y + q.y);
                best[k] = *reinterpret_cast<const uint32_t*>(&r);
            }
        }
        *reinterpret_cast<uint4*>(y + off) = make_uint4(best[0], best[1], best[2], best[3]);
    }
}
}  // namespace dlv3p

extern "C" int dlv3p_maxpool3x3s2_fwd(const void* x, void* y, uint8_t* argmax, int N, int H, int W, int C,
                                      int pad_t, int pad_l, int Ho, int Wo, const void* addend, int dtype,
                                      void* stream) {
    DLV3P_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, DLV3P_ERR_SHAPE, "maxpool_fwd: bad arguments");
    DLV3P_REQUIRE(C % 8 == 0 && aligned16(x) && aligned16(y), DLV3P_ERR_ALIGN, "maxpool_fwd: C %% 8 and alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)N * Ho * Wo * (C / 8);
    if (dtype == DLV3P_BF16 && (addend == nullptr || aligned16(addend)) && (argmax == nullptr || aligned16(argmax)) &&
        (long long)N * cdiv(Ho, kMpRowsF) <= 65535 && (long long)N * H * W * C < 0x7fffffffLL) {
        const int ncvb = cdiv(C / 8, 16);
        maxpool3x3s2_fwd_bf16_kernel<false><<<dim3(ncvb * cdiv(Wo, 16), N * cdiv(Ho, kMpRowsF)), dim3(16, 16), 0, st>>>(
            (const __nv_bfloat16*)x, (__nv_bfloat16*)y, argmax, N, H, W, C, pad_t, pad_l, Ho, Wo,
            (const __nv_bfloat16*)addend, ncvb, nullptr, nullptr, nullptr);
        return check_launch("maxpool3x3s2_fwd");
    }
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        maxpool3x3s2_fwd_kernel<T><<<cdiv(total, 256), 256, 0, st>>>((const T*)x, (T*)y, argmax, N, H, W, C, pad_t,
                                                                     pad_l, Ho, Wo, (const T*)addend, total);
        return check_launch("maxpool_fwd");
    });
    return 0;
}

extern "C" int dlv3p_maxpool3x3s2_bn_fwd(const void* x, const float* scale, const float* shift, void* y, void* ymax,
                                         uint8_t* argmax, int N, int H, int W, int C, int pad_t, int pad_l, int Ho,
                                         int Wo, const void* addend, int dtype, void* stream) {
    DLV3P_REQUIRE(x && scale && shift && y && ymax && argmax && N > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, DLV3P_ERR_SHAPE,
                  "maxpool_bn_fwd: bad arguments");
    DLV3P_REQUIRE(dtype == DLV3P_BF16, DLV3P_ERR_DTYPE, "maxpool_bn_fwd: bf16 only (use bn_train_apply + maxpool3x3s2_fwd)");
    DLV3P_REQUIRE(C % 8 == 0 && aligned16(x) && aligned16(y) && aligned16(ymax) && aligned16(argmax) &&
                  (addend == nullptr || aligned16(addend)), DLV3P_ERR_ALIGN, "maxpool_bn_fwd: C %% 8 and alignment");
    DLV3P_REQUIRE((long long)N * cdiv(Ho, kMpRowsF) <= 65535 && (long long)N * H * W * C < 0x7fffffffLL, DLV3P_ERR_UNSUPPORTED,
                  "maxpool_bn_fwd: tensor too large for the 32-bit kernel");
    const int ncvb = cdiv(C / 8, 16);
    maxpool3x3s2_fwd_bf16_kernel<true><<<dim3(ncvb * cdiv(Wo, 16), N * cdiv(Ho, kMpRowsF)), dim3(16, 16), 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, argmax, N, H, W, C, pad_t, pad_l, Ho, Wo, (const __nv_bfloat16*)addend, ncvb,
        scale, shift, (__nv_bfloat16*)ymax);
    return check_launch("maxpool3x3s2_bn_fwd");
}

namespace dlv3p {
// bf16 specialisation of the max-pool input gradient: the winning-tap test is a packed byte compare (__vcmpeq4 on four
// argmax bytes), the byte masks are widened to bf16x2 lane masks with one PRMT each and applied with AND, and the <= 4
// window contributions are summed with packed bf16 adds — 16 instructions per window instead of ~40 scalar ones (the
// scalar kernel was issue-bound at 1.4 TB/s).  Two terms sum exactly like the fp32 path (one rounding); three or four
// (only even-row/even-column pixels that win several windows) round at most twice more.
constexpr int kMpRows = 4;          // 2x2 pixel blocks walked by one thread along H

// One thread owns 8 channels of a 2x2 block of input pixels {2a, 2a+1} x {2b, 2b+1} (in padded coordinates): the block
// touches only the windows {a-1, a} x {b-1, b}, so four (argmax, dy) loads serve four output pixels and nine
// (pixel, window) pairs — the pixel-per-thread version issued up to four loads per pixel.
// BN: the pooled tensor was max(scale*y+shift) of a training-mode BatchNormalization (maxpool3x3s2_fwd_bf16_kernel
// <AFFINE>): the routed gradient g is not written; the kernel reads the raw conv output y of the same pixels and writes
// the BatchNormalization input gradient  scale*(g - red0/M - xhat*red1/M) = scale*g + y*A + B  directly.
template <bool BN>
__global__ void __launch_bounds__(256)
maxpool3x3s2_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ argmax,
                             __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C, int pad_t, int pad_l, int Ho,
                             int Wo, const __nv_bfloat16* __restrict__ addend, int ncvb,
                             const __nv_bfloat16* __restrict__ yraw, const float* __restrict__ scale,
                             const float* __restrict__ mean, const float* __restrict__ invstd,
                             const float* __restrict__ red, float inv_count) {
    // block = 16 channel-packs x 16 column pairs; grid.x = (channel-pack block, column block), grid.y = (image, row
    // block).  All index arithmetic is 32-bit and the only divisions are per-block (uniform).
    const int cvb = blockIdx.x % ncvb, wb = blockIdx.x / ncvb;
    const int cv = cvb * 16 + (int)threadIdx.x;
    const int b2 = wb * 16 + (int)threadIdx.y;             // column pair index in padded coordinates
    const int PW = (W + pad_l + 1) >> 1;                   // number of column pairs covering [pad_l, W + pad_l)
    const int PH = (H + pad_t + 1) >> 1;
    if (cv >= (C >> 3) || b2 >= PW) return;
    const int c0 = cv << 3;
    const int hblocks = (PH + kMpRows - 1) / kMpRows;
    const int n = blockIdx.y / hblocks;
    const int a0 = (blockIdx.y % hblocks) * kMpRows;
    const int img = n * Ho;
    const __nv_bfloat162 zero2 = __float2bfloat162_rn(0.f);
    float bsc[8], bA[8], bB[8];
    if (BN) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            bsc[k] = __ldg(scale + c0 + k);
            const float t = bsc[k] * __ldg(invstd + c0 + k) * __ldg(red + C + c0 + k) * inv_count;
            bA[k] = -t;
            bB[k] = -bsc[k] * __ldg(red + c0 + k) * inv_count + __ldg(mean + c0 + k) * t;
        }
    }
    for (int a2 = a0; a2 < min(a0 + kMpRows, PH); ++a2) {
        // windows (a2-1+u, b2-1+v), u,v in {0,1}
        uint2 pk[2][2]; uint4 g[2][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const int ho = a2 - 1 + u, wo = b2 - 1 + v;
                const bool ok = ho >= 0 && ho < Ho && wo >= 0 && wo < Wo;
                const int off = ok ? ((img + ho) * Wo + wo) * C + c0 : 0;
                pk[u][v] = ok ? __ldg(reinterpret_cast<const uint2*>(argmax + off)) : make_uint2(0xffffffffu, 0xffffffffu);
                g[u][v] = ok ? __ldg(reinterpret_cast<const uint4*>(dy + off)) : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int hi = 2 * a2 + r - pad_t;
            if (hi < 0 || hi >= H) continue;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int wi = 2 * b2 + c - pad_l;
                if (wi < 0 || wi >= W) continue;
                __nv_bfloat162 acc[4] = {zero2, zero2, zero2, zero2};
                // pixel (r, c) of the block is tap (r, c) of window (a2, b2), tap (r, 2) of window (a2, b2-1) when
                // c == 0, tap (2, c) of window (a2-1, b2) when r == 0, and tap (2, 2) of (a2-1, b2-1) when r == c == 0
#pragma unroll
                for (int u = 0; u < 2; ++u) {
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        if ((u == 0 && r != 0) || (v == 0 && c != 0)) continue;
                        const int ti = (u == 1) ? r : 2, tj = (v == 1) ? c : 2;
                        const uint32_t tapv = (uint32_t)(ti * 3 + tj) * 0x01010101u;
                        const uint32_t mlo = __vcmpeq4(pk[u][v].x, tapv), mhi = __vcmpeq4(pk[u][v].y, tapv);
                        uint32_t w4[4] = {g[u][v].x & __byte_perm(mlo, 0, 0x1100), g[u][v].y & __byte_perm(mlo, 0, 0x3322),
                                          g[u][v].z & __byte_perm(mhi, 0, 0x1100), g[u][v].w & __byte_perm(mhi, 0, 0x3322)};
#pragma unroll
                        for (int k = 0; k < 4; ++k) acc[k] = __hadd2(acc[k], *reinterpret_cast<__nv_bfloat162*>(&w4[k]));
                    }
                }
                const int off = ((n * H + hi) * W + wi) * C + c0;
                if (BN) {
                    const uint4 yv = __ldg(reinterpret_cast<const uint4*>(yraw + off));
                    const uint32_t yy[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 g2 = __bfloat1622float2(acc[k]);
                        const float2 y2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yy[k]));
                        acc[k] = __floats2bfloat162_rn(fmaf(bsc[2 * k], g2.x, fmaf(y2.x, bA[2 * k], bB[2 * k])),
                                                       fmaf(bsc[2 * k + 1], g2.y, fmaf(y2.y, bA[2 * k + 1], bB[2 * k + 1])));
                    }
                }
                if (addend != nullptr) {
                    const uint4 ad = __ldcg(reinterpret_cast<const uint4*>(addend + off));      // may alias dx
                    const uint32_t av[4] = {ad.x, ad.y, ad.z, ad.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 p = __bfloat1622float2(acc[k]);
                        const float2 q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&av[k]));
                        acc[k] = __floats2bfloat162_rn(p.x + q.x, p.y + q.y);
                    }
                }
                uint4 o;
                o.x = *reinterpret_cast<uint32_t*>(&acc[0]); o.y = *reinterpret_cast<uint32_t*>(&acc[1]);
                o.z = *reinterpret_cast<uint32_t*>(&acc[2]); o.w = *reinterpret_cast<uint32_t*>(&acc[3]);
                *reinterpret_cast<uint4*>(dx + off) = o;
            }
        }
    }
}

// im2col for channel counts that are not a multiple of 8 (the 3-channel image of block1_conv1 / Conv1): one thread
// gathers 8 consecutive columns (scalar loads) and issues ONE 16-byte store — the element-per-thread kernel wrote 2
// bytes per thread (400 GB/s).  ld_col % 8 == 0, col 16-byte aligned.
__global__ void __launch_bounds__(256)
im2col3x3_gather8_bf16_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ col, int N, int H, int W,
                              int C, int stride, int dil, int pad_t, int pad_l, int Ho, int Wo, long long ld_col,
                              long long total) {
    const int JV = (int)(ld_col / 8);
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int j0 = (int)(idx % JV) * 8;
    const long long p = idx / JV;
    const int wo = (int)(p % Wo); long long t = p / Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const unsigned short* xs = reinterpret_cast<const unsigned short*>(x);
    unsigned short v[8];
    int tap = j0 / C, c = j0 % C;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = 0;
        if (tap < 9) {
            const int hi = ho * stride - pad_t + (tap / 3) * dil;
            const int wi = wo * stride - pad_l + (tap % 3) * dil;
            if (hi >= 0 && hi < H && wi >= 0 && wi < W) v[k] = __ldg(xs + (((long long)n * H + hi) * W + wi) * C + c);
        }
        if (++c == C) { c = 0; ++tap; }
    }
    uint4 o;
    o.x = (uint32_t)v[0] | ((uint32_t)v[1] << 16); o.y = (uint32_t)v[2] | ((uint32_t)v[3] << 16);
    o.z = (uint32_t)v[4] | ((uint32_t)v[5] << 16); o.w = (uint32_t)v[6] | ((uint32_t)v[7] << 16);
    *reinterpret_cast<uint4*>(col + p * ld_col + j0) = o;
}

// dropout, 8 elements per thread (16-byte accesses): four 64-bit hashes give eight 24-bit uniforms
template <typename T>
__global__ void __launch_bounds__(256)
dropout8_kernel(const T* __restrict__ x, T* __restrict__ y, long long n8, float rate, uint64_t seed,
                const uint64_t* __restrict__ seed_offset, const T* __restrict__ addend) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    if (seed_offset != nullptr) seed += *seed_offset;
    Vec8<T> a; a.load_stream(x + i * 8);
    float f[8]; a.to_float(f);
    const float keep = 1.f / (1.f - rate);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint64_t h = splitmix64(seed * 0xD1342543DE82EF95ull + (uint64_t)(i * 4 + q));
        const float u0 = (float)(h >> 40) * (1.0f / 16777216.0f);
        const float u1 = (float)((h >> 8) & 0xFFFFFFull) * (1.0f / 16777216.0f);
        f[2 * q] = (u0 >= rate) ? f[2 * q] * keep : 0.f;
        f[2 * q + 1] = (u1 >= rate) ? f[2 * q + 1] * keep : 0.f;
    }
    if (addend != nullptr) {
        Vec8<T> d; d.load_rw(addend + i * 8);          // may alias y (gradient accumulated in place)
        float e[8]; d.to_float(e);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] += e[k];
    }
    Vec8<T> o; o.from_float(f);
    o.store(y + i * 8);
}
}  // namespace dlv3p

extern "C" int dlv3p_maxpool3x3s2_bwd(const void* dy, const uint8_t* argmax, void* dx, int N, int H, int W, int C,
                                      int pad_t, int pad_l, int Ho, int Wo, const void* addend, int dtype,
                                      void* stream) {
    DLV3P_REQUIRE(dy && argmax && dx && N > 0, DLV3P_ERR_SHAPE, "maxpool_bwd: bad arguments");
    DLV3P_REQUIRE(C % 8 == 0 && aligned16(dy) && aligned16(dx), DLV3P_ERR_ALIGN, "maxpool_bwd: C %% 8 and alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)N * H * W * (C / 8);
    if (dtype == DLV3P_BF16 && (addend == nullptr || aligned16(addend)) && (long long)N * H <= 65535 &&
        (long long)N * H * W * C < 0x7fffffffLL) {
        const int ncvb = cdiv(C / 8, 16);
        const int PWp = (W + pad_l + 1) / 2, PHp = (H + pad_t + 1) / 2;
        maxpool3x3s2_bwd_bf16_kernel<false><<<dim3(ncvb * cdiv(PWp, 16), N * cdiv(PHp, kMpRows)), dim3(16, 16), 0, st>>>(
            (const __nv_bfloat16*)dy, argmax, (__nv_bfloat16*)dx, N, H, W, C, pad_t, pad_l, Ho, Wo,
            (const __nv_bfloat16*)addend, ncvb, nullptr, nullptr, nullptr, nullptr, nullptr, 0.f);
        return check_launch("maxpool3x3s2_bwd");
    }
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        maxpool3x3s2_bwd_kernel<T><<<cdiv(total, 256), 256, 0, st>>>((const T*)dy, argmax, (T*)dx, N, H, W, C, pad_t,
                                                                     pad_l, Ho, Wo, (const T*)addend, total);
        return check_launch("maxpool_bwd");
    });
    return 0;
}

extern "C" int dlv3p_maxpool3x3s2_bn_bwd(const void* dy, const uint8_t* argmax, const void* yraw, const float* scale,
                                         const float* mean, const float* invstd, const float* red, double count,
                                         void* dx, int N, int H, int W, int C, int pad_t, int pad_l, int Ho, int Wo,
                                         int dtype, void* stream) {
    DLV3P_REQUIRE(dy && argmax && yraw && scale && mean && invstd && red && dx && N > 0 && count > 0, DLV3P_ERR_SHAPE,
                  "maxpool_bn_bwd: bad arguments");
    DLV3P_REQUIRE(dtype == DLV3P_BF16, DLV3P_ERR_DTYPE, "maxpool_bn_bwd: bf16 only (use maxpool3x3s2_bwd + bn_bwd_*)");
    DLV3P_REQUIRE(C % 8 == 0 && aligned16(dy) && aligned16(dx) && aligned16(yraw) && aligned16(argmax), DLV3P_ERR_ALIGN,
                  "maxpool_bn_bwd: C %% 8 and alignment");
    DLV3P_REQUIRE((long long)N * H <= 65535 && (long long)N * H * W * C < 0x7fffffffLL, DLV3P_ERR_UNSUPPORTED,
                  "maxpool_bn_bwd: tensor too large for the 32-bit kernel");
    const int ncvb = cdiv(C / 8, 16);
    const int PWp = (W + pad_l + 1) / 2, PHp = (H + pad_t + 1) / 2;
    maxpool3x3s2_bwd_bf16_kernel<true><<<dim3(ncvb * cdiv(PWp, 16), N * cdiv(PHp, kMpRows)), dim3(16, 16), 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dy, argmax, (__nv_bfloat16*)dx, N, H, W, C, pad_t, pad_l, Ho, Wo, nullptr, ncvb,
        (const __nv_bfloat16*)yraw, scale, mean, invstd, red, (float)(1.0 / count));
    return check_launch("maxpool3x3s2_bn_bwd");
}

extern "C" int dlv3p_avgpool_fwd(const void* x, void* y, int N, int H, int W, int C, int k, int Ho, int Wo,
                                 int dtype, void* stream) {
    DLV3P_REQUIRE(x && y && N > 0 && k > 0 && Ho == H / k && Wo == W / k && Ho > 0 && Wo > 0, DLV3P_ERR_SHAPE,
                  "avgpool_fwd: bad arguments (VALID, stride=k)");
    DLV3P_REQUIRE(C % 8 == 0 && aligned16(x) && aligned16(y), DLV3P_ERR_ALIGN, "avgpool_fwd: C %% 8 and alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)N * Ho * Wo * (C / 8);
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        avgpool_fwd_kernel<T><<<cdiv(total, 256), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, k, Ho, Wo, total);
        return check_launch("avgpool_fwd");
    });
    return 0;
}

extern "C" int dlv3p_avgpool_bwd(const void* dy, void* dx, int N, int H, int W, int C, int k, int Ho, int Wo,
                                 const void* addend, int dtype, void* stream) {
    DLV3P_REQUIRE(dy && dx && N > 0 && k > 0 && Ho == H / k && Wo == W / k, DLV3P_ERR_SHAPE,
                  "avgpool_bwd: bad arguments");
    DLV3P_REQUIRE(C % 8 == 0 && aligned16(dy) && aligned16(dx), DLV3P_ERR_ALIGN, "avgpool_bwd: C %% 8 and alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)N * H * W * (C / 8);
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        avgpool_bwd_kernel<T><<<cdiv(total, 256), 256, 0, st>>>((const T*)dy, (T*)dx, N, H, W, C, k, Ho, Wo,
                                                                (const T*)addend, total);
        return check_launch("avgpool_bwd");
    });
    return 0;
}

#define DLV3P_DISPATCH_2DTYPE(da, TA, db, TB, ...)                                                          \
    do {                                                                                                    \
        if ((da) == DLV3P_F32 && (db) == DLV3P_F32) { using TA = float; using TB = float; __VA_ARGS__; }    \
        else if ((da) == DLV3P_BF16 && (db) == DLV3P_BF16) { using TA = __nv_bfloat16; using TB = __nv_bfloat16; __VA_ARGS__; } \
        else if ((da) == DLV3P_BF16 && (db) == DLV3P_F32) { using TA = __nv_bfloat16; using TB = float; __VA_ARGS__; } \
        else if ((da) == DLV3P_F32 && (db) == DLV3P_BF16) { using TA = float; using TB = __nv_bfloat16; __VA_ARGS__; } \
        else { ::dlv3p::set_error("unsupported dtype pair %d,%d", (int)(da), (int)(db)); return DLV3P_ERR_DTYPE; } \
    } while (0)

extern "C" int dlv3p_bilinear_fwd(const void* x, int64_t ld_x, void* y, int64_t ld_y, int N, int H, int W, int C,
                                  int fh, int fw, int in_dtype, int out_dtype, void* stream) {
    DLV3P_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && fh >= 1 && fw >= 1, DLV3P_ERR_SHAPE,
                  "bilinear_fwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const bool v8 = vec_ok(C, {ld_x, ld_y}, {x, y});
    DLV3P_DISPATCH_2DTYPE(in_dtype, TI, out_dtype, TO, {
        if (v8) {
            const long long total = (long long)N * H * fh * W * fw * (C / 8);
            bilinear_fwd_kernel<TI, TO, 8><<<cdiv(total, 256), 256, 0, st>>>((const TI*)x, ld_x, (TO*)y, ld_y, N, H,
                                                                             W, C, fh, fw, total);
        } else {
            const long long total = (long long)N * H * fh * W * fw * C;
            bilinear_fwd_kernel<TI, TO, 1><<<cdiv(total, 256), 256, 0, st>>>((const TI*)x, ld_x, (TO*)y, ld_y, N, H,
                                                                             W, C, fh, fw, total);
        }
        return check_launch("bilinear_fwd");
    });
    return 0;
}

namespace dlv3p {
// Inference tail (segment(), ss.py:1207-1227): bilinear x(fh, fw) up-sampling of the low-resolution logits fused with
// the channel argmax — softmax is monotone, so the label map needs neither the [N,Ho,Wo,C] logits (1.27 GB fp32 at
// BASELINE cfg-5) nor the probabilities.  One thread per output pixel; the four corner logit rows are read through
// the read-only path (the whole low-resolution tensor is a few MB and stays in L1/L2); the interpolation is the
// expression of bilinear_fwd_kernel, so the labels equal those of the materialised path bit for bit (first maximum wins).
template <typename OT>
__global__ void __launch_bounds__(256)
upsample_argmax_kernel(const float* __restrict__ z, OT* __restrict__ labels, int N, int H, int W, int C, int fh, int fw,
                       long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int Wo = W * fw, Ho = H * fh;
    long long t = idx;
    const int xo = (int)(t % Wo); t /= Wo;
    const int yo = (int)(t % Ho);
    const int n = (int)(t / Ho);
    int y0, y1, x0, x1; float ly, lx;
    hp_src(yo, 1.f / (float)fh, H, y0, y1, ly);
    hp_src(xo, 1.f / (float)fw, W, x0, x1, lx);
    const long long b = (long long)n * H * W;
    const float* ptl = z + (b + (long long)y0 * W + x0) * C;
    const float* ptr = z + (b + (long long)y0 * W + x1) * C;
    const float* pbl = z + (b + (long long)y1 * W + x0) * C;
    const float* pbr = z + (b + (long long)y1 * W + x1) * C;
    float best = -INFINITY; int arg = 0;
    for (int c = 0; c < C; ++c) {
        const float tl = __ldg(ptl + c), tr = __ldg(ptr + c), bl = __ldg(pbl + c), br = __ldg(pbr + c);
        const float top = tl + (tr - tl) * lx;
        const float bot = bl + (br - bl) * lx;
        const float v = top + (bot - top) * ly;
        if (v > best) { best = v; arg = c; }
    }
    labels[idx] = (OT)arg;
}
}  // namespace dlv3p

extern "C" int dlv3p_upsample_argmax(const float* z, void* labels, int label_bytes, int N, int H, int W, int C, int fh,
                                     int fw, void* stream) {
    DLV3P_REQUIRE(z && labels && N > 0 && H > 0 && W > 0 && C > 0 && fh >= 1 && fw >= 1, DLV3P_ERR_SHAPE,
                  "upsample_argmax: bad arguments");
    DLV3P_REQUIRE(label_bytes == 4 || (label_bytes == 1 && C <= 256), DLV3P_ERR_DTYPE,
                  "upsample_argmax: labels are int32 (label_bytes 4) or uint8 (label_bytes 1, C <= 256)");
    const long long total = (long long)N * H * fh * W * fw;
    cudaStream_t st = (cudaStream_t)stream;
    if (label_bytes == 4)
        dlv3p::upsample_argmax_kernel<int32_t><<<dlv3p::cdiv(total, 256), 256, 0, st>>>(z, (int32_t*)labels, N, H, W, C, fh, fw, total);
    else
        dlv3p::upsample_argmax_kernel<uint8_t><<<dlv3p::cdiv(total, 256), 256, 0, st>>>(z, (uint8_t*)labels, N, H, W, C, fh, fw, total);
    return dlv3p::check_launch("upsample_argmax");
}

extern "C" int dlv3p_bilinear_bwd(const void* dy, int64_t ld_dy, void* dx, int64_t ld_dx, int N, int H, int W,
                                  int C, int fh, int fw, const void* addend, int dy_dtype, int dx_dtype,
                                  void* stream) {
    DLV3P_REQUIRE(dy && dx && N > 0 && H > 0 && W > 0 && C > 0 && fh >= 1 && fw >= 1, DLV3P_ERR_SHAPE,
                  "bilinear_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const bool v8 = vec_ok(C, {ld_dy, ld_dx}, {dy, dx, addend});
    DLV3P_DISPATCH_2DTYPE(dy_dtype, TI, dx_dtype, TO, {
        if (v8) {
            const long long total = (long long)N * H * W * (C / 8);
            bilinear_bwd_kernel<TI, TO, 8><<<cdiv(total, 256), 256, 0, st>>>((const TI*)dy, ld_dy, (TO*)dx, ld_dx, N,
                                                                             H, W, C, fh, fw, (const TO*)addend,
                                                                             total);
        } else {
            const long long total = (long long)N * H * W * C;
            bilinear_bwd_kernel<TI, TO, 1><<<cdiv(total, 256), 256, 0, st>>>((const TI*)dy, ld_dy, (TO*)dx, ld_dx, N,
                                                                             H, W, C, fh, fw, (const TO*)addend,
                                                                             total);
        }
        return check_launch("bilinear_bwd");
    });
    return 0;
}

extern "C" int dlv3p_im2col3x3(const void* x, void* col, int N, int H, int W, int C, int stride, int dil, int pad_t,
                               int pad_l, int Ho, int Wo, int64_t ld_col, int dtype, void* stream) {
    DLV3P_REQUIRE(x && col && N > 0 && C > 0 && Ho > 0 && Wo > 0 && ld_col >= 9 * C, DLV3P_ERR_SHAPE,
                  "im2col3x3: bad arguments (ld_col=%lld, 9C=%d)", (long long)ld_col, 9 * C);
    cudaStream_t st = (cudaStream_t)stream;
    const bool v8 = vec_ok(C, {ld_col}, {x, col});
    const long long P = (long long)N * Ho * Wo;
    if (!v8 && dtype == DLV3P_BF16 && (ld_col % 8) == 0 && aligned16(col)) {
        const long long total = P * (ld_col / 8);
        im2col3x3_gather8_bf16_kernel<<<cdiv(total, 256), 256, 0, st>>>(
            (const __nv_bfloat16*)x, (__nv_bfloat16*)col, N, H, W, C, stride, dil, pad_t, pad_l, Ho, Wo, ld_col, total);
        return check_launch("im2col3x3");
    }
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        if (v8) {
            const long long total = P * (ld_col / 8);
            if (P * ld_col < 0x7fffffffLL && (long long)N * H * W * C < 0x7fffffffLL)
                im2col3x3_kernel<T, 8, int><<<cdiv(total, 256), 256, 0, st>>>((const T*)x, (T*)col, N, H, W, C, stride, dil,
                                                                              pad_t, pad_l, Ho, Wo, ld_col, (int)total);
            else
                im2col3x3_kernel<T, 8, long long><<<cdiv(total, 256), 256, 0, st>>>((const T*)x, (T*)col, N, H, W, C, stride,
                                                                                    dil, pad_t, pad_l, Ho, Wo, ld_col, total);
        } else {
            const long long total = P * ld_col;
            im2col3x3_kernel<T, 1, long long><<<cdiv(total, 256), 256, 0, st>>>((const T*)x, (T*)col, N, H, W, C, stride, dil,
                                                                     pad_t, pad_l, Ho, Wo, ld_col, total);
        }
        return check_launch("im2col3x3");
    });
    return 0;
}

extern "C" int dlv3p_col2im3x3(const void* col, void* dx, int N, int H, int W, int C, int stride, int dil,
                               int pad_t, int pad_l, int Ho, int Wo, int64_t ld_col, const void* addend, int dtype,
                               void* stream) {
    DLV3P_REQUIRE(col && dx && N > 0 && C > 0 && Ho > 0 && Wo > 0 && ld_col >= 9 * C, DLV3P_ERR_SHAPE,
                  "col2im3x3: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const bool v8 = vec_ok(C, {ld_col}, {col, dx, addend});
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        if (v8) {
            const long long total = (long long)N * H * W * (C / 8);
            if ((long long)N * Ho * Wo * ld_col < 0x7fffffffLL && (long long)N * H * W * C < 0x7fffffffLL)
                col2im3x3_kernel<T, 8, int><<<cdiv(total, 256), 256, 0, st>>>((const T*)col, (T*)dx, N, H, W, C, stride, dil,
                                                                              pad_t, pad_l, Ho, Wo, ld_col, (const T*)addend,
                                                                              (int)total);
            else
                col2im3x3_kernel<T, 8, long long><<<cdiv(total, 256), 256, 0, st>>>((const T*)col, (T*)dx, N, H, W, C, stride,
                                                                                    dil, pad_t, pad_l, Ho, Wo, ld_col,
                                                                                    (const T*)addend, total);
        } else {
            const long long total = (long long)N * H * W * C;
            col2im3x3_kernel<T, 1, long long><<<cdiv(total, 256), 256, 0, st>>>((const T*)col, (T*)dx, N, H, W, C, stride, dil,
                                                                     pad_t, pad_l, Ho, Wo, ld_col, (const T*)addend,
                                                                     total);
        }
        return check_launch("col2im3x3");
    });
    return 0;
}

extern "C" int dlv3p_subsample_fwd(const void* x, void* y, int N, int H, int W, int C, int stride, int Ho, int Wo,
                                   int dtype, void* stream) {
    DLV3P_REQUIRE(x && y && N > 0 && stride >= 1 && (Ho - 1) * stride < H && (Wo - 1) * stride < W,
                  DLV3P_ERR_SHAPE, "subsample_fwd: bad arguments");
    DLV3P_REQUIRE(C % 8 == 0 && aligned16(x) && aligned16(y), DLV3P_ERR_ALIGN, "subsample_fwd: C %% 8 and alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)N * Ho * Wo * (C / 8);
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        if ((long long)N * H * W * C < 0x7fffffffLL)
            subsample_fwd_kernel<T, int><<<cdiv(total, 256), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, stride, Ho, Wo,
                                                                           (int)total);
        else
            subsample_fwd_kernel<T, long long><<<cdiv(total, 256), 256, 0, st>>>((const T*)x, (T*)y, N, H, W, C, stride, Ho,
                                                                                 Wo, total);
        return check_launch("subsample_fwd");
    });
    return 0;
}

extern "C" int dlv3p_subsample_bwd(const void* dy, void* dx, int N, int H, int W, int C, int stride, int Ho, int Wo,
                                   const void* addend, int dtype, void* stream) {
    DLV3P_REQUIRE(dy && dx && N > 0 && stride >= 1, DLV3P_ERR_SHAPE, "subsample_bwd: bad arguments");
    DLV3P_REQUIRE(C % 8 == 0 && aligned16(dy) && aligned16(dx), DLV3P_ERR_ALIGN, "subsample_bwd: C %% 8 and alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)N * H * W * (C / 8);
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        if ((long long)N * H * W * C < 0x7fffffffLL)
            subsample_bwd_kernel<T, int><<<cdiv(total, 256), 256, 0, st>>>((const T*)dy, (T*)dx, N, H, W, C, stride, Ho, Wo,
                                                                           (const T*)addend, (int)total);
        else
            subsample_bwd_kernel<T, long long><<<cdiv(total, 256), 256, 0, st>>>((const T*)dy, (T*)dx, N, H, W, C, stride, Ho,
                                                                                 Wo, (const T*)addend, total);
        return check_launch("subsample_bwd");
    });
    return 0;
}

extern "C" int dlv3p_weight_prep(const float* w, int K, int N, void* wt, int64_t ldt, void* wn, int64_t ldn,
                                 void* stream) {
    DLV3P_REQUIRE(w && wt && K > 0 && N > 0 && ldt >= K && (wn == nullptr || ldn >= N), DLV3P_ERR_SHAPE,
                  "weight_prep: bad arguments");
    launch_pdl(weight_prep_kernel, dim3(cdiv(N, 32), cdiv(K, 32)), dim3(256), 0, (cudaStream_t)stream, w, K, N,
               (__nv_bfloat16*)wt, (long long)ldt, (__nv_bfloat16*)wn, (long long)ldn);
    return check_launch("weight_prep");
}

extern "C" int dlv3p_weight_prep_batch(const void* table, int count, int blocks_per_entry, void* stream) {
    DLV3P_REQUIRE(table && count > 0 && blocks_per_entry > 0, DLV3P_ERR_SHAPE, "weight_prep_batch: bad arguments");
    launch_pdl(weight_prep_batch_kernel, dim3(blocks_per_entry, count), dim3(256), 0, (cudaStream_t)stream,
               (const WeightPrepEntry*)table);
    return check_launch("weight_prep_batch");
}

extern "C" int dlv3p_dropout(const void* x, void* y, int64_t n, float rate, uint64_t seed,
                             const uint64_t* seed_offset, const void* addend, int dtype, void* stream) {
    DLV3P_REQUIRE(x && y && n > 0 && rate >= 0.f && rate < 1.f, DLV3P_ERR_SHAPE, "dropout: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const bool v8 = (n % 8 == 0) && aligned16(x) && aligned16(y) && (addend == nullptr || aligned16(addend));
    DLV3P_DISPATCH_DTYPE(dtype, T, {
        if (v8) dropout8_kernel<T><<<cdiv(n / 8, 256), 256, 0, st>>>((const T*)x, (T*)y, n / 8, rate, seed, seed_offset, (const T*)addend);
        else dropout_kernel<T><<<cdiv(n, 256), 256, 0, st>>>((const T*)x, (T*)y, n, rate, seed, seed_offset, (const T*)addend);
        return check_launch("dropout");
    });
    return 0;
}

extern "C" int dlv3p_adam(float* w, const float* g, float* m, float* v, int64_t n, float lr_t, float beta1,
                          float beta2, float eps, float grad_scale, float l2, void* stream) {
    DLV3P_REQUIRE(w && g && m && v && n > 0, DLV3P_ERR_SHAPE, "adam: bad arguments");
    adam_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(w, g, m, v, n, lr_t, beta1, beta2, eps, grad_scale,
                                                                l2);
    return check_launch("adam");
}

extern "C" int dlv3p_sumsq(const float* w, int64_t n, float* out, void* stream) {
    DLV3P_REQUIRE(w && out && n > 0, DLV3P_ERR_SHAPE, "sumsq: bad arguments");
    int blocks = cdiv(n, 256 * 8); if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    sumsq_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, n, out);
    return check_launch("sumsq");
}

extern "C" int dlv3p_cast(const void* x, int x_dtype, void* y, int y_dtype, int64_t n, void* stream) {
    DLV3P_REQUIRE(x && y && n > 0, DLV3P_ERR_SHAPE, "cast: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    DLV3P_DISPATCH_2DTYPE(x_dtype, TI, y_dtype, TO, {
        cast_kernel<TI, TO><<<cdiv(n, 256), 256, 0, st>>>((const TI*)x, (TO*)y, n);
        return check_launch("cast");
    });
    return 0;
}

extern "C" int dlv3p_cast2d(const void* x, int64_t ld_x, int x_dtype, void* y, int64_t ld_y, int y_dtype, int64_t M,
                            int C, void* stream) {
    DLV3P_REQUIRE(x && y && M > 0 && C > 0 && ld_x >= C && ld_y >= C, DLV3P_ERR_SHAPE, "cast2d: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    DLV3P_DISPATCH_2DTYPE(x_dtype, TI, y_dtype, TO, {
        cast2d_kernel<TI, TO><<<cdiv(M * C, 256), 256, 0, st>>>((const TI*)x, ld_x, (TO*)y, ld_y, M, C);
        return check_launch("cast2d");
    });
    return 0;
}
