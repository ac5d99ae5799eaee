// K3 (part 2) — decoder tail: softmax, class-balanced cross-entropy (reference ss.py:438-447), argmax label map
// (MeanIoUExt, ss.py:310-311) and the fused bilinear-upsample -> softmax -> loss forward/backward.
//
// One warp owns one pixel at a time: lane c holds class c (C <= 32), so the 21 logits of a pixel are one coalesced
// load and the max / sum / dot reductions are warp shuffles.  The fused kernels never materialise the
// [N,H*f,W*f,C] logits / probabilities / gradients the reference writes three times (352 MB each at cfg-2):
// their HBM traffic is the label map (4 B/pixel) plus the low-resolution logits.
#include "common.cuh"

namespace dlv3p {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// softmax of one pixel held one-class-per-lane; inactive lanes (lane >= C) return 0
__device__ __forceinline__ float lane_softmax(float z, bool active) {
    const float m = warp_max(active ? z : -INFINITY);
    // ex2.approx / rcp.approx based intrinsics (<= 2 ulp): these kernels are issue-bound, not HBM-bound — the
    // accurate expf/logf/div expansions cost ~4x the instructions for no measurable change in the loss (parity
    // tests: loss within 1e-4, gradient within 1e-3 of the fp64 oracle)
    const float e = active ? __expf(z - m) : 0.f;
    const float s = warp_sum(e);
    return __fdividef(e, s);
}

// per-lane loss term and dL/dp for class `lane`
__device__ __forceinline__ float cb_term(float p, float y, float pw, float nw, float eps) {
    return -(pw * y * __logf(p + eps) + nw * (1.f - y) * __logf(1.f - p + eps));
}
__device__ __forceinline__ float cb_dterm(float p, float y, float pw, float nw, float eps) {
    return -__fdividef(pw * y, p + eps) + __fdividef(nw * (1.f - y), 1.f - p + eps);
}

__device__ __forceinline__ void block_atomic_sum(float v, float* out) {
    v = warp_sum(v);
    __shared__ float part[32];
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) part[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int k = 0; k < nw; ++k) s += part[k];
        atomicAdd(out, s);
    }
}

__global__ void __launch_bounds__(256)
softmax_cbloss_fwd_kernel(const float* __restrict__ z, const int32_t* __restrict__ labels,
                          const float* __restrict__ pw, const float* __restrict__ nw, float eps, long long P, int C,
                          float* __restrict__ loss_sum, float* __restrict__ probs) {
    const int lane = threadIdx.x & 31;
    const bool active = lane < C;
    const float mypw = active ? pw[lane] : 0.f, mynw = active ? nw[lane] : 0.f;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (long long p = warp; p < P; p += nwarps) {
        const float zz = active ? __ldg(z + p * C + lane) : 0.f;
        const float pr = lane_softmax(zz, active);
        const int lab = __ldg(labels + p);
        if (active) {
            acc += cb_term(pr, lane == lab ? 1.f : 0.f, mypw, mynw, eps);
            if (probs != nullptr) probs[p * C + lane] = pr;
        }
    }
    block_atomic_sum(acc, loss_sum);
}

__global__ void __launch_bounds__(256)
softmax_cbloss_bwd_kernel(const float* __restrict__ z, const int32_t* __restrict__ labels,
                          const float* __restrict__ pw, const float* __restrict__ nw, float eps, long long P, int C,
                          float gscale, float* __restrict__ dz) {
    const int lane = threadIdx.x & 31;
    const bool active = lane < C;
    const float mypw = active ? pw[lane] : 0.f, mynw = active ? nw[lane] : 0.f;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < P; p += nwarps) {
        const float zz = active ? __ldg(z + p * C + lane) : 0.f;
        const float pr = lane_softmax(zz, active);
        const int lab = __ldg(labels + p);
        const float g = active ? cb_dterm(pr, lane == lab ? 1.f : 0.f, mypw, mynw, eps) : 0.f;
        const float dot = warp_sum(g * pr);
        if (active) dz[p * C + lane] = gscale * pr * (g - dot);
    }
}

__global__ void __launch_bounds__(256)
softmax_argmax_kernel(const float* __restrict__ z, long long P, int C, float* __restrict__ probs,
                      int32_t* __restrict__ labels) {
    const int lane = threadIdx.x & 31;
    const bool active = lane < C;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p = warp; p < P; p += nwarps) {
        const float zz = active ? __ldg(z + p * C + lane) : -INFINITY;
        if (probs != nullptr) {
            const float pr = lane_softmax(zz, active);
            if (active) probs[p * C + lane] = pr;
        }
        if (labels != nullptr) {
            float bv = zz; int bi = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (lane == 0) labels[p] = bi;
        }
    }
}

__global__ void __launch_bounds__(256)
cbloss_dense_fwd_kernel(const float* __restrict__ yt, const float* __restrict__ yp, const float* __restrict__ pw,
                        const float* __restrict__ nw, float eps, long long P, int C, float* __restrict__ loss_sum) {
    float acc = 0.f;
    const long long n = P * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        acc += cb_term(yp[i], yt[i], pw[c], nw[c], eps);
    }
    block_atomic_sum(acc, loss_sum);
}

__global__ void __launch_bounds__(256)
cbloss_dense_bwd_kernel(const float* __restrict__ yt, const float* __restrict__ yp, const float* __restrict__ pw,
                        const float* __restrict__ nw, float eps, long long P, int C, float gscale,
                        float* __restrict__ dyp) {
    const long long n = P * C;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % C);
    dyp[i] = gscale * cb_dterm(yp[i], yt[i], pw[c], nw[c], eps);
}

__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp, long long P, int C,
                   float* __restrict__ dz) {
    const int lane = threadIdx.x & 31;
    const bool active = lane < C;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long q = warp; q < P; q += nwarps) {
        const float pr = active ? p[q * C + lane] : 0.f;
        const float g = active ? dp[q * C + lane] : 0.f;
        const float dot = warp_sum(g * pr);
        if (active) dz[q * C + lane] = pr * (g - dot);
    }
}

__global__ void __launch_bounds__(256)
confusion_kernel(const int32_t* __restrict__ yt, const int32_t* __restrict__ yp, long long P, int C,
                 double* __restrict__ cm) {
    extern __shared__ unsigned int hist[];
    for (int i = threadIdx.x; i < C * C; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long)gridDim.x * blockDim.x) {
        const int t = yt[i], p = yp[i];
        if (t >= 0 && t < C && p >= 0 && p < C) atomicAdd(&hist[t * C + p], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * C; i += blockDim.x)
        if (hist[i]) atomicAdd(cm + i, (double)hist[i]);
}

// ---- fused bilinear xf upsample -> softmax -> class-balanced loss -------------------------------------------
// Output pixels are grouped in f x f tiles shifted by f/2 so that every pixel of a tile interpolates between the
// same 2x2 low-resolution logits: tile (ty,tx), ty in [-1,H-1], covers rows [f*ty + f/2, f*(ty+1) + f/2) and uses
// low-res rows clamp(ty), clamp(ty+1).  One warp per tile; lane c = class c.
__device__ __forceinline__ void tile_setup(long long tile, int H, int W, int f, int& n, int& ty, int& tx) {
    const int TH = H + 1, TW = W + 1;
    tx = (int)(tile % TW) - 1; tile /= TW;
    ty = (int)(tile % TH) - 1;
    n = (int)(tile / TH);
}

template <bool BWD>
__global__ void __launch_bounds__(256)
upsample_softmax_cbloss_kernel(const float* __restrict__ zl, const int32_t* __restrict__ labels,
                               const float* __restrict__ pw, const float* __restrict__ nw, float eps, int N, int H,
                               int W, int C, int f, float gscale, float* __restrict__ loss_sum,
                               float* __restrict__ dzl, long long ntiles) {
    const int lane = threadIdx.x & 31;
    const bool active = lane < C;
    const float mypw = active ? pw[lane] : 0.f, mynw = active ? nw[lane] : 0.f;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int Ho = H * f, Wo = W * f;
    const float inv_f = 1.f / (float)f;
    float loss_acc = 0.f;
    for (long long tile = warp; tile < ntiles; tile += nwarps) {
        int n, ty, tx;
        tile_setup(tile, H, W, f, n, ty, tx);
        const int y_lo = max(ty, 0), y_hi = min(ty + 1, H - 1);
        const int x_lo = max(tx, 0), x_hi = min(tx + 1, W - 1);
        const float* base = zl + (long long)n * H * W * C + lane;
        float z00 = 0.f, z01 = 0.f, z10 = 0.f, z11 = 0.f;
        if (active) {
            z00 = __ldg(base + ((long long)y_lo * W + x_lo) * C);
            z01 = __ldg(base + ((long long)y_lo * W + x_hi) * C);
            z10 = __ldg(base + ((long long)y_hi * W + x_lo) * C);
            z11 = __ldg(base + ((long long)y_hi * W + x_hi) * C);
        }
        float g00 = 0.f, g01 = 0.f, g10 = 0.f, g11 = 0.f;
        const int ya = max(f * ty + f / 2, 0), yb = min(f * (ty + 1) + f / 2, Ho);
        const int xa = max(f * tx + f / 2, 0), xb = min(f * (tx + 1) + f / 2, Wo);
        for (int yo = ya; yo < yb; ++yo) {
            // half-pixel source coordinate; inside a tile floor(src) == ty, so lerp = src - ty (TF: in - floor(in))
            const float sy = ((float)yo + 0.5f) * inv_f - 0.5f;
            const float ly = sy - floorf(sy);
            const int32_t* lrow = labels + ((long long)n * Ho + yo) * Wo;
            for (int xo = xa; xo < xb; ++xo) {
                const float sx = ((float)xo + 0.5f) * inv_f - 0.5f;
                const float lx = sx - floorf(sx);
                const float top = z00 + (z01 - z00) * lx;
                const float bot = z10 + (z11 - z10) * lx;
                const float zz = top + (bot - top) * ly;
                const float pr = lane_softmax(zz, active);
                const int lab = __ldg(lrow + xo);
                const float y = (lane == lab) ? 1.f : 0.f;
                if (!BWD) {
                    if (active) loss_acc += cb_term(pr, y, mypw, mynw, eps);
                } else {
                    const float g = active ? cb_dterm(pr, y, mypw, mynw, eps) : 0.f;
                    const float dot = warp_sum(g * pr);
                    const float d = pr * (g - dot);
                    g00 = fmaf(d, (1.f - ly) * (1.f - lx), g00);
                    g01 = fmaf(d, (1.f - ly) * lx, g01);
                    g10 = fmaf(d, ly * (1.f - lx), g10);
                    g11 = fmaf(d, ly * lx, g11);
                }
            }
        }
        if (BWD && active) {
            float* gb = dzl + (long long)n * H * W * C + lane;
            atomicAdd(gb + ((long long)y_lo * W + x_lo) * C, gscale * g00);
            atomicAdd(gb + ((long long)y_lo * W + x_hi) * C, gscale * g01);
            atomicAdd(gb + ((long long)y_hi * W + x_lo) * C, gscale * g10);
            atomicAdd(gb + ((long long)y_hi * W + x_hi) * C, gscale * g11);
        }
    }
    if (!BWD) block_atomic_sum(loss_acc, loss_sum);
}


// ---- fused tail, fast path (f in {2,4,8,16}): one THREAD per output pixel ----------------------------------------
// The warp-per-pixel kernel above is issue-bound (4.2 M warp iterations of ~100 instructions at cfg-2).  Here a block
// of 256 threads owns a 16x16 patch of output pixels aligned to the half-factor-shifted grid, i.e. S x S (S = 16/f)
// sub-tiles whose pixels interpolate between the same 2x2 low-resolution logits.  The (S+1)^2 x C corner logits are
// staged in shared memory; each thread keeps its pixel's C logits / probabilities in registers (softmax without any
// shuffle), and the transposed-resize gradient is reduced separably through shared memory (over x per sub-tile, then
// over y by the corner that gathers its up-to-four sub-tiles) and leaves with one global RED per (corner, class).  Forward and backward share one
// pass (`fwd_bwd`), so the softmax is evaluated once per pixel per step.
// ex2 / lg2 without the denormal fix-up sequences of __expf / __logf (16 % of the kernel's issue slots were ISETP/BRA and
// @p FMUL range handling, profiles/r1_tail_ncu.md); flushing denormal probabilities to zero is far below the loss's
// epsilon of 1e-7
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_log(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * 0.6931471805599453f;
}

__device__ __forceinline__ float fast_lg2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One 16x16 tile of output pixels.  EXACT: C == CMAX, the class loops carry no `c < C` predicates.  F: the resize factor as
// a compile-time constant (2 / 4 / 8 / 16) — S = 16/F sub-tiles per side, (S+1)^2 corners; with F a runtime value the two
// reduction loops below (trip count f) and the divisions by S+1 were 45 % of the kernel's instructions at x2 and the ALU
// pipe its busiest unit (profiles/r2_tail_ncu.md).
template <int CMAX, bool FWD, bool BWD, bool EXACT, int F>
__device__ __forceinline__ void
tail_pixel_tile(float* __restrict__ sm, const float* __restrict__ zl, const int32_t* __restrict__ labels,
                const float* __restrict__ pw, const float* __restrict__ nw, float eps, int H, int W, int C, float gscale,
                float* __restrict__ loss_sum, float* __restrict__ dzl, int nby, int nbx) {
    constexpr int CS = CMAX | 1;                    // odd class stride: conflict-free per-pixel rows in smem
    constexpr int S = 16 / F, PS = S + 1;
    constexpr int LF = (F == 2) ? 1 : (F == 4) ? 2 : (F == 8) ? 3 : 4, LS = 4 - LF;
    float* patch = sm;                              // [PS][PS][CS]   corner logits
    float* spw = patch + PS * PS * CS;              // [32]
    float* snw = spw + 32;                          // [32]
    float* wtab = snw + 32;                         // [2][16]        lerp weights of the gradient reduction: 1-l, l
    float* D = wtab + 32;                           // [256][CS]      per-pixel logit gradients
    float* R = D + 256 * CS;                        // [16][S][2][CS] x-reduced partials

    const int tid = threadIdx.x;
    int b = blockIdx.x;
    const int bx = b % nbx; b /= nbx;
    const int by = b % nby;
    const int n = b / nby;
    const int Ho = H * F, Wo = W * F;
    const int ly0 = S * by - 1, lx0 = S * bx - 1;   // low-res coordinates of patch row/col 0 (before clamping)

    for (int i = tid; i < PS * PS * CS; i += 256) {
        const int c = i % CS;
        const int r = i / CS;
        const int pr = r / PS, pc = r - pr * PS;
        const int yy = min(max(ly0 + pr, 0), H - 1), xx = min(max(lx0 + pc, 0), W - 1);
        patch[i] = (c < C) ? __ldg(zl + (((long long)n * H + yy) * W + xx) * C + c) : 0.f;
    }
    if (tid < 32) {
        spw[tid] = tid < C ? __ldg(pw + tid) : 0.f; snw[tid] = tid < C ? __ldg(nw + tid) : 0.f;
        const float l = ((float)(tid & 15) + 0.5f) * (1.f / (float)F);
        wtab[tid] = (tid & 16) ? l : 1.f - l;
    }
    __syncthreads();

    const int py = tid >> 4, px = tid & 15;
    const int yo = 16 * by + py - F / 2, xo = 16 * bx + px - F / 2;
    const bool valid = (yo >= 0 && yo < Ho && xo >= 0 && xo < Wo);
    const int sy = py >> LF, sx = px >> LF;
    constexpr float inv_f = 1.f / (float)F;
    const float ly = ((float)(py & (F - 1)) + 0.5f) * inv_f;     // TF half-pixel lerp: frac((o+0.5)/f - 0.5)
    const float lx = ((float)(px & (F - 1)) + 0.5f) * inv_f;
    float loss_acc = 0.f;
    float d[CMAX];
    if (valid) {
        const float* p00 = patch + (sy * PS + sx) * CS;
        const float* p01 = p00 + CS;
        const float* p10 = p00 + PS * CS;
        const float* p11 = p10 + CS;
        float z[CMAX];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (EXACT || c < C) {
                const float top = p00[c] + (p01[c] - p00[c]) * lx;
                const float bot = p10[c] + (p11[c] - p10[c]) * lx;
                z[c] = top + (bot - top) * ly;
                m = fmaxf(m, z[c]);
            } else z[c] = -INFINITY;
        }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            z[c] = (EXACT || c < C) ? fast_exp2((z[c] - m) * 1.4426950408889634f) : 0.f;
            s += z[c];
        }
        const float inv = fast_rcp(s);              // s in [1, C]
        const int lab = __ldg(labels + ((long long)n * Ho + yo) * Wo + xo);
        float dot = 0.f, lg_acc = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (EXACT || c < C) {
                // -(pw*y*log(p+eps) + nw*(1-y)*log(1-p+eps)), y one-hot: one weight, one log argument per class; the
                // gradient w.r.t. p is -+w / arg — ONE reciprocal of the selected argument (arg in [eps, 1+eps]; the two
                // __fdividef of both branches were 21 % of the x16 kernel's instructions)
                const float p = z[c] * inv;
                const bool hit = (c == lab);
                const float w = hit ? spw[c] : snw[c];
                const float arg = (hit ? p : 1.f - p) + eps;
                if (FWD) lg_acc = fmaf(w, fast_lg2(arg), lg_acc);
                if (BWD) {
                    const float g = (hit ? -w : w) * fast_rcp(arg);
                    dot = fmaf(g, p, dot);
                    d[c] = g;
                    z[c] = p;
                }
            }
        }
        if (FWD) loss_acc = -0.6931471805599453f * lg_acc;
        if (BWD) {
#pragma unroll
            for (int c = 0; c < CMAX; ++c) d[c] = (EXACT || c < C) ? gscale * z[c] * (d[c] - dot) : 0.f;
        }
    } else if (BWD) {
#pragma unroll
        for (int c = 0; c < CMAX; ++c) d[c] = 0.f;
    }
    if (FWD) block_atomic_sum(loss_acc, loss_sum);
    if (!BWD) return;

#pragma unroll
    for (int c = 0; c < CMAX; ++c) D[tid * CS + c] = d[c];
    __syncthreads();
    // phase A: reduce over x inside each sub-tile: R[py][sx][kx][c] = sum_px D[py][px][c] * (kx ? lx : 1-lx)
    const int Cc = EXACT ? CMAX : C;
    const int itemsA = 16 * S * 2 * Cc;
    for (int i = tid; i < itemsA; i += 256) {
        const int c = i % Cc;
        int r = i / Cc;
        const int kx = r & 1; r >>= 1;
        const int ssx = r & (S - 1);
        const int ppy = r >> LS;
        float acc = 0.f;
        const float* dp = D + (ppy * 16 + (ssx << LF)) * CS + c;
        const float* wp = wtab + kx * 16;
#pragma unroll
        for (int j = 0; j < F; ++j) acc = fmaf(dp[j * CS], wp[j], acc);
        R[((ppy * S + ssx) * 2 + kx) * CS + c] = acc;
    }
    __syncthreads();
    // phase B: every corner GATHERS the y-reductions of the (up to four) sub-tiles around it and goes straight to the
    // global accumulator.  (The first version scattered 4 x S^2 x C partial sums into a shared-memory table with
    // atomicAdd(float) — CAS loops, four-way contended — which at f = 2, S = 8 was 5376 of them per block: the x2 tail of
    // the boundary-refinement configuration took 1.12 ms against 0.26 ms for the same number of output pixels at x16.)
    const int itemsB = PS * PS * Cc;
    for (int i = tid; i < itemsB; i += 256) {
        const int c = i % Cc;
        const int r = i / Cc;
        const int cy = r / PS, cx = r - cy * PS;
        float v = 0.f;
#pragma unroll
        for (int ky = 0; ky < 2; ++ky) {
            const int ssy = cy - ky;
            if (ssy < 0 || ssy >= S) continue;
            const float* wp = wtab + ky * 16;
#pragma unroll
            for (int kx = 0; kx < 2; ++kx) {
                const int ssx = cx - kx;
                if (ssx < 0 || ssx >= S) continue;
                const float* rp = R + ((((ssy << LF) * S) + ssx) * 2 + kx) * CS + c;
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < F; ++j) acc = fmaf(rp[j * S * 2 * CS], wp[j], acc);
                v += acc;
            }
        }
        if (v != 0.f) {
            const int yy = min(max(ly0 + cy, 0), H - 1), xx = min(max(lx0 + cx, 0), W - 1);
            atomicAdd(dzl + (((long long)n * H + yy) * W + xx) * C + c, v);
        }
    }
}

template <int CMAX, bool FWD, bool BWD, bool EXACT>
__global__ void __launch_bounds__(256)
tail_pixel_kernel(const float* __restrict__ zl, const int32_t* __restrict__ labels, const float* __restrict__ pw,
                  const float* __restrict__ nw, float eps, int N, int H, int W, int C, int f, float gscale,
                  float* __restrict__ loss_sum, float* __restrict__ dzl, int nby, int nbx) {
    extern __shared__ float sm[];
#define DLV3P_TILE(F_)                                                                                              \
    tail_pixel_tile<CMAX, FWD, BWD, EXACT, F_>(sm, zl, labels, pw, nw, eps, H, W, C, gscale, loss_sum, dzl, nby, nbx)
    switch (f) {                                    // block-uniform; launch_tail_pixel admits 2 / 4 / 8 / 16 only
    case 16: DLV3P_TILE(16); break;
    case 8: DLV3P_TILE(8); break;
    case 4: DLV3P_TILE(4); break;
    default: DLV3P_TILE(2); break;
    }
#undef DLV3P_TILE
}

template <bool FWD, bool BWD>
static int launch_tail_pixel(const float* zl, const int32_t* labels, const float* pw, const float* nw, float eps, int N,
                             int H, int W, int C, int f, float gscale, float* loss_sum, float* dzl, cudaStream_t st) {
    const int S = 16 / f, PS = S + 1;
    const int nby = (H * f + f / 2 + 15) / 16, nbx = (W * f + f / 2 + 15) / 16;
    const long long blocks = (long long)N * nby * nbx;
    DLV3P_REQUIRE(blocks < 0x7fffffffLL, DLV3P_ERR_SHAPE, "fused tail: too many tiles");
#define DLV3P_TAIL(CM)                                                                                             \
    do {                                                                                                           \
        constexpr int CS = (CM) | 1;                                                                               \
        const int smem = (PS * PS * CS + 96 + (BWD ? 256 * CS + 16 * S * 2 * CS : 0)) * 4;                     \
        static int configured = 0;                                                                                 \
        if (smem > configured) {                                                                                   \
            cudaError_t e = cudaFuncSetAttribute(tail_pixel_kernel<CM, FWD, BWD, false>,                           \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);               \
            if (e == cudaSuccess) e = cudaFuncSetAttribute(tail_pixel_kernel<CM, FWD, BWD, true>,                  \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, smem);               \
            DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "fused tail smem=%d: %s", smem, cudaGetErrorString(e)); \
            configured = smem;                                                                                     \
        }                                                                                                          \
        if (C == (CM))                                                                                             \
            tail_pixel_kernel<CM, FWD, BWD, true><<<(int)blocks, 256, smem, st>>>(zl, labels, pw, nw, eps, N, H, W, C, f, \
                                                                                  gscale, loss_sum, dzl, nby, nbx); \
        else                                                                                                       \
            tail_pixel_kernel<CM, FWD, BWD, false><<<(int)blocks, 256, smem, st>>>(zl, labels, pw, nw, eps, N, H, W, C, f, \
                                                                                   gscale, loss_sum, dzl, nby, nbx); \
    } while (0)
    if (C <= 8) DLV3P_TAIL(8);
    else if (C <= 16) DLV3P_TAIL(16);
    else if (C <= 21) DLV3P_TAIL(21);
    else if (C <= 24) DLV3P_TAIL(24);
    else DLV3P_TAIL(32);
#undef DLV3P_TAIL
    return check_launch("upsample_softmax_cbloss (pixel)");
}

static inline bool tail_fast_ok(int f) { return f == 2 || f == 4 || f == 8 || f == 16; }

static int warp_grid(long long work_items) {
    long long blocks = (work_items + 7) / 8;          // 8 warps per block
    const long long cap = (long long)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace dlv3p

using namespace dlv3p;

#define DLV3P_CHECK_PC(name)                                                                          \
    DLV3P_REQUIRE(P > 0 && C > 0 && C <= 32, DLV3P_ERR_SHAPE, name ": need P > 0 and 1 <= C <= 32 (C=%d)", C)

extern "C" int dlv3p_softmax_cbloss_fwd(const float* z, const int32_t* labels, const float* pw, const float* nw,
                                        float eps, int64_t P, int C, float* loss_sum, float* probs, void* stream) {
    DLV3P_CHECK_PC("softmax_cbloss_fwd");
    DLV3P_REQUIRE(z && labels && pw && nw && loss_sum, DLV3P_ERR_SHAPE, "softmax_cbloss_fwd: null pointer");
    softmax_cbloss_fwd_kernel<<<warp_grid(P), 256, 0, (cudaStream_t)stream>>>(z, labels, pw, nw, eps, P, C, loss_sum,
                                                                            probs);
    return check_launch("softmax_cbloss_fwd");
}

extern "C" int dlv3p_softmax_cbloss_bwd(const float* z, const int32_t* labels, const float* pw, const float* nw,
                                        float eps, int64_t P, int C, float grad_scale, float* dz, void* stream) {
    DLV3P_CHECK_PC("softmax_cbloss_bwd");
    DLV3P_REQUIRE(z && labels && pw && nw && dz, DLV3P_ERR_SHAPE, "softmax_cbloss_bwd: null pointer");
    softmax_cbloss_bwd_kernel<<<warp_grid(P), 256, 0, (cudaStream_t)stream>>>(z, labels, pw, nw, eps, P, C,
                                                                            grad_scale, dz);
    return check_launch("softmax_cbloss_bwd");
}

extern "C" int dlv3p_upsample_softmax_cbloss_fwd(const float* zl, const int32_t* labels, const float* pw,
                                                 const float* nw, float eps, int N, int H, int W, int C, int f,
                                                 float* loss_sum, void* stream) {
    DLV3P_REQUIRE(zl && labels && pw && nw && loss_sum, DLV3P_ERR_SHAPE, "upsample_softmax_cbloss_fwd: null pointer");
    DLV3P_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C <= 32 && f >= 1 && (f == 1 || f % 2 == 0), DLV3P_ERR_SHAPE,
                  "upsample_softmax_cbloss_fwd: need C <= 32 and an even (or unit) factor, got C=%d f=%d", C, f);
    if (tail_fast_ok(f))
        return launch_tail_pixel<true, false>(zl, labels, pw, nw, eps, N, H, W, C, f, 0.f, loss_sum, nullptr,
                                              (cudaStream_t)stream);
    const long long ntiles = (long long)N * (H + 1) * (W + 1);
    upsample_softmax_cbloss_kernel<false><<<warp_grid(ntiles), 256, 0, (cudaStream_t)stream>>>(
        zl, labels, pw, nw, eps, N, H, W, C, f, 0.f, loss_sum, nullptr, ntiles);
    return check_launch("upsample_softmax_cbloss_fwd");
}

extern "C" int dlv3p_upsample_softmax_cbloss_bwd(const float* zl, const int32_t* labels, const float* pw,
                                                 const float* nw, float eps, int N, int H, int W, int C, int f,
                                                 float grad_scale, float* dzl, void* stream) {
    DLV3P_REQUIRE(zl && labels && pw && nw && dzl, DLV3P_ERR_SHAPE, "upsample_softmax_cbloss_bwd: null pointer");
    DLV3P_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C <= 32 && f >= 1 && (f == 1 || f % 2 == 0), DLV3P_ERR_SHAPE,
                  "upsample_softmax_cbloss_bwd: need C <= 32 and an even (or unit) factor, got C=%d f=%d", C, f);
    if (tail_fast_ok(f))
        return launch_tail_pixel<false, true>(zl, labels, pw, nw, eps, N, H, W, C, f, grad_scale, nullptr, dzl,
                                              (cudaStream_t)stream);
    const long long ntiles = (long long)N * (H + 1) * (W + 1);
    upsample_softmax_cbloss_kernel<true><<<warp_grid(ntiles), 256, 0, (cudaStream_t)stream>>>(
        zl, labels, pw, nw, eps, N, H, W, C, f, grad_scale, nullptr, dzl, ntiles);
    return check_launch("upsample_softmax_cbloss_bwd");
}

extern "C" int dlv3p_upsample_softmax_cbloss_fwd_bwd(const float* zl, const int32_t* labels, const float* pw,
                                                     const float* nw, float eps, int N, int H, int W, int C, int f,
                                                     float grad_scale, float* loss_sum, float* dzl, void* stream) {
    DLV3P_REQUIRE(zl && labels && pw && nw && loss_sum && dzl, DLV3P_ERR_SHAPE,
                  "upsample_softmax_cbloss_fwd_bwd: null pointer");
    DLV3P_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C <= 32 && f >= 1 && (f == 1 || f % 2 == 0), DLV3P_ERR_SHAPE,
                  "upsample_softmax_cbloss_fwd_bwd: need C <= 32 and an even (or unit) factor, got C=%d f=%d", C, f);
    if (tail_fast_ok(f))
        return launch_tail_pixel<true, true>(zl, labels, pw, nw, eps, N, H, W, C, f, grad_scale, loss_sum, dzl,
                                             (cudaStream_t)stream);
    const long long ntiles = (long long)N * (H + 1) * (W + 1);
    upsample_softmax_cbloss_kernel<false><<<warp_grid(ntiles), 256, 0, (cudaStream_t)stream>>>(
        zl, labels, pw, nw, eps, N, H, W, C, f, 0.f, loss_sum, nullptr, ntiles);
    upsample_softmax_cbloss_kernel<true><<<warp_grid(ntiles), 256, 0, (cudaStream_t)stream>>>(
        zl, labels, pw, nw, eps, N, H, W, C, f, grad_scale, nullptr, dzl, ntiles);
    return check_launch("upsample_softmax_cbloss_fwd_bwd");
}

extern "C" int dlv3p_softmax_argmax(const float* z, int64_t P, int C, float* probs, int32_t* labels, void* stream) {
    DLV3P_CHECK_PC("softmax_argmax");
    DLV3P_REQUIRE(z && (probs || labels), DLV3P_ERR_SHAPE, "softmax_argmax: null pointer");
    softmax_argmax_kernel<<<warp_grid(P), 256, 0, (cudaStream_t)stream>>>(z, P, C, probs, labels);
    return check_launch("softmax_argmax");
}

extern "C" int dlv3p_cbloss_dense_fwd(const float* y_true, const float* y_pred, const float* pw, const float* nw,
                                      float eps, int64_t P, int C, float* loss_sum, void* stream) {
    DLV3P_REQUIRE(y_true && y_pred && pw && nw && loss_sum && P > 0 && C > 0, DLV3P_ERR_SHAPE,
                  "cbloss_dense_fwd: bad arguments");
    int blocks = cdiv(P * C, 256 * 4); if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    cbloss_dense_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(y_true, y_pred, pw, nw, eps, P, C, loss_sum);
    return check_launch("cbloss_dense_fwd");
}

extern "C" int dlv3p_cbloss_dense_bwd(const float* y_true, const float* y_pred, const float* pw, const float* nw,
                                      float eps, int64_t P, int C, float grad_scale, float* dy_pred, void* stream) {
    DLV3P_REQUIRE(y_true && y_pred && pw && nw && dy_pred && P > 0 && C > 0, DLV3P_ERR_SHAPE,
                  "cbloss_dense_bwd: bad arguments");
    cbloss_dense_bwd_kernel<<<cdiv(P * C, 256), 256, 0, (cudaStream_t)stream>>>(y_true, y_pred, pw, nw, eps, P, C,
                                                                              grad_scale, dy_pred);
    return check_launch("cbloss_dense_bwd");
}

extern "C" int dlv3p_softmax_bwd(const float* p, const float* dp, int64_t P, int C, float* dz, void* stream) {
    DLV3P_CHECK_PC("softmax_bwd");
    DLV3P_REQUIRE(p && dp && dz, DLV3P_ERR_SHAPE, "softmax_bwd: null pointer");
    softmax_bwd_kernel<<<warp_grid(P), 256, 0, (cudaStream_t)stream>>>(p, dp, P, C, dz);
    return check_launch("softmax_bwd");
}

extern "C" int dlv3p_confusion_matrix(const int32_t* y_true, const int32_t* y_pred, int64_t P, int C, double* cm,
                                      void* stream) {
    DLV3P_REQUIRE(y_true && y_pred && cm && P > 0 && C > 0 && C <= 64, DLV3P_ERR_SHAPE,
                  "confusion_matrix: bad arguments");
    int blocks = cdiv(P, 256 * 16); if (blocks > kNumSMs * 4) blocks = kNumSMs * 4;
    confusion_kernel<<<blocks, 256, C * C * sizeof(unsigned int), (cudaStream_t)stream>>>(y_true, y_pred, P, C, cm);
    return check_launch("confusion_matrix");
}
