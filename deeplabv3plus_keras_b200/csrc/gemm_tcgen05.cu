// K2 — pointwise / projection convolutions as bf16 tcgen05 tensor-core GEMMs (sm_100a).
//
// Replaces TF Conv2D(1x1) (cuDNN/cuBLAS in the reference's TF 2.4 runtime) for the pointwise half of every
// SeparableConv2D, the ASPP / decoder projections (ss.py:814-818,833-838,843-847,865-869,931-935) and, after
// im2col, the dense 3x3 convs (ss.py:893-897).
//
// Persistent kernel: one CTA per SM (192 threads) walks a static list of 128 x BLOCK_N output tiles.
//   warp 0, one lane : TMA producer  — cp.async.bulk.tensor 2D tiles (128B swizzle) into a 4-stage smem ring that
//                                      runs ahead across tile boundaries; completion on mbarriers (expect_tx)
//   warp 1, one lane : MMA issuer    — tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N, K=16 per instruction,
//                                      fp32 accumulators in TMEM, TWO accumulator stages (2*BLOCK_N columns) so the
//                                      epilogue of tile i overlaps the main loop of tile i+1
//   warps 2..5       : epilogue      — tcgen05.ld 32x32b (one accumulator row per lane) -> fused per-column
//                                      scale/shift (folded BatchNorm) + ReLU/ReLU6 -> bf16 rows into a 128B-swizzled
//                                      smem staging tile (32 rows x 64 cols, double buffered per warp) -> ONE TMA
//                                      tensor store per tile (full-line coalesced writes, OOB clipping); training-mode
//                                      BatchNorm statistics (per-column sum / sum of squares) are read back
//                                      column-per-lane from the same staging tile (conflict-free, shuffle-free).
//                                      The filter gradient stages fp32 rows and issues TMA reduce-adds
//                                      (cp.reduce.async.bulk.tensor ... add.f32) instead of per-element REDs.
//                                      [A lane-per-row direct store touches 32 different 128 B lines per STG: ncu
//                                      showed the K=64..288 GEMMs bound by exactly that, profiles/r1_*]
//                                      Fallback (fp32 C, residual addend, unaligned ldc): direct stores / REDs.
// Forward / input-gradient use K-major operands (A[M,K], B[N,K], K contiguous).  The filter gradient
// dW = X^T dY contracts over the pixel axis, which is NOT contiguous in NHWC: both operands are fed MN-major
// (64-element x 64-row TMA boxes, same 128B swizzle) so no transposed copy of the activations is ever written;
// its work list is (pixel-range split) x (tile), split-major so that concurrently running CTAs share operands in L2.
#include <stdlib.h>
#include <string.h>

#include "tma.cuh"

namespace dlv3p {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // 64 bf16 = 128 B = one swizzle atom row
constexpr int kEpiWarps = 8;             // two per TMEM lane quadrant: they split the column chunks of a tile
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kEpiStageFloats = 32 * 33;          // fallback path, per epilogue warp: 32 rows x 32 cols, padded
constexpr int kEpiBufBytes = 32 * 128;            // TMA-store path: one staging tile = 32 rows x 128 B
constexpr int kEpiBytes = kEpiWarps * 2 * kEpiBufBytes;   // per epilogue warp 2 buffers (also holds the fallback transposes)
constexpr int kMaxStatCols = 1024;                // per-CTA smem accumulators for the BN column statistics
// The statistics region holds either ONE [2][kMaxStatCols] table updated with shared-memory atomics (mode 1) or, when the
// columns fit, one PRIVATE [2][SC] table per TMEM lane quadrant (mode 2): the same lane always owns the same column of
// its quadrant's table, so the per-tile update is a plain load-add-store.  atomicAdd(float) on shared memory has no
// native instruction (SASS: ATOMS.CAST.SPIN loops behind a generic-address dispatch); four quadrant warps hammering the
// same columns made the statistics the slowest part of the short-K GEMMs' epilogue.
constexpr int kStatFloats1 = 2 * kMaxStatCols;    // 1-CTA kernels: 8 KB  -> 4 private tables of 256 columns
constexpr int kStatFloats2 = 8 * kMaxStatCols;    // 2-CTA kernel: 32 KB -> 4 private tables of 1024 columns
__host__ __device__ constexpr int stat_capacity(int region_floats, int slots) { return region_floats / (2 * slots); }
static_assert(kEpiBytes >= 4 * kEpiStageFloats * 4, "staging area must hold the fallback transposes");
template <int BLOCK_N> struct GemmCfg {
    static constexpr int kStageBytes = kBlockM * kBlockK * 2 + BLOCK_N * kBlockK * 2;
    // operand ring: 3 x 48 KB (N tile 256), 4 x 32 KB (128), 6 x 24 KB (64), 7 x 20 KB (32).  The narrow tiles are the
    // K-deep, latency-bound ones (the 21-class logits convolution: 45 k-blocks of 20 KB per 128-pixel tile, tensor pipe
    // 8 % active, nothing saturated in the ncu capture profiles/r2_new_kernels_ncu.md): more k-blocks in flight
    static constexpr int kStages = BLOCK_N == 256 ? 3 : (BLOCK_N == 128 ? 4 : (BLOCK_N == 64 ? 6 : 7));
    static constexpr int kSmem = kStages * kStageBytes + kEpiBytes + 2 * kMaxStatCols * 4 + 1024 /*align*/ + 256 /*barriers*/;
    static_assert(kSmem <= 232448, "shared memory budget");
};

struct GemmParams {
    int M, N, K;                       // logical GEMM extents (for WGRAD: rows=K(cin), cols=N(cout), reduction=M)
    void* C; long long ldc; int c_dtype;
    const float* col_scale; const float* col_shift; int act;
    const void* addend; long long ld_add;
    float* col_stats;
    int kb_per_split;                  // WGRAD: reduction blocks (of 64 rows) handled by one work item
    int splits;                        // WGRAD: number of pixel-range splits
    int n_tiles, m_tiles;              // output tile grid
    int tma_store;                     // epilogue through the smem staging tile + TMA store / reduce-add
    // implicit-GEMM 3x3 VALID stride-1 convolution (no im2col buffer): 0 = plain GEMM, 1 = forward, 2 = input gradient,
    // 3 = filter gradient.  Output/M tiles never cross an image row; A comes straight from the NHWC tensor.
    // 3x3 SAME stride-1 convolution (the logits / boundary-refinement convolution, ss.py:893-897): 4 = forward, 5 = input
    // gradient, 6 = filter gradient — every tap is a rank-4 window (channel block, 128 / 64 pixels of one row, row, image)
    // shifted by the tap; what falls outside the image is zero-filled by TMA, which IS the SAME padding.
    int conv_mode;
    int cv_rows_in;                    // image rows of the A-side tensor per image (fwd/wgrad: H of x; dgrad: Ho of dy)
    int cv_rows_out;                   // rows per image of the M-side index (fwd/wgrad: Ho; dgrad: H)
    int cv_width;                      // filter gradient: Wo (pixel index of a dy reduction block)
    int cv_rlimit;                     // epilogue clip inside a C "row": fwd Wo, dgrad W, wgrad 3*Cin
    int cv_tpr;                        // fwd/dgrad: 128-pixel M tiles per row; wgrad: 64-pixel reduction blocks per row
    int cv_kbr;                        // fwd/wgrad: 64-element k-blocks per filter row; dgrad: 64-channel blocks per tap
    int cv_run;                        // fwd: 3*Cin, the length of one filter row's run in the [Cout, 9*Cin] filter matrix
    int dbg;                           // DLV3P_DIAG builds only (libdlv3p_diag.so): 1 = no operand loads, 2 = no epilogue work, 4 = no MMAs, 8 = no C stores, 16 = no TMEM reads
};

// Bottleneck-decomposition switches exist only in the diagnostics build (csrc/build.sh with DLV3P_DIAG=1 ->
// libdlv3p_diag.so, used by scripts/gemm_decompose.py); the shipped library has no path that skips work.
#ifdef DLV3P_DIAG
#define DLV3P_DBG(p, bit) ((p).dbg & (bit))
#else
#define DLV3P_DBG(p, bit) false
#endif

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

// decode a work item: forward -> output tile (n fastest, so neighbouring CTAs share the A rows in L2);
// filter gradient -> (split, tile), split-major
template <int BLOCK_N, bool WGRAD>
__device__ __forceinline__ void decode_work(const GemmParams& p, int w, int& row0, int& col0, int& kb_begin,
                                            int& kb_end) {
    const int tiles = p.n_tiles * p.m_tiles;
    const int tile = WGRAD ? (w % tiles) : w;
    row0 = (tile / p.n_tiles) * kBlockM;
    col0 = (tile % p.n_tiles) * BLOCK_N;
    if (WGRAD) {
        const int total_kb = (p.M + kBlockK - 1) / kBlockK;
        kb_begin = (w / tiles) * p.kb_per_split;
        kb_end = min(kb_begin + p.kb_per_split, total_kb);
    } else {
        kb_begin = 0;
        kb_end = (p.K + kBlockK - 1) / kBlockK;
    }
}

// ---- staged epilogue of ONE accumulator tile (TMEM -> registers -> swizzled smem tile -> TMA store / reduce-add) ----
// `acc_tmem` = TMEM address of this warp's lane quadrant at the accumulator stage; `rbase` = first output row of the
// warp; `empty_bar` is arrived on once every TMEM read of the tile has completed (REMOTE: it is a shared::cluster
// address in the leader CTA of a 2-CTA pair).
// SC: 0 = per-column scale / shift (if any) fetched with global loads (the implicit-convolution kernels), 1 = from the
// CTA's shared-memory table `sc_tab`, 2 = the launch has none (compiled out: the training GEMMs)
template <int BLOCK_N, bool WGRAD, bool REMOTE, int SC = 0>
__device__ __forceinline__ void staged_tile_epilogue(const GemmParams& p, const CUtensorMap* tmC, uint32_t acc_tmem,
                                                     int rbase, int col0, int lane, int half, uint32_t stg0, uint32_t& buf,
                                                     uint32_t empty_bar, float* stat_smem, int stat_mode, int stat_slot,
                                                     int stat_cap, int row_limit, int c2, const float* sc_tab = nullptr) {
    constexpr int CW = WGRAD ? 32 : 64;                 // columns per staging tile (128-byte rows)
    const uint32_t lane_row = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    const int ncols = min(BLOCK_N, p.N - col0);
    const int n_chunks = (ncols + CW - 1) / CW;
    if (half >= n_chunks) {
        // the two warps of a lane quadrant take alternate column chunks; this one has none in this tile
        if (lane == 0) {
            if (REMOTE) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(empty_bar) : "memory");
            else mbar_arrive(empty_bar);
        }
        return;
    }
#pragma unroll 1
    for (int ch = half; ch < n_chunks; ch += 2) {
        const int n_base = col0 + ch * CW;
        const uint32_t stg = stg0 + buf * kEpiBufBytes;
        if (lane == 0) tma_wait_group_read<1>();    // the store that last read this buffer has drained it
        __syncwarp();
#pragma unroll
        for (int h = 0; h < CW / 32; ++h) {
            uint32_t raw[32];
            if (DLV3P_DBG(p, 16)) {
#pragma unroll
                for (int j = 0; j < 32; ++j) raw[j] = 0x3f800000u + (uint32_t)(lane + j);
            } else {
                tc_ld_32x32b_x32(acc_tmem + (uint32_t)(ch * CW + h * 32), raw);
            }
            if (ch + 2 >= n_chunks && h == CW / 32 - 1) {
                // every TMEM read of this warp in this accumulator stage has completed: hand it back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (REMOTE) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(empty_bar) : "memory");
                    else mbar_arrive(empty_bar);
                }
            }
            if (WGRAD) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane_row + (((uint32_t)j ^ sw) << 4)),
                                 "r"(raw[4 * j]), "r"(raw[4 * j + 1]), "r"(raw[4 * j + 2]), "r"(raw[4 * j + 3]) : "memory");
            } else {
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                if (SC == 1) {
                    // per-column scale / shift (folded BatchNormalization) from the CTA's shared-memory table: 16 broadcast
                    // LDS.128 per 32 columns instead of 64 global loads per lane (the ncu source page of the MobileNetV2
                    // projections showed the epilogue warps serialised on them, profiles/r2_skinny_gemm.md)
                    const float4* sc4 = reinterpret_cast<const float4*>(sc_tab + n_base + h * 32);
                    const float4* sh4 = reinterpret_cast<const float4*>(sc_tab + kMaxStatCols + n_base + h * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 a = sc4[j], b = sh4[j];
                        v[4 * j] = fmaf(v[4 * j], a.x, b.x); v[4 * j + 1] = fmaf(v[4 * j + 1], a.y, b.y);
                        v[4 * j + 2] = fmaf(v[4 * j + 2], a.z, b.z); v[4 * j + 3] = fmaf(v[4 * j + 3], a.w, b.w);
                    }
                } else if (SC == 0 && p.col_scale != nullptr) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int n = min(n_base + h * 32 + j, p.N - 1);
                        v[j] = fmaf(v[j], __ldg(p.col_scale + n), __ldg(p.col_shift + n));
                    }
                }
                if (p.act != DLV3P_ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], p.act);
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        __nv_bfloat162 pr = __floats2bfloat162_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
                        o[e] = *reinterpret_cast<uint32_t*>(&pr);
                    }
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane_row + (((uint32_t)(h * 4 + g) ^ sw) << 4)),
                                 "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
                }
            }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && rbase < row_limit && !DLV3P_DBG(p, 8)) {
            // c2 >= 0: implicit-convolution tile, rank-3 C map (channel, column in the row, row) clips the ragged row end
            if (WGRAD) { if (c2 >= 0) tma_reduce_add_3d(tmC, stg, n_base, rbase, c2); else tma_reduce_add_2d(tmC, stg, n_base, rbase); }
            else { if (c2 >= 0) tma_store_3d(tmC, stg, n_base, rbase, c2); else tma_store_2d(tmC, stg, n_base, rbase); }
            tma_commit_group();
        }
        if (!WGRAD && p.col_stats != nullptr) {
            // BatchNormalization batch statistics of the STORED (bf16) conv output: lane L owns columns
            // n_base+2L, n_base+2L+1 and walks the 32 rows of the staging tile (one 32-bit word per row,
            // all 32 banks distinct); rows >= M hold exact zeros (TMA zero-filled A)
            float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
            const uint32_t cj = (uint32_t)(lane >> 2), cw = (uint32_t)(lane & 3) << 2;
            // (the halo-staged convolution kernel computes real values for the pixels beyond a ragged row end: they are
            // clipped from the store and must be kept out of the statistics as well)
            const int rv = row_limit - rbase;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
                uint32_t word;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(word)
                             : "r"(stg + (uint32_t)rr * 128u + ((cj ^ (uint32_t)(rr & 7)) << 4) + cw));
                if (rr >= rv) word = 0u;
                const float a = __uint_as_float(word << 16), b = __uint_as_float(word & 0xffff0000u);
                s1a += a; s2a = fmaf(a, a, s2a);
                s1b += b; s2b = fmaf(b, b, s2b);
            }
            const int col = n_base + 2 * lane;
            if (stat_mode == 2) {
                float* mine = stat_smem + stat_slot * 2 * stat_cap;          // this quadrant's private table
                if (col < p.N) { mine[col] += s1a; mine[stat_cap + col] += s2a; }
                if (col + 1 < p.N) { mine[col + 1] += s1b; mine[stat_cap + col + 1] += s2b; }
            } else if (stat_mode == 1) {
                if (col < p.N) { atomicAdd(stat_smem + col, s1a); atomicAdd(stat_smem + kMaxStatCols + col, s2a); }
                if (col + 1 < p.N) { atomicAdd(stat_smem + col + 1, s1b); atomicAdd(stat_smem + kMaxStatCols + col + 1, s2b); }
            } else {
                if (col < p.N) { atomicAdd(p.col_stats + col, s1a); atomicAdd(p.col_stats + p.N + col, s2a); }
                if (col + 1 < p.N) { atomicAdd(p.col_stats + col + 1, s1b); atomicAdd(p.col_stats + p.N + col + 1, s2b); }
            }
        }
        buf ^= 1u;
    }
}

// Per-column epilogue operands (col_scale / col_shift = the folded BatchNormalization of an inference GEMM) staged once
// per CTA in the statistics region ([kMaxStatCols] scale, then [kMaxStatCols] shift; the region is free because a GEMM
// either folds a BN or collects its batch statistics, never both).  Columns past N read as scale 0 / shift 0.  Called by
// ALL threads right after the programmatic-dependency wait (the vectors may come from the preceding dlv3p_bn_fold).
__device__ __forceinline__ const float* fill_scale_table(const GemmParams& p, float* stat_smem, int nthreads) {
    const int n_pad = min((p.N + 255) & ~255, kMaxStatCols);        // whole 256-column tiles are read unpredicated
    for (int i = threadIdx.x; i < n_pad; i += nthreads) {
        stat_smem[i] = i < p.N ? __ldg(p.col_scale + i) : 0.f;
        stat_smem[kMaxStatCols + i] = i < p.N ? __ldg(p.col_shift + i) : 0.f;
    }
    __syncthreads();
    return stat_smem;
}

// SC (see staged_tile_epilogue): 2 = no col_scale / col_shift (the training GEMMs: compiled out), 1 = from the
// shared-memory table (inference GEMM with a folded BatchNormalization), 0 = global loads (scale together with col_stats,
// or N > kMaxStatCols: the table lives in the statistics region)
template <int BLOCK_N, bool WGRAD, int SC>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    constexpr int A_BYTES = kBlockM * kBlockK * 2;
    constexpr int B_BYTES = BLOCK_N * kBlockK * 2;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int kStages = GemmCfg<BLOCK_N>::kStages;
    // accumulator stages in TMEM: two 256 / 128-column tiles, four 64-column or eight 32-column ones (256 columns): the
    // narrow-N GEMMs (MobileNetV2's 16..96-channel projections, the 21-class logits convolution) have ONE k-block per tile,
    // so a tile lives ~3 us of pure latency (TMA -> MMA -> TMEM read -> store); with two stages the issuer stalls on the
    // epilogue every other tile, with eight the chain is pipelined and the kernel becomes bandwidth-bound
    constexpr uint32_t kAcc = BLOCK_N >= 128 ? 2u : (BLOCK_N == 64 ? 4u : 8u);
    constexpr uint32_t TMEM_COLS = kAcc * BLOCK_N;         // 256 .. 512, power of two
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* epi_bytes = smem + kStages * STAGE_BYTES;                       // 1024-aligned (stage sizes are)
    float* epi_stage = reinterpret_cast<float*>(epi_bytes);
    float* stat_smem = reinterpret_cast<float*>(epi_bytes + kEpiBytes);      // [2][kMaxStatCols]
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_bytes + kEpiBytes + 2 * kMaxStatCols * 4);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_bar = smem_u32(bars);                      // kStages
    const uint32_t empty_bar = smem_u32(bars + kStages);           // kStages
    const uint32_t tmem_full_bar = smem_u32(bars + 2 * kStages);          // kAcc
    const uint32_t tmem_empty_bar = smem_u32(bars + 2 * kStages + kAcc);  // kAcc
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2 * kAcc);
    static_assert((2 * kStages + 2 * kAcc + 1) * 8 <= 256, "barrier area");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();

    if (!WGRAD && p.col_stats != nullptr)
        for (int i = threadIdx.x; i < 2 * kMaxStatCols; i += kThreads) stat_smem[i] = 0.f;
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        for (int s = 0; s < (int)kAcc; ++s) { mbar_init(tmem_full_bar + 8 * s, 1); mbar_init(tmem_empty_bar + 8 * s, kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    pdl_wait();                              // prologue done; the operands may still be in flight from the previous kernel
    const float* sc_tab = SC == 1 ? fill_scale_table(p, stat_smem, kThreads) : nullptr;

    const int num_work = p.n_tiles * p.m_tiles * (WGRAD ? p.splits : 1);

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            uint32_t it = 0;                                   // ring position, continues across tiles
            for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
                int row0, col0, kb_begin, kb_end;
                decode_work<BLOCK_N, WGRAD>(p, w, row0, col0, kb_begin, kb_end);
                for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                    const uint32_t s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1u;
                    mbar_wait(empty_bar + 8 * s, ph ^ 1u);
                    const uint32_t a_dst = smem_base + s * STAGE_BYTES;
                    const uint32_t b_dst = a_dst + A_BYTES;
                    const uint32_t fb = full_bar + 8 * s;
                    if (DLV3P_DBG(p, 1)) { mbar_arrive(fb); continue; }
                    mbar_expect_tx(fb, STAGE_BYTES);
                    const int kk = kb * kBlockK;
                    if (p.conv_mode != 0) {
                        if (WGRAD && p.conv_mode == 6) {
                            // SAME filter gradient: output rows = (tap, 128 input channels), reduction block kb = 64
                            // consecutive output pixels of one image row; x window shifted by the tap
                            const int mt = row0 / kBlockM;
                            const int tap = mt / p.cv_kbr, ci0 = (mt % p.cv_kbr) * kBlockM;
                            const int r = kb / p.cv_tpr, qd = kb % p.cv_tpr;
                            const int n = r / p.cv_rows_out, ho = r % p.cv_rows_out;
#pragma unroll
                            for (int h = 0; h < kBlockM / 64; ++h)
                                tma_load_4d(a_dst + h * 8192, &tmA, fb, ci0 + 64 * h, qd * 64 + tap % 3 - 1, ho + tap / 3 - 1, n);
                            // dy through a rank-3 map (channel, pixel of the row, row): pixels past the end of the row are
                            // zero-filled — the shifted x window is NOT zero there (it still covers real pixels)
#pragma unroll
                            for (int h = 0; h < BLOCK_N / 64; ++h)
                                tma_load_3d(b_dst + h * 8192, &tmB, fb, col0 + 64 * h, qd * 64, r);
                        } else if (WGRAD) {
                            // filter gradient: output rows = the 128-element padded K-run of filter row i, reduction
                            // block kb = 64 consecutive output pixels of one image row
                            const int i = row0 / kBlockM;
                            const int r = kb / p.cv_tpr, qd = kb % p.cv_tpr;
                            const int n = r / p.cv_rows_out, ho = r % p.cv_rows_out;
#pragma unroll
                            for (int h = 0; h < kBlockM / 64; ++h)
                                tma_load_3d(a_dst + h * 8192, &tmA, fb, 64 * h, qd * 64, n * p.cv_rows_in + ho + i);
#pragma unroll
                            for (int h = 0; h < BLOCK_N / 64; ++h)
                                tma_load_2d(b_dst + h * 8192, &tmB, fb, col0 + 64 * h, r * p.cv_width + qd * 64);
                        } else {
                            const int mt = row0 / kBlockM;
                            const int r = mt / p.cv_tpr, w0 = (mt % p.cv_tpr) * kBlockM;
                            const int n = r / p.cv_rows_out, hh = r % p.cv_rows_out;
                            if (p.conv_mode == 1) {
                                // forward: k-block = 64 elements of the 3*Cin-element run under filter row i (elements
                                // beyond the run and columns beyond the row are zero-filled by TMA)
                                const int i = kb / p.cv_kbr, part = kb % p.cv_kbr;
                                tma_load_3d(a_dst, &tmA, fb, part * 64, w0, n * p.cv_rows_in + hh + i);
                                // filter columns of the same 64 run elements; what the box takes beyond the run belongs
                                // to the next filter row (or is zero-filled) and meets the zero-filled A elements
                                tma_load_2d(b_dst, &tmB, fb, i * p.cv_run + part * 64, col0);
                            } else if (p.conv_mode == 2) {
                                // input gradient: k-block = 64 output channels of dy under tap (i, j), window shifted by
                                // (-i, -j); positions outside dy are zero-filled (full correlation)
                                const int tap = kb / p.cv_kbr, part = kb % p.cv_kbr;
                                tma_load_4d(a_dst, &tmA, fb, part * 64, w0 - tap % 3, hh - tap / 3, n);
                                tma_load_2d(b_dst, &tmB, fb, kk, col0);
                            } else {
                                // SAME forward (4) / input gradient (5): k-block = 64 channels of the operand image under
                                // tap (i, j), window shifted by +-(i-1, j-1); outside the image = zero = the SAME padding.
                                // B: the tap's slice of the filter matrix (cv_run K-elements per tap); a channel block
                                // that runs past the tap's channels meets zero-filled A columns
                                const int tap = kb / p.cv_kbr, part = kb % p.cv_kbr;
                                const int sg = p.conv_mode == 4 ? 1 : -1;
                                tma_load_4d(a_dst, &tmA, fb, part * 64, w0 + sg * (tap % 3 - 1), hh + sg * (tap / 3 - 1), n);
                                tma_load_2d(b_dst, &tmB, fb, tap * p.cv_run + part * 64, col0);
                            }
                        }
                    } else if (WGRAD) {
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
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            constexpr uint32_t idesc = (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ |
                                       ((WGRAD ? 1u : 0u) << 15) | ((WGRAD ? 1u : 0u) << 16) |
                                       ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
            uint32_t it = 0, t = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++t) {
                int row0, col0, kb_begin, kb_end;
                decode_work<BLOCK_N, WGRAD>(p, w, row0, col0, kb_begin, kb_end);
                const uint32_t as = t % kAcc;                  // accumulator stage
                mbar_wait(tmem_empty_bar + 8 * as, ((t / kAcc) & 1u) ^ 1u);   // epilogue has drained this stage
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BLOCK_N;
                for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
                    const uint32_t s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1u;
                    mbar_wait(full_bar + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t a_src = smem_base + s * STAGE_BYTES;
                    const uint32_t b_src = a_src + A_BYTES;
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        uint64_t ad, bd;
                        if (WGRAD) {
                            // MN-major SW128: LBO = next 64-element MN chunk (one 8 KB box), SBO = next 8 K-rows
                            // (1 KB); a K=16 step is 16 rows of 128 B
                            ad = umma_desc(a_src + k * 2048, 8192, 1024);
                            bd = umma_desc(b_src + k * 2048, 8192, 1024);
                        } else {
                            // K-major SW128: SBO = next 8 rows (1 KB); a K=16 step is 32 B inside the swizzle atom
                            ad = umma_desc(a_src + k * 32, 16, 1024);
                            bd = umma_desc(b_src + k * 32, 16, 1024);
                        }
                        if (!DLV3P_DBG(p, 4)) tc_mma_bf16(d_tmem, ad, bd, idesc, (kb > kb_begin || k > 0) ? 1u : 0u);
                    }
                    tc_commit(empty_bar + 8 * s);              // smem slot reusable once these MMAs retire
                }
                tc_commit(tmem_full_bar + 8 * as);             // accumulator stage complete
            }
        }
    } else {
        // ===== epilogue: warps 2..9, TMEM lane quadrant = warp % 4, two warps per quadrant =====
        const int q = warp & 3;
        const int ew = warp - 2, half = ew >> 2;
        float* stage = epi_stage + q * kEpiStageFloats;
        const int row_limit = WGRAD ? p.K : p.M;
        constexpr int kCap1 = stat_capacity(kStatFloats1, 4);
        const bool use_smem_stats = (!WGRAD) && (p.col_stats != nullptr) && (p.N <= kMaxStatCols);
        const int stat_mode = !use_smem_stats ? 0 : (p.N <= kCap1 ? 2 : 1);
        if (p.tma_store) {
            const uint32_t stg0 = smem_u32(epi_bytes) + (uint32_t)ew * 2u * kEpiBufBytes;
            uint32_t buf = 0, t = 0;
            for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++t) {
                int row0, col0, kb_begin, kb_end;
                decode_work<BLOCK_N, WGRAD>(p, w, row0, col0, kb_begin, kb_end);
                const uint32_t as = t % kAcc;
                mbar_wait(tmem_full_bar + 8 * as, (t / kAcc) & 1u);
                tc_fence_after();
                if (DLV3P_DBG(p, 2)) { if (lane == 0) mbar_arrive(tmem_empty_bar + 8 * as); continue; }
                int rb = row0 + q * 32, rl = row_limit, c2 = -1;
                if (p.conv_mode != 0) {
                    const int mt = row0 / kBlockM;
                    if (WGRAD && p.conv_mode == 6) { rb = (mt % p.cv_kbr) * kBlockM + q * 32; rl = p.cv_rlimit; c2 = mt / p.cv_kbr; }   // (cout, cin, tap)
                    else if (WGRAD) { rb = q * 32; rl = p.cv_rlimit; c2 = mt; }          // (cout, element of the run, filter row)
                    else { rb = (mt % p.cv_tpr) * kBlockM + q * 32; rl = p.cv_rlimit; c2 = mt / p.cv_tpr; }   // (ch, column, row)
                }
                staged_tile_epilogue<BLOCK_N, WGRAD, false, SC>(p, &tmC, tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N,
                                                            rb, col0, lane, half, stg0, buf,
                                                            tmem_empty_bar + 8 * as, stat_smem, stat_mode, q, kCap1, rl, c2,
                                                            sc_tab);
            }
            if (lane == 0) tma_wait_group_read<0>();      // smem may be released; the writes drain before the grid completes
        } else {
        uint32_t t = 0;
        for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++t) {
            int row0, col0, kb_begin, kb_end;
            decode_work<BLOCK_N, WGRAD>(p, w, row0, col0, kb_begin, kb_end);
            const uint32_t as = t % kAcc;
            mbar_wait(tmem_full_bar + 8 * as, (t / kAcc) & 1u);
            tc_fence_after();
            if (DLV3P_DBG(p, 2)) { if (lane == 0) mbar_arrive(tmem_empty_bar + 8 * as); continue; }
            // (the statistics / filter-gradient variants share per-quadrant tables and a staging buffer: one warp only)
            const int owner = (WGRAD || p.col_stats != nullptr) ? 0 : (int)(t & 1u);
            if (half != owner) {
                // direct-store fallback: ONE warp per lane quadrant per tile, the two warps of a quadrant take alternate
                // tiles (the narrow-N GEMMs that come here are epilogue-paced: scripts/skinny_decompose.py)
                if (lane == 0) mbar_arrive(tmem_empty_bar + 8 * as);
                continue;
            }
            int rbase = row0 + q * 32;                         // first output row of this warp
            int rlim = row_limit;
            if (p.conv_mode != 0) {
                // implicit-convolution tiles: the tile index is (image row, 128-pixel block) / (tap, 128-channel block);
                // rows past the end of the image row / of the tap's channels are padding of the tile
                const int mt = row0 / kBlockM;
                if (WGRAD) {
                    const int tap = mt / p.cv_kbr;
                    rbase = tap * p.cv_rlimit + (mt % p.cv_kbr) * kBlockM + q * 32;
                    rlim = (tap + 1) * p.cv_rlimit;
                } else {
                    const int line = mt / p.cv_tpr;
                    rbase = line * p.cv_width + (mt % p.cv_tpr) * kBlockM + q * 32;
                    rlim = line * p.cv_width + p.cv_rlimit;
                }
            }
            const int r = rbase + lane;                        // output row owned by this lane (row-per-lane layout)
            const bool row_ok = r < rlim;
#pragma unroll 1
            for (int ch = 0; ch < BLOCK_N / 32; ++ch) {
                const int n_base = col0 + ch * 32;
                const bool last_chunk = (ch == BLOCK_N / 32 - 1) || (n_base + 32 >= p.N);
                uint32_t raw[32];
                tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N + (uint32_t)(ch * 32), raw);
                if (last_chunk) {
                    // every TMEM read of this accumulator stage has completed: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tmem_empty_bar + 8 * as);
                }
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);

                if (WGRAD || p.col_stats != nullptr) {
                    // 32x33 smem transpose (row-per-lane -> column-per-lane): coalesced fp32 REDs for the filter
                    // gradient, shuffle-free column sums for the BatchNormalization statistics
#pragma unroll
                    for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = v[j];
                    __syncwarp();
                    const int col = n_base + lane;
                    if (WGRAD) {
                        if (col < p.N) {
                            float* dst = reinterpret_cast<float*>(p.C) + (long long)rbase * p.ldc + col;
#pragma unroll 8
                            for (int rr = 0; rr < 32; ++rr)
                                if (rbase + rr < rlim) atomicAdd(dst + (long long)rr * p.ldc, stage[rr * 33 + lane]);
                        }
                    } else {
                        // rows beyond M were zero-filled by TMA and add nothing; the ghost pixels past the end of an image
                        // row of a SAME-convolution tile are NOT zero (their shifted windows cover real pixels): skipped
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
                        for (int rr = 0; rr < 32; ++rr) {
                            const float u = (rbase + rr < rlim) ? stage[rr * 33 + lane] : 0.f;
                            s1 += u; s2 = fmaf(u, u, s2);
                        }
                        if (col < p.N) {
                            if (stat_mode == 2) {
                                float* mine = stat_smem + q * 2 * kCap1;       // this quadrant's private table
                                mine[col] += s1;
                                mine[kCap1 + col] += s2;
                            } else if (stat_mode == 1) {
                                atomicAdd(stat_smem + col, s1);
                                atomicAdd(stat_smem + kMaxStatCols + col, s2);
                            } else {
                                atomicAdd(p.col_stats + col, s1);
                                atomicAdd(p.col_stats + p.N + col, s2);
                            }
                        }
                    }
                    __syncwarp();                              // stage buffer is reused by the next chunk
                    if (WGRAD) { if (last_chunk) break; continue; }
                }

                if (row_ok) {
                    if (SC == 1) {
                        const float4* sc4 = reinterpret_cast<const float4*>(sc_tab + n_base);
                        const float4* sh4 = reinterpret_cast<const float4*>(sc_tab + kMaxStatCols + n_base);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 a = sc4[j], b = sh4[j];
                            v[4 * j] = fmaf(v[4 * j], a.x, b.x); v[4 * j + 1] = fmaf(v[4 * j + 1], a.y, b.y);
                            v[4 * j + 2] = fmaf(v[4 * j + 2], a.z, b.z); v[4 * j + 3] = fmaf(v[4 * j + 3], a.w, b.w);
                        }
                    } else if (SC == 0 && p.col_scale != nullptr) {
                        const int jmax = p.N - n_base;             // columns of this chunk that exist (predicated loads)
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < jmax) v[j] = fmaf(v[j], __ldg(p.col_scale + n_base + j), __ldg(p.col_shift + n_base + j));
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
                        // 16-byte stores per group of 8 columns; a narrow output (MobileNetV2's 16 / 24 / 32-channel projections)
                        // stores its 2 / 3 / 4 whole groups instead of 16 / 24 / 32 two-byte elements per lane
                        const bool vec = (full || (p.N & 7) == 0) && ((p.ldc & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                                         (add == nullptr || (((p.ld_add & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.addend) & 15) == 0)));
                        if (vec) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                if (n_base + g * 8 >= p.N) break;
                                float f[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) f[j] = v[g * 8 + j];
                                if (add != nullptr) {
                                    Vec8<__nv_bfloat16> a8; a8.load_rw(add + g * 8);      // may alias C
                                    float af[8]; a8.to_float(af);
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
                                    if (add != nullptr) f += __bfloat162float(__ldcg(add + j));
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
                            if (add != nullptr && n_base + j < p.N) v[j] += __ldcg(add + j);
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
                if (last_chunk) break;
            }
        }
        }   // direct-store fallback
        if (use_smem_stats) {
            // one flush per CTA: contended global atomics cost ~35 ns per cache line (serialised at L2)
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");    // the epilogue warps only
            const int e = threadIdx.x - 64;
            for (int c = e; c < p.N; c += 32 * kEpiWarps) {
                float a1, a2;
                if (stat_mode == 2) {
                    a1 = 0.f; a2 = 0.f;
#pragma unroll
                    for (int sl = 0; sl < 4; ++sl) { a1 += stat_smem[sl * 2 * kCap1 + c]; a2 += stat_smem[sl * 2 * kCap1 + kCap1 + c]; }
                } else { a1 = stat_smem[c]; a2 = stat_smem[kMaxStatCols + c]; }
                if (a1 != 0.f || a2 != 0.f) {
                    atomicAdd(p.col_stats + c, a1);
                    atomicAdd(p.col_stats + p.N + c, a2);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// =====================================================================================================================
// 2-CTA variant (forward / input-gradient GEMMs with N > 128): a cluster of two CTAs on one TPC computes a 256 x 256
// output tile with tcgen05.mma.cta_group::2 (UMMA M = 256).  Each CTA stages its own 128 rows of A and HALF of the B
// tile (128 of the 256 output channels), so per CTA a k-block costs 16 + 16 KB of TMA writes and 4 + 4 KB of tensor-core
// operand reads instead of 16 + 32 and 4 + 8: the single-CTA kernel is bound by shared-memory bandwidth (ncu: tensor
// pipe 35 % active, MMA queue always full, L2 / DRAM far from saturated — profiles/r1_gemm728_v2_*).
//   * both CTAs run a TMA producer (own A rows, own B half) that signals the LEADER's full barrier;
//   * only the leader issues MMAs; tcgen05.commit multicasts the "slot free" / "accumulator ready" arrivals to the
//     barriers of both CTAs;
//   * each CTA's epilogue warps drain their own 128 accumulator rows from their own TMEM and arrive (remotely for the
//     peer) on the leader's "accumulator free" barrier.
// =====================================================================================================================
constexpr int kStages2 = 4;            // 4 x 32 KB operand ring per CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int BLOCK_N, bool WGRAD, int SC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    constexpr int A_BYTES = kBlockM * kBlockK * 2;               // this CTA's 128 rows
    constexpr int B_BYTES = (BLOCK_N / 2) * kBlockK * 2;         // this CTA's half of the output channels
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* epi_bytes = smem + kStages2 * STAGE_BYTES;
    float* stat_smem = reinterpret_cast<float*>(epi_bytes + kEpiBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_bytes + kEpiBytes + kStatFloats2 * 4);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_bar = smem_u32(bars);                        // kStages2 (used in the leader only)
    const uint32_t empty_bar = smem_u32(bars + kStages2);            // kStages2 (one per CTA, multicast arrivals)
    const uint32_t tmem_full_bar = smem_u32(bars + 2 * kStages2);    // 2 (per CTA, multicast arrivals)
    const uint32_t tmem_empty_bar = smem_u32(bars + 2 * kStages2 + 2);   // 2 (leader only, 8 arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages2 + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    pdl_launch_dependents();

    if (!WGRAD && p.col_stats != nullptr)
        for (int i = threadIdx.x; i < kStatFloats2; i += kThreads) stat_smem[i] = 0.f;
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
        for (int s = 0; s < kStages2; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar + 8 * s, 1); mbar_init(tmem_empty_bar + 8 * s, 2 * kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // barriers of both CTAs initialised, TMEM allocated in both
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    pdl_wait();
    const float* sc_tab = SC == 1 ? fill_scale_table(p, stat_smem, kThreads) : nullptr;

    // forward: 256 x BLOCK_N pair tiles, n fastest.  filter gradient: (pixel-range split) x (tile), split-major; rows =
    // input channels (p.K), columns = output channels (p.N), reduction over the p.M pixels
    const int tiles = p.n_tiles * p.m_tiles;
    const int num_work = tiles * (WGRAD ? p.splits : 1);
    const int total_kb = ((WGRAD ? p.M : p.K) + kBlockK - 1) / kBlockK;
    auto kb_range = [&](int w, int& kb0, int& kb1) {
        if (WGRAD) { kb0 = (w / tiles) * p.kb_per_split; kb1 = min(kb0 + p.kb_per_split, total_kb); }
        else { kb0 = 0; kb1 = total_kb; }
    };

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer (both CTAs): own A rows + own B half, completion on the LEADER's full barrier =====
            const uint32_t leader_full = mapa_rank(full_bar, 0);
            uint32_t it = 0;
            for (int w = pair; w < num_work; w += num_pairs) {
                const int tile = w % tiles;
                const int row0 = (tile / p.n_tiles) * (2 * kBlockM) + (int)rank * kBlockM;
                const int col0 = (tile % p.n_tiles) * BLOCK_N + (int)rank * (BLOCK_N / 2);
                int kb0, kb1;
                kb_range(w, kb0, kb1);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const uint32_t s = it % kStages2;
                    const uint32_t ph = (it / kStages2) & 1u;
                    mbar_wait(empty_bar + 8 * s, ph ^ 1u);
                    const uint32_t a_dst = smem_base + s * STAGE_BYTES;
                    const uint32_t b_dst = a_dst + A_BYTES;
                    if (DLV3P_DBG(p, 1)) { if (leader) mbar_arrive(full_bar + 8 * s); continue; }
                    if (leader) mbar_expect_tx(full_bar + 8 * s, 2 * STAGE_BYTES);     // both CTAs' boxes
                    const uint32_t fb = leader_full + 8 * s;
                    if (WGRAD) {
                        // MN-major operands: boxes of 64 channels (inner, 128 B) x 64 pixels
#pragma unroll
                        for (int h = 0; h < kBlockM / 64; ++h)
                            tma_load_2d_2sm(a_dst + h * 8192, &tmA, fb, row0 + 64 * h, kb * kBlockK);
#pragma unroll
                        for (int h = 0; h < BLOCK_N / 2 / 64; ++h)
                            tma_load_2d_2sm(b_dst + h * 8192, &tmB, fb, col0 + 64 * h, kb * kBlockK);
                    } else {
                        tma_load_2d_2sm(a_dst, &tmA, fb, kb * kBlockK, row0);
                        tma_load_2d_2sm(b_dst, &tmB, fb, kb * kBlockK, col0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            // ===== MMA issuer (leader only) =====
            constexpr uint32_t idesc = (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ |
                                       ((WGRAD ? 1u : 0u) << 15) | ((WGRAD ? 1u : 0u) << 16) |
                                       ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)((2 * kBlockM) >> 4) << 24);
            uint32_t it = 0, t = 0;
            for (int w = pair; w < num_work; w += num_pairs, ++t) {
                int kb0, kb1;
                kb_range(w, kb0, kb1);
                const uint32_t as = t & 1u;
                mbar_wait(tmem_empty_bar + 8 * as, ((t >> 1) & 1u) ^ 1u);   // both CTAs' epilogues drained this stage
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BLOCK_N;
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const uint32_t s = it % kStages2;
                    const uint32_t ph = (it / kStages2) & 1u;
                    mbar_wait(full_bar + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t a_src = smem_base + s * STAGE_BYTES;
                    const uint32_t b_src = a_src + A_BYTES;
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {
                        // K-major SW128: a K=16 step is 32 B inside the swizzle atom; MN-major SW128: LBO = next
                        // 64-element MN chunk (one 8 KB box), SBO = next 8 K-rows, a K=16 step is 16 rows of 128 B
                        const uint64_t ad = WGRAD ? umma_desc(a_src + k * 2048, 8192, 1024) : umma_desc(a_src + k * 32, 16, 1024);
                        const uint64_t bd = WGRAD ? umma_desc(b_src + k * 2048, 8192, 1024) : umma_desc(b_src + k * 32, 16, 1024);
                        if (!DLV3P_DBG(p, 4)) tc_mma_bf16_2sm(d_tmem, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    tc_commit_2sm(empty_bar + 8 * s);          // slot reusable in BOTH CTAs once these MMAs retire
                }
                tc_commit_2sm(tmem_full_bar + 8 * as);         // accumulator stage complete in BOTH CTAs
            }
        }
    } else {
        // ===== epilogue (both CTAs): warps 2..9, TMEM lane quadrant = warp % 4, two warps per quadrant =====
        const int q = warp & 3;
        const int ew = warp - 2, half = ew >> 2;
        constexpr int kCap2 = stat_capacity(kStatFloats2, 4);
        const bool use_smem_stats = (!WGRAD) && (p.col_stats != nullptr) && (p.N <= kCap2);
        const int stat_mode = use_smem_stats ? 2 : 0;
        const uint32_t stg0 = smem_u32(epi_bytes) + (uint32_t)ew * 2u * kEpiBufBytes;
        const uint32_t leader_tmem_empty = mapa_rank(tmem_empty_bar, 0);
        uint32_t buf = 0, t = 0;
        for (int w = pair; w < num_work; w += num_pairs, ++t) {
            const int tile = w % tiles;
            const int row0 = (tile / p.n_tiles) * (2 * kBlockM) + (int)rank * kBlockM;
            const int col0 = (tile % p.n_tiles) * BLOCK_N;
            const uint32_t as = t & 1u;
            mbar_wait(tmem_full_bar + 8 * as, (t >> 1) & 1u);
            tc_fence_after();
            if (DLV3P_DBG(p, 2)) {
                if (lane == 0) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(leader_tmem_empty + 8 * as) : "memory");
                continue;
            }
            staged_tile_epilogue<BLOCK_N, WGRAD, true, SC>(p, &tmC, tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N,
                                                       row0 + q * 32, col0, lane, half, stg0, buf,
                                                       leader_tmem_empty + 8 * as, stat_smem, stat_mode, q, kCap2,
                                                       WGRAD ? p.K : p.M, -1, sc_tab);
        }
        if (lane == 0) tma_wait_group_read<0>();      // smem may be released; the writes drain before the grid completes
        if (use_smem_stats) {
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");    // the epilogue warps only
            const int e = threadIdx.x - 64;
            for (int c = e; c < p.N; c += 32 * kEpiWarps) {
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int sl = 0; sl < 4; ++sl) { a1 += stat_smem[sl * 2 * kCap2 + c]; a2 += stat_smem[sl * 2 * kCap2 + kCap2 + c]; }
                if (a1 != 0.f || a2 != 0.f) {
                    atomicAdd(p.col_stats + c, a1);
                    atomicAdd(p.col_stats + p.N + c, a2);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // neither CTA may exit (or free TMEM) while its peer still uses it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------------------------------
template <int BLOCK_N, bool WGRAD, int SC>
static int launch_gemm_inst(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                            cudaStream_t st) {
    constexpr int smem = GemmCfg<BLOCK_N>::kSmem;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N, WGRAD, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", smem, cudaGetErrorString(e));
        configured = true;
    }
    const int work = p.n_tiles * p.m_tiles * (WGRAD ? p.splits : 1);
    const int grid = work < kNumSMs ? work : kNumSMs;         // persistent: one CTA per SM
    launch_pdl(gemm_tc_kernel<BLOCK_N, WGRAD, SC>, dim3(grid), dim3(kThreads), smem, st, tmA, tmB, tmC, p);
    return check_launch(WGRAD ? "gemm_wgrad_bf16" : "gemm_bf16");
}

// the shared-memory scale / shift table lives in the statistics region: usable when the launch collects no statistics
static inline bool scale_table_ok(const GemmParams& p) { return p.col_stats == nullptr && p.N <= kMaxStatCols; }

template <int BLOCK_N, bool WGRAD>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                       cudaStream_t st) {
    if constexpr (!WGRAD) {
        if (p.col_scale != nullptr)
            return scale_table_ok(p) ? launch_gemm_inst<BLOCK_N, false, 1>(tmA, tmB, tmC, p, st)
                                     : launch_gemm_inst<BLOCK_N, false, 0>(tmA, tmB, tmC, p, st);
    }
    return launch_gemm_inst<BLOCK_N, WGRAD, 2>(tmA, tmB, tmC, p, st);
}

template <int BLOCK_N, bool WGRAD, int SC>
static int launch_gemm2_inst(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, GemmParams p,
                             cudaStream_t st) {
    constexpr int smem = kStages2 * (kBlockM * kBlockK * 2 + (BLOCK_N / 2) * kBlockK * 2) + kEpiBytes +
                         kStatFloats2 * 4 + 1024 + 256;
    static_assert(smem <= 232448, "shared memory budget");
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc2_kernel<BLOCK_N, WGRAD, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "cudaFuncSetAttribute(2-CTA smem=%d): %s", smem, cudaGetErrorString(e));
        configured = true;
    }
    const int work = p.n_tiles * p.m_tiles * (WGRAD ? p.splits : 1);
    const int pairs = work < kNumSMs / 2 ? work : kNumSMs / 2;
    launch_pdl(gemm_tc2_kernel<BLOCK_N, WGRAD, SC>, dim3(2 * pairs), dim3(kThreads), smem, st, tmA, tmB, tmC, p);
    return check_launch("gemm_bf16 (2-CTA)");
}

template <int BLOCK_N, bool WGRAD>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const GemmParams& p,
                        cudaStream_t st) {
    if constexpr (!WGRAD) {
        if (p.col_scale != nullptr)
            return scale_table_ok(p) ? launch_gemm2_inst<BLOCK_N, false, 1>(tmA, tmB, tmC, p, st)
                                     : launch_gemm2_inst<BLOCK_N, false, 0>(tmA, tmB, tmC, p, st);
    }
    return launch_gemm2_inst<BLOCK_N, WGRAD, 2>(tmA, tmB, tmC, p, st);
}

#ifdef DLV3P_DIAG
// diagnostics build: DLV3P_GEMM_2CTA=0 disables the 2-CTA path (A/B measurements); DLV3P_GEMM_DBG (re-read on every call
// once DLV3P_GEMM_DBG_ENABLE is set at first use) switches off operand loads (1), epilogue work (2), MMAs (4), C stores
// (8), TMEM reads (16) of the 2-CTA kernel for the bottleneck decomposition; results are garbage by design
static int g_gemm_2cta = -1;
static int gemm_dbg_mode() {
    static const bool enabled = getenv("DLV3P_GEMM_DBG_ENABLE") != nullptr;
    if (!enabled) return 0;
    const char* e = getenv("DLV3P_GEMM_DBG");
    return e ? atoi(e) : 0;
}
static bool gemm_2cta_enabled() {
    return g_gemm_2cta != 0;
}
#else
static int gemm_dbg_mode() { return 0; }
static bool gemm_2cta_enabled() { return true; }
#endif

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
    p.addend = addend; p.ld_add = ld_addend; p.col_stats = col_stats; p.kb_per_split = 0; p.splits = 1;
    p.dbg = gemm_dbg_mode(); p.conv_mode = 0;
    p.n_tiles = cdiv(N, bn); p.m_tiles = cdiv(M, kBlockM);
    p.tma_store = (c_dtype == DLV3P_BF16 && addend == nullptr && (ldc % 8) == 0 && aligned16(C) && bn >= 64) ? 1 : 0;
    CUtensorMap tmC = tmA;
    if (p.tma_store) {
        rc = make_tmap(&tmC, C, N, M, ldc, 64, 32);
        if (rc) return rc;
    }
    // (short reductions are epilogue-bound and gain nothing from pairing: measured 64 vs 51 us at M=258064 N=K=256)
    if (gemm_2cta_enabled() && bn == 256 && p.tma_store && M >= 2 * kBlockM && K >= 512) {
        // 2-CTA pairs: every CTA loads half of the B tile (128 rows of the [N,K] operand)
        rc = make_tmap(&tmB, B, K, N, ldb, kBlockK, bn / 2);
        if (rc) return rc;
        p.m_tiles = cdiv(M, 2 * kBlockM);
        return launch_gemm2<256, false>(tmA, tmB, tmC, p, st);
    }
    switch (bn) {
        case 32: return launch_gemm<32, false>(tmA, tmB, tmC, p, st);
        case 64: return launch_gemm<64, false>(tmA, tmB, tmC, p, st);
        case 128: return launch_gemm<128, false>(tmA, tmB, tmC, p, st);
        default: return launch_gemm<256, false>(tmA, tmB, tmC, p, st);
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
    GemmParams p;
    p.M = M; p.N = N; p.K = K; p.C = dW; p.ldc = ldw; p.c_dtype = DLV3P_F32;
    p.col_scale = nullptr; p.col_shift = nullptr; p.act = 0; p.addend = nullptr; p.ld_add = 0; p.col_stats = nullptr;
    p.dbg = gemm_dbg_mode(); p.conv_mode = 0;
    p.n_tiles = cdiv(N, bn); p.m_tiles = cdiv(K, kBlockM);
    const int tiles = p.n_tiles * p.m_tiles;
    const int total_kb = cdiv(M, kBlockK);
    // one wave of work items: every extra split costs a full tile of fp32 REDs
    int splits = kNumSMs / tiles;
    if (splits < 1) splits = 1;
    if (splits > total_kb) splits = total_kb;
    p.kb_per_split = cdiv(total_kb, splits);
    p.splits = cdiv(total_kb, p.kb_per_split);
    p.tma_store = ((ldw % 4) == 0 && aligned16(dW)) ? 1 : 0;
    CUtensorMap tmC = tmA;
    if (p.tma_store) {
        rc = make_tmap(&tmC, dW, N, K, ldw, 32, 32, /*f32=*/true);
        if (rc) return rc;
    }
    if (gemm_2cta_enabled() && bn == 256 && p.tma_store && K > kBlockM && M >= 4096) {
        // 2-CTA pairs: 256 input channels x 256 output channels per pair tile, one wave of (split, tile) items
        p.m_tiles = cdiv(K, 2 * kBlockM);
        const int tiles2 = p.n_tiles * p.m_tiles;
        int sp = (kNumSMs / 2) / tiles2; if (sp < 1) sp = 1; if (sp > total_kb) sp = total_kb;
        p.kb_per_split = cdiv(total_kb, sp);
        p.splits = cdiv(total_kb, p.kb_per_split);
        return launch_gemm2<256, true>(tmA, tmB, tmC, p, st);
    }
    switch (bn) {
        case 64: return launch_gemm<64, true>(tmA, tmB, tmC, p, st);
        case 128: return launch_gemm<128, true>(tmA, tmB, tmC, p, st);
        default: return launch_gemm<256, true>(tmA, tmB, tmC, p, st);
    }
}


// =====================================================================================================================
// Implicit-GEMM 3x3 VALID stride-1 convolution (Xception block1_conv2, keras.applications.xception; the reference
// reaches it through ss.py:512-515).  The im2col matrix [N*Ho*Wo, 9*Cin] is never written: under one filter row the
// three taps of an output pixel are 3*Cin CONTIGUOUS NHWC elements, so a rank-3 tensor map whose rows overlap (row
// length 3*Cin elements, row pitch Cin elements) hands the GEMM its A tiles directly; the run is cut into 64-element
// k-blocks and TMA zero-fills what lies beyond the run / the image row, so K = 3 * 64*ceil(3*Cin/64) with matching zero
// columns in the prepared filter matrix.  M tiles are 128 output pixels of ONE image row (rank-3 C map clips the
// ragged row end), so BatchNormalization statistics see exact zeros for the clipped pixels.
//   wt (forward B operand)  bf16 [Cout, 9*Cin] K-major (row pitch ldw), wt[o, (i*3+j)*Cin + c] = W[i,j,c,o] — the same
//                           matrix the im2col GEMM uses
//   wd (dgrad B operand)    bf16 [Cin, 9*Cout], wd[c, (i*3+j)*Cout + o] = W[i,j,c,o]
// =====================================================================================================================
// ---- forward 3x3 VALID convolution with 32 input channels: halo-staged implicit GEMM ---------------------------------
// Same idea as the input-gradient kernel below, on 64-byte pixels (Cin = 32): one stage holds the three input rows
// ho, ho+1, ho+2 as 130-pixel boxes under the 64B swizzle, tap (i, j) is the view that starts j pixels (64 B each) into
// row buffer i (K = 32 per tap, two UMMA K-steps), the nine [Cout x 32] filter slices stay resident.  Per 128-pixel tile
// 25 KB arrive instead of the 144 KB of the run-based path (measured 163 us, feed-bound).  Epilogue = the GEMM's staged
// TMA-store epilogue (bf16 tile, BatchNormalization statistics, rank-3 C map clips the ragged row end).
namespace dlv3p {
constexpr int kDgBoxPix = 130;                          // 128 + 2 halo pixels
constexpr int kFwRowBytes = 9 * 1024;                   // 130 x 64 B rounded up to 1 KB
constexpr int kFwStageBytes = 3 * kFwRowBytes;
constexpr int kFwStages = 4;
constexpr int kFwBTapBytes = 64 * 64;                   // [64 output channels x 32 input channels] bf16, K-major, SW64
constexpr int kFwSmem = kFwStages * kFwStageBytes + 9 * kFwBTapBytes + kEpiBytes + 2 * kMaxStatCols * 4 + 1024 + 256;
static_assert(kFwSmem <= 232448, "shared memory budget");

// UMMA smem descriptor, K-major, 64B swizzle (layout type 4): 8-row groups are 512 B apart
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) |
           (4ull << 61);
}

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_valid_fwd32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t b_base = smem_base + kFwStages * kFwStageBytes;
    uint8_t* epi_bytes = smem + kFwStages * kFwStageBytes + 9 * kFwBTapBytes;
    float* stat_smem = reinterpret_cast<float*>(epi_bytes + kEpiBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_bytes + kEpiBytes + 2 * kMaxStatCols * 4);
    const uint32_t full_bar = smem_u32(bars);
    const uint32_t empty_bar = smem_u32(bars + kFwStages);
    const uint32_t tmem_full_bar = smem_u32(bars + 2 * kFwStages);
    const uint32_t tmem_empty_bar = smem_u32(bars + 2 * kFwStages + 2);
    const uint32_t b_bar = smem_u32(bars + 2 * kFwStages + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kFwStages + 5);
    constexpr uint32_t TMEM_COLS = 128;                    // two accumulator stages of 64 columns
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    if (p.col_stats != nullptr)
        for (int i = threadIdx.x; i < 2 * kMaxStatCols; i += kThreads) stat_smem[i] = 0.f;
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmC)) : "memory");
        for (int s = 0; s < kFwStages; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar + 8 * s, 1); mbar_init(tmem_empty_bar + 8 * s, kEpiWarps); }
        mbar_init(b_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    pdl_wait();
    const int tiles = p.m_tiles;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(b_bar, 9 * kFwBTapBytes);
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(b_base + tap * kFwBTapBytes, &tmB, b_bar, tap * 32, 0);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const uint32_t s = it % kFwStages, ph = (it / kFwStages) & 1u;
                const int r = t / p.cv_tpr, w0 = (t % p.cv_tpr) * kBlockM;
                const int n = r / p.cv_rows_out, ho = r % p.cv_rows_out;
                mbar_wait(empty_bar + 8 * s, ph ^ 1u);
                mbar_expect_tx(full_bar + 8 * s, 3 * kDgBoxPix * 64);
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    tma_load_4d(smem_base + s * kFwStageBytes + i * kFwRowBytes, &tmA, full_bar + 8 * s, 0, w0, ho + i, n);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) |
                                       ((uint32_t)(kBlockM >> 4) << 24);
            mbar_wait(b_bar, 0);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const uint32_t s = it % kFwStages, ph = (it / kFwStages) & 1u;
                const uint32_t as = it & 1u;
                mbar_wait(tmem_empty_bar + 8 * as, ((it >> 1) & 1u) ^ 1u);
                mbar_wait(full_bar + 8 * s, ph);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 64;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const uint32_t a_src = smem_base + s * kFwStageBytes + (tap / 3) * kFwRowBytes + (uint32_t)(tap % 3) * 64u;
#pragma unroll
                    for (int k = 0; k < 2; ++k)
                        tc_mma_bf16(d_tmem, umma_desc_sw64(a_src + k * 32), umma_desc_sw64(b_base + tap * kFwBTapBytes + k * 32),
                                    idesc, (tap > 0 || k > 0) ? 1u : 0u);
                }
                tc_commit(empty_bar + 8 * s);
                tc_commit(tmem_full_bar + 8 * as);
            }
        }
    } else {
        const int q = warp & 3;
        const int ew = warp - 2, half = ew >> 2;
        constexpr int kCapF = stat_capacity(kStatFloats1, kEpiWarps);       // one private table per epilogue WARP here
        const bool use_smem_stats = (p.col_stats != nullptr) && (p.N <= kCapF);
        const int stat_mode = use_smem_stats ? 2 : 0;
        const uint32_t stg0 = smem_u32(epi_bytes) + (uint32_t)ew * 2u * kEpiBufBytes;
        uint32_t buf = 0, it = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
            const uint32_t as = it & 1u;
            const int r = t / p.cv_tpr, w0 = (t % p.cv_tpr) * kBlockM;
            mbar_wait(tmem_full_bar + 8 * as, (it >> 1) & 1u);
            tc_fence_after();
            // one 64-column chunk per tile: the two warps of a TMEM lane quadrant take alternate TILES, so the epilogues of
            // both accumulator stages run concurrently (the tile loop is epilogue-paced: 18 short MMAs per tile)
            staged_tile_epilogue<64, false, false>(p, &tmC, tmem_base + ((uint32_t)(q * 32) << 16) + as * 64, w0 + q * 32, 0,
                                                   lane, half ^ (int)(it & 1u), stg0, buf, tmem_empty_bar + 8 * as, stat_smem,
                                                   stat_mode, ew, kCapF, p.cv_rlimit, r);
        }
        if (lane == 0) tma_wait_group_read<0>();      // smem may be released; the writes drain before the grid completes
        if (use_smem_stats) {
            asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
            const int e = threadIdx.x - 64;
            for (int c = e; c < p.N; c += 32 * kEpiWarps) {
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int sl = 0; sl < kEpiWarps; ++sl) { a1 += stat_smem[sl * 2 * kCapF + c]; a2 += stat_smem[sl * 2 * kCapF + kCapF + c]; }
                if (a1 != 0.f || a2 != 0.f) {
                    atomicAdd(p.col_stats + c, a1);
                    atomicAdd(p.col_stats + p.N + c, a2);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}
}  // namespace dlv3p

static int conv_check(const void* a, const void* b, const void* c, int N, int H, int W, int Cin, int Cout) {
    DLV3P_REQUIRE(a && b && c && N > 0 && H >= 3 && W >= 3 && Cin > 0 && Cout > 0, DLV3P_ERR_SHAPE,
                  "conv3x3_valid: bad arguments N=%d H=%d W=%d Cin=%d Cout=%d", N, H, W, Cin, Cout);
    DLV3P_REQUIRE((Cin % 8) == 0 && (Cout % 8) == 0 && aligned16(a) && aligned16(b) && aligned16(c), DLV3P_ERR_ALIGN,
                  "conv3x3_valid: Cin/Cout must be multiples of 8 and the pointers 16-byte aligned");
    return 0;
}

extern "C" int dlv3p_conv3x3_valid_fwd_bf16(const void* x, const void* wt, int64_t ldw, void* y, int N, int H, int W,
                                            int Cin, int Cout, const float* col_scale, const float* col_shift, int act,
                                            float* col_stats, void* stream) {
    int rc = conv_check(x, wt, y, N, H, W, Cin, Cout);
    if (rc) return rc;
    DLV3P_REQUIRE(Cout <= 256 && Cout >= 64, DLV3P_ERR_UNSUPPORTED, "conv3x3_valid_fwd: 64 <= Cout <= 256 (got %d)", Cout);
    DLV3P_REQUIRE(ldw >= 9LL * Cin && (ldw % 8) == 0, DLV3P_ERR_ALIGN, "conv3x3_valid_fwd: ldw >= 9*Cin and a multiple of 8");
    DLV3P_REQUIRE((col_scale == nullptr) == (col_shift == nullptr), DLV3P_ERR_SHAPE, "conv3x3_valid_fwd: scale/shift mismatch");
    const int Ho = H - 2, Wo = W - 2;
    const int kbr = cdiv(3 * Cin, kBlockK), KR = kbr * kBlockK;
    const int bn = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
    const bool fast = (Cin == 32 && Cout == 64);             // 64-byte pixels: halo-staged kernel, taps as shifted views
    CUtensorMap tmA, tmB, tmC;
    if (fast) {
        const long long dims[4] = {Cin, W, H, N};
        const long long str[3] = {2LL * Cin, 2LL * W * Cin, 2LL * H * W * Cin};
        const int box[4] = {32, kDgBoxPix, 1, 1};
        rc = make_tmap_nd(&tmA, x, 4, dims, str, box, false, /*swizzle64=*/true);
        if (rc) return rc;
        const long long bdims[2] = {9LL * Cin, Cout};
        const long long bstr[1] = {2LL * ldw};
        const int bbox[2] = {32, 64};
        rc = make_tmap_nd(&tmB, wt, 2, bdims, bstr, bbox, false, /*swizzle64=*/true);
        if (rc) return rc;
    } else {
        const long long dims[3] = {3LL * Cin, Wo, (long long)N * H};
        const long long str[2] = {2LL * Cin, 2LL * W * Cin};
        const int box[3] = {kBlockK, kBlockM, 1};
        rc = make_tmap_nd(&tmA, x, 3, dims, str, box);
        if (rc) return rc;
        rc = make_tmap(&tmB, wt, 9LL * Cin, Cout, ldw, kBlockK, bn);
        if (rc) return rc;
    }
    {
        const long long dims[3] = {Cout, Wo, (long long)N * Ho};
        const long long str[2] = {2LL * Cout, 2LL * Wo * Cout};
        const int box[3] = {64, 32, 1};
        rc = make_tmap_nd(&tmC, y, 3, dims, str, box);
        if (rc) return rc;
    }
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.cv_tpr = cdiv(Wo, kBlockM);
    p.m_tiles = N * Ho * p.cv_tpr; p.n_tiles = cdiv(Cout, bn);
    p.M = p.m_tiles * kBlockM; p.N = Cout; p.K = 3 * KR;
    p.C = y; p.ldc = Cout; p.c_dtype = DLV3P_BF16; p.col_scale = col_scale; p.col_shift = col_shift; p.act = act;
    p.col_stats = col_stats; p.splits = 1; p.tma_store = 1; p.dbg = 0;
    p.conv_mode = 1; p.cv_rows_in = H; p.cv_rows_out = Ho; p.cv_width = Wo; p.cv_rlimit = Wo; p.cv_kbr = kbr;
    p.cv_run = 3 * Cin;
    cudaStream_t st = (cudaStream_t)stream;
    if (fast) {
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(conv3x3_valid_fwd32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwSmem);
            DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "cudaFuncSetAttribute(conv fwd smem=%d): %s", kFwSmem, cudaGetErrorString(e));
            configured = true;
        }
        const int grid = p.m_tiles < kNumSMs ? p.m_tiles : kNumSMs;
        launch_pdl(conv3x3_valid_fwd32_kernel, dim3(grid), dim3(kThreads), kFwSmem, st, tmA, tmB, tmC, p);
        return check_launch("conv3x3_valid_fwd (halo-staged)");
    }
    switch (bn) {
        case 64: return launch_gemm<64, false>(tmA, tmB, tmC, p, st);
        case 128: return launch_gemm<128, false>(tmA, tmB, tmC, p, st);
        default: return launch_gemm<256, false>(tmA, tmB, tmC, p, st);
    }
}

// ---- input gradient of the 3x3 VALID convolution: halo-staged implicit GEMM --------------------------------------
// dx[n,h,w,c] = sum_{i,j,o} dy[n,h-i,w-j,o] * W[i,j,c,o].  M tile = 128 consecutive pixels of one dx row, K = 9 taps x 64
// output channels, N = Cin (<= 32).  The generic implicit path above would fetch nine shifted 16 KB windows of dy per
// tile (measured 226 us at [16,256,256,32]: 1.8 GB of L2 -> SM boxes, feed-bound).  Here ONE stage holds the three dy
// rows h, h-1, h-2 as 130-pixel boxes (pixels w0-2 .. w0+127, zero-filled outside dy = full correlation), and the nine
// taps are nine VIEWS of that stage: the A descriptor of tap (i, j) simply starts (2-j) pixel rows (128 B each) into row
// buffer i.  The 128B swizzle of both TMA and UMMA is a function of the absolute shared-memory address (bits 4..6 ^= bits
// 7..9), so a start address that is off the 1024-byte swizzle period needs nothing else: measured on B200, descriptor
// base_offset (bits 49..51) = 0 reproduces the fp64 reference bit-for-bit-in-bf16, (start >> 7) & 7 does not.
// The 9 x [Cin x 64] filter slices stay resident in shared memory for the whole kernel.
namespace dlv3p {
constexpr int kDgRowBytes = 17 * 1024;                  // 130 x 128 B rounded up to the swizzle period
constexpr int kDgStageBytes = 3 * kDgRowBytes;
constexpr int kDgStages = 3;
constexpr int kDgBTapBytes = 32 * 128;                  // [32 input channels x 64 output channels] bf16, K-major
constexpr int kDgThreads = 320;                         // producer, issuer, 8 epilogue warps (two per TMEM lane quadrant)
constexpr int kDgSmem = kDgStages * kDgStageBytes + 9 * kDgBTapBytes + 1024 + 256;
static_assert(kDgSmem <= 232448, "shared memory budget");

struct ConvDgradParams {
    __nv_bfloat16* dx;
    int N, H, W, Cin, tpr, tiles;
};

__global__ void __launch_bounds__(kDgThreads, 1)
conv3x3_valid_dgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const ConvDgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t b_base = smem_base + kDgStages * kDgStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDgStages * kDgStageBytes + 9 * kDgBTapBytes);
    const uint32_t full_bar = smem_u32(bars);                      // kDgStages
    const uint32_t empty_bar = smem_u32(bars + kDgStages);         // kDgStages
    const uint32_t tmem_full_bar = smem_u32(bars + 2 * kDgStages);     // 2
    const uint32_t tmem_empty_bar = smem_u32(bars + 2 * kDgStages + 2);    // 2
    const uint32_t b_bar = smem_u32(bars + 2 * kDgStages + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kDgStages + 5);
    constexpr uint32_t TMEM_COLS = 64;                     // two accumulator stages of 32 columns
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < kDgStages; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar + 8 * s, 1); mbar_init(tmem_empty_bar + 8 * s, 4); }
        mbar_init(b_bar, 1);                                  // (accumulator stage `as` is drained by the 4 warps of one half)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            // filter slices: resident for the whole kernel
            mbar_expect_tx(b_bar, 9 * kDgBTapBytes);
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(b_base + tap * kDgBTapBytes, &tmB, b_bar, tap * 64, 0);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
                const uint32_t s = it % kDgStages, ph = (it / kDgStages) & 1u;
                const int r = t / p.tpr, w0 = (t % p.tpr) * kBlockM;
                const int n = r / p.H, h = r % p.H;
                mbar_wait(empty_bar + 8 * s, ph ^ 1u);
                mbar_expect_tx(full_bar + 8 * s, 3 * kDgBoxPix * 128);
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    tma_load_4d(smem_base + s * kDgStageBytes + i * kDgRowBytes, &tmA, full_bar + 8 * s, 0, w0 - 2, h - i, n);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) |
                                       ((uint32_t)(kBlockM >> 4) << 24);
            mbar_wait(b_bar, 0);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
                const uint32_t s = it % kDgStages, ph = (it / kDgStages) & 1u;
                const uint32_t as = it & 1u;
                mbar_wait(tmem_empty_bar + 8 * as, ((it >> 1) & 1u) ^ 1u);
                mbar_wait(full_bar + 8 * s, ph);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 32;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int i = tap / 3, j = tap % 3;
                    const uint32_t a_src = smem_base + s * kDgStageBytes + i * kDgRowBytes + (uint32_t)(2 - j) * 128u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t ad = umma_desc(a_src + k * 32, 16, 1024);
                        const uint64_t bd = umma_desc(b_base + tap * kDgBTapBytes + k * 32, 16, 1024);
                        tc_mma_bf16(d_tmem, ad, bd, idesc, (tap > 0 || k > 0) ? 1u : 0u);
                    }
                }
                tc_commit(empty_bar + 8 * s);
                tc_commit(tmem_full_bar + 8 * as);
            }
        }
    } else {
        const int q = warp & 3;                              // TMEM lane quadrant of this warp
        const uint32_t half = (uint32_t)(warp - 2) >> 2;     // warps 2..5 drain the even tiles, 6..9 the odd ones
        uint32_t it = 0;
        for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
            const uint32_t as = it & 1u;
            if (as != half) continue;
            const int r = t / p.tpr, w = (t % p.tpr) * kBlockM + q * 32 + lane;
            mbar_wait(tmem_full_bar + 8 * as, (it >> 1) & 1u);
            tc_fence_after();
            uint32_t raw[32];
            tc_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + as * 32, raw);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar + 8 * as);
            if (w < p.W) {
                // one pixel per lane: Cin (<= 32) contiguous bf16 = up to four 16-byte stores
                __nv_bfloat16* dst = p.dx + ((long long)r * p.W + w) * p.Cin;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (g * 8 >= p.Cin) break;
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        __nv_bfloat162 pr = __floats2bfloat162_rn(__uint_as_float(raw[g * 8 + 2 * e]),
                                                                  __uint_as_float(raw[g * 8 + 2 * e + 1]));
                        o[e] = *reinterpret_cast<uint32_t*>(&pr);
                    }
                    *reinterpret_cast<uint4*>(dst + g * 8) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}
}  // namespace dlv3p

extern "C" int dlv3p_conv3x3_valid_dgrad_bf16(const void* dy, const void* wd, void* dx, int N, int H, int W, int Cin,
                                              int Cout, void* stream) {
    int rc = conv_check(dy, wd, dx, N, H, W, Cin, Cout);
    if (rc) return rc;
    DLV3P_REQUIRE(Cout == 64 && Cin <= 32, DLV3P_ERR_UNSUPPORTED,
                  "conv3x3_valid_dgrad: Cout == 64 and Cin <= 32 (got Cin=%d Cout=%d)", Cin, Cout);
    const int Ho = H - 2, Wo = W - 2;
    CUtensorMap tmA, tmB;
    {
        const long long dims[4] = {Cout, Wo, Ho, N};
        const long long str[3] = {2LL * Cout, 2LL * Wo * Cout, 2LL * Ho * Wo * Cout};
        const int box[4] = {kBlockK, kDgBoxPix, 1, 1};
        rc = make_tmap_nd(&tmA, dy, 4, dims, str, box);
        if (rc) return rc;
    }
    rc = make_tmap(&tmB, wd, 9LL * Cout, Cin, 9LL * Cout, kBlockK, 32);
    if (rc) return rc;
    ConvDgradParams p;
    p.dx = (__nv_bfloat16*)dx; p.N = N; p.H = H; p.W = W; p.Cin = Cin;
    p.tpr = cdiv(W, kBlockM); p.tiles = N * H * p.tpr;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_valid_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDgSmem);
        DLV3P_REQUIRE(e == cudaSuccess, DLV3P_ERR_CUDA, "cudaFuncSetAttribute(conv dgrad smem=%d): %s", kDgSmem, cudaGetErrorString(e));
        configured = true;
    }
    const int grid = p.tiles < kNumSMs ? p.tiles : kNumSMs;
    launch_pdl(conv3x3_valid_dgrad_kernel, dim3(grid), dim3(kDgThreads), kDgSmem, (cudaStream_t)stream, tmA, tmB, p);
    return check_launch("conv3x3_valid_dgrad");
}

extern "C" int dlv3p_conv3x3_valid_wgrad_bf16(const void* x, const void* dy, float* dw, int N, int H, int W, int Cin,
                                              int Cout, void* stream) {
    int rc = conv_check(x, dy, dw, N, H, W, Cin, Cout);
    if (rc) return rc;
    DLV3P_REQUIRE(3 * Cin > 64 && 3 * Cin <= 128 && Cout <= 256, DLV3P_ERR_UNSUPPORTED,
                  "conv3x3_valid_wgrad: 64 < 3*Cin <= 128 and Cout <= 256 (got Cin=%d Cout=%d)", Cin, Cout);
    const int Ho = H - 2, Wo = W - 2;
    const int bn = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
    CUtensorMap tmA, tmB, tmC;
    {
        const long long dims[3] = {3LL * Cin, Wo, (long long)N * H};
        const long long str[2] = {2LL * Cin, 2LL * W * Cin};
        const int box[3] = {64, kBlockK, 1};
        rc = make_tmap_nd(&tmA, x, 3, dims, str, box);
        if (rc) return rc;
    }
    rc = make_tmap(&tmB, dy, Cout, (long long)N * Ho * Wo, Cout, 64, kBlockK);
    if (rc) return rc;
    {
        // dw = HWIO [3][3*Cin][Cout] fp32: rank-3 map clips the zero-padded tail of every 128-element run
        const long long dims[3] = {Cout, 3LL * Cin, 3};
        const long long str[2] = {4LL * Cout, 4LL * 3 * Cin * Cout};
        const int box[3] = {32, 32, 1};
        rc = make_tmap_nd(&tmC, dw, 3, dims, str, box, /*f32=*/true);
        if (rc) return rc;
    }
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.cv_tpr = cdiv(Wo, kBlockK);                         // 64-pixel reduction blocks per output row
    const int total_kb = N * Ho * p.cv_tpr;
    p.m_tiles = 3; p.n_tiles = cdiv(Cout, bn);
    p.M = total_kb * kBlockK; p.N = Cout; p.K = 3 * kBlockM;
    p.C = dw; p.ldc = Cout; p.c_dtype = DLV3P_F32; p.tma_store = 1;
    const int tiles = p.m_tiles * p.n_tiles;
    int splits = kNumSMs / tiles; if (splits < 1) splits = 1; if (splits > total_kb) splits = total_kb;
    p.kb_per_split = cdiv(total_kb, splits);
    p.splits = cdiv(total_kb, p.kb_per_split);
    p.conv_mode = 3; p.cv_rows_in = H; p.cv_rows_out = Ho; p.cv_width = Wo; p.cv_rlimit = 3 * Cin; p.cv_kbr = 2;
    cudaStream_t st = (cudaStream_t)stream;
    switch (bn) {
        case 64: return launch_gemm<64, true>(tmA, tmB, tmC, p, st);
        case 128: return launch_gemm<128, true>(tmA, tmB, tmC, p, st);
        default: return launch_gemm<256, true>(tmA, tmB, tmC, p, st);
    }
}


// =====================================================================================================================
// 3x3 SAME stride-1 convolution as implicit GEMMs (conv modes 4 / 5 / 6 of gemm_tc_kernel): the logits convolution of
// the decoder (ss.py:893-897; after boundary refinement its input is the 304-channel concat at 256 x 256, ss.py:915-954)
// without the [pixels, 9*Cin] column matrix (5.7 GB per step at BASELINE cfg-4) that im2col + GEMM + col2im move.
// =====================================================================================================================
static int same_check(const char* who, const void* a, const void* b, const void* c, int N, int H, int W, int Cin, int Cout) {
    DLV3P_REQUIRE(a && b && c && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, DLV3P_ERR_SHAPE,
                  "%s: bad arguments N=%d H=%d W=%d Cin=%d Cout=%d", who, N, H, W, Cin, Cout);
    DLV3P_REQUIRE((Cin % 8) == 0 && aligned16(a) && aligned16(b), DLV3P_ERR_ALIGN,
                  "%s: Cin must be a multiple of 8 and the operand pointers 16-byte aligned (Cin=%d)", who, Cin);
    DLV3P_REQUIRE((long long)N * H * cdiv(W, kBlockM) < (1LL << 23), DLV3P_ERR_UNSUPPORTED, "%s: too many tiles", who);
    return 0;
}

extern "C" int dlv3p_conv3x3_same_fwd_bf16(const void* x, const void* wt, int64_t ldw, void* y, int c_dtype, int N, int H,
                                           int W, int Cin, int Cout, const float* col_scale, const float* col_shift,
                                           int act, float* col_stats, void* stream) {
    int rc = same_check("conv3x3_same_fwd", x, wt, y, N, H, W, Cin, Cout);
    if (rc) return rc;
    DLV3P_REQUIRE(Cout <= 256, DLV3P_ERR_UNSUPPORTED, "conv3x3_same_fwd: Cout <= 256 (got %d)", Cout);
    DLV3P_REQUIRE(ldw >= 9LL * Cin && (ldw % 8) == 0, DLV3P_ERR_ALIGN, "conv3x3_same_fwd: ldw >= 9*Cin and a multiple of 8");
    DLV3P_REQUIRE((col_scale == nullptr) == (col_shift == nullptr), DLV3P_ERR_SHAPE, "conv3x3_same_fwd: scale/shift mismatch");
    DLV3P_REQUIRE(c_dtype == DLV3P_BF16 || c_dtype == DLV3P_F32, DLV3P_ERR_DTYPE, "conv3x3_same_fwd: bad c_dtype %d", c_dtype);
    const int bn = Cout <= 32 ? 32 : (Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256));
    const int kbr = cdiv(Cin, kBlockK);
    CUtensorMap tmA, tmB, tmC;
    {
        const long long dims[4] = {Cin, W, H, N};
        const long long str[3] = {2LL * Cin, 2LL * W * Cin, 2LL * H * W * Cin};
        const int box[4] = {kBlockK, kBlockM, 1, 1};
        rc = make_tmap_nd(&tmA, x, 4, dims, str, box);
        if (rc) return rc;
    }
    rc = make_tmap(&tmB, wt, 9LL * Cin, Cout, ldw, kBlockK, bn);
    if (rc) return rc;
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.tma_store = (c_dtype == DLV3P_BF16 && (Cout % 8) == 0 && aligned16(y) && bn >= 64) ? 1 : 0;
    tmC = tmA;
    if (p.tma_store) {
        const long long dims[3] = {Cout, W, (long long)N * H};
        const long long str[2] = {2LL * Cout, 2LL * W * Cout};
        const int box[3] = {64, 32, 1};
        rc = make_tmap_nd(&tmC, y, 3, dims, str, box);
        if (rc) return rc;
    }
    p.cv_tpr = cdiv(W, kBlockM);
    p.m_tiles = N * H * p.cv_tpr; p.n_tiles = cdiv(Cout, bn);
    p.M = p.m_tiles * kBlockM; p.N = Cout; p.K = 9 * kbr * kBlockK;
    p.C = y; p.ldc = Cout; p.c_dtype = c_dtype; p.col_scale = col_scale; p.col_shift = col_shift; p.act = act;
    p.col_stats = col_stats; p.splits = 1;
    p.conv_mode = 4; p.cv_rows_in = H; p.cv_rows_out = H; p.cv_width = W; p.cv_rlimit = W; p.cv_kbr = kbr; p.cv_run = Cin;
    cudaStream_t st = (cudaStream_t)stream;
    switch (bn) {
        case 32: return launch_gemm<32, false>(tmA, tmB, tmC, p, st);
        case 64: return launch_gemm<64, false>(tmA, tmB, tmC, p, st);
        case 128: return launch_gemm<128, false>(tmA, tmB, tmC, p, st);
        default: return launch_gemm<256, false>(tmA, tmB, tmC, p, st);
    }
}

// dx[n,h,w,ci] = sum_{i,j,co} dy[n, h-i+1, w-j+1, co] * W[i,j,ci,co].  dy: bf16 [N,H,W,ld_dy] (ld_dy >= Cout, the engine's
// 8-element padded cast of the fp32 logits gradient); wd: bf16 [Cin, 9*kp] K-major with kp = 64*ceil(Cout/64) columns per
// tap, wd[ci, tap*kp + co] = W[tap, ci, co], zero beyond Cout.
extern "C" int dlv3p_conv3x3_same_dgrad_bf16(const void* dy, int64_t ld_dy, const void* wd, void* dx, int N, int H, int W,
                                             int Cin, int Cout, void* stream) {
    int rc = same_check("conv3x3_same_dgrad", dy, wd, dx, N, H, W, Cin, Cout);
    if (rc) return rc;
    DLV3P_REQUIRE(ld_dy >= Cout && (ld_dy % 8) == 0 && aligned16(dx), DLV3P_ERR_ALIGN,
                  "conv3x3_same_dgrad: ld_dy >= Cout, a multiple of 8, dx 16-byte aligned (ld_dy=%lld)", (long long)ld_dy);
    const int kbr = cdiv(Cout, kBlockK), kp = kbr * kBlockK;
    // fewest padded output columns: 304 input channels = 3 x 128 rather than 2 x 256
    int bn = 256;
    if (Cin <= 32) bn = 32; else if (Cin <= 64) bn = 64; else if (Cin <= 128) bn = 128;
    else if (cdiv(Cin, 128) * 128 < cdiv(Cin, 256) * 256) bn = 128;
    CUtensorMap tmA, tmB, tmC;
    {
        const long long dims[4] = {Cout, W, H, N};
        const long long str[3] = {2LL * ld_dy, 2LL * W * ld_dy, 2LL * H * W * ld_dy};
        const int box[4] = {kBlockK, kBlockM, 1, 1};
        rc = make_tmap_nd(&tmA, dy, 4, dims, str, box);
        if (rc) return rc;
    }
    rc = make_tmap(&tmB, wd, 9LL * kp, Cin, 9LL * kp, kBlockK, bn);
    if (rc) return rc;
    {
        const long long dims[3] = {Cin, W, (long long)N * H};
        const long long str[2] = {2LL * Cin, 2LL * W * Cin};
        const int box[3] = {64, 32, 1};
        rc = make_tmap_nd(&tmC, dx, 3, dims, str, box);
        if (rc) return rc;
    }
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.cv_tpr = cdiv(W, kBlockM);
    p.m_tiles = N * H * p.cv_tpr; p.n_tiles = cdiv(Cin, bn);
    p.M = p.m_tiles * kBlockM; p.N = Cin; p.K = 9 * kp;
    p.C = dx; p.ldc = Cin; p.c_dtype = DLV3P_BF16; p.splits = 1;
    p.tma_store = bn >= 64 ? 1 : 0;
    p.conv_mode = 5; p.cv_rows_in = H; p.cv_rows_out = H; p.cv_width = W; p.cv_rlimit = W; p.cv_kbr = kbr; p.cv_run = kp;
    cudaStream_t st = (cudaStream_t)stream;
    switch (bn) {
        case 32: return launch_gemm<32, false>(tmA, tmB, tmC, p, st);
        case 64: return launch_gemm<64, false>(tmA, tmB, tmC, p, st);
        case 128: return launch_gemm<128, false>(tmA, tmB, tmC, p, st);
        default: return launch_gemm<256, false>(tmA, tmB, tmC, p, st);
    }
}

// dw[tap, ci, co] += sum_{n,h,w} x[n, h+i-1, w+j-1, ci] * dy[n,h,w,co]   (fp32 HWIO [3,3,Cin,Cout], accumulated)
extern "C" int dlv3p_conv3x3_same_wgrad_bf16(const void* x, const void* dy, int64_t ld_dy, float* dw, int N, int H, int W,
                                             int Cin, int Cout, void* stream) {
    int rc = same_check("conv3x3_same_wgrad", x, dy, dw, N, H, W, Cin, Cout);
    if (rc) return rc;
    DLV3P_REQUIRE(ld_dy >= Cout && (ld_dy % 8) == 0 && Cout <= 256, DLV3P_ERR_ALIGN,
                  "conv3x3_same_wgrad: ld_dy >= Cout, a multiple of 8, Cout <= 256 (ld_dy=%lld Cout=%d)", (long long)ld_dy, Cout);
    const int bn = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
    const int cit = cdiv(Cin, kBlockM);                   // 128-channel row tiles per tap
    CUtensorMap tmA, tmB, tmC;
    {
        const long long dims[4] = {Cin, W, H, N};
        const long long str[3] = {2LL * Cin, 2LL * W * Cin, 2LL * H * W * Cin};
        const int box[4] = {64, kBlockK, 1, 1};
        rc = make_tmap_nd(&tmA, x, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        const long long dims[3] = {Cout, W, (long long)N * H};
        const long long str[2] = {2LL * ld_dy, 2LL * W * ld_dy};
        const int box[3] = {64, kBlockK, 1};
        rc = make_tmap_nd(&tmB, dy, 3, dims, str, box);
        if (rc) return rc;
    }
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.tma_store = ((Cout % 4) == 0 && aligned16(dw)) ? 1 : 0;      // TMA strides are multiples of 16 bytes
    tmC = tmA;
    if (p.tma_store) {
        const long long dims[3] = {Cout, Cin, 9};
        const long long str[2] = {4LL * Cout, 4LL * Cin * Cout};
        const int box[3] = {32, 32, 1};
        rc = make_tmap_nd(&tmC, dw, 3, dims, str, box, /*f32=*/true);
        if (rc) return rc;
    }
    p.cv_tpr = cdiv(W, kBlockK);                          // 64-pixel reduction blocks per image row
    const int total_kb = N * H * p.cv_tpr;
    p.m_tiles = 9 * cit; p.n_tiles = cdiv(Cout, bn);
    p.M = total_kb * kBlockK; p.N = Cout; p.K = p.m_tiles * kBlockM;
    p.C = dw; p.ldc = Cout; p.c_dtype = DLV3P_F32;
    const int tiles = p.m_tiles * p.n_tiles;
    int splits = (2 * kNumSMs) / tiles; if (splits < 1) splits = 1; if (splits > total_kb) splits = total_kb;
    p.kb_per_split = cdiv(total_kb, splits);
    p.splits = cdiv(total_kb, p.kb_per_split);
    p.conv_mode = 6; p.cv_rows_in = H; p.cv_rows_out = H; p.cv_width = W; p.cv_rlimit = Cin; p.cv_kbr = cit;
    cudaStream_t st = (cudaStream_t)stream;
    switch (bn) {
        case 64: return launch_gemm<64, true>(tmA, tmB, tmC, p, st);
        case 128: return launch_gemm<128, true>(tmA, tmB, tmC, p, st);
        default: return launch_gemm<256, true>(tmA, tmB, tmC, p, st);
    }
}
