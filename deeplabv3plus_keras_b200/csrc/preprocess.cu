// Input pipeline of the reference's keras Sequences on the GPU: the per-sample work of
// TrainingSequencePascalVOC2012Ext.__getitem__ (ss.py:1528-1560) — normalise the decoded uint8 image to (-1, 1),
// aspect-preserving LINEAR resize to the square network input with scipy.ndimage.affine_transform semantics
// (resize(), ss.py:130-195: matrix diag(1/fy, 1/fx), no offset, order=1, mode='nearest'), zero padding to size x size
// (resize_image_to_target_symmeric_size, ss.py:198-280, including its swapped left/right padding for portrait
// images), and the same for the uint8 label map (classes above num_classes-1 cleared before and after, linear
// interpolation of class ids rounded as scipy rounds integer outputs).  In the reference this is a single Python
// thread per step (workers=0) plus a 262 144-iteration Python loop per image for the one-hot (ss.py:357-358) — its
// real bottleneck (4 s/step, nb cell 29).  Here: one launch per batch, raw uint8 over PCIe instead of fp32/one-hot.
// Coordinates and interpolation are evaluated in fp64 exactly as scipy does, so label maps are bit-exact.
#include "common.cuh"

namespace dlv3p {

struct PreprocEntry {              // 48 bytes, part of the C-ABI (dlv3p_preprocess_*_batch)
    const uint8_t* src;            // decoded sample, HWC uint8 (C = 3 for images, 1 for labels), device memory
    double inv_fy, inv_fx;         // 1/fy, 1/fx with fy = h_p / h, fx = w_p / w (evaluated by the caller in fp64)
    int h, w;                      // source extent
    int hp, wp;                    // resized extent (before padding)
    int off_y, off_x;              // zero padding above / left of the resized sample
};
static_assert(sizeof(PreprocEntry) == 48, "table layout is part of the C-ABI");

// scipy order-1 spline sampling with mode='nearest' along one axis: clamp the coordinate, two taps, weights (1-t, t)
__device__ __forceinline__ void axis_taps(double inv_f, int o, int len, int& i0, int& i1, double& t) {
    double c = inv_f * (double)o;
    if (c > (double)(len - 1)) c = (double)(len - 1);
    if (c < 0.0) c = 0.0;
    const double f = floor(c);
    i0 = (int)f;
    i1 = min(i0 + 1, len - 1);
    t = c - f;
}

template <typename TO>
__global__ void __launch_bounds__(256)
preprocess_image_kernel(const PreprocEntry* __restrict__ table, TO* __restrict__ out, int S) {
    const PreprocEntry e = table[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * S) return;
    const int oy = idx / S, ox = idx % S;
    TO* dst = out + ((long long)blockIdx.y * S * S + idx) * 3;
    const int py = oy - e.off_y, px = ox - e.off_x;
    if (py < 0 || py >= e.hp || px < 0 || px >= e.wp) {
        dst[0] = from_f<TO>(0.f); dst[1] = from_f<TO>(0.f); dst[2] = from_f<TO>(0.f);
        return;
    }
    int y0, y1, x0, x1; double ty, tx;
    axis_taps(e.inv_fy, py, e.h, y0, y1, ty);
    axis_taps(e.inv_fx, px, e.w, x0, x1, tx);
    const uint8_t* r0 = e.src + (long long)y0 * e.w * 3;
    const uint8_t* r1 = e.src + (long long)y1 * e.w * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // 2.0 * (image / 255 - 0.5) in fp64 before the resize, as at ss.py:1532
        const double a = 2.0 * ((double)r0[x0 * 3 + c] / 255.0 - 0.5), b = 2.0 * ((double)r0[x1 * 3 + c] / 255.0 - 0.5);
        const double p = 2.0 * ((double)r1[x0 * 3 + c] / 255.0 - 0.5), q = 2.0 * ((double)r1[x1 * 3 + c] / 255.0 - 0.5);
        const double v = (1.0 - ty) * ((1.0 - tx) * a + tx * b) + ty * ((1.0 - tx) * p + tx * q);
        dst[c] = from_f<TO>((float)v);
    }
}

__global__ void __launch_bounds__(256)
preprocess_label_kernel(const PreprocEntry* __restrict__ table, int32_t* __restrict__ out, int S, int num_classes) {
    const PreprocEntry e = table[blockIdx.y];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * S) return;
    const int oy = idx / S, ox = idx % S;
    int32_t* dst = out + (long long)blockIdx.y * S * S + idx;
    const int py = oy - e.off_y, px = ox - e.off_x;
    if (py < 0 || py >= e.hp || px < 0 || px >= e.wp) { *dst = 0; return; }
    int y0, y1, x0, x1; double ty, tx;
    axis_taps(e.inv_fy, py, e.h, y0, y1, ty);
    axis_taps(e.inv_fx, px, e.w, x0, x1, tx);
    auto cls = [&](int y, int x) {                       // label[label > num_classes - 1] = 0   (ss.py:1540)
        const int v = e.src[(long long)y * e.w + x];
        return (double)(v > num_classes - 1 ? 0 : v);
    };
    const double v = (1.0 - ty) * ((1.0 - tx) * cls(y0, x0) + tx * cls(y0, x1)) +
                     ty * ((1.0 - tx) * cls(y1, x0) + tx * cls(y1, x1));
    // scipy writes integer outputs as (uint8)(v + 0.5) for v > 0 (ni_interpolation.c, CASE_INTERP_OUT_UINT)
    int r = v > 0.0 ? (int)(v + 0.5) : 0;
    if (r > 255) r = 255;
    if (r > num_classes - 1) r = 0;                      // the second clear, ss.py:1553
    *dst = r;
}

}  // namespace dlv3p

using namespace dlv3p;

extern "C" int dlv3p_preprocess_image_batch(const void* table, int count, void* out, int S, int out_dtype, void* stream) {
    DLV3P_REQUIRE(table && out && count > 0 && S > 0, DLV3P_ERR_SHAPE, "preprocess_image_batch: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(cdiv((long long)S * S, 256), count);
    if (out_dtype == DLV3P_F32)
        preprocess_image_kernel<float><<<grid, 256, 0, st>>>((const PreprocEntry*)table, (float*)out, S);
    else if (out_dtype == DLV3P_BF16)
        preprocess_image_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const PreprocEntry*)table, (__nv_bfloat16*)out, S);
    else { set_error("preprocess_image_batch: unsupported dtype %d", out_dtype); return DLV3P_ERR_DTYPE; }
    return check_launch("preprocess_image_batch");
}

extern "C" int dlv3p_preprocess_label_batch(const void* table, int count, int32_t* out, int S, int num_classes,
                                            void* stream) {
    DLV3P_REQUIRE(table && out && count > 0 && S > 0 && num_classes > 0 && num_classes <= 256, DLV3P_ERR_SHAPE,
                  "preprocess_label_batch: bad arguments");
    const dim3 grid(cdiv((long long)S * S, 256), count);
    preprocess_label_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const PreprocEntry*)table, out, S, num_classes);
    return check_launch("preprocess_label_batch");
}
