// TensorFlow custom-op registration of the libdlv3p C-ABI (include/dlv3p.h) — the `tf.load_op_library` route named
// in BASELINE.json:north_star.  Each OpKernel forwards the raw device pointers of its inputs/outputs and TensorFlow's
// own CUDA stream to one C entry point; no arithmetic happens here.
//
// NOT BUILT IN THIS IMAGE: TensorFlow (headers + libtensorflow_framework) is not installable offline.
// Build on a machine with TF >= 2.4 with tf_ops/build.sh; the Python side (gradients, Keras layer subclasses that keep
// the reference's constructor signatures, ss.py:795-954) is tf_ops/dlv3p_tf.py.
//
// Ops registered (the ones the reference's layers reach on the hot path):
//   Dlv3pDepthwiseConv3x3 / ...BackpropInput / ...BackpropFilter   <- DepthwiseConv2dNative* (SeparableConv2D, ss.py:823-830)
//   Dlv3pPointwiseConv                                              <- Conv2D 1x1 (+ folded BN + ReLU) (ss.py:814-818, 833-838)
//   Dlv3pPointwiseConvBackpropFilter                                <- Conv2DBackpropFilter of the 1x1
//   Dlv3pUpsampleSoftmaxCbLoss / ...Grad                            <- ResizeBilinear + Softmax + class_balanced_loss
//                                                                      (ss.py:904-909, 438-447)
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/util/gpu_kernel_helper.h"

#include "../include/dlv3p.h"

namespace tf = tensorflow;
using tf::shape_inference::InferenceContext;

namespace {

inline void* stream_of(tf::OpKernelContext* ctx) {
    return reinterpret_cast<void*>(ctx->eigen_gpu_device().stream());
}
inline int dtype_of(const tf::Tensor& t) { return t.dtype() == tf::DT_BFLOAT16 ? DLV3P_BF16 : DLV3P_F32; }
inline const void* ptr(const tf::Tensor& t) { return t.tensor_data().data(); }
inline void* mptr(tf::Tensor* t) { return const_cast<char*>(t->tensor_data().data()); }

#define DLV3P_TF_CALL(ctx, expr)                                                                       \
    do {                                                                                               \
        int rc__ = (expr);                                                                             \
        OP_REQUIRES(ctx, rc__ == 0, tf::errors::InvalidArgument("libdlv3p: ", dlv3p_last_error()));    \
    } while (0)

// TF 'SAME': out = ceil(n/s), pad_before = max((out-1)*s + (k-1)*d + 1 - n, 0) / 2
inline void same_geometry(int n, int s, int d, int* out, int* before) {
    *out = (n + s - 1) / s;
    int total = (*out - 1) * s + 2 * d + 1 - n;
    if (total < 0) total = 0;
    *before = total / 2;
}
inline void geometry(const std::string& padding, int n, int s, int d, int* out, int* before) {
    if (padding == "SAME") { same_geometry(n, s, d, out, before); }
    else { *out = (n - (2 * d + 1)) / s + 1; *before = 0; }
}

struct DwAttrs {
    int stride, dil_h, dil_w, act;
    std::string padding;
    explicit DwAttrs(tf::OpKernelConstruction* c) {
        OP_REQUIRES_OK(c, c->GetAttr("stride", &stride));
        OP_REQUIRES_OK(c, c->GetAttr("dilation_h", &dil_h));
        OP_REQUIRES_OK(c, c->GetAttr("dilation_w", &dil_w));
        OP_REQUIRES_OK(c, c->GetAttr("padding", &padding));
        OP_REQUIRES_OK(c, c->GetAttr("pre_activation", &act));
    }
};

}  // namespace

#define DW_ATTRS                                                                                        \
    .Attr("T: {float, bfloat16}").Attr("stride: int = 1").Attr("dilation_h: int = 1")                   \
    .Attr("dilation_w: int = 1").Attr("padding: {'SAME', 'VALID'} = 'SAME'").Attr("pre_activation: int = 0")

REGISTER_OP("Dlv3pDepthwiseConv3x3")
    .Input("x: T").Input("filter: float") DW_ATTRS.Output("y: T")
    .SetShapeFn([](InferenceContext* c) { c->set_output(0, c->UnknownShapeOfRank(4)); return tf::Status(); });
REGISTER_OP("Dlv3pDepthwiseConv3x3BackpropInput")
    .Input("x: T").Input("filter: float").Input("dy: T") DW_ATTRS.Output("dx: T")
    .SetShapeFn([](InferenceContext* c) { c->set_output(0, c->input(0)); return tf::Status(); });
REGISTER_OP("Dlv3pDepthwiseConv3x3BackpropFilter")
    .Input("x: T").Input("dy: T") DW_ATTRS.Output("dfilter: float")
    .SetShapeFn([](InferenceContext* c) { c->set_output(0, c->UnknownShapeOfRank(4)); return tf::Status(); });

class DwFwdOp : public tf::OpKernel {
  public:
    explicit DwFwdOp(tf::OpKernelConstruction* c) : OpKernel(c), a_(c) {}
    void Compute(tf::OpKernelContext* ctx) override {
        const tf::Tensor& x = ctx->input(0);
        const tf::Tensor& w = ctx->input(1);              // [3,3,C,1] fp32 (Keras depthwise_kernel)
        const int N = x.dim_size(0), H = x.dim_size(1), W = x.dim_size(2), C = x.dim_size(3);
        int Ho, Wo, pt, pl;
        geometry(a_.padding, H, a_.stride, a_.dil_h, &Ho, &pt);
        geometry(a_.padding, W, a_.stride, a_.dil_w, &Wo, &pl);
        tf::Tensor* y = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({N, Ho, Wo, C}), &y));
        DLV3P_TF_CALL(ctx, dlv3p_dwconv3x3_fwd(ptr(x), w.flat<float>().data(), mptr(y), N, H, W, C, a_.stride, a_.dil_h,
                                               a_.dil_w, pt, pl, Ho, Wo, nullptr, nullptr, a_.act, dtype_of(x),
                                               stream_of(ctx)));
    }
  private:
    DwAttrs a_;
};

class DwBackpropInputOp : public tf::OpKernel {
  public:
    explicit DwBackpropInputOp(tf::OpKernelConstruction* c) : OpKernel(c), a_(c) {}
    void Compute(tf::OpKernelContext* ctx) override {
        const tf::Tensor& x = ctx->input(0);
        const tf::Tensor& w = ctx->input(1);
        const tf::Tensor& dy = ctx->input(2);
        const int N = x.dim_size(0), H = x.dim_size(1), W = x.dim_size(2), C = x.dim_size(3);
        int Ho, Wo, pt, pl;
        geometry(a_.padding, H, a_.stride, a_.dil_h, &Ho, &pt);
        geometry(a_.padding, W, a_.stride, a_.dil_w, &Wo, &pl);
        tf::Tensor* dx = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &dx));
        DLV3P_TF_CALL(ctx, dlv3p_dwconv3x3_dgrad(ptr(dy), w.flat<float>().data(), mptr(dx), N, H, W, C, a_.stride,
                                                 a_.dil_h, a_.dil_w, pt, pl, Ho, Wo, a_.act ? ptr(x) : nullptr, nullptr,
                                                 nullptr, a_.act, nullptr, dtype_of(x), stream_of(ctx)));
    }
  private:
    DwAttrs a_;
};

class DwBackpropFilterOp : public tf::OpKernel {
  public:
    explicit DwBackpropFilterOp(tf::OpKernelConstruction* c) : OpKernel(c), a_(c) {}
    void Compute(tf::OpKernelContext* ctx) override {
        const tf::Tensor& x = ctx->input(0);
        const tf::Tensor& dy = ctx->input(1);
        const int N = x.dim_size(0), H = x.dim_size(1), W = x.dim_size(2), C = x.dim_size(3);
        int Ho, Wo, pt, pl;
        geometry(a_.padding, H, a_.stride, a_.dil_h, &Ho, &pt);
        geometry(a_.padding, W, a_.stride, a_.dil_w, &Wo, &pl);
        tf::Tensor* dw = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({3, 3, C, 1}), &dw));
        float* dwp = dw->flat<float>().data();
        cudaMemsetAsync(dwp, 0, sizeof(float) * 9 * C, (cudaStream_t)stream_of(ctx));   // the entry point accumulates
        DLV3P_TF_CALL(ctx, dlv3p_dwconv3x3_wgrad(ptr(x), ptr(dy), dwp, N, H, W, C, a_.stride, a_.dil_h, a_.dil_w, pt, pl,
                                                 Ho, Wo, nullptr, nullptr, a_.act, dtype_of(x), stream_of(ctx)));
    }
  private:
    DwAttrs a_;
};

REGISTER_KERNEL_BUILDER(Name("Dlv3pDepthwiseConv3x3").Device(tf::DEVICE_GPU), DwFwdOp);
REGISTER_KERNEL_BUILDER(Name("Dlv3pDepthwiseConv3x3BackpropInput").Device(tf::DEVICE_GPU), DwBackpropInputOp);
REGISTER_KERNEL_BUILDER(Name("Dlv3pDepthwiseConv3x3BackpropFilter").Device(tf::DEVICE_GPU), DwBackpropFilterOp);

// ---- pointwise 1x1 conv (+ folded BatchNormalization + ReLU/ReLU6) as a bf16 tcgen05 GEMM ------------------------
// x [N,H,W,K] bf16; w_t [Cout,K] bf16 (the Keras kernel [1,1,K,Cout] transposed once on the Python side);
// scale/shift [Cout] fp32 (gamma*rsqrt(var+eps), beta - mean*scale) or empty tensors for "no BN".
REGISTER_OP("Dlv3pPointwiseConv")
    .Input("x: bfloat16").Input("w_t: bfloat16").Input("scale: float").Input("shift: float")
    .Attr("activation: int = 0").Output("y: bfloat16")
    .SetShapeFn([](InferenceContext* c) { c->set_output(0, c->UnknownShapeOfRank(4)); return tf::Status(); });
// dW[K,Cout] fp32 = X^T dY
REGISTER_OP("Dlv3pPointwiseConvBackpropFilter")
    .Input("x: bfloat16").Input("dy: bfloat16").Output("dw: float")
    .SetShapeFn([](InferenceContext* c) { c->set_output(0, c->UnknownShapeOfRank(2)); return tf::Status(); });

class PointwiseOp : public tf::OpKernel {
  public:
    explicit PointwiseOp(tf::OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("activation", &act_)); }
    void Compute(tf::OpKernelContext* ctx) override {
        const tf::Tensor& x = ctx->input(0);
        const tf::Tensor& wt = ctx->input(1);
        const tf::Tensor& sc = ctx->input(2);
        const tf::Tensor& sh = ctx->input(3);
        const int N = x.dim_size(0), H = x.dim_size(1), W = x.dim_size(2), K = x.dim_size(3);
        const int Cout = wt.dim_size(0);
        OP_REQUIRES(ctx, wt.dim_size(1) == K && K % 8 == 0, tf::errors::InvalidArgument("w_t must be [Cout,K], K % 8 == 0"));
        tf::Tensor* y = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({N, H, W, Cout}), &y));
        const bool bn = sc.NumElements() == Cout;
        DLV3P_TF_CALL(ctx, dlv3p_gemm_bf16(ptr(x), K, ptr(wt), K, mptr(y), Cout, N * H * W, Cout, K, DLV3P_BF16,
                                           bn ? sc.flat<float>().data() : nullptr, bn ? sh.flat<float>().data() : nullptr,
                                           act_, nullptr, 0, nullptr, stream_of(ctx)));
    }
  private:
    int act_;
};

class PointwiseBackpropFilterOp : public tf::OpKernel {
  public:
    explicit PointwiseBackpropFilterOp(tf::OpKernelConstruction* c) : OpKernel(c) {}
    void Compute(tf::OpKernelContext* ctx) override {
        const tf::Tensor& x = ctx->input(0);
        const tf::Tensor& dy = ctx->input(1);
        const int K = x.dim_size(3), Cout = dy.dim_size(3);
        const int M = x.NumElements() / K;
        tf::Tensor* dw = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({K, Cout}), &dw));
        float* p = dw->flat<float>().data();
        cudaMemsetAsync(p, 0, sizeof(float) * K * Cout, (cudaStream_t)stream_of(ctx));
        DLV3P_TF_CALL(ctx, dlv3p_gemm_wgrad_bf16(ptr(x), K, ptr(dy), Cout, p, Cout, M, K, Cout, stream_of(ctx)));
    }
};
REGISTER_KERNEL_BUILDER(Name("Dlv3pPointwiseConv").Device(tf::DEVICE_GPU), PointwiseOp);
REGISTER_KERNEL_BUILDER(Name("Dlv3pPointwiseConvBackpropFilter").Device(tf::DEVICE_GPU), PointwiseBackpropFilterOp);

// ---- fused decoder tail ------------------------------------------------------------------------------------------
REGISTER_OP("Dlv3pUpsampleSoftmaxCbLoss")
    .Input("logits: float").Input("labels: int32").Input("pos_weights: float").Input("neg_weights: float")
    .Attr("factor: int").Attr("epsilon: float = 1e-7").Output("loss_sum: float").Output("dlogits: float")
    .SetShapeFn([](InferenceContext* c) { c->set_output(0, c->Scalar()); c->set_output(1, c->input(0)); return tf::Status(); });

class TailOp : public tf::OpKernel {
  public:
    explicit TailOp(tf::OpKernelConstruction* c) : OpKernel(c) {
        OP_REQUIRES_OK(c, c->GetAttr("factor", &f_));
        OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
    }
    void Compute(tf::OpKernelContext* ctx) override {
        const tf::Tensor& z = ctx->input(0);
        const tf::Tensor& lab = ctx->input(1);
        const int N = z.dim_size(0), H = z.dim_size(1), W = z.dim_size(2), C = z.dim_size(3);
        tf::Tensor *loss = nullptr, *dz = nullptr;
        OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({}), &loss));
        OP_REQUIRES_OK(ctx, ctx->allocate_output(1, z.shape(), &dz));
        cudaStream_t st = (cudaStream_t)stream_of(ctx);
        cudaMemsetAsync(loss->flat<float>().data(), 0, sizeof(float), st);
        cudaMemsetAsync(dz->flat<float>().data(), 0, sizeof(float) * z.NumElements(), st);
        const float gscale = 1.0f / (float)((long long)N * H * f_ * W * f_);        // Keras mean over B*H*W
        DLV3P_TF_CALL(ctx, dlv3p_upsample_softmax_cbloss_fwd_bwd(
                               z.flat<float>().data(), lab.flat<tf::int32>().data(), ctx->input(2).flat<float>().data(),
                               ctx->input(3).flat<float>().data(), eps_, N, H, W, C, f_, gscale,
                               loss->flat<float>().data(), dz->flat<float>().data(), st));
    }
  private:
    int f_;
    float eps_;
};
REGISTER_KERNEL_BUILDER(Name("Dlv3pUpsampleSoftmaxCbLoss").Device(tf::DEVICE_GPU), TailOp);
