// Test-only stand-in for the slice of the TensorFlow C++ op API that tf_ops/*.cc uses.  TensorFlow cannot be installed in
// the build image, so tests/test_tf_ops.py type-checks the custom-op sources against these declarations
// (g++ -fsyntax-only): every call into libdlv3p is checked against include/dlv3p.h's prototypes, every TF construct
// against the signature it has in TF 2.4.  Nothing here is compiled into the product.
#pragma once
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <string>

namespace tensorflow {
typedef std::int64_t int64;
typedef std::uint64_t uint64;
typedef std::int32_t int32;
typedef std::uint8_t uint8;
struct bfloat16 { std::uint16_t v; };
enum DataType { DT_FLOAT = 1, DT_BFLOAT16 = 14 };
extern const char* const DEVICE_GPU;

class Status {
  public:
    Status() {}
    bool ok() const { return true; }
};
namespace errors {
template <typename... A> Status InvalidArgument(A...) { return Status(); }
}  // namespace errors

struct StringPiece { const char* data() const; };
class TensorShape {
  public:
    TensorShape() {}
    TensorShape(std::initializer_list<int64> dims) { (void)dims; }
};
template <typename T> struct FlatView { T* data() const; };
template <typename T> struct ScalarView { T& operator()() const; };
class Tensor {
  public:
    int64 NumElements() const;
    int64 dim_size(int i) const;
    DataType dtype() const;
    const TensorShape& shape() const;
    StringPiece tensor_data() const;
    template <typename T> FlatView<T> flat();
    template <typename T> FlatView<const T> flat() const;
    template <typename T> ScalarView<const T> scalar() const;
};

struct GpuStreamHolder { void* stream() const; };
class OpKernelConstruction {
  public:
    template <typename T> Status GetAttr(const char* name, T* value) const;
    void CtxFailure(const Status&);
    void CtxFailureWithWarning(const Status&);
};
class OpKernelContext {
  public:
    const Tensor& input(int i);
    void set_output(int i, const Tensor& t);
    Status allocate_output(int i, const TensorShape& s, Tensor** out);
    const GpuStreamHolder& eigen_gpu_device() const;
    void CtxFailure(const Status&);
};
class OpKernel {
  public:
    explicit OpKernel(OpKernelConstruction*) {}
    virtual ~OpKernel() {}
    virtual void Compute(OpKernelContext* ctx) = 0;
};

#define OP_REQUIRES(CTX, EXP, STATUS) do { if (!(EXP)) { (CTX)->CtxFailure((STATUS)); return; } } while (0)
#define OP_REQUIRES_OK(CTX, ...) do { ::tensorflow::Status s__(__VA_ARGS__); if (!s__.ok()) { (CTX)->CtxFailure(s__); return; } } while (0)

struct KernelDefBuilder {
    KernelDefBuilder& Device(const char*);
    KernelDefBuilder& HostMemory(const char*);
};
KernelDefBuilder Name(const char*);
#define TF_STUB_CAT2(a, b) a##b
#define TF_STUB_CAT(a, b) TF_STUB_CAT2(a, b)
#define REGISTER_KERNEL_BUILDER(builder, ...) \
    static ::tensorflow::OpKernel* TF_STUB_CAT(make_kernel_, __COUNTER__)(::tensorflow::OpKernelConstruction* c) { using namespace ::tensorflow; (void)(builder); return new __VA_ARGS__(c); }
}  // namespace tensorflow
