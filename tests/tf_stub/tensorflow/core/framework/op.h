#pragma once
#include "tensorflow/core/framework/shape_inference.h"
namespace tensorflow {
struct OpDefBuilderStub {
    OpDefBuilderStub& Attr(const char*);
    OpDefBuilderStub& Input(const char*);
    OpDefBuilderStub& Output(const char*);
    OpDefBuilderStub& SetShapeFn(std::function<Status(shape_inference::InferenceContext*)>);
};
OpDefBuilderStub RegisterOpStub(const char*);
#define REGISTER_OP(name) static ::tensorflow::OpDefBuilderStub TF_STUB_CAT(op_reg_, __COUNTER__) = ::tensorflow::RegisterOpStub(name)
}  // namespace tensorflow
