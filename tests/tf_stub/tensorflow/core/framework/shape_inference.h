#pragma once
#include "tensorflow/core/framework/op_kernel.h"
namespace tensorflow { namespace shape_inference {
struct ShapeHandle {};
class InferenceContext {
  public:
    ShapeHandle input(int i);
    void set_output(int i, ShapeHandle s);
    ShapeHandle UnknownShapeOfRank(int r);
    ShapeHandle Scalar();
};
}}  // namespace tensorflow::shape_inference
