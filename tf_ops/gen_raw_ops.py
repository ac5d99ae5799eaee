"""Generate tf_ops/dlv3p_tf_raw_ops.cc: one TensorFlow op per C entry point of include/dlv3p.h.

`tf_ops/dlv3p_tf_ops.cc` holds the hand-written, TF-shaped ops (DepthwiseConv2dNative-like signatures with shape
inference and gradients).  This generator completes the op list mechanically so that EVERY libdlv3p entry point is
reachable from a TensorFlow graph through `tf.load_op_library`, with the C-ABI's own calling convention:

  * `const T*` argument        -> op input (device tensor); an empty tensor means NULL (optional operand);
  * non-const pointer argument -> op input AND output: the caller allocates the buffer (the C-ABI never allocates),
                                  the kernel writes / accumulates in place and the op forwards the same buffer as
                                  output `<name>_out`, which is what downstream ops must consume (ordering);
  * int / int64_t              -> op attribute (Keras graphs have static shapes: geometry is known at build time);
  * float / double / uint64_t  -> scalar input in HOST memory (learning rate, dropout seed ... change per step
                                  without retracing);
  * `void* stream`             -> TensorFlow's own CUDA stream of the op's device context.

Op name: dlv3p_dwconv3x3_fwd -> "Dlv3pRawDwconv3x3Fwd" (tf python: _lib.dlv3p_raw_dwconv3x3_fwd).
Run `python tf_ops/gen_raw_ops.py` after editing the header; tests/test_tf_ops.py checks the committed file is current.
TensorFlow is not installable in this image, so the output is delivered as source (tf_ops/build.sh compiles it
where TensorFlow exists).
"""
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, "..", "include", "dlv3p.h")
OUT = os.path.join(HERE, "dlv3p_tf_raw_ops.cc")
SKIP = {"dlv3p_last_error", "dlv3p_version", "dlv3p_device_arch", "dlv3p_set_pdl"}   # host-side queries, no tensors

PTR_TF = {"float": "float", "int32_t": "int32", "uint8_t": "uint8", "uint64_t": "uint64", "double": "double"}


def prototypes():
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for m in re.finditer(r"\bint\s+(dlv3p_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, raw = m.group(1), " ".join(m.group(2).split())
        if name in SKIP:
            continue
        args = []
        for a in raw.split(","):
            mm = re.match(r"\s*(const\s+)?(\w+)\s*(\*?)\s*(\w+)\s*$", a)
            assert mm, (name, a)
            args.append(dict(const=bool(mm.group(1)), base=mm.group(2), ptr=bool(mm.group(3)), name=mm.group(4)))
        yield name, args


def camel(name):
    return "Dlv3pRaw" + "".join(p.capitalize() for p in name[len("dlv3p_"):].split("_"))


def wrap(items, indent, sep="", width=118):
    """Join items, breaking lines at `width` columns (continuation lines indented by `indent`)."""
    out, cur = [], ""
    for it in items:
        if cur and len(indent) + len(cur) + len(sep) + len(it) > width:
            out.append(cur)
            cur = it
        else:
            cur = cur + (sep if cur else "") + it
    out.append(cur)
    first = out[0] if indent == "    " and out[0].startswith("REGISTER_OP") else indent + out[0]
    return "\n".join([first] + [indent + o for o in out[1:]])


def emit(name, args):
    op = camel(name)
    reg, body, call = [f'REGISTER_OP("{op}")'], [], []
    host, n_in, n_out, fwd = [], 0, 0, []
    lowered = [a["name"].lower() for a in args]
    for a in args:
        an = a["name"]
        # TensorFlow wants [a-z][a-z0-9_]* for tensor arguments; keep them unique case-insensitively (gemm: C vs c_dtype ok,
        # but M/N/K attrs next to pointers A/B/C)
        tn = an.lower() + ("_t" if lowered.count(an.lower()) > 1 or an != an.lower() else "")
        if a["ptr"] and an == "stream":
            call.append("stream_of(ctx)")
        elif a["ptr"]:
            if a["base"] == "void":
                reg.append(f'.Attr("T_{an}: type").Input("{tn}: T_{an}")')
                cast = "const void*" if a["const"] else "void*"
            else:
                reg.append(f'.Input("{tn}: {PTR_TF[a["base"]]}")')
                cast = ("const " if a["const"] else "") + a["base"] + "*"
            body.append(f"        {cast} p_{an} = reinterpret_cast<{cast}>(dev_ptr(ctx->input({n_in})));")
            if not a["const"]:
                tname = f"T_{an}" if a["base"] == "void" else PTR_TF[a["base"]]
                reg.append(f'.Output("{tn}_out: {tname}")')
                fwd.append((n_out, n_in))
                n_out += 1
            call.append(f"p_{an}")
            n_in += 1
        elif a["base"] in ("int", "int64_t"):
            reg.append(f'.Attr("{an}: int")')
            call.append(f"({a['base']}){an}_")
        else:                                         # float / double / uint64_t: host-memory scalar input
            tf_t = {"float": "float", "double": "double", "uint64_t": "uint64"}[a["base"]]
            c_t = {"float": "float", "double": "double", "uint64_t": "tensorflow::uint64"}[a["base"]]
            reg.append(f'.Input("{tn}: {tf_t}")')
            host.append(tn)
            body.append(f"        const {a['base']} s_{an} = ({a['base']})ctx->input({n_in}).scalar<{c_t}>()();")
            call.append(f"s_{an}")
            n_in += 1
    shape = "".join(f" c->set_output({o}, c->input({i}));" for o, i in fwd)
    reg.append(f".SetShapeFn([](InferenceContext* c) {{{shape} return tf::Status(); }});")
    attrs = [a for a in args if not a["ptr"] and a["base"] in ("int", "int64_t")]
    cls = op + "Op"
    lines = [wrap(reg, "    "), f"class {cls} : public tf::OpKernel {{", "  public:",
             f"    explicit {cls}(tf::OpKernelConstruction* c) : OpKernel(c) {{"]
    if attrs:
        lines.append(wrap([f'GA(c, "{a["name"]}", &{a["name"]}_);' for a in attrs], "        ", sep=" "))
    lines += ["    }", "    void Compute(tf::OpKernelContext* ctx) override {"]
    lines += body
    lines.append(wrap([f"DLV3P_TF_CALL(ctx, {name}({call[0]},"] + [c + "," for c in call[1:-1]] + [call[-1] + "));"],
                      "        ", sep=" "))
    for o, i in fwd:
        lines.append(f"        ctx->set_output({o}, ctx->input({i}));       // written in place: same buffer, new edge")
    lines += ["    }", "  private:"]
    if attrs:
        lines.append(wrap(["tf::int64"] + [a["name"] + "_," for a in attrs[:-1]] + [attrs[-1]["name"] + "_;"], "    ", sep=" "))
    lines.append("};")
    kb = f'REGISTER_KERNEL_BUILDER(Name("{op}").Device(tf::DEVICE_GPU)'
    for h in host:
        kb += f'.HostMemory("{h}")'
    lines.append(kb + f", {cls});")
    return "\n".join(lines)


PRELUDE = '''// GENERATED by tf_ops/gen_raw_ops.py from include/dlv3p.h — do not edit.
// One TensorFlow op per libdlv3p C entry point (the complete op list of the tf.load_op_library route; the TF-shaped
// ops with shape inference and gradients are in dlv3p_tf_ops.cc).  Calling convention: see gen_raw_ops.py.
// NOT BUILT IN THIS IMAGE (no TensorFlow); tf_ops/build.sh compiles it where TensorFlow >= 2.4 exists.
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/util/gpu_kernel_helper.h"

#include "../include/dlv3p.h"

namespace tf = tensorflow;
using tf::shape_inference::InferenceContext;

namespace {
inline void* stream_of(tf::OpKernelContext* ctx) { return reinterpret_cast<void*>(ctx->eigen_gpu_device().stream()); }
// an empty tensor stands for a NULL (optional) operand
inline void* dev_ptr(const tf::Tensor& t) {
    return t.NumElements() == 0 ? nullptr : const_cast<char*>(t.tensor_data().data());
}
#define GA(c, name, ptr) OP_REQUIRES_OK(c, c->GetAttr(name, ptr))
#define DLV3P_TF_CALL(ctx, expr)                                                                       \\
    do {                                                                                               \\
        int rc__ = (expr);                                                                             \\
        OP_REQUIRES(ctx, rc__ == 0, tf::errors::InvalidArgument("libdlv3p: ", dlv3p_last_error()));    \\
    } while (0)
}  // namespace
'''


def generate() -> str:
    parts = [PRELUDE]
    for name, args in prototypes():
        parts.append(f"// ---- {name} " + "-" * max(4, 100 - len(name)))
        parts.append(emit(name, args))
    return "\n\n".join(parts) + "\n"


if __name__ == "__main__":
    open(OUT, "w").write(generate())
    print("wrote", OUT, "with", sum(1 for _ in prototypes()), "ops")
