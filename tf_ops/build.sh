#!/usr/bin/env bash
# Build the TensorFlow custom-op library over libdlv3p.so.  Needs a machine with TensorFlow >= 2.4 (this image has
# none: the script exits 3 and says so).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
if ! python -c "import tensorflow" 2>/dev/null; then
  echo "tf_ops/build.sh: TensorFlow is not importable here; nothing built" >&2
  exit 3
fi
CFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))')
LFLAGS=$(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))')
LIBDIR="${HERE}/../deeplabv3plus_keras_b200"
[ -f "${LIBDIR}/libdlv3p.so" ] || bash "${LIBDIR}/csrc/build.sh"
g++ -std=c++17 -shared -fPIC -O2 -DGOOGLE_CUDA=1 "${HERE}/dlv3p_tf_ops.cc" "${HERE}/dlv3p_tf_raw_ops.cc" -o "${HERE}/libdlv3p_tf_ops.so" \
    ${CFLAGS} -I/usr/local/cuda/include ${LFLAGS} -L"${LIBDIR}" -ldlv3p -L/usr/local/cuda/lib64 -lcudart \
    -Wl,-rpath,"${LIBDIR}"
echo "built ${HERE}/libdlv3p_tf_ops.so"
