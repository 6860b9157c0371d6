#!/usr/bin/env bash
# Builds _roi3d_tf_ops.so (TensorFlow op registration + GPU kernels that forward to libroi3d_b200.so)
# and lays out the four top-level packages of the reference wheel (top_level.txt) so that
# core/custom_op/custom_op.py:22-25 imports unchanged.  Needs TensorFlow built for CUDA >= 12.8.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
PKG="$(dirname "$HERE")"
ROOT="$(dirname "$PKG")"
OUT="${1:-$HERE/dist}"
TF_CFLAGS=( $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_compile_flags()))') )
TF_LFLAGS=( $(python -c 'import tensorflow as tf; print(" ".join(tf.sysconfig.get_link_flags()))') )
CUDA_HOME="${CUDA_HOME:-/usr/local/cuda}"
python -c "import sys; sys.path.insert(0, '$ROOT'); import __graft_entry__ as g; g.build()"
mkdir -p "$OUT"
g++ -std=c++17 -O2 -shared -fPIC "$HERE/roi3d_tf_ops.cc" -o "$OUT/_roi3d_tf_ops.so" \
    -I"$ROOT/include" -I"$CUDA_HOME/include" "${TF_CFLAGS[@]}" -DGOOGLE_CUDA=1 \
    -L"$PKG/lib" -lroi3d_b200 -Wl,-rpath,'$ORIGIN' -L"$CUDA_HOME/lib64" -lcudart "${TF_LFLAGS[@]}"
cp "$PKG/lib/libroi3d_b200.so" "$OUT/"
for pkg in crop_and_resize_3d crop_and_resize_3d_grad_image crop_and_resize_3d_grad_boxes non_max_suppression_3d; do
  mkdir -p "$OUT/$pkg"
  cp "$HERE/packages/$pkg.py" "$OUT/$pkg/__init__.py"
done
cp "$HERE/packages/_roi3d_loader.py" "$OUT/"
# ppc64le loader (core/custom_op/ppc64le_custom_op.py:21-30): four libraries, each registering ONE op, under the file
# names it tf.load_op_library()s.  Copy (or symlink) $OUT/ppc64le/*.so into the reference's core/custom_op/.
mkdir -p "$OUT/ppc64le"
n=0
for lib in crop_and_resize_3d_op crop_and_resize_3d_grad_image_op crop_and_resize_3d_grad_boxes_op non_max_suppression_3d_op; do
  n=$((n + 1))
  g++ -std=c++17 -O2 -shared -fPIC "$HERE/roi3d_tf_ops.cc" -o "$OUT/ppc64le/$lib.so" -DROI3D_TF_SINGLE_OP=$n \
      -I"$ROOT/include" -I"$CUDA_HOME/include" "${TF_CFLAGS[@]}" -DGOOGLE_CUDA=1 \
      -L"$PKG/lib" -lroi3d_b200 -Wl,-rpath,'$ORIGIN/..' -L"$CUDA_HOME/lib64" -lcudart "${TF_LFLAGS[@]}"
done
echo "built: add $OUT to PYTHONPATH (it replaces tensorflow_nms_car_3d-0.1.0)"
