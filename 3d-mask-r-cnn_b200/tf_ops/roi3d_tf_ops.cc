// roi3d_tf_ops.cc -- TensorFlow registration of the B200 ROI hot path: the drop-in for the
// reference's wheel (core/custom_op/tensorflow_nms_car_3d-0.1.0-cp36-cp36m-linux_x86_64.whl).
//
// Same op names, inputs, outputs, attrs and validation messages as the reference's
// REGISTER_OP blocks (recovered from the wheel's .rodata, SURVEY.md section 8 rows a1, a4, a6,
// a7), so core/custom_op/custom_op.py:22-65 (imports + @ops.RegisterGradient("CropAndResize3D"))
// and core/models.py work unchanged.  Differences, all below the op registry:
//   * kernels are registered for DEVICE_GPU (float32) and only forward to the C ABI of
//     libroi3d_b200.so (include/roi3d.h) on TF's compute stream -- no CPU kernel, no fallback;
//   * crop_size / image_size / max_output_size are HostMemory inputs;
//   * NonMaxSuppression3D synchronises the stream once to learn the output length, like TF's own
//     GPU NonMaxSuppression.
//
// The development image has no TensorFlow: tests/test_tf_ops_source.py compiles this file with -fsyntax-only against
// the minimal stand-in headers of tests/tf_stub (current-TF signatures: absl::Status, int64_t), which catches removed
// aliases, typos and C-ABI argument mismatches but proves nothing about linking.  Build it for real where TF >= 2.15
// built for CUDA >= 12.8 is installed, with tf_ops/build_tf_ops.sh.
#define EIGEN_USE_GPU
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"
#include "tensorflow/core/framework/common_shape_fns.h"
#include "tensorflow/core/util/gpu_kernel_helper.h"   // GetGpuStream
#include "absl/status/status.h"                      // absl::OkStatus (tensorflow::OkStatus is gone from current TF)

#include <algorithm>
#include <cstdint>
#include <string>

#include "roi3d.h"

// ROI3D_TF_SINGLE_OP = 1..4 compiles the registration of ONE op only (CropAndResize3D, ...GradImage, ...GradBoxes,
// NonMaxSuppression3D): the reference's ppc64le loader tf.load_op_library()s four separate files and takes each op's
// wrapper from "its" module (core/custom_op/ppc64le_custom_op.py:21-30); a library that registered all four would give
// the 2nd-4th module an empty op list.  Undefined (default): all four ops in one library.
#if defined(ROI3D_TF_SINGLE_OP)
#define ROI3D_TF_HAS(n) (ROI3D_TF_SINGLE_OP == (n))
#else
#define ROI3D_TF_HAS(n) 1
#endif

namespace tf = tensorflow;
using tf::shape_inference::DimensionHandle;
using tf::shape_inference::InferenceContext;
using tf::shape_inference::ShapeHandle;

namespace {

tf::Status SetOutputToSizedImage3D(InferenceContext* c, DimensionHandle batch, int size_input_idx,
                                   DimensionHandle channels) {
  ShapeHandle size;
  TF_RETURN_IF_ERROR(c->WithRank(c->input(size_input_idx), 1, &size));
  DimensionHandle unused;
  TF_RETURN_IF_ERROR(c->WithValue(c->Dim(size, 0), 3, &unused));
  const tf::Tensor* size_tensor = c->input_tensor(size_input_idx);
  DimensionHandle h, w, d;
  if (size_tensor == nullptr) {
    h = c->UnknownDim(); w = c->UnknownDim(); d = c->UnknownDim();
  } else {
    if (size_tensor->dtype() != tf::DT_INT32)
      return tf::errors::InvalidArgument("Bad size input type for SetOutputToSizedImage: Expected DT_INT32 but got ",
                                         tf::DataTypeString(size_tensor->dtype()));
    auto v = size_tensor->vec<int32_t>();
    h = c->MakeDim(v(0)); w = c->MakeDim(v(1)); d = c->MakeDim(v(2));
  }
  c->set_output(0, c->MakeShape({batch, h, w, d, channels}));
  return absl::OkStatus();
}

int MethodFromName(const std::string& m) { return m == "nearest" ? ROI3D_METHOD_NEAREST : ROI3D_METHOD_TRILINEAR; }

tf::Status Roi3dStatus(int code, const char* what) {
  if (code == ROI3D_OK) return absl::OkStatus();
  if (code == ROI3D_EINVAL) return tf::errors::InvalidArgument(what, ": ", roi3d_strerror(code));
  if (code == ROI3D_EUNSUPPORTED) return tf::errors::Unimplemented(what, ": ", roi3d_strerror(code));
  return tf::errors::Internal(what, ": ", roi3d_strerror(code), " cudaError=", roi3d_last_cuda_error());
}

}  // namespace

// ----------------------------------------------------------------------------------------------
// Op registry -- byte-identical signatures (CAR.so / GI.so / GB.so / NMS.so .rodata)
// ----------------------------------------------------------------------------------------------
#if ROI3D_TF_HAS(1)
REGISTER_OP("CropAndResize3D")
    .Input("image: T")
    .Input("boxes: float")
    .Input("box_index: int32")
    .Input("crop_size: int32")
    .Output("crops: float")
    .Attr("T: {uint8, uint16, int8, int16, int32, int64, half, float, double}")
    .Attr("method_name: {'trilinear', 'nearest'} = 'trilinear'")
    .Attr("extrapolation_value: float = 0")
    .SetShapeFn([](InferenceContext* c) {
      ShapeHandle input, boxes, box_ind;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 5, &input));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 2, &boxes));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(2), 1, &box_ind));
      DimensionHandle num_boxes, unused;
      TF_RETURN_IF_ERROR(c->Merge(c->Dim(boxes, 0), c->Dim(box_ind, 0), &num_boxes));
      TF_RETURN_IF_ERROR(c->WithValue(c->Dim(boxes, 1), 6, &unused));
      return SetOutputToSizedImage3D(c, num_boxes, 3, c->Dim(input, 4));
    });
#endif

#if ROI3D_TF_HAS(2)
REGISTER_OP("CropAndResize3DGradImage")
    .Input("grads: float")
    .Input("boxes: float")
    .Input("box_ind: int32")
    .Input("image_size: int32")
    .Output("output: T")
    .Attr("T: {float, half, double}")
    .Attr("method_name: {'trilinear', 'nearest'} = 'trilinear'")
    .SetShapeFn([](InferenceContext* c) {
      ShapeHandle out;
      TF_RETURN_IF_ERROR(c->MakeShapeFromShapeTensor(3, &out));
      TF_RETURN_IF_ERROR(c->WithRank(out, 5, &out));
      c->set_output(0, out);
      return absl::OkStatus();
    });
#endif

#if ROI3D_TF_HAS(3)
REGISTER_OP("CropAndResize3DGradBoxes")
    .Input("grads: float")
    .Input("image: T")
    .Input("boxes: float")
    .Input("box_ind: int32")
    .Output("output: float")
    .Attr("T: {uint8, uint16, int8, int16, int32, int64, half, float, double}")
    .Attr("method_name: {'trilinear'} = 'trilinear'")
    .SetShapeFn([](InferenceContext* c) {
      c->set_output(0, c->input(2));
      return absl::OkStatus();
    });
#endif

#if ROI3D_TF_HAS(4)
REGISTER_OP("NonMaxSuppression3D")
    .Input("boxes: float")
    .Input("scores: float")
    .Input("max_output_size: int32")
    .Output("selected_indices: int32")
    .Attr("iou_threshold: float = 0.5")
    .SetShapeFn([](InferenceContext* c) {
      ShapeHandle boxes, scores, max_output_size;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 2, &boxes));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 1, &scores));
      TF_RETURN_IF_ERROR(c->WithRank(c->input(2), 0, &max_output_size));
      DimensionHandle unused;
      TF_RETURN_IF_ERROR(c->Merge(c->Dim(boxes, 0), c->Dim(scores, 0), &unused));
      TF_RETURN_IF_ERROR(c->WithValue(c->Dim(boxes, 1), 6, &unused));
      c->set_output(0, c->Vector(c->UnknownDim()));
      return absl::OkStatus();
    });
#endif

// ----------------------------------------------------------------------------------------------
// GPU kernels: validate like the reference (same messages), then call the C ABI
// ----------------------------------------------------------------------------------------------
class CropAndResize3DGpuOp : public tf::OpKernel {
 public:
  explicit CropAndResize3DGpuOp(tf::OpKernelConstruction* ctx) : tf::OpKernel(ctx) {
    std::string method;
    OP_REQUIRES_OK(ctx, ctx->GetAttr("method_name", &method));
    OP_REQUIRES(ctx, method == "trilinear" || method == "nearest",
                tf::errors::InvalidArgument("method must be 'trilinear' or 'nearest'"));
    method_ = MethodFromName(method);
    OP_REQUIRES_OK(ctx, ctx->GetAttr("extrapolation_value", &extrapolation_value_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& image = ctx->input(0);
    const tf::Tensor& boxes = ctx->input(1);
    const tf::Tensor& box_index = ctx->input(2);
    const tf::Tensor& crop_size = ctx->input(3);          // HostMemory
    OP_REQUIRES(ctx, image.dims() == 5, tf::errors::InvalidArgument("input image must be 5-D", image.shape().DebugString()));
    OP_REQUIRES(ctx, boxes.dims() == 2, tf::errors::InvalidArgument("boxes must be 2-D", boxes.shape().DebugString()));
    OP_REQUIRES(ctx, boxes.dim_size(1) == 6, tf::errors::InvalidArgument("boxes must have 6 columns"));
    OP_REQUIRES(ctx, box_index.dims() == 1, tf::errors::InvalidArgument("box_index must be 1-D", box_index.shape().DebugString()));
    OP_REQUIRES(ctx, box_index.dim_size(0) == boxes.dim_size(0), tf::errors::InvalidArgument("box_index has incompatible shape"));
    OP_REQUIRES(ctx, crop_size.dims() == 1, tf::errors::InvalidArgument("crop_size must be 1-D", crop_size.shape().DebugString()));
    OP_REQUIRES(ctx, crop_size.dim_size(0) == 3, tf::errors::InvalidArgument("crop_size must have three elements", crop_size.shape().DebugString()));
    const int B = image.dim_size(0), H = image.dim_size(1), W = image.dim_size(2), D = image.dim_size(3), C = image.dim_size(4);
    OP_REQUIRES(ctx, H > 0 && W > 0 && D > 0, tf::errors::InvalidArgument("image dimensions must be positive"));
    auto cs = crop_size.vec<int32_t>();
    const int ph = cs(0), pw = cs(1), pd = cs(2);
    OP_REQUIRES(ctx, ph > 0 && pw > 0 && pd > 0, tf::errors::InvalidArgument("crop dimensions must be positive"));
    const int n = boxes.dim_size(0);
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({n, ph, pw, pd, C}), &out));
    if (n == 0) return;
    // scratch for the ROI processing order (include/roi3d.h: the *_ws entry points); a temp of the op, like NMS's workspace
    tf::Tensor ws;
    const size_t ws_bytes = roi3d_car3d_workspace_bytes(n);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_INT32, tf::TensorShape({static_cast<int64_t>(ws_bytes / 4)}), &ws));
    const int rc = roi3d_car3d_fwd_ws(image.flat<float>().data(), B, H, W, D, C, boxes.flat<float>().data(),
                                      box_index.flat<int32_t>().data(), n, ph, pw, pd, method_, extrapolation_value_,
                                      out->flat<float>().data(), ws.flat<int32_t>().data(), ws_bytes, tf::GetGpuStream(ctx));
    OP_REQUIRES_OK(ctx, Roi3dStatus(rc, "CropAndResize3D"));
  }
 private:
  int method_;
  float extrapolation_value_;
};

class CropAndResize3DGradImageGpuOp : public tf::OpKernel {
 public:
  explicit CropAndResize3DGradImageGpuOp(tf::OpKernelConstruction* ctx) : tf::OpKernel(ctx) {
    std::string method;
    OP_REQUIRES_OK(ctx, ctx->GetAttr("method_name", &method));
    OP_REQUIRES(ctx, method == "trilinear" || method == "nearest",
                tf::errors::InvalidArgument("method must be 'trilinear' or 'nearest'"));
    method_ = MethodFromName(method);
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& grads = ctx->input(0);
    const tf::Tensor& boxes = ctx->input(1);
    const tf::Tensor& box_index = ctx->input(2);
    const tf::Tensor& image_size = ctx->input(3);         // HostMemory
    OP_REQUIRES(ctx, grads.dims() == 5, tf::errors::InvalidArgument("grads image must be 5-D", grads.shape().DebugString()));
    OP_REQUIRES(ctx, boxes.dims() == 2, tf::errors::InvalidArgument("boxes must be 2-D", boxes.shape().DebugString()));
    OP_REQUIRES(ctx, boxes.dim_size(1) == 6, tf::errors::InvalidArgument("boxes must have 6 columns"));
    OP_REQUIRES(ctx, box_index.dims() == 1, tf::errors::InvalidArgument("box_index must be 1-D", box_index.shape().DebugString()));
    OP_REQUIRES(ctx, box_index.dim_size(0) == boxes.dim_size(0), tf::errors::InvalidArgument("box_index has incompatible shape"));
    OP_REQUIRES(ctx, image_size.dims() == 1, tf::errors::InvalidArgument("image_size must be 1-D", image_size.shape().DebugString()));
    OP_REQUIRES(ctx, image_size.dim_size(0) == 5, tf::errors::InvalidArgument("image_size must have five elements", image_size.shape().DebugString()));
    const int n = grads.dim_size(0), ph = grads.dim_size(1), pw = grads.dim_size(2), pd = grads.dim_size(3);
    if (n > 0) OP_REQUIRES(ctx, ph > 0 && pw > 0 && pd > 0, tf::errors::InvalidArgument("grads dimensions must be positive"));
    // n comes from grads: a shorter boxes tensor would be read out of bounds on the device
    OP_REQUIRES(ctx, boxes.dim_size(0) == n, tf::errors::InvalidArgument("boxes and grads have incompatible shape"));
    auto sz = image_size.vec<int32_t>();
    const int B = sz(0), H = sz(1), W = sz(2), D = sz(3), C = sz(4);
    OP_REQUIRES(ctx, H > 0 && W > 0 && D > 0, tf::errors::InvalidArgument("image dimensions must be positive"));
    OP_REQUIRES(ctx, grads.dim_size(4) == C, tf::errors::InvalidArgument("image_size and grads are incompatible"));
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({B, H, W, D, C}), &out));
    tf::Tensor ws;
    const size_t ws_bytes = roi3d_car3d_workspace_bytes(n);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_INT32, tf::TensorShape({static_cast<int64_t>(ws_bytes / 4)}), &ws));
    const int rc = roi3d_car3d_grad_image_ws(grads.flat<float>().data(), boxes.flat<float>().data(),
                                             box_index.flat<int32_t>().data(), n, ph, pw, pd, B, H, W, D, C, method_,
                                             out->flat<float>().data(), ws.flat<int32_t>().data(), ws_bytes, tf::GetGpuStream(ctx));
    OP_REQUIRES_OK(ctx, Roi3dStatus(rc, "CropAndResize3DGradImage"));
  }
 private:
  int method_;
};

class CropAndResize3DGradBoxesGpuOp : public tf::OpKernel {
 public:
  explicit CropAndResize3DGradBoxesGpuOp(tf::OpKernelConstruction* ctx) : tf::OpKernel(ctx) {
    std::string method;
    OP_REQUIRES_OK(ctx, ctx->GetAttr("method_name", &method));
    OP_REQUIRES(ctx, method == "trilinear", tf::errors::InvalidArgument("method must be 'trilinear' or 'nearest'"));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& grads = ctx->input(0);
    const tf::Tensor& image = ctx->input(1);
    const tf::Tensor& boxes = ctx->input(2);
    const tf::Tensor& box_index = ctx->input(3);
    OP_REQUIRES(ctx, grads.dims() == 5, tf::errors::InvalidArgument("grads image must be 5-D", grads.shape().DebugString()));
    OP_REQUIRES(ctx, image.dims() == 5, tf::errors::InvalidArgument("input image must be 5-D", image.shape().DebugString()));
    OP_REQUIRES(ctx, boxes.dims() == 2, tf::errors::InvalidArgument("boxes must be 2-D", boxes.shape().DebugString()));
    OP_REQUIRES(ctx, boxes.dim_size(1) == 6, tf::errors::InvalidArgument("boxes must have 6 columns"));
    OP_REQUIRES(ctx, box_index.dims() == 1, tf::errors::InvalidArgument("box_index must be 1-D", box_index.shape().DebugString()));
    const int n = grads.dim_size(0), ph = grads.dim_size(1), pw = grads.dim_size(2), pd = grads.dim_size(3);
    if (n > 0) OP_REQUIRES(ctx, ph > 0 && pw > 0 && pd > 0, tf::errors::InvalidArgument("grads dimensions must be positive"));
    const int B = image.dim_size(0), H = image.dim_size(1), W = image.dim_size(2), D = image.dim_size(3), C = image.dim_size(4);
    OP_REQUIRES(ctx, H > 0 && W > 0 && D > 0, tf::errors::InvalidArgument("image dimensions must be positive"));
    OP_REQUIRES(ctx, grads.dim_size(4) == C, tf::errors::InvalidArgument("image and grads depths are incompatible"));
    OP_REQUIRES(ctx, box_index.dim_size(0) == boxes.dim_size(0), tf::errors::InvalidArgument("box_index has incompatible shape"));
    OP_REQUIRES(ctx, boxes.dim_size(0) == n, tf::errors::InvalidArgument("boxes and grads have incompatible shape"));
    tf::Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({n, 6}), &out));
    if (n == 0) return;
    const int rc = roi3d_car3d_grad_boxes(grads.flat<float>().data(), image.flat<float>().data(), B, H, W, D, C,
                                          boxes.flat<float>().data(), box_index.flat<int32_t>().data(), n, ph, pw, pd,
                                          out->flat<float>().data(), tf::GetGpuStream(ctx));
    OP_REQUIRES_OK(ctx, Roi3dStatus(rc, "CropAndResize3DGradBoxes"));
  }
};

class NonMaxSuppression3DGpuOp : public tf::OpKernel {
 public:
  explicit NonMaxSuppression3DGpuOp(tf::OpKernelConstruction* ctx) : tf::OpKernel(ctx) {
    OP_REQUIRES_OK(ctx, ctx->GetAttr("iou_threshold", &iou_threshold_));
  }
  void Compute(tf::OpKernelContext* ctx) override {
    const tf::Tensor& boxes = ctx->input(0);
    const tf::Tensor& scores = ctx->input(1);
    const tf::Tensor& max_output_size = ctx->input(2);    // HostMemory
    OP_REQUIRES(ctx, boxes.dims() == 2, tf::errors::InvalidArgument("boxes must be 2-D", boxes.shape().DebugString()));
    OP_REQUIRES(ctx, boxes.dim_size(1) == 6, tf::errors::InvalidArgument("boxes must have 6 columns"));
    OP_REQUIRES(ctx, scores.dims() == 1, tf::errors::InvalidArgument("scores must be 1-D", scores.shape().DebugString()));
    OP_REQUIRES(ctx, scores.dim_size(0) == boxes.dim_size(0), tf::errors::InvalidArgument("scores has incompatible shape"));
    OP_REQUIRES(ctx, tf::TensorShapeUtils::IsScalar(max_output_size.shape()),
                tf::errors::InvalidArgument("max_output_size must be 0-D, got shape ", max_output_size.shape().DebugString()));
    OP_REQUIRES(ctx, iou_threshold_ >= 0 && iou_threshold_ <= 1, tf::errors::InvalidArgument("iou_threshold must be in [0, 1]"));
    const int n = boxes.dim_size(0);
    const int max_out = std::max(0, max_output_size.scalar<int32_t>()());
    tf::Tensor* out = nullptr;
    if (n == 0 || max_out == 0) {
      OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({0}), &out));
      return;
    }
    // scratch: workspace + the selected indices at full capacity + the device-written count (pinned host)
    const size_t ws_bytes = roi3d_nms3d_workspace_bytes(n);
    tf::Tensor ws, keep, count;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_INT8, tf::TensorShape({static_cast<int64_t>(ws_bytes + 256)}), &ws));
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_INT32, tf::TensorShape({max_out}), &keep));
    tf::AllocatorAttributes pinned;
    pinned.set_on_host(true);
    pinned.set_gpu_compatible(true);
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(tf::DT_INT32, tf::TensorShape({1}), &count, pinned));
    auto stream = tf::GetGpuStream(ctx);
    char* ws_ptr = reinterpret_cast<char*>(ws.flat<int8_t>().data());
    ws_ptr += (256 - reinterpret_cast<uintptr_t>(ws_ptr) % 256) % 256;
    const int rc = roi3d_nms3d(boxes.flat<float>().data(), scores.flat<float>().data(), n, max_out, iou_threshold_,
                               keep.flat<int32_t>().data(), count.flat<int32_t>().data(), ws_ptr, ws_bytes, stream);
    OP_REQUIRES_OK(ctx, Roi3dStatus(rc, "NonMaxSuppression3D"));
    OP_REQUIRES(ctx, cudaStreamSynchronize(stream) == cudaSuccess, tf::errors::Internal("NonMaxSuppression3D: stream sync failed"));
    const int m = count.flat<int32_t>()(0);
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, tf::TensorShape({m}), &out));
    if (m > 0)
      OP_REQUIRES(ctx, cudaMemcpyAsync(out->flat<int32_t>().data(), keep.flat<int32_t>().data(), sizeof(int) * m,
                                       cudaMemcpyDeviceToDevice, stream) == cudaSuccess,
                  tf::errors::Internal("NonMaxSuppression3D: copy failed"));
  }
 private:
  float iou_threshold_;
};

#if ROI3D_TF_HAS(1)
REGISTER_KERNEL_BUILDER(Name("CropAndResize3D").Device(tf::DEVICE_GPU).TypeConstraint<float>("T").HostMemory("crop_size"),
                        CropAndResize3DGpuOp);
#endif
#if ROI3D_TF_HAS(2)
REGISTER_KERNEL_BUILDER(Name("CropAndResize3DGradImage").Device(tf::DEVICE_GPU).TypeConstraint<float>("T").HostMemory("image_size"),
                        CropAndResize3DGradImageGpuOp);
#endif
#if ROI3D_TF_HAS(3)
REGISTER_KERNEL_BUILDER(Name("CropAndResize3DGradBoxes").Device(tf::DEVICE_GPU).TypeConstraint<float>("T"),
                        CropAndResize3DGradBoxesGpuOp);
#endif
#if ROI3D_TF_HAS(4)
REGISTER_KERNEL_BUILDER(Name("NonMaxSuppression3D").Device(tf::DEVICE_GPU).HostMemory("max_output_size"),
                        NonMaxSuppression3DGpuOp);
#endif
