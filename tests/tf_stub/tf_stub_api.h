// tf_stub_api.h -- a MINIMAL stand-in for the TensorFlow C++ op API, only for `g++ -fsyntax-only` of
// 3d-mask-r-cnn_b200/tf_ops/roi3d_tf_ops.cc in an image without TensorFlow (tests/test_tf_ops_source.py).
// It declares exactly the names the shim uses, with the signatures of current TensorFlow (>= 2.15: absl::Status,
// int64_t, no tensorflow::OkStatus), so that a use of a removed alias, a wrong argument count to the C ABI of
// include/roi3d.h or a typo fails the CPU test suite.  It is NOT TensorFlow and proves nothing about linking.
#pragma once
#include <cstdint>
#include <functional>
#include <initializer_list>
#include <sstream>
#include <string>
#include <vector>
#include <cuda_runtime_api.h>

namespace absl {
class Status {
 public:
  Status() = default;
  Status(int code, std::string msg) : code_(code), msg_(std::move(msg)) {}
  bool ok() const { return code_ == 0; }
 private:
  int code_ = 0;
  std::string msg_;
};
inline Status OkStatus() { return Status(); }
}  // namespace absl

namespace tensorflow {
using Status = absl::Status;
using int32 = std::int32_t;      // still provided by tensorflow/core/platform/types.h
using int8 = std::int8_t;
enum DataType { DT_FLOAT = 1, DT_INT32 = 3, DT_INT8 = 6 };
inline std::string DataTypeString(DataType) { return std::string(); }
extern const char* const DEVICE_GPU;

namespace errors {
namespace internal { template <typename... A> std::string cat(const A&... a) { std::ostringstream s; (void)std::initializer_list<int>{((s << a), 0)...}; return s.str(); } }
template <typename... A> Status InvalidArgument(const A&... a) { return Status(3, internal::cat(a...)); }
template <typename... A> Status Unimplemented(const A&... a) { return Status(12, internal::cat(a...)); }
template <typename... A> Status Internal(const A&... a) { return Status(13, internal::cat(a...)); }
}  // namespace errors

class TensorShape {
 public:
  TensorShape() = default;
  TensorShape(std::initializer_list<int64_t> d) : d_(d) {}
  std::string DebugString() const { return std::string(); }
  int dims() const { return (int)d_.size(); }
 private:
  std::vector<int64_t> d_;
};
struct TensorShapeUtils { static bool IsScalar(const TensorShape& s) { return s.dims() == 0; } };

template <typename T> struct FlatView { T* data() const { return nullptr; } T& operator()(int64_t) const { return *data(); } };
template <typename T> struct ScalarView { T& operator()() const { static T v{}; return v; } };
class Tensor {
 public:
  int dims() const { return 0; }
  int64_t dim_size(int) const { return 0; }
  const TensorShape& shape() const { static TensorShape s; return s; }
  DataType dtype() const { return DT_FLOAT; }
  template <typename T> FlatView<const T> flat() const { return {}; }
  template <typename T> FlatView<T> flat() { return {}; }
  template <typename T> FlatView<const T> vec() const { return {}; }
  template <typename T> ScalarView<const T> scalar() const { return {}; }
};

struct AllocatorAttributes { void set_on_host(bool) {} void set_gpu_compatible(bool) {} };

class OpKernelConstruction {
 public:
  template <typename T> Status GetAttr(const char*, T*) const { return absl::OkStatus(); }
  void CtxFailure(const char*, int, const Status&) {}
  void CtxFailureWithWarning(const char*, int, const Status&) {}
};
class OpKernelContext {
 public:
  const Tensor& input(int) { static Tensor t; return t; }
  Status allocate_output(int, const TensorShape&, Tensor**) { return absl::OkStatus(); }
  Status allocate_temp(DataType, const TensorShape&, Tensor*) { return absl::OkStatus(); }
  Status allocate_temp(DataType, const TensorShape&, Tensor*, AllocatorAttributes) { return absl::OkStatus(); }
  void CtxFailure(const char*, int, const Status&) {}
  void CtxFailureWithWarning(const char*, int, const Status&) {}
};
class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() = default;
  virtual void Compute(OpKernelContext*) = 0;
};
cudaStream_t GetGpuStream(OpKernelContext*);

#define OP_REQUIRES(CTX, EXP, STATUS) do { if (!(EXP)) { (CTX)->CtxFailure(__FILE__, __LINE__, (STATUS)); return; } } while (0)
#define OP_REQUIRES_OK(CTX, ...) do { const ::tensorflow::Status _s(__VA_ARGS__); if (!_s.ok()) { (CTX)->CtxFailureWithWarning(__FILE__, __LINE__, _s); return; } } while (0)
#define TF_RETURN_IF_ERROR(...) do { const ::tensorflow::Status _s = (__VA_ARGS__); if (!_s.ok()) return _s; } while (0)

namespace shape_inference {
struct DimensionHandle {};
struct ShapeHandle {};
class InferenceContext {
 public:
  ShapeHandle input(int) { return {}; }
  const Tensor* input_tensor(int) { return nullptr; }
  Status WithRank(ShapeHandle, int64_t, ShapeHandle*) { return absl::OkStatus(); }
  Status WithValue(DimensionHandle, int64_t, DimensionHandle*) { return absl::OkStatus(); }
  Status Merge(DimensionHandle, DimensionHandle, DimensionHandle*) { return absl::OkStatus(); }
  Status MakeShapeFromShapeTensor(int, ShapeHandle*) { return absl::OkStatus(); }
  DimensionHandle Dim(ShapeHandle, int64_t) { return {}; }
  DimensionHandle UnknownDim() { return {}; }
  DimensionHandle MakeDim(int64_t) { return {}; }
  ShapeHandle MakeShape(std::initializer_list<DimensionHandle>) { return {}; }
  ShapeHandle Vector(DimensionHandle) { return {}; }
  void set_output(int, ShapeHandle) {}
};
}  // namespace shape_inference

namespace register_op {
struct OpDefBuilderWrapper {
  explicit OpDefBuilderWrapper(const char*) {}
  OpDefBuilderWrapper& Input(const char*) { return *this; }
  OpDefBuilderWrapper& Output(const char*) { return *this; }
  OpDefBuilderWrapper& Attr(const char*) { return *this; }
  OpDefBuilderWrapper& SetShapeFn(std::function<Status(shape_inference::InferenceContext*)>) { return *this; }
};
}  // namespace register_op
struct KernelDefBuilder {
  explicit KernelDefBuilder(const char*) {}
  KernelDefBuilder& Device(const char*) { return *this; }
  template <typename T> KernelDefBuilder& TypeConstraint(const char*) { return *this; }
  KernelDefBuilder& HostMemory(const char*) { return *this; }
};
namespace register_kernel {              // the real macro prefixes its first argument with this namespace
struct Name : KernelDefBuilder { explicit Name(const char* n) : KernelDefBuilder(n) {} };
}  // namespace register_kernel
#define TF_STUB_CAT2(a, b) a##b
#define TF_STUB_CAT(a, b) TF_STUB_CAT2(a, b)
#define REGISTER_OP(name) static ::tensorflow::register_op::OpDefBuilderWrapper TF_STUB_CAT(tf_stub_op_, __COUNTER__) = ::tensorflow::register_op::OpDefBuilderWrapper(name)
#define REGISTER_KERNEL_BUILDER(builder, ...) static ::tensorflow::KernelDefBuilder TF_STUB_CAT(tf_stub_kb_, __COUNTER__) = ::tensorflow::register_kernel::builder; \
  static ::tensorflow::OpKernel* TF_STUB_CAT(tf_stub_mk_, __COUNTER__)(::tensorflow::OpKernelConstruction* c) { return new __VA_ARGS__(c); }
}  // namespace tensorflow
