// roi3d_common.cuh -- shared device helpers for the ROI hot-path kernels (sm_100a).
//
// Arithmetic contract: the reference's kernels are scalar SSE code with one
// rounding per fp32 operation (no FMA).  Every helper here that feeds a
// bit-exact comparison (sample coordinates, IoU, the forward lerps) spells the
// operations with __f*_rn intrinsics so nvcc can neither contract them into
// FMAs nor reassociate them.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/roi3d.h"

namespace roi3d {

// ---- host-side bookkeeping ------------------------------------------------
extern thread_local int g_last_cuda_error;
extern thread_local long long g_launches;
int option_value(int which);
enum { OPT_CAR_FWD_VARIANT = 0, OPT_CAR_BWD_VARIANT = 1, OPT_NMS_VARIANT = 2, OPT_CAR_V = 3, OPT_KSPLIT = 4, OPT_NMS_SORT = 5, OPT_PDL = 6,
       OPT_OS_TZ = 7, OPT_OS_STAGES = 8, OPT_OS_STAGE_KIB = 9, OPT_OS_DEBUG = 10, OPT_BWD_SPLIT = 11, OPT_SEP_RC = 12, OPT_SEP_NS = 13, OPT_FILL_CTAS = 14, OPT_BWD_STAGE_KIB = 15, OPT_EXPERIMENT = 16,
       OPT_COUNT };
// cudaFuncAttributeMaxDynamicSharedMemorySize, issued once per (kernel, device, size) instead of on every call
cudaError_t ensure_dyn_smem(const void *kernel, size_t bytes);
// multiprocessor count of the current device (cached per device)
int num_sms();

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return ROI3D_ECUDA;
}
#define ROI3D_CUDA_TRY(expr)                                        \
    do {                                                            \
        cudaError_t e__ = (expr);                                   \
        if (e__ != cudaSuccess) return ::roi3d::cuda_fail(e__);     \
    } while (0)
#define ROI3D_LAUNCH_CHECK()                                        \
    do {                                                            \
        ++::roi3d::g_launches;                                      \
        cudaError_t e__ = cudaGetLastError();                       \
        if (e__ != cudaSuccess) return ::roi3d::cuda_fail(e__);     \
    } while (0)


// ---- sample coordinates (CAR.so@0x499a-0x4ae5, 0x533a) ---------------------
// scale = ((a2 - a1) * f(dim-1)) / f(p-1)            (p > 1), else 0
// in(k) = a1 * f(dim-1) + f(k) * scale               (p > 1)
//       = float(double(a1 + a2) * 0.5 * double(dim-1))   (p == 1)
__device__ __forceinline__ float axis_scale(float a1, float a2, int dim, int p) {
    return (p > 1) ? __fdiv_rn(__fmul_rn(__fsub_rn(a2, a1), (float)(dim - 1)), (float)(p - 1)) : 0.0f;
}
__device__ __forceinline__ float axis_coord(float a1, float a2, int dim, int p, int k, float scale) {
    if (p > 1) return __fadd_rn(__fmul_rn(a1, (float)(dim - 1)), __fmul_rn((float)k, scale));
    return (float)__dmul_rn(__dmul_rn((double)__fadd_rn(a1, a2), 0.5), (double)(dim - 1));
}
__device__ __forceinline__ bool axis_invalid(float in, int dim) {
    return in < 0.0f || in > (float)(dim - 1);
}

// a + (b - a) * t with three roundings, as the reference's subss/mulss/addss
__device__ __forceinline__ float lerp_rn(float a, float b, float t) {
    return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));
}
__device__ __forceinline__ float4 lerp_rn(const float4 a, const float4 b, float t) {
    return make_float4(lerp_rn(a.x, b.x, t), lerp_rn(a.y, b.y, t), lerp_rn(a.z, b.z, t), lerp_rn(a.w, b.w, t));
}

// ---- memory helpers ---------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float *p) {
    return __ldg(reinterpret_cast<const float4 *>(p));
}
// streaming (evict-first) 128-bit store: crops / grads are written once and not re-read here
__device__ __forceinline__ void st_stream4(float *p, const float4 v) {
    __stcs(reinterpret_cast<float4 *>(p), v);
}
// float16 flavour: round to nearest even (ndarray.astype(float16) for finite inputs), one 8-byte streaming store
__device__ __forceinline__ void st_stream4(__half *p, const float4 v) {
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" :: "l"(p), "r"(*reinterpret_cast<const unsigned *>(&lo)),
                 "r"(*reinterpret_cast<const unsigned *>(&hi)) : "memory");
}
// vectorised fire-and-forget global reduction (sm_90+): one 16-byte RED per 4 channels
__device__ __forceinline__ void red_add4(float *p, const float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add1(float *p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

// pull one 128-byte line towards L2 ahead of its use (no register, no scoreboard entry)
__device__ __forceinline__ void prefetch_l2(const void *p) {
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier --------------------------------
__device__ __forceinline__ void mbar_init(unsigned mbar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst_smem, const void *src, unsigned bytes, unsigned mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// be scheduled while its predecessor still runs; it must execute pdl_wait() before touching anything the predecessor
// wrote (the wait returns when the predecessor grid has completed and flushed).  pdl_trigger() in the predecessor lets
// the successor's launch latency overlap with the predecessor's execution.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                    Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- shared memory through 32-bit addresses (no generic->shared window math in the loops) ----
__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128u(unsigned a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64u(unsigned a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds32u(unsigned a) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, const float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- geometry + kernel launchers (defined in roi3d_car_*.cu / roi3d_nms.cu) ------
struct CarGeom {
    int B, H, W, D, C;      // volume [B,H,W,D,C]
    int n, ph, pw, pd;      // n boxes, crop size
};

int launch_car3d_fwd_direct(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                            int method, float ext, float *crops, cudaStream_t stream);
int launch_car3d_grad_image_direct(const float *grads, const float *boxes, const int *box_ind, const CarGeom &g,
                                   int method, float *grad_image, cudaStream_t stream);
int launch_car3d_grad_boxes(const float *grads, const float *image, const float *boxes, const int *box_ind,
                            const CarGeom &g, float *grad_boxes, cudaStream_t stream);
int launch_car3d_fwd_plane_tma(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                               float ext, float *crops, cudaStream_t stream);
int launch_car3d_fwd_plane_g4(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                              float ext, float *crops, cudaStream_t stream);
int launch_car3d_fwd_plane(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                           float ext, float *crops, cudaStream_t stream, const int *perm = nullptr);
int launch_car3d_order(const float *boxes, const int *box_index, int n, int rois_per_image, int *perm, cudaStream_t stream);
int launch_car3d_grad_image_plane(const float *grads, const float *boxes, const int *box_ind, const CarGeom &g,
                                  float *grad_image, cudaStream_t stream, bool zero_fill = false, bool tma = false,
                                  const int *perm = nullptr);
struct PyrParams;
int launch_car3d_fwd_sep(const float *image, const float *boxes, const int *box_index, const CarGeom &g,
                         float ext, void *crops, const PyrParams *pyr, bool half_out, cudaStream_t stream);
bool car3d_grad_image_os_supported(const CarGeom &g);
long long car3d_grad_image_os_ctas(const CarGeom &g);
int launch_car3d_grad_image_os(const float *grads, const float *boxes, const int *box_ind, const CarGeom &g,
                               float *grad_image, cudaStream_t stream);
int launch_pyramid_fwd(const float *const images[4], const int H[4], const int W[4], const int D[4], int B, int C,
                       const float *boxes, int rois_per_image, float imH, float imW, float imD,
                       int ph, int pw, int pd, void *crops, bool half_out, cudaStream_t stream, const int *perm = nullptr);
int launch_pyramid_grad(const float *grads, float *const grad_images[4], const int H[4], const int W[4], const int D[4],
                        int B, int C, const float *boxes, int rois_per_image, float imH, float imW, float imD,
                        int ph, int pw, int pd, cudaStream_t stream, const int *perm = nullptr);
int launch_overlaps3d(const float *boxes1, int n, const float *boxes2, int m, float *out, cudaStream_t stream);
int launch_decode_proposals(const float *anchors, const float *deltas, const int *index, int n, const float std_dev[6],
                            float image_depth, float *boxes, cudaStream_t stream);
size_t topk_workspace_bytes(int n);
int launch_topk(const float *scores, int n, int k, int *idx_out, float *scores_out, void *ws, size_t ws_bytes,
                cudaStream_t stream);
int launch_gather_pad_boxes(const float *boxes, const int *keep, const int *count, int p, float *out, cudaStream_t stream);
size_t nms3d_workspace_bytes(int n, int segments);
size_t refine_detections_workspace_bytes(int images, int rois, int max_inst);
int launch_refine_detections(const float *rois, const float *probs, const float *deltas, int images, int rois_per_image,
                             int num_classes, const float image_shape[3], const float std_dev[6], float min_conf,
                             float nms_thr, int nms_mode, int max_inst, float *detections, int *det_count, void *ws,
                             size_t ws_bytes, cudaStream_t stream);
int launch_mask_targets(const void *masks, int mask_dtype, int H, int W, int D, const float *boxes, const int *assignment,
                        int n, int mh, int mw, int md, float *targets, unsigned char *bits, cudaStream_t stream);
int launch_f32_to_f16(const float *x, long long n, void *y, cudaStream_t stream);
int launch_f16_to_f32(const void *x, long long n, float *y, cudaStream_t stream);
int launch_pack_bits(const float *x, long long n, unsigned char *bits, cudaStream_t stream);
int launch_unpack_bits(const unsigned char *bits, long long n, float *y, cudaStream_t stream);
int launch_nms3d(const float *boxes, const float *scores, const int *seg_offsets, int segments, int n, int max_out,
                 float thr, int *keep_idx, int *keep_count, void *ws, size_t ws_bytes, cudaStream_t stream);

}  // namespace roi3d
