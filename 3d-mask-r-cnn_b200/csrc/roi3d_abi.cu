// roi3d_abi.cu -- the C ABI of libroi3d_b200.so (see include/roi3d.h): argument
// validation, variant selection and kernel launches.  No torch / TF types, no
// allocation, no device or stream synchronisation, no CPU fallback.
#include "roi3d_common.cuh"
#include <map>
#include <mutex>
#include <string.h>
#include <utility>

namespace roi3d {
thread_local int g_last_cuda_error = 0;
thread_local long long g_launches = 0;
// Tuning knobs are per calling thread: nothing another thread (a second TF inter-op worker, a second torch stream
// thread) does can change the kernels this thread's calls select.
static thread_local int g_options[OPT_COUNT];
int option_value(int which) { return g_options[which]; }

// The only process-wide state: two caches of facts about the device (the opt-in shared-memory size already granted
// to a kernel, the SM count).  Both are idempotent -- a lost race repeats a query, never changes a result.
cudaError_t ensure_dyn_smem(const void *kernel, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> granted;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t &have = granted[std::make_pair(kernel, dev)];
    if (bytes <= have) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

int num_sms()
{
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cached[dev];
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev] = v;
    }
    return v;
}

static int option_index(const char *name) {
    if (!name) return -1;
    if (!strcmp(name, "car_fwd_variant")) return OPT_CAR_FWD_VARIANT;
    if (!strcmp(name, "car_bwd_variant")) return OPT_CAR_BWD_VARIANT;
    if (!strcmp(name, "nms_variant")) return OPT_NMS_VARIANT;
    if (!strcmp(name, "car_lanes_v")) return OPT_CAR_V;
    if (!strcmp(name, "car_ctas_per_sm_target")) return OPT_KSPLIT;
    if (!strcmp(name, "nms_sort_variant")) return OPT_NMS_SORT;
    if (!strcmp(name, "pdl") || !strcmp(name, "nms_pdl")) return OPT_PDL;     // "nms_pdl": round-1 name, kept as an alias
    if (!strcmp(name, "car_os_tile_depth")) return OPT_OS_TZ;
    if (!strcmp(name, "car_os_ring_stages")) return OPT_OS_STAGES;
    if (!strcmp(name, "car_os_stage_kib")) return OPT_OS_STAGE_KIB;
    if (!strcmp(name, "car_os_debug")) return OPT_OS_DEBUG;
    if (!strcmp(name, "car_bwd_image_split")) return OPT_BWD_SPLIT;
    if (!strcmp(name, "car_fill_ctas_per_sm")) return OPT_FILL_CTAS;
    if (!strcmp(name, "car_bwd_stage_kib")) return OPT_BWD_STAGE_KIB;
    if (!strcmp(name, "car_experiment")) return OPT_EXPERIMENT;         // bit mask of A/B switches used by profiles/*.py
    if (!strcmp(name, "car_sep_rows")) return OPT_SEP_RC;
    if (!strcmp(name, "car_sep_ring")) return OPT_SEP_NS;
    return -1;
}

static bool geom_ok(const CarGeom &g) {
    return g.B > 0 && g.H > 0 && g.W > 0 && g.D > 0 && g.C > 0 && g.n >= 0 && g.ph > 0 && g.pw > 0 && g.pd > 0;
}
// the kernels index one batch item with 32-bit element offsets and per-axis tables of <= 64 samples
static bool geom_supported(const CarGeom &g) {
    const long long per_image = (long long)g.H * g.W * g.D * g.C;
    return per_image < (1ll << 31);
}
}  // namespace roi3d

using namespace roi3d;

extern "C" {

const char *roi3d_version(void) { return "roi3d-b200 0.1.0 (sm_100a)"; }

const char *roi3d_strerror(int code) {
    switch (code) {
    case ROI3D_OK: return "ok";
    case ROI3D_EINVAL: return "invalid argument";
    case ROI3D_EWORKSPACE: return "workspace missing, too small or misaligned";
    case ROI3D_EUNSUPPORTED: return "shape not supported by the sm_100a kernels";
    case ROI3D_ECUDA: return "CUDA error (see roi3d_last_cuda_error)";
    default: return "unknown roi3d error";
    }
}

int roi3d_last_cuda_error(void) { return g_last_cuda_error; }
long long roi3d_kernel_launches(void) { return g_launches; }
void roi3d_reset_kernel_launches(void) { g_launches = 0; }

int roi3d_set_option(const char *name, int value) {
    const int i = option_index(name);
    if (i < 0) return ROI3D_EINVAL;
    g_options[i] = value;
    return ROI3D_OK;
}
int roi3d_get_option(const char *name, int *value) {
    const int i = option_index(name);
    if (i < 0 || !value) return ROI3D_EINVAL;
    *value = g_options[i];
    return ROI3D_OK;
}

size_t roi3d_nms3d_workspace_bytes(int n) { return nms3d_workspace_bytes(n, 1); }
size_t roi3d_nms3d_batched_workspace_bytes(int n_max, int segments) { return nms3d_workspace_bytes(n_max, segments); }

int roi3d_nms3d_batched(const float *boxes, const float *scores, const int *seg_offsets, int segments, int n_max,
                        int max_out, float iou_thr, int *keep_idx, int *keep_count, void *workspace,
                        size_t workspace_bytes, roi3d_stream_t stream)
{
    if (segments < 0 || n_max < 0 || max_out < 0 || !keep_count) return ROI3D_EINVAL;
    if (!(iou_thr >= 0.0f && iou_thr <= 1.0f)) return ROI3D_EINVAL;
    if (segments == 0) return ROI3D_OK;
    if (!seg_offsets) return ROI3D_EINVAL;
    if (n_max > 0 && max_out > 0 && (!boxes || !scores || !keep_idx)) return ROI3D_EINVAL;
    return launch_nms3d(boxes, scores, seg_offsets, segments, n_max, max_out, iou_thr, keep_idx, keep_count, workspace,
                        workspace_bytes, static_cast<cudaStream_t>(stream));
}

int roi3d_nms3d(const float *boxes, const float *scores, int n, int max_out, float iou_thr,
                int *keep_idx, int *keep_count, void *workspace, size_t workspace_bytes,
                roi3d_stream_t stream)
{
    if (n < 0 || max_out < 0 || !keep_count) return ROI3D_EINVAL;
    if (!(iou_thr >= 0.0f && iou_thr <= 1.0f)) return ROI3D_EINVAL;       // "iou_threshold must be in [0, 1]"
    if (n > 0 && max_out > 0 && (!boxes || !scores || !keep_idx)) return ROI3D_EINVAL;
    return launch_nms3d(boxes, scores, nullptr, 0, n, max_out, iou_thr, keep_idx, keep_count, workspace, workspace_bytes,
                        static_cast<cudaStream_t>(stream));
}

// ROI processing order (see car3d_order_kernel): worth its ~3 us kernel when a call moves enough bytes
static const int *order_rois(const float *boxes, const int *box_index, int n, int rois_per_image, int pool_voxels, int C,
                             void *workspace, size_t workspace_bytes, cudaStream_t s, int *rc)
{
    *rc = ROI3D_OK;
    if (!workspace || workspace_bytes < roi3d_car3d_workspace_bytes(n) || (reinterpret_cast<uintptr_t>(workspace) & 3)) return nullptr;
    if (option_value(OPT_EXPERIMENT) & 16) return nullptr;                 // A/B switch: ROIs in the order given
    if (n < 32 || (long long)n * pool_voxels * C < (1ll << 26)) return nullptr;   // < 256 MB of crops: not worth a launch
    int *perm = static_cast<int *>(workspace);
    *rc = launch_car3d_order(boxes, box_index, n, rois_per_image, perm, s);
    return *rc == ROI3D_OK ? perm : nullptr;
}

size_t roi3d_car3d_workspace_bytes(int n) { return n > 0 ? ((size_t)n * sizeof(int) + 255) & ~size_t(255) : 256; }

int roi3d_car3d_fwd(const float *image, int B, int H, int W, int D, int C,
                    const float *boxes, const int *box_index, int n,
                    int ph, int pw, int pd, int method, float extrapolation_value,
                    float *crops, roi3d_stream_t stream)
{
    return roi3d_car3d_fwd_ws(image, B, H, W, D, C, boxes, box_index, n, ph, pw, pd, method, extrapolation_value, crops,
                              nullptr, 0, stream);
}

int roi3d_car3d_fwd_ws(const float *image, int B, int H, int W, int D, int C,
                       const float *boxes, const int *box_index, int n,
                       int ph, int pw, int pd, int method, float extrapolation_value,
                       float *crops, void *workspace, size_t workspace_bytes, roi3d_stream_t stream)
{
    const CarGeom g{B, H, W, D, C, n, ph, pw, pd};
    if (!geom_ok(g)) return ROI3D_EINVAL;
    if (method != ROI3D_METHOD_TRILINEAR && method != ROI3D_METHOD_NEAREST) return ROI3D_EINVAL;
    if (n == 0) return ROI3D_OK;
    if (!image || !boxes || !box_index || !crops) return ROI3D_EINVAL;
    if (!geom_supported(g)) return ROI3D_EUNSUPPORTED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    int variant = option_value(OPT_CAR_FWD_VARIANT);
    const bool plane_ok = method == ROI3D_METHOD_TRILINEAR && C % 4 == 0 && ph <= 64 && pw <= 64 && pd <= 64 &&
                          ((reinterpret_cast<uintptr_t>(image) | reinterpret_cast<uintptr_t>(crops)) & 15) == 0;
    // measured on B200 (profiles/fwd_rule_sweep.py, variants_r1.txt): the plane-staged kernel wins once a depth slice
    // has enough outputs to amortise its per-CTA tables -- from 10 x 10 at any channel count >= 32, from 7 x 7 when
    // the channel chunks are full (C >= 128); below that the direct gather wins or ties
    if (variant == 0) variant = (plane_ok && ((C >= 32 && ph * pw >= 100) || (C >= 128 && ph * pw >= 49))) ? 2 : 1;
    if (variant == 5 && plane_ok) {                                     // plane kernel fed by TMA gather4 row copies
        const int rc = launch_car3d_fwd_plane_g4(image, boxes, box_index, g, extrapolation_value, crops, s);
        if (rc != ROI3D_EUNSUPPORTED) return rc;
        variant = 2;
    }
    if (variant == 4 && plane_ok) {                                     // row-walk separable kernel
        const int rc = launch_car3d_fwd_sep(image, boxes, box_index, g, extrapolation_value, crops, nullptr, false, s);
        if (rc != ROI3D_EUNSUPPORTED) return rc;
        variant = 2;
    }
    if (variant == 3 && plane_ok) {                                     // TMA-fed plane kernel (opt-in)
        const int rc = launch_car3d_fwd_plane_tma(image, boxes, box_index, g, extrapolation_value, crops, s);
        if (rc != ROI3D_EUNSUPPORTED) return rc;
        variant = 2;
    }
    if (variant == 2 && plane_ok) {
        int rc;
        const int *perm = order_rois(boxes, box_index, n, 0, ph * pw * pd, C, workspace, workspace_bytes, s, &rc);
        if (rc != ROI3D_OK) return rc;
        return launch_car3d_fwd_plane(image, boxes, box_index, g, extrapolation_value, crops, s, perm);
    }
    return launch_car3d_fwd_direct(image, boxes, box_index, g, method, extrapolation_value, crops, s);
}

int roi3d_car3d_grad_image(const float *grads, const float *boxes, const int *box_ind, int n,
                           int ph, int pw, int pd, int B, int H, int W, int D, int C, int method,
                           float *grad_image, roi3d_stream_t stream)
{
    return roi3d_car3d_grad_image_ws(grads, boxes, box_ind, n, ph, pw, pd, B, H, W, D, C, method, grad_image, nullptr, 0, stream);
}

int roi3d_car3d_grad_image_ws(const float *grads, const float *boxes, const int *box_ind, int n,
                              int ph, int pw, int pd, int B, int H, int W, int D, int C, int method,
                              float *grad_image, void *workspace, size_t workspace_bytes, roi3d_stream_t stream)
{
    const CarGeom g{B, H, W, D, C, n, ph, pw, pd};
    if (!geom_ok(g)) return ROI3D_EINVAL;
    if (method != ROI3D_METHOD_TRILINEAR && method != ROI3D_METHOD_NEAREST) return ROI3D_EINVAL;
    if (!grad_image) return ROI3D_EINVAL;
    if (!geom_supported(g)) return ROI3D_EUNSUPPORTED;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // the op's contract: the whole output is defined (GI.so@0x3ec5 zero-fills it)
    int variant = option_value(OPT_CAR_BWD_VARIANT);
    const bool aligned = ((reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(grad_image)) & 15) == 0;
    const bool plane_ok = method == ROI3D_METHOD_TRILINEAR && C % 4 == 0 && ph <= 64 && pw <= 64 && pd <= 64 && aligned;
    const bool os_ok = method == ROI3D_METHOD_TRILINEAR && aligned && car3d_grad_image_os_supported(g);
    const bool have_in = n > 0 && grads && boxes && box_ind;
    // auto: the plane-staged RED scatter with its grads slices staged by TMA tensor tile copies (variant 4; cfg2 14^3:
    // 0.318 vs 0.371 ms with per-thread loads, profiles/bwd_tma_sweep.py); variant 2 keeps the LDG staging.  The
    // output-stationary kernel (variant 3: every voxel stored once, no zero-fill, no atomics, deterministic; DRAM traffic
    // 1.03 vs 1.74 GB at cfg2 14^3) is opt-in: it is latency-bound on per-(box, tile) bookkeeping and measures
    // 0.55-0.64 ms (profiles/r2_os_grad_image_ncu.txt)
    if (variant == 0) variant = (plane_ok && C >= 32) ? 4 : 1;
    if (variant == 3 && os_ok && have_in) return launch_car3d_grad_image_os(grads, boxes, box_ind, g, grad_image, s);
    if (variant == 3) variant = (plane_ok && C >= 32) ? 2 : 1;
    bool tma = variant == 4 && plane_ok;
    if (variant == 4) variant = plane_ok ? 2 : 1;
    const bool plane = have_in && variant == 2 && plane_ok;
    // plane path: the zero-fill is a kernel that the scatter kernel overlaps with (programmatic dependent launch;
    // option "pdl" = 1 restores memset + plain stream order)
    const bool fused_fill = plane && option_value(OPT_PDL) == 0;
    if (!fused_fill) ROI3D_CUDA_TRY(cudaMemsetAsync(grad_image, 0, sizeof(float) * (size_t)B * H * W * D * C, s));
    if (n == 0) return ROI3D_OK;
    if (!grads || !boxes || !box_ind) return ROI3D_EINVAL;
    const int *perm = nullptr;
    if (plane) {                                         // the order kernel runs ahead of the zero-fill / scatter pair
        int rc;
        perm = order_rois(boxes, box_ind, n, 0, ph * pw * pd, C, workspace, workspace_bytes, s, &rc);
        if (rc != ROI3D_OK) return rc;
    }
    if (plane && tma) {
        const int rc = launch_car3d_grad_image_plane(grads, boxes, box_ind, g, grad_image, s, fused_fill, true, perm);
        if (rc != ROI3D_EUNSUPPORTED) return rc;        // (nothing has been launched when the TMA build declines a shape)
    }
    if (plane) return launch_car3d_grad_image_plane(grads, boxes, box_ind, g, grad_image, s, fused_fill, false, perm);
    return launch_car3d_grad_image_direct(grads, boxes, box_ind, g, method, grad_image, s);
}

int roi3d_car3d_grad_boxes(const float *grads, const float *image, int B, int H, int W, int D, int C,
                           const float *boxes, const int *box_ind, int n, int ph, int pw, int pd,
                           float *grad_boxes, roi3d_stream_t stream)
{
    const CarGeom g{B, H, W, D, C, n, ph, pw, pd};
    if (!geom_ok(g)) return ROI3D_EINVAL;
    if (n == 0) return ROI3D_OK;
    if (!grads || !image || !boxes || !box_ind || !grad_boxes) return ROI3D_EINVAL;
    if (!geom_supported(g)) return ROI3D_EUNSUPPORTED;
    return launch_car3d_grad_boxes(grads, image, boxes, box_ind, g, grad_boxes, static_cast<cudaStream_t>(stream));
}

int roi3d_overlaps3d(const float *boxes1, int n, const float *boxes2, int m, float *overlaps, roi3d_stream_t stream)
{
    if (n < 0 || m < 0) return ROI3D_EINVAL;
    if (n == 0 || m == 0) return ROI3D_OK;
    if (!boxes1 || !boxes2 || !overlaps) return ROI3D_EINVAL;
    return launch_overlaps3d(boxes1, n, boxes2, m, overlaps, static_cast<cudaStream_t>(stream));
}

int roi3d_decode_proposals(const float *anchors, const float *deltas, const int *index, int n, const float std_dev[6],
                           float image_depth, float *boxes, roi3d_stream_t stream)
{
    if (n < 0 || !std_dev) return ROI3D_EINVAL;
    if (n == 0) return ROI3D_OK;
    if (!anchors || !deltas || !boxes) return ROI3D_EINVAL;
    return launch_decode_proposals(anchors, deltas, index, n, std_dev, image_depth, boxes, static_cast<cudaStream_t>(stream));
}

size_t roi3d_topk_workspace_bytes(int n) { return topk_workspace_bytes(n > 0 ? n : 1); }

int roi3d_topk(const float *scores, int n, int k, int *idx_out, float *scores_out, void *workspace, size_t workspace_bytes,
               roi3d_stream_t stream)
{
    if (n < 0 || k < 0 || k > n) return ROI3D_EINVAL;
    if (k == 0) return ROI3D_OK;
    if (!scores || !idx_out) return ROI3D_EINVAL;
    return launch_topk(scores, n, k, idx_out, scores_out, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int roi3d_gather_pad_boxes(const float *boxes, const int *keep_idx, const int *keep_count, int proposal_count,
                           float *proposals, roi3d_stream_t stream)
{
    if (proposal_count < 0) return ROI3D_EINVAL;
    if (proposal_count == 0) return ROI3D_OK;
    if (!boxes || !keep_idx || !keep_count || !proposals) return ROI3D_EINVAL;
    return launch_gather_pad_boxes(boxes, keep_idx, keep_count, proposal_count, proposals, static_cast<cudaStream_t>(stream));
}

size_t roi3d_refine_detections_workspace_bytes(int images, int rois_per_image, int max_instances)
{
    if (images <= 0 || rois_per_image < 0 || max_instances < 0) return 256;
    return refine_detections_workspace_bytes(images, rois_per_image, max_instances);
}

int roi3d_refine_detections(const float *rois, const float *probs, const float *deltas, int images, int rois_per_image,
                            int num_classes, const float image_shape[3], const float std_dev[6], float min_confidence,
                            float nms_threshold, int nms_mode, int max_instances, float *detections, int *det_count,
                            void *workspace, size_t workspace_bytes, roi3d_stream_t stream)
{
    if (nms_mode != ROI3D_NMS_REFERENCE_2D && nms_mode != ROI3D_NMS_3D) return ROI3D_EINVAL;
    if (images < 0 || rois_per_image < 0 || num_classes < 2 || max_instances < 0 || !image_shape || !std_dev) return ROI3D_EINVAL;
    if (!(nms_threshold >= 0.0f && nms_threshold <= 1.0f)) return ROI3D_EINVAL;
    if (!(image_shape[0] > 0.0f && image_shape[1] > 0.0f && image_shape[2] > 0.0f)) return ROI3D_EINVAL;
    if (images == 0 || max_instances == 0) return ROI3D_OK;
    if (!detections || (rois_per_image > 0 && (!rois || !probs || !deltas))) return ROI3D_EINVAL;
    if ((long long)images * max_instances * 8 > 0x7fffffffll || (long long)images * rois_per_image > 0x7fffffffll) return ROI3D_EUNSUPPORTED;
    return launch_refine_detections(rois, probs, deltas, images, rois_per_image, num_classes, image_shape, std_dev,
                                    min_confidence, nms_threshold, nms_mode, max_instances, detections, det_count, workspace,
                                    workspace_bytes, static_cast<cudaStream_t>(stream));
}

int roi3d_mask_targets(const void *masks, int mask_dtype, int G, int H, int W, int D, const float *boxes,
                       const int *assignment, int n, int mh, int mw, int md, float *targets, unsigned char *bits,
                       roi3d_stream_t stream)
{
    if (G < 0 || H <= 0 || W <= 0 || D <= 0 || n < 0 || mh <= 0 || mw <= 0 || md <= 0) return ROI3D_EINVAL;
    if (mask_dtype != ROI3D_MASK_F32 && mask_dtype != ROI3D_MASK_U8) return ROI3D_EINVAL;
    if (!targets && !bits) return ROI3D_EINVAL;
    if (n == 0) return ROI3D_OK;
    if (!masks || !boxes || G == 0) return ROI3D_EINVAL;
    if (reinterpret_cast<uintptr_t>(bits) & 3) return ROI3D_EINVAL;
    if ((long long)H * W * D > 0x7fffffffll) return ROI3D_EUNSUPPORTED;
    return launch_mask_targets(masks, mask_dtype, H, W, D, boxes, assignment, n, mh, mw, md, targets, bits,
                               static_cast<cudaStream_t>(stream));
}

int roi3d_pack_f16(const float *x, long long n, void *half_out, roi3d_stream_t stream)
{
    if (n < 0) return ROI3D_EINVAL;
    if (n == 0) return ROI3D_OK;
    if (!x || !half_out || (reinterpret_cast<uintptr_t>(half_out) & 1) || (reinterpret_cast<uintptr_t>(x) & 3)) return ROI3D_EINVAL;
    return launch_f32_to_f16(x, n, half_out, static_cast<cudaStream_t>(stream));
}

int roi3d_unpack_f16(const void *half_in, long long n, float *y, roi3d_stream_t stream)
{
    if (n < 0) return ROI3D_EINVAL;
    if (n == 0) return ROI3D_OK;
    if (!half_in || !y || (reinterpret_cast<uintptr_t>(half_in) & 1) || (reinterpret_cast<uintptr_t>(y) & 3)) return ROI3D_EINVAL;
    return launch_f16_to_f32(half_in, n, y, static_cast<cudaStream_t>(stream));
}

int roi3d_pack_bits(const float *x, long long n, unsigned char *bits, roi3d_stream_t stream)
{
    if (n < 0) return ROI3D_EINVAL;
    if (n == 0) return ROI3D_OK;
    if (!x || !bits || (reinterpret_cast<uintptr_t>(bits) & 3)) return ROI3D_EINVAL;
    return launch_pack_bits(x, n, bits, static_cast<cudaStream_t>(stream));
}

int roi3d_unpack_bits(const unsigned char *bits, long long n, float *y, roi3d_stream_t stream)
{
    if (n < 0) return ROI3D_EINVAL;
    if (n == 0) return ROI3D_OK;
    if (!bits || !y) return ROI3D_EINVAL;
    return launch_unpack_bits(bits, n, y, static_cast<cudaStream_t>(stream));
}

// ---- ProposalLayer for one image in a single call (top-k -> decode -> NMS3D -> gather/pad), no host sync ----
struct ProposalLayout { size_t idx, val, boxes, keep, topk, nms, total; };
static ProposalLayout proposal_layout(int n, int k, int proposal_count) {
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    ProposalLayout L;
    size_t off = 0;
    L.idx = off;   off += up(sizeof(int) * (size_t)k);
    L.val = off;   off += up(sizeof(float) * (size_t)k);
    L.boxes = off; off += up(sizeof(float) * 6 * (size_t)k);
    L.keep = off;  off += up(sizeof(int) * (size_t)(proposal_count > 0 ? proposal_count : 1));
    L.topk = off;  off += up(topk_workspace_bytes(n));
    L.nms = off;   off += up(nms3d_workspace_bytes(k, 0));
    L.total = off;
    return L;
}

size_t roi3d_proposal_layer_workspace_bytes(int n_anchors, int pre_nms_limit, int proposal_count)
{
    if (n_anchors <= 0 || pre_nms_limit <= 0) return 256;
    return proposal_layout(n_anchors, pre_nms_limit < n_anchors ? pre_nms_limit : n_anchors, proposal_count).total;
}

int roi3d_proposal_layer(const float *scores, const float *deltas, const float *anchors, int n_anchors,
                         const float std_dev[6], float image_depth, int pre_nms_limit, int proposal_count,
                         float nms_threshold, float *proposals, int *count, void *workspace, size_t workspace_bytes,
                         roi3d_stream_t stream)
{
    if (n_anchors < 0 || pre_nms_limit < 0 || proposal_count < 0 || !std_dev || !count) return ROI3D_EINVAL;
    if (!(nms_threshold >= 0.0f && nms_threshold <= 1.0f)) return ROI3D_EINVAL;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int k = pre_nms_limit < n_anchors ? pre_nms_limit : n_anchors;
    if (proposal_count == 0) { ROI3D_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int), s)); return ROI3D_OK; }
    if (!proposals) return ROI3D_EINVAL;
    if (k == 0) {
        ROI3D_CUDA_TRY(cudaMemsetAsync(count, 0, sizeof(int), s));
        ROI3D_CUDA_TRY(cudaMemsetAsync(proposals, 0, sizeof(float) * 6 * (size_t)proposal_count, s));
        return ROI3D_OK;
    }
    if (!scores || !deltas || !anchors) return ROI3D_EINVAL;
    const ProposalLayout L = proposal_layout(n_anchors, k, proposal_count);
    if (!workspace || workspace_bytes < L.total || (reinterpret_cast<uintptr_t>(workspace) & 255)) return ROI3D_EWORKSPACE;
    char *base = static_cast<char *>(workspace);
    int *idx = reinterpret_cast<int *>(base + L.idx);
    float *val = reinterpret_cast<float *>(base + L.val);
    float *boxes = reinterpret_cast<float *>(base + L.boxes);
    int *keep = reinterpret_cast<int *>(base + L.keep);
    int rc = launch_topk(scores, n_anchors, k, idx, val, base + L.topk, L.nms - L.topk, s);
    if (rc != ROI3D_OK) return rc;
    rc = launch_decode_proposals(anchors, deltas, idx, k, std_dev, image_depth, boxes, s);
    if (rc != ROI3D_OK) return rc;
    rc = launch_nms3d(boxes, val, nullptr, 0, k, proposal_count, nms_threshold, keep, count, base + L.nms, L.total - L.nms, s);
    if (rc != ROI3D_OK) return rc;
    return launch_gather_pad_boxes(boxes, keep, count, proposal_count, proposals, s);
}

static int pyramid_check(const int level_shapes[4][3], int B, int C, const float *boxes, int rois_per_image,
                         const float image_shape[3], int ph, int pw, int pd)
{
    if (!level_shapes || !image_shape || B <= 0 || C <= 0 || rois_per_image < 0 || ph <= 0 || pw <= 0 || pd <= 0) return ROI3D_EINVAL;
    if (rois_per_image > 0 && !boxes) return ROI3D_EINVAL;
    for (int l = 0; l < 4; ++l) {
        if (level_shapes[l][0] <= 0 || level_shapes[l][1] <= 0 || level_shapes[l][2] <= 0) return ROI3D_EINVAL;
        if ((long long)level_shapes[l][0] * level_shapes[l][1] * level_shapes[l][2] * C >= (1ll << 31)) return ROI3D_EUNSUPPORTED;
    }
    if (C % 4 != 0 || ph > 64 || pw > 64 || pd > 64) return ROI3D_EUNSUPPORTED;     // the fused path is the plane kernel
    return ROI3D_OK;
}

static int pyramid_fwd_any(const float *const feature_maps[4], const int level_shapes[4][3], int B, int C,
                           const float *boxes, int rois_per_image, const float image_shape[3],
                           int ph, int pw, int pd, void *pooled, bool half_out, void *workspace, size_t workspace_bytes,
                           roi3d_stream_t stream)
{
    const int rc = pyramid_check(level_shapes, B, C, boxes, rois_per_image, image_shape, ph, pw, pd);
    if (rc != ROI3D_OK) return rc;
    if (rois_per_image == 0) return ROI3D_OK;
    if (!feature_maps || !pooled) return ROI3D_EINVAL;
    int H[4], W[4], D[4];
    for (int l = 0; l < 4; ++l) {
        if (!feature_maps[l] || (reinterpret_cast<uintptr_t>(feature_maps[l]) & 15)) return ROI3D_EINVAL;
        H[l] = level_shapes[l][0]; W[l] = level_shapes[l][1]; D[l] = level_shapes[l][2];
    }
    if (reinterpret_cast<uintptr_t>(pooled) & (half_out ? 7 : 15)) return ROI3D_EINVAL;
    int orc;
    const int *perm = order_rois(boxes, nullptr, B * rois_per_image, rois_per_image, ph * pw * pd, C, workspace, workspace_bytes,
                                 static_cast<cudaStream_t>(stream), &orc);
    if (orc != ROI3D_OK) return orc;
    return launch_pyramid_fwd(feature_maps, H, W, D, B, C, boxes, rois_per_image, image_shape[0], image_shape[1],
                              image_shape[2], ph, pw, pd, pooled, half_out, static_cast<cudaStream_t>(stream), perm);
}

int roi3d_pyramid_roi_align_fwd(const float *const feature_maps[4], const int level_shapes[4][3], int B, int C,
                                const float *boxes, int rois_per_image, const float image_shape[3],
                                int ph, int pw, int pd, float *pooled, roi3d_stream_t stream)
{
    return pyramid_fwd_any(feature_maps, level_shapes, B, C, boxes, rois_per_image, image_shape, ph, pw, pd, pooled, false, nullptr, 0, stream);
}

int roi3d_pyramid_roi_align_fwd_ws(const float *const feature_maps[4], const int level_shapes[4][3], int B, int C,
                                   const float *boxes, int rois_per_image, const float image_shape[3],
                                   int ph, int pw, int pd, float *pooled, void *workspace, size_t workspace_bytes,
                                   roi3d_stream_t stream)
{
    return pyramid_fwd_any(feature_maps, level_shapes, B, C, boxes, rois_per_image, image_shape, ph, pw, pd, pooled, false,
                           workspace, workspace_bytes, stream);
}

int roi3d_pyramid_roi_align_fwd_f16(const float *const feature_maps[4], const int level_shapes[4][3], int B, int C,
                                    const float *boxes, int rois_per_image, const float image_shape[3],
                                    int ph, int pw, int pd, void *pooled_f16, roi3d_stream_t stream)
{
    return pyramid_fwd_any(feature_maps, level_shapes, B, C, boxes, rois_per_image, image_shape, ph, pw, pd, pooled_f16, true, nullptr, 0, stream);
}

int roi3d_pyramid_roi_align_fwd_f16_ws(const float *const feature_maps[4], const int level_shapes[4][3], int B, int C,
                                       const float *boxes, int rois_per_image, const float image_shape[3],
                                       int ph, int pw, int pd, void *pooled_f16, void *workspace, size_t workspace_bytes,
                                       roi3d_stream_t stream)
{
    return pyramid_fwd_any(feature_maps, level_shapes, B, C, boxes, rois_per_image, image_shape, ph, pw, pd, pooled_f16, true,
                           workspace, workspace_bytes, stream);
}

int roi3d_pyramid_roi_align_grad(const float *grads, float *const grad_maps[4], const int level_shapes[4][3], int B, int C,
                                 const float *boxes, int rois_per_image, const float image_shape[3],
                                 int ph, int pw, int pd, roi3d_stream_t stream)
{
    return roi3d_pyramid_roi_align_grad_ws(grads, grad_maps, level_shapes, B, C, boxes, rois_per_image, image_shape, ph, pw, pd,
                                           nullptr, 0, stream);
}

int roi3d_pyramid_roi_align_grad_ws(const float *grads, float *const grad_maps[4], const int level_shapes[4][3], int B, int C,
                                    const float *boxes, int rois_per_image, const float image_shape[3],
                                    int ph, int pw, int pd, void *workspace, size_t workspace_bytes, roi3d_stream_t stream)
{
    const int rc = pyramid_check(level_shapes, B, C, boxes, rois_per_image, image_shape, ph, pw, pd);
    if (rc != ROI3D_OK) return rc;
    if (!grad_maps) return ROI3D_EINVAL;
    if (rois_per_image > 0 && !grads) return ROI3D_EINVAL;
    int H[4], W[4], D[4];
    for (int l = 0; l < 4; ++l) {
        if (!grad_maps[l] || (reinterpret_cast<uintptr_t>(grad_maps[l]) & 15)) return ROI3D_EINVAL;
        H[l] = level_shapes[l][0]; W[l] = level_shapes[l][1]; D[l] = level_shapes[l][2];
    }
    if (reinterpret_cast<uintptr_t>(grads) & 15) return ROI3D_EINVAL;
    int orc = ROI3D_OK;
    const int *perm = rois_per_image > 0 ? order_rois(boxes, nullptr, B * rois_per_image, rois_per_image, ph * pw * pd, C, workspace,
                                                      workspace_bytes, static_cast<cudaStream_t>(stream), &orc) : nullptr;
    if (orc != ROI3D_OK) return orc;
    return launch_pyramid_grad(grads, grad_maps, H, W, D, B, C, boxes, rois_per_image, image_shape[0], image_shape[1],
                               image_shape[2], ph, pw, pd, static_cast<cudaStream_t>(stream), perm);
}

}  // extern "C"
