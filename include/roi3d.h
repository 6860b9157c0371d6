/*
 * roi3d.h -- C ABI of the B200-native ROI hot path of 3D Mask R-CNN
 * (libroi3d_b200.so, hand-written sm_100a CUDA; no CPU fallback).
 *
 * This is the drop-in boundary.  Each entry point is what the reference's
 * TensorFlow custom-op kernels for this path would bind instead of their own
 * CPU loops.  The reference ships those kernels only as a prebuilt wheel
 * (core/custom_op/tensorflow_nms_car_3d-0.1.0-cp36-cp36m-linux_x86_64.whl);
 * "replaces" cites the op registration / OpKernel::Compute inside the wheel's
 * shared objects (notation <lib>@0xADDR) and the Python
 * call site in the reference tree.
 *
 * Conventions (identical to the reference ops):
 *   boxes   float32 [N,6] = (y1,x1,z1,y2,x2,z2), normalized to [0,1];
 *           normalized -> voxel is coord * (dim - 1).
 *   image   float32 [B,H,W,D,C], channel-last; y<->H, x<->W, z<->D.
 *   crops   float32 [N,ph,pw,pd,C].
 *   method  0 = "trilinear", 1 = "nearest".
 *
 * Ownership / threading: every pointer is a DEVICE pointer on the current CUDA
 * device unless stated otherwise and is owned by the caller (TF allocator,
 * torch, cudaMalloc ...).  The library allocates nothing and keeps no pointer; all calls are re-entrant.  Its only
 * state is (i) per calling thread: the last CUDA error, the launch counter and the tuning options of
 * roi3d_set_option, (ii) process-wide, write-once caches of device facts (SM count, shared-memory opt-in already
 * granted to a kernel).  No call can change what another thread's call computes.  All work is enqueued
 * on `stream` (a cudaStream_t passed as void*); no call synchronizes the device
 * or the stream.  Return value: ROI3D_OK or a negative ROI3D_E* code; nothing
 * is thrown and nothing aborts.
 */
#ifndef ROI3D_H_
#define ROI3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define ROI3D_API
#else
#define ROI3D_API __attribute__((visibility("default")))
#endif

typedef void *roi3d_stream_t;            /* cudaStream_t */

enum {
    ROI3D_OK            = 0,
    ROI3D_EINVAL        = -1,            /* bad argument (NULL, negative size, bad method ...) */
    ROI3D_EWORKSPACE    = -2,            /* workspace missing, too small or misaligned */
    ROI3D_EUNSUPPORTED  = -3,            /* shape beyond what the kernels index */
    ROI3D_ECUDA         = -4             /* a CUDA launch/runtime error; see roi3d_last_cuda_error() */
};

enum { ROI3D_METHOD_TRILINEAR = 0, ROI3D_METHOD_NEAREST = 1 };

ROI3D_API const char *roi3d_version(void);
ROI3D_API const char *roi3d_strerror(int code);
/* cudaError_t of the most recent ROI3D_ECUDA on the calling thread (0 if none). */
ROI3D_API int roi3d_last_cuda_error(void);

/* ---------------------------------------------------------------------------
 * NonMaxSuppression3D
 * replaces: REGISTER_OP("NonMaxSuppression3D") + NonMaxSuppression3DOp<CPUDevice>::Compute
 *           (NMS.so@0xe4e0) -> DoNonMaxSuppressionOp<float> (NMS.so@0xd0c0) with
 *           IOU<float> (NMS.so@0xb500); Python: core/custom_op/custom_op.py:25,
 *           called from core/models.py:453-455 and core/utils.py:498-500.
 *
 * Greedy hard NMS over `n` boxes: candidates are visited by descending score
 * (ties: lower index first; scores <= -FLT_MAX or NaN are never candidates); a
 * candidate is dropped when its 3-D IoU with an already selected box is
 * >= iou_thr; stops after max_out selections.  keep_idx receives the selected
 * ORIGINAL indices in selection order (capacity >= min(n,max_out) ints... see
 * note), *keep_count their number.  The reference's zero-volume quirk (a
 * selected box with volume <= 0 is emitted repeatedly until max_out, SURVEY.md
 * section 8 row a2) is reproduced, so keep_idx needs capacity max_out when that can
 * happen; capacity max(1, max_out) is always safe.
 *
 * keep_count may be a device pointer or a pinned/mapped host pointer (it is
 * written by a kernel).  The caller must synchronize `stream` before reading
 * it -- the one sync this path needs, because the output length is data
 * dependent.
 * workspace: roi3d_nms3d_workspace_bytes(n) bytes, 256-byte aligned.
 * ------------------------------------------------------------------------- */
ROI3D_API size_t roi3d_nms3d_workspace_bytes(int n);
ROI3D_API int roi3d_nms3d(const float *boxes, const float *scores, int n, int max_out, float iou_thr,
                          int *keep_idx, int *keep_count,
                          void *workspace, size_t workspace_bytes, roi3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Batched NonMaxSuppression3D: `segments` independent NMS problems in one set of launches.
 * replaces: the per-image loop of utils.batch_slice around the op (core/utils.py:1459-1544, called from
 *           ProposalLayer core/models.py:487-490) and the per-class loop of the upstream DetectionLayer design that
 *           utils.non_max_suppression_3d_graph (core/utils.py:467-503) exists for.
 * Segment z covers boxes/scores [seg_offsets[z], seg_offsets[z+1]) (device int32 [segments+1], ascending); every
 * segment has at most n_max boxes.  keep_idx is [segments, max_out] (row z: indices LOCAL to segment z, selection
 * order), keep_count [segments] (device or pinned host).  Each segment's result is bit-identical to roi3d_nms3d on
 * that segment alone.  workspace: roi3d_nms3d_batched_workspace_bytes(n_max, segments), 256-byte aligned.
 * ------------------------------------------------------------------------- */
ROI3D_API size_t roi3d_nms3d_batched_workspace_bytes(int n_max, int segments);
ROI3D_API int roi3d_nms3d_batched(const float *boxes, const float *scores, const int *seg_offsets, int segments,
                                  int n_max, int max_out, float iou_thr, int *keep_idx, int *keep_count,
                                  void *workspace, size_t workspace_bytes, roi3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * CropAndResize3D (forward)
 * replaces: REGISTER_OP("CropAndResize3D") + CropAndResize3DOp::Compute (CAR.so@0x4370);
 *           Python: core/custom_op/custom_op.py:22, called from core/models.py:663-664
 *           (PyramidROIAlign) and core/models.py:992-994 (mask targets).
 * box_index[i] selects the batch item of box i.  The reference does not check it (an index outside [0,B) reads
 * out of bounds there); here such a box reads nothing: its crop is filled with extrapolation_value, it scatters
 * nothing in CropAndResize3DGradImage and its CropAndResize3DGradBoxes row is zero.  n == 0 is valid (nothing is
 * launched).
 * ------------------------------------------------------------------------- */
ROI3D_API int roi3d_car3d_fwd(const float *image, int B, int H, int W, int D, int C,
                              const float *boxes, const int *box_index, int n,
                              int ph, int pw, int pd, int method, float extrapolation_value,
                              float *crops, roi3d_stream_t stream);
/* The same ops with a small caller-owned device workspace (roi3d_car3d_workspace_bytes(n) bytes, 4-byte aligned; what a
 * TF kernel gets from ctx->allocate_temp): the call first sorts the ROIs by (image, y centre) on the device -- one tiny
 * kernel -- and processes them in that order, so that ROIs whose footprints are neighbours in memory are in flight
 * together and share L2 lines (cfg2 14^3: forward -6 %, grad-image -6 %).  Results are those of the plain entry points
 * (the forward bit for bit).  workspace == NULL or too small: ROIs are processed in the order given. */
ROI3D_API size_t roi3d_car3d_workspace_bytes(int n);
ROI3D_API int roi3d_car3d_fwd_ws(const float *image, int B, int H, int W, int D, int C,
                                 const float *boxes, const int *box_index, int n,
                                 int ph, int pw, int pd, int method, float extrapolation_value,
                                 float *crops, void *workspace, size_t workspace_bytes, roi3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * CropAndResize3DGradImage
 * replaces: REGISTER_OP("CropAndResize3DGradImage") + CropAndResize3DGradImageOp::Compute
 *           (GI.so@0x3a80); Python: core/custom_op/custom_op.py:24,52-53.
 * grad_image [B,H,W,D,C] is zero-filled by the call, then every in-range crop
 * voxel scatter-adds its gradient to its 8 (trilinear) or 1 (nearest) taps.
 * ------------------------------------------------------------------------- */
ROI3D_API int roi3d_car3d_grad_image(const float *grads, const float *boxes, const int *box_ind, int n,
                                     int ph, int pw, int pd,
                                     int B, int H, int W, int D, int C, int method,
                                     float *grad_image, roi3d_stream_t stream);
ROI3D_API int roi3d_car3d_grad_image_ws(const float *grads, const float *boxes, const int *box_ind, int n,
                                     int ph, int pw, int pd,
                                     int B, int H, int W, int D, int C, int method,
                                     float *grad_image, void *workspace, size_t workspace_bytes, roi3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * CropAndResize3DGradBoxes (trilinear only, like the reference)
 * replaces: REGISTER_OP("CropAndResize3DGradBoxes") + CropAndResize3DGradBoxesOp::Compute
 *           (GB.so@0x3980); Python: core/custom_op/custom_op.py:23,63.
 * grad_boxes [N,6] is fully written by the call.
 * ------------------------------------------------------------------------- */
ROI3D_API int roi3d_car3d_grad_boxes(const float *grads, const float *image,
                                     int B, int H, int W, int D, int C,
                                     const float *boxes, const int *box_ind, int n,
                                     int ph, int pw, int pd,
                                     float *grad_boxes, roi3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Fused PyramidROIAlign3D (offered IN ADDITION to the drop-in ops)
 * replaces: PyramidROIAlign.call, core/models.py:604-685 -- box clip + min sizes (:615-632), level routing
 *           (:637-649), the four per-level crop_and_resize_3d calls (:663-664), the concat / top_k / gather order
 *           restore (:667-676) and the non-finite scrub (:683) -- by one launch (forward) and four zero-fills plus one
 *           launch (backward w.r.t. the feature maps; boxes are stop_gradient'ed upstream, :660).
 * feature_maps / grad_maps: P2..P5, float32 [B, H_l, W_l, D_l, C], level_shapes[l] = {H_l, W_l, D_l};
 * boxes [B, R, 6] normalized; image_shape = {H, W, D} of the input volume (image_meta); pooled / grads
 * [B, R, ph, pw, pd, C] in the boxes' own order.  Needs C % 4 == 0 and crop dims <= 64 (ROI3D_EUNSUPPORTED
 * otherwise: use the per-level ops).  The level index comes from fp32 cbrt/log2 like the TF graph; a ROI whose
 * level expression falls within an ulp of x.5 may round differently from another libm.
 * ------------------------------------------------------------------------- */
ROI3D_API int roi3d_pyramid_roi_align_fwd(const float *const feature_maps[4], const int level_shapes[4][3], int B, int C,
                                          const float *boxes, int rois_per_image, const float image_shape[3],
                                          int ph, int pw, int pd, float *pooled, roi3d_stream_t stream);
/* float16 output: the `rois_aligned` payload of the head-target files (core/models.py:3613 `ra.astype(np.float16)`)
 * written by the crop kernel itself -- bit-identical to roi3d_pack_f16 of the float32 result, half the bytes. */
ROI3D_API int roi3d_pyramid_roi_align_fwd_f16(const float *const feature_maps[4], const int level_shapes[4][3], int B, int C,
                                              const float *boxes, int rois_per_image, const float image_shape[3],
                                              int ph, int pw, int pd, void *pooled_f16, roi3d_stream_t stream);
ROI3D_API int roi3d_pyramid_roi_align_grad(const float *grads, float *const grad_maps[4], const int level_shapes[4][3],
                                           int B, int C, const float *boxes, int rois_per_image,
                                           const float image_shape[3], int ph, int pw, int pd, roi3d_stream_t stream);
/* ... and with the workspace of roi3d_car3d_workspace_bytes(B * rois_per_image) bytes: ROIs processed by (image, y centre) */
ROI3D_API int roi3d_pyramid_roi_align_fwd_ws(const float *const feature_maps[4], const int level_shapes[4][3], int B, int C,
                                             const float *boxes, int rois_per_image, const float image_shape[3],
                                             int ph, int pw, int pd, float *pooled, void *workspace, size_t workspace_bytes,
                                             roi3d_stream_t stream);
ROI3D_API int roi3d_pyramid_roi_align_fwd_f16_ws(const float *const feature_maps[4], const int level_shapes[4][3], int B, int C,
                                                 const float *boxes, int rois_per_image, const float image_shape[3],
                                                 int ph, int pw, int pd, void *pooled_f16, void *workspace,
                                                 size_t workspace_bytes, roi3d_stream_t stream);
ROI3D_API int roi3d_pyramid_roi_align_grad_ws(const float *grads, float *const grad_maps[4], const int level_shapes[4][3],
                                              int B, int C, const float *boxes, int rois_per_image,
                                              const float image_shape[3], int ph, int pw, int pd, void *workspace,
                                              size_t workspace_bytes, roi3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Box-space helpers either side of the ops
 * roi3d_overlaps3d        replaces overlaps_graph, core/models.py:695-733: overlaps [n,m] float32, fp32 arithmetic of
 *                         the TF graph (union clamped with 1e-10; NOT the NMS op's IOU<float>).
 * roi3d_decode_proposals  replaces ProposalLayer's per-anchor chain between top_k and NMS, core/models.py:397-447:
 *                         deltas * std_dev, clip +-3, apply_box_deltas_graph (:280-337), clip to [0,1], min sizes.
 *                         index (int32 [n], may be NULL) gathers anchors/deltas rows first (the top_k indices).
 * ------------------------------------------------------------------------- */
ROI3D_API int roi3d_overlaps3d(const float *boxes1, int n, const float *boxes2, int m, float *overlaps,
                               roi3d_stream_t stream);
/* roi3d_topk              replaces tf.nn.top_k(scores, k) ahead of NMS (core/models.py:403-404): the selected SET equals
 *                         TF's (threshold ties -> lower indices); idx_out [k] is in ASCENDING INDEX order (the NMS that
 *                         follows orders by (score, position) itself, so ties resolve as with top_k's sorted output);
 *                         scores_out [k] optional.  3-pass radix select + ordered compaction, 5 launches (the CTA that finishes last closes each pass), no host sync.
 * roi3d_gather_pad_boxes  replaces tf.gather(boxes, idx) + tf.pad to proposal_count (core/models.py:476-484); the keep
 *                         count is read on the device, so ProposalLayer needs no host synchronisation at all. */
ROI3D_API size_t roi3d_topk_workspace_bytes(int n);
ROI3D_API int roi3d_topk(const float *scores, int n, int k, int *idx_out, float *scores_out,
                         void *workspace, size_t workspace_bytes, roi3d_stream_t stream);
ROI3D_API int roi3d_gather_pad_boxes(const float *boxes, const int *keep_idx, const int *keep_count, int proposal_count,
                                     float *proposals, roi3d_stream_t stream);
/* roi3d_proposal_layer    one image of ProposalLayer.call (core/models.py:382-500) in a single call: the four entry points
 *                         above chained on `stream` (top-k of the foreground scores, decode, NMS3D, gather + zero-pad),
 *                         no host synchronisation; count (device int32) receives the number of real proposals. */
ROI3D_API size_t roi3d_proposal_layer_workspace_bytes(int n_anchors, int pre_nms_limit, int proposal_count);
ROI3D_API int roi3d_proposal_layer(const float *scores, const float *deltas, const float *anchors, int n_anchors,
                                   const float std_dev[6], float image_depth, int pre_nms_limit, int proposal_count,
                                   float nms_threshold, float *proposals, int *count,
                                   void *workspace, size_t workspace_bytes, roi3d_stream_t stream);
ROI3D_API int roi3d_decode_proposals(const float *anchors, const float *deltas, const int *index, int n,
                                     const float std_dev[6], float image_depth, float *boxes, roi3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * DetectionLayer on the device
 * replaces: refine_detections_graph + the utils.batch_slice loop of DetectionLayer.call, core/models.py:1415-1575, for
 *           the whole batch in one set of launches.
 * nms_mode  ROI3D_NMS_REFERENCE_2D: the NMS this fork's graph performs -- tf.image.non_max_suppression on the (y, x)
 *           projection of the refined boxes, suppressing on IoU > threshold (core/models.py:1496-1501).  It runs on the
 *           3-D kernels with every box given the depth interval [0, 1], which makes IoU3D == IoU2D bit for bit, and the
 *           threshold moved up by one ulp (iou > t <=> iou >= nextafter(t)).  This is the drop-in behaviour.
 *           ROI3D_NMS_3D: the 3-D op (IoU over the volume, suppressing on IoU >= threshold): the upstream design that
 *           utils.non_max_suppression_3d_graph (core/utils.py:467-503) was written for and BASELINE cfg3 names; opt-in,
 *           its detections differ from the fork's whenever boxes overlap in y/x but not in z.
 * rois [B,R,6] normalised, probs [B,R,num_classes], deltas [B,R,num_classes,6] (device, float32).  As in this fork
 * the class is always 1 (fg_probs = probs[:,1], :1441).  Per ROI: score >= min_confidence, deltas * std_dev,
 * apply_box_deltas_3d_graph in pixels (core/utils.py:412-464), clip to [0,H]x[0,W]x[0,D], sizes >= (1,1,0.5) px; NMS
 * with max_instances outputs; rows in selection order = descending score (the graph's top_k is then the identity);
 * boxes back to normalised coordinates clipped to [0,1]; rows past the kept count are zero.
 * detections [B,max_instances,8] = (y1,x1,z1,y2,x2,z2,class_id,score); det_count [B] (device, optional, may be NULL).
 * image_shape / std_dev are host arrays.  No host synchronisation.  workspace:
 * roi3d_refine_detections_workspace_bytes(B, R, max_instances), 256-byte aligned.
 * ------------------------------------------------------------------------- */
enum { ROI3D_NMS_REFERENCE_2D = 0, ROI3D_NMS_3D = 1 };
ROI3D_API size_t roi3d_refine_detections_workspace_bytes(int images, int rois_per_image, int max_instances);
ROI3D_API int roi3d_refine_detections(const float *rois, const float *probs, const float *deltas, int images,
                                      int rois_per_image, int num_classes, const float image_shape[3],
                                      const float std_dev[6], float min_confidence, float nms_threshold, int nms_mode,
                                      int max_instances, float *detections, int *det_count,
                                      void *workspace, size_t workspace_bytes, roi3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Mask targets and the target files' wire format
 * roi3d_mask_targets  replaces detection_targets_graph._get_masks, core/models.py:972-1005: tf.gather of the assigned
 *                     ground-truth masks, cast to float32, CropAndResize3D (C = 1, trilinear, extrapolation 0) and
 *                     tf.round, in one kernel.  masks [G,H,W,D] (the transposed layout of :973), mask_dtype
 *                     ROI3D_MASK_F32 / ROI3D_MASK_U8; boxes [n,6]; assignment int32 [n] (NULL = identity, the
 *                     box_ids = range(n) of :989).  targets float32 [n,mh,mw,md] and/or bits (the packed form below,
 *                     4-byte aligned); either may be NULL, not both.
 * roi3d_pack_f16 / roi3d_unpack_f16   ndarray.astype(float16) (round to nearest even) and back: the `rois_aligned`
 *                     payload of the target files, core/models.py:3613.
 * roi3d_pack_bits / roi3d_unpack_bits numpy.packbits((x > 0.5).reshape(-1)) (MSB first, zero padded) and
 *                     numpy.unpackbits(...)[:n] as float32: the `mask_bits` / `tm_bits` payloads, :3585-3595.
 * n counts ELEMENTS.  The npz container around the payloads stays on the host.
 * ------------------------------------------------------------------------- */
#define ROI3D_MASK_F32 0
#define ROI3D_MASK_U8 1
ROI3D_API int roi3d_mask_targets(const void *masks, int mask_dtype, int G, int H, int W, int D, const float *boxes,
                                 const int *assignment, int n, int mh, int mw, int md, float *targets,
                                 unsigned char *bits, roi3d_stream_t stream);
ROI3D_API int roi3d_pack_f16(const float *x, long long n, void *half_out, roi3d_stream_t stream);
ROI3D_API int roi3d_unpack_f16(const void *half_in, long long n, float *y, roi3d_stream_t stream);
ROI3D_API int roi3d_pack_bits(const float *x, long long n, unsigned char *bits, roi3d_stream_t stream);
ROI3D_API int roi3d_unpack_bits(const unsigned char *bits, long long n, float *y, roi3d_stream_t stream);

/* ---------------------------------------------------------------------------
 * Tuning / introspection (not part of the reference surface).
 * roi3d_set_option: knobs of the CALLING THREAD (thread-local; another thread's calls are unaffected), used by the
 * benchmarks and tests to select a kernel variant; the defaults (0) are the production choice.
 *   "car_fwd_variant"   0 = auto, 1 = direct gather, 2 = plane-staged separable, 3 = plane-staged fed by TMA bulk
 *                       copies, 4 = row-walk separable (z / x / y lerps each evaluated once), 5 = plane-staged fed by
 *                       TMA gather4 row copies (cp.async.bulk.tensor.2d ... tile::gather4); 3-5 are bit-exact; 3 and 4
 *                       are slower, 5 is level with 2 at 14^3 and slower at 7^3: opt-in
 *   "car_bwd_variant"   0 = auto (= 4), 1 = direct scatter, 2 = plane-staged RED scatter with per-thread-load staging,
 *                       3 = output-stationary (every voxel stored once, no zero-fill, no atomics, deterministic;
 *                       slower: opt-in), 4 = plane-staged RED scatter whose grads slices are staged by TMA tensor
 *                       tile copies (falls back to 2 where the driver cannot encode the tensor map)
 *   "car_lanes_v"       0 = auto, 1 / 2 = float4 channel groups per thread in the plane / row-walk kernels
 *   "car_ctas_per_sm_target"   grid sizing of the plane kernels: depth-sample splits are chosen so that about this
 *                       many CTAs per SM exist (0 = default 16)
 *   "pdl" (alias "nms_pdl")    0 = dependent kernels (NMS mask / scan chain, grad-image scatter behind its zero-fill)
 *                       are launched with programmatic dependent launch, 1 = plain stream order
 *   "nms_variant"       0 = speculative head/tail schedule, 1 = single phase (full mask, then one scan)
 *   "nms_sort_variant"  0 = auto (bucketed from 8192 boxes), 1 = rank by counting, 2 = bucketed sort
 *   experiment knobs read by profiles/*.py only: "car_os_tile_depth", "car_os_ring_stages", "car_os_stage_kib",
 *   "car_os_debug", "car_bwd_image_split", "car_sep_rows", "car_sep_ring", "car_fill_ctas_per_sm",
 *   "car_bwd_stage_kib", "car_experiment"
 * Returns ROI3D_EINVAL for an unknown name.  roi3d_kernel_launches() returns
 * the number of kernel launches this library has enqueued on the calling
 * thread since the last roi3d_reset_kernel_launches().
 * ------------------------------------------------------------------------- */
ROI3D_API int roi3d_set_option(const char *name, int value);
ROI3D_API int roi3d_get_option(const char *name, int *value);
ROI3D_API long long roi3d_kernel_launches(void);
ROI3D_API void roi3d_reset_kernel_launches(void);

#ifdef __cplusplus
}
#endif
#endif /* ROI3D_H_ */
