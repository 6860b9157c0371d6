"""CPU oracle for the ROI hot path -- TEST INFRASTRUCTURE, not product code.

ctypes front-end of ``oracle/roi3d_oracle.c`` (a restatement of the reference's
four native ops, see the header of that file for the binary addresses each
function follows).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package; the product package ``3d-mask-r-cnn_b200/`` never does.

All functions take and return numpy arrays in the reference's layouts:
boxes ``float32 [N,6] = (y1,x1,z1,y2,x2,z2)`` normalized, volumes
``float32 [B,H,W,D,C]`` channel-last, crops ``float32 [N,ph,pw,pd,C]``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)

METHODS = {"trilinear": 0, "nearest": 1}


def build(force=False):
    """Compile oracle/roi3d_oracle.c (plain + OpenMP flavours) with gcc."""
    plain = os.path.join(_BUILD, "libroi3d_oracle.so")
    omp = os.path.join(_BUILD, "libroi3d_oracle_omp.so")
    src = os.path.join(_HERE, "roi3d_oracle.c")
    stale = force or not (os.path.exists(plain) and os.path.exists(omp)) or \
        os.path.getmtime(src) > min(os.path.getmtime(plain), os.path.getmtime(omp))
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return plain, omp


_libs = {}


def _lib(omp=False):
    key = bool(omp)
    if key not in _libs:
        plain, ompp = build()
        lib = ctypes.CDLL(ompp if omp else plain)
        lib.roi3d_oracle_iou3d.restype = ctypes.c_float
        lib.roi3d_oracle_iou3d.argtypes = [_f32p, ctypes.c_int, ctypes.c_int]
        lib.roi3d_oracle_nms3d.restype = ctypes.c_int
        lib.roi3d_oracle_nms3d.argtypes = [_f32p, _f32p, ctypes.c_int, ctypes.c_int, ctypes.c_float, _i32p]
        lib.roi3d_oracle_car3d_fwd.restype = None
        lib.roi3d_oracle_car3d_fwd.argtypes = [_f32p] + [ctypes.c_int] * 5 + [_f32p, _i32p] + \
            [ctypes.c_int] * 5 + [ctypes.c_float, _f32p, ctypes.c_int]
        lib.roi3d_oracle_car3d_grad_image.restype = None
        lib.roi3d_oracle_car3d_grad_image.argtypes = [_f32p, _f32p, _i32p] + [ctypes.c_int] * 10 + \
            [_f32p, ctypes.c_int]
        lib.roi3d_oracle_car3d_grad_boxes.restype = None
        lib.roi3d_oracle_car3d_grad_boxes.argtypes = [_f32p, _f32p] + [ctypes.c_int] * 5 + [_f32p, _i32p] + \
            [ctypes.c_int] * 4 + [_f32p]
        lib.roi3d_oracle_iou_matrix.restype = None
        lib.roi3d_oracle_iou_matrix.argtypes = [_f32p, ctypes.c_int, _f32p]
        lib.roi3d_oracle_max_threads.restype = ctypes.c_int
        _libs[key] = lib
    return _libs[key]


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _fp(a):
    return a.ctypes.data_as(_f32p)


def _ip(a):
    return a.ctypes.data_as(_i32p)


def max_threads():
    return int(_lib(True).roi3d_oracle_max_threads())


def iou3d(boxes, i, j):
    boxes = _f32(boxes)
    return float(_lib().roi3d_oracle_iou3d(_fp(boxes), int(i), int(j)))


def iou_matrix(boxes):
    boxes = _f32(boxes).reshape(-1, 6)
    n = boxes.shape[0]
    out = np.empty((n, n), np.float32)
    _lib().roi3d_oracle_iou_matrix(_fp(boxes), n, _fp(out))
    return out


def non_max_suppression_3d(boxes, scores, max_output_size, iou_threshold=0.5):
    """Reference NonMaxSuppression3D (NMS.so@0xe4e0 -> 0xd0c0): int32 [M]."""
    boxes = _f32(boxes).reshape(-1, 6)
    scores = _f32(scores).reshape(-1)
    n = boxes.shape[0]
    assert scores.shape[0] == n
    out = np.empty(max(int(max_output_size), 1), np.int32)
    m = _lib().roi3d_oracle_nms3d(_fp(boxes), _fp(scores), n, int(max_output_size),
                                  float(iou_threshold), _ip(out))
    return out[:m].copy()


def crop_and_resize_3d(image, boxes, box_index, crop_size, method_name="trilinear",
                       extrapolation_value=0.0, threads=1):
    """Reference CropAndResize3D (CAR.so@0x4370): float32 [N,ph,pw,pd,C]."""
    image = _f32(image)
    boxes = _f32(boxes).reshape(-1, 6)
    box_index = _i32(box_index).reshape(-1)
    B, H, W, D, C = image.shape
    ph, pw, pd = (int(v) for v in crop_size)
    n = boxes.shape[0]
    out = np.empty((n, ph, pw, pd, C), np.float32)
    _lib(threads > 1).roi3d_oracle_car3d_fwd(_fp(image), B, H, W, D, C, _fp(boxes), _ip(box_index), n,
                                            ph, pw, pd, METHODS[method_name], float(extrapolation_value),
                                            _fp(out), int(threads))
    return out


def crop_and_resize_3d_grad_image(grads, boxes, box_ind, image_size, method_name="trilinear", threads=1):
    """Reference CropAndResize3DGradImage (GI.so@0x3a80): float32 [B,H,W,D,C]."""
    grads = _f32(grads)
    boxes = _f32(boxes).reshape(-1, 6)
    box_ind = _i32(box_ind).reshape(-1)
    B, H, W, D, C = (int(v) for v in image_size)
    n, ph, pw, pd, Cg = grads.shape
    assert Cg == C and boxes.shape[0] == n
    out = np.empty((B, H, W, D, C), np.float32)
    _lib(threads > 1).roi3d_oracle_car3d_grad_image(_fp(grads), _fp(boxes), _ip(box_ind), n, ph, pw, pd,
                                                   B, H, W, D, C, METHODS[method_name], _fp(out), int(threads))
    return out


def crop_and_resize_3d_grad_boxes(grads, image, boxes, box_ind):
    """Reference CropAndResize3DGradBoxes (GB.so@0x3980): float32 [N,6]."""
    grads = _f32(grads)
    image = _f32(image)
    boxes = _f32(boxes).reshape(-1, 6)
    box_ind = _i32(box_ind).reshape(-1)
    B, H, W, D, C = image.shape
    n, ph, pw, pd, _ = grads.shape
    out = np.empty((n, 6), np.float32)
    _lib().roi3d_oracle_car3d_grad_boxes(_fp(grads), _fp(image), B, H, W, D, C, _fp(boxes), _ip(box_ind), n,
                                         ph, pw, pd, _fp(out))
    return out


# ---------------------------------------------------------------------------------------------
# PyramidROIAlign layer (core/models.py:604-685) restated in numpy on top of the op oracle -- checker for the
# fused kernel.  fp32 arithmetic like the TF graph.
# ---------------------------------------------------------------------------------------------
def pyramid_prepare(boxes, image_shape):
    """Clip to [0,1], min sizes (core/models.py:615-632) and level routing (:637-649). boxes [B,R,6]."""
    b = np.clip(np.asarray(boxes, np.float32), np.float32(0), np.float32(1)).copy()
    H, W, D = (np.float32(v) for v in image_shape)
    eps = np.float32(1e-6)
    b[..., 3] = np.maximum(b[..., 3], b[..., 0] + eps)
    b[..., 4] = np.maximum(b[..., 4], b[..., 1] + eps)
    b[..., 5] = np.maximum(b[..., 5], b[..., 2] + np.float32(1.0) / np.maximum(D, np.float32(1.0)))
    vol = ((b[..., 3] - b[..., 0]) * (b[..., 4] - b[..., 1]) * (b[..., 5] - b[..., 2])).astype(np.float32)
    area = np.float32(H * W * D)
    third = np.float32(1.0 / 3.0)
    ratio = (np.power(vol, third) / (np.float32(224.0) / np.power(area, third))).astype(np.float32)
    lvl = (np.log(ratio) / np.log(np.float32(2.0))).astype(np.float32)
    level = np.minimum(5, np.maximum(2, 4 + np.rint(lvl).astype(np.int32)))
    return b, level


def pyramid_roi_align(boxes, image_shape, feature_maps, pool_shape):
    """[B,R,ph,pw,pd,C]: per-level CropAndResize3D, original order, non-finite -> 0."""
    b, level = pyramid_prepare(boxes, image_shape)
    B, R = b.shape[:2]
    C = feature_maps[0].shape[4]
    out = np.zeros((B, R) + tuple(pool_shape) + (C,), np.float32)
    for i, lv in enumerate(range(2, 6)):
        ib, ir = np.nonzero(level == lv)
        if len(ib):
            out[ib, ir] = crop_and_resize_3d(feature_maps[i], b[ib, ir], ib.astype(np.int32), pool_shape)
    return np.where(np.isfinite(out), out, np.float32(0))


def pyramid_roi_align_grad(grads, boxes, image_shape, level_shapes):
    """Gradients w.r.t. P2..P5 (list of [B,H_l,W_l,D_l,C])."""
    b, level = pyramid_prepare(boxes, image_shape)
    outs = []
    for i, lv in enumerate(range(2, 6)):
        ib, ir = np.nonzero(level == lv)
        outs.append(crop_and_resize_3d_grad_image(np.asarray(grads, np.float32)[ib, ir], b[ib, ir], ib.astype(np.int32),
                                                  level_shapes[i]))
    return outs


# ---------------------------------------------------------------------------------------------
# box-space helpers: numpy restatements of the TF graph code, fp32
# ---------------------------------------------------------------------------------------------
def overlaps_graph(boxes1, boxes2):
    """core/models.py:695-733."""
    b1 = np.asarray(boxes1, np.float32)[:, None, :]
    b2 = np.asarray(boxes2, np.float32)[None, :, :]
    z = np.float32(0)
    y1 = np.maximum(b1[..., 0], b2[..., 0]); x1 = np.maximum(b1[..., 1], b2[..., 1]); z1 = np.maximum(b1[..., 2], b2[..., 2])
    y2 = np.minimum(b1[..., 3], b2[..., 3]); x2 = np.minimum(b1[..., 4], b2[..., 4]); z2 = np.minimum(b1[..., 5], b2[..., 5])
    inter = np.maximum(y2 - y1, z) * np.maximum(x2 - x1, z) * np.maximum(z2 - z1, z)
    v1 = (b1[..., 3] - b1[..., 0]) * (b1[..., 4] - b1[..., 1]) * (b1[..., 5] - b1[..., 2])
    v2 = (b2[..., 3] - b2[..., 0]) * (b2[..., 4] - b2[..., 1]) * (b2[..., 5] - b2[..., 2])
    union = v1 + v2 - inter
    return (inter / np.maximum(union, np.float32(1e-10))).astype(np.float32)


def decode_proposals(anchors, deltas, std_dev, image_depth, index=None):
    """core/models.py:397-447 with apply_box_deltas_graph (:280-337) and clip_boxes_graph (:340-364)."""
    a = np.asarray(anchors, np.float32)
    d = np.asarray(deltas, np.float32) * np.asarray(std_dev, np.float32)[None, :]
    d = np.clip(d, np.float32(-3), np.float32(3))
    if index is not None:
        a, d = a[index], d[index]
    half = np.float32(0.5)
    h = a[:, 3] - a[:, 0]; w = a[:, 4] - a[:, 1]; dp = a[:, 5] - a[:, 2]
    cy = a[:, 0] + half * h; cx = a[:, 1] + half * w; cz = a[:, 2] + half * dp
    cy = cy + d[:, 0] * h; cx = cx + d[:, 1] * w; cz = cz + d[:, 2] * dp
    h = h * np.exp(d[:, 3]); w = w * np.exp(d[:, 4]); dp = dp * np.exp(d[:, 5])
    y1 = cy - half * h; x1 = cx - half * w; z1 = cz - half * dp
    out = np.stack([y1, x1, z1, y1 + h, x1 + w, z1 + dp], axis=1).astype(np.float32)
    out = np.clip(out, np.float32(0), np.float32(1))
    depth = np.float32(max(float(image_depth), 1.0))
    min_dz = np.maximum(np.float32(1.0) / depth, np.float32(1e-4))
    out[:, 3] = np.maximum(out[:, 3], out[:, 0] + np.float32(1e-6))
    out[:, 4] = np.maximum(out[:, 4], out[:, 1] + np.float32(1e-6))
    out[:, 5] = np.maximum(out[:, 5], out[:, 2] + min_dz)
    return out


# ---------------------------------------------------------------------------------------------
# DetectionLayer and the target files: numpy restatements, fp32
# ---------------------------------------------------------------------------------------------
def refine_decode(rois, probs, deltas, image_shape, min_confidence, bbox_std_dev=(0.1, 0.1, 0.1, 0.2, 0.2, 0.2)):
    """The per-ROI half of refine_detections_graph (core/models.py:1440-1488): pixel boxes after deltas + clip, the
    class-1 score, and the mask of ROIs that pass the confidence and min-size filters."""
    rois = np.asarray(rois, np.float32)
    score = np.asarray(probs, np.float32)[:, 1]
    d = np.asarray(deltas, np.float32)[:, 1, :] * np.asarray(bbox_std_dev, np.float32)[None, :]
    dim = np.asarray(image_shape, np.float32)[:3]
    scale = np.concatenate([dim, dim])
    b = rois * scale[None, :]                                              # denorm_boxes_3d_graph, core/utils.py:401
    half, lim = np.float32(0.5), np.float32(np.log(np.float32(62.5)))      # apply_box_deltas_3d_graph, :412-464
    out = np.empty_like(b)
    for a in range(3):
        ln = b[:, a + 3] - b[:, a]
        ctr = b[:, a] + half * ln
        ds = np.clip(d[:, a + 3], -lim, lim)
        ctr2 = ctr + d[:, a] * ln
        ln2 = ln * np.exp(ds)
        lo = ctr2 - half * ln2
        out[:, a] = np.clip(lo, np.float32(0), dim[a])
        out[:, a + 3] = np.clip(lo + ln2, np.float32(0), dim[a])
    ok = (score >= np.float32(min_confidence)) & (out[:, 3] - out[:, 0] >= 1.0) & (out[:, 4] - out[:, 1] >= 1.0) & \
        (out[:, 5] - out[:, 2] >= 0.5)
    return out.astype(np.float32), score, ok


def non_max_suppression_2d_tf(boxes4, scores, max_output_size, iou_threshold):
    """tf.image.non_max_suppression (NonMaxSuppressionV3, TF r2.2 core/kernels/non_max_suppression_op.cc) restated from
    the published algorithm -- TensorFlow is absent here, so this restatement is UNPINNED by execution: candidates by
    descending score (ties -> lower index), a candidate is dropped when its IoU with a selected box is > threshold
    (strict), IoU in float32 as area_i = (ymax - ymin) * (xmax - xmin), inter / (area_i + area_j - inter), 0 for an
    empty box.  Called like the reference does (core/models.py:1496-1501), i.e. with (x1, y1, x2, y2) columns."""
    b = np.ascontiguousarray(boxes4, np.float32).reshape(-1, 4)
    s = np.ascontiguousarray(scores, np.float32).reshape(-1)
    order = np.lexsort((np.arange(len(s)), -s.astype(np.float64)))
    order = [i for i in order if s[i] > -np.finfo(np.float32).max]
    f = np.float32
    lo0, hi0 = np.minimum(b[:, 0], b[:, 2]), np.maximum(b[:, 0], b[:, 2])
    lo1, hi1 = np.minimum(b[:, 1], b[:, 3]), np.maximum(b[:, 1], b[:, 3])
    area = ((hi0 - lo0).astype(f) * (hi1 - lo1).astype(f)).astype(f)
    thr = f(iou_threshold)
    sel = []
    for i in order:
        if len(sel) >= int(max_output_size):
            break
        if sel:
            j = np.asarray(sel)
            d0 = np.maximum((np.minimum(hi0[i], hi0[j]) - np.maximum(lo0[i], lo0[j])).astype(f), f(0))
            d1 = np.maximum((np.minimum(hi1[i], hi1[j]) - np.maximum(lo1[i], lo1[j])).astype(f), f(0))
            inter = (d0 * d1).astype(f)
            with np.errstate(divide="ignore", invalid="ignore"):
                iou = (inter / ((area[i] + area[j]).astype(f) - inter).astype(f)).astype(f)
            iou = np.where((area[i] <= 0) | (area[j] <= 0), f(0), iou)
            if np.any(iou > thr):
                continue
        sel.append(int(i))
    return np.asarray(sel, np.int32)


def refine_detections(rois, probs, deltas, image_shape, min_confidence, nms_threshold,
                      bbox_std_dev=(0.1, 0.1, 0.1, 0.2, 0.2, 0.2), max_instances=200, boxes_px=None, nms_mode="reference_2d"):
    """refine_detections_graph (core/models.py:1415-1524) for one image, restated from the graph source (TensorFlow is
    absent: unpinned by execution).  ``nms_mode="reference_2d"``: the graph's own tf.image.non_max_suppression on the
    (y, x) projection; ``"3d"``: the 3-D op as the NMS (the upstream design, row f3).
    ``boxes_px`` overrides the decoded pixel boxes (tests pass the device's, whose expf may differ in the last ulp,
    to compare the selection exactly)."""
    px, score, ok = refine_decode(rois, probs, deltas, image_shape, min_confidence, bbox_std_dev)
    if boxes_px is not None:
        px = np.asarray(boxes_px, np.float32)
    det = np.zeros((int(max_instances), 8), np.float32)
    ix = np.nonzero(ok)[0]
    if len(ix) == 0:
        return det
    if nms_mode == "3d":
        sel = non_max_suppression_3d(px[ix], score[ix], int(max_instances), float(nms_threshold))
    else:
        p = px[ix]
        sel = non_max_suppression_2d_tf(np.stack([p[:, 1], p[:, 0], p[:, 4], p[:, 3]], axis=1), score[ix],
                                        int(max_instances), float(nms_threshold))
    fb, fs = px[ix][sel], score[ix][sel]
    order = np.argsort(-fs, kind="stable")                                 # tf.nn.top_k: descending, ties -> lower index
    fb, fs = fb[order], fs[order]
    dim = np.asarray(image_shape, np.float32)[:3]
    scale = np.concatenate([dim, dim])
    k = len(fs)
    det[:k, :6] = np.clip(fb / scale[None, :], np.float32(0), np.float32(1))
    det[:k, 6] = 1.0
    det[:k, 7] = fs
    return det


def mask_targets(masks, boxes, assignment, mask_shape):
    """detection_targets_graph._get_masks (core/models.py:972-1005): gather, cast, CropAndResize3D (C = 1), tf.round.
    ``masks [G,H,W,D]`` any dtype; returns float32 ``[N,mh,mw,md]``."""
    m = np.asarray(masks).astype(np.float32)[..., None]
    n = len(boxes)
    ids = np.arange(n, dtype=np.int32) if assignment is None else np.asarray(assignment, np.int32)
    crops = crop_and_resize_3d(m, boxes, ids, mask_shape, "trilinear", 0.0)
    return np.rint(crops[..., 0]).astype(np.float32)                       # tf.round = half to even


def pack_f16(x):
    """core/models.py:3613: ``ra.astype(np.float16)``."""
    with np.errstate(over="ignore"):
        return np.asarray(x, np.float32).astype(np.float16)


def pack_bits(x):
    """core/models.py:3585-3595 ``_bitpack``: ``(packed uint8, shape int32)``."""
    a = np.asarray(x)
    if a.dtype != np.uint8:
        a = (a > 0.5).astype(np.uint8)
    return np.packbits(a.reshape(-1)), np.array(a.shape, dtype=np.int32)


def unpack_bits(bits, shape):
    """The reader side: ``np.unpackbits(bits)[:prod(shape)].reshape(shape)`` as float32."""
    n = int(np.prod(shape))
    return np.unpackbits(np.asarray(bits, np.uint8))[:n].reshape(shape).astype(np.float32)
