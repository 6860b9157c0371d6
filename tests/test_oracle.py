"""CPU tests of the oracle (oracle/roi3d_oracle.c): known answers, properties, golden vectors,
and the pin against the reference's own IOU<float> machine code (when the wheel is present)."""
import glob
import os

import numpy as np
import pytest

import oracle
import roi3d_synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- IoU / NMS known answers ------------------------------------------------------------
def test_iou_known_answers():
    b = np.array([[0, 0, 0, 1, 1, 1],
                  [0, 0, 0, 1, 1, 0.5],          # half of box 0
                  [2, 2, 2, 3, 3, 3],            # disjoint
                  [1, 1, 1, 0, 0, 0],            # box 0 with swapped corners
                  [0.5, 0.5, 0.5, 0.5, 1, 1]],   # zero volume
                 np.float32)
    assert oracle.iou3d(b, 0, 0) == 1.0
    assert oracle.iou3d(b, 0, 1) == 0.5
    assert oracle.iou3d(b, 0, 2) == 0.0
    assert oracle.iou3d(b, 0, 3) == 1.0
    assert oracle.iou3d(b, 0, 4) == 0.0 and oracle.iou3d(b, 4, 4) == 0.0
    m = oracle.iou_matrix(b)
    assert np.array_equal(m, m.T)


def test_nms_known_answers():
    boxes = np.array([[0, 0, 0, 1, 1, 1], [0, 0, 0, 1, 1, 0.9], [0, 0, 0, 1, 1, 0.5], [2, 2, 2, 3, 3, 3]], np.float32)
    scores = np.array([0.9, 0.8, 0.7, 0.1], np.float32)
    assert oracle.non_max_suppression_3d(boxes, scores, 10, 0.6).tolist() == [0, 2, 3]   # iou(0,1)=.9, iou(0,2)=.5
    assert oracle.non_max_suppression_3d(boxes, scores, 10, 0.5).tolist() == [0, 3]      # >= threshold suppresses
    assert oracle.non_max_suppression_3d(boxes, scores, 2, 0.6).tolist() == [0, 2]       # max_output_size
    assert oracle.non_max_suppression_3d(boxes, scores, 0, 0.6).tolist() == []
    assert oracle.non_max_suppression_3d(boxes, scores, 10, 0.0).tolist() == [0]         # iou >= 0 always
    assert oracle.non_max_suppression_3d(boxes, scores, 10, 1.0).tolist() == [0, 1, 2, 3]


def test_nms_ties_lower_index_first():
    boxes = np.array([[0, 0, 0, 1, 1, 1], [5, 5, 5, 6, 6, 6], [0, 0, 0, 1, 1, 1]], np.float32)
    scores = np.array([0.5, 0.5, 0.5], np.float32)
    assert oracle.non_max_suppression_3d(boxes, scores, 3, 0.5).tolist() == [0, 1]
    scores = np.array([0.0, -0.0, 0.0], np.float32)      # -0.0 == +0.0
    assert oracle.non_max_suppression_3d(boxes, scores, 3, 0.5).tolist() == [0, 1]


def test_nms_zero_volume_quirk_and_invalid_scores():
    # TF r2.0-2.2 re-push quirk (NMS.so@0xdc55): a selected zero-volume box is selected again until max_out
    boxes = np.array([[0, 0, 0, 1, 1, 1], [0.2, 0.2, 0.2, 0.2, 0.5, 0.5], [3, 3, 3, 4, 4, 4]], np.float32)
    scores = np.array([0.9, 0.8, 0.7], np.float32)
    assert oracle.non_max_suppression_3d(boxes, scores, 5, 0.5).tolist() == [0, 1, 1, 1, 1]
    assert oracle.non_max_suppression_3d(boxes, scores, 5, 0.0).tolist() == [0]
    # candidates need score > -FLT_MAX (NMS.so@0xd64a): -inf and NaN never enter the queue
    scores = np.array([-np.inf, np.nan, 0.1], np.float32)
    assert oracle.non_max_suppression_3d(boxes, scores, 5, 0.5).tolist() == [2]


def _numpy_nms(boxes, scores, max_out, thr):
    order = np.lexsort((np.arange(len(scores)), -scores.astype(np.float64)))
    iou = oracle.iou_matrix(boxes)
    keep = []
    for i in order:
        if len(keep) >= max_out:
            break
        if all(iou[i, j] < thr for j in keep):
            keep.append(i)
    return keep


@pytest.mark.parametrize("n,thr,max_out", [(200, 0.3, 50), (500, 0.7, 500), (350, 0.5, 20)])
def test_nms_equals_sorted_greedy_form(n, thr, max_out):
    """For volumes > 0 the queue algorithm equals stable-sort + greedy keep."""
    boxes, scores = roi3d_synth.nms_boxes(n, (64, 64, 64), seed=n)
    got = oracle.non_max_suppression_3d(boxes, scores, max_out, thr).tolist()
    assert got == _numpy_nms(boxes, scores, max_out, thr)


# ---- crop-and-resize known answers --------------------------------------------------------
def test_car_identity_crop():
    rng = np.random.default_rng(0)
    img = rng.standard_normal((1, 4, 5, 6, 3), dtype=np.float32)
    box = np.array([[0, 0, 0, 1, 1, 1]], np.float32)
    out = oracle.crop_and_resize_3d(img, box, [0], (4, 5, 6))
    assert np.array_equal(out[0], img[0])
    # reversed box flips every axis
    out = oracle.crop_and_resize_3d(img, box[:, [3, 4, 5, 0, 1, 2]], [0], (4, 5, 6))
    assert np.array_equal(out[0], img[0, ::-1, ::-1, ::-1])


def test_car_extrapolation_and_crop1():
    img = np.arange(2 * 3 * 3 * 3, dtype=np.float32).reshape(2, 3, 3, 3, 1)
    box = np.array([[-0.5, 0, 0, 1.5, 1, 1]], np.float32)       # y from -1 to 3 over 5 samples: -1,0,1,2,3
    out = oracle.crop_and_resize_3d(img, box, [1], (5, 3, 3), extrapolation_value=-7.0)
    assert np.all(out[0, 0] == -7.0) and np.all(out[0, 4] == -7.0)
    assert np.array_equal(out[0, 1:4], img[1])
    # crop == 1 samples the box centre in double (CAR.so@0x533a)
    out = oracle.crop_and_resize_3d(img, np.array([[0, 0, 0, 1, 1, 1]], np.float32), [0], (1, 1, 1))
    assert out.shape == (1, 1, 1, 1, 1) and out[0, 0, 0, 0, 0] == img[0, 1, 1, 1, 0]
    # nearest rounds half away from zero (roundf)
    out = oracle.crop_and_resize_3d(img, np.array([[0.25, 0, 0, 0.25, 0, 0]], np.float32), [0], (1, 1, 1), "nearest")
    assert out[0, 0, 0, 0, 0] == img[0, 1, 0, 0, 0]             # in_y = 0.5 -> 1
    assert oracle.crop_and_resize_3d(img, np.zeros((0, 6), np.float32), np.zeros(0, np.int32), (2, 2, 2)).shape == (0, 2, 2, 2, 1)


def test_car_matches_grid_sample():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(3)
    B, H, W, D, C = 2, 7, 8, 9, 4
    img = rng.standard_normal((B, H, W, D, C), dtype=np.float32)
    n, crop = 6, (3, 4, 5)
    lo = rng.uniform(0, 0.5, (n, 3))
    boxes = np.concatenate([lo, lo + rng.uniform(0.1, 0.5, (n, 3))], 1).astype(np.float32)
    bi = rng.integers(0, B, n).astype(np.int32)
    out = oracle.crop_and_resize_3d(img, boxes, bi, crop)
    t = torch.from_numpy(img).permute(0, 4, 1, 2, 3)
    for k in range(n):
        axes = [np.linspace(boxes[k, a], boxes[k, a + 3], crop[a]) for a in range(3)]
        g = np.stack(np.meshgrid(*axes, indexing="ij"), -1) * 2 - 1
        grid = torch.from_numpy(g[..., [2, 1, 0]].astype(np.float32))[None]
        ref = torch.nn.functional.grid_sample(t[bi[k]:bi[k] + 1], grid, mode="bilinear", align_corners=True)
        assert np.abs(ref[0].permute(1, 2, 3, 0).numpy() - out[k]).max() < 1e-5


def test_grad_image_is_adjoint_of_forward():
    rng = np.random.default_rng(5)
    B, H, W, D, C = 2, 5, 6, 7, 3
    img = rng.standard_normal((B, H, W, D, C), dtype=np.float32)
    boxes = roi3d_synth.rois(9, (20, 24, 7), 5, side_px=(3, 20))
    boxes[0] = [-0.2, 0.1, 0.1, 0.7, 1.3, 0.9]
    bi = rng.integers(0, B, 9).astype(np.int32)
    for method in ("trilinear", "nearest"):
        out = oracle.crop_and_resize_3d(img, boxes, bi, (3, 2, 4), method, 0.0)
        g = rng.standard_normal(out.shape, dtype=np.float32)
        gi = oracle.crop_and_resize_3d_grad_image(g, boxes, bi, img.shape, method)
        lhs = float((out.astype(np.float64) * g).sum())
        rhs = float((img.astype(np.float64) * gi).sum())
        assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1.0)


def test_grad_boxes_matches_finite_differences():
    rng = np.random.default_rng(6)
    B, H, W, D, C = 1, 9, 9, 9, 2
    img = rng.standard_normal((B, H, W, D, C), dtype=np.float32)
    # (coordinates chosen off the integer grid: at a kink the op returns a one-sided derivative; z1 == y1 with
    # H == D and ph == pd, the one configuration in which the reference's depth-step slip -- (z2 - y1) * ratio_h,
    # GB.so@0x4059 -- coincides with the true step, so all six columns are real derivatives)
    boxes = np.array([[0.11, 0.23, 0.11, 0.71, 0.83, 0.67], [0.31, 0.13, 0.31, 0.93, 0.64, 0.83]], np.float32)
    bi = np.zeros(2, np.int32)
    crop = (3, 3, 3)
    g = rng.standard_normal((2,) + crop + (C,), dtype=np.float32)
    gb = oracle.crop_and_resize_3d_grad_boxes(g, img, boxes, bi)
    eps = 2e-4
    bad = 0
    for n in range(2):
        for k in range(6):
            bp, bm = boxes.copy(), boxes.copy()
            bp[n, k] += eps
            bm[n, k] -= eps
            fp = (oracle.crop_and_resize_3d(img, bp, bi, crop).astype(np.float64) * g).sum()
            fm = (oracle.crop_and_resize_3d(img, bm, bi, crop).astype(np.float64) * g).sum()
            bad += abs((fp - fm) / (2 * eps) - gb[n, k]) > 5e-2 * max(1.0, abs(gb[n, k]))
    assert bad <= 1          # the crop is piecewise linear: a difference may straddle one kink


def test_openmp_flavour_is_bit_identical():
    rng = np.random.default_rng(7)
    img = rng.standard_normal((2, 6, 6, 8, 16), dtype=np.float32)
    boxes = roi3d_synth.rois(12, (24, 24, 8), 7, side_px=(3, 20))
    bi = rng.integers(0, 2, 12).astype(np.int32)
    a = oracle.crop_and_resize_3d(img, boxes, bi, (4, 4, 4), threads=1)
    b = oracle.crop_and_resize_3d(img, boxes, bi, (4, 4, 4), threads=4)
    assert np.array_equal(a, b)
    g = rng.standard_normal(a.shape, dtype=np.float32)
    ga = oracle.crop_and_resize_3d_grad_image(g, boxes, bi, img.shape, threads=1)
    gb = oracle.crop_and_resize_3d_grad_image(g, boxes, bi, img.shape, threads=4)
    assert np.array_equal(ga, gb)


# ---- golden vectors -------------------------------------------------------------------------
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "car_*.npz"))))
def test_oracle_reproduces_golden_car(path):
    z = np.load(path)
    crop = tuple(int(v) for v in z["crop"])
    for method in ("trilinear", "nearest"):
        fwd = oracle.crop_and_resize_3d(z["image"], z["boxes"], z["box_index"], crop, method, 0.25)
        assert np.array_equal(fwd, z["fwd_" + method])
        gi = oracle.crop_and_resize_3d_grad_image(z["grads"], z["boxes"], z["box_index"], z["image"].shape, method)
        assert np.array_equal(gi, z["gi_" + method])
    gb = oracle.crop_and_resize_3d_grad_boxes(z["grads"], z["image"], z["boxes"], z["box_index"])
    assert np.array_equal(gb, z["gb"])


def test_oracle_reproduces_golden_nms():
    z = np.load(os.path.join(GOLDEN, "nms.npz"))
    for n in (300, 1000, 64):
        mo, thr = z["args_%d" % n]
        keep = oracle.non_max_suppression_3d(z["boxes_%d" % n], z["scores_%d" % n], int(mo), float(thr))
        assert np.array_equal(keep, z["keep_%d" % n])
    assert len(set(z["keep_64"].tolist())) < len(z["keep_64"])      # the quirk case really repeats an index


# ---- synthetic generator sanity ------------------------------------------------------------------
def test_synth_contract():
    boxes, scores = roi3d_synth.nms_boxes(6000, (128, 128, 128))
    assert boxes.shape == (6000, 6) and boxes.dtype == np.float32 and scores.dtype == np.float32
    assert (boxes >= 0).all() and (boxes <= 1).all() and (boxes[:, 3:] > boxes[:, :3]).all()
    assert len(np.unique(scores)) < 6000                          # tied scores exist
    keep = oracle.non_max_suppression_3d(boxes, scores, 1000, 0.7)
    assert len(keep) == 1000                                      # cfg1 reaches its early exit
    lv = roi3d_synth.roi_levels(roi3d_synth.rois(1000, (128, 128, 128), 2001), (128, 128, 128))
    assert lv.min() >= 2 and lv.max() <= 5
    b = roi3d_synth.rois(4, (128, 128, 128), 1)
    assert roi3d_synth.car_algorithmic_bytes(b, (1, 32, 32, 128, 256), (7, 7, 7)) > 4 * 343 * 1024


# ---------------------------------------------------------------------------------------------
# rows f3 / f4: the numpy restatements of DetectionLayer, mask targets and the target files' payloads
# ---------------------------------------------------------------------------------------------
def test_refine_detections_restatement_properties():
    rng = np.random.default_rng(11)
    R = 400
    rois = roi3d_synth.nms_boxes(R, (128, 128, 64), seed=5)[0]
    probs = rng.random((R, 2)).astype(np.float32)
    deltas = (rng.standard_normal((R, 2, 6)) * 0.5).astype(np.float32)
    shape = (128.0, 128.0, 64.0)
    det = oracle.refine_detections(rois, probs, deltas, shape, 0.6, 0.3, max_instances=50)
    k = int((det[:, 6] > 0).sum())
    assert det.shape == (50, 8) and 0 < k <= 50 and not det[k:].any()
    assert np.all(det[:k, 7] >= 0.6) and np.all(np.diff(det[:k, 7]) <= 0)
    assert np.all((det[:k, :6] >= 0) & (det[:k, :6] <= 1))
    # the kept boxes are mutually below the threshold under the op's own IoU, in pixel space
    kept_px = det[:k, :6] * np.array(shape + shape, np.float32)
    for i in range(min(k, 20)):
        for j in range(i):
            assert oracle.iou3d(kept_px, i, j) < 0.3 + 1e-4
    # zero deltas, threshold 1: every ROI that passes the size filter comes back, as itself, in score order
    px0, _, ok0 = oracle.refine_decode(rois, probs, np.zeros_like(deltas), shape, 0.0)
    det0 = oracle.refine_detections(rois, probs, np.zeros_like(deltas), shape, 0.0, 1.0, max_instances=R)
    order = np.argsort(-probs[:, 1], kind="stable")
    order = order[ok0[order]]
    assert np.array_equal(det0[: len(order), 7], probs[order, 1])
    assert np.allclose(det0[: len(order), :6], np.clip(rois[order], 0, 1), atol=1e-6)
    # nothing passes the confidence filter -> all zeros (the graph's _empty branch)
    assert not oracle.refine_detections(rois, probs, deltas, shape, 2.0, 0.3, max_instances=7).any()


def test_mask_targets_and_payload_restatements():
    rng = np.random.default_rng(12)
    masks = (rng.random((3, 8, 8, 8)) > 0.5)
    boxes = np.array([[0, 0, 0, 1, 1, 1], [0.25, 0.25, 0.25, 0.75, 0.75, 0.75]], np.float32)
    out = oracle.mask_targets(masks, boxes, np.array([2, 0], np.int32), (8, 8, 8))
    assert np.array_equal(out[0], masks[2].astype(np.float32))               # identity crop of a binary mask
    assert set(np.unique(out)) <= {0.0, 1.0}
    bits, shape = oracle.pack_bits(out)
    assert bits.dtype == np.uint8 and len(bits) == out.size // 8 and tuple(shape) == out.shape
    assert np.array_equal(oracle.unpack_bits(bits, shape), out)
    x = np.array([0.5, 0.50001, 1.0, 0.0, 7.0, -1.0, 0.4, 0.6, 0.9], np.float32)
    b, _ = oracle.pack_bits(x)
    assert list(b) == [0b01101001, 0b10000000]                               # MSB first, zero padded
    assert oracle.pack_f16(np.array([65520.0, 1e-8, 1.0009765625], np.float32)).tolist() == [np.inf, 0.0, 1.0009765625]


def test_restated_tf_image_nms_2d_known_answers():
    """oracle.non_max_suppression_2d_tf restates tf.image.non_max_suppression (the NMS of this fork's DetectionLayer,
    core/models.py:1496-1501): strict '>' at the threshold, ties -> lower index, empty boxes never suppress."""
    b = np.array([[0, 0, 64, 64], [0, 0, 32, 64], [100, 100, 110, 110], [0, 0, 64, 64]], np.float32)
    s = np.array([0.9, 0.8, 0.7, 0.9], np.float32)
    assert oracle.non_max_suppression_2d_tf(b, s, 10, 0.5).tolist() == [0, 1, 2]      # IoU(0,1) == 0.5 is not > 0.5; box 3 duplicates 0
    assert oracle.non_max_suppression_2d_tf(b, s, 10, 0.49).tolist() == [0, 2]
    assert oracle.non_max_suppression_2d_tf(b, s, 1, 0.5).tolist() == [0]
    z = np.array([[5, 5, 5, 9], [5, 5, 5, 9]], np.float32)                             # zero-area boxes: IoU 0 with everything
    assert oracle.non_max_suppression_2d_tf(z, np.array([0.5, 0.4], np.float32), 5, 0.0).tolist() == [0, 1]
    # agreement with the 3-D op on boxes that share one z interval, away from threshold ties
    boxes, scores = roi3d_synth.nms_boxes(400, (64, 64, 64), seed=31)
    boxes[:, 2], boxes[:, 5] = 0.0, 1.0
    ref3 = oracle.non_max_suppression_3d(boxes, scores, 400, 0.4)
    ref2 = oracle.non_max_suppression_2d_tf(boxes[:, [1, 0, 4, 3]], scores, 400, 0.4)
    assert np.array_equal(ref2, ref3)
