"""GPU parity tests: the sm_100a kernels (called through the C ABI via the host-side mirror of
core/custom_op) against the CPU oracle and the committed golden vectors.

Tolerances (BASELINE.json north_star):
  NMS3D kept indices + order ............ bit-exact
  CropAndResize3D forward ............... <= 1e-5 relative  (and, in fact, bit-exact: asserted)
  CropAndResize3DGradImage .............. <= 1e-4 relative  (atomic re-ordering)
"relative" = |a-b| <= tol * |b| + tol * max|b| elementwise (allclose with rtol = tol, atol = tol * |b|_inf,
SURVEY.md section 8c: a pure elementwise ratio is meaningless where sums cancel).
"""
import glob
import os

import numpy as np
import pytest

import oracle
import roi3d_synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FWD_TOL, BWD_TOL, GB_TOL = 1e-5, 1e-4, 2e-3


def rel_ok(a, b, tol):
    b = np.asarray(b, np.float64)
    a = np.asarray(a, np.float64)
    bmax = np.abs(b).max() if b.size else 0.0
    return bool(np.all(np.abs(a - b) <= tol * np.abs(b) + tol * bmax + 1e-30))


def dev(x, cuda_device):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda_device)


@pytest.fixture(autouse=True)
def _reset_variants(rb):
    yield
    rb.custom_op.set_option("car_fwd_variant", 0)
    rb.custom_op.set_option("car_bwd_variant", 0)


# =============================================================================================
# NMS3D
# =============================================================================================
def run_nms(rb, cuda_device, boxes, scores, max_out, thr):
    keep = rb.non_max_suppression_3d(dev(boxes, cuda_device), dev(scores, cuda_device), max_out, thr)
    assert keep.dtype.is_floating_point is False and keep.dim() == 1
    return keep.cpu().numpy()


@pytest.mark.parametrize("n,thr,max_out,presorted", [
    (1, 0.5, 10, False), (2, 0.5, 1, False), (31, 0.5, 31, False), (32, 0.3, 100, False), (33, 0.7, 5, True),
    (257, 0.5, 64, False), (1000, 0.3, 1000, False), (1000, 0.7, 100, True), (4097, 0.5, 300, False),
    (6000, 0.7, 1000, True), (6000, 0.7, 1000, False), (6000, 0.3, 6000, False), (20000, 0.7, 2000, False),
])
def test_nms_matches_oracle(rb, cuda_device, n, thr, max_out, presorted):
    boxes, scores = roi3d_synth.nms_boxes(n, (128, 128, 128), seed=900 + n, presorted=presorted)
    got = run_nms(rb, cuda_device, boxes, scores, max_out, thr)
    ref = oracle.non_max_suppression_3d(boxes, scores, max_out, thr)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("n,max_out,thr", [(6000, 1000, 0.7), (20000, 2000, 0.7), (20000, 2000, 0.5), (9000, 300, 0.7), (33000, 4000, 0.6)])
def test_nms_dense_clusters_deep_scan(rb, cuda_device, n, max_out, thr):
    """Dense proposals (clusters of 64 near-duplicates, what a trained RPN emits): the scan goes several times deeper than
    max_out, through the windowed tail phases (>= 8192 boxes).  Every schedule must give the oracle's indices."""
    boxes, scores = roi3d_synth.nms_boxes(n, (256, 256, 256), seed=4100 + n, cluster=64, jitter=0.05)
    ref = oracle.non_max_suppression_3d(boxes, scores, max_out, thr)
    try:
        for variant in (0, 1, 2):                           # windowed tail, single phase, head + one tail
            rb.custom_op.set_option("nms_variant", variant)
            assert np.array_equal(run_nms(rb, cuda_device, boxes, scores, max_out, thr), ref), variant
    finally:
        rb.custom_op.set_option("nms_variant", 0)


def test_nms_golden(rb, cuda_device):
    z = np.load(os.path.join(GOLDEN, "nms.npz"))
    for n in (300, 1000, 64):
        mo, thr = z["args_%d" % n]
        got = run_nms(rb, cuda_device, z["boxes_%d" % n], z["scores_%d" % n], int(mo), float(thr))
        assert np.array_equal(got, z["keep_%d" % n])


def test_nms_edge_cases(rb, cuda_device):
    boxes = np.array([[0, 0, 0, 1, 1, 1], [0, 0, 0, 1, 1, 0.9], [0, 0, 0, 1, 1, 0.5], [2, 2, 2, 3, 3, 3]], np.float32)
    scores = np.array([0.9, 0.8, 0.7, 0.1], np.float32)
    for thr, mo in ((0.6, 10), (0.5, 10), (0.6, 2), (0.0, 10), (1.0, 10), (0.6, 0)):
        assert run_nms(rb, cuda_device, boxes, scores, mo, thr).tolist() == \
            oracle.non_max_suppression_3d(boxes, scores, mo, thr).tolist()
    # empty input
    assert run_nms(rb, cuda_device, np.zeros((0, 6), np.float32), np.zeros(0, np.float32), 5, 0.5).tolist() == []
    # ties (incl. -0.0 == +0.0) -> lower index first
    b3 = np.array([[0, 0, 0, 1, 1, 1], [5, 5, 5, 6, 6, 6], [0, 0, 0, 1, 1, 1]], np.float32)
    assert run_nms(rb, cuda_device, b3, np.array([0.5, 0.5, 0.5], np.float32), 3, 0.5).tolist() == [0, 1]
    assert run_nms(rb, cuda_device, b3, np.array([0.0, -0.0, 0.0], np.float32), 3, 0.5).tolist() == [0, 1]
    # scores <= -FLT_MAX / NaN are never candidates
    s = np.array([-np.inf, np.nan, 0.1], np.float32)
    assert run_nms(rb, cuda_device, b3, s, 5, 0.5).tolist() == oracle.non_max_suppression_3d(b3, s, 5, 0.5).tolist() == [2]
    # negative scores, reversed corners
    bx, sc = roi3d_synth.nms_boxes(500, (64, 64, 64), seed=77)
    bx[::3] = bx[::3][:, [3, 4, 5, 0, 1, 2]]
    sc = sc - 0.5
    assert np.array_equal(run_nms(rb, cuda_device, bx, sc, 200, 0.4), oracle.non_max_suppression_3d(bx, sc, 200, 0.4))


def test_nms_zero_volume_quirk(rb, cuda_device):
    boxes = np.array([[0, 0, 0, 1, 1, 1], [0.2, 0.2, 0.2, 0.2, 0.5, 0.5], [3, 3, 3, 4, 4, 4]], np.float32)
    scores = np.array([0.9, 0.8, 0.7], np.float32)
    for thr, mo in ((0.5, 5), (0.0, 5), (0.5, 2), (0.5, 1)):
        assert run_nms(rb, cuda_device, boxes, scores, mo, thr).tolist() == \
            oracle.non_max_suppression_3d(boxes, scores, mo, thr).tolist()
    bx, sc = roi3d_synth.nms_boxes(300, (64, 64, 64), seed=78)
    bx[40, 3] = bx[40, 0]                                     # a zero-volume box somewhere in the middle
    assert np.array_equal(run_nms(rb, cuda_device, bx, sc, 250, 0.5), oracle.non_max_suppression_3d(bx, sc, 250, 0.5))


def test_nms_many_duplicates(rb, cuda_device):
    bx, _ = roi3d_synth.nms_boxes(3000, (128, 128, 128), seed=79)
    sc = np.random.default_rng(79).integers(0, 7, 3000).astype(np.float32) / 7.0     # heavy ties
    assert np.array_equal(run_nms(rb, cuda_device, bx, sc, 3000, 0.5), oracle.non_max_suppression_3d(bx, sc, 3000, 0.5))


def test_nms_large_properties(rb, cuda_device):
    """100k boxes (BASELINE sweep max): properties that do not need the oracle."""
    n, thr = 100000, 0.5
    boxes, scores = roi3d_synth.nms_boxes(n, (256, 256, 256), seed=80)
    keep = run_nms(rb, cuda_device, boxes, scores, 2000, thr)
    assert len(keep) == 2000 and len(set(keep.tolist())) == 2000
    ks = scores[keep]
    assert np.all(ks[:-1] >= ks[1:])                            # selection order = score order
    iou = oracle.iou_matrix(boxes[keep])
    np.fill_diagonal(iou, 0.0)
    assert iou.max() < thr                                      # kept set is conflict free
    again = run_nms(rb, cuda_device, boxes[keep], scores[keep], 2000, thr)
    assert np.array_equal(again, np.arange(2000))               # idempotent
    # every dropped box ranked above the last kept one is suppressed by a kept box with higher priority
    order = np.lexsort((np.arange(n), -scores.astype(np.float64)))
    rank = np.empty(n, np.int64)
    rank[order] = np.arange(n)
    cut = rank[keep[-1]]
    dropped = order[:cut][~np.isin(order[:cut], keep)]
    sample = dropped[:: max(1, len(dropped) // 200)]
    both = np.concatenate([boxes[keep], boxes[sample]])
    m = oracle.iou_matrix(both)[2000:, :2000]
    for r, d in enumerate(sample):
        assert np.any((m[r] >= thr) & (rank[keep] < rank[d]))


# =============================================================================================
# CropAndResize3D
# =============================================================================================
def car_inputs(seed, B, H, W, D, C, n, crop, wild=True):
    rng = np.random.default_rng(seed)
    image = rng.standard_normal((B, H, W, D, C), dtype=np.float32)
    boxes = roi3d_synth.rois(n, (H * 4, W * 4, D), seed, side_px=(4.0, 3.0 * max(H, W)))
    if wild and n >= 6:
        boxes[0] = [-0.2, 0.1, 0.1, 0.7, 1.3, 0.9]            # partly outside -> extrapolation
        boxes[1] = [0.8, 0.7, 0.9, 0.2, 0.1, 0.3]             # reversed corners
        boxes[2] = [0.5, 0.5, 0.5, 0.5, 0.5, 0.5]             # zero size
        boxes[3] = [0.0, 0.0, 0.0, 1.0, 1.0, 1.0]             # whole volume, integer coordinates
        boxes[4] = [1.2, 1.2, 1.2, 1.5, 1.5, 1.5]             # fully outside
        boxes[5] = [0.25, 0.25, 0.25, 0.75, 0.75, 0.75]
    box_index = rng.integers(0, B, n).astype(np.int32)
    grads = rng.standard_normal((n,) + tuple(crop) + (C,), dtype=np.float32)
    return image, boxes, box_index, grads


CAR_CASES = [
    # B, H, W, D, C, n, crop
    (2, 8, 8, 16, 64, 24, (7, 7, 7)),
    (2, 8, 8, 16, 256, 12, (14, 14, 14)),
    (1, 16, 16, 32, 32, 16, (7, 7, 7)),
    (2, 6, 7, 9, 8, 10, (3, 4, 5)),
    (2, 5, 4, 6, 12, 8, (1, 2, 1)),
    (1, 4, 4, 8, 128, 8, (28, 28, 28)),
    (3, 9, 5, 7, 1, 9, (5, 3, 4)),                            # C = 1 (mask targets, core/models.py:992)
    (2, 7, 6, 5, 3, 9, (2, 2, 2)),                            # C % 4 != 0
    (1, 32, 32, 16, 64, 6, (14, 14, 14)),                     # large footprints -> several y-tiles
    (1, 2, 2, 2, 4, 7, (64, 3, 2)),
]


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5])            # direct, plane-staged, fed by TMA bulk copies, row-walk, fed by TMA gather4
@pytest.mark.parametrize("case", CAR_CASES)
def test_car_forward_matches_oracle(rb, cuda_device, case, variant):
    B, H, W, D, C, n, crop = case
    image, boxes, bidx, _ = car_inputs(100 + C + n, B, H, W, D, C, n, crop)
    rb.custom_op.set_option("car_fwd_variant", variant)
    out = rb.crop_and_resize_3d(dev(image, cuda_device), dev(boxes, cuda_device), dev(bidx, cuda_device), crop,
                                extrapolation_value=0.5).cpu().numpy()
    ref = oracle.crop_and_resize_3d(image, boxes, bidx, crop, "trilinear", 0.5)
    assert out.shape == ref.shape
    assert rel_ok(out, ref, FWD_TOL)
    assert np.array_equal(out, ref)                             # same operations, same order: bit-exact


@pytest.mark.parametrize("case", CAR_CASES[:5] + CAR_CASES[6:8])
def test_car_forward_nearest(rb, cuda_device, case):
    B, H, W, D, C, n, crop = case
    image, boxes, bidx, _ = car_inputs(200 + C + n, B, H, W, D, C, n, crop)
    out = rb.crop_and_resize_3d(dev(image, cuda_device), dev(boxes, cuda_device), dev(bidx, cuda_device), crop,
                                method_name="nearest", extrapolation_value=-1.0).cpu().numpy()
    assert np.array_equal(out, oracle.crop_and_resize_3d(image, boxes, bidx, crop, "nearest", -1.0))


@pytest.mark.parametrize("variant", [1, 2, 3, 4])               # direct scatter, plane-staged scatter, output-stationary, TMA-staged scatter
@pytest.mark.parametrize("case", CAR_CASES)
def test_car_grad_image_matches_oracle(rb, cuda_device, case, variant):
    B, H, W, D, C, n, crop = case
    image, boxes, bidx, grads = car_inputs(300 + C + n, B, H, W, D, C, n, crop)
    rb.custom_op.set_option("car_bwd_variant", variant)
    out = rb.crop_and_resize_3d_grad_image(dev(grads, cuda_device), dev(boxes, cuda_device), dev(bidx, cuda_device),
                                           image.shape).cpu().numpy()
    ref = oracle.crop_and_resize_3d_grad_image(grads, boxes, bidx, image.shape)
    assert out.shape == ref.shape
    assert rel_ok(out, ref, BWD_TOL)


@pytest.mark.parametrize("case", CAR_CASES[:5] + CAR_CASES[6:8])
def test_car_grad_image_nearest(rb, cuda_device, case):
    B, H, W, D, C, n, crop = case
    image, boxes, bidx, grads = car_inputs(400 + C + n, B, H, W, D, C, n, crop)
    out = rb.crop_and_resize_3d_grad_image(dev(grads, cuda_device), dev(boxes, cuda_device), dev(bidx, cuda_device),
                                           image.shape, method_name="nearest").cpu().numpy()
    assert rel_ok(out, oracle.crop_and_resize_3d_grad_image(grads, boxes, bidx, image.shape, "nearest"), BWD_TOL)


@pytest.mark.parametrize("case", CAR_CASES[:5] + CAR_CASES[6:8])
def test_car_grad_boxes_matches_oracle(rb, cuda_device, case):
    B, H, W, D, C, n, crop = case
    image, boxes, bidx, grads = car_inputs(500 + C + n, B, H, W, D, C, n, crop)
    out = rb.crop_and_resize_3d_grad_boxes(dev(grads, cuda_device), dev(image, cuda_device), dev(boxes, cuda_device),
                                           dev(bidx, cuda_device)).cpu().numpy()
    ref = oracle.crop_and_resize_3d_grad_boxes(grads, image, boxes, bidx)
    # the reference accumulates ~1e5 fp32 terms sequentially; compare at the accuracy of that sum
    assert np.all(np.abs(out - ref) <= GB_TOL * np.abs(ref).max() + 1e-6)


def _fuzz_case(seed):
    """Random geometry: ragged volumes, crops up to 20 per axis (non-cubic), any channel count (C % 4 != 0 included),
    boxes partly / fully outside, reversed, degenerate, on integer coordinates; random extrapolation value."""
    rng = np.random.default_rng(seed)
    B = int(rng.integers(1, 4))
    H, W, D = (int(rng.integers(1, 20)) for _ in range(3))
    C = int(rng.choice([1, 3, 4, 8, 20, 36, 64, 68, 132, 256]))
    crop = tuple(int(rng.integers(1, 21)) for _ in range(3))
    n = int(rng.integers(1, 14))
    image = rng.standard_normal((B, H, W, D, C), dtype=np.float32)
    c = rng.uniform(-0.2, 1.2, (n, 3))
    side = rng.uniform(0.0, 1.3, (n, 3)) * rng.choice([1.0, 1.0, 1.0, -1.0], (n, 3))      # some corners reversed
    boxes = np.concatenate([c - side / 2, c + side / 2], axis=1).astype(np.float32)
    for i in range(0, n, 5):                                                              # integer-coordinate boxes
        boxes[i] = np.array([0, 0, 0, 1, 1, 1], np.float32) * rng.choice([0.5, 1.0])
    bidx = rng.integers(0, B, n).astype(np.int32)
    grads = rng.standard_normal((n,) + crop + (C,), dtype=np.float32)
    return image, boxes, bidx, grads, crop, float(rng.choice([0.0, -3.5, 0.25]))


@pytest.mark.parametrize("seed", range(40))
def test_car_fuzz_all_variants(rb, cuda_device, seed):
    """Every forward variant bit-exact and every backward variant within 1e-4 of the oracle on random geometry."""
    image, boxes, bidx, grads, crop, ext = _fuzz_case(7000 + seed)
    ref = oracle.crop_and_resize_3d(image, boxes, bidx, crop, "trilinear", ext)
    gref = oracle.crop_and_resize_3d_grad_image(grads, boxes, bidx, image.shape)
    t = [dev(x, cuda_device) for x in (image, boxes, bidx, grads)]
    for fv in (0, 1, 2, 3, 4, 5):
        rb.custom_op.set_option("car_fwd_variant", fv)
        out = rb.crop_and_resize_3d(t[0], t[1], t[2], crop, extrapolation_value=ext).cpu().numpy()
        assert np.array_equal(out, ref), ("forward variant", fv, image.shape, crop)
    for bv in (0, 1, 2, 3, 4):
        rb.custom_op.set_option("car_bwd_variant", bv)
        gi = rb.crop_and_resize_3d_grad_image(t[3], t[1], t[2], image.shape).cpu().numpy()
        assert rel_ok(gi, gref, BWD_TOL), ("backward variant", bv, image.shape, crop)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "car_*.npz"))))
def test_car_golden(rb, cuda_device, path):
    z = np.load(path)
    crop = tuple(int(v) for v in z["crop"])
    image, boxes, bidx, grads = (dev(z[k], cuda_device) for k in ("image", "boxes", "box_index", "grads"))
    for method in ("trilinear", "nearest"):
        out = rb.crop_and_resize_3d(image, boxes, bidx, crop, method_name=method, extrapolation_value=0.25)
        assert np.array_equal(out.cpu().numpy(), z["fwd_" + method])
        gi = rb.crop_and_resize_3d_grad_image(grads, boxes, bidx, z["image"].shape, method_name=method)
        assert rel_ok(gi.cpu().numpy(), z["gi_" + method], BWD_TOL)
    gb = rb.crop_and_resize_3d_grad_boxes(grads, image, boxes, bidx).cpu().numpy()
    assert np.all(np.abs(gb - z["gb"]) <= GB_TOL * np.abs(z["gb"]).max() + 1e-6)


def test_car_empty_and_single(rb, cuda_device):
    import torch
    image = dev(np.ones((1, 4, 4, 4, 8), np.float32), cuda_device)
    e6 = torch.zeros((0, 6), device=cuda_device)
    ei = torch.zeros((0,), dtype=torch.int32, device=cuda_device)
    assert tuple(rb.crop_and_resize_3d(image, e6, ei, (7, 7, 7)).shape) == (0, 7, 7, 7, 8)
    g = torch.zeros((0, 7, 7, 7, 8), device=cuda_device)
    gi = rb.crop_and_resize_3d_grad_image(g, e6, ei, (1, 4, 4, 4, 8))
    assert tuple(gi.shape) == (1, 4, 4, 4, 8) and float(gi.abs().sum()) == 0.0     # zero-filled like GI.so@0x3ec5
    assert tuple(rb.crop_and_resize_3d_grad_boxes(g, image, e6, ei).shape) == (0, 6)
    # host buffers, no boxes: the zero-filled result is produced in host memory (numpy in -> numpy out)
    hz = rb.crop_and_resize_3d_grad_image(np.zeros((0, 7, 7, 7, 8), np.float32), np.zeros((0, 6), np.float32),
                                          np.zeros(0, np.int32), (2, 4, 4, 4, 8))
    assert isinstance(hz, np.ndarray) and hz.shape == (2, 4, 4, 4, 8) and hz.dtype == np.float32 and not hz.any()


def test_roi_processing_order_does_not_change_results(rb, cuda_device):
    """The workspace entry points process the ROIs sorted by (image, y centre): the forward is bit-identical to the plain
    entry point and to the oracle, the backward stays within 1e-4; an undersized workspace falls back to the given order."""
    import ctypes
    import torch
    lib = rb._lib.load()
    vp = ctypes.c_void_p
    B, H, W, D, C, n, crop = 3, 16, 16, 24, 256, 96, (14, 14, 14)     # 96 x 2744 x 256 floats: above the ordering threshold
    image, boxes, bidx, grads = car_inputs(5150, B, H, W, D, C, n, crop)
    bidx = np.random.default_rng(1).integers(0, B, n).astype(np.int32)            # images interleaved
    t_img, t_b, t_i, t_g = (dev(x, cuda_device) for x in (image, boxes, bidx, grads))
    stream = vp(torch.cuda.current_stream().cuda_stream)
    ref = oracle.crop_and_resize_3d(image, boxes, bidx, crop, threads=max(1, oracle.max_threads()))
    gref = oracle.crop_and_resize_3d_grad_image(grads, boxes, bidx, image.shape, threads=max(1, oracle.max_threads()))
    full = int(lib.roi3d_car3d_workspace_bytes(n))
    assert full >= 4 * n and int(lib.roi3d_car3d_workspace_bytes(0)) > 0
    for nbytes in (full, 8, 0):
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=cuda_device)
        out = torch.empty((n,) + crop + (C,), device=cuda_device)
        rb._lib.check(lib.roi3d_car3d_fwd_ws(vp(t_img.data_ptr()), B, H, W, D, C, vp(t_b.data_ptr()), vp(t_i.data_ptr()), n, *crop, 0, 0.0,
                                             vp(out.data_ptr()), vp(ws.data_ptr()) if nbytes else None, nbytes, stream))
        assert np.array_equal(out.cpu().numpy(), ref), nbytes
        gi = torch.empty(image.shape, device=cuda_device)
        rb._lib.check(lib.roi3d_car3d_grad_image_ws(vp(t_g.data_ptr()), vp(t_b.data_ptr()), vp(t_i.data_ptr()), n, *crop, B, H, W, D, C, 0,
                                                    vp(gi.data_ptr()), vp(ws.data_ptr()) if nbytes else None, nbytes, stream))
        assert rel_ok(gi.cpu().numpy(), gref, BWD_TOL), nbytes


@pytest.mark.parametrize("uploads_first", [False, True])
def test_host_buffer_pipeline_policies(rb, cuda_device, uploads_first):
    """Host buffers in, host results out, several independent ops inside one `deferred` block: both copy policies
    (full duplex / uploads first) return the oracle's results once the block has exited."""
    B, H, W, D, C, n, crop = 2, 8, 8, 16, 256, 12, (14, 14, 14)
    image, boxes, bidx, grads = car_inputs(4242, B, H, W, D, C, n, crop)
    big = np.tile(image, (1, 4, 4, 2, 1))                       # > 1 MiB: goes through the upload stream
    try:
        rb.host_pipeline(uploads_first=uploads_first)
        with rb.deferred():
            d_img = rb.upload(big)
            outs = [rb.crop_and_resize_3d(d_img, boxes, bidx, crop) for _ in range(2)]
            gis = [rb.crop_and_resize_3d_grad_image(grads, boxes, bidx, big.shape) for _ in range(2)]
    finally:
        rb.host_pipeline(uploads_first=False)
    ref = oracle.crop_and_resize_3d(big, boxes, bidx, crop)
    gref = oracle.crop_and_resize_3d_grad_image(grads, boxes, bidx, big.shape)
    to_np = lambda t: t if isinstance(t, np.ndarray) else t.numpy()      # noqa: E731  (host results: numpy or pinned CPU tensors)
    for o in outs:
        assert not getattr(o, "is_cuda", False) and np.array_equal(to_np(o), ref)
    for g in gis:
        assert not getattr(g, "is_cuda", False) and rel_ok(to_np(g), gref, BWD_TOL)


@pytest.mark.parametrize("case", [CAR_CASES[0], CAR_CASES[1], CAR_CASES[7]])
def test_car_box_index_out_of_range_is_guarded(rb, cuda_device, case):
    """The reference reads out of bounds for a box_index outside [0, B) (no check in CAR.so / GI.so / GB.so).  Here
    such a box reads nothing: crop = extrapolation value, no scatter, zero grad-boxes row; the other boxes are untouched."""
    B, H, W, D, C, n, crop = case
    image, boxes, bidx, grads = car_inputs(900 + C, B, H, W, D, C, n, crop, wild=False)
    bad = np.zeros(n, bool)
    bad[[1, n // 2, n - 1]] = True
    bidx_bad = bidx.copy()
    bidx_bad[bad] = [-1, B, 2 ** 30]
    good = ~bad
    t = [dev(x, cuda_device) for x in (image, boxes, bidx_bad, grads)]
    for fv in (1, 2, 3, 4, 5):
        rb.custom_op.set_option("car_fwd_variant", fv)
        out = rb.crop_and_resize_3d(t[0], t[1], t[2], crop, extrapolation_value=7.0).cpu().numpy()
        assert np.array_equal(out[good], oracle.crop_and_resize_3d(image, boxes[good], bidx[good], crop, "trilinear", 7.0))
        assert np.all(out[bad] == 7.0)
    ref = oracle.crop_and_resize_3d_grad_image(grads[good], boxes[good], bidx[good], image.shape)
    for bv in (1, 2, 3, 4):
        rb.custom_op.set_option("car_bwd_variant", bv)
        gi = rb.crop_and_resize_3d_grad_image(t[3], t[1], t[2], image.shape).cpu().numpy()
        assert rel_ok(gi, ref, BWD_TOL)
    gb = rb.crop_and_resize_3d_grad_boxes(t[3], t[0], t[1], t[2]).cpu().numpy()
    rgb = oracle.crop_and_resize_3d_grad_boxes(grads[good], image, boxes[good], bidx[good])
    assert np.all(gb[bad] == 0.0)
    assert np.all(np.abs(gb[good] - rgb) <= GB_TOL * np.abs(rgb).max() + 1e-6)


def test_autograd_matches_registered_gradient(rb, cuda_device):
    """backward() == (CropAndResize3DGradImage, CropAndResize3DGradBoxes, None, None) -- the wiring of
    _CropAndResize3DGrad, core/custom_op/custom_op.py:28-65."""
    B, H, W, D, C, n, crop = 2, 8, 8, 12, 16, 10, (5, 5, 5)
    image, boxes, bidx, grads = car_inputs(600, B, H, W, D, C, n, crop)
    t_img = dev(image, cuda_device).requires_grad_(True)
    t_box = dev(boxes, cuda_device).requires_grad_(True)
    out = rb.crop_and_resize_3d(t_img, t_box, dev(bidx, cuda_device), crop)
    out.backward(dev(grads, cuda_device))
    assert rel_ok(t_img.grad.cpu().numpy(), oracle.crop_and_resize_3d_grad_image(grads, boxes, bidx, image.shape), BWD_TOL)
    ref = oracle.crop_and_resize_3d_grad_boxes(grads, image, boxes, bidx)
    assert np.all(np.abs(t_box.grad.cpu().numpy() - ref) <= GB_TOL * np.abs(ref).max())


def test_host_buffers_round_trip(rb, cuda_device):
    """numpy in -> numpy out through pinned H2D/D2H: the end-to-end path bench.py times."""
    B, H, W, D, C, n, crop = 2, 8, 8, 16, 64, 12, (7, 7, 7)
    image, boxes, bidx, grads = car_inputs(700, B, H, W, D, C, n, crop)
    out = rb.crop_and_resize_3d(image, boxes, bidx, crop)
    assert isinstance(out, np.ndarray) and np.array_equal(out, oracle.crop_and_resize_3d(image, boxes, bidx, crop))
    gi = rb.crop_and_resize_3d_grad_image(grads, boxes, bidx, image.shape)
    assert isinstance(gi, np.ndarray) and rel_ok(gi, oracle.crop_and_resize_3d_grad_image(grads, boxes, bidx, image.shape), BWD_TOL)
    bx, sc = roi3d_synth.nms_boxes(800, (64, 64, 64), seed=5)
    keep = rb.non_max_suppression_3d(bx, sc, 100, 0.5)
    assert isinstance(keep, np.ndarray) and np.array_equal(keep, oracle.non_max_suppression_3d(bx, sc, 100, 0.5))


# =============================================================================================
# BASELINE shapes (cfg2: B=2, 128^3, 128 ROIs/image, C=256, P2..P5) -- size-independent properties
# =============================================================================================
@pytest.mark.parametrize("crop", [(7, 7, 7), (14, 14, 14)])
def test_cfg2_full_size_properties(rb, cuda_device, crop):
    import torch
    vol, B = (128, 128, 128), 2
    routed = roi3d_synth.pyramid_rois(128, B, vol, seed=2002)
    assert sum(len(v[0]) for v in routed.values()) == 256
    for level, (boxes, bidx, _) in routed.items():
        if len(boxes) == 0:
            continue
        shape = roi3d_synth.level_shape(vol, level, batch=B)
        torch.manual_seed(1000 + level)
        image = torch.randn(shape, device=cuda_device)
        tb, ti = dev(boxes, cuda_device), dev(bidx, cuda_device)
        rb.custom_op.set_option("car_fwd_variant", 2)
        out2 = rb.crop_and_resize_3d(image, tb, ti, crop)
        rb.custom_op.set_option("car_fwd_variant", 1)
        out1 = rb.crop_and_resize_3d(image, tb, ti, crop)
        assert torch.equal(out1, out2)                          # direct gather == plane-staged, bit for bit
        # a handful of ROIs against the oracle at full feature-map size
        pick = np.arange(0, len(boxes), max(1, len(boxes) // 4))[:4]
        ref = oracle.crop_and_resize_3d(image.cpu().numpy(), boxes[pick], bidx[pick], crop)
        assert np.array_equal(out2[torch.from_numpy(pick).to(cuda_device)].cpu().numpy(), ref)
        # adjoint identity <fwd(image), g> == <image, bwd(g)> ties backward to forward at full size
        g = torch.randn_like(out2)
        for variant in (1, 2, 3):
            rb.custom_op.set_option("car_bwd_variant", variant)
            gi = rb.crop_and_resize_3d_grad_image(g, tb, ti, shape)
            lhs = float((out2.double() * g.double()).sum())
            rhs = float((image.double() * gi.double()).sum())
            assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), float(out2.double().norm() * g.double().norm()) * 1e-3)
        # linearity of the backward: bwd(2g) == 2 bwd(g) up to atomic re-ordering
        gi2 = rb.crop_and_resize_3d_grad_image(2 * g, tb, ti, shape)
        assert torch.allclose(gi2, 2 * gi, rtol=1e-4, atol=1e-4 * float(gi.abs().max()))
        del image, out1, out2, g, gi, gi2
        torch.cuda.empty_cache()


def _full_size_backward_vs_oracle(rb, cuda_device, batch, rois_per_image, crops, seed):
    """Every ROI of a BASELINE-sized step, P2..P5 of a 128^3 volume at C = 256: forward bit-exact against the oracle
    over ALL ROIs, grad-image <= 1e-4 (SURVEY.md 8c rule) for every backward variant, on the inputs bench.py uses."""
    import torch
    vol = (128, 128, 128)
    nt = max(1, oracle.max_threads())      # OpenMP over boxes (fwd) / channels (bwd): per-voxel order unchanged
    routed = roi3d_synth.pyramid_rois(rois_per_image, batch, vol, seed=seed)
    assert sum(len(v[0]) for v in routed.values()) == batch * rois_per_image
    for level, (boxes, bidx, _) in routed.items():
        shape = roi3d_synth.level_shape(vol, level, batch=batch)
        tb, ti = dev(boxes, cuda_device), dev(bidx, cuda_device)
        image = roi3d_synth.feature_map(vol, level, batch=batch)
        t_image = dev(image, cuda_device)
        for crop in crops:
            n = len(boxes)
            if n:
                out = rb.crop_and_resize_3d(t_image, tb, ti, crop).cpu().numpy()
                assert np.array_equal(out, oracle.crop_and_resize_3d(image, boxes, bidx, crop, threads=nt)), (level, crop)
                del out
            grads = roi3d_synth.grads_like((n,) + crop + (shape[4],), 4000 + level * 10 + crop[0])
            ref = oracle.crop_and_resize_3d_grad_image(grads, boxes, bidx, shape, threads=nt)
            tg = dev(grads, cuda_device)
            for variant in (0, 2, 3, 4):
                rb.custom_op.set_option("car_bwd_variant", variant)
                gi = rb.crop_and_resize_3d_grad_image(tg, tb, ti, shape).cpu().numpy()
                assert rel_ok(gi, ref, BWD_TOL), (level, crop, variant)
            if n:                                               # the output-stationary kernel is deterministic
                rb.custom_op.set_option("car_bwd_variant", 3)
                a = rb.crop_and_resize_3d_grad_image(tg, tb, ti, shape)
                b = rb.crop_and_resize_3d_grad_image(tg, tb, ti, shape)
                assert torch.equal(a, b)
                del a, b
            del tg, ref, grads
        del t_image, image
        torch.cuda.empty_cache()


def test_cfg2_full_size_vs_oracle(rb, cuda_device):
    """BASELINE configs[1]: batch 2 x 128 ROIs, 7^3 and 14^3 heads, all 256 ROIs against the oracle (GI.so@0x3a80)."""
    _full_size_backward_vs_oracle(rb, cuda_device, 2, 128, ((7, 7, 7), (14, 14, 14)), seed=2002)


def test_cfg4_full_size_vs_oracle(rb, cuda_device):
    """BASELINE configs[3]: one image, 1000 ROIs x 14^3 x 256 channels, all ROIs against the oracle."""
    _full_size_backward_vs_oracle(rb, cuda_device, 1, 1000, ((14, 14, 14),), seed=2002)


# =============================================================================================
# CUDA path against the reference's OWN binaries (oracle/_ref travels to the GPU box)
# =============================================================================================
def _reference():
    try:
        from oracle import refrun
        return refrun.load()
    except Exception as e:  # noqa: BLE001
        pytest.skip("reference binaries unavailable: %s" % e)


def test_cuda_vs_reference_binaries(rb, cuda_device):
    ref = _reference()
    # NMS3D: bit-exact against NonMaxSuppression3DOp<CPUDevice>::Compute itself
    for n, thr, mo in ((6000, 0.7, 1000), (2500, 0.4, 2500), (64, 0.5, 64)):
        boxes, scores = roi3d_synth.nms_boxes(n, (128, 128, 128), seed=1200 + n)
        if n == 64:
            boxes[7] = [0.4, 0.4, 0.4, 0.4, 0.7, 0.7]
            scores[7] = 3.0
        assert np.array_equal(run_nms(rb, cuda_device, boxes, scores, mo, thr), ref.non_max_suppression_3d(boxes, scores, mo, thr))
    # CropAndResize3D forward: bit-exact against CropAndResize3DOp::Compute; grads within tolerance
    for case in ((2, 8, 8, 16, 64, 24, (7, 7, 7)), (2, 8, 8, 16, 256, 12, (14, 14, 14)), (3, 9, 5, 7, 1, 9, (5, 3, 4))):
        B, H, W, D, C, n, crop = case
        image, boxes, bidx, grads = car_inputs(1300 + C, B, H, W, D, C, n, crop)
        t = [dev(x, cuda_device) for x in (image, boxes, bidx, grads)]
        out = rb.crop_and_resize_3d(t[0], t[1], t[2], crop, extrapolation_value=1.5).cpu().numpy()
        assert np.array_equal(out, ref.crop_and_resize_3d(image, boxes, bidx, crop, "trilinear", 1.5))
        gi = rb.crop_and_resize_3d_grad_image(t[3], t[1], t[2], image.shape).cpu().numpy()
        assert rel_ok(gi, ref.crop_and_resize_3d_grad_image(grads, boxes, bidx, image.shape), BWD_TOL)
        gb = rb.crop_and_resize_3d_grad_boxes(t[3], t[0], t[1], t[2]).cpu().numpy()
        rgb = ref.crop_and_resize_3d_grad_boxes(grads, image, boxes, bidx)
        assert np.all(np.abs(gb - rgb) <= GB_TOL * np.abs(rgb).max() + 1e-6)


def test_deferred_host_pipeline(rb, cuda_device):
    """rb.deferred(): host results are valid after the block; copies and kernels of independent ops overlap."""
    B, H, W, D, C, n, crop = 2, 8, 8, 16, 64, 12, (7, 7, 7)
    cases = [car_inputs(800 + i, B, H, W, D, C, n, crop) for i in range(4)]
    with rb.deferred():
        outs = [rb.crop_and_resize_3d(im, bx, bi, crop) for im, bx, bi, _ in cases]
        gis = [rb.crop_and_resize_3d_grad_image(g, bx, bi, im.shape) for im, bx, bi, g in cases]
    for (im, bx, bi, g), out, gi in zip(cases, outs, gis):
        assert np.array_equal(out, oracle.crop_and_resize_3d(im, bx, bi, crop))
        assert rel_ok(gi, oracle.crop_and_resize_3d_grad_image(g, bx, bi, im.shape), BWD_TOL)


# =============================================================================================
# batched / per-class NMS3D
# =============================================================================================
def test_nms_batched_matches_per_segment_oracle(rb, cuda_device):
    rng = np.random.default_rng(21)
    sizes = [0, 1, 37, 2000, 640, 0, 6000, 33]
    parts = [roi3d_synth.nms_boxes(max(n, 1), (128, 128, 128), seed=2100 + i) for i, n in enumerate(sizes)]
    boxes = np.concatenate([p[0][:n] for p, n in zip(parts, sizes)])
    scores = np.concatenate([p[1][:n] for p, n in zip(parts, sizes)])
    offs = np.concatenate([[0], np.cumsum(sizes)])
    for thr, mo in ((0.7, 1000), (0.3, 50)):
        got = rb.non_max_suppression_3d_batched(dev(boxes, cuda_device), dev(scores, cuda_device), offs, mo, thr)
        assert len(got) == len(sizes)
        for z, n in enumerate(sizes):
            ref = oracle.non_max_suppression_3d(boxes[offs[z]:offs[z + 1]], scores[offs[z]:offs[z + 1]], mo, thr)
            assert np.array_equal(got[z].cpu().numpy(), ref), (z, n)
    # ProposalLayer shape: a batch of images, 6000 boxes each (core/models.py:487-490)
    B = 4
    parts = [roi3d_synth.nms_boxes(6000, (128, 128, 128), seed=2200 + i, presorted=True) for i in range(B)]
    got = rb.non_max_suppression_3d_batched(dev(np.concatenate([p[0] for p in parts]), cuda_device),
                                            dev(np.concatenate([p[1] for p in parts]), cuda_device),
                                            np.arange(B + 1) * 6000, 1000, 0.7)
    for z in range(B):
        assert np.array_equal(got[z].cpu().numpy(), oracle.non_max_suppression_3d(parts[z][0], parts[z][1], 1000, 0.7))


def test_nms_per_class_and_graph_wrapper(rb, cuda_device):
    """BASELINE cfg3: per-class NMS over 2000 refined boxes; wrapper signature of core/utils.py:467-503."""
    boxes, scores = roi3d_synth.nms_boxes(2000, (256, 256, 256), seed=23)
    cls = np.random.default_rng(23).integers(1, 3, 2000)
    keep, classes = rb.non_max_suppression_3d_per_class(dev(boxes, cuda_device), dev(scores, cuda_device),
                                                        dev(cls, cuda_device), 100, 0.3)
    assert classes == [1, 2]
    for k, c in zip(keep, classes):
        ix = np.nonzero(cls == c)[0]
        ref = ix[oracle.non_max_suppression_3d(boxes[ix], scores[ix], 100, 0.3)]
        assert np.array_equal(k.cpu().numpy(), ref)
    sel, kidx = rb.non_max_suppression_3d_graph(dev(boxes, cuda_device), dev(scores, cuda_device), 0.3, 200)
    ref = oracle.non_max_suppression_3d(boxes, scores, 200, 0.3)
    assert np.array_equal(kidx.cpu().numpy(), ref) and np.array_equal(sel.cpu().numpy(), boxes[ref])


# =============================================================================================
# fused PyramidROIAlign3D
# =============================================================================================
@pytest.mark.parametrize("pool", [(7, 7, 7), (14, 14, 14), (3, 5, 4)])
def test_fused_pyramid_roi_align(rb, cuda_device, pool):
    import torch
    rng = np.random.default_rng(31)
    B, R, C = 2, 40, 32
    image_shape = (512, 512, 64)
    level_hw = [(32, 32), (16, 16), (8, 8), (4, 4)]
    fms = [rng.standard_normal((B, h, w, 64, C), dtype=np.float32) for h, w in level_hw]
    boxes = np.stack([roi3d_synth.rois(R, image_shape, seed=3100 + i, side_px=(6.0, 600.0)) for i in range(B)])
    boxes[0, 0] = 0.0                                           # ProposalLayer zero padding (core/models.py:476-484)
    boxes[0, 1] = [-0.3, 0.2, 0.1, 0.4, 1.7, 0.9]               # clipped by the layer
    fms[1][0, 3, 3, 5, 0] = np.inf                              # non-finite features are scrubbed to 0 (:683)
    _, level = oracle.pyramid_prepare(boxes, image_shape)
    assert len(np.unique(level)) >= 3                           # several levels exercised
    t_fms = [dev(f, cuda_device).requires_grad_(True) for f in fms]
    out = rb.pyramid_roi_align_3d(dev(boxes, cuda_device), image_shape, t_fms, pool)
    ref = oracle.pyramid_roi_align(boxes, image_shape, fms, pool)
    assert out.shape == ref.shape
    assert np.array_equal(out.detach().cpu().numpy(), ref)
    # float16 output mode = the target files' `rois_aligned` payload: bit-identical to astype(float16) of the above
    fms[0][0, 5, 5, 7, 1] = 1e6                                 # overflows float16 -> inf, like numpy
    with torch.no_grad():
        h = rb.pyramid_roi_align_3d(dev(boxes, cuda_device), image_shape, [dev(f, cuda_device) for f in fms], pool,
                                    out_dtype=torch.float16)
    with np.errstate(over="ignore"):
        href = oracle.pyramid_roi_align(boxes, image_shape, fms, pool).astype(np.float16)
    assert h.dtype == torch.float16 and np.array_equal(h.cpu().numpy().view(np.uint16), href.view(np.uint16))
    fms[0][0, 5, 5, 7, 1] = 0.5
    grads = rng.standard_normal(ref.shape, dtype=np.float32)
    out.backward(dev(grads, cuda_device))
    gref = oracle.pyramid_roi_align_grad(grads, boxes, image_shape, [f.shape for f in fms])
    for t, g in zip(t_fms, gref):
        assert rel_ok(t.grad.cpu().numpy(), g, BWD_TOL)


def test_fused_pyramid_equals_per_level_ops_cfg2(rb, cuda_device):
    """cfg2 shapes: the fused launch equals routing on the host + the four drop-in ops, bit for bit."""
    import torch
    vol, B, R = (128, 128, 128), 2, 128
    boxes = np.stack([roi3d_synth.rois(R, vol, seed=2002 * 131 + b) for b in range(B)])
    torch.manual_seed(5)
    fms = [torch.randn(roi3d_synth.level_shape(vol, lv, batch=B), device=cuda_device) for lv in roi3d_synth.LEVELS]
    out = rb.pyramid_roi_align_3d(dev(boxes, cuda_device), vol, fms, (7, 7, 7))
    bp, level = oracle.pyramid_prepare(boxes, vol)
    for i, lv in enumerate(roi3d_synth.LEVELS):
        ib, ir = np.nonzero(level == lv)
        if len(ib) == 0:
            continue
        per = rb.crop_and_resize_3d(fms[i], dev(bp[ib, ir], cuda_device), dev(ib.astype(np.int32), cuda_device), (7, 7, 7))
        assert torch.equal(out[torch.from_numpy(ib).to(cuda_device), torch.from_numpy(ir).to(cuda_device)], per)


# =============================================================================================
# box-space helpers
# =============================================================================================
def test_overlaps_graph_bit_exact(rb, cuda_device):
    b1 = roi3d_synth.rois(700, (128, 128, 128), seed=41, side_px=(8, 64))
    b2 = roi3d_synth.rois(37, (128, 128, 128), seed=42, side_px=(8, 64))
    b2[:20] = b1[:20] + np.float32(0.01)                        # real overlaps
    b1[5] = b1[5][[3, 4, 5, 0, 1, 2]]                           # negative volume: overlaps_graph does not reorder corners
    out = rb.overlaps_3d(dev(b1, cuda_device), dev(b2, cuda_device)).cpu().numpy()
    ref = oracle.overlaps_graph(b1, b2)
    assert out.shape == (700, 37) and (ref > 0).sum() > 20
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))


def test_decode_proposals_matches_graph(rb, cuda_device):
    rng = np.random.default_rng(43)
    n = 20000
    anchors = roi3d_synth.rois(n, (128, 128, 128), seed=44, side_px=(4, 96))
    deltas = (rng.standard_normal((n, 6)) * 3).astype(np.float32)
    std = (0.1, 0.1, 0.1, 0.2, 0.2, 0.2)
    index = rng.permutation(n)[:6000].astype(np.int32)
    for ix in (None, index):
        out = rb.decode_proposals(dev(anchors, cuda_device), dev(deltas, cuda_device), std, 128,
                                  None if ix is None else dev(ix, cuda_device)).cpu().numpy()
        ref = oracle.decode_proposals(anchors, deltas, std, 128, ix)
        assert out.shape == ref.shape
        assert np.allclose(out, ref, rtol=2e-6, atol=2e-7)      # expf vs numpy exp: <= 2 ulp, then clipped to [0,1]
        assert (out[:, 3:] > out[:, :3]).all() and out.min() >= 0 and out.max() <= 1 + 1.0 / 128 + 1e-6   # min-size may exceed 1


def _topk_ref(scores, k):
    """tf.nn.top_k's set: k largest, threshold ties -> lower indices."""
    order = np.lexsort((np.arange(len(scores)), -scores.astype(np.float64)))
    return np.sort(order[:k])


@pytest.mark.parametrize("n,k", [(1, 1), (1000, 1), (1000, 1000), (5000, 37), (300000, 6000), (1 << 20, 6000), (4100, 4096)])
def test_top_k_set(rb, cuda_device, n, k):
    rng = np.random.default_rng(n + k)
    for kind in ("uniform", "ties", "constant"):
        if kind == "uniform":
            s = rng.random(n).astype(np.float32)
        elif kind == "ties":
            s = (rng.integers(0, 50, n) / 50.0).astype(np.float32)     # the threshold always falls inside a tie class
            s[::97] = -s[::97]
        else:
            s = np.full(n, 0.25, np.float32)
        idx, val = rb.top_k_set(dev(s, cuda_device), k)
        ref = _topk_ref(s, k)
        assert np.array_equal(idx.cpu().numpy(), ref)
        assert np.array_equal(val.cpu().numpy(), s[ref])


def test_proposal_layer_matches_restated_graph(rb, cuda_device):
    """top_k -> apply deltas -> clip -> min size -> NMS3D -> gather -> pad, against numpy + the NMS oracle."""
    rng = np.random.default_rng(51)
    n, pre, P, thr, depth = 120000, 6000, 1000, 0.7, 128
    anchors = roi3d_synth.nms_boxes(n, (128, 128, 128), seed=52, side_px=(8.0, 64.0))[0]
    deltas = (rng.standard_normal((n, 6)) * 0.5).astype(np.float32)
    scores = rng.random(n).astype(np.float32)
    scores[rng.integers(0, n, 2000)] = scores[rng.integers(0, n, 2000)]
    std = (0.1, 0.1, 0.1, 0.2, 0.2, 0.2)
    out, count = rb.proposal_layer(dev(scores, cuda_device), dev(deltas, cuda_device), dev(anchors, cuda_device), std, depth, pre, P, thr)
    sel = _topk_ref(scores, pre)
    # reference order: top_k sorted by (score desc, index asc); NMS ties follow that order
    order = sel[np.lexsort((sel, -scores[sel].astype(np.float64)))]
    boxes = oracle.decode_proposals(anchors, deltas, std, depth, order)
    keep = oracle.non_max_suppression_3d(boxes, scores[order], P, thr)
    ref = np.zeros((P, 6), np.float32)
    ref[:len(keep)] = boxes[keep]
    got = out.cpu().numpy()
    assert int(count.item()) == len(keep)
    assert np.allclose(got, ref, rtol=2e-6, atol=2e-7)          # decode uses expf (<= 2 ulp from numpy's exp)


def test_cfg3_large_volume_256(rb, cuda_device):
    """BASELINE cfg3 scale: P2 of a 256^3 volume is [1,64,64,256,256] = 1.07 GB (32-bit element offsets must hold)."""
    import torch
    vol = (256, 256, 256)
    shape = roi3d_synth.level_shape(vol, 2, batch=1)
    assert int(np.prod(shape)) == 64 * 64 * 256 * 256
    torch.manual_seed(3)
    image = torch.randn(shape, device=cuda_device)
    boxes = roi3d_synth.rois(48, vol, seed=3003, side_px=(8.0, 200.0))
    boxes[0] = [0.97, 0.97, 0.97, 1.0, 1.0, 1.0]                # the far corner: largest offsets
    bidx = np.zeros(48, np.int32)
    tb, ti = dev(boxes, cuda_device), dev(bidx, cuda_device)
    for crop in ((7, 7, 7), (14, 14, 14)):
        out = rb.crop_and_resize_3d(image, tb, ti, crop)
        pick = np.array([0, 1, 17, 47])
        img_np = image.cpu().numpy()
        ref = oracle.crop_and_resize_3d(img_np, boxes[pick], bidx[pick], crop)
        assert np.array_equal(out[torch.from_numpy(pick).to(cuda_device)].cpu().numpy(), ref)
        g = torch.randn_like(out)
        gi = rb.crop_and_resize_3d_grad_image(g, tb, ti, shape)
        lhs = float((out.double() * g.double()).sum())
        rhs = float((image.double() * gi.double()).sum())
        assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), float(out.double().norm() * g.double().norm()) * 1e-3)
        del out, g, gi, img_np
    # NMS3D at cfg3 size with the oracle (20000 -> 2000 @0.7)
    nb, ns = roi3d_synth.nms_boxes(20000, vol)
    assert np.array_equal(run_nms(rb, cuda_device, nb, ns, 2000, 0.7), oracle.non_max_suppression_3d(nb, ns, 2000, 0.7))


def test_car_random_shapes(rb, cuda_device):
    """80 seeded random geometries (channels, volume dims, crop dims, weird boxes) through both kernel variants."""
    rng = np.random.default_rng(2024)
    for it in range(80):
        C = int(rng.choice([1, 3, 4, 8, 12, 32, 36, 64, 96, 128, 256]))
        B = int(rng.integers(1, 3))
        H, W, D = (int(v) for v in rng.integers(1, 24, 3))
        crop = tuple(int(v) for v in rng.integers(1, 18 if C > 64 else 30, 3))
        n = int(rng.integers(0, 10))
        spread = float(rng.choice([0.0, 0.4, 1.5]))
        image = rng.standard_normal((B, H, W, D, C), dtype=np.float32)
        boxes = (rng.random((n, 6)) * (1 + spread) - spread / 2).astype(np.float32)
        if it % 3 == 0 and n:
            boxes[:, 3:] = np.maximum(boxes[:, 3:], boxes[:, :3])           # ordered corners
        bidx = rng.integers(0, B, n).astype(np.int32)
        grads = rng.standard_normal((n,) + crop + (C,), dtype=np.float32)
        ref = oracle.crop_and_resize_3d(image, boxes, bidx, crop, "trilinear", -2.0)
        gref = oracle.crop_and_resize_3d_grad_image(grads, boxes, bidx, image.shape)
        t = [dev(x, cuda_device) for x in (image, boxes, bidx, grads)]
        for variant in (1, 2):
            rb.custom_op.set_option("car_fwd_variant", variant)
            rb.custom_op.set_option("car_bwd_variant", variant)
            out = rb.crop_and_resize_3d(t[0], t[1], t[2], crop, extrapolation_value=-2.0).cpu().numpy()
            assert np.array_equal(out, ref), (it, variant, C, (H, W, D), crop, n)
            if variant == 2:
                rb.custom_op.set_option("car_fwd_variant", 3)
                out = rb.crop_and_resize_3d(t[0], t[1], t[2], crop, extrapolation_value=-2.0).cpu().numpy()
                assert np.array_equal(out, ref), (it, 3, C, (H, W, D), crop, n)
            gi = rb.crop_and_resize_3d_grad_image(t[3], t[1], t[2], image.shape).cpu().numpy()
            assert rel_ok(gi, gref, BWD_TOL), (it, variant, C, (H, W, D), crop, n)


@pytest.mark.parametrize("n,max_out,thr,kw", [
    (6000, 1000, 0.7, {}),                                  # head phase suffices (scan depth ~1020)
    (6000, 1000, 0.3, {}),                                  # tail phase needed
    (6000, 1000, 0.7, dict(cluster=64, jitter=0.05)),       # dense clusters: deep scan
    (6000, 100, 0.5, dict(cluster=64, jitter=0.05)),
    (3000, 2500, 0.5, {}),                                  # head covers everything
    (2049, 1, 0.5, {}), (2049, 1400, 0.0, {}), (5000, 1300, 1.0, {}),
    (20000, 2000, 0.7, dict(cluster=64, jitter=0.05)),
])
def test_nms_head_tail_schedule(rb, cuda_device, n, max_out, thr, kw):
    """The speculative head/tail split (mask triangle + scan of the first ~1.25 max_out boxes, the rest behind a done
    flag) returns exactly what the single-phase schedule and the oracle return, whichever phase finishes the job."""
    boxes, scores = roi3d_synth.nms_boxes(n, (128, 128, 128), seed=77 + n + max_out, **kw)
    ref = oracle.non_max_suppression_3d(boxes, scores, max_out, thr)
    try:
        for variant in (0, 1):
            rb.custom_op.set_option("nms_variant", variant)
            assert np.array_equal(run_nms(rb, cuda_device, boxes, scores, max_out, thr), ref), variant
    finally:
        rb.custom_op.set_option("nms_variant", 0)
    # degenerate (zero-volume) boxes in the tail region keep the reference's refill quirk across the hand-over
    b2 = boxes.copy()
    order = np.argsort(-scores, kind="stable")
    pos = order[min(n - 1, max_out + max_out // 4 + 900)]
    b2[pos, 3:] = b2[pos, :3]
    assert np.array_equal(run_nms(rb, cuda_device, b2, scores, max_out, thr),
                          oracle.non_max_suppression_3d(b2, scores, max_out, thr))


@pytest.mark.parametrize("n,max_out,dist", [
    (1, 1, "uniform"), (100, 50, "uniform"), (6000, 1000, "uniform"), (20000, 2000, "uniform"), (50000, 3000, "uniform"),
    (5000, 5000, "constant"), (6000, 700, "peaked"), (6000, 6000, "wide"), (4097, 300, "ties"),
])
def test_nms_bucketed_sort(rb, cuda_device, n, max_out, dist):
    """The bucketed sort (large n) orders exactly like the rank-by-counting sort: same NMS result for uniform scores,
    constant scores (one bucket), scores crowded near 1, scores outside [0, 1] with NaN / -inf / -FLT_MAX non-candidates,
    and heavy ties."""
    boxes, scores = roi3d_synth.nms_boxes(n, (128, 128, 128), seed=31 + n)
    rng = np.random.default_rng(n)
    if dist == "constant":
        scores[:] = 0.5
    elif dist == "peaked":
        scores = (1.0 - rng.random(n) ** 6 * 1e-3).astype(np.float32)
    elif dist == "wide":
        scores = (rng.standard_normal(n) * 3).astype(np.float32)
        scores[::97] = np.nan
        scores[5::101] = -np.inf
        scores[7::103] = -np.finfo(np.float32).max
        scores[9::107] = np.inf
        scores[11::109] = -0.0
        scores[13::113] = 0.0
    elif dist == "ties":
        scores = (rng.integers(0, 12, n) / 11.0).astype(np.float32)
    ref = oracle.non_max_suppression_3d(boxes, scores, max_out, 0.5)
    try:
        for variant in (2, 1):
            rb.custom_op.set_option("nms_sort_variant", variant)
            assert np.array_equal(run_nms(rb, cuda_device, boxes, scores, max_out, 0.5), ref), variant
    finally:
        rb.custom_op.set_option("nms_sort_variant", 0)


def test_nms_unaligned_scores_and_tail_vectors(rb, cuda_device):
    """The rank sort stages scores with 16-byte loads when it can; views that are not 16-byte aligned and lengths that
    are not multiples of four take the pre-computed-key path / the padded tail and return the same result."""
    for n in (5, 6, 7, 1001, 1002, 1003, 6001):
        boxes, scores = roi3d_synth.nms_boxes(n + 1, (128, 128, 128), seed=500 + n)
        tb, ts = dev(boxes, cuda_device), dev(scores, cuda_device)
        for off in (0, 1):
            ref = oracle.non_max_suppression_3d(boxes[off:], scores[off:], 300, 0.5)
            got = rb.non_max_suppression_3d(tb[off:], ts[off:], 300, 0.5).cpu().numpy()
            assert np.array_equal(got, ref), (n, off)
