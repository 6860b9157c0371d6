"""GPU parity tests for the rows SURVEY.md section 8 marks "next" after the hot ops: DetectionLayer
(refine_detections, row f3), mask targets and the target files' wire format (row f4).

Bars: bit-exact for the byte / index work (packbits, float16 payloads, rounded mask targets, the selected ROIs and
their order, scores); the refined boxes go through expf, so they are compared at 1e-5 relative (float tolerance).
"""
import numpy as np
import pytest

import oracle
import roi3d_synth

pytestmark = pytest.mark.gpu


def dev(x, cuda_device):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda_device)


def detection_inputs(seed, B, R, K, vol=(128, 128, 128)):
    rng = np.random.default_rng(seed)
    rois = np.stack([roi3d_synth.nms_boxes(R, vol, seed=seed + 10 * b)[0] for b in range(B)]).astype(np.float32)
    logits = rng.standard_normal((B, R, K)).astype(np.float32) * 2
    probs = np.exp(logits) / np.exp(logits).sum(-1, keepdims=True)
    deltas = (rng.standard_normal((B, R, K, 6)) * 1.5).astype(np.float32)
    deltas[:, ::17, 1, 3:] *= 40.0                       # some log-scale deltas beyond the +-log(62.5) clip
    return rois, probs.astype(np.float32), deltas


@pytest.mark.parametrize("B,R,K,min_conf,thr,max_inst", [
    (1, 1000, 2, 0.5, 0.3, 100), (2, 2000, 2, 0.7, 0.3, 200), (3, 257, 3, 0.0, 0.5, 400), (1, 1, 2, 0.0, 0.3, 5),
    (2, 64, 2, 0.999999, 0.3, 10),
])
@pytest.mark.parametrize("nms_mode", ["reference_2d", "3d"])
def test_refine_detections_matches_restated_graph(rb, cuda_device, B, R, K, min_conf, thr, max_inst, nms_mode):
    rois, probs, deltas = detection_inputs(4000 + R, B, R, K)
    shape = (128.0, 128.0, 64.0)
    det, cnt = rb.refine_detections(dev(rois, cuda_device), dev(probs, cuda_device), dev(deltas, cuda_device), shape,
                                    min_conf, thr, None, max_inst, return_counts=True, nms_mode=nms_mode)
    det, cnt = det.cpu().numpy(), cnt.cpu().numpy()
    assert det.shape == (B, max_inst, 8)
    for b in range(B):
        # the device's own decoded boxes (expf may differ from numpy's exp in the last ulp) -> the selection is compared exactly
        ref = oracle.refine_detections(rois[b], probs[b], deltas[b], shape, min_conf, thr, max_instances=max_inst, nms_mode=nms_mode)
        k = int((ref[:, 6] > 0).sum())
        assert cnt[b] == k
        assert np.array_equal(det[b, :, 7], ref[:, 7])                     # same ROIs, same order (scores are copied)
        assert np.array_equal(det[b, :, 6], ref[:, 6])
        assert np.allclose(det[b, :, :6], ref[:, :6], rtol=1e-5, atol=1e-6)
        assert not det[b, k:].any()                                        # tf.pad rows
        assert np.all(np.diff(det[b, :k, 7]) <= 0)                         # descending score
    # single-image call == row of the batched call
    one = rb.refine_detections(dev(rois[0], cuda_device), dev(probs[0], cuda_device), dev(deltas[0], cuda_device), shape,
                               min_conf, thr, None, max_inst, nms_mode=nms_mode).cpu().numpy()
    assert np.array_equal(one, det[0])


def test_refine_detections_2d_vs_3d_and_threshold_tie(rb, cuda_device):
    """Boxes stacked along z overlap fully in (y, x) and not at all in z: the fork's 2-D NMS keeps one, the 3-D op keeps
    all.  And the compare rules differ exactly at IoU == threshold: tf.image suppresses on >, the 3-D op on >=."""
    n = 6
    rois = np.zeros((n, 6), np.float32)
    for i in range(n):                                       # same (y, x) square, disjoint z slabs
        rois[i] = [0.25, 0.25, i / 8.0, 0.5, 0.5, (i + 0.9) / 8.0]
    probs = np.stack([np.full(n, 0.1, np.float32), np.linspace(0.9, 0.6, n).astype(np.float32)], axis=1)
    deltas = np.zeros((n, 2, 6), np.float32)
    shape = (128.0, 128.0, 64.0)
    args = (dev(rois, cuda_device), dev(probs, cuda_device), dev(deltas, cuda_device), shape, 0.5, 0.3, None, 10)
    d2 = rb.refine_detections(*args, nms_mode="reference_2d").cpu().numpy()
    d3 = rb.refine_detections(*args, nms_mode="3d").cpu().numpy()
    assert int((d2[:, 6] > 0).sum()) == 1 and int((d3[:, 6] > 0).sum()) == n
    for mode, d in (("reference_2d", d2), ("3d", d3)):
        assert np.array_equal(d[:, 7], oracle.refine_detections(rois, probs, deltas, shape, 0.5, 0.3, max_instances=10, nms_mode=mode)[:, 7])
    # IoU exactly 0.5: box A = [0,64]x[0,64], box B = [0,64]x[0,32] in pixels (area ratio 1/2, all values exact in fp32)
    rois = np.array([[0, 0, 0, 0.5, 0.5, 0.5], [0, 0, 0, 0.5, 0.25, 0.5]], np.float32)
    probs = np.array([[0.1, 0.9], [0.2, 0.8]], np.float32)
    deltas = np.zeros((2, 2, 6), np.float32)
    args = (dev(rois, cuda_device), dev(probs, cuda_device), dev(deltas, cuda_device), shape, 0.5, 0.5, None, 4)
    assert int((rb.refine_detections(*args, nms_mode="reference_2d").cpu().numpy()[:, 6] > 0).sum()) == 2   # 0.5 > 0.5 is false
    assert int((rb.refine_detections(*args, nms_mode="3d").cpu().numpy()[:, 6] > 0).sum()) == 1             # 0.5 >= 0.5


def test_refine_detections_host_buffers_and_errors(rb, cuda_device):
    rois, probs, deltas = detection_inputs(4100, 1, 300, 2)
    out = rb.refine_detections(rois[0], probs[0], deltas[0], (128, 128, 64), 0.5, 0.3, [0.1] * 3 + [0.2] * 3, 50)
    assert isinstance(out, np.ndarray) and out.shape == (50, 8)
    ref = oracle.refine_detections(rois[0], probs[0], deltas[0], (128, 128, 64), 0.5, 0.3, max_instances=50)
    assert np.array_equal(out[:, 7], ref[:, 7])
    with pytest.raises(rb.InvalidArgumentError):
        rb.refine_detections(rois[0], probs[0][:, :1], deltas[0][:, :1], (128, 128, 64), 0.5, 0.3)
    with pytest.raises(rb.InvalidArgumentError):
        rb.refine_detections(rois[0], probs[0], deltas[0], (128, 128, 64), 0.5, 1.5)


@pytest.mark.parametrize("dtype", [np.float32, np.uint8, np.bool_])
@pytest.mark.parametrize("G,H,W,D,n,mshape", [(5, 24, 20, 16, 37, (28, 28, 28)), (3, 9, 7, 5, 4, (3, 5, 2)), (2, 16, 16, 16, 1, (7, 7, 7))])
def test_mask_targets_bit_exact(rb, cuda_device, dtype, G, H, W, D, n, mshape):
    rng = np.random.default_rng(G * 100 + n)
    masks = (rng.random((G, H, W, D)) > 0.6).astype(dtype)
    if dtype == np.float32:
        masks = masks * rng.random((G, H, W, D)).astype(np.float32) * 2            # soft masks exercise the rounding
    boxes = np.asarray(roi3d_synth.rois(n, (H, W, D), seed=n, side_px=(3.0, 20.0)), np.float32).reshape(n, 6)
    boxes[0] = (-0.2, 0.1, 0.1, 0.9, 1.3, 0.8)                                      # leaves the volume: extrapolation 0
    assign = rng.integers(0, G, n).astype(np.int32)
    ref = oracle.mask_targets(masks, boxes, assign, mshape)
    out, bits = rb.mask_targets(dev(masks, cuda_device), dev(boxes, cuda_device), dev(assign, cuda_device), mshape, packed=True)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert np.array_equal(bits.cpu().numpy(), oracle.pack_bits(ref)[0])
    if n <= G:                                                                      # assignment None == range(n)
        out2 = rb.mask_targets(dev(masks, cuda_device), dev(boxes, cuda_device), None, mshape).cpu().numpy()
        assert np.array_equal(out2, oracle.mask_targets(masks, boxes, None, mshape))


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 31, 32, 33, 255, 1000, 4099, (1 << 20) + 5])
def test_wire_format_payloads_bit_exact(rb, cuda_device, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 10).astype(np.float32)
    if n >= 9:
        x[:9] = [0.5, np.nextafter(np.float32(0.5), np.float32(1)), np.inf, -np.inf, np.nan, 65519.9, 65520.0, 1e-8, -6e-8]
    t = dev(x, cuda_device)
    h = rb.pack_f16(t)
    assert np.array_equal(h.cpu().numpy().view(np.uint16), oracle.pack_f16(x).view(np.uint16))
    assert np.array_equal(rb.unpack_f16(h).cpu().numpy().view(np.uint32), oracle.pack_f16(x).astype(np.float32).view(np.uint32))
    bits, shape = rb.pack_bits(t)
    ref_bits, ref_shape = oracle.pack_bits(x)
    assert np.array_equal(bits.cpu().numpy(), ref_bits) and np.array_equal(shape, ref_shape)
    back = rb.unpack_bits(bits, (n,)).cpu().numpy()
    assert np.array_equal(back, oracle.unpack_bits(ref_bits, (n,)))
    assert np.array_equal(back, (x > 0.5).astype(np.float32))
    if n > 33:                                                   # unaligned views take the scalar paths
        assert np.array_equal(rb.pack_f16(t[1:]).cpu().numpy().view(np.uint16), oracle.pack_f16(x[1:]).view(np.uint16))
        assert np.array_equal(rb.unpack_bits(bits, (n - 3,)).cpu().numpy(), oracle.unpack_bits(ref_bits, (n - 3,)))


def test_target_files_interoperate_with_the_reference_format(rb, cuda_device, tmp_path):
    """Files written here load with the reference reader's numpy code (core/data_generators.py:1908-1921) and
    files written by the reference's numpy writer (core/models.py:3585-3636) load here."""
    rng = np.random.default_rng(5)
    T, P, M, C = 12, 7, 14, 16
    rois = rng.random((T, 6)).astype(np.float32)
    ra = rng.standard_normal((T, P, P, P, C)).astype(np.float32)
    ma = rng.random((T, M, M, M, 3)).astype(np.float32)
    tci = rng.integers(0, 2, T).astype(np.int32)
    tb = rng.standard_normal((T, 6)).astype(np.float32)
    tm = (rng.random((T, 28, 28, 28)) > 0.5).astype(np.float32)
    paths = rb.target_files.save_head_targets(str(tmp_path), "vol_001", dev(rois, cuda_device), dev(ra, cuda_device),
                                              dev(ma, cuda_device), tci, tb, dev(tm, cuda_device))
    # the reference reader
    z = np.load(paths[1]); assert z["rois_aligned"].dtype == np.float16
    assert np.array_equal(z["rois_aligned"].view(np.uint16), ra.astype(np.float16).view(np.uint16))
    z = np.load(paths[2])
    assert np.array_equal(z["mask_bits"], np.packbits((ma > 0.5).astype(np.uint8).reshape(-1)))
    assert np.array_equal(z["mask_shape"], np.array(ma.shape, np.int32))
    z = np.load(paths[5])
    flat = np.unpackbits(z["tm_bits"])[: tm.size].reshape(tuple(z["tm_shape"]))
    assert np.array_equal(flat.astype(np.float32), tm)
    assert np.array_equal(np.load(paths[0])["rois"], rois) and np.array_equal(np.load(paths[3])["tci"], tci)
    assert np.array_equal(np.load(paths[4])["bbox"], tb)
    # the reference writer -> our reader
    ref_dir = tmp_path / "ref"
    ref_paths = []
    for sub, arrays in (("rois", dict(rois=rois)), ("rois_aligned", dict(rois_aligned=ra.astype(np.float16))),
                        ("mask_aligned", dict(zip(("mask_bits", "mask_shape"), oracle.pack_bits(ma)))),
                        ("target_class_ids", dict(tci=tci)), ("target_bbox", dict(bbox=tb)),
                        ("target_mask", dict(zip(("tm_bits", "tm_shape"), oracle.pack_bits(tm))))):
        (ref_dir / sub).mkdir(parents=True)
        p = str(ref_dir / sub / "vol_001.npz")
        np.savez_compressed(p, **arrays)
        ref_paths.append(p)
    for src in (paths, ref_paths):
        r2, ra2, ma2, tci2, tb2, tm2 = rb.target_files.load_head_targets(src)
        assert np.array_equal(r2.cpu().numpy(), rois) and np.array_equal(tci2.cpu().numpy(), tci)
        assert np.array_equal(tb2.cpu().numpy(), tb)
        assert np.array_equal(ra2.cpu().numpy(), ra.astype(np.float16).astype(np.float32))
        assert np.array_equal(ma2.cpu().numpy(), (ma > 0.5).astype(np.float32))
        assert np.array_equal(tm2.cpu().numpy(), tm)
