"""Regenerates tests/golden/*.npz: small seeded input/output vectors of the four ops.

Outputs come from the CPU oracle (oracle/roi3d_oracle.c).  When the stub TF runtime is built
(oracle/refrun, needs /root/reference) `python tests/golden/make_golden.py --check-ref` also
runs the reference's own binaries on the same inputs and asserts they agree bit for bit
before writing, which is how the committed files were produced.

    python tests/golden/make_golden.py [--check-ref]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle          # noqa: E402
import roi3d_synth     # noqa: E402


def car_case(seed, B, H, W, D, C, n, crop, wild=False):
    rng = np.random.default_rng(seed)
    image = rng.standard_normal((B, H, W, D, C), dtype=np.float32)
    boxes = roi3d_synth.rois(n, (H * 4, W * 4, D), seed, side_px=(4.0, 3.0 * max(H, W)))
    if wild:                         # out-of-range, reversed and degenerate boxes
        boxes[0] = [-0.2, 0.1, 0.1, 0.7, 1.3, 0.9]
        boxes[1] = [0.8, 0.7, 0.9, 0.2, 0.1, 0.3]
        boxes[2] = [0.5, 0.5, 0.5, 0.5, 0.5, 0.5]
        boxes[3] = [0.0, 0.0, 0.0, 1.0, 1.0, 1.0]
    box_index = rng.integers(0, B, n).astype(np.int32)
    grads = rng.standard_normal((n,) + crop + (C,), dtype=np.float32)
    return image, boxes, box_index, grads


def main():
    check_ref = "--check-ref" in sys.argv
    ref = None
    if check_ref:
        from oracle import refrun
        ref = refrun.load()
    cases = {
        "car_a": (car_case(11, 2, 6, 7, 9, 8, 6, (3, 4, 5), wild=True), (3, 4, 5)),
        "car_b": (car_case(12, 1, 8, 8, 16, 4, 5, (7, 7, 7)), (7, 7, 7)),
        "car_c": (car_case(13, 2, 5, 4, 6, 3, 4, (1, 2, 1), wild=True), (1, 2, 1)),
    }
    for name, ((image, boxes, box_index, grads), crop) in cases.items():
        out = {"image": image, "boxes": boxes, "box_index": box_index, "grads": grads,
               "crop": np.array(crop, np.int32)}
        for method in ("trilinear", "nearest"):
            out["fwd_" + method] = oracle.crop_and_resize_3d(image, boxes, box_index, crop, method, 0.25)
            out["gi_" + method] = oracle.crop_and_resize_3d_grad_image(grads, boxes, box_index, image.shape, method)
        out["gb"] = oracle.crop_and_resize_3d_grad_boxes(grads, image, boxes, box_index)
        if ref is not None:
            for method in ("trilinear", "nearest"):
                # the reference's `nearest` forward is undefined for crop_width != crop_depth (CAR.so@0x5570)
                if method == "trilinear" or crop[1] == crop[2]:
                    assert np.array_equal(ref.crop_and_resize_3d(image, boxes, box_index, crop, method, 0.25),
                                          out["fwd_" + method], equal_nan=True), (name, method, "fwd")
                assert np.array_equal(ref.crop_and_resize_3d_grad_image(grads, boxes, box_index, image.shape, method),
                                      out["gi_" + method], equal_nan=True), (name, method, "gi")
            assert np.array_equal(ref.crop_and_resize_3d_grad_boxes(grads, image, boxes, box_index), out["gb"]), name
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    nms = {}
    for n, thr, mo in ((300, 0.3, 40), (1000, 0.7, 200), (64, 0.5, 64)):
        boxes, scores = roi3d_synth.nms_boxes(n, (64, 64, 64), seed=500 + n)
        if n == 64:                  # a zero-volume box with the top score: the r2.2 re-push quirk
            boxes[5] = [0.3, 0.3, 0.3, 0.3, 0.6, 0.6]
            scores[5] = 2.0
        keep = oracle.non_max_suppression_3d(boxes, scores, mo, thr)
        if ref is not None:
            assert np.array_equal(ref.non_max_suppression_3d(boxes, scores, mo, thr), keep), n
        nms.update({"boxes_%d" % n: boxes, "scores_%d" % n: scores, "keep_%d" % n: keep,
                    "args_%d" % n: np.array([mo, thr], np.float64)})
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **nms)
    print("golden vectors written to", HERE, "(reference-checked)" if ref else "(oracle only)")


if __name__ == "__main__":
    main()
