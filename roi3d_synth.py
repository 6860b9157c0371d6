"""Seeded synthetic inputs of the ROI hot path, shared by the tests,
bench.py and the CPU baseline so that the oracle and the CUDA kernels always see the same
data.  Pure numpy, no GPU.

Shapes follow the reference graph: this fork's backbone/FPN never stride depth
(core/models.py:242-267, 3193-3202), so pyramid level l of an HxWxD volume is
``[B, H/2^l, W/2^l, D, 256]``.
"""
import numpy as np

TOP_DOWN_PYRAMID_SIZE = 256          # core/config.py TOP_DOWN_PYRAMID_SIZE
LEVELS = (2, 3, 4, 5)


def level_shape(volume, level, batch=1, channels=TOP_DOWN_PYRAMID_SIZE, isotropic=False):
    H, W, D = volume
    s = 1 << level
    return (batch, max(H // s, 1), max(W // s, 1), max(D // s, 1) if isotropic else D, channels)


def feature_map(volume, level, batch=1, channels=TOP_DOWN_PYRAMID_SIZE, isotropic=False, seed=None):
    """N(0,1) float32 pyramid level, rng seed 1000 + level."""
    rng = np.random.default_rng(1000 + level if seed is None else seed)
    return rng.standard_normal(level_shape(volume, level, batch, channels, isotropic), dtype=np.float32)


def min_size_rule(boxes, depth):
    """ProposalLayer / PyramidROIAlign clip + min-size (core/models.py:429-447, 615-632)."""
    b = np.clip(boxes.astype(np.float32), 0.0, 1.0)
    eps = np.float32(1e-6)
    b[:, 3] = np.maximum(b[:, 3], b[:, 0] + eps)
    b[:, 4] = np.maximum(b[:, 4], b[:, 1] + eps)
    b[:, 5] = np.maximum(b[:, 5], b[:, 2] + np.float32(max(1.0 / max(depth, 1), 1e-4)))
    return b


def rois(n, volume, seed, side_px=(8.0, 96.0)):
    """n ROIs: centres U(0,1)^3, per-axis side log-uniform in `side_px` voxels, normalized."""
    rng = np.random.default_rng(seed)
    H, W, D = volume
    c = rng.uniform(0.0, 1.0, (n, 3))
    side = np.exp(rng.uniform(np.log(side_px[0]), np.log(side_px[1]), (n, 3))) / np.array([H, W, D], np.float64)
    boxes = np.concatenate([c - side / 2, c + side / 2], axis=1).astype(np.float32)
    return min_size_rule(boxes, D)


def roi_levels(boxes, volume):
    """PyramidROIAlign level routing (core/models.py:637-649); fp32 like the TF graph."""
    H, W, D = (np.float32(v) for v in volume)
    h = boxes[:, 3] - boxes[:, 0]
    w = boxes[:, 4] - boxes[:, 1]
    d = boxes[:, 5] - boxes[:, 2]
    vol = (h * w * d).astype(np.float32)
    image_area = np.float32(H * W * D)
    ratio = np.power(vol, np.float32(1.0 / 3.0)) / (np.float32(224.0) / np.power(image_area, np.float32(1.0 / 3.0)))
    lvl = np.log(ratio.astype(np.float32)) / np.float32(np.log(2.0))
    return np.minimum(5, np.maximum(2, 4 + np.rint(lvl).astype(np.int32)))   # tf.round = half to even


def pyramid_rois(n_per_image, batch, volume, seed):
    """ROIs for a batch, routed to levels: {level: (boxes [N_l,6], box_index [N_l], order [N_l])}."""
    all_boxes, all_idx = [], []
    for b in range(batch):
        all_boxes.append(rois(n_per_image, volume, seed * 131 + b))
        all_idx.append(np.full(n_per_image, b, np.int32))
    boxes = np.concatenate(all_boxes)
    idx = np.concatenate(all_idx)
    lv = roi_levels(boxes, volume)
    out = {}
    for level in LEVELS:
        sel = np.nonzero(lv == level)[0]
        out[level] = (np.ascontiguousarray(boxes[sel]), np.ascontiguousarray(idx[sel]), sel)
    return out


def nms_boxes(n, volume, seed=None, presorted=False, side_px=(8.0, 64.0), jitter=0.15, cluster=8):
    """Clustered boxes + scores for NMS3D: n/cluster seeds jittered `cluster` times
    (sigma = jitter * side), scores U(0,1) with 1 % exact duplicates."""
    rng = np.random.default_rng(3000 + n if seed is None else seed)
    H, W, D = volume
    dims = np.array([H, W, D], np.float64)
    ns = (n + cluster - 1) // cluster
    c = rng.uniform(0.0, 1.0, (ns, 3))
    side = np.exp(rng.uniform(np.log(side_px[0]), np.log(side_px[1]), (ns, 3))) / dims
    c = np.repeat(c, cluster, axis=0)[:n]
    side = np.repeat(side, cluster, axis=0)[:n]
    c = c + rng.standard_normal((n, 3)) * jitter * side
    side = side * np.exp(rng.standard_normal((n, 3)) * 0.1)
    boxes = min_size_rule(np.concatenate([c - side / 2, c + side / 2], axis=1), D)
    scores = rng.uniform(0.0, 1.0, n).astype(np.float32)
    ndup = max(n // 100, 1 if n > 1 else 0)
    if ndup:
        src = rng.integers(0, n, ndup)
        dst = rng.integers(0, n, ndup)
        scores[dst] = scores[src]
    perm = rng.permutation(n)
    boxes, scores = boxes[perm], scores[perm]
    if presorted:                                   # ProposalLayer feeds top_k(sorted=True) output
        order = np.argsort(-scores, kind="stable")
        boxes, scores = boxes[order], scores[order]
    return np.ascontiguousarray(boxes), np.ascontiguousarray(scores)


def grads_like(shape, seed):
    return np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)


# ---- algorithmic bytes ------------------------------------------
def _axis_footprint(a1, a2, dim, p):
    """Distinct integer taps {floor(in), ceil(in)} over the p in-range samples of one axis (fp32 math)."""
    a1, a2 = np.float32(a1), np.float32(a2)
    if p > 1:
        scale = np.float32(np.float32((a2 - a1) * np.float32(dim - 1)) / np.float32(p - 1))
        ins = np.float32(a1 * np.float32(dim - 1)) + np.arange(p, dtype=np.float32) * scale
        ins = ins.astype(np.float32)
    else:
        ins = np.array([np.float32(np.float64(np.float32(a1 + a2)) * 0.5 * np.float64(dim - 1))], np.float32)
    ok = ~((ins < 0) | (ins > np.float32(dim - 1)))
    ins = ins[ok]
    return len(np.unique(np.concatenate([np.floor(ins), np.ceil(ins)]))), int(ok.sum())


def car_algorithmic_bytes(boxes, image_shape, crop, backward=False):
    """fwd: sum[(p^3 + ny*nx*nz) * C * 4] + 28 N ; bwd adds the footprint once more and the zero-fill."""
    B, H, W, D, C = image_shape
    ph, pw, pd = crop
    total = 28 * len(boxes)
    for bx in boxes:
        ny, _ = _axis_footprint(bx[0], bx[3], H, ph)
        nx, _ = _axis_footprint(bx[1], bx[4], W, pw)
        nz, _ = _axis_footprint(bx[2], bx[5], D, pd)
        fp = ny * nx * nz
        total += (ph * pw * pd + (2 * fp if backward else fp)) * C * 4
    if backward:
        total += B * H * W * D * C * 4
    return int(total)
