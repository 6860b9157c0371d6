"""Host-side mirror of the reference's ``core/custom_op`` surface on B200.

The reference imports four callables (core/custom_op/custom_op.py:22-25) and
registers one gradient (core/custom_op/custom_op.py:28-65):

    crop_and_resize_3d(image, boxes, box_index, crop_size)                 core/models.py:663-664, 992-994
    crop_and_resize_3d_grad_image(grads, boxes, box_ind, image_size, T=, method_name=)
    crop_and_resize_3d_grad_boxes(grads, image, boxes, box_ind)
    non_max_suppression_3d(boxes, scores, max_output_size, iou_threshold)   core/models.py:453-455

This module exposes the same names with the same argument meaning, layouts
(boxes ``[N,6] = (y1,x1,z1,y2,x2,z2)`` normalized; volumes ``[B,H,W,D,C]``
channel-last float32) and the same validation messages as the reference ops'
``OP_REQUIRES`` checks (raised as :class:`InvalidArgumentError`).  TensorFlow is
not available in this image, so tensors are ``torch`` CUDA tensors (device
memory + streams only; every computation happens in the hand-written sm_100a
kernels behind the C ABI of ``include/roi3d.h``).  Host buffers (numpy arrays or
CPU tensors) are also accepted: they are copied host->device through pinned
memory, computed on the GPU and copied back -- the end-to-end path a CPU-op
drop-in sees.  There is no CPU fallback: without a CUDA device or without the
built library every call raises.

The TF-side registration that makes ``core/models.py`` load this library
unchanged is in ``tf_ops/``.
"""
import ctypes

import numpy as np
import torch

from . import _lib

__all__ = [
    "InvalidArgumentError", "crop_and_resize_3d", "crop_and_resize_3d_grad_image",
    "crop_and_resize_3d_grad_boxes", "non_max_suppression_3d", "CropAndResize3DFunction",
    "non_max_suppression_3d_batched", "non_max_suppression_3d_per_class", "non_max_suppression_3d_graph",
    "pyramid_roi_align_3d", "PyramidROIAlign3DFunction", "overlaps_3d", "decode_proposals", "top_k_set", "proposal_layer",
    "set_option", "get_option", "kernel_launches", "reset_kernel_launches", "deferred", "synchronize", "upload", "host_pipeline",
]

METHODS = {"trilinear": 0, "nearest": 1}


class InvalidArgumentError(ValueError):
    """Mirror of tf.errors.InvalidArgumentError raised by the reference ops' OP_REQUIRES."""


def _require(cond, msg):
    if not cond:
        raise InvalidArgumentError(msg)


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("roi3d_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t.numel() else 0)


def _shape(x):
    """Shape of a torch tensor / numpy array / nested list without touching the device."""
    return tuple(x.shape) if hasattr(x, "shape") else tuple(np.asarray(x).shape)


class _HostPipe:
    """Streams of the host-buffer path: H2D copies, kernels and D2H copies run on three streams chained by
    events, so the upload of one op's inputs overlaps the download of the previous op's result (PCIe is
    full-duplex) and both overlap the kernels."""

    _pipes = {}

    def __init__(self, device):
        self.h2d = torch.cuda.Stream(device)
        self.d2h = torch.cuda.Stream(device)
        self.pending = []                  # events of D2H copies not yet waited for (deferred mode)
        self.deferred = 0
        self.held = []                     # (pinned result, device result) pairs whose download waits for the block's end

    @classmethod
    def get(cls, device):
        p = cls._pipes.get(device.index)
        if p is None:
            p = cls._pipes[device.index] = cls(device)
        return p


class deferred:
    """Context manager: inside it, calls on host buffers return their (pinned) result tensors / arrays
    immediately; the data is valid after the block exits (or after :func:`synchronize`).  Lets a step of
    many independent ops -- the 8 + 8 CropAndResize3D nodes of PyramidROIAlign -- keep both PCIe directions
    and the GPU busy, the way TF's executor overlaps independent nodes."""

    def __enter__(self):
        self.pipe = _HostPipe.get(_device())
        self.pipe.deferred += 1
        return self

    def __exit__(self, *exc):
        self.pipe.deferred -= 1
        if self.pipe.deferred == 0:
            pipe = self.pipe
            if pipe.held:                  # "uploads first" policy: every download of the block starts after its last upload
                dev = pipe.held[0][1].device
                pipe.d2h.wait_stream(pipe.h2d)
                pipe.d2h.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(pipe.d2h):
                    for res, out in pipe.held:
                        res.copy_(out, non_blocking=True)
                        out.record_stream(pipe.d2h)
                    ev = torch.cuda.Event()
                    ev.record(pipe.d2h)
                pipe.pending.append(ev)
                pipe.held.clear()
            synchronize()
        return False


_UPLOADS_FIRST = [False]


def host_pipeline(uploads_first=None):
    """Copy policy of `deferred` blocks.  Default (False): full duplex -- each result is downloaded as soon as its kernel
    has finished, while later inputs go up; right for one process per host link.  `uploads_first=True`: all downloads of
    a block start after its last upload -- an experiment for hosts shared by several ranks whose memory system gives more
    one-directional than duplex bandwidth (this pool's 8-GPU box: 233 GB/s up alone, 128 GB/s down alone, 82 + 82 GB/s
    duplex, profiles/pcie_probe.py); measured there at 174.0 vs 169.8 ms per 8-rank step, i.e. no gain.  Returns the
    policy in force."""
    if uploads_first is not None:
        _UPLOADS_FIRST[0] = bool(uploads_first)
    return _UPLOADS_FIRST[0]


def synchronize():
    """Wait for every outstanding host-buffer result of the current device."""
    pipe = _HostPipe.get(_device())
    for ev in pipe.pending:
        ev.synchronize()
    pipe.pending.clear()


class _Arg:
    """Brings one argument to the device; remembers whether the caller passed a host buffer."""

    __slots__ = ("dev", "host", "numpy")

    def __init__(self, x, dtype, device):
        self.numpy = isinstance(x, np.ndarray) or not isinstance(x, torch.Tensor)
        t = torch.as_tensor(x) if self.numpy else x
        self.host = t.device.type != "cuda"
        if self.host:
            if t.dtype != dtype:
                t = t.to(dtype)
            t = t.contiguous()
            if not t.is_pinned() and t.numel() * t.element_size() >= (1 << 20):
                t = t.pin_memory()                     # pageable -> pinned staging, then one async H2D
            if t.numel() * t.element_size() < (1 << 20):
                self.dev = t.to(device, non_blocking=True)     # small: not worth a second stream
            else:
                pipe = _HostPipe.get(device)
                cur = torch.cuda.current_stream(device)
                with torch.cuda.stream(pipe.h2d):
                    self.dev = t.to(device, non_blocking=True)
                cur.wait_stream(pipe.h2d)              # kernels start when the upload has landed
                self.dev.record_stream(cur)
        else:
            if t.dtype != dtype:
                t = t.to(dtype)
            self.dev = t.contiguous()


def upload(x, dtype=torch.float32):
    """Bring a host tensor / array to the current device through the host-buffer pipeline's upload stream, in order
    with the uploads of the op calls around it (uploads issued from different streams share the copy engine and delay
    each other); returns a device tensor that is ready on the current stream.  For inputs several op calls share --
    PyramidROIAlign's feature maps feed the 7^3 and the 14^3 crop."""
    return _Arg(x, dtype, _device()).dev


def _finish(out, host, as_numpy):
    """Return `out` where the caller's inputs lived (device stays device, host gets a pinned copy)."""
    if not host:
        return out
    pipe = _HostPipe.get(out.device)
    if not pipe.deferred and out.numel() * out.element_size() < (1 << 20):
        res = out.cpu()                                # small and synchronous: one blocking copy
        return res.numpy() if as_numpy else res
    res = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
    if pipe.deferred and _UPLOADS_FIRST[0]:
        pipe.held.append((res, out))
        return res.numpy() if as_numpy else res
    pipe.d2h.wait_stream(torch.cuda.current_stream(out.device))
    with torch.cuda.stream(pipe.d2h):
        res.copy_(out, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(pipe.d2h)
    out.record_stream(pipe.d2h)
    if pipe.deferred:
        pipe.pending.append(ev)
    else:
        ev.synchronize()
    return res.numpy() if as_numpy else res


# ---------------------------------------------------------------------------------------
# NonMaxSuppression3D
# ---------------------------------------------------------------------------------------
_ws_cache = {}


def _workspace(nbytes, device):
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def non_max_suppression_3d(boxes, scores, max_output_size, iou_threshold=0.5, name=None):
    """Greedy 3-D NMS; returns int32 ``[M]`` selected indices in selection order, M <= max_output_size.

    Mirrors REGISTER_OP("NonMaxSuppression3D") (NMS.so@0xe4e0): ``iou_threshold`` is an attr
    (python float), ``max_output_size`` a 0-D int.  Bit-exact with the reference (same IoU
    arithmetic, suppression on ``iou >= threshold``, ties -> lower index first).
    """
    del name
    bs, ss = _shape(boxes), _shape(scores)
    _require(len(bs) == 2, "boxes must be 2-D")
    _require(bs[1] == 6, "boxes must have 6 columns")
    _require(len(ss) == 1, "scores must be 1-D")
    _require(ss[0] == bs[0], "scores has incompatible shape")
    mos = np.asarray(max_output_size.cpu() if isinstance(max_output_size, torch.Tensor) else max_output_size)
    _require(mos.ndim == 0, "max_output_size must be 0-D, got shape %s" % (list(mos.shape),))
    thr = float(iou_threshold)
    _require(0.0 <= thr <= 1.0, "iou_threshold must be in [0, 1]")
    dev = _device()
    b, s = _Arg(boxes, torch.float32, dev), _Arg(scores, torch.float32, dev)
    n, max_out = int(b.dev.shape[0]), max(int(mos), 0)
    host = b.host or s.host
    lib = _lib.load()
    keep = torch.empty(max(max_out, 1), dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int32, pin_memory=True)
    if n > 0 and max_out > 0:
        nbytes = lib.roi3d_nms3d_workspace_bytes(n)
        ws = _workspace(nbytes, dev)
        _lib.check(lib.roi3d_nms3d(_ptr(b.dev), _ptr(s.dev), n, max_out, thr, _ptr(keep), _ptr(count),
                                   _ptr(ws), ws.numel(), _stream_ptr()))
        torch.cuda.current_stream().synchronize()      # output length is data dependent
    m = int(count[0])
    return _finish(keep[:m], host, b.numpy)


def non_max_suppression_3d_batched(boxes, scores, seg_offsets, max_output_size, iou_threshold=0.5):
    """`S` independent 3-D NMS problems in one set of launches.

    ``boxes [T,6]``, ``scores [T]`` hold the segments back to back; ``seg_offsets`` is an ascending int sequence of
    length S+1 (host list / numpy / tensor).  Returns a list of S int32 tensors with indices LOCAL to each segment,
    each bit-identical to :func:`non_max_suppression_3d` on that segment alone.  This is what replaces the
    per-image ``utils.batch_slice`` loop around the op in ProposalLayer (core/models.py:487-490) and a per-class
    loop in DetectionLayer.
    """
    bs, ss = _shape(boxes), _shape(scores)
    _require(len(bs) == 2, "boxes must be 2-D")
    _require(bs[1] == 6, "boxes must have 6 columns")
    _require(len(ss) == 1, "scores must be 1-D")
    _require(ss[0] == bs[0], "scores has incompatible shape")
    thr = float(iou_threshold)
    _require(0.0 <= thr <= 1.0, "iou_threshold must be in [0, 1]")
    offs = np.asarray(seg_offsets.cpu() if isinstance(seg_offsets, torch.Tensor) else seg_offsets, dtype=np.int64)
    _require(offs.ndim == 1 and len(offs) >= 1, "seg_offsets must be 1-D with at least one element")
    _require(offs[0] >= 0 and offs[-1] <= bs[0] and np.all(np.diff(offs) >= 0), "seg_offsets must ascend within [0, N]")
    S, max_out = len(offs) - 1, max(int(max_output_size), 0)
    dev = _device()
    b, s = _Arg(boxes, torch.float32, dev), _Arg(scores, torch.float32, dev)
    if S == 0:
        return []
    n_max = int(np.diff(offs).max())
    host = b.host or s.host
    lib = _lib.load()
    keep = torch.empty((S, max(max_out, 1)), dtype=torch.int32, device=dev)
    count = torch.zeros(S, dtype=torch.int32, pin_memory=True)
    if n_max > 0 and max_out > 0:
        d_offs = torch.as_tensor(offs.astype(np.int32)).to(dev, non_blocking=True)
        nbytes = lib.roi3d_nms3d_batched_workspace_bytes(n_max, S)
        ws = _workspace(nbytes, dev)
        _lib.check(lib.roi3d_nms3d_batched(_ptr(b.dev), _ptr(s.dev), _ptr(d_offs), S, n_max, max_out, thr, _ptr(keep),
                                           _ptr(count), _ptr(ws), ws.numel(), _stream_ptr()))
        torch.cuda.current_stream().synchronize()
    out = [keep[z, :int(count[z])] for z in range(S)]
    if host:
        out = [t.cpu() for t in out]
        if b.numpy:
            out = [t.numpy() for t in out]
    return out


def non_max_suppression_3d_per_class(boxes, scores, class_ids, max_output_size, iou_threshold=0.5):
    """Per-class 3-D NMS (the upstream DetectionLayer design, BASELINE cfg3): boxes of different classes never
    suppress each other.  Returns ``(keep, classes)``: for each class present (ascending id) the ORIGINAL indices
    kept, in selection order.  Grouping is a stable device sort by class id; the NMS itself is one batched call."""
    dev = _device()
    c = _Arg(class_ids, torch.int64, dev)
    b, s = _Arg(boxes, torch.float32, dev), _Arg(scores, torch.float32, dev)
    _require(c.dev.dim() == 1 and c.dev.shape[0] == b.dev.shape[0], "class_ids has incompatible shape")
    order = torch.sort(c.dev, stable=True).indices
    classes, counts = torch.unique_consecutive(c.dev[order], return_counts=True)
    offs = np.concatenate([[0], np.cumsum(counts.cpu().numpy())])
    kept = non_max_suppression_3d_batched(b.dev[order], s.dev[order], offs, max_output_size, iou_threshold)
    keep = [order[int(offs[z]):int(offs[z + 1])][k.long()].to(torch.int32) for z, k in enumerate(kept)]
    if b.host:
        keep = [k.cpu().numpy() if b.numpy else k.cpu() for k in keep]
    return keep, classes.cpu().tolist()


def non_max_suppression_3d_graph(boxes, scores, threshold, max_boxes):
    """Mirror of ``utils.non_max_suppression_3d_graph`` (core/utils.py:467-503): ``(selected_boxes [M,6],
    keep_indices [M])`` with python-scalar threshold / limit."""
    keep = non_max_suppression_3d(boxes, scores, int(max_boxes), float(threshold))
    if isinstance(keep, np.ndarray):
        return np.asarray(boxes, np.float32)[keep], keep
    src = boxes if isinstance(boxes, torch.Tensor) else torch.as_tensor(boxes)
    return src.to(torch.float32)[keep.long().to(src.device)], keep


# ---------------------------------------------------------------------------------------
# CropAndResize3D family
# ---------------------------------------------------------------------------------------
def _check_boxes(boxes, box_index):
    b, bi = _shape(boxes), _shape(box_index)
    _require(len(b) == 2, "boxes must be 2-D")
    _require(b[1] == 6, "boxes must have 6 columns")
    _require(len(bi) == 1, "box_index must be 1-D")
    _require(bi[0] == b[0], "box_index has incompatible shape")


def _method(method_name):
    _require(method_name in METHODS, "method must be 'trilinear' or 'nearest'")
    return METHODS[method_name]


def _crop_size(crop_size):
    cs = np.asarray(crop_size.cpu() if isinstance(crop_size, torch.Tensor) else crop_size)
    _require(cs.ndim == 1, "crop_size must be 1-D")
    _require(cs.shape[0] == 3, "crop_size must have three elements")
    ph, pw, pd = (int(v) for v in cs)
    _require(ph > 0 and pw > 0 and pd > 0, "crop dimensions must be positive")
    return ph, pw, pd


_car_ws_cache = {}


def _car_workspace(n, device):
    """Device scratch of the crop-and-resize calls (the ROI processing order, roi3d_car3d_workspace_bytes): one buffer
    per (device, stream), grown on demand; calls on one stream are stream-ordered, so they can share it."""
    nbytes = int(_lib.load().roi3d_car3d_workspace_bytes(int(n)))
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    buf = _car_ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _car_ws_cache[key] = buf
    return buf


def _fwd_device(image, boxes, box_index, crop, method, ext):
    B, H, W, D, C = image.shape
    n = boxes.shape[0]
    out = torch.empty((n,) + crop + (C,), dtype=torch.float32, device=image.device)
    lib = _lib.load()
    ws = _car_workspace(n, image.device)
    _lib.check(lib.roi3d_car3d_fwd_ws(_ptr(image), B, H, W, D, C, _ptr(boxes), _ptr(box_index), n,
                                      crop[0], crop[1], crop[2], method, float(ext), _ptr(out), _ptr(ws), ws.numel(), _stream_ptr()))
    return out


def _grad_image_device(grads, boxes, box_ind, image_size, method):
    B, H, W, D, C = image_size
    n, ph, pw, pd = grads.shape[:4]
    out = torch.empty((B, H, W, D, C), dtype=torch.float32, device=grads.device)
    lib = _lib.load()
    ws = _car_workspace(n, grads.device)
    _lib.check(lib.roi3d_car3d_grad_image_ws(_ptr(grads), _ptr(boxes), _ptr(box_ind), n, ph, pw, pd,
                                             B, H, W, D, C, method, _ptr(out), _ptr(ws), ws.numel(), _stream_ptr()))
    return out


def _grad_boxes_device(grads, image, boxes, box_ind):
    B, H, W, D, C = image.shape
    n, ph, pw, pd = grads.shape[:4]
    out = torch.zeros((n, 6), dtype=torch.float32, device=grads.device)
    lib = _lib.load()
    _lib.check(lib.roi3d_car3d_grad_boxes(_ptr(grads), _ptr(image), B, H, W, D, C, _ptr(boxes), _ptr(box_ind), n,
                                          ph, pw, pd, _ptr(out), _stream_ptr()))
    return out


class CropAndResize3DFunction(torch.autograd.Function):
    """Autograd wiring equal to ``_CropAndResize3DGrad`` (core/custom_op/custom_op.py:28-65):
    d/d image through CropAndResize3DGradImage, d/d boxes through CropAndResize3DGradBoxes
    (always the trilinear approximation), None for box_index and crop_size."""

    @staticmethod
    def forward(ctx, image, boxes, box_index, crop, method, ext):
        ctx.save_for_backward(image, boxes, box_index)
        ctx.method = method
        return _fwd_device(image, boxes, box_index, crop, method, ext)

    @staticmethod
    def backward(ctx, grad):
        image, boxes, box_index = ctx.saved_tensors
        grad = grad.contiguous()
        grad0 = grad1 = None
        if ctx.needs_input_grad[0]:
            grad0 = _grad_image_device(grad, boxes, box_index, tuple(image.shape), ctx.method)
        if ctx.needs_input_grad[1]:
            grad1 = _grad_boxes_device(grad, image, boxes, box_index)
        return grad0, grad1, None, None, None, None


def crop_and_resize_3d(image, boxes, box_index, crop_size, method_name="trilinear",
                       extrapolation_value=0.0, name=None):
    """Trilinear (or nearest) crop of ``image [B,H,W,D,C]`` by ``boxes [N,6]`` -> ``[N,ph,pw,pd,C]``.

    Mirrors REGISTER_OP("CropAndResize3D") (CAR.so@0x4370).  Differentiable w.r.t. ``image``
    and ``boxes`` when they are CUDA tensors requiring grad.
    """
    del name
    method = _method(method_name)
    ims = _shape(image)
    _require(len(ims) == 5, "input image must be 5-D")
    _check_boxes(boxes, box_index)
    crop = _crop_size(crop_size)
    _require(all(int(d) > 0 for d in ims[1:4]), "image dimensions must be positive")
    dev = _device()
    im, b, bi = _Arg(image, torch.float32, dev), _Arg(boxes, torch.float32, dev), _Arg(box_index, torch.int32, dev)
    host = im.host or b.host or bi.host
    if not host and (im.dev.requires_grad or b.dev.requires_grad):
        return CropAndResize3DFunction.apply(im.dev, b.dev, bi.dev, crop, method, float(extrapolation_value))
    out = _fwd_device(im.dev, b.dev, bi.dev, crop, method, extrapolation_value)
    return _finish(out, host, im.numpy)


def crop_and_resize_3d_grad_image(grads, boxes, box_ind, image_size, T=None, method_name="trilinear", name=None):
    """Gradient of :func:`crop_and_resize_3d` w.r.t. the image: ``[B,H,W,D,C]`` float32.

    Mirrors REGISTER_OP("CropAndResize3DGradImage") (GI.so@0x3a80).  ``T`` is accepted for
    signature compatibility; like the reference kernel only float32 is computed.
    """
    del name
    if T is not None and T not in (torch.float32, np.float32, "float32", "float"):
        raise InvalidArgumentError("CropAndResize3DGradImage: only T=float32 is implemented")
    method = _method(method_name)
    gs = _shape(grads)
    _require(len(gs) == 5, "grads image must be 5-D")
    _check_boxes(boxes, box_ind)
    isz = np.asarray(image_size.cpu() if isinstance(image_size, torch.Tensor) else image_size)
    _require(isz.ndim == 1, "image_size must be 1-D")
    _require(isz.shape[0] == 5, "image_size must have five elements")
    size = tuple(int(v) for v in isz)
    if gs[0] > 0:
        _require(all(int(d) > 0 for d in gs[1:4]), "grads dimensions must be positive")
    _require(all(d > 0 for d in size[1:4]), "image dimensions must be positive")
    _require(size[4] == gs[4], "image_size and grads are incompatible")
    _require(_shape(boxes)[0] == gs[0], "boxes and grads have incompatible shape")
    dev = _device()
    if gs[0] == 0 and not any(isinstance(x, torch.Tensor) and x.device.type == "cuda" for x in (grads, boxes, box_ind)):
        # no boxes, host buffers: the op's result is its zero-fill (GI.so@0x3ec5).  It is produced where the caller
        # wants it -- in (pinned) host memory -- instead of being memset on the device and carried over PCIe.
        res = torch.empty(size, dtype=torch.float32, pin_memory=True).zero_()
        return res.numpy() if isinstance(grads, np.ndarray) or not isinstance(grads, torch.Tensor) else res
    g, b, bi = _Arg(grads, torch.float32, dev), _Arg(boxes, torch.float32, dev), _Arg(box_ind, torch.int32, dev)
    host = g.host or b.host or bi.host
    out = _grad_image_device(g.dev, b.dev, bi.dev, size, method)
    return _finish(out, host, g.numpy)


def crop_and_resize_3d_grad_boxes(grads, image, boxes, box_ind, method_name="trilinear", name=None):
    """Gradient of :func:`crop_and_resize_3d` w.r.t. the boxes: ``[N,6]`` float32 (trilinear only).

    Mirrors REGISTER_OP("CropAndResize3DGradBoxes") (GB.so@0x3980).
    """
    del name
    _require(method_name == "trilinear", "method must be 'trilinear' or 'nearest'")
    gs, ims = _shape(grads), _shape(image)
    _require(len(gs) == 5, "grads image must be 5-D")
    _require(len(ims) == 5, "input image must be 5-D")
    _check_boxes(boxes, box_ind)
    if gs[0] > 0:
        _require(all(int(d) > 0 for d in gs[1:4]), "grads dimensions must be positive")
    _require(all(int(d) > 0 for d in ims[1:4]), "image dimensions must be positive")
    _require(ims[4] == gs[4], "image and grads depths are incompatible")
    _require(_shape(boxes)[0] == gs[0], "boxes and grads have incompatible shape")
    dev = _device()
    g, im = _Arg(grads, torch.float32, dev), _Arg(image, torch.float32, dev)
    b, bi = _Arg(boxes, torch.float32, dev), _Arg(box_ind, torch.int32, dev)
    host = g.host or im.host or b.host or bi.host
    out = _grad_boxes_device(g.dev, im.dev, b.dev, bi.dev)
    return _finish(out, host, g.numpy)


# ---------------------------------------------------------------------------------------
# fused PyramidROIAlign3D
# ---------------------------------------------------------------------------------------
def _pyr_args(feature_maps, image_shape):
    assert len(feature_maps) == 4, "feature_maps = [P2, P3, P4, P5]"
    shapes = (ctypes.c_int * 12)(*[int(d) for fm in feature_maps for d in fm.shape[1:4]])
    ptrs = (ctypes.c_void_p * 4)(*[fm.data_ptr() for fm in feature_maps])
    ishape = (ctypes.c_float * 3)(*[float(v) for v in image_shape])
    return ptrs, shapes, ishape


class PyramidROIAlign3DFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, boxes, image_shape, pool_shape, p2, p3, p4, p5):
        fms = [t.contiguous() for t in (p2, p3, p4, p5)]
        boxes = boxes.contiguous()
        B, R = boxes.shape[:2]
        C = fms[0].shape[4]
        out = torch.empty((B, R) + tuple(pool_shape) + (C,), dtype=torch.float32, device=boxes.device)
        ptrs, shapes, ishape = _pyr_args(fms, image_shape)
        ws = _car_workspace(B * R, boxes.device)
        _lib.check(_lib.load().roi3d_pyramid_roi_align_fwd_ws(ptrs, shapes, B, C, _ptr(boxes), R, ishape, pool_shape[0],
                                                              pool_shape[1], pool_shape[2], _ptr(out), _ptr(ws), ws.numel(), _stream_ptr()))
        ctx.save_for_backward(boxes)
        ctx.meta = (tuple(image_shape), tuple(pool_shape), [tuple(t.shape) for t in fms])
        return out

    @staticmethod
    def backward(ctx, grad):
        (boxes,) = ctx.saved_tensors
        image_shape, pool_shape, shapes = ctx.meta
        grad = grad.contiguous()
        B, R = boxes.shape[:2]
        gms = [torch.empty(sh, dtype=torch.float32, device=grad.device) for sh in shapes]
        ptrs, cshapes, ishape = _pyr_args(gms, image_shape)
        ws = _car_workspace(B * R, grad.device)
        _lib.check(_lib.load().roi3d_pyramid_roi_align_grad_ws(_ptr(grad), ptrs, cshapes, B, shapes[0][4], _ptr(boxes), R, ishape,
                                                               pool_shape[0], pool_shape[1], pool_shape[2], _ptr(ws), ws.numel(), _stream_ptr()))
        return (None, None, None) + tuple(gms)


def pyramid_roi_align_3d(boxes, image_shape, feature_maps, pool_shape, out_dtype=torch.float32):
    """Fused ``PyramidROIAlign(pool_shape)([boxes, image_meta, P2, P3, P4, P5])`` (core/models.py:604-685).

    ``boxes [B,R,6]`` normalized, ``image_shape = (H, W, D)`` of the input volume, ``feature_maps`` = four CUDA
    tensors ``[B,H_l,W_l,D_l,C]``.  Returns ``[B,R,ph,pw,pd,C]`` in the boxes' order; differentiable w.r.t. the
    feature maps (boxes are stop_gradient'ed upstream, core/models.py:660).  ``out_dtype=torch.float16`` writes the
    target files' ``rois_aligned`` payload directly (core/models.py:3613; inference only, bit-identical to
    ``pack_f16`` of the float32 result)."""
    dev = _device()
    for t in list(feature_maps) + [boxes]:
        if not (isinstance(t, torch.Tensor) and t.device.type == "cuda" and t.dtype == torch.float32):
            raise InvalidArgumentError("pyramid_roi_align_3d takes float32 CUDA tensors")
    _require(boxes.dim() == 3 and boxes.shape[2] == 6, "boxes must be [B, R, 6]")
    _require(all(fm.dim() == 5 and fm.shape[0] == boxes.shape[0] for fm in feature_maps), "feature maps must be [B,H,W,D,C]")
    if out_dtype == torch.float16:
        fms = [t.contiguous() for t in feature_maps]
        boxes = boxes.contiguous()
        B, R = boxes.shape[:2]
        C = fms[0].shape[4]
        ps = tuple(int(v) for v in pool_shape)
        out = torch.empty((B, R) + ps + (C,), dtype=torch.float16, device=dev)
        ptrs, shapes, ishape = _pyr_args(fms, image_shape)
        ws = _car_workspace(B * R, dev)
        _lib.check(_lib.load().roi3d_pyramid_roi_align_fwd_f16_ws(ptrs, shapes, B, C, _ptr(boxes), R, ishape, ps[0], ps[1], ps[2],
                                                                  _ptr(out), _ptr(ws), ws.numel(), _stream_ptr()))
        return out
    _require(out_dtype == torch.float32, "out_dtype must be float32 or float16")
    return PyramidROIAlign3DFunction.apply(boxes, tuple(image_shape), tuple(int(v) for v in pool_shape), *feature_maps)


# ---------------------------------------------------------------------------------------
# box-space helpers
# ---------------------------------------------------------------------------------------
def overlaps_3d(boxes1, boxes2):
    """``overlaps_graph`` (core/models.py:695-733): IoU matrix ``[N, M]`` float32 of two CUDA box sets."""
    dev = _device()
    b1, b2 = _Arg(boxes1, torch.float32, dev), _Arg(boxes2, torch.float32, dev)
    _require(b1.dev.dim() == 2 and b1.dev.shape[1] == 6 and b2.dev.dim() == 2 and b2.dev.shape[1] == 6, "boxes must be [N, 6]")
    n, m = b1.dev.shape[0], b2.dev.shape[0]
    out = torch.empty((n, m), dtype=torch.float32, device=dev)
    _lib.check(_lib.load().roi3d_overlaps3d(_ptr(b1.dev), n, _ptr(b2.dev), m, _ptr(out), _stream_ptr()))
    return _finish(out, b1.host or b2.host, b1.numpy)


def decode_proposals(anchors, deltas, std_dev, image_depth, index=None):
    """ProposalLayer's box front-end (core/models.py:397-447): ``deltas * std_dev`` -> clip +-3 ->
    ``apply_box_deltas_graph`` -> clip to [0,1] -> min sizes.  ``index`` (the ``tf.nn.top_k`` indices) gathers rows."""
    dev = _device()
    a, d = _Arg(anchors, torch.float32, dev), _Arg(deltas, torch.float32, dev)
    _require(a.dev.dim() == 2 and a.dev.shape[1] == 6 and tuple(d.dev.shape) == tuple(a.dev.shape), "anchors / deltas must be [N, 6]")
    ix = _Arg(index, torch.int32, dev).dev if index is not None else None
    n = int(ix.shape[0]) if ix is not None else int(a.dev.shape[0])
    out = torch.empty((n, 6), dtype=torch.float32, device=dev)
    std = (ctypes.c_float * 6)(*[float(v) for v in std_dev])
    _lib.check(_lib.load().roi3d_decode_proposals(_ptr(a.dev), _ptr(d.dev), _ptr(ix) if ix is not None else None, n, std,
                                                  float(image_depth), _ptr(out), _stream_ptr()))
    return _finish(out, a.host or d.host, a.numpy)


def top_k_set(scores, k):
    """The index SET of ``tf.nn.top_k(scores, k)`` (threshold ties -> lower indices), in ascending index order, and
    the matching scores -- device radix select, no host sync.  CUDA float32 ``scores [N]`` only."""
    dev = _device()
    s = _Arg(scores, torch.float32, dev)
    _require(s.dev.dim() == 1, "scores must be 1-D")
    n, k = int(s.dev.shape[0]), int(k)
    _require(0 <= k <= n, "k must be in [0, N]")
    idx = torch.empty(k, dtype=torch.int32, device=dev)
    val = torch.empty(k, dtype=torch.float32, device=dev)
    if k:
        lib = _lib.load()
        ws = _workspace(lib.roi3d_topk_workspace_bytes(n), dev)
        _lib.check(lib.roi3d_topk(_ptr(s.dev), n, k, _ptr(idx), _ptr(val), _ptr(ws), ws.numel(), _stream_ptr()))
    return idx, val


def proposal_layer(scores, deltas, anchors, std_dev, image_depth, pre_nms_limit, proposal_count, nms_threshold):
    """One image of ``ProposalLayer.call`` (core/models.py:382-500) entirely on the device and without a host
    synchronisation: top-k -> delta decode + clip + min sizes -> NMS3D -> gather -> zero-pad to ``proposal_count``.

    ``scores [N]`` (foreground probabilities), ``deltas [N,6]``, ``anchors [N,6]`` CUDA float32.  Returns
    ``proposals [proposal_count, 6]`` and the device int32 count of real proposals."""
    dev = _device()
    sc, dl, an = _Arg(scores, torch.float32, dev), _Arg(deltas, torch.float32, dev), _Arg(anchors, torch.float32, dev)
    n = int(sc.dev.shape[0])
    P = int(proposal_count)
    lib = _lib.load()
    out = torch.empty((P, 6), dtype=torch.float32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    nbytes = lib.roi3d_proposal_layer_workspace_bytes(n, int(pre_nms_limit), P)
    key = ("proposal", dev.index, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = _ws_cache[key] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    std = (ctypes.c_float * 6)(*[float(v) for v in std_dev])
    _lib.check(lib.roi3d_proposal_layer(_ptr(sc.dev), _ptr(dl.dev), _ptr(an.dev), n, std, float(image_depth),
                                        int(pre_nms_limit), P, float(nms_threshold), _ptr(out), _ptr(count), _ptr(ws),
                                        ws.numel(), _stream_ptr()))
    return out, count


# ---------------------------------------------------------------------------------------
# DetectionLayer, mask targets and the target files' payloads
# ---------------------------------------------------------------------------------------
NMS_MODES = {"reference_2d": 0, "3d": 1}


def refine_detections(rois, probs, deltas, image_shape, detection_min_confidence, detection_nms_threshold,
                      bbox_std_dev=None, detection_max_instances=None, return_counts=False, nms_mode="reference_2d"):
    """``refine_detections_graph`` (core/models.py:1415-1524) / ``DetectionLayer.call`` (:1552-1575) on the device.

    ``rois [R,6]`` / ``probs [R,K]`` / ``deltas [R,K,6]`` for one image, or with a leading batch axis for the whole
    layer (one set of launches, no ``batch_slice`` loop, no host sync).  ``image_shape`` = (H, W, D) in pixels.
    Returns ``detections [max_instances, 8]`` (or ``[B, max_instances, 8]``) =
    ``(y1,x1,z1,y2,x2,z2,class_id,score)`` normalised, zero padded.

    ``nms_mode="reference_2d"`` (default) is the graph's own NMS: ``tf.image.non_max_suppression`` on the (y, x)
    projection, suppressing on IoU > threshold (:1496-1501).  ``nms_mode="3d"`` runs the 3-D op instead (IoU over the
    volume, >= threshold): the upstream design ``utils.non_max_suppression_3d_graph`` exists for; opt-in."""
    _require(nms_mode in NMS_MODES, "nms_mode must be 'reference_2d' or '3d'")
    dev = _device()
    r, p, d = _Arg(rois, torch.float32, dev), _Arg(probs, torch.float32, dev), _Arg(deltas, torch.float32, dev)
    single = r.dev.dim() == 2
    _require(r.dev.dim() in (2, 3) and r.dev.shape[-1] == 6, "rois must be [R, 6] or [B, R, 6]")
    _require(p.dev.dim() == r.dev.dim() and tuple(p.dev.shape[:-1]) == tuple(r.dev.shape[:-1]), "probs has incompatible shape")
    K = int(p.dev.shape[-1])
    _require(K >= 2, "probs needs a foreground class column")
    _require(tuple(d.dev.shape) == tuple(r.dev.shape[:-1]) + (K, 6), "deltas has incompatible shape")
    B, R = (1, int(r.dev.shape[0])) if single else (int(r.dev.shape[0]), int(r.dev.shape[1]))
    std = [0.1, 0.1, 0.1, 0.2, 0.2, 0.2] if bbox_std_dev is None else [float(v) for v in bbox_std_dev]
    M = 200 if detection_max_instances is None else int(detection_max_instances)
    thr = float(detection_nms_threshold)
    _require(0.0 <= thr <= 1.0, "iou_threshold must be in [0, 1]")
    lib = _lib.load()
    det = torch.empty((B, M, 8), dtype=torch.float32, device=dev)
    cnt = torch.zeros(B, dtype=torch.int32, device=dev)
    if B and M:
        ws = torch.empty(lib.roi3d_refine_detections_workspace_bytes(B, R, M), dtype=torch.uint8, device=dev)
        shp = (ctypes.c_float * 3)(*[float(v) for v in list(image_shape)[:3]])
        _lib.check(lib.roi3d_refine_detections(_ptr(r.dev), _ptr(p.dev), _ptr(d.dev), B, R, K, shp, (ctypes.c_float * 6)(*std),
                                               float(detection_min_confidence), thr, NMS_MODES[nms_mode], M, _ptr(det), _ptr(cnt), _ptr(ws),
                                               ws.numel(), _stream_ptr()))
    out = det[0] if single else det
    host = r.host or p.host or d.host
    out = _finish(out, host, r.numpy)
    if return_counts:
        return out, _finish(cnt, host, r.numpy)
    return out


def mask_targets(gt_masks, boxes, assignment, mask_shape, packed=False):
    """``detection_targets_graph._get_masks`` (core/models.py:972-1005) in one kernel: the ground-truth mask assigned
    to each positive ROI (``gt_masks [G,H,W,D]``, float32 or uint8/bool; ``assignment`` int32 ``[N]`` or None for
    ``range(N)``) is cropped to ``boxes [N,6]`` at ``mask_shape`` and rounded.  Returns float32 ``[N,mh,mw,md]``;
    with ``packed=True`` also the ``numpy.packbits`` payload of the target file (uint8 ``[ceil(N*mh*mw*md/8)]``)."""
    dev = _device()
    g = gt_masks if isinstance(gt_masks, torch.Tensor) else torch.as_tensor(np.asarray(gt_masks))
    if g.dtype == torch.bool:
        g = g.to(torch.uint8)
    m = _Arg(g, torch.uint8 if g.dtype == torch.uint8 else torch.float32, dev)
    b = _Arg(boxes, torch.float32, dev)
    _require(m.dev.dim() == 4, "gt_masks must be [G, H, W, D]")
    _require(b.dev.dim() == 2 and b.dev.shape[1] == 6, "boxes must have 6 columns")
    a = _Arg(assignment, torch.int32, dev).dev if assignment is not None else None
    n = int(b.dev.shape[0])
    _require(a is None or tuple(a.shape) == (n,), "assignment has incompatible shape")
    mh, mw, md = _crop_size(mask_shape)
    G, H, W, D = (int(v) for v in m.dev.shape)
    out = torch.empty((n, mh, mw, md), dtype=torch.float32, device=dev)
    bits = torch.empty((n * mh * mw * md + 7) // 8, dtype=torch.uint8, device=dev) if packed else None
    _lib.check(_lib.load().roi3d_mask_targets(_ptr(m.dev), 1 if m.dev.dtype == torch.uint8 else 0, G, H, W, D, _ptr(b.dev),
                                              _ptr(a) if a is not None else None, n, mh, mw, md, _ptr(out),
                                              _ptr(bits) if packed else None, _stream_ptr()))
    host = m.host or b.host
    res = _finish(out, host, b.numpy)
    return (res, _finish(bits, host, b.numpy)) if packed else res


def pack_f16(x):
    """``x.astype(np.float16)`` on the device (round to nearest even): the ``rois_aligned`` payload, core/models.py:3613."""
    dev = _device()
    a = _Arg(x, torch.float32, dev)
    out = torch.empty(a.dev.shape, dtype=torch.float16, device=dev)
    _lib.check(_lib.load().roi3d_pack_f16(_ptr(a.dev), a.dev.numel(), _ptr(out), _stream_ptr()))
    return _finish(out, a.host, a.numpy)


def unpack_f16(x):
    """float16 -> float32 on the device (the reader's ``astype(np.float32)``, core/data_generators.py:245)."""
    dev = _device()
    a = _Arg(x, torch.float16, dev)
    out = torch.empty(a.dev.shape, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().roi3d_unpack_f16(_ptr(a.dev), a.dev.numel(), _ptr(out), _stream_ptr()))
    return _finish(out, a.host, a.numpy)


def pack_bits(x):
    """``_bitpack`` (core/models.py:3585-3595): ``(numpy.packbits((x > 0.5).reshape(-1)), shape int32)``."""
    dev = _device()
    a = _Arg(x, torch.float32, dev)
    n = a.dev.numel()
    bits = torch.empty((n + 7) // 8, dtype=torch.uint8, device=dev)
    _lib.check(_lib.load().roi3d_pack_bits(_ptr(a.dev), n, _ptr(bits), _stream_ptr()))
    return _finish(bits, a.host, a.numpy), np.array(tuple(a.dev.shape), dtype=np.int32)


def unpack_bits(bits, shape):
    """``_unbit`` (core/data_generators.py:1908-1921): packed uint8 -> float32 0/1 array of ``shape`` on the device."""
    dev = _device()
    a = _Arg(bits, torch.uint8, dev)
    shape = tuple(int(v) for v in shape)
    n = int(np.prod(shape)) if shape else 1
    _require(a.dev.numel() * 8 >= n, "bits too short for shape")
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().roi3d_unpack_bits(_ptr(a.dev), n, _ptr(out), _stream_ptr()))
    return _finish(out, a.host, a.numpy)


# ---------------------------------------------------------------------------------------
# tuning / introspection passthroughs
# ---------------------------------------------------------------------------------------
def set_option(name, value):
    _lib.check(_lib.load().roi3d_set_option(name.encode(), int(value)))


def get_option(name):
    v = ctypes.c_int(0)
    _lib.check(_lib.load().roi3d_get_option(name.encode(), ctypes.byref(v)))
    return v.value


def kernel_launches():
    return int(_lib.load().roi3d_kernel_launches())


def reset_kernel_launches():
    _lib.load().roi3d_reset_kernel_launches()
