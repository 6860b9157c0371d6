"""The C-ABI library loads without a GPU and exports exactly what include/roi3d.h declares;
argument validation returns error codes without touching the device."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    import roi3d_b200
    return roi3d_b200._lib.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "roi3d.h")).read()
    return sorted(set(re.findall(r"ROI3D_API[^;(]*?\b(roi3d_\w+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", lib._name], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (roi3d_\w+)", out))
    assert exported == set(names)          # nothing undeclared leaks out, nothing declared is missing


def test_library_is_sm100a_only(lib):
    out = subprocess.run(["cuobjdump", "-lelf", lib._name], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, out


def test_version_and_errors(lib):
    assert b"sm_100a" in lib.roi3d_version()
    assert lib.roi3d_strerror(0) == b"ok"
    assert b"invalid" in lib.roi3d_strerror(-1)
    assert lib.roi3d_nms3d_workspace_bytes(0) > 0
    n = 6000
    need = lib.roi3d_nms3d_workspace_bytes(n)
    assert need >= n * ((n + 31) // 32) * 4


def test_argument_validation_without_gpu(lib):
    # every call below must be rejected before any CUDA work
    assert lib.roi3d_nms3d(None, None, -1, 10, 0.5, None, None, None, 0, None) == -1
    cnt = (ctypes.c_int * 1)()
    assert lib.roi3d_nms3d(None, None, 5, 10, 1.5, None, cnt, None, 0, None) == -1        # iou_threshold range
    assert lib.roi3d_nms3d(None, None, 5, 10, float("nan"), None, cnt, None, 0, None) == -1
    assert lib.roi3d_nms3d(None, None, 5, 10, 0.5, None, cnt, None, 0, None) == -1        # NULL boxes
    buf = np.zeros(64, np.float32).ctypes.data_as(ctypes.c_void_p)
    assert lib.roi3d_nms3d(buf, buf, 5, 10, 0.5, buf, cnt, None, 0, None) == -2           # no workspace
    assert lib.roi3d_car3d_fwd(buf, 1, 2, 2, 2, 1, buf, buf, 1, 2, 2, 2, 7, 0.0, buf, None) == -1   # bad method
    assert lib.roi3d_car3d_fwd(buf, 1, 0, 2, 2, 1, buf, buf, 1, 2, 2, 2, 0, 0.0, buf, None) == -1   # bad dims
    assert lib.roi3d_car3d_fwd(buf, 1, 2, 2, 2, 1, buf, buf, 0, 2, 2, 2, 0, 0.0, buf, None) == 0    # n == 0 is a no-op
    assert lib.roi3d_car3d_fwd(None, 1, 2, 2, 2, 1, buf, buf, 1, 2, 2, 2, 0, 0.0, buf, None) == -1
    assert lib.roi3d_car3d_fwd(buf, 1, 4096, 4096, 512, 256, buf, buf, 1, 2, 2, 2, 0, 0.0, buf, None) == -3
    assert lib.roi3d_car3d_grad_image(buf, buf, buf, 1, 2, 2, 2, 1, 2, 2, 2, 1, 0, None, None) == -1
    assert lib.roi3d_car3d_grad_boxes(buf, buf, 1, 2, 2, 2, 1, buf, buf, 0, 2, 2, 2, buf, None) == 0
    assert lib.roi3d_set_option(b"no_such_option", 1) == -1
    assert lib.roi3d_set_option(b"car_fwd_variant", 0) == 0


def test_host_api_validation_messages():
    """Same messages as the reference ops' OP_REQUIRES (strings in the wheel's .rodata)."""
    import roi3d_b200 as rb
    z = np.zeros
    cases = [
        (lambda: rb.non_max_suppression_3d(z((3, 6, 1)), z(3), 1, 0.5), "boxes must be 2-D"),
        (lambda: rb.non_max_suppression_3d(z((3, 4)), z(3), 1, 0.5), "boxes must have 6 columns"),
        (lambda: rb.non_max_suppression_3d(z((3, 6)), z((3, 1)), 1, 0.5), "scores must be 1-D"),
        (lambda: rb.non_max_suppression_3d(z((3, 6)), z(4), 1, 0.5), "scores has incompatible shape"),
        (lambda: rb.non_max_suppression_3d(z((3, 6)), z(3), [1, 2], 0.5), "max_output_size must be 0-D"),
        (lambda: rb.non_max_suppression_3d(z((3, 6)), z(3), 1, 1.5), "iou_threshold must be in [0, 1]"),
        (lambda: rb.crop_and_resize_3d(z((1, 2, 2, 2)), z((1, 6)), [0], (2, 2, 2)), "input image must be 5-D"),
        (lambda: rb.crop_and_resize_3d(z((1, 2, 2, 2, 1)), z((1, 5)), [0], (2, 2, 2)), "boxes must have 6 columns"),
        (lambda: rb.crop_and_resize_3d(z((1, 2, 2, 2, 1)), z((1, 6)), [[0]], (2, 2, 2)), "box_index must be 1-D"),
        (lambda: rb.crop_and_resize_3d(z((1, 2, 2, 2, 1)), z((1, 6)), [0, 0], (2, 2, 2)), "box_index has incompatible shape"),
        (lambda: rb.crop_and_resize_3d(z((1, 2, 2, 2, 1)), z((1, 6)), [0], (2, 2)), "crop_size must have three elements"),
        (lambda: rb.crop_and_resize_3d(z((1, 2, 2, 2, 1)), z((1, 6)), [0], (2, 0, 2)), "crop dimensions must be positive"),
        (lambda: rb.crop_and_resize_3d(z((1, 2, 2, 2, 1)), z((1, 6)), [0], (2, 2, 2), method_name="cubic"),
         "method must be 'trilinear' or 'nearest'"),
        (lambda: rb.crop_and_resize_3d_grad_image(z((1, 2, 2, 2)), z((1, 6)), [0], (1, 2, 2, 2, 1)), "grads image must be 5-D"),
        (lambda: rb.crop_and_resize_3d_grad_image(z((1, 2, 2, 2, 1)), z((1, 6)), [0], (1, 2, 2, 2)), "image_size must have five elements"),
        (lambda: rb.crop_and_resize_3d_grad_image(z((1, 2, 2, 2, 1)), z((1, 6)), [0], (1, 2, 2, 2, 3)), "image_size and grads are incompatible"),
        (lambda: rb.crop_and_resize_3d_grad_boxes(z((1, 2, 2, 2, 1)), z((1, 2, 2, 2, 3)), z((1, 6)), [0]),
         "image and grads depths are incompatible"),
        (lambda: rb.crop_and_resize_3d_grad_boxes(z((2, 2, 2, 2, 1)), z((1, 2, 2, 2, 1)), z((1, 6)), [0]),
         "boxes and grads have incompatible shape"),
    ]
    for fn, msg in cases:
        with pytest.raises(rb.InvalidArgumentError, match=re.escape(msg)):
            fn()


def test_no_cpu_fallback():
    """Without a CUDA device the product path raises instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import roi3d_b200 as rb
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rb.non_max_suppression_3d(np.zeros((3, 6), np.float32), np.zeros(3, np.float32), 1, 0.5)
    src = "".join(open(os.path.join(ROOT, "3d-mask-r-cnn_b200", f)).read()
                  for f in ("__init__.py", "custom_op.py", "_lib.py"))
    assert "import oracle" not in src and "from oracle" not in src
