"""The TensorFlow registration source (3d-mask-r-cnn_b200/tf_ops/roi3d_tf_ops.cc) cannot be compiled here (no TF);
what can be checked is that every REGISTER_OP signature string and validation message in it is byte-identical to the
ones inside the reference's shared objects (.rodata of the wheel's libraries)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "3d-mask-r-cnn_b200", "tf_ops", "roi3d_tf_ops.cc")
LIBS = {"CropAndResize3D": "_crop_and_resize_3d_ops.so", "CropAndResize3DGradImage": "_crop_and_resize_3d_grad_image_ops.so",
        "CropAndResize3DGradBoxes": "_crop_and_resize_3d_grad_boxes_ops.so", "NonMaxSuppression3D": "_non_max_suppression_3d_ops.so"}


def _ref_bytes(name):
    path = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(path):
        try:
            from oracle import refrun
            refrun.build()
        except Exception as e:  # noqa: BLE001
            pytest.skip("reference libraries unavailable: %s" % e)
    return open(path, "rb").read()


def test_register_op_strings_match_the_wheel():
    src = open(SRC).read()
    blocks = re.findall(r'REGISTER_OP\("(\w+)"\)(.*?)\.SetShapeFn', src, re.S)
    assert sorted(b[0] for b in blocks) == sorted(LIBS)
    for op, body in blocks:
        blob = _ref_bytes(LIBS[op])
        assert op.encode() + b"\0" in blob
        sigs = re.findall(r'\.(?:Input|Output|Attr)\("([^"]+)"\)', body)
        assert len(sigs) >= 4
        for sig in sigs:
            assert sig.encode() + b"\0" in blob, (op, sig)


def test_validation_messages_match_the_wheel():
    src = open(SRC).read()
    msgs = set(re.findall(r'InvalidArgument\("([^"]+)"', src))
    blob = b"".join(_ref_bytes(n) for n in LIBS.values())
    ours_only = {"NonMaxSuppression3D: stream sync failed"}
    for m in msgs - ours_only:
        assert m.encode() in blob, m


def test_package_shims_export_reference_names():
    pk = os.path.join(ROOT, "3d-mask-r-cnn_b200", "tf_ops", "packages")
    for pkg, alias in (("crop_and_resize_3d", "crop_and_resize3d"), ("crop_and_resize_3d_grad_image", "crop_and_resize3d_grad_image"),
                       ("crop_and_resize_3d_grad_boxes", "crop_and_resize3d_grad_boxes"), ("non_max_suppression_3d", "non_max_suppression3d")):
        text = open(os.path.join(pk, pkg + ".py")).read()
        assert "%s = _ops.%s" % (pkg, alias) in text


@pytest.mark.parametrize("single_op", [None, 1, 2, 3, 4])
def test_tf_shim_compiles_against_stub_headers(single_op):
    """g++ -fsyntax-only of the TensorFlow registration against tests/tf_stub (a minimal stand-in for the TF op API with
    current-TF signatures): removed aliases (tensorflow::OkStatus, tensorflow::int64), typos and wrong argument counts
    to the C ABI of include/roi3d.h fail here.  ROI3D_TF_SINGLE_OP = 1..4 are the per-op libraries of the ppc64le loader."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    cuda_inc = next((p for p in ("/usr/local/cuda/include",) if os.path.exists(os.path.join(p, "cuda_runtime_api.h"))), None)
    if cuda_inc is None:
        pytest.skip("no CUDA headers")
    cmd = [gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror=return-type", "-I" + os.path.join(ROOT, "tests", "tf_stub"),
           "-I" + os.path.join(ROOT, "include"), "-I" + cuda_inc, "-DGOOGLE_CUDA=1", SRC]
    if single_op:
        cmd.insert(-1, "-DROI3D_TF_SINGLE_OP=%d" % single_op)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-3000:]
    src = open(SRC).read()
    assert "tf::OkStatus" not in src and "tf::int64" not in src          # gone from current TensorFlow
