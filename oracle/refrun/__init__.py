"""oracle.refrun -- the reference's OWN compiled ops, executed here without TensorFlow.

TEST INFRASTRUCTURE ONLY (see oracle/roi3d_oracle.c for the rules).  ``refrun.cc`` maps the four
shared objects of the reference wheel
(/root/reference/core/custom_op/tensorflow_nms_car_3d-0.1.0-cp36-cp36m-linux_x86_64.whl) into
memory, resolves their ~25 TensorFlow call-outs to stand-ins and calls the reference's own
``Compute`` functions.  Used to (1) pin the C oracle bit-for-bit against the real reference
(tests/test_oracle_pin.py, tests/golden/make_golden.py --check-ref) and (2) time the real
reference on the GPU box's host cores (bench.py ``cpu_baseline.kind == "reference"``).

Files: ``build()`` writes only into ``oracle/_ref/`` (git-ignored, not gpurun-ignored): the
runner ``librefrun.so`` and a byte-for-byte extraction of the wheel's four ``.so`` files, so the
GPU box -- which has no /root/reference -- can run them too.  Nothing of the reference is
committed.
"""
import ctypes
import os
import subprocess
import zipfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(os.path.dirname(_HERE), "_ref")
WHEEL = "/root/reference/core/custom_op/tensorflow_nms_car_3d-0.1.0-cp36-cp36m-linux_x86_64.whl"
MEMBERS = {
    "car": "crop_and_resize_3d/python/ops/_crop_and_resize_3d_ops.so",
    "gi": "crop_and_resize_3d_grad_image/python/ops/_crop_and_resize_3d_grad_image_ops.so",
    "gb": "crop_and_resize_3d_grad_boxes/python/ops/_crop_and_resize_3d_grad_boxes_ops.so",
    "nms": "non_max_suppression_3d/python/ops/_non_max_suppression_3d_ops.so",
}
RUNNER = os.path.join(REF_DIR, "librefrun.so")
# Provenance pin: SHA-256 of the four members of the reference wheel.  Anything else in oracle/_ref/ is refused --
# this package executes those files' machine code in-process, so they must be exactly the reference's own binaries.
PINNED_SHA256 = {
    "car": "5e37f9b340e19be8665dea48009474d094c1e2f621ce7cdd997ae2edaec60fc0",
    "gi": "59aa4c2314a15ecfaab139a6fa4675e7c8488ac07b9590f77b334a1cbfaf52ae",
    "gb": "3b7b24cee72c711788d41d19540cf2fffd29c296e23b498855393beb821fc24f",
    "nms": "bc5da3c474b8a9775bae8bed45ba071f898e9a0660e6ba6bfbeed4d2354ae40e",
}


def enabled():
    """ROI3D_REFRUN=0 switches the reference runner off everywhere (tests skip, bench.py falls back to the C port)."""
    return os.environ.get("ROI3D_REFRUN", "1").lower() not in ("0", "no", "off", "false")


def verify():
    """Raise unless every library in oracle/_ref/ is byte-identical (SHA-256) to the pinned reference wheel member."""
    import hashlib
    for key, path in lib_paths().items():
        with open(path, "rb") as f:
            got = hashlib.sha256(f.read()).hexdigest()
        if got != PINNED_SHA256[key]:
            raise RuntimeError("refrun: %s does not match the pinned reference binary (sha256 %s); refusing to load it" % (path, got))


def harden_process():
    """For the dedicated child processes that run the reference (bench.py's CPU baseline): no new privileges, no core
    dumps, no file writes, own network namespace when the kernel allows it.  Best effort, never fatal."""
    import resource
    for lim, val in ((resource.RLIMIT_CORE, 0), (resource.RLIMIT_FSIZE, 0)):
        try:
            resource.setrlimit(lim, (val, val))
        except (ValueError, OSError):
            pass
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        libc.prctl(38, 1, 0, 0, 0)                     # PR_SET_NO_NEW_PRIVS
        libc.unshare(0x40000000)                       # CLONE_NEWNET: the child has no network at all
    except Exception:  # noqa: BLE001
        pass
_f32p, _i32p = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)


def lib_paths():
    return {k: os.path.join(REF_DIR, os.path.basename(v)) for k, v in MEMBERS.items()}


def available():
    return enabled() and os.path.exists(RUNNER) and all(os.path.exists(p) for p in lib_paths().values())


def build(force=False):
    """Extract the wheel's libraries (when /root/reference is present) and compile the runner."""
    os.makedirs(REF_DIR, exist_ok=True)
    paths = lib_paths()
    if os.path.exists(WHEEL):
        with zipfile.ZipFile(WHEEL) as z:
            for key, member in MEMBERS.items():
                if force or not os.path.exists(paths[key]):
                    with open(paths[key], "wb") as f:
                        f.write(z.read(member))
    missing = [p for p in paths.values() if not os.path.exists(p)]
    if missing:
        raise RuntimeError("reference libraries unavailable (no wheel at %s and no extraction in %s)" % (WHEEL, REF_DIR))
    verify()
    src = os.path.join(_HERE, "refrun.cc")
    if force or not os.path.exists(RUNNER) or os.path.getmtime(src) > os.path.getmtime(RUNNER):
        cmd = ["g++", "-O1", "-g", "-std=c++17", "-shared", "-fPIC", "-D_GLIBCXX_USE_CXX11_ABI=0",
               "-fvisibility=hidden", "-o", RUNNER, src, "-ldl"]
        env = dict(os.environ)
        env.pop("CXX", None)
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError("g++ failed for refrun.cc:\n" + res.stdout)
    return RUNNER


class Reference:
    """numpy front-end with the same signatures as the ``oracle`` package."""

    def __init__(self, lib):
        self._lib = lib

    def _check(self, rc):
        if rc < 0:
            raise RuntimeError("refrun: " + self._lib.refrun_last_error().decode())
        return rc

    def non_max_suppression_3d(self, boxes, scores, max_output_size, iou_threshold=0.5):
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        scores = np.ascontiguousarray(scores, np.float32).reshape(-1)
        cap = max(int(max_output_size), 1)
        out = np.empty(cap, np.int32)
        m = self._check(self._lib.refrun_nms3d(boxes.ctypes.data_as(_f32p), scores.ctypes.data_as(_f32p), len(boxes),
                                               int(max_output_size), float(iou_threshold), out.ctypes.data_as(_i32p), cap))
        return out[:m].copy()

    def iou_pairs(self, boxes, ii, jj):
        """The reference's IOU<float> (NMS.so@0xb500) on index pairs."""
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        ii = np.ascontiguousarray(ii, np.int32)
        jj = np.ascontiguousarray(jj, np.int32)
        out = np.empty(len(ii), np.float32)
        self._check(self._lib.refrun_iou_pairs(boxes.ctypes.data_as(_f32p), len(boxes), ii.ctypes.data_as(_i32p),
                                               jj.ctypes.data_as(_i32p), len(ii), out.ctypes.data_as(_f32p)))
        return out

    def crop_and_resize_3d(self, image, boxes, box_index, crop_size, method_name="trilinear", extrapolation_value=0.0):
        image = np.ascontiguousarray(image, np.float32)
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        box_index = np.ascontiguousarray(box_index, np.int32).reshape(-1)
        B, H, W, D, C = image.shape
        ph, pw, pd = (int(v) for v in crop_size)
        out = np.empty((len(boxes), ph, pw, pd, C), np.float32)
        self._check(self._lib.refrun_car3d_fwd(image.ctypes.data_as(_f32p), B, H, W, D, C, boxes.ctypes.data_as(_f32p),
                                               box_index.ctypes.data_as(_i32p), len(boxes), ph, pw, pd,
                                               {"trilinear": 0, "nearest": 1}[method_name], float(extrapolation_value),
                                               out.ctypes.data_as(_f32p)))
        return out

    def crop_and_resize_3d_grad_image(self, grads, boxes, box_ind, image_size, method_name="trilinear"):
        grads = np.ascontiguousarray(grads, np.float32)
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        box_ind = np.ascontiguousarray(box_ind, np.int32).reshape(-1)
        B, H, W, D, C = (int(v) for v in image_size)
        n, ph, pw, pd, _ = grads.shape
        out = np.empty((B, H, W, D, C), np.float32)
        self._check(self._lib.refrun_car3d_grad_image(grads.ctypes.data_as(_f32p), boxes.ctypes.data_as(_f32p),
                                                      box_ind.ctypes.data_as(_i32p), n, ph, pw, pd, B, H, W, D, C,
                                                      {"trilinear": 0, "nearest": 1}[method_name], out.ctypes.data_as(_f32p)))
        return out

    def crop_and_resize_3d_grad_boxes(self, grads, image, boxes, box_ind):
        grads = np.ascontiguousarray(grads, np.float32)
        image = np.ascontiguousarray(image, np.float32)
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        box_ind = np.ascontiguousarray(box_ind, np.int32).reshape(-1)
        B, H, W, D, C = image.shape
        n, ph, pw, pd, _ = grads.shape
        out = np.empty((n, 6), np.float32)
        self._check(self._lib.refrun_car3d_grad_boxes(grads.ctypes.data_as(_f32p), image.ctypes.data_as(_f32p), B, H, W, D, C,
                                                      boxes.ctypes.data_as(_f32p), box_ind.ctypes.data_as(_i32p), n, ph, pw, pd,
                                                      out.ctypes.data_as(_f32p)))
        return out


_ref = None


def load():
    """Build if needed, map the reference libraries and return a :class:`Reference`."""
    global _ref
    if not enabled():
        raise RuntimeError("refrun disabled (ROI3D_REFRUN=0)")
    if _ref is None:
        build()
        verify()
        lib = ctypes.CDLL(RUNNER, mode=ctypes.RTLD_GLOBAL)      # its stand-in symbols must be visible to dlsym
        lib.refrun_last_error.restype = ctypes.c_char_p
        lib.refrun_open.argtypes = [ctypes.c_char_p] * 4
        i, f = ctypes.c_int, ctypes.c_float
        lib.refrun_car3d_fwd.argtypes = [_f32p, i, i, i, i, i, _f32p, _i32p, i, i, i, i, i, f, _f32p]
        lib.refrun_car3d_grad_image.argtypes = [_f32p, _f32p, _i32p, i, i, i, i, i, i, i, i, i, i, _f32p]
        lib.refrun_car3d_grad_boxes.argtypes = [_f32p, _f32p, i, i, i, i, i, _f32p, _i32p, i, i, i, i, _f32p]
        lib.refrun_nms3d.argtypes = [_f32p, _f32p, i, i, f, _i32p, i]
        lib.refrun_iou_pairs.argtypes = [_f32p, i, _i32p, _i32p, i, _f32p]
        p = lib_paths()
        rc = lib.refrun_open(p["car"].encode(), p["gi"].encode(), p["gb"].encode(), p["nms"].encode())
        if rc != 0:
            raise RuntimeError("refrun_open failed (%d): %s" % (rc, lib.refrun_last_error().decode()))
        _ref = Reference(lib)
    return _ref
