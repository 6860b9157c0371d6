"""ctypes binding of libroi3d_b200.so (the C ABI declared in include/roi3d.h).

The library is built in-tree by :func:`build` (nvcc, sm_100a only) into
``3d-mask-r-cnn_b200/lib/``.  Importing the package without it raises -- there is
no CPU or eager fallback on this path (the reference's ProposalLayer swallows
exceptions at graph-build time, core/models.py:451-474, so a silent fallback
would degrade NMS to top-k unnoticed).
"""
import ctypes
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libroi3d_b200.so")
HEADER = os.path.join(_ROOT, "include", "roi3d.h")

SOURCES = ["roi3d_abi.cu", "roi3d_car_direct.cu", "roi3d_car_plane.cu", "roi3d_car_sep.cu", "roi3d_car_os.cu", "roi3d_nms.cu", "roi3d_boxes.cu",
           "roi3d_detect.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",      # Blackwell B200 only
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                                      # fp32 ops stay separately rounded (parity)
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]

EXPORTS = [
    "roi3d_version", "roi3d_strerror", "roi3d_last_cuda_error",
    "roi3d_nms3d_workspace_bytes", "roi3d_nms3d", "roi3d_nms3d_batched_workspace_bytes", "roi3d_nms3d_batched",
    "roi3d_car3d_fwd", "roi3d_car3d_grad_image", "roi3d_car3d_grad_boxes",
    "roi3d_car3d_workspace_bytes", "roi3d_car3d_fwd_ws", "roi3d_car3d_grad_image_ws",
    "roi3d_pyramid_roi_align_fwd", "roi3d_pyramid_roi_align_fwd_f16", "roi3d_pyramid_roi_align_grad",
    "roi3d_pyramid_roi_align_fwd_ws", "roi3d_pyramid_roi_align_fwd_f16_ws", "roi3d_pyramid_roi_align_grad_ws", "roi3d_overlaps3d", "roi3d_decode_proposals",
    "roi3d_topk_workspace_bytes", "roi3d_topk", "roi3d_gather_pad_boxes",
    "roi3d_proposal_layer_workspace_bytes", "roi3d_proposal_layer",
    "roi3d_refine_detections_workspace_bytes", "roi3d_refine_detections", "roi3d_mask_targets",
    "roi3d_pack_f16", "roi3d_unpack_f16", "roi3d_pack_bits", "roi3d_unpack_bits",
    "roi3d_set_option", "roi3d_get_option", "roi3d_kernel_launches", "roi3d_reset_kernel_launches",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libroi3d_b200.so")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into lib/libroi3d_b200.so."""
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)           # use nvcc's default host compiler (g++ on PATH)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB_PATH


_c_float_p = ctypes.POINTER(ctypes.c_float)


def _declare(lib):
    vp, i, f, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
    lib.roi3d_version.restype = ctypes.c_char_p
    lib.roi3d_version.argtypes = []
    lib.roi3d_strerror.restype = ctypes.c_char_p
    lib.roi3d_strerror.argtypes = [i]
    lib.roi3d_last_cuda_error.restype = i
    lib.roi3d_last_cuda_error.argtypes = []
    lib.roi3d_nms3d_workspace_bytes.restype = sz
    lib.roi3d_nms3d_workspace_bytes.argtypes = [i]
    lib.roi3d_nms3d.restype = i
    lib.roi3d_nms3d.argtypes = [vp, vp, i, i, f, vp, vp, vp, sz, vp]
    lib.roi3d_nms3d_batched_workspace_bytes.restype = sz
    lib.roi3d_nms3d_batched_workspace_bytes.argtypes = [i, i]
    lib.roi3d_nms3d_batched.restype = i
    lib.roi3d_nms3d_batched.argtypes = [vp, vp, vp, i, i, i, f, vp, vp, vp, sz, vp]
    lib.roi3d_car3d_fwd.restype = i
    lib.roi3d_car3d_fwd.argtypes = [vp, i, i, i, i, i, vp, vp, i, i, i, i, i, f, vp, vp]
    lib.roi3d_car3d_grad_image.restype = i
    lib.roi3d_car3d_grad_image.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, i, i, i, vp, vp]
    lib.roi3d_car3d_workspace_bytes.restype = sz
    lib.roi3d_car3d_workspace_bytes.argtypes = [i]
    lib.roi3d_car3d_fwd_ws.restype = i
    lib.roi3d_car3d_fwd_ws.argtypes = [vp, i, i, i, i, i, vp, vp, i, i, i, i, i, f, vp, vp, sz, vp]
    lib.roi3d_car3d_grad_image_ws.restype = i
    lib.roi3d_car3d_grad_image_ws.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, i, i, i, vp, vp, sz, vp]
    lib.roi3d_car3d_grad_boxes.restype = i
    lib.roi3d_car3d_grad_boxes.argtypes = [vp, vp, i, i, i, i, i, vp, vp, i, i, i, i, vp, vp]
    lib.roi3d_pyramid_roi_align_fwd.restype = i
    lib.roi3d_pyramid_roi_align_fwd.argtypes = [vp, vp, i, i, vp, i, vp, i, i, i, vp, vp]
    lib.roi3d_pyramid_roi_align_fwd_f16.restype = i
    lib.roi3d_pyramid_roi_align_fwd_f16.argtypes = [vp, vp, i, i, vp, i, vp, i, i, i, vp, vp]
    lib.roi3d_pyramid_roi_align_grad.restype = i
    lib.roi3d_pyramid_roi_align_grad.argtypes = [vp, vp, vp, i, i, vp, i, vp, i, i, i, vp]
    lib.roi3d_pyramid_roi_align_fwd_ws.restype = i
    lib.roi3d_pyramid_roi_align_fwd_ws.argtypes = [vp, vp, i, i, vp, i, vp, i, i, i, vp, vp, sz, vp]
    lib.roi3d_pyramid_roi_align_fwd_f16_ws.restype = i
    lib.roi3d_pyramid_roi_align_fwd_f16_ws.argtypes = [vp, vp, i, i, vp, i, vp, i, i, i, vp, vp, sz, vp]
    lib.roi3d_pyramid_roi_align_grad_ws.restype = i
    lib.roi3d_pyramid_roi_align_grad_ws.argtypes = [vp, vp, vp, i, i, vp, i, vp, i, i, i, vp, sz, vp]
    lib.roi3d_overlaps3d.restype = i
    lib.roi3d_overlaps3d.argtypes = [vp, i, vp, i, vp, vp]
    lib.roi3d_decode_proposals.restype = i
    lib.roi3d_decode_proposals.argtypes = [vp, vp, vp, i, vp, f, vp, vp]
    lib.roi3d_topk_workspace_bytes.restype = sz
    lib.roi3d_topk_workspace_bytes.argtypes = [i]
    lib.roi3d_topk.restype = i
    lib.roi3d_topk.argtypes = [vp, i, i, vp, vp, vp, sz, vp]
    lib.roi3d_proposal_layer_workspace_bytes.restype = sz
    lib.roi3d_proposal_layer_workspace_bytes.argtypes = [i, i, i]
    lib.roi3d_proposal_layer.restype = i
    lib.roi3d_proposal_layer.argtypes = [vp, vp, vp, i, vp, f, i, i, f, vp, vp, vp, sz, vp]
    lib.roi3d_gather_pad_boxes.restype = i
    lib.roi3d_gather_pad_boxes.argtypes = [vp, vp, vp, i, vp, vp]
    ll = ctypes.c_longlong
    lib.roi3d_refine_detections_workspace_bytes.restype = sz
    lib.roi3d_refine_detections_workspace_bytes.argtypes = [i, i, i]
    lib.roi3d_refine_detections.restype = i
    lib.roi3d_refine_detections.argtypes = [vp, vp, vp, i, i, i, vp, vp, f, f, i, i, vp, vp, vp, sz, vp]
    lib.roi3d_mask_targets.restype = i
    lib.roi3d_mask_targets.argtypes = [vp, i, i, i, i, i, vp, vp, i, i, i, i, vp, vp, vp]
    for name in ("roi3d_pack_f16", "roi3d_unpack_f16", "roi3d_pack_bits", "roi3d_unpack_bits"):
        getattr(lib, name).restype = i
        getattr(lib, name).argtypes = [vp, ll, vp, vp]
    lib.roi3d_set_option.restype = i
    lib.roi3d_set_option.argtypes = [ctypes.c_char_p, i]
    lib.roi3d_get_option.restype = i
    lib.roi3d_get_option.argtypes = [ctypes.c_char_p, ctypes.POINTER(i)]
    lib.roi3d_kernel_launches.restype = ctypes.c_longlong
    lib.roi3d_kernel_launches.argtypes = []
    lib.roi3d_reset_kernel_launches.restype = None
    lib.roi3d_reset_kernel_launches.argtypes = []
    return lib


_lib = None


def load():
    """Load libroi3d_b200.so; raises ImportError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libroi3d_b200.so is missing (%s). Build it with `python -c \"import __graft_entry__ as g; "
                "g.build()\"`; this package has no CPU fallback." % LIB_PATH)
        _lib = _declare(ctypes.CDLL(LIB_PATH))
    return _lib


class Roi3dError(RuntimeError):
    """A negative return code from the C ABI."""


def check(code):
    if code != 0:
        lib = load()
        msg = lib.roi3d_strerror(code).decode()
        if code == -4:
            msg += " [cudaError %d]" % lib.roi3d_last_cuda_error()
        raise Roi3dError("roi3d: %s (code %d)" % (msg, code))
