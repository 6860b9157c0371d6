"""3d-mask-r-cnn_b200 -- B200-native (sm_100a) ROI hot path of 3D Mask R-CNN.

Scope: NonMaxSuppression3D and CropAndResize3D forward /
grad-image / grad-boxes behind the reference's own custom-op surface
(core/custom_op/custom_op.py).  Layout:

    csrc/        hand-written CUDA kernels + the C ABI (include/roi3d.h)
    lib/         libroi3d_b200.so, built in-tree by _lib.build()
    custom_op.py host-side mirror of the reference's four callables + gradient, and of the layers around them
    target_files.py  the head-target .npz files with device-side fp16 / bit packing
    tf_ops/      TensorFlow op registration sources for the real drop-in

The directory name is not a Python identifier; import it through the
``roi3d_b200`` shim at the repository root (``import roi3d_b200``).
Importing fails loudly when the CUDA library has not been built: there is no
CPU fallback on this path.
"""
from . import _lib

_lib.load()                     # ImportError if libroi3d_b200.so is missing

from . import custom_op         # noqa: E402
from . import sharding          # noqa: E402,F401
from . import target_files      # noqa: E402,F401
from .custom_op import (        # noqa: E402,F401
    InvalidArgumentError,
    crop_and_resize_3d,
    crop_and_resize_3d_grad_boxes,
    crop_and_resize_3d_grad_image,
    decode_proposals,
    deferred,
    synchronize,
    get_option,
    host_pipeline,
    kernel_launches,
    mask_targets,
    non_max_suppression_3d,
    non_max_suppression_3d_batched,
    non_max_suppression_3d_graph,
    non_max_suppression_3d_per_class,
    overlaps_3d,
    pack_bits,
    pack_f16,
    proposal_layer,
    top_k_set,
    pyramid_roi_align_3d,
    refine_detections,
    reset_kernel_launches,
    set_option,
    unpack_bits,
    unpack_f16,
    upload,
)

__version__ = "0.1.0"
