"""Drop-in for `crop_and_resize_3d_grad_image` (core/custom_op/custom_op.py:24)."""
from _roi3d_loader import _ops

crop_and_resize_3d_grad_image = _ops.crop_and_resize3d_grad_image
