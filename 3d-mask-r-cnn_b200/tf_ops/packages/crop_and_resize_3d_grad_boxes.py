"""Drop-in for `crop_and_resize_3d_grad_boxes` (core/custom_op/custom_op.py:23)."""
from _roi3d_loader import _ops

crop_and_resize_3d_grad_boxes = _ops.crop_and_resize3d_grad_boxes
