"""Drop-in for the reference wheel's top-level package `crop_and_resize_3d`
(core/custom_op/custom_op.py:22: `from crop_and_resize_3d import crop_and_resize_3d`)."""
from _roi3d_loader import _ops

crop_and_resize_3d = _ops.crop_and_resize3d
