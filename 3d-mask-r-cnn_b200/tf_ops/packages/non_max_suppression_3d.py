"""Drop-in for `non_max_suppression_3d` (core/custom_op/custom_op.py:25)."""
from _roi3d_loader import _ops

non_max_suppression_3d = _ops.non_max_suppression3d
