"""Loads _roi3d_tf_ops.so once (all four ops live in one library) -- the equivalent of the
per-package `load_library.load_op_library(resource_loader.get_path_to_datafile(...))` calls of the
reference wheel (crop_and_resize_3d/python/ops/crop_and_resize_3d_ops.py:1-8)."""
import os

from tensorflow.python.framework import load_library

_ops = load_library.load_op_library(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_roi3d_tf_ops.so"))
