"""Head-target files of the reference's 3-stage pipeline, written and read with the payload work on the device.

The reference's TARGET_GENERATION stage predicts per image and saves six artefacts as ``.npz``
(core/models.py:3585-3636, ``_save_arrays``); HEAD training reads them back (core/data_generators.py:1886-1960).
The bulky ones are ``rois_aligned`` (float32 -> float16) and the two mask arrays (``> 0.5`` -> ``numpy.packbits``).
Here those conversions run as CUDA kernels *before* the download (writer) and *after* the upload (reader), so only
the 2x / 32x smaller payloads cross PCIe.  File names, ``.npz`` keys and dtypes are the reference's: files written
here load in the reference's reader and vice versa.  The ``.npz`` container (zip + deflate) is host-side numpy.

    rois/<name>.npz              rois            float32 [T,6]
    rois_aligned/<name>.npz      rois_aligned    float16 [T,P,P,P,C]
    mask_aligned/<name>.npz      mask_bits uint8 [ceil(n/8)], mask_shape int32
    target_class_ids/<name>.npz  tci             int32 [T]
    target_bbox/<name>.npz       bbox            float32 [T,6]
    target_mask/<name>.npz       tm_bits uint8, tm_shape int32
"""
import os

import numpy as np
import torch

from . import custom_op

DIRS = ("rois", "rois_aligned", "mask_aligned", "target_class_ids", "target_bbox", "target_mask")


def _host(x, dtype):
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.asarray(x).astype(dtype, copy=False)


def _dev(x):
    dev = custom_op._device()
    t = x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x))
    return t.to(dev, torch.float32, non_blocking=True)


def save_head_targets(base_path, name, rois=None, rois_aligned=None, mask_aligned=None, target_class_ids=None,
                      target_bbox=None, target_mask=None):
    """``_save_arrays(..., use_npz=True)`` (core/models.py:3604-3636).  Tensors may live on the device (the
    predictions of the targeting model) or the host.  Returns the six paths in the CSV's column order
    (``None`` for artefacts not given)."""
    paths = []

    def save(sub, **arrays):
        d = os.path.join(base_path, sub)
        os.makedirs(d, exist_ok=True)
        p = os.path.join(d, name + ".npz")
        np.savez_compressed(p, **arrays)
        paths.append(p)

    def bits(x):
        payload, shape = custom_op.pack_bits(_dev(x))
        return payload.cpu().numpy(), shape

    if rois is not None:
        save("rois", rois=_host(rois, np.float32))
    else:
        paths.append(None)
    if rois_aligned is not None:
        save("rois_aligned", rois_aligned=custom_op.pack_f16(_dev(rois_aligned)).cpu().numpy())
    else:
        paths.append(None)
    if mask_aligned is not None:
        b, s = bits(mask_aligned)
        save("mask_aligned", mask_bits=b, mask_shape=s)
    else:
        paths.append(None)
    if target_class_ids is not None:
        save("target_class_ids", tci=_host(target_class_ids, np.int32))
    else:
        paths.append(None)
    if target_bbox is not None:
        save("target_bbox", bbox=_host(target_bbox, np.float32))
    else:
        paths.append(None)
    if target_mask is not None:
        b, s = bits(target_mask)
        save("target_mask", tm_bits=b, tm_shape=s)
    else:
        paths.append(None)
    return tuple(paths)


def load_head_targets(paths):
    """The reader (core/data_generators.py:1886-1960 ``load_data``): returns device tensors
    ``(rois, rois_aligned float32, mask_aligned float32 0/1, target_class_ids int32, target_bbox, target_mask)``;
    the float16 and bit payloads are uploaded as stored and expanded on the device."""
    dev = custom_op._device()
    r_path, ra_path, ma_path, tci_path, tb_path, tm_path = paths

    def plain(p, key, dtype):
        if p is None:
            return None
        with np.load(p, allow_pickle=False) as z:
            return torch.from_numpy(np.asarray(z[key], dtype)).to(dev)

    def unbit(p, bits_key, shape_key):
        if p is None:
            return None
        with np.load(p, allow_pickle=False) as z:
            payload, shape = z[bits_key], z[shape_key].astype(np.int64)
        return custom_op.unpack_bits(torch.from_numpy(payload).to(dev), tuple(shape))

    ra = None
    if ra_path is not None:
        with np.load(ra_path, allow_pickle=False) as z:
            ra = custom_op.unpack_f16(torch.from_numpy(z["rois_aligned"]).to(dev))
    return (plain(r_path, "rois", np.float32), ra, unbit(ma_path, "mask_bits", "mask_shape"),
            plain(tci_path, "tci", np.int32), plain(tb_path, "bbox", np.float32), unbit(tm_path, "tm_bits", "tm_shape"))
