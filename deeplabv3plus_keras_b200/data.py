"""GPU input pipeline — the per-sample work of the reference's keras Sequences (ss.py:1528-1560) as one launch per
batch: raw decoded uint8 images / label maps in, network-ready tensors out.

`preprocess_batch` mirrors what `TrainingSequencePascalVOC2012Ext.__getitem__` returns for a batch — images
`[B,S,S,3]` in (-1, 1) and labels — except that labels stay an index map `[B,S,S]` (what the fused loss consumes; the
reference's one-hot tensor is `get_one_hot` of exactly this map).  JPEG/PNG decoding and dataset listing stay on the
host (out of scope)."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import BF16, F32, call

_REC = np.dtype([("src", "<u8"), ("inv_fy", "<f8"), ("inv_fx", "<f8"), ("h", "<i4"), ("w", "<i4"), ("hp", "<i4"),
                 ("wp", "<i4"), ("off_y", "<i4"), ("off_x", "<i4")])
assert _REC.itemsize == 48


def target_geometry(h: int, w: int, size: int) -> Tuple[int, int, int, int]:
    """(h_p, w_p, off_y, off_x) of resize_image_to_target_symmeric_size (ss.py:224-278): aspect-preserving fit into
    size x size; the reference pads portrait images with (pad_r, pad_l) — right amount on the left — kept as is."""
    if w >= h:
        h_p = int(h / w * size)
        return h_p, size, (size - h_p) // 2, 0
    w_p = int(w / h * size)
    pad = size - w_p
    return size, w_p, 0, pad - pad // 2


def _table(samples: Sequence[torch.Tensor], size: int, channels: int) -> torch.Tensor:
    rec = np.zeros(len(samples), dtype=_REC)
    for i, t in enumerate(samples):
        if t.dtype != torch.uint8 or not t.is_cuda or not t.is_contiguous():
            raise ValueError("preprocess: samples must be contiguous CUDA uint8 tensors (decoded HWC arrays)")
        h, w = int(t.shape[0]), int(t.shape[1])
        if t.numel() != h * w * channels:
            raise ValueError(f"preprocess: expected {channels} channel(s), got shape {tuple(t.shape)}")
        hp, wp, oy, ox = target_geometry(h, w, size)
        fy, fx = hp / float(h), wp / float(w)                      # ss.py:152-153
        rec[i] = (t.data_ptr(), 1.0 / fy, 1.0 / fx, h, w, hp, wp, oy, ox)
    return torch.from_numpy(rec.view(np.uint8).copy()).to(samples[0].device)


def preprocess_batch(images: Sequence[torch.Tensor], labels: Optional[Sequence[torch.Tensor]], size: int,
                     num_classes: int, dtype: torch.dtype = torch.float32,
                     out_images: Optional[torch.Tensor] = None, out_labels: Optional[torch.Tensor] = None):
    """images: list of uint8 [h,w,3] CUDA tensors; labels: list of uint8 [h,w] CUDA tensors (or None)."""
    dev = images[0].device
    B = len(images)
    st = torch.cuda.current_stream().cuda_stream
    x = out_images if out_images is not None else torch.empty((B, size, size, 3), dtype=dtype, device=dev)
    tab = _table(images, size, 3)
    call("dlv3p_preprocess_image_batch", tab.data_ptr(), B, x.data_ptr(), size,
         F32 if x.dtype == torch.float32 else BF16, st)
    y = None
    if labels is not None:
        if len(labels) != B:
            raise ValueError("preprocess: images and labels differ in count")
        y = out_labels if out_labels is not None else torch.empty((B, size, size), dtype=torch.int32, device=dev)
        ltab = _table(labels, size, 1)
        call("dlv3p_preprocess_label_batch", ltab.data_ptr(), B, y.data_ptr(), size, int(num_classes), st)
    # the descriptor tables are freed when this function returns; the caching allocator hands their memory out again
    # in stream order only, i.e. after the two launches above have read them
    return x, y
