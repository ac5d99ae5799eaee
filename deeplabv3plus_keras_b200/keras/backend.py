"""tf.keras.backend subset used by the reference: K.resize_images (ss.py:852,904,941,946), K.int_shape (ss.py:883)."""
from __future__ import annotations

from .base import KTensor, reset_uids
from .layers import ResizeImages


def int_shape(x: KTensor):
    return tuple(x.shape)


def resize_images(x: KTensor, height_factor, width_factor, data_format="channels_last", interpolation="nearest"):
    if data_format != "channels_last":
        raise ValueError("resize_images: channels_last only")
    return ResizeImages(height_factor, width_factor, interpolation)(x)


def clear_session():
    reset_uids()
