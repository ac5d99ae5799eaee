"""Host-side mirror of the tf.keras surface the reference builds its model with."""
from . import backend, layers
from .base import Input, reset_uids, set_random_seed
from .layers import (Activation, Add, AveragePooling2D, BatchNormalization, Concatenate, Conv2D, DepthwiseConv2D,
                     Dropout, Lambda, MaxPooling2D, ReLU, SeparableConv2D, ZeroPadding2D, initializers, regularizers)
from .models import Model
