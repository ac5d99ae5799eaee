"""deeplabv3plus_keras_b200 — B200-native (sm_100a) DeepLabV3+ encoder/decoder hot path behind the Keras-style
model-building surface of tonandr/deeplabv3plus_keras.

    from deeplabv3plus_keras_b200 import SemanticSegmentation      # mirrors bodhi.deeplabv3plus_keras (__init__.py:1)
"""
__version__ = "0.1.0"

_EXPORTS = ("SemanticSegmentation", "ClassBalancedLoss", "class_balanced_loss", "MeanIoUExt", "ss_pw", "ss_nw")


def __getattr__(name):          # lazy: `import deeplabv3plus_keras_b200` stays cheap (no torch import)
    if name in _EXPORTS:
        from . import deeplab
        return getattr(deeplab, name)
    raise AttributeError(name)
