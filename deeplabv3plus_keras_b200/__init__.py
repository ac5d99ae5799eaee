"""deeplabv3plus_keras_b200 — B200-native (sm_100a) DeepLabV3+ encoder/decoder hot path behind the Keras-style
model-building surface of tonandr/deeplabv3plus_keras."""
__version__ = "0.1.0"
