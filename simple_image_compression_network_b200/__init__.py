"""B200-native FINN-style quantized non-square convolution layer (hot path of
shengjie-chen/simple_image_compression_network): host-side mirror of the reference's layer interface
over the C ABI in include/finnconv_b200.h.  The compute path is CUDA only (sm_100a)."""
from .desc import (ACT_BIAS_RELU, ACT_PASSTHROUGH, ACT_THRESHOLDS, CMP_GREATER, CMP_GREATER_EQUAL, CMP_LESS,
                   CMP_LESS_EQUAL, KIND_CONV, KIND_DECONV522, W_BINARY_PM1, W_BINARY_XNOR, W_FIXED, LayerDesc)

__all__ = ["LayerDesc", "KIND_CONV", "KIND_DECONV522", "W_FIXED", "W_BINARY_XNOR", "W_BINARY_PM1",
           "ACT_PASSTHROUGH", "ACT_BIAS_RELU", "ACT_THRESHOLDS", "CMP_LESS", "CMP_GREATER", "CMP_LESS_EQUAL",
           "CMP_GREATER_EQUAL"]
