from .activation import GaussianActivation
from .feature_pyramid import ConvNeXtExtractor, FeaturePyramidContext, FeaturePyramidExtractor
from .linear_lift import LinearLift
from .mlp import MLP
from .normalization import AdaGN
from .ray import GroupNormBNC, RayNetwork
from .set_transformer import AttentionPool, Broadcast, BroadcastingLayer, SetTransformer
