"""Drop-in for the reference's ``src/model.py``: put ``<repo>/compat`` (and ``<repo>``) ahead of the
reference checkout on PYTHONPATH and ``from src.model import LineRefineNet`` (reference train.py:7,
train_dist.py:9, inference.py:12, inference_whole_scene.py:13) resolves to the B200-native module."""
from pointnet_refine_b200.model import (DetrTransformerDecoderLayer, LineRefineNet,  # noqa: F401
                                        MultiScalePointNetEncoder, PositionalEncoding)
