"""B200-native ViT-VQGAN hot path.  Import as ``enhancing_transformers_b200``."""
from . import _lib, ops  # noqa: F401

__version__ = "0.1.0"
