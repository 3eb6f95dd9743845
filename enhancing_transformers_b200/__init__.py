"""B200-native ViT-VQGAN hot path.  Import as ``enhancing_transformers_b200``.

    import enhancing_transformers_b200 as etb
    etb.patch()          # rebind Encoder / Decoder / VectorQuantizer inside the reference's vitvqgan.py
    # ... then run the reference's main.py / ViTVQ unchanged
"""
import sys
import types

from . import _lib, functional, ops  # noqa: F401
from .layers import (Attention, FeedForward, PreNorm, Transformer, ViTDecoder, ViTEncoder,  # noqa: F401
                     sincos_table)
from .quantizers import BaseQuantizer, VectorQuantizer  # noqa: F401

__version__ = "0.1.0"
__all__ = ["ViTEncoder", "ViTDecoder", "VectorQuantizer", "BaseQuantizer", "Transformer", "Attention", "FeedForward",
           "PreNorm", "patch", "install_as_reference_modules", "ops", "functional"]

_REF_PKG = "enhancing.modules.stage1"


def patch(vitvqgan_module=None):
    """Rebind the three names ``ViTVQ.__init__`` looks up at construction time
    (reference vitvqgan.py:20-21,35-37).  Pass the already-imported module, or leave None to
    import ``enhancing.modules.stage1.vitvqgan``.  The reference file itself is not edited."""
    if vitvqgan_module is None:
        import importlib
        vitvqgan_module = importlib.import_module(_REF_PKG + ".vitvqgan")
    vitvqgan_module.Encoder = ViTEncoder
    vitvqgan_module.Decoder = ViTDecoder
    vitvqgan_module.VectorQuantizer = VectorQuantizer
    return vitvqgan_module


def install_as_reference_modules():
    """Alternative to patch(): pre-register this package's classes under the module names
    ``enhancing.modules.stage1.layers`` / ``.quantizers`` so that the reference's
    ``from .layers import ViTEncoder as Encoder`` resolves here.  Call before importing the
    reference's vitvqgan."""
    from . import layers as _layers
    from . import quantizers as _quantizers
    lay = types.ModuleType(_REF_PKG + ".layers")
    lay.__dict__.update({k: getattr(_layers, k) for k in ("ViTEncoder", "ViTDecoder", "Transformer", "Attention",
                                                           "FeedForward", "PreNorm")})
    qua = types.ModuleType(_REF_PKG + ".quantizers")
    qua.VectorQuantizer = _quantizers.VectorQuantizer
    qua.BaseQuantizer = _quantizers.BaseQuantizer

    class _GumbelUnavailable:  # vitvqgan.py:21 imports the name; ViTVQGumbel is out of scope
        def __init__(self, *a, **k):
            raise NotImplementedError("GumbelQuantizer is not part of the B200 hot path; use the reference class")
    qua.GumbelQuantizer = _GumbelUnavailable
    sys.modules[lay.__name__] = lay
    sys.modules[qua.__name__] = qua
    return lay, qua
