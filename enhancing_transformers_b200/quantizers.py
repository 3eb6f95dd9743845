"""Drop-in VectorQuantizer for the reference's ``enhancing/modules/stage1/quantizers.py``.

Same constructor (``VectorQuantizer(embed_dim, n_embed, beta=0.25, use_norm=True,
use_residual=False, num_quantizers=None, **kwargs)``, quantizers.py:67-68), same return triple
``(z_q with straight-through gradient, loss, int64 indices)`` and the attributes the unchanged
LightningModule reads (``embedding``, ``norm``, ``use_residual`` -- vitvqgan.py:82-85).  The
residual mode the task calls "ResidualQuantizer" is ``use_residual=True, num_quantizers=T``
exactly as in the reference (there is no separate class there).

Forward and backward each run as one fused CUDA launch sequence (csrc/vq.cu); the
[tokens, n_embed] distance matrix of quantizers.py:78-80 is never materialised."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn
from . import ops


class BaseQuantizer(nn.Module):
    def __init__(self, embed_dim: int, n_embed: int, straight_through: bool = True, use_norm: bool = True,
                 use_residual: bool = False, num_quantizers: Optional[int] = None) -> None:
        super().__init__()
        self.straight_through = straight_through
        self.use_norm = use_norm
        self.norm = (lambda x: F.normalize(x, dim=-1)) if use_norm else (lambda x: x)
        self.use_residual = use_residual
        self.num_quantizers = num_quantizers
        self.embed_dim = embed_dim
        self.n_embed = n_embed
        self.embedding = nn.Embedding(self.n_embed, self.embed_dim)
        self.embedding.weight.data.normal_()


class VectorQuantizer(BaseQuantizer):
    def __init__(self, embed_dim: int, n_embed: int, beta: float = 0.25, use_norm: bool = True,
                 use_residual: bool = False, num_quantizers: Optional[int] = None, **kwargs) -> None:
        super().__init__(embed_dim, n_embed, True, use_norm, use_residual, num_quantizers)
        self.beta = beta
        if use_residual and not num_quantizers:
            raise ValueError("use_residual=True needs num_quantizers")

    @property
    def depth(self) -> int:
        return int(self.num_quantizers) if self.use_residual else 1

    def forward(self, z: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        zc = z.contiguous()
        out, loss, idx = Fn.VectorQuantizeFn.apply(zc.view(-1, self.embed_dim), self.embedding.weight, self.depth,
                                                   float(self.beta), bool(self.use_residual), bool(self.use_norm))
        out = out.view_as(zc)
        idx = idx.view(*z.shape[:-1], self.depth) if self.use_residual else idx.view(*z.shape[:-1])
        return out, loss, idx

    def quantize(self, z: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """single-depth lookup (quantizers.py:74-92): (normalised code, loss, indices).  Inference helper:
        the returned code carries no autograd graph; training goes through forward()."""
        zc = z.detach().contiguous().view(-1, self.embed_dim)
        _, loss, idx = ops.vq_fwd(zc, self.embedding.weight.detach(), 1, float(self.beta), bool(self.use_norm))
        q = ops.vq_embed(self.embedding.weight.detach(), idx, 1, bool(self.use_norm)).view_as(z)
        return q, loss, idx.view(*z.shape[:-1])

    def embed_codes(self, code: torch.Tensor) -> torch.Tensor:
        """decode_codes fast path (vitvqgan.py:81-86 in one kernel): sum_t normalize(E[code[..., t]])"""
        depth = self.depth
        lead = code.shape[:-1] if self.use_residual else code.shape
        return ops.vq_embed(self.embedding.weight.detach(), code.contiguous().view(-1, depth), depth,
                            bool(self.use_norm)).view(*lead, self.embed_dim)
