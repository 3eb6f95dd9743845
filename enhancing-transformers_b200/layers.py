"""Drop-in ViTEncoder / ViTDecoder for the reference's ``enhancing/modules/stage1/layers.py``.

Same constructor keywords (``vitvqgan.py:35-36`` splats the YAML dicts into them), same
``forward`` signatures, same ``state_dict`` keys and shapes (checkpoints load both ways), same
sub-module tree (``transformer.layers[i][0].fn.to_qkv`` ...), but every FLOP runs in
libb200vq.so: the sub-modules are parameter containers whose ``forward`` dispatch to the CUDA
kernels, and ``Transformer.forward`` runs each pre-norm block as one fused autograd unit.

There is no CPU path: CPU tensors raise (use the reference classes on CPU)."""
from __future__ import annotations

import math
from typing import Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import functional as Fn

Size2 = Union[Tuple[int, int], int]


def _pair(v: Size2) -> Tuple[int, int]:
    return v if isinstance(v, tuple) else (v, v)


def sincos_table(dim: int, grid_hw: Tuple[int, int]) -> np.ndarray:
    """Fixed 2-D sin-cos positional table, float64 math then cast to float32, as the reference
    builds it (layers.py:21-68): channels [0, dim/2) encode the column index, [dim/2, dim) the
    row index, each half laid out as [sin | cos] over frequencies 10000^(-j/(dim/4))."""
    assert dim % 4 == 0, "embedding dim must be divisible by 4 for the 2-D sin-cos table"
    gh, gw = grid_hw
    freq = 1.0 / (10000.0 ** (np.arange(dim // 4, dtype=np.float64) / (dim / 4.0)))
    cols = np.tile(np.arange(gw, dtype=np.float64), gh)          # column index of token (r, c)
    rows = np.repeat(np.arange(gh, dtype=np.float64), gw)
    parts = []
    for coord in (cols, rows):
        ang = coord[:, None] * freq[None, :]
        parts += [np.sin(ang), np.cos(ang)]
    return np.concatenate(parts, axis=1).astype(np.float32)


def _xavier_(w: torch.Tensor) -> None:
    flat = w.view(w.shape[0], -1)
    bound = math.sqrt(6.0 / (flat.shape[0] + flat.shape[1]))
    with torch.no_grad():
        w.uniform_(-bound, bound)


def _init_like_reference(module: nn.Module) -> None:
    """Same distributions as reference init_weights (layers.py:71-82): xavier-uniform matrices,
    zero Linear biases, unit LayerNorm; conv biases keep torch's default."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            _xavier_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
        elif isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            _xavier_(m.weight)


def _flat2d(x: torch.Tensor) -> torch.Tensor:
    x = x.contiguous()
    return x.view(-1, x.shape[-1])


class PreNorm(nn.Module):
    def __init__(self, dim: int, fn: nn.Module) -> None:
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    def forward(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        y = Fn.LayerNormFn.apply(_flat2d(x), self.norm.weight, self.norm.bias, False).view_as(x)
        return self.fn(y, **kwargs)


class FeedForward(nn.Module):
    def __init__(self, dim: int, hidden_dim: int) -> None:
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.Tanh(), nn.Linear(hidden_dim, dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = Fn.LinearFn.apply(_flat2d(x), self.net[0].weight, self.net[0].bias, 1, False)
        y = Fn.LinearFn.apply(h, self.net[2].weight, self.net[2].bias, 0, False)
        return y.view(*x.shape[:-1], y.shape[-1])


class Attention(nn.Module):
    def __init__(self, dim: int, heads: int = 8, dim_head: int = 64) -> None:
        super().__init__()
        inner_dim = dim_head * heads
        project_out = not (heads == 1 and dim_head == dim)
        self.heads = heads
        self.dim_head = dim_head
        self.scale = dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)          # kept for module-tree parity; the kernel fuses it
        self.to_qkv = nn.Linear(dim, inner_dim * 3, bias=False)
        self.to_out = nn.Linear(inner_dim, dim) if project_out else nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, N, _ = x.shape
        qkv = Fn.LinearFn.apply(_flat2d(x), self.to_qkv.weight, None, 0, False)
        o = Fn.AttentionCoreFn.apply(qkv, B, N, self.heads, self.dim_head)
        if isinstance(self.to_out, nn.Linear):
            o = Fn.LinearFn.apply(o, self.to_out.weight, self.to_out.bias, 0, False)
        return o.view(B, N, -1)


class Transformer(nn.Module):
    def __init__(self, dim: int, depth: int, heads: int, dim_head: int, mlp_dim: int) -> None:
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head)),
                                              PreNorm(dim, FeedForward(dim, mlp_dim))]))
        self.norm = nn.LayerNorm(dim)
        self.heads, self.dim_head = heads, dim_head
        self.round_final = False    # decoder sets it: its final LN only feeds the to_pixel GEMM

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, N, D = x.shape
        h = _flat2d(x)
        for attn, ff in self.layers:
            a, f = attn.fn, ff.fn
            if not isinstance(a.to_out, nn.Linear):
                raise NotImplementedError("heads == 1 and dim_head == dim (no output projection) is not built")
            h = Fn.TransformerLayerFn.apply(h, attn.norm.weight, attn.norm.bias, a.to_qkv.weight, a.to_out.weight,
                                            a.to_out.bias, ff.norm.weight, ff.norm.bias, f.net[0].weight, f.net[0].bias,
                                            f.net[2].weight, f.net[2].bias, B, N, a.heads, a.dim_head)
        h = Fn.LayerNormFn.apply(h, self.norm.weight, self.norm.bias, self.round_final)
        return h.view(B, N, D)


class ViTEncoder(nn.Module):
    """reference layers.py:153-182"""

    def __init__(self, image_size: Size2, patch_size: Size2, dim: int, depth: int, heads: int, mlp_dim: int,
                 channels: int = 3, dim_head: int = 64) -> None:
        super().__init__()
        image_height, image_width = _pair(image_size)
        patch_height, patch_width = _pair(patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, \
            'Image dimensions must be divisible by the patch size.'
        if patch_height != patch_width or patch_height % 4:
            raise NotImplementedError("b200vq: square patches with side % 4 == 0 only")
        grid = (image_height // patch_height, image_width // patch_width)
        self.num_patches = grid[0] * grid[1]
        self.patch_dim = channels * patch_height * patch_width
        self.patch = patch_height
        # nn.Sequential(conv, rearrange) in the reference; index 0 keeps the checkpoint key
        self.to_patch_embedding = nn.Sequential(nn.Conv2d(channels, dim, kernel_size=patch_size, stride=patch_size),
                                                nn.Identity())
        self.en_pos_embedding = nn.Parameter(torch.from_numpy(sincos_table(dim, grid)).unsqueeze(0), requires_grad=False)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim)
        _init_like_reference(self)

    def forward(self, img: torch.Tensor) -> torch.Tensor:
        conv = self.to_patch_embedding[0]
        B = img.shape[0]
        x = Fn.PatchEmbedFn.apply(img.contiguous(), conv.weight, conv.bias, self.en_pos_embedding, self.patch)
        return self.transformer(x.view(B, self.num_patches, -1))


class ViTDecoder(nn.Module):
    """reference layers.py:185-217"""

    def __init__(self, image_size: Size2, patch_size: Size2, dim: int, depth: int, heads: int, mlp_dim: int,
                 channels: int = 3, dim_head: int = 64) -> None:
        super().__init__()
        image_height, image_width = _pair(image_size)
        patch_height, patch_width = _pair(patch_size)
        assert image_height % patch_height == 0 and image_width % patch_width == 0, \
            'Image dimensions must be divisible by the patch size.'
        if patch_height != patch_width or patch_height % 4:
            raise NotImplementedError("b200vq: square patches with side % 4 == 0 only")
        grid = (image_height // patch_height, image_width // patch_width)
        self.num_patches = grid[0] * grid[1]
        self.patch_dim = channels * patch_height * patch_width
        self.patch = patch_height
        self.image_hw = (image_height, image_width)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim)
        self.transformer.round_final = True
        self.de_pos_embedding = nn.Parameter(torch.from_numpy(sincos_table(dim, grid)).unsqueeze(0), requires_grad=False)
        self.to_pixel = nn.Sequential(nn.Identity(),
                                      nn.ConvTranspose2d(dim, channels, kernel_size=patch_size, stride=patch_size))
        _init_like_reference(self)

    def forward(self, token: torch.Tensor) -> torch.Tensor:
        B, N, D = token.shape
        x = Fn.AddPosFn.apply(_flat2d(token), self.de_pos_embedding)
        x = self.transformer(x.view(B, N, D))
        convt = self.to_pixel[1]
        return Fn.ToPixelFn.apply(_flat2d(x), convt.weight, convt.bias, B, self.image_hw[0], self.image_hw[1], self.patch)

    def get_last_layer(self) -> nn.Parameter:
        return self.to_pixel[-1].weight
