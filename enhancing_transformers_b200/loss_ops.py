"""Drop-in replacements for the reference's ``enhancing/losses/op`` package: the two StyleGAN2
ops its discriminator uses, bound to libb200vq.so instead of the CUDA extensions the reference JIT-compiles at import
time (``torch.utils.cpp_extension.load`` in losses/op/fused_act.py:11 and upfirdn2d.py:11 -- which needs nvcc, ninja and
a visible GPU on the training host).

    fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5)      losses/op/fused_act.py:110-127
    FusedLeakyReLU(channel, bias=True, negative_slope=0.2, scale=2 ** 0.5)      losses/op/fused_act.py:93-107
    upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0))                          losses/op/upfirdn2d.py:149-165

Same signatures, same first- and second-order autograd structure (R1-style gradient penalties differentiate through the
backward).  CPU tensors take the reference's own CPU formulas (plain torch), as the reference does.
``install_as_reference_ops()`` registers this module as ``enhancing.losses.op`` before the reference imports it."""
from __future__ import annotations

import sys
import types
from collections import abc

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function

from . import _lib
from .ops import _p, _req, _stream


def _bias_act(x, bias, ref, act, grad, alpha, scale):
    _req(x, "input"); _req(bias, "bias"); _req(ref, "ref")
    out = torch.empty_like(x)
    step_b = 1
    size_b = 0
    if bias is not None:
        size_b = bias.numel()
        for d in x.shape[2:]:
            step_b *= d
    _lib.check(_lib.lib().b200vq_bias_act(_p(x), _p(bias), _p(ref), _p(out), x.numel(), step_b, size_b, act, grad, float(alpha),
                                          float(scale), _stream()), "bias_act")
    return out


class _FusedLeakyReLUBackward(Function):
    """d/dx of the fused op, itself differentiable (second derivative of a piecewise-linear function: the same mask)"""

    @staticmethod
    def forward(ctx, grad_output, out, has_bias, negative_slope, scale):
        ctx.save_for_backward(out)
        ctx.cfg = (negative_slope, scale)
        grad_input = _bias_act(grad_output.contiguous(), None, out, 3, 1, negative_slope, scale)
        dims = [0] + list(range(2, grad_input.ndim))
        grad_bias = grad_input.sum(dims).detach() if has_bias else grad_input.new_empty(0)
        return grad_input, grad_bias

    @staticmethod
    def backward(ctx, gradgrad_input, gradgrad_bias):
        out, = ctx.saved_tensors
        negative_slope, scale = ctx.cfg
        bias = gradgrad_bias.contiguous() if gradgrad_bias is not None and gradgrad_bias.numel() else None
        return _bias_act(gradgrad_input.contiguous(), bias, out, 3, 1, negative_slope, scale), None, None, None, None


class _FusedLeakyReLU(Function):
    @staticmethod
    def forward(ctx, input, bias, negative_slope, scale):
        ctx.has_bias = bias is not None
        out = _bias_act(input, bias.contiguous() if bias is not None else None, None, 3, 0, negative_slope, scale)
        ctx.save_for_backward(out)
        ctx.cfg = (negative_slope, scale)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        out, = ctx.saved_tensors
        negative_slope, scale = ctx.cfg
        grad_input, grad_bias = _FusedLeakyReLUBackward.apply(grad_output, out, ctx.has_bias, negative_slope, scale)
        return grad_input, (grad_bias if ctx.has_bias else None), None, None


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    if input.device.type == "cpu":          # the reference's CPU branch, verbatim semantics (fused_act.py:111-122; slope fixed at 0.2 there)
        if bias is not None:
            rest_dim = [1] * (input.ndim - bias.ndim - 1)
            return F.leaky_relu(input + bias.view(1, bias.shape[0], *rest_dim), negative_slope=0.2) * scale
        return F.leaky_relu(input, negative_slope=0.2) * scale
    return _FusedLeakyReLU.apply(input.contiguous(), bias, negative_slope, scale)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def _upfirdn2d_raw(x, kernel, up, down, pad):
    """x [planes..., H, W] contiguous fp32 -> [planes..., out_h, out_w]"""
    _req(x, "input"); _req(kernel, "kernel")
    in_h, in_w = x.shape[-2:]
    kh, kw = kernel.shape
    planes = x.numel() // (in_h * in_w)
    out_h = (in_h * up[1] + pad[2] + pad[3] - kh + down[1]) // down[1]
    out_w = (in_w * up[0] + pad[0] + pad[1] - kw + down[0]) // down[0]
    out = torch.empty(*x.shape[:-2], out_h, out_w, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().b200vq_upfirdn2d(_p(x), _p(kernel), _p(out), planes, in_h, in_w, kh, kw, up[0], up[1], down[0], down[1],
                                           pad[0], pad[1], pad[2], pad[3], _stream()), "upfirdn2d")
    return out


class _UpFirDn2dBackward(Function):
    """the adjoint is another upfirdn2d (flipped kernel, up <-> down, complementary padding: upfirdn2d.py:107-112)"""

    @staticmethod
    def forward(ctx, grad_output, kernel, flipped, up, down, pad, g_pad, in_size):
        ctx.save_for_backward(kernel)
        ctx.cfg = (up, down, pad)
        gi = _upfirdn2d_raw(grad_output.contiguous(), flipped, down, up, g_pad)
        return gi.view(in_size)

    @staticmethod
    def backward(ctx, gradgrad_input):
        kernel, = ctx.saved_tensors
        up, down, pad = ctx.cfg
        return _upfirdn2d_raw(gradgrad_input.contiguous(), kernel, up, down, pad), None, None, None, None, None, None, None


class _UpFirDn2d(Function):
    @staticmethod
    def forward(ctx, input, kernel, up, down, pad):
        up_x, up_y = up
        down_x, down_y = down
        pad_x0, pad_x1, pad_y0, pad_y1 = pad
        kh, kw = kernel.shape
        in_h, in_w = input.shape[-2:]
        out = _upfirdn2d_raw(input, kernel, up, down, pad)
        out_h, out_w = out.shape[-2:]
        ctx.save_for_backward(kernel, torch.flip(kernel, [0, 1]).contiguous())
        ctx.cfg = (up, down, pad, (kw - pad_x0 - 1, in_w * up_x - out_w * down_x + pad_x0 - up_x + 1,
                                   kh - pad_y0 - 1, in_h * up_y - out_h * down_y + pad_y0 - up_y + 1), input.shape)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        kernel, flipped = ctx.saved_tensors
        up, down, pad, g_pad, in_size = ctx.cfg
        gi = None
        if ctx.needs_input_grad[0]:
            gi = _UpFirDn2dBackward.apply(grad_output, kernel, flipped, up, down, pad, g_pad, in_size)
        return gi, None, None, None, None


def upfirdn2d_native(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1):
    """CPU path: zero-insert, pad / crop, convolve, decimate -- the reference's own CPU formula (upfirdn2d.py:168-206)"""
    _, channel, in_h, in_w = input.shape
    x = input.reshape(-1, in_h, 1, in_w, 1)
    x = F.pad(x, [0, up_x - 1, 0, 0, 0, up_y - 1]).reshape(-1, in_h * up_y, in_w * up_x)
    x = F.pad(x, [max(pad_x0, 0), max(pad_x1, 0), max(pad_y0, 0), max(pad_y1, 0)])
    x = x[:, max(-pad_y0, 0):x.shape[1] - max(-pad_y1, 0), max(-pad_x0, 0):x.shape[2] - max(-pad_x1, 0)]
    kh, kw = kernel.shape
    x = F.conv2d(x.unsqueeze(1), torch.flip(kernel, [0, 1]).view(1, 1, kh, kw))[:, 0, ::down_y, ::down_x]
    return x.reshape(-1, channel, x.shape[-2], x.shape[-1])


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    if not isinstance(up, abc.Iterable):
        up = (up, up)
    if not isinstance(down, abc.Iterable):
        down = (down, down)
    if len(pad) == 2:
        pad = (pad[0], pad[1], pad[0], pad[1])
    if input.device.type == "cpu":
        return upfirdn2d_native(input, kernel, *up, *down, *pad)
    return _UpFirDn2d.apply(input.contiguous(), kernel.contiguous(), tuple(up), tuple(down), tuple(pad))


def install_as_reference_ops():
    """register this module's ops as ``enhancing.losses.op`` (call before importing ``enhancing.losses``): the
    reference's ``from .op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d, conv2d_gradfix`` (losses/layers.py:19) then
    resolves here and no extension is JIT-compiled.  ``conv2d_gradfix`` is a thin wrapper over ``F.conv2d`` /
    ``F.conv_transpose2d`` in the reference (library calls); the stand-in forwards to them."""
    mod = types.ModuleType("enhancing.losses.op")
    mod.FusedLeakyReLU, mod.fused_leaky_relu, mod.upfirdn2d = FusedLeakyReLU, fused_leaky_relu, upfirdn2d
    gradfix = types.ModuleType("enhancing.losses.op.conv2d_gradfix")
    gradfix.conv2d, gradfix.conv_transpose2d = F.conv2d, F.conv_transpose2d
    mod.conv2d_gradfix = gradfix
    sys.modules[mod.__name__] = mod
    sys.modules[gradfix.__name__] = gradfix
    return mod
