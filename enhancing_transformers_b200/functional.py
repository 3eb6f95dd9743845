"""autograd.Functions that stitch the C-ABI kernels into the forward/backward of the ViT blocks
and the quantiser.  Each Function is one fused unit whose saved tensors were chosen for HBM
footprint (SURVEY.md section 7 'Memory at config 2'): LayerNorm statistics, one copy of qkv, the
attention output and its log-sum-exp (never the N x N probabilities), the tanh output.

Precision.  Tensor-core GEMMs are tcgen05 kind::tf32 (fp32 accumulate).  The hardware truncates
fp32 operands to tf32, which would bias every product low; operands are therefore rounded to
nearest once, where they are produced (`round_out=True` on the producing kernel) or, for
parameters, in a cached shadow copy that is refreshed when the parameter's version counter
changes."""
from __future__ import annotations


import os

import torch

from . import ops

Tensor = torch.Tensor

# cta_group used for the big GEMMs (2 = CTA pair per 256-row tile); a module-level switch so
# that tests and bench.py can pin it
GEMM_CTA_GROUP = 2

_SHADOW_ATTR = "_b200vq_tf32_shadow"


def tf32_shadow(w: Tensor) -> Tensor:
    """tf32-rounded copy of a parameter.  The copy hangs off the parameter object itself (so it lives and
    dies with it -- a global cache keyed by address would hand a new parameter allocated at a recycled
    address the previous owner's values) and is refreshed when the parameter's version counter moves
    (in-place optimizer steps, ``load_state_dict`` and ``copy_`` all bump it)."""
    ent = getattr(w, _SHADOW_ATTR, None)
    ver = w._version
    if ent is not None and ent[0] == ver and ent[1].device == w.device:
        return ent[1]
    src = w.detach()
    if not src.is_contiguous():
        src = src.contiguous()
    out = ops.round_tf32(src, ent[1] if ent is not None and ent[1].shape == src.shape and ent[1].device == src.device else None)
    try:
        setattr(w, _SHADOW_ATTR, (ver, out))
    except AttributeError:      # exotic tensor subclasses without a __dict__: just do not cache
        pass
    return out


def clear_shadow_cache() -> None:
    """kept for API compatibility: shadows are per-parameter attributes now, nothing global to clear"""


def _wgrad(dy: Tensor, x: Tensor, rows: int, cols: int) -> Tensor:
    """dW[rows, cols] = dy[M, rows]^T . x[M, cols]   (both operands MN-major, split-K)"""
    M = dy.shape[0]
    splits = ops.pick_splits(M, rows, cols)
    part = ops.gemm(dy, x, rows, cols, M // splits, a_major=1, b_major=1, splits=splits, cta_group=GEMM_CTA_GROUP)
    return ops.splitk_reduce(part) if splits > 1 else part


GEMM_COLSUM = os.environ.get("B200VQ_GEMM_COLSUM", "1") != "0"   # bias-gradient column sums from the dgrad epilogue
_COLSUM_ATTR = "_b200vq_colsum"


def _attach_colsum(t: torch.Tensor, colsum: torch.Tensor) -> torch.Tensor:
    """Remember the column sums of a gradient tensor the LayerNorm-backward kernel produced for free.
    The next backward node down the residual stream needs exactly colsum(t) for a bias gradient; it
    picks the value up with `_colsum_of` if (and only if) the very same, unmodified tensor reaches it."""
    setattr(t, _COLSUM_ATTR, (colsum, t._version, t.data_ptr()))
    return t


def _colsum_of(t: torch.Tensor) -> torch.Tensor:
    tag = getattr(t, _COLSUM_ATTR, None)
    if tag is not None:
        colsum, version, ptr = tag
        if version == t._version and ptr == t.data_ptr() and colsum.shape[0] == t.shape[-1]:
            return colsum
    return ops.colsum(t)


class TransformerLayerFn(torch.autograd.Function):
    """x -> attn(LN(x)) + x -> ff(LN(.)) + .   (reference layers.py:145-148 with :85-132)"""

    @staticmethod
    def forward(ctx, x, ln1_w, ln1_b, w_qkv, w_out, b_out, ln2_w, ln2_b, w1, b1, w2, b2, B, N, heads, dh):
        cg = GEMM_CTA_GROUP
        M, D = x.shape
        inner = heads * dh
        mlp = w1.shape[0]
        scale = dh ** -0.5
        wq, wo, w1r, w2r = tf32_shadow(w_qkv), tf32_shadow(w_out), tf32_shadow(w1), tf32_shadow(w2)
        h1, mean1, rstd1 = ops.layernorm_fwd(x, ln1_w, ln1_b, True)
        qkv = ops.gemm(h1, wq, M, 3 * inner, D, round_out=True, cta_group=cg)
        o, lse = ops.attention_fwd(qkv, B, N, heads, dh, scale, True)
        x1 = ops.gemm(o, wo, M, D, inner, bias=b_out, res=x, cta_group=cg)
        h2, mean2, rstd2 = ops.layernorm_fwd(x1, ln2_w, ln2_b, True)
        t = ops.gemm(h2, w1r, M, mlp, D, bias=b1, act=1, round_out=True, cta_group=cg)
        x2 = ops.gemm(t, w2r, M, D, mlp, bias=b2, res=x1, cta_group=cg)
        ctx.save_for_backward(x, mean1, rstd1, h1, qkv, o, lse, x1, mean2, rstd2, h2, t, ln1_w, ln2_w, wq, wo, w1r, w2r)
        ctx.dims = (B, N, heads, dh)
        return x2

    @staticmethod
    def backward(ctx, g):
        cg = GEMM_CTA_GROUP
        x, mean1, rstd1, h1, qkv, o, lse, x1, mean2, rstd2, h2, t, ln1_w, ln2_w, wq, wo, w1r, w2r = ctx.saved_tensors
        B, N, heads, dh = ctx.dims
        M, D = x.shape
        inner = heads * dh
        mlp = w1r.shape[0]
        scale = dh ** -0.5
        g = g.contiguous()
        # ---- feed-forward branch
        db2 = _colsum_of(g)          # free when g came out of a LayerNorm-backward kernel (next block / final norm)
        dw2 = _wgrad(g, t, D, mlp)
        if GEMM_COLSUM:
            dt, db1 = ops.gemm(g, w2r, M, mlp, D, b_major=1, aux=t, round_out=True, cta_group=cg,
                               want_colsum=True)                                        # (g W2) * (1 - t^2), colsum(dt)
        else:
            dt = ops.gemm(g, w2r, M, mlp, D, b_major=1, aux=t, round_out=True, cta_group=cg)
            db1 = ops.colsum(dt)
        dw1 = _wgrad(dt, h2, mlp, D)
        dh2 = ops.gemm(dt, w1r, M, D, mlp, b_major=1, cta_group=cg)
        del dt
        g1, dln2_w, dln2_b, dbo = ops.layernorm_bwd(dh2, x1, mean2, rstd2, ln2_w, g, want_colsum=True)
        del dh2
        # ---- attention branch (dbo = colsum(g1) came with the kernel above)
        dwo = _wgrad(g1, o, D, inner)
        do = ops.gemm(g1, wo, M, inner, D, b_major=1, round_out=True, cta_group=cg)
        dqkv = ops.attention_bwd(qkv, o, lse, do, B, N, heads, dh, scale, True)
        del do
        dwq = _wgrad(dqkv, h1, 3 * inner, D)
        dh1 = ops.gemm(dqkv, wq, M, D, 3 * inner, b_major=1, cta_group=cg)
        del dqkv
        gx, dln1_w, dln1_b, gx_sum = ops.layernorm_bwd(dh1, x, mean1, rstd1, ln1_w, g1, want_colsum=True)
        _attach_colsum(gx, gx_sum)
        return gx, dln1_w, dln1_b, dwq, dwo, dbo, dln2_w, dln2_b, dw1, db1, dw2, db2, None, None, None, None


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm(dim) (reference layers.py:143,150)"""

    @staticmethod
    def forward(ctx, x, w, b, round_out):
        y, mean, rstd = ops.layernorm_fwd(x, w, b, bool(round_out))
        ctx.save_for_backward(x, mean, rstd, w)
        return y

    @staticmethod
    def backward(ctx, g):
        x, mean, rstd, w = ctx.saved_tensors
        dx, dw, db, dx_sum = ops.layernorm_bwd(g.contiguous(), x, mean, rstd, w, None, want_colsum=True)
        _attach_colsum(dx, dx_sum)
        return dx, dw, db, None


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) (+ res): stand-alone GEMM unit used by the sub-modules when they are
    called outside the fused layer (reference layers.py:99-101,118,120)"""

    @staticmethod
    def forward(ctx, x, w, b, act, round_out):
        M, K = x.shape
        N = w.shape[0]
        xr = ops.round_tf32(x)
        wr = tf32_shadow(w)
        y = ops.gemm(xr, wr, M, N, K, bias=b, act=int(act), round_out=bool(round_out), cta_group=GEMM_CTA_GROUP)
        ctx.save_for_backward(xr, wr, y if act else None)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, g):
        xr, wr, y = ctx.saved_tensors
        M, K = xr.shape
        N = wr.shape[0]
        g = g.contiguous()
        if y is not None:   # tanh backward folded into an elementwise pass of the dgrad epilogue is
            g = g * (1 - y * y)   # only available fused (TransformerLayerFn); stand-alone path keeps it simple
        gr = ops.round_tf32(g)
        db = ops.colsum(gr) if ctx.has_bias else None
        dw = _wgrad(gr, xr, N, K)
        dx = ops.gemm(gr, wr, M, K, N, b_major=1, cta_group=GEMM_CTA_GROUP)
        return dx, dw, db, None, None


class AttentionCoreFn(torch.autograd.Function):
    """softmax(q k^T * scale) v on the packed qkv matrix (reference layers.py:124-130)"""

    @staticmethod
    def forward(ctx, qkv, B, N, heads, dh):
        scale = dh ** -0.5
        qr = ops.round_tf32(qkv)
        o, lse = ops.attention_fwd(qr, B, N, heads, dh, scale, False)
        ctx.save_for_backward(qr, o, lse)
        ctx.dims = (B, N, heads, dh)
        return o

    @staticmethod
    def backward(ctx, g):
        qr, o, lse = ctx.saved_tensors
        B, N, heads, dh = ctx.dims
        dqkv = ops.attention_bwd(qr, o, lse, ops.round_tf32(g.contiguous()), B, N, heads, dh, dh ** -0.5, False)
        return dqkv, None, None, None, None


class PatchEmbedFn(torch.autograd.Function):
    """Conv2d(C, D, k=s=p) + 'b c h w -> b (h w) c' + positional table
    (reference layers.py:168-172,178-179) as one GEMM over the im2col view."""

    @staticmethod
    def forward(ctx, img, w, b, pos, p):
        B, C, H, W = img.shape
        D = w.shape[0]
        n_tok = (H // p) * (W // p)
        patches = ops.patchify(img, p, True)
        M, pd = patches.shape
        wr = tf32_shadow(w)
        x = ops.gemm(patches, wr.view(D, pd), M, D, pd, bias=b, res=pos.view(n_tok, D), res_row_mod=n_tok,
                     cta_group=GEMM_CTA_GROUP)
        ctx.save_for_backward(patches, wr)
        ctx.geom = (B, C, H, W, p)
        return x

    @staticmethod
    def backward(ctx, g):
        patches, w = ctx.saved_tensors
        B, C, H, W, p = ctx.geom
        D = w.shape[0]
        M, pd = patches.shape
        g = g.contiguous()
        db = _colsum_of(g)
        dw = _wgrad(g, patches, D, pd).view_as(w)
        dimg = None
        if ctx.needs_input_grad[0]:
            dpat = ops.gemm(g, w.view(D, pd), M, pd, D, b_major=1, cta_group=GEMM_CTA_GROUP)
            dimg = ops.unpatchify(dpat, None, B, C, H, W, p)
        return dimg, dw, db, None, None


class ToPixelFn(torch.autograd.Function):
    """'b (h w) c -> b c h w' + ConvTranspose2d(D, C, k=s=p) (reference layers.py:202-205,212)
    as one GEMM plus a pixel-shuffle store."""

    @staticmethod
    def forward(ctx, x, w, b, B, H, W, p):
        M, D = x.shape
        C = w.shape[1]
        pd = C * p * p
        wr = tf32_shadow(w)
        y = ops.gemm(x, wr.view(D, pd), M, pd, D, b_major=1, cta_group=GEMM_CTA_GROUP)
        img = ops.unpatchify(y, b, B, C, H, W, p)
        ctx.save_for_backward(x, wr)
        ctx.geom = (B, C, H, W, p)
        return img

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        B, C, H, W, p = ctx.geom
        M, D = x.shape
        pd = C * p * p
        dy = ops.patchify(g.contiguous(), p, True)
        db = ops.colsum(dy).view(C, p * p).sum(dim=1)
        dw = _wgrad(x, dy, D, pd).view_as(w)
        dx = ops.gemm(dy, w.view(D, pd), M, D, pd, cta_group=GEMM_CTA_GROUP)
        return dx, dw, db, None, None, None, None


class AddPosFn(torch.autograd.Function):
    """token + de_pos_embedding (reference layers.py:210)"""

    @staticmethod
    def forward(ctx, x, pos):
        return ops.add_rows_mod(x, pos)

    @staticmethod
    def backward(ctx, g):
        return g, None


class VectorQuantizeFn(torch.autograd.Function):
    """BaseQuantizer.forward + VectorQuantizer.quantize (reference quantizers.py:38-63,74-92)."""

    @staticmethod
    def forward(ctx, z, E, depth, beta, residual):
        out, loss, idx = ops.vq_fwd(z, E, depth, beta)
        ctx.save_for_backward(z, E, idx)
        ctx.cfg = (depth, beta, residual)
        ctx.mark_non_differentiable(idx)
        return out, loss, idx

    @staticmethod
    def backward(ctx, g_out, g_loss, _g_idx):
        z, E, idx = ctx.saved_tensors
        depth, beta, residual = ctx.cfg
        if g_out is not None:
            g_out = g_out.contiguous()
        if g_loss is not None:
            g_loss = g_loss.contiguous()
        gz, gE = ops.vq_bwd(z, E, idx, g_out, g_loss, residual, beta)
        return gz, gE, None, None, None
