"""TEST INFRASTRUCTURE ONLY: torch-CPU stand-ins for the `enhancing_transformers_b200.ops` wrappers the modules call,
following the contracts written in include/b200vq.h.  `install(monkeypatch)` swaps them in so that the *host logic* of
functional.py / layers.py / quantizers.py / stage2.py (autograd wiring, argument order, operand majors, gradient scaling,
packed-qkv layout, row windows, KV-cache bookkeeping) can be checked against the reference goldens on a machine without a
GPU.  The product never imports this file and has no CPU path; the kernels themselves are checked on the GPU (-m gpu)."""

import torch


def _mat(t, major, rows, k_total):
    """operand as a [rows, k_total] matrix: major 0 = stored [rows, K]; 1 = stored [K_total, rows]"""
    return t[:rows, :k_total] if major == 0 else t[:k_total, :rows].t()


def gemm(a, b, M, N, K, *, a_major=0, b_major=0, lda=None, ldb=None, out=None, bias=None, res=None, res_row_mod=0, aux=None,
         act=0, round_out=False, splits=1, cta_group=1, bn=0, want_colsum=False, a_lo=None, b_lo=None, out_half=False, alpha=None):
    assert out is None
    outs = []
    for z in range(splits):
        A = _mat(a, a_major, M, K * splits)[:, z * K:(z + 1) * K]
        B = _mat(b, b_major, N, K * splits)[:, z * K:(z + 1) * K]
        c = A.double() @ B.double().t()
        if alpha is not None:
            c = c * alpha.double()
        if bias is not None:
            c = c + bias.double()
        if act == 1:
            c = torch.tanh(c)
        if aux is not None:
            c = c * (1 - aux.double() ** 2)
        if res is not None:
            r = res.double().view(-1, N)
            c = c + (r[torch.arange(M) % res_row_mod] if res_row_mod else r)
        outs.append(c.half() if out_half else c.float())
    c = outs[0] if splits == 1 else torch.stack(outs)
    return (c, c.sum(0)) if want_colsum else c


def splitk_reduce(part, out=None, alpha=None):
    s = part.sum(0)
    return s * alpha if alpha is not None else s


def round_tf32(x, out=None):
    return x.clone()


def split_tf32_lo(x):
    return torch.zeros_like(x)


def to_half(x, scale=None):
    return (x if scale is None else x * scale).clamp(-65504, 65504).half()


def grad_scale(g, target_log2=6):
    m = g.abs().max().clamp_min(1e-30)
    S = torch.exp2(target_log2 - torch.ceil(torch.log2(m)))
    return torch.stack([S, 1 / S]).float()


def layernorm_fwd(x, gamma, beta, round_out, out_half=False):
    mean = x.mean(-1)
    var = x.var(-1, unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    y = ((x - mean[:, None]) * rstd[:, None]) * gamma + beta
    return (y.half() if out_half else y), mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dres, round_out=False, want_colsum=False, half_scale=None, dy_scale=None):
    dy = dy.float()
    if dy_scale is not None:                 # the fp16 output of a dgrad GEMM still carries the gradient scale
        dy = dy * dy_scale
    xh = (x - mean[:, None]) * rstd[:, None]
    g = dy * gamma
    dx = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    if dres is not None:
        dx = dx + dres
    out = (dx, (dy * xh).sum(0), dy.sum(0))
    out = out + ((dx.sum(0),) if want_colsum else ())
    return out + ((to_half(dx, half_scale),) if half_scale is not None else ())


def colsum(x):
    return x.view(-1, x.shape[-1]).sum(0)


def _visible(T, cond):
    m = torch.tril(torch.ones(T, T, dtype=torch.bool))
    m[:cond, :cond] = True
    return m


def _attn(qkv, B, N, heads, dh, scale, cond):
    q, k, v = (t.reshape(B, N, heads, dh).transpose(1, 2) for t in qkv.view(B, N, 3, heads * dh).unbind(2))
    s = ((q @ k.transpose(-2, -1)) * scale).masked_fill(~_visible(N, cond), float("-inf"))
    return (s.softmax(-1) @ v).transpose(1, 2).reshape(B * N, heads * dh), torch.logsumexp(s, -1).reshape(-1)


def attention_causal_fwd(qkv, B, N, heads, dh, scale, cond_len, exact, round_out=False):
    return _attn(qkv, B, N, heads, dh, scale, cond_len)


def attention_causal_bwd(qkv, out, lse, dout, B, N, heads, dh, scale, cond_len, exact, round_out=False):
    with torch.enable_grad():
        q = qkv.detach().clone().requires_grad_(True)
        o, _ = _attn(q, B, N, heads, dh, scale, cond_len)
        o.backward(dout)
    return q.grad


def _attn_grad(qkv, dout, B, N, heads, dh, scale, cond):
    with torch.enable_grad():
        q = qkv.detach().float().clone().requires_grad_(True)
        o, _ = _attn(q, B, N, heads, dh, scale, cond)
        o.backward(dout.float())
    return q.grad


# stage-1 cores: no mask == a fully visible prefix of N tokens
def attention_fwd(qkv, B, N, heads, dh, scale, round_out, out_half=False):
    o, lse = _attn(qkv.float(), B, N, heads, dh, scale, N)
    return (o.half() if out_half else o), lse


def attention_bwd(qkv, out, lse, dout, B, N, heads, dh, scale, round_out, half_scale=None):
    dq = _attn_grad(qkv, dout, B, N, heads, dh, scale, N)
    return to_half(dq, half_scale) if half_scale is not None else dq


def attention_f16_fwd(qkv, B, N, heads, dh, scale):
    o, lse = _attn(qkv.float(), B, N, heads, dh, scale, N)
    return o.half(), lse


def attention_f16_bwd(qkv, out, lse, dout, B, N, heads, dh, scale):
    return _attn_grad(qkv, dout, B, N, heads, dh, scale, N).clamp(-65504, 65504).half()     # linear in dout: carries its scale


def attention_exact_fwd(qkv, B, N, heads, dh, scale):
    return _attn(qkv, B, N, heads, dh, scale, N)


def attention_exact_bwd(qkv, out, lse, dout, B, N, heads, dh, scale):
    return _attn_grad(qkv, dout, B, N, heads, dh, scale, N)


def _pair(p):
    return (int(p[0]), int(p[1])) if isinstance(p, (tuple, list)) else (int(p), int(p))


def patchify(img, p, round_out):
    B, C, H, W = img.shape
    ph, pw = _pair(p)
    gh, gw = H // ph, W // pw
    return img.reshape(B, C, gh, ph, gw, pw).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, C * ph * pw).contiguous()


def unpatchify(tok, bias, B, C, H, W, p):
    ph, pw = _pair(p)
    gh, gw = H // ph, W // pw
    img = tok.reshape(B, gh, gw, C, ph, pw).permute(0, 3, 1, 4, 2, 5).reshape(B, C, H, W)
    return (img + bias.view(1, C, 1, 1) if bias is not None else img).contiguous()


def add_rows_mod(x, table):
    D = x.shape[-1]
    flat = x.reshape(-1, D)
    t = table.reshape(-1, D)
    return (flat + t[torch.arange(flat.shape[0]) % t.shape[0]]).view_as(x)


def _vq_oracle():
    from oracle import vitvq_oracle as O
    return O


def vq_fwd(z, E, depth, beta, use_norm=True):
    assert use_norm
    out, loss, idx = _vq_oracle().vq_forward(z.detach(), E.detach(), beta, depth > 1, depth)
    return out, loss, idx.reshape(-1, depth)


def vq_bwd(z, E, idx, g_out, g_loss, residual, beta, use_norm=True):
    depth = idx.shape[-1] if idx.dim() == 2 else 1
    with torch.enable_grad():
        z_, E_ = z.detach().clone().requires_grad_(True), E.detach().clone().requires_grad_(True)
        out, loss, _ = _vq_oracle().vq_forward(z_, E_, beta, bool(residual), depth)
        total = (out * g_out).sum() if g_out is not None else out.sum() * 0
        if g_loss is not None:
            total = total + loss * g_loss
        total.backward()
    return z_.grad, (E_.grad if E_.grad is not None else torch.zeros_like(E))


def vq_embed(E, codes, depth, use_norm=True):
    q = torch.nn.functional.normalize(E[codes.reshape(-1, depth)], dim=-1)
    return q.sum(-2)


def time_mix_fwd(x, w, T, round_out=False):
    xs = x.view(-1, T, x.shape[-1])
    sh = torch.cat([torch.zeros_like(xs[:, :1]), xs[:, :-1]], 1)
    return (xs * w + sh * (1 - w)).view_as(x)


def time_mix_bwd(g, x, w, T):
    C = x.shape[-1]
    gs, xs = g.view(-1, T, C), x.view(-1, T, C)
    gn = torch.cat([gs[:, 1:], torch.zeros_like(gs[:, :1])], 1)
    sh = torch.cat([torch.zeros_like(xs[:, :1]), xs[:, :-1]], 1)
    return (gs * w + gn * (1 - w)).view_as(x), (gs * (xs - sh)).sum((0, 1))


def sqrelu(x, g=None, round_out=False):
    r = torch.relu(x)
    return r * r if g is None else g * 2 * r


def token_embed_fwd(conds, codes, Wc, pos_c, Wi, pos_i):
    B, Tc = conds.shape
    Ti = codes.shape[1]
    C = Wi.shape[1]
    parts = []
    if Tc:
        parts.append(Wc[conds] + pos_c.reshape(-1, C)[:Tc])
    if Ti:
        parts.append(Wi[codes] + pos_i.reshape(-1, C)[:Ti])
    return torch.cat(parts, 1).reshape(B * (Tc + Ti), C)


def token_embed_bwd(conds, codes, g, Vc, Vi):
    B, Tc = conds.shape
    Ti = codes.shape[1]
    C = g.shape[-1]
    g3 = g.view(B, Tc + Ti, C)
    gWc = torch.zeros(Vc, C).index_add_(0, conds.reshape(-1), g3[:, :Tc].reshape(-1, C))
    gWi = torch.zeros(Vi, C).index_add_(0, codes.reshape(-1), g3[:, Tc:].reshape(-1, C))
    return gWc, g3[:, :Tc].sum(0), gWi, g3[:, Tc:].sum(0)


def copy_rows(src, B, T_src, T_dst, off_src, off_dst, n):
    C = src.shape[-1]
    dst = torch.zeros(B, T_dst, C)
    dst[:, off_dst:off_dst + n] = src.view(B, T_src, C)[:, off_src:off_src + n]
    return dst.view(B * T_dst, C)


def decode_attention(qkv, cache_k, cache_v, heads, hs, pos, scale):
    B, C = qkv.shape[0], heads * hs
    cache_k[:, pos], cache_v[:, pos] = qkv[:, C:2 * C], qkv[:, 2 * C:]
    q = qkv[:, :C].view(B, heads, 1, hs)
    K = cache_k[:, :pos + 1].view(B, pos + 1, heads, hs).transpose(1, 2)
    V = cache_v[:, :pos + 1].view(B, pos + 1, heads, hs).transpose(1, 2)
    return (((q @ K.transpose(-2, -1)) * scale).softmax(-1) @ V).reshape(B, C)


_COUNT = [0]


def launch_count():
    _COUNT[0] += 100
    return _COUNT[0]


NAMES = ("attention_fwd", "attention_bwd", "attention_f16_fwd", "attention_f16_bwd", "attention_exact_fwd", "attention_exact_bwd",
         "patchify", "unpatchify", "add_rows_mod", "vq_fwd", "vq_bwd", "vq_embed",
         "gemm", "splitk_reduce", "round_tf32", "split_tf32_lo", "to_half", "grad_scale", "layernorm_fwd", "layernorm_bwd", "colsum", "attention_causal_fwd",
         "attention_causal_bwd", "time_mix_fwd", "time_mix_bwd", "sqrelu", "token_embed_fwd", "token_embed_bwd", "copy_rows",
         "decode_attention", "launch_count")


def install(monkeypatch):
    from enhancing_transformers_b200 import functional, ops
    for n in NAMES:
        monkeypatch.setattr(ops, n, globals()[n])
    # the weight-shadow makers captured the real wrappers at import time
    monkeypatch.setitem(functional._SHADOW_MAKERS, "tf32", round_tf32)
    monkeypatch.setitem(functional._SHADOW_MAKERS, "lo", split_tf32_lo)
    monkeypatch.setitem(functional._SHADOW_MAKERS, "f16", to_half)
    monkeypatch.setattr(ops, "pick_splits", lambda *a, **k: 1)
