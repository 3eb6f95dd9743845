"""TEST INFRASTRUCTURE ONLY: torch-CPU stand-ins for the `enhancing_transformers_b200.ops` wrappers the stage-2 modules call,
following the contracts written in include/b200vq.h.  `install(monkeypatch)` swaps them in so that the *host logic* of
stage2.py (autograd wiring, argument order, packed-qkv layout, row windows, KV-cache bookkeeping) can be checked against the
reference golden on a machine without a GPU.  The product never imports this file and has no CPU path; the kernels
themselves are checked on the GPU (tests/test_stage2.py -m gpu)."""

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
    return ((x - mean[:, None]) * rstd[:, None]) * gamma + beta, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dres, round_out=False, want_colsum=False, half_scale=None, dy_scale=None):
    xh = (x - mean[:, None]) * rstd[:, None]
    g = dy * gamma
    dx = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    if dres is not None:
        dx = dx + dres
    out = (dx, (dy * xh).sum(0), dy.sum(0))
    return out + ((dx.sum(0),) if want_colsum else ())


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


NAMES = ("gemm", "splitk_reduce", "round_tf32", "split_tf32_lo", "to_half", "grad_scale", "layernorm_fwd", "layernorm_bwd", "colsum", "attention_causal_fwd",
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
