"""Tiny launch scripts for `ncu --set full` captures (one kernel family per invocation):
    python tools/ncu_target.py gemm|gemm2|gemm16|attn|attn16|vq|ln"""
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import enhancing_transformers_b200 as etb  # noqa: E402

ops = etb.ops
what = sys.argv[1] if len(sys.argv) > 1 else "gemm"
torch.manual_seed(0)
if what in ("gemm", "gemm2"):
    M, N, K = 131072, 2304, 768
    a = torch.randn(M, K, device="cuda"); b = torch.randn(N, K, device="cuda")
    a, b = ops.round_tf32(a), ops.round_tf32(b)
    out = torch.empty(M, N, device="cuda")
    for _ in range(4):
        ops.gemm(a, b, M, N, K, out=out, cta_group=2 if what == "gemm2" else 1, round_out=True)
elif what == "attn":
    B, N, heads, dh = 32, 1024, 12, 64
    qkv = ops.round_tf32(torch.randn(B * N, 3 * heads * dh, device="cuda"))
    for _ in range(3):
        o, lse = ops.attention_fwd(qkv, B, N, heads, dh, 0.125, True)
    do = ops.round_tf32(torch.randn(B * N, heads * dh, device="cuda"))
    for _ in range(2):
        ops.attention_bwd(qkv, o, lse, do, B, N, heads, dh, 0.125, True)
elif what == "attn16":
    B, N, heads, dh = 32, 1024, 12, 64
    qkv = torch.randn(B * N, 3 * heads * dh, device="cuda").half()
    for _ in range(2):
        o, lse = ops.attention_f16_fwd(qkv, B, N, heads, dh, 0.125)
    do = torch.randn(B * N, heads * dh, device="cuda").half()
    for _ in range(2):
        ops.attention_f16_bwd(qkv, o, lse, do, B, N, heads, dh, 0.125)
elif what == "gemm16":
    M, N, K = 131072, 2304, 768
    a = torch.randn(M, K, device="cuda").half(); b = torch.randn(N, K, device="cuda").half()
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for _ in range(3):
        ops.gemm(a, b, M, N, K, out=out, cta_group=2, out_half=True)
    a2 = torch.randn(M, 3072, device="cuda").half(); b2 = torch.randn(768, 3072, device="cuda").half()
    out2 = torch.empty(M, 768, device="cuda")
    for _ in range(3):
        ops.gemm(a2, b2, M, 768, 3072, out=out2, cta_group=2)
elif what == "vq":
    z = torch.randn(131072, 32, device="cuda"); E = torch.randn(8192, 32, device="cuda")
    for _ in range(3):
        ops.vq_fwd(z, E, 1, 0.25)
elif what == "ln":
    x = torch.randn(131072, 768, device="cuda"); g = torch.ones(768, device="cuda"); b = torch.zeros(768, device="cuda")
    dres = torch.randn_like(x)
    sc = ops.grad_scale(dres)
    dy = ops.to_half(torch.randn_like(x), sc[0:1])                      # what the fp16 dgrad GEMM hands over
    for _ in range(3):
        y, m, r = ops.layernorm_fwd(x, g, b, False, out_half=True)       # the fp16 data path's configuration
        ops.layernorm_bwd(dy, x, m, r, g, dres, want_colsum=True, half_scale=sc[0:1], dy_scale=sc[1:2])
torch.cuda.synchronize()
