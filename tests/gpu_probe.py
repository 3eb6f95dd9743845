"""Diagnostic sweep run on the GPU box during development (not a pytest file):
    python tests/gpu_probe.py [group ...]
Each group runs in its own subprocess so that a trapped kernel cannot poison the rest."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def tf32_rn(t):
    import torch
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1fff).view(torch.float32)


def tf32_trunc(t):
    import torch
    i = t.contiguous().view(torch.int32)
    return (i & ~0x1fff).view(torch.float32)


def relerr(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def time_ms(fn, iters=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def g_gemm_basic(cg=1):
    import torch
    import enhancing_transformers_b200 as etb
    ops = etb.ops
    torch.manual_seed(0)
    dev = "cuda"
    for (M, N, K, bn) in [(128 * cg, 64, 32, 64), (256, 256, 64, 256), (512, 768, 768, 256), (384, 192, 96, 192),
                          (256, 96, 160, 64), (1024, 2304, 768, 0), (256, 128, 3072, 128)]:
        a = tf32_rn(torch.randn(M, K, device=dev))
        b = tf32_rn(torch.randn(N, K, device=dev))
        ref = a.double() @ b.double().t()
        c = ops.gemm(a, b, M, N, K, cta_group=cg, bn=bn)
        torch.cuda.synchronize()
        print(f"gemm NT cg={cg} M={M} N={N} K={K} bn={bn}: relerr {relerr(c.double(), ref):.3e}", flush=True)
    # epilogue
    M, N, K = 512, 768, 256
    a = tf32_rn(torch.randn(M, K, device=dev)); b = tf32_rn(torch.randn(N, K, device=dev))
    bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev); aux = torch.tanh(torch.randn(M, N, device=dev))
    pos = torch.randn(128, N, device=dev)
    base = (a.double() @ b.double().t())
    c = ops.gemm(a, b, M, N, K, bias=bias, cta_group=cg)
    print("  +bias", relerr(c.double(), base + bias.double()))
    c = ops.gemm(a, b, M, N, K, bias=bias, act=1, cta_group=cg)
    print("  +bias tanh", relerr(c.double(), torch.tanh(base + bias.double())))
    c = ops.gemm(a, b, M, N, K, bias=bias, res=res, cta_group=cg)
    print("  +bias +res", relerr(c.double(), base + bias.double() + res.double()))
    c = ops.gemm(a, b, M, N, K, bias=bias, res=pos, res_row_mod=128, cta_group=cg)
    print("  +bias +pos(mod 128)", relerr(c.double(), base + bias.double() + pos.double().repeat(M // 128, 1)))
    c = ops.gemm(a, b, M, N, K, aux=aux, cta_group=cg)
    print("  *tanh'", relerr(c.double(), base * (1 - aux.double() ** 2)))
    c = ops.gemm(a, b, M, N, K, round_out=True, cta_group=cg)
    print("  round_out", relerr(c.double(), tf32_rn(base.float()).double()), "lowbits", int((c.view(torch.int32) & 0x1fff).abs().max()))
    # hardware rounding mode of un-rounded operands
    a = torch.randn(M, K, device=dev); b = torch.randn(N, K, device=dev)
    c = ops.gemm(a, b, M, N, K, cta_group=cg).double()
    for nm, f in (("trunc", tf32_trunc), ("rn", tf32_rn)):
        print(f"  hw-vs-{nm}: {relerr(c, f(a).double() @ f(b).double().t()):.3e}")
    print(f"  hw-vs-fp32: {relerr(c, a.double() @ b.double().t()):.3e}")


def g_gemm_cg2():
    g_gemm_basic(2)


def g_gemm_major():
    import torch
    import enhancing_transformers_b200 as etb
    ops = etb.ops
    torch.manual_seed(1)
    dev = "cuda"
    for cg in (1, 2):
        for (M, N, K) in [(256, 256, 64), (512, 768, 3072), (256, 96, 160), (1024, 192, 768)]:
            # NN (dgrad): C[M,N] = A[M,K] . Bs[K,N]
            a = tf32_rn(torch.randn(M, K, device=dev)); bs = tf32_rn(torch.randn(K, N, device=dev))
            ref = a.double() @ bs.double()
            c = ops.gemm(a, bs, M, N, K, b_major=1, cta_group=cg)
            torch.cuda.synchronize()
            print(f"gemm NN cg={cg} M={M} N={N} K={K}: relerr {relerr(c.double(), ref):.3e}", flush=True)
        for (M, N, K, splits) in [(256, 256, 128, 1), (768, 768, 4096, 4), (128 * cg, 96, 2048, 2), (3072, 768, 2048, 2)]:
            # TN (wgrad): C[M,N] = As[Ktot,M]^T . Bs[Ktot,N]
            As = tf32_rn(torch.randn(K, M, device=dev)); Bs = tf32_rn(torch.randn(K, N, device=dev))
            ref = As.double().t() @ Bs.double()
            part = ops.gemm(As, Bs, M, N, K // splits, a_major=1, b_major=1, splits=splits, cta_group=cg)
            c = ops.splitk_reduce(part) if splits > 1 else part
            torch.cuda.synchronize()
            print(f"gemm TN cg={cg} M={M} N={N} Ktot={K} splits={splits}: relerr {relerr(c.double(), ref):.3e}", flush=True)


def g_gemm_perf():
    import torch
    import enhancing_transformers_b200 as etb
    ops = etb.ops
    dev = "cuda"
    M = 131072
    for (N, K) in [(2304, 768), (768, 768), (3072, 768), (768, 3072)]:
        a = tf32_rn(torch.randn(M, K, device=dev)); b = tf32_rn(torch.randn(N, K, device=dev))
        out = torch.empty(M, N, device=dev)
        for cg in (1, 2):
            for bn in (256, 128):
                ms = time_ms(lambda: ops.gemm(a, b, M, N, K, out=out, cta_group=cg, bn=bn), iters=5, warm=2)
                print(f"perf NT M={M} N={N} K={K} cg={cg} bn={bn}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        torch.backends.cuda.matmul.allow_tf32 = True
        ms = time_ms(lambda: torch.matmul(a, b.t(), out=out), iters=5, warm=2)
        print(f"   cuBLAS tf32: {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TFLOP/s")
        torch.backends.cuda.matmul.allow_tf32 = False
        del a, b, out
    # dgrad / wgrad shapes
    N, K = 3072, 768
    dy = tf32_rn(torch.randn(M, N, device=dev)); w = tf32_rn(torch.randn(N, K, device=dev)); x = tf32_rn(torch.randn(M, K, device=dev))
    for cg in (1, 2):
        ms = time_ms(lambda: ops.gemm(dy, w, M, K, N, b_major=1, cta_group=cg), iters=5, warm=2)
        print(f"perf NN dgrad cg={cg}: {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TFLOP/s")
        for splits in (4, 8, 16):
            ms = time_ms(lambda: ops.splitk_reduce(ops.gemm(dy, x, N, K, M // splits, a_major=1, b_major=1, splits=splits, cta_group=cg)), iters=5, warm=2)
            print(f"perf TN wgrad cg={cg} splits={splits}: {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TFLOP/s")


def g_vq():
    import numpy as np
    import torch
    import enhancing_transformers_b200 as etb
    from oracle import vitvq_oracle as O
    ops = etb.ops
    g = np.load(os.path.join(ROOT, "tests", "golden", "vq_cases.npz"))
    for tag, depth in (("plain", 1), ("res4", 4), ("res2", 2), ("clustered", 1)):
        z = torch.from_numpy(g[f"{tag}.z"]).cuda(); E = torch.from_numpy(g[f"{tag}.E"]).cuda()
        out, loss, idx = ops.vq_fwd(z, E, depth, 0.25)
        idx_ref = g[f"{tag}.idx"].reshape(-1, depth)
        print(f"vq {tag}: idx mismatches {(idx.cpu().numpy() != idx_ref).sum()} / {idx_ref.size}; out maxdiff "
              f"{np.abs(out.cpu().numpy() - g[f'{tag}.zq']).max():.3e}; loss {loss.item():.8f} ref {float(g[f'{tag}.loss']):.8f}")
        g_out = torch.from_numpy(g[f"{tag}.g_out"]).cuda() if f"{tag}.g_out" in g.files else torch.zeros_like(z)
        gl = float(g[f"{tag}.g_loss"]) if f"{tag}.g_loss" in g.files else 1.0
        gz, gE = ops.vq_bwd(z, E, idx, g_out, torch.tensor(gl, device="cuda"), depth > 1, 0.25)
        print(f"   bwd: gz relerr {relerr(gz.cpu(), torch.from_numpy(g[f'{tag}.gz'])):.3e} gE relerr {relerr(gE.cpu(), torch.from_numpy(g[f'{tag}.gE'])):.3e}")
    # larger random case vs numpy oracle
    torch.manual_seed(0)
    M, K = 8192, 8192
    z = torch.randn(M, 32); E = torch.randn(K, 32)
    ref = O.vq_lookup_np(z.numpy(), E.numpy())
    out, loss, idx = ops.vq_fwd(z.cuda(), E.cuda(), 1, 0.25)
    mism = np.nonzero(idx.cpu().numpy()[:, 0] != ref)[0]
    print(f"vq random M={M} K={K}: mismatches {len(mism)}")
    if len(mism):
        gap = O.vq_top2_gap_f64(z.numpy()[mism], E.numpy())
        print("   f64 top-2 gaps of mismatching rows:", gap[:10])
    for (M, depth) in ((131072, 1), (131072, 4)):
        z = torch.randn(M, 32, device="cuda"); Eg = torch.randn(8192, 32, device="cuda")
        ms = time_ms(lambda: ops.vq_fwd(z, Eg, depth, 0.25), iters=5, warm=2)
        print(f"vq perf M={M} depth={depth}: {ms:.3f} ms  {depth*2*32*8192*M/ms/1e9:.1f} TFLOP/s  {(M*(256+8*depth)+2**20)/ms/1e6:.1f} GB/s")
        out, loss, idx = ops.vq_fwd(z, Eg, depth, 0.25)
        gl = torch.tensor(1.0, device="cuda")
        ms = time_ms(lambda: ops.vq_bwd(z, Eg, idx, out, gl, depth > 1, 0.25), iters=5, warm=2)
        print(f"vq bwd perf M={M} depth={depth}: {ms:.3f} ms")


def g_rowwise():
    import torch
    import torch.nn.functional as F
    import enhancing_transformers_b200 as etb
    ops = etb.ops
    torch.manual_seed(0)
    for (M, D) in ((64, 64), (1000, 96), (4096, 768), (512, 1280), (256, 512)):
        x = torch.randn(M, D, device="cuda") * 2 + 0.5
        g = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
        y, mean, rstd = ops.layernorm_fwd(x, g, b, False)
        xr = x.double().requires_grad_(True); gr = g.double().requires_grad_(True); br = b.double().requires_grad_(True)
        yr = F.layer_norm(xr, (D,), gr, br, 1e-5)
        dy = torch.randn(M, D, device="cuda"); dres = torch.randn(M, D, device="cuda")
        yr.backward(dy.double())
        dx, dg, db = ops.layernorm_bwd(dy, x, mean, rstd, g, dres)
        print(f"ln M={M} D={D}: y {relerr(y.double(), yr.detach()):.2e} dx {relerr(dx.double(), xr.grad + dres.double()):.2e} "
              f"dg {relerr(dg.double(), gr.grad):.2e} db {relerr(db.double(), br.grad):.2e}")
    img = torch.rand(3, 3, 64, 32, device="cuda")
    from oracle import vitvq_oracle as O
    p = ops.patchify(img, 8, False)
    print("patchify", (p.cpu() - O.patchify(img.cpu(), 8).reshape(p.shape)).abs().max().item())
    bias = torch.randn(3, device="cuda")
    u = ops.unpatchify(p, bias, 3, 3, 64, 32, 8)
    print("unpatchify", (u - (img + bias.view(1, 3, 1, 1))).abs().max().item())
    x = torch.randn(5000, 768, device="cuda")
    print("colsum", relerr(ops.colsum(x).double(), x.double().sum(0)))
    x = torch.randn(4096, device="cuda")
    print("round_tf32", (ops.round_tf32(x) - tf32_rn(x)).abs().max().item())
    M, D = 131072, 768
    x = torch.randn(M, D, device="cuda"); g = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
    ms = time_ms(lambda: ops.layernorm_fwd(x, g, b, True))
    print(f"ln fwd perf: {ms:.3f} ms {2*M*D*4/ms/1e6:.0f} GB/s")
    y, mean, rstd = ops.layernorm_fwd(x, g, b, True)
    ms = time_ms(lambda: ops.layernorm_bwd(y, x, mean, rstd, g, y))
    print(f"ln bwd perf: {ms:.3f} ms {4*M*D*4/ms/1e6:.0f} GB/s")
    ms = time_ms(lambda: ops.colsum(x))
    print(f"colsum perf: {ms:.3f} ms {M*D*4/ms/1e6:.0f} GB/s")


GROUPS = {k[2:]: v for k, v in list(globals().items()) if k.startswith("g_")}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        GROUPS[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(GROUPS)
    for n in names:
        t0 = time.time()
        print(f"===== {n}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", n], timeout=300)
            print(f"===== {n}: exit {r.returncode} in {time.time()-t0:.1f}s", flush=True)
        except subprocess.TimeoutExpired:
            print(f"===== {n}: TIMEOUT", flush=True)
