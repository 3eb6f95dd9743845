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
    dy = torch.randn_like(x); dres = torch.randn_like(x)
    ms = time_ms(lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dres))
    print(f"ln bwd perf: {ms:.3f} ms {4*M*D*4/ms/1e6:.0f} GB/s")
    ms = time_ms(lambda: ops.colsum(x))
    print(f"colsum perf: {ms:.3f} ms {M*D*4/ms/1e6:.0f} GB/s")


def g_attention():
    import torch
    import enhancing_transformers_b200 as etb
    ops = etb.ops
    torch.manual_seed(0)
    for (B, N, heads, dh) in [(2, 16, 2, 32), (1, 24, 3, 64), (2, 200, 2, 64), (2, 1024, 4, 64), (1, 130, 1, 32)]:
        inner = heads * dh
        qkv = tf32_rn(torch.randn(B * N, 3 * inner, device="cuda"))
        scale = dh ** -0.5
        o, lse = ops.attention_fwd(qkv, B, N, heads, dh, scale, False)
        q, k, v = (t.reshape(B, N, heads, dh).permute(0, 2, 1, 3).double() for t in qkv.split(inner, dim=-1))
        q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
        s = (q @ k.transpose(-1, -2)) * scale
        p = torch.softmax(s, -1)
        oref = (p @ v).permute(0, 2, 1, 3).reshape(B * N, inner)
        lref = torch.logsumexp(s, -1).reshape(-1)
        do = tf32_rn(torch.randn(B * N, inner, device="cuda"))
        oref.backward(do.double())
        dqkv = ops.attention_bwd(qkv, o, lse, do, B, N, heads, dh, scale, False)
        dref = torch.cat([t.grad.permute(0, 2, 1, 3).reshape(B * N, inner) for t in (q, k, v)], dim=-1)
        print(f"attn B={B} N={N} h={heads} dh={dh}: o {relerr(o.double(), oref.detach()):.2e} lse {relerr(lse.double(), lref.detach()):.2e} "
              f"dq {relerr(dqkv[:, :inner].double(), dref[:, :inner]):.2e} dk {relerr(dqkv[:, inner:2*inner].double(), dref[:, inner:2*inner]):.2e} "
              f"dv {relerr(dqkv[:, 2*inner:].double(), dref[:, 2*inner:]):.2e}", flush=True)
    B, N, heads, dh = 64, 1024, 12, 64
    inner = heads * dh
    qkv = tf32_rn(torch.randn(B * N, 3 * inner, device="cuda"))
    ms = time_ms(lambda: ops.attention_fwd(qkv, B, N, heads, dh, 0.125, True), iters=5, warm=2)
    fl = 4 * B * heads * N * N * dh
    print(f"attn fwd perf B={B}: {ms:.3f} ms {fl/ms/1e9:.1f} TFLOP/s")
    o, lse = ops.attention_fwd(qkv, B, N, heads, dh, 0.125, True)
    do = tf32_rn(torch.randn(B * N, inner, device="cuda"))
    ms = time_ms(lambda: ops.attention_bwd(qkv, o, lse, do, B, N, heads, dh, 0.125, True), iters=5, warm=2)
    print(f"attn bwd perf B={B}: {ms:.3f} ms {2.5*fl/ms/1e9:.1f} TFLOP/s (5 GEMM-units counted)")


def _load_modules(cfg, sd, device="cuda"):
    import torch
    import enhancing_transformers_b200 as etb
    e, d, q = cfg["encoder"], cfg["decoder"], cfg["quantizer"]
    enc = etb.ViTEncoder(cfg["image_size"], cfg["patch_size"], **e)
    dec = etb.ViTDecoder(cfg["image_size"], cfg["patch_size"], **d)
    vq = etb.VectorQuantizer(**q)
    pre = etb.QuantLinear(e["dim"], q["embed_dim"]); post = etb.QuantLinear(q["embed_dim"], d["dim"])
    for pfx, m in (("encoder.", enc), ("decoder.", dec), ("quantizer.", vq), ("pre_quant.", pre), ("post_quant.", post)):
        m.load_state_dict({k[len(pfx):]: v for k, v in sd.items() if k.startswith(pfx)}, strict=True)
        m.to(device)
    return enc, dec, vq, pre, post


def _step(mods, img):
    enc, dec, vq, pre, post = mods
    h = enc(img)
    z = pre(h)
    zq, qloss, idx = vq(z)
    rec = dec(post(zq))
    loss = ((rec - img) ** 2).mean() + qloss
    return loss, rec, idx, h, z


def g_model_tiny():
    import numpy as np
    import torch
    g = np.load(os.path.join(ROOT, "tests", "golden", "vit_tiny.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    cfg = dict(image_size=32, patch_size=8, encoder=dict(dim=64, depth=2, heads=2, mlp_dim=128),
               decoder=dict(dim=96, depth=2, heads=3, mlp_dim=160, dim_head=32), quantizer=dict(embed_dim=32, n_embed=256))
    mods = _load_modules(cfg, sd)
    img = torch.from_numpy(g["img"]).cuda()
    loss, rec, idx, h, z = _step(mods, img)
    print("tiny: enc", relerr(h.cpu(), torch.from_numpy(g["enc_out"])), "z", relerr(z.cpu(), torch.from_numpy(g["z"])),
          "idx mismatch", int((idx.cpu().numpy() != g["idx"]).sum()), "rec", relerr(rec.cpu(), torch.from_numpy(g["rec"])),
          "loss", loss.item(), float(g["loss"]))
    loss.backward()
    names = dict(encoder=mods[0], decoder=mods[1], quantizer=mods[2], pre_quant=mods[3], post_quant=mods[4])
    worst = 0
    for k in g.files:
        if k.startswith("grad."):
            mod, _, pname = k[5:].partition(".")
            p = dict(names[mod].named_parameters())[pname]
            e = relerr(p.grad.cpu(), torch.from_numpy(g[k]))
            worst = max(worst, e)
            if e > 3e-3:
                print("   grad", k, e)
    print("tiny: worst grad relerr", worst)


def g_model_small():
    import torch
    from oracle import vitvq_oracle as O
    import enhancing_transformers_b200 as etb
    for name, B in (("small", 2), ("base", 1)):
        cfg = O.CONFIGS[name]
        sd = O.init_vitvq_sd(cfg, seed=0)
        torch.manual_seed(0)
        img = torch.rand(B, 3, 256, 256)
        sdg = {k: v.clone().requires_grad_(v.is_floating_point() and "pos_embedding" not in k) for k, v in sd.items()}
        torch.set_num_threads(os.cpu_count())
        t0 = time.time()
        loss_ref, rec_ref, idx_ref = O.vitvq_loss(sdg, img, cfg)
        loss_ref.backward()
        print(f"{name}: oracle fwd+bwd B={B} took {time.time()-t0:.1f}s on {os.cpu_count()} threads")
        mods = _load_modules(cfg, sd)
        loss, rec, idx, h, z = _step(mods, img.cuda())
        loss.backward()
        agree = (idx.cpu() == idx_ref).float().mean().item()
        print(f"{name}: loss {loss.item():.6f} ref {loss_ref.item():.6f}; rec maxabs-rel {relerr(rec.cpu(), rec_ref.detach()):.3e} "
              f"rel-l2 {((rec.cpu()-rec_ref.detach()).norm()/rec_ref.detach().norm()).item():.3e}; idx agree {agree:.5f}")
        names = dict(encoder=mods[0], decoder=mods[1], quantizer=mods[2], pre_quant=mods[3], post_quant=mods[4])
        worst = []
        for k, v in sdg.items():
            if v.grad is None:
                continue
            mod, _, pname = k.partition(".")
            p = dict(names[mod].named_parameters())[pname]
            worst.append((((p.grad.cpu() - v.grad).norm() / v.grad.norm().clamp_min(1e-30)).item(), k))
        worst.sort(reverse=True)
        print("   worst grad rel-l2:", [(f"{e:.2e}", k) for e, k in worst[:5]])
        del mods
        torch.cuda.empty_cache()


def g_model_perf():
    import torch
    from oracle import vitvq_oracle as O
    import enhancing_transformers_b200 as etb
    cfg = O.CONFIGS["base"]
    sd = O.init_vitvq_sd(cfg, seed=0)
    mods = _load_modules(cfg, sd)
    for cg in (1, 2):
        etb.functional.GEMM_CTA_GROUP = cg
        for B in (16, 64, 128):
            img = torch.rand(B, 3, 256, 256, device="cuda")

            def step():
                for m in mods:
                    m.zero_grad(set_to_none=True)
                loss = _step(mods, img)[0]
                loss.backward()
            try:
                ms = time_ms(step, iters=3, warm=2)
                fl = 3 * O.flops_per_image(cfg) * B
                print(f"base fwd+bwd cg={cg} B={B}: {ms:.1f} ms  {B/ms*1e3:.1f} img/s  {fl/ms/1e9:.1f} TFLOP/s  mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
            except Exception as ex:
                print(f"base B={B} cg={cg} failed: {ex}")
                break


def g_precision():
    """error of the tf32 path vs the oracle evaluated in fp64 on the GPU, per stage"""
    import torch
    from oracle import vitvq_oracle as O
    import enhancing_transformers_b200 as etb
    torch.backends.cuda.matmul.allow_tf32 = False
    import sys
    which = [a for a in sys.argv[3:]] or ["small", "base"]
    for name in which:
        B = 2 if name == "large" else 4
        cfg = O.CONFIGS[name]
        sd = O.init_vitvq_sd(cfg, seed=0)
        img = torch.rand(B, 3, 256, 256, generator=torch.Generator().manual_seed(0))
        sd64 = {k: v.double().cuda() for k, v in sd.items()}
        p, g = cfg["patch_size"], cfg["image_size"] // cfg["patch_size"]
        e, d = cfg["encoder"], cfg["decoder"]
        with torch.no_grad():
            h64 = O.vit_encoder(sd64, img.double().cuda(), patch=p, depth=e["depth"], heads=e["heads"], prefix="encoder.")
            z64 = h64 @ sd64["pre_quant.weight"].t() + sd64["pre_quant.bias"]
            mods = _load_modules(cfg, sd)
            enc, dec, vq, pre, post = mods
            h = enc(img.cuda()); z = pre(h)
            zq, _, idx = vq(z)
            qc = cfg["quantizer"]
            zq64, _, idx64 = O.vq_forward(z64.float(), sd["quantizer.embedding.weight"].cuda(), qc.get("beta", 0.25),
                                          qc.get("use_residual", False), qc.get("num_quantizers"))
            rec = dec(post(zq))
            t64 = zq.double() @ sd64["post_quant.weight"].t() + sd64["post_quant.bias"]
            rec64 = O.vit_decoder(sd64, t64, patch=p, depth=d["depth"], heads=d["heads"], grid_hw=(g, g), prefix="decoder.")
        print(f"{name}: enc relmax {relerr(h.double(), h64):.2e} rel-l2 {((h.double()-h64).norm()/h64.norm()).item():.2e} | z relmax {relerr(z.double(), z64):.2e} | "
              f"idx agree {(idx == idx64).float().mean().item():.5f} | dec-only (same codes) relmax {relerr(rec.double(), rec64):.2e} rel-l2 {((rec.double()-rec64).norm()/rec64.norm()).item():.2e}", flush=True)
        del mods
        torch.cuda.empty_cache()


def g_f16perf():
    """GEMM throughput of the two operand kinds at the config-2 shapes (M = 131072 tokens)"""
    import torch
    import enhancing_transformers_b200 as etb
    ops = etb.ops
    M = 131072
    H = torch.float16
    alpha = torch.tensor([0.5], device="cuda")
    shapes = [("to_qkv fwd", 2304, 768, {}), ("to_out fwd", 768, 768, {}), ("mlp1 fwd", 3072, 768, {}), ("mlp2 fwd", 768, 3072, {})]
    for nm, N, K, _ in shapes:
        for kind in ("tf32", "f16"):
            dt = H if kind == "f16" else torch.float32
            a = torch.randn(M, K, device="cuda").to(dt); b = torch.randn(N, K, device="cuda").to(dt)
            ms = time_ms(lambda: ops.gemm(a, b, M, N, K, cta_group=2), iters=5, warm=2)
            line = f"{nm:12s} N={N:5d} K={K:5d} {kind:5s} out32: {ms:.3f} ms {2*M*N*K/ms/1e9:7.1f} TFLOP/s"
            if kind == "f16":
                bias = torch.randn(N, device="cuda")
                ms2 = time_ms(lambda: ops.gemm(a, b, M, N, K, bias=bias, act=1, out_half=True, cta_group=2), iters=5, warm=2)
                line += f" | out16+tanh: {ms2:.3f} ms {2*M*N*K/ms2/1e9:7.1f}"
            print(line, flush=True)
            del a, b
    # dgrad (B MN-major) and wgrad (both MN-major, split-K)
    for nm, N, K in [("mlp2 dgrad", 3072, 768), ("mlp1 dgrad", 768, 3072), ("qkv dgrad", 768, 2304)]:
        for kind in ("tf32", "f16"):
            dt = H if kind == "f16" else torch.float32
            a = torch.randn(M, K, device="cuda").to(dt); b = torch.randn(K, N, device="cuda").to(dt)
            ms = time_ms(lambda: ops.gemm(a, b, M, N, K, b_major=1, cta_group=2), iters=5, warm=2)
            print(f"{nm:12s} N={N:5d} K={K:5d} {kind:5s}: {ms:.3f} ms {2*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
            del a, b
    for nm, R, C in [("mlp wgrad", 768, 3072), ("qkv wgrad", 2304, 768), ("out wgrad", 768, 768)]:
        for kind in ("tf32", "f16"):
            dt = H if kind == "f16" else torch.float32
            dy = torch.randn(M, R, device="cuda").to(dt); x = torch.randn(M, C, device="cuda").to(dt)
            splits = ops.pick_splits(M, R, C, k_atom=64 if kind == "f16" else 32)
            def run():
                part = ops.gemm(dy, x, R, C, M // splits, a_major=1, b_major=1, splits=splits, cta_group=2)
                return ops.splitk_reduce(part, alpha=alpha) if splits > 1 else part
            ms = time_ms(run, iters=5, warm=2)
            print(f"{nm:12s} [{R}x{C}] splits={splits} {kind:5s}: {ms:.3f} ms {2*M*R*C/ms/1e9:7.1f} TFLOP/s", flush=True)
            del dy, x


def g_modes():
    """fwd+bwd step time of the hot path per precision mode (base, B from argv, default 32)"""
    import sys
    import torch
    from oracle import vitvq_oracle as O
    import enhancing_transformers_b200 as etb
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    cfg = O.CONFIGS["base"]
    sd = O.init_vitvq_sd(cfg, seed=0)
    mods = _load_modules(cfg, sd)
    params = [p for m in mods for p in m.parameters()]
    img = torch.rand(B, 3, 256, 256, device="cuda")
    for mode in ("tf32", "fp16") + (("parity",) if B <= 8 else ()):
        etb.set_precision(mode)

        def step():
            for p in params:
                p.grad = None
            loss = _step(mods, img)[0]
            loss.backward()
            return loss
        ms = time_ms(step, iters=3, warm=2)
        print(f"mode {mode:6s} base B={B}: {ms:.1f} ms/step  {B/ms*1e3:.1f} img/s  loss {step().item():.6f} "
              f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)


def g_ln16():
    """LayerNorm forward / backward in the fp16 data path's configuration (M=131072, D=768): ms and effective HBM GB/s"""
    import torch
    import enhancing_transformers_b200 as etb
    ops = etb.ops
    M, D = 131072, 768
    x = torch.randn(M, D, device="cuda"); g = torch.ones(D, device="cuda"); b = torch.zeros(D, device="cuda")
    dres = torch.randn_like(x)
    sc = ops.grad_scale(dres)
    dy = ops.to_half(torch.randn_like(x), sc[0:1])
    y, m, r = ops.layernorm_fwd(x, g, b, False, out_half=True)
    ms = time_ms(lambda: ops.layernorm_fwd(x, g, b, False, out_half=True), iters=10, warm=3)
    print(f"ln_fwd  fp16 out : {ms*1e3:7.1f} us  {M*D*6/ms/1e6:7.0f} GB/s")
    ms = time_ms(lambda: ops.layernorm_bwd(dy, x, m, r, g, dres, want_colsum=True, half_scale=sc[0:1], dy_scale=sc[1:2]), iters=10, warm=3)
    print(f"ln_bwd  fp16 dy, fp32+fp16 dx, colsum : {ms*1e3:7.1f} us  {M*D*16/ms/1e6:7.0f} GB/s (incl. the parameter reduce launch)")
    dy32 = torch.randn_like(x)
    for name, d_, kw, nbytes in (("fp32 dy, fp32 dx", dy32, {}, 16), ("fp32 dy, +colsum", dy32, dict(want_colsum=True), 16),
                                 ("fp32 dy, +fp16 dx", dy32, dict(half_scale=sc[0:1]), 18), ("fp16 dy, fp32 dx", dy, dict(dy_scale=sc[1:2]), 14),
                                 ("fp16 dy, +fp16 dx", dy, dict(dy_scale=sc[1:2], half_scale=sc[0:1]), 16),
                                 ("fp16 dy, +colsum", dy, dict(dy_scale=sc[1:2], want_colsum=True), 14)):
        ms = time_ms(lambda: ops.layernorm_bwd(d_, x, m, r, g, dres, **kw), iters=10, warm=3)
        print(f"ln_bwd  {name:20s}: {ms*1e3:7.1f} us  {M*D*nbytes/ms/1e6:7.0f} GB/s")


def g_attn16():
    """per-kernel time of the fp16 attention core vs the tf32 one (B from argv, default 64; base heads)"""
    import sys
    import torch
    import enhancing_transformers_b200 as etb
    ops = etb.ops
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    N, heads, dh = 1024, 12, 64
    inner = heads * dh
    qkv32 = tf32_rn(torch.randn(B * N, 3 * inner, device="cuda"))
    qkv16 = qkv32.half()
    fl = 4 * B * heads * N * N * dh
    ms = time_ms(lambda: ops.attention_fwd(qkv32, B, N, heads, dh, 0.125, True), iters=5, warm=2)
    print(f"tf32 fwd B={B}: {ms:.3f} ms {fl/ms/1e9:.1f} TFLOP/s")
    ms = time_ms(lambda: ops.attention_f16_fwd(qkv16, B, N, heads, dh, 0.125), iters=5, warm=2)
    print(f"f16  fwd B={B}: {ms:.3f} ms {fl/ms/1e9:.1f} TFLOP/s")
    o32, lse = ops.attention_fwd(qkv32, B, N, heads, dh, 0.125, True)
    o16, lse16 = ops.attention_f16_fwd(qkv16, B, N, heads, dh, 0.125)
    do32 = tf32_rn(torch.randn(B * N, inner, device="cuda"))
    do16 = do32.half()
    ms = time_ms(lambda: ops.attention_bwd(qkv32, o32, lse, do32, B, N, heads, dh, 0.125, True), iters=5, warm=2)
    print(f"tf32 bwd B={B}: {ms:.3f} ms {2.5*fl/ms/1e9:.1f} TFLOP/s (algorithmic)")
    ms = time_ms(lambda: ops.attention_f16_bwd(qkv16, o16, lse16, do16, B, N, heads, dh, 0.125), iters=5, warm=2)
    print(f"f16  bwd B={B}: {ms:.3f} ms {2.5*fl/ms/1e9:.1f} TFLOP/s (algorithmic)")


def g_attn16_trace():
    """timeline of the fp16 dKV kernel's two issuer warps and one softmax warp (needs B200VQ_LIB=.../libb200vq_trace16.so,
    `make -C enhancing_transformers_b200/csrc trace16`)"""
    import ctypes
    import torch
    import enhancing_transformers_b200 as etb
    ops = etb.ops
    L = ctypes.CDLL(etb._lib.LIB_PATH)
    B, N, heads, dh = 32, 1024, 12, 64
    inner = heads * dh
    qkv = torch.randn(B * N, 3 * inner, device="cuda").half()
    o, lse = ops.attention_f16_fwd(qkv, B, N, heads, dh, 0.125)
    do = torch.randn(B * N, inner, device="cuda").half()
    cap = 64
    buf = (ctypes.c_longlong * (3 * 2 * cap))()
    ops.attention_f16_bwd(qkv, o, lse, do, B, N, heads, dh, 0.125)
    if os.environ.get("TRACE_FWD"):      # a library built with -DB200_ATTN16_TRACE_FWD dumps the forward kernel's timeline instead
        ops.attention_f16_fwd(qkv, B, N, heads, dh, 0.125)
    torch.cuda.synchronize()
    L.b200vq_trace16_read(buf)
    ev = []
    for r in range(3):
        for i in range(cap):
            e, t = buf[r * 2 * cap + 2 * i], buf[r * 2 * cap + 2 * i + 1]
            if t:
                ev.append((t, r, e))
    ev.sort()
    names = {100: "A: wait qd_full", 101: "A: got qd_full, wait sfree", 102: "A: got sfree", 103: "A: issued S^T, dP^T",
             110: "B: got qd_full, wait p_full", 111: "B: got p_full", 112: "B: issued dV, dK",
             120: "sm: wait s_full", 121: "sm: got s_full", 122: "sm: tcgen05.ld done", 123: "sm: computed", 124: "sm: stored + arrived",
             200: "A: wait k_full+sfree", 201: "A: got them", 202: "A: issued S", 210: "B: wait v_full+o_empty+p_full", 211: "B: got them",
             212: "B: issued PV", 220: "sm: S in registers", 221: "sm: row max exchanged", 222: "sm: exps done", 223: "sm: P stored + arrived",
             224: "sm: got next S / prev PV", 225: "sm: PV folded"}
    t0 = ev[0][0] if ev else 0
    prev = t0
    for t, r, e in ev:
        print(f"{t - t0:9d} (+{t - prev:5d})  {names.get(e, e)}")
        prev = t


def g_attn_trace():
    """timeline of the dQ kernel's MMA warp and two softmax warps (needs B200VQ_LIB=.../libb200vq_trace.so)"""
    import ctypes
    import torch
    import enhancing_transformers_b200 as etb
    ops = etb.ops
    L = ctypes.CDLL(etb._lib.LIB_PATH)
    B, N, heads, dh = 32, 1024, 12, 64
    inner = heads * dh
    qkv = tf32_rn(torch.randn(B * N, 3 * inner, device="cuda"))
    o, lse = ops.attention_fwd(qkv, B, N, heads, dh, 0.125, True)
    do = tf32_rn(torch.randn(B * N, inner, device="cuda"))
    buf = (ctypes.c_longlong * (3 * 72))()
    ops.attention_bwd(qkv, o, lse, do, B, N, heads, dh, 0.125, True)
    torch.cuda.synchronize()
    L.b200vq_trace_read(buf)
    ev = []
    for r in range(3):
        for i in range(36):
            e, t = buf[r * 72 + 2 * i], buf[r * 72 + 2 * i + 1]
            if t:
                ev.append((t, r, e))
    ev.sort()
    t0 = ev[0][0]
    names = {100: "wait k_full", 107: "got k, wait sfree", 101: "got sfree", 102: "issued S", 103: "wait v/o_empty", 105: "got them, wait p_full", 104: "got p_full", 250: "max exchanged", 290: "pv accumulated",
             106: "issued PV", 200: "wait s_full", 220: "got s_full", 240: "ld+max done", 260: "compute done", 280: "arrived p_full"}
    role = {0: "mmaA ", 1: "mmaB ", 2: "sm2  "}
    start = 0
    prev = ev[start][0]
    for t, r, e in ev[start:start + 100]:
        print(f"{t - t0:9d} (+{t - prev:5d})  {role[r]}{names.get(e, e)}")
        prev = t


def g_fwd_trace():
    """timeline of the forward kernel's two partner softmax warps and issuer B (needs B200VQ_LIB=.../libb200vq_trace.so)"""
    import ctypes
    import torch
    import enhancing_transformers_b200 as etb
    ops = etb.ops
    L = ctypes.CDLL(etb._lib.LIB_PATH)
    B, N, heads, dh = 32, 1024, 12, 64
    inner = heads * dh
    qkv = tf32_rn(torch.randn(B * N, 3 * inner, device="cuda"))
    for _ in range(3):
        o, lse = ops.attention_fwd(qkv, B, N, heads, dh, 0.125, True)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (3 * 72))()
    L.b200vq_trace_read(buf)
    ev = []
    for r in range(3):
        for i in range(36):
            e, t = buf[r * 72 + 2 * i], buf[r * 72 + 2 * i + 1]
            if t:
                ev.append((t, r, e))
    ev.sort()
    t0 = ev[0][0]
    names = {103: "wait v/o_empty", 105: "got them, wait p_full", 104: "got p_full", 106: "issued PV",
             200: "wait s_full", 220: "got s_full", 230: "S loaded", 240: "max done", 250: "max exchanged", 260: "exp done",
             280: "P stored + arrived", 290: "pv accumulated"}
    role = {0: "smA  ", 1: "smB  ", 2: "mmaB "}
    start = 0
    last = {}
    for t, r, e in ev[start:start + 90]:
        print(f"{t - t0:9d} (+{t - last.get(r, t):5d})  {'      ' * r}{role[r]}{names.get(e, e)}")
        last[r] = t


GROUPS = {k[2:]: v for k, v in list(globals().items()) if k.startswith("g_")}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        GROUPS[sys.argv[2]]()
        sys.exit(0)
    names = [a for a in sys.argv[1:] if a in GROUPS] or list(GROUPS)
    extra = [a for a in sys.argv[1:] if a not in GROUPS]
    for n in names:
        t0 = time.time()
        print(f"===== {n}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", n] + extra, timeout=300)
            print(f"===== {n}: exit {r.returncode} in {time.time()-t0:.1f}s", flush=True)
        except subprocess.TimeoutExpired:
            print(f"===== {n}: TIMEOUT", flush=True)
