"""Shapes of the reference's shipped stage-1 configs and the algorithmic FLOP model used by bench.py.

The dicts restate ``configs/imagenet_vitvq_{small,base,large}.yaml:7-19`` of the reference
(``model.params.{image_size,patch_size,encoder,decoder,quantizer}``): they are splatted into the
constructors exactly as ``vitvqgan.py:35-37`` does.  ``base_rq4`` is BASELINE.json config 3 (base with
``use_residual=True, num_quantizers=4``); ``tiny`` is not a reference config (test miniature)."""
from __future__ import annotations

CONFIGS = {
    "small": dict(image_size=256, patch_size=8,
                  encoder=dict(dim=512, depth=8, heads=8, mlp_dim=2048),
                  decoder=dict(dim=512, depth=8, heads=8, mlp_dim=2048),
                  quantizer=dict(embed_dim=32, n_embed=8192)),
    "base": dict(image_size=256, patch_size=8,
                 encoder=dict(dim=768, depth=12, heads=12, mlp_dim=3072),
                 decoder=dict(dim=768, depth=12, heads=12, mlp_dim=3072),
                 quantizer=dict(embed_dim=32, n_embed=8192)),
    "base_rq4": dict(image_size=256, patch_size=8,
                     encoder=dict(dim=768, depth=12, heads=12, mlp_dim=3072),
                     decoder=dict(dim=768, depth=12, heads=12, mlp_dim=3072),
                     quantizer=dict(embed_dim=32, n_embed=8192, use_residual=True, num_quantizers=4)),
    "large": dict(image_size=256, patch_size=8,
                  encoder=dict(dim=512, depth=8, heads=8, mlp_dim=2048),
                  decoder=dict(dim=1280, depth=32, heads=16, mlp_dim=5120),
                  quantizer=dict(embed_dim=32, n_embed=8192)),
    "tiny": dict(image_size=64, patch_size=8,
                 encoder=dict(dim=128, depth=2, heads=2, mlp_dim=256),
                 decoder=dict(dim=192, depth=2, heads=3, mlp_dim=384),
                 quantizer=dict(embed_dim=32, n_embed=512)),
}


def quantizer_depth(cfg: dict) -> int:
    q = cfg["quantizer"]
    return int(q.get("num_quantizers") or 1) if q.get("use_residual") else 1


def flops_per_image(cfg: dict) -> float:
    """Algorithmic *forward* FLOPs per image (fwd+bwd = 3x, recompute not counted):
    1024 * [L_e blk_e + L_d blk_d + 2*192*D_e + 2*192*D_d + 2*D_e*32 + 2*32*D_d + T*2*32*8192],
    blk(D, inner, mlp) = 6*D*inner + 4*N*inner + 2*inner*D + 4*D*mlp, inner = 64*heads."""
    n = (cfg["image_size"] // cfg["patch_size"]) ** 2
    pd = 3 * cfg["patch_size"] ** 2
    q = cfg["quantizer"]

    def blk(c):
        dim, inner, mlp = c["dim"], c.get("dim_head", 64) * c["heads"], c["mlp_dim"]
        return 6 * dim * inner + 4 * n * inner + 2 * inner * dim + 4 * dim * mlp
    e, d = cfg["encoder"], cfg["decoder"]
    per_tok = (e["depth"] * blk(e) + d["depth"] * blk(d) + 2 * pd * e["dim"] + 2 * pd * d["dim"]
               + 2 * e["dim"] * q["embed_dim"] + 2 * q["embed_dim"] * d["dim"]
               + quantizer_depth(cfg) * 2 * q["embed_dim"] * q["n_embed"])
    return float(n * per_tok)


def gemm_flops_per_image(cfg: dict) -> float:
    """forward FLOPs of the dense projections only (everything `gemm_*_kernel` executes)"""
    n = (cfg["image_size"] // cfg["patch_size"]) ** 2
    q = cfg["quantizer"]
    attn_core = sum(c["depth"] * 4 * n * c.get("dim_head", 64) * c["heads"] for c in (cfg["encoder"], cfg["decoder"]))
    return flops_per_image(cfg) - n * (attn_core + quantizer_depth(cfg) * 2 * q["embed_dim"] * q["n_embed"])
