"""world_size-2 gloo test of the multi-GPU host logic: images are
sharded evenly across ranks, every rank takes the *per-rank* mean loss, and the gradient
all-reduce(mean) reproduces the single-process gradient on the concatenated batch -- including
the dense codebook gradient.  The oracle stands in for the CUDA modules (no GPU here); on the GPU
the same wiring runs through bench.py's DistributedDataParallel wrapper."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = dict(image_size=32, patch_size=8, encoder=dict(dim=32, depth=1, heads=1, mlp_dim=32, dim_head=32),
           decoder=dict(dim=32, depth=1, heads=1, mlp_dim=32, dim_head=32), quantizer=dict(embed_dim=32, n_embed=64))


class OracleNet(torch.nn.Module):
    def __init__(self, sd):
        super().__init__()
        self.names = list(sd)
        self.params = torch.nn.ParameterList([torch.nn.Parameter(v.clone(), requires_grad="pos_embedding" not in k)
                                              for k, v in sd.items()])

    def forward(self, img):
        from oracle import vitvq_oracle as O
        sd = dict(zip(self.names, self.params))
        return O.vitvq_loss(sd, img, CFG)[0]


def _worker(rank, world, port, q, flat):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import vitvq_oracle as O
    torch.set_num_threads(1)
    sd = O.init_vitvq_sd(CFG, seed=0)
    imgs = torch.rand(4, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    shard = imgs[rank * 2:(rank + 1) * 2]                       # even split, as main.py's DDP does
    if flat:      # the package's own reduction (what bench.py uses on the GPUs): one flat all-reduce after backward
        from enhancing_transformers_b200.parallel import FlatGradients, allreduce_gradients
        local = OracleNet(sd)
        if flat == "views":       # .grad are views into one flat buffer, reduced in place
            fg = FlatGradients(local.parameters())
            fg.zero_()
            local(shard).backward()
            fg.allreduce()
        else:
            local(shard).backward()
            allreduce_gradients(local.parameters())
        names, params = local.names, local.params
    else:
        net = torch.nn.parallel.DistributedDataParallel(OracleNet(sd))
        net(shard).backward()
        names, params = net.module.names, net.module.params
    grads = {n: p.grad.clone() for n, p in zip(names, params) if p.grad is not None}
    if rank == 0:
        single = OracleNet(sd)
        single(imgs).backward()
        worst = 0.0
        for n, p in zip(single.names, single.params):
            if p.grad is None:
                continue
            err = (grads[n] - p.grad).abs().max().item() / (p.grad.abs().max().item() + 1e-30)
            worst = max(worst, err)
        q.put(worst)
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("flat", [False, True, "views"], ids=["torch_ddp", "flat_allreduce", "flat_views"])
def test_sharded_ddp_gradients_equal_single_process(flat):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, flat)) for r in range(2)]
    for p in procs:
        p.start()
    worst = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert worst < 1e-4, worst


# ------------------------------------------------------------------------------------------------
# the PRODUCT modules (stage-1 hot path and stage-2 GPT) under the same sharding, with the C-ABI calls replaced by the
# torch stand-ins of tests/emulated_ops.py (no GPU here): their autograd Functions hand gradients to parameters whose .grad
# are views into FlatGradients' flat buffer, which is what bench.py runs on N > 1 GPUs
# ------------------------------------------------------------------------------------------------
class _Patcher:
    """the two monkeypatch methods emulated_ops.install uses, without pytest (spawned worker processes)"""

    @staticmethod
    def setattr(obj, name, value):
        setattr(obj, name, value)

    @staticmethod
    def setitem(mapping, key, value):
        mapping[key] = value


PCFG = dict(image_size=32, patch_size=8, encoder=dict(dim=64, depth=1, heads=2, mlp_dim=64, dim_head=32),
            decoder=dict(dim=64, depth=1, heads=2, mlp_dim=64, dim_head=32), quantizer=dict(embed_dim=32, n_embed=64))
GCFG = dict(vocab_cond_size=5, vocab_img_size=64, embed_dim=64, cond_num_tokens=1, img_num_tokens=12, n_heads=2, n_layers=1)


def _product_net(kind):
    import enhancing_transformers_b200 as etb
    etb.set_precision("parity")
    torch.manual_seed(0)
    if kind == "gpt":
        gpt = etb.GPT(**GCFG)
        with torch.no_grad():
            gpt.pos_emb_code.normal_(0, 0.2)

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.gpt = gpt

            def forward(self, batch):
                codes, conds = batch
                logits = self.gpt(codes, conds)
                return torch.nn.functional.cross_entropy(logits.view(-1, logits.shape[-1]), codes.view(-1))
        return Net()
    e, d, q = PCFG["encoder"], PCFG["decoder"], PCFG["quantizer"]

    class HotPath(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder = etb.ViTEncoder(PCFG["image_size"], PCFG["patch_size"], **e)
            self.decoder = etb.ViTDecoder(PCFG["image_size"], PCFG["patch_size"], **d)
            self.quantizer = etb.VectorQuantizer(**q)
            self.pre_quant = etb.QuantLinear(e["dim"], q["embed_dim"])
            self.post_quant = etb.QuantLinear(q["embed_dim"], d["dim"])
            etb.fuse_post_quant_pos(self)

        def forward(self, x):
            quant, qloss, _ = self.quantizer(self.pre_quant(self.encoder(x)))
            return ((self.decoder(self.post_quant(quant)) - x) ** 2).mean() + qloss
    return HotPath()


def _product_worker(rank, world, port, q, kind):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emulated_ops
    emulated_ops.install(_Patcher)
    from enhancing_transformers_b200.parallel import FlatGradients
    g = torch.Generator().manual_seed(1)
    if kind == "gpt":
        full = (torch.randint(0, 64, (4, 12), generator=g), torch.randint(0, 5, (4, 1), generator=g))
        shard = tuple(t[rank * 2:(rank + 1) * 2] for t in full)
    else:
        full = torch.rand(4, 3, 32, 32, generator=g)
        shard = full[rank * 2:(rank + 1) * 2]
    net = _product_net(kind)
    fg = FlatGradients(net.parameters())
    fg.zero_()
    net(shard).backward()
    fg.allreduce()
    grads = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    if rank == 0:
        single = _product_net(kind)
        single(full).backward()
        worst = 0.0
        for n, p in single.named_parameters():
            if p.grad is None or p.grad.abs().max().item() < 1e-7:      # e.g. attn.key.bias: zero up to rounding
                continue
            worst = max(worst, (grads[n] - p.grad).abs().max().item() / p.grad.abs().max().item())
        q.put(worst)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["vitvq", "gpt"])
def test_product_modules_sharded_gradients_equal_single_process(kind):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_product_worker, args=(r, 2, port, q, kind)) for r in range(2)]
    for p in procs:
        p.start()
    worst = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert worst < 1e-4, worst
