"""world_size-2 gloo test (CPU) of the caller path (msfno_b200.trainer): freezing, DDP over a tiny
stand-in network with the FiLMed net's call signature, gradient accumulation with no_sync, the two validation
collectives.  Reference semantics: /root/reference MSFNO/Models/train.py:201-298,318-339,533-654; sfno/model.py:1011-1023."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

import msfno_b200
from msfno_b200.trainer import Trainer, TrainerConfig, freeze_backbone


class TinyFilmed(nn.Module):
    """net(x, cond, scale) with a frozen 'backbone' and a trainable film_gen head, like FourierNeuralOperatorNet_Filmed."""

    def __init__(self, C=3):
        super().__init__()
        self.backbone = nn.Conv2d(C, C, 1)
        self.film_gen = nn.Linear(4, 2 * C)
        self.C = C

    def forward(self, x, cond, scale=1.0):
        gb = self.film_gen(cond)
        g, b = gb[:, :self.C, None, None], gb[:, self.C:, None, None]
        return (1 + g * scale) * self.backbone(x) + b * scale


def _data(rank, n, steps, B=2, C=3, H=4, W=8):
    g = torch.Generator().manual_seed(100 + rank)
    return [[(torch.randn(B, C, H, W, generator=g), torch.randn(B, 4, generator=g)) for _ in range(steps + 2)] for _ in range(n)]


def test_freeze_backbone_names():
    net = TinyFilmed()
    params = freeze_backbone(net)
    assert {n for n, p in net.named_parameters() if p.requires_grad} == {"film_gen.weight", "film_gen.bias"}
    assert len(params) == 2
    freeze_backbone(net, retrain_film=True, grad_layers=("backbone",))
    assert {n for n, p in net.named_parameters() if p.requires_grad} == {"backbone.weight", "backbone.bias"}


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = TinyFilmed()
        cfg = TrainerConfig(ddp=True, rank=rank, world_size=world, accumulation_steps=1, multi_step_training=1,
                            discount_factor=0.9, multi_step_validation=1, learning_rate=1e-2)
        tr = Trainer(net, cfg, device=torch.device("cpu"))
        tr.ready_model()
        w0 = net.backbone.weight.detach().clone()
        losses = tr.train_epoch(_data(rank, 4, 1))
        assert tr.iter == 2 and losses.numel() == 2            # 4 micro-batches, update every 2
        assert torch.equal(net.backbone.weight, w0)            # frozen
        val, pervar = tr.validation(_data(rank, 2, 1))
        q.put((rank, net.film_gen.weight.detach().tolist(), val.tolist(), pervar.tolist()))   # plain lists: no shared-memory handles
    finally:
        dist.destroy_process_group()


def test_trainer_ddp_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w_a, val_a, pv_a), (_, w_b, val_b, pv_b) = [(r, torch.tensor(w), torch.tensor(v), torch.tensor(pv)) for r, w, v, pv in res]
    assert torch.allclose(w_a, w_b)                  # DDP kept the replicas in step (gradients were averaged)
    assert torch.allclose(val_a, val_b) and torch.allclose(pv_a, pv_b)   # the two all-reduces
    assert val_a.shape == (2,) and pv_a.shape == (2, 3)

    # single-process reference of the same two updates: gradients averaged over the two ranks' micro-batches
    torch.manual_seed(0)
    net = TinyFilmed()
    freeze_backbone(net)
    opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=1e-2)
    data = [_data(r, 4, 1) for r in range(2)]
    lossf = nn.MSELoss()
    for it in range(2):
        opt.zero_grad()
        for r in range(2):
            for mb in (2 * it, 2 * it + 1):
                d = data[r][mb]
                out0 = net(d[0][0], d[0][1], 1.0)
                out1 = net(out0, d[1][1], 1.0)
                loss = (lossf(out0, d[1][0]) / 2 / 2 + lossf(out1, d[2][0]) / 2 / 2 * 0.9) / 2   # / world: DDP averages
                loss.backward()
        opt.step()
    assert torch.allclose(net.film_gen.weight, w_a, atol=1e-6), float((net.film_gen.weight - w_a).abs().max())
