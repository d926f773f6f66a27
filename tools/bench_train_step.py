"""BASELINE config 3: MSFNO (FiLM-modulated SFNO) training step -- forward + backward + Adam on the FiLM head, synthetic
ERA5-shaped batch, data-parallel (DDP over NCCL) when launched with torchrun.  Prints one JSON line from rank 0.
    python tools/bench_train_step.py [--batch 8] [--film-layers 1|12] [--steps 5] [--precision tf32|fp32]
Reference semantics: all non-film_gen parameters are frozen (sfno/model.py:1021-1023); the encoder and the
un-FiLMed blocks run under no_grad (sfnonet.py:817-827,843-844), so with film_layers=1 the backward pass stops at block 11;
film_layers=12 sends it through every SHT / ISHT / spectral-MLP adjoint kernel."""
import argparse, json, os, sys
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200


class Cfg:
    film_gen_type, cls, embed_dim, mlp_dim, dropout, scale_weight, repeat_film = "mae", "x", 512, 1024, 0.0, 1, False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--film-layers", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precision", default="tf32")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    msfno_b200.set_precision(a.precision)
    torch.manual_seed(0)
    cfg = Cfg()
    cfg.film_layers, cfg.batch_size = a.film_layers, a.batch
    net = msfno_b200.FourierNeuralOperatorNet_Filmed(dev, cfg, advanced_logging=False, film_layers=a.film_layers, model_depth=6).to(dev)
    for n, p in net.named_parameters():
        p.requires_grad_(n.startswith("film_gen"))
    model = net
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], broadcast_buffers=False)
    opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=1e-4)
    g = torch.Generator().manual_seed(rank)
    x = torch.randn(a.batch, 73, 721, 1440, generator=g).to(dev)
    y = torch.randn(a.batch, 73, 721, 1440, generator=g).to(dev)
    cond = torch.randn(a.batch, 512, generator=g).to(dev)
    lossf = torch.nn.MSELoss()

    def step():
        opt.zero_grad(set_to_none=True)
        with msfno_b200.precision.library_scope():
            out = model(x, cond, 1.0)
            loss = lossf(out, y)
            loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        loss = step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        gn = sum(float(p.grad.norm()) for p in net.parameters() if p.grad is not None)
        print(json.dumps({"config": "3: MSFNO fwd+bwd+Adam, film_layers=%d" % a.film_layers, "n_gpus": world,
                          "batch_per_gpu": a.batch, "precision": a.precision, "ms_per_step": float(t),
                          "samples_per_sec": world * a.batch / float(t) * 1e3, "loss": float(loss), "grad_norm_sum": gn,
                          "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
