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


def run(dev, rank, world, max_over_ranks, sync_all, batch=8, film_layers=1, steps=5, warmup=2, precision="tf32"):
    """One data-parallel MSFNO training step through msfno_b200.Trainer (the reference's loop structure, train.py:201-298):
    frozen backbone, DistributedDataParallel over the FiLM head when world > 1.  Returns the result dict (all ranks)."""
    msfno_b200.set_precision(precision)
    try:
        torch.manual_seed(0)
        cfg = Cfg()
        cfg.film_layers, cfg.batch_size = film_layers, batch
        net = msfno_b200.FourierNeuralOperatorNet_Filmed(dev, cfg, advanced_logging=False, film_layers=film_layers, model_depth=6).to(dev)
        tcfg = msfno_b200.TrainerConfig(ddp=world > 1, rank=rank, world_size=world, model_version="film")
        trainer = msfno_b200.Trainer(net, tcfg, loss_fn=torch.nn.MSELoss(), device=dev)
        trainer.ready_model()
        g = torch.Generator().manual_seed(rank)
        x = torch.randn(batch, 73, 721, 1440, generator=g).to(dev)
        y = torch.randn(batch, 73, 721, 1440, generator=g).to(dev)
        cond = torch.randn(batch, 512, generator=g).to(dev)
        data = [(x, cond), (y, cond)]

        def loader(n):
            for _ in range(n):
                yield data

        with msfno_b200.precision.library_scope():
            trainer.train_epoch(loader(warmup))
            sync_all()
            torch.cuda.reset_peak_memory_stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            losses = trainer.train_epoch(loader(steps))
            e1.record()
            sync_all()
        ms = max_over_ranks(e0.elapsed_time(e1) / steps)
        return {"config": "configs[2]: MSFNO fwd+bwd+Adam on the FiLM head, film_layers=%d, DDP" % film_layers, "n_gpus": world,
                "batch_per_gpu": batch, "precision": precision, "ms_per_step": ms, "samples_per_sec": world * batch / ms * 1e3,
                "loss": float(losses[-1]), "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9,
                "trainable_params": sum(p.numel() for p in net.parameters() if p.requires_grad)}
    finally:
        msfno_b200.set_precision("fp32")


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

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    res = run(dev, rank, world, max_over_ranks, sync_all, a.batch, a.film_layers, a.steps, a.warmup, a.precision)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
