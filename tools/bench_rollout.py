"""BASELINE config 4: autoregressive rollout x <- net(x) of an ensemble, members sharded across GPUs with no data-path
collective (SURVEY.md section 8(d) "Config 4", 8(e) "Ensemble / batch").  Reference loop: MSFNO/Models/sfno/model.py:327-331
(one forward per 6 h step, output fed back as the next input).
    python tools/bench_rollout.py [--members 8] [--steps 112] [--batch 1|2|4|8] [--precision tf32|fp32]
    torchrun --nproc-per-node N ... tools/bench_rollout.py      (each rank rolls out its own --members members)
The members of a rank are forecast in groups of --batch (one CUDA graph of the batched forward, replayed 112 times per
group).  Device time is taken with CUDA events around every group's rollout; rank 0 prints one JSON line with the
member-steps/s of the whole job (max time over ranks) and the per-member rate."""
import argparse, json, os, sys
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200.graph import GraphedForward


def run(dev, rank, world, max_over_ranks, sync_all, members=8, batch=1, steps=112, precision="tf32"):
    """Rolls out this rank's `members` members for `steps` 6 h steps on an initialised process group (world may be 1);
    returns the result dict on every rank (device time: CUDA events, max over ranks)."""
    assert members % batch == 0, "members must be a multiple of batch"
    msfno_b200.set_precision(precision)
    try:
        torch.manual_seed(0)
        net = msfno_b200.FourierNeuralOperatorNet(dev, None, filter_type="non-linear").to(dev).eval()
        g = torch.Generator().manual_seed(1000 + rank)
        ens = torch.randn(members, 73, 721, 1440, generator=g).to(dev)
        final = torch.empty_like(ens)
        with msfno_b200.precision.library_scope():
            gf = GraphedForward(net, ens[:batch])
            gf.rollout(ens[:batch], 2)  # warm-up replays
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(0, members, batch):
            final[i:i + batch].copy_(gf.rollout(ens[i:i + batch], steps), non_blocking=True)
        e1.record()
        sync_all()
        ms = max_over_ranks(e0.elapsed_time(e1))
        finite = bool(torch.isfinite(final).all())
        total = world * members * steps
        # the same rollout with EVERY step's fields copied to pinned host memory (what the reference's running() loop does,
        # sfno/model.py:345-370), for one batch of members: device-side snapshot + double-buffered D2H under the next steps
        host = [torch.empty(ens[:batch].shape, dtype=ens.dtype).pin_memory() for _ in range(steps)]
        gf.rollout(ens[:batch], 4, host_out=host[:4])
        sync_all()
        e0.record()
        gf.rollout(ens[:batch], steps, host_out=host)
        e1.record()
        sync_all()
        ms_host = max_over_ranks(e0.elapsed_time(e1))
        d2h_bytes = host[0].numel() * 4
        return {"metric": "sfno_rollout_member_steps_per_sec_721x1440x73", "value": total / (ms * 1e-3), "unit": "member-steps/s",
                "n_gpus": world, "members_per_gpu": members, "members_total": world * members, "batch": batch,
                "steps_per_member": steps, "ms_per_member_step": ms / (members * steps),
                "rollout_seconds_per_member": ms * 1e-3 / members, "wall_ms": ms, "scaling": "weak", "dtype": precision,
                "with_every_step_copied_to_host": {"ms_per_member_step": ms_host / (batch * steps), "d2h_bytes_per_step": d2h_bytes,
                                                   "d2h_GBps_per_gpu": d2h_bytes * steps / (ms_host * 1e-3) / 1e9,
                                                   "member_steps_per_sec": world * batch * steps / (ms_host * 1e-3)},
                "data": "synthetic (random-init weights: the iterated map is not a forecast, only its cost is meaningful)",
                "output_finite_this_rank": finite,
                "config": {"workload": "configs[3]: 112-step autoregressive rollout, ensemble members sharded across GPUs, "
                                       "no inter-GPU traffic", "filter_type": "non-linear", "cuda_graph": True}}
    finally:
        msfno_b200.set_precision("fp32")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=8, help="ensemble members per GPU (config 4: 64 members over 8 GPUs)")
    ap.add_argument("--batch", type=int, default=1, help="members forecast together in one batched forward")
    ap.add_argument("--steps", type=int, default=112, help="6 h steps per member (28 days = 112)")
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

    res = run(dev, rank, world, max_over_ranks, sync_all, a.members, a.batch, a.steps, a.precision)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
