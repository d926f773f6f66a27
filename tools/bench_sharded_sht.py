"""BASELINE config 5A: spatially sharded SHT + ISHT at 0.125 deg (1441 x 2880, 256 channels, lmax 240 / mmax 241) on
N GPUs (torchrun), lat<->m all-to-all over NCCL/NVLink.  Prints one JSON line from rank 0.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded_sht.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msfno_b200
from msfno_b200 import distributed as D


def run(dev, rank, world, max_over_ranks, sync_all, steps=10, nlat=1441, nlon=2880, L=240, M=241, B=1, C=256):
    """SHT + ISHT round trip on an initialised process group (world may be 1); returns the result dict (all ranks)."""
    sht = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid="equiangular").float().to(dev)
    isht = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid="equiangular").float().to(dev)
    dsht = D.DistributedSHT(nlat, nlon, L, M, lambda nloc: D.CudaStages(nlat, nloc, nlon, L, M, sht.weights, isht.pct, dev))
    x = torch.randn(B, C, dsht.nlat_loc, nlon, device=dev)
    with torch.no_grad():
        for _ in range(3):
            pm = dsht.forward_packed(x)
            y = dsht.inverse_packed(pm.transpose(1, 2).contiguous())
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            pm = dsht.forward_packed(x)
            y = dsht.inverse_packed(pm.transpose(1, 2).contiguous())
        e1.record()
        sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1) / steps)
    gb = 2 * (4 * B * C * nlat * nlon + 8 * B * C * L * M + 4 * M * L * nlat) / 1e9
    # all-to-all payload of one direction: the truncated spectrum [B][mlim][2C][nlat] fp32; (world - 1) / world of it
    # crosses NVLink
    payload = 4.0 * B * min(L, M) * 2 * C * nlat
    wire = payload * (world - 1) / world
    return {"config": "configs[4] (A): sharded SHT + ISHT round trip, %d x %d, C = %d, lmax = %d" % (nlat, nlon, C, L), "n_gpus": world,
            "ms_per_roundtrip": ms, "algorithmic_GB": gb, "aggregate_GBps": gb / ms * 1e3,
            "all_to_all_payload_MB_per_direction": payload / 1e6, "nvlink_MB_per_direction": wire / 1e6,
            "finite": bool(torch.isfinite(y).all())}


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    def sync_all():
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    res = run(dev, rank, world, max_over_ranks, sync_all)
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
