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


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nlat, nlon, L, M, B, C = 1441, 2880, 240, 241, 1, 256
    sht = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid="equiangular").float().to(dev)
    isht = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid="equiangular").float().to(dev)
    dsht = D.DistributedSHT(nlat, nlon, L, M, lambda nloc: D.CudaStages(nlat, nloc, nlon, L, M, sht.weights, isht.pct, dev))
    x = torch.randn(B, C, dsht.nlat_loc, nlon, device=dev)
    p0, p1 = dsht.pos_range()
    with torch.no_grad():
        for _ in range(3):
            pm = dsht.forward_packed(x)
            y = dsht.inverse_packed(pm.transpose(1, 2).contiguous())
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 10
        e0.record()
        for _ in range(K):
            pm = dsht.forward_packed(x)
            y = dsht.inverse_packed(pm.transpose(1, 2).contiguous())
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        gb = 2 * (4 * B * C * nlat * nlon + 8 * B * C * L * M + 4 * M * L * nlat) / 1e9
        print(json.dumps({"config": "5A sharded SHT+ISHT 1441x2880 C=256 lmax=240", "n_gpus": world, "ms_per_roundtrip": float(t),
                          "algorithmic_GB": gb, "aggregate_GBps": gb / float(t) * 1e3, "finite": bool(torch.isfinite(y).all())}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
