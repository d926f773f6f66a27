"""Multi-GPU (NCCL, >= 2 devices) parity of the spatially sharded SHT against the single-GPU transform.
Skipped on single-GPU boxes; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_distributed.py -m gpu`."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, nlat, nlon, L, M, grid, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import msfno_b200
        from msfno_b200 import _lib
        from msfno_b200 import distributed as D
        from msfno_b200.sht import relayout
        sht = msfno_b200.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().to(dev)
        isht = msfno_b200.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid).float().to(dev)
        g = torch.Generator().manual_seed(0)
        B, C = 1, 8
        x = torch.randn(B, C, nlat, nlon, generator=g).to(dev)
        with torch.no_grad():
            pm_ref = sht.forward_packed(x)
            dsht = D.DistributedSHT(nlat, nlon, L, M,
                                    lambda nloc: D.CudaStages(nlat, nloc, nlon, L, M, sht.weights, isht.pct, dev))
            pm_loc = dsht.forward_packed(x[:, :, dsht.lat_lo:dsht.lat_hi].contiguous())
            pm = dsht.gather_pm(pm_loc)
            e_f = float((pm - pm_ref).norm() / pm_ref.norm())
            cm_ref = relayout(pm_ref, sht, _lib.LAYOUT_PM, _lib.LAYOUT_CM, B, C)
            y_ref = isht.inverse_packed(cm_ref)
            p0, p1 = dsht.pos_range()
            y_loc = dsht.inverse_packed(cm_ref[:, :, p0:p1].contiguous())
            yr = y_ref[:, :, dsht.lat_lo:dsht.lat_hi]
            e_i = float((y_loc - yr).norm() / yr.norm())
        torch.cuda.synchronize()
        q.put((rank, e_f, e_i))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("nlat,nlon,L,M,grid", [(721, 1440, 120, 121, "equiangular"), (120, 240, 120, 121, "legendre-gauss")])
def test_sharded_sht_matches_single_gpu(nlat, nlon, L, M, grid):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nlat, nlon, L, M, grid, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    for _ in range(world):
        rank, e_f, e_i = q.get(timeout=5)
        assert e_f < 1e-5 and e_i < 1e-5, (rank, e_f, e_i)
