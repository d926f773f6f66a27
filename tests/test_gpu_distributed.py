"""Multi-GPU (NCCL, >= 2 devices) parity of the spatially sharded SHT against the single-GPU transform.
Skipped on single-GPU boxes; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_distributed.py -m gpu`."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, nlat, nlon, L, M, grid, B, q):
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
        C = 8
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
            if B == 1:
                # the same transform with the lat<->m transpose over NVLink peer memory (PeerExchange): True = one block-copy
                # launch per direction, "fused" = the stores / loads are issued by the FFT kernels themselves; each twice, so
                # that buffer re-use across calls goes through the flag barriers
                for engine in (True, "fused"):
                    dp = D.DistributedSHT(nlat, nlon, L, M,
                                          lambda nloc: D.CudaStages(nlat, nloc, nlon, L, M, sht.weights, isht.pct, dev),
                                          peer_exchange=engine)
                    for it in range(2):
                        xs = x * (1.0 + it)
                        pm2 = dp.gather_pm(dp.forward_packed(xs[:, :, dp.lat_lo:dp.lat_hi].contiguous()))
                        e_f = max(e_f, float((pm2 - pm_ref * (1.0 + it)).norm() / pm_ref.norm()))
                        y2 = dp.inverse_packed((cm_ref * (1.0 + it))[:, :, p0:p1].contiguous())
                        e_i = max(e_i, float((y2 - yr * (1.0 + it)).norm() / yr.norm()))
                    assert dp.peer is not None and dp.peer.fused == (engine == "fused")
                    dp.peer.check()
                    dp.peer.close()
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


# B = 1: all_to_all_single straight out of / into the stage buffers + one msfno_lat_segments launch per direction;
# B = 2: the per-peer list exchange (a destination's orders are not one contiguous run of a batched intermediate)
@pytest.mark.parametrize("B", [1, 2])
@pytest.mark.parametrize("nlat,nlon,L,M,grid", [(721, 1440, 120, 121, "equiangular"), (120, 240, 120, 121, "legendre-gauss")])
def test_sharded_sht_matches_single_gpu(nlat, nlon, L, M, grid, B):
    world = torch.cuda.device_count()
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(world, 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nlat, nlon, L, M, grid, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    for _ in range(world):
        rank, e_f, e_i = q.get(timeout=5)
        assert e_f < 1e-5 and e_i < 1e-5, (rank, e_f, e_i)


def test_lat_segments_gather_scatter_single_gpu():
    """msfno_lat_segments (both ends of the lat<->m exchange) against index arithmetic: uneven latitude split, padded
    pitches, zeroed tails.  One GPU."""
    import ctypes
    from msfno_b200._lib import check, lib
    nlat, rows = 181, 37
    bounds = [0, 46, 91, 136, 181]                       # 46 / 45 / 45 / 45 latitudes
    pad = lambda n: (n + 31) // 32 * 32
    g = torch.Generator().manual_seed(2)
    blocks = [torch.randn(rows, pad(bounds[i + 1] - bounds[i]), generator=g) for i in range(4)]
    flat = torch.cat([b.reshape(-1) for b in blocks]).cuda()
    full = torch.full((rows, pad(nlat)), float("nan"), device="cuda")
    n = 4
    lo = (ctypes.c_int * n)(*bounds[:-1])
    cnt = (ctypes.c_int * n)(*[bounds[i + 1] - bounds[i] for i in range(n)])
    st = torch.cuda.current_stream().cuda_stream
    check(lib.msfno_lat_segments(1, flat.data_ptr(), full.data_ptr(), rows, pad(nlat), nlat, n, lo, cnt, st), "lat_segments")
    want = torch.zeros(rows, pad(nlat))
    for i in range(4):
        want[:, bounds[i]:bounds[i + 1]] = blocks[i][:, :bounds[i + 1] - bounds[i]]
    assert torch.equal(full.cpu(), want)
    back = torch.full_like(flat, float("nan"))
    check(lib.msfno_lat_segments(0, back.data_ptr(), full.data_ptr(), rows, pad(nlat), nlat, n, lo, cnt, st), "lat_segments")
    off = 0
    for i in range(4):
        ns, pw = bounds[i + 1] - bounds[i], pad(bounds[i + 1] - bounds[i])
        got = back[off:off + rows * pw].view(rows, pw).cpu()
        assert torch.equal(got[:, :ns], blocks[i][:, :ns]) and float(got[:, ns:].abs().max()) == 0.0
        off += rows * pw
