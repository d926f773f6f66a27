"""world_size-2/3 gloo tests (CPU) of the multi-GPU host logic: member sharding, order balancing and the lat<->m
all-to-all plumbing of the spatially sharded SHT (msfno_b200.distributed).  The per-rank stages are stand-ins
written with torch CPU ops in the kernels' data layouts; the exchange / assembly code under test is the product's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import th_shim

import msfno_b200
from msfno_b200 import distributed as D


def test_shard_members_and_bounds():
    for n, w in ((64, 8), (8, 8), (10, 4), (3, 4), (1, 2)):
        slices = [D.shard_members(n, w, r) for r in range(w)]
        got = [i for s in slices for i in range(s.start, s.stop)]
        assert got == list(range(n))
        sizes = [s.stop - s.start for s in slices]
        assert max(sizes) - min(sizes) <= 1
    assert D.split_even(721, 4) == [0, 181, 361, 541, 721]
    for L, M, w in ((120, 121, 2), (120, 121, 8), (240, 241, 4), (12, 13, 3), (4, 5, 8)):
        b = D.split_orders(L, M, w)
        assert len(b) == w + 1 and b[0] == 0 and b[-1] == min(L, M) and all(x <= y for x, y in zip(b, b[1:]))
        # cost of an order to its owner: exchange / operand traffic (one unit) + 128-row tiles of its Legendre GEMM
        cost = [sum(1 + (L - m + 127) // 128 for m in range(b[r], b[r + 1])) for r in range(w)]
        if min(L, M) >= 4 * w:
            assert max(cost) <= 1.25 * sum(cost) / w + 3
            n_orders = [b[r + 1] - b[r] for r in range(w)]
            assert max(n_orders) <= 2 * min(n_orders)      # no rank receives a multiple of another's share of the spectrum
    poff, plen4, P = D.packed_offsets(120, 121)
    assert P == 7440 and poff[0] == 0 and poff[1] == 120 and plen4[119] == 4


class TorchStages:
    """Stand-in stages in the kernels' layouts (Xt/Yt: [B, m, 2C, lat]; PM: [B, p, 2C]; CM: [B, 2C, p])."""

    def __init__(self, nlat, nlat_loc, nlon, lmax, mmax, weights, pct):
        self.nlat, self.nlat_loc, self.nlon, self.lmax, self.mmax = nlat, nlat_loc, nlon, lmax, mmax
        self.mlim = min(lmax, mmax)
        self.w, self.pct = weights, pct
        self.poff, self.plen4, self.P = D.packed_offsets(lmax, mmax)

    def pad(self, n):
        return n

    def fft_fwd(self, x):
        B, C = x.shape[:2]
        X = 2.0 * torch.pi * torch.fft.rfft(x, dim=-1, norm="forward")[..., :self.mlim]   # [B,C,k,m]
        Xr = torch.view_as_real(X).permute(0, 3, 1, 4, 2)                                 # [B,m,C,2,k]
        return Xr.reshape(B, self.mlim, 2 * C, x.shape[2]).contiguous()

    def fft_inv(self, Yt, B, C):
        Y = Yt.reshape(B, self.mlim, C, 2, -1).permute(0, 2, 4, 1, 3).contiguous()         # [B,C,k,m,2]
        return torch.fft.irfft(torch.view_as_complex(Y), n=self.nlon, dim=-1, norm="forward")

    def legendre_fwd(self, Xt, m_lo, m_hi, Ploc):
        B, _, C2, _ = Xt.shape
        out = torch.zeros(B, Ploc, C2, dtype=Xt.dtype)
        p0 = self.poff[m_lo]
        for m in range(m_lo, m_hi):
            n = self.lmax - m
            out[:, self.poff[m] - p0:self.poff[m] - p0 + n] = torch.einsum("lk,bck->blc", self.w[m, m:], Xt[:, m - m_lo])
        return out

    def legendre_inv(self, cm, m_lo, m_hi):
        B, C2, _ = cm.shape
        out = torch.zeros(B, m_hi - m_lo, C2, self.nlat, dtype=cm.dtype)
        p0 = self.poff[m_lo]
        for m in range(m_lo, m_hi):
            n = self.lmax - m
            out[:, m - m_lo] = torch.einsum("lk,bcl->bck", self.pct[m, m:], cm[:, :, self.poff[m] - p0:self.poff[m] - p0 + n])
        return out


def _worker(rank, world, port, nlat, nlon, L, M, grid, q, B=2):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        o_s = th_shim.RealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid)
        o_i = th_shim.InverseRealSHT(nlat, nlon, lmax=L, mmax=M, grid=grid)
        C = 3
        x = torch.randn(B, C, nlat, nlon, dtype=torch.float64)
        dsht = D.DistributedSHT(nlat, nlon, L, M, lambda nloc: TorchStages(nlat, nloc, nlon, L, M, o_s.weights, o_i.pct))
        pm_loc = dsht.forward_packed(x[:, :, dsht.lat_lo:dsht.lat_hi])
        pm = dsht.gather_pm(pm_loc)                                     # [B, P, 2C]
        want = torch.view_as_real(o_s(x))                               # [B,C,L,M,2]
        poff, _, P = D.packed_offsets(L, M)
        err = 0.0
        for m in range(min(L, M)):
            got_m = pm[:, poff[m]:poff[m] + (L - m)].reshape(B, L - m, C, 2)            # [B, l, C, ri]
            err = max(err, float((got_m.permute(0, 2, 1, 3) - want[:, :, m:, m]).abs().max()))
        # inverse: feed the local positions (CM layout) of a random band-limited spectrum
        cin = torch.randn(B, C, L, M, 2, dtype=torch.float64)
        ii, jj = torch.triu_indices(L, M, offset=1)
        cin[:, :, ii, jj] = 0
        if M > L:
            cin[:, :, :, L:] = 0
        y_want = o_i(torch.view_as_complex(cin))
        p0, p1 = dsht.pos_range()
        cm_loc = torch.zeros(B, 2 * C, p1 - p0, dtype=torch.float64)
        for m in range(dsht.m_lo, dsht.m_hi):
            blk = cin[:, :, m:, m]                                                     # [B, C, l, ri]
            cm_loc[:, :, poff[m] - p0:poff[m] - p0 + (L - m)] = blk.permute(0, 1, 3, 2).reshape(B, 2 * C, L - m)
        y_loc = dsht.inverse_packed(cm_loc)
        err_i = float((y_loc - y_want[:, :, dsht.lat_lo:dsht.lat_hi]).abs().max())
        q.put((rank, err, err_i))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,nlat,nlon,L,M,grid,B", [(2, 24, 48, 12, 13, "equiangular", 2), (3, 25, 48, 10, 13, "equiangular", 2),
                                                        (2, 16, 32, 16, 9, "legendre-gauss", 2),
                                                        (3, 25, 48, 10, 13, "equiangular", 1)])   # B = 1: receive straight into views
def test_spatially_sharded_sht_gloo(world, nlat, nlon, L, M, grid, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nlat, nlon, L, M, grid, q, B)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    res = [q.get(timeout=5) for _ in range(world)]
    for rank, err, err_i in res:
        assert err < 1e-10 and err_i < 1e-10, (rank, err, err_i)
