"""Multi-GPU forms of the hot path.

1. Member sharding (BASELINE configs 3 and 4): ensemble / batch members are independent units -- `shard_members`
   assigns each rank a contiguous slice, no data-path collective exists.
2. Spatially sharded SHT (BASELINE config 5): latitude rows are split across ranks for the longitude FFT, azimuthal
   orders m are split across ranks for the Legendre contraction, and ONE all-to-all (lat <-> m transpose of the
   truncated spectrum) sits between the two stages, in each direction.  The reference has no such transform (it only
   uses DDP, /root/reference main.py:39-49, MSFNO/Models/train.py:370-374); this is the new capability config 5 asks for.

      forward : x[lat_r] --FFT--> Xt[lat_r][all m] ==all-to-all==> Xt[all lat][m_r] --Legendre--> coef[m_r]
      inverse : coef[m_r] --Legendre--> Yt[all lat][m_r] ==all-to-all==> Yt[lat_r][all m] --FFT--> y[lat_r]

   Orders are balanced by Legendre work (order m costs lmax - m degrees).  The coefficient layouts are m-major, so a
   rank's orders are one contiguous range of packed positions: the spectral ops need no further communication
   (SpectralAttentionS2 is position-wise; SpectralConvS2 weights shard by mode along with m).

The exchange code is backend-agnostic (`all_to_all` on NCCL, batched isend/irecv on gloo); the per-rank stages are
pluggable so tests/test_distributed_cpu.py can drive the identical plumbing under gloo on CPU with stand-in stages.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib


def shard_members(n_members, world_size, rank):
    """Contiguous, balanced slice of the member (batch / ensemble) dimension owned by `rank`."""
    base, rem = divmod(n_members, world_size)
    lo = rank * base + min(rank, rem)
    return slice(lo, lo + base + (1 if rank < rem else 0))


def split_even(n, parts):
    """Boundaries [b_0=0, ..., b_parts=n] of a contiguous split with sizes differing by at most one."""
    base, rem = divmod(n, parts)
    out = [0]
    for r in range(parts):
        out.append(out[-1] + base + (1 if r < rem else 0))
    return out


def packed_offsets(lmax, mmax):
    """poff[m] (start of order m in the packed position layout), plen4[m] and P -- must match msfno_plan_create."""
    mlim = min(lmax, mmax)
    poff, plen4, P = [], [], 0
    for m in range(mlim):
        poff.append(P)
        n = (lmax - m + 3) // 4 * 4
        plen4.append(n)
        P += n
    return poff, plen4, P


def split_orders(lmax, mmax, parts):
    """Boundaries of a contiguous split of the orders [0, mlim) balanced by Legendre work sum(lmax - m)."""
    mlim = min(lmax, mmax)
    cost = [lmax - m for m in range(mlim)]
    total = sum(cost)
    bounds, acc, nxt = [0], 0, 1
    for m in range(mlim):
        acc += cost[m]
        while nxt < parts and acc >= total * nxt / parts and len(bounds) < parts:
            bounds.append(m + 1)
            nxt += 1
    while len(bounds) < parts:
        bounds.append(mlim)
    bounds.append(mlim)
    return bounds


def exchange(outs, ins, group=None):
    """outs[s] <- what rank s put in its ins[my_rank]  (one all-to-all)."""
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_to_all(outs, ins, group=group)
        return
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    ops = []
    for s in range(world):
        if s == rank:
            outs[s].copy_(ins[s])
            continue
        peer = dist.get_global_rank(group, s) if group is not None else s
        ops.append(dist.P2POp(dist.isend, ins[s], peer, group=group))
        ops.append(dist.P2POp(dist.irecv, outs[s], peer, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class CudaStages:
    """The per-rank stages on the sm_100a kernels: longitude FFT on the local latitude rows (msfno_fft_stage),
    Legendre contraction on the local orders (msfno_legendre_stage)."""

    def __init__(self, nlat, nlat_loc, nlon, lmax, mmax, weights, pct, device):
        from .sht import _Plan
        self.device = device
        self.nlat, self.nlat_loc, self.lmax, self.mmax = nlat, nlat_loc, lmax, mmax
        self.fft_plan = _Plan(nlat_loc, nlon, lmax, mmax, device)
        self.leg_a = _Plan(nlat, nlon, lmax, mmax, device)
        self.leg_s = _Plan(nlat, nlon, lmax, mmax, device)
        self.weights = weights.float().contiguous().to(device) if weights is not None else None
        self.pct = pct.float().contiguous().to(device) if pct is not None else None
        if self.weights is not None:
            self.leg_a.set_table(self.weights, True)
        if self.pct is not None:
            self.leg_s.set_table(self.pct, False)
        self.mlim = self.fft_plan.mlim
        self.nlon = nlon

    @staticmethod
    def _st():
        return torch.cuda.current_stream().cuda_stream

    def pad(self, n):
        return (n + 31) // 32 * 32

    def fft_fwd(self, x):
        B, C = x.shape[0], x.shape[1]
        out = torch.empty((B, self.mlim, 2 * C, self.pad(self.nlat_loc)), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib.msfno_fft_stage(self.fft_plan.h, 0, 0, x.data_ptr(), out.data_ptr(), B, C, self._st()), "fft_stage")
        return out

    def fft_inv(self, Yt, B, C):
        y = torch.empty((B, C, self.nlat_loc, self.nlon), dtype=torch.float32, device=Yt.device)
        _lib.check(_lib.lib.msfno_fft_stage(self.fft_plan.h, 1, 0, Yt.data_ptr(), y.data_ptr(), B, C, self._st()), "fft_stage")
        return y

    def legendre_fwd(self, Xt, m_lo, m_hi, Ploc):
        B, C = Xt.shape[0], Xt.shape[2] // 2
        pm = torch.empty((B, Ploc, 2 * C), dtype=torch.float32, device=Xt.device)
        _lib.check(_lib.lib.msfno_legendre_stage(self.leg_a.h, 0, Xt.data_ptr(), pm.data_ptr(), m_lo, m_hi, B, C, self._st()),
                   "legendre_stage")
        return pm

    def legendre_inv(self, cm, m_lo, m_hi):
        B, C = cm.shape[0], cm.shape[1] // 2
        Yt = torch.empty((B, m_hi - m_lo, 2 * C, self.pad(self.nlat)), dtype=torch.float32, device=cm.device)
        _lib.check(_lib.lib.msfno_legendre_stage(self.leg_s.h, 2, cm.data_ptr(), Yt.data_ptr(), m_lo, m_hi, B, C, self._st()),
                   "legendre_stage")
        return Yt


class DistributedSHT:
    """Forward / inverse SHT with latitude sharded for the FFT and orders sharded for the Legendre stage.

    forward_packed(x_loc [B,C,nlat_loc,nlon])  -> coefficients of the local orders, PM layout [B, Ploc, 2C]
    inverse_packed(cm_loc [B,2C,Ploc])         -> y_loc [B,C,nlat_loc,nlon]
    """

    def __init__(self, nlat, nlon, lmax, mmax, stages_factory, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.nlat, self.nlon, self.lmax, self.mmax = nlat, nlon, lmax, mmax
        self.mlim = min(lmax, mmax)
        self.lat_bounds = split_even(nlat, self.world)
        self.m_bounds = split_orders(lmax, mmax, self.world)
        self.poff, self.plen4, self.P = packed_offsets(lmax, mmax)
        self.lat_lo, self.lat_hi = self.lat_bounds[self.rank], self.lat_bounds[self.rank + 1]
        self.m_lo, self.m_hi = self.m_bounds[self.rank], self.m_bounds[self.rank + 1]
        self.nlat_loc = self.lat_hi - self.lat_lo
        self.stages = stages_factory(self.nlat_loc)

    def pos_range(self, r=None):
        r = self.rank if r is None else r
        lo, hi = self.m_bounds[r], self.m_bounds[r + 1]
        p0 = self.poff[lo] if lo < self.mlim else self.P
        p1 = self.poff[hi] if hi < self.mlim else self.P
        return p0, p1

    def forward_packed(self, x_loc):
        st = self.stages
        B, C = x_loc.shape[0], x_loc.shape[1]
        Xt_loc = st.fft_fwd(x_loc.contiguous())                       # [B, mlim, 2C, pad(nlat_loc)]
        mloc = self.m_hi - self.m_lo
        if self.world == 1:
            Xt = Xt_loc                                               # the stage's own padded layout: nothing to move
        else:
            # send side: a destination's orders are one contiguous run of the m-major intermediate (for B = 1 the slices
            # below are views -- no copy), sent with their padded latitude pitch; receive side: one strided scatter per
            # source into the full-latitude operand of the Legendre stage
            ins, outs = [], []
            for s in range(self.world):
                lo, hi = self.m_bounds[s], self.m_bounds[s + 1]
                ins.append(Xt_loc[:, lo:hi].contiguous())
                pw = st.pad(self.lat_bounds[s + 1] - self.lat_bounds[s])     # the SENDER's padded pitch
                outs.append(torch.empty((B, mloc, 2 * C, pw), dtype=x_loc.dtype, device=x_loc.device))
            exchange(outs, ins, self.group)
            Xt = torch.empty((B, mloc, 2 * C, st.pad(self.nlat)), dtype=x_loc.dtype, device=x_loc.device)
            Xt[..., self.nlat:] = 0
            for s in range(self.world):
                n_s = self.lat_bounds[s + 1] - self.lat_bounds[s]
                Xt[..., self.lat_bounds[s]:self.lat_bounds[s + 1]] = outs[s][..., :n_s]
        p0, p1 = self.pos_range()
        if mloc == 0:
            return torch.zeros((B, 0, 2 * C), dtype=x_loc.dtype, device=x_loc.device)
        return st.legendre_fwd(Xt, self.m_lo, self.m_hi, p1 - p0)

    def inverse_packed(self, cm_loc):
        st = self.stages
        B, C = cm_loc.shape[0], cm_loc.shape[1] // 2
        mloc = self.m_hi - self.m_lo
        if mloc > 0:
            Yt = st.legendre_inv(cm_loc.contiguous(), self.m_lo, self.m_hi)  # [B, mloc, 2C, pad(nlat)]
        else:
            Yt = torch.zeros((B, 0, 2 * C, st.pad(self.nlat)), dtype=cm_loc.dtype, device=cm_loc.device)
        if self.world == 1:
            return st.fft_inv(Yt, B, C)
        # send side: one strided gather per destination, written with the destination's padded latitude pitch; receive
        # side: a source's orders are one contiguous run of the local m-major operand -- for B = 1 the receive buffers ARE
        # views of it (no assembly copy)
        padl = st.pad(self.nlat_loc)
        Yt_loc = torch.empty((B, self.mlim, 2 * C, padl), dtype=cm_loc.dtype, device=cm_loc.device)
        direct = B == 1
        ins, outs = [], []
        for s in range(self.world):
            n_s = self.lat_bounds[s + 1] - self.lat_bounds[s]
            pw = st.pad(n_s)
            blk = torch.empty((B, mloc, 2 * C, pw), dtype=cm_loc.dtype, device=cm_loc.device)
            blk[..., :n_s] = Yt[..., self.lat_bounds[s]:self.lat_bounds[s + 1]]
            if pw > n_s:
                blk[..., n_s:] = 0
            ins.append(blk)
            view = Yt_loc[:, self.m_bounds[s]:self.m_bounds[s + 1]]
            outs.append(view if direct else torch.empty_like(view))
        exchange(outs, ins, self.group)
        if not direct:
            for s in range(self.world):
                Yt_loc[:, self.m_bounds[s]:self.m_bounds[s + 1]] = outs[s]
        return st.fft_inv(Yt_loc, B, C)

    # -- helpers for tests / the public boundary ---------------------------------------------------
    def gather_pm(self, pm_loc):
        """All ranks' PM coefficients concatenated along the position axis -> [B, P, 2C] on every rank."""
        B, C2 = pm_loc.shape[0], pm_loc.shape[2]
        parts = []
        for r in range(self.world):
            p0, p1 = self.pos_range(r)
            parts.append(torch.empty((B, p1 - p0, C2), dtype=pm_loc.dtype, device=pm_loc.device))
        # all_gather needs equal shapes: pad to the maximum
        pmax = max(p.shape[1] for p in parts)
        buf = torch.zeros((B, pmax, C2), dtype=pm_loc.dtype, device=pm_loc.device)
        buf[:, :pm_loc.shape[1]] = pm_loc
        gathered = [torch.empty_like(buf) for _ in range(self.world)]
        dist.all_gather(gathered, buf, group=self.group)
        return torch.cat([g[:, :p.shape[1]] for g, p in zip(gathered, parts)], dim=1)
