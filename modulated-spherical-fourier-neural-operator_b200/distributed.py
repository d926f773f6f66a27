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


def split_even(n, parts, align=1):
    """Boundaries [b_0=0, ..., b_parts=n] of a contiguous split with sizes differing by at most one -- in units of `align`
    (inner boundaries are multiples of it; the last part takes the remainder)."""
    if n < align * parts:        # too few rows for aligned shares: nobody may end up empty
        align = 1
    units = (n + align - 1) // align
    base, rem = divmod(units, parts)
    out = [0]
    for r in range(parts):
        out.append(min(n, out[-1] + align * (base + (1 if r < rem else 0))))
    out[-1] = n
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
    """Boundaries of a contiguous split of the orders [0, mlim) balanced by what an order really costs its owner: one unit
    for its share of the lat<->m exchange and of the operand traffic (every order is 2C rows of nlat latitudes, whatever
    its degree range) plus one unit per 128-row tile of its Legendre GEMM (lmax - m rows).  Balancing by flops
    sum(lmax - m) alone gave the last of 8 ranks 85 of 240 orders: 35 % of the spectrum arrived through ONE rank's NVLink
    port and its GEMMs ran 85 tiles against 30 on rank 0 (profiles/r02_sharded_sht_n8_peer.json: exchange phases
    0.47 / 0.62 ms of a 2.1 ms round trip)."""
    mlim = min(lmax, mmax)
    cost = [1 + (lmax - m + 127) // 128 for m in range(mlim)]
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

    def lat_segments(self, gather, flat, full, lat_bounds):
        """One launch for either end of the lat<->m exchange (msfno_lat_segments): `flat` = the peers' blocks back to back,
        each [rows][pad(n_s)]; `full` = [..., pad(nlat)] with every latitude."""
        n = len(lat_bounds) - 1
        lo = (ctypes.c_int * n)(*lat_bounds[:-1])
        cnt = (ctypes.c_int * n)(*[lat_bounds[i + 1] - lat_bounds[i] for i in range(n)])
        rows = full.numel() // full.shape[-1]
        _lib.check(_lib.lib.msfno_lat_segments(1 if gather else 0, flat.data_ptr(), full.data_ptr(), rows, full.shape[-1], self.nlat, n,
                                               lo, cnt, self._st()), "lat_segments")

    def legendre_inv(self, cm, m_lo, m_hi, out=None):
        B, C = cm.shape[0], cm.shape[1] // 2
        Yt = out if out is not None else torch.empty((B, m_hi - m_lo, 2 * C, self.pad(self.nlat)), dtype=torch.float32,
                                                     device=cm.device)
        _lib.check(_lib.lib.msfno_legendre_stage(self.leg_s.h, 2, cm.data_ptr(), Yt.data_ptr(), m_lo, m_hi, B, C, self._st()),
                   "legendre_stage")
        return Yt


class _DevArray:
    """__cuda_array_interface__ view of raw device memory (own cudaMalloc or a peer mapping) for torch.as_tensor."""

    def __init__(self, ptr, n, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class _PeerMap(ctypes.Structure):        # mirrors msfno_peer_map (include/msfno_b200.h)
    _fields_ = [("world", ctypes.c_int), ("m_bounds", ctypes.c_int * 17), ("buf", ctypes.c_void_p * 16), ("pitch", ctypes.c_int),
                ("lat_lo", ctypes.c_int)]


class _PeerBlock(ctypes.Structure):      # mirrors msfno_peer_block (include/msfno_b200.h)
    _fields_ = [("dst", ctypes.c_void_p), ("rows", ctypes.c_longlong), ("cols", ctypes.c_int), ("zero_tail", ctypes.c_int),
                ("src_row0", ctypes.c_longlong), ("src_col0", ctypes.c_longlong), ("src_pitch", ctypes.c_longlong),
                ("dst_row0", ctypes.c_longlong), ("dst_col0", ctypes.c_longlong), ("dst_pitch", ctypes.c_longlong)]


class PeerExchange:
    """The lat<->m transpose as direct NVLink stores into the consumers' operand buffers (csrc/peer.cu): every rank owns
         A [mloc][2C][pad(nlat)]     operand of ITS Legendre analysis -- filled by all ranks' longitude stages
         Bf [mlim][2C][pad(nlat_loc)] operand of ITS inverse longitude stage -- filled by all ranks' Legendre syntheses
    both cudaMalloc'ed, exported through CUDA IPC and mapped by every peer (one process per GPU, one node).  A direction
    is barrier -> ONE block-copy launch -> barrier on the caller's stream; no library collective, no staging copy.
    fused=True goes one step further (msfno_fft_stage_peer): the forward FFT kernel's epilogue stores its row segments
    straight into the owners' A, and the inverse FFT kernel's staging fill loads them straight from the owners' Legendre
    synthesis output (which then lives in A as well) -- the exchange has no kernel of its own."""

    def __init__(self, dsht, C, device, pad, fused=False):
        self.d, self.C, self.device, self.group = dsht, C, device, dsht.group
        self.fused = bool(fused)
        self.world, self.rank = dsht.world, dsht.rank
        if self.world > 16:
            raise ValueError("PeerExchange supports up to 16 ranks (MSFNO_MAX_PEERS)")
        lib, check = _lib.lib, _lib.check
        d = dsht
        self.padN = pad(d.nlat)
        self.pads = [pad(d.lat_bounds[s + 1] - d.lat_bounds[s]) for s in range(self.world)]
        mloc = d.m_hi - d.m_lo
        self.nA = max(mloc, 1) * 2 * C * self.padN
        self.nB = d.mlim * 2 * C * self.pads[self.rank]
        own, handles = [], []
        with torch.cuda.device(device):
            for nbytes in (4 * self.nA, 4 * self.nB, 4 * 64):
                p, h = ctypes.c_void_p(), ctypes.create_string_buffer(64)
                check(lib.msfno_peer_alloc(nbytes, ctypes.byref(p), h), "peer_alloc")
                own.append(p.value)
                handles.append(h.raw)
        self.own = own
        gathered = [None] * self.world
        dist.all_gather_object(gathered, handles, group=self.group)
        self.maps = []          # maps[r] = [A, Bf, flags] pointers of rank r as seen from this process
        with torch.cuda.device(device):
            for r in range(self.world):
                if r == self.rank:
                    self.maps.append(list(own))
                    continue
                ptrs = []
                for raw in gathered[r]:
                    q = ctypes.c_void_p()
                    check(lib.msfno_peer_open(ctypes.create_string_buffer(raw, 64), ctypes.byref(q)), "peer_open")
                    ptrs.append(q.value)
                self.maps.append(ptrs)
        self.A = torch.as_tensor(_DevArray(own[0], self.nA), device=device)
        self.Bf = torch.as_tensor(_DevArray(own[1], self.nB), device=device)
        self.flag_ptrs = (ctypes.c_void_p * self.world)(*[m[2] for m in self.maps])
        self.state = torch.zeros(2, dtype=torch.int32, device=device)     # [a barrier timed out, barriers so far]
        # the block lists are static: built once, so a direction costs three launches and no Python object churn
        d, C2 = dsht, 2 * C
        n_me, padl, mloc = d.lat_hi - d.lat_lo, self.pads[self.rank], d.m_hi - d.m_lo
        fwd = [_PeerBlock(self.maps[s][0], (d.m_bounds[s + 1] - d.m_bounds[s]) * C2, n_me, 0, d.m_bounds[s] * C2, 0, padl,
                          0, d.lat_lo, self.padN) for s in range(self.world)]
        inv = [_PeerBlock(self.maps[s][1], mloc * C2, d.lat_bounds[s + 1] - d.lat_bounds[s],
                          self.pads[s] - (d.lat_bounds[s + 1] - d.lat_bounds[s]), 0, d.lat_bounds[s], self.padN, d.m_lo * C2, 0,
                          self.pads[s]) for s in range(self.world)]
        self.blocks_fwd = (_PeerBlock * self.world)(*fwd)
        self.blocks_inv = (_PeerBlock * self.world)(*inv)
        self.A_view = self.A.view(1, max(mloc, 1), C2, self.padN)[:, :mloc]
        self.B_view = self.Bf.view(1, d.mlim, C2, padl)
        self.mloc = mloc
        pm = _PeerMap()
        pm.world, pm.pitch, pm.lat_lo = self.world, self.padN, d.lat_lo
        for s_ in range(self.world + 1):
            pm.m_bounds[s_] = d.m_bounds[s_]
        for s_ in range(self.world):
            pm.buf[s_] = self.maps[s_][0]
        self.peer_map = pm
        dist.barrier(group=self.group)      # every mapping exists before anyone stores through one

    # ---- fused engine: the exchange inside the FFT kernels -------------------------------------------------------------
    def fft_fwd_into_peers(self, stages, x_loc):
        """x_loc [1][C][nlat_loc][nlon] -> view of A (all latitudes of this rank's orders), filled by every rank's FFT kernel."""
        st = torch.cuda.current_stream().cuda_stream
        self._barrier(st)                    # every peer has finished with its A (analysis / inverse-FFT loads of the last call)
        _lib.check(_lib.lib.msfno_fft_stage_peer(stages.fft_plan.h, 0, x_loc.data_ptr(), None, ctypes.byref(self.peer_map), self.C, st),
                   "fft_stage_peer")
        self._barrier(st)                    # every peer's stores into MY A are visible
        return self.A_view

    def synthesis_target(self):
        """Where this rank's Legendre synthesis must write (A: the peers' inverse FFTs load from it)."""
        self._barrier(torch.cuda.current_stream().cuda_stream)     # nobody still loads the previous contents
        return self.A_view

    def fft_inv_from_peers(self, stages, C):
        st = torch.cuda.current_stream().cuda_stream
        self._barrier(st)                    # every rank's synthesis output is complete and visible
        d = self.d
        y = torch.empty((1, C, d.nlat_loc, d.nlon), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib.msfno_fft_stage_peer(stages.fft_plan.h, 1, None, y.data_ptr(), ctypes.byref(self.peer_map), C, st),
                   "fft_stage_peer")
        return y

    def _barrier(self, st):
        _lib.check(_lib.lib.msfno_peer_barrier(self.flag_ptrs, self.rank, self.world, self.state.data_ptr(), st), "peer_barrier")

    def forward(self, Xt_loc):
        """Xt_loc [1][mlim][2C][pad(nlat_loc)] (this rank's latitudes, every order) -> view of A [1][mloc][2C][pad(nlat)]."""
        st = torch.cuda.current_stream().cuda_stream
        self._barrier(st)                    # every peer has finished reading its A of the previous call
        _lib.check(_lib.lib.msfno_peer_block_copy(Xt_loc.data_ptr(), self.world, self.blocks_fwd, st), "peer_block_copy")
        self._barrier(st)                    # every peer's stores into MY A are visible
        return self.A_view

    def inverse(self, Yt):
        """Yt [1][mloc][2C][pad(nlat)] (this rank's orders, every latitude) -> view of Bf [1][mlim][2C][pad(nlat_loc)]."""
        st = torch.cuda.current_stream().cuda_stream
        self._barrier(st)
        if self.mloc > 0:
            _lib.check(_lib.lib.msfno_peer_block_copy(Yt.data_ptr(), self.world, self.blocks_inv, st), "peer_block_copy")
        self._barrier(st)
        return self.B_view

    def check(self):
        """Synchronises; raises if a barrier gave up waiting for a peer."""
        if int(self.state[0].item()) != 0:
            raise RuntimeError("PeerExchange: a rank never arrived at a peer barrier")

    def close(self):
        lib = _lib.lib
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)       # nobody unmaps while a peer may still store
        for r, ptrs in enumerate(self.maps):
            if r != self.rank:
                for p in ptrs:
                    lib.msfno_peer_close(p)
        self.A = self.Bf = self.A_view = self.B_view = None
        for p in self.own:
            lib.msfno_peer_free(p)
        self.maps, self.own = [], []


class DistributedSHT:
    """Forward / inverse SHT with latitude sharded for the FFT and orders sharded for the Legendre stage.

    forward_packed(x_loc [B,C,nlat_loc,nlon])  -> coefficients of the local orders, PM layout [B, Ploc, 2C]
    inverse_packed(cm_loc [B,2C,Ploc])         -> y_loc [B,C,nlat_loc,nlon]
    """

    def __init__(self, nlat, nlon, lmax, mmax, stages_factory, group=None, peer_exchange=False):
        """peer_exchange=True (NCCL process group on one node, B = 1): the lat<->m transpose runs as direct NVLink stores
        into the peers' operand buffers (PeerExchange) instead of all_to_all_single + pack / unpack launches;
        peer_exchange="fused": those stores / loads are issued by the FFT kernels themselves (msfno_fft_stage_peer)."""
        self.group = group
        self.use_peer, self.peer = (peer_exchange if peer_exchange == "fused" else bool(peer_exchange)), None
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.nlat, self.nlon, self.lmax, self.mmax = nlat, nlon, lmax, mmax
        self.mlim = min(lmax, mmax)
        # latitude boundaries on multiples of four rows: every block of the lat<->m exchange is then 16-byte aligned
        self.lat_bounds = split_even(nlat, self.world, align=4)
        self.m_bounds = split_orders(lmax, mmax, self.world)
        self.poff, self.plen4, self.P = packed_offsets(lmax, mmax)
        self.lat_lo, self.lat_hi = self.lat_bounds[self.rank], self.lat_bounds[self.rank + 1]
        self.m_lo, self.m_hi = self.m_bounds[self.rank], self.m_bounds[self.rank + 1]
        self.nlat_loc = self.lat_hi - self.lat_lo
        self.stages = stages_factory(self.nlat_loc)

    def _peer(self, B, C, st):
        if not self.use_peer or B != 1 or not hasattr(st, "lat_segments") or dist.get_backend(self.group) != "nccl":
            return None
        if self.peer is None or self.peer.C != C:
            if self.peer is not None:
                self.peer.close()
            self.peer = PeerExchange(self, C, st.device, st.pad, fused=self.use_peer == "fused")
        return self.peer

    def _fused(self, B, C, st):
        """peer_exchange="fused": the exchange inside the four-step FFT kernels (nlon 240 / 1440 / 2880)."""
        return self.use_peer == "fused" and self.nlon in (240, 1440, 2880) and self._peer(B, C, st) is not None

    def _single_exchange(self, B, st):
        """all_to_all_single + one pack / unpack launch (NCCL, B = 1, stages that provide lat_segments)."""
        return (B == 1 and hasattr(st, "lat_segments") and self.world <= 16 and dist.get_backend(self.group) == "nccl")

    def pos_range(self, r=None):
        r = self.rank if r is None else r
        lo, hi = self.m_bounds[r], self.m_bounds[r + 1]
        p0 = self.poff[lo] if lo < self.mlim else self.P
        p1 = self.poff[hi] if hi < self.mlim else self.P
        return p0, p1

    def forward_packed(self, x_loc):
        st = self.stages
        B, C = x_loc.shape[0], x_loc.shape[1]
        mloc = self.m_hi - self.m_lo
        if self.world > 1 and self._fused(B, C, st):
            Xt = self.peer.fft_fwd_into_peers(st, x_loc.contiguous())
            if mloc == 0:
                return torch.zeros((B, 0, 2 * C), dtype=x_loc.dtype, device=x_loc.device)
            p0, p1 = self.pos_range()
            return st.legendre_fwd(Xt, self.m_lo, self.m_hi, p1 - p0)
        Xt_loc = st.fft_fwd(x_loc.contiguous())                       # [B, mlim, 2C, pad(nlat_loc)]
        if self.world == 1:
            Xt = Xt_loc                                               # the stage's own padded layout: nothing to move
        elif self._peer(B, C, st) is not None:
            Xt = self.peer.forward(Xt_loc)
        elif self._single_exchange(B, st):
            # B = 1 on the CUDA stages: a destination's orders are ONE contiguous run of the m-major intermediate, so
            # all_to_all_single sends straight out of the FFT stage's output (no send-side copy) and one gather launch
            # assembles the full-latitude operand of the Legendre stage from the received blocks
            padl = Xt_loc.shape[-1]
            in_splits = [(self.m_bounds[s + 1] - self.m_bounds[s]) * 2 * C * padl for s in range(self.world)]
            out_splits = [mloc * 2 * C * st.pad(self.lat_bounds[s + 1] - self.lat_bounds[s]) for s in range(self.world)]
            recv = torch.empty(max(sum(out_splits), 1), dtype=x_loc.dtype, device=x_loc.device)
            dist.all_to_all_single(recv[:sum(out_splits)], Xt_loc.view(-1), out_splits, in_splits, group=self.group)
            Xt = torch.empty((B, mloc, 2 * C, st.pad(self.nlat)), dtype=x_loc.dtype, device=x_loc.device)
            if mloc > 0:
                st.lat_segments(True, recv, Xt, self.lat_bounds)
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
        if self.world > 1 and self._fused(B, C, st):
            out = self.peer.synthesis_target()
            if mloc > 0:
                st.legendre_inv(cm_loc.contiguous(), self.m_lo, self.m_hi, out=out)
            return self.peer.fft_inv_from_peers(st, C)
        if mloc > 0:
            Yt = st.legendre_inv(cm_loc.contiguous(), self.m_lo, self.m_hi)  # [B, mloc, 2C, pad(nlat)]
        else:
            Yt = torch.zeros((B, 0, 2 * C, st.pad(self.nlat)), dtype=cm_loc.dtype, device=cm_loc.device)
        if self.world == 1:
            return st.fft_inv(Yt, B, C)
        if self._peer(B, C, st) is not None:
            return st.fft_inv(self.peer.inverse(Yt), B, C)
        if self._single_exchange(B, st):
            # one scatter launch writes every destination's block with that destination's padded pitch; the received
            # orders land directly in the FFT stage's m-major operand (a source's orders are one contiguous run of it)
            padl = st.pad(self.nlat_loc)
            in_splits = [mloc * 2 * C * st.pad(self.lat_bounds[s + 1] - self.lat_bounds[s]) for s in range(self.world)]
            out_splits = [(self.m_bounds[s + 1] - self.m_bounds[s]) * 2 * C * padl for s in range(self.world)]
            send = torch.empty(max(sum(in_splits), 1), dtype=cm_loc.dtype, device=cm_loc.device)
            if mloc > 0:
                st.lat_segments(False, send, Yt, self.lat_bounds)
            Yt_loc = torch.empty((B, self.mlim, 2 * C, padl), dtype=cm_loc.dtype, device=cm_loc.device)
            dist.all_to_all_single(Yt_loc.view(-1), send[:sum(in_splits)], out_splits, in_splits, group=self.group)
            return st.fft_inv(Yt_loc, B, C)
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
