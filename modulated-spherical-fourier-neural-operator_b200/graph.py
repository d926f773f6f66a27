"""CUDA-graph capture of a whole forward step.

One SFNO forward is ~200 kernel launches of 5-500 us each; replaying them as one CUDA graph removes the per-launch
host cost and the gaps between dependent kernels (BASELINE config 4: 112-step autoregressive rollouts replay the
same graph with the output fed back as input).  Every entry point of the C ABI is capture-safe once the
per-shape descriptor tables exist, i.e. after one eager warm-up call -- which `GraphedForward` performs.
"""
import torch


class GraphedForward:
    """Wraps `net(x, *args)` (inference, fixed shapes) into a replayable CUDA graph with static I/O buffers."""

    def __init__(self, net, example_x, *example_args, warmup=2):
        self.net = net
        self.warmup = warmup
        self.static_x = example_x.clone()
        self.static_args = tuple(a.clone() if torch.is_tensor(a) else a for a in example_args)
        self._capture()

    def _capture(self):
        """The graph bakes in the addresses of the derived weight copies the launches read (padded / TF32-packed
        weights, packed spectral-MLP weights) and skips their re-packing: `self._refs` keeps every one of them alive for
        the life of the graph, so a later cache eviction can never leave a replay reading freed memory.  The graph does
        NOT see later parameter changes -- call recapture() after changing weights."""
        from . import _lib
        net = self.net
        _lib._capture_refs = refs = []
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.no_grad(), torch.cuda.stream(side):
                for _ in range(self.warmup):
                    self.static_y = net(self.static_x, *self.static_args)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(self.graph):
                self.static_y = net(self.static_x, *self.static_args)
        finally:
            _lib._capture_refs = None
        self._refs = refs
        if hasattr(self, "inplace_graph"):
            del self.inplace_graph

    def recapture(self):
        """Forget every derived weight copy and capture again (after the net's parameters changed)."""
        from . import _lib
        _lib.invalidate_caches()
        self._capture()

    def __call__(self, x=None, *args):
        """Copies x (and tensor args) into the static buffers, replays, returns the static output tensor
        (valid until the next call)."""
        if x is not None and x.data_ptr() != self.static_x.data_ptr():
            self.static_x.copy_(x, non_blocking=True)
        for dst, src in zip(self.static_args, args):
            if torch.is_tensor(dst) and src is not None and src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_y

    def _capture_inplace(self):
        """Second graph for rollouts: the decoder writes the forecast straight into the input buffer (`_decode_out`
        hook of the fused forward), so a 6 h step needs no feedback copy (303 MB read + write, 0.1 ms of a 5 ms step).
        Nets without that hook keep the copying loop."""
        self.inplace_graph = None
        if (not hasattr(self.net, "_decode_fused") or self.static_args
                or getattr(self.net, "in_chans", None) != getattr(self.net, "out_chans", -1)):
            return
        from . import _lib
        self.net._decode_out = self.static_x
        _lib._capture_refs = self._refs
        try:
            keep = self.static_x.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.no_grad(), torch.cuda.stream(side):
                y = self.net(self.static_x)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if y.data_ptr() == self.static_x.data_ptr():
                g = torch.cuda.CUDAGraph()
                with torch.no_grad(), torch.cuda.graph(g, pool=self.graph.pool()):
                    y = self.net(self.static_x)
                if y.data_ptr() == self.static_x.data_ptr():
                    self.inplace_graph = g
            self.static_x.copy_(keep)
        finally:
            _lib._capture_refs = None
            del self.net._decode_out

    def rollout(self, x0, steps, host_out=None, every=1):
        """Autoregressive rollout x <- net(x) (reference: /root/reference MSFNO/Models/sfno/model.py:327-331).
        Returns the state after `steps` steps (valid until the next call).

        host_out: optional sequence of (pinned) host tensors; host_out[k] receives the state after (k + 1) * every steps.
        The reference copies every step's fields to the host synchronously before the next step starts
        (sfno/model.py:345-370); here a step's state is snapshotted on the device (0.1 ms: the in-place graph overwrites
        its input) and drained to the host on a side stream under the following steps, double-buffered -- the rollout runs
        at max(step, device->host copy) instead of their sum."""
        if not hasattr(self, "inplace_graph"):
            self._capture_inplace()
        self.static_x.copy_(x0, non_blocking=True)
        inplace = self.inplace_graph is not None
        if host_out is None:
            if inplace:
                for _ in range(steps):
                    self.inplace_graph.replay()
                return self.static_x
            for _ in range(steps):
                self.graph.replay()
                self.static_x.copy_(self.static_y, non_blocking=True)
            return self.static_y
        main = torch.cuda.current_stream()
        if not hasattr(self, "_snap"):
            self._snap = [torch.empty_like(self.static_x) for _ in range(2)]
            self._s_out = torch.cuda.Stream()
            self._ev_out = [torch.cuda.Event() for _ in range(2)]
        k = 0
        for t in range(1, steps + 1):
            if inplace:
                self.inplace_graph.replay()
                state = self.static_x
            else:
                self.graph.replay()
                state = self.static_y
            if t % every == 0 and k < len(host_out):
                j = k % 2
                if k >= 2:
                    main.wait_event(self._ev_out[j])           # staging buffer j has been drained to the host
                self._snap[j].copy_(state, non_blocking=True)   # the next step may now overwrite the state
                ev = torch.cuda.Event()
                ev.record(main)
                self._s_out.wait_event(ev)
                with torch.cuda.stream(self._s_out):
                    host_out[k].copy_(self._snap[j], non_blocking=True)
                    self._ev_out[j].record(self._s_out)
                k += 1
            if not inplace:
                self.static_x.copy_(self.static_y, non_blocking=True)
        main.wait_stream(self._s_out)
        return self.static_x if inplace else self.static_y
