"""Host-buffer serving pipeline: forecasts whose inputs and outputs live in (pinned) HOST memory.

The reference's inference loop copies every step's fields device->host synchronously before the next step
(/root/reference MSFNO/Models/sfno/model.py:345-370) and `Trainer.save_forecast` does the same for zarr output
(train.py:942-1022).  Here the H2D copy of member i+1 and the D2H copy of member i-1 overlap the forward of
member i on separate CUDA streams (double-buffered device staging), so throughput is bounded by
max(compute, PCIe) instead of their sum.
"""
import os

import torch


def bind_host_to_device(device):
    """Pin this process (and so its later pinned-memory allocations, by first touch) to the CPUs of the NUMA node the GPU
    hangs off: with one rank per GPU on an 8-GPU box every rank otherwise allocates its staging buffers wherever the kernel
    puts them, and half of the ranks' 600 MB per step cross the socket interconnect.  Reads
    /sys/bus/pci/devices/<bdf>/local_cpulist; returns the CPU list it bound to, or None when the topology is not exposed
    (containers without sysfs PCI entries) -- never raises."""
    try:
        idx = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
        bdf = torch.cuda.get_device_properties(idx).pci_bus_id if hasattr(torch.cuda.get_device_properties(idx), "pci_bus_id") else None
        if bdf is None:
            import ctypes
            buf = ctypes.create_string_buffer(32)
            rt = ctypes.CDLL("libcudart.so.12")
            if rt.cudaDeviceGetPCIBusId(buf, 32, idx) != 0:
                return None
            bdf = buf.value.decode()
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:      # cudart prints an 8-digit domain, sysfs uses 4
            bdf = bdf[4:]
        path = "/sys/bus/pci/devices/%s/local_cpulist" % bdf
        with open(path) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return spec
    except Exception:
        return None


class HostPipeline:
    """net: a CUDA nn.Module mapping [B, Cin, H, W] -> [B, Cout, H, W]; run() consumes / fills host tensors."""

    def __init__(self, net, device=None, depth=2):
        self.net = net
        self.device = device if device is not None else next(net.parameters()).device
        self.depth = depth
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self.xd, self.yd = [None] * depth, [None] * depth
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]
        self.ev_compute = [torch.cuda.Event() for _ in range(depth)]
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]

    @torch.no_grad()
    def run(self, xs_host, ys_host, *net_args):
        """xs_host / ys_host: equal-length sequences of host tensors (pinned for true overlap); ys_host[i] receives
        net(xs_host[i], *net_args).  Returns after all outputs are enqueued; the caller's current stream is made to
        wait for the last device->host copy, so a stream/device synchronize afterwards guarantees completion."""
        main = torch.cuda.current_stream(self.device)
        for i, (xh, yh) in enumerate(zip(xs_host, ys_host)):
            j = i % self.depth
            if self.xd[j] is None or self.xd[j].shape != xh.shape:
                self.xd[j] = torch.empty(xh.shape, dtype=xh.dtype, device=self.device)
            if i >= self.depth:
                self.s_in.wait_event(self.ev_compute[j])   # the forward that read xd[j] has finished
            else:
                self.s_in.wait_stream(main)
            with torch.cuda.stream(self.s_in):
                self.xd[j].copy_(xh, non_blocking=True)
                self.ev_in[j].record(self.s_in)
            main.wait_event(self.ev_in[j])
            y = self.net(self.xd[j], *net_args)
            if self.yd[j] is None or self.yd[j].shape != y.shape:
                self.yd[j] = torch.empty_like(y)
            if i >= self.depth:
                main.wait_event(self.ev_out[j])            # yd[j] has been drained to the host
            self.yd[j].copy_(y)
            self.ev_compute[j].record(main)
            self.s_out.wait_event(self.ev_compute[j])
            with torch.cuda.stream(self.s_out):
                yh.copy_(self.yd[j], non_blocking=True)
                self.ev_out[j].record(self.s_out)
        main.wait_stream(self.s_out)
        main.wait_stream(self.s_in)
