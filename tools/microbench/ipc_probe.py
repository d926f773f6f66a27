"""Probe: can one process per GPU write into a peer's cudaMalloc'd buffer through CUDA IPC (NVLink P2P)?  torchrun, >= 2 GPUs.
Each rank allocates a buffer, all ranks open all handles, every rank writes its id into its slot of EVERY peer's buffer,
barrier, every rank checks its own buffer; then a bandwidth number for a large peer write."""
import ctypes
import os

import torch
import torch.distributed as dist


class Handle(ctypes.Structure):      # cudaIpcMemHandle_t travels BY VALUE into cudaIpcOpenMemHandle
    _fields_ = [("reserved", ctypes.c_char * 64)]


class Arr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rt = ctypes.CDLL("libcudart.so.12")
    n = 64 << 20   # floats (256 MB)
    p = ctypes.c_void_p()
    assert rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(4 * n)) == 0
    h = Handle()
    rc = rt.cudaIpcGetMemHandle(ctypes.byref(h), p)
    rt.cudaIpcOpenMemHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), Handle, ctypes.c_uint]
    handles = [None] * world
    dist.all_gather_object(handles, (rc, bytes(bytearray(h))))
    peers = []
    for r, (rc_r, raw) in enumerate(handles):
        if r == rank:
            peers.append(p.value)
            continue
        q = ctypes.c_void_p()
        hb = Handle.from_buffer_copy(raw)
        rc2 = rt.cudaIpcOpenMemHandle(ctypes.byref(q), hb, 1)
        if rc2 != 0:
            print("rank %d: cudaIpcOpenMemHandle(rank %d) failed rc=%d (export rc=%d)" % (rank, r, rc2, rc_r), flush=True)
            peers.append(None)
        else:
            peers.append(q.value)
    ok = all(x is not None for x in peers)
    if ok:
        for r in range(world):
            t = torch.as_tensor(Arr(peers[r], n), device=dev)
            t[rank * 1024:(rank + 1) * 1024].fill_(float(rank + 1))
        torch.cuda.synchronize()
        dist.barrier()
        mine = torch.as_tensor(Arr(p.value, n), device=dev)
        good = all(bool((mine[r * 1024:(r + 1) * 1024] == r + 1).all()) for r in range(world))
        src = torch.randn(n, device=dev)
        peer = torch.as_tensor(Arr(peers[(rank + 1) % world], n), device=dev)
        for _ in range(2):
            peer.copy_(src)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            peer.copy_(src)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("rank %d: peer writes visible=%s, 256 MB peer copy %.3f ms = %.0f GB/s" % (rank, good, ms, 4 * n / ms / 1e6), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
