#!/usr/bin/env python
"""Host->device copy bandwidth from default pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined), and device->host
into default pinned memory, one process per GPU.   [torchrun ...] python tools/dbg/wc_probe.py"""
import ctypes as ct, json, os, time
import numpy as np, torch
import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rt = ct.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else ct.CDLL("libcudart.so")
N = 327680000   # bytes: the int16 samples of one configs[1] step
dev = torch.empty(N, dtype=torch.uint8, device="cuda")
res = {}
for tag, flags in (("pinned_default", 0), ("pinned_write_combined", 4), ("pinned_portable_wc", 4 | 1)):
    p = ct.c_void_p()
    assert rt.cudaHostAlloc(ct.byref(p), ct.c_size_t(N), ct.c_uint(flags)) == 0
    ct.memset(p, 1, N)
    for it in range(2):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            assert rt.cudaMemcpyAsync(ct.c_void_p(dev.data_ptr()), p, ct.c_size_t(N), ct.c_int(1), ct.c_void_p(0)) == 0
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        dt = (time.perf_counter() - t0) / 5
    res[tag] = N / dt / 1e9
    rt.cudaFreeHost(p)
t = torch.tensor([res[k] for k in sorted(res)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print(json.dumps({"n_gpus": world, "h2d_GBps_min_over_ranks": dict(zip(sorted(res), [float(v) for v in t]))}))
if world > 1:
    dist.destroy_process_group()
