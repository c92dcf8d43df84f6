#!/usr/bin/env python3
"""Host <-> device copy bandwidth of this box, through the library's own staging pipeline (mlkem_b200_copy_probe: the chunks,
slots and streams of a host-memory call, no kernels).  One JSON line (rank 0): per-GPU and box-wide GB/s for H2D alone, D2H
alone and both at once, with ordinary pinned input buffers and with write-combined ones.

    python tools/pcie_bw.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_bw.py

Under torchrun every rank drives its own GPU and all ranks copy at the same time: the aggregate is what bounds bench.py's
e2e leg at N GPUs (the host's PCIe / memory fabric, not a kernel)."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import crystals_kyber_b200 as ck
from crystals_kyber_b200.lib import MEM_HOST, Opts

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
try:
    import pynvml

    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
except Exception:
    pass
dist = None
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = ck.load()
ITEM, N = 4096, 1 << 18  # 1 GiB per direction
o = Opts(local, MEM_HOST, None, 0, 0, 0)
bufs = {"pinned_in": lib.mlkem_b200_host_alloc(ITEM * N), "wc_in": lib.mlkem_b200_host_alloc_wc(ITEM * N), "out": lib.mlkem_b200_host_alloc(ITEM * N)}
for p in bufs.values():
    C.memset(p, 1, ITEM * N)  # first touch on this rank's NUMA node


def probe(ins, outs):
    ip, ib = (C.c_void_p * len(ins))(*ins), (C.c_size_t * len(ins))(*([ITEM] * len(ins)))
    op, ob = (C.c_void_p * len(outs))(*outs), (C.c_size_t * len(outs))(*([ITEM] * len(outs)))
    assert lib.mlkem_b200_copy_probe(N, len(ins), ip, ib, len(outs), op, ob, C.byref(o)) == 0, lib.mlkem_b200_last_error()


def timed(fn, reps=4):
    fn()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return ITEM * N * reps / dt / 1e9


res = {"ranks": world, "bytes_per_direction": ITEM * N}
for name, ins, outs in (("h2d_pinned", [bufs["pinned_in"]], []), ("h2d_write_combined", [bufs["wc_in"]], []), ("d2h", [], [bufs["out"]]),
                        ("both_pinned", [bufs["pinned_in"]], [bufs["out"]]), ("both_write_combined", [bufs["wc_in"]], [bufs["out"]])):
    g = timed(lambda: probe(ins, outs))
    res[name] = {"GBps_per_gpu_each_direction": g, "GBps_box_each_direction": g * world}
if rank == 0:
    print(json.dumps(res))
if dist is not None:
    dist.destroy_process_group()
