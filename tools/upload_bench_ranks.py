"""Concurrent host -> device upload rates with one process per GPU (torchrun): every rank uploads its
1/world share of a C4-sized CSR (two arrays) at the same time, as the sharded front end does.  Prints
the slowest rank's time per method.  Diagnostic."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genevector_b200 import _lib  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    total = int(float(sys.argv[1]) * 1e9) if len(sys.argv) > 1 else 3_200_000_000
    n = total // world // 2
    arrays = [np.random.default_rng(rank * 2 + k).integers(0, 255, size=n, dtype=np.uint8) for k in range(2)]
    dsts = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(2)]
    st = torch.cuda.current_stream(dev).cuda_stream

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            best = min(best, float(dt.item()))
        return best

    def staged(th):
        for d, a in zip(dsts, arrays):
            _lib.call("gv_upload_pageable", d.data_ptr(), a.ctypes.data, a.nbytes, th, st)

    def registered():
        toks = []
        for d, a in zip(dsts, arrays):
            t = C.c_void_p()
            _lib.call("gv_upload_registered", d.data_ptr(), a.ctypes.data, a.nbytes, st, C.byref(t))
            toks.append(t)
        return toks

    def reg_and_release():
        for t in registered():
            _lib.call("gv_upload_release", t.value)

    res = {}
    for th in (1, 2, 4, 8):
        res[f"staged x{th}"] = timed(lambda: staged(th))
    pending = []

    def reg_only():
        pending.extend(registered())

    res["registered (copy done, pages still locked)"] = timed(reg_only, reps=1)
    for t in pending:
        _lib.call("gv_upload_release", t.value)
    res["registered + release"] = timed(reg_and_release, reps=2)
    if rank == 0:
        gb = total / 1e9
        print(f"world={world} cpus={os.cpu_count()} stage copy={'memcpy' if os.environ.get('GV_STAGE_MEMCPY') == '1' else 'nt-store'} "
              f"reg chunk MB={os.environ.get('GV_REG_CHUNK_MB', 'whole')}: {gb:.1f} GB in all, {2 * n / 1e6:.0f} MB per rank")
        for k, v in res.items():
            print(f"  {k:45s} {v * 1e3:8.1f} ms   {gb / v:7.1f} GB/s aggregate")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
