"""Host -> device upload rates of a pageable array on this box: gv_upload_pageable by thread count,
plain torch copies, a pinned copy, host memcpy alone, cudaHostRegister in place.  Diagnostic."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genevector_b200 import _lib  # noqa: E402


def main():
    nbytes = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else (3 << 29)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    a = np.random.default_rng(0).integers(0, 255, size=nbytes, dtype=np.uint8)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    gb = nbytes / 1e9

    def t(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best

    for th in (1, 2, 4, 8, 12, 16):
        s = t(lambda: _lib.call("gv_upload_pageable", d.data_ptr(), a.ctypes.data, nbytes, th, st))
        print(f"gv_upload_pageable threads={th:2d}: {s*1e3:7.1f} ms  {gb/s:6.1f} GB/s", flush=True)
    assert np.array_equal(d[:1 << 20].cpu().numpy(), a[:1 << 20])
    ta = torch.from_numpy(a)
    s = t(lambda: d.copy_(ta))
    print(f"torch pageable copy_           : {s*1e3:7.1f} ms  {gb/s:6.1f} GB/s")
    tp = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    s = t(lambda: tp.copy_(ta))
    print(f"host memcpy pageable->pinned (torch, {torch.get_num_threads()} thr): {s*1e3:7.1f} ms  {gb/s:6.1f} GB/s")
    b = np.empty_like(a)
    s = t(lambda: np.copyto(b, a))
    print(f"host memcpy numpy 1 thread     : {s*1e3:7.1f} ms  {gb/s:6.1f} GB/s")
    s = t(lambda: d.copy_(tp, non_blocking=True))
    print(f"pinned -> device               : {s*1e3:7.1f} ms  {gb/s:6.1f} GB/s")
    s = t(lambda: tp.copy_(d, non_blocking=True))
    print(f"device -> pinned               : {s*1e3:7.1f} ms  {gb/s:6.1f} GB/s")
    rt = torch.cuda.cudart()
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(a.ctypes.data, nbytes, 0)
    t1 = time.perf_counter()
    print(f"cudaHostRegister in place rc={rc}: {(t1-t0)*1e3:7.1f} ms  {gb/(t1-t0):6.1f} GB/s")
    if int(rc) == 0:
        s = t(lambda: d.copy_(ta, non_blocking=True))
        print(f"registered -> device           : {s*1e3:7.1f} ms  {gb/s:6.1f} GB/s")
        t0 = time.perf_counter()
        rt.cudaHostUnregister(a.ctypes.data)
        print(f"cudaHostUnregister             : {(time.perf_counter()-t0)*1e3:7.1f} ms")


if __name__ == "__main__":
    main()
