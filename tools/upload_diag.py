"""Where the time of Engine.upload_csr goes at a bench workload (diagnostic)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genevector_b200 import _lib  # noqa: E402
from genevector_b200.engine import get_engine  # noqa: E402
from genevector_b200.synth import CONFIGS  # noqa: E402
import bench  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
    G, N = CONFIGS[wl]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    eng = get_engine(dev)
    X, _ = bench.synthetic_input(N, G, dev, keep_device=False)
    st = torch.cuda.current_stream(dev).cuda_stream
    for it in range(3):
        t = [time.perf_counter()]
        data, cols = np.ascontiguousarray(X.data), np.ascontiguousarray(X.indices, dtype=np.int32)
        ptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
        t.append(time.perf_counter())
        dv = torch.empty(data.shape, dtype=torch.float32, device=dev)
        dc = torch.empty(cols.shape, dtype=torch.int32, device=dev)
        dp = torch.empty(ptr.shape, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        t.append(time.perf_counter())
        for dst, src in ((dv, data), (dc, cols), (dp, ptr)):
            _lib.call("gv_upload_pageable", dst.data_ptr(), src.ctypes.data, src.nbytes, eng.copy_threads, st)
            t.append(time.perf_counter())
        torch.cuda.synchronize()
        t.append(time.perf_counter())
        ok = np.zeros(1, dtype=np.int32)
        scratch = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.call("gv_csr_check", dc.data_ptr(), _lib.GV_F32, dp.data_ptr(), N, G, scratch.data_ptr(), ok.ctypes.data, st)
        t.append(time.perf_counter())
        names = ["views", "torch.empty", "upload data", "upload cols", "upload ptr", "sync", "csr_check"]
        print(f"iter {it} threads={eng.copy_threads}: " +
              ", ".join(f"{n} {1e3*(b-a):.1f}" for n, a, b in zip(names, t, t[1:])), flush=True)
        t0 = time.perf_counter()
        csr = eng.upload_csr(X)
        torch.cuda.synchronize()
        print(f"   Engine.upload_csr total {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
        del csr, dv, dc, dp
    # the public call, as bench.py's e2e leg makes it
    from genevector_b200 import metrics as gm
    from genevector_b200.synth import gene_names
    genes = gene_names(G)
    last = None
    for it in range(5):
        eng.host_timings = {}
        t0 = time.perf_counter()
        last = gm.target_mi(X, genes, signed=True, backend="b200", device=dev)
        dt = time.perf_counter() - t0
        print(f"target_mi call {it}: {1e3*dt:.1f} ms  " +
              ", ".join(f"{k}: {v[0]:.1f}" for k, v in eng.host_timings.items()), flush=True)


if __name__ == "__main__":
    main()
