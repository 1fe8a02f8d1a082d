"""Is the FP4 MI kernel limited by the L2 -> SM feed?  Run it with a schedule in which every CTA pair
works on the SAME tile (all loads of a wave coalesce on a few L2 lines) and with the real schedule."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from genevector_b200 import _lib
from genevector_b200.engine import Engine, CsrDevice
from genevector_b200.synth import synth_counts_torch

G, N = 6000, 200000
dev = torch.device("cuda", 0)
eng = Engine(dev)
v, c, p = synth_counts_torch(N, G, seed=0, device=dev)
csr = CsrDevice(v, c, p, N, G)
blk = eng.discretize(csr, 10)
for operand in ("fp4", "int8"):
    onehot, B = eng.expand_onehot(blk, operand)
    rows, row_bytes = onehot.shape
    fmt = _lib.GV_OPERAND_FP4 if operand == "fp4" else _lib.GV_OPERAND_INT8
    tiles, n, _ = eng.schedule(rows)
    out = torch.zeros((G, G), dtype=torch.float32, device=dev)
    n_use = (n // 74) * 74
    variants = {"real schedule": tiles[:n_use].contiguous()}
    same = torch.zeros((n_use, 2), dtype=torch.int32, device=dev)
    variants["all pairs on tile (0,0)"] = same
    rowshare = tiles[:n_use].clone(); rowshare[:, 1] = (torch.arange(n_use, device=dev) % 8 * 256).int()
    rowshare[:, 0] = 0
    variants["same A panel, 8 B panels"] = rowshare.contiguous()
    for name, tl in variants.items():
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("gv_mi_tiles", onehot.data_ptr(), rows, row_bytes, B, blk.marg.data_ptr(), None, 0, G,
                      tl.data_ptr(), n_use, out.data_ptr(), out.stride(0), 2, fmt, eng._sync_ws.data_ptr(),
                      torch.cuda.current_stream().cuda_stream)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        issued = n_use * 2.0 * 256 * 240 * (blk.kp)
        print(f"{operand:5s} {name:28s}: {ms:8.2f} ms for {n_use} tiles -> issued {issued/ms/1e9:8.1f} TOP/s")
    # main loop alone: accumulators discarded (no epilogue work, no TMEM reads)
    for name, tl in variants.items():
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("gv_gram_dump", onehot.data_ptr(), rows, row_bytes, tl.data_ptr(), n_use, None, rows, 2, 30, fmt,
                      torch.cuda.current_stream().cuda_stream)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        issued = n_use * 2.0 * 256 * 240 * (blk.kp)
        print(f"{operand:5s} no epilogue, no lockstep, {name:24s}: {ms:8.2f} ms -> issued {issued/ms/1e9:8.1f} TOP/s")
    del onehot
