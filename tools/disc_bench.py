"""Time the discretisation entry point alone (not a bench line; used under ncu for the per-kernel
numbers in profiles/).  Usage: python tools/disc_bench.py [workload] [reps] [val_mode]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genevector_b200.engine import Engine, CsrDevice  # noqa: E402
from genevector_b200.synth import synth_counts_torch, CONFIGS  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    val_mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    G, N = CONFIGS[wl]
    dev = torch.device("cuda", 0)
    eng = Engine(dev)
    vals, cols, ptr = synth_counts_torch(N, G, seed=0, device=dev)
    csr = CsrDevice(vals, cols, ptr, N, G)
    buf = None
    for i in range(reps + 1):
        eng.kernel_events = []
        blk = eng.discretize(csr, 10, val_mode=val_mode, buf=buf)
        buf = blk.buf
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for (_, a, b, _) in eng.kernel_events]
        if i:
            print(f"{wl} val_mode={val_mode} nnz={csr.nnz} discretize {ms[0]:.3f} ms  path={blk.path_used}")


if __name__ == "__main__":
    main()
