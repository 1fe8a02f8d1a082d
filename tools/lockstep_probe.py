"""Diagnostic: C3 signed MI with the wave lockstep on / off (and cooperative launch on / off)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genevector_b200.engine import Engine, CsrDevice
from genevector_b200.synth import synth_counts_torch
mode = sys.argv[1]
dev = torch.device("cuda", 0)
eng = Engine(dev)
eng.lockstep = {"on": True, "off": False, "auto": "auto"}[mode]
v, c, p = synth_counts_torch(50000, 5000, seed=0, device=dev)
csr = CsrDevice(v, c, p, 50000, 5000)
for signed in (False, True):
    t0 = time.perf_counter()
    out = eng.mi_from_csr(csr, 50000, 5000, signed=signed)
    torch.cuda.synchronize()
    print(mode, "signed" if signed else "unsigned", "ok", f"{time.perf_counter()-t0:.3f}s", float(out.abs().sum()), flush=True)
