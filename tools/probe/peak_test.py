import sys, torch
sys.path.insert(0, '.')
from genevector_b200.engine import Engine
e = Engine()
for op in ("int8", "fp4"):
    for mode in (0, 5, 6, 7):
        for iters in (60000, 400000):
            print(op, "mode", mode, "iters", iters, "TOP/s", round(e.mma_peak(op, iters, mode), 1))
