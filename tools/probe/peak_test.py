import sys, torch
sys.path.insert(0, '.')
from genevector_b200.engine import Engine
e = Engine()
for op in ("int8", "fp4"):
    for cg in (1, 2):
        e.cta_group = cg
        for it in (20000, 60000):
            print(op, "cg", cg, "iters", it, "TOP/s", round(e.mma_peak(op, it), 1))
