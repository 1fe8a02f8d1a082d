import sys, torch
sys.path.insert(0, '.')
from genevector_b200.engine import Engine
e = Engine()
for op in ("int8", "fp4"):
    for mode in (0, 4, 5):
        print(op, "mode", mode, "TOP/s", round(e.mma_peak(op, 60000, mode), 1))
