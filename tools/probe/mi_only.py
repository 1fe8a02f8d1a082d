import os, sys, torch
sys.path.insert(0, '.')
from genevector_b200.engine import Engine, CsrDevice
from genevector_b200.synth import synth_counts_torch
G, N = 20000, 200000
dev = torch.device("cuda", 0)
eng = Engine(dev)
v, c, p = synth_counts_torch(N, G, seed=0, device=dev)
csr = CsrDevice(v, c, p, N, G)
blk = eng.discretize(csr, 10)
onehot, B = eng.expand_onehot(blk, "fp4")
out = torch.zeros((G, G), dtype=torch.float32, device=dev)
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.mi_scores(blk, onehot, B, out=out); e1.record(); torch.cuda.synchronize()
    print("mi_tiles ms", e0.elapsed_time(e1))
