"""Wall-clock breakdown of the public call compute_mi_b200(scipy CSR) at a given workload (pageable
host arrays in, numpy matrix out).  Diagnostic; bench.py's e2e leg uses pinned buffers."""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genevector_b200.engine import get_engine  # noqa: E402
from genevector_b200.metrics import compute_mi_b200  # noqa: E402
from genevector_b200.synth import synth_counts_torch, CONFIGS, gene_names  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C4"
    G, N = CONFIGS[wl]
    dev = torch.device("cuda", 0)
    v, c, p = synth_counts_torch(N, G, seed=0, device=dev)
    X = sp.csr_matrix((v.cpu().numpy(), c.cpu().numpy(), p.cpu().numpy()), shape=(N, G))
    del v, c, p
    genes = gene_names(G)
    eng = get_engine()
    for it in range(3):
        t0 = time.perf_counter()
        canon = X.has_canonical_format
        t1 = time.perf_counter()
        csr = eng.upload_csr(X)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out = eng.mi_from_csr(csr, N, G, signed=True)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        m = out.cpu().numpy()
        t4 = time.perf_counter()
        del csr, out, m
        t5 = time.perf_counter()
        sm = compute_mi_b200(X, genes, signed=True)
        t6 = time.perf_counter()
        print(f"{wl} iter {it}: canonical check {1e3*(t1-t0):.0f} ms ({canon}), upload {1e3*(t2-t1):.0f} ms, "
              f"compute {1e3*(t3-t2):.0f} ms, result to numpy {1e3*(t4-t3):.0f} ms | "
              f"compute_mi_b200 total {1e3*(t6-t5):.0f} ms")
        del sm


if __name__ == "__main__":
    main()
