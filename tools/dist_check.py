"""Multi-rank correctness check on real GPUs (torchrun, NCCL): the sharded signed-MI and matrix-target
paths must reproduce the single-rank result exactly.  Usage:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
from genevector_b200.engine import Engine  # noqa: E402
from genevector_b200.metrics import compute_mi_b200, target_pearson, target_spearman  # noqa: E402
from genevector_b200.synth import synth_counts, gene_names  # noqa: E402
from genevector_b200 import _lib  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    X = synth_counts(3000, 1500, seed=3)
    Xf = X.copy().astype(np.float32)
    Xf.data = np.log1p(Xf.data * 0.37).astype(np.float32)          # non-count data: fixed-point rows
    genes = gene_names(1500)
    eng = Engine(torch.device("cuda", local))
    results = {}
    for name, fn in (("signed_mi", lambda M: compute_mi_b200(M, genes, signed=True)),
                     ("pearson", lambda M: target_pearson(M, genes)),
                     ("spearman", lambda M: target_spearman(M, genes))):
        for tag, M in (("counts", X), ("lognorm", Xf)):
            results[(name, tag)] = fn(M).matrix
    if rank == 0:
        # single-rank reference on the same device, outside the process group's sharding
        csr, csrf = eng.upload_csr(X), eng.upload_csr(Xf)
        ok = True
        for (name, tag), got in results.items():
            c = csr if tag == "counts" else csrf
            if name == "signed_mi":
                want = eng.mi_from_csr(c, 3000, 1500, signed=True)
            else:
                want = eng.target_from_csr(c, _lib.GV_MOM_PEARSON, ranks=(name == "spearman"))
            same = np.array_equal(got, want.cpu().numpy(), equal_nan=True)
            ok &= same
            print(f"world={world} {name:10s} {tag:8s} identical to single rank: {same}")
        print("DIST CHECK", "OK" if ok else "FAILED")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
