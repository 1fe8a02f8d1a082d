"""Multi-rank correctness check on real GPUs (torchrun, NCCL): the sharded public calls (signed MI with
both front ends, Pearson, Spearman; count and log-normalised data) must hand EVERY rank the full
matrix, identical bit for bit to the single-rank result.  Exit code 1 on any difference.  Usage:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py [G N]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
from genevector_b200.engine import get_engine  # noqa: E402
from genevector_b200.metrics import compute_mi_b200, target_cosine, target_pearson, target_spearman  # noqa: E402
from genevector_b200.synth import synth_counts, gene_names  # noqa: E402
from genevector_b200 import _lib  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    X = synth_counts(N, G, seed=3)
    Xf = X.copy().astype(np.float32)
    Xf.data = np.log1p(Xf.data * 0.37).astype(np.float32)          # non-count data: fixed-point rows
    genes = gene_names(G)
    eng = get_engine(dev)
    # "auto" takes the sharded front end where the data qualify (counts) and the broadcast otherwise
    fronts = ["broadcast", "auto"]
    calls = [(f"signed_mi[{f}]", (lambda M, f=f: compute_mi_b200(M, genes, signed=True, front=f))) for f in fronts]
    calls += [("mi_unsigned", lambda M: compute_mi_b200(M, genes, signed=False)),
              ("pearson", lambda M: target_pearson(M, genes)),
              ("cosine", lambda M: target_cosine(M, genes)),
              ("spearman", lambda M: target_spearman(M, genes))]
    ok = True
    for name, fn in calls:
        for tag, M in (("counts", X), ("lognorm", Xf)):
            for rep in range(2):                       # twice: the second call reuses the shared host matrix
                got = fn(M).matrix
                c = eng.upload_csr(M)
                if name.startswith("signed_mi") or name == "mi_unsigned":
                    want = eng.mi_from_csr(c, N, G, signed=name.startswith("signed"))
                else:
                    want = eng.target_from_csr(c, _lib.GV_MOM_COSINE if name == "cosine" else _lib.GV_MOM_PEARSON,
                                               ranks=(name == "spearman"))
                same = np.array_equal(np.asarray(got).view(np.uint32), want.cpu().numpy().view(np.uint32))
                ok &= same
                print(f"world={world} rank={rank} {name:22s} {tag:8s} call {rep}: identical to single rank: {same}",
                      flush=True)
                del got
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("DIST CHECK", "OK" if int(flag.item()) == 0 else "FAILED", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
