"""TEST INFRASTRUCTURE (lives under oracle/ because it runs the oracle).  BASELINE.md section 4: the reference's own Numba pair kernel (unmodified, JIT-warmed, driven through the
dummy-column wrapper that neutralises its index bug) timed beside the C/OpenMP port of src/lib.rs that
bench.py uses as its CPU arm.  Build container only (needs /root/reference); the GPU box has no reference
tree, so this is a builder-run number kept in profiles/."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_golden  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from genevector_b200.synth import synth_counts  # noqa: E402


def main():
    import numba
    refm = gen_golden.import_reference_metrics()
    ncpu = os.cpu_count()
    orc.set_num_threads(ncpu)
    numba.set_num_threads(min(ncpu, numba.config.NUMBA_NUM_THREADS))
    print(f"cpu_count={ncpu} numba threads={numba.get_num_threads()} OMP threads={orc.num_threads()}")
    for n_genes, n_cells in ((300, 5000), (200, 50000), (96, 200000)):
        X = synth_counts(n_cells, n_genes, seed=0)
        t0 = time.perf_counter()
        xd, nb = refm.discretize_genes(X, n_bins=10)
        t_disc = time.perf_counter() - t0
        xd = np.ascontiguousarray(xd)
        pairs = n_genes * (n_genes - 1) // 2
        Xp = np.concatenate([np.zeros((n_cells, 1), np.int32), xd], axis=1)
        nbp = np.concatenate([[1], nb]).astype(np.int32)
        refm._compute_all_mi_numba(np.ascontiguousarray(Xp[:64]), nbp, n_genes + 1)      # JIT warm-up
        t0 = time.perf_counter()
        flat = refm._compute_all_mi_numba(np.ascontiguousarray(Xp), nbp, n_genes + 1)
        t_numba = time.perf_counter() - t0
        t0 = time.perf_counter()
        M, n = orc.mi_pairs_c(xd, nb)
        t_port = time.perf_counter() - t0
        t0 = time.perf_counter()
        orc.discretize_genes_c(X, 10)
        t_disc_port = time.perf_counter() - t0
        # same numbers (the dummy column shifts the slots: pair (i, j) sits in the slot of (i, j + 1))
        Gp = n_genes + 1
        chk = [(0, 1), (3, 17), (n_genes - 2, n_genes - 1)]
        err = max(abs(flat[i * Gp - i * (i + 1) // 2 + (j + 1 - i - 1)] - M[i, j]) for i, j in chk)
        print(f"{n_genes} genes x {n_cells} cells ({pairs} pairs): reference Numba kernel {pairs / t_numba:10.0f} pairs/s "
              f"({t_numba:.2f} s) | C/OpenMP port of src/lib.rs {pairs / t_port:10.0f} pairs/s ({t_port:.2f} s) | "
              f"discretize_genes: reference {t_disc:.2f} s, port {t_disc_port:.2f} s | max |diff| on spot pairs {err:.1e}")


if __name__ == "__main__":
    main()
