"""genevector_b200.rust_compat -- drop-in for the reference's native module `genevector._rust`.

The reference's only FFI is one PyO3 function (src/lib.rs:9-15, stub genevector/_rust.pyi:1-5)

    compute_mi_pairs(x_disc: int32 [n_cells, n_genes], n_bins_per_gene: int32 [n_genes],
                     corr_signs: float32 [n_genes, n_genes] | None) -> list[(i, j, mi)]

called from compute_mi_rust (genevector/metrics.py:312-339) after the CPU-side discretize_genes and
np.corrcoef.  `compute_mi_pairs` here has the same signature and return value and runs on B200 through
the C entry gv_compute_mi_pairs (include/gvb200.h): host arrays in, the pair list out.  A maintainer who
wants to swap only the Rust module writes, in genevector/metrics.py:16-20,

    from genevector_b200.rust_compat import compute_mi_pairs as _rust_mi_pairs

and keeps everything else.  `compute_mi_pairs_matrix` returns the dense float32 matrix instead of the
O(G^2) Python list (which at 20,000 genes is 2*10^8 tuples).
"""
from __future__ import annotations

import numpy as np

from . import _lib


def compute_mi_pairs_matrix(x_disc, n_bins_per_gene, corr_signs=None, operand: str = "auto", device=None):
    """Dense float32 [G][G] scores (both triangles, zero diagonal, zero for the pairs the Rust function
    skips) and the number of triples it would return."""
    import torch
    x = np.ascontiguousarray(x_disc, dtype=np.int32)
    if x.ndim != 2:
        raise ValueError("x_disc must be [n_cells, n_genes]")
    n_cells, n_genes = x.shape
    nb = np.ascontiguousarray(n_bins_per_gene, dtype=np.int32)
    if nb.shape != (n_genes,):
        raise ValueError("n_bins_per_gene must have one entry per gene")
    corr = None
    if corr_signs is not None:
        corr = np.ascontiguousarray(corr_signs, dtype=np.float32)
        if corr.shape != (n_genes, n_genes):
            raise ValueError("corr_signs must be [n_genes, n_genes]")
    if not torch.cuda.is_available():
        raise _lib.GvError("compute_mi_pairs", -3, "no CUDA device: genevector_b200 has no CPU fallback")
    dev = torch.device(device if device is not None else "cuda")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    fmt = {"auto": -1, "int8": _lib.GV_OPERAND_INT8, "fp4": _lib.GV_OPERAND_FP4}[operand]
    out = np.empty((n_genes, n_genes), dtype=np.float32)
    n_pairs = np.zeros(1, dtype=np.int64)
    with torch.cuda.device(idx):
        _lib.call("gv_compute_mi_pairs", x.ctypes.data, n_cells, n_genes, nb.ctypes.data,
                  corr.ctypes.data if corr is not None else None, out.ctypes.data, n_pairs.ctypes.data, fmt,
                  torch.cuda.current_stream().cuda_stream)
    return out, int(n_pairs[0])


def compute_mi_pairs(x_disc, n_bins_per_gene, corr_signs=None):
    """genevector._rust.compute_mi_pairs: list of (i, j, mi) for every i < j whose genes both have more
    than one bin (src/lib.rs:30-38, 97)."""
    out, _ = compute_mi_pairs_matrix(x_disc, n_bins_per_gene, corr_signs)
    nb = np.asarray(n_bins_per_gene)
    keep = np.flatnonzero(nb > 1)
    iu, ju = np.triu_indices(len(keep), 1)
    i, j = keep[iu], keep[ju]
    vals = out[i, j].astype(np.float64)
    return list(zip(i.tolist(), j.tolist(), vals.tolist()))
