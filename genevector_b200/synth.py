"""Synthetic sparse count matrices of the benchmark shapes.

Per gene: mean mu_g ~ Gamma(shape=2, scale=1), detection rate p_g = clip(0.10 * mu_g / 2, 0.005, 0.6)
(overall density ~10 %, typical of scRNA-seq); x_cg = Bernoulli(p_g) * (1 + Poisson(mu_g)).
Returned as scipy CSR [n_cells, n_genes], float32 values, int32 indices.  `synth_counts` is the
numpy generator (tests, small benches); `synth_counts_torch` draws the same distribution on a
torch device so the 20,000 x 200,000 case is generated in seconds, not minutes.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

CONFIGS = {
    # name: (n_genes, n_cells)   BASELINE.json configs[1..4]
    "C2": (2000, 5000),
    "C3": (5000, 50000),
    "C4": (20000, 200000),
    "C5": (10000, 100000),
}


def gene_params(n_genes: int, rng: np.random.Generator):
    mu = rng.gamma(shape=2.0, scale=1.0, size=n_genes)
    p = np.clip(0.10 * mu / 2.0, 0.005, 0.6)
    return mu, p


def synth_counts(n_cells: int, n_genes: int, seed: int = 0, dtype=np.float32,
                 block_rows: int = 8192) -> sp.csr_matrix:
    rng = np.random.default_rng(seed)
    mu, p = gene_params(n_genes, rng)
    blocks = []
    for r0 in range(0, n_cells, block_rows):
        r = min(block_rows, n_cells - r0)
        det = rng.random((r, n_genes)) < p[None, :]
        cnt = 1 + rng.poisson(np.broadcast_to(mu[None, :], (r, n_genes)))
        blocks.append(sp.csr_matrix(np.where(det, cnt, 0).astype(dtype)))
    X = sp.vstack(blocks, format="csr")
    X.indices = X.indices.astype(np.int32)
    X.indptr = X.indptr.astype(np.int32)
    return X


def gene_names(n_genes: int):
    return [f"GENE{i}" for i in range(n_genes)]


def synth_counts_torch(n_cells: int, n_genes: int, seed: int = 0, device="cuda",
                       block_rows: int = 4096):
    """Same distribution drawn with torch on `device`.  Returns CSR pieces as torch tensors on
    the device: (values float32 [nnz], col_idx int32 [nnz], row_ptr int64 [n_cells+1])."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    # Gamma(2,1) = sum of two Exp(1)
    u = torch.rand(2, n_genes, generator=g, device=device).clamp_min(1e-12)
    mu = -(u[0].log() + u[1].log())
    p = (0.10 * mu / 2.0).clamp(0.005, 0.6)
    vals, cols, counts = [], [], []
    for r0 in range(0, n_cells, block_rows):
        r = min(block_rows, n_cells - r0)
        det = torch.rand(r, n_genes, generator=g, device=device) < p[None, :]
        idx = det.nonzero()                      # row-major order: CSR order
        lam = mu[idx[:, 1]]
        cnt = 1.0 + torch.poisson(lam, generator=g)
        vals.append(cnt.to(torch.float32))
        cols.append(idx[:, 1].to(torch.int32))
        counts.append(det.sum(dim=1))
    values = torch.cat(vals)
    col_idx = torch.cat(cols)
    row_ptr = torch.zeros(n_cells + 1, dtype=torch.int64, device=device)
    row_ptr[1:] = torch.cumsum(torch.cat(counts), dim=0)
    return values, col_idx, row_ptr
