"""genevector_b200.data -- the score-building step of GeneVectorDataset on the matrix path.

Mirrors the two methods of the reference's dataset that sit on the co-expression path:

    compute_target_scores(...)   = GeneVectorDataset._compute_target_scores     data.py:292-334
    build_training_tensors(...)  = GeneVectorDataset._build_training_tensors    data.py:377-411
    save_target_scores / load_target_scores                                      data.py:235-251
    get_gene_entropy(...)        = GeneVectorDataset.get_gene_entropy            data.py:186-203

The first dispatches to the B200 targets (with the reference's cache protocol); the second turns a
ScoreMatrix into the (i_idx, j_idx, xij) training tensors with one CUDA kernel
(gv_build_training_tensors) instead of the reference's double Python loop over the nested dict.
Everything else of the dataset (Context, entropy, batching) is unchanged reference code and not
part of this package.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .engine import get_engine
from .scores import ScoreMatrix


def compute_target_scores(X, gene_names, target="mi", target_kwargs=None, signed_mi=True,
                          mi_backend="b200", device="cuda", use_cache=False):
    """data.py:292-334: cache probe -> target function -> cache save.  Returns a ScoreMatrix."""
    from .metrics import get_target_function
    target_kwargs = target_kwargs or {}
    cache_key = None
    if use_cache:
        from .cache import compute_cache_key, load_scores
        target_name = target if isinstance(target, str) else "custom"
        cache_key = compute_cache_key(X, gene_names, target_name, target_kwargs, signed_mi)
        cached, _ = load_scores(cache_key)
        if cached is not None:
            return cached
    if callable(target):
        scores = target(X, gene_names, **target_kwargs)
    elif isinstance(target, str):
        fn = get_target_function(target)
        kwargs = {"signed": signed_mi, "backend": mi_backend, "device": device, **target_kwargs}
        scores = fn(X, gene_names, **kwargs)
    else:
        raise TypeError(f"target must be str or callable, got {type(target)}")
    if use_cache:
        # one process per GPU: every rank holds the same full matrix (metrics._run_to_host); rank 0
        # alone writes the cache file
        from .cache import save_scores
        from .metrics import _dist_info
        if _dist_info()[0] == 0:
            save_scores(cache_key, scores, gene_names)
    return scores


def build_training_tensors(scores, genes, c=100.0, signed_mi=True, device="cuda"):
    """data.py:377-411 on the matrix path: (i_idx int64, j_idx int64, xij float32) on `device`,
    all off-diagonal ordered pairs in row-major order, xij = round(score, 5) * c**2, negatives
    clamped to 0 when not signed_mi.  `scores` is a ScoreMatrix (or a device/host float32 matrix
    whose rows follow `genes`)."""
    genes = list(genes)
    eng = get_engine(device if device not in (None, "cpu") else None)
    if isinstance(scores, ScoreMatrix):
        if scores.genes != genes:
            order = [scores._index[g] for g in genes]
            mat = scores.matrix[np.ix_(order, order)]
        else:
            mat = scores.matrix
    else:
        mat = scores
    m = torch.as_tensor(mat, dtype=torch.float32).to(eng.device).contiguous()
    n = len(genes)
    if m.shape != (n, n):
        raise ValueError("score matrix does not match the gene list")
    i_idx = torch.empty(n * (n - 1), dtype=torch.int64, device=eng.device)
    j_idx = torch.empty_like(i_idx)
    xij = torch.empty(n * (n - 1), dtype=torch.float32, device=eng.device)
    with torch.cuda.device(eng.device):
        _lib.call("gv_build_training_tensors", m.data_ptr(), m.stride(0), n, float(c),
                  0 if signed_mi else 1, 1, i_idx.data_ptr(), j_idx.data_ptr(), xij.data_ptr(),
                  torch.cuda.current_stream(eng.device).cuda_stream)
    return i_idx, j_idx, xij


def save_target_scores(scores, genes, filepath):
    """GeneVectorDataset.save_target_scores (data.py:235-239) on the matrix path.  The reference derives
    a cache key from the path and writes into the cache directory (its defect 9 in SURVEY Appendix D is
    kept: same file, same schema), so a file written here is found by the reference and vice versa."""
    from .cache import save_scores
    key = filepath.replace(".npz", "").replace("/", "_")
    return save_scores(key, scores, list(genes))


def load_target_scores(filepath):
    """GeneVectorDataset.load_target_scores (data.py:241-251) without the G^2 Python loop: the .npz
    schema of the cache (`scores` float32 [G][G], `genes`) straight into a ScoreMatrix.  Exact zeros
    read 0.0 through the mapping, which is what the reference's defaultdict yields for the pairs it
    drops (data.py:250)."""
    data = np.load(filepath, allow_pickle=False)
    matrix = np.array(data["scores"], dtype=np.float32)
    np.fill_diagonal(matrix, 0.0)
    return ScoreMatrix(matrix, [str(g) for g in data["genes"]])


def get_gene_entropy(X, gene_names, device="cuda"):
    """GeneVectorDataset.get_gene_entropy (data.py:186-203) on the device: {gene: entropy in nats of
    the counts of its unique values after dropping the first}.  Integer counts <= 255 go through a
    per-gene value histogram, any other non-negative data through a per-gene sort; negative values
    raise GvError (no CPU fallback)."""
    eng = get_engine(device if device not in (None, "cpu") else None)
    csr = eng.upload_csr(X)
    if len(gene_names) != csr.n_genes:
        raise ValueError("gene_names must have one entry per column of X")
    with torch.cuda.device(eng.device):
        dt = _lib.GV_F32 if csr.values.dtype == torch.float32 else _lib.GV_F64
        ws_bytes = eng.lib.gv_discretize_workspace_bytes(csr.nnz, csr.n_cells, csr.n_genes, dt, 1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=eng.device)
        ent = torch.empty(csr.n_genes, dtype=torch.float64, device=eng.device)
        _lib.call("gv_gene_entropy", csr.values.data_ptr(), dt, csr.col_idx.data_ptr(),
                  csr.row_ptr.data_ptr(), csr.nnz, csr.n_cells, csr.n_genes, ent.data_ptr(),
                  ws.data_ptr(), ws_bytes, torch.cuda.current_stream(eng.device).cuda_stream)
        vals = ent.cpu().numpy()
    return {g: float(v) for g, v in zip(gene_names, vals)}


# ------------------------------------------------------------------ methods for the reference's dataset
# metrics.install_into_reference binds these onto genevector.data.GeneVectorDataset; they only touch the
# attributes the reference's own methods touch (adata.X, adata.var.index, data.genes, mi_scores,
# signed_mi, device, _i_idx / _j_idx / _xij), so any object with those works.
def dataset_get_gene_entropy(adata):
    """GeneVectorDataset.get_gene_entropy (staticmethod, data.py:186-203)."""
    return get_gene_entropy(adata.X, adata.var.index.tolist())


def dataset_build_training_tensors(self, c=100.):
    """GeneVectorDataset._build_training_tensors (data.py:377-411)."""
    if not isinstance(self.mi_scores, ScoreMatrix):
        ref = getattr(type(self), "_build_training_tensors_reference", None)
        if ref is None:
            raise TypeError("mi_scores is not a ScoreMatrix and the reference implementation is not available")
        return ref(self, c=c)
    dev = self.device if str(self.device).startswith("cuda") else "cuda"
    i_idx, j_idx, xij = build_training_tensors(self.mi_scores, self.data.genes, c=c, signed_mi=self.signed_mi,
                                               device=dev)
    if not str(self.device).startswith("cuda"):
        i_idx, j_idx, xij = i_idx.to(self.device), j_idx.to(self.device), xij.to(self.device)
    self._i_idx, self._j_idx, self._xij = i_idx, j_idx, xij


def dataset_load_target_scores(self, filepath):
    """GeneVectorDataset.load_target_scores (data.py:241-251)."""
    self.mi_scores = load_target_scores(filepath)


def dataset_save_target_scores(self, filepath):
    """GeneVectorDataset.save_target_scores (data.py:235-239)."""
    return save_target_scores(self.mi_scores, self.data.genes, filepath)
