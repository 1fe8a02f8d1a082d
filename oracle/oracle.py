"""oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's all-pairs co-expression path (nceglia/genevector; all
citations are into /root/reference).  Two flavours of every function:

  *_np   numpy, written against the same third-party calls the reference makes
         (np.percentile / np.searchsorted / np.unique / np.corrcoef), so third-party arithmetic
         is identical by construction;
  *_c    the plain-C restatement in gv_oracle.c (arithmetic restated from scratch; OpenMP),
         used for larger shapes, for raw uint32 joint counts and as the CPU baseline.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (genevector_b200/) never does.  Pinned against the reference
itself through tests/golden/*.npz (oracle/gen_golden.py runs the reference's own functions).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
import scipy.sparse as _sp
from scipy.sparse import issparse

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgvoracle.so")
_lib = None


def build() -> str:
    """Compile gv_oracle.c (gcc, OpenMP).  Building the checker is not using it."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "gv_oracle.c")
        if not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
            build()
        L = C.CDLL(_SO)
        L.gvo_num_threads.restype = C.c_int
        L.gvo_quantile_fraction.restype = C.c_double
        L.gvo_quantile_fraction.argtypes = [C.c_int, C.c_int]
        L.gvo_discretize_dense.restype = C.c_int
        L.gvo_discretize_dense.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p]
        L.gvo_discretize_csr.restype = C.c_int
        L.gvo_discretize_csr.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gvo_joint.restype = C.c_uint32
        L.gvo_joint.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                C.c_int, C.c_void_p]
        L.gvo_mi_from_joint_u32.restype = C.c_double
        L.gvo_mi_from_joint_u32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32]
        L.gvo_mi_pairs.restype = C.c_int64
        L.gvo_mi_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int,
                                   C.c_void_p]
        L.gvo_mi_pairs_gm.restype = C.c_int64
        L.gvo_mi_pairs_gm.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                      C.c_void_p]
        L.gvo_sign_int.restype = None
        L.gvo_sign_int.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().gvo_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the C oracle (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().gvo_set_num_threads(int(n))
    return num_threads()


def _dense(X):
    return np.asarray(X.todense()) if issparse(X) else np.asarray(X)


# --------------------------------------------------------------------------- discretisation
def quantile_table(n_bins_max: int = 31) -> np.ndarray:
    """q[k][t-1] = np.linspace(0,100,k+1)[1:-1][t-1] / 100 (metrics.py:64 + np.percentile's /100)."""
    q = np.zeros((32, 32), dtype=np.float64)
    for k in range(2, n_bins_max + 1):
        q[k, :k - 1] = np.linspace(0, 100, k + 1)[1:-1] / 100
    return q


def discretize_genes_np(X, n_bins: int = 10):
    """metrics.py:25-69, statement by statement."""
    X = _dense(X)
    n_cells, n_genes = X.shape
    X_disc = np.zeros((n_cells, n_genes), dtype=np.int32)
    n_bins_per_gene = np.zeros(n_genes, dtype=np.int32)
    for g in range(n_genes):
        col = X[:, g].ravel()
        nz = col > 0                                              # :53
        vals = col[nz]
        if len(vals) == 0:
            n_bins_per_gene[g] = 1                                # :55-57
            continue
        k = min(n_bins, len(np.unique(vals)))                     # :58-59
        if k <= 1:
            X_disc[nz, g] = 1                                     # :60-62
            n_bins_per_gene[g] = 2
        else:
            edges = np.percentile(vals, np.linspace(0, 100, k + 1)[1:-1])   # :64-65
            X_disc[nz, g] = np.searchsorted(edges, vals, side="right") + 1  # :66
            n_bins_per_gene[g] = k + 1                            # :67
    return X_disc, n_bins_per_gene


def discretize_genes_c(X, n_bins: int = 10, return_edges: bool = False):
    """gv_oracle.c gvo_discretize_dense (dense, cell-major int32 result like the reference)."""
    X = np.ascontiguousarray(_dense(X))
    if X.dtype not in (np.float32, np.float64):
        X = X.astype(np.float64)
    n_cells, n_genes = X.shape
    X_disc = np.zeros((n_cells, n_genes), dtype=np.int32)
    nb = np.zeros(n_genes, dtype=np.int32)
    edges = np.zeros((n_genes, 32), dtype=np.float64)
    rc = lib().gvo_discretize_dense(X.ctypes.data, 0 if X.dtype == np.float32 else 1, n_cells,
                                    n_genes, n_bins, X_disc.ctypes.data, nb.ctypes.data,
                                    edges.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"gvo_discretize_dense rc={rc}")
    return (X_disc, nb, edges) if return_edges else (X_disc, nb)


def discretize_csr_c(X, n_bins: int = 10):
    """gv_oracle.c gvo_discretize_csr: gene-major uint8 bins [G][N] (no dense temporary)."""
    X = X.tocsr()
    X.sum_duplicates()
    vals = np.ascontiguousarray(X.data)
    if vals.dtype not in (np.float32, np.float64):
        vals = vals.astype(np.float64)
    idx = np.ascontiguousarray(X.indices, dtype=np.int32)
    ptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
    n_cells, n_genes = X.shape
    bins = np.zeros((n_genes, n_cells), dtype=np.uint8)
    nb = np.zeros(n_genes, dtype=np.int32)
    edges = np.zeros((n_genes, 32), dtype=np.float64)
    rc = lib().gvo_discretize_csr(vals.ctypes.data, 0 if vals.dtype == np.float32 else 1,
                                  idx.ctypes.data, ptr.ctypes.data, n_cells, n_genes, n_bins,
                                  bins.ctypes.data, nb.ctypes.data, edges.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"gvo_discretize_csr rc={rc}")
    return bins, nb, edges


# --------------------------------------------------------------------------- joint counts + MI
def joint_np(col_a, col_b, na, nb):
    """metrics.py:122-129: joint histogram over cells where either bin is non-zero."""
    mask = (col_a > 0) | (col_b > 0)
    joint = np.zeros((na, nb), dtype=np.float64)
    np.add.at(joint, (col_a[mask], col_b[mask]), 1)
    return joint


def mi_from_joint_np(pxy):
    """metrics.py:74-81 (_mi_from_joint)."""
    pxy = pxy / pxy.sum()
    px = pxy.sum(axis=1)
    py = pxy.sum(axis=0)
    px_py = np.outer(px, py)
    nz = (pxy > 0) & (px_py > 0)
    return np.sum(pxy[nz] * np.log2(pxy[nz] / px_py[nz]))


def pearson_sign_np(X):
    """metrics.py:106-111,134: np.sign(nan_to_num(np.corrcoef(X_dense.T)))."""
    with np.errstate(invalid="ignore", divide="ignore"):
        corr = np.nan_to_num(np.corrcoef(_dense(X).T), nan=0.0)
    return corr


def mi_matrix_np(X, n_bins: int = 10, signed: bool = False):
    """compute_mi_vectorized (metrics.py:86-142) returning the raw (unrounded) float64 matrix.
    Pure-Python pair loop: small cases only."""
    X_disc, nb = discretize_genes_np(X, n_bins)
    G = X_disc.shape[1]
    corr = pearson_sign_np(X) if signed else None
    out = np.zeros((G, G), dtype=np.float64)
    for i in range(G):
        for j in range(i + 1, G):
            if nb[i] <= 1 or nb[j] <= 1:                          # :119-120
                continue
            ca, cb = X_disc[:, i], X_disc[:, j]
            if ((ca > 0) | (cb > 0)).sum() == 0:                  # :125-126
                continue
            mi = mi_from_joint_np(joint_np(ca, cb, nb[i], nb[j]))
            if signed:
                mi = np.sign(corr[i, j]) * mi                     # :132-135
            out[i, j] = out[j, i] = mi
    return out


def joint_c(X_disc, nb, i, j):
    """Raw uint32 joint table of one pair and its count (lib.rs:46-56)."""
    X_disc = np.ascontiguousarray(X_disc, dtype=np.int32)
    na, nbj = int(nb[i]), int(nb[j])
    joint = np.zeros((na, nbj), dtype=np.uint32)
    n_cells, n_genes = X_disc.shape
    cnt = lib().gvo_joint(X_disc.ctypes.data, n_cells, n_genes, i, j, na, nbj, joint.ctypes.data)
    return joint, int(cnt)


def mi_pairs_c(X_disc, nb, corr=None, sign_mode: int = 1):
    """gv_oracle.c gvo_mi_pairs == src/lib.rs:30-98; returns (float64 [G][G], n_pairs_computed)."""
    X_disc = np.ascontiguousarray(X_disc, dtype=np.int32)
    nb = np.ascontiguousarray(nb, dtype=np.int32)
    n_cells, n_genes = X_disc.shape
    out = np.zeros((n_genes, n_genes), dtype=np.float64)
    cp = None
    if corr is not None:
        corr = np.ascontiguousarray(corr, dtype=np.float64)
        cp = corr.ctypes.data
    n = lib().gvo_mi_pairs(X_disc.ctypes.data, nb.ctypes.data, n_cells, n_genes, cp, sign_mode,
                           out.ctypes.data)
    return out, int(n)


def mi_pairs_gm_c(bins_gm, nb, sign=None):
    """Gene-major uint8 variant for larger shapes; `sign` optional int8 [G][G]."""
    bins_gm = np.ascontiguousarray(bins_gm, dtype=np.uint8)
    nb = np.ascontiguousarray(nb, dtype=np.int32)
    n_genes, n_cells = bins_gm.shape
    out = np.zeros((n_genes, n_genes), dtype=np.float64)
    sp = None
    if sign is not None:
        sign = np.ascontiguousarray(sign, dtype=np.int8)
        sp = sign.ctypes.data
    n = lib().gvo_mi_pairs_gm(bins_gm.ctypes.data, nb.ctypes.data, n_cells, n_genes, sp,
                              out.ctypes.data)
    return out, int(n)


def sign_int_c(X):
    """Exact integer Pearson sign for count matrices (int8 [G][G])."""
    Xd = _dense(X)
    vals = np.ascontiguousarray(Xd.T.astype(np.int32))
    if not np.array_equal(vals, Xd.T):
        raise ValueError("sign_int_c needs integer-valued data")
    n_genes, n_cells = vals.shape
    sign = np.zeros((n_genes, n_genes), dtype=np.int8)
    lib().gvo_sign_int(vals.ctypes.data, n_cells, n_genes, sign.ctypes.data)
    return sign


# --------------------------------------------------------------------------- matrix targets
def pearson_np(X):
    """target_pearson (metrics.py:414-420) without the dict: np.corrcoef(X.T)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.corrcoef(_dense(X).T)


def spearman_np(X):
    """target_spearman (metrics.py:423-431) without the dict: scipy.stats.spearmanr(X_dense)."""
    from scipy.stats import spearmanr
    with np.errstate(invalid="ignore", divide="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            corr, _ = spearmanr(_dense(X))
    if np.ndim(corr) == 0:
        corr = np.array([[1.0, corr], [corr, 1.0]])
    return np.asarray(corr)


def rank_rows_np(X):
    """What the B200 path contracts for Spearman (checker for the value rows, not reference code):
    per gene #{y < x} + #{y <= x} = 2 * average rank - 1 as integers [n_genes][n_cells]; genes without
    negative values are shifted by their number of zero cells, so that those cells hold 0."""
    from scipy.stats import rankdata
    Xd = _dense(X)
    out = np.zeros(Xd.T.shape, dtype=np.int64)
    for g in range(Xd.shape[1]):
        col = Xd[:, g]
        r2m1 = np.rint(2 * rankdata(col, method="average")).astype(np.int64) - 1
        if (col < 0).any():
            out[g] = r2m1
        else:
            n_zero = int((col == 0).sum())
            out[g] = np.where(col > 0, r2m1 - n_zero, 0)
    return out


def fixed_point_rows_np(X, total_bits):
    """What the B200 path contracts for Pearson / cosine / the sign of non-count data (checker for
    the value rows): per gene rint((x - o) * 2^e) with o = min(0, min x) (0 for non-negative genes),
    e = total_bits - exponent(top), top = max(max x, 0) - o = f * 2^exponent, f in [0.5, 1); capped at
    2^total_bits - 1.  For data without negative values only x > 0 contributes.  [n_genes][n_cells]
    int64; also returns nothing else: use fixed_point_offsets_np for the per-gene zero-cell value."""
    Xd = _dense(X).astype(np.float64)
    has_neg = bool((Xd < 0).any())
    out = np.zeros(Xd.T.shape, dtype=np.int64)
    for g in range(Xd.shape[1]):
        col = Xd[:, g] if has_neg else np.where(Xd[:, g] > 0, Xd[:, g], 0.0)
        o = min(0.0, float(col.min())) if col.size else 0.0
        top = max(float(col.max()), 0.0) - o if col.size else 0.0
        if top <= 0:
            continue
        _, ex = np.frexp(top)
        out[g] = np.clip(np.rint(np.ldexp(col - o, int(total_bits - ex))).astype(np.int64), 0,
                         (1 << total_bits) - 1)
    return out


def int_moment_matrices(V):
    """Exact integer co-moment statistics of an integer matrix V [n_genes][n_cells] (Python ints):
    returns (cov, var) with cov[i,j] = n*Sum(v_i v_j) - Sum(v_i) Sum(v_j) as object arrays."""
    V = np.asarray(V).astype(object)
    n = V.shape[1]
    s = V.sum(axis=1)
    cov = n * (V @ V.T) - np.outer(s, s)
    return cov, np.diag(cov).copy()


def int_sign_exact(V):
    """np.sign of the exact integer covariance, 0 where either variance is 0, zero diagonal."""
    cov, var = int_moment_matrices(V)
    G = cov.shape[0]
    sg = np.zeros((G, G), dtype=np.int8)
    for i in range(G):
        for j in range(G):
            if i != j and var[i] != 0 and var[j] != 0:
                sg[i, j] = (cov[i, j] > 0) - (cov[i, j] < 0)
    return sg


def int_pearson_exact(V):
    """Pearson correlation from the exact integer moments (float64 quotient), NaN where a variance is 0."""
    cov, var = int_moment_matrices(V)
    with np.errstate(invalid="ignore", divide="ignore"):
        c = cov.astype(np.float64) / np.sqrt(np.outer(var.astype(np.float64), var.astype(np.float64)))
    return np.clip(c, -1.0, 1.0)


def seeded_graph(n_cells, seed):
    """Sparse neighbour graph with 0..6 weighted neighbours per cell (some cells isolated)."""
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    for c in range(n_cells):
        deg = int(rng.integers(0, 7))
        nb = rng.choice(n_cells, size=deg, replace=False)
        rows += [c] * deg
        cols += nb.tolist()
        vals += rng.choice([1.0, 1.0, 0.5, 2.0], size=deg).tolist()
    return _sp.csr_matrix((vals, (rows, cols)), shape=(n_cells, n_cells), dtype=np.float64)


def graph_xcorr_np(X, graph, include_self=False):
    """target_graph_xcorr with aggr="mean" (genevector/_graph_targets.py:38-52 and
    genevector/_aggregation.py:48-109) without the dict: float64 [G][G], zero diagonal."""
    from scipy.sparse import diags
    X_dense = np.asarray(_dense(X), dtype=np.float64)                   # _aggregation.py:66-84
    row_sums = np.asarray(graph.sum(axis=1)).ravel()                    # :59-63
    row_sums[row_sums == 0] = 1.0
    W = diags(1.0 / row_sums) @ graph
    X_agg = W @ X_dense                                                 # :105-106
    if include_self:
        X_agg = 0.5 * X_dense + 0.5 * X_agg                             # :107-108
    n_cells = X_dense.shape[0]
    X_std = (X_dense - X_dense.mean(axis=0)) / (X_dense.std(axis=0) + 1e-8)      # _graph_targets.py:45
    X_agg_std = (X_agg - X_agg.mean(axis=0)) / (X_agg.std(axis=0) + 1e-8)        # :46
    xcorr = (X_std.T @ X_agg_std) / n_cells                             # :48
    sym = (xcorr + xcorr.T) / 2                                         # :49
    np.fill_diagonal(sym, 0)
    return np.asarray(sym)


def cosine_np(X):
    """target_cosine (metrics.py:452-461): sklearn cosine_similarity(X.T), diagonal zeroed."""
    from sklearn.metrics.pairwise import cosine_similarity
    sim = cosine_similarity(X.T)
    sim = np.asarray(sim)
    np.fill_diagonal(sim, 0)
    return sim


def jaccard_np(X):
    """target_jaccard (metrics.py:434-449)."""
    if issparse(X):
        binary_dense = np.asarray((X > 0).astype(np.float32).todense())
    else:
        binary_dense = (np.asarray(X) > 0).astype(np.float32)
    intersection = binary_dense.T @ binary_dense
    sums = binary_dense.sum(axis=0)
    union = sums[:, None] + sums[None, :] - intersection
    union[union == 0] = 1
    jaccard = intersection / union
    np.fill_diagonal(jaccard, 0)
    return np.array(jaccard)


def round5_dict_matrix(M):
    """What the reference stores: round(float(v), 5) for i != j (metrics.py:139-140, 464-472)."""
    R = np.round(np.asarray(M, dtype=np.float64), 5)
    np.fill_diagonal(R, 0.0)
    return R


# --------------------------------------------------------------------------- gene entropy
def gene_entropy_np(X):
    """GeneVectorDataset.get_gene_entropy (data.py:186-203): per gene, scipy entropy of the counts
    of the unique values with the first count dropped."""
    from scipy.stats import entropy
    Xd = _dense(X).T
    out = np.zeros(Xd.shape[0])
    for g, exp in enumerate(Xd):
        counts = np.unique(exp, return_counts=True)
        out[g] = entropy(counts[1][1:])
    return out
