"""genevector_b200.cache -- the reference's score cache (genevector/cache.py) with a matrix path.

Same key, same directory, same `.npz` schema (`scores` float32 [G][G] + `genes`) as the
reference (cache.py:29-44, 47-61, 64-87), so files written by either side load on the other.
The O(G^2) Python loops of save_scores / load_scores (cache.py:74-82, 104-107) are replaced by
array operations on the ScoreMatrix; load_scores returns a ScoreMatrix instead of a nested dict.
"""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np
from scipy.sparse import issparse

from .scores import ScoreMatrix

CACHE_DIR = os.path.expanduser("~/.genevector/cache")


def _hash_array(X):
    """cache.py:13-26."""
    h = hashlib.sha256()
    if issparse(X):
        h.update(np.ascontiguousarray(X.data).tobytes())
        h.update(np.ascontiguousarray(X.indices).tobytes())
        h.update(np.ascontiguousarray(X.indptr).tobytes())
    else:
        h.update(np.ascontiguousarray(X).tobytes())
    h.update(str(X.shape).encode())
    return h.hexdigest()[:16]


def compute_cache_key(X, gene_names, target_name, target_kwargs, signed_mi):
    """cache.py:29-44 (same bytes hashed in the same order, so keys are interchangeable)."""
    h = hashlib.sha256()
    h.update(_hash_array(X).encode())
    h.update(json.dumps(sorted(gene_names)).encode())
    h.update(target_name.encode() if isinstance(target_name, str) else b"custom")
    h.update(json.dumps(target_kwargs, sort_keys=True, default=str).encode())
    h.update(str(signed_mi).encode())
    return h.hexdigest()[:32]


def get_cache_path(cache_key):
    os.makedirs(CACHE_DIR, exist_ok=True)
    return os.path.join(CACHE_DIR, f"scores_{cache_key}.npz")


def save_scores(cache_key, mi_scores, gene_names):
    """cache.py:64-87.  `mi_scores` may be a ScoreMatrix (matrix path) or any nested mapping."""
    gene_names = list(gene_names)
    if isinstance(mi_scores, ScoreMatrix) and mi_scores.genes == gene_names:
        matrix = mi_scores.rounded().astype(np.float32)     # the dict values are round(score, 5)
    else:
        n = len(gene_names)
        idx = {g: i for i, g in enumerate(gene_names)}
        matrix = np.zeros((n, n), dtype=np.float32)
        for g1, inner in mi_scores.items():
            i = idx.get(g1)
            if i is None:
                continue
            for g2, val in inner.items():
                j = idx.get(g2)
                if j is not None:
                    matrix[i, j] = val
    path = get_cache_path(cache_key)
    # written under a private name and moved into place: a concurrent reader (another rank, another
    # process) never sees a partial file
    tmp = f"{path[:-4]}.tmp{os.getpid()}.npz"
    np.savez_compressed(tmp, scores=matrix, genes=np.array(gene_names))
    os.replace(tmp, path)
    return path


def load_scores(cache_key):
    """cache.py:90-110.  Returns (ScoreMatrix, gene_names) or (None, None)."""
    path = get_cache_path(cache_key)
    if not os.path.exists(path):
        return None, None
    data = np.load(path, allow_pickle=False)
    gene_names = [str(g) for g in data["genes"]]
    matrix = np.array(data["scores"], dtype=np.float32)
    np.fill_diagonal(matrix, 0.0)                           # self pairs are never read (:106)
    return ScoreMatrix(matrix, gene_names), gene_names


def clear_cache():
    if os.path.exists(CACHE_DIR):
        import shutil
        shutil.rmtree(CACHE_DIR)
