"""ScoreMatrix -- the pairwise score structure handed back to genevector.cache and
GeneVector.train.

The reference returns `defaultdict(lambda: defaultdict(float))` with `round(score, 5)` values
for every computed pair, both orders, no self pairs (genevector/metrics.py:113,137-140,
464-472).  At 20,000 genes that is 4*10^8 Python floats; this class is a read-only Mapping
*view* over one dense float32 [G][G] matrix that supports exactly what the reference's
consumers do with the dict --

    g in scores, scores[g1][g2], g2 in scores[g1], .items() / .keys() / len() at both levels
    (data.py:383-388, cache.py:74-82)

-- and exposes `.matrix` / `.genes` for the matrix fast paths (genevector_b200/data.py).
Values read through the mapping are `round(float(v), 5)` like the reference's; `.matrix` holds
the unrounded float32 scores.  A pair the reference skips (a gene without positive values,
metrics.py:119-120) reads 0.0, which is what `defaultdict(float)` gives its consumers.
"""
from __future__ import annotations

from collections.abc import Mapping

import numpy as np


class _Row(Mapping):
    __slots__ = ("_sm", "_i")

    def __init__(self, sm: "ScoreMatrix", i: int):
        self._sm, self._i = sm, i

    def __getitem__(self, gene):
        j = self._sm._index.get(gene)
        if j is None or j == self._i:
            # defaultdict(float) semantics for absent keys (self pairs are never stored)
            return 0.0
        return round(float(self._sm.matrix[self._i, j]), 5)

    def __contains__(self, gene):
        j = self._sm._index.get(gene)
        return j is not None and j != self._i

    def __iter__(self):
        i = self._i
        return (g for j, g in enumerate(self._sm.genes) if j != i)

    def __len__(self):
        return max(len(self._sm.genes) - 1, 0)

    def items(self):
        i = self._i
        row = np.round(self._sm.matrix[i].astype(np.float64), 5)
        return ((g, float(row[j])) for j, g in enumerate(self._sm.genes) if j != i)


class ScoreMatrix(Mapping):
    """Mapping[str, Mapping[str, float]] view over a dense symmetric score matrix."""

    def __init__(self, matrix, genes):
        matrix = np.asarray(matrix)
        if matrix.ndim != 2 or matrix.shape[0] != matrix.shape[1] or matrix.shape[0] != len(genes):
            raise ValueError("matrix must be [G][G] with one gene name per row")
        self.matrix = matrix
        self.genes = list(genes)
        self._index = {g: i for i, g in enumerate(self.genes)}
        if len(self._index) != len(self.genes):
            raise ValueError("gene names must be unique")

    def __getitem__(self, gene):
        return _Row(self, self._index[gene])

    def __contains__(self, gene):
        return gene in self._index

    def __iter__(self):
        return iter(self.genes)

    def __len__(self):
        return len(self.genes)

    def rounded(self) -> np.ndarray:
        """float64 matrix of the values the mapping yields (5 dp, zero diagonal)."""
        R = np.round(self.matrix.astype(np.float64), 5)
        np.fill_diagonal(R, 0.0)
        return R

    def to_dict(self):
        """Materialise the reference's nested dict (small gene counts only)."""
        import collections
        d = collections.defaultdict(lambda: collections.defaultdict(float))
        R = self.rounded()
        for i, g1 in enumerate(self.genes):
            for j, g2 in enumerate(self.genes):
                if i != j:
                    d[g1][g2] = float(R[i, j])
        return d
