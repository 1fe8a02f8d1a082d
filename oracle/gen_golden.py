"""gen_golden.py -- TEST INFRASTRUCTURE.  Generates tests/golden/*.npz from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference):  python oracle/gen_golden.py
It imports the reference's genevector/metrics.py with the package __init__ bypassed (the
__init__ pulls matplotlib/scanpy, absent here), runs the reference's own functions
(discretize_genes, compute_mi_vectorized, _mi_from_joint, _compute_all_mi_numba, target_pearson,
target_spearman, target_cosine, target_jaccard) on seeded inputs and stores inputs + outputs.  It also checks the
oracle (oracle/oracle.py, gv_oracle.c) against those outputs before writing, so a golden file is
only ever produced by a run in which oracle == reference.
"""
import importlib
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference_metrics():
    pkg = types.ModuleType("genevector")
    pkg.__path__ = [os.path.join(REF, "genevector")]
    sys.modules["genevector"] = pkg
    return importlib.import_module("genevector.metrics")


def import_reference_graph_targets():
    return importlib.import_module("genevector._graph_targets")


def import_reference_data():
    return importlib.import_module("genevector.data")


class _FakeAnnData:
    """The two attributes GeneVectorDataset.get_gene_entropy touches (data.py:196-198)."""
    def __init__(self, X, genes):
        import pandas
        self.X = X
        self.var = pandas.DataFrame(index=genes)


def dict_to_matrix(scores, genes):
    idx = {g: i for i, g in enumerate(genes)}
    M = np.zeros((len(genes), len(genes)), dtype=np.float64)
    for g1, inner in scores.items():
        for g2, v in inner.items():
            M[idx[g1], idx[g2]] = v
    return M


def numba_raw_mi(refm, X_disc, nb):
    """Unmodified reference Numba kernel through the dummy-column trick: its
    flat-index recovery (metrics.py:158-159) is off by one, so prepend an all-zero gene and read
    pair (i,j) from the slot of (i, j+1)."""
    n_cells, G = X_disc.shape
    Xp = np.concatenate([np.zeros((n_cells, 1), np.int32), X_disc], axis=1)
    nbp = np.concatenate([[1], nb]).astype(np.int32)
    flat = refm._compute_all_mi_numba(np.ascontiguousarray(Xp), nbp, G + 1)
    M = np.zeros((G, G))
    # row-major upper triangle of the (G+1)-gene problem: slot of (p,q), p<q
    Gp = G + 1
    def slot(p, q):
        return p * Gp - p * (p + 1) // 2 + (q - p - 1)
    for i in range(G):
        for j in range(i + 1, G):
            M[i, j] = M[j, i] = flat[slot(i, j + 1)]
    return M


def make_case(refm, name, X, n_bins, with_targets=True, with_numba=True):
    from oracle import oracle as orc
    genes = [f"GENE{i}" for i in range(X.shape[1])]
    X = sp.csr_matrix(X)
    X_disc, nb = refm.discretize_genes(X, n_bins=n_bins)
    # ---- oracle == reference (discretisation)
    Xd_np, nb_np = orc.discretize_genes_np(X, n_bins)
    Xd_c, nb_c, edges_c = orc.discretize_genes_c(X, n_bins, return_edges=True)
    assert np.array_equal(X_disc, Xd_np) and np.array_equal(nb, nb_np), name
    assert np.array_equal(X_disc, Xd_c) and np.array_equal(nb, nb_c), name
    bins_gm, nb_csr, _ = orc.discretize_csr_c(X, n_bins)
    assert np.array_equal(bins_gm.T.astype(np.int32), X_disc) and np.array_equal(nb_csr, nb), name
    # ---- reference MI (NumPy tier = semantic oracle), rounded dicts
    mi_u = dict_to_matrix(refm.compute_mi_vectorized(X, genes, n_bins=n_bins, signed=False), genes)
    mi_s = dict_to_matrix(refm.compute_mi_vectorized(X, genes, n_bins=n_bins, signed=True), genes)
    corr = np.nan_to_num(np.corrcoef(np.asarray(X.todense()).T), nan=0.0)   # metrics.py:111
    # ---- raw (unrounded) MI from the reference's own functions
    G = X.shape[1]
    raw = np.zeros((G, G))
    for i in range(G):
        for j in range(i + 1, G):
            if nb[i] <= 1 or nb[j] <= 1:
                continue
            ca, cb = X_disc[:, i], X_disc[:, j]
            mask = (ca > 0) | (cb > 0)
            if mask.sum() == 0:
                continue
            joint = np.zeros((nb[i], nb[j]))
            np.add.at(joint, (ca[mask], cb[mask]), 1)
            raw[i, j] = raw[j, i] = refm._mi_from_joint(joint)
    assert np.allclose(np.round(raw, 5), mi_u, atol=1e-12), name
    if with_numba:
        raw_nb = numba_raw_mi(refm, X_disc, nb)
        assert np.abs(raw_nb - raw).max() < 1e-12, (name, np.abs(raw_nb - raw).max())
    # ---- oracle == reference (MI)
    raw_c, _ = orc.mi_pairs_c(X_disc, nb)
    assert np.abs(raw_c - raw).max() < 1e-12, (name, np.abs(raw_c - raw).max())
    raw_gm, _ = orc.mi_pairs_gm_c(bins_gm, nb)
    assert np.array_equal(raw_gm, raw_c), name
    raw_np = orc.mi_matrix_np(X, n_bins, signed=False)
    assert np.abs(raw_np - raw).max() < 1e-12, name
    sgn_c, _ = orc.mi_pairs_c(X_disc, nb, corr=corr, sign_mode=1)
    assert np.allclose(np.round(sgn_c, 5), mi_s, atol=1e-12), name
    # integer sign == np.sign(corr) wherever the covariance is not exactly zero
    fsign = np.sign(corr).astype(np.int8)
    np.fill_diagonal(fsign, 0)
    Xd = np.asarray(X.todense())
    if np.array_equal(Xd, np.rint(Xd)):
        isign = orc.sign_int_c(X)
        diff = (isign != fsign)
        n_amb = int(diff.sum())
        assert np.all(isign[diff] == 0), (name, "integer sign disagrees away from zero covariance")
    else:
        isign, n_amb = fsign, 0
    # ---- a few raw joint tables
    pairs = [(0, 1), (0, G - 1), (G // 2, G // 2 + 1)]
    joints = {}
    for (i, j) in pairs:
        if nb[i] > 1 and nb[j] > 1:
            jt, cnt = orc.joint_c(X_disc, nb, i, j)
            jn = orc.joint_np(X_disc[:, i], X_disc[:, j], nb[i], nb[j])
            assert np.array_equal(jt.astype(np.float64), jn) and cnt == int(jn.sum()), name
            joints[f"joint_{i}_{j}"] = jt
    out = dict(
        data=X.data, indices=X.indices.astype(np.int32), indptr=X.indptr.astype(np.int64),
        shape=np.array(X.shape, dtype=np.int64), n_bins=np.int32(n_bins),
        x_disc=X_disc.astype(np.int8), n_bins_per_gene=nb, edges=edges_c,
        mi_raw=raw, mi_unsigned_round5=mi_u, mi_signed_round5=mi_s, corr_sign=fsign,
        int_sign=isign, n_sign_ambiguous=np.int64(n_amb), **joints)
    # gene entropy straight from the reference's GeneVectorDataset.get_gene_entropy (data.py:186-203)
    refd = import_reference_data()
    ent = refd.GeneVectorDataset.get_gene_entropy(_FakeAnnData(X, genes))
    out["gene_entropy"] = np.array([ent[g] for g in genes])
    assert np.array_equal(out["gene_entropy"], orc.gene_entropy_np(X), equal_nan=True), name
    if with_targets:
        out["pearson"] = dict_to_matrix(refm.target_pearson(X, genes), genes)
        out["cosine"] = dict_to_matrix(refm.target_cosine(X, genes), genes)
        out["jaccard"] = dict_to_matrix(refm.target_jaccard(X, genes), genes)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out["spearman"] = dict_to_matrix(refm.target_spearman(X, genes), genes)
        assert np.allclose(orc.round5_dict_matrix(orc.spearman_np(X)), out["spearman"], atol=1e-12, equal_nan=True)
        # the integer rank rows reproduce spearmanr exactly (up to the float64 quotient)
        sp_int = orc.int_pearson_exact(orc.rank_rows_np(X))
        ref_sp = orc.spearman_np(X)
        ok = ~np.isnan(ref_sp)
        assert np.abs(sp_int[ok] - ref_sp[ok]).max() < 1e-12 and np.array_equal(np.isnan(sp_int), np.isnan(ref_sp)), name
        assert np.allclose(orc.round5_dict_matrix(orc.pearson_np(X)), out["pearson"], atol=1e-12, equal_nan=True)
        assert np.allclose(orc.round5_dict_matrix(orc.cosine_np(X)), out["cosine"], atol=1e-12)
        assert np.allclose(orc.round5_dict_matrix(orc.jaccard_np(X)), out["jaccard"], atol=1e-12)
    if with_targets:
        # graph cross-correlation straight from the reference's target_graph_xcorr (aggr="mean")
        refg = import_reference_graph_targets()
        Wg = orc.seeded_graph(X.shape[0], seed=len(name))
        out["graph_data"], out["graph_indices"], out["graph_indptr"] = (
            Wg.data, Wg.indices.astype(np.int32), Wg.indptr.astype(np.int64))
        out["graph_xcorr"] = dict_to_matrix(refg.target_graph_xcorr(X, genes, graph=Wg), genes)
        out["graph_xcorr_self"] = dict_to_matrix(
            refg.target_graph_xcorr(X, genes, graph=Wg, aggr_params={"include_self": True}), genes)
        assert np.allclose(orc.round5_dict_matrix(orc.graph_xcorr_np(X, Wg)), out["graph_xcorr"], atol=1e-12)
        assert np.allclose(orc.round5_dict_matrix(orc.graph_xcorr_np(X, Wg, True)), out["graph_xcorr_self"], atol=1e-12)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: shape={X.shape} nnz={X.nnz} n_bins={n_bins} sum(nbpg)={int(nb.sum())} "
          f"sign-ambiguous={n_amb} -> {os.path.getsize(path)/1024:.0f} KB")
    return out


def main():
    refm = import_reference_metrics()
    from genevector_b200.synth import synth_counts
    # 1. the reference's own test generator (tests/test_metrics.py:15-20): tie-heavy Poisson(2)
    rng = np.random.RandomState(42)
    X = rng.poisson(2, size=(500, 50)).astype(np.float32)
    X[X < 1] = 0
    g = make_case(refm, "ref_poisson_500x50_b10", X, 10)
    # SURVEY Appendix C known answers
    assert g["n_bins_per_gene"][:12].tolist() == [8, 9, 8, 9, 9, 9, 8, 8, 9, 9, 7, 10]
    assert abs(g["mi_raw"][0, 1] - 0.04174775719939659) < 1e-15
    assert g["mi_signed_round5"][0, 1] == -0.04175
    # 2. same generator, n_bins=5, 20 genes (tests/test_metrics.py:34-47)
    rng = np.random.RandomState(42)
    X = rng.poisson(2, size=(500, 20)).astype(np.float32)
    X[X < 1] = 0
    make_case(refm, "ref_poisson_500x20_b5", X, 5)
    # 3. synthetic scRNA-like counts incl. an all-zero gene, a constant gene,
    #    a gene detected in one cell and a gene with explicit zeros / a negative entry
    X = synth_counts(700, 96, seed=0).toarray()
    X[:, 5] = 0
    X[:, 6] = np.where(X[:, 6] > 0, 3, 0)
    X[:, 7] = 0; X[13, 7] = 4
    X[:, 8] = 2
    make_case(refm, "synth_counts_700x96_b10", X, 10)
    # 4. float64 log-normalised data (non-integer values; discretisation arithmetic in float64)
    Xc = synth_counts(300, 40, seed=3).toarray().astype(np.float64)
    tot = Xc.sum(axis=1, keepdims=True) + 1.0
    Xl = np.log1p(Xc / tot * 1e4)
    make_case(refm, "lognorm_f64_300x40_b10", Xl, 10)
    # 5. float32 log-normalised (float32 subtraction inside the lerp)
    make_case(refm, "lognorm_f32_300x40_b7", Xl.astype(np.float32), 7)
    # 6. wider bins
    X = synth_counts(400, 30, seed=5).toarray()
    make_case(refm, "synth_counts_400x30_b15", X * 3.0, 15)
    # 7. n_bins 16..31 (one byte per cell in the bin matrix, 32 operand rows per gene)
    make_case(refm, "synth_counts_400x30_b16", X * 3.0, 16)
    make_case(refm, "lognorm_f32_300x40_b20", Xl.astype(np.float32), 20)
    rng = np.random.RandomState(7)
    Xp = rng.poisson(20, size=(500, 24)).astype(np.float32)
    Xp[rng.rand(500, 24) < 0.35] = 0
    Xp[:, 3] = 0                                        # an all-zero gene
    Xp[:, 4] = np.where(Xp[:, 4] > 0, 7, 0)             # one unique value
    make_case(refm, "poisson20_500x24_b31", Xp, 31)
    # 8. gene pairs whose covariance is EXACTLY zero while np.corrcoef returns rounding noise
    #    (-4.4e-17 and the like): the reference signs those scores with the noise, the exact integer
    #    covariance here gives sign 0 ("Sign at zero covariance")
    a = np.array([1, 2, 3, 3, 2, 3, 3, 2, 2]); b = np.array([0, 4, 3, 2, 2, 0, 2, 4, 4])
    c = np.array([4, 1, 3, 2, 4, 4, 3, 0, 0]); d = np.array([0, 3, 2, 0, 4, 2, 0, 0, 0])
    e = np.array([1, 2, 1, 2, 0, 0, 0, 0, 0]); f = np.array([1, 1, 2, 2, 0, 0, 0, 0, 0])
    base = np.stack([a, b, c, d, e, f], axis=1)
    rng = np.random.RandomState(11)
    Xz = np.concatenate([np.tile(base, (30, 1)), rng.poisson(2, size=(270, 6))], axis=1).astype(np.float32)
    g = make_case(refm, "zero_cov_270x12_b10", Xz, 10)
    assert int(g["n_sign_ambiguous"]) > 0, "the constructed case lost its zero-covariance pairs"


if __name__ == "__main__":
    main()
