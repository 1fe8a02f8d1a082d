"""GPU parity tests of the value-row (co-moment) paths beyond plain small counts:

  * count digits with three INT8 rows per gene (counts up to 2^21 - 1), exact;
  * fixed-point rows for data that are not counts (log-normalised floats): the rows are bit-exact
    against the checker's restatement of the quantisation, the sign map is the exact sign of the
    integer covariance, and Pearson / cosine / signed MI agree with the reference's float64 results
    (golden vectors written by the reference itself) inside the stated tolerances;
  * rank rows and the Spearman target (genevector/metrics.py:423-431) against scipy.stats.spearmanr
    as called by the reference.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_golden

pytestmark = pytest.mark.gpu

MI_TOL = 1e-6
ROUND_TOL = 1.1e-5
F32_TOL = 2e-6         # float32 output of a correlation in [-1, 1] + fixed-point quantisation


def lognorm(n_cells, n_genes, seed, dtype=np.float32):
    from genevector_b200.synth import synth_counts
    Xc = synth_counts(n_cells, n_genes, seed=seed).toarray().astype(np.float64)
    tot = Xc.sum(axis=1, keepdims=True) + 1.0
    return sp.csr_matrix(np.log1p(Xc / tot * 1e4).astype(dtype))


# ------------------------------------------------------------------ three-digit counts
@pytest.mark.parametrize("cg", [1, 2])
def test_three_digit_counts_exact(engine, cg):
    from genevector_b200 import _lib
    from genevector_b200.synth import synth_counts
    from oracle import oracle as orc
    X = synth_counts(700, 210, seed=41)
    rng = np.random.default_rng(41)
    X.data = (X.data * rng.integers(1, 60000, size=X.data.shape)).astype(np.float32)
    assert 16384 <= X.data.max() < (1 << 21)
    csr = engine.upload_csr(X)
    blk = engine.discretize(csr, 10, val_mode=0)
    assert (blk.val_mode, blk.n_limbs, blk.limb_bits) == (0, 3, 7)
    V = np.asarray(X.todense()).astype(np.int64).T
    assert np.array_equal(blk.gene_values(), V)
    assert np.array_equal(blk.gsum.cpu().numpy(), V.sum(axis=1))
    assert np.array_equal(blk.gsq.cpu().numpy(), (V * V).sum(axis=1))
    old = engine.cta_group
    engine.cta_group = cg
    try:
        sgn = engine.sign_matrix(blk).cpu().numpy()[:, :210]
        p = engine.moment_scores(blk, _lib.GV_MOM_PEARSON).cpu().numpy().astype(np.float64)
        c = engine.moment_scores(blk, _lib.GV_MOM_COSINE).cpu().numpy().astype(np.float64)
    finally:
        engine.cta_group = old
    assert np.array_equal(sgn, orc.int_sign_exact(V))
    want = orc.pearson_np(X); np.fill_diagonal(want, 0.0)
    assert np.abs(p - want).max() <= F32_TOL
    assert np.abs(c - orc.cosine_np(X)).max() <= F32_TOL


# ------------------------------------------------------------------ fixed-point rows (non-count data)
@pytest.mark.parametrize("case", ["lognorm_f32_300x40_b7", "lognorm_f64_300x40_b10"])
def test_fixed_point_rows_and_targets_vs_reference(engine, case):
    from genevector_b200 import _lib
    from genevector_b200.metrics import target_pearson, target_cosine, compute_mi_b200
    from oracle import oracle as orc
    z = load_golden(case)
    X = z["X"]
    G = X.shape[1]
    genes = [f"GENE{i}" for i in range(G)]
    csr = engine.upload_csr(X)
    blk = engine.discretize(csr, z["n_bins"], val_mode=0)       # auto: not counts -> fixed point
    assert (blk.val_mode, blk.n_limbs, blk.limb_bits) == (2, 3, 7)
    assert blk.path_used == _lib.GV_PATH_GENERAL
    assert np.array_equal(blk.x_disc(), z["x_disc"].astype(np.int32))       # bins unaffected
    V = orc.fixed_point_rows_np(X, 21)
    assert np.array_equal(blk.gene_values(), V)
    assert np.array_equal(blk.gsum.cpu().numpy(), V.sum(axis=1))
    assert np.array_equal(blk.gsq.cpu().numpy(), (V * V).sum(axis=1))
    # the sign map is the exact sign of the integer covariance of the fixed-point values ...
    sgn = engine.sign_matrix(blk).cpu().numpy()[:, :G]
    assert np.array_equal(sgn, orc.int_sign_exact(V))
    # ... which is the reference's np.sign(corrcoef) on these data
    assert np.array_equal(sgn, z["corr_sign"])
    # Pearson / cosine against the reference's own outputs
    p = target_pearson(X, genes)
    assert np.allclose(p.rounded(), z["pearson"], atol=ROUND_TOL, equal_nan=True)
    assert np.array_equal(np.isnan(p.rounded()), np.isnan(z["pearson"]))
    raw = orc.pearson_np(X); np.fill_diagonal(raw, 0.0)
    ok = ~np.isnan(raw)
    assert np.abs(p.matrix.astype(np.float64)[ok] - raw[ok]).max() <= F32_TOL
    c = target_cosine(X, genes)
    assert np.abs(c.rounded() - z["cosine"]).max() <= ROUND_TOL
    # signed MI on non-count data
    sm = compute_mi_b200(X, genes, n_bins=z["n_bins"], signed=True)
    assert np.abs(sm.matrix.astype(np.float64) - z["mi_raw"] * z["corr_sign"]).max() <= MI_TOL
    assert np.abs(sm.rounded() - z["mi_signed_round5"]).max() <= ROUND_TOL


@pytest.mark.parametrize("cg", [1, 2])
def test_fixed_point_several_tiles(engine, cg):
    """Three digits per gene: 80 genes per 256-row tile side, so 300 genes span 4 x 4 tiles."""
    from genevector_b200 import _lib
    from oracle import oracle as orc
    X = lognorm(1500, 300, seed=7)
    csr = engine.upload_csr(X)
    blk = engine.discretize(csr, 10, val_mode=0)
    assert (blk.val_mode, blk.n_limbs) == (2, 3)
    V = orc.fixed_point_rows_np(X, 21)
    assert np.array_equal(blk.gene_values(), V)
    old = engine.cta_group
    engine.cta_group = cg
    try:
        sgn = engine.sign_matrix(blk).cpu().numpy()[:, :300]
        p = engine.moment_scores(blk, _lib.GV_MOM_PEARSON).cpu().numpy().astype(np.float64)
    finally:
        engine.cta_group = old
    assert np.array_equal(sgn, orc.int_sign_exact(V))
    want = orc.pearson_np(X); np.fill_diagonal(want, 0.0)
    ok = ~np.isnan(want)
    assert np.abs(p[ok] - want[ok]).max() <= F32_TOL
    # signed MI end to end vs the C oracle with the reference's float64 sign
    bins_gm, nb, _ = orc.discretize_csr_c(X, 10)
    corr = orc.pearson_sign_np(X)
    fs = np.sign(corr).astype(np.int8); np.fill_diagonal(fs, 0)
    assert np.array_equal(sgn, fs)
    want_mi, _ = orc.mi_pairs_gm_c(bins_gm, nb, sign=fs)
    out = engine.mi_from_csr(csr, 1500, 300, signed=True).cpu().numpy()
    assert np.abs(out.astype(np.float64) - want_mi).max() <= MI_TOL


def test_fixed_point_sharded_ranks_cover_the_matrix(engine):
    """Emulated 3-rank run on non-count data: each rank computes only the sign tiles its MI tiles
    read (three-digit value rows: different genes-per-tile than the one-hot operand)."""
    import torch
    X = lognorm(1200, 700, seed=9)
    csr = engine.upload_csr(X)
    full = engine.mi_from_csr(csr, 1200, 700, signed=True)
    acc = torch.zeros_like(full)
    for rank in range(3):
        acc += engine.mi_from_csr(csr, 1200, 700, signed=True, rank=rank, world=3)
    assert torch.equal(acc, full)
    assert int((full < 0).sum()) > 0


# ------------------------------------------------------------------ Spearman
@pytest.mark.parametrize("case", ["ref_poisson_500x50_b10", "synth_counts_700x96_b10",
                                  "synth_counts_400x30_b15", "lognorm_f32_300x40_b7",
                                  "lognorm_f64_300x40_b10"])
def test_spearman_matches_reference(engine, case):
    from genevector_b200 import _lib
    from genevector_b200.metrics import target_spearman, get_target_function
    from oracle import oracle as orc
    z = load_golden(case)
    X = z["X"]
    G = X.shape[1]
    genes = [f"GENE{i}" for i in range(G)]
    csr = engine.upload_csr(X)
    # rank rows: both discretisation strategies give the same doubled average ranks
    R = orc.rank_rows_np(X)
    is_counts = bool(np.array_equal(X.data, np.rint(X.data)) and X.data.max() < 256 and X.data.min() >= 0)
    for path in ([_lib.GV_PATH_COUNTS] if is_counts else []) + [_lib.GV_PATH_GENERAL]:
        blk = engine.discretize(csr, 10, val_mode=3, path=path)
        assert blk.path_used == path and blk.val_mode == 3
        assert np.array_equal(blk.gene_values(), R), (case, path)
        assert np.array_equal(blk.gsum.cpu().numpy(), R.sum(axis=1))
        assert np.array_equal(blk.gsq.cpu().numpy(), (R * R).sum(axis=1))
    s = target_spearman(X, genes)
    assert get_target_function("spearman") is target_spearman
    assert np.allclose(s.rounded(), z["spearman"], atol=ROUND_TOL, equal_nan=True)
    assert np.array_equal(np.isnan(s.rounded()), np.isnan(z["spearman"]))
    raw = orc.spearman_np(X); np.fill_diagonal(raw, 0.0)
    ok = ~np.isnan(raw)
    assert np.abs(s.matrix.astype(np.float64)[ok] - raw[ok]).max() <= 1e-7    # float32 output only


def test_spearman_seeded_larger(engine):
    from genevector_b200.metrics import target_spearman
    from oracle import oracle as orc
    for X in (lognorm(2100, 170, seed=12), None):
        if X is None:
            from genevector_b200.synth import synth_counts
            X = synth_counts(9000, 90, seed=13)         # 2 * 9000 needs three 7-bit digits
        genes = [f"G{i}" for i in range(X.shape[1])]
        s = target_spearman(X, genes).matrix.astype(np.float64)
        raw = orc.spearman_np(X); np.fill_diagonal(raw, 0.0)
        ok = ~np.isnan(raw)
        assert np.abs(s[ok] - raw[ok]).max() <= 1e-7
        assert np.array_equal(np.isnan(s), np.isnan(raw))


# ------------------------------------------------------------------ graph cross-correlation
@pytest.mark.parametrize("case", ["ref_poisson_500x50_b10", "synth_counts_700x96_b10",
                                  "synth_counts_400x30_b15", "lognorm_f32_300x40_b7",
                                  "lognorm_f64_300x40_b10"])
def test_graph_xcorr_matches_reference(engine, case):
    """target_graph_xcorr (genevector/_graph_targets.py:10-55, aggr="mean" with and without
    include_self) against the reference's own output on a seeded neighbour graph with weights and
    isolated cells."""
    from genevector_b200.metrics import target_graph_xcorr, get_target_function
    from oracle import oracle as orc
    z = load_golden(case)
    X, W = z["X"], z["graph"]
    genes = [f"GENE{i}" for i in range(X.shape[1])]
    assert get_target_function("graph_xcorr") is target_graph_xcorr
    for key, params in (("graph_xcorr", None), ("graph_xcorr_self", {"include_self": True})):
        s = target_graph_xcorr(X, genes, graph=W, aggr_params=params)
        assert np.abs(s.rounded() - z[key]).max() <= ROUND_TOL, (case, key)
        raw = orc.graph_xcorr_np(X, W, include_self=bool(params))
        assert np.abs(s.matrix.astype(np.float64) - raw).max() <= F32_TOL, (case, key)
        assert np.array_equal(s.matrix, s.matrix.T) and not np.diag(s.matrix).any()


def test_graph_xcorr_reference_test_cases(engine):
    """The structural checks of the reference's tests/test_graph_targets.py on non-negative data."""
    from genevector_b200.metrics import target_graph_xcorr
    from oracle import oracle as orc
    adj = sp.csr_matrix(np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]], dtype=np.float64))
    X = np.array([[1, 0], [0, 2], [3, 1]], dtype=np.float64)
    for Xin in (X, sp.csr_matrix(X)):
        s = target_graph_xcorr(Xin, ["g0", "g1"], graph=adj)
        assert "g0" in s and "g1" in s and "g0" not in s["g0"]
        assert s["g0"]["g1"] == s["g1"]["g0"]
        assert abs(s["g0"]["g1"] - round(float(orc.graph_xcorr_np(X, adj)[0, 1]), 5)) <= ROUND_TOL
    # chain graph: A high in even cells, B in odd cells (the neighbours) -> |xcorr(A,B)| is large
    n = 50
    row = list(range(n - 1)) + list(range(1, n))
    col = list(range(1, n)) + list(range(n - 1))
    chain = sp.csr_matrix(([1.0] * len(row), (row, col)), shape=(n, n))
    rng = np.random.default_rng(0)
    Xc = np.column_stack([[5.0 if i % 2 == 0 else 0.0 for i in range(n)],
                          [5.0 if i % 2 == 1 else 0.0 for i in range(n)], rng.random(n)])
    s = target_graph_xcorr(Xc, ["A", "B", "C"], graph=chain)
    assert abs(s["A"]["B"]) > abs(s["A"]["C"])
    assert np.abs(s.matrix.astype(np.float64) - orc.graph_xcorr_np(Xc, chain)).max() <= F32_TOL
    # values stay in [-1, 1]
    Xr = rng.random((60, 12))
    Wr = sp.csr_matrix(rng.integers(0, 2, size=(60, 60)).astype(np.float64))
    s = target_graph_xcorr(Xr, [f"g{i}" for i in range(12)], graph=Wr)
    assert np.abs(s.matrix).max() <= 1.0 + 1e-6
    assert np.abs(s.matrix.astype(np.float64) - orc.graph_xcorr_np(Xr, Wr)).max() <= F32_TOL


@pytest.mark.parametrize("cg", [1, 2])
def test_graph_xcorr_several_tiles(engine, cg):
    """350 genes: 2 x 5 tiles of the rectangle (256 count rows per A tile, 80 genes per B tile)."""
    from genevector_b200.metrics import target_graph_xcorr
    from genevector_b200.synth import synth_counts
    from oracle import oracle as orc
    X = synth_counts(1800, 350, seed=17)
    W = orc.seeded_graph(1800, seed=5)
    old = engine.cta_group
    engine.cta_group = cg
    try:
        s = target_graph_xcorr(X, [f"g{i}" for i in range(350)], graph=W)
    finally:
        engine.cta_group = old
    assert np.abs(s.matrix.astype(np.float64) - orc.graph_xcorr_np(X, W)).max() <= F32_TOL


# ------------------------------------------------------------------ data with negative values
def _signed_sparse(n_cells, n_genes, seed, dtype=np.float32):
    """Sparse matrix with normal-distributed stored values (negative, positive, explicit zeros)."""
    rng = np.random.default_rng(seed)
    dense = rng.standard_normal((n_cells, n_genes)) * (rng.random((n_cells, n_genes)) < 0.3)
    dense[:, 3] = np.abs(dense[:, 3])           # a gene without negative values
    dense[:, 4] = -np.abs(dense[:, 4])          # a gene without positive values
    dense[:, 5] = 0.0                           # an empty gene
    X = sp.csr_matrix(dense.astype(dtype))
    X.data[7] = 0.0                             # an explicit stored zero
    return X


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_negative_values_signed_fixed_point(engine, dtype):
    """Centred / scaled data: bins treat x <= 0 as not expressed (metrics.py:53) while Pearson, its
    sign, cosine and signed MI use the raw values (shifted fixed-point rows, offsets undone for
    cosine)."""
    from genevector_b200 import _lib
    from genevector_b200.engine import GvError
    from genevector_b200.metrics import target_pearson, target_cosine, target_spearman, compute_mi_b200
    from genevector_b200.data import get_gene_entropy
    from oracle import oracle as orc
    X = _signed_sparse(900, 120, seed=3, dtype=dtype)
    G = X.shape[1]
    genes = [f"g{i}" for i in range(G)]
    csr = engine.upload_csr(X)
    blk = engine.discretize(csr, 10, val_mode=0)
    assert (blk.val_mode, blk.n_limbs) == (2, 3) and blk.data_flags[0] & 1
    xd, nb = orc.discretize_genes_np(X, 10)
    assert np.array_equal(blk.x_disc(), xd) and np.array_equal(blk.n_bins_per_gene.cpu().numpy(), nb)
    V = orc.fixed_point_rows_np(X, 21)
    assert np.array_equal(blk.gene_values(), V)
    assert np.array_equal(blk.gsum.cpu().numpy(), V.sum(axis=1))
    assert np.array_equal(blk.gsq.cpu().numpy(), (V * V).sum(axis=1))
    sgn = engine.sign_matrix(blk).cpu().numpy()[:, :G]
    assert np.array_equal(sgn, orc.int_sign_exact(V))
    ref_sign = np.sign(orc.pearson_sign_np(X)).astype(np.int8); np.fill_diagonal(ref_sign, 0)
    assert np.array_equal(sgn, ref_sign)
    p = target_pearson(X, genes).matrix.astype(np.float64)
    want = orc.pearson_np(X); np.fill_diagonal(want, 0.0)
    ok = ~np.isnan(want)
    assert np.abs(p[ok] - want[ok]).max() <= F32_TOL and np.array_equal(np.isnan(p), np.isnan(want))
    c = target_cosine(X, genes).matrix.astype(np.float64)
    assert np.abs(c - orc.cosine_np(X)).max() <= F32_TOL
    want_mi, _ = orc.mi_pairs_c(xd, nb, corr=ref_sign.astype(np.float64), sign_mode=1)
    got_mi = compute_mi_b200(X, genes, signed=True).matrix.astype(np.float64)
    assert np.abs(got_mi - want_mi).max() <= MI_TOL
    # ranks over every stored entry (negatives, explicit zeros, positives) and the zero cells
    blk3 = engine.discretize(csr, 10, val_mode=3)
    R = orc.rank_rows_np(X)
    assert blk3.val_mode == 3 and np.array_equal(blk3.gene_values(), R)
    assert np.array_equal(blk3.gsum.cpu().numpy(), R.sum(axis=1))
    assert np.array_equal(blk3.gsq.cpu().numpy(), (R * R).sum(axis=1))
    assert np.array_equal(blk3.x_disc(), xd)
    sp_got = target_spearman(X, genes).matrix.astype(np.float64)
    sp_want = orc.spearman_np(X); np.fill_diagonal(sp_want, 0.0)
    oks = ~np.isnan(sp_want)
    assert np.abs(sp_got[oks] - sp_want[oks]).max() <= 1e-7 and np.array_equal(np.isnan(sp_got), np.isnan(sp_want))
    # unique-value entropy (data.py:186-203): the dropped first value is now the smallest negative one
    ent = get_gene_entropy(X, genes)
    assert np.abs(np.array([ent[g] for g in genes]) - orc.gene_entropy_np(X)).max() < 1e-12
    # count digits cannot hold negative values
    with pytest.raises(GvError, match="non-negative"):
        engine.discretize(csr, 10, val_mode=0, n_limbs=1)


def test_graph_xcorr_standard_normal_like_the_reference_tests(engine):
    """tests/test_graph_targets.py:52-77 of the reference draws X from a standard normal."""
    from genevector_b200.metrics import target_graph_xcorr
    from oracle import oracle as orc
    rng = np.random.default_rng(42)
    for n, d in ((50, 10), (300, 90)):
        X = rng.standard_normal((n, d))
        adj = sp.csr_matrix(rng.integers(0, 2, size=(n, n)).astype(np.float64))
        genes = [f"g{i}" for i in range(d)]
        s = target_graph_xcorr(X, genes, graph=adj)
        assert np.abs(s.matrix).max() <= 1.0 + 1e-6
        assert np.array_equal(s.matrix, s.matrix.T)
        assert np.abs(s.matrix.astype(np.float64) - orc.graph_xcorr_np(X, adj)).max() <= F32_TOL
        s2 = target_graph_xcorr(X, genes, graph=adj, aggr_params={"include_self": True})
        assert np.abs(s2.matrix.astype(np.float64) - orc.graph_xcorr_np(X, adj, True)).max() <= F32_TOL


def test_nan_values_are_not_silently_turned_into_zeros(engine):
    """NaN > 0 is False, so the reference bins NaN as "not expressed" (metrics.py:53) -- reproduced --
    while its np.corrcoef turns the whole gene into NaN; the co-moment paths refuse such input loudly
    instead of treating NaN as 0."""
    from genevector_b200.engine import GvError
    from genevector_b200.metrics import compute_mi_b200, target_pearson
    from genevector_b200.synth import synth_counts, gene_names
    from oracle import oracle as orc
    X = synth_counts(400, 30, seed=4)
    X.data[11] = np.nan
    genes = gene_names(30)
    csr = engine.upload_csr(X)
    blk = engine.discretize(csr, 10)
    with np.errstate(invalid="ignore"):
        xd, nb = orc.discretize_genes_np(X, 10)
    assert np.array_equal(blk.x_disc(), xd) and np.array_equal(blk.n_bins_per_gene.cpu().numpy(), nb)
    assert blk.data_flags[0] & 8
    got = compute_mi_b200(X, genes, signed=False).matrix.astype(np.float64)
    assert np.abs(got - orc.mi_pairs_c(xd, nb)[0]).max() <= MI_TOL
    for fn in (lambda: compute_mi_b200(X, genes, signed=True), lambda: target_pearson(X, genes)):
        with pytest.raises(GvError, match="NaN"):
            fn()
