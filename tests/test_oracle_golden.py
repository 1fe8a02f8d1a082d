"""CPU tests: the oracle (numpy + C restatements) against the golden vectors that
oracle/gen_golden.py produced by running the reference's own functions."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden
from oracle import oracle as orc


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_discretize_oracles_match_reference(case):
    z = load_golden(case)
    want = z["x_disc"].astype(np.int32)
    xd_np, nb_np = orc.discretize_genes_np(z["X"], z["n_bins"])
    xd_c, nb_c, edges = orc.discretize_genes_c(z["X"], z["n_bins"], return_edges=True)
    bins_gm, nb_csr, edges_csr = orc.discretize_csr_c(z["X"], z["n_bins"])
    for xd, nb in ((xd_np, nb_np), (xd_c, nb_c), (bins_gm.T.astype(np.int32), nb_csr)):
        assert np.array_equal(xd, want)
        assert np.array_equal(nb, z["n_bins_per_gene"])
    assert np.array_equal(edges, z["edges"]) and np.array_equal(edges_csr, z["edges"])


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_mi_oracles_match_reference(case):
    z = load_golden(case)
    xd, nb = z["x_disc"].astype(np.int32), z["n_bins_per_gene"]
    raw_c, n = orc.mi_pairs_c(xd, nb)
    assert np.abs(raw_c - z["mi_raw"]).max() < 1e-12
    valid = int(((nb > 1)[:, None] & (nb > 1)[None, :])[np.triu_indices(len(nb), 1)].sum())
    assert n == valid
    raw_gm, _ = orc.mi_pairs_gm_c(np.ascontiguousarray(xd.T).astype(np.uint8), nb)
    assert np.array_equal(raw_gm, raw_c)
    assert np.array_equal(np.round(raw_c, 5), z["mi_unsigned_round5"])
    signed, _ = orc.mi_pairs_c(xd, nb, corr=z["corr_sign"].astype(np.float64), sign_mode=1)
    assert np.array_equal(np.round(signed, 5), z["mi_signed_round5"])


def test_numpy_tier_small():
    z = load_golden("ref_poisson_500x20_b5")
    got = orc.mi_matrix_np(z["X"], 5, signed=True)
    assert np.array_equal(np.round(got, 5), z["mi_signed_round5"])


def test_known_answers_survey_appendix_c():
    """SURVEY.md Appendix C values (produced from the reference's NumPy tier at survey time)."""
    z = load_golden("ref_poisson_500x50_b10")
    assert z["n_bins_per_gene"][:12].tolist() == [8, 9, 8, 9, 9, 9, 8, 8, 9, 9, 7, 10]
    assert int(z["n_bins_per_gene"].sum()) == 417
    J = z["joint_0_1"]
    assert J.shape == (8, 9) and J[0, 0] == 0 and int(J.sum()) == 492
    assert J[0].tolist() == [0, 0, 0, 17, 0, 18, 0, 11, 10]
    assert J[3].tolist() == [22, 0, 0, 36, 0, 37, 0, 22, 11]
    assert abs(z["mi_raw"][0, 1] - 0.04174775719939659) < 1e-15
    s = z["mi_signed_round5"]
    assert (s[0, 1], s[0, 2], s[10, 20], s[48, 49]) == (-0.04175, 0.03263, 0.04031, 0.04036)
    assert abs(s.sum() - -6.00824) < 1e-9 and int((s < 0).sum()) == 1290 and s.max() == 0.08053
    assert z["pearson"][0, 1] == -0.01513 and z["jaccard"][0, 1] == 0.75407 and z["cosine"][0, 1] == 0.66345


def test_joint_excludes_both_zero_cells():
    z = load_golden("synth_counts_700x96_b10")
    xd, nb = z["x_disc"].astype(np.int32), z["n_bins_per_gene"]
    J, cnt = orc.joint_c(xd, nb, 0, 1)
    assert J[0, 0] == 0 and cnt == int(((xd[:, 0] > 0) | (xd[:, 1] > 0)).sum())
    assert np.array_equal(J.astype(np.float64), orc.joint_np(xd[:, 0], xd[:, 1], nb[0], nb[1]))
    assert abs(orc.mi_from_joint_np(J.astype(np.float64)) - z["mi_raw"][0, 1]) < 1e-12


def test_integer_sign_matches_corrcoef_sign():
    for case in ("ref_poisson_500x50_b10", "synth_counts_700x96_b10"):
        z = load_golden(case)
        isign = orc.sign_int_c(z["X"])
        assert np.array_equal(isign, z["int_sign"])
        differ = isign != z["corr_sign"]
        assert np.all(isign[differ] == 0)      # only where the covariance is exactly zero


def test_quantile_fractions_match_numpy():
    q = orc.quantile_table()
    L = orc.lib()
    for k in range(2, 16):
        for t in range(1, k):
            assert L.gvo_quantile_fraction(k, t) == q[k, t - 1]


def test_matrix_target_oracles():
    z = load_golden("ref_poisson_500x50_b10")
    assert np.allclose(orc.round5_dict_matrix(orc.pearson_np(z["X"])), z["pearson"], atol=1e-12)
    assert np.allclose(orc.round5_dict_matrix(orc.cosine_np(z["X"])), z["cosine"], atol=1e-12)
    assert np.allclose(orc.round5_dict_matrix(orc.jaccard_np(z["X"])), z["jaccard"], atol=1e-12)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_gene_entropy_oracle_matches_reference(case):
    z = load_golden(case)
    assert np.array_equal(orc.gene_entropy_np(z["X"]), z["gene_entropy"], equal_nan=True)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_spearman_and_graph_xcorr_oracles_match_reference(case):
    """scipy.stats.spearmanr as the reference calls it (metrics.py:423-431) and target_graph_xcorr
    (_graph_targets.py:10-55, aggr mean with / without include_self) against the reference's output;
    the integer rank rows the B200 path contracts reproduce spearmanr."""
    z = load_golden(case)
    if "spearman" not in z:
        pytest.skip("no matrix targets in this case")
    sp_np = orc.spearman_np(z["X"])
    assert np.allclose(orc.round5_dict_matrix(sp_np), z["spearman"], atol=1e-12, equal_nan=True)
    ri = orc.int_pearson_exact(orc.rank_rows_np(z["X"]))
    ok = ~np.isnan(sp_np)
    assert np.abs(ri[ok] - sp_np[ok]).max() < 1e-12 and np.array_equal(np.isnan(ri), np.isnan(sp_np))
    assert np.allclose(orc.round5_dict_matrix(orc.graph_xcorr_np(z["X"], z["graph"])), z["graph_xcorr"], atol=1e-12)
    assert np.allclose(orc.round5_dict_matrix(orc.graph_xcorr_np(z["X"], z["graph"], True)),
                       z["graph_xcorr_self"], atol=1e-12)
