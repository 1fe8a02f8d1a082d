"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden
vectors generated from the reference.  Bars: bins / n_bins_per_gene /
joint counts bit-exact; raw MI within 1e-6 absolute (fp32 epilogue vs the reference's fp64);
rounded scores within 1e-5 (+1e-6: a 1e-7 difference can flip the fifth decimal); Pearson sign
equal wherever the covariance is not exactly zero."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN_CASES, load_golden

pytestmark = pytest.mark.gpu

MI_TOL = 1e-6          # north_star: MI / signed MI within 1e-6 absolute
ROUND_TOL = 1.1e-5     # 5-dp rounding of two values 1e-6 apart can differ by one unit


def onehot_row(g, a, B):
    gpb = 32 // B
    return (g // gpb) * 32 + (g % gpb) * B + (a - 1)


# ------------------------------------------------------------------ discretisation
@pytest.mark.parametrize("path", ["auto", "general"])
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_discretize_bit_exact_vs_reference(engine, case, path):
    """Both discretisation strategies (count-data histogram/LUT, general radix select) give the
    reference's bins, n_bins_per_gene, edges and marginals bit for bit."""
    from genevector_b200 import _lib
    z = load_golden(case)
    csr = engine.upload_csr(z["X"])
    blk = engine.discretize(csr, z["n_bins"], path=_lib.GV_PATH_AUTO if path == "auto" else _lib.GV_PATH_GENERAL)
    is_counts = bool(np.array_equal(z["X"].data, np.rint(z["X"].data)) and z["X"].data.max() < 256)
    want_path = _lib.GV_PATH_COUNTS if (path == "auto" and is_counts) else _lib.GV_PATH_GENERAL
    assert blk.path_used == want_path
    got = blk.x_disc()
    want = z["x_disc"].astype(np.int32)
    bad = np.argwhere(got != want)
    assert bad.size == 0, f"{len(bad)} bins differ, first {bad[:5].tolist()}"
    assert np.array_equal(blk.n_bins_per_gene.cpu().numpy(), z["n_bins_per_gene"])
    # edges: float64, identical bits (unused slots differ: oracle uses +inf)
    e_got, e_want = blk.edges.cpu().numpy(), z["edges"]
    for g, nb in enumerate(z["n_bins_per_gene"]):
        ne = max(int(nb) - 2, 0)
        assert np.array_equal(e_got[g, :ne], e_want[g, :ne]), (g, e_got[g, :ne], e_want[g, :ne])
    # marginals: bin counts and the non-zero count
    marg = blk.marg.cpu().numpy()
    for a in range(1, 32):
        assert np.array_equal(marg[:, a], (want == a).sum(axis=0)), a
    assert np.array_equal(marg[:, 0], (want > 0).sum(axis=0))


def test_discretize_matches_reference_api(engine):
    """Same checks as the reference's tests/test_metrics.py:23-31 through the mirrored API."""
    from genevector_b200.metrics import discretize_genes
    z = load_golden("ref_poisson_500x50_b10")
    X_disc, n_bins = discretize_genes(z["X"], n_bins=10)
    assert X_disc.shape == z["X"].shape and len(n_bins) == z["X"].shape[1]
    assert X_disc.dtype == np.int32
    assert np.all(X_disc[np.asarray(z["X"].todense()) == 0] == 0)


# ------------------------------------------------------------------ joint counts
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("case", ["ref_poisson_500x50_b10", "ref_poisson_500x20_b5",
                                  "synth_counts_700x96_b10", "synth_counts_400x30_b15",
                                  "synth_counts_400x30_b16", "poisson20_500x24_b31", "lognorm_f32_300x40_b20"])
def test_joint_counts_bit_exact(engine, case, cg):
    """Raw tcgen05 accumulators == the reference's joint histogram (non-zero bins) for every
    gene pair, and the zero-bin row/column + total follow from the marginals exactly."""
    from oracle import oracle as orc
    z = load_golden(case)
    csr = engine.upload_csr(z["X"])
    blk = engine.discretize(csr, z["n_bins"])
    onehot, B = engine.expand_onehot(blk, "int8")
    C = engine.gram_dump(onehot, cta_group=cg).cpu().numpy()
    if (32 // B) * B == 30:
        # the production MI kernel stages only the 30 data rows of each block on the B side
        # (MMA N = 240): same counts wherever a data row meets a data column
        C30 = engine.gram_dump(onehot, cta_group=cg, active_rows=30).cpu().numpy()
        data = (np.arange(C.shape[0]) % 32) < 30
        assert np.array_equal(C30[:, data], C[:, data])
        assert not C30[:, ~data].any()
        C = C30
    x_disc = z["x_disc"].astype(np.int32)
    nb = z["n_bins_per_gene"]
    marg = blk.marg.cpu().numpy()
    G = x_disc.shape[1]
    rng = np.random.default_rng(0)
    pairs = [(i, j) for i in range(G) for j in range(i + 1, G)]
    if len(pairs) > 600:
        pairs = [pairs[k] for k in rng.choice(len(pairs), 600, replace=False)]
    for i, j in pairs:
        if nb[i] <= 1 or nb[j] <= 1:
            continue
        J, count = orc.joint_c(x_disc, nb, i, j)
        rows = [onehot_row(i, a, B) for a in range(1, nb[i])]
        cols = [onehot_row(j, b, B) for b in range(1, nb[j])]
        Gij = C[np.ix_(rows, cols)]
        assert np.array_equal(Gij, J[1:, 1:].astype(np.int64)), (i, j)
        S = int(Gij.sum())
        assert np.array_equal(marg[i, 1:nb[i]] - Gij.sum(axis=1), J[1:, 0]), (i, j)
        assert np.array_equal(marg[j, 1:nb[j]] - Gij.sum(axis=0), J[0, 1:]), (i, j)
        assert marg[i, 0] + marg[j, 0] - S == count and J[0, 0] == 0, (i, j)
    for name in z:
        if name.startswith("joint_"):
            _, i, j = name.split("_")
            i, j = int(i), int(j)
            rows = [onehot_row(i, a, B) for a in range(1, nb[i])]
            cols = [onehot_row(j, b, B) for b in range(1, nb[j])]
            assert np.array_equal(C[np.ix_(rows, cols)], z[name][1:, 1:].astype(np.int64))


# ------------------------------------------------------------------ MI
@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_mi_matches_reference(engine, case, cg):
    z = load_golden(case)
    csr = engine.upload_csr(z["X"])
    blk = engine.discretize(csr, z["n_bins"])
    onehot, B = engine.expand_onehot(blk)
    old = engine.cta_group
    engine.cta_group = cg
    try:
        out = engine.mi_scores(blk, onehot, B).cpu().numpy().astype(np.float64)
    finally:
        engine.cta_group = old
    err = np.abs(out - z["mi_raw"])
    assert err.max() <= MI_TOL, f"max |MI - ref| = {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"
    assert np.array_equal(out, out.T) and np.all(np.diag(out) == 0)
    assert np.abs(np.round(out, 5) - z["mi_unsigned_round5"]).max() <= ROUND_TOL


@pytest.mark.parametrize("case", ["ref_poisson_500x50_b10", "ref_poisson_500x20_b5",
                                  "synth_counts_700x96_b10", "synth_counts_400x30_b15",
                                  "synth_counts_400x30_b16", "poisson20_500x24_b31"])
def test_signed_mi_matches_reference(engine, case):
    from genevector_b200.metrics import compute_mi_b200
    z = load_golden(case)
    genes = [f"GENE{i}" for i in range(z["X"].shape[1])]
    sm = compute_mi_b200(z["X"], genes, n_bins=z["n_bins"], signed=True)
    got = sm.matrix.astype(np.float64)
    want = z["mi_raw"] * z["corr_sign"]
    amb = z["int_sign"] != z["corr_sign"]          # exactly-zero covariance: fp64 sign is noise
    err = np.abs(got - want)
    err[amb] = 0.0     # the reference's sign is rounding noise there; ours is exactly 0 by rule
    assert err.max() <= MI_TOL, f"max err {err.max():.3e}"
    assert np.abs(sm.rounded() - z["mi_signed_round5"])[~amb].max() <= ROUND_TOL
    # mapping semantics the reference's consumers rely on (data.py:383-388)
    assert genes[0] in sm and genes[1] in sm[genes[0]] and genes[0] not in sm[genes[0]]
    assert abs(sm[genes[0]][genes[1]] - z["mi_signed_round5"][0, 1]) <= ROUND_TOL


def test_sign_at_exactly_zero_covariance(engine):
    """Constructed golden case: three gene pairs whose covariance is exactly zero while the reference's
    float64 np.corrcoef returns rounding noise (-4e-17 ...) and signs the score with it.  The exact integer
    covariance here gives sign 0, so the default rule (np.sign semantics) scores those pairs 0 -- the
    documented deviation -- and |MI| is what BASELINE.md section 5 compares there: the
    "rust" rule (+1 unless negative) returns it.  Every other pair equals the reference."""
    from genevector_b200.metrics import compute_mi_b200
    z = load_golden("zero_cov_270x12_b10")
    genes = [f"GENE{i}" for i in range(12)]
    amb = (z["int_sign"] == 0) & (z["corr_sign"] != 0)
    assert int(amb.sum()) == 6 and set(map(tuple, np.argwhere(np.triu(amb)))) == {(0, 1), (0, 4), (1, 5)}
    assert (z["mi_raw"][amb] > 0.4).all()
    dflt = compute_mi_b200(z["X"], genes, signed=True).matrix.astype(np.float64)
    assert not dflt[amb].any()
    rust = compute_mi_b200(z["X"], genes, signed=True, sign_rule="rust").matrix.astype(np.float64)
    assert np.abs(rust[amb] - z["mi_raw"][amb]).max() <= MI_TOL
    want = z["corr_sign"] * z["mi_raw"]
    assert np.abs(dflt - want)[~amb].max() <= MI_TOL
    blk = engine.discretize(engine.upload_csr(z["X"]), 10, val_mode=0)
    assert np.array_equal(engine.sign_matrix(blk).cpu().numpy()[:, :12], z["int_sign"])


def test_sign_matrix_exact(engine):
    z = load_golden("synth_counts_700x96_b10")
    csr = engine.upload_csr(z["X"])
    blk = engine.discretize(csr, 10, val_mode=0)
    sgn = engine.sign_matrix(blk).cpu().numpy()[:, :96]
    assert np.array_equal(sgn, z["int_sign"])


# ------------------------------------------------------------------ matrix targets
@pytest.mark.parametrize("case", ["ref_poisson_500x50_b10", "synth_counts_700x96_b10"])
def test_matrix_targets_match_reference(engine, case):
    from genevector_b200.metrics import target_pearson, target_cosine, target_jaccard
    from oracle import oracle as orc
    z = load_golden(case)
    genes = [f"GENE{i}" for i in range(z["X"].shape[1])]
    p = target_pearson(z["X"], genes)
    assert np.allclose(p.rounded(), z["pearson"], atol=ROUND_TOL, equal_nan=True)
    assert np.array_equal(np.isnan(p.rounded()), np.isnan(z["pearson"]))
    c = target_cosine(z["X"], genes)
    assert np.abs(c.rounded() - z["cosine"]).max() <= ROUND_TOL
    j = target_jaccard(z["X"], genes)
    assert np.abs(j.rounded() - z["jaccard"]).max() <= ROUND_TOL
    # jaccard is a float32 quotient of exact integers on both sides: bit-exact before rounding
    jr = orc.jaccard_np(z["X"]).astype(np.float32)
    assert np.array_equal(j.matrix, jr)


# ------------------------------------------------------------------ larger seeded cases vs the C oracle
@pytest.mark.parametrize("n_cells,n_genes,n_bins,seed", [
    (1000, 130, 10, 1),      # K not a multiple of 128, several tiles
    (3000, 333, 10, 2),      # > 1 tile row, ragged last block
    (2500, 64, 5, 3),        # 6 genes per 32-row block
    (777, 50, 8, 4),
    (129, 7, 10, 5),         # tiny
])
def test_mi_seeded_vs_oracle(engine, n_cells, n_genes, n_bins, seed):
    from oracle import oracle as orc
    from genevector_b200.synth import synth_counts
    X = synth_counts(n_cells, n_genes, seed=seed)
    bins_gm, nb, _ = orc.discretize_csr_c(X, n_bins)
    want, _ = orc.mi_pairs_gm_c(bins_gm, nb, sign=orc.sign_int_c(X))
    csr = engine.upload_csr(X)
    out = engine.mi_from_csr(csr, n_cells, n_genes, n_bins=n_bins, signed=True).cpu().numpy()
    from genevector_b200 import _lib
    for path in (_lib.GV_PATH_COUNTS, _lib.GV_PATH_GENERAL):
        blk = engine.discretize(csr, n_bins, val_mode=0, path=path)
        assert blk.path_used == path
        assert np.array_equal(blk.x_disc(), bins_gm.T.astype(np.int32))
        assert np.array_equal(blk.n_bins_per_gene.cpu().numpy(), nb)
        Xd = np.asarray(X.todense())
        assert np.array_equal(blk.gsum.cpu().numpy(), Xd.sum(axis=0).astype(np.int64))
        assert np.array_equal(blk.gsq.cpu().numpy(), (Xd.astype(np.int64) ** 2).sum(axis=0))
        assert np.array_equal(blk.val_rows.cpu().numpy()[:n_genes, :n_cells], Xd.T.astype(np.int8))
    err = np.abs(out.astype(np.float64) - want)
    assert err.max() <= MI_TOL, f"max err {err.max():.3e}"


def test_edge_cases(engine):
    """Empty matrix, one gene, all-zero genes, a constant gene, explicit zeros and negatives."""
    from oracle import oracle as orc
    # explicit zero and negative entries are 'not expressed' (metrics.py:53)
    X = sp.csr_matrix(np.array([[0, 2, -1, 5], [1, 0, 3, 5], [2, 2, 0, 5], [0, 0, 0, 5], [4, 1, -2, 5]],
                               dtype=np.float32))
    X.data[0] = 0.0  # stored zero
    xd, nb = orc.discretize_genes_np(X, 10)
    csr = engine.upload_csr(X)
    blk = engine.discretize(csr, 10)
    assert np.array_equal(blk.x_disc(), xd) and np.array_equal(blk.n_bins_per_gene.cpu().numpy(), nb)
    want = orc.mi_matrix_np(X, 10)
    out = engine.mi_from_csr(csr, 5, 4, n_bins=10).cpu().numpy()
    assert np.abs(out - want).max() <= MI_TOL
    # all-zero matrix: every gene has one bin, every pair is skipped (metrics.py:119-120)
    Z = sp.csr_matrix((40, 9), dtype=np.float32)
    csr = engine.upload_csr(Z)
    blk = engine.discretize(csr, 10)
    assert np.all(blk.n_bins_per_gene.cpu().numpy() == 1) and not blk.x_disc().any()
    assert not engine.mi_from_csr(csr, 40, 9).cpu().numpy().any()
    # a single gene: no pairs
    one = sp.csr_matrix(np.arange(10, dtype=np.float32).reshape(10, 1))
    csr = engine.upload_csr(one)
    assert engine.mi_from_csr(csr, 10, 1).cpu().numpy().shape == (1, 1)


def test_backend_dispatch_errors():
    """Unknown backends raise like the reference (metrics.py:410-411); no CPU tiers exist."""
    from genevector_b200.metrics import target_mi, get_target_function
    z = load_golden("ref_poisson_500x20_b5")
    genes = [f"GENE{i}" for i in range(20)]
    for bad in ("numpy", "numba", "rust", "gpu", "nope"):
        with pytest.raises(ValueError, match="Unknown MI backend"):
            target_mi(z["X"], genes, backend=bad)
    with pytest.raises(ValueError, match="Unknown target"):
        get_target_function("spearman_typo")
    sm = get_target_function("mi")(z["X"], genes, signed=False, backend="b200", device="cuda", n_bins=5)
    assert np.abs(sm.rounded() - z["mi_unsigned_round5"]).max() <= ROUND_TOL


def test_wave_lockstep_gives_identical_scores(engine):
    """The wave barrier only changes when tiles start, never what they compute."""
    from genevector_b200.synth import synth_counts
    X = synth_counts(2000, 700, seed=11)        # 234 blocks -> 30 tile rows, 465 tiles > 74 pairs
    csr = engine.upload_csr(X)
    old = engine.lockstep
    try:
        engine.lockstep = False
        a = engine.mi_from_csr(csr, 2000, 700, signed=True).cpu().numpy()
        engine.lockstep = True
        b = engine.mi_from_csr(csr, 2000, 700, signed=True).cpu().numpy()
    finally:
        engine.lockstep = old
    assert np.array_equal(a, b) and np.abs(a).max() > 0


def test_build_training_tensors_matches_reference_loop(engine):
    """gv_build_training_tensors vs the reference's _build_training_tensors (data.py:377-411)
    restated with its own dict loop, for signed and unsigned training."""
    import torch
    from genevector_b200.data import build_training_tensors, compute_target_scores
    z = load_golden("ref_poisson_500x50_b10")
    genes = [f"GENE{i}" for i in range(50)]
    sm = compute_target_scores(z["X"], genes, target="mi", target_kwargs={"n_bins": 10}, signed_mi=True,
                               mi_backend="b200", device="cuda", use_cache=False)
    d = sm.to_dict()
    for signed in (True, False):
        n = len(genes)
        score_matrix = np.zeros((n, n), dtype=np.float32)          # data.py:383-388
        for i, g1 in enumerate(genes):
            if g1 in d:
                for j, g2 in enumerate(genes):
                    if i != j and g2 in d[g1]:
                        score_matrix[i, j] = d[g1][g2]
        score_matrix *= 100.0 ** 2                                  # :390
        if not signed:
            score_matrix[score_matrix < 0] = 0.0                    # :392-393
        idx = np.arange(n)
        ig, jg = np.meshgrid(idx, idx, indexing="ij")
        off = ig != jg
        i_idx, j_idx, xij = build_training_tensors(sm, genes, c=100.0, signed_mi=signed, device="cuda")
        assert i_idx.dtype == torch.int64 and xij.dtype == torch.float32 and xij.is_cuda
        assert np.array_equal(i_idx.cpu().numpy(), ig[off].ravel())
        assert np.array_equal(j_idx.cpu().numpy(), jg[off].ravel())
        assert np.array_equal(xij.cpu().numpy(), score_matrix[off].ravel())


@pytest.mark.parametrize("world", [2, 3, 8])
def test_rank_slices_assemble_to_the_single_rank_result(engine, world):
    """Multi-rank path emulated rank by rank on one GPU (no process group): disjoint per-rank
    tiles (MI and the sharded sign map) sum to exactly the single-rank matrix."""
    import torch
    from genevector_b200.synth import synth_counts
    X = synth_counts(1500, 900, seed=21)
    csr = engine.upload_csr(X)
    full = engine.mi_from_csr(csr, 1500, 900, signed=True)
    acc = torch.zeros_like(full)
    nz_owner = torch.zeros_like(full, dtype=torch.int32)
    for rank in range(world):
        part = engine.mi_from_csr(csr, 1500, 900, signed=True, rank=rank, world=world)
        acc += part
        nz_owner += (part != 0).int()
    assert torch.equal(acc, full)
    assert int(nz_owner.max()) == 1


def test_count_path_falls_back_and_refuses(engine):
    """Values >= 256 or fractional values are outside the histogram path: AUTO falls back to the
    general path (same result), COUNTS refuses loudly."""
    from genevector_b200 import _lib
    from genevector_b200.engine import GvError
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    X = sp.csr_matrix((rng.integers(0, 4, size=(300, 20)) * rng.integers(1, 400, size=(300, 20))).astype(np.float32))
    assert X.data.max() >= 256
    xd, nb = orc.discretize_genes_np(X, 10)
    csr = engine.upload_csr(X)
    blk = engine.discretize(csr, 10)
    assert blk.path_used == _lib.GV_PATH_GENERAL
    assert np.array_equal(blk.x_disc(), xd) and np.array_equal(blk.n_bins_per_gene.cpu().numpy(), nb)
    with pytest.raises(GvError, match="count-data path"):
        engine.discretize(csr, 10, path=_lib.GV_PATH_COUNTS)
    # signed MI on such data takes two 7-bit digits per gene (counts up to 16383)
    blk2 = engine.discretize(csr, 10, val_mode=0)
    assert blk2.n_limbs == 2 and blk2.val_rows.shape[0] == 64
    Xd = np.asarray(X.todense()).astype(np.int64)
    vr = blk2.val_rows.cpu().numpy()[:, :300].astype(np.int64)
    assert np.array_equal(vr[0:40:2] + (vr[1:40:2] << blk2.limb_bits), Xd.T)
    assert np.array_equal(blk2.gene_values(), Xd.T)
    want, _ = orc.mi_pairs_c(xd, nb, corr=orc.sign_int_c(X).astype(np.float64), sign_mode=1)
    out = engine.mi_from_csr(csr, 300, 20, signed=True).cpu().numpy()
    assert np.abs(out - want).max() <= MI_TOL
    # larger counts take a third digit (still exact); negative values are refused loudly
    Xbig = X.copy(); Xbig.data[0] = 20000.0
    blk3 = engine.discretize(engine.upload_csr(Xbig), 10, val_mode=0)
    assert blk3.n_limbs == 3 and blk3.val_mode == 0
    assert np.array_equal(blk3.gene_values(), np.asarray(Xbig.todense()).astype(np.int64).T)
    # a negative value: "not expressed" for the bins (metrics.py:53), the raw value for the sign
    Xn = X.copy(); Xn.data[0] = -2.0
    xdn, nbn = orc.discretize_genes_np(Xn, 10)
    sgn_n = np.sign(orc.pearson_sign_np(Xn)).astype(np.int8); np.fill_diagonal(sgn_n, 0)
    want_n, _ = orc.mi_pairs_c(xdn, nbn, corr=sgn_n.astype(np.float64), sign_mode=1)
    out_n = engine.mi_from_csr(engine.upload_csr(Xn), 300, 20, signed=True).cpu().numpy()
    assert np.abs(out_n - want_n).max() <= MI_TOL


@pytest.mark.parametrize("n_cells,n_genes,seed", [(900, 300, 31), (1300, 150, 32)])
def test_two_digit_counts_matrix_targets_and_sign(engine, n_cells, n_genes, seed):
    """Counts up to a few thousand (two INT8 digits per gene): sign map exact, Pearson / cosine
    against np.corrcoef / sklearn, on a case with several 128-gene tiles."""
    from genevector_b200 import _lib
    from genevector_b200.synth import synth_counts
    from oracle import oracle as orc
    X = synth_counts(n_cells, n_genes, seed=seed)
    rng = np.random.default_rng(seed)
    X.data = (X.data * rng.integers(1, 300, size=X.data.shape)).astype(np.float32)
    assert 127 < X.data.max() < 16384
    csr = engine.upload_csr(X)
    blk = engine.discretize(csr, 10, val_mode=0)
    assert blk.n_limbs == 2
    assert np.array_equal(engine.sign_matrix(blk).cpu().numpy()[:, :n_genes], orc.sign_int_c(X))
    p = engine.moment_scores(blk, _lib.GV_MOM_PEARSON).cpu().numpy().astype(np.float64)
    want = orc.pearson_np(X); np.fill_diagonal(want, 0.0)
    assert np.abs(p - want).max() <= 2e-6
    c = engine.moment_scores(blk, _lib.GV_MOM_COSINE).cpu().numpy().astype(np.float64)
    assert np.abs(c - orc.cosine_np(X)).max() <= 2e-6


def test_gene_entropy_matches_reference(engine):
    """data.py:186-203 incl. its quirk of dropping the first unique value even when it is not 0."""
    from genevector_b200.data import get_gene_entropy
    from genevector_b200.engine import GvError
    from oracle import oracle as orc
    z = load_golden("synth_counts_700x96_b10")      # has an all-zero gene, constants, a dense gene
    genes = [f"GENE{i}" for i in range(96)]
    got = get_gene_entropy(z["X"], genes)
    want = orc.gene_entropy_np(z["X"])
    assert np.array_equal(want, z["gene_entropy"])          # the reference's own output
    assert list(got) == genes
    assert np.abs(np.array([got[g] for g in genes]) - want).max() < 1e-12
    # data that are not small counts go through the per-gene sort
    for case in ("lognorm_f32_300x40_b7", "lognorm_f64_300x40_b10", "synth_counts_400x30_b15"):
        zl = load_golden(case)
        gl = [f"G{i}" for i in range(zl["X"].shape[1])]
        e = get_gene_entropy(zl["X"], gl)
        assert np.abs(np.array([e[g] for g in gl]) - zl["gene_entropy"]).max() < 1e-12, case
    Xn = zl["X"].copy(); Xn.data[3] = -1.0; Xn.data[5] = 0.0       # a negative value, a stored zero
    en = get_gene_entropy(Xn, gl)
    assert np.abs(np.array([en[g] for g in gl]) - orc.gene_entropy_np(Xn)).max() < 1e-12


@pytest.mark.parametrize("cg", [1, 2])
@pytest.mark.parametrize("case", ["ref_poisson_500x50_b10", "ref_poisson_500x20_b5", "synth_counts_700x96_b10"])
def test_fp4_operand_counts_and_scores_identical_to_int8(engine, case, cg):
    """The FP4 (E2M1, unit-scale mxf4) operand gives the same joint counts bit for bit, hence the
    same scores, as the INT8 operand."""
    from genevector_b200 import _lib
    z = load_golden(case)
    csr = engine.upload_csr(z["X"])
    blk = engine.discretize(csr, z["n_bins"])
    oh8, B = engine.expand_onehot(blk, "int8")
    oh4, _ = engine.expand_onehot(blk, "fp4")
    assert oh4.shape == (oh8.shape[0], oh8.shape[1] // 2)
    # operand bytes: nibble c&1 of byte c>>1 is 0x2 exactly where the INT8 operand is 1
    a8 = oh8.cpu().numpy()
    a4 = oh4.cpu().numpy()
    assert np.array_equal(a4 & 0x0F, a8[:, 0::2] * 2) and np.array_equal(a4 >> 4, a8[:, 1::2] * 2)
    c8 = engine.gram_dump(oh8, cta_group=cg, active_rows=30).cpu().numpy()
    c4 = engine.gram_dump(oh4, cta_group=cg, active_rows=30, fmt=_lib.GV_OPERAND_FP4).cpu().numpy()
    assert np.array_equal(c4, c8)
    old = engine.cta_group
    engine.cta_group = cg
    try:
        m8 = engine.mi_scores(blk, oh8, B).cpu().numpy()
        m4 = engine.mi_scores(blk, oh4, B).cpu().numpy()
    finally:
        engine.cta_group = old
    assert np.array_equal(m4, m8)
    assert np.abs(m4.astype(np.float64) - z["mi_raw"]).max() <= MI_TOL


def test_fp4_operand_seeded_vs_oracle(engine):
    from oracle import oracle as orc
    from genevector_b200.synth import synth_counts
    X = synth_counts(3000, 333, seed=2)
    bins_gm, nb, _ = orc.discretize_csr_c(X, 10)
    want, _ = orc.mi_pairs_gm_c(bins_gm, nb, sign=orc.sign_int_c(X))
    csr = engine.upload_csr(X)
    old = engine.operand
    engine.operand = "fp4"
    try:
        out = engine.mi_from_csr(csr, 3000, 333, signed=True).cpu().numpy()
    finally:
        engine.operand = old
    assert np.abs(out.astype(np.float64) - want).max() <= MI_TOL


def test_streamed_host_result_equals_device_result(engine):
    """mi_from_csr(out_host=...): the result rows are copied band by band under the computation;
    the pinned host matrix must equal the device matrix and the one-launch result."""
    import torch
    from genevector_b200.synth import synth_counts
    X = synth_counts(1500, 2600, seed=23)          # 867 blocks -> 109 tile rows = 14 bands
    csr = engine.upload_csr(X)
    ref = engine.mi_from_csr(csr, 1500, 2600, signed=True)
    host = torch.full((2600, 2600), float("nan"), dtype=torch.float32).pin_memory()
    out = engine.mi_from_csr(csr, 1500, 2600, signed=True, out_host=host)
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
    assert torch.equal(host, ref.cpu())


def test_input_formats_match_the_reference_semantics(engine):
    """The reference takes whatever `adata.X` is (metrics.py:45, data.py:171): dense arrays, CSC, integer
    or bool dtypes, CSR with unsorted indices or duplicate entries (which scipy's todense sums).  All
    must give the result of the canonical float32 CSR."""
    from genevector_b200.metrics import compute_mi_b200, discretize_genes
    from genevector_b200.synth import synth_counts, gene_names
    from oracle import oracle as orc
    X = synth_counts(600, 60, seed=11)
    genes = gene_names(60)
    ref = compute_mi_b200(X, genes, signed=True).matrix
    want, _ = orc.mi_pairs_c(*orc.discretize_genes_np(X, 10), corr=orc.sign_int_c(X).astype(np.float64), sign_mode=1)
    assert np.abs(ref.astype(np.float64) - want).max() <= MI_TOL
    dense = np.asarray(X.todense())
    variants = {
        "dense float32": dense, "dense float64": dense.astype(np.float64), "csc": sp.csc_matrix(X),
        "coo": sp.coo_matrix(X), "int32": sp.csr_matrix(dense.astype(np.int32)),
        "int64 dense": dense.astype(np.int64), "uint8": sp.csr_matrix(dense.astype(np.uint8)),
    }
    # unsorted column indices within rows
    U = X.copy()
    for r in range(U.shape[0]):
        a, b = U.indptr[r], U.indptr[r + 1]
        U.indices[a:b] = U.indices[a:b][::-1].copy()
        U.data[a:b] = U.data[a:b][::-1].copy()
    U.has_sorted_indices = False
    variants["unsorted csr"] = U
    # duplicate entries: every value split into two stored entries that sum to it
    half = np.floor(X.data / 2)
    data2 = np.empty(2 * X.nnz, dtype=X.data.dtype)
    data2[0::2], data2[1::2] = half, X.data - half
    D = sp.csr_matrix((data2, np.repeat(X.indices, 2), X.indptr * 2), shape=X.shape)
    assert D.nnz == 2 * X.nnz and not D.has_canonical_format
    variants["duplicates"] = D
    for name, V in variants.items():
        got = compute_mi_b200(V, genes, signed=True).matrix
        assert np.array_equal(got, ref), name
    xd, nb = discretize_genes(dense, n_bins=10)
    xr, nr = orc.discretize_genes_np(X, 10)
    assert np.array_equal(xd, xr) and np.array_equal(nb, nr)
    # bool input: every expressed gene has one unique value -> one non-zero bin (metrics.py:60-62)
    B = sp.csr_matrix(dense > 0)
    xb, nbb = discretize_genes(B, n_bins=10)
    xbr, nbr = orc.discretize_genes_np(B.astype(np.float32), 10)
    assert np.array_equal(xb, xbr) and np.array_equal(nbb, nbr)


def test_rust_sign_rule(engine):
    """sign_rule="rust" (src/lib.rs:91-94: +1 unless the correlation is negative) against the C port of
    the Rust tier; where the covariance is exactly zero the Rust tier signs with float64 rounding
    noise, the exact integer covariance here gives +1."""
    from genevector_b200.metrics import compute_mi_b200, target_mi
    from oracle import oracle as orc
    z = load_golden("synth_counts_700x96_b10")        # has constant and all-zero genes
    genes = [f"GENE{i}" for i in range(96)]
    corr = orc.pearson_sign_np(z["X"])                # nan_to_num(corrcoef) as compute_mi_rust passes it
    want, _ = orc.mi_pairs_c(z["x_disc"].astype(np.int32), z["n_bins_per_gene"], corr=corr, sign_mode=2)
    got = compute_mi_b200(z["X"], genes, signed=True, sign_rule="rust").matrix.astype(np.float64)
    exact0 = z["int_sign"] == 0
    assert np.abs(got - want)[~exact0].max() <= MI_TOL
    assert np.abs(got - np.abs(want))[exact0].max() <= MI_TOL         # +1 there
    assert (got[exact0 & (z["mi_raw"] > 0)] > 0).any()
    # default rule zeroes those pairs; the kwarg travels through the target function
    dflt = compute_mi_b200(z["X"], genes, signed=True).matrix
    assert not dflt[exact0].any()
    via = target_mi(z["X"], genes, signed=True, backend="b200", sign_rule="rust").matrix
    assert np.array_equal(via, got.astype(np.float32))
    with pytest.raises(ValueError, match="sign rule"):
        compute_mi_b200(z["X"], genes, signed=True, sign_rule="julia")


@pytest.mark.parametrize("extra", [[], ["--operand", "int8"], ["--unsigned"], ["--target", "pearson"],
                                   ["--target", "spearman"], ["--target", "jaccard"]])
def test_bench_lines_carry_the_contract_keys(extra):
    """bench.py on a small workload: one JSON line on stdout with every key of the bench contract."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "C2", "--steps", "2",
                        "--warmup", "3", "--no-cpu-baseline"] + extra, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"):
        assert key in d, key
    assert d["metric"] == "gene-pairs/s" and d["value"] > 0 and d["gpu_launches"] > 0
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in d["e2e"], key
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in d["roofline"], key
    assert 0 < d["roofline"]["frac"] < 1.05 and "workload" in d["config"]
    if "--target" not in extra:
        # the e2e leg is the reference-facing plugin call
        assert "target_mi" in d["e2e"]["note"] and d["e2e"]["h2d_bytes_per_step"] < d["e2e"]["host_csr_bytes"]
    if not extra:
        r8 = d["roofline_int8"]                      # the metric's own dtype, from a short INT8 pass
        assert r8["achieved"] > 0 and 0 < r8["frac"] < 1.05 and "kind::i8" in r8["peak_note"]
