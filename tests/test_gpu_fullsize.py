"""GPU tests at BASELINE.json's full shapes.  C2 is compared with the oracle exhaustively; the
larger configurations through size-independent properties (symmetry, zero diagonal, range,
idempotence) plus the oracle on a seeded sample of genes / gene pairs taken from the very buffers
the kernels produced (bins, scores)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MI_TOL = 1e-6
LOG2_11 = float(np.log2(11.0))


def _unpack_rows(blk, genes):
    """Bins of the given genes from the packed device block -> uint8 [len(genes)][n_cells]."""
    b = blk.bins[torch.as_tensor(genes, device=blk.bins.device)].cpu().numpy()
    full = np.empty((len(genes), b.shape[1] * 2), dtype=np.uint8)
    full[:, 0::2], full[:, 1::2] = b & 0x0F, b >> 4
    return full[:, :blk.n_cells]


def _dense_columns(csr, genes):
    """Dense float32 [n_cells][len(genes)] of the given gene columns of a device CSR."""
    gsel = torch.as_tensor(genes, device=csr.col_idx.device, dtype=torch.int32)
    hit = torch.isin(csr.col_idx, gsel)
    e = hit.nonzero().flatten()
    rows = torch.searchsorted(csr.row_ptr, e, right=True) - 1
    lut = torch.full((csr.n_genes,), -1, dtype=torch.int64, device=e.device)
    lut[gsel.long()] = torch.arange(len(genes), device=e.device)
    cols = lut[csr.col_idx[e].long()]
    D = torch.zeros((csr.n_cells, len(genes)), dtype=torch.float32, device=e.device)
    D[rows, cols] = csr.values[e].float()
    return D.cpu().numpy()


def _make(engine, n_cells, n_genes, seed=0):
    from genevector_b200.engine import CsrDevice
    from genevector_b200.synth import synth_counts_torch
    v, c, p = synth_counts_torch(n_cells, n_genes, seed=seed, device=engine.device)
    return CsrDevice(v, c, p, n_cells, n_genes)


def _check_sampled(engine, csr, out, blk, n_gene_sample, n_pair_sample, signed, seed=0):
    from oracle import oracle as orc
    rng = np.random.default_rng(seed)
    G, N = csr.n_genes, csr.n_cells
    genes = np.sort(rng.choice(G, size=min(n_gene_sample, G), replace=False))
    # discretisation of the sampled genes: bit-exact vs the oracle on the same columns
    D = _dense_columns(csr, genes)
    xd, nb = orc.discretize_genes_c(D, blk.n_bins)
    got = _unpack_rows(blk, genes)
    assert np.array_equal(got.T.astype(np.int32), xd), "bins differ from the oracle"
    assert np.array_equal(blk.n_bins_per_gene[torch.as_tensor(genes, device=engine.device)].cpu().numpy(), nb)
    # scores of sampled pairs among those genes
    sign = orc.sign_int_c(D) if signed else None
    pairs = [(a, b) for a in range(len(genes)) for b in range(a + 1, len(genes))]
    sel = rng.choice(len(pairs), size=min(n_pair_sample, len(pairs)), replace=False)
    worst = 0.0
    sub = out[torch.as_tensor(genes, device=out.device)][:, torch.as_tensor(genes, device=out.device)].cpu().numpy()
    for k in sel:
        a, b = pairs[k]
        if nb[a] <= 1 or nb[b] <= 1:
            want = 0.0
        else:
            J = orc.joint_np(xd[:, a], xd[:, b], nb[a], nb[b])
            want = orc.mi_from_joint_np(J) if J.sum() > 0 else 0.0
            if signed:
                want *= float(sign[a, b])
        worst = max(worst, abs(float(sub[a, b]) - want))
    assert worst <= MI_TOL, f"max |score - oracle| over sampled pairs = {worst:.3e}"
    return worst


def _structural(out, signed):
    G = out.shape[0]
    # symmetric, zero diagonal, |MI| <= log2(11); checked on strided blocks to stay cheap at 20k genes
    idx = torch.arange(0, G, max(1, G // 2048), device=out.device)
    sub = out[idx][:, idx]
    assert torch.equal(sub, sub.T)
    assert float(out.diagonal().abs().max()) == 0.0
    assert torch.isfinite(out).all()
    amax = float(out.abs().max())
    assert amax <= LOG2_11 + 1e-4 and amax > 0
    if not signed:
        assert float(out.min()) >= 0.0


def test_c2_full_vs_oracle(engine):
    """configs[1]: 2,000 genes x 5,000 cells, every one of the 1,999,000 pairs against the oracle."""
    from oracle import oracle as orc
    from genevector_b200.synth import synth_counts
    X = synth_counts(5000, 2000, seed=0)
    bins_gm, nb, _ = orc.discretize_csr_c(X, 10)
    sign = orc.sign_int_c(X)
    want, n = orc.mi_pairs_gm_c(bins_gm, nb, sign=sign)
    csr = engine.upload_csr(X)
    blk = engine.discretize(csr, 10, val_mode=0)
    assert np.array_equal(blk.x_disc(), bins_gm.T.astype(np.int32))
    assert np.array_equal(blk.n_bins_per_gene.cpu().numpy(), nb)
    assert np.array_equal(engine.sign_matrix(blk).cpu().numpy()[:, :2000], sign)
    out = engine.mi_from_csr(csr, 5000, 2000, n_bins=10, signed=True)
    err = np.abs(out.cpu().numpy().astype(np.float64) - want)
    assert err.max() <= MI_TOL, f"max err {err.max():.3e}"
    _structural(out, signed=True)


def test_c3_sampled(engine):
    """configs[2]: 5,000 genes x 50,000 cells."""
    csr = _make(engine, 50000, 5000)
    blk = engine.discretize(csr, 10, val_mode=0)
    out = engine.mi_from_csr(csr, 50000, 5000, n_bins=10, signed=True)
    _structural(out, signed=True)
    _check_sampled(engine, csr, out, blk, n_gene_sample=48, n_pair_sample=400, signed=True)
    # idempotence: a second run gives the same bits
    out2 = engine.mi_from_csr(csr, 50000, 5000, n_bins=10, signed=True)
    assert torch.equal(out, out2)


def test_c4_sampled(engine):
    """configs[3], the headline shape on ONE GPU: 20,000 genes x 200,000 cells (~200M pairs)."""
    free, _ = torch.cuda.mem_get_info()
    if free < 90 * 2 ** 30:
        pytest.skip("needs ~60 GB of free HBM")
    csr = _make(engine, 200000, 20000)
    blk = engine.discretize(csr, 10, val_mode=0)
    out = engine.mi_from_csr(csr, 200000, 20000, n_bins=10, signed=True)
    _structural(out, signed=True)
    _check_sampled(engine, csr, out, blk, n_gene_sample=40, n_pair_sample=300, signed=True)
    # unsigned run: same magnitudes, no sign
    out_u = engine.mi_from_csr(csr, 200000, 20000, n_bins=10, signed=False)
    a, b = out_u[:4096, :4096], out[:4096, :4096]
    nz = b != 0                       # sign 0 (exactly zero covariance / constant gene) zeroes the score
    assert torch.equal(a[nz], b[nz].abs()) and float(nz.float().mean()) > 0.99
    del out, out_u, blk, csr
    torch.cuda.empty_cache()


def test_wide_matrix_beyond_2g_result_elements(engine):
    """48,000 genes x 20,000 cells (a whole-transcriptome gene axis): 2.3e9 result elements, i.e. every
    64-bit index of the result / sign-map / tile arithmetic beyond 2^31, checked on genes sampled from
    the far end of the gene axis as well as at random, plus the streamed host copy."""
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~40 GB of free HBM")
    G, N = 48000, 20000
    csr = _make(engine, N, G, seed=8)
    blk = engine.discretize(csr, 10, val_mode=0)
    out = engine.mi_from_csr(csr, N, G, n_bins=10, signed=True)
    assert out.numel() > 2 ** 31
    _structural(out, signed=True)
    _check_sampled(engine, csr, out, blk, n_gene_sample=40, n_pair_sample=300, signed=True, seed=1)
    # the last genes of the axis (largest row / column offsets)
    from oracle import oracle as orc
    genes = np.arange(G - 24, G)
    D = _dense_columns(csr, genes)
    xd, nb = orc.discretize_genes_c(D, 10)
    want, _ = orc.mi_pairs_c(xd, nb, corr=orc.sign_int_c(D).astype(np.float64), sign_mode=1)
    gi = torch.as_tensor(genes, device=out.device)
    got = out[gi][:, gi].cpu().numpy().astype(np.float64)
    assert np.abs(got - want).max() <= MI_TOL
    del out, blk, csr
    torch.cuda.empty_cache()


def test_atlas_scale_cell_axis(engine):
    """1,048,593 cells x 1,200 genes: a cell axis beyond 2^20 (6-bit digits for the exact co-moments,
    8,193 cell blocks in the transposition, K extent not a multiple of anything)."""
    from genevector_b200.engine import limb_bits_for
    G, N = 1200, (1 << 20) + 17
    assert limb_bits_for(N) == 6
    csr = _make(engine, N, G, seed=9)
    blk = engine.discretize(csr, 10, val_mode=0)
    assert blk.limb_bits == 6 and blk.val_mode == 0
    out = engine.mi_from_csr(csr, N, G, n_bins=10, signed=True)
    _structural(out, signed=True)
    _check_sampled(engine, csr, out, blk, n_gene_sample=16, n_pair_sample=60, signed=True, seed=2)
    del out, blk, csr
    torch.cuda.empty_cache()


def test_c5_matrix_targets_sampled(engine):
    """configs[4]: 10,000 genes x 100,000 cells, pearson / cosine / jaccard on the shared skeleton."""
    from genevector_b200 import _lib
    from oracle import oracle as orc
    csr = _make(engine, 100000, 10000, seed=5)
    rng = np.random.default_rng(5)
    genes = np.sort(rng.choice(10000, size=40, replace=False))
    D = _dense_columns(csr, genes)
    gi = torch.as_tensor(genes, device=engine.device)
    for kind, ref in ((_lib.GV_MOM_PEARSON, orc.pearson_np), (_lib.GV_MOM_COSINE, orc.cosine_np),
                      (_lib.GV_MOM_JACCARD, orc.jaccard_np)):
        out = engine.target_from_csr(csr, kind)
        idx = torch.arange(0, 10000, 5, device=out.device)
        sub = out[idx][:, idx]
        assert torch.equal(sub, sub.T) and float(out.diagonal().abs().max()) == 0.0
        got = out[gi][:, gi].cpu().numpy().astype(np.float64)
        want = np.array(ref(D), dtype=np.float64)
        np.fill_diagonal(want, 0.0)
        tol = 0.0 if kind == _lib.GV_MOM_JACCARD else 2e-6     # float32 output of an exact quotient
        if kind == _lib.GV_MOM_JACCARD:
            assert np.array_equal(got.astype(np.float32), want.astype(np.float32))
        else:
            assert np.abs(got - want).max() <= tol + 1e-6, np.abs(got - want).max()
        del out


def test_c5_spearman_and_lognorm_targets_sampled(engine):
    """configs[4] shape with the value-row modes beyond single-digit counts: Spearman (three rank
    digits, 2*N = 200,000) and Pearson / signed-MI sign on log-normalised float32 input (three
    fixed-point digits): 128-bit covariances and 63-bit co-moment sums at full size."""
    from genevector_b200 import _lib
    from genevector_b200.engine import CsrDevice
    from oracle import oracle as orc
    csr = _make(engine, 100000, 10000, seed=6)
    rng = np.random.default_rng(6)
    genes = np.sort(rng.choice(10000, size=32, replace=False))
    gi = torch.as_tensor(genes, device=engine.device)
    D = _dense_columns(csr, genes)
    out = engine.target_from_csr(csr, _lib.GV_MOM_PEARSON, ranks=True)
    got = out[gi][:, gi].cpu().numpy().astype(np.float64)
    want = orc.spearman_np(D); np.fill_diagonal(want, 0.0)
    assert np.abs(got - want).max() <= 1e-6, np.abs(got - want).max()
    idx = torch.arange(0, 10000, 7, device=out.device)
    assert torch.equal(out[idx][:, idx], out[idx][:, idx].T)
    del out
    # log-normalised values: log1p(count / cell total * 1e4), computed on the device
    tot = torch.zeros(csr.n_cells, dtype=torch.float32, device=engine.device)
    rows = torch.repeat_interleave(torch.arange(csr.n_cells, device=engine.device),
                                   csr.row_ptr[1:] - csr.row_ptr[:-1])
    tot.index_add_(0, rows, csr.values)
    vals = torch.log1p(csr.values / (tot[rows] + 1.0) * 1e4)
    csr_l = CsrDevice(vals, csr.col_idx, csr.row_ptr, csr.n_cells, csr.n_genes)
    blk = engine.discretize(csr_l, 10, val_mode=0)
    assert (blk.val_mode, blk.n_limbs) == (2, 3) and blk.path_used == _lib.GV_PATH_GENERAL
    Dl = _dense_columns(csr_l, genes)
    xd, nb = orc.discretize_genes_c(Dl, 10)
    assert np.array_equal(_unpack_rows(blk, genes).T.astype(np.int32), xd)
    p = engine.moment_scores(blk, _lib.GV_MOM_PEARSON)
    got = p[gi][:, gi].cpu().numpy().astype(np.float64)
    want = orc.pearson_np(Dl); np.fill_diagonal(want, 0.0)
    assert np.abs(got - want).max() <= 2e-6, np.abs(got - want).max()
    sgn = engine.sign_matrix(blk)[gi][:, gi].cpu().numpy()
    ref_sign = np.sign(orc.pearson_sign_np(Dl)).astype(np.int8); np.fill_diagonal(ref_sign, 0)
    assert np.array_equal(sgn, ref_sign)


def test_c1_pbmc_like_full_vs_oracle_and_training_handoff(engine):
    """configs[0] stand-in (example/PBMC.h5ad is a missing blob): 1,000 genes x 13,000 cells,
    seed 42; every pair against the oracle, then the hand-off to the trainer: the (i_idx, j_idx,
    xij) tensors drive two steps of the reference's training objective (model.py:82-85,149-150,
    172-199: two embedding tables, row-wise dot product, MSE, Adadelta) restated here."""
    from oracle import oracle as orc
    from genevector_b200.data import build_training_tensors, compute_target_scores, get_gene_entropy
    from genevector_b200.synth import synth_counts, gene_names
    G, N = 1000, 13000
    X = synth_counts(N, G, seed=42)
    genes = gene_names(G)
    bins_gm, nb, _ = orc.discretize_csr_c(X, 10)
    want, _ = orc.mi_pairs_gm_c(bins_gm, nb, sign=orc.sign_int_c(X))
    sm = compute_target_scores(X, genes, target="mi", signed_mi=True, mi_backend="b200", device="cuda")
    assert np.abs(sm.matrix.astype(np.float64) - want).max() <= MI_TOL
    ent = get_gene_entropy(X, genes)
    assert np.abs(np.array([ent[g] for g in genes]) - orc.gene_entropy_np(X)).max() < 1e-12
    i_idx, j_idx, xij = build_training_tensors(sm, genes, c=100.0, signed_mi=True, device="cuda")
    assert i_idx.numel() == G * (G - 1) and xij.is_cuda
    torch.manual_seed(0)
    wi = torch.nn.Embedding(G, 100).cuda()
    wj = torch.nn.Embedding(G, 100).cuda()
    opt = torch.optim.Adadelta(list(wi.parameters()) + list(wj.parameters()))
    losses = []
    for _ in range(2):
        opt.zero_grad()
        pred = (wi(i_idx) * wj(j_idx)).sum(dim=1)
        loss = torch.nn.functional.mse_loss(pred, xij)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert np.isfinite(losses).all() and losses[1] < losses[0]
