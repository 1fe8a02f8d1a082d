"""CPU tests of the boundary and the host logic: the C-ABI library loads and exports every
symbol include/gvb200.h declares, fails loudly without a B200, and the host-side pieces (tile
schedule, broadcast block layout, ScoreMatrix, score-building fast paths) behave."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol():
    from genevector_b200 import _lib
    header = open(os.path.join(ROOT, "include", "gvb200.h")).read()
    declared = set(re.findall(r"\b(gv_[a-z0-9_]+)\s*\(", header))
    declared -= {"gv_status", "gv_dtype", "gv_moment_kind"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gv_abi_version() == _lib.GV_ABI_VERSION


def test_ctypes_signatures_have_the_arity_the_header_declares():
    """Every function of include/gvb200.h is bound with as many arguments as it declares (a mismatch would
    shift every later argument of a call)."""
    from genevector_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "gvb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decls = re.findall(r"\b(gv_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    assert len(decls) >= 40
    for name, params in decls:
        params = params.strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert len(_lib.SIGNATURES[name][1]) == n, name


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point refuses; nothing routes to the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from genevector_b200 import _lib
    from genevector_b200.engine import Engine, GvError
    lib = _lib.load()
    assert lib.gv_device_check() < 0 and lib.gv_last_error()
    with pytest.raises(GvError):
        Engine()
    src = "".join(open(os.path.join(ROOT, "genevector_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "genevector_b200")) if f.endswith(".py"))
    assert "oracle" not in src.replace("oracle/", "")   # the product never imports the oracle


def test_argument_checks_do_not_need_a_gpu():
    from genevector_b200 import _lib
    lib = _lib.load()
    assert lib.gv_rows_per_gene(10) == 10 and lib.gv_rows_per_gene(5) == 5
    assert lib.gv_rows_per_gene(7) == 8 and lib.gv_rows_per_gene(15) == 16
    assert lib.gv_rows_per_gene(16) == 32 and lib.gv_rows_per_gene(31) == 32
    assert lib.gv_rows_per_gene(32) < 0 and b"n_bins" in lib.gv_last_error()
    assert lib.gv_onehot_rows(100, 32) == 3200
    assert lib.gv_onehot_rows(20000, 10) == 6667 * 32
    assert lib.gv_onehot_rows(7, 5) == 64
    assert lib.gv_tile_m(1) == 128 and lib.gv_tile_m(2) == 256 and lib.gv_tile_m(3) < 0


@pytest.mark.parametrize("op_rows,tile_m,band", [(256, 256, 8), (1312, 256, 8), (5000, 256, 3),
                                                 (1312, 128, 16), (640, 128, 1), (213376, 256, 8)])
def test_tile_schedule_covers_upper_triangle_once(op_rows, tile_m, band):
    from genevector_b200 import _lib
    lib = _lib.load()
    n = lib.gv_tile_schedule(op_rows, tile_m, band, None, 0)
    tiles = np.zeros((n, 2), dtype=np.int32)
    assert lib.gv_tile_schedule(op_rows, tile_m, band, tiles.ctypes.data, n) == n
    keys = tiles[:, 0].astype(np.int64) * (1 << 32) + tiles[:, 1]
    assert len(np.unique(keys)) == n                      # no tile twice
    assert np.all(tiles[:, 0] % tile_m == 0) and np.all(tiles[:, 1] % 256 == 0)
    tr, tc = -(-op_rows // tile_m), -(-op_rows // 256)
    r, c = np.meshgrid(np.arange(tr), np.arange(tc), indexing="ij")
    need = r * tile_m <= c * 256 + 255                    # tile holds some (row <= col)
    assert n == int(need.sum())
    got = np.zeros_like(need)
    got[tiles[:, 0] // tile_m, tiles[:, 1] // 256] = True
    assert np.array_equal(got, need)
    if op_rows <= 5000:
        # every (row <= col) operand pair is inside exactly one scheduled tile
        cover = np.zeros((op_rows, op_rows), dtype=np.int8)
        for r0, c0 in tiles:
            cover[r0:r0 + tile_m, c0:c0 + 256] += 1
        iu = np.triu_indices(op_rows)
        assert np.all(cover[iu] == 1)
    # L2 locality: tiles that run together (148/2 clusters) touch few distinct panels
    if n >= 74 * 4 and tile_m == 256 and band == 8:
        w = tiles[74:148]
        assert len(np.unique(w[:, 0])) + len(np.unique(w[:, 1])) <= 24


def test_rank_slices_partition_the_schedule():
    from genevector_b200.engine import schedule_host, rank_bounds
    th, tile_m = schedule_host(213344, 2, 8)          # the one-hot operand of 20,000 genes
    n = len(th)
    for world in (1, 2, 4, 8):
        b = rank_bounds(th, tile_m, 8, world)
        assert b[0] == 0 and b[-1] == n and all(b[r] <= b[r + 1] for r in range(world))
        sizes = [b[r + 1] - b[r] for r in range(world)]
        # equal shares snapped to whole tile columns of a band (at most 8 tiles) or a diagonal block
        assert max(sizes) - min(sizes) <= 2 * 36, sizes


@pytest.mark.parametrize("n_genes,B,cg,band,world", [(700, 10, 2, 8, 1), (700, 10, 2, 8, 3), (1500, 10, 2, 3, 4),
                                                     (1100, 5, 1, 8, 2), (2600, 16, 2, 8, 8), (900, 8, 1, 2, 5),
                                                     (96, 10, 2, 8, 2)])
@pytest.mark.parametrize("chunks", [3, 8])
def test_host_pieces_are_exactly_what_the_tiles_write(n_genes, B, cg, band, world, chunks):
    """The rectangles a rank copies to the shared host matrix (engine.host_pieces) are exactly the
    entries its tiles write (pair (i, j), its mirror and the diagonal), every entry of the matrix is
    owned by exactly one rank, and a piece's rectangles are final once its tiles have run."""
    from genevector_b200 import _lib
    from genevector_b200.engine import host_pieces, tile_slice
    lib = _lib.load()
    gpb = 32 // B
    op_rows = lib.gv_onehot_rows(n_genes, B)
    tile_m = lib.gv_tile_m(cg)
    owner = np.zeros((n_genes, n_genes), dtype=np.int32)
    for rank in range(world):
        tiles, _ = tile_slice(op_rows, cg, band, rank, world)
        pieces = host_pieces(op_rows, cg, band, gpb, n_genes, rank, world, chunks=chunks)
        assert not pieces or (pieces[0][0] == 0 and pieces[-1][1] == len(tiles))
        assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
        assert all(t1 > t0 for t0, t1, _ in pieces) and len(pieces) <= chunks
        written = np.zeros((n_genes, n_genes), dtype=bool)      # by the tiles run so far
        copied = np.zeros((n_genes, n_genes), dtype=np.int32)
        for t0, t1, rects in pieces:
            for r0, c0 in tiles[t0:t1]:
                gi = np.arange(r0 // 32 * gpb, min(n_genes, (r0 + tile_m) // 32 * gpb))
                gj = np.arange(c0 // 32 * gpb, min(n_genes, (c0 + 256) // 32 * gpb))
                I, J = np.meshgrid(gi, gj, indexing="ij")
                m = I <= J
                written[I[m], J[m]] = True
                written[J[m], I[m]] = True
            for r0, r1, c0, c1 in rects:
                assert 0 <= r0 < r1 <= n_genes and 0 <= c0 < c1 <= n_genes
                if world > 1:
                    assert written[r0:r1, c0:c1].all(), "a rectangle is copied before its tiles ran"
                copied[r0:r1, c0:c1] += 1
        assert copied.max() <= 1, "a rank copies an entry twice"
        if world > 1:
            assert np.array_equal(copied.astype(bool), written), "rectangles != entries written by the tiles"
        owner += copied
    assert np.array_equal(owner, np.ones_like(owner)), "an entry of the matrix has zero or two owners"


def test_block_layout_and_limbs():
    from genevector_b200.engine import block_layout, limb_bits_for, quantile_table
    kp, pitch, lay, total = block_layout(5000, 2000, 0)
    assert kp == 5120 and pitch == 2560 and pitch % 16 == 0
    offs = sorted(lay.values())
    for (o1, n1), (o2, _) in zip(offs, offs[1:]):
        assert o1 % 256 == 0 and o1 + n1 <= o2
    assert total >= offs[-1][0] + offs[-1][1]
    assert block_layout(5000, 2000, -1)[2]["val"][1] == 0
    # a two-digit block is the one-digit block plus a tail (what the second broadcast sends)
    _, _, lay2, total2 = block_layout(5000, 2000, 0, 2)
    assert all(lay2[k] == lay[k] for k in lay if k != "val")
    assert lay2["val"][0] == lay["val"][0] and total2 > total
    # value rows: 32-row blocks of 32 // n_limbs genes (three digits leave two zero rows per block)
    from genevector_b200.engine import value_rows, max_limbs_for
    assert [value_rows(2000, L) for L in (1, 2, 3, 4)] == [2016, 4000, 6400, 8000]
    assert lay["val"][1] == 2016 * kp and lay2["val"][1] == 4000 * kp
    assert max_limbs_for(200000, 7) == 3 and max_limbs_for(5000, 7) == 3 and max_limbs_for(100, 7) == 4
    assert limb_bits_for(200000) == 7 and limb_bits_for(266287) == 7 and limb_bits_for(300000) == 6
    q = quantile_table()
    assert np.array_equal(q[10, :9], np.linspace(0, 100, 11)[1:-1] / 100)


def test_score_matrix_has_the_reference_dict_semantics():
    from genevector_b200.scores import ScoreMatrix
    z = load_golden("ref_poisson_500x20_b5")
    genes = [f"GENE{i}" for i in range(20)]
    sm = ScoreMatrix(z["mi_raw"].astype(np.float32), genes)
    ref = z["mi_signed_round5"] * 0 + z["mi_unsigned_round5"]
    assert len(sm) == 20 and list(sm)[:2] == ["GENE0", "GENE1"]
    assert "GENE3" in sm and "NOPE" not in sm
    assert "GENE3" in sm["GENE2"] and "GENE2" not in sm["GENE2"]
    assert sm["GENE2"]["GENE3"] == ref[2, 3] and sm["GENE2"]["GENE2"] == 0.0
    assert sm["GENE2"]["NOPE"] == 0.0                          # defaultdict(float) read
    assert len(sm["GENE0"]) == 19 and len(list(sm["GENE0"].items())) == 19
    d = sm.to_dict()
    assert d["GENE5"]["GENE9"] == ref[5, 9] and "GENE5" not in d["GENE5"]
    assert np.array_equal(sm.rounded(), ref)
    # the reference's consumer loop (data.py:383-388) over the view
    n = len(genes)
    M = np.zeros((n, n), dtype=np.float32)
    for i, g1 in enumerate(genes):
        if g1 in sm:
            for j, g2 in enumerate(genes):
                if i != j and g2 in sm[g1]:
                    M[i, j] = sm[g1][g2]
    assert np.array_equal(M, ref.astype(np.float32))


def test_synth_generator_shape_and_density():
    from genevector_b200.synth import synth_counts
    X = synth_counts(2000, 300, seed=0)
    assert X.shape == (2000, 300) and X.dtype == np.float32 and X.indices.dtype == np.int32
    dens = X.nnz / (2000 * 300)
    assert 0.05 < dens < 0.2
    assert X.data.min() >= 1 and np.array_equal(X.data, np.rint(X.data))
    Y = synth_counts(2000, 300, seed=0)
    assert (X != Y).nnz == 0


def test_cache_roundtrip_and_key_compat(tmp_path, monkeypatch):
    """Same key as the reference (cache.py:29-44) and a round trip through its .npz schema
    (tests/test_cache.py:7-20 of the reference, tolerance 1e-4 there)."""
    import hashlib, json
    from genevector_b200 import cache
    from genevector_b200.scores import ScoreMatrix
    monkeypatch.setattr(cache, "CACHE_DIR", str(tmp_path))
    z = load_golden("ref_poisson_500x20_b5")
    genes = [f"GENE{i}" for i in range(20)]
    key = cache.compute_cache_key(z["X"], genes, "mi", {"n_bins": 5}, True)
    # the reference's recipe, restated
    X = z["X"]
    h0 = hashlib.sha256()
    for a in (X.data, X.indices, X.indptr):
        h0.update(np.ascontiguousarray(a).tobytes())
    h0.update(str(X.shape).encode())
    h = hashlib.sha256()
    h.update(h0.hexdigest()[:16].encode())
    h.update(json.dumps(sorted(genes)).encode())
    h.update(b"mi")
    h.update(json.dumps({"n_bins": 5}, sort_keys=True, default=str).encode())
    h.update(b"True")
    assert key == h.hexdigest()[:32] and len(key) == 32
    assert key != cache.compute_cache_key(z["X"], genes, "pearson", {"n_bins": 5}, True)
    sm = ScoreMatrix(z["mi_raw"].astype(np.float32), genes)
    assert cache.load_scores(key) == (None, None)
    path = cache.save_scores(key, sm, genes)
    raw = np.load(path)
    assert raw["scores"].dtype == np.float32 and raw["scores"].shape == (20, 20)
    assert [str(g) for g in raw["genes"]] == genes
    back, g2 = cache.load_scores(key)
    assert g2 == genes and np.abs(back.rounded() - z["mi_unsigned_round5"]).max() < 1e-6
    # a nested dict (what the reference passes) is accepted too
    cache.save_scores(key + "d", sm.to_dict(), genes)
    back2, _ = cache.load_scores(key + "d")
    assert np.array_equal(back2.matrix, back.matrix)


def test_sign_tiles_cover_every_pair_of_the_rank(monkeypatch):
    """Host logic of the sharded sign pass: the sign tiles chosen for a rank's MI tiles contain
    every gene pair those MI tiles produce."""
    from genevector_b200 import engine as E
    from genevector_b200 import _lib
    lib = _lib.load()

    class Fake:                      # Engine without a device: only the pure host method is used
        cta_group, band = 2, 8
        device = "cpu"
    G, B = 5000, 10
    gpb = 32 // B
    rows = lib.gv_onehot_rows(G, B)
    for world in (2, 8):
        for rank in (0, world - 1):
            mine, _ = E.tile_slice(rows, 2, 8, rank, world)
            st = E.Engine.sign_tiles_for(Fake(), mine, B).numpy()
            cover = np.zeros((-(-G // 256), -(-G // 256)), dtype=bool)
            cover[st[:, 0] // 256, st[:, 1] // 256] = True
            for r0, c0 in mine[:: max(1, len(mine) // 200)]:
                gi = np.arange(r0 // 32 * gpb, min(G, (r0 + 256) // 32 * gpb))
                gj = np.arange(c0 // 32 * gpb, min(G, (c0 + 256) // 32 * gpb))
                ii, jj = np.meshgrid(gi, gj, indexing="ij")
                m = ii < jj
                assert cover[ii[m] // 256, jj[m] // 256].all()
            assert len(st) < 0.75 * len(E.tile_slice(G, 2, 8, 0, 1)[0]) or world == 2


def test_graph_xcorr_argument_errors_match_the_reference():
    """genevector/_graph_targets.py:38-42 and _aggregation.py:40-43: same exceptions, raised before
    any device work."""
    import pytest
    from genevector_b200.metrics import target_graph_xcorr, TARGETS
    assert "graph_xcorr" in TARGETS and "spearman" in TARGETS
    X = np.array([[1, 2], [3, 4]], dtype=np.float64)
    with pytest.raises(ValueError, match="graph required"):
        target_graph_xcorr(X, ["a", "b"])
    import scipy.sparse as sp
    with pytest.raises(ValueError, match="Unknown aggregation"):
        target_graph_xcorr(X, ["a", "b"], graph=sp.identity(2, format="csr"), aggr="nonexistent")
    with pytest.raises(ValueError, match="callable"):
        target_graph_xcorr(X, ["a", "b"], graph=sp.identity(2, format="csr"), aggr=lambda X, G, **kw: X)


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """bench.py --impl reference (the CPU arm: runs without a GPU): exactly one line on stdout, a JSON
    object with the keys of the bench contract."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "300x2000", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "impl", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "gene-pairs/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def test_install_into_reference_uses_the_registry_hook():
    """INTEGRATION.md route A: the B200 targets are re-registered through the reference's own
    register_target decorator (metrics.py:347-352), replacing the entries of its TARGETS table."""
    import types
    from genevector_b200 import metrics as ours
    ref = types.ModuleType("genevector.metrics")
    ref.TARGETS = {"mi": object(), "pearson": object(), "custom_user_target": object()}

    def register_target(name):
        def wrapper(fn):
            ref.TARGETS[name] = fn
            return fn
        return wrapper
    ref.register_target = register_target
    keep = ref.TARGETS["custom_user_target"]
    assert ours.install_into_reference(ref) is ref
    for name in ("mi", "pearson", "spearman", "cosine", "jaccard", "graph_xcorr"):
        assert ref.TARGETS[name] is ours.TARGETS[name]
    assert ref.TARGETS["custom_user_target"] is keep
    # the call GeneVectorDataset._compute_target_scores makes (data.py:319-326): unknown backends raise
    with pytest.raises(ValueError, match="Unknown MI backend"):
        ref.TARGETS["mi"](None, [], signed=True, backend="numba", device="cpu")


REF_DIR = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_DIR, "genevector")),
                    reason="reference sources only exist in the build container")
def test_reference_consumers_accept_the_score_matrix(tmp_path, monkeypatch):
    """Drop-in check against the reference's own consumer code (build container only): its
    GeneVectorDataset._build_training_tensors (data.py:377-411), cache.save_scores / load_scores
    (cache.py:64-110) and compute_cache_key run unmodified on a ScoreMatrix and give what they give on
    the nested dict the reference's tiers return."""
    import collections
    import importlib
    import sys
    import types
    from genevector_b200 import cache as our_cache
    from genevector_b200.scores import ScoreMatrix
    pkg = types.ModuleType("genevector")
    pkg.__path__ = [os.path.join(REF_DIR, "genevector")]
    monkeypatch.setitem(sys.modules, "genevector", pkg)
    for sub in ("genevector.data", "genevector.cache", "genevector.metrics", "genevector._aggregation",
                "genevector._graph_targets"):
        monkeypatch.delitem(sys.modules, sub, raising=False)
    refd = importlib.import_module("genevector.data")
    refc = importlib.import_module("genevector.cache")
    monkeypatch.setattr(refc, "CACHE_DIR", str(tmp_path))
    z = load_golden("ref_poisson_500x20_b5")
    genes = [f"GENE{i}" for i in range(20)]
    M = z["mi_signed_round5"]
    ours = ScoreMatrix(M.astype(np.float32), genes)
    theirs = collections.defaultdict(lambda: collections.defaultdict(float))
    for i, g1 in enumerate(genes):
        for j, g2 in enumerate(genes):
            if i != j and M[i, j] != 0:
                theirs[g1][g2] = float(M[i, j])

    def build(scores, signed):
        ds = object.__new__(refd.GeneVectorDataset)
        ds.data = types.SimpleNamespace(genes=genes)
        ds.mi_scores, ds.signed_mi, ds.device = scores, signed, "cpu"
        ds._build_training_tensors(100.0)
        return ds._i_idx.numpy(), ds._j_idx.numpy(), ds._xij.numpy()

    for signed in (True, False):
        a, b = build(ours, signed), build(theirs, signed)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert np.abs(a[2] - b[2]).max() <= 100.0 ** 2 * 1e-6        # float32 storage of the 5-dp values
    # the reference's cache writer / reader on our Mapping, and key compatibility
    key = refc.compute_cache_key(z["X"], genes, "mi", {}, True)
    assert key == our_cache.compute_cache_key(z["X"], genes, "mi", {}, True)
    refc.save_scores(key, ours, genes)
    refc.save_scores(key + "t", theirs, genes)
    m1 = np.load(refc.get_cache_path(key))["scores"]
    m2 = np.load(refc.get_cache_path(key + "t"))["scores"]
    assert np.abs(m1 - m2).max() <= 1e-6
    back, g2 = refc.load_scores(key)
    assert g2 == genes and abs(back[genes[0]][genes[1]] - M[0, 1]) <= 1.1e-5
    # save_target_scores / load_target_scores (data.py:235-251): files are interchangeable
    from genevector_b200 import data as our_data
    monkeypatch.setattr(our_cache, "CACHE_DIR", str(tmp_path))
    ds = object.__new__(refd.GeneVectorDataset)
    ds.data, ds.mi_scores = types.SimpleNamespace(genes=genes), theirs
    ds.save_target_scores("run1/scores.npz")
    path_ours = our_data.save_target_scores(ours, genes, "run2/scores.npz")
    assert os.path.basename(path_ours) == "scores_run2_scores.npz"
    got = our_data.load_target_scores(refc.get_cache_path("run1_scores"))
    assert got.genes == genes and np.abs(got.rounded() - M).max() <= 1.1e-5
    ds.load_target_scores(path_ours)
    assert abs(ds.mi_scores[genes[0]][genes[1]] - M[0, 1]) <= 1.1e-5 and genes[0] not in ds.mi_scores[genes[0]]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_DIR, "genevector")),
                    reason="reference sources only exist in the build container")
def test_install_into_reference_replaces_the_python_loops_of_the_dataset(tmp_path, monkeypatch):
    """install_into_reference(metrics, data, cache) on the reference's own modules (build container
    only): GeneVectorDataset.get_gene_entropy / _build_training_tensors / load / save_target_scores and
    cache.save_scores / load_scores become the matrix fast paths, the originals stay reachable, and a
    nested dict a user side-loads still builds its tensors through the reference's own loop."""
    import collections
    import importlib
    import sys
    import types
    from genevector_b200 import cache as our_cache, data as our_data, metrics as ours
    from genevector_b200.scores import ScoreMatrix
    pkg = types.ModuleType("genevector")
    pkg.__path__ = [os.path.join(REF_DIR, "genevector")]
    monkeypatch.setitem(sys.modules, "genevector", pkg)
    for sub in ("genevector.data", "genevector.cache", "genevector.metrics", "genevector._aggregation",
                "genevector._graph_targets"):
        monkeypatch.delitem(sys.modules, sub, raising=False)
    refm = importlib.import_module("genevector.metrics")
    refd = importlib.import_module("genevector.data")
    refc = importlib.import_module("genevector.cache")
    D = refd.GeneVectorDataset
    orig_btt, orig_save = D.__dict__["_build_training_tensors"], refc.save_scores
    ours.install_into_reference(refm, refd, refc)
    assert refm.TARGETS["mi"] is ours.TARGETS["mi"]
    assert D.__dict__["_build_training_tensors"] is our_data.dataset_build_training_tensors
    assert D.__dict__["_build_training_tensors_reference"] is orig_btt
    assert D.get_gene_entropy is our_data.dataset_get_gene_entropy
    assert D.__dict__["load_target_scores"] is our_data.dataset_load_target_scores
    assert refc.save_scores is our_cache.save_scores and refc.save_scores_reference is orig_save
    assert refc.load_scores is our_cache.load_scores
    ours.install_into_reference(refm, refd, refc)                       # idempotent: originals are kept
    assert D.__dict__["_build_training_tensors_reference"] is orig_btt and refc.save_scores_reference is orig_save
    # a side-loaded nested dict (GeneVectorDataset.load_targets) still takes the reference's own loop
    genes = ["A", "B", "C"]
    d = collections.defaultdict(lambda: collections.defaultdict(float))
    d["A"]["B"] = d["B"]["A"] = 0.5
    ds = object.__new__(D)
    ds.data = types.SimpleNamespace(genes=genes)
    ds.mi_scores, ds.signed_mi, ds.device = d, True, "cpu"
    ds._build_training_tensors(10.0)
    assert ds._xij.tolist() == [50.0, 0.0, 50.0, 0.0, 0.0, 0.0]
    # the cache round trip of the installed functions on a ScoreMatrix, through the reference's own
    # _compute_target_scores protocol (data.py:300-334): key -> save -> load
    monkeypatch.setattr(our_cache, "CACHE_DIR", str(tmp_path))
    M = np.array([[0, .12345678, -.5], [.12345678, 0, .25], [-.5, .25, 0]], dtype=np.float32)
    key = refc.compute_cache_key(np.eye(3), genes, "mi", {}, True)
    refc.save_scores(key, ScoreMatrix(M, genes), genes)
    back, g2 = refc.load_scores(key)
    assert g2 == genes and isinstance(back, ScoreMatrix) and back["A"]["B"] == 0.12346 and back["A"]["C"] == -0.5


def test_count_packer_matches_numpy_and_refuses_what_does_not_fit():
    """gv_pack_counts_host (the staging pass of gv_upload_packed_counts, host to host): float32 counts ->
    uint8, int32 columns -> uint16, in order, for lengths around the 32-entry vector width; anything that
    is not an integer in [0, 255] / a column index in [0, 65535] makes it report "does not fit"."""
    import ctypes as C
    from genevector_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)

    def pack(v, c):
        n = len(v)
        dv = np.zeros(n + 64, dtype=np.uint8)
        dc = np.zeros(n + 64, dtype=np.uint16)
        ov = (-dv.ctypes.data) % 32
        oc = ((-dc.ctypes.data) % 32) // 2
        ok = np.zeros(1, dtype=np.int32)
        _lib.check("gv_pack_counts_host", lib.gv_pack_counts_host(
            v.ctypes.data, c.ctypes.data, n, dv.ctypes.data + ov, dc.ctypes.data + 2 * oc, ok.ctypes.data))
        return int(ok[0]), dv[ov:ov + n], dc[oc:oc + n]

    for n in (0, 1, 31, 32, 33, 95, 1000, 100003):
        v = rng.integers(0, 256, size=n).astype(np.float32)
        c = rng.integers(0, 65536, size=n).astype(np.int32)
        ok, dv, dc = pack(v, c)
        assert ok == 1
        assert np.array_equal(dv, v.astype(np.uint8)) and np.array_equal(dc, c.astype(np.uint16))
    base_v = rng.integers(0, 256, size=1000).astype(np.float32)
    base_c = rng.integers(0, 65536, size=1000).astype(np.int32)
    for pos in (0, 17, 500, 991, 999):
        for bad in (0.5, -1.0, 256.0, np.nan, np.inf, 3e9, -0.25):
            v = base_v.copy()
            v[pos] = bad
            assert pack(v, base_c)[0] == 0, (pos, bad)
        for badc in (65536, -1, 2 ** 31 - 1):
            c = base_c.copy()
            c[pos] = badc
            assert pack(base_v, c)[0] == 0, (pos, badc)


def test_rounding_conventions_agree():
    """The three places a score is rounded to 5 decimals -- the mapping (Python round, like the reference's
    round(mi, 5), metrics.py:139), ScoreMatrix.items() / rounded() / the cache writer (np.round) and
    gv_build_training_tensors (rint(x * 1e5) / 1e5 on the device) -- give the same float64 for every float32
    score: x * 1e5 is exact in float64 for a float32 x, so each is the correctly rounded decimal."""
    from genevector_b200.scores import ScoreMatrix
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.uniform(-3.5, 3.5, 200000).astype(np.float32),
                           (np.arange(-2000, 2000) * 1e-5 + 5e-6).astype(np.float32),      # near the ties
                           np.float32([0.5, 0.125, 1e-6, -1e-6, 0.000005, 2.5e-5, 3.4599998])])
    x64 = vals.astype(np.float64)
    a = np.array([round(float(v), 5) for v in x64])
    b = np.round(x64, 5)
    c = np.rint(x64 * 1e5) / 1e5
    assert np.array_equal(a, b) and np.array_equal(a, c)
    n = 300
    M = vals[:n * n].reshape(n, n).copy()
    M = ((M + M.T) / 2).astype(np.float32)
    np.fill_diagonal(M, 0)
    genes = [f"G{i}" for i in range(n)]
    sm = ScoreMatrix(M, genes)
    row = dict(sm["G7"].items())
    assert all(sm["G7"][g] == row[g] for g in genes if g != "G7")
    assert np.array_equal(sm.rounded()[7], np.array([0.0 if g == "G7" else sm["G7"][g] for g in genes]))


def test_shared_result_pool_slots_are_held_recycled_grown_and_bounded():
    """hostpool.SharedResultPool (one rank, not page-locked: no CUDA here): a result slot stays held while
    any numpy view of it is alive, is handed out again after the last view dies, is replaced by a larger
    segment when the next matrix does not fit, and at most eight results can be alive at once."""
    import gc
    import os
    from genevector_b200.hostpool import SharedResultPool
    base = f"/gvb200_test_{os.getpid()}"
    pool = SharedResultPool(0, 1, base, cuda=False, timeout_ms=2000)
    try:
        seq, slot, m = pool.begin(100 * 100 * 4)
        m.array(np.float32, (100, 100))[:] = 3.0
        pool.finish(seq, slot)
        a = pool.matrix(slot, 100)
        assert a.shape == (100, 100) and float(a.sum()) == 30000.0
        row = a[7]                                  # a derived view keeps the slot held
        del a
        gc.collect()
        seq2, slot2, _ = pool.begin(100 * 100 * 4)
        assert slot2 != slot, "a slot with a live view was handed out again"
        pool.finish(seq2, slot2)
        pool.release(slot2, pool.slots[slot2][0])   # nobody took a view of it
        assert float(row.sum()) == 300.0
        del row
        gc.collect()
        seq3, slot3, m3 = pool.begin(100 * 100 * 4)
        assert slot3 == slot, "the released slot is recycled"
        pool.finish(seq3, slot3)
        pool.release(slot3, pool.slots[slot3][0])
        gen_before = pool.slots[slot][0]
        seq4, slot4, m4 = pool.begin(300 * 300 * 4)  # does not fit: same slot, new generation
        assert slot4 == slot and pool.slots[slot][0] == gen_before + 1 and m4.nbytes >= 300 * 300 * 4
        pool.finish(seq4, slot4)
        held = [pool.matrix(slot4, 300)]
        for _ in range(7):
            sq, sl, _ = pool.begin(64)
            pool.finish(sq, sl)
            held.append(pool.matrix(sl, 4))
        with pytest.raises(RuntimeError, match="still referenced"):
            pool.begin(64)
        del held
        gc.collect()
        sq, sl, _ = pool.begin(64)                   # everything was released again
        pool.finish(sq, sl)
        assert not [f for f in os.listdir("/dev/shm") if f.startswith(base[1:] + ".s")], "result segments are unlinked"
    finally:
        pool.close()
    assert not [f for f in os.listdir("/dev/shm") if f.startswith(base[1:])]
