"""GPU tests of the host runtime around the kernels: pageable upload, canonical-CSR check, result
rectangles into one host matrix (the multi-rank assembly, ranks emulated one after another on one
GPU), the public call's shared-matrix path under a real process group when several GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_upload_pageable_is_byte_exact_while_the_stream_is_busy(engine):
    """Several arrays of many staging chunks each, uploaded back to back while earlier DMAs still sit
    behind queued GPU work: every staging slot is reused many times and must never be overwritten
    before its DMA has read it."""
    import torch
    from genevector_b200 import _lib
    rng = np.random.default_rng(0)
    arrays = [rng.integers(0, 2 ** 63 - 1, size=n, dtype=np.int64) for n in (9_000_011, 5_000_003, 12_345_678)]
    dev = engine.device
    busy = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        for threads in (1, 3, 8):
            outs = []
            for _ in range(20):
                busy.mul_(1.0001)                          # keeps the stream busy: DMAs queue up behind it
            for a in arrays:
                d = torch.empty(a.shape, dtype=torch.int64, device=dev)
                _lib.call("gv_upload_pageable", d.data_ptr(), a.ctypes.data, a.nbytes, threads, st)
                outs.append(d)
            torch.cuda.synchronize(dev)
            for a, d in zip(arrays, outs):
                assert np.array_equal(d.cpu().numpy(), a), f"upload with {threads} threads corrupted the data"


def test_csr_check_finds_unsorted_duplicate_and_out_of_range(engine):
    import scipy.sparse as sp
    import torch
    from genevector_b200 import _lib
    from genevector_b200.synth import synth_counts
    X = synth_counts(400, 60, seed=1)
    dev = engine.device

    def check(indices, indptr, n_genes=60):
        with torch.cuda.device(dev):
            c = torch.from_numpy(np.ascontiguousarray(indices, dtype=np.int32)).to(dev)
            p = torch.from_numpy(np.ascontiguousarray(indptr, dtype=np.int64)).to(dev)
            ok = np.zeros(1, dtype=np.int32)
            scratch = torch.zeros(1, dtype=torch.int32, device=dev)
            _lib.call("gv_csr_check", c.data_ptr(), _lib.GV_F32, p.data_ptr(), len(indptr) - 1, n_genes, scratch.data_ptr(),
                      ok.ctypes.data, torch.cuda.current_stream(dev).cuda_stream)
            return int(ok[0])

    assert check(X.indices, X.indptr) == 0
    row = int(np.flatnonzero(np.diff(X.indptr) >= 2)[0])
    e = X.indptr[row]
    swapped, sdata = X.indices.copy(), X.data.copy()
    swapped[e], swapped[e + 1] = swapped[e + 1], swapped[e]
    sdata[e], sdata[e + 1] = sdata[e + 1], sdata[e]              # the same matrix, one row unsorted
    assert check(swapped, X.indptr) == 1
    dup = X.indices.copy()
    dup[e + 1] = dup[e]
    assert check(dup, X.indptr) == 1
    oob = X.indices.copy()
    oob[e] = 60
    assert check(oob, X.indptr) == 2
    # the public path repairs case 1 on the host and refuses case 2
    from genevector_b200.metrics import discretize_genes
    Xs = sp.csr_matrix((sdata, swapped, X.indptr.copy()), shape=X.shape)
    got, nb = discretize_genes(Xs, 10)
    want, nb_w = discretize_genes(X, 10)
    assert np.array_equal(got, want) and np.array_equal(nb, nb_w)
    Xo = sp.csr_matrix(X.shape, dtype=np.float32)
    Xo.data, Xo.indices, Xo.indptr = X.data.copy(), oob, X.indptr.copy()
    with pytest.raises(ValueError):
        engine.upload_csr(Xo)


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("target", ["signed_mi", "pearson"])
def test_rank_rectangles_assemble_the_single_rank_matrix_bit_for_bit(engine, world, target):
    """Every emulated rank streams the rectangles of its own tiles into ONE host matrix (what the
    shared host matrix is under a process group); the result equals the single-rank matrix bit for
    bit and nothing is left untouched."""
    import torch
    from genevector_b200 import _lib
    from genevector_b200.engine import HostSink
    from genevector_b200.synth import synth_counts
    n_cells, n_genes = 1200, 2100
    X = synth_counts(n_cells, n_genes, seed=33)
    csr = engine.upload_csr(X)
    host = torch.full((n_genes, n_genes), float("nan"), dtype=torch.float32).pin_memory()
    if target == "signed_mi":
        full = engine.mi_from_csr(csr, n_cells, n_genes, signed=True)
    else:
        full = engine.target_from_csr(csr, _lib.GV_MOM_PEARSON)
    for rank in range(world):
        sink = HostSink(ptr=host.data_ptr(), pitch=n_genes * 4)
        if target == "signed_mi":
            engine.mi_from_csr(csr, n_cells, n_genes, signed=True, rank=rank, world=world, sink=sink)
        else:
            engine.target_from_csr(csr, _lib.GV_MOM_PEARSON, rank=rank, world=world, sink=sink)
    torch.cuda.synchronize()
    assert np.array_equal(host.numpy().view(np.uint32), full.cpu().numpy().view(np.uint32))


@pytest.mark.parametrize("world,signed", [(2, True), (3, False), (8, True)])
def test_sharded_front_end_kernels_reproduce_the_bin_block(engine, world, signed):
    """gv_shard_transpose / gv_shard_bins with the ranks emulated one after another on one GPU (the
    "peer" buffers are plain buffers of this device): every rank's dense rows, bins, marginals, edges
    and sums equal what gv_discretize_csr produces from the whole matrix."""
    import ctypes as C
    import torch
    from genevector_b200 import _lib
    from genevector_b200.engine import Engine, CsrDevice, block_layout, quantile_table, value_rows, limb_bits_for
    from genevector_b200.synth import synth_counts
    n_cells, n_genes, n_bins = 2999, 777, 10
    X = synth_counts(n_cells, n_genes, seed=44)
    csr = engine.upload_csr(X)
    want = engine.discretize(csr, n_bins, 0 if signed else -1)
    dev = engine.device
    kp, pitch, lay, total = block_layout(n_cells, n_genes, 0, 1)
    vrows = value_rows(n_genes, 1)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        raws = [torch.full((vrows, kp), 0x55, dtype=torch.uint8, device=dev) for _ in range(world)]
        hists = [torch.empty((n_genes, 256), dtype=torch.int32, device=dev) for _ in range(world)]
        fl = torch.zeros(64, dtype=torch.uint8, device=dev)
        dests = (C.c_void_p * world)(*[r.data_ptr() for r in raws])
        f0 = f1 = 0
        for rank in range(world):
            b0, b1, r0, r1 = Engine.shard_cells(n_cells, rank, world)
            e0, e1 = int(csr.row_ptr[r0]), int(csr.row_ptr[r1])
            loc = CsrDevice(csr.values[e0:e1], csr.col_idx[e0:e1], (csr.row_ptr[r0:r1 + 1] - e0).contiguous(),
                            r1 - r0, n_genes)
            flags = np.zeros(2, dtype=np.uint64)
            _lib.call("gv_shard_transpose", loc.values.data_ptr(), _lib.GV_F32, loc.col_idx.data_ptr(),
                      loc.row_ptr.data_ptr(), loc.n_cells, n_genes, b0, b1, dests, world, rank, kp,
                      hists[rank].data_ptr(), fl.data_ptr(), flags.ctypes.data, 3, st)
            f0 |= int(flags[0])
            f1 = max(f1, int(flags[1]))
        assert f0 == 0 and f1 == int(X.data.max())
        hp = (C.c_void_p * world)(*[h.data_ptr() for h in hists])
        lb = limb_bits_for(n_cells)
        q = quantile_table()
        for rank in range(world):
            bins = torch.empty((n_genes, pitch), dtype=torch.uint8, device=dev)
            nbpg = torch.empty(n_genes, dtype=torch.int32, device=dev)
            marg = torch.empty((n_genes, 32), dtype=torch.int32, device=dev)
            edges = torch.empty((n_genes, 32), dtype=torch.float64, device=dev)
            gsum = torch.empty(n_genes, dtype=torch.int64, device=dev)
            gsq = torch.empty_like(gsum)
            tot = torch.empty((n_genes, 256), dtype=torch.int32, device=dev)
            qd = torch.empty(8192, dtype=torch.uint8, device=dev)
            luts = torch.empty(n_genes * 256, dtype=torch.uint8, device=dev)
            _lib.call("gv_shard_bins", _lib.GV_F32, n_cells, n_genes, n_bins, q.ctypes.data, raws[rank].data_ptr(),
                      hp, world, tot.data_ptr(), qd.data_ptr(), luts.data_ptr(), bins.data_ptr(), pitch, nbpg.data_ptr(),
                      marg.data_ptr(), edges.data_ptr(), gsum.data_ptr(), gsq.data_ptr(), 1 if signed else 0,
                      vrows, lb, st)
            torch.cuda.synchronize(dev)
            assert torch.equal(bins, want.bins), f"rank {rank}: packed bins differ"
            assert torch.equal(nbpg, want.n_bins_per_gene) and torch.equal(marg, want.marg)
            assert torch.equal(edges.view(torch.int64), want.edges.view(torch.int64))
            assert torch.equal(gsum, want.gsum) and torch.equal(gsq, want.gsq)
            if signed:
                assert torch.equal(raws[rank].view(torch.int8), want.val_rows), f"rank {rank}: dense rows differ"
            else:
                dense = torch.from_numpy(np.asarray(X.todense()).T.astype(np.uint8)).to(dev)
                assert torch.equal(raws[rank][:n_genes, :n_cells], dense)
                assert not raws[rank][:n_genes, n_cells:].any()


@pytest.mark.parametrize("vmax", [9, 255])
def test_packed_count_upload_gives_the_same_bin_block(engine, vmax):
    """upload_csr(pack=True): uint8 values + uint16 columns on the device; the discretisation (bins,
    marginals, edges, value rows, sums) is bit-identical to the one from the float32 / int32 arrays.
    Data that are not counts are uploaded as they are."""
    import torch
    from genevector_b200.synth import synth_counts
    X = synth_counts(3000, 500, seed=51)
    if vmax == 255:
        X.data[::7] = 255.0
        X.data[3::11] = 200.0
    a = engine.upload_csr(X)
    b = engine.upload_csr(X, pack=True)
    assert b.packed and b.values.dtype == torch.uint8 and b.col_idx.dtype == torch.uint16 and not a.packed
    assert engine.last_upload_bytes == X.nnz * 3 + (X.shape[0] + 1) * 8
    for mode in (-1, 0, 1, 3):
        A, B = engine.discretize(a, 10, mode), engine.discretize(b, 10, mode)
        assert B.path_used == 1 and A.n_limbs == B.n_limbs and A.val_mode == B.val_mode
        for f in ("bins", "marg", "n_bins_per_gene", "gsum", "gsq"):
            assert torch.equal(getattr(A, f), getattr(B, f)), (mode, f)
        assert torch.equal(A.edges.view(torch.int64), B.edges.view(torch.int64))
        if mode >= 0:
            assert torch.equal(A.val_rows, B.val_rows), mode
    Xf = X.copy()
    Xf.data = np.log1p(Xf.data).astype(np.float32)
    assert not engine.upload_csr(Xf, pack=True).packed
    X64 = X.astype(np.float64)
    assert not engine.upload_csr(X64, pack=True).packed


@pytest.mark.parametrize("case", ["ref_poisson_500x50_b10", "ref_poisson_500x20_b5", "synth_counts_700x96_b10",
                                  "synth_counts_400x30_b15", "lognorm_f32_300x40_b7"])
def test_compute_mi_pairs_entry_matches_the_reference_goldens(engine, case):
    """gv_compute_mi_pairs -- the reference's FFI shape (src/lib.rs:9-15): the golden x_disc and
    n_bins_per_gene (outputs of the reference's discretize_genes) in, mi_raw (the reference's own MI) out
    within 1e-6; with a correlation matrix the Rust tier's sign rule (C port of src/lib.rs, oracle)."""
    from conftest import load_golden
    from genevector_b200.rust_compat import compute_mi_pairs, compute_mi_pairs_matrix
    from oracle import oracle as orc
    z = load_golden(case)
    xd, nb = z["x_disc"].astype(np.int32), z["n_bins_per_gene"].astype(np.int32)
    for operand in ("auto", "int8"):
        got, n_pairs = compute_mi_pairs_matrix(xd, nb, operand=operand)
        assert np.abs(got.astype(np.float64) - z["mi_raw"]).max() <= 1e-6, operand
        k = int((nb > 1).sum())
        assert n_pairs == k * (k - 1) // 2
    corr = orc.pearson_sign_np(z["X"]).astype(np.float32)
    want, n_want = orc.mi_pairs_c(xd, nb, corr=corr.astype(np.float64), sign_mode=2)
    got, n_pairs = compute_mi_pairs_matrix(xd, nb, corr)
    assert np.abs(got.astype(np.float64) - want).max() <= 1e-6 and n_pairs == n_want
    triples = compute_mi_pairs(xd, nb, corr)
    assert len(triples) == n_want and all(i < j for i, j, _ in triples)
    i, j, v = triples[len(triples) // 2]
    assert abs(v - want[i, j]) <= 1e-6
    bad = xd.copy()
    bad[3, 1] = nb[1]                                    # a bin index the gene cannot have
    with pytest.raises(Exception, match="bin index"):
        compute_mi_pairs_matrix(bad, nb)


def test_reference_shaped_dataset_flow_at_5000_genes(engine, tmp_path, monkeypatch):
    """The flow of GeneVectorDataset.create_inputs_outputs (data.py:336-375) with the functions
    install_into_reference binds onto the reference's classes, at 5,000 genes x 50,000 cells (C3): gene
    entropy -> target scores through the registry and the cache protocol -> training tensors -> cache hit
    -> .npz round trip, in seconds and without a Python loop over gene pairs."""
    import time
    import types
    import torch
    from genevector_b200 import cache as our_cache, data as our_data
    from genevector_b200.scores import ScoreMatrix
    from genevector_b200.synth import synth_counts_torch, gene_names
    import scipy.sparse as sp
    monkeypatch.setattr(our_cache, "CACHE_DIR", str(tmp_path))
    G, N = 5000, 50000
    v, c, p = synth_counts_torch(N, G, seed=2, device=engine.device)
    X = sp.csr_matrix((v.cpu().numpy(), c.cpu().numpy(), p.cpu().numpy()), shape=(N, G))
    genes = gene_names(G)
    adata = types.SimpleNamespace(X=X, var=types.SimpleNamespace(index=types.SimpleNamespace(tolist=lambda: genes)))
    ds = types.SimpleNamespace(adata=adata, data=types.SimpleNamespace(genes=genes), device="cuda", signed_mi=True,
                               mi_scores=None)
    t0 = time.perf_counter()
    ent = our_data.dataset_get_gene_entropy(adata)                                   # data.py:349
    assert len(ent) == G and all(np.isfinite(e) for e in list(ent.values())[:50])
    ds.mi_scores = our_data.compute_target_scores(X, genes, target="mi", signed_mi=True, mi_backend="b200",
                                                  device="cuda", use_cache=True)     # data.py:353-354
    assert isinstance(ds.mi_scores, ScoreMatrix)
    our_data.dataset_build_training_tensors(ds, c=100.0)                             # data.py:372
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert ds._xij.shape == (G * (G - 1),) and ds._i_idx.dtype == torch.int64 and ds._xij.is_cuda
    # spot check against the mapping semantics the reference's loop reads (data.py:383-390)
    for k in (0, 12345, G * (G - 1) - 1):
        i, j = int(ds._i_idx[k]), int(ds._j_idx[k])
        assert abs(float(ds._xij[k]) - np.float32(ds.mi_scores[genes[i]][genes[j]]) * np.float32(1e4)) <= 1e-2
    # second call: served by the cache file the first one wrote (same key as the reference's)
    t1 = time.perf_counter()
    again = our_data.compute_target_scores(X, genes, target="mi", signed_mi=True, mi_backend="b200",
                                           device="cuda", use_cache=True)
    dt_hit = time.perf_counter() - t1
    assert np.abs(again.matrix - ds.mi_scores.rounded()).max() <= 1e-6
    our_data.dataset_save_target_scores(ds, "run/scores.npz")
    ds2 = types.SimpleNamespace(mi_scores=None)
    our_data.dataset_load_target_scores(ds2, our_cache.get_cache_path("run_scores"))
    assert np.array_equal(ds2.mi_scores.matrix, again.matrix)
    print(f"entropy + scores + cache save + training tensors: {dt:.2f} s; cache hit: {dt_hit:.2f} s")
    assert dt < 60.0


def test_public_call_reuses_its_pinned_result_buffers(engine):
    """compute_mi_b200 twice: results are independent objects (the first stays valid after the second)."""
    from genevector_b200.metrics import compute_mi_b200
    from genevector_b200.synth import synth_counts, gene_names
    Xa, Xb = synth_counts(600, 300, seed=1), synth_counts(600, 300, seed=2)
    genes = gene_names(300)
    a = compute_mi_b200(Xa, genes, signed=True)
    a_copy = a.matrix.copy()
    b = compute_mi_b200(Xb, genes, signed=True)
    assert np.array_equal(a.matrix, a_copy) and not np.array_equal(a.matrix, b.matrix)


def test_sharded_public_calls_on_all_visible_gpus():
    """The real thing: one process per visible GPU (torch.distributed.run, NCCL), signed MI / Pearson /
    Spearman through the public calls, every rank's full matrix compared bitwise with the single-rank
    result (tools/dist_check.py).  Skipped on a one-GPU box."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    env = dict(os.environ, GV_FRONT_CHUNK_NNZ="20000")       # small matrices too go up in several chunks
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29731",
                        os.path.join(ROOT, "tools", "dist_check.py")], capture_output=True, text=True, timeout=900,
                       env=env)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"dist_check_world{n}.log"), "w") as f:
        f.write(r.stdout + "\n---- stderr ----\n" + r.stderr[-20000:])
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "DIST CHECK OK" in r.stdout
