"""genevector_b200.metrics -- the reference's operator interface for the co-expression path,
served by the B200 engine.

Mirrors genevector/metrics.py of nceglia/genevector for the part on the hot path: same names,
same argument meaning, same error behaviour, so that `GeneVectorDataset(mi_backend="b200")`
(data.py:320-326 forwards `backend=`) and the reference's own tests read the same:

    discretize_genes(X, n_bins)                         metrics.py:25-69
    compute_mi_b200(X, gene_names, n_bins, signed)      new tier next to compute_mi_rust (:312-339)
    target_mi(X, gene_names, signed, backend, device, n_bins)     :381-411
    target_pearson / target_cosine / target_jaccard     :414-420, :452-461, :434-449
    target_spearman                                     :423-431
    target_graph_xcorr                                  genevector/_graph_targets.py:10-55
    TARGETS / register_target / get_target_function     :344-376

There is one backend and no fallback: backend must be "b200" (or "auto", which here can only
mean "b200"); the reference's CPU tier names raise the same ValueError an unknown name raises
in the reference (metrics.py:410-411).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .engine import get_engine
from .scores import ScoreMatrix

TARGETS = {}


def register_target(name):
    """Decorator to register a target function (metrics.py:347-352)."""
    def wrapper(fn):
        TARGETS[name] = fn
        return fn
    return wrapper


def get_target_function(name):
    """Look up a registered target (metrics.py:355-376); ValueError if unknown."""
    if name not in TARGETS:
        available = ", ".join(sorted(TARGETS.keys()))
        raise ValueError(f"Unknown target '{name}'. Available: {available}")
    return TARGETS[name]


def _device_of(device):
    if device in (None, "cpu", "auto"):
        return None            # the engine's current CUDA device; "cpu" cannot be honoured
    return device


def discretize_genes(X, n_bins=10, device=None):
    """GPU discretisation with the reference's return convention (metrics.py:25-69):
    (X_disc int32 [n_cells, n_genes], n_bins_per_gene int32 [n_genes])."""
    eng = get_engine(_device_of(device))
    csr = eng.upload_csr(X)
    blk = eng.discretize(csr, n_bins)
    return blk.x_disc(), blk.n_bins_per_gene.cpu().numpy()


def _dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def _pinned_result(n_genes):
    import torch
    return lambda: torch.empty((n_genes, n_genes), dtype=torch.float32, pin_memory=True)


def _run_to_host(eng, n_genes, gene_names, run, front=None):
    """Shared body of the public calls: run(sink) enqueues the whole path on this rank with its
    result rectangles streaming into the host matrix behind `sink`; returns the ScoreMatrix.

    One rank: a pinned matrix (allocated while the GPU already computes).  Several ranks (one process
    per GPU, torch.distributed initialised): the one shared, page-locked host matrix of the box
    (hostpool.py); every rank copies the rectangles of its own tiles into it and, after a per-rank
    completion flag, every rank returns the full matrix.  No reduce, no gather."""
    import time
    rank, world = _dist_info()
    t0 = time.perf_counter()
    eng.defer_release = True
    try:
        return _run_to_host_inner(eng, n_genes, gene_names, run, rank, world, t0)
    finally:
        eng.defer_release = False
        eng.release_uploads()


def _run_to_host_inner(eng, n_genes, gene_names, run, rank, world, t0):
    import time
    import torch
    from .engine import HostSink
    if world == 1:
        run(HostSink(tensor=_pinned_result(n_genes)))
        eng._timed_host("enqueue (upload + launches + host buffer)", t0)
        t1 = time.perf_counter()
        eng.release_uploads()
        torch.cuda.synchronize(eng.device)
        eng._timed_host("wait for the device", t1)
        host, eng.last_host_result = eng.last_host_result, None
        return ScoreMatrix(host.numpy(), gene_names)
    from .hostpool import get_pool
    pool = get_pool()
    seq, slot, mapping = pool.begin(n_genes * n_genes * 4)
    eng._timed_host("agree on the shared host matrix", t0)
    t1 = time.perf_counter()
    try:
        run(HostSink(ptr=mapping.ptr, pitch=n_genes * 4))
        eng._timed_host("enqueue (upload + launches)", t1)
        t2 = time.perf_counter()
        eng.release_uploads()
        torch.cuda.synchronize(eng.device)
        eng._timed_host("wait for the device", t2)
    except BaseException:
        eng.release_uploads()
        pool.release(slot, pool.slots[slot][0])
        raise
    t3 = time.perf_counter()
    pool.finish(seq, slot)
    eng._timed_host("wait for the other ranks", t3)
    return ScoreMatrix(pool.matrix(slot, n_genes), gene_names)


def compute_mi_b200(X, gene_names, n_bins=10, signed=False, device=None, sign_rule="numpy", front="auto"):
    """All-pairs (signed) mutual information on B200; returns a ScoreMatrix (Mapping view with
    the reference's dict semantics, see scores.py).

    Under torch.distributed (one process per GPU of one box, every rank calling with the same X) the
    tile space is split over the ranks; every rank writes its own result tiles into one shared host
    matrix and every rank returns the full ScoreMatrix (see _run_to_host).

    front: how the discretised matrix reaches every rank.  "broadcast": rank 0 uploads and
    discretises, one NCCL broadcast replicates the bin block (north_star's prescription).  "sharded":
    every rank uploads its own block of cells over its own PCIe link and transposes it straight into
    every peer's memory over NVLink (peer stores, no NCCL call), then bins all genes locally.  "auto":
    sharded when the data qualify (count data on one box), else broadcast.

    sign_rule: "numpy" (default) signs with np.sign of the Pearson correlation like the reference's
    NumPy / Numba / torch tiers and its documentation (metrics.py:134, 227, 302): a zero or undefined
    correlation zeroes the score; "rust" reproduces the Rust tier (src/lib.rs:91-94): +1 unless the
    correlation is negative."""
    eng = get_engine(_device_of(device))
    rank, world = _dist_info()
    n_cells, n_genes = X.shape
    if len(gene_names) != n_genes:
        raise ValueError("gene_names must have one entry per column of X")
    if sign_rule not in ("numpy", "rust"):
        raise ValueError(f"unknown sign rule {sign_rule!r} (numpy or rust)")

    def run(sink):
        if world > 1 and front != "broadcast":
            done = eng.mi_from_host_sharded(X, n_bins=n_bins, signed=signed, rank=rank, world=world,
                                            sign_rule=sign_rule, sink=sink, required=(front == "sharded"))
            if done:
                return
        csr = eng.upload_csr(X, pack=True) if rank == 0 else None
        eng.mi_from_csr(csr, n_cells, n_genes, n_bins=n_bins, signed=signed, rank=rank, world=world,
                        sign_rule=sign_rule, sink=sink)

    return _run_to_host(eng, n_genes, gene_names, run)


@register_target("mi")
def target_mi(X, gene_names, signed=False, backend="b200", device="cuda", n_bins=10, **kwargs):
    """Mutual information, optionally signed by the Pearson correlation (metrics.py:381-411).
    target_kwargs={"sign_rule": "rust"} selects the Rust tier's sign convention."""
    if backend in ("b200", "auto"):
        return compute_mi_b200(X, gene_names, n_bins=n_bins, signed=signed, device=device,
                               sign_rule=kwargs.get("sign_rule", "numpy"), front=kwargs.get("front", "auto"))
    raise ValueError(f"Unknown MI backend: {backend}")


def _moment_target(X, gene_names, kind, device=None, ranks=False):
    """Shared body of the matrix targets; under torch.distributed the tile space is split over the
    ranks like compute_mi_b200 (one broadcast of the value rows, result rectangles into the shared
    host matrix)."""
    eng = get_engine(_device_of(device))
    rank, world = _dist_info()
    n_cells, n_genes = X.shape
    if len(gene_names) != n_genes:
        raise ValueError("gene_names must have one entry per column of X")

    def run(sink):
        if world > 1 and not ranks and eng.target_from_host_sharded(X, kind, rank=rank, world=world, sink=sink):
            return
        csr = eng.upload_csr(X, pack=True) if rank == 0 else None
        eng.target_from_csr(csr, kind, ranks=ranks, n_cells=n_cells, n_genes=n_genes, rank=rank, world=world,
                            sink=sink)

    return _run_to_host(eng, n_genes, gene_names, run)


@register_target("pearson")
def target_pearson(X, gene_names, device=None, **kwargs):
    """Pearson correlation between all gene pairs (metrics.py:414-420)."""
    return _moment_target(X, gene_names, _lib.GV_MOM_PEARSON, device)


@register_target("spearman")
def target_spearman(X, gene_names, device=None, **kwargs):
    """Spearman rank correlation between all gene pairs (metrics.py:423-431): the Pearson
    correlation of the per-gene average ranks, computed from exact integer rank co-moments."""
    return _moment_target(X, gene_names, _lib.GV_MOM_PEARSON, device, ranks=True)


@register_target("jaccard")
def target_jaccard(X, gene_names, device=None, **kwargs):
    """Jaccard index on detected / not detected (metrics.py:434-449)."""
    return _moment_target(X, gene_names, _lib.GV_MOM_JACCARD, device)


@register_target("cosine")
def target_cosine(X, gene_names, device=None, **kwargs):
    """Cosine similarity between gene vectors (metrics.py:452-461)."""
    return _moment_target(X, gene_names, _lib.GV_MOM_COSINE, device)


@register_target("graph_xcorr")
def target_graph_xcorr(X, gene_names, graph=None, aggr="mean", aggr_params=None, device=None, **kwargs):
    """Cross-correlation between self-expression and graph-neighbour-aggregated expression, symmetrised
    (genevector/_graph_targets.py:10-55).  The built-in "mean" aggregation (with its include_self
    option, _aggregation.py:88-109) runs on the device; host callables cannot."""
    import torch
    if graph is None:
        raise ValueError(
            "graph required. Pass any scipy sparse adjacency matrix "
            "via target_kwargs={'graph': G}"
        )
    if callable(aggr):
        raise ValueError("the b200 backend runs the built-in 'mean' aggregation on the device; "
                         "callable aggregations are host functions and are not supported")
    if aggr != "mean":
        raise ValueError(f"Unknown aggregation '{aggr}'. Available: mean")
    params = dict(aggr_params or {})
    include_self = bool(params.pop("include_self", False))
    eng = get_engine(_device_of(device))
    if len(gene_names) != X.shape[1]:
        raise ValueError("gene_names must have one entry per column of X")
    csr = eng.upload_csr(X)
    out = eng.graph_xcorr(csr, eng.upload_graph(graph, X.shape[0]), include_self=include_self)
    torch.cuda.synchronize(eng.device)
    return ScoreMatrix(out.cpu().numpy(), gene_names)


def install_into_reference(ref_metrics_module, ref_data_module=None, ref_cache_module=None):
    """Plug the B200 tier into an imported reference package.

    ref_metrics_module (genevector.metrics): the targets are overridden through the reference's own
    register_target hook.
    ref_data_module (genevector.data) / ref_cache_module (genevector.cache), optional: the O(G^2) Python
    around the targets is replaced by the matrix fast paths, so that the reference's own
    GeneVectorDataset.create_inputs_outputs (data.py:336-375) runs at 5,000+ genes --
      GeneVectorDataset.get_gene_entropy      data.py:186-203  (densifies X)   -> gv_gene_entropy
      GeneVectorDataset._build_training_tensors  data.py:377-411 (double loop) -> gv_build_training_tensors
      GeneVectorDataset.load_target_scores / save_target_scores  data.py:235-251
      cache.save_scores / cache.load_scores   cache.py:64-110 (double loops)   -> array operations
    The originals stay reachable as <name>_reference; a nested dict handed in by the user (load_targets)
    still goes through them."""
    for name in ("mi", "pearson", "spearman", "cosine", "jaccard", "graph_xcorr"):
        ref_metrics_module.register_target(name)(TARGETS[name])
    if ref_cache_module is not None:
        from . import cache as fast_cache
        for name in ("save_scores", "load_scores"):
            if not hasattr(ref_cache_module, name + "_reference"):
                setattr(ref_cache_module, name + "_reference", getattr(ref_cache_module, name))
            setattr(ref_cache_module, name, getattr(fast_cache, name))
    if ref_data_module is not None:
        from . import data as fast_data
        D = ref_data_module.GeneVectorDataset
        for name in ("get_gene_entropy", "_build_training_tensors", "load_target_scores", "save_target_scores"):
            if name + "_reference" not in D.__dict__:
                setattr(D, name + "_reference", D.__dict__[name])
        D.get_gene_entropy = staticmethod(fast_data.dataset_get_gene_entropy)
        D._build_training_tensors = fast_data.dataset_build_training_tensors
        D.load_target_scores = fast_data.dataset_load_target_scores
        D.save_target_scores = fast_data.dataset_save_target_scores
    return ref_metrics_module
