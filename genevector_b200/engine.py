"""Host orchestration of the B200 co-expression engine.

Python stays the host language (as in the reference); PyTorch tensors are used only as device
buffers and for the NCCL broadcast; every compute step is a call into libgvb200.so
(include/gvb200.h).  No step has a CPU or PyTorch fallback.

Pipeline for all-pairs (signed) MI on one rank:
    CSR (pageable host arrays) --gv_upload_packed_counts / gv_upload_pageable--> device
                        gv_discretize_csr  -> packed bins + marginals (+ INT8 value rows)
                        gv_moment_tiles    -> int8 Pearson-sign matrix           (signed only)
                        gv_expand_onehot   -> FP4 / INT8 one-hot operand (K-major)
                        gv_mi_tiles        -> float32 score matrix [G][G], launched in a few pieces whose
                                              result rectangles stream to the host matrix (gv_copy_rect_d2h)
With several ranks (one process per GPU of one box) every rank takes a contiguous slice of the
upper-triangular tile schedule and writes its result rectangles into the one shared host matrix
(hostpool.py).  The discretised matrix reaches every rank by one of two front ends: "broadcast"
(rank 0 discretises, one NCCL broadcast of the bin block) or "sharded" (every rank uploads and
transposes its own block of cells, the slabs are exchanged by peer copies over NVLink, every rank bins
all genes locally: gv_shard_transpose / gv_shard_bins, no NCCL call).
"""
from __future__ import annotations

import dataclasses
import os
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import GvError  # noqa: F401  (re-export)

_A = 256  # byte alignment of the sections of the broadcast block


def _align(n: int, a: int = _A) -> int:
    return (n + a - 1) // a * a


def quantile_table() -> np.ndarray:
    """numpy's own quantile fractions (metrics.py:64 + np.percentile's division by 100)."""
    q = np.zeros((_lib.GV_QTABLE_DIM, _lib.GV_QTABLE_DIM), dtype=np.float64)
    for k in range(2, _lib.GV_MAX_BINS + 1):
        q[k, :k - 1] = np.linspace(0, 100, k + 1)[1:-1] / 100
    return q


def limb_bits_for(n_cells: int) -> int:
    """Largest digit width (<= 7) whose squared maximum summed over all cells fits in uint32."""
    for lb in range(7, 0, -1):
        if n_cells * ((1 << lb) - 1) ** 2 < (1 << 32):
            return lb
    raise GvError("limb_bits_for", _lib.GV_OK - 4, f"n_cells={n_cells} too large for the exact co-moment path")


def value_rows(n_genes: int, n_limbs: int) -> int:
    """Operand rows of the value matrix: 32-row blocks of 32 // n_limbs genes (gv_value_rows)."""
    gpb = 32 // n_limbs
    return (n_genes + gpb - 1) // gpb * 32


def max_limbs_for(n_cells: int, limb_bits: int) -> int:
    """Most digits per gene (<= 4) for which Sum x_i*x_j over all cells still fits 63 bits."""
    tb_max = (63 - max(int(n_cells), 1).bit_length()) // 2
    return max(0, min(4, tb_max // limb_bits))


@dataclasses.dataclass
class CsrDevice:
    """A cells x genes CSR matrix resident on the device."""
    values: torch.Tensor     # float32 / float64 [nnz]   (uint8 when packed)
    col_idx: torch.Tensor    # int32 [nnz]               (uint16 when packed)
    row_ptr: torch.Tensor    # int64 [n_cells + 1]
    n_cells: int
    n_genes: int
    packed: bool = False     # count data packed on the way up (gv_upload_packed_counts)

    @property
    def dtype_code(self) -> int:
        if self.packed:
            return _lib.GV_COUNTS_U8_COL16
        return _lib.GV_F32 if self.values.dtype == torch.float32 else _lib.GV_F64

    @property
    def nnz(self) -> int:
        return int(self.values.numel())


@dataclasses.dataclass
class BinBlock:
    """Result of the discretisation: one contiguous device buffer (the broadcast payload) with
    typed views into it."""
    buf: torch.Tensor            # uint8, everything below lives in here
    n_cells: int
    n_genes: int
    n_bins: int
    kp: int                      # K extent padded to 128
    bins_pitch: int              # bytes per gene row of the bins: kp / 2 (two cells per byte) or kp (n_bins > 15)
    bins: torch.Tensor           # uint8 [G][bins_pitch]
    marg: torch.Tensor           # int32 [G][32]
    n_bins_per_gene: torch.Tensor  # int32 [G]
    edges: torch.Tensor          # float64 [G][32]
    gsum: torch.Tensor           # int64 [G]
    gsq: torch.Tensor            # int64 [G]
    val_rows: Optional[torch.Tensor]   # int8 [value_rows(G, n_limbs)][kp] or None
    val_mode: int                # -1 none, 0 count digits, 1 detection indicator, 2 fixed point, 3 ranks
    n_limbs: int                 # INT8 rows per gene in val_rows (1..4)
    limb_bits: int
    data_flags: tuple            # (flag bits, max value)
    path_used: int = 0           # gv_disc_path the discretisation took (1 counts, 2 general)
    pending: object = None       # handle of a value-row broadcast still in flight (Engine._replicate)

    def gene_values(self) -> np.ndarray:
        """The integers behind the value rows, int64 [n_genes][n_cells] (tests)."""
        vr = self.val_rows.cpu().numpy().astype(np.int64)
        L, gpb = self.n_limbs, 32 // self.n_limbs
        g = np.arange(self.n_genes)
        row0 = (g // gpb) * 32 + (g % gpb) * L
        out = np.zeros((self.n_genes, self.n_cells), dtype=np.int64)
        for d in range(L):
            out += vr[row0 + d, :self.n_cells] << (d * self.limb_bits)
        return out

    def x_disc(self) -> np.ndarray:
        """Unpack to the reference's layout: int32 [n_cells][n_genes] (tests / small cases)."""
        b = self.bins.cpu().numpy()
        if self.n_bins > 15:                      # one byte per cell
            return np.ascontiguousarray(b[:, :self.n_cells].T.astype(np.int32))
        lo, hi = b & 0x0F, b >> 4
        full = np.empty((self.n_genes, b.shape[1] * 2), dtype=np.int32)
        full[:, 0::2], full[:, 1::2] = lo, hi
        return np.ascontiguousarray(full[:, :self.n_cells].T)


HEADER_BYTES = 256   # int64[0] = n_limbs, [1] = limb_bits, [2] = data flag bits, [3] = max value, [4] = val_mode


def block_layout(n_cells: int, n_genes: int, val_mode: int, n_limbs: int = 1, n_bins: int = 10):
    """Byte offsets of the sections of the broadcast block.  The value rows come last, so a
    two-digit block is the one-digit block plus a tail.  Bins take half a byte per cell, a whole one
    when n_bins > 15."""
    kp = _align(max(n_cells, 1), 256)      # a 128-byte K slice of the FP4 operand is 256 cells
    pitch = kp if n_bins > 15 else kp // 2
    ms, es = _lib.GV_MARG_STRIDE, _lib.GV_EDGE_STRIDE
    off, lay = 0, {}
    for name, nbytes in (("header", HEADER_BYTES), ("bins", n_genes * pitch),
                         ("marg", n_genes * ms * 4), ("nbpg", n_genes * 4),
                         ("edges", n_genes * es * 8), ("gsum", n_genes * 8), ("gsq", n_genes * 8),
                         ("val", (value_rows(n_genes, n_limbs) * kp) if val_mode >= 0 else 0)):
        lay[name] = (off, nbytes)
        off += _align(nbytes)
    return kp, pitch, lay, off


def _views(buf, n_cells, n_genes, n_bins, val_mode, n_limbs, limb_bits, flags=(0, 0)) -> BinBlock:
    kp, pitch, lay, _ = block_layout(n_cells, n_genes, val_mode, n_limbs, n_bins)

    def v(name, dtype, shape):
        o, nb = lay[name]
        return buf[o:o + nb].view(dtype).view(*shape)

    return BinBlock(
        buf=buf, n_cells=n_cells, n_genes=n_genes, n_bins=n_bins, kp=kp, bins_pitch=pitch,
        bins=v("bins", torch.uint8, (n_genes, pitch)),
        marg=v("marg", torch.int32, (n_genes, _lib.GV_MARG_STRIDE)),
        n_bins_per_gene=v("nbpg", torch.int32, (n_genes,)),
        edges=v("edges", torch.float64, (n_genes, _lib.GV_EDGE_STRIDE)), gsum=v("gsum", torch.int64, (n_genes,)),
        gsq=v("gsq", torch.int64, (n_genes,)),
        val_rows=v("val", torch.int8, (value_rows(n_genes, n_limbs), kp)) if val_mode >= 0 else None,
        val_mode=val_mode, n_limbs=n_limbs, limb_bits=limb_bits, data_flags=flags)


_host_schedules = {}


def schedule_host(op_rows: int, cta_group: int, band: int):
    """The whole band-ordered upper-triangular tile schedule (host, int32 [n,2]) and tile_m."""
    key = (op_rows, cta_group, band)
    if key not in _host_schedules:
        lib = _lib.load()
        tile_m = _lib.call("gv_tile_m", cta_group)
        n = _lib.check("gv_tile_schedule", lib.gv_tile_schedule(op_rows, tile_m, band, None, 0))
        host = np.zeros((max(n, 1), 2), dtype=np.int32)
        _lib.check("gv_tile_schedule", lib.gv_tile_schedule(op_rows, tile_m, band, host.ctypes.data, n))
        _host_schedules[key] = (host[:n], tile_m)
    return _host_schedules[key]


def _cut_points(th: np.ndarray, tile_m: int, band: int):
    """Tile indices at which the schedule may be cut between ranks (or between host-copy pieces):
    the first tile of a band, or the first tile of a tile column that lies entirely to the right of
    the band's own rows.  The tiles of a band's diagonal block (columns overlapping the band's rows)
    so stay with one owner, and what an owner writes is a set of full rectangles of the result
    matrix: (band rows) x (its columns) and the mirror image."""
    n = len(th)
    if n == 0:
        return np.zeros(1, dtype=np.int64), np.zeros(0, dtype=np.int64)
    row0, col0 = th[:, 0].astype(np.int64), th[:, 1].astype(np.int64)
    band_id = row0 // (tile_m * band)
    band_end = (band_id + 1) * tile_m * band
    first_of_band = np.diff(band_id, prepend=-1) != 0
    first_of_col = first_of_band | (np.diff(col0, prepend=-1) != 0)
    ok = first_of_band | (first_of_col & (col0 >= band_end))
    return np.append(np.flatnonzero(ok), n), band_id


def rank_bounds(th: np.ndarray, tile_m: int, band: int, world: int):
    """world + 1 tile indices: rank r owns tiles [b[r], b[r+1]).  Equal shares snapped to the nearest
    allowed cut (a tile column is at most `band` tiles, so the shares differ by a few tiles)."""
    n = len(th)
    cuts, _ = _cut_points(th, tile_m, band)
    b = [0]
    for r in range(1, world):
        want = n * r // world
        c = int(cuts[np.abs(cuts - want).argmin()])
        b.append(max(c, b[-1]))
    b.append(n)
    return b


def tile_slice(op_rows: int, cta_group: int, band: int, rank: int = 0, world: int = 1):
    """This rank's contiguous slice of the band-ordered upper-triangular tile schedule
    (host, int32 [n,2]) and the total tile count.  Tiles cost the same (same K), so equal counts
    are equal work; the slices are disjoint and cover the schedule."""
    th, tile_m = schedule_host(op_rows, cta_group, band)
    b = rank_bounds(th, tile_m, band, world)
    return np.ascontiguousarray(th[b[rank]:b[rank + 1]]), len(th)


def host_pieces(op_rows: int, cta_group: int, band: int, genes_per_block: int, n_genes: int,
                rank: int = 0, world: int = 1, chunks: int = 8):
    """How a rank's slice of the schedule is launched so that its results can stream to the host
    under the computation: a list of (t0, t1, rects) with t0 / t1 tile offsets inside the slice and
    rects = [(row0, row1, col0, col1), ...] the gene rectangles of the result matrix that are final
    once tiles [0, t1) of the slice have run and that no other piece (or rank) writes.

    world == 1: pieces end at band boundaries and a tile of band b only writes rows of bands >= b, so
    the rows of the finished bands are final: one contiguous block of full rows per piece.
    world > 1: a piece is a run of whole tile columns of bands (see _cut_points); per band it owns
    the rectangle (band rows) x (its columns) and the mirror image below the band.

    Pieces halve in size (1/2, 1/4, ... of the slice): the copy of a piece runs under the next piece's
    kernel, only the last piece's copy is exposed, and late bands are short, so an equal split would
    leave ~40 % of the matrix rows to the exposed copy; the halving split leaves a few per cent."""
    th, tile_m = schedule_host(op_rows, cta_group, band)
    b = rank_bounds(th, tile_m, band, world)
    lo, hi = b[rank], b[rank + 1]
    n = hi - lo
    if n == 0:
        return []
    cuts, band_id = _cut_points(th, tile_m, band)
    gpb = genes_per_block

    def genes(rows):
        return min(n_genes, int(rows) // 32 * gpb)

    if world == 1:
        starts = np.flatnonzero(np.diff(band_id, prepend=-1))
        inner = sorted({int(starts[np.abs(starts - (n - (n >> k))).argmin()]) for k in range(1, chunks)} - {0, n})
        ends = inner + [n]
        out, t0, g0 = [], 0, 0
        for t1 in ends:
            g1 = n_genes if t1 == n else genes(th[t1, 0])
            out.append((t0, t1, [(g0, g1, 0, n_genes)] if g1 > g0 else []))
            t0, g0 = t1, g1
        return out
    inside = cuts[(cuts > lo) & (cuts < hi)]
    inner = sorted({int(inside[np.abs(inside - (hi - (n >> k))).argmin()])
                    for k in range(1, chunks)}) if len(inside) else []
    ends = inner + [hi]
    out, t0 = [], lo
    for t1 in ends:
        if t1 <= t0:
            continue
        rects = []
        seg = th[t0:t1].astype(np.int64)
        bid = band_id[t0:t1]
        for bnd in np.unique(bid):
            cols = seg[bid == bnd, 1]
            r0, r1 = genes(bnd * band * tile_m), genes((bnd + 1) * band * tile_m)
            c0, c1 = max(genes(cols.min()), r0), genes(cols.max() + 256)
            if r1 > r0 and c1 > c0:
                rects.append((r0, r1, c0, c1))
                m0 = max(c0, r1)
                if c1 > m0:
                    rects.append((m0, c1, r0, r1))
        out.append((t0 - lo, t1 - lo, rects))
        t0 = t1
    return out


def broadcast_block(buf: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """The single collective of the path: replicate rank 0's bin block (NCCL over NVLink on
    GPUs; any torch.distributed backend works, the CPU tests use gloo)."""
    import torch.distributed as dist
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(buf[:total], src=src, group=group)
    return buf


class _DevPtr:
    """A device allocation made outside torch (gv_ipc_alloc), exposed through
    __cuda_array_interface__ so that torch can view it without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerBlock:
    """The bin block of one (n_cells, n_genes) shape in peer-mapped device memory: this rank's own
    allocation plus the mappings of every other rank's (same layout), for the sharded front end."""

    def __init__(self, own_ptr, total, ptrs, offsets, tensor):
        self.own_ptr, self.total, self.ptrs, self.off, self.tensor = own_ptr, total, ptrs, offsets, tensor


class HostSink:
    """Where a rank's result rectangles go on the host: a page-locked float32 matrix given as
    (address, bytes per row), as a pinned tensor, or as a callable returning a pinned tensor (called
    after the kernels are enqueued, so its allocation overlaps the computation)."""

    def __init__(self, ptr: int = 0, pitch: int = 0, tensor=None):
        self.ptr, self.pitch, self.tensor = ptr, pitch, tensor

    def resolve(self, engine=None):
        if self.tensor is not None:
            t = self.tensor() if callable(self.tensor) else self.tensor
            self.tensor = t
            if engine is not None:
                engine.last_host_result = t
            self.ptr, self.pitch = t.data_ptr(), t.stride(0) * t.element_size()
        return self.ptr, self.pitch


class Engine:
    """One engine per process / device."""

    def __init__(self, device=None, cta_group: int = 2, band: int = 8):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise GvError("Engine", _lib.GV_OK - 3, "no CUDA device: genevector_b200 has no CPU fallback")
        dev = torch.device(device if device is not None else "cuda")
        if dev.type != "cuda":
            raise GvError("Engine", _lib.GV_OK - 3, f"device {device!r} is not a CUDA device")
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        with torch.cuda.device(self.device):
            _lib.call("gv_device_check")
        self.cta_group = cta_group
        self.band = band
        self._sched_cache = {}
        # storage of the MI operand: "int8" (tcgen05 kind::i8, INT32 accumulators), "fp4" (kind::mxf4
        # with unit scales, FP32 accumulators: half the bytes, twice the cells per MMA, the same exact
        # counts below 2^24 cells) or "auto" (fp4 where it applies: 10 or 5 rows per gene)
        self.operand = "auto"
        # wave lockstep of the persistent tile kernels: "auto" = when the operand exceeds L2
        self.lockstep = "auto"
        # barrier words of the wave lockstep: one 64-byte slot per launch, taken in turn, so launches
        # that overlap on different streams never share one
        self._sync_ws = torch.zeros(32 * 64, dtype=torch.uint8, device=self.device)
        self._sync_next = 0
        self._flag_ws = torch.zeros(16, dtype=torch.int32, device=self.device)
        # bench.py sets this to a list to collect (entry point, start event, end event, kernels)
        self.kernel_events = None
        # multi-rank: send the value rows asynchronously and expand the one-hot operand meanwhile
        self.defer_values = os.environ.get("GV_DEFER_VALUES", "1") != "0"
        # pinned host matrix a callable `out_host` produced during the last mi_scores call
        self.last_host_result = None
        self.host_timings = None
        # worker threads of the pageable upload (0 = from the core count; one process per GPU shares
        # the host cores with its peers)
        lw = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
        ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)
        env_threads = int(os.environ.get("GV_COPY_THREADS", "0"))
        # every rank of the box uploads at once (sharded front end): a share of the cores each;
        # one process alone (single GPU, or rank 0 of the broadcast front end): up to 16 of them
        self.copy_threads = env_threads or max(2, min(8, ncpu // max(lw, 1)))
        self.copy_threads_alone = env_threads or max(2, min(16, ncpu - 2))
        # how pageable arrays are read: "staged" (worker threads copy into a pinned ring) or
        # "registered" (the caller's pages are page-locked in place: no CPU pass over the data, which is
        # what the ranks of a box want when they all upload at once)
        self.upload_mode = os.environ.get("GV_UPLOAD_MODE", "") or "staged"
        self._upload_tokens = []
        # metrics._run_to_host sets this: registered pages are unlocked after the whole call is enqueued
        self.defer_release = False
        # pack count data while staging (GV_PACK_UPLOADS=0 turns it off); bytes of the last upload
        self.pack_uploads = os.environ.get("GV_PACK_UPLOADS", "1") != "0"
        # chunks a rank's block of cells is uploaded in by the sharded front end (1 = one piece; more
        # overlap a chunk's exchange with the next chunk's upload -- measured: no gain at 2 or 8 GPUs,
        # the per-chunk hand-overs cost what the overlap saves)
        self.front_chunks = int(os.environ.get("GV_FRONT_CHUNKS", "1"))
        self.front_chunk_nnz = max(1, int(os.environ.get("GV_FRONT_CHUNK_NNZ", str(4 << 20))))   # entries per chunk, at least
        self.last_upload_bytes = 0

    # kernels launched by each entry point (memsets / copies not counted)
    _LAUNCHES = {"gv_discretize_csr": 2, "gv_expand_onehot": 1, "gv_moment_tiles": 1,
                 "gv_mi_tiles": 1, "gv_gram_dump": 1, "gv_graph_aggregate": 2, "gv_xcorr_tiles": 1,
                 "gv_symmetrize": 1}

    def _call(self, name, *args):
        """Call a kernel-launching entry point, optionally bracketed by CUDA events recorded on
        the launching stream."""
        if self.kernel_events is None:
            return _lib.call(name, *args)
        stream = torch.cuda.current_stream(self.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rc = _lib.call(name, *args)
        e1.record(stream)
        self.kernel_events.append((name, e0, e1, self._LAUNCHES.get(name, 1)))
        return rc

    # ------------------------------------------------------------------ helpers
    def _sync_ptr(self, operand_bytes: int):
        """Device scratch for the wave barrier, or None when the operand fits L2 anyway."""
        on = self.lockstep is True or (self.lockstep == "auto" and operand_bytes > 96 * 1024 * 1024)
        if not on:
            return None
        self._sync_next = (self._sync_next + 1) % 32
        return self._sync_ws.data_ptr() + 64 * self._sync_next

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def upload_csr(self, X, pinned: bool = False, rows: Optional[tuple] = None, pack: bool = False,
                   alone: bool = True) -> CsrDevice:
        """scipy CSR (host) -> device tensors.  Integer dtypes are converted to the smallest float
        type that holds them exactly (discretisation arithmetic is identical: see DESIGN.md).
        The arrays are pageable host memory (the reference hands over adata.X, data.py:171): they go
        through gv_upload_pageable (pinned staging ring filled by worker threads).  Whether the matrix
        is canonical (ascending, unique, in-range column indices per row) is taken from scipy when it
        already knows and otherwise checked on the device (gv_csr_check); only a matrix that is not
        is canonicalised on the host and uploaded again.
        rows = (r0, r1): upload that block of cells only (the sharded front end).
        pack: count data (float32 integers <= 255, at most 65,536 genes) are packed to uint8 values +
        uint16 column indices by the staging pass itself (3 bytes per entry cross the PCIe link instead
        of 8); data that do not fit are uploaded as they are.  Packed matrices feed the count path of the
        discretisation only (the MI / matrix-target calls); other consumers upload with pack=False.
        alone: no other process of the box is uploading at the same time (more staging threads)."""
        import time
        import scipy.sparse as sp
        t_up = time.perf_counter()
        if not sp.issparse(X) or X.format != "csr":
            X = sp.csr_matrix(X)
        known = getattr(X, "_has_canonical_format", None)
        if getattr(X, "_has_sorted_indices", True) is False:
            known = False
        if known is False:
            X = X.copy()
            X.sum_duplicates()
            known = True
        n_cells, n_genes = X.shape
        r0, r1 = rows if rows is not None else (0, n_cells)
        e0, e1 = int(X.indptr[r0]), int(X.indptr[r1])
        data = X.data[e0:e1]
        if data.dtype == np.float32 or data.dtype == np.float64:
            pass
        elif np.issubdtype(data.dtype, np.integer) or data.dtype == np.bool_:
            big = data.size and np.abs(data).max() >= (1 << 24)
            data = data.astype(np.float64 if big else np.float32)
        else:
            data = data.astype(np.float64)
        cols = np.ascontiguousarray(X.indices[e0:e1], dtype=np.int32)
        ptr = np.ascontiguousarray(X.indptr[r0:r1 + 1], dtype=np.int64)
        if e0:
            ptr = ptr - e0
        data = np.ascontiguousarray(data)
        threads = self.copy_threads_alone if alone else self.copy_threads
        with torch.cuda.device(self.device):
            st = self._stream()
            dp = torch.empty(ptr.shape, dtype=torch.int64, device=self.device)
            packed = False
            if pack and data.dtype == np.float32 and n_genes <= 65536 and self.pack_uploads:
                t_st = time.perf_counter()
                dv = torch.empty(data.shape, dtype=torch.uint8, device=self.device)
                dc = torch.empty(cols.shape, dtype=torch.uint16, device=self.device)
                fits = np.zeros(1, dtype=np.int32)
                _lib.call("gv_upload_packed_counts", dv.data_ptr(), dc.data_ptr(), data.ctypes.data, cols.ctypes.data,
                          data.size, threads, fits.ctypes.data, st)
                packed = bool(fits[0])
                if packed:
                    _lib.call("gv_upload_pageable", dp.data_ptr(), ptr.ctypes.data, ptr.nbytes, threads, st)
                    self.last_upload_bytes = int(data.size * 3 + ptr.nbytes)
                    self._timed_host("upload_csr: host side of the copies (packed counts)", t_st)
            if not packed:
                dv = torch.empty(data.shape, dtype=torch.float32 if data.dtype == np.float32 else torch.float64,
                                 device=self.device)
                dc = torch.empty(cols.shape, dtype=torch.int32, device=self.device)
                self.last_upload_bytes = int(data.nbytes + cols.nbytes + ptr.nbytes)
            self._timed_host("upload_csr: views + device buffers", t_up)
            t_st = time.perf_counter()
            for dst, src in (() if packed else ((dv, data), (dc, cols), (dp, ptr))):
                if self.upload_mode == "registered" and src.nbytes >= (32 << 20):
                    import ctypes as C
                    tok = C.c_void_p()
                    _lib.call("gv_upload_registered", dst.data_ptr(), src.ctypes.data, src.nbytes, st, C.byref(tok))
                    if tok.value:
                        self._upload_tokens.append((tok.value, src))      # src stays alive until released
                else:
                    _lib.call("gv_upload_pageable", dst.data_ptr(), src.ctypes.data, src.nbytes,
                              threads, st)
            if not packed:
                self._timed_host("upload_csr: host side of the copies", t_st)
            csr = CsrDevice(dv, dc, dp, r1 - r0, n_genes, packed)
            if known is not True:
                ok = np.zeros(1, dtype=np.int32)
                _lib.call("gv_csr_check", dc.data_ptr(), csr.dtype_code, dp.data_ptr(), csr.n_cells, n_genes,
                          self._flag_ws.data_ptr(), ok.ctypes.data, st)
                if int(ok[0]) == 2:
                    raise ValueError("column index out of range in the CSR matrix")
                if int(ok[0]) != 0:
                    Xc = X.copy()
                    Xc.sum_duplicates()             # sorts the indices and merges duplicate entries
                    Xc._has_canonical_format = True
                    return self.upload_csr(Xc, rows=rows, pack=pack, alone=alone)
        if not self.defer_release:
            self.release_uploads()
        self._timed_host("upload_csr (host side: staging + canonical check)", t_up)
        return csr

    def release_uploads(self):
        """Unlock the pages of registered uploads (waits for their DMAs).  Called once the rest of a
        call's work is enqueued, so the host does it while the device computes."""
        toks, self._upload_tokens = self._upload_tokens, []
        for tok, _src in toks:
            _lib.call("gv_upload_release", tok)

    def schedule(self, op_rows: int, cta_group: Optional[int] = None, rank: int = 0, world: int = 1):
        """Device tile list (int32 [n,2]) for this rank: a contiguous slice of the band-ordered
        upper-triangular schedule."""
        cg = cta_group or self.cta_group
        key = (op_rows, cg, self.band, rank, world)
        if key not in self._sched_cache:
            part, n = tile_slice(op_rows, cg, self.band, rank, world)
            self._sched_cache[key] = (torch.from_numpy(part).to(self.device), len(part), n)
        return self._sched_cache[key]

    # ------------------------------------------------------------------ discretisation
    def discretize(self, csr: CsrDevice, n_bins: int = 10, val_mode: int = -1,
                   buf: Optional[torch.Tensor] = None, path: int = _lib.GV_PATH_AUTO,
                   n_limbs: Optional[int] = None) -> BinBlock:
        """gv_discretize_csr.  val_mode: -1 no value rows; 0 the values themselves (sign / pearson /
        cosine): count digits, one per gene, retried with more digits when the counts need them and
        with fixed-point rows (mode 2) when the data are not integers; 1 detection indicator
        (jaccard); 2 fixed point; 3 doubled average ranks (spearman).  n_limbs fixes the digit count
        (no retry).  The block's val_mode / n_limbs say what was produced."""
        if not 1 <= n_bins <= _lib.GV_MAX_BINS:
            raise GvError("discretize", -1, f"n_bins must be in [1, {_lib.GV_MAX_BINS}]")
        if csr.values.dtype not in (torch.float32, torch.float64) and not csr.packed:
            raise GvError("discretize", -1, "values must be float32 or float64")
        with torch.cuda.device(self.device):
            lb = limb_bits_for(max(csr.n_cells, 1)) if val_mode in (0, 2, 3) else 7
            l_max = max_limbs_for(csr.n_cells, lb)
            l_fixed = min(3, l_max)
            auto = n_limbs is None and val_mode == 0
            if val_mode in (-1, 1):
                L = 1
            elif n_limbs is not None:
                L = n_limbs
            elif val_mode == 2:
                L = l_fixed
            elif val_mode == 3:
                L = -(-(2 * max(csr.n_cells, 1)).bit_length() // lb)
            else:
                L = 1
            if val_mode >= 0 and not 1 <= L <= max(l_max, 1):
                raise GvError("discretize", -4, f"value rows would need {L} digits per gene; at most "
                              f"{l_max} keep the co-moments of {csr.n_cells} cells exact")
            dt = csr.dtype_code
            sorted_ws = 1 if val_mode == 3 else 0
            ws_bytes = self.lib.gv_discretize_workspace_bytes(csr.nnz, csr.n_cells, csr.n_genes,
                                                              _lib.GV_F32 if csr.packed else dt, sorted_ws)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
            q = quantile_table()
            mode = val_mode
            while True:
                kp, pitch, lay, total = block_layout(csr.n_cells, csr.n_genes, mode, L, n_bins)
                b = buf if (buf is not None and buf.numel() >= total) else \
                    torch.empty(total, dtype=torch.uint8, device=self.device)
                assert b.device == self.device
                blk = _views(b, csr.n_cells, csr.n_genes, n_bins, mode, L, lb)
                flags = np.zeros(2, dtype=np.uint64)
                used = np.zeros(1, dtype=np.int32)
                vr = blk.val_rows
                try:
                    self._call("gv_discretize_csr", csr.values.data_ptr(), dt, csr.col_idx.data_ptr(),
                               csr.row_ptr.data_ptr(), csr.n_cells, csr.n_genes, csr.nnz, n_bins,
                               q.ctypes.data, blk.bins.data_ptr(), pitch, blk.n_bins_per_gene.data_ptr(),
                               blk.marg.data_ptr(), blk.edges.data_ptr(), blk.gsum.data_ptr(),
                               blk.gsq.data_ptr(), vr.data_ptr() if vr is not None else None, kp,
                               vr.shape[0] if vr is not None else 0, max(mode, 0), L, lb,
                               flags.ctypes.data, ws.data_ptr(), ws_bytes, path, used.ctypes.data,
                               self._stream())
                except GvError as e:
                    negative, fractional = int(flags[0]) & 1, int(flags[0]) & 2
                    if not (auto and mode == 0 and e.status == -4):
                        raise
                    need = -(-int(flags[1]).bit_length() // lb)      # digits the largest count needs
                    if fractional or negative or need > l_max:
                        if l_fixed < 1:
                            raise
                        mode, L = 2, l_fixed       # not counts (or too large): fixed-point rows
                    elif need > L:
                        L = need                   # larger counts: more INT8 digits per gene, still exact
                    else:
                        raise
                    continue
                break
            blk.data_flags = (int(flags[0]), int(flags[1]))
            blk.path_used = int(used[0])
            hdr = torch.tensor([L, lb, int(flags[0]), int(flags[1]) & 0x7FFFFFFFFFFFFFFF, mode],
                               dtype=torch.int64)
            o, _ = lay["header"]
            b[o:o + 40].view(torch.int64).copy_(hdr)
            return blk

    def _fmt(self, operand: Optional[str], blk: Optional[BinBlock] = None) -> int:
        op = operand or self.operand
        if op == "auto":
            ok = blk is not None and _lib.call("gv_rows_per_gene", blk.n_bins) in (5, 10) \
                and blk.n_cells < (1 << 24)
            op = "fp4" if ok else "int8"
        if op == "int8":
            return _lib.GV_OPERAND_INT8
        if op == "fp4":
            return _lib.GV_OPERAND_FP4
        raise GvError("operand", -1, f"unknown operand format {op!r} (int8 or fp4)")

    def expand_onehot(self, blk: BinBlock, operand: Optional[str] = None) -> tuple:
        """gv_expand_onehot: packed bins -> one-hot operand [rows][row_bytes] (int8: one byte per
        cell; fp4: two cells per byte).  Returns (operand tensor, rows per gene)."""
        with torch.cuda.device(self.device):
            fmt = self._fmt(operand, blk)
            if fmt == _lib.GV_OPERAND_FP4 and blk.n_cells >= (1 << 24):
                raise GvError("expand_onehot", -4, "fp4 operand: counts are exact only below 2^24 cells")
            B = _lib.call("gv_rows_per_gene", blk.n_bins)
            rows = _lib.check("gv_onehot_rows", self.lib.gv_onehot_rows(blk.n_genes, B))
            row_bytes = blk.kp if fmt == _lib.GV_OPERAND_INT8 else blk.kp // 2
            onehot = torch.empty((rows, row_bytes), dtype=torch.int8 if fmt == 0 else torch.uint8,
                                 device=self.device)
            self._call("gv_expand_onehot", blk.bins.data_ptr(), blk.bins_pitch, blk.n_genes, B,
                       onehot.data_ptr(), blk.kp, fmt, self._stream())
            return onehot, B

    # ------------------------------------------------------------------ tile kernels
    def gram_dump(self, operand: torch.Tensor, cta_group: Optional[int] = None,
                  active_rows: int = 32, fmt: int = 0) -> torch.Tensor:
        """Test-only: raw INT32 accumulators of every scheduled tile."""
        with torch.cuda.device(self.device):
            cg = cta_group or self.cta_group
            rows, kp = operand.shape
            tiles, n, _ = self.schedule(rows, cg)
            out = torch.zeros((rows, rows), dtype=torch.int32, device=self.device)
            self._call("gv_gram_dump", operand.data_ptr(), rows, kp, tiles.data_ptr(), n,
                      out.data_ptr(), rows, cg, active_rows, fmt, self._stream())
            return out

    def sign_tiles_for(self, mi_tiles: np.ndarray, B: int, n_limbs: int = 1) -> torch.Tensor:
        """Sign-map tiles (operand rows of the value matrix: 32 // n_limbs genes per 32 rows) that cover the
        gene pairs of the given MI tiles (one-hot operand rows), so a rank only computes the part
        of the sign map it reads."""
        gpb = 32 // B
        tm = _lib.call("gv_tile_m", self.cta_group)
        vgpb = 32 // n_limbs                            # genes per 32-row block of the value rows
        gr, gc = tm // 32 * vgpb, 8 * vgpb              # genes per sign tile side
        r_lo = mi_tiles[:, 0].astype(np.int64) // 32 * gpb
        r_hi = (mi_tiles[:, 0].astype(np.int64) + tm) // 32 * gpb - 1
        c_lo = mi_tiles[:, 1].astype(np.int64) // 32 * gpb
        c_hi = (mi_tiles[:, 1].astype(np.int64) + 256) // 32 * gpb - 1
        keys = set()
        for rr in (r_lo, r_hi):
            for cc in (c_lo, c_hi):
                # an MI tile spans fewer genes per side than a sign tile, so its corners hit every
                # sign tile it touches
                keys.update(np.unique((rr // gr) * (1 << 32) + (cc // gc)).tolist())
        arr = np.array(sorted(keys), dtype=np.int64)
        st = np.stack([(arr >> 32) * tm, (arr & 0xFFFFFFFF) * 256], axis=1)
        st = st[st[:, 0] <= st[:, 1] + 255]          # tiles holding some (row <= col)
        return torch.from_numpy(np.ascontiguousarray(st.astype(np.int32))).to(self.device)

    def sign_matrix(self, blk: BinBlock, tiles: Optional[torch.Tensor] = None,
                    rule: str = "numpy") -> torch.Tensor:
        """int8 [G][ld] Pearson sign from the INT8 value rows (exact integer co-moments).
        `tiles`: optional subset of sign tiles (see sign_tiles_for); default the whole triangle.
        rule "numpy": np.sign of the correlation, 0 where it is 0 or undefined (metrics.py:111,134, the
        NumPy / Numba / torch tiers); rule "rust": +1 unless the correlation is negative (lib.rs:91-94)."""
        if rule not in ("numpy", "rust"):
            raise ValueError(f"unknown sign rule {rule!r} (numpy or rust)")
        assert blk.val_rows is not None and blk.val_mode in (0, 2)
        with torch.cuda.device(self.device):
            G = blk.n_genes
            ld = _align(G, 16)
            sgn = torch.zeros((G, ld), dtype=torch.int8, device=self.device)
            rows = value_rows(G, blk.n_limbs)
            if tiles is None:
                tiles, n, _ = self.schedule(rows)
            else:
                n = int(tiles.shape[0])
            self._call("gv_moment_tiles", _lib.GV_MOM_SIGN, blk.val_rows.data_ptr(), rows, blk.kp,
                       blk.gsum.data_ptr(), blk.gsq.data_ptr(), blk.marg.data_ptr(), blk.edges.data_ptr(),
                       blk.n_cells, G,
                       blk.n_limbs, blk.limb_bits, tiles.data_ptr(), n, sgn.data_ptr(), ld, None, 0,
                       self.cta_group, self._sync_ptr(rows * blk.kp), 1 if rule == "rust" else 0,
                       self._stream())
            return sgn

    def _tiles_to_host(self, launch, n_tiles, pieces, out, sink):
        """Run a rank's tile slice with launch(t0, t1) in `pieces` (host_pieces) and copy every
        finished piece's rectangles of `out` (device float32 [G][G]) into the host matrix of `sink`
        (HostSink) on a second stream, under the next piece's computation."""
        if sink is None or n_tiles == 0 or not pieces:
            if n_tiles:
                launch(0, n_tiles)
            if sink is not None:
                sink.resolve(self)
            return
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        # all pieces are enqueued first (an event after each), then the copies: a host buffer that
        # is only allocated now (lazy sink) is paid for while the GPU is already computing
        evs = []
        for t0, t1, rects in pieces:
            launch(t0, t1)
            ev = torch.cuda.Event()
            ev.record(main)
            evs.append(ev)
        ptr, pitch = sink.resolve(self)
        cs = self._copy_stream
        esz, ld = out.element_size(), out.stride(0) * out.element_size()
        for ev, (t0, t1, rects) in zip(evs, pieces):
            cs.wait_event(ev)
            for r0, r1, c0, c1 in rects:
                _lib.call("gv_copy_rect_d2h", ptr + r0 * pitch + c0 * esz, pitch,
                          out.data_ptr() + r0 * ld + c0 * esz, ld, (c1 - c0) * esz, r1 - r0, cs.cuda_stream)
        done = torch.cuda.Event()
        done.record(cs)
        main.wait_event(done)

    def moment_scores(self, blk: BinBlock, kind: int, out: Optional[torch.Tensor] = None,
                      rank: int = 0, world: int = 1, sink=None, host_chunks: int = 6) -> torch.Tensor:
        """float32 [G][G] Pearson / cosine / Jaccard matrix (zero diagonal).  sink: see mi_scores."""
        assert blk.val_rows is not None
        with torch.cuda.device(self.device):
            G = blk.n_genes
            if out is None:
                out = torch.zeros((G, G), dtype=torch.float32, device=self.device)
            rows = value_rows(G, blk.n_limbs)
            tiles, n, _ = self.schedule(rows, rank=rank, world=world)

            def launch(t0, t1):
                self._call("gv_moment_tiles", kind, blk.val_rows.data_ptr(), rows, blk.kp,
                           blk.gsum.data_ptr(), blk.gsq.data_ptr(), blk.marg.data_ptr(), blk.edges.data_ptr(),
                           blk.n_cells, G, blk.n_limbs, blk.limb_bits, tiles.data_ptr() + 8 * t0, t1 - t0,
                           None, 0, out.data_ptr(), out.stride(0), self.cta_group,
                           self._sync_ptr(rows * blk.kp), 0, self._stream())

            pieces = None
            if sink is not None:
                pieces = host_pieces(rows, self.cta_group, self.band, 32 // blk.n_limbs, G, rank, world,
                                     host_chunks)
            self._tiles_to_host(launch, n, pieces, out, sink)
            return out

    def mi_scores(self, blk: BinBlock, onehot: torch.Tensor, B: int,
                  sgn: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                  rank: int = 0, world: int = 1, out_host=None, host_chunks: int = 8,
                  sink=None) -> torch.Tensor:
        """float32 [G][G] MI (times the Pearson sign when `sgn` is given).

        sink (HostSink: a host matrix of pitch bytes per row, page-locked): the result is streamed to
        the host while it is computed.  The rank's slice of the tile schedule is launched in
        `host_chunks` pieces and the rectangles of the result matrix a finished piece owns
        (host_pieces) are copied on a second stream under the next piece.  With one rank these are
        blocks of full rows; with several every rank writes the rectangles of its own tiles into the
        one host matrix all ranks of the box share (hostpool.py), so the full matrix is assembled
        without a collective.  out_host (pinned float32 [G][G] tensor, or a callable returning one)
        is the single-rank shorthand."""
        with torch.cuda.device(self.device):
            G = blk.n_genes
            if out is None:
                out = torch.zeros((G, G), dtype=torch.float32, device=self.device)
            rows, row_bytes = onehot.shape
            fmt = _lib.GV_OPERAND_FP4 if onehot.dtype == torch.uint8 else _lib.GV_OPERAND_INT8
            tiles, n, _ = self.schedule(rows, rank=rank, world=world)

            def launch(t0, t1):
                self._call("gv_mi_tiles", onehot.data_ptr(), rows, row_bytes, B, blk.marg.data_ptr(),
                           sgn.data_ptr() if sgn is not None else None,
                           sgn.stride(0) if sgn is not None else 0, G, tiles.data_ptr() + 8 * t0, t1 - t0,
                           out.data_ptr(), out.stride(0), self.cta_group, fmt,
                           self._sync_ptr(rows * row_bytes), self._stream())

            if sink is None and out_host is not None:
                if world > 1:
                    raise GvError("mi_scores", -1, "out_host is single-rank; pass a HostSink over the shared matrix")
                sink = HostSink(tensor=out_host)
            pieces = None
            if sink is not None:
                pieces = host_pieces(rows, self.cta_group, self.band, 32 // B, G, rank, world, host_chunks)
            self._tiles_to_host(launch, n, pieces, out, sink)
            return out

    # ------------------------------------------------------------------ graph cross-correlation
    def upload_graph(self, graph, n_cells: int):
        """scipy sparse adjacency [n_cells, n_cells] -> device CSR (int64 row_ptr, int32 col, float64 val)."""
        import scipy.sparse as sp
        W = sp.csr_matrix(graph)
        if W.shape != (n_cells, n_cells):
            raise ValueError(f"graph must be {n_cells} x {n_cells}, got {W.shape}")
        if W.nnz and float(W.data.min()) < 0:
            raise GvError("upload_graph", -4, "graph weights must be non-negative")
        ptr = torch.from_numpy(np.ascontiguousarray(W.indptr, dtype=np.int64)).to(self.device)
        col = torch.from_numpy(np.ascontiguousarray(W.indices, dtype=np.int32)).to(self.device)
        val = torch.from_numpy(np.ascontiguousarray(W.data, dtype=np.float64)).to(self.device)
        return ptr, col, val

    def rect_schedule(self, rows_a: int, rows_b: int):
        key = ("rect", rows_a, rows_b, self.cta_group)
        if key not in self._sched_cache:
            tm = _lib.call("gv_tile_m", self.cta_group)
            n = _lib.check("gv_tile_schedule_rect", self.lib.gv_tile_schedule_rect(rows_a, rows_b, tm, None, 0))
            host = np.zeros((max(n, 1), 2), dtype=np.int32)
            _lib.check("gv_tile_schedule_rect",
                       self.lib.gv_tile_schedule_rect(rows_a, rows_b, tm, host.ctypes.data, n))
            self._sched_cache[key] = (torch.from_numpy(host).to(self.device), n)
        return self._sched_cache[key]

    def graph_xcorr(self, csr: CsrDevice, graph_dev, include_self: bool = False) -> torch.Tensor:
        """float32 [G][G] symmetrised cross-correlation between expression and graph-neighbour mean
        expression (_graph_targets.py:10-55, aggregation "mean")."""
        w_ptr, w_col, w_val = graph_dev
        with torch.cuda.device(self.device):
            blk = self.discretize(csr, 10, val_mode=0)        # count digits, or fixed point for floats
            G, N, kp, lb = blk.n_genes, blk.n_cells, blk.kp, blk.limb_bits
            ly = min(3, max_limbs_for(N, lb))
            if ly != 3:
                raise GvError("graph_xcorr", -4, f"{N} cells leave fewer than three exact digits per value")
            rows_x, rows_y = value_rows(G, blk.n_limbs), value_rows(G, ly)
            y_rows = torch.empty((rows_y, kp), dtype=torch.int8, device=self.device)
            ysum = torch.empty(G, dtype=torch.int64, device=self.device)
            ysq = torch.empty_like(ysum)
            yscale = torch.empty(G, dtype=torch.float64, device=self.device)
            w_inv = torch.empty(max(N, 1), dtype=torch.float64, device=self.device)
            xscale_ptr = blk.edges.data_ptr() + 8 * _lib.GV_EDGE_SCALE_SLOT
            xoff_ptr = blk.edges.data_ptr() + 8 * _lib.GV_EDGE_OFFSET_SLOT
            self._call("gv_graph_aggregate", blk.val_rows.data_ptr(), rows_x, kp, blk.n_limbs, lb, xscale_ptr,
                       xoff_ptr, _lib.GV_EDGE_STRIDE, w_ptr.data_ptr(), w_col.data_ptr(), w_val.data_ptr(),
                       w_inv.data_ptr(), 1 if include_self else 0, N, G, y_rows.data_ptr(), rows_y, ly,
                       ysum.data_ptr(), ysq.data_ptr(), yscale.data_ptr(), self._stream())
            tiles, n = self.rect_schedule(rows_x, rows_y)
            xc = torch.empty((G, G), dtype=torch.float32, device=self.device)
            self._call("gv_xcorr_tiles", blk.val_rows.data_ptr(), rows_x, blk.n_limbs, y_rows.data_ptr(), rows_y,
                       ly, kp, lb, blk.gsum.data_ptr(), blk.gsq.data_ptr(), xscale_ptr, _lib.GV_EDGE_STRIDE,
                       ysum.data_ptr(), ysq.data_ptr(), yscale.data_ptr(), N, G, tiles.data_ptr(), n,
                       xc.data_ptr(), xc.stride(0), self.cta_group,
                       self._sync_ptr(max(rows_x, rows_y) * kp), self._stream())
            out = torch.empty_like(xc)
            self._call("gv_symmetrize", xc.data_ptr(), xc.stride(0), G, out.data_ptr(), out.stride(0),
                       self._stream())
            return out

    def mma_peak(self, operand: str = "fp4", iters: int = 60000, mode: int = 0) -> float:
        """Diagnostic: TOP/s the tensor pipe sustains on the MI contraction's own MMA instruction
        issued back to back (gv_mma_peak); bench.py's live roofline denominator."""
        with torch.cuda.device(self.device):
            out = np.zeros(1, dtype=np.float64)
            _lib.call("gv_mma_peak", self._fmt(operand), self.cta_group, iters, mode, out.ctypes.data,
                      self._stream())
            return float(out[0])

    # ------------------------------------------------------------------ whole paths
    def _replicate(self, blk_or_none, n_cells, n_genes, n_bins, val_mode, group,
                   defer_values: bool = False) -> BinBlock:
        """The single collective of the path: broadcast the bin block from rank 0.  The header at
        the front of the block says how many digits per gene the value rows have; only when there is
        more than one (counts > 2^limb_bits - 1, or data that are not counts) is the extra tail sent
        by a further call.  defer_values: the value rows (the last section of the block) are sent by
        an asynchronous call whose handle is left in blk.pending, so that the caller can expand the
        one-hot operand (which needs the bins only) while they are still on the wire."""
        import torch.distributed as dist
        _, _, lay, total1 = block_layout(n_cells, n_genes, val_mode, 1, n_bins)
        if blk_or_none is not None:
            buf = blk_or_none.buf
        else:
            buf = torch.empty(total1, dtype=torch.uint8, device=self.device)
        ev = None
        if self.kernel_events is not None:
            stream = torch.cuda.current_stream(self.device)
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(stream)
        src = dist.get_global_rank(group, 0) if group is not None else 0
        val_off, val_bytes = lay["val"]
        pending = None
        if defer_values and val_bytes > 0:
            broadcast_block(buf, val_off, group)                       # header, bins, per-gene arrays
        else:
            broadcast_block(buf, total1, group)
        o, _ = lay["header"]
        hdr = buf[o:o + 40].view(torch.int64).cpu()
        L, lb, flags, mode = int(hdr[0]), int(hdr[1]), (int(hdr[2]), int(hdr[3])), int(hdr[4])
        if val_mode < 0:
            mode = val_mode
        if defer_values and val_bytes > 0:
            if L == 1:
                pending = dist.broadcast(buf[val_off:total1], src=src, group=group, async_op=True)
            else:
                dist.broadcast(buf[val_off:total1], src=src, group=group)
        if L > 1:
            _, _, _, total_l = block_layout(n_cells, n_genes, val_mode, L, n_bins)
            if blk_or_none is None:
                big = torch.empty(total_l, dtype=torch.uint8, device=self.device)
                big[:total1].copy_(buf)
                buf = big
            dist.broadcast(buf[total1:total_l], src=src, group=group)
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(self.device))
            self.kernel_events.append(("broadcast", ev[0], ev[1], 0))
        blk = _views(buf, n_cells, n_genes, n_bins, mode, L, lb, flags)
        blk.pending = pending
        return blk

    def mi_from_csr(self, csr: Optional[CsrDevice], n_cells: int, n_genes: int, n_bins: int = 10,
                    signed: bool = False, rank: int = 0, world: int = 1, group=None,
                    out: Optional[torch.Tensor] = None,
                    out_host: Optional[torch.Tensor] = None, sign_rule: str = "numpy",
                    sink=None) -> torch.Tensor:
        """All-pairs (signed) MI.  With world > 1 only rank 0 needs `csr`; each rank returns a
        [G][G] matrix in which the entries of its own tiles are filled."""
        import torch.distributed as dist
        val_mode = 0 if signed else -1
        distributed = world > 1 and dist.is_available() and dist.is_initialized()
        # without an initialised process group the ranks are emulated one after another in this
        # process (tests): each "rank" then discretises for itself
        blk = self.discretize(csr, n_bins, val_mode) if (rank == 0 or not distributed) else None
        if distributed:
            blk = self._replicate(blk, n_cells, n_genes, n_bins, val_mode, group,
                                  defer_values=signed and self.defer_values)
        sgn = None
        B = _lib.call("gv_rows_per_gene", n_bins)
        onehot, B = self.expand_onehot(blk)        # needs the bins only: runs under the value-row transfer
        if getattr(blk, "pending", None) is not None:
            blk.pending.wait()                     # the current stream now waits for the value rows
            blk.pending = None
        if signed:
            # each rank computes only the part of the sign map its own MI tiles read (1 % of the
            # work, sharded like the MI tiles); no second collective
            st = None
            if world > 1:
                rows = _lib.check("gv_onehot_rows", self.lib.gv_onehot_rows(n_genes, B))
                key = ("sign", rows, self.cta_group, self.band, rank, world, blk.n_limbs)
                if key not in self._sched_cache:
                    mine, _ = tile_slice(rows, self.cta_group, self.band, rank, world)
                    self._sched_cache[key] = self.sign_tiles_for(mine, B, blk.n_limbs)
                st = self._sched_cache[key]
            sgn = self.sign_matrix(blk, tiles=st, rule=sign_rule)
        return self.mi_scores(blk, onehot, B, sgn=sgn, out=out, rank=rank, world=world, out_host=out_host,
                              sink=sink)

    # ------------------------------------------------------------------ sharded front end
    # scratch words of hostpool (per rank): 0 handle epoch, 1..8 the 64-byte IPC handle, 9 idle epoch,
    # 10 transposition-done epoch, 11 / 12 data flags, 13 peers-opened epoch * 2 + ok
    def resolve_front(self, front: str, world: int) -> str:
        """"broadcast" or "sharded" for this process group (see metrics.compute_mi_b200)."""
        if world <= 1 or front == "broadcast":
            return "broadcast"
        if front not in ("auto", "sharded"):
            raise ValueError(f"unknown front {front!r} (auto, broadcast or sharded)")
        if getattr(self, "_shard_disabled", False) or world > 8:
            return "broadcast"
        env = os.environ.get("GV_FRONT", "")
        if front == "auto" and env in ("broadcast", "sharded"):
            return env
        return "sharded"

    def _peer_block(self, n_cells, n_genes, rank, world, pool):
        """Allocate (once per shape) this rank's peer-visible block and map the other ranks'."""
        import ctypes as C
        key = (n_cells, n_genes, world, self.device.index)
        if key in pool.peer_blocks:
            return pool.peer_blocks[key]
        if len(pool.peer_blocks) >= 4:
            return None
        kp, pitch, lay, total1 = block_layout(n_cells, n_genes, 0, 1)
        off = dict(lay)
        cur = total1
        for name, nbytes in (("ghist", n_genes * 256 * 4), ("ghist_total", n_genes * 256 * 4),
                             ("flags", 256), ("qtab", 8192), ("luts", n_genes * 256)):
            off[name] = (cur, nbytes)
            cur += _align(nbytes)
        pool.peer_epoch += 1
        epoch = pool.peer_epoch
        own, ok = C.c_void_p(), 1
        handle = (C.c_uint8 * 64)()
        try:
            with torch.cuda.device(self.device):
                _lib.call("gv_ipc_alloc", cur, C.byref(own), handle)
        except GvError:
            ok = 0
        words = np.frombuffer(bytes(handle), dtype=np.uint64)
        for k in range(8):
            pool.scratch_store(1 + k, int(words[k]))
        pool.scratch_store(0, epoch * 2 + ok)
        ptrs = [None] * world
        for r in range(world):
            _lib.call("gv_flag_wait_ge", pool.scratch_addr(r, 0), epoch * 2, pool.timeout_ms)
            ok &= pool.scratch_load(r, 0) & 1
        if ok:
            for r in range(world):
                if r == rank:
                    ptrs[r] = own.value
                    continue
                h = np.array([pool.scratch_load(r, 1 + k) for k in range(8)], dtype=np.uint64).tobytes()
                p = C.c_void_p()
                try:
                    with torch.cuda.device(self.device):
                        _lib.call("gv_ipc_open", h, C.byref(p))
                    ptrs[r] = p.value
                except GvError:
                    ok = 0
                    break
        pool.scratch_store(13, epoch * 2 + ok)
        for r in range(world):
            _lib.call("gv_flag_wait_ge", pool.scratch_addr(r, 13), epoch * 2, pool.timeout_ms)
            ok &= pool.scratch_load(r, 13) & 1
        if not ok:
            # some rank cannot allocate or map peer memory: every rank drops the sharded front end
            self._shard_disabled = True
            pool.peer_blocks[key] = None
            return None
        tensor = torch.as_tensor(_DevPtr(own.value, cur), device=self.device)
        pb = PeerBlock(own.value, cur, ptrs, off, tensor)
        pool.peer_blocks[key] = pb
        return pb

    def _shard_front(self, chunks, n_cells, n_genes, n_bins, signed, rank, world, pool):
        """Sharded discretisation: returns the BinBlock (views into this rank's peer block), or None
        when the data do not qualify (every rank decides the same from the combined data flags).
        chunks: iterable of (CsrDevice of a run of this rank's cells, (first block, end block)), consumed
        lazily -- a chunk's transposition and peer copies are enqueued before the next chunk is produced,
        so they run under its upload."""
        import ctypes as C
        if not 1 <= n_bins <= 15:
            return None                      # one byte per cell (n_bins 16..31): the broadcast front end
        pb = self._peer_block(n_cells, n_genes, rank, world, pool)
        if pb is None:
            return None
        pool.front_seq += 1
        step = pool.front_seq
        kp, pitch, lay, _ = block_layout(n_cells, n_genes, 0, 1)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            # nobody may still be reading what this step overwrites (peers' dense rows, my histogram)
            stream.synchronize()
            pool.scratch_store(9, step)
            pool.scratch_wait_all(9, step)
            ev = None
            if self.kernel_events is not None:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record(stream)
            dests = (C.c_void_p * world)(*[p + pb.off["val"][0] for p in pb.ptrs])
            flags = np.zeros(2, dtype=np.uint64)
            dt, first, prev = _lib.GV_F32, True, None

            def transpose(item, phase):
                c, (c0, c1) = item
                _lib.call("gv_shard_transpose", c.values.data_ptr(), c.dtype_code, c.col_idx.data_ptr(),
                          c.row_ptr.data_ptr(), c.n_cells, n_genes, c0, c1, dests, world, rank, kp,
                          pb.own_ptr + pb.off["ghist"][0], pb.own_ptr + pb.off["flags"][0], flags.ctypes.data,
                          phase, stream.cuda_stream)
                return _lib.GV_F64 if c.dtype_code == _lib.GV_F64 else _lib.GV_F32

            for item in chunks:                  # producing the next chunk (its upload) overlaps the previous one
                if prev is not None:
                    dt = transpose(prev, 1 if first else 0)
                    first = False
                prev = item
            dt = transpose(prev, (1 if first else 0) | 2)
            pool.scratch_store(11, int(flags[0]))
            pool.scratch_store(12, int(flags[1]))
            pool.scratch_store(10, step)
            pool.scratch_wait_all(10, step)
            f0 = f1 = 0
            for r in range(world):
                f0 |= pool.scratch_load(r, 11)
                f1 = max(f1, pool.scratch_load(r, 12))
            if ev is not None:
                ev[1].record(stream)
                self.kernel_events.append(("gv_shard_transpose (+ peer stores)", ev[0], ev[1], 1))
            # signed MI: the dense count rows are the one-digit INT8 operand of the sign pass
            lb = limb_bits_for(max(n_cells, 1)) if signed else 7
            limit = ((1 << lb) - 1) if signed else 255
            if (f0 & 14) or (signed and (f0 & 1)) or f1 > limit:
                return None
            hists = (C.c_void_p * world)(*[p + pb.off["ghist"][0] for p in pb.ptrs])
            val_mode = 0 if signed else -1
            blk = _views(pb.tensor, n_cells, n_genes, n_bins, val_mode, 1, lb, (f0, f1))
            q = quantile_table()
            vrows = value_rows(n_genes, 1)
            self._call("gv_shard_bins", dt, n_cells, n_genes, n_bins, q.ctypes.data,
                       pb.own_ptr + pb.off["val"][0], hists, world, pb.own_ptr + pb.off["ghist_total"][0],
                       pb.own_ptr + pb.off["qtab"][0], pb.own_ptr + pb.off["luts"][0], blk.bins.data_ptr(), pitch,
                       blk.n_bins_per_gene.data_ptr(),
                       blk.marg.data_ptr(), blk.edges.data_ptr(), blk.gsum.data_ptr(), blk.gsq.data_ptr(),
                       1 if signed else 0, vrows, lb, stream.cuda_stream)
            blk.path_used = _lib.GV_PATH_COUNTS
            return blk

    @staticmethod
    def shard_cells(n_cells, rank, world):
        """(first block, end block, first cell, end cell) of a rank: contiguous runs of 128-cell blocks of
        the padded cell axis."""
        kp = _align(max(n_cells, 1), 256)
        blocks = kp // 128
        b0, b1 = blocks * rank // world, blocks * (rank + 1) // world
        return b0, b1, min(n_cells, b0 * 128), min(n_cells, b1 * 128)

    def _mi_after_front(self, blk, signed, rank, world, out, sink, sign_rule):
        B = _lib.call("gv_rows_per_gene", blk.n_bins)
        onehot, B = self.expand_onehot(blk)
        sgn = None
        if signed:
            st = None
            if world > 1:
                rows = _lib.check("gv_onehot_rows", self.lib.gv_onehot_rows(blk.n_genes, B))
                key = ("sign", rows, self.cta_group, self.band, rank, world, blk.n_limbs)
                if key not in self._sched_cache:
                    mine, _ = tile_slice(rows, self.cta_group, self.band, rank, world)
                    self._sched_cache[key] = self.sign_tiles_for(mine, B, blk.n_limbs)
                st = self._sched_cache[key]
            sgn = self.sign_matrix(blk, tiles=st, rule=sign_rule)
        return self.mi_scores(blk, onehot, B, sgn=sgn, out=out, rank=rank, world=world, sink=sink)

    def mi_device_step(self, csr, n_cells, n_genes, n_bins=10, signed=False, rank=0, world=1, out=None,
                       front="broadcast"):
        """One pass of the hot path with the input already resident in HBM (bench.py's `value` leg).
        `csr` is the whole matrix on this rank's device; the sharded front end uses this rank's block
        of cells of it."""
        if front == "sharded" and world > 1:
            from .hostpool import get_pool
            pool = get_pool()
            b0, b1, r0, r1 = self.shard_cells(n_cells, rank, world)
            key = ("slice", csr.values.data_ptr(), rank, world)
            if key not in self._sched_cache:
                e0, e1 = int(csr.row_ptr[r0]), int(csr.row_ptr[r1])
                self._sched_cache[key] = CsrDevice(csr.values[e0:e1], csr.col_idx[e0:e1],
                                                   (csr.row_ptr[r0:r1 + 1] - e0).contiguous(), r1 - r0, n_genes)
            blk = self._shard_front([(self._sched_cache[key], (b0, b1))], n_cells, n_genes, n_bins, signed, rank,
                                    world, pool)
            if blk is not None:
                return self._mi_after_front(blk, signed, rank, world, out, None, "numpy")
        return self.mi_from_csr(csr if rank == 0 else None, n_cells, n_genes, n_bins=n_bins, signed=signed,
                                rank=rank, world=world, out=out)

    def _timed_host(self, name, t0):
        """Host wall-clock phases of the public call (bench.py sets host_timings to a dict)."""
        if getattr(self, "host_timings", None) is not None:
            import time
            self.host_timings.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)

    def _front_from_host_sharded(self, X, n_bins, value_rows_wanted, rank, world, required):
        """Upload this rank's block of cells and run the sharded front end; the BinBlock, or None when
        the data do not qualify (the caller then takes the broadcast path on every rank alike)."""
        from .hostpool import get_pool
        if self.resolve_front("sharded" if required else "auto", world) != "sharded":
            if required:
                raise GvError("sharded front end", -4, "the sharded front end is not available here")
            return None
        pool = get_pool()
        n_cells, n_genes = X.shape
        b0, b1, r0, r1 = self.shard_cells(n_cells, rank, world)
        # the rank's cells may go up in a few chunks of whole 128-cell blocks (the exchange of a chunk
        # then runs under the upload of the next; GV_FRONT_CHUNKS, default one piece)
        nnz_local = int(X.indptr[r1]) - int(X.indptr[r0])
        n_chunks = max(1, min(self.front_chunks, (b1 - b0), nnz_local // self.front_chunk_nnz))

        def chunks():
            for c in range(n_chunks):
                c0, c1 = b0 + (b1 - b0) * c // n_chunks, b0 + (b1 - b0) * (c + 1) // n_chunks
                if c1 > c0 or n_chunks == 1:
                    yield (self.upload_csr(X, rows=(min(n_cells, c0 * 128), min(n_cells, c1 * 128)), pack=True,
                                           alone=False), (c0, c1))

        blk = self._shard_front(chunks(), n_cells, n_genes, n_bins, value_rows_wanted, rank, world, pool)
        if blk is None and required:
            raise GvError("sharded front end", -4, "the data do not qualify for the sharded front end "
                          "(count data <= 127 with value rows / 255 without, in canonical CSR)")
        return blk

    def mi_from_host_sharded(self, X, n_bins=10, signed=False, rank=0, world=1, sign_rule="numpy",
                             sink=None, required=False) -> bool:
        """Sharded front end from a host scipy CSR: every rank uploads its own block of cells over its
        own PCIe link, transposes it, the slabs are exchanged by peer copies over NVLink and every rank
        bins all genes locally; no NCCL call.  Returns False when the data do not qualify (not count data
        <= 127 / 255, peer memory unavailable): the caller then takes the broadcast path."""
        blk = self._front_from_host_sharded(X, n_bins, signed, rank, world, required)
        if blk is None:
            return False
        self._mi_after_front(blk, signed, rank, world, None, sink, sign_rule)
        return True

    def target_from_host_sharded(self, X, kind: int, rank=0, world=1, sink=None) -> bool:
        """Pearson / cosine on the sharded front end: its dense count rows (counts <= 127) are the
        one-digit value-row operand of the co-moment contraction.  False = take the broadcast path
        (Jaccard and Spearman need detection / rank rows, larger counts more digits)."""
        if kind not in (_lib.GV_MOM_PEARSON, _lib.GV_MOM_COSINE):
            return False
        blk = self._front_from_host_sharded(X, 10, True, rank, world, False)
        if blk is None:
            return False
        self.moment_scores(blk, kind, rank=rank, world=world, sink=sink)
        return True

    def target_from_csr(self, csr: Optional[CsrDevice], kind: int, ranks: bool = False,
                        n_cells: Optional[int] = None, n_genes: Optional[int] = None,
                        rank: int = 0, world: int = 1, group=None,
                        out: Optional[torch.Tensor] = None, sink=None) -> torch.Tensor:
        """Pearson / cosine / Jaccard matrix; ranks=True with GV_MOM_PEARSON is Spearman.  With
        world > 1 only rank 0 needs `csr` (pass n_cells / n_genes elsewhere): rank 0 builds the value
        rows, one broadcast replicates the block, every rank computes its slice of the tile schedule."""
        import torch.distributed as dist
        mode = 1 if kind == _lib.GV_MOM_JACCARD else (3 if ranks else 0)
        if csr is not None:
            n_cells, n_genes = csr.n_cells, csr.n_genes
        distributed = world > 1 and dist.is_available() and dist.is_initialized()
        blk = self.discretize(csr, 10, mode) if (rank == 0 or not distributed) else None
        if distributed:
            blk = self._replicate(blk, n_cells, n_genes, 10, mode, group)
        return self.moment_scores(blk, kind, out=out, rank=rank, world=world, sink=sink)


_engines = {}


def get_engine(device=None) -> Engine:
    dev = torch.device(device if device is not None else "cuda")
    idx = dev.index if dev.index is not None else (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    key = (dev.type, idx)
    if key not in _engines:
        _engines[key] = Engine(device)
    return _engines[key]
