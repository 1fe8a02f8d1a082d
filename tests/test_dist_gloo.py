"""world_size-2 gloo test of the multi-rank host logic on CPU: rank 0 owns the bin block, one
broadcast replicates it, each rank takes its slice of the tile schedule, the partial score
matrices are disjoint and assemble to the full result.  The per-tile arithmetic is stood in for
by the oracle (CPU test infrastructure); the CUDA kernels are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from genevector_b200 import engine as E
    from genevector_b200 import _lib
    from oracle import oracle as orc
    z = load_golden("synth_counts_700x96_b10")
    n_cells, n_genes = z["X"].shape
    n_bins, val_mode = 10, -1
    kp, pitch, lay, total = E.block_layout(n_cells, n_genes, val_mode)
    buf = torch.zeros(total, dtype=torch.uint8)
    if rank == 0:
        # rank 0 "discretises": fill the block exactly as gv_discretize_csr lays it out
        blk = E._views(buf, n_cells, n_genes, n_bins, val_mode, 1, 7)
        bins_gm, nb, _ = orc.discretize_csr_c(z["X"], n_bins)
        packed = np.zeros((n_genes, pitch), dtype=np.uint8)
        padded = np.zeros((n_genes, pitch * 2), dtype=np.uint8)
        padded[:, :n_cells] = bins_gm
        packed[:] = padded[:, 0::2] | (padded[:, 1::2] << 4)
        blk.bins.copy_(torch.from_numpy(packed))
        blk.n_bins_per_gene.copy_(torch.from_numpy(nb))
        marg = np.zeros((n_genes, 32), dtype=np.int32)
        for a in range(1, 32):
            marg[:, a] = (bins_gm == a).sum(axis=1)
        marg[:, 0] = (bins_gm > 0).sum(axis=1)
        blk.marg.copy_(torch.from_numpy(marg))
    E.broadcast_block(buf, total)                       # the one collective
    blk = E._views(buf, n_cells, n_genes, n_bins, val_mode, 1, 7)
    x_disc = blk.x_disc()
    assert np.array_equal(x_disc, z["x_disc"].astype(np.int32)), "block did not survive the broadcast"
    assert np.array_equal(blk.n_bins_per_gene.numpy(), z["n_bins_per_gene"])
    # this rank's tiles -> the gene pairs it owns
    lib = _lib.load()
    B = lib.gv_rows_per_gene(n_bins)
    op_rows = lib.gv_onehot_rows(n_genes, B)
    # 96 genes are a single 256-row tile; use 32-row granularity bands by testing a larger virtual
    # operand too (schedule logic only)
    tiles, n_all = E.tile_slice(op_rows, 2, 8, rank, world)
    big, n_big = E.tile_slice(213376, 2, 8, rank, world)
    counts = torch.tensor([len(tiles), len(big)], dtype=torch.int64)
    dist.all_reduce(counts)
    assert int(counts[0]) == n_all and int(counts[1]) == n_big
    gpb = 32 // B
    full, _ = orc.mi_pairs_c(x_disc, blk.n_bins_per_gene.numpy())
    part = np.zeros((n_genes, n_genes))
    owned = np.zeros((n_genes, n_genes), dtype=np.int64)
    for r0, c0 in tiles:
        gi = np.arange(r0 // 32 * gpb, min(n_genes, (r0 + 256) // 32 * gpb))
        gj = np.arange(c0 // 32 * gpb, min(n_genes, (c0 + 256) // 32 * gpb))
        for i in gi:
            for j in gj:
                if i < j:
                    part[i, j] = part[j, i] = full[i, j]
                    owned[i, j] += 1
    t_part, t_owned = torch.from_numpy(part), torch.from_numpy(owned)
    dist.reduce(t_part, dst=0)
    dist.reduce(t_owned, dst=0)
    if rank == 0:
        iu = np.triu_indices(n_genes, 1)
        assert np.all(t_owned.numpy()[iu] == 1), "a pair is owned by zero or two ranks"
        assert np.array_equal(t_part.numpy(), full)

    # ---- result assembly without a collective: every rank writes the rectangles of its own tiles
    # into the one shared host matrix (hostpool.py; not page-locked here: no CUDA on the CPU box)
    from genevector_b200.hostpool import get_pool
    pool = get_pool(cuda=False)
    full32 = full.astype(np.float32)
    G2 = 1500                                             # several bands and tile columns per rank
    rng = np.random.default_rng(5)
    big_full = rng.standard_normal((G2, G2)).astype(np.float32)
    rows2 = lib.gv_onehot_rows(G2, B)
    held = []
    for it, (G, M, rows) in enumerate([(n_genes, full32, op_rows), (G2, big_full, rows2), (G2, big_full, rows2),
                                       (n_genes, full32, op_rows)]):
        seq, slot, mapping = pool.begin(G * G * 4)
        dst = mapping.array(np.float32, (G, G))
        if it != 2:
            dst[:] = -7.0 if rank == 0 else dst             # stale contents must be overwritten
        dist.barrier()                                      # (test only: rank 0's fill before the writes)
        for _, _, rects in E.host_pieces(rows, 2, 8, gpb, G, rank, world, chunks=3):
            for r0, r1, c0, c1 in rects:
                dst[r0:r1, c0:c1] = M[r0:r1, c0:c1]
        pool.finish(seq, slot)
        got = pool.matrix(slot, G)
        assert np.array_equal(got, M), f"rank {rank}: assembled matrix differs (call {it})"
        if it == 1:
            held.append((slot, got))                        # keep this result alive: its slot is not reused
        elif it == 2:
            assert slot != held[0][0], "a slot still referenced on some rank was handed out again"
            assert np.array_equal(held[0][1], big_full)
        del got, dst
        dist.barrier()
    assert not os.path.exists("/dev/shm" + pool.base + ".s0.g1"), "segment names are unlinked once mapped"
    if rank == 0:
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_partition_broadcast_and_assembly(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
