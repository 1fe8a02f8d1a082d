#!/usr/bin/env python
"""bench.py -- all-pairs signed mutual information throughput (gene-pairs/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C4|C3|C2|GxN] [--impl reference]

One step = one full pass of the hot path over the synthetic count matrix of the workload:
discretise -> (N > 1: replicate the discretised matrix) -> one-hot expansion -> Pearson-sign tiles ->
tcgen05 contraction with the fused MI epilogue over this rank's slice of the upper-triangular tile
schedule.  `value` times it with the CSR already in HBM; `e2e` times the reference-facing plugin call
itself, genevector_b200.metrics.target_mi(scipy CSR on the host, gene names, signed=True,
backend="b200") (what data.py:319-326 calls), from pageable host arrays to the full ScoreMatrix on
rank 0 -- host -> device copies, result assembly and device -> host copies inside the timed region.
Prints ONE JSON line (rank 0).  The default workload is BASELINE.json's headline configuration,
20,000 genes x 200,000 cells (C4), which fits one B200.

--impl reference times the reference's CPU implementation of the path instead: the C/OpenMP port
of src/lib.rs + discretize_genes + np.corrcoef (oracle/, the Rust original cannot be built here),
all host threads, on a bounded gene sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


class StdoutGuard:
    """Rank 0's stdout carries exactly one JSON line: while the bench runs, file descriptor 1 points
    at stderr (NCCL prints its version banner straight to fd 1), and emit() restores it for the line."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)


N_BINS = 10
L2_BYTES = 126 * 1024 * 1024


def workload_name(wl, G, N, signed=True):
    """The one workload string both arms print (the gene subsample of the CPU arm is in `sample`)."""
    return f"{wl}: all-pairs {'signed ' if signed else ''}MI, {G} genes x {N} cells, n_bins={N_BINS}"


def parse_workload(name):
    from genevector_b200.synth import CONFIGS
    if name in CONFIGS:
        g, n = CONFIGS[name]
        return name, g, n
    g, n = name.lower().split("x")
    return name, int(g), int(n)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"],
                "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured",
                "sm_mhz_sustained": (p.get("clocks_under_load") or {}).get("sm_mhz_median")}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback",
            "sm_mhz_sustained": 1300.0}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.proc.wait()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        # the sampler runs only while the timed region does, so every sample is under load
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_median": float(np.median(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(n_cells, n_genes_workload, steps, warmup, target_s=8.0, seed=0):
    """The reference's CPU path (port) on a bounded gene sample of the workload, all host threads.
    Per step: discretize_genes + np.corrcoef sign + all-pairs MI (lib.rs:30-98) on S genes x all cells."""
    from oracle import oracle as orc
    from genevector_b200.synth import synth_counts
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm sets its own thread count
    # (all host cores) for the C port and for numpy's BLAS
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = orc.set_num_threads(ncpu)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=ncpu)
    except Exception:
        pass

    phases = {"discretize_genes": [], "corrcoef_sign": [], "mi_pairs": []}

    def one_step(Xd):
        t0 = time.perf_counter()
        xd, nb = orc.discretize_genes_c(Xd, N_BINS)
        t1 = time.perf_counter()
        with np.errstate(invalid="ignore", divide="ignore"):
            corr = np.nan_to_num(np.corrcoef(Xd.T), nan=0.0)
        t2 = time.perf_counter()
        _, npairs = orc.mi_pairs_c(xd, nb, corr=corr, sign_mode=1)
        t3 = time.perf_counter()
        for k, v in zip(phases, (t1 - t0, t2 - t1, t3 - t2)):
            phases[k].append(v * 1e3)
        return t3 - t0

    # calibrate on 48 genes, then size the sample for ~target_s per step
    S0 = min(48, n_genes_workload)
    X0 = synth_counts(n_cells, S0, seed=seed).toarray()
    one_step(X0)
    t_cal = one_step(X0)
    rate = (S0 * (S0 - 1) / 2) / max(t_cal, 1e-6)
    S = int(min(n_genes_workload, max(S0, min(1024, np.sqrt(2 * rate * target_s)))))
    X = synth_counts(n_cells, S, seed=seed).toarray() if S != S0 else X0
    for _ in range(max(warmup, 0)):
        one_step(X)
    for v in phases.values():
        v.clear()
    times = [one_step(X) for _ in range(max(steps, 1))]
    pairs = S * (S - 1) / 2
    t = float(np.mean(times))
    return {"value": pairs / t, "unit": "gene-pairs/s", "cores": cores, "kind": "port",
            "phases_ms": {k: float(np.mean(v)) for k, v in phases.items()},
            "sample": f"{S} genes x {n_cells} cells ({int(pairs)} pairs) of the workload's generator, "
                      f"discretize + corrcoef sign + all-pairs MI per step, mean of {len(times)} steps; "
                      "C/OpenMP port of src/lib.rs (oracle/gv_oracle.c): the Rust tier cannot be built here and "
                      "the reference's Numba tier (BASELINE.md 4(i)) is not on the GPU box -- its builder-run "
                      "timing beside this port is in profiles/r02_cpu_numba_vs_port.txt",
            "ms_per_step": t * 1e3}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    guard = StdoutGuard()
    wl, G, N = parse_workload(args.workload)
    r = cpu_reference(N, G, args.steps, args.warmup)
    line = {"metric": "gene-pairs/s", "value": r["value"], "unit": "gene-pairs/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": workload_name(wl, G, N, True),
                       "note": "CPU path timed on a gene sample; pairs/s is per-pair throughput at this cell count"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "phases_ms")},
            "e2e": {"value": r["value"], "unit": "gene-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    guard.emit(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ shared pieces
def setup_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, local, dev


def synthetic_input(N, G, dev, keep_device):
    """The workload's matrix as every caller of the reference holds it: a scipy CSR in pageable host
    memory (drawn on this rank's GPU, same seed on every rank).  keep_device: also return it resident
    on the device for the device-timed leg."""
    import scipy.sparse as sp
    import torch
    from genevector_b200.engine import CsrDevice
    from genevector_b200.synth import synth_counts_torch
    vals, cols, ptr = synth_counts_torch(N, G, seed=0, device=dev)
    X = sp.csr_matrix((vals.cpu().numpy(), cols.cpu().numpy(), ptr.cpu().numpy().astype(np.int32)
                       if int(ptr[-1]) < 2 ** 31 else ptr.cpu().numpy()), shape=(N, G))
    csr = CsrDevice(vals, cols, ptr, N, G) if keep_device else None
    torch.cuda.synchronize()
    return X, csr


def make_timed(world, dev):
    import torch
    import torch.distributed as dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps between barriers; CUDA events on the launching stream and the host clock around
        them (a step that ends in host work is not over when its last kernel is), max over ranks."""
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ms = torch.tensor([max(ev0.elapsed_time(ev1), 0.0), (t1 - t0) * 1e3], dtype=torch.float64, device=dev)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0].item()), float(ms[1].item())

    return barrier, timed


def csr_bytes(X):
    return int(X.data.nbytes + X.indices.nbytes + (X.shape[0] + 1) * 8)


# ------------------------------------------------------------------------------ matrix targets
def run_matrix_target(args):
    """BASELINE configs[4]: pearson / cosine / jaccard (and spearman) on the shared contraction
    skeleton.  Same contract as the MI line; not the default line."""
    guard = StdoutGuard()
    import torch
    import torch.distributed as dist
    from genevector_b200 import _lib
    from genevector_b200 import metrics as gm
    from genevector_b200.engine import get_engine
    from genevector_b200.synth import gene_names

    rank, world, local, dev = setup_dist()
    wl, G, N = parse_workload(args.workload if args.workload != "C4" else "C5")
    pairs = G * (G - 1) // 2
    kind = {"pearson": _lib.GV_MOM_PEARSON, "spearman": _lib.GV_MOM_PEARSON, "cosine": _lib.GV_MOM_COSINE,
            "jaccard": _lib.GV_MOM_JACCARD}[args.target]
    ranks = args.target == "spearman"
    eng = get_engine(dev)
    eng.cta_group = args.cta_group
    X, csr = synthetic_input(N, G, dev, keep_device=(rank == 0))
    genes = gene_names(G)
    out = torch.zeros((G, G), dtype=torch.float32, device=dev)
    flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)
    barrier, timed = make_timed(world, dev)
    public = {"pearson": gm.target_pearson, "spearman": gm.target_spearman, "cosine": gm.target_cosine,
              "jaccard": gm.target_jaccard}[args.target]

    def step_device():
        flush.fill_(1)
        eng.target_from_csr(csr, kind, ranks=ranks, n_cells=N, n_genes=G, rank=rank, world=world, out=out)

    last = [None]

    def step_e2e():
        flush.fill_(1)
        last[0] = public(X, genes, device=dev)

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    eng.kernel_events = []
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, _ = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    kev = eng.kernel_events
    eng.kernel_events = None
    phase_ms = {}
    for name, a, b, _ in kev:
        phase_ms.setdefault(name, []).append(a.elapsed_time(b))
    for _ in range(max(3, args.warmup)):
        step_e2e()
    ms_e2e_ev, ms_e2e_wall = timed(step_e2e, args.steps)
    ms_e2e = max(ms_e2e_ev, ms_e2e_wall)

    mom_ms = float(np.mean(phase_ms["gv_moment_tiles"]))
    alg_ops = pairs / world * 2.0 * N                      # 2*N ops per pair
    achieved = alg_ops / (mom_ms * 1e-3) / 1e12
    peak = eng.mma_peak("int8")
    line = {
        "metric": "gene-pairs/s", "value": pairs / (ms_dev / args.steps * 1e-3), "unit": "gene-pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int8", "data": "synthetic",
        "config": {"workload": f"{wl}: all-pairs {args.target}, {G} genes x {N} cells", "pairs": pairs,
                   "cta_group": args.cta_group, "l2": "L2 flushed between steps",
                   "parallelism": f"tile-sharded x{world}, 1 broadcast" if world > 1 else "single GPU"},
        "e2e": {"value": pairs / (ms_e2e / args.steps * 1e-3), "unit": "gene-pairs/s",
                "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": csr_bytes(X), "d2h_bytes_per_step": int(G * G * 4),
                "note": f"genevector_b200.metrics.target_{args.target}(scipy CSR on the host, gene names) -> full "
                        "ScoreMatrix on rank 0 (pageable host arrays in, shared page-locked host matrix out)"},
        "gpu_launches": int(sum(k for (_, _, _, k) in kev)),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": None, "kernel_ms": mom_ms,
                     "kernel": "gram_tiles_kernel<2, MomentEpi<kind, L>> (gv_moment_tiles)",
                     "peak_note": "tcgen05 kind::i8 issued back to back, measured live (gv_mma_peak); achieved counts "
                                  "2*N ops per gene pair: digit rows beyond the first, padding and the redundant half "
                                  "of the diagonal tiles are not counted"},
        "phases_ms": {k: float(np.mean(v)) for k, v in phase_ms.items()},
        "clocks": clocks,
    }
    if rank == 0:
        guard.emit(json.dumps(line))
    last[0] = None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------------------ B200 arm
def parity_check(eng, csr, sm, G, N, signed, rank, world, n_sample=12, seed=0):
    """Rank 0 recomputes, single rank, a seeded sample of tiles that OTHER ranks own and compares them
    bit for bit with the matrix the ranks assembled in the shared host matrix."""
    import torch
    from genevector_b200 import _lib
    from genevector_b200.engine import schedule_host, rank_bounds
    B = eng.lib.gv_rows_per_gene(N_BINS)
    gpb = 32 // B
    op_rows = eng.lib.gv_onehot_rows(G, B)
    th, tile_m = schedule_host(op_rows, eng.cta_group, eng.band)
    b = rank_bounds(th, tile_m, eng.band, world)
    rng = np.random.default_rng(seed)
    pick = []
    for r in range(1, world):
        if b[r + 1] > b[r]:
            pick += list(rng.integers(b[r], b[r + 1], size=max(1, n_sample // (world - 1))))
    pick = np.array(sorted(set(int(t) for t in pick)), dtype=np.int64)
    tiles_h = np.ascontiguousarray(th[pick])
    blk = eng.discretize(csr, N_BINS, 0 if signed else -1)
    onehot, B = eng.expand_onehot(blk)
    sgn = eng.sign_matrix(blk, tiles=eng.sign_tiles_for(tiles_h, B, blk.n_limbs)) if signed else None
    out = torch.zeros((G, G), dtype=torch.float32, device=eng.device)
    tiles_d = torch.from_numpy(tiles_h).to(eng.device)
    fmt = _lib.GV_OPERAND_FP4 if onehot.dtype == torch.uint8 else _lib.GV_OPERAND_INT8
    _lib.call("gv_mi_tiles", onehot.data_ptr(), onehot.shape[0], onehot.shape[1], B, blk.marg.data_ptr(),
              sgn.data_ptr() if sgn is not None else None, sgn.stride(0) if sgn is not None else 0, G,
              tiles_d.data_ptr(), len(tiles_h), out.data_ptr(), out.stride(0), eng.cta_group, fmt, None,
              torch.cuda.current_stream(eng.device).cuda_stream)
    torch.cuda.synchronize()
    identical, entries = True, 0
    for r0, c0 in tiles_h:
        g0, g1 = r0 // 32 * gpb, min(G, (r0 + tile_m) // 32 * gpb)
        h0, h1 = c0 // 32 * gpb, min(G, (c0 + 256) // 32 * gpb)
        want = out[g0:g1, h0:h1].cpu().numpy()
        got = sm.matrix[g0:g1, h0:h1]
        upper = np.arange(g0, g1)[:, None] < np.arange(h0, h1)[None, :]
        identical &= bool(np.array_equal(want[upper].view(np.uint32), np.asarray(got)[upper].view(np.uint32)))
        mirror = sm.matrix[h0:h1, g0:g1].T
        identical &= bool(np.array_equal(want[upper].view(np.uint32), np.asarray(mirror)[upper].view(np.uint32)))
        entries += int(upper.sum())
    return {"tiles": int(len(tiles_h)), "entries": entries, "identical": bool(identical),
            "note": "tiles owned by ranks 1..N-1, recomputed single-rank on rank 0, compared bitwise (both "
                    "triangles) with the assembled host matrix"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C4")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--unsigned", action="store_true", help="plain MI (no Pearson sign pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-int8-pass", action="store_true", help="skip the short INT8-operand pass (roofline_int8)")
    ap.add_argument("--cta-group", type=int, default=2)
    ap.add_argument("--operand", default="fp4", choices=["int8", "fp4"],
                    help="storage of the one-hot operand (fp4: tcgen05 kind::mxf4, exact counts)")
    ap.add_argument("--front", default="auto", choices=["auto", "broadcast", "sharded"],
                    help="N > 1: how the discretised matrix reaches every rank (DESIGN.md section 5)")
    ap.add_argument("--target", default="mi", choices=["mi", "pearson", "spearman", "cosine", "jaccard"],
                    help="mi (default, the headline) or one of the matrix targets of BASELINE configs[4]")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.target != "mi":
        return run_matrix_target(args)

    guard = StdoutGuard()
    import torch
    import torch.distributed as dist
    from genevector_b200 import metrics as gm
    from genevector_b200.engine import get_engine
    from genevector_b200.synth import gene_names

    rank, world, local, dev = setup_dist()
    signed = not args.unsigned
    wl, G, N = parse_workload(args.workload)
    pairs = G * (G - 1) // 2
    eng = get_engine(dev)
    eng.cta_group = args.cta_group
    eng.operand = args.operand

    # ---- synthetic input: a scipy CSR in pageable host memory on every rank (what the reference's
    # callers hold); the device-timed leg keeps a resident copy
    X, csr = synthetic_input(N, G, dev, keep_device=True)
    nnz = int(X.nnz)
    genes = gene_names(G)
    B = eng.lib.gv_rows_per_gene(N_BINS)
    op_rows = eng.lib.gv_onehot_rows(G, B)
    _, n_my_tiles, n_all_tiles = eng.schedule(op_rows, rank=rank, world=world)
    out = torch.zeros((G, G), dtype=torch.float32, device=dev)
    flush = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)
    operand_bytes = op_rows * ((N + 255) // 256 * 256) // (2 if args.operand == "fp4" else 1)
    flush_needed = operand_bytes <= 2 * L2_BYTES
    barrier, timed = make_timed(world, dev)
    front = eng.resolve_front(args.front, world)

    def step_device():
        if flush_needed:
            flush.fill_(1)
        eng.mi_device_step(csr, N, G, n_bins=N_BINS, signed=signed, rank=rank, world=world, out=out, front=front)

    last = [None]

    def step_e2e():
        if flush_needed:
            flush.fill_(1)
        # the reference-facing plugin call: data.py:319-326 -> TARGETS["mi"](X, gene_names, signed=, backend=, device=)
        last[0] = gm.target_mi(X, genes, signed=signed, backend="b200", device=dev, n_bins=N_BINS, front=front)

    def device_leg(steps, warmup):
        for _ in range(warmup):
            step_device()
        torch.cuda.synchronize()
        eng.kernel_events = []
        ms, _ = timed(step_device, steps)
        kev = eng.kernel_events
        eng.kernel_events = None
        return ms, kev

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    eng.kernel_events = []
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, _ = timed(step_device, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    kev = eng.kernel_events
    eng.kernel_events = None
    mi_ms = [a.elapsed_time(b) for (name, a, b, _) in kev if name == "gv_mi_tiles"]
    n_mi_launches = max(1, len(mi_ms) // max(args.steps, 1))
    phase_ms = {}
    for name, a, b, _ in kev:
        phase_ms.setdefault(name, []).append(a.elapsed_time(b))

    # ---- end to end through the public call (first calls allocate the page-locked buffers)
    for _ in range(max(3, args.warmup)):
        step_e2e()
    eng.host_timings = {}
    ms_e2e_ev, ms_e2e_wall = timed(step_e2e, args.steps)
    ms_e2e = max(ms_e2e_ev, ms_e2e_wall)
    host_ms = {k: float(np.mean(v)) for k, v in (eng.host_timings or {}).items()}
    eng.host_timings = None
    sm = last[0]
    # bytes that crossed the PCIe links in one step: every rank's own upload (count data are packed to
    # 3 bytes per entry by the staging pass) and the result rectangles (G x G float32 in all)
    h2d = torch.tensor([eng.last_upload_bytes], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(h2d)
    h2d_bytes = int(h2d.item())

    # correctness guards inside the bench: symmetric, finite, bounded by log2(11); with several ranks a
    # sample of tiles of the other ranks is recomputed on rank 0 and compared bitwise
    parity = None
    if rank == 0:
        sub = np.asarray(sm.matrix[:512, :512])
        assert np.isfinite(sub).all() and float(np.abs(sub).max()) <= 3.46 + 1e-3
        assert np.array_equal(sub, sub.T)
        if world > 1:
            parity = parity_check(eng, csr, sm, G, N, signed, rank, world)
    if world > 1:
        dist.barrier()

    # ---- roofline of the dominant kernel (contraction + fused MI epilogue), this rank's share
    peaks = load_peaks()

    sm_mhz = (clocks or {}).get("sm_mhz") if rank == 0 else None

    def roofline(operand, mi_list, step_ms):
        fp4 = operand == "fp4"
        my_frac = n_my_tiles / max(n_all_tiles, 1)
        alg_ops = pairs * my_frac * 2.0 * N_BINS * N_BINS * N        # 2*B^2*N per pair
        per_step = float(np.sum(mi_list)) / max(args.steps if mi_list is mi_ms else 1, 1) if mi_list else float("nan")
        achieved = alg_ops / (per_step * 1e-3) / 1e12 if mi_list else None
        long_step = step_ms > 500.0
        rate = 4.0 if fp4 else 2.0
        nominal = 9000.0 if fp4 else 4500.0
        bf16_scaled = rate * (peaks["bf16_sustained"] if long_step else peaks["bf16_burst"])
        peak = eng.mma_peak(operand)
        opb = op_rows * ((N + 255) // 256 * 256) // (2 if fp4 else 1)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            ent = json.load(open(tpath)).get(f"{wl}:{world}:{operand}")
            if ent and signed and args.cta_group == 2:
                traffic = ent["traffic_bytes"]
        return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "algorithmic_operand_bytes": int(opb),
                "kernel": f"gram_tiles_kernel<{args.cta_group}, MiEpi<10{', fp4' if fp4 else ''}>> (gv_mi_tiles)",
                "kernel_ms": per_step, "launches_per_step": n_mi_launches if mi_list is mi_ms else None,
                "peak_note": f"peak = tcgen05 {'kind::mxf4' if fp4 else 'kind::i8'} M256 N240 issued back to back from "
                             "resident shared memory, measured live in this run (gv_mma_peak); "
                             f"{rate:.0f} x {peaks['src']} cuBLAS bf16 ({'sustained' if long_step else 'burst'}) "
                             f"would be {bf16_scaled:.0f} TOP/s, nominal dense {'FP4 9000' if fp4 else 'INT8 4500'} TOP/s; "
                             "achieved counts algorithmic work only (issued work is 128/120 of it)",
                "peak_bf16_scaled": bf16_scaled,
                # the driver's cuBLAS bf16 loop ran power-throttled (dense random operands); the one-hot
                # operands here are 90 % zeros and hold a much higher SM clock under the same power cap:
                # the same tensor-pipe rate at this run's clock is the comparable figure
                "peak_bf16_scaled_at_this_clock": (bf16_scaled * sm_mhz / peaks["sm_mhz_sustained"])
                if (long_step and sm_mhz and peaks.get("sm_mhz_sustained")) else None,
                "clock_note": (f"cuBLAS bf16 sustained was measured at {peaks.get('sm_mhz_sustained')} MHz; this kernel ran at "
                               f"{sm_mhz} MHz (median under load)") if (long_step and sm_mhz) else None,
                "frac_of_nominal": (achieved / nominal) if achieved else None,
                "frac_of_nominal_int8": (achieved / 4500.0) if achieved else None}

    roof = roofline(args.operand, mi_ms, ms_dev / args.steps)
    roof_int8 = None
    if args.operand != "int8" and not args.no_int8_pass:
        # the metric's own dtype ("% INT8 TC peak"): a short pass with the INT8 one-hot operand
        eng.operand = "int8"
        ms8, kev8 = device_leg(1, 1)
        mi8 = [a.elapsed_time(b) for (name, a, b, _) in kev8 if name == "gv_mi_tiles"]
        roof_int8 = roofline("int8", mi8, ms8)
        roof_int8["ms_per_step"] = ms8
        roof_int8["value"] = pairs / (ms8 * 1e-3)
        eng.operand = args.operand

    fp4 = args.operand == "fp4"
    line = {
        "metric": "gene-pairs/s", "value": pairs / (ms_dev / args.steps * 1e-3), "unit": "gene-pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        # arithmetic type of the contraction: one-hot operands in E2M1 (fp4) or INT8, exact integer counts
        "dtype": "fp4" if fp4 else "int8",
        "data": "synthetic",
        "config": {"workload": workload_name(wl, G, N, signed), "nnz": nnz,
                   "pairs": pairs, "cta_group": args.cta_group, "operand": args.operand,
                   "parallelism": (f"tile-sharded x{world}, front={front}" if world > 1 else "single GPU"),
                   "l2": "operand larger than L2" if not flush_needed else "L2 flushed between steps"},
        "e2e": {"value": pairs / (ms_e2e / args.steps * 1e-3), "unit": "gene-pairs/s",
                "ms_per_step": ms_e2e / args.steps,
                "ms_per_step_events": ms_e2e_ev / args.steps, "ms_per_step_wall": ms_e2e_wall / args.steps,
                "h2d_bytes_per_step": h2d_bytes, "host_csr_bytes": csr_bytes(X),
                "d2h_bytes_per_step": int(G) * int(G) * 4,
                "host_phases_ms": host_ms,
                "note": "genevector_b200.metrics.target_mi(scipy CSR in pageable host memory, gene names, signed=True, "
                        "backend='b200') -> full ScoreMatrix on rank 0; all ranks' host->device and device->host "
                        "copies are inside the timed region (bytes are totals over the ranks)"},
        "gpu_launches": int(sum(k for (_, _, _, k) in kev)),
        "roofline": roof,
        "phases_ms": {k: float(np.sum(v)) / args.steps for k, v in phase_ms.items()},
        "clocks": clocks,
    }
    if roof_int8 is not None:
        line["roofline_int8"] = roof_int8
    if parity is not None:
        line["parity_check"] = parity
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference(N, G, steps=1, warmup=0, target_s=10.0)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "phases_ms")}
    if rank == 0:
        guard.emit(json.dumps(line))
    last[0] = sm = None
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
