"""Bottom-up GPU bring-up probe for the tcgen05 Gram-tile kernel (not a test, not a benchmark).

Runs gv_gram_dump on random 0/1 INT8 operands for cta_group 1 and 2 and compares the raw
INT32 accumulators with an exact float matmul on the same device.  Prints a diagnosis that is
readable from the tail of a gpurun log.  Usage: python tools/gpu_probe.py [cg ...]
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genevector_b200 import _lib  # noqa: E402


def schedule(op_rows, tile_m, band=8):
    lib = _lib.load()
    n = lib.gv_tile_schedule(op_rows, tile_m, band, None, 0)
    buf = np.zeros((n, 2), dtype=np.int32)
    lib.gv_tile_schedule(op_rows, tile_m, band, buf.ctypes.data, n)
    return buf


def probe(cg, rows, kp, density=0.3, seed=0, vmax=1):
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(seed)
    if vmax == 1:
        a = (torch.rand(rows, kp, generator=g) < density).to(torch.int8)
    else:
        a = torch.randint(0, vmax + 1, (rows, kp), generator=g).to(torch.int8)
    a_d = a.to(dev).contiguous()
    tiles = schedule(rows, lib.gv_tile_m(cg))
    tiles_d = torch.from_numpy(tiles).to(dev)
    out = torch.full((rows, rows), -7, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    t0 = time.time()
    st = lib.gv_gram_dump(a_d.data_ptr(), rows, kp, tiles_d.data_ptr(), len(tiles), out.data_ptr(),
                          rows, cg, 32, 0, torch.cuda.current_stream().cuda_stream)
    if st != 0:
        print(f"[cg={cg}] launch status {st}: {lib.gv_last_error().decode()}")
        return False
    try:
        torch.cuda.synchronize()
    except RuntimeError as e:  # trap from the mbarrier watchdog, illegal address, ...
        print(f"[cg={cg}] kernel fault: {e}")
        return False
    dt = time.time() - t0
    ref = (a_d.double() @ a_d.double().T).to(torch.int64)
    got = out.to(torch.int64)
    tm = lib.gv_tile_m(cg)
    # only tiles of the schedule are written: compare there
    mask = torch.zeros(rows, rows, dtype=torch.bool, device=dev)
    for r0, c0 in tiles:
        mask[r0:r0 + tm, c0:c0 + 256] = True
    bad = (got != ref) & mask
    nbad = int(bad.sum())
    print(f"[cg={cg}] rows={rows} kp={kp} tiles={len(tiles)} time={dt*1e3:.2f} ms "
          f"mismatches={nbad}/{int(mask.sum())}")
    if nbad:
        idx = bad.nonzero()[:12].tolist()
        for i, j in idx:
            print(f"   ({i},{j}) got {int(got[i, j])} want {int(ref[i, j])}")
        # structure of the error: which rows / cols are affected
        print("   bad rows (first 16):", sorted(set(int(x) for x in bad.any(1).nonzero().flatten()[:16].tolist())))
        print("   bad cols (first 16):", sorted(set(int(x) for x in bad.any(0).nonzero().flatten()[:16].tolist())))
        unwritten = int(((got == -7) & mask).sum())
        print("   unwritten entries inside scheduled tiles:", unwritten)
    return nbad == 0


def main():
    lib = _lib.load()
    print("device:", torch.cuda.get_device_name(0), "check:", lib.gv_device_check(),
          lib.gv_last_error().decode())
    cgs = [int(x) for x in sys.argv[1:]] or [1, 2]
    ok = True
    for cg in cgs:
        # one tile, one K block, then many tiles / many K blocks (pipeline wrap, both accumulators)
        for rows, kp in [(256, 128), (256, 512), (512, 1024), (1312, 4096)]:
            ok &= probe(cg, rows, kp)
        ok &= probe(cg, 512, 512, vmax=127)
    print("PROBE", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
