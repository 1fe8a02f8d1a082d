"""Shared host result matrices for one process per GPU on one box.

north_star: "each rank writes its own result tiles" and there is no collective besides the one
broadcast.  The G x G score matrix the reference hands to its consumers (metrics.py:137-140) lives
on the host, so with several ranks it is ONE POSIX shared-memory segment that every process maps
and page-locks (gv_shm_open -> cudaHostRegister); each rank DMAs the rectangles of its own tiles
straight into it (Engine._tiles_to_host) and every rank then reads the whole matrix through its own
mapping.  Completion is a flag per rank in a small control segment (release / acquire stores and
loads, gv_flag_*): no reduction, no gather, no NCCL call.

Protocol (all words are uint64 in the control segment):
  rank 0, entering call number `seq`, picks a result slot that no rank still holds a view of (or
  creates / enlarges one) and publishes (slot, generation, bytes) followed by announce = seq;
  every rank waits for announce >= seq, maps the slot if it has not mapped that generation yet,
  marks the slot held, computes and copies, synchronises its stream and stores done[rank] = seq;
  every rank waits for all done[*] >= seq and wraps its mapping in the ScoreMatrix it returns; when
  that object is garbage collected the rank clears its hold, which makes the slot reusable.
Segment names are unlinked as soon as every rank has mapped them, so a crashed run leaves nothing
behind in /dev/shm.
"""
from __future__ import annotations

import atexit
import ctypes as C
import os
import weakref

import numpy as np

from . import _lib

_CTL_WORDS = 4096
_MAX_SLOTS = 8
_W_MAGIC, _W_WORLD, _W_ANNOUNCE, _W_SLOT, _W_GEN, _W_BYTES = 0, 1, 2, 3, 4, 5
_W_DONE = 64              # + rank
_W_HELD = 1024            # + slot * 64 + rank     (world <= 64)
_W_MAPPED = 2048          # + slot * 64 + rank: generation of the slot this rank has mapped
_W_SCRATCH = 3072         # + rank * 16 + k: small per-rank words other ranks read (front-end hand-shakes)
_MAGIC = 0x67766232303068


class _Mapping:
    def __init__(self, name, nbytes, create, cuda):
        self.name, self.nbytes, self.cuda = name, int(nbytes), bool(cuda)
        p = C.c_void_p()
        _lib.call("gv_shm_open", name.encode(), self.nbytes, 1 if create else 0, 1 if cuda else 0, C.byref(p))
        self.ptr = p.value
        self.closed = False

    def array(self, dtype, shape, offset=0):
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        a = np.frombuffer(buf, dtype=dtype, count=n // np.dtype(dtype).itemsize, offset=offset)
        return a.reshape(shape)

    def unlink(self):
        try:
            _lib.call("gv_shm_close", None, 0, 0, self.name.encode())     # shm_unlink only
        except Exception:
            pass

    def close(self):
        if not self.closed and self.ptr:
            self.closed = True
            try:
                _lib.call("gv_shm_close", self.ptr, self.nbytes, 1 if self.cuda else 0, None)
            except Exception:
                pass


class SharedResultPool:
    """One per process group; see the module docstring."""

    def __init__(self, rank: int, world: int, base: str, cuda: bool = True, timeout_ms: int = 600000):
        if world > 64:
            raise ValueError("the shared result pool supports up to 64 ranks per box")
        self.rank, self.world, self.base, self.cuda, self.timeout_ms = rank, world, base, cuda, timeout_ms
        self.seq = 0
        self.slots = {}           # slot -> (generation, _Mapping)
        self._unlinked = set()
        # state of the sharded front end that lives and dies with this control segment (engine.py)
        self.peer_blocks, self.peer_epoch, self.front_seq = {}, 0, 0
        self.ctl = None
        if rank == 0:
            self.ctl = _Mapping(base + ".ctl", _CTL_WORDS * 8, True, False)
            w = self.ctl.array(np.uint64, (_CTL_WORDS,))
            w[:] = 0
            w[_W_WORLD] = world
            self._store(_W_MAGIC, _MAGIC)
        atexit.register(self.close)

    # the control segment of the other ranks is opened lazily: rank 0 creates it in its constructor and
    # the name reaches the others only after that (setup hand-shake in get_pool)
    def _ctl(self):
        if self.ctl is None:
            self.ctl = _Mapping(self.base + ".ctl", _CTL_WORDS * 8, False, False)
        return self.ctl

    def _addr(self, word):
        return self._ctl().ptr + 8 * word

    def _store(self, word, value):
        _lib.call("gv_flag_store", self._addr(word), int(value))

    def _load(self, word):
        return int(_lib.load().gv_flag_load(self._addr(word)))

    def _wait_ge(self, word, value):
        _lib.call("gv_flag_wait_ge", self._addr(word), int(value), self.timeout_ms)

    def _slot_name(self, slot, gen):
        return f"{self.base}.s{slot}.g{gen}"

    # ------------------------------------------------------------------ protocol
    def begin(self, nbytes: int):
        """Start call number seq: agree on the result slot and return (seq, slot, mapping)."""
        self.seq += 1
        seq = self.seq
        nbytes = int(nbytes)
        if self.rank == 0:
            slot = gen = None
            n_slots = len(self.slots)
            for s, (g, m) in sorted(self.slots.items()):
                held = any(self._load(_W_HELD + s * 64 + r) for r in range(self.world))
                if not held:
                    slot, gen = s, g
                    if m.nbytes < nbytes:
                        gen = None            # free but too small: replace it by a new generation
                    break
            if slot is None:
                if n_slots >= _MAX_SLOTS:
                    raise RuntimeError(f"{_MAX_SLOTS} shared result matrices are still referenced; "
                                       "release earlier ScoreMatrix objects first")
                slot = n_slots
            if gen is None:
                old = self.slots.pop(slot, None)
                gen = (old[0] + 1) if old else 1
                if old:
                    old[1].close()
                m = _Mapping(self._slot_name(slot, gen), nbytes, True, self.cuda)
                self.slots[slot] = (gen, m)
                self._store(_W_MAPPED + slot * 64, gen)
            self._store(_W_SLOT, slot)
            self._store(_W_GEN, gen)
            self._store(_W_BYTES, self.slots[slot][1].nbytes)
            self._store(_W_ANNOUNCE, seq)
        else:
            self._wait_ge(_W_ANNOUNCE, seq)
            slot, gen, sb = self._load(_W_SLOT), self._load(_W_GEN), self._load(_W_BYTES)
            cur = self.slots.get(slot)
            if cur is None or cur[0] != gen:
                if cur is not None:
                    cur[1].close()
                self.slots[slot] = (gen, _Mapping(self._slot_name(slot, gen), sb, False, self.cuda))
                self._store(_W_MAPPED + slot * 64 + self.rank, gen)
        self._store(_W_HELD + slot * 64 + self.rank, 1)
        return seq, slot, self.slots[slot][1]

    def finish(self, seq: int, slot: int):
        """This rank's copies into the slot are complete (its stream is synchronised): publish that and
        wait for every other rank."""
        self._store(_W_DONE + self.rank, seq)
        for r in range(self.world):
            self._wait_ge(_W_DONE + r, seq)
        if self.rank == 0:
            gen, m = self.slots[slot]
            key = (slot, gen)
            if key not in self._unlinked and all(self._load(_W_MAPPED + slot * 64 + r) == gen
                                                 for r in range(self.world)):
                m.unlink()                    # every rank has it mapped: the name is no longer needed
                self._unlinked.add(key)

    def matrix(self, slot: int, n: int) -> np.ndarray:
        """float32 [n][n] view of the slot; the slot stays held until the view (and every array derived
        from it) is gone."""
        gen, m = self.slots[slot]
        # numpy keeps the ctypes buffer alive as the base of every derived view, so the finalizer (which
        # clears this rank's hold on the slot) runs when the data are no longer referenced anywhere
        buf = (C.c_uint8 * m.nbytes).from_address(m.ptr)
        weakref.finalize(buf, self.release, slot, gen)
        return np.frombuffer(buf, dtype=np.float32, count=n * n).reshape(n, n)

    def release(self, slot: int, gen: int):
        cur = self.slots.get(slot)
        if self.ctl is not None and not self.ctl.closed and cur is not None and cur[0] == gen:
            try:
                self._store(_W_HELD + slot * 64 + self.rank, 0)
            except Exception:
                pass

    # small per-rank words for other hand-shakes (the sharded front end)
    def scratch_store(self, k: int, value: int):
        self._store(_W_SCRATCH + self.rank * 16 + k, value)

    def scratch_addr(self, rank: int, k: int) -> int:
        return self._addr(_W_SCRATCH + rank * 16 + k)

    def scratch_wait_all(self, k: int, value: int):
        for r in range(self.world):
            self._wait_ge(_W_SCRATCH + r * 16 + k, value)

    def scratch_load(self, rank: int, k: int) -> int:
        return self._load(_W_SCRATCH + rank * 16 + k)

    def close(self):
        for s, (g, m) in list(self.slots.items()):
            if self.rank == 0 and (s, g) not in self._unlinked:
                m.unlink()
            m.close()
        self.slots.clear()
        if self.ctl is not None:
            if self.rank == 0:
                self.ctl.unlink()
            self.ctl.close()


_pools = {}


def get_pool(group=None, cuda: bool = True) -> SharedResultPool:
    """The pool of a process group.  The first call is a setup hand-shake (one broadcast of the segment
    base name, like the creation of the NCCL communicator itself); no later call uses the group."""
    import torch.distributed as dist
    if group is not None:
        key = id(group)
    else:
        try:                                  # a re-initialised default group gets a pool of its own
            key = ("world", id(dist.distributed_c10d._get_default_group()))
        except Exception:
            key = ("world", dist.get_world_size())
    if key in _pools:
        return _pools[key]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lw = os.environ.get("LOCAL_WORLD_SIZE")
    if lw is not None and int(lw) != world and group is None:
        raise RuntimeError("genevector_b200 shards over the GPUs of ONE box (shared host result matrix); "
                           f"LOCAL_WORLD_SIZE={lw} != WORLD_SIZE={world}")
    base = [None]
    pool = None
    if rank == 0:
        base[0] = f"/gvb200_{os.getuid()}_{os.getpid()}_{int.from_bytes(os.urandom(4), 'little'):08x}"
        pool = SharedResultPool(rank, world, base[0], cuda=cuda)
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast_object_list(base, src=src, group=group)
    if rank != 0:
        pool = SharedResultPool(rank, world, base[0], cuda=cuda)
    _pools[key] = pool
    return pool
