"""Host-side logic of the peer-memory data plane that needs no GPU: the NVLS multicast (re)build / reuse agreement across
quorums (fake driver calls, two ranks as threads over a dict store) and the HSDP gradient arena allocator."""
import json
import threading
from datetime import timedelta
from types import SimpleNamespace

import torch

from torchft_b200.parallel.hsdp import SymmetricShardGrads
from torchft_b200.parallel.symm_mem import SymmetricComm


class _Store:
    def __init__(self):
        self.d, self.cv = {}, threading.Condition()

    def set(self, k, v):
        with self.cv:
            self.d[k] = v.encode() if isinstance(v, str) else v
            self.cv.notify_all()

    def get(self, k):
        with self.cv:
            assert self.cv.wait_for(lambda: k in self.d, timeout=10), k
            return self.d[k]

    def add(self, k, n):
        with self.cv:
            self.d[k] = self.d.get(k, 0) + n
            return self.d[k]


class _FakeDriver:
    """Counts multicast driver calls; handles are just integers."""

    def __init__(self):
        self.created = self.imported = self.bound = self.released = 0
        self._next = 100

    def mc_create(self, world, nbytes):
        self.created += 1
        self._next += 1
        return self._next, self._next + 1000

    def fetch_fd(self, server, key):
        return hash((server, key)) & 0xffff

    def mc_import(self, fd):
        self.imported += 1
        return fd

    def mc_add_device(self, mc):
        pass

    def mc_bind_and_map(self, mc, mem, nbytes):
        self.bound += 1
        return 0x7000_0000 + mc

    def mc_release(self, mc, va, size):
        self.released += 1


def _rank(rank):
    c = SymmetricComm.__new__(SymmetricComm)
    c._K = _FakeDriver()
    c._segments = {n: SimpleNamespace(nbytes=1 << 20, mem_handle=7) for n in ("core", "z1_grad", "z1_param")}
    c._mc, c._mc_gen, c._mc_members = {}, "", ""
    c._epoch, c._flag = 1, 16
    c._timeout = timedelta(seconds=10)
    c._fdserver = SimpleNamespace(publish=lambda k, fd: None, unpublish=lambda k: None)
    return c


def _configure(comms, store, pids):
    descs = [{"host": "h", "pid": p, "fd_server": "srv0", "mc_gen": c._mc_gen} for c, p in zip(comms, pids)]
    errs = []

    def run(r):
        try:
            comms[r]._setup_multicast(store, descs, ["z1_grad", "z1_param"], r, len(comms))
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=run, args=(r,)) for r in range(len(comms))]
    [t.start() for t in ts]
    [t.join(20) for t in ts]
    assert not errs, errs


def test_multicast_objects_are_built_once_and_kept_while_members_are_unchanged():
    comms = [_rank(0), _rank(1)]
    _configure(comms, _Store(), [11, 22])
    assert comms[0]._K.created == 2 and comms[1]._K.imported == 2 and all(c._K.bound == 2 for c in comms)
    assert comms[0]._mc_gen == comms[1]._mc_gen != "" and sorted(comms[0]._mc) == ["z1_grad", "z1_param"]
    # same members, everybody still holds the same creation: nothing is rebuilt (no store traffic either)
    for c in comms:
        c._epoch += 1
    empty = _Store()
    _configure(comms, empty, [11, 22])
    assert comms[0]._K.created == 2 and all(c._K.bound == 2 and c._K.released == 0 for c in comms) and not empty.d


def test_multicast_objects_are_rebuilt_by_everybody_when_one_rank_lost_them_or_members_changed():
    comms = [_rank(0), _rank(1)]
    _configure(comms, _Store(), [11, 22])
    gen1 = comms[0]._mc_gen
    comms[1]._release_multicast()  # e.g. rank 1 sat in a quorum of its own in between
    for c in comms:
        c._epoch += 1
    _configure(comms, _Store(), [11, 22])
    assert comms[0]._K.created == 4 and comms[0]._K.released == 2 and all(c._K.bound == 4 for c in comms)
    assert comms[0]._mc_gen == comms[1]._mc_gen != gen1
    # a member was replaced by a new process (different pid, no objects): rebuild again
    fresh = _rank(1)
    for c in comms:
        c._epoch += 1
    _configure([comms[0], fresh], _Store(), [11, 33])
    assert comms[0]._K.created == 6 and fresh._K.imported == 2 and json.loads(comms[0]._mc_members)[1] == ["h", 33]


def test_hsdp_gradient_arena_hands_out_outputs_only_and_rewinds():
    arena = SymmetricShardGrads(torch.zeros(4096, dtype=torch.uint8))
    base = arena._arena.untyped_storage().data_ptr()
    cpu = torch.device("cpu")
    for step in range(2):
        arena.begin_step()
        inp = arena.allocate((200,), dtype=torch.float32, device=cpu)   # reduce-scatter input: ordinary memory
        out = arena.allocate((100,), dtype=torch.float32, device=cpu)   # output: carved from the arena, 256 B aligned
        inp2 = arena.allocate((64,), dtype=torch.float32, device=cpu)
        out2 = arena.allocate((32,), dtype=torch.float32, device=cpu)
        assert inp.untyped_storage().data_ptr() != base and inp2.untyped_storage().data_ptr() != base
        assert out.data_ptr() == base and out2.data_ptr() == base + 512 and out.shape == (100,) and out.dtype == torch.float32
    # too large for what is left: falls back to an ordinary allocation and counts it
    arena.allocate((8,), dtype=torch.float32, device=cpu)
    big = arena.allocate((4096,), dtype=torch.float32, device=cpu)
    assert arena.fallbacks == 1 and big.untyped_storage().data_ptr() != base
