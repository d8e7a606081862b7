"""CPU suite: host-side logic of the device exchanges (`sharding.NcclComm`, `sharding.PeerWindow`) against a stand-in for the
device context -- host arrays behind integer "device pointers", the library calls of a one-rank world implemented in numpy.
Covers what needs no GPU: the staging blocks of the host-staged all-gather, and that a window either opens on every rank or
raises `PeerUnavailable` everywhere with nothing left allocated.  The exchanges themselves run in tests/test_gpu_multi.py."""
import ctypes as C

import numpy as np
import pytest

from pydeseq2_b200 import _lib
from pydeseq2_b200.sharding import NcclComm, PeerUnavailable, PeerWindow


class _FakeLib:
    """The C entry points the two classes call, for world size 1."""

    def __init__(self, owner, fail_open=False, fail_alloc=False):
        self.o, self.fail_open, self.fail_alloc = owner, fail_open, fail_alloc
        self.pushes = []

    def pdq_comm_init(self, h, buf, world, rank):
        return 0

    def pdq_comm_destroy(self, h):
        return 0

    def pdq_allgather_f64_dev(self, h, send, recv, n):
        self.o.mem[recv.value][:n] = self.o.mem[send.value][:n]
        return 0

    def pdq_peer_window_alloc(self, h, nbytes, w, d, hbuf):
        if self.fail_alloc:
            return -1
        p = self.o.malloc(nbytes + 4096)
        C.cast(w, C.POINTER(C.c_void_p))[0] = p
        C.cast(d, C.POINTER(C.c_void_p))[0] = p + 4096
        C.memmove(hbuf, bytes(range(150, 150 + _lib.PEER_HANDLE_BYTES)), _lib.PEER_HANDLE_BYTES)
        return 0

    def pdq_peer_window_free(self, h, w):
        self.o.free(w.value)
        return 0

    def pdq_peer_group_open(self, h, own, world, rank, handles, out):
        assert bytes(handles.raw[: _lib.PEER_HANDLE_BYTES]) == bytes(range(150, 150 + _lib.PEER_HANDLE_BYTES))
        if self.fail_open:
            return -1
        C.cast(out, C.POINTER(C.c_void_p))[0] = 0xBEEF
        return 0

    def pdq_peer_push_dev(self, h, group, k, send, offs, count):
        self.pushes.append((k, [send[i] for i in range(k)], [offs[i] for i in range(k)], count))
        return 0

    def pdq_peer_status(self, h, group, out):
        C.cast(out, C.POINTER(C.c_uint64))[0] = self.o.status
        return 0

    def pdq_peer_group_close(self, h, group):
        self.o.closed_groups += 1
        return 0

    def pdq_last_error(self, h):
        return b"stand-in failure"


class _FakeCtx:
    def __init__(self, **kw):
        self.h, self.mem, self._next, self.status, self.closed_groups = None, {}, 1 << 20, 0, 0
        self.lib = _FakeLib(self, **kw)

    def check(self, rc):
        if rc != 0:
            raise _lib.B200Error("stand-in failure")

    def malloc(self, nbytes):
        p, self._next = self._next, self._next + ((int(nbytes) + 255) // 256 + 1) * 256
        self.mem[p] = np.full(int(nbytes) // 8 + 1, -7.0)
        return p

    def free(self, p):
        if p:
            del self.mem[p]

    def pinned_empty(self, shape):
        return np.full(shape, -3.0)

    def h2d(self, p, arr):
        self.mem[p][: arr.size] = arr.ravel()

    def d2h(self, arr, p):
        arr.ravel()[:] = self.mem[p][: arr.size]

    def sync(self):
        pass


def _comm(**kw):
    ctx = _FakeCtx(**kw)
    return ctx, NcclComm(ctx, [5], 0, bytes(_lib.UNIQUE_ID_BYTES))


def test_staged_all_gather_returns_exactly_the_sent_length():
    """The staging blocks keep the largest size seen; a later, shorter send must not return their stale tail."""
    ctx, comm = _comm()
    big = np.arange(64, dtype=np.float64)
    np.testing.assert_array_equal(comm._gather_equal(big), big)
    small = np.array([1.0])
    out = comm._gather_equal(small)
    assert out.shape == (1,) and out[0] == 1.0
    v = np.array([4.0, 5.0, 6.0, 7.0, 8.0])
    np.testing.assert_array_equal(comm.allgather(v), v)


def test_peer_window_opens_pushes_and_closes():
    ctx, comm = _comm()
    n0 = len(ctx.mem)
    win = comm.open_window(1024)
    assert isinstance(win, PeerWindow) and win.data == win.window + 4096
    win.push([(111, 0), (222, 512)], 17)
    win.push([], 0)  # bare barrier
    assert ctx.lib.pushes == [(2, [111, 222], [0, 512], 17), (0, [], [], 0)]
    win.check()
    ctx.status = 2  # the push kernel gave up waiting for rank 1
    with pytest.raises(RuntimeError, match="rank 1"):
        win.check()
    win.close()
    assert ctx.closed_groups == 1 and win.window is None
    assert len(ctx.mem) == n0 + 2  # only the communicator's staging pair is left


@pytest.mark.parametrize("kw", [{"fail_open": True}, {"fail_alloc": True}])
def test_peer_window_unavailable_leaves_nothing_behind(kw):
    ctx, comm = _comm(**kw)
    with pytest.raises(PeerUnavailable):
        comm.open_window(1024)
    live = [p for p, a in ctx.mem.items() if a.size > 200]  # a window would be > 4096 bytes
    assert live == []
