"""Gene shards across GPUs: one process per GPU, contiguous gene blocks, the design matrix and
size factors replicated.  Every hot-path method is independent per gene; the ONLY cross-gene dependency on
the path is the dispersion trend + prior (``dds.py:799-884``), which need the genewise dispersions and
normalised means of all genes -- one small all-gather of two float64 vectors.

``NcclComm`` runs that all-gather with NCCL on device buffers through the C ABI (``pdq_allgather_f64_dev``);
``PeerWindow`` is the single-node fast path on top of it: one kernel per exchange that stores the shard's vectors straight into
every peer's HBM over NVLink and waits for the peers' stores (copy + barrier, no NCCL on the data path);
``TorchDistComm`` does the same through ``torch.distributed`` (used with the ``gloo`` backend by the CPU tests
of the multi-rank host logic).
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_bounds(G: int, world: int, rank: int):
    """Contiguous block ``[lo, hi)`` of rank ``rank``: blocks of ceil(G / world) genes, the tail may be short."""
    per = -(-G // world)
    lo = min(G, rank * per)
    return lo, min(G, lo + per)


def shard_sizes(G: int, world: int):
    return [shard_bounds(G, world, r)[1] - shard_bounds(G, world, r)[0] for r in range(world)]


class _PaddedGather:
    """All-gather of ragged per-rank vectors through an equal-count primitive (pad with NaN, strip after)."""

    sizes: list
    rank: int

    def local_slice(self) -> slice:
        """Position of this rank's genes in a gathered (full-length) vector."""
        lo = int(sum(self.sizes[: self.rank]))
        return slice(lo, lo + int(self.sizes[self.rank]))

    def _gather_equal(self, send: np.ndarray) -> np.ndarray:  # (world * len(send),)
        raise NotImplementedError

    def allgather(self, v: np.ndarray) -> np.ndarray:
        m = max(self.sizes)
        send = np.full(m, np.nan)
        send[: len(v)] = v
        out = self._gather_equal(send).reshape(len(self.sizes), m)
        return np.concatenate([out[r, :n] for r, n in enumerate(self.sizes)])

    def allgather_pair(self, a: np.ndarray, b: np.ndarray):
        m = max(self.sizes)
        send = np.full(2 * m, np.nan)
        send[: len(a)] = a
        send[m: m + len(b)] = b
        out = self._gather_equal(send).reshape(len(self.sizes), 2, m)
        return (np.concatenate([out[r, 0, :n] for r, n in enumerate(self.sizes)]),
                np.concatenate([out[r, 1, :n] for r, n in enumerate(self.sizes)]))


    def allgather_table(self, cols: dict) -> dict:
        """All-gather of several per-gene arrays at once (``(n_local,)`` or ``(n_local, k)``): ONE collective on a packed
        buffer; returns full-length arrays in rank order.  This is the end-of-call exchange of SURVEY.md §8(e): after it every
        rank holds the complete per-gene result tables."""
        m = max(self.sizes)
        names = list(cols)
        widths = [1 if np.ndim(cols[k]) == 1 else int(np.shape(cols[k])[1]) for k in names]
        send = np.full((sum(widths), m), np.nan)
        row = 0
        for k, w in zip(names, widths):
            v = np.asarray(cols[k], dtype=np.float64)
            send[row:row + w, : v.shape[0]] = v.reshape(v.shape[0], w).T
            row += w
        out = self._gather_equal(send.ravel()).reshape(len(self.sizes), sum(widths), m)
        res, row = {}, 0
        for k, w in zip(names, widths):
            full = np.concatenate([out[r, row:row + w, :n].T for r, n in enumerate(self.sizes)])
            res[k] = full[:, 0] if np.ndim(cols[k]) == 1 else full
            row += w
        return res


class TorchDistComm(_PaddedGather):
    """torch.distributed flavour (gloo on CPU for tests; also works with nccl + cuda tensors)."""

    def __init__(self, sizes, group=None, device="cpu"):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.sizes = list(sizes)
        self.device = device
        self.rank = dist.get_rank(group)
        self.world = len(self.sizes)

    def _gather_equal(self, send):
        import torch

        t = torch.from_numpy(np.ascontiguousarray(send)).to(self.device)
        outs = [torch.empty_like(t) for _ in self.sizes]
        self.dist.all_gather(outs, t, group=self.group)
        return np.concatenate([o.cpu().numpy() for o in outs])


class NcclComm(_PaddedGather):
    """NCCL all-gather on device buffers through the C ABI.  ``unique_id`` must be the same 128 bytes on all ranks
    (rank 0: ``NcclComm.make_unique_id(ctx)``; ship it with any out-of-band channel, e.g. a torch.distributed
    broadcast or a file)."""

    def __init__(self, ctx, sizes, rank: int, unique_id: bytes):
        from . import _lib

        self.ctx = ctx
        self.sizes = list(sizes)
        self.rank = rank
        self._c_dptr = _lib.c_dptr
        buf = C.create_string_buffer(bytes(unique_id), _lib.UNIQUE_ID_BYTES)
        ctx.check(ctx.lib.pdq_comm_init(ctx.h, buf, len(self.sizes), rank))
        self._cap = 0
        self._send = self._recv = None
        self.world = len(self.sizes)
        self.max_size = max(self.sizes)

    @staticmethod
    def make_unique_id(ctx) -> bytes:
        from . import _lib

        buf = C.create_string_buffer(_lib.UNIQUE_ID_BYTES)
        ctx.check(ctx.lib.pdq_comm_unique_id(ctx.h, buf))
        return buf.raw

    def _gather_equal(self, send):
        n = len(send)
        world = len(self.sizes)
        if n > self._cap:
            if self._send:
                self.ctx.free(self._send)
                self.ctx.free(self._recv)
            self._send = self.ctx.malloc(n * 8)
            self._recv = self.ctx.malloc(n * 8 * world)
            self._cap = n
            self._hs = self.ctx.pinned_empty((n,))
            self._hr = self.ctx.pinned_empty((n * world,))
        hs, hr = self._hs[:n], self._hr[: n * world]  # the staging blocks keep the largest size seen
        hs[:] = send
        self.ctx.h2d(self._send, hs)
        self.ctx.check(self.ctx.lib.pdq_allgather_f64_dev(self.ctx.h, self._c_dptr(self._send), self._c_dptr(self._recv), n))
        self.ctx.d2h(hr, self._recv)
        self.ctx.sync()
        return hr.copy()

    def allgather_dev(self, pairs, count):
        """Device-to-device all-gather of several equal-length vectors as ONE NCCL group (a single fused launch on the context's
        stream, capturable into a CUDA graph).  ``pairs`` = [(send_ptr, recv_ptr), ...], every send buffer holds ``count``
        doubles -- the largest shard's length; shorter shards keep NaN behind their last gene (written once at upload time,
        never per step) -- and every recv buffer ``world * count``."""
        k = len(pairs)
        send = (C.c_void_p * k)(*[p[0] for p in pairs])
        recv = (C.c_void_p * k)(*[p[1] for p in pairs])
        self.ctx.check(self.ctx.lib.pdq_allgather_multi_f64_dev(self.ctx.h, k, send, recv, int(count)))

    def barrier(self):
        self._gather_equal(np.zeros(1))

    def open_window(self, data_bytes: int) -> "PeerWindow":
        """Collective: every rank allocates a receive window of ``data_bytes`` and maps its peers' (``PeerWindow``)."""
        return PeerWindow(self, data_bytes)

    def close(self):
        self.ctx.check(self.ctx.lib.pdq_comm_destroy(self.ctx.h))


class PeerUnavailable(RuntimeError):
    """The ranks cannot map each other's device memory (no CUDA IPC / peer access): callers keep the NCCL exchange."""


class PeerWindow:
    """Peer-memory exchange of the gene shards of ONE node (``pdq_peer_*`` in ``include/pydeseq2_b200.h``): every rank owns a
    device window that all peers map through CUDA IPC; :meth:`push` is one kernel that stores this rank's segments into every
    rank's window over NVLink and returns -- on the stream -- when all ranks' segments have arrived here.  Construction and
    :meth:`close` are collective over ``comm`` (the IPC handles travel through its host-staged all-gather); either every rank
    gets a window or every rank raises :class:`PeerUnavailable`."""

    def __init__(self, comm: NcclComm, data_bytes: int):
        from . import _lib

        ctx = self.ctx = comm.ctx
        self.comm = comm
        self.data_bytes = int(data_bytes)
        self._group = None
        w, d = C.c_void_p(), C.c_void_p()
        hbuf = C.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        err = None
        try:
            ctx.check(ctx.lib.pdq_peer_window_alloc(ctx.h, self.data_bytes, C.byref(w), C.byref(d), hbuf))
        except RuntimeError as e:  # keep the collective sequence alive: the other ranks are waiting in the all-gather
            err = e
        self.window, self.data = w.value, d.value
        # one double per handle byte: the bits of an IPC handle are not a float64 anybody should normalise
        mine = np.frombuffer(hbuf.raw, dtype=np.uint8).astype(np.float64)
        handles = comm._gather_equal(mine).astype(np.uint8).tobytes()
        if err is None:
            g = C.c_void_p()
            hb = C.create_string_buffer(handles, len(handles))
            try:
                ctx.check(ctx.lib.pdq_peer_group_open(ctx.h, C.c_void_p(self.window), comm.world, comm.rank, hb, C.byref(g)))
                self._group = g
            except RuntimeError as e:
                err = e
        ok = comm._gather_equal(np.array([0.0 if err is not None else 1.0]))
        if not bool(np.all(ok == 1.0)):
            self.close()
            raise PeerUnavailable(str(err) if err is not None else "a peer rank could not map the windows")

    def push(self, segments, count: int):
        """``segments`` = [(send_ptr, byte offset of the gathered vector in the window's payload), ...] (at most 4), ``count``
        doubles per rank each; segment of rank r lands at ``offset + r * count * 8`` in every window.  No segments: barrier."""
        k = len(segments)
        send = (C.c_void_p * max(k, 1))(*[s[0] for s in segments])
        offs = (C.c_uint64 * max(k, 1))(*[int(s[1]) for s in segments])
        self.ctx.check(self.ctx.lib.pdq_peer_push_dev(self.ctx.h, self._group, k, send, offs, int(count)))

    def check(self):
        """After a stream synchronisation: raise when a push gave up waiting for a peer."""
        st = C.c_uint64(0)
        self.ctx.check(self.ctx.lib.pdq_peer_status(self.ctx.h, self._group, C.byref(st)))
        if st.value:
            raise RuntimeError(f"peer exchange timed out waiting for rank {st.value - 1}")

    def close(self):
        """Collective: unmap the peers' windows, wait until every rank has done so, free the own window."""
        if self._group is not None:
            self.ctx.check(self.ctx.lib.pdq_peer_group_close(self.ctx.h, self._group))
            self._group = None
        self.comm.barrier()
        if self.window:
            self.ctx.check(self.ctx.lib.pdq_peer_window_free(self.ctx.h, C.c_void_p(self.window)))
            self.window = self.data = None
