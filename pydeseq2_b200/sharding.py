"""Gene shards across GPUs: one process per GPU, contiguous gene blocks, the design matrix and
size factors replicated.  Every hot-path method is independent per gene; the ONLY cross-gene dependency on
the path is the dispersion trend + prior (``dds.py:799-884``), which need the genewise dispersions and
normalised means of all genes -- one small all-gather of two float64 vectors.

``NcclComm`` runs that all-gather with NCCL on device buffers through the C ABI (``pdq_allgather_f64_dev``);
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
        self._hs[:] = send
        self.ctx.h2d(self._send, self._hs)
        self.ctx.check(self.ctx.lib.pdq_allgather_f64_dev(self.ctx.h, self._c_dptr(self._send), self._c_dptr(self._recv), n))
        self.ctx.d2h(self._hr, self._recv)
        self.ctx.sync()
        return self._hr.copy()

    def allgather_dev(self, pairs, count):
        """Device-to-device all-gather of several equal-length vectors as ONE NCCL group (a single fused launch on the context's
        stream, capturable into a CUDA graph).  ``pairs`` = [(send_ptr, recv_ptr), ...], every send buffer holds ``count``
        doubles -- the largest shard's length; shorter shards keep NaN behind their last gene (written once at upload time,
        never per step) -- and every recv buffer ``world * count``."""
        k = len(pairs)
        send = (C.c_void_p * k)(*[p[0] for p in pairs])
        recv = (C.c_void_p * k)(*[p[1] for p in pairs])
        self.ctx.check(self.ctx.lib.pdq_allgather_multi_f64_dev(self.ctx.h, k, send, recv, int(count)))

    def close(self):
        self.ctx.check(self.ctx.lib.pdq_comm_destroy(self.ctx.h))
