"""Peer-mapped gather buffers for the fused mask all-gather (include/sqmp_b200.h: sqmp_check_states_gather_dev).

One process per GPU (torch.distributed): every rank allocates [flags | buffer 0 | buffer 1 | ...] with cudaMalloc, the
CUDA IPC handles are exchanged through the process group, and every rank maps every other rank's allocation
(cudaIpcOpenMemHandle enables peer access over NVLink lazily).  The validity kernel then stores a rank's mask words
straight into all `world` gather buffers; no collective is launched in the step.  Plumbing only — no arithmetic.
"""
from __future__ import annotations

import ctypes as C

from . import api

FLAG_BYTES = 256


def _ck(ret):
    err = ret[0] if isinstance(ret, tuple) else ret
    if int(err) != 0:
        raise RuntimeError("CUDA runtime error %s" % (err,))
    if isinstance(ret, tuple):
        return ret[1] if len(ret) == 2 else ret[1:]
    return None


class PeerGather:
    def __init__(self, rank: int, world: int, words_per_rank: int, n_buffers: int = 2, group=None):
        import torch.distributed as dist
        from cuda.bindings import runtime as cudart
        if world > api.MAX_PEERS:
            raise ValueError("at most %d ranks per node" % api.MAX_PEERS)
        self.cudart, self.rank, self.world, self.words, self.n_buffers = cudart, rank, world, words_per_rank, n_buffers
        self.buf_bytes = world * words_per_rank * 4
        total = FLAG_BYTES + n_buffers * self.buf_bytes
        self.base = int(_ck(cudart.cudaMalloc(total)))
        _ck(cudart.cudaMemset(self.base, 0, total))
        _ck(cudart.cudaDeviceSynchronize())
        handle = _ck(cudart.cudaIpcGetMemHandle(self.base))
        blobs = [None] * world
        dist.all_gather_object(blobs, bytes(handle.reserved), group=group)
        self.peer_base, self._opened = [], []
        for p in range(world):
            if p == rank:
                self.peer_base.append(self.base)
                continue
            h = cudart.cudaIpcMemHandle_t()
            h.reserved = blobs[p]
            ptr = int(_ck(cudart.cudaIpcOpenMemHandle(h, cudart.cudaIpcMemLazyEnablePeerAccess)))
            self.peer_base.append(ptr)
            self._opened.append(ptr)
        dist.barrier(group=group)
        self.flags_ptr = self.base                      # this rank's flag array (uint32[world])
        self.structs = []
        for b in range(n_buffers):
            g = api.PeerGatherC()
            g.n_peers, g.rank, g.word_offset = world, rank, rank * words_per_rank
            for p in range(world):
                g.mask[p] = self.peer_base[p] + FLAG_BYTES + b * self.buf_bytes
                g.flag[p] = self.peer_base[p]
            self.structs.append(g)

    def buffer_ptr(self, b: int) -> int:
        """Device address of this rank's gather buffer b (world * words uint32)."""
        return self.base + FLAG_BYTES + b * self.buf_bytes

    def copy_buffer_to(self, b: int, tensor):
        """Device-to-device copy of gather buffer b into a torch int32 tensor of world * words elements."""
        _ck(self.cudart.cudaMemcpy(tensor.data_ptr(), self.buffer_ptr(b), self.buf_bytes,
                                   self.cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice))

    def close(self):
        for ptr in self._opened:
            self.cudart.cudaIpcCloseMemHandle(ptr)
        self._opened = []
        if self.base:
            self.cudart.cudaFree(self.base)
            self.base = 0
