"""Peer-memory BatchNorm statistics exchange (csrc/peer_exchange.cu): host-side set-up.

One mailbox per process, allocated by the library with cudaMalloc, exported as a CUDA-IPC handle, gathered over
the process group and mapped into every peer.  ``allreduce_bn`` then replaces, per BN layer, the
``dist.all_reduce`` of the 2C fp64 sums AND (forward) the finalize launch with one single-CTA kernel whose traffic
goes straight over NVLink.  Ranks must call it in the same order (they do: same model, same step)."""
import ctypes

import torch

from .. import _lib
from .._lib import call

MAX_WORLD = 8
MAX_VALUES = 4096


class PeerExchange:
    def __init__(self, group=None):
        import torch.distributed as dist
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        if self.world > MAX_WORLD:
            raise ValueError('peer exchange supports at most %d ranks' % MAX_WORLD)
        self.seq = 0
        own = ctypes.c_void_p()
        call('pxl_peer_alloc', ctypes.byref(own))
        self._own = own
        handle = (ctypes.c_ubyte * 64)()
        call('pxl_peer_export', own, handle)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=self.group)
        self._opened = []
        ptrs = (ctypes.c_void_p * MAX_WORLD)()
        for r in range(self.world):
            if r == self.rank:
                ptrs[r] = own.value
            else:
                p = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(handles[r])
                call('pxl_peer_open', buf, ctypes.byref(p))
                self._opened.append(p)
                ptrs[r] = p.value
        self._ptrs = ptrs
        dist.barrier(group=self.group)          # every mailbox is mapped everywhere before the first exchange

    def allreduce_bn(self, sums, finalize=None, param_grads=None):
        """In-place sum of ``sums`` (fp64, <= 4096 values) over the ranks.  finalize = (count, C, gamma, beta,
        running_mean, running_var, momentum, eps, clamp, mean, invstd, scale, shift) also finishes the layer;
        param_grads = (dgamma_acc, dbeta_acc) receive += the local sums before the exchange (backward)."""
        n = sums.numel()
        if n > MAX_VALUES or sums.dtype != torch.float64 or not sums.is_cuda or not sums.is_contiguous():
            raise ValueError('peer exchange takes a contiguous CUDA fp64 vector of at most %d values' % MAX_VALUES)
        self.seq += 1
        from .. import ops
        P = ops._p                              # plain ints / None: ~310 exchanges per step, keep the host side cheap
        stream = ops._stream()
        if finalize is None:
            z = None
            dg, db = (P(param_grads[0]), P(param_grads[1])) if param_grads is not None else (z, z)
            call('pxl_peer_allreduce_bn', P(sums), n, self._ptrs, self.rank, self.world, self.seq, 0.0, 0,
                 z, z, z, z, 0.0, 0.0, 0, z, z, z, z, dg, db, stream)
        else:
            count, C, gamma, beta, rm, rv, momentum, eps, clamp, mean, invstd, scale, shift = finalize
            call('pxl_peer_allreduce_bn', P(sums), n, self._ptrs, self.rank, self.world, self.seq, float(count), int(C),
                 P(gamma), P(beta), P(rm), P(rv), float(momentum), float(eps), int(clamp), P(mean), P(invstd), P(scale),
                 P(shift), None, None, stream)
        return sums

    def status(self):
        return int(_lib.load().pxl_peer_status())

    def close(self):
        torch.cuda.synchronize()
        for p in self._opened:
            _lib.load().pxl_peer_close(p)
        self._opened = []
        if self._own is not None:
            _lib.load().pxl_peer_free(self._own)
            self._own = None
