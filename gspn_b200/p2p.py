"""Peer-memory all-reduce of small vectors over NVLink (csrc/p2p.cu) for the data-parallel training form: the whole-batch
batch-norm statistics -- 52 dependent collectives of a few hundred bytes per train step, which through NCCL + c10d are pure latency.

  grp = p2p.PeerGroup(device)            # once per process, after torch.distributed is initialised (any backend: it only carries
                                         # the 64-byte CUDA IPC handles); one process per GPU, all on one node
  grp.allreduce_(buf)                    # buf: contiguous float64 CUDA tensor (<= max_doubles), summed over ranks in place, on the
                                         # current stream -- no NCCL, no host synchronisation, CUDA-graph capturable
  train.use_peer_moments(grp)            # SyncBN through it

Every rank must issue the same sequence of calls (like any collective).  Sums are taken in rank order, so every rank gets
bit-identical results."""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check


class PeerGroup:
    def __init__(self, device, group=None, max_doubles=4096):
        if not dist.is_initialized():
            raise RuntimeError("PeerGroup needs torch.distributed (any backend) to exchange the CUDA IPC handles")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.max_doubles = int(max_doubles)
        self.device = torch.device(device)
        L = _lib.lib()
        with torch.cuda.device(self.device):
            box = ctypes.c_void_p()
            handle = (ctypes.c_ubyte * 64)()
            check(L.gspn_p2p_mailbox_create(self.world, self.max_doubles, ctypes.byref(box), ctypes.cast(handle, ctypes.c_void_p)), "p2p_mailbox_create")
            self._mine = box
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self._peers = []
            ptrs = (ctypes.c_void_p * self.world)()
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs[r] = box.value
                    self._peers.append(None)
                    continue
                hb = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                pp = ctypes.c_void_p()
                check(L.gspn_p2p_mailbox_open(ctypes.cast(hb, ctypes.c_void_p), ctypes.byref(pp)), "p2p_mailbox_open (CUDA IPC: all ranks on one node?)")
                ptrs[r] = pp.value
                self._peers.append(pp)
            self._ptrs = ptrs
        dist.barrier(group=group)  # every mailbox is mapped everywhere before the first store lands

    def allreduce_(self, buf):
        if not (isinstance(buf, torch.Tensor) and buf.is_cuda and buf.dtype == torch.float64 and buf.is_contiguous()):
            raise TypeError("PeerGroup.allreduce_ takes a contiguous float64 CUDA tensor")
        if buf.numel() > self.max_doubles:
            raise ValueError("vector of %d doubles exceeds the mailbox slot (%d)" % (buf.numel(), self.max_doubles))
        check(_lib.lib().gspn_p2p_allreduce_f64(self.rank, self.world, self.max_doubles, ctypes.cast(self._ptrs, ctypes.c_void_p), buf.numel(),
                                                buf.data_ptr(), torch.cuda.current_stream().cuda_stream), "p2p_allreduce")
        return buf

    def close(self):
        L = _lib.lib()
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            dist.barrier(group=self.group)  # nobody unmaps while a peer may still store
        for pp in self._peers:
            if pp is not None:
                L.gspn_p2p_mailbox_close(pp)
        L.gspn_p2p_mailbox_destroy(self._mine)
        self._peers, self._mine = [], None
