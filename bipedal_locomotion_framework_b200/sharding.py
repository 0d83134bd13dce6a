"""Multi-GPU host logic of the sampling-MPC path.

One process per GPU.  Rollouts (samples) are block-partitioned across ranks so that every
evaluation of a rollout lives on one GPU: the contact evaluation itself needs NO collective.  The
only exchange is the 16-byte (cost, index) pair each rank's arg-min produces: an all-gather over
NCCL (NVLink 5 / NVSwitch) on the GPU box, gloo in the CPU tests, followed by a lowest-index
tie-break selection (blf_ccm_argmin_pairs on the device).
"""
from __future__ import annotations

import struct


def shard_rollouts(n_rollouts: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition: returns (first_rollout, count) owned by `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    first = n_rollouts * rank // world
    last = n_rollouts * (rank + 1) // world
    return first, last - first


def pack_pair(cost: float, index: int) -> tuple[int, int]:
    """(cost, index) as the two int64 words the device writes (double bits, index)."""
    return struct.unpack("<q", struct.pack("<d", cost))[0], int(index)


def unpack_pair(word0: int, word1: int) -> tuple[float, int]:
    return struct.unpack("<d", struct.pack("<q", int(word0)))[0], int(word1)


def all_gather_pairs(best, world: int, dist=None):
    """All-gather every rank's (2,) int64 `best` tensor into a (world, 2) tensor on the same
    device (16 bytes per rank; pure latency on NVSwitch)."""
    import torch
    if world == 1:
        return best.view(1, 2)
    dist = dist or torch.distributed
    gathered = torch.empty((world, 2), dtype=torch.int64, device=best.device)
    dist.all_gather_into_tensor(gathered.view(-1), best)
    return gathered


class PeerArgmin:
    """Arg-min exchange over NVLink peer memory (blf_ccm_p2p_mailbox_* / blf_ccm_argmin_exchange_p2p):
    one single-warp kernel per rank instead of an NCCL all-gather.  `dist` (torch.distributed, any
    backend) is used ONCE, to exchange the 64-byte CUDA IPC handles."""

    def __init__(self, batch, world: int, rank: int, dist=None):
        import ctypes as C

        import torch

        from . import _capi
        self._capi, self._C, self._torch = _capi, C, torch
        self._batch = batch
        self.world = world
        handle = (C.c_ubyte * 64)()
        _capi.check(_capi.lib().blf_ccm_p2p_mailbox_create(batch.handle.ptr, world, rank, handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8)
        dist = dist or torch.distributed
        if dist.get_backend() == "nccl":
            gathered = torch.empty((world, 64), dtype=torch.uint8, device=batch.device)
            dist.all_gather_into_tensor(gathered.view(-1), mine.to(batch.device))
            gathered = gathered.cpu()
        else:
            gathered = torch.empty((world, 64), dtype=torch.uint8)
            dist.all_gather_into_tensor(gathered.view(-1), mine)
        blob = (C.c_ubyte * (64 * world))(*gathered.view(-1).tolist())
        _capi.check(_capi.lib().blf_ccm_p2p_mailbox_connect(batch.handle.ptr, blob))
        dist.barrier()   # every mailbox is mapped everywhere before the first exchange
        self._out = torch.empty(2, dtype=torch.int64, device=batch.device)
        # self-check: one exchange of (cost = rank, index = rank) must yield (0.0, 0) everywhere
        probe = torch.tensor(list(pack_pair(float(rank), rank)), dtype=torch.int64, device=batch.device)
        got = unpack_pair(*self.exchange(probe).cpu().tolist())
        if got != (0.0, 0):
            raise RuntimeError(f"peer-memory exchange self-check failed on rank {rank}: {got}")

    def exchange(self, best, out=None):
        """best: this rank's (2,) int64 pair tensor -> (2,) global best (same on every rank)."""
        out = self._out if out is None else out
        st = self._torch.cuda.current_stream(self._batch.device).cuda_stream
        rc = self._capi.lib().blf_ccm_argmin_exchange_p2p(self._batch.handle.ptr, best.data_ptr(),
                                                          out.data_ptr(), st)
        if rc:
            self._capi.check(rc)
        return out

    def fuse_into_rollouts(self, on: bool = True):
        """Every later rollout call of the batch also runs the exchange inside its own reduction
        kernel (blf_ccm_rollout_set_exchange); the global best lands in `self.global_best`."""
        self._capi.check(self._capi.lib().blf_ccm_rollout_set_exchange(
            self._batch.handle.ptr, self._out.data_ptr() if on else None))
        return self._out

    @property
    def global_best(self):
        return self._out

    def close(self):
        self._capi.check(self._capi.lib().blf_ccm_p2p_mailbox_destroy(self._batch.handle.ptr))
