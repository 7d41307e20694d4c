"""Multi-GPU XC build: grid leaves sharded over ranks, density matrix replicated, partial K / Exc / N summed.

Reference analogue: the thread-local accumulators of DFT::compute_K_dft and their serial reduction
(include/occ/dft/dft.h:221-232,375-380) — here one process per GPU (torch.distributed, NCCL over NVLink) instead of
TBB threads, and ONE all-reduce of the packed buffer [K | exc, n_alpha, n_beta] instead of nthreads matrix adds.
The partition itself (contiguous Morton runs of equal predicted cost n*nb^2) lives in the native library
(occ_xc_create, options.rank / world_size).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .dft import DFT, RESTRICTED, UNRESTRICTED, GridHierarchy


def pack_len(nspin: int, nbf: int) -> int:
    """K (nspin*nbf x nbf, column-major) followed by exc, n_alpha, n_beta and one pad double"""
    return nspin * nbf * nbf + 4


def allreduce_partial(buf: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the packed partial results over ranks (NCCL on device tensors, gloo on CPU tensors in the tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def unpack(buf: torch.Tensor, nspin: int, nbf: int):
    n = nspin * nbf * nbf
    K = buf[:n].view(nbf, nspin * nbf).t()  # column-major (nspin*nbf, nbf) seen as a strided view
    return K, buf[n:n + 3]


class DistributedDFT:
    """One DFT handle per rank on its own GPU; build() leaves the SUMMED packed result on every rank's device."""

    def __init__(self, method: str, basis, grid: GridHierarchy, spinorbital_kind: str = RESTRICTED, group=None,
                 device: int | None = None, **kw):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = torch.cuda.current_device() if device is None else device
        self.dft = DFT(method, basis, grid, spinorbital_kind, device=self.device, rank=self.rank,
                       world_size=self.world, **kw)
        self.nspin, self.nbf = self.dft.nspin, self.dft.nbf
        self._sc_host = torch.zeros(3, dtype=torch.float64).pin_memory()
        dev = torch.device("cuda", self.device)
        self.d_D = torch.empty(self.nspin * self.nbf * self.nbf, dtype=torch.float64, device=dev)
        self.d_out = torch.zeros(pack_len(self.nspin, self.nbf), dtype=torch.float64, device=dev)
        # a dedicated non-default stream: the C ABI reads a NULL stream as "the handle's own stream", so the legacy
        # default stream (value 0) can never be passed through it; copies, build and all-reduce all run here
        self.stream = torch.cuda.Stream(device=dev)

    def build_device(self) -> torch.Tensor:
        """D already in self.d_D (column-major). Enqueues build + all-reduce on self.stream; returns the packed
        device buffer [K | exc, n_alpha, n_beta] (valid once self.stream has been synchronised / waited on)."""
        n = self.nspin * self.nbf * self.nbf
        self.dft.build_device(self.d_D.data_ptr(), self.d_out.data_ptr(), self.d_out.data_ptr() + 8 * n,
                              self.stream.cuda_stream)
        if self.world > 1:
            with torch.cuda.stream(self.stream):
                allreduce_partial(self.d_out, self.group)
        return self.d_out

    def build(self, D2_host: torch.Tensor | None, K_host: torch.Tensor | None):
        """Host-to-host build as the reference's single-process SCF driver would see it: rank 0 owns the host
        matrices (pinned CPU tensors, column-major).  Rank 0 copies D to its GPU, D is broadcast to the other GPUs over
        NVLink (NCCL), every rank builds its shard, the packed result is all-reduced and rank 0 copies K back.
        Other ranks pass None (or buffers that are ignored). Returns (exc, n_alpha, n_beta) on every rank."""
        n = self.nspin * self.nbf * self.nbf
        with torch.cuda.stream(self.stream):
            if self.rank == 0:
                self.d_D.copy_(D2_host.view(-1), non_blocking=True)
            if self.world > 1:
                dist.broadcast(self.d_D, src=0, group=self.group)
            out = self.build_device()
            if self.rank == 0 and K_host is not None:
                K_host.view(-1)[:n].copy_(out[:n], non_blocking=True)
            self._sc_host.copy_(out[n:n + 3], non_blocking=True)
        self.stream.synchronize()
        return float(self._sc_host[0]), float(self._sc_host[1]), float(self._sc_host[2])

    def close(self):
        self.dft.close()
