"""Multi-GPU XC build: grid leaves sharded over ranks, density matrix replicated, partial K / Exc / N summed.

Reference analogue: the thread-local accumulators of DFT::compute_K_dft and their serial reduction
(include/occ/dft/dft.h:221-232,375-380) — here one process per GPU (torch.distributed, NCCL over NVLink) instead of
TBB threads, and ONE all-reduce of the packed buffer [K | exc, n_alpha, n_beta] instead of nthreads matrix adds.
The partition itself (contiguous Morton runs of equal predicted cost n*nb^2) lives in the native library
(occ_xc_create, options.rank / world_size).
"""
from __future__ import annotations

import os
import time
from multiprocessing import resource_tracker, shared_memory

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


def slice_len(nspin: int, nbf: int, world: int) -> int:
    """length of one rank's slice of the packed buffer padded to a multiple of world (reduce-scatter granularity)"""
    return -(-pack_len(nspin, nbf) // world)


class SharedHostArray:
    """A float64 (or int64) array in POSIX shared memory, mapped by every rank of the node and registered with CUDA as
    pinned memory in each process: the HOST side of the multi-GPU build.  The reference's SCF is one host process that
    owns D and K (include/occ/qm/scf_impl.h:504-572); here that process is rank 0 and the other ranks map the same
    pages, so each GPU copies D over its OWN PCIe link and writes its slice of K back the same way — no GPU broadcast."""

    def __init__(self, name: str, nelem: int, create: bool, dtype=np.float64, register: bool = True):
        nbytes = max(1, nelem) * np.dtype(dtype).itemsize
        self.shm = shared_memory.SharedMemory(name=name, create=create, size=nbytes)
        if not create:   # the creating rank unlinks; keep this process's resource tracker out of it (Python < 3.13)
            try:
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.owner = create
        self.array = np.ndarray((nelem,), dtype=dtype, buffer=self.shm.buf)
        self.tensor = torch.from_numpy(self.array)
        self.registered = False
        if register and torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self.tensor.data_ptr(), nbytes, 1)   # cudaHostRegisterPortable
            if int(rc) != 0:
                raise RuntimeError(f"cudaHostRegister failed with code {int(rc)}")
            self.registered = True

    def close(self):
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self.registered = False
        self.tensor = None
        self.array = None
        try:
            self.shm.close()
            if self.owner:
                self.shm.unlink()
        except Exception:
            pass


class DistributedDFT:
    """One DFT handle per rank on its own GPU; build() leaves the SUMMED packed result on every rank's device."""

    def __init__(self, method: str, basis, grid: GridHierarchy, spinorbital_kind: str = RESTRICTED, group=None,
                 device: int | None = None, shared_host: bool = True, shared_tag: str = "0", **kw):
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
        # occupied orbitals for the factorised (MO-form) density: nspin blocks nbf x nocc_s, packed
        self.nocc = self.dft.nocc
        self.nC = self.nbf * (self.nocc[0] + (self.nocc[1] if self.nspin == 2 else 0)) if self.nocc else 0
        self.d_C = torch.empty(max(self.nC, 1), dtype=torch.float64, device=dev)
        self.host_C = None
        self.d_out = torch.zeros(pack_len(self.nspin, self.nbf), dtype=torch.float64, device=dev)
        # a dedicated non-default stream: the C ABI reads a NULL stream as "the handle's own stream", so the legacy
        # default stream (value 0) can never be passed through it; copies, build and all-reduce all run here
        self.stream = torch.cuda.Stream(device=dev)
        # ---- host side of the N > 1 end-to-end path: D, K and a few control words in node-shared pinned memory ----
        self.host_D = self.host_K = self._ctl = None
        self._epoch = 0
        if self.world > 1 and shared_host:
            n = self.nspin * self.nbf * self.nbf
            self._slice = slice_len(self.nspin, self.nbf, self.world)
            # /dev/shm must hold both matrices (a container may cap it at 64 MB; writing past the cap is a SIGBUS, not an
            # exception): rank 0 checks, every rank takes rank 0's verdict; without room the broadcast path stays
            need = 8 * (n + self._slice * self.world + self.nC) + (1 << 20)
            try:
                st = os.statvfs("/dev/shm")
                ok = st.f_bavail * st.f_frsize >= need
            except OSError:
                ok = False
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.broadcast(flag, src=0, group=group)
            shared_host = bool(int(flag.item()))
        if self.world > 1 and shared_host:
            tag = f"occb200_{os.environ.get('MASTER_PORT', '0')}_{shared_tag}"
            create = self.rank == 0
            if not create:
                dist.barrier(group)          # rank 0 has created the segments
            self.host_D = SharedHostArray(tag + "_D", n, create)
            self.host_K = SharedHostArray(tag + "_K", self._slice * self.world, create)
            if self.nC:
                self.host_C = SharedHostArray(tag + "_C", self.nC, create)
            self._ctl = SharedHostArray(tag + "_c", 2 + self.world, create, dtype=np.int64, register=False)
            if create:
                self._ctl.array[:] = 0
                dist.barrier(group)
            dist.barrier(group)              # everyone has attached
            self.d_red = torch.zeros(self._slice * self.world, dtype=torch.float64, device=dev)
            self.d_slice = torch.zeros(self._slice, dtype=torch.float64, device=dev)

    def _launch(self, out: torch.Tensor, mo: bool, d2_resident: bool = True):
        n = self.nspin * self.nbf * self.nbf
        if mo:   # d2_resident False: the library forms D2 = C C^T itself (only the orbitals were uploaded)
            self.dft.build_mo_device(self.d_D.data_ptr() if d2_resident else 0, self.d_C.data_ptr(), out.data_ptr(),
                                     out.data_ptr() + 8 * n, self.stream.cuda_stream)
        else:
            self.dft.build_device(self.d_D.data_ptr(), out.data_ptr(), out.data_ptr() + 8 * n, self.stream.cuda_stream)

    def build_device(self, mo: bool = False, d2_resident: bool = True) -> torch.Tensor:
        """D already in self.d_D (and, with mo=True, the orbitals in self.d_C), column-major. Enqueues build +
        all-reduce on self.stream; returns the packed device buffer [K | exc, n_alpha, n_beta] (valid once self.stream
        has been synchronised / waited on)."""
        self._launch(self.d_out, mo, d2_resident)
        if self.world > 1:
            with torch.cuda.stream(self.stream):
                allreduce_partial(self.d_out, self.group)
        return self.d_out

    def capture_device_build(self, mo: bool = False, d2_resident: bool = True) -> "torch.cuda.CUDAGraph":
        """CUDA graph of one device-resident build (one GPU).  occ_xc_build(_mo)_device is a fixed sequence of kernel
        launches and memsets on the caller's stream, with no host synchronisation and every buffer owned by the handle or
        by this object, so it can be captured once and replayed: for small systems (water: 9 launches of a few
        microseconds each) the replay removes the launch gaps.  `with torch.cuda.stream(self.stream): g.replay()`
        leaves the packed result in self.d_out, computed from whatever self.d_D / self.d_C hold at replay time."""
        if self.world > 1:
            raise RuntimeError("capture_device_build: one GPU only (the all-reduce is not captured)")
        self._launch(self.d_out, mo, d2_resident)        # warm-up outside the capture (shared-memory opt-ins, lazy module loads)
        self.stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self.stream, capture_error_mode="thread_local"):
            self._launch(self.d_out, mo, d2_resident)
        return g

    def build(self, D2_host: torch.Tensor | None, K_host: torch.Tensor | None, C_host: torch.Tensor | None = None,
              mo: bool = False):
        """Host-to-host build as the reference's single-process SCF driver would see it: rank 0 owns the host
        matrices (pinned CPU tensors, column-major).  Rank 0 copies the inputs to its GPU — with mo=True ONLY the
        occupied orbitals, D2 = C C^T is formed on each device — they are broadcast to the other GPUs over NVLink (NCCL),
        every rank builds its shard, the packed result is all-reduced and rank 0 copies K back.
        Other ranks pass None (or buffers that are ignored). Returns (exc, n_alpha, n_beta) on every rank."""
        n = self.nspin * self.nbf * self.nbf
        with torch.cuda.stream(self.stream):
            if self.rank == 0:
                if mo:
                    self.d_C.copy_(C_host.view(-1), non_blocking=True)
                else:
                    self.d_D.copy_(D2_host.view(-1), non_blocking=True)
            if self.world > 1:
                dist.broadcast(self.d_C if mo else self.d_D, src=0, group=self.group)
            out = self.build_device(mo, d2_resident=False)
            if self.rank == 0 and K_host is not None:
                K_host.view(-1)[:n].copy_(out[:n], non_blocking=True)
            self._sc_host.copy_(out[n:n + 3], non_blocking=True)
        self.stream.synchronize()
        return float(self._sc_host[0]), float(self._sc_host[1]), float(self._sc_host[2])

    def _spin_until(self, index: int, value: int, timeout: float = 120.0):
        a = self._ctl.array
        t0 = time.perf_counter()
        while a[index] < value:
            if time.perf_counter() - t0 > timeout:
                raise RuntimeError("DistributedDFT: a rank did not reach the shared-memory rendezvous")

    def build_shared(self, mo: bool = False):
        """Host-to-host build for N > 1 with the host matrices in node-shared pinned memory (self.host_D / self.host_K,
        column-major, owned by rank 0's caller — the SCF process).  Every rank: H2D of D (mo: of C_occ only) over its own PCIe link ->
        sharded build -> NCCL reduce-scatter of the packed [K | exc, n_alpha, n_beta] -> D2H of its 1/N slice into
        host_K.  Rank 0 publishes "D is ready" and collects "slice written" through control words in the same shared
        segment (no collective on the host path).  Returns (exc, n_alpha, n_beta) on every rank."""
        n = self.nspin * self.nbf * self.nbf
        self._epoch += 1
        ep = self._epoch
        if self.rank == 0:
            self._ctl.array[0] = ep                      # D in host_D is complete
        else:
            self._spin_until(0, ep)
        with torch.cuda.stream(self.stream):
            if mo:   # only the orbitals cross PCIe; D2 = C C^T is formed on the device
                self.d_C.copy_(self.host_C.tensor, non_blocking=True)
            else:
                self.d_D.copy_(self.host_D.tensor, non_blocking=True)
            self._launch(self.d_red, mo, d2_resident=False)
            dist.reduce_scatter_tensor(self.d_slice, self.d_red, op=dist.ReduceOp.SUM, group=self.group)
            lo = self.rank * self._slice
            self.host_K.tensor[lo:lo + self._slice].copy_(self.d_slice, non_blocking=True)
        self.stream.synchronize()
        self._ctl.array[2 + self.rank] = ep              # this rank's slice of K is in host memory
        for r in range(self.world):
            self._spin_until(2 + r, ep)
        sc = self.host_K.array[n:n + 3]
        return float(sc[0]), float(sc[1]), float(sc[2])

    def close(self):
        for b in (self.host_D, self.host_K, self.host_C, self._ctl):
            if b is not None:
                b.close()
        self.host_D = self.host_K = self.host_C = self._ctl = None
        self.dft.close()
