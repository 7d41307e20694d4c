#!/usr/bin/env python
"""bench.py — the Vxc+Exc build per SCF iteration (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

A step = one DFT::compute_K_dft-equivalent build (include/occ/dft/dft.h:190-388) on a fixed synthetic density matrix.
  value  : grid-points/s with D resident in HBM at the start and K left in HBM (device events, max over ranks,
           the N>1 number includes the NCCL all-reduce of the packed [K | Exc, N_a, N_b] buffer)
  e2e    : the same through the host-facing call: pinned host D -> H2D -> build -> (all-reduce) -> D2H K + scalars
  roofline: dominant kernel (the FP64 DMMA contractions) against the FP64 tensor peak measured on this pool's B200
           (profiles/r01_fp64_peaks_b200.jsonl: 36.9 TFLOP/s DMMA.8x8x4; MEASURED_PEAKS.json carries no FP64 entry)
  cpu_baseline / --impl reference: the CPU restatement of the reference algorithm (oracle/, OpenMP over leaves +
           OpenBLAS) on the box's host cores, on a bounded sample of the same workload (every s-th Morton leaf).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP64_TENSOR_PEAK_TFLOPS = 36.9  # profiles/r01_fp64_peaks_b200.jsonl (dmma_m8n8k4, measured on this pool's B200)

WORKLOADS = {
    # name: (molecule, basis, method, nspin, grid quality or None for the reference default)
    "water_b3lyp_svp": ("water", "def2-svp", "b3lyp", 1, None),
    "benzene_pbe_tzvp": ("benzene", "def2-tzvp", "pbe", 1, None),
    "h2o32_b3lyp_svp_uks": ("h2o_32", "def2-svp", "b3lyp", 2, None),
    "valinomycin_b3lyp_tzvp": ("valinomycin", "def2-tzvp", "b3lyp", 1, None),
    "taxol_r2scan_tzvp": ("taxol", "def2-tzvp", "r2scan", 1, None),
}
DEFAULT_WORKLOAD = "valinomycin_b3lyp_tzvp"
KERNEL_NAMES = {"density": "k_density", "vxc": "k_vxc_resident"}


def make_inputs(workload):
    from occ_b200 import aobasis, molecules
    mol, bas, method, nspin, quality = WORKLOADS[workload]
    Z, pos = molecules.MOLECULES[mol]()
    basis = aobasis.AOBasis.load(Z, pos, bas)
    D2 = molecules.synthetic_density(basis, nspin)
    return basis, method, nspin, quality, D2


def make_orbitals(basis, nspin):
    """the orbitals behind the synthetic density (D2_s = C_s C_s^T), packed like occ_xc_build_mo wants them"""
    from occ_b200 import molecules
    Cs = molecules.synthetic_orbitals(basis, nspin)
    return Cs[0].shape[1], np.concatenate([c.ravel(order="F") for c in Cs])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_build_sample(basis, method, nspin, D2, hierarchy, target_seconds=15.0, nthreads=None, repeats=1):
    """Time the CPU restatement of the reference algorithm on every `stride`-th Morton leaf; returns points/s etc."""
    from occ_b200 import dft
    from oracle import oracle as O
    ob = O.Basis(basis.as_arrays())
    names = [n for i in dft.get_dft_method(method) for n, v in O.FUNC_IDS.items() if v == i]
    leaves = hierarchy.leaves
    nthreads = nthreads or os.cpu_count() or 1
    # probe with a thin sample to pick the stride
    stride = max(1, len(leaves) // 64)
    probe = leaves[::stride]
    t0 = time.perf_counter()
    O.xc_build(ob, hierarchy.sorted_points, hierarchy.sorted_weights, probe, names, nspin, D2, nthreads=nthreads)
    t_probe = time.perf_counter() - t0
    full_est = t_probe * len(leaves) / max(1, len(probe))
    stride = max(1, int(np.ceil(full_est / target_seconds)))
    sample = leaves[::stride]
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = O.xc_build(ob, hierarchy.sorted_points, hierarchy.sorted_weights, sample, names, nspin, D2, nthreads=nthreads)
        times.append(time.perf_counter() - t0)
    npts = int(sample["count"].sum())
    return {"seconds": times, "npts": npts, "stride": stride, "nleaves": len(sample), "cores": nthreads,
            "flops": (16.0 if O.method_deriv(names) >= 2 else 4.0) * nspin * r["stats"]["sum_n_nb2"]}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle restatement; the reference binary cannot be built
    here — Eigen/libxc/oneTBB are network-fetched) on the host cores, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from occ_b200 import grid
    basis, method, nspin, quality, D2 = make_inputs(args.workload)
    try:
        g = grid.MolecularGrid(basis, grid.GridSettings.from_grid_quality(quality) if quality else None)
        h = g.hierarchy()
    except RuntimeError:
        # no usable GPU for the (GPU-accelerated) grid producer: fall back to the oracle's own grid code
        from occ_b200 import dft
        from oracle import oracle as O
        ob = O.Basis(basis.as_arrays())
        pts, w, _ = O.molecular_grid(basis.atomic_numbers, basis.positions, ob)
        spts, sw, perm, leaves = O.hierarchy(pts, w)
        h = dft.GridHierarchy(spts, sw, leaves)
    cores = os.cpu_count() or 1
    budget = 150.0 / max(1, args.steps + args.warmup)
    res = cpu_build_sample(basis, method, nspin, D2, h, target_seconds=min(20.0, budget), nthreads=cores,
                           repeats=args.steps + args.warmup)
    t = res["seconds"][args.warmup:]
    ms = 1e3 * sum(t) / len(t)
    value = res["npts"] / (ms * 1e-3)
    sample = f"every {res['stride']}-th Morton leaf ({res['nleaves']} leaves, {res['npts']} points) of {args.workload}"
    line = {"impl": "reference", "metric": "xc_build_grid_points_per_s", "value": value, "unit": "grid-points/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "basis_functions": basis.nbf, "grid_points": len(h.sorted_weights),
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": "grid-points/s", "cores": cores, "kind": "port", "sample": sample,
                             "gflops": res["flops"] / (ms * 1e-3) / 1e9},
            "e2e": {"value": value, "unit": "grid-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from occ_b200 import dft, grid, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    t_setup = time.perf_counter()
    basis, method, nspin, quality, D2 = make_inputs(args.workload)
    g = grid.MolecularGrid(basis, grid.GridSettings.from_grid_quality(quality) if quality else None)
    h = g.hierarchy()
    npts_total = len(h.sorted_weights)
    kind = dft.UNRESTRICTED if nspin == 2 else dft.RESTRICTED
    mo = args.density_form == "mo"
    nocc, Cpacked = make_orbitals(basis, nspin)
    ddft = parallel.DistributedDFT(method, basis, h, kind, device=local, batch_bytes=args.batch_mb << 20,
                                   nocc=nocc if mo else None)
    info = ddft.dft.info()
    nbf = basis.nbf
    n = nspin * nbf * nbf
    # normalise the synthetic density to N_electrons with one build (the electron count is linear in D)
    hD = torch.from_numpy(np.ascontiguousarray(D2.ravel(order="F"))).pin_memory()
    hK = torch.empty(n, dtype=torch.float64).pin_memory()
    hC = torch.from_numpy(np.ascontiguousarray(Cpacked)).pin_memory()
    exc, na, nb_ = ddft.build(hD, hK)
    nel = float(basis.num_electrons())
    scale = nel / (na + nb_ if nspin == 2 else na)
    hD.mul_(scale)
    hC.mul_(scale ** 0.5)
    t_setup = time.perf_counter() - t_setup

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ddft.stream)          # the stream the library launches on (not torch's current stream)
        for _ in range(steps):
            fn()
        e1.record(ddft.stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        sync_all()
        return float(ms.item()) / steps

    # ---- device-resident arm ----
    with torch.cuda.stream(ddft.stream):
        ddft.d_D.copy_(hD, non_blocking=True)
        if mo:
            ddft.d_C.copy_(hC, non_blocking=True)

    def step_dev():
        ddft.build_device(mo)

    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(step_dev, args.steps)
    ms_dev_dform = None
    if mo:      # the D-only build of the same handle, for reference beside the factorised form
        for _ in range(2):
            ddft.build_device(False)
        ms_dev_dform = timed(lambda: ddft.build_device(False), args.steps)

    # ---- end-to-end arm (host buffers through the public call) ----
    # N = 1: pinned host D -> H2D -> build -> D2H K.  N > 1: the host matrices live in node-shared pinned memory owned
    # by rank 0 (the SCF process); every rank copies D over its own PCIe link, the packed result is reduce-scattered over
    # NVLink and every rank writes its 1/N slice of K back to the host buffer (occ_b200/parallel.py: build_shared).
    shared = world > 1 and ddft.host_D is not None     # False when /dev/shm cannot hold the two matrices
    if shared:
        if rank == 0:
            ddft.host_D.tensor.copy_(hD)
            if mo:
                ddft.host_C.tensor.copy_(hC)

        def step_e2e():
            ddft.build_shared(mo)
    else:
        def step_e2e():
            ddft.build(hD if rank == 0 else None, hK if rank == 0 else None, hC if rank == 0 else None, mo)

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    exc, na, nb_ = ddft.build_shared(mo) if shared else ddft.build(hD, hK, hC, mo)
    if shared and rank == 0:        # the shared-memory path must reproduce the broadcast / all-reduce path
        exc_b, na_b, nb_b = ddft.build(hD, hK)
        kdiff = float((ddft.host_K.tensor[:n] - hK).abs().max())
        if not (kdiff < 1e-9 and abs(exc - exc_b) < 1e-8):
            raise RuntimeError(f"shared-host e2e path disagrees with the all-reduce path: dK={kdiff} dExc={abs(exc - exc_b)}")
    elif shared:
        ddft.build(None, None)

    # ---- per-kernel times of this rank's shard (CUDA events on the build stream, profiled pass) ----
    ddft.dft.set_profiling(True)
    phases = {}
    nprof = max(1, min(3, args.steps))
    for _ in range(nprof):
        ddft.build_device(mo)
        torch.cuda.synchronize()
        for k, v in ddft.dft.phase_ms().items():
            phases[k] = phases.get(k, 0.0) + v / nprof
    ddft.dft.set_profiling(False)

    flops_total = info["flops_per_build"]
    exec_total = 0.5 * info["flops_per_build"] + (info["mo_density_flops"] if mo else 0.5 * info["flops_per_build"])
    nlaunch = float(info["launches_per_build_mo"] if mo else info["launches_per_build"])
    if world > 1:
        ft = torch.tensor([flops_total, nlaunch, exec_total], dtype=torch.float64, device=dev)
        dist.all_reduce(ft)
        flops_total, launches_total, exec_total = float(ft[0]), float(ft[1]), float(ft[2])
    else:
        launches_total = nlaunch

    if rank == 0:
        g_fac = 16.0 if info["deriv"] >= 2 else 4.0
        # the two DMMA contractions do g/2 * nspin * sum(n nb^2) algorithmic flops each
        kern = {}
        executed = {"density": info["mo_density_flops"] if mo else 0.5 * info["flops_per_build"],
                    "vxc": 0.5 * info["flops_per_build"]}
        for name in ("density", "vxc"):
            fl = executed[name]      # tensor flops the phase really issues (the factorised density issues fewer)
            ms_launch = phases[name] / max(1, info["nbatches"])
            kern[name] = {"ms_per_build": phases[name], "launches_per_build": info["nbatches"],
                          "avg_launch_ms": ms_launch, "tflops": fl / (phases[name] * 1e-3) / 1e12 if phases[name] > 0 else 0.0}
        dom = max(kern, key=lambda k: kern[k]["ms_per_build"])
        achieved = kern[dom]["tflops"]
        traffic = None
        try:   # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json"))).get(args.workload)
            if tj and tj.get("batch_mb") == args.batch_mb and world == 1:
                traffic = tj.get(KERNEL_NAMES[dom])
        except Exception:
            pass
        cpu = None
        if not args.no_cpu_baseline:
            res = cpu_build_sample(basis, method, nspin, np.asfortranarray(hD.numpy().reshape(nbf, nspin * nbf).T), h,
                                   target_seconds=args.cpu_seconds)
            t = res["seconds"][0]
            cpu = {"value": res["npts"] / t, "unit": "grid-points/s", "cores": res["cores"], "kind": "port",
                   "sample": f"every {res['stride']}-th Morton leaf ({res['nleaves']} leaves, {res['npts']} points) of "
                             f"{args.workload}, one build, {t:.1f} s",
                   "gflops": res["flops"] / t / 1e9}
        cosx_row = None
        if world == 1 and not args.no_cosx:
            try:     # SURVEY 8(f) rank 2, reported beside the headline (never allowed to break it)
                cosx_row = cosx_next_row(basis, WORKLOADS[args.workload][0], WORKLOADS[args.workload][1])
            except Exception as e:  # noqa: BLE001
                cosx_row = {"error": str(e)[:200]}
        line = {
            "metric": "xc_build_grid_points_per_s", "value": npts_total / (ms_dev * 1e-3), "unit": "grid-points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "molecule": WORKLOADS[args.workload][0],
                       "basis": WORKLOADS[args.workload][1], "functional": method, "nspin": nspin,
                       "basis_functions": nbf, "grid_points": npts_total, "leaves": int(h.num_leaves),
                       "nb_local_min_max": [info["nb_min"], info["nb_max"]], "batch_mb": args.batch_mb,
                       "parallelism": f"grid leaves sharded over {world} GPU(s), D replicated, NCCL all-reduce of K",
                       "l2": "per-step working set (AO workspace, %.1f GB streamed per build) exceeds the 126 MB L2; "
                             "no explicit flush" % _workspace_gb(info, nspin)},
            "xc_energy": exc, "n_electrons": na + (nb_ if nspin == 2 else 0.0),
            "algorithmic_tflops": flops_total / (ms_dev * 1e-3) / 1e12,
            "executed_tensor_tflops": exec_total / (ms_dev * 1e-3) / 1e12,
            # the whole-build fraction is quoted against the D-based algorithmic count even when the density
            # runs in the factorised form; the fraction of the peak the tensor pipe really sustained is given beside it
            "fp64_tensor_frac_whole_build": flops_total / (ms_dev * 1e-3) / 1e12 / (FP64_TENSOR_PEAK_TFLOPS * world),
            "executed_tensor_frac_whole_build": exec_total / (ms_dev * 1e-3) / 1e12 / (FP64_TENSOR_PEAK_TFLOPS * world),
            "density_form": ("factorised where cheaper: psi = phi C, X = psi C^T (4 n nb nocc instead of 2 n nb^2 flops; "
                             "occ_xc_build_mo, the seam's `mo` carries C); algorithmic_tflops / fp64_tensor_frac_whole_build count the "
                             "D-based 4 n nb^2 of SURVEY 8d, executed_tensor_* the DMMA flops really issued"
                             if mo else "D form (phi D)"),
            "d_form": None if ms_dev_dform is None else {
                "ms_per_step": ms_dev_dform, "value": npts_total / (ms_dev_dform * 1e-3),
                "algorithmic_tflops": flops_total / (ms_dev_dform * 1e-3) / 1e12,
                "fp64_tensor_frac_whole_build": flops_total / (ms_dev_dform * 1e-3) / 1e12 / (FP64_TENSOR_PEAK_TFLOPS * world)},
            "roofline": {"bound": "tensor", "kernel": KERNEL_NAMES[dom], "achieved": achieved, "peak": FP64_TENSOR_PEAK_TFLOPS,
                         "unit": "TFLOP/s", "frac": achieved / FP64_TENSOR_PEAK_TFLOPS, "traffic": traffic,
                         "algorithmic_flops_per_launch": 0.5 * info["flops_per_build"] / max(1, info["nbatches"]),
                         "peak_source": "measured DMMA.8x8x4 microbenchmark on this pool's B200 "
                                        "(profiles/r01_fp64_peaks_b200.jsonl); MEASURED_PEAKS.json has no FP64 entry",
                         "kernels": kern},
            "phases_ms": phases,
            "cpu_baseline": cpu,
            "e2e": {"value": npts_total / (ms_e2e * 1e-3), "unit": "grid-points/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": 8 * (n + (hC.numel() if mo else 0)) * (world if shared else 1),
                    "d2h_bytes_per_step": 8 * n + 32 * world,
                    "path": ("pinned host D -> H2D -> build -> D2H of K and the 3 scalars" if world == 1 else
                             "rank 0: pinned host D -> H2D; NCCL broadcast of D; sharded build; NCCL all-reduce of [K|Exc,Na,Nb]; "
                             "rank 0: D2H of K (no room in /dev/shm for the shared-host path)" if not shared else
                             "host D / K in node-shared pinned memory owned by rank 0; every rank: H2D of D over its own "
                             "PCIe link; sharded build; NCCL reduce-scatter of [K|Exc,Na,Nb]; every rank: D2H of its 1/N "
                             "slice into the host K")},
            "gpu_launches": int(launches_total * args.steps),
            "clocks": clocks, "setup_seconds": t_setup,
            "next_rows": {"cosx": cosx_row},
        }
        print(json.dumps(line), flush=True)
    ddft.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cosx_next_row(basis, mol, bas, reps=2):
    """SemiNumericalExchange.compute_K (host D [+ C_occ] -> K, H2D / D2H inside the timed region) of the same molecule and
    basis on a 50-point sgx grid; restricted, full Coulomb operator. Headline: the factorised-F entry
    occ_cosx_compute_K_mo; the D-only occ_cosx_compute_K beside it."""
    import time

    from occ_b200 import cosx, molecules
    t0 = time.perf_counter()
    sgx = cosx.SemiNumericalExchange(basis, cosx.sgx_grid_settings(50, 1e-5))
    t_setup = time.perf_counter() - t0
    D = 0.5 * molecules.synthetic_density(basis, 1)
    Cocc = molecules.synthetic_orbitals(basis, 1)[0]      # 2 D = Cocc Cocc^T: the orbitals behind synthetic_density
    out = {}
    for name, call in (("occ_cosx_compute_K_mo", lambda: sgx.compute_K_mo(D, Cocc)),
                       ("occ_cosx_compute_K", lambda: sgx.compute_K(D))):
        call()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            K = call()
            ts.append(time.perf_counter() - t0)
        info = sgx.info()
        out[name] = {"ms_per_call": 1e3 * min(ts), "gemm_flops": info["gemm_flops"], "kernel_launches": info["launches"],
                     "library_gemm_calls": info["gemm_calls"], "K_absmax": float(np.abs(K).max())}
    n = sgx.num_grid_points()
    sgx.close()
    head = out["occ_cosx_compute_K_mo"]
    return {"entry": "occ_cosx_compute_K_mo", "workload": f"{mol}/{bas} RKS, sgx grid 50 angular / radial precision 1e-5",
            "grid_points": n, "batches": info["nbatches"], "shell_pairs": info["npairs"], "nocc": int(Cocc.shape[1]),
            "ms_per_call": head["ms_per_call"], "grid_points_per_s": n / (1e-3 * head["ms_per_call"]),
            "gemm_flops": head["gemm_flops"], "kernel_launches": head["kernel_launches"],
            "library_gemm_calls": head["library_gemm_calls"],
            "h2d_bytes": 8 * basis.nbf ** 2 + 8 * basis.nbf * int(Cocc.shape[1]), "d2h_bytes": 8 * basis.nbf ** 2,
            "K_absmax": head["K_absmax"], "d_form": dict(out["occ_cosx_compute_K"], entry="occ_cosx_compute_K"),
            "setup_seconds": t_setup}


def _workspace_gb(info, nspin):
    # AO slabs streamed per build: (ncomp + nspin) slabs of 128 x nb doubles per tile ~ sum over leaves
    ncomp = 4 if info["deriv"] >= 1 else 1
    mean_nb = (info["sum_n_nb2"] / max(1.0, float(info["npts"]))) ** 0.5
    return 8.0 * (ncomp + nspin) * info["npts"] * mean_nb / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--batch-mb", type=int, default=4096)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cosx", action="store_true", help="skip the COSX next-row measurement at N = 1")
    ap.add_argument("--density-form", default="mo", choices=["mo", "d"],
                    help="mo: factorised density where cheaper (occ_xc_build_mo, needs the occupied orbitals); d: D only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
