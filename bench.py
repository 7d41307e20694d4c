#!/usr/bin/env python
"""bench.py — the Vxc+Exc build per SCF iteration (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]
  (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

A step = one DFT::compute_K_dft-equivalent build (include/occ/dft/dft.h:190-388) on a fixed synthetic density.
  value   : grid-points/s with the inputs (D2 and C_occ) resident in HBM at the start and K left in HBM (CUDA events on
            the library's stream, max over ranks; the N > 1 number includes the NCCL all-reduce of [K | Exc, N_a, N_b])
  e2e     : the same through the host-facing call, host buffers in and out.  The seam's argument `mo` carries the occupied
            orbitals (include/occ/qm/mo.h), so the factorised entry uploads ONLY C_occ (8 nbf nocc bytes) and forms
            D2 = C C^T on the device; K and the three scalars come back over PCIe.  `e2e.pageable` is the same call on
            plain (unpinned) host arrays, the way an Eigen-backed caller would make it.
  roofline: the dominant DMMA contraction kernel against the FP64 tensor-pipe rate MEASURED IN THIS RUN
            (occ_xc_measure_fp64_tensor_peak: register-resident mma.sync.m8n8k4.f64 at the clocks of the run;
            MEASURED_PEAKS.json carries no FP64 entry)
  configs : device / e2e ms and tensor fractions of all five BASELINE.json configurations (N = 1)
  cpu_baseline / --impl reference: the CPU restatement of the reference algorithm (oracle/: OpenMP over leaves +
            OpenBLAS), on the box's host cores.  A step processes every s-th Morton leaf; the per-build fixed cost
            (shell extents, the per-thread full-size K copies of dft.h:221-232 and their serial reduction) is timed
            separately from the leaf loop, and ONLY the leaf loop is extrapolated to the whole grid:
            full_build_s = fixed_s + leaf_loop_s * (all leaves / sampled leaves).  The reference arm builds its grid
            with the oracle as well: it never loads the product library.
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

FP64_TENSOR_PEAK_R01 = 36.9  # round-1 figure (profiles/r01_fp64_peaks_b200.jsonl), kept for comparison only

WORKLOADS = {
    # name: (molecule, basis, method, nspin, grid quality or None for the reference default); BASELINE.json configs[0..4]
    "water_b3lyp_svp": ("water", "def2-svp", "b3lyp", 1, None),
    "benzene_pbe_tzvp": ("benzene", "def2-tzvp", "pbe", 1, None),
    "h2o32_b3lyp_svp_uks": ("h2o_32", "def2-svp", "b3lyp", 2, None),
    "valinomycin_b3lyp_tzvp": ("valinomycin", "def2-tzvp", "b3lyp", 1, None),
    # "fine grid" = 590 angular points (GridQuality::VeryFine: Fine equals the default)
    "taxol_r2scan_tzvp": ("taxol", "def2-tzvp", "r2scan", 1, "veryfine"),
    "taxol_r2scan_tzvp_default_grid": ("taxol", "def2-tzvp", "r2scan", 1, None),
}
BASELINE_CONFIGS = ["water_b3lyp_svp", "benzene_pbe_tzvp", "h2o32_b3lyp_svp_uks", "valinomycin_b3lyp_tzvp",
                    "taxol_r2scan_tzvp"]
DEFAULT_WORKLOAD = "valinomycin_b3lyp_tzvp"
KERNEL_NAMES = {"density": "k_mo_psi + k_mo_density + k_density", "density_d": "k_density", "vxc": "k_vxc_resident"}


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
    return [c.shape[1] for c in Cs], np.concatenate([c.ravel(order="F") for c in Cs])


def common_config(workload, basis, method, nspin, npts, nleaves):
    """identical in both arms (the driver compares the two `config` objects)"""
    mol, bas, _, _, quality = WORKLOADS[workload]
    return {"workload": workload, "molecule": mol, "basis": bas, "functional": method, "nspin": nspin,
            "grid": "reference default (302/110 angular, LMG 1e-12, NWChem pruning)" if quality is None else
                    f"GridQuality::{quality} (590/302 angular, LMG 1e-12, NWChem pruning)",
            "basis_functions": int(basis.nbf), "grid_points": int(npts), "leaves": int(nleaves),
            "l2": "no explicit flush: one build streams its AO workspace (GBs) and the K / D matrices through HBM, "
                  "far beyond the 126 MB L2"}


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


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference algorithm, host cores only
# ---------------------------------------------------------------------------------------------------------------
def oracle_grid(basis, quality):
    """the molecular grid + Morton hierarchy from the ORACLE's grid code (no product library involved)"""
    from oracle import oracle as O
    ob = O.Basis(basis.as_arrays())
    s = O.default_grid_settings() if quality is None else \
        O.default_grid_settings(max_angular_points=590, min_angular_points=302)
    pts, w, _ = O.molecular_grid(basis.atomic_numbers, basis.positions, ob, s)
    spts, sw, perm, leaves = O.hierarchy(pts, w)
    return ob, spts, sw, leaves


def cpu_pick_threads(ob, spts, sw, leaves, names, nspin, D2):
    """thread count with the best leaf-loop throughput on a thin probe (more threads was slower on some boxes)"""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    cands = sorted({cores, max(1, cores // 2)}, reverse=True)
    # >= 40 leaves per thread: with fewer, the imbalance of a dynamic schedule over unequal leaves decides the probe
    probe = leaves[::max(1, len(leaves) // max(96, 40 * cores))]
    best, rows = None, []
    for nt in cands:
        O.xc_build(ob, spts, sw, probe[:8], names, nspin, D2, nthreads=nt)        # touch the per-thread buffers
        r = O.xc_build(ob, spts, sw, probe, names, nspin, D2, nthreads=nt)
        rows.append({"threads": nt, "leaf_loop_s": r["stats"]["t_leaf_loop"], "fixed_s": r["stats"]["t_fixed"]})
        if best is None or r["stats"]["t_leaf_loop"] < best[1]:
            best = (nt, r["stats"]["t_leaf_loop"])
    return best[0], rows


def cpu_build_sample(ob, spts, sw, leaves, names, nspin, D2, target_seconds, repeats=1, whole_build_budget_s=0.0):
    """Time the CPU restatement on every `stride`-th Morton leaf.  Returns the per-step wall times with the leaf loop and
    the fixed part separated, and the whole-build time extrapolated from the leaf loop alone."""
    from oracle import oracle as O
    nthreads, thread_probe = cpu_pick_threads(ob, spts, sw, leaves, names, nspin, D2)
    probe = leaves[::max(1, len(leaves) // 64)]
    r = O.xc_build(ob, spts, sw, probe, names, nspin, D2, nthreads=nthreads)
    per_leaf = r["stats"]["t_leaf_loop"] / max(1, len(probe))
    fixed = r["stats"]["t_fixed"]
    full_est = fixed + per_leaf * len(leaves)
    stride = max(1, int(np.ceil(per_leaf * len(leaves) / max(1e-3, target_seconds - fixed))))
    sample = leaves[::stride]
    steps = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = O.xc_build(ob, spts, sw, sample, names, nspin, D2, nthreads=nthreads)
        wall = time.perf_counter() - t0
        steps.append({"wall_s": wall, "leaf_loop_s": r["stats"]["t_leaf_loop"], "fixed_s": r["stats"]["t_fixed"]})
    whole = None
    if stride > 1 and full_est < whole_build_budget_s:      # one build of the WHOLE grid anchors the extrapolation
        t0 = time.perf_counter()
        rw = O.xc_build(ob, spts, sw, leaves, names, nspin, D2, nthreads=nthreads)
        whole = {"wall_s": time.perf_counter() - t0, "leaf_loop_s": rw["stats"]["t_leaf_loop"],
                 "fixed_s": rw["stats"]["t_fixed"], "exc": rw["exc"], "n_electrons": rw["n_alpha"] + rw["n_beta"] * (nspin == 2)}
    g = 16.0 if O.method_deriv(names) >= 2 else 4.0
    return {"steps": steps, "stride": stride, "nleaves_sample": len(sample), "nleaves": len(leaves),
            "npts_sample": int(sample["count"].sum()), "threads": nthreads, "thread_probe": thread_probe,
            "flops_sample": g * nspin * r["stats"]["sum_n_nb2"], "whole_build": whole}


def cpu_summary(res, npts_total, skip=0):
    st = res["steps"][skip:] or res["steps"]
    loop = statistics.median(s["leaf_loop_s"] for s in st)
    fixed = statistics.median(s["fixed_s"] for s in st)
    ratio = res["nleaves"] / max(1, res["nleaves_sample"])
    full = fixed + loop * ratio
    return {"leaf_loop_s": loop, "fixed_s": fixed, "sample_wall_s": statistics.median(s["wall_s"] for s in st),
            "extrapolated_full_build_s": full, "value": npts_total / full,
            "gflops": res["flops_sample"] / loop / 1e9 if loop > 0 else 0.0}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle restatement; the reference itself cannot be built here —
    Eigen / libxc / oneTBB are network-fetched) on the host cores; never touches occ_b200/libocc_xc_b200.so."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    basis, method, nspin, quality, D2 = make_inputs(args.workload)
    names = O.METHODS[method]
    t0 = time.perf_counter()
    ob, spts, sw, leaves = oracle_grid(basis, quality)
    t_grid = time.perf_counter() - t0
    npts = len(sw)
    # scale the synthetic density to N_electrons like the GPU arm does (the count is linear in D); here the integral is
    # estimated from every 16th leaf, which fixes the magnitude of the input to a fraction of a percent
    thin = leaves[::16]
    r0 = O.xc_build(ob, spts, sw, thin, names, nspin, D2)
    n_est = (r0["n_alpha"] + (r0["n_beta"] if nspin == 2 else 0.0)) * len(leaves) / max(1, len(thin))
    D2 = np.asfortranarray(D2 * (float(basis.num_electrons()) / n_est))
    nsteps = args.steps + args.warmup
    budget = 170.0
    res = cpu_build_sample(ob, spts, sw, leaves, names, nspin, D2, target_seconds=max(1.0, min(20.0, budget / nsteps)),
                           repeats=nsteps, whole_build_budget_s=45.0)
    s = cpu_summary(res, npts, skip=args.warmup)
    sample = (f"every {res['stride']}-th Morton leaf ({res['nleaves_sample']} of {res['nleaves']} leaves, "
              f"{res['npts_sample']} points) per step; value = N_grid / (fixed_s + leaf_loop_s * {res['nleaves']}/"
              f"{res['nleaves_sample']}): only the leaf loop is extrapolated")
    line = {"impl": "reference", "metric": "xc_build_grid_points_per_s", "value": s["value"], "unit": "grid-points/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * s["sample_wall_s"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": common_config(args.workload, basis, method, nspin, npts, len(leaves)),
            "cpu_baseline": {"value": s["value"], "unit": "grid-points/s", "cores": res["threads"], "kind": "port",
                             "sample": sample, "leaf_loop_s": s["leaf_loop_s"], "fixed_s": s["fixed_s"],
                             "extrapolated_full_build_s": s["extrapolated_full_build_s"], "gflops": s["gflops"],
                             "host_cores": os.cpu_count(), "thread_probe": res["thread_probe"],
                             "whole_build": res["whole_build"], "grid_setup_s": t_grid},
            "e2e": {"value": s["value"], "unit": "grid-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
def measure_peak(device):
    import ctypes as C
    from occ_b200 import _lib
    tf, ms = C.c_double(), C.c_double()
    _lib.check(_lib.lib().occ_xc_measure_fp64_tensor_peak(C.c_int(device), C.c_double(60.0), C.byref(tf), C.byref(ms)))
    return tf.value, ms.value


def setup_case(workload, local, world_opts, batch_mb, mo=True):
    import torch
    from occ_b200 import dft, grid, parallel
    t0 = time.perf_counter()
    basis, method, nspin, quality, D2 = make_inputs(workload)
    g = grid.MolecularGrid(basis, grid.GridSettings.from_grid_quality(quality) if quality else None)
    h = g.hierarchy()
    kind = dft.UNRESTRICTED if nspin == 2 else dft.RESTRICTED
    nocc, Cpacked = make_orbitals(basis, nspin)
    ddft = parallel.DistributedDFT(method, basis, h, kind, device=local, batch_bytes=batch_mb << 20,
                                   nocc=nocc if mo else None, **world_opts)
    nbf = basis.nbf
    n = nspin * nbf * nbf
    hD = torch.from_numpy(np.ascontiguousarray(D2.ravel(order="F"))).pin_memory()
    hK = torch.empty(n, dtype=torch.float64).pin_memory()
    hC = torch.from_numpy(np.ascontiguousarray(Cpacked)).pin_memory()
    # scale the synthetic density to N_electrons with one build (the electron count is linear in D); this only fixes
    # the magnitude of the inputs — correctness is the business of tests/, which compare with the CPU oracle
    exc, na, nb_ = ddft.build(hD, hK)
    scale = float(basis.num_electrons()) / (na + nb_ if nspin == 2 else na)
    hD.mul_(scale)
    hC.mul_(scale ** 0.5)
    return {"basis": basis, "method": method, "nspin": nspin, "h": h, "ddft": ddft, "hD": hD, "hK": hK, "hC": hC,
            "n": n, "nbf": nbf, "nocc": nocc, "setup_s": time.perf_counter() - t0, "info": ddft.dft.info()}


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    mo = args.density_form == "mo"

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(ddft, fn, steps):
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

    def measure(c, steps, warmup, with_dform):
        """device-resident and end-to-end ms per build of one prepared case"""
        ddft, hD, hK, hC = c["ddft"], c["hD"], c["hK"], c["hC"]
        with torch.cuda.stream(ddft.stream):
            ddft.d_D.copy_(hD, non_blocking=True)
            if mo:
                ddft.d_C.copy_(hC, non_blocking=True)
        for _ in range(warmup):
            ddft.build_device(mo)
        out = {"device_ms": timed(ddft, lambda: ddft.build_device(mo), steps)}
        if world == 1:                   # the same build replayed from a CUDA graph (no launch gaps: matters for small systems)
            try:
                g = ddft.capture_device_build(mo)

                def replay():
                    with torch.cuda.stream(ddft.stream):
                        g.replay()
                for _ in range(2):
                    replay()
                out["device_graph_ms"] = timed(ddft, replay, steps)
                del g
            except Exception as e:                       # noqa: BLE001 - reported, the plain launch path is the headline
                out["device_graph_ms"] = None
                out["device_graph_error"] = str(e)[:200]
        if mo and with_dform:      # the D-only build of the same handle, for reference beside the factorised form
            for _ in range(2):
                ddft.build_device(False)
            out["device_ms_d_form"] = timed(ddft, lambda: ddft.build_device(False), steps)
        # end-to-end: N = 1 pinned host C (or D) -> H2D -> build -> D2H of K and the scalars.  N > 1: host matrices in
        # node-shared pinned memory owned by rank 0 (the SCF process); every rank copies the inputs over its own PCIe
        # link, the packed result is reduce-scattered over NVLink, every rank writes its 1/N slice of K to the host
        shared = world > 1 and ddft.host_K is not None
        if shared:
            if rank == 0:
                ddft.host_D.tensor.copy_(hD)
                if mo:
                    ddft.host_C.tensor.copy_(hC)
            step = lambda: ddft.build_shared(mo)                                   # noqa: E731
        else:
            step = lambda: ddft.build(hD if rank == 0 else None, hK if rank == 0 else None,   # noqa: E731
                                      hC if rank == 0 else None, mo)
        for _ in range(min(warmup, 2)):
            step()
        out["e2e_ms"] = timed(ddft, step, steps)
        out["shared"] = shared
        out["scalars"] = step()
        return out

    c = setup_case(args.workload, local, {}, args.batch_mb, mo)
    ddft, info, basis, nspin, nbf, n = c["ddft"], c["info"], c["basis"], c["nspin"], c["nbf"], c["n"]
    npts_total = len(c["h"].sorted_weights)
    peak0, _ = measure_peak(local)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    m = measure(c, args.steps, args.warmup, True)
    clocks = sampler.stop() if rank == 0 else None
    peak1, _ = measure_peak(local)
    peak = max(peak0, peak1)
    ms_dev, ms_e2e, shared = m["device_ms"], m["e2e_ms"], m["shared"]
    exc, na, nb_ = m["scalars"]
    if shared and rank == 0:        # the shared-memory path must reproduce the broadcast / all-reduce path
        exc_b, na_b, nb_b = ddft.build(c["hD"], c["hK"], c["hC"], mo)
        kdiff = float((ddft.host_K.tensor[:n] - c["hK"]).abs().max())
        if not (kdiff < 1e-9 and abs(exc - exc_b) < 1e-8):
            raise RuntimeError(f"shared-host e2e path disagrees with the all-reduce path: dK={kdiff} dExc={abs(exc - exc_b)}")
    elif shared:
        ddft.build(None, None, None, mo)

    # e2e on plain pageable host arrays at N = 1 (what an Eigen-backed caller hands over without registering its storage)
    pageable_ms = None
    if world == 1:
        Dp, Cp = np.array(c["hD"].numpy()), np.array(c["hC"].numpy())
        Kp = np.empty(n)
        call = (lambda: ddft.dft.build_mo_raw(0, Cp.ctypes.data, Kp.ctypes.data)) if mo else \
               (lambda: ddft.dft.build_raw(Dp.ctypes.data, Kp.ctypes.data))
        call()
        ts = []
        for _ in range(max(3, min(args.steps, 5))):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        pageable_ms = 1e3 * statistics.median(ts)

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
    exec_total = info["vxc_flops"] + (info["mo_density_flops"] if mo else 0.5 * info["flops_per_build"])
    nlaunch = float(info["launches_per_build_mo"] if mo else info["launches_per_build"])
    if world > 1:
        ft = torch.tensor([flops_total, nlaunch, exec_total], dtype=torch.float64, device=dev)
        dist.all_reduce(ft)
        flops_total, launches_total, exec_total = float(ft[0]), float(ft[1]), float(ft[2])
    else:
        launches_total = nlaunch

    # ---- all five BASELINE configurations (N = 1): device / e2e ms, tensor fractions ----
    configs = None
    if world == 1 and not args.no_configs:
        configs = {}
        for wl in BASELINE_CONFIGS:
            try:
                if wl == args.workload:
                    cc, mm = c, m
                else:
                    cc = setup_case(wl, local, {}, args.batch_mb, mo)
                    mm = measure(cc, max(3, min(args.steps, 10)), 3, False)
                ii = cc["info"]
                ex = ii["vxc_flops"] + (ii["mo_density_flops"] if mo else 0.5 * ii["flops_per_build"])
                npt = len(cc["h"].sorted_weights)
                configs[wl] = {"config": common_config(wl, cc["basis"], cc["method"], cc["nspin"], npt, cc["h"].num_leaves),
                               "device_ms": mm["device_ms"], "e2e_ms": mm["e2e_ms"],
                               "device_graph_ms": mm.get("device_graph_ms"),
                               "grid_points_per_s": npt / (mm["device_ms"] * 1e-3),
                               "nb_local_min_max": [ii["nb_min"], ii["nb_max"]], "nocc": cc["nocc"],
                               "algorithmic_flops": ii["flops_per_build"], "executed_tensor_flops": ex,
                               "algorithmic_tensor_frac": ii["flops_per_build"] / (mm["device_ms"] * 1e-3) / 1e12 / peak,
                               "executed_tensor_frac": ex / (mm["device_ms"] * 1e-3) / 1e12 / peak,
                               "launches_per_build": ii["launches_per_build_mo"] if mo else ii["launches_per_build"],
                               "xc_energy": mm["scalars"][0], "setup_s": cc["setup_s"]}
                if cc is not c:
                    cc["ddft"].close()
            except Exception as e:  # noqa: BLE001
                configs[wl] = {"error": str(e)[:300]}

    if rank == 0:
        kern = {}
        executed = {"density": info["mo_density_flops"] if mo else 0.5 * info["flops_per_build"],
                    "vxc": info["vxc_flops"]}
        for name in ("density", "vxc"):
            fl = executed[name]      # tensor flops the phase really issues (the factorised density issues fewer)
            kern[name] = {"kernel": KERNEL_NAMES[name if (mo or name == "vxc") else "density_d"],
                          "ms_per_build": phases[name], "launches_per_build": info["nbatches"],
                          "avg_launch_ms": phases[name] / max(1, info["nbatches"]),
                          "tflops": fl / (phases[name] * 1e-3) / 1e12 if phases[name] > 0 else 0.0}
        dom = max(kern, key=lambda k: kern[k]["ms_per_build"])
        achieved = kern[dom]["tflops"]
        traffic = None
        try:   # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(args.workload)
            if tj and tj.get("batch_mb") == args.batch_mb and world == 1:
                traffic = tj.get("k_vxc_resident" if dom == "vxc" else "k_density")
        except Exception:
            pass
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import oracle as O
            names = O.METHODS[c["method"]]
            ob = O.Basis(basis.as_arrays())
            Dn = np.asfortranarray(c["hD"].numpy().reshape(nbf, nspin * nbf).T)
            h = c["h"]
            res = cpu_build_sample(ob, h.sorted_points, h.sorted_weights, h.leaves, names, nspin, Dn,
                                   target_seconds=args.cpu_seconds)
            s = cpu_summary(res, npts_total)
            cpu = {"value": s["value"], "unit": "grid-points/s", "cores": res["threads"], "kind": "port",
                   "sample": f"every {res['stride']}-th Morton leaf ({res['nleaves_sample']} of {res['nleaves']} leaves, "
                             f"{res['npts_sample']} points) of {args.workload}, one build; only the leaf loop is extrapolated",
                   "leaf_loop_s": s["leaf_loop_s"], "fixed_s": s["fixed_s"],
                   "extrapolated_full_build_s": s["extrapolated_full_build_s"], "gflops": s["gflops"],
                   "host_cores": os.cpu_count(), "thread_probe": res["thread_probe"]}
        cosx_row = None
        if world == 1 and not args.no_cosx:
            try:     # SURVEY 8(f) rank 2, reported beside the headline (never allowed to break it)
                cosx_row = cosx_next_row(basis, WORKLOADS[args.workload][0], WORKLOADS[args.workload][1], peak)
            except Exception as e:  # noqa: BLE001
                cosx_row = {"error": str(e)[:200]}
        grad_row = vv10_row = None
        if world == 1 and not args.no_cosx and nspin == 1:
            try:     # SURVEY 8(f) ranks 1 and 4 on the same molecule (host D -> host result, one call each)
                grad_row, vv10_row = gradient_and_vv10_next_rows(basis, c["method"], args.workload)
            except Exception as e:  # noqa: BLE001
                grad_row = {"error": str(e)[:200]}
        nC = int(c["hC"].numel())
        h2d = 8 * (nC if mo else n)
        line = {
            "metric": "xc_build_grid_points_per_s", "value": npts_total / (ms_dev * 1e-3), "unit": "grid-points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": common_config(args.workload, basis, c["method"], nspin, npts_total, c["h"].num_leaves),
            "gpu_config": {"nb_local_min_max": [info["nb_min"], info["nb_max"]], "batch_mb": args.batch_mb,
                           "nocc": c["nocc"],
                           "parallelism": f"grid leaves sharded over {world} GPU(s), inputs replicated, NCCL all-reduce of K",
                           "workspace_streamed_gb_per_build": _workspace_gb(info, nspin)},
            "xc_energy": exc, "n_electrons": na + (nb_ if nspin == 2 else 0.0),
            "algorithmic_tflops": flops_total / (ms_dev * 1e-3) / 1e12,
            "executed_tensor_tflops": exec_total / (ms_dev * 1e-3) / 1e12,
            # the whole-build fraction is quoted against the D-based algorithmic count even when the density
            # runs in the factorised form; the fraction of the peak the tensor pipe really sustained is given beside it
            "fp64_tensor_frac_whole_build": flops_total / (ms_dev * 1e-3) / 1e12 / (peak * world),
            "executed_tensor_frac_whole_build": exec_total / (ms_dev * 1e-3) / 1e12 / (peak * world),
            "density_form": ("factorised where cheaper: psi = phi C, X = psi C^T (4 n nb nocc instead of 2 n nb^2 flops; "
                             "occ_xc_build_mo, the seam's `mo` carries C); algorithmic_* count the D-based 4 n nb^2 of "
                             "SURVEY 8d, executed_tensor_* the DMMA flops really issued" if mo else "D form (phi D)"),
            "d_form": None if "device_ms_d_form" not in m else {
                "ms_per_step": m["device_ms_d_form"], "value": npts_total / (m["device_ms_d_form"] * 1e-3),
                "algorithmic_tflops": flops_total / (m["device_ms_d_form"] * 1e-3) / 1e12,
                "fp64_tensor_frac_whole_build": flops_total / (m["device_ms_d_form"] * 1e-3) / 1e12 / (peak * world)},
            "roofline": {"bound": "tensor", "kernel": kern[dom]["kernel"], "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                         "algorithmic_flops_per_launch": executed[dom] / max(1, info["nbatches"]),
                         "peak_source": "occ_xc_measure_fp64_tensor_peak in this run: register-resident DMMA.8x8x4, 2 CTAs x 8 "
                                        "warps per SM, ~60 ms, max of the runs before and after the timed region",
                         "peak_before_after": [peak0, peak1], "peak_round1_constant": FP64_TENSOR_PEAK_R01,
                         "kernels": kern},
            "phases_ms": phases,
            "configs": configs,
            "cpu_baseline": cpu,
            "e2e": {"value": npts_total / (ms_e2e * 1e-3), "unit": "grid-points/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d * (world if shared else 1),
                    "d2h_bytes_per_step": 8 * n + 32 * world,
                    "pageable_ms_per_step": pageable_ms,
                    "path": (("pinned host C_occ -> H2D -> D2 = C C^T on the device" if mo else "pinned host D2 -> H2D") +
                             " -> build -> D2H of K and the 3 scalars" if world == 1 else
                             "rank 0: pinned host inputs -> H2D; NCCL broadcast; sharded build; NCCL all-reduce of [K|Exc,Na,Nb]; "
                             "rank 0: D2H of K (no room in /dev/shm for the shared-host path)" if not shared else
                             "host " + ("C_occ" if mo else "D2") + " / K in node-shared pinned memory owned by rank 0; every "
                             "rank: H2D over its own PCIe link" + (", D2 = C C^T on the device" if mo else "") +
                             "; sharded build; NCCL reduce-scatter of [K|Exc,Na,Nb]; every rank: D2H of its 1/N slice "
                             "into the host K")},
            "gpu_launches": int(launches_total * args.steps),
            "clocks": clocks, "setup_seconds": c["setup_s"],
            "next_rows": {"cosx": cosx_row, "xc_gradient": grad_row, "vv10": vv10_row},
        }
        print(json.dumps(line), flush=True)
    ddft.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cosx_next_row(basis, mol, bas, peak, reps=2):
    """SemiNumericalExchange.compute_K (host D [+ C_occ] -> K, H2D / D2H inside the timed region) of the same molecule and
    basis on a 50-point sgx grid; restricted, full Coulomb operator. Headline: the factorised-F entry
    occ_cosx_compute_K_mo; the D-only occ_cosx_compute_K beside it."""
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


def gradient_and_vv10_next_rows(basis, method, workload, reps=2):
    """occ_xc_gradient_mo on the workload's own grid and occ_xc_vv10 / occ_xc_vv10_gradient on the reference's default
    NLC grid (host D2 -> host result; restricted)."""
    from occ_b200 import dft, grid, molecules, nlc
    quality = WORKLOADS[workload][4]
    h = grid.MolecularGrid(basis, grid.GridSettings.from_grid_quality(quality) if quality else None).hierarchy()
    D2 = molecules.synthetic_density(basis, 1)
    Cs = molecules.synthetic_orbitals(basis, 1)

    def best(call):
        call()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            r = call()
            ts.append(time.perf_counter() - t0)
        return 1e3 * min(ts), r

    d = dft.DFT(method, basis, h, gradient=True, nocc=Cs[0].shape[1])
    ms_g, g = best(lambda: d.gradient_mo(D2, Cs))
    ms_gd, _ = best(lambda: d.gradient(D2))
    d.close()
    grad_row = {"entry": "occ_xc_gradient_mo", "workload": workload, "grid_points": len(h.sorted_weights),
                "ms_per_call": ms_g, "d_form_ms_per_call": ms_gd, "gradient_absmax": float(np.abs(g).max())}
    f = nlc.NonLocalCorrelationFunctional()
    f.set_integration_grid(basis, screening_threshold=1e-10, gradient=True)
    ms_e, (e, _) = best(lambda: f.vv10(D2))
    ms_vg, (_, gv) = best(lambda: f.vv10_gradient(D2))
    n = f.num_grid_points
    f.close()
    vv10_row = {"entry": "occ_xc_vv10 / occ_xc_vv10_gradient", "workload": f"{workload} density, default NLC grid (110/50, 1e-7)",
                "nlc_grid_points": n, "pairs": n * (n - 1) // 2, "energy_vxc_ms_per_call": ms_e,
                "gradient_ms_per_call": ms_vg, "pairs_per_s": n * (n - 1) / 2 / (1e-3 * ms_e), "energy": e,
                "gradient_absmax": float(np.abs(gv).max())}
    return grad_row, vv10_row


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
    ap.add_argument("--batch-mb", type=int, default=16384)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cosx", action="store_true", help="skip the COSX next-row measurement at N = 1")
    ap.add_argument("--no-configs", action="store_true", help="skip the block over all five BASELINE configurations")
    ap.add_argument("--density-form", default="mo", choices=["mo", "d"],
                    help="mo: factorised density where cheaper (occ_xc_build_mo, needs the occupied orbitals); d: D only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
