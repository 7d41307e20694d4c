#!/usr/bin/env python
"""COSX timing on the GPU box: SemiNumericalExchange.compute_K per call, phase split, work counters,
and the CPU oracle on a bounded sample of the same batches.  usage: gpu_cosx_bench.py [mol basis pure angular [cpu_batches]]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, cosx


def run(mol, bas, pure, angular, cpu_batches=0, reps=3, radial_precision=1e-5):
    lib_k = os.environ.get("COSX_KGEMM_LIBRARY", "0") == "1"   # A/B: K' += Phi^T G through cuBLAS per batch
    Z, pos = molecules.MOLECULES[mol]()
    b = aobasis.AOBasis.load(Z, pos, bas, pure=pure)
    D = molecules.synthetic_density(b, 1) * 0.5
    t0 = time.perf_counter()
    sgx = cosx.SemiNumericalExchange(b, cosx.sgx_grid_settings(angular, radial_precision), collect_stats=True,
                                     )
    t_setup = time.perf_counter() - t0
    K = sgx.compute_K(D)
    info = sgx.info()
    Cocc = molecules.synthetic_orbitals(b, 1)[0]          # 2 D = Cocc Cocc^T (the orbitals behind synthetic_density)
    Kmo = sgx.compute_K_mo(D, Cocc)
    info_mo = sgx.info()
    sgx.close()
    # timed handle without the per-pair counters (one atomic per evaluated pair) and without per-chunk event syncs
    sgx = cosx.SemiNumericalExchange(b, atom_grids=(sgx.atom_offsets, sgx.points, sgx.weights),
                                     projector=sgx.overlap_projector, )
    sgx.compute_K(D)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); K2 = sgx.compute_K(D); ts.append(time.perf_counter() - t0)
    sgx.compute_K_mo(D, Cocc)
    ts_mo = []
    for _ in range(reps):
        t0 = time.perf_counter(); sgx.compute_K_mo(D, Cocc); ts_mo.append(time.perf_counter() - t0)
    n = sgx.num_grid_points()
    out = {"row": "f-2 COSX compute_K", "workload": f"{mol}/{bas} {'spherical' if pure else 'cartesian'} RKS, sgx grid {angular} angular, radial precision {radial_precision:g}",
           "nbf": b.nbf, "grid_points": n, "batches": info["nbatches"], "shell_pairs": info["npairs"],
           "primitive_pairs": info["nprimitive_pairs"], "pair_evals": info["pair_evals"],
           "ms_per_call": 1e3 * min(ts), "grid_points_per_s": n / min(ts),
           "phase_ms_profiled_call": {k: info[k] for k in ("ms_collocate", "ms_fgemm", "ms_esp", "ms_kgemm")},
           "dense_equivalent_esp_tflops": info["dense_flops"] / min(ts) / 1e12, "dense_esp_flops": info["dense_flops"],
           "gemm_flops": info["gemm_flops"], "launches": info["launches"], "gemm_calls": info["gemm_calls"],
           "kgemm": "cublas per batch" if lib_k else "k_cosx_kgemm", "kgemm_blocks_executed": info["kgemm_blocks"],
           "kgemm_blocks_total": info["nbatches"] * ((b.nbf + 63) // 64),
           "mo_form": {"entry": "occ_cosx_compute_K_mo (F = (Phi.(P C)).C^T)", "nocc": int(Cocc.shape[1]),
                       "ms_per_call": 1e3 * min(ts_mo), "grid_points_per_s": n / min(ts_mo),
                       "phase_ms_profiled_call": {k: info_mo[k] for k in ("ms_collocate", "ms_fgemm", "ms_esp", "ms_kgemm")},
                       "gemm_flops": info_mo["gemm_flops"], "max_abs_diff_vs_D_form": float(np.abs(Kmo - K).max())},
           "setup_s": t_setup, "device_bytes": info["device_bytes"], "K_absmax": float(np.abs(K).max()),
           "repeatability": float(np.abs(K - K2).max())}
    if cpu_batches:
        from oracle import oracle as O
        ob = O.Basis(b.as_arrays())
        offs, pts, w = sgx.atom_offsets, sgx.points, sgx.weights
        # every s-th batch (a block of 128 points of one atom grid) as its own "atom grid"
        starts = [(int(offs[a]) + s, min(128, int(offs[a + 1]) - int(offs[a]) - s)) for a in range(len(offs) - 1)
                  for s in range(0, int(offs[a + 1] - offs[a]), 128)]
        stride = max(1, len(starts) // cpu_batches)
        sel = starts[::stride]
        spts = np.concatenate([pts[s:s + c] for s, c in sel])
        sw = np.concatenate([w[s:s + c] for s, c in sel])
        soff = np.concatenate([[0], np.cumsum([c for _, c in sel])]).astype(np.uint64)
        D2q = sgx.overlap_projector @ (2.0 * D)
        t0 = time.perf_counter(); Kc, st = O.cosx_K_for_density(ob, soff, spts, sw, D2q); tc = time.perf_counter() - t0
        out["cpu_oracle"] = {"sample": f"every {stride}-th batch ({len(sel)} of {len(starts)} batches, {len(sw)} points)",
                             "seconds": tc, "grid_points_per_s": len(sw) / tc, "cores": os.cpu_count(),
                             "note": "includes the oracle's per-call shell-pair setup"}
        out["speedup_vs_cpu_oracle"] = out["grid_points_per_s"] / out["cpu_oracle"]["grid_points_per_s"]
    print(json.dumps(out), flush=True)
    sgx.close()


if __name__ == "__main__":
    a = sys.argv[1:]
    if a:
        run(a[0], a[1], a[2] == "1", int(a[3]), int(a[4]) if len(a) > 4 else 0, reps=int(a[5]) if len(a) > 5 else 3)
    else:
        run("benzene", "def2-tzvp", True, 50, cpu_batches=40)
        run("h2o_32", "def2-svp", False, 50, cpu_batches=40)
