#!/usr/bin/env python
"""Developer check (GPU box): CUDA build vs the CPU oracle on small cases, with per-phase times."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, dft
from oracle import oracle as O

def case(mol, bas, method, nspin, pure=False, batch_bytes=0, reps=3):
    Z, pos = molecules.MOLECULES[mol]()
    b = aobasis.AOBasis.load(Z, pos, bas, pure=pure)
    ob = O.Basis(b.as_arrays())
    pts, w, _ = O.molecular_grid(Z, pos, ob)
    spts, sw, perm, leaves = O.hierarchy(pts, w)
    D2 = molecules.synthetic_density(b, nspin)
    t0 = time.time()
    ref = O.xc_build(ob, spts, sw, leaves, [k for k, v in O.FUNC_IDS.items() if v in dft.get_dft_method(method)] if False else
                     [n for i in dft.get_dft_method(method) for n, v in O.FUNC_IDS.items() if v == i], nspin, D2)
    t_cpu = time.time() - t0
    d = dft.DFT(method, b, dft.GridHierarchy(spts, sw, leaves), dft.UNRESTRICTED if nspin == 2 else dft.RESTRICTED,
                batch_bytes=batch_bytes)
    K = d.build(D2)
    d.set_profiling(True)
    K = d.build(D2)
    ph = d.phase_ms()
    d.set_profiling(False)
    ts = []
    for _ in range(reps):
        t0 = time.time(); K = d.build(D2); ts.append(time.time() - t0)
    info = d.info()
    dK = np.abs(K - ref["K"]).max()
    print(f"{mol:12s} {bas:10s} {method:7s} nspin={nspin} pure={int(pure)} nbf={b.nbf} npts={len(sw)} "
          f"dK={dK:.2e} dExc={abs(d.m_exc_dft - ref['exc']):.2e} dNa={abs(d.total_density[0]-ref['n_alpha']):.2e} "
          f"dNb={abs(d.total_density[1]-ref['n_beta']):.2e} |K|max={np.abs(ref['K']).max():.3f} Exc={ref['exc']:.8f} "
          f"gpu={min(ts)*1e3:.2f}ms cpu={t_cpu*1e3:.0f}ms batches={info['nbatches']} "
          f"TF={info['flops_per_build']/min(ts)/1e12:.2f}", flush=True)
    print("    phases(ms):", {k: round(v, 3) for k, v in ph.items()}, "skipped:", ref["stats"]["nskipped"], flush=True)
    d.close()
    return dK

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "small"
    case("water", "def2-svp", "svwn", 1)
    case("water", "def2-svp", "pbe", 1)
    case("water", "def2-svp", "b3lyp", 1)
    case("water", "def2-svp", "pbe", 1, pure=True)
    case("water", "def2-svp", "pbe", 2)
    case("water", "def2-svp", "svwn5", 2)
    case("water", "def2-svp", "r2scan", 1)
    case("water", "def2-svp", "r2scan", 2)
    case("benzene", "def2-tzvp", "pbe", 1)
    case("benzene", "def2-tzvp", "pbe", 1, batch_bytes=32 << 20)
    case("benzene", "def2-tzvp", "pbe", 1, batch_bytes=2 << 30)
    case("benzene", "def2-tzvp", "b3lyp", 2, pure=True)
    case("benzene", "def2-tzvp", "r2scan", 1)
    if which == "big":
        case("h2o_32", "def2-svp", "b3lyp", 2)
