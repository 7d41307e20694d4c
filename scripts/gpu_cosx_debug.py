#!/usr/bin/env python
"""ESP-phase time of one COSX call under the kernel's timing switches (env OCC_COSX_DEBUG: 1 = no G atomics,
2 = no primitive loop, 3 = both).  usage: gpu_cosx_debug.py [mol basis angular]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occ_b200 import aobasis, molecules, cosx

mol, bas, ang = (sys.argv[1:4] + ["valinomycin", "def2-tzvp", "50"][len(sys.argv) - 1:])[:3]
Z, pos = molecules.MOLECULES[mol]()
b = aobasis.AOBasis.load(Z, pos, bas, pure=False)
D = molecules.synthetic_density(b, 1) * 0.5
C = molecules.synthetic_orbitals(b, 1)[0]
import time
# no per-pair counters: they are two same-address atomics per evaluated pair and become the floor once the math is removed
sgx = cosx.SemiNumericalExchange(b, cosx.sgx_grid_settings(int(ang), 1e-5), collect_stats=False)
sgx.compute_K_mo(D, C)
ts = []
for _ in range(2):
    t0 = time.perf_counter(); sgx.compute_K_mo(D, C); ts.append(time.perf_counter() - t0)
print(json.dumps({"debug": os.environ.get("OCC_COSX_DEBUG", "0"), "ms_per_call": 1e3 * min(ts)}))
