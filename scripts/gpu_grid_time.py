import time, sys
sys.path.insert(0, '.')
from occ_b200 import aobasis, grid, molecules
Z, pos = molecules.MOLECULES["valinomycin"]()
b = aobasis.AOBasis.load(Z, pos, "def2-tzvp")
for i in range(2):
    t0 = time.perf_counter(); mg = grid.MolecularGrid(b); t1 = time.perf_counter(); mg.close()
    print("valinomycin grid create s:", round(t1 - t0, 3))
