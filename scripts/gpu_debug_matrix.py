#!/usr/bin/env python
"""Developer experiment: where the Vxc / density kernels lose time (debug switches of options.reserved[0])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.gpu_experiments import run
for dbg in [int(x) for x in (sys.argv[1:] or ["0", "1", "8", "16", "25", "2"])]:
    run("valinomycin", "def2-tzvp", "b3lyp", 1, 4096, debug=dbg, lanes=1, reps=1)
