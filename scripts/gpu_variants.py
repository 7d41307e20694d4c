#!/usr/bin/env python
"""Developer experiment: time the DMMA kernels for each prebuilt kernel variant (OCC_XC_B200_LIB override)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys; sys.path.insert(0, %r)
from scripts.gpu_experiments import run
run("valinomycin", "def2-tzvp", "b3lyp", 1, 4096, lanes=1, reps=1)
run("valinomycin", "def2-tzvp", "b3lyp", 1, 4096, lanes=1, reps=1)
''' % ROOT
vdir = os.path.join(ROOT, "occ_b200", "variants")
libs = [None] + sorted(os.path.join(vdir, f) for f in os.listdir(vdir)) + [None]
for lib in libs:
    env = dict(os.environ)
    if lib: env["OCC_XC_B200_LIB"] = lib
    print("=== variant", lib or "default", flush=True)
    subprocess.run([sys.executable, "-c", code], env=env)
