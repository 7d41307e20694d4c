#!/usr/bin/env python
"""Phase times of every BASELINE config with the current kernels (developer experiment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.gpu_experiments import run
run("water", "def2-svp", "b3lyp", 1, 4096, lanes=1)
run("benzene", "def2-tzvp", "pbe", 1, 4096, lanes=1)
run("h2o_32", "def2-svp", "b3lyp", 2, 4096, lanes=1)
run("valinomycin", "def2-tzvp", "b3lyp", 1, 4096, lanes=1)
run("taxol", "def2-tzvp", "r2scan", 1, 4096, lanes=1)
