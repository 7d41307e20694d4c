#!/usr/bin/env python
"""Generate the committed data fixtures from the reference tree (run in the build container only).

  * occ_b200/data/basis/*.json   — subset (Z <= 18) of the reference's share/basis/*.json (MolSSI BSE
                                   schema, public basis-set data) for the basis sets the configs use;
  * occ_b200/data/lebedev_tables.npz and tests/golden/lebedev_tables.npz
                                 — Lebedev-Laikov tables produced by the reference's own
                                   src/numint/lebedev.cpp compiled into oracle/_ref/libocc_ref.so.

/root/reference does not exist on the GPU box, which is why these are committed.
"""
import ctypes
import json
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OCC_REFERENCE", "/root/reference")

BASIS_SETS = ["3-21g", "def2-svp", "def2-tzvp", "def2-qzvp"]
MAX_Z = 18
LEBEDEV_ORDERS = [6, 14, 26, 38, 50, 74, 86, 110, 146, 170, 194, 230, 266, 302, 350, 434, 590, 770, 974]


def main():
    out_basis = os.path.join(ROOT, "occ_b200", "data", "basis")
    os.makedirs(out_basis, exist_ok=True)
    for name in BASIS_SETS:
        src = os.path.join(REF, "share", "basis", name + ".json")
        j = json.load(open(src))
        sub = {
            "molssi_bse_schema": j.get("molssi_bse_schema"),
            "name": j.get("name"),
            "description": j.get("description"),
            "source": "subset (Z<=%d) of peterspackman/occ share/basis/%s.json" % (MAX_Z, name),
            "elements": {},
        }
        for z, el in j["elements"].items():
            if int(z) <= MAX_Z:
                sub["elements"][z] = {"electron_shells": [
                    {k: s[k] for k in ("angular_momentum", "exponents", "coefficients")}
                    for s in el["electron_shells"]]}
        with open(os.path.join(out_basis, name + ".json"), "w") as f:
            json.dump(sub, f, separators=(",", ":"))
        print("basis", name, len(sub["elements"]), "elements")

    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libocc_ref.so"))
    lib.occ_ref_lebedev.argtypes = [ctypes.c_int, ctypes.c_void_p]
    tables = {}
    for n in LEBEDEV_ORDERS:
        a = np.zeros((n, 4))
        lib.occ_ref_lebedev(n, a.ctypes.data)
        assert abs(a[:, 3].sum() - 1.0) < 1e-13, n
        assert np.allclose((a[:, :3] ** 2).sum(1), 1.0, atol=1e-14), n
        tables["n%d" % n] = a
    p = os.path.join(ROOT, "occ_b200", "data", "lebedev_tables.npz")
    np.savez_compressed(p, **tables)
    shutil.copy(p, os.path.join(ROOT, "tests", "golden", "lebedev_tables.npz"))
    print("lebedev", sorted(int(k[1:]) for k in tables), os.path.getsize(p), "bytes")


if __name__ == "__main__":
    sys.exit(main())
