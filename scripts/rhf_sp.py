#!/usr/bin/env python
"""Tiny restricted Hartree-Fock for s/p Gaussian basis sets (McMurchie-Davidson integrals, pure Python/numpy).

Purpose: reproduce the fixture of the reference's tests/isolated_xc_tests.cpp:17-69 — the converged RHF/3-21G
(spherical) density of water — so that the four PySCF exchange-correlation energies that file pins
(Slater+VWN3, PBE, B88+LYP, r2SCAN; :71-294) can be checked against the CPU oracle and the CUDA build.
Running this file writes tests/golden/water_321g_rhf.npz (density matrix in the library's basis-function order).
"""
import os
import sys

import numpy as np
from scipy.special import hyp1f1

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def boys(n, x):
    return hyp1f1(n + 0.5, n + 1.5, -x) / (2.0 * n + 1.0)


def E(i, j, t, Qx, a, b):
    """Hermite expansion coefficient of the product of two 1-D Gaussians"""
    p = a + b
    q = a * b / p
    if t < 0 or t > i + j:
        return 0.0
    if i == j == t == 0:
        return np.exp(-q * Qx * Qx)
    if j == 0:
        return (1 / (2 * p)) * E(i - 1, j, t - 1, Qx, a, b) - (q * Qx / a) * E(i - 1, j, t, Qx, a, b) + \
            (t + 1) * E(i - 1, j, t + 1, Qx, a, b)
    return (1 / (2 * p)) * E(i, j - 1, t - 1, Qx, a, b) + (q * Qx / b) * E(i, j - 1, t, Qx, a, b) + \
        (t + 1) * E(i, j - 1, t + 1, Qx, a, b)


def R(t, u, v, n, p, PCx, PCy, PCz, RPC):
    T = p * RPC * RPC
    val = 0.0
    if t == u == v == 0:
        return (-2 * p) ** n * boys(n, T)
    if t == u == 0:
        if v > 1:
            val += (v - 1) * R(t, u, v - 2, n + 1, p, PCx, PCy, PCz, RPC)
        return val + PCz * R(t, u, v - 1, n + 1, p, PCx, PCy, PCz, RPC)
    if t == 0:
        if u > 1:
            val += (u - 1) * R(t, u - 2, v, n + 1, p, PCx, PCy, PCz, RPC)
        return val + PCy * R(t, u - 1, v, n + 1, p, PCx, PCy, PCz, RPC)
    if t > 1:
        val += (t - 1) * R(t - 2, u, v, n + 1, p, PCx, PCy, PCz, RPC)
    return val + PCx * R(t - 1, u, v, n + 1, p, PCx, PCy, PCz, RPC)


class BF:
    """one contracted Cartesian function: sum_k c_k x^l y^m z^n exp(-a_k r^2) (coefficients include every norm)"""

    def __init__(self, origin, lmn, exps, coefs):
        self.o, self.lmn, self.a, self.c = np.asarray(origin, float), lmn, np.asarray(exps, float), np.asarray(coefs, float)


def _prim_overlap(a, l1, A, b, l2, B):
    p = a + b
    return (E(l1[0], l2[0], 0, A[0] - B[0], a, b) * E(l1[1], l2[1], 0, A[1] - B[1], a, b) *
            E(l1[2], l2[2], 0, A[2] - B[2], a, b) * (np.pi / p) ** 1.5)


def _prim_kinetic(a, l1, A, b, l2, B):
    l2 = tuple(l2)
    t0 = b * (2 * sum(l2) + 3) * _prim_overlap(a, l1, A, b, l2, B)
    t1 = -2 * b * b * sum(_prim_overlap(a, l1, A, b, tuple(l2[k] + 2 * (k == d) for k in range(3)), B) for d in range(3))
    t2 = -0.5 * sum(l2[d] * (l2[d] - 1) * _prim_overlap(a, l1, A, b, tuple(l2[k] - 2 * (k == d) for k in range(3)), B)
                    for d in range(3) if l2[d] >= 2)
    return t0 + t1 + t2


def _prim_nuclear(a, l1, A, b, l2, B, C):
    p = a + b
    P = (a * A + b * B) / p
    RPC = np.linalg.norm(P - C)
    val = 0.0
    for t in range(l1[0] + l2[0] + 1):
        for u in range(l1[1] + l2[1] + 1):
            for v in range(l1[2] + l2[2] + 1):
                val += (E(l1[0], l2[0], t, A[0] - B[0], a, b) * E(l1[1], l2[1], u, A[1] - B[1], a, b) *
                        E(l1[2], l2[2], v, A[2] - B[2], a, b) * R(t, u, v, 0, p, P[0] - C[0], P[1] - C[1], P[2] - C[2], RPC))
    return val * 2 * np.pi / p


def _prim_eri(a, l1, A, b, l2, B, c, l3, C, d, l4, D):
    p, q = a + b, c + d
    alpha = p * q / (p + q)
    P, Q = (a * A + b * B) / p, (c * C + d * D) / q
    RPQ = np.linalg.norm(P - Q)
    Eab = [[E(l1[k], l2[k], t, A[k] - B[k], a, b) for t in range(l1[k] + l2[k] + 1)] for k in range(3)]
    Ecd = [[E(l3[k], l4[k], t, C[k] - D[k], c, d) for t in range(l3[k] + l4[k] + 1)] for k in range(3)]
    val = 0.0
    for t, et in enumerate(Eab[0]):
        for u, eu in enumerate(Eab[1]):
            for v, ev in enumerate(Eab[2]):
                for tau, ft in enumerate(Ecd[0]):
                    for nu, fu in enumerate(Ecd[1]):
                        for phi, fv in enumerate(Ecd[2]):
                            val += (et * eu * ev * ft * fu * fv * (-1) ** (tau + nu + phi) *
                                    R(t + tau, u + nu, v + phi, 0, alpha, P[0] - Q[0], P[1] - Q[1], P[2] - Q[2], RPQ))
    return val * 2 * np.pi ** 2.5 / (p * q * np.sqrt(p + q))


def contracted(fn, *bfs, extra=()):
    import itertools
    val = 0.0
    for idx in itertools.product(*[range(len(b.a)) for b in bfs]):
        coef = np.prod([b.c[i] for b, i in zip(bfs, idx)])
        args = []
        for b, i in zip(bfs, idx):
            args += [b.a[i], b.lmn, b.o]
        val += coef * fn(*args, *extra)
    return val


def basis_functions(basis):
    """Cartesian s/p functions in shell order, x,y,z within p; coefficients = common_fac * normalised coefficients"""
    fac = {0: 0.282094791773878143, 1: 0.488602511902919921}
    out = []
    for sh in basis.shells:
        assert sh.l <= 1, "s/p only"
        lmns = [(0, 0, 0)] if sh.l == 0 else [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
        for lmn in lmns:
            out.append(BF(sh.origin, lmn, sh.exponents, fac[sh.l] * sh.coefficients))
    return out


def rhf(basis, Z, pos, verbose=False):
    bfs = basis_functions(basis)
    n = len(bfs)
    S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            S[i, j] = S[j, i] = contracted(_prim_overlap, bfs[i], bfs[j])
            T[i, j] = T[j, i] = contracted(_prim_kinetic, bfs[i], bfs[j])
            V[i, j] = V[j, i] = sum(-z * contracted(_prim_nuclear, bfs[i], bfs[j], extra=(c,)) for z, c in zip(Z, pos))
    assert np.allclose(np.diag(S), 1.0, atol=1e-10), np.diag(S)
    eri = np.zeros((n, n, n, n))
    for i in range(n):
        for j in range(i + 1):
            for k in range(n):
                for l in range(k + 1):
                    if i * (i + 1) // 2 + j < k * (k + 1) // 2 + l:
                        continue
                    v = contracted(_prim_eri, bfs[i], bfs[j], bfs[k], bfs[l])
                    for a, b, c, d in ((i, j, k, l), (j, i, k, l), (i, j, l, k), (j, i, l, k), (k, l, i, j), (l, k, i, j),
                                       (k, l, j, i), (l, k, j, i)):
                        eri[a, b, c, d] = v
    enuc = sum(Z[a] * Z[b] / np.linalg.norm(pos[a] - pos[b]) for a in range(len(Z)) for b in range(a))
    H = T + V
    nocc = int(sum(Z)) // 2
    s, U = np.linalg.eigh(S)
    X = U @ np.diag(s ** -0.5) @ U.T
    P = np.zeros((n, n))
    e_old, Fs, es = 0.0, [], []
    for it in range(200):
        J = np.einsum("pqrs,rs->pq", eri, P)
        K = np.einsum("prqs,rs->pq", eri, P)
        F = H + J - 0.5 * K
        e = 0.5 * np.sum(P * (H + F)) + enuc
        err = X.T @ (F @ P @ S - S @ P @ F) @ X
        Fs.append(F)
        es.append(err)
        Fs, es = Fs[-8:], es[-8:]
        if len(Fs) > 1:   # DIIS
            m = len(Fs)
            B = -np.ones((m + 1, m + 1))
            B[m, m] = 0.0
            for a in range(m):
                for b in range(m):
                    B[a, b] = np.sum(es[a] * es[b])
            rhs = np.zeros(m + 1)
            rhs[m] = -1.0
            try:
                c = np.linalg.solve(B, rhs)[:m]
                F = sum(ci * Fi for ci, Fi in zip(c, Fs))
            except np.linalg.LinAlgError:
                pass
        eps, C = np.linalg.eigh(X.T @ F @ X)
        C = X @ C
        Pn = 2.0 * C[:, :nocc] @ C[:, :nocc].T
        dP = np.abs(Pn - P).max()
        P = Pn
        if verbose:
            print(it, e, dP, np.abs(err).max())
        if dP < 1e-11 and abs(e - e_old) < 1e-12:
            break
        e_old = e
    return {"P": P, "energy": e, "S": S, "iterations": it + 1}


def main():
    from occ_b200 import aobasis, molecules
    Z, pos = molecules.water_isolated_xc()
    basis = aobasis.AOBasis.load(Z, pos, "3-21g", pure=True)
    r = rhf(basis, Z, pos, verbose=True)
    # Cartesian p order (x, y, z) -> the library's spherical order for l = 1 (m = -1, 0, 1) = (y, z, x)
    perm, col = [], 0
    for sh in basis.shells:
        if sh.l == 0:
            perm.append(col)
            col += 1
        else:
            perm += [col + 1, col + 2, col]
            col += 3
    P = r["P"][np.ix_(perm, perm)]
    print("E(RHF/3-21G) =", r["energy"], "iterations", r["iterations"], "N_el =", np.sum(r["P"] * r["S"]))
    out = os.path.join(ROOT, "tests", "golden", "water_321g_rhf.npz")
    np.savez_compressed(out, P=P, energy=r["energy"], S=r["S"][np.ix_(perm, perm)])
    print("wrote", out)


if __name__ == "__main__":
    main()
