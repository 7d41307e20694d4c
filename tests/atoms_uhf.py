"""TEST HELPER: unrestricted Hartree-Fock for atoms whose occupied orbitals are all of s type (He, Li, Be) in an
uncontracted even-tempered s-Gaussian basis — closed-form one-centre integrals, numpy only.  It supplies spin
densities with 0 < zeta < 1 (lithium) so that published atomic exchange / correlation energies can pin the
partially-polarized branches of the functionals, which no test of the reference covers."""
import numpy as np


def s_basis(n=20, a0=0.012, ratio=2.1):
    return a0 * ratio ** np.arange(n)


def integrals(a, Z):
    p = a[:, None] + a[None, :]
    S = (2.0 * np.sqrt(a[:, None] * a[None, :]) / p) ** 1.5
    T = 3.0 * a[:, None] * a[None, :] / p * S
    V = -Z * 2.0 * np.sqrt(p / np.pi) * S
    P = p[:, :, None, None]
    Q = p[None, None, :, :]
    eri = S[:, :, None, None] * S[None, None, :, :] * 2.0 / np.sqrt(np.pi) * np.sqrt(P * Q / (P + Q))
    return S, T + V, eri


def uhf(Z, nalpha, nbeta, a=None, iters=200):
    a = s_basis() if a is None else a
    S, H, eri = integrals(a, Z)
    w, U = np.linalg.eigh(S)
    X = U / np.sqrt(w)
    Da = Db = np.zeros_like(S)
    E_old = 0.0
    for it in range(iters):
        J = np.einsum("ijkl,kl->ij", eri, Da + Db)
        Ka = np.einsum("ikjl,kl->ij", eri, Da)
        Kb = np.einsum("ikjl,kl->ij", eri, Db)
        Fa, Fb = H + J - Ka, H + J - Kb
        E = 0.5 * (np.sum((Da + Db) * H) + np.sum(Da * Fa) + np.sum(Db * Fb))
        new = []
        for F, nocc in ((Fa, nalpha), (Fb, nbeta)):
            e, C = np.linalg.eigh(X.T @ F @ X)
            C = X @ C
            new.append(C[:, :nocc] @ C[:, :nocc].T)
        mix = 0.0 if it == 0 else 0.3
        Da, Db = (1 - mix) * new[0] + mix * Da, (1 - mix) * new[1] + mix * Db
        if it > 5 and abs(E - E_old) < 1e-11:
            break
        E_old = E
    Ex = -0.5 * (np.einsum("ij,ikjl,kl->", Da, eri, Da) + np.einsum("ij,ikjl,kl->", Db, eri, Db))
    return {"energy": E, "exchange": Ex, "a": a, "Da": Da, "Db": Db}


def radial_densities(res, r):
    """rho_s(r), |grad rho_s|^2, grad rho_a . grad rho_b, tau_s on the radii r for s = alpha, beta"""
    a = res["a"]
    N = (2.0 * a / np.pi) ** 0.75
    g = N[None, :] * np.exp(-a[None, :] * r[:, None] ** 2)
    dg = -2.0 * a[None, :] * r[:, None] * g
    out = []
    for D in (res["Da"], res["Db"]):
        rho = np.einsum("pi,ij,pj->p", g, D, g)
        drho = 2.0 * np.einsum("pi,ij,pj->p", dg, D, g)
        tau = 0.5 * np.einsum("pi,ij,pj->p", dg, D, dg)
        out.append((rho, drho, tau))
    (ra, da, ta), (rb, db, tb) = out
    return (np.stack([ra, rb], 1), np.stack([da * da, da * db, db * db], 1), np.stack([ta, tb], 1))
