"""Generates tests/golden/*.npz from the CPU oracle (oracle/reach_oracle.py).

The reference (pure Julia, LazySets) cannot run in this container or on the GPU box, so these are NOT outputs of
the reference itself: they freeze the pinned oracle on the reference's own test systems so that (a) the oracle
cannot drift silently and (b) the GPU box has byte-identical inputs/expected outputs without /root/reference.
Systems: the 4x4 LCS of /root/reference/test/Reachability/solve_continuous.jl:6-12, its affine CLCCS variant
(:140-162, B 4x2, U = Interval x Interval) and the 5x5 docstring example of BFFPSV18.jl:187-197.

Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import reach_oracle as orc  # noqa: E402

OUT = Path(__file__).resolve().parent

A4 = np.array([[0.0509836, 0.168159, 0.95246, 0.33644],
               [0.42377, 0.67972, 0.129232, 0.126662],
               [0.518654, 0.981313, 0.489854, 0.588326],
               [0.38318, 0.616014, 0.518412, 0.778765]])
# BFFPSV18.jl:187-191 docstring matrix
A5 = np.array([[-1., -4, 0, 0, 0], [4, -1, 0, 0, 0], [0, 0, -3, 1, 0], [0, 0, -1, 3, 0], [0, 0, 0, 0, -2]])


def pack_zonos(Z):
    p = np.array([z.ngens for z in Z], dtype=np.int32)
    n = len(Z[0].c)
    C = np.stack([z.c for z in Z], axis=1)
    G = np.zeros((n, int(p.sum())))
    o = 0
    for z in Z:
        G[:, o:o + z.ngens] = z.G
        o += z.ngens
    return p, C, G


def case(name, A, X0, delta, T, B=None, U=None, max_order=10):
    out = {"A": A, "c": X0.c, "r": X0.r, "delta": delta, "T": T}
    if B is not None:
        out.update(B=B, cU=U.c, rU=U.r)
    for alg in ("forward", "backward"):
        D = orc.discretize_box(A, X0, delta, B, U, algorithm=alg)
        N = orc.n_steps_bffpsv18(T, delta)
        C, R = orc.bffpsv18_reach(D, N)
        out.update({f"{alg}_Phi": D.Phi, f"{alg}_Phi2abs": D.Phi2abs, f"{alg}_c0": D.c0, f"{alg}_r0": D.r0,
                    f"{alg}_rE": D.rE, f"{alg}_centers": C, f"{alg}_radii": R})
        if B is not None:
            out.update({f"{alg}_rpsi": D.rpsi, f"{alg}_MU": D.MU})
    Dz = orc.discretize_zonotope(A, X0, delta, B, U)
    Nz = orc.n_steps_glgm06(T, delta)
    Z, sels = orc.glgm06_reach(Dz, Nz, max_order=max_order, return_selections=True)
    p, C, G = pack_zonos(Z)
    out.update(z_Omega0_c=Dz.Omega0.c, z_Omega0_G=Dz.Omega0.G, z_p=p, z_C=C, z_G=G, z_max_order=max_order)
    if Dz.V is not None:
        out.update(z_V_c=Dz.V.c, z_V_G=Dz.V.G)
    np.savez_compressed(OUT / f"{name}.npz", **out)
    print(name, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    case("lcs4", A4, orc.Box(np.ones(4), np.full(4, 0.1)), 0.01, 0.1)
    Bm = np.array([[0.866688, 0.771231], [0.065935, 0.349839], [0.109643, 0.904222], [0.292474, 0.227857]])  # :146-149
    case("clccs4", A4, orc.Box(np.ones(4), np.full(4, 0.1)), 0.01, 0.1, Bm,
         orc.Box(np.array([1.0, 1.0]), np.array([0.01, 0.01])), max_order=2)    # U = [0.99,1.01]^2, :152
    case("doc5", A5, orc.Box(np.ones(5), np.full(5, 0.1)), 0.04, 5.0, max_order=3)
