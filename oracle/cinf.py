"""ORACLE (test infrastructure) -- terminal invariant set C_inf of mpc.py:88-104.

Restates common/utils.py:588-627 (`Cinf`, `precursor`, `minkowski_sum`) without the
`polytope` package (polytope==0.2.5 is pinned in common/environment.yml:83 but not
installed here): vertex enumeration + convex hull via scipy.spatial, exactly the
"precursor ∩ Omega until fixed point" iteration of utils.py:588-597.

System (mpc.py:88-103):  [v; a]+ = [[1, dt], [0, 1]] [v; a] + [0; 1] da,
X = {v <= 5, a <= 3, -v <= 1, -a <= 4},  |da| <= dt * jerk_limit = 0.09.
Result: rows (A[m,2], b[m]) with unit-norm A rows, A [v; a] <= b.
"""
import numpy as np
from scipy.spatial import ConvexHull, HalfspaceIntersection


def _hrep_from_vertices(V):
    hull = ConvexHull(V)
    eq = hull.equations            # [n0, n1, off]:  n.x + off <= 0 inside
    A = eq[:, :2].copy()
    b = -eq[:, 2].copy()
    # merge coplanar duplicates produced by qhull triangulation (none in 2-D, kept for safety)
    keep = []
    for i in range(len(b)):
        dup = False
        for j in keep:
            if abs(A[i] @ A[j] - 1.0) < 1e-12 and abs(b[i] - b[j]) < 1e-12:
                dup = True
                break
        if not dup:
            keep.append(i)
    return A[keep], b[keep]


def _vertices_from_hrep(A, b):
    # interior point by Chebyshev-like averaging: solve small LP via scipy
    from scipy.optimize import linprog
    nrm = np.linalg.norm(A, axis=1)
    res = linprog(c=[0, 0, -1], A_ub=np.hstack([A, nrm[:, None]]), b_ub=b,
                  bounds=[(None, None), (None, None), (0, None)])
    xc = res.x[:2]
    hs = HalfspaceIntersection(np.hstack([A, -b[:, None]]), xc)
    V = hs.intersections
    hull = ConvexHull(V)
    return V[hull.vertices]


def _reduce(A, b, tol=1e-9):
    """Drop redundant rows (those not touching the polygon)."""
    V = _vertices_from_hrep(A, b)
    keep = []
    for i in range(len(b)):
        touching = np.sum(np.abs(V @ A[i] - b[i]) < 1e-7)
        if touching >= 2:
            keep.append(i)
    A2, b2 = A[keep], b[keep]
    # de-duplicate parallel identical rows
    out = []
    for i in range(len(b2)):
        if not any(np.linalg.norm(A2[i] - A2[j]) < tol and abs(b2[i] - b2[j]) < tol for j in out):
            out.append(i)
    return A2[out], b2[out]


def _normalize(A, b):
    n = np.linalg.norm(A, axis=1)
    return A / n[:, None], b / n


def _contains(A, b, V, tol=1e-7):
    return np.all(V @ A.T <= b[None, :] + tol)


def cinf_vertices_hrep(dt=0.1, jerk_limit=0.9, v_hi=5.0, a_hi=3.0, v_lo=-1.0, a_lo=-4.0,
                       max_iter=200):
    """Returns (A[m,2], b[m], n_iterations)."""
    Asys = np.array([[1.0, dt], [0.0, 1.0]])
    Bsys = np.array([[0.0], [1.0]])
    XA = np.array([[1.0, 0], [0, 1.0], [-1.0, 0], [0, -1.0]])
    Xb = np.array([v_hi, a_hi, -v_lo, -a_lo])
    Uv = np.array([[dt * jerk_limit], [-dt * jerk_limit]])       # extreme(U)
    OA, Ob = XA.copy(), Xb.copy()
    for it in range(max_iter):
        # precursor: minkowski_sum(Omega, extreme(U) @ -B.T) then pre-image under A
        V1 = _vertices_from_hrep(OA, Ob)
        V2 = Uv @ (-Bsys.T)
        Vs = np.array([p + q for p in V1 for q in V2])
        TA, Tb = _hrep_from_vertices(Vs)
        PA, Pb = _normalize(TA @ Asys, Tb)
        # intersect with Omega
        NA, Nb = _reduce(np.vstack([PA, OA]), np.concatenate([Pb, Ob]))
        # fixed point test (polytope.__eq__: mutual containment, abs_tol 1e-7)
        Vn = _vertices_from_hrep(NA, Nb)
        if _contains(NA, Nb, V1) and _contains(OA, Ob, Vn):
            return NA, Nb, it
        OA, Ob = NA, Nb
    raise RuntimeError("C_inf iteration did not reach a fixed point")
