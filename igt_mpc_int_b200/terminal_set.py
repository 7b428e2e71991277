"""Terminal invariant set C_inf of the (v, a) double integrator (host side, numpy only).

The reference builds it in MPC_Planner.__init__ (mpc.py:88-104) with the `polytope` package:
Cinf(A, B, X, U) of common/utils.py:588-627 iterates  Omega <- Pre(Omega) ∩ Omega  until the
polytope stops changing, where Pre(Omega) = {x : exists u in U, A x + B u in Omega}.  For this
2-D system with a 1-D input, Pre(Omega) is the pre-image under A of Omega swept along the
a-axis by +-da_max, so the whole iteration is plain convex-polygon arithmetic: sweep the
vertices, take the hull, map the vertices by A^-1, clip against Omega.

System (mpc.py:88-103):  [v; a]+ = [[1, dt], [0, 1]] [v; a] + [0; 1] da,  |da| <= dt * jerk_limit,
X = {-1 <= v <= 5, -4 <= a <= 3}.  Result: unit-normal rows A [v; a] <= b (74 rows for the
reference constants).
"""
import numpy as np


def _hull(points):
    """Andrew monotone chain; returns CCW vertices without collinear points.  Coordinates are
    snapped to 1e-12 first so that round-off cannot split a vertex or an axis-aligned edge."""
    pts = sorted(set((round(float(x), 12) + 0.0, round(float(y), 12) + 0.0) for x, y in points))
    if len(pts) <= 2:
        return np.array(pts)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and cross(lower[-2], lower[-1], p) <= 1e-13:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and cross(upper[-2], upper[-1], p) <= 1e-13:
            upper.pop()
        upper.append(p)
    return np.array(lower[:-1] + upper[:-1])


def _clip(poly, n, b, eps=1e-12):
    """Sutherland-Hodgman clip of a CCW convex polygon against the half-plane n.x <= b.
    Points within eps of the line count as inside (edges of Pre(Omega) often coincide with
    edges of Omega)."""
    out = []
    m = len(poly)
    for i in range(m):
        p, q = poly[i], poly[(i + 1) % m]
        dp, dq = n @ p - b, n @ q - b
        if dp <= eps:
            out.append(p)
        if (dp < -eps and dq > eps) or (dq < -eps and dp > eps):
            t = dp / (dp - dq)
            out.append(p + t * (q - p))
    return np.array(out)


def _dedupe(poly, tol=1e-10):
    out = []
    for p in poly:
        if not out or np.linalg.norm(p - out[-1]) > tol:
            out.append(p)
    if len(out) > 1 and np.linalg.norm(out[0] - out[-1]) <= tol:
        out.pop()
    return np.array(out)


def _hrep(poly):
    """Unit outward normals and offsets of a CCW polygon."""
    m = len(poly)
    A = np.empty((m, 2))
    b = np.empty(m)
    for i in range(m):
        p, q = poly[i], poly[(i + 1) % m]
        e = q - p
        n = np.array([e[1], -e[0]])
        n /= np.linalg.norm(n)
        A[i], b[i] = n, n @ p
    return A, b


def _inside(A, b, pts, tol):
    return bool(np.all(pts @ A.T <= b[None, :] + tol))


def cinf(dt=0.1, jerk_limit=0.9, v_lo=-1.0, v_hi=5.0, a_lo=-4.0, a_hi=3.0, max_iter=500, tol=1e-7):
    """Returns (A[m,2], b[m]) with A [v; a] <= b describing C_inf (utils.py:588-597)."""
    da = dt * jerk_limit
    XA = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0]])
    Xb = np.array([v_hi, a_hi, -v_lo, -a_lo])
    omega = np.array([[v_lo, a_lo], [v_hi, a_lo], [v_hi, a_hi], [v_lo, a_hi]])
    Ainv = np.array([[1.0, -dt], [0.0, 1.0]])
    for _ in range(max_iter):
        swept = np.vstack([omega + [0.0, da], omega - [0.0, da]])        # Omega (+) (-B U), utils.py:603
        pre = _hull(swept) @ Ainv.T                                       # {x : A x in swept}
        new = _hull(pre)
        OA, Ob = _hrep(omega)
        for n, b in zip(OA, Ob):                                          # ... ∩ Omega, utils.py:592
            new = _clip(new, n, b)
        new = _hull(_dedupe(new))
        NA, Nb = _hrep(new)
        if _inside(NA, Nb, omega, tol) and _inside(OA, Ob, new, tol):     # polytope.__eq__, abs_tol 1e-7
            return NA, Nb
        omega = new
    raise RuntimeError("C_inf iteration did not converge")
