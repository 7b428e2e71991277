"""ORACLE (test infrastructure) -- groundwork for the OBCA collision mode (not yet in the solver).

The reference's `ca_type='obca'` (mpc.py:170-175, :211-221) adds, per step k >= 1 and obstacle, dual
variables lambda, mu in R^4 with
    -g' mu + (A p - b)' lambda >= d_min + 1e-6,   G' mu + R(psi)' A' lambda = 0,   |A' lambda|^2 <= 1,   lambda, mu >= 0
where {y: A y <= b} is the obstacle rectangle (utils.rotation_translation of its forecast pose), {y: G y <= g} the
ego rectangle in its own frame, p = (x, y) and R(psi) the ego pose.  (lambda, mu) appear in no other row and in
no cost term, so they can be maximised out stage by stage: the largest value of the first row's left-hand side
over the other three is the Euclidean distance between the two rectangles (0 if they intersect) -- strong
duality of the distance problem between two polytopes (Zhang, Liniger, Borrelli, "Optimization-based collision
avoidance").  The five OBCA rows of a stage are therefore equivalent to ONE row

    d_min + 1e-6 - dist(ego rectangle(x, y, psi), obstacle rectangle_k) <= 0

with gradient (n, n . (dR/dpsi a_local)) from the closest-point pair, which is what a shooting solver with
stage-local rows can use directly.  This module states both sides -- `dual_value` solves the reference's dual
problem with scipy, `rect_distance` is the closed form with its gradient -- and tests/test_oracle_obca.py checks
that they agree.
"""
import numpy as np
from scipy.optimize import minimize

VEH_LENGTH, VEH_WIDTH = 4.47, 2.0           # mpc.py:48-51


def rotation_translation(x0, theta, h=VEH_LENGTH, w=VEH_WIDTH):
    """H-representation {y: A y <= b} of a rectangle of length h, width w centred at x0 with heading theta
    (common/utils.py:412-442, numpy branch)."""
    R = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
    A = np.vstack([R.T, -R.T])
    b = np.array([h / 2, w / 2, h / 2, w / 2]) + A @ np.asarray(x0, dtype=float)
    return A, b


def dual_value(ego, obs, h=VEH_LENGTH, w=VEH_WIDTH, return_arg=False):
    """max of -g' mu + (A p - b)' lambda over the reference's dual feasible set (mpc.py:216-221), by SLSQP from
    several starts.  ego, obs = (x, y, heading)."""
    A, b = rotation_translation(obs[:2], obs[2], h, w)
    G, g = rotation_translation([0.0, 0.0], 0.0, h, w)                                   # mpc.py:106
    p = np.asarray(ego[:2], dtype=float)
    c, s = np.cos(ego[2]), np.sin(ego[2])
    R = np.array([[c, -s], [s, c]])
    r = A @ p - b

    def neg(v):
        return g @ v[4:] - r @ v[:4]

    cons = [{'type': 'eq', 'fun': lambda v: G.T @ v[4:] + R.T @ A.T @ v[:4]},
            {'type': 'ineq', 'fun': lambda v: 1.0 - np.sum((A.T @ v[:4]) ** 2)}]
    best, arg = 0.0, np.zeros(8)                                                         # lambda = mu = 0 is feasible
    rng = np.random.default_rng(0)
    for _ in range(12):
        v0 = rng.uniform(0, 0.5, 8)
        res = minimize(neg, v0, method='SLSQP', bounds=[(0, None)] * 8, constraints=cons,
                       options={'ftol': 1e-12, 'maxiter': 300})
        if res.success:
            v = res.x
            feas = np.max(np.abs(G.T @ v[4:] + R.T @ A.T @ v[:4])) < 1e-7 and np.sum((A.T @ v[:4]) ** 2) <= 1 + 1e-7
            if feas and -res.fun > best:
                best, arg = -res.fun, v
    return (best, arg) if return_arg else best


def _corners(pose, h, w):
    x, y, th = pose
    c, s = np.cos(th), np.sin(th)
    loc = np.array([[h / 2, w / 2], [-h / 2, w / 2], [-h / 2, -w / 2], [h / 2, -w / 2]])
    return np.array([x, y]) + loc @ np.array([[c, s], [-s, c]]), loc


def _point_rect(q, pose, h, w):
    """distance from point q to the rectangle `pose`, the closest point on it, and whether q is inside"""
    x, y, th = pose
    c, s = np.cos(th), np.sin(th)
    d = np.array([c * (q[0] - x) + s * (q[1] - y), -s * (q[0] - x) + c * (q[1] - y)])      # q in the rectangle's frame
    cl = np.clip(d, [-h / 2, -w / 2], [h / 2, w / 2])
    inside = bool(np.all(np.abs(d) <= np.array([h / 2, w / 2])))
    world = np.array([x, y]) + np.array([c * cl[0] - s * cl[1], s * cl[0] + c * cl[1]])
    return float(np.linalg.norm(d - cl)), world, cl, inside


def _segments_cross(P, Q):
    def cross(a, b):
        return a[0] * b[1] - a[1] * b[0]
    for i in range(4):
        a0, a1 = P[i], P[(i + 1) % 4]
        for j in range(4):
            b0, b1 = Q[j], Q[(j + 1) % 4]
            d1, d2 = a1 - a0, b1 - b0
            den = cross(d1, d2)
            if abs(den) < 1e-14:
                continue
            t = cross(b0 - a0, d2) / den
            u = cross(b0 - a0, d1) / den
            if 0 <= t <= 1 and 0 <= u <= 1:
                return True
    return False


def _penetration(ego, obs, h, w):
    """Intersecting rectangles: minus the penetration depth (separating-axis test over the four face normals)
    and its gradient w.r.t. the ego pose -- the continuation of the distance into the infeasible region, so that
    an iterate that jumped into the obstacle sees how to get out (the clamped distance is flat there)."""
    hl, hw = h / 2, w / 2
    c, s = np.cos(ego[2]), np.sin(ego[2])
    uxe, uye = np.array([c, s]), np.array([-s, c])
    co, so = np.cos(obs[2]), np.sin(obs[2])
    uxo, uyo = np.array([co, so]), np.array([-so, co])
    dvec = np.asarray(ego[:2], dtype=float) - np.asarray(obs[:2], dtype=float)
    sg = lambda v: 1.0 if v >= 0 else -1.0
    best = None
    for a, da, he in ((uxe, uye, hl), (uye, -uxe, hw)):           # ego face normals (turn with psi)
        ho = hl * abs(a @ uxo) + hw * abs(a @ uyo)
        ov = he + ho - abs(a @ dvec)
        g_p = -sg(a @ dvec) * a
        g_psi = hl * sg(a @ uxo) * (da @ uxo) + hw * sg(a @ uyo) * (da @ uyo) - sg(a @ dvec) * (da @ dvec)
        if best is None or ov < best[0]:
            best = (ov, g_p, g_psi)
    for a, ho in ((uxo, hl), (uyo, hw)):                          # obstacle face normals (fixed)
        he = hl * abs(a @ uxe) + hw * abs(a @ uye)
        ov = he + ho - abs(a @ dvec)
        g_p = -sg(a @ dvec) * a
        g_psi = hl * sg(a @ uxe) * (a @ uye) + hw * sg(a @ uye) * (a @ -uxe)
        if ov < best[0]:
            best = (ov, g_p, g_psi)
    ov, g_p, g_psi = best
    return -ov, -np.array([g_p[0], g_p[1], g_psi])


def rect_distance(ego, obs, h=VEH_LENGTH, w=VEH_WIDTH):
    """Signed distance between the two rectangles -- the Euclidean distance if they are disjoint (which is what
    the reference's dual rows certify), minus the penetration depth if they intersect -- and its gradient with
    respect to the ego pose (x, y, psi).  For disjoint convex polygons the minimum is attained between a vertex of
    one and the boundary of the other, so the eight vertex-to-rectangle distances cover it."""
    Pe, loc_e = _corners(ego, h, w)
    Po, _ = _corners(obs, h, w)
    best = (np.inf, None)
    for i in range(4):                                            # ego vertex -> obstacle rectangle
        d, cl_world, _, inside = _point_rect(Pe[i], obs, h, w)
        if inside:
            return _penetration(ego, obs, h, w)
        if d < best[0]:
            best = (d, ('ego_vertex', Pe[i], cl_world, loc_e[i]))
    for j in range(4):                                            # obstacle vertex -> ego rectangle
        d, cl_world, cl_local, inside = _point_rect(Po[j], ego, h, w)
        if inside:
            return _penetration(ego, obs, h, w)
        if d < best[0]:
            best = (d, ('obs_vertex', cl_world, Po[j], cl_local))
    if _segments_cross(Pe, Po):                                   # crossing edges without a vertex inside
        return _penetration(ego, obs, h, w)
    d, (_, a, bpt, a_local) = best
    n = (a - bpt) / d                                             # from the obstacle's closest point to the ego's
    c, s = np.cos(ego[2]), np.sin(ego[2])
    da_dpsi = np.array([-s * a_local[0] - c * a_local[1], c * a_local[0] - s * a_local[1]])   # d(R a_local)/dpsi
    return d, np.array([n[0], n[1], n @ da_dpsi])
