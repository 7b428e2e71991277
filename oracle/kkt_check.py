"""ORACLE (test infrastructure) -- independent optimality checks of a returned solution.

Only tests/, tools/ audits, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Nothing here shares code with the interior-point iLQR (oracle/solver.py, oracle/igt_oracle.c,
the CUDA kernels): the NLP is taken in the reduced (single-shooting) form
    min_u J(u)   s.t.  c_i(u) <= 0
with J = nlp.cost and c = nlp.inequality_rows -- the rows exactly as mpc.py:177-321 writes them
(collision in squared-distance units) evaluated on a rollout of the pinned dynamics -- and
  * kkt_residual():  derivatives by central finite differences of those two functions,
    multipliers by non-negative least squares on the near-active rows;
  * reduced_derivatives(): analytic dZ/dU from the stage Jacobians (chain rule over the
    horizon), used to hand scipy's SLSQP exact first derivatives at N = 40;
  * slsqp(): scipy.optimize.minimize(SLSQP) on that reduced NLP;
  * phase1(): minimise the sum of squared row violations (is there ANY feasible point?).
"""
import numpy as np
from scipy.optimize import least_squares, lsq_linear, minimize

from . import c_oracle, dynamics as D, nlp


def _problem(pb, i, gt=False):
    return nlp.Problem(x0=pb.x0[i], u_prev=pb.u_prev[i], curv=tuple(pb.curv[i]), obs=pb.obs[i],
                       nn_ctx=tuple(pb.nn_ctx[i]) if gt else None)


class Reduced:
    """J(u), c(u) of one problem with the rollout done by the C oracle's dynamics (pinned against the
    reference's numpy model by tests/test_oracle_dynamics.py)."""

    def __init__(self, P, prob, mlp=None):
        self.P, self.prob, self.mlp = P, prob, mlp
        self.co = c_oracle.COracle(P)
        self.N = P.N

    def rollout(self, Us):
        Us = np.asarray(Us, dtype=np.float64).reshape(-1, self.N, 2)
        n = Us.shape[0]
        return self.co.rollout(np.repeat(self.prob.x0[None], n, 0), Us, np.repeat(np.asarray(self.prob.curv)[None], n, 0))

    def J(self, Z, U):
        return nlp.cost(self.P, self.prob, Z, U, self.mlp)

    def c(self, Z, U):
        return nlp.inequality_rows(self.P, self.prob, Z, U)

    def fd(self, U, h=1e-6):
        """central differences of (J, c) w.r.t. the 2N controls: gJ[2N], Jc[m, 2N]"""
        n = 2 * self.N
        u = U.reshape(-1)
        pts = np.repeat(u[None], 2 * n, 0)
        for j in range(n):
            pts[2 * j, j] += h
            pts[2 * j + 1, j] -= h
        Zs = self.rollout(pts)
        Us = pts.reshape(-1, self.N, 2)
        Jv = np.array([self.J(Zs[i], Us[i]) for i in range(2 * n)])
        cv = np.array([self.c(Zs[i], Us[i]) for i in range(2 * n)])
        return (Jv[0::2] - Jv[1::2]) / (2 * h), ((cv[0::2] - cv[1::2]) / (2 * h)).T


def kkt_residual(P, prob, U, mlp=None, comp_tol=1e-7, h=1e-6, analytic=False, lam_floor=1e-6):
    """Independent first-order check at U[N,2].  Looks for multipliers that satisfy dual feasibility and
    complementarity by construction,
        0 <= lam_i <= comp_tol / |c_i(U)|      (so lam_i |c_i| <= comp_tol; a row with slack 1e-3 may carry 1e-4;
                                                rows whose bound is below lam_floor get lam_i = 0)
    and minimise the stationarity residual |grad J + sum lam_i grad c_i| (bounded least squares, scipy
    lsq_linear / BVLS).  Derivatives: central finite differences of nlp.cost / nlp.inequality_rows
    (`analytic`: the chain-rule Jacobians of reduced_derivatives instead).
    Returns dict(stat = |residual|_inf, comp = max lam_i |c_i|, viol = max(0, max c_i), lam_max, n_active, cost)."""
    R = Reduced(P, prob, mlp)
    Z = R.rollout(U)[0]
    c = R.c(Z, U)
    if analytic:
        _, G = reduced_derivatives(P, prob, U, R.co)
        gJ, Jc = cost_gradient(P, prob, Z, U, G, mlp), rows_jacobian(P, prob, Z, U, G)
    else:
        gJ, Jc = R.fd(U, h)
    ub = comp_tol / np.maximum(np.abs(c), 1e-12)
    keep = np.where(ub >= lam_floor)[0]
    A = Jc[keep].T
    sol = lsq_linear(A, -gJ, bounds=(np.zeros(len(keep)), ub[keep]), method="bvls", tol=1e-14, max_iter=2000)
    lam = sol.x
    r = gJ + A @ lam
    return dict(stat=float(np.max(np.abs(r))), comp=float(np.max(lam * np.abs(c[keep]))), viol=float(max(0.0, c.max())),
                n_active=int(np.sum(lam > 1e-9)), lam_max=float(lam.max()), cost=R.J(Z, U))


# ---------------------------------------------------------------------------------------------
def reduced_derivatives(P, prob, U, co=None):
    """Z[N+1,7] and G[N+1,7,N,2] = dZ_k/dU_j by the chain rule over the stage Jacobians A_k, B_k
    (the C oracle's rollout Jacobians, finite-difference checked in tests/test_oracle_dynamics.py)."""
    co = co or c_oracle.COracle(P)
    N = P.N
    Z, A, Bm = co.rollout(prob.x0[None], U[None], np.asarray(prob.curv)[None], jac=True)
    Z, A, Bm = Z[0], A[0], Bm[0]
    G = np.zeros((N + 1, 7, N, 2))
    for k in range(N):
        G[k + 1] = np.einsum("ab,bjc->ajc", A[k], G[k])
        G[k + 1, :, k, :] += Bm[k]
    return Z, G


def rows_jacobian(P, prob, Z, U, G):
    """analytic Jacobian [m, 2N] of nlp.inequality_rows w.r.t. the controls (same row order)"""
    N = P.N
    n = 2 * N
    Gf = G.reshape(N + 1, 7, n)
    dv, dey = Gf[:, D.IV, :], Gf[:, D.IEY, :]
    I = np.eye(n)
    da, ddf = I[0::2], I[1::2]
    rows = [-dv[:N], dv[:N], -da, da, -ddf, ddf]
    dda = da - np.vstack([np.zeros((1, n)), da[:-1]])
    dddf = ddf - np.vstack([np.zeros((1, n)), ddf[:-1]])
    rows += [-dda, dda, -dddf, dddf, dey, -dey]
    rows += [P.cinf_A[:, :1] * dv[N - 1][None] + P.cinf_A[:, 1:2] * da[N - 1][None]]
    dp = Z[1:, :2] - prob.obs[1:]
    rows += [-2.0 * (dp[:, 0:1] * Gf[1:, D.IX, :] + dp[:, 1:2] * Gf[1:, D.IY, :])]
    return np.vstack(rows)


def cost_gradient(P, prob, Z, U, G, mlp=None):
    N = P.N
    Gf = G.reshape(N + 1, 7, 2 * N)
    g = 2.0 * (Z[:, D.IEPSI] @ Gf[:, D.IEPSI, :] + Z[:, D.IEY] @ Gf[:, D.IEY, :]) + 2.0 * P.w_u * U.reshape(-1)
    if mlp is None:
        g -= Gf[N, D.IS, :]
    else:
        _, gv = mlp.value(Z[N, D.IS], Z[N, D.IV], prob.nn_ctx, order=1)
        g -= gv[0] * Gf[N, D.IS, :] + gv[1] * Gf[N, D.IV, :]
    return g


def slsqp(P, prob, U0, mlp=None, maxiter=500, ftol=1e-12):
    """scipy SLSQP on the reduced NLP with analytic first derivatives.  Returns (res, Z, U)."""
    co = c_oracle.COracle(P)
    N = P.N
    cache = {}

    def ev(u):
        key = u.tobytes()
        if cache.get("key") != key:
            U = u.reshape(N, 2)
            Z, G = reduced_derivatives(P, prob, U, co)
            cache.update(key=key, Z=Z, U=U.copy(), G=G)
        return cache["Z"], cache["U"], cache["G"]

    def f(u):
        Z, U, _ = ev(u)
        return nlp.cost(P, prob, Z, U, mlp)

    def fj(u):
        Z, U, G = ev(u)
        return cost_gradient(P, prob, Z, U, G, mlp)

    def g(u):
        Z, U, _ = ev(u)
        return -nlp.inequality_rows(P, prob, Z, U)

    def gj(u):
        Z, U, G = ev(u)
        return -rows_jacobian(P, prob, Z, U, G)

    res = minimize(f, np.asarray(U0, dtype=np.float64).ravel(), jac=fj, method="SLSQP",
                   constraints=[{"type": "ineq", "fun": g, "jac": gj}], options={"maxiter": maxiter, "ftol": ftol})
    Z, U, _ = ev(res.x)
    return res, Z, U


# ---------------------------------------------------------------------------------------------
def scaled_rows(P, prob, Z, U):
    """nlp.inequality_rows with the collision rows in metres (d_min - dist) so that every row is O(1)."""
    c = nlp.inequality_rows(P, prob, Z, U).copy()
    dist = np.sqrt(np.sum((Z[1:, :2] - prob.obs[1:]) ** 2, axis=1))
    c[-P.N:] = P.d_min - dist
    return c


def phase1(P, prob, starts, max_nfev=200):
    """Is there a control sequence satisfying every row?  Minimise sum max(0, c_i)^2 (rows scaled as in
    scaled_rows) with scipy least_squares (trust-region reflective, analytic Jacobian, box rows as bounds)
    from each start.  Returns (best max violation in reference units, U_best)."""
    co = c_oracle.COracle(P)
    N = P.N
    lo = np.tile([P.a_min, -P.df_max], N)
    hi = np.tile([P.a_max, P.df_max], N)

    def resid(u):
        U = u.reshape(N, 2)
        Z = co.rollout(prob.x0[None], U[None], np.asarray(prob.curv)[None])[0]
        return np.maximum(0.0, scaled_rows(P, prob, Z, U))

    def jac(u):
        U = u.reshape(N, 2)
        Z, G = reduced_derivatives(P, prob, U, co)
        Jc = rows_jacobian(P, prob, Z, U, G)
        dist = np.maximum(np.sqrt(np.sum((Z[1:, :2] - prob.obs[1:]) ** 2, axis=1)), 1e-9)
        Jc[-N:] = Jc[-N:] / (2.0 * dist[:, None])              # d(d_min - dist) = d(d_min^2 - dist^2) / (2 dist)
        c = scaled_rows(P, prob, Z, U)
        return Jc * (c > 0)[:, None]

    best = (np.inf, None)
    for U0 in starts:
        u0 = np.clip(np.asarray(U0, dtype=np.float64).ravel(), lo + 1e-9, hi - 1e-9)
        try:
            r = least_squares(resid, u0, jac=jac, bounds=(lo, hi), method="trf", max_nfev=max_nfev, xtol=1e-14,
                              ftol=1e-14, gtol=1e-14)
        except Exception:
            continue
        U = r.x.reshape(N, 2)
        Z = co.rollout(prob.x0[None], U[None], np.asarray(prob.curv)[None])[0]
        v = nlp.max_violation(P, prob, Z, U)
        if v < best[0]:
            best = (v, U)
        if v <= 1e-9:
            break
    return best
