"""ORACLE (test infrastructure) -- fp64 interior-point iLQR for the reference NLP.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

PARITY UNPINNED against IPOPT: the reference solves the NLP of mpc.py:21-160 with
CasADi Opti + IPOPT (mpc.py:129-146, :392); casadi==3.6.5 is not installed and not
installable offline, and the reference has no test that pins a solver result
(SURVEY 4, 8(c)).  This file restates the *problem* exactly (oracle/nlp.py) and solves it
to a KKT residual <= 1e-8 -- far tighter than IPOPT's tol=1e-3 (mpc.py:133-135) -- so that
any KKT point the reference could accept within its tolerance is bracketed.  The solution
is cross-checked against scipy.optimize (SLSQP / trust-constr) in
tests/test_oracle_solver.py.

Algorithm (shared, by design, with the CUDA solver so both reach the same local minimum of
this non-convex problem; DESIGN.md "solver"): single shooting on u with the previous input
appended to the state (zeta = [z; u_prev], so the rate rows of mpc.py:301-312 become
stage-local), Gauss-Newton Riccati backward pass on the perturbed KKT system of an
infeasible-start primal-dual interior point method (slack y > 0 with c + y = 0,
multiplier s > 0 with s*y = mu), closed-loop nonlinear forward pass with
fraction-to-boundary and a two-entry filter, monotone mu decrease.
"""
import numpy as np

from . import dynamics as D
from . import nlp


class Options:
    tol = 1e-6           # stationarity: |adjoint-based Lagrangian gradient|_inf <= tol * max(1, |s|_inf)
    tol_rp = 1e-8        # primal residual |c + y|_inf
    tol_comp = 1e-7      # max s*y at termination
    max_iter = 300
    mu0 = 0.3
    mu_floor = 1e-8
    kappa_eps = 10.0     # shrink mu when barrier-KKT error <= kappa_eps * mu
    kappa_mu = 0.2
    theta_mu = 1.5
    y_init_min = 0.3
    tau_min = 0.99
    reg_min = 1e-4
    reg_up = 10.0
    reg_down = 10.0
    reg_max = 1e10
    reg_jump = 1.1       # after a non-positive-definite Quu: reg >= reg_jump * (-lambda_min(Quu)) of that stage
    n_alpha = 6
    eps_phi = 1e-12
    gamma_theta = 1e-6
    theta_small = 1e-10
    second_order = True
    predict_alpha = True
    stall_iter = 16
    stall_rp = 1e-2
    alpha_safety = 0.99
    # reference-tolerance exit (status 6): no further progress at a point within the reference's IPOPT tolerances
    # (mpc.py:133-135 tol = dual_inf_tol = constr_viol_tol = 1e-3, IPOPT default compl_inf_tol 1e-4; primal bound 1e-6)
    acc_tol = 1e-3
    acc_rp = 1e-6
    acc_comp = 1e-4


# --------------------------------------------------------------------------------------
# stage-wise constraint rows  c(zeta, u) <= 0 with gradients, zeta = [z(7); up(2)]
# --------------------------------------------------------------------------------------
def stage_rows(P, prob, k, z, up, u):
    """Rows attached to stage k < N.  Returns c[m], Cx[m,9], Cu[m,2], hxy[m] where hxy is the
    coefficient h such that the row's Hessian is h*I on (x, y) (collision rows: -2)."""
    c, Cx, Cu, hxy = [], [], [], []

    def add(val, gx=None, gu=None, h=0.0):
        c.append(val)
        Cx.append(np.zeros(9) if gx is None else gx)
        Cu.append(np.zeros(2) if gu is None else gu)
        hxy.append(h)

    def ex(i, sgn=1.0):
        g = np.zeros(9)
        g[i] = sgn
        return g

    def eu(i, sgn=1.0):
        g = np.zeros(2)
        g[i] = sgn
        return g

    if k >= 1:
        # state-only rows; at k = 0 they involve x0 only and are pre-checked instead
        add(z[D.IV] - P.v_max, ex(D.IV))                     # mpc.py:317
        add(P.v_min - z[D.IV], ex(D.IV, -1.0))               # mpc.py:316
        add(z[D.IEY] - P.ey_lim, ex(D.IEY))                  # mpc.py:298
        add(-P.ey_lim - z[D.IEY], ex(D.IEY, -1.0))           # mpc.py:299
        cval, g, hc = collision_row(P, z, prob.obs[k], None if prob.obs_psi is None else prob.obs_psi[k])
        add(cval, g, None, hc)                                   # mpc.py:226 (distance form) / :211-221 (OBCA)
    add(u[0] - P.a_max, None, eu(0))                          # mpc.py:319
    add(P.a_min - u[0], None, eu(0, -1.0))                    # mpc.py:318
    add(u[1] - P.df_max, None, eu(1))                         # mpc.py:321
    add(-P.df_max - u[1], None, eu(1, -1.0))                  # mpc.py:320
    add((u[0] - up[0]) - P.da_max, ex(7, -1.0), eu(0))        # mpc.py:303-311
    add(-(u[0] - up[0]) - P.da_max, ex(7), eu(0, -1.0))
    add((u[1] - up[1]) - P.ddf_max, ex(8, -1.0), eu(1))
    add(-(u[1] - up[1]) - P.ddf_max, ex(8), eu(1, -1.0))
    if k == P.N - 1:
        for m in range(len(P.cinf_b)):                        # mpc.py:177-180
            gx = np.zeros(9)
            gx[D.IV] = P.cinf_A[m, 0]
            add(P.cinf_A[m, 0] * z[D.IV] + P.cinf_A[m, 1] * u[0] - P.cinf_b[m], gx,
                np.array([P.cinf_A[m, 1], 0.0]))
    return np.array(c), np.array(Cx), np.array(Cu), hxy


def collision_row(P, z, o, o_psi=None):
    """mpc.py:226 writes d_min^2 - |p - o|^2 <= 0; the solvers use the equivalent, better scaled
    d_min - |p - o| <= 0 (same feasible set, multiplier scaled by 2|p - o|).  Returns the row
    value, its gradient w.r.t. zeta and its 2x2 Hessian block on (x, y).
    With an obstacle heading: the OBCA rows of mpc.py:211-221 with their duals maximised out
    (oracle/obca.py): margin - rectangle distance, gradient on (x, y, psi), Gauss-Newton (no Hessian)."""
    if o_psi is not None:
        from . import obca
        d, gd = obca.rect_distance(np.array([z[D.IX], z[D.IY], z[D.IPSI]]), np.array([o[0], o[1], o_psi]))
        g = np.zeros(9)
        g[D.IX], g[D.IY], g[D.IPSI] = -gd[0], -gd[1], -gd[2]
        return nlp.OBCA_MARGIN - d, g, 0.0
    dx, dy = z[D.IX] - o[0], z[D.IY] - o[1]
    dist = max(np.hypot(dx, dy), 1e-9)
    nx_, ny_ = dx / dist, dy / dist
    g = np.zeros(9)
    g[D.IX], g[D.IY] = -nx_, -ny_
    H = -np.array([[1 - nx_ * nx_, -nx_ * ny_], [-nx_ * ny_, 1 - ny_ * ny_]]) / dist
    return P.d_min - dist, g, H


def add_row_hessians(M, s, hxy):
    """M[(x,y),(x,y)] += sum_i s_i * H_i for rows carrying a 2x2 Hessian block."""
    for si, h in zip(s, hxy):
        if not np.isscalar(h):
            M[np.ix_([D.IX, D.IY], [D.IX, D.IY])] += si * h


def terminal_rows(P, prob, z):
    N = P.N
    c, Cx, hxy = [], [], []
    g = np.zeros(9); g[D.IEY] = 1.0
    c.append(z[D.IEY] - P.ey_lim); Cx.append(g); hxy.append(0.0)
    g = np.zeros(9); g[D.IEY] = -1.0
    c.append(-P.ey_lim - z[D.IEY]); Cx.append(g); hxy.append(0.0)
    cval, g, hc = collision_row(P, z, prob.obs[N], None if prob.obs_psi is None else prob.obs_psi[N])
    c.append(cval); Cx.append(g); hxy.append(hc)
    return np.array(c), np.array(Cx), np.zeros((3, 2)), hxy


def all_rows_values(P, prob, Z, U):
    """c values only, list over stages 0..N (used by the forward pass)."""
    out = []
    up = prob.u_prev
    for k in range(P.N):
        out.append(stage_rows(P, prob, k, Z[k], up, U[k])[0])
        up = U[k]
    out.append(terminal_rows(P, prob, Z[P.N])[0])
    return out


def x0_feasible(P, prob, tol=1e-6):
    """Rows that involve only the fixed initial state (mpc.py:316-317 at k=0, :298-299 at
    k=0).  If x0 violates them the reference NLP is infeasible."""
    v, ey = prob.x0[D.IV], prob.x0[D.IEY]
    return (P.v_min - tol <= v <= P.v_max + tol) and (abs(ey) <= P.ey_lim + tol)



def dyn_hessian(z, u, curv, lam, dt):
    """Approximate second-order dynamics term  sum_i lam_i d2F_i/d(zeta,u)^2  ~=  dt * Hess(lam . f)
    (one-point quadrature of the continuous right-hand side at (z, u); the exact term does not
    exist in the reference either -- CasADi builds the exact Lagrangian Hessian by AD).  It only
    shapes the Newton direction: the KKT point reached does not depend on it.
    Returns Hxx[9,9], Hux[2,9], Huu[2,2]."""
    ey, epsi, v, psi = z[D.IEY], z[D.IEPSI], z[D.IV], z[D.IPSI]
    df = u[1]
    K = D.curvature(z[D.IS], curv)
    rho = D.L_R / (D.L_F + D.L_R)
    t = np.tan(df)
    beta = np.arctan(rho * t)
    q = 1.0 + rho * rho * t * t
    b1 = rho * (1.0 + t * t) / q
    b2 = rho * 2.0 * t * (1.0 - rho * rho) / (q * q) * (1.0 + t * t)
    th, ph = beta + epsi, psi + beta
    cth, sth, cph, sph, cb, sb = np.cos(th), np.sin(th), np.cos(ph), np.sin(ph), np.cos(beta), np.sin(beta)
    Dn = 1.0 - K * ey
    m = lam[D.IS] - K * lam[D.IEPSI]
    n = (lam[D.IEPSI] + lam[D.IPSI]) / D.L_R
    ley, lx, ly = lam[D.IEY], lam[D.IX], lam[D.IY]
    G_ey = m * cth * K / Dn ** 2
    G_epsi = -m * sth / Dn + ley * cth
    G_psi = -lx * sph + ly * cph
    G_b = G_epsi + n * cb + G_psi
    G_eyey = m * cth * 2 * K * K / Dn ** 3
    G_eyepsi = -m * sth * K / Dn ** 2
    G_epsiepsi = -m * cth / Dn - ley * sth
    G_psipsi = -lx * cph - ly * sph
    G_bb = G_epsiepsi - n * sb + G_psipsi
    Hxx = np.zeros((9, 9)); Hux = np.zeros((2, 9)); Huu = np.zeros((2, 2))
    def setx(i, j, val):
        Hxx[i, j] = val; Hxx[j, i] = val
    setx(D.IEY, D.IEY, v * G_eyey)
    setx(D.IEY, D.IEPSI, v * G_eyepsi)
    setx(D.IEPSI, D.IEPSI, v * G_epsiepsi)
    setx(D.IPSI, D.IPSI, v * G_psipsi)
    setx(D.IV, D.IEY, G_ey)
    setx(D.IV, D.IEPSI, G_epsi)
    setx(D.IV, D.IPSI, G_psi)
    Hux[1, D.IEY] = v * G_eyepsi * b1          # d2/(dbeta dey) = G_ey,beta = G_ey,epsi
    Hux[1, D.IEPSI] = v * G_epsiepsi * b1
    Hux[1, D.IPSI] = v * G_psipsi * b1
    Hux[1, D.IV] = G_b * b1
    Huu[1, 1] = v * (G_bb * b1 * b1 + G_b * b2)
    return dt * Hxx, dt * Hux, dt * Huu

# --------------------------------------------------------------------------------------
class Result:
    pass


def stage_cost_grad(P, z, u):
    lx = np.zeros(9)
    lx[D.IEY], lx[D.IEPSI] = 2 * z[D.IEY], 2 * z[D.IEPSI]
    lu = 2 * P.w_u * u if u is not None else None
    return lx, lu


def solve(P: nlp.Params, prob: nlp.Problem, mlp: nlp.MLPTerm = None, opt: Options = None,
          verbose=False):
    """Returns a Result with status (0 converged, 1 iteration limit, 2 x0 violates its own
    rows, 3 regularisation limit in the backward pass, 4 line search failed, 5 stalled at an
    infeasible point, 6 no further progress at a point within the reference's IPOPT tolerances), iters, Z[N+1,7],
    U[N,2], cost, viol (max inequality-row violation in the reference's units)."""
    opt = opt or Options()
    N = P.N
    res = Result()
    res.status, res.iters = 1, 0
    if not x0_feasible(P, prob):
        res.status = 2
        res.Z = res.U = None
        res.cost = np.nan
        res.viol = np.inf
        return res

    U = np.array(prob.u_init, dtype=np.float64) if prob.u_init is not None else nlp.initial_guess(P, prob)
    Z = D.frenet_rollout(prob.x0, U, prob.curv, P.dt, P.n_rk)
    mu = opt.mu0
    C = all_rows_values(P, prob, Z, U)
    Y = [np.maximum(-c, opt.y_init_min) for c in C]
    S = [mu / y for y in Y]
    mu_floor = opt.mu_floor
    reg = 0.0

    def barrier_cost(Zt, Ut, Yt, mu_):
        return nlp.cost(P, prob, Zt, Ut, mlp) - mu_ * sum(np.sum(np.log(y)) for y in Yt)

    def infeas(Ct, Yt):
        return sum(np.sum(np.abs(c + y)) for c, y in zip(Ct, Yt))

    def terminal_cost_derivs(zN):
        lx = np.zeros(9); lxx = np.zeros((9, 9))
        lx[D.IEY], lx[D.IEPSI] = 2 * zN[D.IEY], 2 * zN[D.IEPSI]
        lxx[D.IEY, D.IEY] = lxx[D.IEPSI, D.IEPSI] = 2.0
        if mlp is None:
            lx[D.IS] -= 1.0
        else:
            _, g, H = mlp.value(zN[D.IS], zN[D.IV], prob.nn_ctx, order=2)
            lx[D.IS] -= g[0]; lx[D.IV] -= g[1]
            lxx[D.IS, D.IS] -= H[0, 0]; lxx[D.IS, D.IV] -= H[0, 1]
            lxx[D.IV, D.IS] -= H[0, 1]; lxx[D.IV, D.IV] -= H[1, 1]
        return lx, lxx

    need_jac = True
    for it in range(opt.max_iter + 1):
        res.iters = it
        # ---- sweep 1: sensitivities of the current trajectory --------------------------
        if need_jac:
            _, A, B = D.frenet_rollout(prob.x0, U, prob.curv, P.dt, P.n_rk, want_jac=True)
            rows = []
            for k in range(N):
                up = prob.u_prev if k == 0 else U[k - 1]
                rows.append(stage_rows(P, prob, k, Z[k], up, U[k]))
            rows.append(terminal_rows(P, prob, Z[N]))
            lxN, lxxN = terminal_cost_derivs(Z[N])
            # ---- sweep 2 (backward, vectors only): adjoint + KKT residuals -------------
            lam = [None] * (N + 1)
            c, Cx, Cu, _ = rows[N]
            lam[N] = lxN + Cx.T @ S[N]
            stat = 0.0
            rp = np.max(np.abs(c + Y[N]))
            s_max = np.max(S[N])
            sy_min, sy_max = np.min(S[N] * Y[N]), np.max(S[N] * Y[N])
            for k in range(N - 1, -1, -1):
                c, Cx, Cu, _ = rows[k]
                lx, lu = stage_cost_grad(P, Z[k], U[k])
                gu = lu + B[k].T @ lam[k + 1][:7] + lam[k + 1][7:] + Cu.T @ S[k]
                stat = max(stat, np.max(np.abs(gu)))
                lam[k] = lx + Cx.T @ S[k]
                lam[k][:7] += A[k].T @ lam[k + 1][:7]
                rp = max(rp, np.max(np.abs(c + Y[k])))
                sy = S[k] * Y[k]
                s_max = max(s_max, np.max(S[k]))
                sy_min, sy_max = min(sy_min, sy.min()), max(sy_max, sy.max())
            need_jac = False
        if verbose:
            print(f"it {it:3d} mu {mu:.1e} J {nlp.cost(P, prob, Z, U, mlp):+.9f} stat {stat:.2e} rp {rp:.2e} "
                  f"sy [{sy_min:.1e},{sy_max:.1e}] reg {reg:.1e}")
        if stat <= opt.tol * max(1.0, s_max) and rp <= opt.tol_rp and sy_max <= opt.tol_comp:
            res.status = 0
            break
        acceptable = opt.acc_tol > 0 and stat <= opt.acc_tol * max(1.0, s_max) and rp <= opt.acc_rp and sy_max <= opt.acc_comp
        if it == opt.max_iter:
            if acceptable:
                res.status = 6
            break
        if it >= opt.stall_iter and rp > opt.stall_rp:      # stalled at an infeasible point
            res.status = 5
            break
        # ---- barrier update (monotone, IPOPT-style) -------------------------------------
        while mu > mu_floor and max(stat, rp, abs(sy_max - mu), abs(sy_min - mu)) <= opt.kappa_eps * mu:
            mu = max(mu_floor, min(opt.kappa_mu * mu, mu ** opt.theta_mu))
        phi, theta = barrier_cost(Z, U, Y, mu), infeas(C, Y)
        # ---- sweep 3 (backward): Riccati on the perturbed KKT system --------------------
        while True:
            ok = True
            reg_hint = 0.0
            c, Cx, Cu, hxy = rows[N]
            s, y = S[N], Y[N]
            rhat = s * c + mu
            sig = s / y
            Vx = lxN + Cx.T @ (s + rhat / y)
            Vxx = lxxN + Cx.T @ (sig[:, None] * Cx)
            add_row_hessians(Vxx, s, hxy)
            ks = [None] * (N + 1); Ks = [None] * (N + 1); ky = [None] * (N + 1); Ky = [None] * (N + 1)
            ku = [None] * N; Ku = [None] * N
            ks[N] = rhat / y; Ks[N] = sig[:, None] * Cx
            ky[N] = -(c + y); Ky[N] = -Cx
            for k in range(N - 1, -1, -1):
                c, Cx, Cu, hxy = rows[k]
                s, y = S[k], Y[k]
                At = np.zeros((9, 9)); At[:7, :7] = A[k]
                Bt = np.zeros((9, 2)); Bt[:7] = B[k]; Bt[7:] = np.eye(2)
                lx, lu = stage_cost_grad(P, Z[k], U[k])
                lxx = np.zeros((9, 9)); lxx[D.IEY, D.IEY] = lxx[D.IEPSI, D.IEPSI] = 2.0
                luu = 2 * P.w_u * np.eye(2)
                Qx = lx + At.T @ Vx + Cx.T @ s
                Qu = lu + Bt.T @ Vx + Cu.T @ s
                Qxx = lxx + At.T @ Vxx @ At
                add_row_hessians(Qxx, s, hxy)
                Quu = luu + Bt.T @ Vxx @ Bt
                Qux = Bt.T @ Vxx @ At
                if opt.second_order:
                    Hxx, Hux, Huu = dyn_hessian(Z[k], U[k], prob.curv, lam[k + 1][:7], P.dt)
                    Qxx += Hxx; Qux += Hux; Quu += Huu
                rhat = s * c + mu
                sig = s / y
                Qxh = Qx + Cx.T @ (rhat / y)
                Quh = Qu + Cu.T @ (rhat / y)
                Qxxh = Qxx + Cx.T @ (sig[:, None] * Cx)
                Quuh = Quu + Cu.T @ (sig[:, None] * Cu)
                Quxh = Qux + Cu.T @ (sig[:, None] * Cx)
                Qr = Quuh + reg * np.eye(2)
                det = Qr[0, 0] * Qr[1, 1] - Qr[0, 1] * Qr[1, 0]
                if not (Qr[0, 0] > 0 and det > 1e-12 * Qr[0, 0] * Qr[1, 1]):
                    ok = False
                    # the shift that makes this stage's Quu positive definite (inertia-correction shortcut
                    # instead of climbing reg_up by reg_up)
                    a_, b_, c_ = Quuh[0, 0], Quuh[1, 1], Quuh[0, 1]
                    lmin = 0.5 * (a_ + b_) - np.sqrt(0.25 * (a_ - b_) ** 2 + c_ * c_)
                    reg_hint = -lmin * opt.reg_jump if lmin < 0 else 0.0
                    break
                Qinv = np.array([[Qr[1, 1], -Qr[0, 1]], [-Qr[1, 0], Qr[0, 0]]]) / det
                ku[k] = -Qinv @ Quh
                Ku[k] = -Qinv @ Quxh
                ks[k] = (rhat + s * (Cu @ ku[k])) / y
                Ks[k] = sig[:, None] * (Cx + Cu @ Ku[k])
                ky[k] = -(c + y) - Cu @ ku[k]
                Ky[k] = -Cx - Cu @ Ku[k]
                Vx = Qxh + Ku[k].T @ Quh + Ku[k].T @ (Quuh @ ku[k]) + Quxh.T @ ku[k]
                Vxx = Qxxh + Ku[k].T @ Quuh @ Ku[k] + Ku[k].T @ Quxh + Quxh.T @ Ku[k]
                Vxx = 0.5 * (Vxx + Vxx.T)
            if ok:
                break
            reg = max(reg * opt.reg_up, opt.reg_min, reg_hint)
            if reg > opt.reg_max:
                break
        if reg > opt.reg_max:
            res.status = 6 if acceptable else 3
            break
        # ---- sweeps 4..: closed-loop forward pass with fraction-to-boundary -------------
        tau = max(opt.tau_min, 1.0 - mu)
        accepted = False
        alpha = 1.0
        if opt.predict_alpha:
            # largest step keeping every slack inside the fraction-to-boundary rule on the
            # linearised closed-loop model
            dz = np.zeros(9)
            for k in range(N + 1):
                c = rows[k][0]
                dy = ky[k] + Ky[k] @ dz
                neg = dy < 0
                if np.any(neg):
                    alpha = min(alpha, np.min(tau * Y[k][neg] / (-dy[neg])))
                if k == N:
                    break
                du = ku[k] + Ku[k] @ dz
                dzn = np.zeros(9)
                dzn[:7] = A[k] @ dz[:7] + B[k] @ du
                dzn[7:] = du
                dz = dzn
            alpha *= opt.alpha_safety
        for _ in range(opt.n_alpha):
            Zn = np.empty_like(Z); Un = np.empty_like(U)
            Yn = [None] * (N + 1); Sn = [None] * (N + 1)
            Zn[0] = Z[0]
            fail = False
            upn = prob.u_prev
            for k in range(N + 1):
                up = prob.u_prev if k == 0 else U[k - 1]
                dz = np.concatenate([Zn[k] - Z[k], upn - up])
                Yn[k] = Y[k] + alpha * ky[k] + Ky[k] @ dz
                Sn[k] = S[k] + alpha * ks[k] + Ks[k] @ dz
                if np.any(Yn[k] < (1 - tau) * Y[k]):             # fraction to the boundary
                    fail = True
                    break
                Sn[k] = np.maximum(Sn[k], (1 - tau) * S[k])      # multiplier safeguard
                if k == N:
                    break
                Un[k] = U[k] + alpha * ku[k] + Ku[k] @ dz
                Zn[k + 1] = D.frenet_rk4_step(Zn[k], Un[k], prob.curv, P.dt, P.n_rk)
                if not np.all(np.isfinite(Zn[k + 1])):
                    fail = True
                    break
                upn = Un[k]
            if not fail:
                Cn = all_rows_values(P, prob, Zn, Un)
                phin, thetan = barrier_cost(Zn, Un, Yn, mu), infeas(Cn, Yn)
                if verbose > 1:
                    print(f"      trial alpha {alpha:.3e}: dphi {phin - phi:+.3e} (need < {-opt.eps_phi * abs(phi):.1e}) theta {theta:.3e} -> {thetan:.3e}")
                if np.isfinite(phin) and (phin < phi - opt.eps_phi * abs(phi) or thetan < theta * (1 - opt.gamma_theta)
                                          or (thetan <= opt.theta_small and phin <= phi + opt.eps_phi * max(1, abs(phi)))):
                    accepted = True
                    break
            elif verbose > 1:
                print(f"      trial alpha {alpha:.3e}: fraction-to-boundary / non-finite")
            alpha *= 0.5
        if accepted:
            Z, U, Y, S, C = Zn, Un, Yn, Sn, Cn
            reg = reg / opt.reg_down if reg > opt.reg_min else 0.0
            need_jac = True
        else:
            reg = max(reg * opt.reg_up, opt.reg_min)
            if reg > opt.reg_max:
                res.status = 6 if acceptable else 4
                break
    res.Z, res.U = Z, U
    res.cost = nlp.cost(P, prob, Z, U, mlp)
    res.viol = nlp.max_violation(P, prob, Z, U)
    res.mu = mu
    res.S, res.Y = S, Y
    res.stat, res.rp = stat, rp
    return res
