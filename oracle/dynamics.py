"""ORACLE (test infrastructure, not product code) -- fp64 numpy restatement of the
reference vehicle models and their analytic Jacobians.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Follows (all paths relative to the reference repo):
  * common/kinematic_bicycle_model_frenet.py:69-185  -- 7-state kinematic bicycle,
    RK4 with `num_rk4_steps` sub-steps (numpy branch :70-127, casadi branch :129-185
    is the same arithmetic).  Internal k-vector order is (s, ey, epsi, v, x, y, psi)
    (:108); the k4 stage evaluates xdot, ydot at psi + h/2*k3[6] (:111, :169) -- that
    quirk is reproduced here on purpose.
  * common/kinematic_bicycle_model.py:15-50 -- 4-state Cartesian Euler model.
  * mpc.py:183-200 -- curvature K(s) as casadi `pw_const`, i.e.
    K(s) = Kval * (s >= b0) - Kval * (s >= b1), derivative identically zero.

State layout used at the planner boundary (mpc.py:163): z = [x, y, s, ey, epsi, v, psi],
input u = [a, df] (mpc.py:164).  Jacobians do not exist in the reference (CasADi AD);
they are derived here by forward sensitivities through the RK stages and FD-checked in
tests/test_oracle_dynamics.py.
"""
import numpy as np

L_R = 2.235   # mpc.py:50-51  (l_f = l_r = 4.47/2)
L_F = 2.235

# planner-order index of each state
IX, IY, IS, IEY, IEPSI, IV, IPSI = range(7)


def curvature(s, curv):
    """pw_const curvature, mpc.py:199.  curv = (b0, b1, Kval)."""
    b0, b1, kval = curv
    return kval * (1.0 if s >= b0 else 0.0) - kval * (1.0 if s >= b1 else 0.0)


def beta_of(df):
    """Slip angle, kinematic_bicycle_model_frenet.py:72; returns (beta, dbeta/ddf)."""
    rho = L_R / (L_F + L_R)
    t = np.tan(df)
    beta = np.arctan(rho * t)
    dbeta = rho * (1.0 + t * t) / (1.0 + rho * rho * t * t)
    return beta, dbeta


def rhs(z, u, curv, want_jac=False):
    """Continuous-time right-hand side zdot = f(z, u) in PLANNER order.

    kinematic_bicycle_model_frenet.py:71-91.  Returns zdot[7] and, if asked, the dense
    Jacobians fz[7,7], fu[7,2] (13 + 7 structural non-zeros).
    """
    ey, epsi, v, psi = z[IEY], z[IEPSI], z[IV], z[IPSI]
    a, df = u
    K = curvature(z[IS], curv)
    beta, dbeta = beta_of(df)
    c1, s1 = np.cos(beta + epsi), np.sin(beta + epsi)
    cp, sp = np.cos(psi + beta), np.sin(psi + beta)
    sb = np.sin(beta)
    den = 1.0 - K * ey
    sdot = v * c1 / den
    zd = np.empty(7)
    zd[IS] = sdot
    zd[IEY] = v * s1
    zd[IEPSI] = v * sb / L_R - sdot * K
    zd[IV] = a
    zd[IX] = v * cp
    zd[IY] = v * sp
    zd[IPSI] = v * sb / L_R
    if not want_jac:
        return zd
    fz = np.zeros((7, 7))
    fu = np.zeros((7, 2))
    # sdot
    ds_dey = v * c1 * K / (den * den)
    ds_depsi = -v * s1 / den
    ds_dv = c1 / den
    ds_dbeta = -v * s1 / den
    fz[IS, IEY], fz[IS, IEPSI], fz[IS, IV] = ds_dey, ds_depsi, ds_dv
    fu[IS, 1] = ds_dbeta * dbeta
    # eydot
    fz[IEY, IEPSI], fz[IEY, IV] = v * c1, s1
    fu[IEY, 1] = v * c1 * dbeta
    # epsidot = v sin(beta)/lr - K sdot
    cb = np.cos(beta)
    fz[IEPSI, IEY] = -K * ds_dey
    fz[IEPSI, IEPSI] = -K * ds_depsi
    fz[IEPSI, IV] = sb / L_R - K * ds_dv
    fu[IEPSI, 1] = (v * cb / L_R - K * ds_dbeta) * dbeta
    # vdot
    fu[IV, 0] = 1.0
    # xdot, ydot
    fz[IX, IV], fz[IX, IPSI] = cp, -v * sp
    fu[IX, 1] = -v * sp * dbeta
    fz[IY, IV], fz[IY, IPSI] = sp, v * cp
    fu[IY, 1] = v * cp * dbeta
    # psidot
    fz[IPSI, IV] = sb / L_R
    fu[IPSI, 1] = v * cb / L_R * dbeta
    return zd, fz, fu


def frenet_rk4_step(z, u, curv, dt=0.1, n_rk=4, want_jac=False):
    """One MPC step z+ = F(z, u): n_rk RK4 sub-steps, with the reference's k4 psi quirk.

    kinematic_bicycle_model_frenet.py:93-127.  With want_jac returns (z+, A[7,7], B[7,2]).
    """
    h = dt / n_rk
    z = np.array(z, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    half = np.full(7, 0.5)
    full = np.ones(7)
    full[IPSI] = 0.5          # :111 -- psi + h/2*k3[6] inside the k4 evaluation
    if want_jac:
        S = np.zeros((7, 9))
        S[:, :7] = np.eye(7)
        E = np.zeros((2, 9))
        E[:, 7:] = np.eye(2)
    for _ in range(n_rk):
        if not want_jac:
            k1 = rhs(z, u, curv)
            k2 = rhs(z + h * half * k1, u, curv)
            k3 = rhs(z + h * half * k2, u, curv)
            k4 = rhs(z + h * full * k3, u, curv)
        else:
            k1, fz, fu = rhs(z, u, curv, True)
            d1 = fz @ S + fu @ E
            k2, fz, fu = rhs(z + h * half * k1, u, curv, True)
            d2 = fz @ (S + h * half[:, None] * d1) + fu @ E
            k3, fz, fu = rhs(z + h * half * k2, u, curv, True)
            d3 = fz @ (S + h * half[:, None] * d2) + fu @ E
            k4, fz, fu = rhs(z + h * full * k3, u, curv, True)
            d4 = fz @ (S + h * full[:, None] * d3) + fu @ E
            S = S + h / 6.0 * (d1 + 2 * d2 + 2 * d3 + d4)
        z = z + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    if want_jac:
        return z, S[:, :7].copy(), S[:, 7:].copy()
    return z


def frenet_rollout(z0, U, curv, dt=0.1, n_rk=4, want_jac=False):
    """Roll N steps.  U[N,2] -> Z[N+1,7] (and A[N,7,7], B[N,7,2])."""
    U = np.asarray(U, dtype=np.float64)
    N = U.shape[0]
    Z = np.empty((N + 1, 7))
    Z[0] = z0
    if want_jac:
        A = np.empty((N, 7, 7))
        B = np.empty((N, 7, 2))
    for k in range(N):
        if want_jac:
            Z[k + 1], A[k], B[k] = frenet_rk4_step(Z[k], U[k], curv, dt, n_rk, True)
        else:
            Z[k + 1] = frenet_rk4_step(Z[k], U[k], curv, dt, n_rk)
    return (Z, A, B) if want_jac else Z


def cartesian_euler_step(z, u, dt=0.1, want_jac=False):
    """4-state Euler model z = (x, y, psi, v), kinematic_bicycle_model.py:27-31."""
    x, y, psi, v = z
    a, df = u
    beta, dbeta = beta_of(df)
    lsum = L_R + L_F
    t = np.tan(df)
    cb = np.cos(beta)
    zn = np.array([x + dt * v * np.cos(psi + beta),
                   y + dt * v * np.sin(psi + beta),
                   psi + dt * (v * cb / lsum * t),
                   v + dt * a])
    if not want_jac:
        return zn
    A = np.eye(4)
    B = np.zeros((4, 2))
    A[0, 2] = -dt * v * np.sin(psi + beta)
    A[0, 3] = dt * np.cos(psi + beta)
    A[1, 2] = dt * v * np.cos(psi + beta)
    A[1, 3] = dt * np.sin(psi + beta)
    A[2, 3] = dt * cb / lsum * t
    B[0, 1] = -dt * v * np.sin(psi + beta) * dbeta
    B[1, 1] = dt * v * np.cos(psi + beta) * dbeta
    B[2, 1] = dt * v / lsum * (-np.sin(beta) * dbeta * t + cb * (1.0 + t * t))
    B[3, 0] = dt
    return zn, A, B


def cartesian_rollout(z0, U, dt=0.1, want_jac=False):
    U = np.asarray(U, dtype=np.float64)
    N = U.shape[0]
    Z = np.empty((N + 1, 4))
    Z[0] = z0
    if want_jac:
        A = np.empty((N, 4, 4))
        B = np.empty((N, 4, 2))
    for k in range(N):
        if want_jac:
            Z[k + 1], A[k], B[k] = cartesian_euler_step(Z[k], U[k], dt, True)
        else:
            Z[k + 1] = cartesian_euler_step(Z[k], U[k], dt)
    return (Z, A, B) if want_jac else Z
