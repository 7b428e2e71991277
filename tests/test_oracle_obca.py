"""Groundwork for the OBCA collision mode (mpc.py:211-221): the reference's dual formulation maximised over its
dual variables equals the rectangle-to-rectangle distance, whose closed form and gradient are what a stage-local
row needs (oracle/obca.py)."""
import numpy as np

from oracle import obca


def _poses(rng, n, lo, hi):
    out = []
    while len(out) < n:
        ego = np.array([rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-np.pi, np.pi)])
        ang, dist = rng.uniform(-np.pi, np.pi), rng.uniform(lo, hi)
        obs = np.array([ego[0] + dist * np.cos(ang), ego[1] + dist * np.sin(ang), rng.uniform(-np.pi, np.pi)])
        out.append((ego, obs))
    return out


def test_h_representation_matches_reference_convention():
    A, b = obca.rotation_translation([3.0, -1.0], 0.4)
    c, s = np.cos(0.4), np.sin(0.4)
    front = np.array([3.0, -1.0]) + 2.235 * np.array([c, s])          # the front bumper's centre lies on face 0
    assert abs(A[0] @ front - b[0]) < 1e-12 and np.all(A @ np.array([3.0, -1.0]) < b)
    assert np.allclose(A[:2], -A[2:])


def test_dual_formulation_equals_rectangle_distance():
    rng = np.random.default_rng(7)
    for ego, obs in _poses(rng, 40, 3.0, 12.0):
        d, _ = obca.rect_distance(ego, obs)
        v = obca.dual_value(ego, obs)
        assert abs(max(d, 0.0) - v) < 1e-5 * max(1.0, d), (ego, obs, d, v)      # the dual certifies distances, not depths
    # intersecting rectangles: the dual cannot certify any positive distance (lambda = mu = 0 is its optimum)
    for ego, obs in _poses(rng, 10, 0.0, 0.9):
        d, g = obca.rect_distance(ego, obs)
        assert d <= 0.0                                            # minus the penetration depth
        assert obca.dual_value(ego, obs) < 1e-6


def test_rectangle_distance_gradient():
    rng = np.random.default_rng(3)
    checked = 0
    for ego, obs in _poses(rng, 60, 3.5, 10.0) + _poses(rng, 30, 0.2, 1.5):      # disjoint and intersecting
        d, g = obca.rect_distance(ego, obs)
        if abs(d) < 0.05:
            continue
        fd = np.zeros(3)
        smooth = True
        for i in range(3):
            e = np.zeros(3); e[i] = 1e-6
            dp, gp = obca.rect_distance(ego + e, obs)
            dm, gm = obca.rect_distance(ego - e, obs)
            fd[i] = (dp - dm) / 2e-6
            smooth = smooth and np.allclose(gp, gm, atol=1e-3)      # skip the kinks (the closest pair of features changes)
        if smooth:
            assert np.allclose(g, fd, atol=2e-5), (ego, obs, g, fd)
            checked += 1
    assert checked >= 55
    # a hand-checkable case: two axis-aligned cars 7 m apart nose to tail: distance 7 - 4.47, gradient (+-1, 0, 0)
    d, g = obca.rect_distance(np.array([7.0, 0.0, 0.0]), np.array([0.0, 0.0, 0.0]))
    assert abs(d - (7.0 - 4.47)) < 1e-12 and np.allclose(g[:2], [1.0, 0.0])


def test_obca_mpc_eliminated_rows_vs_reference_formulation_with_explicit_duals():
    """A small MPC problem in OBCA mode, solved (a) by the python oracle with the ONE eliminated row per stage and
    (b) by scipy SLSQP on the reference's own formulation -- lambda_k, mu_k in R^4 as decision variables with the
    rows of mpc.py:216-221: the same optimum, with the collision row active."""
    from scipy.optimize import minimize
    from oracle import nlp, solver, dynamics as D
    N = 8
    base = nlp.Params(N=40)
    P = nlp.Params(N=N, cinf_A=base.cinf_A, cinf_b=base.cinf_b)
    x0 = np.array([5.0, 2.8, 5.0, 0.0, 0.0, 3.0, 0.0])
    opose = np.array([11.4, 3.6, 1.1])                             # a car standing askew across the ego's lane: the
                                                                   # ego slows down and swerves to the lane's edge
    prob = nlp.Problem(x0=x0, u_prev=np.array([0.0, 0.0]), curv=(1e30, 1e30, 0.0),
                       obs=np.tile(opose[:2], (N + 1, 1)), obs_psi=np.full(N + 1, opose[2]))
    r = solver.solve(P, prob)
    assert r.status == 0 and r.viol <= 1e-9
    d_end = obca.rect_distance(r.Z[N, [0, 1, 6]], opose)[0]
    assert abs(d_end - nlp.OBCA_MARGIN) < 1e-7                       # the row is active at the end of the horizon
    free = solver.solve(P, nlp.Problem(x0=x0, u_prev=prob.u_prev, curv=prob.curv, obs=np.full((N + 1, 2), -50.0)))
    assert free.cost < r.cost - 1e-3                               # and it costs progress

    # (b) the reference's formulation: variables u[N,2], lambda[N,4], mu[N,4] (k = 1..N), single shooting for x
    G, g = obca.rotation_translation([0.0, 0.0], 0.0)
    A, b = obca.rotation_translation(opose[:2], opose[2])
    other = nlp.Problem(x0=x0, u_prev=prob.u_prev, curv=prob.curv, obs=np.full((N + 1, 2), -50.0))   # rows without collision

    def unpack(v):
        return v[:2 * N].reshape(N, 2), v[2 * N:6 * N].reshape(N, 4), v[6 * N:].reshape(N, 4)

    def roll(U):
        return D.frenet_rollout(x0, U, prob.curv, P.dt, P.n_rk)

    def f(v):
        U = unpack(v)[0]
        return nlp.cost(P, other, roll(U), U)

    def ineq(v):
        U, lam, mu = unpack(v)
        Z = roll(U)
        rows = [-nlp.inequality_rows(P, other, Z, U)]
        for k in range(1, N + 1):
            p = Z[k, :2]
            rows.append([-g @ mu[k - 1] + (A @ p - b) @ lam[k - 1] - nlp.OBCA_MARGIN,          # mpc.py:216
                         1.0 - np.sum((A.T @ lam[k - 1]) ** 2)])                                # mpc.py:218
        return np.concatenate([np.ravel(x) for x in rows])

    def eq(v):
        U, lam, mu = unpack(v)
        Z = roll(U)
        out = []
        for k in range(1, N + 1):
            c, s = np.cos(Z[k, 6]), np.sin(Z[k, 6])
            R = np.array([[c, -s], [s, c]])
            out.append(G.T @ mu[k - 1] + R.T @ A.T @ lam[k - 1])                               # mpc.py:217
        return np.concatenate(out)

    # started next to the oracle's solution (controls perturbed by 0.02, duals at their per-stage optimum): the
    # bilinear dual formulation is too degenerate for SLSQP from a cold start, but it must come back to the same point
    rng = np.random.default_rng(1)
    U0 = r.U + 0.02 * rng.normal(size=r.U.shape)
    duals = [obca.dual_value(r.Z[k, [0, 1, 6]], opose, return_arg=True)[1] for k in range(1, N + 1)]
    v_star = np.concatenate([r.U.ravel(), np.array([d[:4] for d in duals]).ravel(), np.array([d[4:] for d in duals]).ravel()])
    assert ineq(v_star).min() > -1e-9 and np.abs(eq(v_star)).max() < 1e-9      # the oracle's point is feasible for the reference's rows
    v0 = v_star.copy(); v0[:2 * N] = U0.ravel()
    bounds = [(None, None)] * (2 * N) + [(0, None)] * (8 * N)
    res = minimize(f, v0, method="SLSQP", bounds=bounds,
                   constraints=[{"type": "ineq", "fun": ineq}, {"type": "eq", "fun": eq}],
                   options={"maxiter": 300, "ftol": 1e-10})
    assert res.success, res.message
    U = unpack(res.x)[0]
    assert abs(res.fun - r.cost) < 1e-5 * max(1.0, abs(r.cost)), (res.fun, r.cost)
    assert np.max(np.abs(U - r.U)) < 2e-3


def test_c_port_of_the_rectangle_distance_equals_the_python_statement():
    """oracle/igt_oracle.c restates oracle/obca.rect_distance (incl. the penetration branch)."""
    from oracle import c_oracle
    rng = np.random.default_rng(11)
    pairs = _poses(rng, 200, 2.5, 12.0) + _poses(rng, 100, 0.1, 2.5)
    ego = np.array([p[0] for p in pairs]); obs = np.array([p[1] for p in pairs])
    d, g = c_oracle.rect_sdist(ego, obs)
    for i in range(len(pairs)):
        do, go = obca.rect_distance(ego[i], obs[i])
        assert abs(do - d[i]) < 1e-12 and np.max(np.abs(go - g[i])) < 1e-11
    assert (d < 0).sum() > 20 and (d > 0).sum() > 150


def test_c_oracle_obca_mode_equals_python_oracle():
    """same algorithm, same rows: the C port reproduces the python oracle's OBCA solves"""
    from oracle import nlp, solver, c_oracle
    from igt_mpc_int_b200 import scenarios as S
    N = 20
    base = nlp.Params(N=40)
    P = nlp.Params(N=N, d_min=0.0, cinf_A=base.cinf_A, cinf_b=base.cinf_b)        # mpc.py:42-43: d_min = 0 in OBCA mode
    pb = S.mid_episode(16, N=N, seed=5)
    assert np.abs(pb.obs_psi).max() > 0.5
    o = c_oracle.COracle(P, max_iter=60).solve(pb.x0, pb.u_prev, pb.curv, pb.obs, obs_psi=pb.obs_psi)
    opt = solver.Options(); opt.max_iter = 60
    for i in (0, 2, 5):
        prob = nlp.Problem(x0=pb.x0[i], u_prev=pb.u_prev[i], curv=tuple(pb.curv[i]), obs=pb.obs[i], obs_psi=pb.obs_psi[i])
        r = solver.solve(P, prob, opt=opt)
        assert r.status == o["status"][i] and r.iters == o["iters"][i]
        if r.status == 0:
            assert abs(r.cost - o["cost"][i]) < 1e-9 and np.max(np.abs(r.U - o["U"][i])) < 1e-7
            assert abs(r.viol - o["viol"][i]) < 1e-12
