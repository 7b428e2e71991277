"""Pins for the oracle SOLVER (the reference's CasADi/IPOPT cannot run here -- "parity unpinned"
against IPOPT itself; these tests pin the oracle against (a) the NLP optima found by the survey's
independent SLSQP probe, (b) scipy.optimize.minimize(SLSQP) run here on the same NLP rows, (c) the
readable numpy restatement oracle/solver.py)."""
import numpy as np
import pytest
from scipy.optimize import minimize

from oracle import nlp, solver, dynamics as D, c_oracle
from igt_mpc_int_b200 import scenarios as S

FAR = lambda N: np.full((N + 1, 2), -20.0)
CURV_L = (19.3, 19.3 + 8.6 * np.pi / 2, 1 / 8.6)
X0_A = np.array([1.91, 2.8, 1.91, 0.0, 0.0, 0.0, 0.0])
X0_B = np.array([16.0, 2.85, 16.0, 0.05, -0.02, 4.0, -0.02])


def _batch(P, probs):
    return (np.array([p.x0 for p in probs]), np.array([p.u_prev for p in probs]),
            np.array([p.curv for p in probs]), np.array([p.obs for p in probs]))


def test_survey_probe_optima(oracle_params):
    """optimal costs found with scipy SLSQP on a separate single-shooting restatement:
    case A (episode start, straight) N=10 -0.112013, N=40 -5.639325; case B (left turn entering the
    curve) N=10 -4.136490, N=40 -17.193168."""
    for N, JA, JB in ((10, -0.112013, -4.136490), (40, -5.639325, -17.193168)):
        P = oracle_params[N]
        pa = nlp.Problem(x0=X0_A, u_prev=np.array([0.1, 0.0]), curv=(1e30, 1e30, 0.0), obs=FAR(N))
        pb = nlp.Problem(x0=X0_B, u_prev=np.array([0.5, 0.0]), curv=CURV_L, obs=FAR(N))
        r = c_oracle.COracle(P).solve(*_batch(P, [pa, pb]))
        assert list(r["status"]) == [0, 0]
        assert abs(r["cost"][0] - JA) < 2e-6 and abs(r["cost"][1] - JB) < 2e-6
        assert r["viol"].max() <= 1e-9


def _slsqp(P, prob, U0):
    N = P.N

    def roll(u):
        U = u.reshape(N, 2)
        return D.frenet_rollout(prob.x0, U, prob.curv, P.dt, P.n_rk), U

    def f(u):
        Z, U = roll(u)
        return nlp.cost(P, prob, Z, U)

    def g(u):
        Z, U = roll(u)
        return -nlp.inequality_rows(P, prob, Z, U)          # SLSQP wants g >= 0

    res = minimize(f, U0.ravel(), method="SLSQP", constraints=[{"type": "ineq", "fun": g}],
                   options={"maxiter": 400, "ftol": 1e-12})
    Z, U = roll(res.x)
    return res, Z, U


@pytest.mark.parametrize("case", ["straight_start", "left_turn", "right_turn_obstacle", "collision_active"])
def test_oracle_vs_scipy_slsqp_N10(oracle_params, case):
    """Independent solver (scipy SLSQP, finite-difference derivatives) on the rows exactly as written
    in mpc.py, started from the oracle's own cold-start guess: same optimum within 1e-5 in cost and
    2e-3 in u (SLSQP's own accuracy with FD gradients)."""
    N = 10
    P = oracle_params[N]
    k = np.arange(N + 1)
    if case == "straight_start":
        prob = nlp.Problem(x0=X0_A, u_prev=np.array([0.1, 0.0]), curv=(1e30, 1e30, 0.0), obs=FAR(N))
    elif case == "left_turn":
        prob = nlp.Problem(x0=X0_B, u_prev=np.array([0.5, 0.0]), curv=CURV_L, obs=FAR(N))
    elif case == "right_turn_obstacle":
        x0 = np.array([9.5, 2.8, 9.5, -0.03, 0.01, 3.0, 0.01])
        obs = np.stack([np.full(N + 1, 22.1), 25.0 - 3.0 * 0.1 * k], 1)
        prob = nlp.Problem(x0=x0, u_prev=np.array([0.2, 0.0]), curv=(10.7, 10.7 + 11.4 * np.pi / 2, -1 / 11.4), obs=obs)
    else:   # obstacle standing ahead on the ego lane: the collision row is active at the end of the horizon
        x0 = np.array([5.0, 2.8, 5.0, 0.0, 0.0, 4.0, 0.0])
        obs = np.stack([np.full(N + 1, 14.7), np.full(N + 1, 2.8)], 1)
        prob = nlp.Problem(x0=x0, u_prev=np.array([0.0, 0.0]), curv=(1e30, 1e30, 0.0), obs=obs)
    r = c_oracle.COracle(P).solve(*_batch(P, [prob]))
    assert r["status"][0] == 0
    res, Z, U = _slsqp(P, prob, nlp.initial_guess(P, prob))
    assert res.success, res.message
    assert nlp.max_violation(P, prob, Z, U) < 1e-6
    assert abs(res.fun - r["cost"][0]) < 1e-5 * max(1.0, abs(r["cost"][0]))
    assert np.max(np.abs(U - r["U"][0])) < 2e-3
    if case == "collision_active":
        d = np.sqrt(np.sum((r["Z"][0][1:, :2] - prob.obs[1:]) ** 2, axis=1))
        assert abs(d.min() - P.d_min) < 1e-6           # the row really is active at the optimum


def test_c_oracle_equals_numpy_restatement(oracle_params):
    """oracle/igt_oracle.c and oracle/solver.py are the same algorithm: same iterates."""
    P = oracle_params[40]
    pb = S.mid_episode(16, N=40, seed=11)
    r = c_oracle.COracle(P).solve(pb.x0, pb.u_prev, pb.curv, pb.obs)
    for i in (0, 3, 7, 12):
        prob = nlp.Problem(x0=pb.x0[i], u_prev=pb.u_prev[i], curv=tuple(pb.curv[i]), obs=pb.obs[i])
        rp = solver.solve(P, prob)
        assert rp.status == r["status"][i] and rp.iters == r["iters"][i]
        if rp.status == 0:
            assert abs(rp.cost - r["cost"][i]) < 1e-10 and np.max(np.abs(rp.U - r["U"][i])) < 1e-8


def test_oracle_kkt_residuals(oracle_params):
    """At the returned point: dynamics satisfied to round-off, every inequality row <= 1e-9."""
    P = oracle_params[40]
    pb = S.mid_episode(32, N=40, seed=21)
    r = c_oracle.COracle(P).solve(pb.x0, pb.u_prev, pb.curv, pb.obs)
    ok = np.where(r["status"] == 0)[0]
    assert len(ok) >= 24
    for i in ok[:8]:
        prob = nlp.Problem(x0=pb.x0[i], u_prev=pb.u_prev[i], curv=tuple(pb.curv[i]), obs=pb.obs[i])
        assert nlp.dynamics_defect(P, prob, r["Z"][i], r["U"][i]) < 1e-12
        assert nlp.max_violation(P, prob, r["Z"][i], r["U"][i]) <= 1e-9
        assert abs(nlp.cost(P, prob, r["Z"][i], r["U"][i]) - r["cost"][i]) < 1e-12


def test_oracle_mlp_term_matches_torch(oracle_params):
    """gt_mpc value term: oracle forward tangents vs torch autograd on the reference's mlp class
    layout (model.py:14-51: Linear/tanh stack, fp64)."""
    import torch
    torch.manual_seed(3)
    dims = [6, 128, 128, 1]
    layers = []
    for i in range(3):
        layers.append(torch.nn.Linear(dims[i], dims[i + 1], dtype=torch.double))
        if i < 2:
            layers.append(torch.nn.Tanh())
    net = torch.nn.Sequential(*layers)
    weights = [(m.weight.detach().numpy(), m.bias.detach().numpy()) for m in net if isinstance(m, torch.nn.Linear)]
    rng = np.random.default_rng(5)
    A = rng.normal(size=(6, 6)); Wn = A @ A.T / 6 + np.eye(6)
    mu_f = rng.normal(size=6)
    term = nlp.MLPTerm(weights=weights, Wn=Wn, mu_f=mu_f, sigma_t=1.7, mu_t=-0.3)
    ctx = (20.0, 3.0, -2.0, 2.0)
    sv = torch.tensor([33.0, 4.2], dtype=torch.double, requires_grad=True)

    def V(sv_):
        xN = torch.stack([torch.tensor(ctx[0]), torch.tensor(ctx[1]), torch.tensor(ctx[2]), sv_[0] - ctx[0],
                          sv_[1] - ctx[1], torch.tensor(ctx[3] - ctx[2])]).double()
        return net(torch.tensor(Wn) @ (xN - torch.tensor(mu_f)))[0] * 1.7 - 0.3

    v = V(sv)
    g = torch.autograd.grad(v, sv, create_graph=True)[0]
    H = torch.stack([torch.autograd.grad(g[i], sv, retain_graph=True)[0] for i in range(2)])
    Vo, go, Ho = term.value(33.0, 4.2, ctx, order=2)
    assert abs(Vo - v.item()) < 1e-12 and np.max(np.abs(go - g.detach().numpy())) < 1e-12
    assert np.max(np.abs(Ho - H.detach().numpy())) < 1e-11
    # and the C oracle solves a gt_mpc problem to the same point as the numpy restatement
    P = oracle_params[10]
    pb = S.mid_episode(8, N=10, seed=31)
    term2 = nlp.MLPTerm(weights=weights, Wn=np.eye(6), mu_f=np.zeros(6), sigma_t=1.0, mu_t=0.0)
    r = c_oracle.COracle(P, term2).solve(pb.x0, pb.u_prev, pb.curv, pb.obs, nn_ctx=pb.nn_ctx)
    i = int(np.where(r["status"] == 0)[0][0])
    prob = nlp.Problem(x0=pb.x0[i], u_prev=pb.u_prev[i], curv=tuple(pb.curv[i]), obs=pb.obs[i], nn_ctx=tuple(pb.nn_ctx[i]))
    rp = solver.solve(P, prob, mlp=term2)
    assert rp.status == 0 and abs(rp.cost - r["cost"][i]) < 1e-9
