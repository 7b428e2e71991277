"""Solver results pinned against something that is not the solver ("parity first").

The reference's CasADi/IPOPT cannot run here, and the oracle and the CUDA kernels share one algorithm, so
"they agree" proves nothing about optimality.  These tests check the returned points of BASELINE
configs 2 and 3 (N = 40) with machinery that shares no code with the interior-point iLQR
(oracle/kkt_check.py):

  * an epsilon-KKT certificate per problem: finite-difference derivatives of the NLP functions exactly as
    mpc.py:177-373 writes them (oracle/nlp.py), multipliers from a bounded least-squares fit --
    stationarity <= 1e-5, complementarity <= 1e-7, primal violation <= 1e-6 (the north-star bound);
  * scipy.optimize SLSQP with exact first derivatives at N = 40, (a) started from the solver's own cold-start
    guess: same optimum on most problems (the NLP is non-convex: the others are distinct local optima, which
    (b) confirms) and (b) started from the solver's solution: it stays there.

The CPU tests judge the oracle's C port, the `gpu` tests judge the CUDA solver through the C ABI.
"""
import numpy as np
import pytest

from oracle import nlp, c_oracle, kkt_check as K
from igt_mpc_int_b200 import scenarios as S

STAT_TOL, COMP_TOL, VIOL_TOL = 1e-5, 1e-7, 1e-6


def _mlp(hidden=(128, 128), seed=2026):
    import bench
    return bench.random_mlp(hidden, seed)


def _certify(P, pb, U, status, mlp_term=None, gt=False, limit=None):
    worst = dict(stat=0.0, comp=0.0, viol=0.0)
    n = 0
    for i in np.where(status == 0)[0][:limit]:
        r = K.kkt_residual(P, K._problem(pb, i, gt), U[i], mlp=mlp_term, comp_tol=COMP_TOL)
        rel = r["stat"] / max(1.0, r["lam_max"])
        assert rel <= STAT_TOL, (i, r)
        assert r["comp"] <= COMP_TOL * (1 + 1e-9) and r["viol"] <= VIOL_TOL, (i, r)
        worst = dict(stat=max(worst["stat"], rel), comp=max(worst["comp"], r["comp"]), viol=max(worst["viol"], r["viol"]))
        n += 1
    return n, worst


def test_derivatives_used_by_the_checks_agree(oracle_params):
    """chain-rule Jacobians (used for SLSQP) vs central differences of the NLP functions (used for the certificate)"""
    P = oracle_params[40]
    pb = S.mid_episode(8, N=40, seed=3)
    U = c_oracle.COracle(P).initial_guess(pb.x0, pb.u_prev, pb.curv, pb.obs)
    for i in range(4):
        prob = K._problem(pb, i)
        R = K.Reduced(P, prob)
        Z, G = K.reduced_derivatives(P, prob, U[i])
        gJ, Jc = R.fd(U[i])
        assert np.max(np.abs(gJ - K.cost_gradient(P, prob, Z, U[i], G))) < 1e-6
        assert np.max(np.abs(Jc - K.rows_jacobian(P, prob, Z, U[i], G))) < 1e-5


def test_oracle_points_are_kkt_points_cfg2(oracle_params):
    """>= 256 problems of the config-2 distribution, N = 40."""
    P = oracle_params[40]
    pb = S.mid_episode(288, N=40, seed=2026)
    o = c_oracle.COracle(P, max_iter=60).solve(pb.x0, pb.u_prev, pb.curv, pb.obs)
    n, worst = _certify(P, pb, o["U"], o["status"])
    assert n >= 256, n
    print("cfg2 KKT certificate over %d problems: worst %s" % (n, worst))


def test_oracle_points_are_kkt_points_cfg3_gt_mpc(oracle_params):
    """>= 64 problems of config 3 (gt_mpc, random-init 6-128-128-1 value network)."""
    P = oracle_params[40]
    m = _mlp()
    term = nlp.MLPTerm(**m)
    pb = S.mid_episode(80, N=40, seed=2026)
    o = c_oracle.COracle(P, term, max_iter=60).solve(pb.x0, pb.u_prev, pb.curv, pb.obs, nn_ctx=pb.nn_ctx)
    n, worst = _certify(P, pb, o["U"], o["status"], mlp_term=term, gt=True)
    assert n >= 64, n
    print("cfg3 KKT certificate over %d problems: worst %s" % (n, worst))


def _pick_with_active_collision(P, o, pb, n_total, n_active):
    ok = np.where(o["status"] == 0)[0]
    dmin = np.array([np.sqrt(((o["Z"][i][1:, :2] - pb.obs[i][1:]) ** 2).sum(1)).min() for i in ok])
    act = ok[dmin < P.d_min + 1e-6]
    assert len(act) >= n_active, "too few problems with an active collision row in the sample"
    rest = [i for i in ok if i not in set(act[:n_active])]
    return list(act[:n_active]) + rest[:n_total - n_active]


def test_slsqp_agrees_at_N40(oracle_params):
    """32 config-2 problems (8 of them with an active collision row) against scipy SLSQP with exact gradients."""
    P = oracle_params[40]
    pb = S.mid_episode(160, N=40, seed=77)
    co = c_oracle.COracle(P, max_iter=60)
    o = co.solve(pb.x0, pb.u_prev, pb.curv, pb.obs)
    U0 = co.initial_guess(pb.x0, pb.u_prev, pb.curv, pb.obs)
    same = 0
    picks = _pick_with_active_collision(P, o, pb, 32, 8)
    for i in picks:
        prob = K._problem(pb, i)
        J = o["cost"][i]
        # (b) started at the solver's point an independent SQP method does not move: it is a local minimum
        res, Z, U = K.slsqp(P, prob, o["U"][i], maxiter=100)
        assert res.success and nlp.max_violation(P, prob, Z, U) <= 1e-8
        assert abs(res.fun - J) <= 1e-5 * max(1.0, abs(J)) and np.max(np.abs(U - o["U"][i])) <= 1e-3, (i, res.fun - J)
        # (a) from the cold-start guess
        res, Z, U = K.slsqp(P, prob, U0[i])
        if res.success and abs(res.fun - J) <= 1e-4 * max(1.0, abs(J)) and np.max(np.abs(U - o["U"][i])) <= 1e-3:
            same += 1
    assert same >= 0.8 * len(picks), same
    print("SLSQP from the cold-start guess reaches the solver's optimum on %d of %d problems" % (same, len(picks)))


# ------------------------------------------------------------------ the CUDA solver -------------
@pytest.mark.gpu
def test_gpu_points_are_kkt_points_cfg2(oracle_params):
    from igt_mpc_int_b200.planner import BatchSolver
    P = oracle_params[40]
    pb = S.mid_episode(288, N=40, seed=2026)
    s = BatchSolver(N=40)
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    s.close()
    n, worst = _certify(P, pb, np.nan_to_num(r["u"]), r["status"])
    assert n >= 256, n
    print("GPU cfg2 KKT certificate over %d problems: worst %s" % (n, worst))


@pytest.mark.gpu
@pytest.mark.parametrize("hidden,tc", [((128, 128), 1), ((128, 128), 0), ((128, 128, 128), 1)])
def test_gpu_points_are_kkt_points_cfg3_gt_mpc(oracle_params, hidden, tc):
    from igt_mpc_int_b200.planner import BatchSolver
    P = oracle_params[40]
    m = _mlp(hidden)
    pb = S.mid_episode(80, N=40, seed=2026)
    s = BatchSolver(N=40, mlp=m)
    s.set_option("tensor_core_mlp", tc)
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs, nn_ctx=pb.nn_ctx)
    s.close()
    n, worst = _certify(P, pb, np.nan_to_num(r["u"]), r["status"], mlp_term=nlp.MLPTerm(**m), gt=True)
    assert n >= 60, n
    print("GPU cfg3 %s tc=%d KKT certificate over %d problems: worst %s" % (hidden, tc, n, worst))


@pytest.mark.gpu
def test_gpu_slsqp_agrees_at_N40(oracle_params):
    from igt_mpc_int_b200.planner import BatchSolver
    P = oracle_params[40]
    pb = S.mid_episode(160, N=40, seed=77)
    s = BatchSolver(N=40)
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    s.close()
    o = dict(status=r["status"], Z=r["x"], U=r["u"], cost=r["cost"])
    for i in _pick_with_active_collision(P, o, pb, 32, 8):
        prob = K._problem(pb, i)
        res, Z, U = K.slsqp(P, prob, o["U"][i], maxiter=100)
        assert res.success and nlp.max_violation(P, prob, Z, U) <= 1e-8
        assert abs(res.fun - o["cost"][i]) <= 1e-5 * max(1.0, abs(o["cost"][i])) and np.max(np.abs(U - o["U"][i])) <= 1e-3
