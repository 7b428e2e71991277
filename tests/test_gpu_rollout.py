"""CUDA rollout kernels through the C ABI vs the fp64 oracle and the reference golden vectors."""
import numpy as np
import pytest

from tests.util import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solver():
    from igt_mpc_int_b200.planner import BatchSolver
    s = BatchSolver(N=40)
    yield s
    s.close()


def test_frenet_rollout_matches_reference_golden(solver, golden):
    """fp32 kernel vs the reference's fp64 numpy model: 1e-5 relative, metric
    |d| / max(|ref|, 1.0) (north star)."""
    g = golden["frenet_rk4"]
    Z = solver.rollout(g["rz0"], g["ru"], g["rcurv"])
    assert np.max(relerr(Z, g["rz"])) < 1e-5


def test_frenet_rollout_large_batch_vs_oracle(solver, oracle_params):
    from oracle import c_oracle
    from igt_mpc_int_b200 import scenarios as S
    rng = np.random.default_rng(11)
    B, N = 8192, 40
    pb = S.mid_episode(B, N=N, seed=31)
    a = np.clip(np.cumsum(rng.uniform(-0.09, 0.09, (B, N)), 1) + pb.u_prev[:, :1], -4, 3)
    d = np.clip(np.cumsum(rng.uniform(-0.07, 0.07, (B, N)), 1) + pb.u_prev[:, 1:], -0.6, 0.6)
    U = np.stack([a, d], 2)
    # fp32 inputs are what the kernel sees: round first so that the comparison isolates arithmetic
    z0 = pb.x0.astype(np.float32).astype(np.float64); U = U.astype(np.float32).astype(np.float64)
    cv = pb.curv.astype(np.float32).astype(np.float64)
    Zo, Ao, Bo = c_oracle.COracle(oracle_params[N]).rollout(z0, U, cv, jac=True)
    Z, A, Bm = solver.rollout(z0, U, cv, jac=True)
    e = relerr(Z, Zo).reshape(B, -1).max(1)
    # every rollout meets 1e-5: the kernel decides the pw_const branch in double from its compensated arc length, so a
    # stage state next to a curvature break-point takes the fp64 reference's branch (round 1: 0.1 % did not)
    print("fraction below 1e-5: %.6f, max %g" % (np.mean(e < 1e-5), e.max()))
    assert np.max(e) < 1e-5, "fraction below 1e-5: %g, max %g" % (np.mean(e < 1e-5), e.max())
    good = e < 1e-5
    # Jacobians: 1e-4 relative (metric |d| / max(|ref|, 1))
    assert np.max(relerr(A[good], Ao[good])) < 1e-4 and np.max(relerr(Bm[good], Bo[good])) < 1e-4


def test_cartesian_euler_rollout(solver, golden, oracle_params):
    from oracle import c_oracle
    e = golden["cartesian_euler"]
    U = np.zeros((len(e["u"]), 40, 2)); U[:, 0] = e["u"]
    Z, A, Bm = solver.rollout(e["z0"], U, jac=True, model=1)
    assert np.max(relerr(Z[:, 1], e["zn"])) < 1e-5
    Zo, Ao, Bo = c_oracle.COracle(oracle_params[40]).rollout(e["z0"], U, model=1, jac=True)
    assert np.max(np.abs(A - Ao)) < 1e-4 and np.max(np.abs(Bm - Bo)) < 1e-4


def test_rollout_edge_cases(solver):
    Z = solver.rollout(np.zeros((0, 7)), np.zeros((0, 40, 2)), np.zeros((0, 3)))
    assert Z.shape == (0, 41, 7)
    # linearity in v: zero steering, zero accel, straight road => s advances by v*dt*N exactly-ish
    z0 = np.zeros((4, 7)); z0[:, 5] = [0.0, 1.0, 2.5, 5.0]
    Z = solver.rollout(z0, np.zeros((4, 40, 2)), np.tile([1e30, 1e30, 0.0], (4, 1)))
    assert np.max(np.abs(Z[:, -1, 2] - z0[:, 5] * 4.0)) < 1e-5
    assert np.max(np.abs(Z[:, -1, 0] - z0[:, 5] * 4.0)) < 1e-5
