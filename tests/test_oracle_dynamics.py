"""Oracle vs the reference's own numpy model (golden vectors) and FD checks of its Jacobians."""
import numpy as np
import pytest

from oracle import dynamics as D, c_oracle, nlp, ref_loader
from tests.util import relerr

# known answers captured by running the reference's numpy model
KAT_START = [18.0, 2.85, 18.0, 0.05, -0.02, 4.0, -0.02]      # planner order x,y,s,ey,epsi,v,psi
KAT_CURV = (19.3, 19.3 + 8.6 * np.pi / 2, 1 / 8.6)
KAT = {  # u -> (s, ey, epsi, v, x, y, psi)
    (1.0, 0.1): (18.404755297998683, 0.064036451728958749, -0.010920692723190885, 4.1000000000000014,
                 18.404758006271674, 2.8639596529752342, -0.010920692723190885),
    (-2.0, -0.3): (18.38320488496933, -0.02242236834018947, -0.0466719282400368, 3.8000000000000007,
                   18.38324520802603, 2.777789181738221, -0.0466719282400368),
    (0.0, 0.0): (18.399920002666633, 0.04200053332266677, -0.02, 4.0, 18.399920002666633, 2.842000533322667, -0.02),
}


def test_known_answers_python_oracle():
    for u, ref in KAT.items():
        z = D.frenet_rk4_step(KAT_START, u, KAT_CURV)
        got = z[[D.IS, D.IEY, D.IEPSI, D.IV, D.IX, D.IY, D.IPSI]]
        assert np.max(np.abs(got - np.array(ref))) < 1e-13
    Z = D.frenet_rollout(KAT_START, np.tile([0.5, 0.05], (10, 1)), KAT_CURV)
    ref10 = [22.24784989189113, 2.97193930626211, 22.194549678492486, -0.32534868168177816,
             -0.3079707937878214, 4.500000000000007, 0.02756391929916035]
    assert np.max(np.abs(Z[-1] - ref10)) < 1e-12
    zc = D.cartesian_euler_step([1.0, 2.8, 0.3, 4.0], [1.0, 0.2])
    assert np.max(np.abs(zc - [1.3682668732330494, 2.956139393104857, 0.3180471396593471, 4.1])) < 1e-14


def test_python_oracle_matches_golden(golden):
    g = golden["frenet_rk4"]
    for z0, u, c, zn in zip(g["z0"], g["u"], g["curv"], g["zn"]):
        assert np.max(np.abs(D.frenet_rk4_step(z0, u, tuple(c)) - zn)) < 1e-12
    for z0, U, c, Z in zip(g["rz0"][:8], g["ru"], g["rcurv"], g["rz"]):
        assert np.max(np.abs(D.frenet_rollout(z0, U, tuple(c)) - Z)) < 1e-10
    e = golden["cartesian_euler"]
    for z0, u, zn in zip(e["z0"], e["u"], e["zn"]):
        assert np.max(np.abs(D.cartesian_euler_step(z0, u) - zn)) < 1e-13


def test_c_oracle_matches_golden(golden, oracle_params):
    g = golden["frenet_rk4"]
    co = c_oracle.COracle(oracle_params[40])
    Z = co.rollout(g["rz0"], g["ru"], g["rcurv"])
    assert np.max(np.abs(Z - g["rz"])) < 1e-10
    # single steps: roll a horizon whose first input is the golden one
    U = np.zeros((len(g["u"]), 40, 2)); U[:, 0] = g["u"]
    Z1 = co.rollout(g["z0"], U, g["curv"])[:, 1]
    assert np.max(np.abs(Z1 - g["zn"])) < 1e-12
    e = golden["cartesian_euler"]
    U = np.zeros((len(e["u"]), 40, 2)); U[:, 0] = e["u"]
    Zc = co.rollout(e["z0"], U, model=1)[:, 1]
    assert np.max(np.abs(Zc - e["zn"])) < 1e-13


@pytest.mark.skipif(not ref_loader.available(), reason="reference not mounted")
def test_python_oracle_matches_reference_live():
    """Re-run the reference's numpy model in-process (only where /root/reference exists)."""
    ns = ref_loader.load()
    M = ns.kinematic_bicycle_model_frenet.KinematicBicycleModelFrenet(2.235, 2.235, 2.0, 0.1, discretization='rk4',
                                                                       mode='numpy', num_rk4_steps=4)
    VR, VA = ns.VehicleReference.VehicleReference, ns.VehicleAction.VehicleAction
    rng = np.random.default_rng(7)
    for _ in range(20):
        z = np.array([rng.uniform(0, 50), rng.uniform(0, 12), rng.uniform(0, 60), rng.uniform(-0.3, 0.3),
                      rng.uniform(-0.3, 0.3), rng.uniform(0, 5), rng.uniform(-3, 3)])
        u = (rng.uniform(-4, 3), rng.uniform(-0.6, 0.6))
        K = lambda s: D.curvature(s, KAT_CURV)
        o = M(VR({'s': z[2], 'ey': z[3], 'epsi': z[4], 'v': z[5], 'x': z[0], 'y': z[1], 'heading': z[6], 'K': K}),
              VA({'a': u[0], 'df': u[1]}))
        ref = np.array([o.x, o.y, o.s, o.ey, o.epsi, o.v, o.heading])
        assert np.max(np.abs(D.frenet_rk4_step(z, u, KAT_CURV) - ref)) < 1e-12


def test_jacobians_finite_difference(oracle_params):
    rng = np.random.default_rng(3)
    co = c_oracle.COracle(oracle_params[40])
    for trial in range(6):
        c = [(19.3, 32.8, 1 / 8.6), (10.7, 28.6, -1 / 11.4), (1e30, 1e30, 0.0)][trial % 3]
        z0 = np.array([rng.uniform(0, 50), rng.uniform(0, 12), rng.uniform(20.5, 27), rng.uniform(-0.2, 0.2),
                       rng.uniform(-0.2, 0.2), rng.uniform(0.5, 5), rng.uniform(-3, 3)])
        u0 = np.array([rng.uniform(-2, 2), rng.uniform(-0.5, 0.5)])
        _, A, B = D.frenet_rk4_step(z0, u0, c, want_jac=True)
        eps = 1e-6
        Afd = np.zeros((7, 7)); Bfd = np.zeros((7, 2))
        for i in range(7):
            e = np.zeros(7); e[i] = eps
            Afd[:, i] = (D.frenet_rk4_step(z0 + e, u0, c) - D.frenet_rk4_step(z0 - e, u0, c)) / (2 * eps)
        for i in range(2):
            e = np.zeros(2); e[i] = eps
            Bfd[:, i] = (D.frenet_rk4_step(z0, u0 + e, c) - D.frenet_rk4_step(z0, u0 - e, c)) / (2 * eps)
        assert np.max(np.abs(A - Afd)) < 1e-7 and np.max(np.abs(B - Bfd)) < 1e-7
        # C oracle Jacobians == python oracle Jacobians
        U = np.zeros((1, 40, 2)); U[0, 0] = u0
        _, Ac, Bc = co.rollout(z0[None], U, np.array([c]), jac=True)
        assert np.max(np.abs(Ac[0, 0] - A)) < 1e-12 and np.max(np.abs(Bc[0, 0] - B)) < 1e-12
    # Cartesian Euler
    z0 = np.array([1.0, 2.8, 0.3, 4.0]); u0 = np.array([1.0, 0.2])
    _, A, B = D.cartesian_euler_step(z0, u0, want_jac=True)
    for i in range(4):
        e = np.zeros(4); e[i] = 1e-6
        fd = (D.cartesian_euler_step(z0 + e, u0) - D.cartesian_euler_step(z0 - e, u0)) / 2e-6
        assert np.max(np.abs(A[:, i] - fd)) < 1e-8
    for i in range(2):
        e = np.zeros(2); e[i] = 1e-6
        fd = (D.cartesian_euler_step(z0, u0 + e) - D.cartesian_euler_step(z0, u0 - e)) / 2e-6
        assert np.max(np.abs(B[:, i] - fd)) < 1e-8
