"""The reference's shipped value networks (game_theoretic_NN/models/V_GT_sc{1..8}.pt; 2 hidden layers for sc1, 2, 4, 5,
8, 3 for sc3, 6, 7 -- configs/sc*_config.yaml) as golden fixtures: tests/golden/value_nets.npz holds their weights and
the outputs of the reference's own `mlp` module (model.py:14-51) on seeded inputs (tests/golden/make_value_nets.py)."""
import os

import numpy as np
import pytest

from oracle import nlp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "value_nets.npz")


def shipped_net(k, Wn=None, mu_f=None, sigma_t=1.0, mu_t=0.0):
    g = np.load(GOLD)
    n = int(g["sc%d_n_layers" % k])
    weights = [(g["sc%d_W%d" % (k, i)], g["sc%d_b%d" % (k, i)]) for i in range(n)]
    return dict(weights=weights, Wn=np.eye(6) if Wn is None else Wn, mu_f=np.zeros(6) if mu_f is None else mu_f,
                sigma_t=sigma_t, mu_t=mu_t), g["sc%d_x" % k], g["sc%d_y" % k]


def raw_to_args(x):
    """network input x[B,6] -> (sN, vN, ctx[B,4]) with Wn = I, mu_f = 0: x = (s_tv, v_tv, e_tv, sN - s_tv, vN - v_tv, e_ego - e_tv)"""
    return x[:, 3] + x[:, 0], x[:, 4] + x[:, 1], np.stack([x[:, 0], x[:, 1], x[:, 2], x[:, 5] + x[:, 2]], 1)


@pytest.mark.parametrize("k", range(1, 9))
def test_oracle_value_term_reproduces_reference_network_outputs(k):
    net, x, y = shipped_net(k)
    assert len(net["weights"]) == (4 if k in (3, 6, 7) else 3)
    term = nlp.MLPTerm(**net)
    sN, vN, ctx = raw_to_args(x)
    v = np.array([term.value(sN[i], vN[i], ctx[i]) for i in range(len(x))])
    assert np.max(np.abs(v - y)) < 1e-12 * max(1.0, np.abs(y).max())


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(1, 9))
def test_gpu_value_term_on_shipped_networks(k):
    """the solver's cooperative fp64 evaluation (and the tensor-core kernel where it applies) against the reference
    module's outputs, plus derivatives against the oracle's forward tangents -- max error, every shipped network"""
    from igt_mpc_int_b200.planner import BatchSolver
    net, x, y = shipped_net(k)
    sN, vN, ctx = raw_to_args(x)
    s = BatchSolver(N=40, mlp=net)
    term = nlp.MLPTerm(**net)
    ref = np.array([np.concatenate([[v], g, [H[0, 0], H[0, 1], H[1, 1]]]) for v, g, H in
                    (term.value(sN[i], vN[i], ctx[i], order=2) for i in range(len(x)))])
    coop = s.mlp_value(sN, vN, ctx, tensor_cores=2)
    scale = max(1.0, np.abs(ref).max())
    assert np.max(np.abs(coop[:, 0] - y)) < 1e-12 * scale
    assert np.max(np.abs(coop - ref)) < 1e-11 * scale
    per_thread = s.mlp_value(sN, vN, ctx, tensor_cores=0)
    assert np.array_equal(per_thread, coop)                         # same operation order: bit-identical
    if len(net["weights"]) == 3:
        tc = s.mlp_value(sN, vN, ctx, tensor_cores=1)
        assert np.max(np.abs(tc[:, 0] - y)) < 2e-5 * scale
    s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 3, 6, 8])
def test_gt_mpc_solves_with_shipped_networks_match_oracle(oracle_params, k):
    """gt_mpc solves (mpc.py:367-369) with the shipped V_GT_sc{k} weights ("also run with shipped
    V_GT_sc{k}.pt"), exact cooperative value term, against the oracle: max errors, not quantiles."""
    from igt_mpc_int_b200.planner import BatchSolver
    from igt_mpc_int_b200 import scenarios as S
    from oracle import c_oracle
    from tests.util import relerr
    net, _, _ = shipped_net(k)
    N, B = 40, 256
    pb = S.mid_episode(B, N=N, scenarios=(k,), seed=41)
    s = BatchSolver(N=N, mlp=net)
    s.set_option("tensor_core_mlp", 0)
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs, nn_ctx=pb.nn_ctx)
    s.close()
    o = c_oracle.COracle(oracle_params[N], nlp.MLPTerm(**net), max_iter=60).solve(pb.x0, pb.u_prev, pb.curv, pb.obs, nn_ctx=pb.nn_ctx)
    assert np.mean(r["status"] == o["status"]) > 0.98
    ok = (r["status"] == 0) & (o["status"] == 0)
    assert ok.sum() > 0.6 * B
    assert np.max(relerr(r["cost"][ok], o["cost"][ok])) < 1e-4
    assert np.max(np.abs(r["u"][ok] - o["U"][ok])) < 1e-3
    assert np.max(r["viol"][ok]) <= 1e-6
