"""Closed-loop driver (evaluate.py:451-569 semantics).  CPU: host logic with the oracle as the solver.
GPU: identical outcomes (deadlock / collision / goal / infeasible counts) for every scenario variant."""
import numpy as np
import pytest

from igt_mpc_int_b200 import episode, geometry as G
from tests.oracle_backend import OracleBackend


def test_reference_episode_specs_follow_evaluate_py():
    specs = episode.reference_episode_specs()
    assert len(specs) == 8 * 4 * 2
    draws = [0.17893481367543618, 0.6399131657151546, 0.4672684011434851, 0.37050052710804804]
    sp = specs[0]
    assert sp.routes == ['13', '23']
    assert abs(sp.s0[0] - draws[0] * 10.7) < 1e-12 and abs(sp.s0[1] - draws[1] * 10.7) < 1e-12
    for sp in specs:
        assert abs(min(G.scenario_encoding(sp.routes))) in range(1, 9)


def test_closed_loop_host_logic_with_oracle_backend():
    specs = [s for s in episode.reference_episode_specs(scenarios=[1, 3])][:3]
    res = episode.run_closed_loop(OracleBackend(N=10, max_iter=60), specs, steps=12, N=10)
    assert res.z_cl.shape == (3, 2, 13, 7)
    # vehicles start at rest and accelerate at the jerk limit: a_0 = 0.1 + 0.09 (evaluate.py:419, mpc.py:301-312)
    assert np.all(res.solved[:, :, 0])
    assert np.allclose(res.u_cl[:, :, 0, 0], 0.19, atol=1e-5)
    assert np.all(np.diff(res.z_cl[:, :, :, 2], axis=2) >= -1e-12)          # s never decreases
    assert np.allclose(res.z_cl[:, :, 1, 5], 0.019, atol=1e-6)              # v_1 = dt * a_0
    assert not res.collision.any()


@pytest.mark.gpu
def test_closed_loop_outcomes_match_oracle_all_scenarios():
    """All 8 scenarios x 4 rotations x 2 agent orders, 150 steps, N = 40 (BASELINE config 1 enumerated
    over the reference's unseeded choices): same per-episode outcome with the GPU solver and with
    the oracle as the solver."""
    from igt_mpc_int_b200.planner import BatchSolver
    specs = episode.reference_episode_specs()
    gpu = BatchSolver(N=40)
    rg = episode.run_closed_loop(gpu, specs, steps=150, N=40, record_latency=True)
    ro = episode.run_closed_loop(OracleBackend(N=40, max_iter=gpu.params.max_iter, max_trials=gpu.params.max_trials), specs, steps=150, N=40)
    # outcome parity (the north-star bar): every episode ends the same way
    assert np.array_equal(rg.deadlock, ro.deadlock)
    assert np.array_equal(rg.collision, ro.collision)
    assert np.array_equal(rg.goal, ro.goal)
    assert not rg.collision.any()
    # trajectory parity: the two fp64 implementations do not round identically, and a solve that takes 40+
    # iterations through repeated line-search failures can end converged in one and given up in the other
    # (status 4 / iteration cap); from that step on the two closed loops see different plans.  Until then
    # they agree to rounding, and that must be the rule, not the exception.
    dz = np.abs(rg.z_cl - ro.z_cl).reshape(len(specs), -1).max(axis=1)
    assert np.mean(dz < 1e-3) >= 0.85, dz
    assert abs(int(rg.num_infeasible.sum()) - int(ro.num_infeasible.sum())) <= 0.05 * ro.num_infeasible.sum()
    same = rg.solved == ro.solved
    assert same.mean() > 0.99
    gpu.close()
