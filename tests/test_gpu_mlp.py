"""gt_mpc value term: tcgen05 kernel and CUDA-core kernel vs the fp64 oracle; gt_mpc solves vs the oracle."""
import numpy as np
import pytest

from tests.util import relerr

pytestmark = pytest.mark.gpu


def _net(hidden, seed=2026):
    import torch
    torch.manual_seed(seed)
    dims = [6] + list(hidden) + [1]
    layers = [torch.nn.Linear(dims[i], dims[i + 1], dtype=torch.double) for i in range(len(dims) - 1)]
    return [(l.weight.detach().numpy().copy(), l.bias.detach().numpy().copy()) for l in layers]


def _term(weights, seeded_norm):
    from oracle import nlp
    if seeded_norm:
        rng = np.random.default_rng(5)
        A = rng.normal(size=(6, 6))
        Wn = 0.3 * (A @ A.T / 6 + np.eye(6))
        mu_f = rng.normal(size=6)
        return nlp.MLPTerm(weights=weights, Wn=Wn, mu_f=mu_f, sigma_t=1.7, mu_t=-0.3)
    return nlp.MLPTerm(weights=weights, Wn=np.eye(6), mu_f=np.zeros(6), sigma_t=1.0, mu_t=0.0)


def _as_dict(t):
    return dict(weights=t.weights, Wn=t.Wn, mu_f=t.mu_f, sigma_t=t.sigma_t, mu_t=t.mu_t)


@pytest.mark.parametrize("seeded_norm", [False, True])
def test_value_term_tensor_core_kernel_vs_oracle(seeded_norm):
    """fp32-accurate (bf16x3) tensor-core evaluation vs the fp64 oracle: 2e-5 abs on V (|V| = O(1)),
    relative 1e-4 (scale 1) on the derivatives.  Sizes: ragged (not a multiple of 256) and one CTA pass."""
    from igt_mpc_int_b200.planner import BatchSolver
    term = _term(_net((128, 128)), seeded_norm)
    s = BatchSolver(N=40, mlp=_as_dict(term))
    rng = np.random.default_rng(1)
    for B in (1000, 256, 37):
        sN, vN = rng.uniform(0, 70, B), rng.uniform(0, 5, B)
        ctx = np.stack([rng.uniform(0, 70, B), rng.uniform(0, 5, B), rng.integers(-8, 9, B).astype(float),
                        rng.integers(-8, 9, B).astype(float)], 1)
        ref = np.array([np.concatenate([[v], g, [H[0, 0], H[0, 1], H[1, 1]]]) for v, g, H in
                        (term.value(sN[i], vN[i], ctx[i], order=2) for i in range(B))])
        cc = s.mlp_value(sN, vN, ctx, tensor_cores=False)
        assert np.max(np.abs(cc - ref)) < 1e-10
        tc = s.mlp_value(sN, vN, ctx, tensor_cores=True)
        assert np.max(np.abs(tc[:, 0] - ref[:, 0])) < 2e-5
        assert np.max(relerr(tc[:, 1:], ref[:, 1:])) < 1e-4
    s.close()


@pytest.mark.parametrize("hidden", [(128, 128), (128, 128, 128)])
def test_gt_mpc_solve_matches_oracle(oracle_params, hidden):
    """gt_mpc mode (mpc.py:367-369) through the C ABI vs the oracle, random-init fp64 network
    (SURVEY 8(d) config 3): 2 hidden layers (sc1,2,4,5,8) and 3 hidden layers (sc3,6,7)."""
    from igt_mpc_int_b200.planner import BatchSolver
    from igt_mpc_int_b200 import scenarios as S
    from oracle import c_oracle
    term = _term(_net(hidden), False)
    N, B = 40, 256
    pb = S.mid_episode(B, N=N, seed=41)
    s = BatchSolver(N=N, mlp=_as_dict(term))
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs, nn_ctx=pb.nn_ctx)
    o = c_oracle.COracle(oracle_params[N], term, max_iter=s.params.max_iter, max_trials=s.params.max_trials).solve(pb.x0, pb.u_prev, pb.curv, pb.obs,
                                                                                  nn_ctx=pb.nn_ctx)
    # default: the exact (fp64, CTA-cooperative) value term -- north-star tolerances on EVERY converged problem
    ok = (r["status"] == 0) & (o["status"] == 0)
    assert ok.sum() > 0.7 * B
    assert np.mean(r["status"] == o["status"]) > 0.98
    assert np.max(relerr(r["cost"][ok], o["cost"][ok])) < 1e-4
    assert np.max(np.abs(r["u"][ok] - o["U"][ok])) < 1e-3
    assert np.max(r["viol"][ok]) <= 1e-6
    if hidden == (128, 128):
        # the optional tensor-core value term (fp32-accurate): same optima within the tolerances for all but the few
        # problems whose line search it steers elsewhere (documented in igt_mpc.h: not the parity path)
        s.set_option("tensor_core_mlp", 1)
        r2 = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs, nn_ctx=pb.nn_ctx)
        ok2 = (r2["status"] == 0) & (r["status"] == 0)
        assert np.mean((r2["status"] == 0) == (r["status"] == 0)) > 0.97
        assert np.quantile(relerr(r2["cost"][ok2], r["cost"][ok2]), 0.98) < 1e-4
    s.close()


def test_gt_mpc_full_size_batch_properties():
    """BASELINE config 3 at full size (gt_mpc, 16384 problems, random-init 6-128-128-1 network, exact cooperative value
    term with the speculative line search): rows satisfied to 1e-6, cost reproduced by the fp64 evaluation kernel
    to round-off, and bit-identical results for the reversed batch (which warp evaluates which candidate depends on
    the batch; the result does not)."""
    from igt_mpc_int_b200.planner import BatchSolver
    from igt_mpc_int_b200 import scenarios as S
    term = _term(_net((128, 128)), False)
    N, B = 40, 16384
    pb = S.mid_episode(B, N=N, seed=2026)
    s = BatchSolver(N=N, mlp=_as_dict(term))
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs, nn_ctx=pb.nn_ctx)
    ok = r["status"] == 0
    assert ok.mean() > 0.88
    assert np.max(r["viol"][ok]) <= 1e-6
    idx = np.where(ok)[0][::16]
    ev = s.evaluate(pb.x0[idx], pb.u_prev[idx], pb.curv[idx], pb.obs[idx], r["u"][idx], nn_ctx=pb.nn_ctx[idx])
    assert np.max(relerr(ev["cost"], r["cost"][idx])) < 1e-10 and np.max(ev["viol"]) <= 1e-6
    rr = s.solve_batch(pb.x0[::-1], pb.u_prev[::-1], pb.curv[::-1], pb.obs[::-1], nn_ctx=pb.nn_ctx[::-1])
    for k in ("status", "iters", "cost", "u"):
        assert np.array_equal(rr[k][::-1], r[k], equal_nan=True), k
    s.close()


@pytest.mark.parametrize("hidden,B", [((37, 65), 1000), ((128,), 19), ((5, 7, 3), 300), ((127, 128, 33), 517), ((16,), 1)])
def test_cooperative_value_term_ragged_shapes(hidden, B):
    """The CTA-cooperative evaluation streams the weights through shared memory in tiles of 16 inputs, two inputs per
    pass, two evaluations per warp and wave: odd and short layers (padded tile rows), widths that are no multiple of 32,
    batches that leave partial waves and warps without work -- bit-identical to the per-thread evaluation (same order of
    operations) and equal to the oracle's forward tangents."""
    from igt_mpc_int_b200.planner import BatchSolver
    term = _term(_net(hidden, seed=7), True)
    rng = np.random.default_rng(B)
    sN, vN = rng.uniform(0, 70, B), rng.uniform(0, 5, B)
    ctx = np.stack([rng.uniform(0, 70, B), rng.uniform(0, 5, B), rng.integers(-8, 9, B).astype(float),
                    rng.integers(-8, 9, B).astype(float)], 1)
    s = BatchSolver(N=40, mlp=_as_dict(term))
    coop = s.mlp_value(sN, vN, ctx, tensor_cores=2)
    per_thread = s.mlp_value(sN, vN, ctx, tensor_cores=0)
    s.close()
    assert np.array_equal(coop, per_thread)
    n = min(B, 40)
    ref = np.array([np.concatenate([[v], g, [H[0, 0], H[0, 1], H[1, 1]]]) for v, g, H in
                    (term.value(sN[i], vN[i], ctx[i], order=2) for i in range(n))])
    assert np.max(np.abs(coop[:n] - ref)) < 1e-11 * max(1.0, np.abs(ref).max())
