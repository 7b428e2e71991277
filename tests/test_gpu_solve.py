"""CUDA solver through the C ABI vs the fp64 oracle (same seeded problems)."""
import numpy as np
import pytest

from tests.util import relerr

pytestmark = pytest.mark.gpu

TIGHT = dict(max_iter=300, max_trials=0)   # product defaults already use the oracle's tolerances; only the iteration cap differs


def _solver(N, **kw):
    from igt_mpc_int_b200.planner import BatchSolver
    return BatchSolver(N=N, **kw)


@pytest.mark.parametrize("N,B", [(40, 512), (20, 128), (10, 128)])
def test_solver_matches_oracle_tight(oracle_params, N, B):
    """Same algorithm, same options, fp64 on both sides: identical outcomes per problem."""
    from oracle import c_oracle
    from igt_mpc_int_b200 import scenarios as S
    pb = S.mid_episode(B, N=N)
    s = _solver(N, **TIGHT)
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    o = c_oracle.COracle(oracle_params[N], max_iter=300).solve(pb.x0, pb.u_prev, pb.curv, pb.obs)
    assert np.mean(r["status"] == o["status"]) > 0.99
    ok = (o["status"] == 0) & (r["status"] == 0)
    assert ok.sum() > 0.8 * B
    assert np.max(relerr(r["cost"][ok], o["cost"][ok])) < 1e-4          # denominator max(|J|, 1)
    assert np.max(np.abs(r["u"][ok] - o["U"][ok])) < 1e-3
    assert np.max(r["viol"][ok]) <= 1e-6
    assert np.max(relerr(r["x"][ok], o["Z"][ok])) < 1e-3
    s.close()


def test_solver_default_options_parity(oracle_params):
    """Product defaults (what bench.py runs) against the tight oracle: north-star tolerances."""
    from oracle import c_oracle
    from igt_mpc_int_b200 import scenarios as S
    B, N = 1024, 40
    for gen, seed in ((S.mid_episode, 123), (S.episode_start, 2026)):
        pb = gen(B, N=N, seed=seed)
        s = _solver(N)
        r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
        o = c_oracle.COracle(oracle_params[N]).solve(pb.x0, pb.u_prev, pb.curv, pb.obs)
        ok = (o["status"] == 0) & (r["status"] == 0)
        assert ok.sum() > 0.85 * B
        assert np.mean((r["status"] == 0) == (o["status"] == 0)) > 0.98
        assert np.max(relerr(r["cost"][ok], o["cost"][ok])) < 1e-4
        assert np.max(np.abs(r["u"][ok] - o["U"][ok])) < 1e-3
        assert np.max(r["viol"][ok]) <= 1e-6
        # the device-side evaluation of the returned controls agrees with the oracle's
        ev = s.evaluate(pb.x0[ok], pb.u_prev[ok], pb.curv[ok], pb.obs[ok], r["u"][ok])
        oc, ov = c_oracle.COracle(oracle_params[N]).eval(pb.x0[ok], pb.u_prev[ok], pb.curv[ok], pb.obs[ok],
                                                          ev["x"], r["u"][ok])
        assert np.max(relerr(ev["cost"], oc)) < 1e-12 and np.max(np.abs(ev["viol"] - ov)) < 1e-12
        s.close()


def test_solver_edge_cases():
    from igt_mpc_int_b200 import scenarios as S
    from igt_mpc_int_b200 import _lib
    s = _solver(40)
    out = s.solve_batch(np.zeros((0, 7)), np.zeros((0, 2)), np.zeros((0, 3)), np.zeros((0, 41, 2)))
    assert out["x"].shape == (0, 41, 7)
    pb = S.mid_episode(8, N=40, seed=5)
    x0 = pb.x0.copy(); x0[0, 5] = 5.5; x0[1, 3] = 0.3
    r = s.solve_batch(x0, pb.u_prev, pb.curv, pb.obs)
    assert r["status"][0] == 2 and r["status"][1] == 2 and np.isnan(r["cost"][0])
    # ragged batch size (not a multiple of the block size) and warm start
    pb = S.mid_episode(104, N=40, seed=9).slice(0, 97)
    cold = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    ok = cold["status"] == 0
    warm = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs, u_init=np.nan_to_num(cold["u"]))
    assert np.all(warm["status"][ok] == 0)
    # restarted from its own solution a solve returns to it -- or, the NLP being non-convex and the warm start soft
    # (mu0_warm), to a better local minimum: never to a worse one
    e = relerr(warm["cost"][ok], cold["cost"][ok])
    assert np.mean(e < 1e-4) >= 0.95 and np.all(warm["cost"][ok] <= cold["cost"][ok] + 1e-4 * np.maximum(1.0, np.abs(cold["cost"][ok])))
    assert np.median(warm["iters"][ok]) < 0.9 * np.median(cold["iters"][ok])        # (soft warm start: mu0_warm 1e-3, igt_mpc.h)
    with pytest.raises(_lib.IgtError):
        s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs, nn_ctx=pb.nn_ctx)      # no MLP set
    with pytest.raises(ValueError):
        s.solve_batch(pb.x0, pb.u_prev[:, :1], pb.curv, pb.obs)
    s.close()


def test_solver_device_pointer_path_matches_host_path():
    import torch
    from igt_mpc_int_b200 import scenarios as S
    pb = S.mid_episode(256, N=40, seed=17)
    s = _solver(40)
    host = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    dev = s.solve_batch_device(t(pb.x0), t(pb.u_prev), t(pb.curv), t(pb.obs))
    torch.cuda.synchronize()
    assert np.array_equal(dev["status"].cpu().numpy(), host["status"])
    assert np.array_equal(dev["u"].cpu().numpy(), host["u"], equal_nan=True)
    assert s.launches >= 2
    s.close()


def test_fp32_solver_quality(oracle_params):
    """The fp32 arithmetic option: reported, looser (it is not the parity path)."""
    from oracle import c_oracle
    from igt_mpc_int_b200 import scenarios as S
    pb = S.mid_episode(512, N=40, seed=3)
    s = _solver(40, precision="f32")
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    o = c_oracle.COracle(oracle_params[40]).solve(pb.x0, pb.u_prev, pb.curv, pb.obs)
    ok = (o["status"] == 0) & (r["status"] == 0)
    assert ok.sum() > 0.8 * len(pb)
    assert np.median(relerr(r["cost"][ok], o["cost"][ok])) < 1e-3
    assert np.max(r["viol"][ok]) <= 1e-3
    s.close()


def test_planner_protocol_matches_reference_interface(oracle_params):
    """MPC_Planner mirror: constructor keywords, three-method protocol, return shapes
    (mpc.py:21-37, :241-294, :383-406)."""
    from igt_mpc_int_b200.planner import MPC_Planner
    from igt_mpc_int_b200 import geometry as G
    from oracle import c_oracle

    class Bag:
        def __init__(self, **kw):
            self.__dict__.update(kw)

    N = 40
    routes = ['12', '31']
    refs = []
    for r in routes:
        b0, b1, kv = G.curvature_params(r)
        K = np.zeros(151)
        if kv != 0.0:
            K[40:70] = kv
        refs.append({'K': K})
    agents = []
    for r, s0 in zip(routes, (5.0, 3.0)):
        x, y, th = G.frenet2global(s0, r)
        agents.append({'type': 'CAV', 'state': Bag(x=x, y=y, heading=th, v=0.0, s=s0, ey=0.0, epsi=0.0)})
    pl = MPC_Planner(N=N, dt=0.1, ca_radius=2.8, agents=agents, routes=routes, ref=refs, goals=None,
                     road_dim=(11.4, 50.0), ds_right=8.6, index=0, num_rk4_steps=4)
    pl.update_initial_condition(agents[0], Bag(a=0.1, df=0.0))
    preds = []
    for r, ag in zip(routes, agents):
        fc = G.constant_acceleration_forecast(ag['state'].s, 0.0, 0.1, r, N)
        preds.append([Bag(x=p[0], y=p[1], s=p[2], v=p[3], heading=0.0, ey=0.0, epsi=0.0) for p in fc])
    pl.update_predictions(preds, raw_preds=preds)
    x, u, ok = pl.solve()
    assert ok and x.shape == (7, N + 1) and u.shape == (2, N) and x.dtype == np.float64
    assert pl.solve_time > 0 and pl.sol.stats()["success"] and pl.x_sol_prev is x
    # same problem through the oracle
    obs = np.array([[p.x, p.y] for p in preds[1]])[None]
    o = c_oracle.COracle(oracle_params[N]).solve(pl._x0, pl._uprev, pl.curv, obs)
    assert o["status"][0] == 0 and abs(o["cost"][0] + 0.0) < 50
    assert np.max(np.abs(u.T - o["U"][0])) < 1e-3
    # warm start with the previous solution, like evaluate.py:478-482
    x2, u2, ok2 = pl.solve(x_sol_prev=x, u_sol_prev=u)
    assert ok2 and np.max(np.abs(u2 - u)) < 1e-3
    # infeasible initial state -> (None, None, False), never raises (mpc.py:402-406)
    agents[0]['state'].v = 9.0
    pl.update_initial_condition(agents[0], Bag(a=0.1, df=0.0))
    assert pl.solve() == (None, None, False)


def test_full_size_batch_properties(oracle_params):
    """BASELINE config 2 at full size (32768 problems, N = 40), through properties that do not need
    the oracle on every problem: (1) every converged solution satisfies the rows to 1e-6 and its
    cost / violation are reproduced by the independent evaluation kernel; (2) a problem's result
    does not depend on the batch it is solved in or on where it sits in it (persistent lanes, tail
    hand-over, speculative line search and CTA-wide phases only re-order independent work): the
    reversed batch and a 2048-problem slice give bit-identical results; (3) warm-starting from the
    solution converges again to the same cost; (4) a 512-problem sample agrees with the oracle."""
    from oracle import c_oracle
    from igt_mpc_int_b200 import scenarios as S
    B, N = 32768, 40
    pb = S.mid_episode(B, N=N, seed=2026)
    s = _solver(N)
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    ok = r["status"] == 0
    assert ok.mean() > 0.9
    assert np.max(r["viol"][ok]) <= 1e-6
    idx = np.where(ok)[0][::16]
    ev = s.evaluate(pb.x0[idx], pb.u_prev[idx], pb.curv[idx], pb.obs[idx], r["u"][idx])
    assert np.max(relerr(ev["cost"], r["cost"][idx])) < 1e-9 and np.max(ev["viol"]) <= 1e-6
    # (2) order / batch independence, bit for bit
    rr = s.solve_batch(pb.x0[::-1], pb.u_prev[::-1], pb.curv[::-1], pb.obs[::-1])
    for k in ("status", "iters", "cost", "u"):
        assert np.array_equal(rr[k][::-1], r[k], equal_nan=True), k
    sl = slice(5000, 7048)
    rs = s.solve_batch(pb.x0[sl], pb.u_prev[sl], pb.curv[sl], pb.obs[sl])
    for k in ("status", "iters", "cost", "u"):
        assert np.array_equal(rs[k], r[k][sl], equal_nan=True), k
    # (3) warm start from the solution
    w = np.where(ok)[0][:4096]
    rw = s.solve_batch(pb.x0[w], pb.u_prev[w], pb.curv[w], pb.obs[w], u_init=r["u"][w])
    assert np.mean(rw["status"] == 0) > 0.995
    okw = rw["status"] == 0
    e = relerr(rw["cost"][okw], r["cost"][w][okw])          # same KKT point again (a property, not the parity bar)
    assert np.quantile(e, 0.999) < 1e-4 and np.max(e) < 1e-3
    assert np.median(rw["iters"][okw]) <= np.median(r["iters"][w]) 
    # (4) oracle on a sample
    smp = np.arange(0, B, 64)
    o = c_oracle.COracle(oracle_params[N], max_iter=s.params.max_iter, max_trials=s.params.max_trials).solve(
        pb.x0[smp], pb.u_prev[smp], pb.curv[smp], pb.obs[smp])
    both = (o["status"] == 0) & (r["status"][smp] == 0)
    assert np.mean((o["status"] == 0) == (r["status"][smp] == 0)) > 0.98
    assert np.max(relerr(r["cost"][smp][both], o["cost"][both])) < 1e-4
    assert np.max(np.abs(r["u"][smp][both] - o["U"][both])) < 1e-3
    s.close()


@pytest.mark.parametrize("N", [10, 20])
def test_short_horizons_full_path(oracle_params, N):
    """BASELINE config 5 horizons N = 10, 20 (N = 40 is covered above) on 4096 problems vs the oracle."""
    from oracle import c_oracle
    from igt_mpc_int_b200 import scenarios as S
    B = 4096
    pb = S.mid_episode(B, N=N, seed=5)
    s = _solver(N)
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    o = c_oracle.COracle(oracle_params[N], max_iter=s.params.max_iter, max_trials=s.params.max_trials).solve(
        pb.x0, pb.u_prev, pb.curv, pb.obs)
    both = (o["status"] == 0) & (r["status"] == 0)
    assert both.mean() > 0.85
    assert np.mean((o["status"] == 0) == (r["status"] == 0)) > 0.99
    assert np.max(relerr(r["cost"][both], o["cost"][both])) < 1e-4
    assert np.max(np.abs(r["u"][both] - o["U"][both])) < 1e-3
    assert np.max(r["viol"][both]) <= 1e-6
    s.close()


def _vs_oracle(s, P, pb, frac=0.9):
    from oracle import c_oracle
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    o = c_oracle.COracle(P, max_iter=s.params.max_iter, max_trials=s.params.max_trials).solve(pb.x0, pb.u_prev, pb.curv, pb.obs)
    both = (o["status"] == 0) & (r["status"] == 0)
    assert np.mean((o["status"] == 0) == (r["status"] == 0)) >= 0.99
    assert both.mean() >= frac
    assert np.max(relerr(r["cost"][both], o["cost"][both])) < 1e-4
    assert np.max(np.abs(r["u"][both] - o["U"][both])) < 1e-3
    assert np.max(r["viol"][both]) <= 1e-6
    return r


def test_reference_initial_condition_distribution(oracle_params):
    """The reference's own start-of-episode distribution (evaluate.py:91-94, :404-419: s0 ~ U(0, 10.7), v0 = 0,
    u_prev = (0.1, 0)), 2048 problems over scenarios 1-8."""
    from igt_mpc_int_b200 import scenarios as S
    s = _solver(40)
    r = _vs_oracle(s, oracle_params[40], S.episode_start(2048, N=40, seed=2026), frac=0.97)
    assert np.median(r["iters"][r["status"] == 0]) <= 20
    s.close()


@pytest.mark.parametrize("N", [2, 5, 64])
def test_extreme_horizons(N):
    """Shortest and longest horizons igt_create accepts (2 <= N <= 64)."""
    from igt_mpc_int_b200 import scenarios as S
    from oracle import nlp
    base = nlp.Params(N=40)
    P = nlp.Params(N=N, cinf_A=base.cinf_A, cinf_b=base.cinf_b)
    s = _solver(N)
    _vs_oracle(s, P, S.mid_episode(64, N=N, seed=3), frac=0.7)
    s.close()


def test_tiny_and_ragged_batches_and_bad_inputs(oracle_params):
    """B = 1, 2, 33 (one lane, the closed-loop pair, one warp + 1): the same answers as inside a big batch; a
    non-finite state comes back with a failure status instead of hanging or poisoning its neighbours."""
    from igt_mpc_int_b200 import scenarios as S
    pb = S.mid_episode(64, N=40, seed=11)
    s = _solver(40)
    full = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    for B in (1, 2, 33):
        r = s.solve_batch(pb.x0[:B], pb.u_prev[:B], pb.curv[:B], pb.obs[:B])
        for k in ("status", "iters", "cost", "u"):
            assert np.array_equal(r[k], full[k][:B], equal_nan=True), (B, k)
    x0 = pb.x0.copy(); x0[3, 0] = np.nan; x0[5, 2] = np.inf
    r = s.solve_batch(x0, pb.u_prev, pb.curv, pb.obs)
    assert r["status"][3] != 0 and r["status"][5] != 0
    keep = np.ones(64, dtype=bool); keep[[3, 5]] = False
    for k in ("status", "iters", "cost"):
        assert np.array_equal(r[k][keep], full[k][keep], equal_nan=True), k
    s.close()


def test_latency_path_is_bit_identical_to_throughput_kernel():
    """Batches of at most one problem per SM run one CTA per problem with the workspace in shared memory
    (igt_set_option "latency_path", default on): the same arithmetic in a different layout -- identical bits,
    cold and warm-started, fp64 and fp32, including the batch that exactly fills the machine and a short horizon."""
    import torch
    from igt_mpc_int_b200 import scenarios as S
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    for N, prec in ((40, "f64"), (10, "f64"), (40, "f32")):
        pb = S.mid_episode(n_sm // 8 * 8 + 8, N=N, seed=23)
        s = _solver(N, precision=prec)
        for B in (1, 2, 33, n_sm):
            a = (pb.x0[:B], pb.u_prev[:B], pb.curv[:B], pb.obs[:B])
            out = {}
            for flag in (0, 1):
                s.set_option("latency_path", flag)
                cold = s.solve_batch(*a)
                warm = s.solve_batch(*a, u_init=np.nan_to_num(cold["u"]))
                out[flag] = (cold, warm)
            for i in (0, 1):
                for k in ("status", "iters", "cost", "viol", "u", "x"):
                    assert np.array_equal(out[0][i][k], out[1][i][k], equal_nan=True), (N, prec, B, i, k)
        s.close()


def test_concurrent_handles_and_streams_match_sequential():
    """The boundary is re-entrant (include/igt_mpc.h "Concurrency"): two handles with different horizons running at
    the same time on two streams, and two calls on ONE handle issued on two streams, give the results of the same
    calls made one after the other."""
    import torch
    from igt_mpc_int_b200 import scenarios as S
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    cases = {}
    for N, B, seed in ((40, 2048, 5), (20, 2048, 6), (10, 1024, 7)):
        pb = S.mid_episode(B, N=N, seed=seed)
        cases[N] = dict(s=_solver(N), args=tuple(t(a) for a in (pb.x0, pb.u_prev, pb.curv, pb.obs)))
    # sequential reference on the default stream
    ref = {}
    for N, c in cases.items():
        o = c["s"].solve_batch_device(*c["args"])
        torch.cuda.synchronize()
        ref[N] = {k: v.clone() for k, v in o.items()}
    # all three handles at once, each on its own stream, several times over (params of different N live in different slots)
    streams = {N: torch.cuda.Stream() for N in cases}
    outs = {N: [] for N in cases}
    for rep in range(3):
        for N, c in cases.items():
            with torch.cuda.stream(streams[N]):
                outs[N].append(c["s"].solve_batch_device(*c["args"]))
    torch.cuda.synchronize()
    for N in cases:
        for o in outs[N]:
            for k in ("status", "iters"):
                assert torch.equal(o[k], ref[N][k]), (N, k)
            ok = ref[N]["status"] == 0
            assert torch.equal(o["u"][ok], ref[N]["u"][ok]) and torch.equal(o["cost"][ok], ref[N]["cost"][ok])
    # one handle, two streams, different inputs: the second call must not start on the shared workspace early
    c = cases[40]
    pb2 = S.mid_episode(2048, N=40, seed=99)
    args2 = tuple(t(a) for a in (pb2.x0, pb2.u_prev, pb2.curv, pb2.obs))
    ref2 = {k: v.clone() for k, v in c["s"].solve_batch_device(*args2).items()}
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(s1):
        o1 = c["s"].solve_batch_device(*c["args"])
    with torch.cuda.stream(s2):
        o2 = c["s"].solve_batch_device(*args2)
    torch.cuda.synchronize()
    assert torch.equal(o1["status"], ref[40]["status"]) and torch.equal(o2["status"], ref2["status"])
    ok1, ok2 = ref[40]["status"] == 0, ref2["status"] == 0
    assert torch.equal(o1["u"][ok1], ref[40]["u"][ok1]) and torch.equal(o2["u"][ok2], ref2["u"][ok2])
    # more handles than constant-memory slots (8): the ninth shares a slot, still correct
    extra = [_solver(10, max_iter=30 + i) for i in range(9)]
    pb3 = S.mid_episode(256, N=10, seed=3)
    a3 = tuple(t(a) for a in (pb3.x0, pb3.u_prev, pb3.curv, pb3.obs))
    r_first = {k: v.clone() for k, v in extra[0].solve_batch_device(*a3).items()}
    for e in extra[1:]:
        e.solve_batch_device(*a3)
    r_again = extra[0].solve_batch_device(*a3)
    torch.cuda.synchronize()
    assert torch.equal(r_first["status"], r_again["status"]) and torch.equal(r_first["iters"], r_again["iters"])
    for e in extra:
        e.close()
    for c in cases.values():
        c["s"].close()


def test_acceptable_status_matches_oracle(oracle_params):
    """IGT_STATUS_ACCEPTABLE (6): solves that end in a failure at a point within the reference's own tolerances are
    the same on the GPU and in the oracle, meet the north star's violation bound, and never replace a tight
    convergence (with the exit disabled exactly the same problems converge)."""
    from oracle import c_oracle
    from igt_mpc_int_b200 import scenarios as S
    pb = S.mid_episode(4096, N=40, seed=2026)
    s = _solver(40)
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    s.close()
    o = c_oracle.COracle(oracle_params[40], max_iter=60).solve(pb.x0, pb.u_prev, pb.curv, pb.obs)
    assert (o["status"] == 6).sum() >= 3
    assert np.mean(r["status"] == o["status"]) > 0.995
    acc = (r["status"] == 6) & (o["status"] == 6)
    assert acc.sum() >= 0.6 * (o["status"] == 6).sum()
    assert np.max(r["viol"][r["status"] == 6]) <= 1e-6
    s = _solver(40, acc_tol=0.0)
    r0 = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    s.close()
    assert (r0["status"] == 6).sum() == 0 and np.array_equal(r0["status"] == 0, r["status"] == 0)
    assert np.all(np.isin(r0["status"][r["status"] == 6], (1, 3, 4)))


def test_obca_mode_matches_oracle(oracle_params):
    """collision_avoidance_type 'obca' (mpc.py:42-43, :170-175, :211-221) on the GPU: the dual-eliminated row
    (csrc/obca.cuh) through igt_solve_host with obs_psi, against the oracle's C port on 1024 problems, plus a
    parallel-faces case (the kink of the rectangle distance in psi)."""
    from oracle import nlp, c_oracle, obca
    from igt_mpc_int_b200 import scenarios as S
    N, B = 40, 1024
    P = nlp.Params(N=N, d_min=0.0, cinf_A=oracle_params[40].cinf_A, cinf_b=oracle_params[40].cinf_b)
    pb = S.mid_episode(B, N=N, seed=5)
    s = _solver(N, d_min=0.0)
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs, obs_psi=pb.obs_psi)
    o = c_oracle.COracle(P, max_iter=s.params.max_iter).solve(pb.x0, pb.u_prev, pb.curv, pb.obs, obs_psi=pb.obs_psi)
    assert np.mean(r["status"] == o["status"]) > 0.99
    ok = (o["status"] == 0) & (r["status"] == 0)
    assert ok.sum() > 0.85 * B
    assert np.max(relerr(r["cost"][ok], o["cost"][ok])) < 1e-4
    assert np.max(np.abs(r["u"][ok] - o["U"][ok])) < 1e-3
    assert np.max(r["viol"][ok]) <= 1e-6
    # the reported violation is the reference's OBCA row, and the rows do bind
    v = c_oracle.COracle(P).viol_obca(pb.x0[ok], pb.u_prev[ok], pb.curv[ok], pb.obs[ok], pb.obs_psi[ok], r["x"][ok], r["u"][ok])
    assert np.max(np.abs(np.maximum(v, 0) - r["viol"][ok])) < 1e-9
    ev = s.evaluate(pb.x0[ok], pb.u_prev[ok], pb.curv[ok], pb.obs[ok], r["u"][ok], obs_psi=pb.obs_psi[ok])
    assert np.max(np.abs(ev["viol"] - np.maximum(v, 0))) < 1e-9
    idx = np.where(ok)[0]
    dmin = np.array([min(obca.rect_distance(r["x"][i][k, [0, 1, 6]], np.array([pb.obs[i, k, 0], pb.obs[i, k, 1], pb.obs_psi[i, k]]))[0]
                         for k in range(1, N + 1)) for i in idx[:256]])
    assert (dmin < 1e-5).sum() >= 3 and dmin.min() > -1e-6
    # the circle rows (d_min = 5.6) on the same problems give different plans: the mode switch is real
    s2 = _solver(N)
    rc = s2.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    both = ok & (rc["status"] == 0)
    assert np.max(np.abs(rc["u"][both] - r["u"][both])) > 1e-2
    s2.close()
    # parallel faces: a car standing in the ego's lane, same heading -- nose-to-tail distance with a kink in psi
    x0 = np.array([[5.0, 2.8, 5.0, 0.0, 0.0, 3.0, 0.0]]); up = np.zeros((1, 2)); cv = np.array([[1e30, 1e30, 0.0]])
    ob = np.tile(np.array([14.0, 2.8]), (1, N + 1, 1)); op = np.zeros((1, N + 1))
    rk = s.solve_batch(x0, up, cv, ob, obs_psi=op)
    okk = c_oracle.COracle(P, max_iter=s.params.max_iter).solve(x0, up, cv, ob, obs_psi=op)
    assert rk["status"][0] == okk["status"][0]
    if rk["status"][0] == 0:
        assert abs(rk["cost"][0] - okk["cost"][0]) < 1e-4 * max(1.0, abs(okk["cost"][0]))
        assert rk["x"][0][-1, 0] <= 14.0 - 4.47 + 1e-6         # stops behind the obstacle's tail
    s.close()


def test_planner_obca_mode_protocol():
    """MPC_Planner(ca_type='obca') (mpc.py:40-43): same protocol, obstacle heading taken from the forecasts."""
    from types import SimpleNamespace as NS
    from igt_mpc_int_b200.planner import MPC_Planner
    from igt_mpc_int_b200 import reference_track as RT
    N = 20
    routes = ['13', '23']
    ref = [dict(K=np.zeros(151)), dict(K=np.zeros(151))]
    agents = [dict(type='mpc', state=NS(x=5.0, y=2.8, s=5.0, ey=0.0, epsi=0.0, v=1.0, heading=0.0)),
              dict(type='mpc', state=NS(x=0.0, y=-20.0, s=0.0, ey=0.0, epsi=0.0, v=0.0, heading=0.0))]
    pl = MPC_Planner(N=N, dt=0.1, agents=agents, goals=None, ca_radius=2.8, ref=ref, road_dim=(11.4, 50), routes=routes,
                     ds_right=8.6, index=0, num_rk4_steps=4, ca_type='obca')
    assert pl.d_min == 0
    pl.update_initial_condition(agents[0], NS(a=0.0, df=0.0))
    obst = [NS(x=13.0, y=3.4, s=0.0, ey=0.0, epsi=0.0, v=0.0, heading=0.9) for _ in range(N + 1)]
    own = [agents[0]['state']] * (N + 1)
    pl.update_predictions([own, obst])
    x, u, ok = pl.solve()
    assert ok and x.shape == (7, N + 1) and u.shape == (2, N)
    from oracle import obca
    d = min(obca.rect_distance(x[[0, 1, 6], k], np.array([13.0, 3.4, 0.9]))[0] for k in range(1, N + 1))
    assert d >= 1e-6 - 1e-7
