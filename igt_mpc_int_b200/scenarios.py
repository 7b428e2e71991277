"""Synthetic problem batches drawn from the fourwayint.yaml geometry.

The reference samples only the episode start (evaluate.py:91-94: s0 ~ U(0, 10.7), v0 = 0,
u_prev = (0.1, 0)).  Benchmarks and parity tests also need mid-episode problems, so two
generators are provided, both seeded numpy PCG64:

* `episode_start`  -- exactly the reference's initial-condition distribution;
* `mid_episode`    -- Config 2/4/5 of SURVEY 8(d): s0 ~ U(0,45), v0 ~ U(0,5), small lateral
  / heading errors, previous input inside its limits, the other vehicle forecast by the
  constant-acceleration model (common/constant_acceleration_model.py:66-80) and passed
  through filter_preds (common/utils.py:365-388).

Output is a `ProblemBatch` of plain fp64 numpy arrays in the solver's input layout.
"""
from dataclasses import dataclass
import math
import numpy as np

from . import geometry as G


@dataclass
class ProblemBatch:
    x0: np.ndarray        # [B, 7]  (x, y, s, ey, epsi, v, psi)      mpc.py:163
    u_prev: np.ndarray    # [B, 2]  (a, df)                          mpc.py:164
    curv: np.ndarray      # [B, 3]  (b0, b1, Kval)                   mpc.py:183-200
    obs: np.ndarray       # [B, N+1, 2] other vehicle's (x, y) forecast  mpc.py:223-226
    nn_ctx: np.ndarray    # [B, 4]  (s_tv, v_tv, e_tv, e_ego)        mpc.py:326-337
    scenario: np.ndarray  # [B] int32 scenario 1..8
    route: list           # [B] ego route string
    obs_psi: np.ndarray = None   # [B, N+1] other vehicle's heading forecast (OBCA rows, mpc.py:211-221); mid_episode only

    def __len__(self):
        return self.x0.shape[0]

    def slice(self, lo, hi):
        return ProblemBatch(self.x0[lo:hi], self.u_prev[lo:hi], self.curv[lo:hi], self.obs[lo:hi],
                            self.nn_ctx[lo:hi], self.scenario[lo:hi], self.route[lo:hi],
                            None if self.obs_psi is None else self.obs_psi[lo:hi])


def _route_xy(s, route):
    """Vectorised frenet2global on the nominal lane centres: s[...] -> x[...], y[...]."""
    s = np.asarray(s, dtype=np.float64)
    x0, y0, th0 = G.start_pose(route[0])
    t0 = np.array([math.cos(th0), math.sin(th0)])
    n0 = np.array([-t0[1], t0[0]])
    sgn = G.turn_sign(route)
    if sgn == 0:
        return x0 + s * t0[0], y0 + s * t0[1]
    b0, b1, K = G.curvature_params(route)
    r = 1.0 / abs(K)
    c = np.array([x0, y0]) + b0 * t0 + sgn * r * n0
    phi = np.clip((s - b0) / r, 0.0, math.pi / 2)
    pre = np.minimum(s, b0) - b0                       # <= 0 before the arc
    post = np.maximum(s - b1, 0.0)
    t1 = sgn * n0
    x = c[0] + r * (np.sin(phi) * t0[0] - sgn * np.cos(phi) * n0[0]) + pre * t0[0] + post * t1[0]
    y = c[1] + r * (np.sin(phi) * t0[1] - sgn * np.cos(phi) * n0[1]) + pre * t0[1] + post * t1[1]
    return x, y


def _input_feasible(v0, a_prev, dt=0.1, da=0.09, v_max=5.0, margin=0.05):
    """Can v stay inside [0, v_max] when a is driven to 0 at the jerk limit?  (Necessary for
    the NLP rows mpc.py:301-321 to be satisfiable; the terminal set C_inf expresses the same
    condition at the end of the horizon.)"""
    a = a_prev
    v = v0
    sign = 1.0 if a > 0 else -1.0
    while sign * a > 0:
        a = a - sign * da
        if sign * a < 0:
            a = 0.0
        v = v + dt * a
    return (margin <= v <= v_max - margin)


def mid_episode(B, N=40, scenarios=(1, 2, 3, 4, 5, 6, 7, 8), seed=2026, dt=0.1, d_min=5.6):
    """B problems, scenario-major blocks of equal size (B must be divisible by len(scenarios))."""
    per = B // len(scenarios)
    assert per * len(scenarios) == B, "B must be a multiple of the number of scenarios"
    x0 = np.empty((B, 7)); up = np.empty((B, 2)); curv = np.empty((B, 3))
    obs = np.empty((B, N + 1, 2)); ctx = np.empty((B, 4)); scen = np.empty(B, dtype=np.int32)
    obs_psi = np.zeros((B, N + 1))
    routes_out = []
    i = 0
    for sc in scenarios:
        rng = np.random.Generator(np.random.PCG64(seed + int(sc)))
        n_done = 0
        while n_done < per:
            rot = int(rng.integers(0, 4)); ego = int(rng.integers(0, 2))
            routes = G.scenario_routes(sc, rot, 0)
            enc = G.scenario_encoding(routes)
            s0 = rng.uniform(0, 45, 2); v0 = rng.uniform(0, 5, 2)
            ey0 = rng.uniform(-0.15, 0.15, 2); epsi0 = rng.uniform(-0.1, 0.1, 2)
            ap = rng.uniform(-1, 1, 2); dp = rng.uniform(-0.1, 0.1, 2)
            oth = 1 - ego
            if not _input_feasible(v0[ego], ap[ego]):
                continue
            xe, ye, the = G.frenet2global(s0[ego], routes[ego], ey=ey0[ego])
            fc = G.constant_acceleration_forecast(s0[oth], v0[oth], ap[oth], routes[oth], N, dt)
            if math.hypot(fc[0, 0] - xe, fc[0, 1] - ye) < d_min + 0.5:
                continue
            psi = the + epsi0[ego]
            x0[i] = (xe, ye, s0[ego], ey0[ego], epsi0[ego], v0[ego], psi)
            up[i] = (ap[ego], dp[ego])
            curv[i] = G.curvature_params(routes[ego])
            obs[i] = G.filter_obstacle((xe, ye), psi, fc[:, :2])
            if obs[i, 0, 0] != -20.0:                                # heading of the forecast poses (mpc.py:250-260: abs() on 32 / 41)
                hd = np.array([G.frenet2global(fc[k, 2], routes[oth])[2] for k in range(N + 1)])
                obs_psi[i] = np.abs(hd) if routes[oth] in ('32', '41') else hd
            ctx[i] = (fc[N, 2], fc[N, 3], enc[oth], enc[ego])
            scen[i] = sc
            routes_out.append(routes[ego])
            i += 1
            n_done += 1
    return ProblemBatch(x0, up, curv, obs, ctx, scen, routes_out, obs_psi)


def episode_start(B, N=40, scenarios=(1, 2, 3, 4, 5, 6, 7, 8), seed=2026, dt=0.1):
    """The reference's own initial-condition distribution (evaluate.py:91-94, :404-419):
    s0 = U(0,1) * max_start, v0 = 0, ey0 = epsi0 = 0, u_prev = (0.1, 0) in 'mpc' mode."""
    per = B // len(scenarios)
    assert per * len(scenarios) == B
    max_start = (G.ROAD_LENGTH - G.ROAD_WIDTH) / 2 - (G.ROAD_WIDTH - G.CA_RADIUS)   # evaluate.py:49
    x0 = np.empty((B, 7)); up = np.empty((B, 2)); curv = np.empty((B, 3))
    obs = np.empty((B, N + 1, 2)); ctx = np.empty((B, 4)); scen = np.empty(B, dtype=np.int32)
    routes_out = []
    i = 0
    for sc in scenarios:
        rng = np.random.Generator(np.random.PCG64(seed + int(sc)))
        for _ in range(per):
            rot = int(rng.integers(0, 4)); ego = int(rng.integers(0, 2))
            routes = G.scenario_routes(sc, rot, 0)
            enc = G.scenario_encoding(routes)
            s0 = rng.random(2) * max_start
            oth = 1 - ego
            xe, ye, the = G.frenet2global(s0[ego], routes[ego])
            fc = G.constant_acceleration_forecast(s0[oth], 0.0, 0.1, routes[oth], N, dt)
            x0[i] = (xe, ye, s0[ego], 0.0, 0.0, 0.0, the)
            up[i] = (0.1, 0.0)
            curv[i] = G.curvature_params(routes[ego])
            obs[i] = G.filter_obstacle((xe, ye), the, fc[:, :2])
            ctx[i] = (fc[N, 2], fc[N, 3], enc[oth], enc[ego])
            scen[i] = sc
            routes_out.append(routes[ego])
            i += 1
    return ProblemBatch(x0, up, curv, obs, ctx, scen, routes_out)
