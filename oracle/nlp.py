"""ORACLE (test infrastructure) -- fp64 restatement of the NLP assembled by
MPC_Planner.__init__ (reference mpc.py:21-160).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Decision variables (mpc.py:162-164): x[7, N+1] = rows [x, y, s, ey, epsi, v, psi],
u[2, N] = rows [a, df].  Arrays here are stored stage-major: Z[N+1, 7], U[N, 2].

Rows, in the order the reference adds them (mpc.py:147-159):
  box         mpc.py:314-321   k=0..N-1: 0<=v_k<=5, -4<=a_k<=3, -1<=df_k<=1
  initial     mpc.py:228-239   x[:,0] == x0 (fabs on heading for routes 32/41 is applied by
                               the host before the numbers reach the solver)
  rate        mpc.py:301-312   |a_k-a_{k-1}|<=0.09, |df_k-df_{k-1}|<=0.07, u_{-1}=u_prev
  dynamics    mpc.py:201-209   x[:,k+1] == F(x[:,k], u[:,k])
  ey          mpc.py:296-299   |ey_k| <= 0.2, k=0..N
  terminal    mpc.py:177-180   C_inf.A [v_{N-1}; a_{N-1}] <= C_inf.b   (index N-1, not N)
  collision   mpc.py:223-226   d_min^2 - |p_k - o_k|^2 <= 0, k=1..N, d_min = 5.6
Cost (mpc.py:356-373):  sum_{k<=N}(epsi_k^2 + ey_k^2) + 0.05 sum_{k<N}(a_k^2 + df_k^2)
  - (s_N - s_0)                                   ['mpc' mode]
  - (V(W (x_N - mu_f)) * sigma_t + mu_t)          ['gt_mpc' mode, mpc.py:367-369]
"""
from dataclasses import dataclass, field
import numpy as np

from . import dynamics as D
from . import cinf as _cinf


@dataclass
class Params:
    """Effective constants of the reference NLP (mpc.py:47-62 override the YAML files)."""
    N: int = 40                 # mpc.yaml:6
    dt: float = 0.1             # mpc.yaml:7
    n_rk: int = 4               # evaluate.py:109
    v_min: float = 0.0          # mpc.py:56
    v_max: float = 5.0          # mpc.py:57
    a_min: float = -4.0         # mpc.py:58
    a_max: float = 3.0          # mpc.py:59
    df_max: float = 1.0         # mpc.py:61
    ey_lim: float = 0.2         # mpc.py:60
    jerk_limit: float = 0.9     # mpc.py:55
    steer_rate: float = 0.7     # mpc.py:54
    d_min: float = 5.6          # mpc.py:45  (2 * ca_radius)
    w_u: float = 0.05           # mpc.py:362
    cinf_A: np.ndarray = field(default=None, repr=False)
    cinf_b: np.ndarray = field(default=None, repr=False)

    def __post_init__(self):
        if self.cinf_A is None:
            A, b, _ = _cinf.cinf_vertices_hrep(self.dt, self.jerk_limit, 5.0, 3.0, -1.0, -4.0)
            self.cinf_A, self.cinf_b = A, b

    @property
    def da_max(self):
        return self.dt * self.jerk_limit

    @property
    def ddf_max(self):
        return self.dt * self.steer_rate


@dataclass
class MLPTerm:
    """gt_mpc terminal value term, mpc.py:326-354, :367-369 and model.py:14-67.
    weights: list of (W[out,in], b[out]) fp64; tanh between layers.
    feature x_N = [s_tv, v_tv, e_tv, s_N - s_tv, v_N - v_tv, e_ego - e_tv]."""
    weights: list
    Wn: np.ndarray            # 6x6 whitening matrix  (mpc.py:116)
    mu_f: np.ndarray          # feature mean [6]
    sigma_t: float
    mu_t: float

    def value(self, sN, vN, ctx, order=0):
        """ctx = (s_tv, v_tv, e_tv, e_ego).  Returns V*sigma_t + mu_t and, for order>=1 /
        order>=2, its gradient[2] / Hessian[2,2] w.r.t. (s_N, v_N) by forward tangents."""
        s_tv, v_tv, e_tv, e_ego = ctx
        xN = np.array([s_tv, v_tv, e_tv, sN - s_tv, vN - v_tv, e_ego - e_tv])
        h = self.Wn @ (xN - self.mu_f)
        # tangents of h w.r.t. (sN, vN): columns 3 and 4 of Wn
        t1 = [self.Wn[:, 3].copy(), self.Wn[:, 4].copy()]
        t2 = {(0, 0): np.zeros(6), (0, 1): np.zeros(6), (1, 1): np.zeros(6)}
        L = len(self.weights)
        for li, (W, b) in enumerate(self.weights):
            h = W @ h + b
            t1 = [W @ t for t in t1]
            t2 = {k: W @ t for k, t in t2.items()}
            if li < L - 1:
                y = np.tanh(h)
                d1 = 1.0 - y * y
                d2 = -2.0 * y * d1
                t2 = {(i, j): d1 * t2[(i, j)] + d2 * t1[i] * t1[j] for (i, j) in t2}
                t1 = [d1 * t for t in t1]
                h = y
        V = float(h[0]) * self.sigma_t + self.mu_t
        if order == 0:
            return V
        g = np.array([t1[0][0], t1[1][0]]) * self.sigma_t
        if order == 1:
            return V, g
        H = np.array([[t2[(0, 0)][0], t2[(0, 1)][0]], [t2[(0, 1)][0], t2[(1, 1)][0]]]) * self.sigma_t
        return V, g, H


@dataclass
class Problem:
    x0: np.ndarray              # [7]  planner order
    u_prev: np.ndarray          # [2]
    curv: tuple                 # (b0, b1, Kval)
    obs: np.ndarray             # [N+1, 2] obstacle (x, y) forecast; row 0 unused (mpc.py:224)
    nn_ctx: tuple = None        # (s_tv, v_tv, e_tv, e_ego) for gt_mpc
    u_init: np.ndarray = None   # [N, 2] optional warm start (mpc.py:386-389)
    obs_psi: np.ndarray = None  # [N+1] obstacle heading forecast: selects the OBCA collision rows (ca_type='obca',
                                # mpc.py:211-221, d_min = 0) in their dual-eliminated form, see oracle/obca.py;
                                # python oracle only so far (the C port and the CUDA solver implement 'circle')


OBCA_MARGIN = 1e-6              # mpc.py:216: ... >= d_min + 1e-6 with d_min = 0 (mpc.py:42-43)


def obca_rows(prob: Problem, Z):
    """OBCA rows k = 1..N in eliminated form: margin - dist(ego rectangle_k, obstacle rectangle_k) <= 0."""
    from . import obca
    return np.array([OBCA_MARGIN - obca.rect_distance(Z[k, [D.IX, D.IY, D.IPSI]],
                                                      np.array([prob.obs[k, 0], prob.obs[k, 1], prob.obs_psi[k]]))[0]
                     for k in range(1, Z.shape[0])])


def cost(P: Params, prob: Problem, Z, U, mlp: MLPTerm = None):
    """mpc.py:356-373."""
    J = np.sum(Z[:, D.IEPSI] ** 2 + Z[:, D.IEY] ** 2) + P.w_u * np.sum(U ** 2)
    if mlp is None:
        J -= Z[-1, D.IS] - Z[0, D.IS]
    else:
        J -= mlp.value(Z[-1, D.IS], Z[-1, D.IV], prob.nn_ctx)
    return float(J)


def inequality_rows(P: Params, prob: Problem, Z, U):
    """All inequality rows c(x,u) <= 0 of the reference NLP, written exactly as in mpc.py
    (collision row in squared-distance units).  Returns a flat array."""
    N = P.N
    rows = []
    v, ey = Z[:, D.IV], Z[:, D.IEY]
    a, df = U[:, 0], U[:, 1]
    rows += [P.v_min - v[:N], v[:N] - P.v_max, P.a_min - a, a - P.a_max, -P.df_max - df, df - P.df_max]
    up = np.vstack([prob.u_prev[None, :], U[:-1]])
    da, ddf = a - up[:, 0], df - up[:, 1]
    rows += [-P.da_max - da, da - P.da_max, -P.ddf_max - ddf, ddf - P.ddf_max]
    rows += [ey - P.ey_lim, -P.ey_lim - ey]
    rows += [P.cinf_A @ np.array([v[N - 1], a[N - 1]]) - P.cinf_b]
    if prob.obs_psi is not None:
        rows += [obca_rows(prob, Z)]
    else:
        dp = Z[1:, :2] - prob.obs[1:]
        rows += [P.d_min ** 2 - np.sum(dp * dp, axis=1)]
    return np.concatenate(rows)


def max_violation(P: Params, prob: Problem, Z, U):
    return float(max(0.0, np.max(inequality_rows(P, prob, Z, U))))


def dynamics_defect(P: Params, prob: Problem, Z, U):
    """max |x_{k+1} - F(x_k,u_k)| and |x_0 - x0|  (mpc.py:201-209, :228-239)."""
    d = np.max(np.abs(Z[0] - prob.x0))
    for k in range(P.N):
        d = max(d, np.max(np.abs(Z[k + 1] - D.frenet_rk4_step(Z[k], U[k], prob.curv, P.dt, P.n_rk))))
    return float(d)


GUESS_SPEEDS = (5.0, 3.5, 2.0, 1.0, 0.0)     # candidate cruise speeds of the cold-start rule
GUESS_RHO = 100.0                            # weight of constraint violation when ranking them


def tracking_rollout(P: Params, prob: Problem, v_target):
    """Roll a rate-limited path-tracking controller towards cruise speed v_target:
      a_k  = clip(0.5 (v_target - v_k),  a_{k-1} -+ 0.09, [a_min, a_max])
      df_k = clip(df_ff(K(s_k)) - 0.3 ey_k - 0.8 epsi_k, df_{k-1} -+ 0.07, [-df_max, df_max])
    with df_ff = atan(2 tan(asin(K l_r))) the steady-state steering on curvature K.
    Returns U[N,2], Z[N+1,7]."""
    N = P.N
    U = np.zeros((N, 2))
    Z = np.zeros((N + 1, 7))
    z = np.array(prob.x0, dtype=np.float64)
    Z[0] = z
    up = np.array(prob.u_prev, dtype=np.float64)
    for k in range(N):
        K = D.curvature(z[D.IS], prob.curv)
        a_des = 0.5 * (v_target - z[D.IV])
        a = min(max(a_des, up[0] - P.da_max), up[0] + P.da_max)
        a = min(max(a, P.a_min), P.a_max)
        df_ff = np.arctan(2.0 * np.tan(np.arcsin(K * D.L_R)))
        d_des = df_ff - 0.3 * z[D.IEY] - 0.8 * z[D.IEPSI]
        d = min(max(d_des, up[1] - P.ddf_max), up[1] + P.ddf_max)
        d = min(max(d, -P.df_max), P.df_max)
        U[k] = (a, d)
        z = D.frenet_rk4_step(z, U[k], prob.curv, P.dt, P.n_rk)
        Z[k + 1] = z
        up = U[k]
    return U, Z


def guess_merit(P: Params, prob: Problem, Z, U):
    """J('mpc' cost) + GUESS_RHO * sum of row violations, collision measured in metres."""
    J = np.sum(Z[:, D.IEPSI] ** 2 + Z[:, D.IEY] ** 2) + P.w_u * np.sum(U ** 2) - (Z[-1, D.IS] - Z[0, D.IS])
    N = P.N
    v, ey = Z[:, D.IV], Z[:, D.IEY]
    viol = np.sum(np.maximum(0.0, v[1:N] - P.v_max)) + np.sum(np.maximum(0.0, P.v_min - v[1:N]))
    viol += np.sum(np.maximum(0.0, np.abs(ey[1:]) - P.ey_lim))
    if prob.obs_psi is not None:
        viol += np.sum(np.maximum(0.0, obca_rows(prob, Z)))
    else:
        dist = np.sqrt(np.sum((Z[1:, :2] - prob.obs[1:]) ** 2, axis=1))
        viol += np.sum(np.maximum(0.0, P.d_min - dist))
    viol += np.sum(np.maximum(0.0, P.cinf_A @ np.array([v[N - 1], U[N - 1, 0]]) - P.cinf_b))
    return float(J + GUESS_RHO * viol)


def initial_guess(P: Params, prob: Problem):
    """Cold-start rule shared by the oracle and the CUDA solver (DESIGN.md "initial guess").

    The reference cold-starts IPOPT from Opti's all-zero primal guess (mpc.py:386-389 sets an
    initial point only when a previous solution exists).  A shooting method needs a control
    sequence, and the collision rows make the problem non-convex (pass first / yield), so
    both solvers roll the tracking controller above for each cruise speed in GUESS_SPEEDS and
    keep the candidate with the smallest guess_merit (first one wins ties)."""
    best = None
    for vt in GUESS_SPEEDS:
        U, Z = tracking_rollout(P, prob, vt)
        m = guess_merit(P, prob, Z, U)
        if best is None or m < best[0]:
            best = (m, U)
    return best[1]
