"""TEST HELPER: the CPU oracle behind the BatchSolver interface (solve_batch / evaluate), so that the
closed-loop driver can be run with the oracle as the solver and compared with the GPU-backed run."""
import numpy as np

from oracle import nlp, c_oracle


class OracleBackend:
    def __init__(self, N=40, mlp=None, **options):
        self.N = N
        self.P = nlp.Params(N=N)
        self.options = options
        term = None if mlp is None else nlp.MLPTerm(**mlp)
        self.cold = c_oracle.COracle(self.P, term, **options)
        warm = dict(options); warm.update(mu0=1e-3, y_init_min=1e-2)      # igt_params mu0_warm / y_init_min_warm
        self.warm = c_oracle.COracle(self.P, term, **warm)
        self.plain = c_oracle.COracle(self.P, **options)                  # 'mpc' cost: evaluate() without a value-network context

    def solve_batch(self, x0, u_prev, curv, obs_xy, nn_ctx=None, u_init=None):
        co = self.warm if u_init is not None else self.cold
        r = co.solve(x0, u_prev, curv, obs_xy, nn_ctx=nn_ctx, u_init=u_init)
        return dict(x=r["Z"], u=r["U"], cost=r["cost"], viol=r["viol"], status=r["status"], iters=r["iters"])

    def evaluate(self, x0, u_prev, curv, obs_xy, u, nn_ctx=None):
        co = self.cold if nn_ctx is not None else self.plain
        Z = co.rollout(x0, u, curv)
        cost, viol = co.eval(x0, u_prev, curv, obs_xy, Z, u, nn_ctx=nn_ctx)
        return dict(cost=cost, viol=viol, x=Z)
