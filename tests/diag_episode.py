"""Closed-loop parity diagnostic: where do the GPU and the oracle disagree on a solve's success?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from igt_mpc_int_b200 import episode
from igt_mpc_int_b200.planner import BatchSolver
from tests.oracle_backend import OracleBackend

specs = episode.reference_episode_specs()
gpu = BatchSolver(N=40)
rg = episode.run_closed_loop(gpu, specs, steps=150, N=40)
ro = episode.run_closed_loop(OracleBackend(N=40, max_iter=gpu.params.max_iter, max_trials=gpu.params.max_trials), specs, steps=150, N=40)
d = np.argwhere(rg.solved != ro.solved)
print("solved mismatches:", len(d))
for e, i, t in d[:20]:
    print("episode", e, specs[e].routes, "vehicle", i, "step", t, "gpu", rg.solved[e, i, t], "oracle", ro.solved[e, i, t])
print("max |z diff|", np.abs(rg.z_cl - ro.z_cl).max(), "infeasible gpu/oracle", rg.num_infeasible.sum(), ro.num_infeasible.sum())
# status / iteration counts of the first mismatching solve on both sides, from identical inputs (replayed with the oracle's states)
first = {}
for e, i, t in d:
    first.setdefault(e, t)
for e, t in first.items():
    dz = np.abs(rg.z_cl[e, :, :t + 1] - ro.z_cl[e, :, :t + 1]).max()
    print("episode", e, "first mismatch at step", t, "max |z diff| before it", dz)

# replay: feed the oracle run's inputs at the first mismatching step of each episode to both solvers
class Rec:
    def __init__(self, inner): self.inner, self.calls = inner, []
    def solve_batch(self, *a, **k):
        r = self.inner.solve_batch(*a, **k); self.calls.append((a, k, r)); return r
    def evaluate(self, *a, **k): return self.inner.evaluate(*a, **k)
for e, t in list(first.items())[:2]:
    rec = Rec(OracleBackend(N=40, max_iter=gpu.params.max_iter, max_trials=gpu.params.max_trials))
    episode.run_closed_loop(rec, specs[e:e + 1], steps=t + 1, N=40)
    for (a, k, r) in rec.calls[-2:]:
        g = gpu.solve_batch(*a, **k)
        print("episode", e, "step", t, "B", len(a[0]), "warm", k.get("u_init") is not None, "oracle status", r["status"], "iters", r["iters"],
              "| gpu status", g["status"], "iters", g["iters"], "cost diff", np.abs(g["cost"] - r["cost"]))
