"""Closed-loop evaluation at scale: `n_samples` draws of the reference's start offsets (evaluate.py:91-94) x all
64 (scenario, rotation, order) variants, every episode 150 steps of 0.1 s, all episodes advanced together with
one batched GPU solve per step.  usage: closed_loop_scale.py [n_samples=32]  -> one JSON line"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from igt_mpc_int_b200 import episode
from igt_mpc_int_b200.planner import BatchSolver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
specs = [sp for i in range(n) for sp in episode.reference_episode_specs(sample=i)]
s = BatchSolver(N=40)
episode.run_closed_loop(s, specs[:64], steps=3, N=40)                       # warm up (allocations, first launches)
t0 = time.perf_counter()
r = episode.run_closed_loop(s, specs, steps=150, N=40, record_latency=True)
dt = time.perf_counter() - t0
E = len(specs)
print(json.dumps({
    "episodes": E, "vehicles": 2 * E, "steps": 150, "wall_s": round(dt, 2), "episodes_per_s": round(E / dt, 1),
    "mpc_solves_per_s_incl_host_glue": round(2 * E * 150 / dt), "solve_call_ms_p50": round(float(np.percentile(r.step_latency_ms, 50)), 2),
    "solve_call_ms_p90": round(float(np.percentile(r.step_latency_ms, 90)), 2),
    "collisions": int(r.collision.sum()), "deadlocks": int(r.deadlock.sum()), "goals_reached": int(r.goal.sum()),
    "failed_solve_fraction": round(float(r.num_infeasible.sum()) / (2 * E * 150), 4), "min_distance_m": round(float(r.min_distance.min()), 3)}))
s.close()
