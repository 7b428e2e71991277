"""Fixed cost of one small solve call (B = 2, host to host): iteration cap 1 vs the default.  usage: overhead_probe.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from igt_mpc_int_b200 import scenarios as S
from igt_mpc_int_b200.planner import BatchSolver
pb = S.mid_episode(8, N=40, seed=11)
a = (pb.x0[:2], pb.u_prev[:2], pb.curv[:2], pb.obs[:2])
for mi in (1, 2, 4, 60):
    s = BatchSolver(N=40, max_iter=mi)
    r = s.solve_batch(*a)
    ui = np.nan_to_num(r["u"])
    ts = []
    for _ in range(50):
        t0 = time.perf_counter(); r = s.solve_batch(*a, u_init=ui); ts.append(time.perf_counter() - t0)
    print("max_iter=%2d  B=2 warm host-to-host p50 %.3f ms  (iters %s)" % (mi, 1e3 * np.median(ts), r["iters"]), flush=True)
    s.close()
