"""Small-batch latency: throughput kernel vs the shared-memory latency path (igt_set_option "latency_path"),
cold and warm-started, host to host.  usage: latency_probe.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from igt_mpc_int_b200 import scenarios as S
from igt_mpc_int_b200.planner import BatchSolver

pb = S.mid_episode(152, N=40, seed=11)
s = BatchSolver(N=40)
for B in (1, 2, 8, 64, 148):
    res = {}
    for flag in (0, 1):
        s.set_option("latency_path", flag)
        a = (pb.x0[:B], pb.u_prev[:B], pb.curv[:B], pb.obs[:B])
        cold = s.solve_batch(*a)
        ui = np.nan_to_num(cold["u"])
        tc, tw = [], []
        for _ in range(15):
            t0 = time.perf_counter(); cold = s.solve_batch(*a); tc.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); warm = s.solve_batch(*a, u_init=ui); tw.append(time.perf_counter() - t0)
        res[flag] = (cold, warm, 1e3 * np.median(tc), 1e3 * np.median(tw))
    same = all(np.array_equal(res[0][i][k], res[1][i][k], equal_nan=True) for i in (0, 1) for k in ("status", "iters", "cost", "u", "x"))
    print("B=%3d cold %.2f -> %.2f ms, warm %.2f -> %.2f ms (throughput kernel -> latency path), iters cold max %d warm max %d, identical results: %s"
          % (B, res[0][2], res[1][2], res[0][3], res[1][3], res[1][0]["iters"].max(), res[1][1]["iters"].max(), same), flush=True)
s.close()
