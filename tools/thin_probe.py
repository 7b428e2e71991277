"""Thin-regime probe: B problems (one per SM at B = 148) on the THROUGHPUT kernel (latency path off) or on the latency
path.  usage: thin_probe.py [B] [latency_path 0|1] [reps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from igt_mpc_int_b200 import scenarios as S
from igt_mpc_int_b200.planner import BatchSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
flag = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pb = S.mid_episode((B + 7) // 8 * 8, N=40, seed=11)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
x0, up, cv, ob = t(pb.x0[:B]), t(pb.u_prev[:B]), t(pb.curv[:B]), t(pb.obs[:B])
s = BatchSolver(N=40)
s.set_option("latency_path", flag)
out = s.solve_batch_device(x0, up, cv, ob); torch.cuda.synchronize()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = s.solve_batch_device(x0, up, cv, ob, out=out); e1.record(); torch.cuda.synchronize()
    it = out["iters"].cpu().numpy()
    print("B=%d latency_path=%d ms=%.2f iters max %d sum %d  -> %.1f us per round of the slowest" % (B, flag, e0.elapsed_time(e1), it.max(), it.sum(), 1e3 * e0.elapsed_time(e1) / max(it.max(), 1)), flush=True)
s.close()
