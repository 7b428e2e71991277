import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from igt_mpc_int_b200 import scenarios as S
from igt_mpc_int_b200.planner import BatchSolver
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
pb = S.mid_episode(B, N=40, seed=2026)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
x0, up, cv, ob, ctx = t(pb.x0), t(pb.u_prev), t(pb.curv), t(pb.obs), t(pb.nn_ctx)
HID = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (128, 128)
mlp = bench.random_mlp(HID)
for name, use_ctx, tcflag in (("mpc", False, 1), ("gt_tc", True, 1), ("gt_coop", True, 0)):
    s = BatchSolver(N=40, mlp=mlp)
    s.set_option("tensor_core_mlp", tcflag)
    out = s.solve_batch_device(x0, up, cv, ob, nn_ctx=ctx if use_ctx else None); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = s.solve_batch_device(x0, up, cv, ob, nn_ctx=ctx if use_ctx else None, out=out); e1.record(); torch.cuda.synchronize()
    st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
    print(name, "ms", round(e0.elapsed_time(e1), 1), "conv", (st == 0).sum(), "iters mean", it.mean().round(2), "max", it.max(), np.bincount(st, minlength=6), flush=True)
    s.close()
