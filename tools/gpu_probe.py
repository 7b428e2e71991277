"""Quick GPU probe: throughput of the solver kernel.  usage: gpu_probe.py B [f64|f32|both] [reps]"""
import sys, time, os
import numpy as np
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from igt_mpc_int_b200 import scenarios as S
from igt_mpc_int_b200.planner import BatchSolver

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    precs = ("f64", "f32") if len(sys.argv) < 3 or sys.argv[2] == "both" else (sys.argv[2],)
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    N = 40
    cache = f"/tmp/pb_{B}.npz"
    if os.path.exists(cache):
        d = np.load(cache); x0, up, cv, ob = d["x0"], d["up"], d["cv"], d["ob"]
    else:
        pb = S.mid_episode(B, N=N); x0, up, cv, ob = pb.x0, pb.u_prev, pb.curv, pb.obs
        np.savez(cache, x0=x0, up=up, cv=cv, ob=ob)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    x0, up, cv, ob = t(x0), t(up), t(cv), t(ob)
    for prec in precs:
        kw = {}
        if os.environ.get('IGT_MAX_TRIALS'): kw['max_trials'] = int(os.environ['IGT_MAX_TRIALS'])
        if os.environ.get('IGT_MAX_ITER'): kw['max_iter'] = int(os.environ['IGT_MAX_ITER'])
        s = BatchSolver(N=N, precision=prec, **kw)
        out = s.solve_batch_device(x0, up, cv, ob); torch.cuda.synchronize()
        for _ in range(reps):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(); out = s.solve_batch_device(x0, up, cv, ob, out=out); ev1.record(); torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
            conv = (st == 0).sum()
            print(f"{prec} B={B} ms={ms:.1f} converged={conv} ({conv/B:.3f}) solves/s={conv/ms*1e3:.0f} "
                  f"iters med={np.median(it[st==0])} p90={np.percentile(it[st==0],90)} sum={it.sum()} status={np.bincount(st, minlength=5)}", flush=True)
        s.close()

if __name__ == "__main__":
    main()
