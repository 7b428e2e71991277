"""Quick GPU probe: throughput of the solver kernel at several batch sizes / precisions."""
import sys, time, os
import numpy as np
import torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from igt_mpc_int_b200 import scenarios as S
from igt_mpc_int_b200.planner import BatchSolver

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    N = 40
    t0 = time.time(); pb = S.mid_episode(B, N=N); print("gen", time.time() - t0, flush=True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    x0, up, cv, ob = t(pb.x0), t(pb.u_prev), t(pb.curv), t(pb.obs)
    for prec in ("f64", "f32"):
        for bsz in (B,):
            s = BatchSolver(N=N, precision=prec)
            out = s.solve_batch_device(x0[:bsz], up[:bsz], cv[:bsz], ob[:bsz]); torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(); out = s.solve_batch_device(x0[:bsz], up[:bsz], cv[:bsz], ob[:bsz], out=out); ev1.record(); torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1)
            st = out["status"].cpu().numpy(); it = out["iters"].cpu().numpy()
            conv = (st == 0).sum()
            print(f"{prec} B={bsz} ms={ms:.1f} converged={conv} ({conv/bsz:.3f}) solves/s={conv/ms*1e3:.0f} "
                  f"iters med={np.median(it[st==0])} p90={np.percentile(it[st==0],90)} max={it.max()} status={np.bincount(st, minlength=5)}", flush=True)
            s.close()

if __name__ == "__main__":
    main()
