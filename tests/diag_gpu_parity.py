import sys, os
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from igt_mpc_int_b200 import scenarios as S
from igt_mpc_int_b200.planner import BatchSolver
from oracle import nlp, c_oracle
TIGHT = dict(tol=1e-6, tol_rp=1e-8, tol_comp=1e-7, mu_floor=1e-8, max_iter=300)
N, B = 40, int(sys.argv[1]) if len(sys.argv) > 1 else 512
pb = S.mid_episode(B, N=N)
P = nlp.Params(N=N)
for name, kw in (("tight", TIGHT), ("default", {})):
    s = BatchSolver(N=N, **kw)
    r = s.solve_batch(pb.x0, pb.u_prev, pb.curv, pb.obs)
    okw = dict(n_alpha=s.params.n_alpha)
    if name == "tight": okw.update(max_iter=300)
    o = c_oracle.COracle(P, **okw).solve(pb.x0, pb.u_prev, pb.curv, pb.obs)
    same = r["status"] == o["status"]
    ok = (o["status"] == 0) & (r["status"] == 0)
    dc = np.abs(r["cost"] - o["cost"])[ok] / np.maximum(1, np.abs(o["cost"][ok]))
    du = np.abs(r["u"] - o["U"])[ok].reshape(ok.sum(), -1).max(1)
    print(name, "status agree", same.mean(), "both ok", ok.sum(), "gpu ok", (r["status"]==0).sum(), "oracle ok", (o["status"]==0).sum())
    print("  cost rel diff: max", dc.max(), " >1e-4:", (dc > 1e-4).sum(), " du max", du.max(), " >1e-3:", (du > 1e-3).sum(), " viol max", r["viol"][ok].max())
    print("  iters equal frac", (r["iters"][ok] == o["iters"][ok]).mean(), "mismatch idx", np.where(~same)[0][:10], "bad du idx", np.where(ok)[0][du > 1e-3][:10])
    bad = np.where(ok)[0][du > 1e-3]
    for i in bad[:5]:
        print("   prob", i, "cost gpu/oracle", r["cost"][i], o["cost"][i], "iters", r["iters"][i], o["iters"][i])
    s.close()
