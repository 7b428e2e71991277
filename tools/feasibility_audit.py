"""Infeasibility audit: are the problems the solver gives up on really infeasible?

For the non-converged problems (status 1 max-iter / 4 line search / 5 stalled-infeasible) of a config-2 sample, an
independent phase-1 search (oracle/kkt_check.phase1: scipy least_squares on the squared row violations, analytic
Jacobian, several starts) looks for ANY control sequence that satisfies every row of mpc.py:177-321.  A problem on
which it finds one (max violation <= 1e-6, the north-star bound) is feasible, i.e. the solver was wrong to fail.

usage: python tools/feasibility_audit.py [B=4096] [seed=2026] [max_problems=400] > profiles/r2_feasibility_audit.log
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import nlp, c_oracle, kkt_check as K          # noqa: E402
from igt_mpc_int_b200 import scenarios as S               # noqa: E402

P = nlp.Params(N=40)


def starts_for(prob, U_last):
    st = [nlp.tracking_rollout(P, prob, vt)[0] for vt in nlp.GUESS_SPEEDS]
    if U_last is not None and np.all(np.isfinite(U_last)):
        st.insert(0, U_last)
    N = P.N
    for a_target in (P.a_min, P.a_max):                     # hardest braking / acceleration the jerk limit allows
        U = np.zeros((N, 2)); a = prob.u_prev[0]
        for k in range(N):
            a = min(max(a_target, a - P.da_max), a + P.da_max)
            U[k, 0] = a
        U[:, 1] = st[-1][:, 1]
        st.append(U)
    return st


def audit_one(args):
    i, x0, up, cv, ob, U_last = args
    prob = nlp.Problem(x0=x0, u_prev=up, curv=tuple(cv), obs=ob)
    v, U = K.phase1(P, prob, starts_for(prob, U_last))
    return i, v


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2026
    cap = int(sys.argv[3]) if len(sys.argv) > 3 else 400
    pb = S.mid_episode(B, N=40, seed=seed)
    o = c_oracle.COracle(P, max_iter=60).solve(pb.x0, pb.u_prev, pb.curv, pb.obs)
    st = o["status"]
    print("config-2 sample: B=%d seed=%d  status counts (0 converged, 1 cap, 2 x0, 3 reg, 4 line search, 5 stalled infeasible, "
          "6 acceptable = within the reference's IPOPT tolerances): %s" % (B, seed, np.bincount(st, minlength=7).tolist()))
    bad = np.where((st == 1) | (st == 4) | (st == 5) | (st == 3))[0][:cap]
    t0 = time.time()
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(audit_one, [(int(i), pb.x0[i], pb.u_prev[i], pb.curv[i], pb.obs[i], o["U"][i]) for i in bad])
    feas = [(i, v) for i, v in res if v <= 1e-6]
    print("audited %d non-converged problems in %.0f s: phase-1 found a feasible point for %d of them" % (len(bad), time.time() - t0, len(feas)))
    for code, name in ((1, "max_iter"), (3, "reg_limit"), (4, "line_search"), (5, "stalled_infeasible")):
        sel = [(i, v) for i, v in res if st[i] == code]
        if sel:
            vs = np.array([v for _, v in sel])
            print("  status %d %-18s n=%4d  feasible point found: %3d   least max-violation found: min %.2e median %.2e"
                  % (code, name, len(sel), int((vs <= 1e-6).sum()), vs.min(), np.median(vs)))
    for i, v in feas:
        print("  FEASIBLE problem %d (status %d, iters %d): phase-1 max violation %.1e" % (i, st[i], o["iters"][i], v))
    # control: the same search on converged problems must (of course) find feasible points
    good = np.where(st == 0)[0][:32]
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        res_g = pool.map(audit_one, [(int(i), pb.x0[i], pb.u_prev[i], pb.curv[i], pb.obs[i], None) for i in good])
    print("control: phase-1 on %d converged problems (solver's solution NOT given as a start): feasible point found for %d"
          % (len(good), sum(v <= 1e-6 for _, v in res_g)))


if __name__ == "__main__":
    main()
