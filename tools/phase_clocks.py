"""Per-phase cycle counters of one persistent thread (debug build of the library with -DIGT_PHASE_CLOCKS).
usage: phase_clocks.py LIB.so B"""
import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from igt_mpc_int_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
from igt_mpc_int_b200 import scenarios as S
from igt_mpc_int_b200.planner import BatchSolver
B = int(sys.argv[2]); N = 40
GT = len(sys.argv) > 3 and sys.argv[3] == "gt"          # gt_mpc mode with the random-init value network (tensor-core path)
pb = S.mid_episode(B, N=N)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
x0, up, cv, ob = t(pb.x0), t(pb.u_prev), t(pb.curv), t(pb.obs)
kw = {}
if GT:
    import bench
    s = BatchSolver(N=N, mlp=bench.random_mlp())
    kw = dict(nn_ctx=t(pb.nn_ctx))
else:
    s = BatchSolver(N=N)
out = s.solve_batch_device(x0, up, cv, ob, **kw); torch.cuda.synchronize()
lib0 = _lib.load()
if hasattr(lib0, "igt_debug_round_max"):
    lib0.igt_debug_round_max.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib0.igt_debug_round_max(s._h, (C.c_int * 2048)())          # reset after the warm-up solve
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = s.solve_batch_device(x0, up, cv, ob, out=out, **kw); e1.record(); torch.cuda.synchronize()
clk = (C.c_longlong * 16)()
lib = _lib.load()
lib.igt_debug_phase_clocks.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
assert lib.igt_debug_phase_clocks(s._h, clk) == 0
names = ["sched", "node1(sens,rows)", "adjoint+test", "node2(kkt)", "riccati", "step_bound", "sync", "rollout", "collect", "idle", "accept+out", "node3"]
tot = sum(clk)
print("B=%d ms=%.1f total cycles %.3e (%.1f ms @1.965GHz)" % (B, e0.elapsed_time(e1), tot, tot / 1.965e6))
for n, c in zip(names, clk):
    print("%-18s %12d  %5.1f%%" % (n, c, 100.0 * c / max(tot, 1)))

if hasattr(lib, "igt_debug_cta_clocks"):
    cc, cr = (C.c_longlong * 512)(), (C.c_int * 512)()
    lib.igt_debug_cta_clocks.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    assert lib.igt_debug_cta_clocks(s._h, cc, cr) == 0
    cc = np.array(cc[:148]) / 1.965e6; cr = np.array(cr[:148])
    print("per CTA: ms from start to exit min %.1f median %.1f max %.1f; loop passes min %d median %d max %d; ms of the CTA with most passes %.1f"
          % (cc.min(), np.median(cc), cc.max(), cr.min(), np.median(cr), cr.max(), cc[cr.argmax()]))

if hasattr(lib, "igt_debug_round_clocks"):
    rc, rn = (C.c_longlong * 512)(), (C.c_int * 512)()
    lib.igt_debug_round_clocks.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    assert lib.igt_debug_round_clocks(s._h, rc, rn) == 0
    print("round: active problems of CTA 0 / kcycles since the previous pass")
    print(" ".join("%d:%d/%.0f" % (i, rn[i], rc[i] / 1e3) for i in range(512) if rn[i] > 0))

if hasattr(lib, "igt_debug_round_phases"):
    ph = (C.c_int * (512 * 12))()
    lib.igt_debug_round_phases.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    assert lib.igt_debug_round_phases(s._h, ph) == 0
    print("per pass (a pass's phases are recorded one slot late for ticks after the round marker): n | " + " ".join(n[:8] for n in names))
    for i in range(512):
        if rn[i] > 0 and (i < 30 or i % 4 == 0):
            print("%3d n=%3d | " % (i, rn[i]) + " ".join("%7d" % ph[i * 12 + j] for j in range(12)))

if hasattr(lib, "igt_debug_round_max"):
    mx = (C.c_int * 2048)()
    assert lib.igt_debug_round_max(s._h, mx) == 0
    print("per pass, max over CTA 0's threads (kcycles): adjoint+test | riccati (all sweeps) | step bound | riccati sweeps")
    for i in range(512):
        if rn[i] > 0 and (i < 30 or i % 2 == 0):
            print("%3d n=%3d | %6d %6d %6d  x%d" % (i, rn[i], mx[4 * i], mx[4 * i + 1], mx[4 * i + 2], mx[4 * i + 3]))
