"""gt_mpc value term alone: throughput of the tcgen05 kernel vs the fp64 CUDA-core kernel.
usage: mlp_probe.py [B]   (ncu: -k regex:mlp_tc_kernel)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench
from igt_mpc_int_b200.planner import BatchSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
rng = np.random.default_rng(0)
sN, vN = rng.uniform(0, 70, B), rng.uniform(0, 5, B)
ctx = np.stack([rng.uniform(0, 70, B), rng.uniform(0, 5, B), rng.integers(-8, 9, B).astype(float), rng.integers(-8, 9, B).astype(float)], 1)
s = BatchSolver(N=40, mlp=bench.random_mlp())
flop = 6 * 2 * (6 * 128 + 128 * 128 + 128) * B            # value + 5 tangent rows through 6-128-128-1
for tc in (True, False):
    s.mlp_value(sN[:1024], vN[:1024], ctx[:1024], tensor_cores=tc)
    t0 = time.perf_counter(); out = s.mlp_value(sN, vN, ctx, tensor_cores=tc); dt = time.perf_counter() - t0
    print("tensor_cores=%s B=%d host-to-host %.2f ms (%.2f algorithmic TFLOP/s incl. copies)" % (tc, B, dt * 1e3, flop / dt / 1e12), flush=True)
s.close()
