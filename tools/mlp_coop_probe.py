"""The cooperative fp64 value-term kernel alone (igt_mlp_value, mode 2).  usage: mlp_coop_probe.py [B] [hidden, e.g. 128,128]
(ncu: -k regex:mlp_coop_kernel)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench
from igt_mpc_int_b200.planner import BatchSolver
B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
HID = tuple(int(v) for v in sys.argv[2].split(",")) if len(sys.argv) > 2 else (128, 128)
rng = np.random.default_rng(0)
sN, vN = rng.uniform(0, 70, B), rng.uniform(0, 5, B)
ctx = np.stack([rng.uniform(0, 70, B), rng.uniform(0, 5, B), rng.integers(-8, 9, B).astype(float), rng.integers(-8, 9, B).astype(float)], 1)
s = BatchSolver(N=40, mlp=bench.random_mlp(HID))
dims = (6,) + HID + (1,)
flop = 6 * 2 * sum(a * b for a, b in zip(dims[:-1], dims[1:])) * B
s.mlp_value(sN[:1024], vN[:1024], ctx[:1024], tensor_cores=2)
t0 = time.perf_counter(); out = s.mlp_value(sN, vN, ctx, tensor_cores=2); dt = time.perf_counter() - t0
print("coop fp64 B=%d %s host-to-host %.2f ms (%.2f algorithmic TFLOP/s incl. copies)" % (B, HID, dt * 1e3, flop / dt / 1e12), flush=True)
s.close()
