"""igt_mpc_int_b200 -- B200-native batched solver for IGT-MPC-INT's per-timestep MPC solve.

Public surface:
  planner.MPC_Planner      drop-in for the reference's mpc.py MPC_Planner (same constructor
                           keywords, update_initial_condition / update_predictions / solve)
  planner.BatchSolver      the batched API (solve_batch) used by bench.py and episode drivers
  geometry, scenarios      host-side intersection geometry and synthetic problem generators
The compute path is the CUDA library behind include/igt_mpc.h; there is no CPU fallback.
"""
__version__ = "0.1"
