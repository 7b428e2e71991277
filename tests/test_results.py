"""Result files in the reference's formats (evaluate.py:326-369, :577-639) and the evaluate-style entry point,
driven on the CPU with the oracle as the solver."""
import csv
import os
import pickle
from types import SimpleNamespace

import numpy as np

from igt_mpc_int_b200 import evaluate, results
from tests.oracle_backend import OracleBackend


def _args(tmp, **kw):
    d = dict(save_dir=str(tmp), num_samples=1, eval_mode='mpc', sc=1, rotation=0, order=0, all_variants=False, nn_weights=None, steps=150)
    d.update(kw)
    return SimpleNamespace(**d)


def test_u_cl_layout_reproduces_the_reference_quirk():
    u = np.arange(2 * 4 * 2, dtype=float).reshape(2, 4, 2) + 1
    solved = np.array([[True, True, False, True], [False, False, True, True]])
    out = results.u_cl_reference_layout(u, solved)
    assert out.shape == (4, 4)
    assert np.array_equal(out[0:2, 0], u[0, 3])          # column 0: the last successful input (evaluate.py:505-506)
    assert np.array_equal(out[0:2, 2], u[0, 2])          # a brake-fallback step keeps its own column (evaluate.py:540-541)
    assert np.all(out[0:2, [1, 3]] == 0)
    assert np.array_equal(out[2:4, 1], u[1, 1]) and np.array_equal(out[2:4, 0], u[1, 3])


def test_evaluate_writes_reference_files(tmp_path):
    backend = OracleBackend(N=40, max_iter=60)
    run_dir, summary = evaluate.main(_args(tmp_path, num_samples=2, steps=12), solver=backend)
    sub = os.path.join(run_dir, 'mpc')
    for f in ('cl_traj.pkl', 'u_cl.pkl', 'evaluation_data.pkl', 'eval_stats.csv', 'mpc.yaml'):
        assert os.path.isfile(os.path.join(sub, f)), f
    cl = pickle.load(open(os.path.join(sub, 'cl_traj.pkl'), 'rb'))
    ucl = pickle.load(open(os.path.join(sub, 'u_cl.pkl'), 'rb'))
    assert cl.shape == (2, 14, 13) and ucl.shape == (2, 4, 12)                 # [samples, 7 M, T+1], [samples, 2 M, T]
    data = pickle.load(open(os.path.join(sub, 'evaluation_data.pkl'), 'rb'))
    assert data['N'] == 40 and data['x_cl'].shape == (2, 21, 13) and data['u_cl'].shape == (2, 6, 12)
    assert data['refs'].shape == (2, 12, 13) and data['deadlock'].shape == (2, 1) and data['routes'].shape == (2, 2)
    assert np.array_equal(data['x_cl'][0, :14], cl[0])
    # vehicles start at rest, s0 from default_rng(2026), and accelerate at the jerk limit
    assert abs(cl[0, 2, 0] - 0.17893481367543618 * 10.7) < 1e-12 and cl[0, 5, 0] == 0.0
    assert abs(cl[1, 2, 0] - 0.3549173343096512 * 10.7) < 1e-12                # second sample: next four draws
    assert abs(data['u_cl'][0, 0, 0] - 0.19) < 1e-5
    rows = list(csv.DictReader(open(os.path.join(sub, 'eval_stats.csv'))))
    assert len(rows) == 2 and set(rows[0]) == {'avg_sol_times', 'std_solve_times', 'infeasible_ratio', 'deadlock'}
    assert len(summary) == 2 and summary[0]['routes'] == ['13', '23']


import pytest


@pytest.mark.gpu
@pytest.mark.parametrize("device_loop", [False, True])
def test_evaluate_with_the_cuda_solver_writes_the_same_files(tmp_path, device_loop):
    """BASELINE config 1 end to end on the GPU: `evaluate --eval_mode mpc --sc 1` with the CUDA solver (host-driven
    loop and the on-device loop) writes the reference's result files, and they agree with the oracle-driven run."""
    a = _args(tmp_path / "gpu", num_samples=1, steps=150, all_variants=True)
    a.device_loop = device_loop
    run_dir, summary = evaluate.main(a)
    sub = os.path.join(run_dir, 'mpc')
    cl = pickle.load(open(os.path.join(sub, 'cl_traj.pkl'), 'rb'))
    data = pickle.load(open(os.path.join(sub, 'evaluation_data.pkl'), 'rb'))
    rows = list(csv.DictReader(open(os.path.join(sub, 'eval_stats.csv'))))
    assert cl.shape == (8, 14, 151) and data['x_cl'].shape == (8, 21, 151) and len(rows) == 8
    ref_dir, ref_summary = evaluate.main(_args(tmp_path / "cpu", num_samples=1, steps=150, all_variants=True),
                                         solver=OracleBackend(N=40, max_iter=60))
    for g, o in zip(summary, ref_summary):
        assert g['routes'] == o['routes'] and g['deadlock'] == o['deadlock'] and g['collision'] == o['collision'] and g['goal'] == o['goal']
    cl_ref = pickle.load(open(os.path.join(ref_dir, 'mpc', 'cl_traj.pkl'), 'rb'))
    close = np.abs(cl - cl_ref).reshape(8, -1).max(axis=1) < 1e-3
    assert close.mean() >= 0.75
