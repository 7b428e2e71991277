"""Result files of a closed-loop evaluation in the reference's formats, so that the
downstream plotting / statistics scripts written for `evaluate.py` keep working on runs of this solver.

Files of the 'mpc' branch (evaluate.py:577-639), one directory per run:
  mpc/cl_traj.pkl          ndarray [samples, 7 M, T+1]   rows 7i..7i+6 = x, y, s, ey, epsi, v, heading of agent i
  mpc/u_cl.pkl             ndarray [samples, 2 M, T]     rows 2i, 2i+1 = a, df of agent i (see `u_cl_reference_layout`)
  mpc/evaluation_data.pkl  dict: N, refs [samples, 6 M, T+1], x_cl [samples, 21, T+1], u_cl [samples, 6, T],
                           initial_agents, agent_types, weights, routes, goals, initial_default, deadlock
  mpc/eval_stats.csv       avg_sol_times, std_solve_times, infeasible_ratio, deadlock -- one line per sample
  mpc/mpc.yaml             the policy configuration
The 'gt_mpc' branch (evaluate.py:326-369) writes cl_traj.pkl / u_cl.pkl under game_mpc/evaluation/ and
stats.csv with an extra NN_query_time column.  Every writer APPENDS a sample to files that already exist,
as the reference does when num_samples > 1.
"""
import csv
import os
import pickle

import numpy as np


def z_cl_rows(z_cl_e):
    """[2, T+1, 7] states of one episode -> the reference's [7 M, T+1] layout (evaluate.py:430-436)."""
    M, T1, _ = z_cl_e.shape
    out = np.zeros((7 * M, T1))
    for i in range(M):
        out[7 * i:7 * i + 7] = z_cl_e[i].T
    return out


def u_cl_reference_layout(u_cl_e, solved_e, u0=(0.1, 0.0)):
    """[2, T, 2] applied inputs + [2, T] success flags -> the reference's u_cl [2 M, T].

    The reference stores the input of a SUCCESSFUL step in column 0 (`u_cl[i*2+0,0] = u_sol[0,0]`,
    evaluate.py:505-506) and only the brake-fallback inputs in their own column t (evaluate.py:540-541), so
    column 0 ends up holding the last successful input and the columns of successful steps stay zero.  Files
    are reproduced in that layout; `u_data` inside evaluation_data.pkl has every step in its own column."""
    M, T, _ = u_cl_e.shape
    out = np.zeros((2 * M, T))
    for i in range(M):
        out[2 * i, 0], out[2 * i + 1, 0] = u0                      # evaluate.py:437-438
        for t in range(T):
            if solved_e[i, t]:
                out[2 * i:2 * i + 2, 0] = u_cl_e[i, t]
            else:
                out[2 * i:2 * i + 2, t] = u_cl_e[i, t]
    return out


def _append_pickle(path, arr):
    if os.path.isfile(path):
        with open(path, 'rb') as f:
            old = pickle.load(f)
        arr = np.concatenate([old, arr], axis=0)
    with open(path, 'wb') as f:
        pickle.dump(arr, f)


def _append_csv(path, row):
    new = not os.path.isfile(path)
    with open(path, mode='a' if not new else 'w', newline='') as f:
        w = csv.DictWriter(f, fieldnames=list(row.keys()))
        if new:
            w.writeheader()
        w.writerow(row)


def write_episode(run_dir, eval_mode, z_cl_e, u_cl_e, solved_e, deadlock, solve_times, *, N=40, refs=None,
                  initial_agents=None, initial_default=None, routes=None, goals=None, policy_config=None,
                  nn_query_time=-1):
    """Append one episode to the result files under `run_dir` (created if missing).

    z_cl_e [M, T+1, 7], u_cl_e [M, T, 2], solved_e [M, T] bool, solve_times: per-agent lists of seconds."""
    M, T = u_cl_e.shape[0], u_cl_e.shape[1]
    sub = os.path.join(run_dir, 'mpc') if eval_mode == 'mpc' else os.path.join(run_dir, 'game_mpc', 'evaluation')
    os.makedirs(sub, exist_ok=True)
    os.makedirs(os.path.join(run_dir, 'mpc' if eval_mode == 'mpc' else 'game_mpc', 'evaluation_videos'), exist_ok=True)
    if policy_config is not None and not os.path.isfile(os.path.join(sub, 'mpc.yaml')):
        import yaml
        with open(os.path.join(sub, 'mpc.yaml'), 'w') as f:
            yaml.dump(policy_config, f)
    z_rows = z_cl_rows(z_cl_e)
    u0 = (0.1, 0.0) if eval_mode == 'mpc' else (0.0, 0.0)          # evaluate.py:437-438 / :188-189
    _append_pickle(os.path.join(sub, 'cl_traj.pkl'), z_rows[np.newaxis])
    _append_pickle(os.path.join(sub, 'u_cl.pkl'), u_cl_reference_layout(u_cl_e, solved_e, u0)[np.newaxis])
    avg = np.array([np.mean(t) if len(t) else np.nan for t in solve_times])
    std = [np.std(a) for a in avg]                                  # the reference takes np.std of each mean (:606): zeros
    stat = {'avg_sol_times': avg, 'std_solve_times': std,
            'infeasible_ratio': np.array([(~solved_e[i]).sum() / T for i in range(M)]), 'deadlock': bool(deadlock)}
    if eval_mode == 'mpc':
        x_data = np.zeros((7 * 3, T + 1)); u_data = np.zeros((2 * 3, T))         # evaluate.py:425-426: 3 agent slots
        x_data[:7 * M] = z_rows
        for i in range(M):
            u_data[2 * i:2 * i + 2] = u_cl_e[i].T
        ref_data = np.zeros((6 * M, T + 1))
        if refs is not None:
            for i in range(M):
                for j, key in enumerate(('x', 'y', 'heading', 'v', 's', 'K')):
                    ref_data[6 * i + j] = refs[i][key]
        path = os.path.join(sub, 'evaluation_data.pkl')
        new = {'N': N, 'refs': np.array([ref_data]), 'x_cl': np.array([x_data]), 'u_cl': np.array([u_data]),
               'initial_agents': initial_agents, 'agent_types': ['CAV'] * M, 'weights': np.array([[1, 1]]),
               'routes': np.array([routes]), 'goals': np.array([goals]), 'initial_default': initial_default,
               'deadlock': np.array([bool(deadlock)])}
        if os.path.isfile(path):
            with open(path, 'rb') as f:
                old = pickle.load(f)
            for key in ('refs', 'x_cl', 'u_cl', 'weights', 'routes', 'goals'):
                new[key] = np.concatenate([old[key], new[key]], axis=0)
            new['initial_agents'] = np.vstack([old['initial_agents'], initial_agents])
            new['deadlock'] = np.vstack([old['deadlock'], np.array([bool(deadlock)])])
        with open(path, 'wb') as f:
            pickle.dump(new, f, protocol=pickle.HIGHEST_PROTOCOL)
        _append_csv(os.path.join(sub, 'eval_stats.csv'), stat)
    else:
        stat = dict({'NN_query_time': np.array([nn_query_time])}, **stat)
        _append_csv(os.path.join(sub, 'stats.csv'), stat)
    return sub
