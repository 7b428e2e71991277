"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN numpy-mode code.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The reference's solve path (CasADi/IPOPT) cannot run here, so solver outputs are NOT in these
files; what is pinned is everything the reference can execute: the Frenet RK4 model
(common/kinematic_bicycle_model_frenet.py:70-127), the Cartesian Euler model
(common/kinematic_bicycle_model.py:15-50), frenet2global (common/utils.py:532-586),
scenario_encoding (utils.py:141-169), filter_preds (utils.py:365-388), augment_prev_sol
(utils.py:354-363) and the generated reference tracks (common/ReferenceGen.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from igt_mpc_int_b200 import geometry as G  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def pw_const(curv):
    b0, b1, kv = curv
    return lambda s: kv * (1.0 if s >= b0 else 0.0) - kv * (1.0 if s >= b1 else 0.0)


def main():
    ns = ref_loader.load()
    VR = ns.VehicleReference.VehicleReference
    VA = ns.VehicleAction.VehicleAction
    VS = ns.VehicleState.VehicleState
    Frenet = ns.kinematic_bicycle_model_frenet.KinematicBicycleModelFrenet
    Cart = ns.kinematic_bicycle_model.KinematicBicycleModel
    U = ns.utils
    rng = np.random.default_rng(12345)

    # ---- Frenet RK4 single steps and rollouts (n_rk = 4, evaluate.py:109) -------------------
    model = Frenet(2.235, 2.235, 2.0, 0.1, discretization='rk4', mode='numpy', num_rk4_steps=4)
    curvs = [G.curvature_params(r) for r in ('12', '14', '13')]
    n_step = 240
    z0 = np.empty((n_step, 7)); u = np.empty((n_step, 2)); cv = np.empty((n_step, 3)); zn = np.empty((n_step, 7))
    for i in range(n_step):
        c = curvs[i % 3]
        s = rng.uniform(0, 60)
        if i % 7 == 0 and c[2] != 0:                       # sit right at / next to a breakpoint
            s = c[i % 2] + rng.choice([-1e-3, 0.0, 1e-3, -0.05, 0.02])
        z = np.array([rng.uniform(-20, 60), rng.uniform(-20, 40), s, rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3),
                      rng.uniform(0, 5), rng.uniform(-3.2, 3.2)])
        a, df = rng.uniform(-4, 3), rng.uniform(-0.6, 0.6)
        st = VR({'s': z[2], 'ey': z[3], 'epsi': z[4], 'v': z[5], 'x': z[0], 'y': z[1], 'heading': z[6], 'K': pw_const(c)})
        o = model(st, VA({'a': a, 'df': df}))
        z0[i], u[i], cv[i] = z, (a, df), c
        zn[i] = (o.x, o.y, o.s, o.ey, o.epsi, o.v, o.heading)
    n_roll, N = 48, 40
    rz0 = np.empty((n_roll, 7)); ru = np.empty((n_roll, N, 2)); rcv = np.empty((n_roll, 3)); rz = np.empty((n_roll, N + 1, 7))
    for i in range(n_roll):
        c = curvs[i % 3]
        z = np.array([rng.uniform(0, 50), rng.uniform(0, 12), rng.uniform(0, 30), rng.uniform(-0.15, 0.15),
                      rng.uniform(-0.1, 0.1), rng.uniform(0, 5), rng.uniform(-3.2, 3.2)])
        a = np.clip(np.cumsum(rng.uniform(-0.09, 0.09, N)) + rng.uniform(-1, 1), -4, 3)
        d = np.clip(np.cumsum(rng.uniform(-0.07, 0.07, N)) + rng.uniform(-0.1, 0.1), -0.6, 0.6)
        rz0[i], ru[i, :, 0], ru[i, :, 1], rcv[i] = z, a, d, c
        rz[i, 0] = z
        st = VR({'s': z[2], 'ey': z[3], 'epsi': z[4], 'v': z[5], 'x': z[0], 'y': z[1], 'heading': z[6], 'K': pw_const(c)})
        for k in range(N):
            st = model(st, VA({'a': a[k], 'df': d[k]}))
            rz[i, k + 1] = (st.x, st.y, st.s, st.ey, st.epsi, st.v, st.heading)
    np.savez(os.path.join(OUT, "frenet_rk4.npz"), z0=z0, u=u, curv=cv, zn=zn, rz0=rz0, ru=ru, rcurv=rcv, rz=rz)

    # ---- Cartesian Euler steps ---------------------------------------------------------------
    cm = Cart(2.235, 2.235, 2.0, 0.1)
    n = 64
    cz0 = np.empty((n, 4)); cu = np.empty((n, 2)); czn = np.empty((n, 4))
    for i in range(n):
        z = np.array([rng.uniform(-20, 60), rng.uniform(-20, 40), rng.uniform(-3.2, 3.2), rng.uniform(0, 5)])
        a, df = rng.uniform(-4, 3), rng.uniform(-0.6, 0.6)
        o = cm(VR({'x': z[0], 'y': z[1], 'heading': z[2], 'v': z[3], 's': 0, 'ey': 0, 'epsi': 0, 'K': 0}), VA({'a': a, 'df': df}))
        cz0[i], cu[i], czn[i] = z, (a, df), (o.x, o.y, o.heading, o.v)
    np.savez(os.path.join(OUT, "cartesian_euler.npz"), z0=cz0, u=cu, zn=czn)

    # ---- reference tracks, frenet2global, scenario encoding -----------------------------------
    W, L, ca = G.ROAD_WIDTH, G.ROAD_LENGTH, G.CA_RADIUS
    env_init = {r: VS(dict(zip(('x', 'y', 'heading'), G.start_pose(r)), v=0)) for r in '1234'}
    goals = {r: VS(dict(zip(('x', 'y', 'heading'), G.goal_pose(r)), v=5)) for r in '1234'}
    tracks = {}
    f2g_s = np.linspace(0, 70, 141)
    f2g = {}
    for route in G.ROUTES:
        other = '13' if route != '13' else '24'
        routes = [route, other]
        init = [{'type': 'CAV', 'state': env_init[r[0]]} for r in routes]
        rg = ns.ReferenceGen.ReferenceGenerator(N=300, dt=0.1, initial_state=init, goals=[goals[r[1]] for r in routes],
                                                env=None, radius=ca, routes=routes, target_velocity=5, mode='frenet',
                                                road_width=W, road_length=L, fillet_radius=W - ca)
        ref = rg.get_reference(150, initial_states=init, output_type=dict)[0]
        tracks[route] = np.stack([ref['x'], ref['y'], ref['heading'], ref['v'], ref['s'], ref['K']])
        f2g[route] = np.array([U.frenet2global(s, ref, route, L, W, ca).ravel() for s in f2g_s])
    enc_routes, enc_vals = [], []
    for sc in range(1, 9):
        for rot in range(4):
            for order in range(2):
                r = G.scenario_routes(sc, rot, order)
                enc_routes.append(r)
                enc_vals.append(U.scenario_encoding(r))
    np.savez(os.path.join(OUT, "geometry.npz"), routes=np.array(G.ROUTES), tracks=np.stack([tracks[r] for r in G.ROUTES]),
             f2g_s=f2g_s, f2g=np.stack([f2g[r] for r in G.ROUTES]), enc_routes=np.array(enc_routes),
             enc_vals=np.array(enc_vals))

    # ---- filter_preds and augment_prev_sol ------------------------------------------------------
    n = 24
    fp_ego = np.empty((n, 3)); fp_obs = np.empty((n, 5, 2)); fp_out = np.empty((n, 5, 2))
    for i in range(n):
        ex, ey_, eh = rng.uniform(0, 50), rng.uniform(-10, 20), rng.uniform(-3.2, 3.2)
        ob = np.stack([rng.uniform(0, 50, 5), rng.uniform(-10, 20, 5)], 1)
        preds = [[VR({'x': ex, 'y': ey_, 'heading': eh, 'v': 1, 's': 0, 'K': None, 'ey': 0, 'epsi': 0}) for _ in range(5)],
                 [VR({'x': ob[k, 0], 'y': ob[k, 1], 'heading': 0, 'v': 1, 's': 0, 'K': None, 'ey': 0, 'epsi': 0}) for k in range(5)]]
        out = U.filter_preds(preds, 0)
        fp_ego[i], fp_obs[i] = (ex, ey_, eh), ob
        fp_out[i] = [(q.x, q.y) for q in out[1]]
    n = 12
    ap_x = np.empty((n, 7, N + 1)); ap_u = np.empty((n, 2, N)); ap_cv = np.empty((n, 3))
    ap_xo = np.empty((n, 7, N + 1)); ap_uo = np.empty((n, 2, N))
    for i in range(n):
        c = curvs[i % 3]
        x = rz[i].T.copy(); uu = ru[i].T.copy()
        if i % 4 == 0:
            x[5, -1] = 4.99; uu[0, -1] = 0.5          # exercises the v > 5 re-roll branch (utils.py:359-360)
        xo, uo = U.augment_prev_sol((x, uu), model, pw_const(c))
        ap_x[i], ap_u[i], ap_cv[i], ap_xo[i], ap_uo[i] = x, uu, c, xo, uo
    np.savez(os.path.join(OUT, "glue.npz"), fp_ego=fp_ego, fp_obs=fp_obs, fp_out=fp_out, ap_x=ap_x, ap_u=ap_u,
             ap_curv=ap_cv, ap_xo=ap_xo, ap_uo=ap_uo)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
