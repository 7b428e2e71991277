"""Reference tracks of the four-way intersection: host-side restatement of
`ReferenceGenerator` in frenet mode (common/ReferenceGen.py:41-235) without CasADi or the
attribute-bag classes.

A track is the path a vehicle follows from the edge of its entry road to the end of its exit road:
x, y, heading, v, s, K sampled every dt while stepping the Cartesian Euler model
(common/kinematic_bicycle_model.py:26-31) with a = 0, df = 0, i.e. straight along the heading.  Inside
the intersection square the stepped point is projected onto the turn circle and the heading is set
to the circle's tangent (ReferenceGen.py:158-186); the speed is 0.9 v_des (left) / 0.7 v_des (right)
there.  The reference keeps mutating the record it stored one step earlier (`state_copy.v = ...`,
`state_copy.heading = ...` act on the object already appended, ReferenceGen.py:157,185,190-191), so
the stored v / heading of a sample are those its successor's step used -- reproduced here, because
`frenet2global` and the plots consume exactly these arrays.  Checked sample for sample against
tests/golden/geometry.npz (written by the reference's own code).
"""
import math

import numpy as np

from . import geometry as G

X, Y, HEADING, V, S, K = range(6)


def circle_params(route, start_x, start_y, W=G.ROAD_WIDTH, L=G.ROAD_LENGTH, ca=G.CA_RADIUS):
    """Centre (h, k) and radius r of the turn circle, ReferenceGen.py:60-94."""
    fillet = W - ca
    lo, hi = (L - W) / 2, (L + W) / 2
    if route == '12':
        return lo, W, W - start_y
    if route == '21':
        return lo - fillet, W + fillet, fillet + ca
    if route == '14':
        return lo - fillet, 0 - fillet, fillet + ca
    if route == '41':
        return lo, 0, start_x - lo
    if route == '23':
        return hi, W, hi - start_x
    if route == '32':
        return hi + fillet, W + fillet, fillet + ca
    if route == '34':
        return hi, 0, start_y
    if route == '43':
        return hi + fillet, 0 - fillet, fillet + ca
    raise ValueError("not a turning route: %r" % (route,))


def in_intersection(route, x, y, h, k):
    """ReferenceGen.py:95-119."""
    if route in ('12', '21'):
        return x >= h and y <= k
    if route in ('14', '41'):
        return x >= h and y >= k
    if route in ('23', '32'):
        return x <= h and y <= k
    if route in ('34', '43'):
        return x <= h and y >= k
    raise ValueError("invalid route")


def _step(rec, dt):
    """Euler step with a = 0, df = 0 (beta = 0): kinematic_bicycle_model.py:26-31."""
    x, y, psi, v = rec[X], rec[Y], rec[HEADING], rec[V]
    beta = np.arctan(0.5 * np.tan(0.0))
    return [x + dt * v * np.cos(psi + beta), y + dt * v * np.sin(psi + beta),
            psi + dt * (v * np.cos(beta) / (G.L_R + G.L_F) * np.tan(0.0)), v + dt * 0.0]


def generate_track(route, n_steps=300, dt=0.1, v_des=5.0, W=G.ROAD_WIDTH, L=G.ROAD_LENGTH, ca=G.CA_RADIUS):
    """[6, n_steps + 1] array (x, y, heading, v, s, K) from the default start of the route's entry road
    (evaluate.py:59-62), like ReferenceGenerator(N=n_steps, mode='frenet').state2array() for one agent."""
    x0, y0, th0 = G.start_pose(route[0], W, L, ca)
    goal_heading = G.goal_pose(route[1], W, L, ca)[2]
    flip = route in ('32', '41')                                   # ReferenceGen.py:127, :185
    if flip:
        th0, goal_heading = abs(th0), abs(goal_heading)
    cur = [x0, y0, th0, v_des, 0, 0]                               # mutable record, as in the reference
    recs = [cur]
    s = 0
    if route in G.STRAIGHT:
        for _ in range(n_steps):
            nx, ny, nh, nv = _step(cur, dt)
            s += cur[V] * dt
            cur = [nx, ny, nh, nv, s, 0]
            recs.append(cur)
        return np.array(recs, dtype=np.float64).T
    left = route in G.LEFT
    h, k, r = circle_params(route, x0, y0, W, L, ca)
    centre = np.array([h, k])
    passed = False
    psi = None
    for _ in range(n_steps):
        if in_intersection(route, cur[X], cur[Y], h, k):
            curvature = 1 / r if left else -1 / r
            passed = True
            weight = 0.9 if left else 0.7
            cur[V] = v_des * weight                                # mutates the stored sample
            nx, ny, nh, nv = _step(cur, dt)
            vec = np.array([nx, ny]) - centre                      # projection onto the circle, :45-59
            px, py = centre + vec / np.linalg.norm(vec) * r
            dx, dy = px - h, py - k
            if left:                                               # tangent heading by quadrant, :164-180
                if dx > 0 and dy > 0:
                    psi = np.pi / 2 + abs(np.arctan(dy / dx))
                elif dx < 0 and dy > 0:
                    psi = (np.pi / 2 - abs(np.arctan(dy / dx))) + np.pi
                elif dx < 0 and dy < 0:
                    psi = -(np.pi / 2 - abs(np.arctan(dy / dx)))
                elif dx > 0 and dy < 0:
                    psi = np.pi / 2 - abs(np.arctan(dy / dx))
            else:
                psi = np.arctan2(-dx, dy)
            new = [px, py, psi, v_des * weight]
            if not in_intersection(route, px, py, h, k):           # leaving: straight on along the goal heading
                cur[HEADING] = goal_heading
                new = _step(cur, dt)
        else:
            curvature = 0
            if passed:
                cur[HEADING] = goal_heading
            cur[V] = v_des
            new = _step(cur, dt)
        s += np.sqrt((new[0] - cur[X]) ** 2 + (new[1] - cur[Y]) ** 2)
        cur = [new[0], new[1], new[2], new[3], s, curvature]
        recs.append(cur)
    return np.array(recs, dtype=np.float64).T


def crop_track(track, x, y, n=150):
    """`get_reference(n, initial_states)` (ReferenceGen.py:203-235): n + 1 samples from the sample nearest
    to (x, y); zero-padded if the track ends first."""
    ind = int(np.argmin((track[X] - x) ** 2 + (track[Y] - y) ** 2))
    end = min(ind + n + 1, track.shape[1])
    out = np.zeros((6, n + 1))
    out[:, :end - ind] = track[:, ind:end]
    return out


def reference_dict(route, x, y, n=150, **kw):
    """The dict `get_reference(..., output_type=dict)` returns for one agent."""
    t = crop_track(generate_track(route, **kw), x, y, n)
    return {'x': t[X], 'y': t[Y], 'heading': t[HEADING], 'v': t[V], 's': t[S], 'K': t[K]}
