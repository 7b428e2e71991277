"""Groundwork for the OBCA collision mode (mpc.py:211-221): the reference's dual formulation maximised over its
dual variables equals the rectangle-to-rectangle distance, whose closed form and gradient are what a stage-local
row needs (oracle/obca.py)."""
import numpy as np

from oracle import obca


def _poses(rng, n, lo, hi):
    out = []
    while len(out) < n:
        ego = np.array([rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-np.pi, np.pi)])
        ang, dist = rng.uniform(-np.pi, np.pi), rng.uniform(lo, hi)
        obs = np.array([ego[0] + dist * np.cos(ang), ego[1] + dist * np.sin(ang), rng.uniform(-np.pi, np.pi)])
        out.append((ego, obs))
    return out


def test_h_representation_matches_reference_convention():
    A, b = obca.rotation_translation([3.0, -1.0], 0.4)
    c, s = np.cos(0.4), np.sin(0.4)
    front = np.array([3.0, -1.0]) + 2.235 * np.array([c, s])          # the front bumper's centre lies on face 0
    assert abs(A[0] @ front - b[0]) < 1e-12 and np.all(A @ np.array([3.0, -1.0]) < b)
    assert np.allclose(A[:2], -A[2:])


def test_dual_formulation_equals_rectangle_distance():
    rng = np.random.default_rng(7)
    for ego, obs in _poses(rng, 40, 3.0, 12.0):
        d, _ = obca.rect_distance(ego, obs)
        v = obca.dual_value(ego, obs)
        assert abs(d - v) < 1e-5 * max(1.0, d), (ego, obs, d, v)
    # intersecting rectangles: the dual cannot certify any positive distance (lambda = mu = 0 is its optimum)
    for ego, obs in _poses(rng, 10, 0.0, 0.9):
        d, g = obca.rect_distance(ego, obs)
        assert d == 0.0 and not g.any()
        assert obca.dual_value(ego, obs) < 1e-6


def test_rectangle_distance_gradient():
    rng = np.random.default_rng(3)
    checked = 0
    for ego, obs in _poses(rng, 60, 3.5, 10.0):
        d, g = obca.rect_distance(ego, obs)
        if d < 0.05:
            continue
        fd = np.zeros(3)
        smooth = True
        for i in range(3):
            e = np.zeros(3); e[i] = 1e-6
            dp, gp = obca.rect_distance(ego + e, obs)
            dm, gm = obca.rect_distance(ego - e, obs)
            fd[i] = (dp - dm) / 2e-6
            smooth = smooth and np.allclose(gp, gm, atol=1e-3)      # skip the kinks (the closest pair of features changes)
        if smooth:
            assert np.allclose(g, fd, atol=2e-5), (ego, obs, g, fd)
            checked += 1
    assert checked >= 40
    # a hand-checkable case: two axis-aligned cars 7 m apart nose to tail: distance 7 - 4.47, gradient (+-1, 0, 0)
    d, g = obca.rect_distance(np.array([7.0, 0.0, 0.0]), np.array([0.0, 0.0, 0.0]))
    assert abs(d - (7.0 - 4.47)) < 1e-12 and np.allclose(g[:2], [1.0, 0.0])
