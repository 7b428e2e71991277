"""ORACLE tooling (test infrastructure) -- import the reference's own numpy-mode modules.

The reference solve path needs `casadi` and `polytope`, which are not installed and
cannot be installed offline.  Its numpy-mode pieces run unchanged once
those imports are stubbed: KinematicBicycleModelFrenet(mode='numpy'),
KinematicBicycleModel, ReferenceGenerator, utils.filter_preds / scenario_encoding /
augment_prev_sol / frenet2global.  This loader is used ONLY by
tests/golden/make_golden.py (to write the committed fixtures) and by tests that
re-generate them when /root/reference is mounted.  Nothing here is read on the GPU box.
"""
import math
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("IGT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "common"))


def _numeric_casadi_stub():
    """Tiny numeric stand-in for the handful of casadi calls made with plain floats
    (utils.py:532-586 uses ca.if_else / ca.sin / ca.cos on numbers)."""
    ca = types.ModuleType("casadi")
    ca.if_else = lambda c, a, b: a if c else b
    ca.sin = math.sin
    ca.cos = math.cos
    ca.tan = math.tan
    ca.atan = math.atan
    ca.fabs = abs
    return ca


def load():
    """Returns a namespace of the reference modules (numpy mode)."""
    if not available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    saved = {k: sys.modules.get(k) for k in
             ("casadi", "polytope", "matplotlib", "matplotlib.pyplot", "matplotlib.ticker",
              "matplotlib.collections")}
    sys.modules["casadi"] = _numeric_casadi_stub()
    pt = types.ModuleType("polytope")
    pt.Polytope = lambda *a, **k: None
    sys.modules["polytope"] = pt
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.collections"):
        m = types.ModuleType(name)
        m.LineCollection = object
        sys.modules[name] = m
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "common"))
    try:
        import importlib
        ns = types.SimpleNamespace()
        for mod in ("VehicleState", "VehicleAction", "VehicleReference",
                    "kinematic_bicycle_model_frenet", "kinematic_bicycle_model",
                    "ReferenceGen", "utils"):
            if mod in sys.modules:
                del sys.modules[mod]
            setattr(ns, mod, importlib.import_module(mod))
        return ns
    finally:
        sys.path.pop(0)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
