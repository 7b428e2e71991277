"""ORACLE (test infrastructure) -- ctypes binding of oracle/igt_oracle.c (fp64, CPU).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  `build()` compiles the C file with gcc (seconds)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libigt_oracle.so")
MAX_CINF, MAX_LAYERS = 128, 5
_dp = C.POINTER(C.c_double)


class CParams(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("n_rk", C.c_int),
        ("dt", C.c_double), ("l_r", C.c_double), ("l_f", C.c_double),
        ("v_min", C.c_double), ("v_max", C.c_double), ("a_min", C.c_double), ("a_max", C.c_double),
        ("df_max", C.c_double), ("ey_lim", C.c_double), ("da_max", C.c_double), ("ddf_max", C.c_double),
        ("d_min", C.c_double), ("w_u", C.c_double),
        ("n_cinf", C.c_int),
        ("cinf_A", C.c_double * 2 * MAX_CINF), ("cinf_b", C.c_double * MAX_CINF),
        ("tol", C.c_double), ("tol_rp", C.c_double), ("tol_comp", C.c_double), ("mu0", C.c_double),
        ("mu_floor", C.c_double), ("kappa_eps", C.c_double), ("kappa_mu", C.c_double),
        ("theta_mu", C.c_double), ("y_init_min", C.c_double),
        ("tau_min", C.c_double), ("reg_min", C.c_double), ("reg_up", C.c_double), ("reg_down", C.c_double),
        ("reg_max", C.c_double), ("eps_phi", C.c_double), ("gamma_theta", C.c_double),
        ("theta_small", C.c_double),
        ("max_iter", C.c_int), ("n_alpha", C.c_int), ("second_order", C.c_int),
        ("max_ls_fail", C.c_int), ("max_trials", C.c_int), ("predict_alpha", C.c_int),
        ("alpha_safety", C.c_double),
        ("reg_jump", C.c_double),
        ("stall_iter", C.c_int), ("stall_rp", C.c_double),
        ("acc_tol", C.c_double), ("acc_rp", C.c_double), ("acc_comp", C.c_double),
        ("n_layers", C.c_int), ("dims", C.c_int * (MAX_LAYERS + 1)),
        ("W", _dp * MAX_LAYERS), ("b", _dp * MAX_LAYERS),
        ("Wn", C.c_double * 36), ("mu_f", C.c_double * 6), ("sigma_t", C.c_double), ("mu_t", C.c_double),
    ]


def build(force=False):
    src = os.path.join(_HERE, "igt_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        assert _lib.igt_oracle_params_size() == C.sizeof(CParams), "struct layout mismatch"
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class COracle:
    """fp64 CPU solver over a batch.  `P` is an oracle.nlp.Params; `mlp` an oracle.nlp.MLPTerm."""

    def __init__(self, P, mlp=None, **options):
        self.P = P
        self.lib = lib()
        cp = CParams()
        self.lib.igt_oracle_default_options(C.byref(cp))
        cp.N, cp.n_rk, cp.dt = P.N, P.n_rk, P.dt
        cp.l_r = cp.l_f = 2.235
        for f in ("v_min", "v_max", "a_min", "a_max", "df_max", "ey_lim", "d_min", "w_u"):
            setattr(cp, f, getattr(P, f))
        cp.da_max, cp.ddf_max = P.da_max, P.ddf_max
        cp.n_cinf = len(P.cinf_b)
        for m in range(cp.n_cinf):
            cp.cinf_A[m][0], cp.cinf_A[m][1] = P.cinf_A[m]
            cp.cinf_b[m] = P.cinf_b[m]
        self._keep = []
        cp.n_layers = 0
        if mlp is not None:
            cp.n_layers = len(mlp.weights)
            cp.dims[0] = mlp.weights[0][0].shape[1]
            for i, (W, b) in enumerate(mlp.weights):
                W, b = _c(W), _c(b)
                self._keep += [W, b]
                cp.W[i], cp.b[i] = _ptr(W), _ptr(b)
                cp.dims[i + 1] = W.shape[0]
            for i, v in enumerate(np.asarray(mlp.Wn, dtype=np.float64).ravel()):
                cp.Wn[i] = v
            for i, v in enumerate(np.asarray(mlp.mu_f, dtype=np.float64).ravel()):
                cp.mu_f[i] = v
            cp.sigma_t, cp.mu_t = float(mlp.sigma_t), float(mlp.mu_t)
        for k, v in options.items():
            if not hasattr(cp, k):
                raise KeyError(k)
            if k == "max_trials" and v <= 0:       # igt_params convention: 0 = no budget of forward passes
                v = 1000000
            setattr(cp, k, v)
        self.cp = cp

    def rollout(self, z0, U, curv=None, jac=False, model=0):
        z0, U = _c(z0), _c(U)
        B, N = U.shape[0], U.shape[1]
        assert N == self.P.N
        nz = 7 if model == 0 else 4
        curv = _c(curv) if curv is not None else np.zeros((B, 3))
        Z = np.empty((B, N + 1, nz))
        A = np.empty((B, N, nz, nz)) if jac else None
        Bm = np.empty((B, N, nz, 2)) if jac else None
        self.lib.igt_oracle_rollout(C.byref(self.cp), B, _ptr(z0), _ptr(U), _ptr(curv), _ptr(Z), _ptr(A),
                                    _ptr(Bm), model)
        return (Z, A, Bm) if jac else Z

    def eval(self, x0, u_prev, curv, obs, Z, U, nn_ctx=None):
        x0, u_prev, curv, obs, Z, U, nn_ctx = map(_c, (x0, u_prev, curv, obs, Z, U, nn_ctx))
        B = x0.shape[0]
        assert self.cp.n_layers == 0 or nn_ctx is not None, "this oracle has a value network: eval needs nn_ctx"
        cost, viol = np.empty(B), np.empty(B)
        self.lib.igt_oracle_eval(C.byref(self.cp), B, _ptr(x0), _ptr(u_prev), _ptr(curv), _ptr(obs),
                                 _ptr(nn_ctx), _ptr(Z), _ptr(U), _ptr(cost), _ptr(viol))
        return cost, viol

    def initial_guess(self, x0, u_prev, curv, obs):
        x0, u_prev, curv, obs = map(_c, (x0, u_prev, curv, obs))
        B = x0.shape[0]
        U = np.empty((B, self.P.N, 2))
        self.lib.igt_oracle_initial_guess(C.byref(self.cp), B, _ptr(x0), _ptr(u_prev), _ptr(curv), _ptr(obs), _ptr(U))
        return U

    def solve(self, x0, u_prev, curv, obs, nn_ctx=None, u_init=None, n_threads=0, obs_psi=None):
        """obs_psi[B,N+1] (obstacle heading forecast) selects the OBCA collision rows (mpc.py:211-221, d_min = 0:
        build the oracle with nlp.Params(d_min=0.0))."""
        x0, u_prev, curv, obs, nn_ctx, u_init, obs_psi = map(_c, (x0, u_prev, curv, obs, nn_ctx, u_init, obs_psi))
        B, N = x0.shape[0], self.P.N
        assert obs.shape == (B, N + 1, 2) and (obs_psi is None or obs_psi.shape == (B, N + 1))
        if self.cp.n_layers > 0:
            assert nn_ctx is not None
        Z = np.empty((B, N + 1, 7)); U = np.empty((B, N, 2))
        cost, viol = np.empty(B), np.empty(B)
        status = np.empty(B, dtype=np.int32); iters = np.empty(B, dtype=np.int32)
        self.lib.igt_oracle_solve_batch_obca(C.byref(self.cp), B, _ptr(x0), _ptr(u_prev), _ptr(curv), _ptr(obs), _ptr(obs_psi),
                                        _ptr(nn_ctx), _ptr(u_init), _ptr(Z), _ptr(U), _ptr(cost), _ptr(viol),
                                        status.ctypes.data_as(C.POINTER(C.c_int)),
                                        iters.ctypes.data_as(C.POINTER(C.c_int)), n_threads)
        return dict(Z=Z, U=U, cost=cost, viol=viol, status=status, iters=iters)

    def viol_obca(self, x0, u_prev, curv, obs, obs_psi, Z, U):
        x0, u_prev, curv, obs, obs_psi, Z, U = map(_c, (x0, u_prev, curv, obs, obs_psi, Z, U))
        viol = np.empty(x0.shape[0])
        self.lib.igt_oracle_viol_obca(C.byref(self.cp), x0.shape[0], _ptr(x0), _ptr(u_prev), _ptr(curv), _ptr(obs), _ptr(obs_psi),
                                      _ptr(Z), _ptr(U), _ptr(viol))
        return viol


def rect_sdist(ego, obs):
    """C port of oracle/obca.rect_distance for n pose pairs: d[n], g[n,3]"""
    ego, obs = _c(ego), _c(obs)
    d = np.empty(ego.shape[0]); g = np.empty((ego.shape[0], 3))
    lib().igt_oracle_rect_sdist(ego.shape[0], _ptr(ego), _ptr(obs), _ptr(d), _ptr(g))
    return d, g
