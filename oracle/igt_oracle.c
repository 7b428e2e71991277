/* ORACLE -- test infrastructure, NOT product code.
 *
 * fp64 plain-C restatement of the reference's per-timestep MPC solve, used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker
 * and the timed CPU baseline.  The product path (igt_mpc_int_b200/csrc) never links this.
 *
 * PARITY UNPINNED against IPOPT: the reference solves this NLP with CasADi Opti + IPOPT
 * (reference mpc.py:129-146, :392).  casadi==3.6.5 / polytope==0.2.5
 * (common/environment.yml:30,83) are absent from /root/reference and not installable
 * offline, and the reference ships no test or golden vector for a solver result.  What IS
 * pinned: the dynamics below reproduce the reference's own numpy model
 * (common/kinematic_bicycle_model_frenet.py:70-127) on committed golden vectors
 * (tests/golden), and the solver is cross-checked against scipy.optimize on the same NLP.
 *
 * The NLP (reference mpc.py:147-160):
 *   variables x[7,N+1] = [x,y,s,ey,epsi,v,psi], u[2,N] = [a,df]        mpc.py:162-164
 *   box / rate / ey / terminal-set / collision rows                    mpc.py:177-321
 *   dynamics: 4-substep RK4 of the Frenet kinematic bicycle            mpc.py:201-209
 *   cost                                                               mpc.py:356-373
 * Algorithm: identical, step for step, to oracle/solver.py (interior-point iLQR).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <unistd.h>

#define NZ 7
#define NA 9
#define NW 11
enum { IX = 0, IY, IS, IEY, IEPSI, IV, IPSI, IPA, IPD, IUA, IUD };
#define MAX_CINF 128
#define MAX_MLP_LAYERS 5
#define MAX_MLP_WIDTH 256

typedef struct {
    int N, n_rk;
    double dt, l_r, l_f;
    double v_min, v_max, a_min, a_max, df_max, ey_lim, da_max, ddf_max, d_min, w_u;
    int n_cinf;
    double cinf_A[MAX_CINF][2], cinf_b[MAX_CINF];
    /* solver options (oracle/solver.py: Options) */
    double tol, tol_rp, tol_comp, mu0, mu_floor, kappa_eps, kappa_mu, theta_mu, y_init_min;
    double tau_min, reg_min, reg_up, reg_down, reg_max, eps_phi, gamma_theta, theta_small;
    int max_iter, n_alpha, second_order;
    int max_ls_fail, max_trials, predict_alpha;
    double alpha_safety;
    double reg_jump;                   /* after a non-PD Quu: reg >= reg_jump * (-lambda_min(Quu)) of that stage */
    int stall_iter; double stall_rp;   /* local-infeasibility exit: it >= stall_iter and |c + y|_inf > stall_rp -> status 5 */
    /* reference-tolerance exit (status 6): when the solve ends in a failure (iteration cap, regularisation limit,
     * budget of forward passes) at a point that already satisfies the reference's own IPOPT tolerances
     * (mpc.py:133-135: tol = dual_inf_tol = constr_viol_tol = 1e-3; IPOPT's default compl_inf_tol 1e-4) -- with the
     * primal bound tightened to the north star's 1e-6 -- the point is returned as "acceptable"; acc_tol = 0 disables */
    double acc_tol, acc_rp, acc_comp;
    /* gt_mpc value term (mpc.py:326-354,:367-369; model.py:14-67); n_layers = 0 -> 'mpc' mode */
    int n_layers;
    int dims[MAX_MLP_LAYERS + 1];
    const double *W[MAX_MLP_LAYERS], *b[MAX_MLP_LAYERS];
    double Wn[36], mu_f[6], sigma_t, mu_t;
} igt_oracle_params;

/* ------------------------------------------------------------------ dynamics ---------- */
static inline double curvature(double s, const double *curv)
{   /* casadi pw_const, mpc.py:199 */
    return curv[2] * (s >= curv[0] ? 1.0 : 0.0) - curv[2] * (s >= curv[1] ? 1.0 : 0.0);
}

typedef struct { double beta, dbeta, sb, cb; } slip_t;

static inline slip_t slip(const igt_oracle_params *P, double df)
{   /* kinematic_bicycle_model_frenet.py:72 */
    slip_t r;
    double rho = P->l_r / (P->l_f + P->l_r), t = tan(df);
    r.beta = atan(rho * t);
    r.dbeta = rho * (1.0 + t * t) / (1.0 + rho * rho * t * t);
    r.sb = sin(r.beta);
    r.cb = cos(r.beta);
    return r;
}

/* zdot = f(z,u) in planner order; optional dense Jacobians fz[7][7], fu[7][2]
 * kinematic_bicycle_model_frenet.py:71-91 */
static void rhs(const igt_oracle_params *P, const double *z, const double *u, const double *curv,
                const slip_t *sl, double *zd, double (*fz)[NZ], double (*fu)[2])
{
    double ey = z[IEY], epsi = z[IEPSI], v = z[IV], psi = z[IPSI];
    double K = curvature(z[IS], curv);
    double c1 = cos(sl->beta + epsi), s1 = sin(sl->beta + epsi);
    double cp = cos(psi + sl->beta), sp = sin(psi + sl->beta);
    double den = 1.0 - K * ey;
    double sdot = v * c1 / den;
    zd[IS] = sdot;
    zd[IEY] = v * s1;
    zd[IEPSI] = v * sl->sb / P->l_r - sdot * K;
    zd[IV] = u[0];
    zd[IX] = v * cp;
    zd[IY] = v * sp;
    zd[IPSI] = v * sl->sb / P->l_r;
    if (!fz) return;
    memset(fz, 0, sizeof(double) * NZ * NZ);
    memset(fu, 0, sizeof(double) * NZ * 2);
    double db = sl->dbeta;
    double ds_dey = v * c1 * K / (den * den), ds_depsi = -v * s1 / den, ds_dv = c1 / den;
    double ds_db = -v * s1 / den;
    fz[IS][IEY] = ds_dey; fz[IS][IEPSI] = ds_depsi; fz[IS][IV] = ds_dv; fu[IS][1] = ds_db * db;
    fz[IEY][IEPSI] = v * c1; fz[IEY][IV] = s1; fu[IEY][1] = v * c1 * db;
    fz[IEPSI][IEY] = -K * ds_dey; fz[IEPSI][IEPSI] = -K * ds_depsi;
    fz[IEPSI][IV] = sl->sb / P->l_r - K * ds_dv;
    fu[IEPSI][1] = (v * sl->cb / P->l_r - K * ds_db) * db;
    fu[IV][0] = 1.0;
    fz[IX][IV] = cp; fz[IX][IPSI] = -v * sp; fu[IX][1] = -v * sp * db;
    fz[IY][IV] = sp; fz[IY][IPSI] = v * cp; fu[IY][1] = v * cp * db;
    fz[IPSI][IV] = sl->sb / P->l_r; fu[IPSI][1] = v * sl->cb / P->l_r * db;
}

/* one MPC step: n_rk RK4 sub-steps, k4 evaluates xdot,ydot at psi + h/2*k3[psi]
 * (kinematic_bicycle_model_frenet.py:93-127, quirk at :111).  A[7][7], B[7][2] optional. */
static void rk4_step(const igt_oracle_params *P, const double *z0, const double *u, const double *curv,
                     double *zn, double (*A)[NZ], double (*B)[2])
{
    double h = P->dt / P->n_rk;
    double z[NZ], zs[NZ], k1[NZ], k2[NZ], k3[NZ], k4[NZ];
    double S[NZ][NA], T[NZ][NA], d1[NZ][NA], d2[NZ][NA], d3[NZ][NA], d4[NZ][NA];
    double fz[NZ][NZ], fu[NZ][2];
    slip_t sl = slip(P, u[1]);
    int want = (A != NULL);
    memcpy(z, z0, sizeof(z));
    if (want) {
        memset(S, 0, sizeof(S));
        for (int i = 0; i < NZ; i++) S[i][i] = 1.0;
    }
#define STAGE(kprev, dprev, kout, dout, coef, coef_psi)                                      \
    do {                                                                                     \
        for (int i = 0; i < NZ; i++) zs[i] = z[i] + h * (coef) * (kprev)[i];                 \
        zs[IPSI] = z[IPSI] + h * (coef_psi) * (kprev)[IPSI];                                 \
        rhs(P, zs, u, curv, &sl, kout, want ? fz : NULL, want ? fu : NULL);                  \
        if (want) {                                                                          \
            for (int i = 0; i < NZ; i++)                                                     \
                for (int j = 0; j < NA; j++)                                                 \
                    T[i][j] = S[i][j] + h * ((i == IPSI) ? (coef_psi) : (coef)) * (dprev)[i][j]; \
            for (int i = 0; i < NZ; i++)                                                     \
                for (int j = 0; j < NA; j++) {                                               \
                    double acc = (j >= NZ) ? fu[i][j - NZ] : 0.0;                            \
                    for (int a = 0; a < NZ; a++) acc += fz[i][a] * T[a][j];                  \
                    (dout)[i][j] = acc;                                                      \
                }                                                                            \
        }                                                                                    \
    } while (0)
    for (int it = 0; it < P->n_rk; it++) {
        rhs(P, z, u, curv, &sl, k1, want ? fz : NULL, want ? fu : NULL);
        if (want)
            for (int i = 0; i < NZ; i++)
                for (int j = 0; j < NA; j++) {
                    double acc = (j >= NZ) ? fu[i][j - NZ] : 0.0;
                    for (int a = 0; a < NZ; a++) acc += fz[i][a] * S[a][j];
                    d1[i][j] = acc;
                }
        STAGE(k1, d1, k2, d2, 0.5, 0.5);
        STAGE(k2, d2, k3, d3, 0.5, 0.5);
        STAGE(k3, d3, k4, d4, 1.0, 0.5);
        for (int i = 0; i < NZ; i++) z[i] += h / 6.0 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
        if (want)
            for (int i = 0; i < NZ; i++)
                for (int j = 0; j < NA; j++)
                    S[i][j] += h / 6.0 * (d1[i][j] + 2 * d2[i][j] + 2 * d3[i][j] + d4[i][j]);
    }
#undef STAGE
    memcpy(zn, z, sizeof(z));
    if (want)
        for (int i = 0; i < NZ; i++) {
            for (int j = 0; j < NZ; j++) A[i][j] = S[i][j];
            B[i][0] = S[i][NZ];
            B[i][1] = S[i][NZ + 1];
        }
}

/* Cartesian Euler model, z = (x, y, psi, v): kinematic_bicycle_model.py:27-31 */
static void euler_step(const igt_oracle_params *P, const double *z, const double *u, double *zn,
                       double *A /*[16]*/, double *B /*[8]*/)
{
    double x = z[0], y = z[1], psi = z[2], v = z[3], dt = P->dt;
    slip_t sl = slip(P, u[1]);
    double lsum = P->l_r + P->l_f, t = tan(u[1]);
    double cp = cos(psi + sl.beta), sp = sin(psi + sl.beta);
    zn[0] = x + dt * v * cp;
    zn[1] = y + dt * v * sp;
    zn[2] = psi + dt * (v * sl.cb / lsum * t);
    zn[3] = v + dt * u[0];
    if (!A) return;
    memset(A, 0, 16 * sizeof(double));
    memset(B, 0, 8 * sizeof(double));
    A[0] = A[5] = A[10] = A[15] = 1.0;
    A[0 * 4 + 2] = -dt * v * sp; A[0 * 4 + 3] = dt * cp;
    A[1 * 4 + 2] = dt * v * cp;  A[1 * 4 + 3] = dt * sp;
    A[2 * 4 + 3] = dt * sl.cb / lsum * t;
    B[0 * 2 + 1] = -dt * v * sp * sl.dbeta;
    B[1 * 2 + 1] = dt * v * cp * sl.dbeta;
    B[2 * 2 + 1] = dt * v / lsum * (-sl.sb * sl.dbeta * t + sl.cb * (1.0 + t * t));
    B[3 * 2 + 0] = dt;
}

/* ------------------------------------------------------------------ MLP term ---------- */
/* value V*sigma_t + mu_t and its gradient / Hessian w.r.t. (s_N, v_N) by forward tangents */
static void mlp_value(const igt_oracle_params *P, double sN, double vN, const double *ctx,
                      double *V, double *g, double *H)
{
    double xN[6] = { ctx[0], ctx[1], ctx[2], sN - ctx[0], vN - ctx[1], ctx[3] - ctx[2] };
    static __thread double h[6][MAX_MLP_WIDTH], hn[6][MAX_MLP_WIDTH];
    /* rows: 0 value, 1 d/ds, 2 d/dv, 3 d2/dss, 4 d2/dsv, 5 d2/dvv */
    for (int i = 0; i < 6; i++) {
        double acc = 0;
        for (int j = 0; j < 6; j++) acc += P->Wn[i * 6 + j] * (xN[j] - P->mu_f[j]);
        h[0][i] = acc;
        h[1][i] = P->Wn[i * 6 + 3];
        h[2][i] = P->Wn[i * 6 + 4];
        h[3][i] = h[4][i] = h[5][i] = 0.0;
    }
    for (int l = 0; l < P->n_layers; l++) {
        int nin = P->dims[l], nout = P->dims[l + 1];
        for (int r = 0; r < 6; r++)
            for (int o = 0; o < nout; o++) {
                double acc = (r == 0) ? P->b[l][o] : 0.0;
                const double *w = P->W[l] + (size_t)o * nin;
                for (int i = 0; i < nin; i++) acc += w[i] * h[r][i];
                hn[r][o] = acc;
            }
        if (l < P->n_layers - 1) {
            for (int o = 0; o < nout; o++) {
                double y = tanh(hn[0][o]), d1 = 1.0 - y * y, d2 = -2.0 * y * d1;
                double ts = hn[1][o], tv = hn[2][o];
                h[0][o] = y;
                h[3][o] = d1 * hn[3][o] + d2 * ts * ts;
                h[4][o] = d1 * hn[4][o] + d2 * ts * tv;
                h[5][o] = d1 * hn[5][o] + d2 * tv * tv;
                h[1][o] = d1 * ts;
                h[2][o] = d1 * tv;
            }
        } else {
            for (int r = 0; r < 6; r++) h[r][0] = hn[r][0];
        }
    }
    *V = h[0][0] * P->sigma_t + P->mu_t;
    if (g) { g[0] = h[1][0] * P->sigma_t; g[1] = h[2][0] * P->sigma_t; }
    if (H) { H[0] = h[3][0] * P->sigma_t; H[1] = h[4][0] * P->sigma_t; H[2] = h[5][0] * P->sigma_t; }
}

/* ------------------------------------------------------------------ NLP rows ---------- */
typedef struct {
    double c;            /* row value, c <= 0 feasible */
    int i0, i1, i2;      /* indices into w = (zeta[9], u[2]); i1 / i2 < 0 if unused (i2: heading entry of the OBCA row) */
    double g0, g1, g2;   /* gradient entries */
    int has_h;           /* collision row: 2x2 Hessian block on (x, y) */
    double hxx, hxy, hyy;
} row_t;

typedef struct {
    const double *x0, *u_prev, *curv, *obs, *ctx, *u_init;
    const double *obs_psi;   /* [N+1] obstacle heading forecast -> OBCA collision rows (mpc.py:211-221); NULL -> circle */
} prob_t;

static inline int n_rows_stage(const igt_oracle_params *P, int k)
{
    if (k == P->N) return 3;
    return (k >= 1 ? 5 : 0) + 8 + (k == P->N - 1 ? P->n_cinf : 0);
}

static inline void set_row(row_t *r, double c, int i0, double g0, int i1, double g1)
{
    r->c = c; r->i0 = i0; r->g0 = g0; r->i1 = i1; r->g1 = g1; r->i2 = -1; r->g2 = 0.0; r->has_h = 0;
}

/* ---- OBCA rows in dual-eliminated form (oracle/obca.py restated in C; pinned there against the reference's
 * explicit-dual formulation, tests/test_oracle_obca.py): signed distance between two 4.47 x 2.0 rectangles ---- */
#define OBCA_MARGIN 1e-6   /* mpc.py:216 */
#define VEH_HL (4.47 / 2)
#define VEH_HW (2.0 / 2)
static double sgn1(double v) { return v >= 0 ? 1.0 : -1.0; }

static double point_to_rect(const double q[2], const double pose[3], double cl_local[2], double cl_world[2], int *inside)
{
    double c = cos(pose[2]), s = sin(pose[2]);
    double dx = c * (q[0] - pose[0]) + s * (q[1] - pose[1]), dy = -s * (q[0] - pose[0]) + c * (q[1] - pose[1]);
    cl_local[0] = fmin(fmax(dx, -VEH_HL), VEH_HL); cl_local[1] = fmin(fmax(dy, -VEH_HW), VEH_HW);
    *inside = fabs(dx) <= VEH_HL && fabs(dy) <= VEH_HW;
    cl_world[0] = pose[0] + c * cl_local[0] - s * cl_local[1];
    cl_world[1] = pose[1] + s * cl_local[0] + c * cl_local[1];
    return hypot(dx - cl_local[0], dy - cl_local[1]);
}

static double rect_penetration(const double ego[3], const double obs[3], double g[3])
{   /* separating-axis test over the four face normals: minus the smallest overlap, gradient w.r.t. the ego pose */
    double uxe[2] = { cos(ego[2]), sin(ego[2]) }, uye[2] = { -uxe[1], uxe[0] };
    double uxo[2] = { cos(obs[2]), sin(obs[2]) }, uyo[2] = { -uxo[1], uxo[0] };
    double dv[2] = { ego[0] - obs[0], ego[1] - obs[1] };
    double best = 1e300, bg[3] = { 0, 0, 0 };
    for (int f = 0; f < 2; f++) {
        const double *a = f == 0 ? uxe : uye;
        double da[2] = { f == 0 ? uye[0] : -uxe[0], f == 0 ? uye[1] : -uxe[1] }, he = f == 0 ? VEH_HL : VEH_HW;
        double axo = a[0] * uxo[0] + a[1] * uxo[1], ayo = a[0] * uyo[0] + a[1] * uyo[1], ad = a[0] * dv[0] + a[1] * dv[1];
        double ov = he + VEH_HL * fabs(axo) + VEH_HW * fabs(ayo) - fabs(ad);
        if (ov < best) {
            best = ov; bg[0] = -sgn1(ad) * a[0]; bg[1] = -sgn1(ad) * a[1];
            bg[2] = VEH_HL * sgn1(axo) * (da[0] * uxo[0] + da[1] * uxo[1]) + VEH_HW * sgn1(ayo) * (da[0] * uyo[0] + da[1] * uyo[1])
                    - sgn1(ad) * (da[0] * dv[0] + da[1] * dv[1]);
        }
    }
    for (int f = 0; f < 2; f++) {
        const double *a = f == 0 ? uxo : uyo;
        double ho = f == 0 ? VEH_HL : VEH_HW;
        double axe = a[0] * uxe[0] + a[1] * uxe[1], aye = a[0] * uye[0] + a[1] * uye[1], ad = a[0] * dv[0] + a[1] * dv[1];
        double ov = VEH_HL * fabs(axe) + VEH_HW * fabs(aye) + ho - fabs(ad);
        if (ov < best) {
            best = ov; bg[0] = -sgn1(ad) * a[0]; bg[1] = -sgn1(ad) * a[1];
            bg[2] = VEH_HL * sgn1(axe) * aye + VEH_HW * sgn1(aye) * (-axe);
        }
    }
    g[0] = -bg[0]; g[1] = -bg[1]; g[2] = -bg[2];
    return -best;
}

static double rect_signed_distance(const double ego[3], const double obs[3], double g[3])
{
    static const double loc[4][2] = { { VEH_HL, VEH_HW }, { -VEH_HL, VEH_HW }, { -VEH_HL, -VEH_HW }, { VEH_HL, -VEH_HW } };
    double Pe[4][2], Po[4][2];
    double ce = cos(ego[2]), se = sin(ego[2]), co = cos(obs[2]), so = sin(obs[2]);
    for (int i = 0; i < 4; i++) {
        Pe[i][0] = ego[0] + ce * loc[i][0] - se * loc[i][1]; Pe[i][1] = ego[1] + se * loc[i][0] + ce * loc[i][1];
        Po[i][0] = obs[0] + co * loc[i][0] - so * loc[i][1]; Po[i][1] = obs[1] + so * loc[i][0] + co * loc[i][1];
    }
    double best = 1e300, a[2] = { 0, 0 }, b[2] = { 0, 0 }, al[2] = { 0, 0 };
    for (int i = 0; i < 4; i++) {
        double cl[2], cw[2]; int in;
        double d = point_to_rect(Pe[i], obs, cl, cw, &in);
        if (in) return rect_penetration(ego, obs, g);
        if (d < best) { best = d; a[0] = Pe[i][0]; a[1] = Pe[i][1]; b[0] = cw[0]; b[1] = cw[1]; al[0] = loc[i][0]; al[1] = loc[i][1]; }
    }
    for (int j = 0; j < 4; j++) {
        double cl[2], cw[2]; int in;
        double d = point_to_rect(Po[j], ego, cl, cw, &in);
        if (in) return rect_penetration(ego, obs, g);
        if (d < best) { best = d; a[0] = cw[0]; a[1] = cw[1]; b[0] = Po[j][0]; b[1] = Po[j][1]; al[0] = cl[0]; al[1] = cl[1]; }
    }
    for (int i = 0; i < 4; i++)                                   /* crossing edges without a vertex inside */
        for (int j = 0; j < 4; j++) {
            double d1x = Pe[(i + 1) & 3][0] - Pe[i][0], d1y = Pe[(i + 1) & 3][1] - Pe[i][1];
            double d2x = Po[(j + 1) & 3][0] - Po[j][0], d2y = Po[(j + 1) & 3][1] - Po[j][1];
            double den = d1x * d2y - d1y * d2x;
            if (fabs(den) < 1e-14) continue;
            double rx = Po[j][0] - Pe[i][0], ry = Po[j][1] - Pe[i][1];
            double t = (rx * d2y - ry * d2x) / den, u = (rx * d1y - ry * d1x) / den;
            if (t >= 0 && t <= 1 && u >= 0 && u <= 1) return rect_penetration(ego, obs, g);
        }
    double nx = (a[0] - b[0]) / best, ny = (a[1] - b[1]) / best;
    g[0] = nx; g[1] = ny;
    g[2] = nx * (-se * al[0] - ce * al[1]) + ny * (ce * al[0] - se * al[1]);
    return best;
}

/* value of the collision row of node k in the units the reference writes it in */
static double collision_row_ref(const igt_oracle_params *P, const prob_t *pr, int k, const double *z)
{
    if (pr->obs_psi) {
        double ego[3] = { z[IX], z[IY], z[IPSI] }, ob[3] = { pr->obs[2 * k], pr->obs[2 * k + 1], pr->obs_psi[k] }, g[3];
        return P->d_min + OBCA_MARGIN - rect_signed_distance(ego, ob, g);          /* mpc.py:216 */
    }
    double dx = z[IX] - pr->obs[2 * k], dy = z[IY] - pr->obs[2 * k + 1];
    return P->d_min * P->d_min - dx * dx - dy * dy;                                 /* mpc.py:226 */
}

static void collision_row(const igt_oracle_params *P, const prob_t *pr, int k, const double *z, row_t *r)
{
    const double *o = pr->obs + 2 * k;
    if (pr->obs_psi) {   /* mpc.py:211-221 with the duals maximised out: margin - signed rectangle distance, Gauss-Newton */
        double ego[3] = { z[IX], z[IY], z[IPSI] }, ob[3] = { o[0], o[1], pr->obs_psi[k] }, g[3];
        double sd = rect_signed_distance(ego, ob, g);
        set_row(r, P->d_min + OBCA_MARGIN - sd, IX, -g[0], IY, -g[1]);
        r->i2 = IPSI; r->g2 = -g[2];
        return;
    }
    /* mpc.py:226 in the equivalent distance form d_min - |p - o| <= 0 */
    double dx = z[IX] - o[0], dy = z[IY] - o[1];
    double dist = hypot(dx, dy);
    if (dist < 1e-9) dist = 1e-9;
    double nx = dx / dist, ny = dy / dist;
    set_row(r, P->d_min - dist, IX, -nx, IY, -ny);
    r->has_h = 1;
    r->hxx = -(1 - nx * nx) / dist; r->hxy = nx * ny / dist; r->hyy = -(1 - ny * ny) / dist;
}

/* rows of stage k (k == N: terminal).  Same order as oracle/solver.py stage_rows/terminal_rows */
static int rows_eval(const igt_oracle_params *P, const prob_t *pr, int k, const double *z,
                     const double *up, const double *u, row_t *r)
{
    int n = 0, N = P->N;
    if (k == N) {
        set_row(&r[n++], z[IEY] - P->ey_lim, IEY, 1.0, -1, 0);
        set_row(&r[n++], -P->ey_lim - z[IEY], IEY, -1.0, -1, 0);
        collision_row(P, pr, N, z, &r[n++]);
        return n;
    }
    if (k >= 1) {
        set_row(&r[n++], z[IV] - P->v_max, IV, 1.0, -1, 0);       /* mpc.py:317 */
        set_row(&r[n++], P->v_min - z[IV], IV, -1.0, -1, 0);      /* mpc.py:316 */
        set_row(&r[n++], z[IEY] - P->ey_lim, IEY, 1.0, -1, 0);    /* mpc.py:298 */
        set_row(&r[n++], -P->ey_lim - z[IEY], IEY, -1.0, -1, 0);  /* mpc.py:299 */
        collision_row(P, pr, k, z, &r[n++]);                      /* mpc.py:226 / :211-221 */
    }
    set_row(&r[n++], u[0] - P->a_max, IUA, 1.0, -1, 0);           /* mpc.py:319 */
    set_row(&r[n++], P->a_min - u[0], IUA, -1.0, -1, 0);          /* mpc.py:318 */
    set_row(&r[n++], u[1] - P->df_max, IUD, 1.0, -1, 0);          /* mpc.py:321 */
    set_row(&r[n++], -P->df_max - u[1], IUD, -1.0, -1, 0);        /* mpc.py:320 */
    set_row(&r[n++], (u[0] - up[0]) - P->da_max, IPA, -1.0, IUA, 1.0);    /* mpc.py:303-311 */
    set_row(&r[n++], -(u[0] - up[0]) - P->da_max, IPA, 1.0, IUA, -1.0);
    set_row(&r[n++], (u[1] - up[1]) - P->ddf_max, IPD, -1.0, IUD, 1.0);
    set_row(&r[n++], -(u[1] - up[1]) - P->ddf_max, IPD, 1.0, IUD, -1.0);
    if (k == N - 1)
        for (int m = 0; m < P->n_cinf; m++)                       /* mpc.py:177-180 */
            set_row(&r[n++], P->cinf_A[m][0] * z[IV] + P->cinf_A[m][1] * u[0] - P->cinf_b[m],
                    IV, P->cinf_A[m][0], IUA, P->cinf_A[m][1]);
    return n;
}

static double cost_fn(const igt_oracle_params *P, const prob_t *pr, const double *Z, const double *U)
{   /* mpc.py:356-373 */
    int N = P->N;
    double J = 0;
    for (int k = 0; k <= N; k++) J += Z[k * NZ + IEPSI] * Z[k * NZ + IEPSI] + Z[k * NZ + IEY] * Z[k * NZ + IEY];
    double su = 0;
    for (int k = 0; k < N; k++) su += U[2 * k] * U[2 * k] + U[2 * k + 1] * U[2 * k + 1];
    J += P->w_u * su;
    if (P->n_layers == 0) J -= Z[N * NZ + IS] - Z[IS];
    else { double V; mlp_value(P, Z[N * NZ + IS], Z[N * NZ + IV], pr->ctx, &V, NULL, NULL); J -= V; }
    return J;
}

/* max over all inequality rows as written in mpc.py (collision in squared-distance units) */
static double max_violation(const igt_oracle_params *P, const prob_t *pr, const double *Z, const double *U)
{
    int N = P->N;
    double m = 0;
#define UPD(val) do { double _v = (val); if (_v > m) m = _v; } while (0)
    for (int k = 0; k < N; k++) {
        const double *z = Z + k * NZ, *u = U + 2 * k;
        const double *up = k == 0 ? pr->u_prev : U + 2 * (k - 1);
        UPD(P->v_min - z[IV]); UPD(z[IV] - P->v_max);
        UPD(P->a_min - u[0]); UPD(u[0] - P->a_max); UPD(-P->df_max - u[1]); UPD(u[1] - P->df_max);
        UPD(fabs(u[0] - up[0]) - P->da_max); UPD(fabs(u[1] - up[1]) - P->ddf_max);
    }
    for (int k = 0; k <= N; k++) UPD(fabs(Z[k * NZ + IEY]) - P->ey_lim);
    for (int m_ = 0; m_ < P->n_cinf; m_++)
        UPD(P->cinf_A[m_][0] * Z[(N - 1) * NZ + IV] + P->cinf_A[m_][1] * U[2 * (N - 1)] - P->cinf_b[m_]);
    for (int k = 1; k <= N; k++) UPD(collision_row_ref(P, pr, k, Z + k * NZ));
#undef UPD
    return m;
}

/* ------------------------------------------------------------------ initial guess ----- */
static const double GUESS_SPEEDS[5] = { 5.0, 3.5, 2.0, 1.0, 0.0 };
#define GUESS_RHO 100.0

static void tracking_rollout(const igt_oracle_params *P, const prob_t *pr, double vt, double *U, double *Z)
{   /* oracle/nlp.py: tracking_rollout */
    int N = P->N;
    double z[NZ], up[2];
    memcpy(z, pr->x0, sizeof(z));
    memcpy(Z, z, sizeof(z));
    up[0] = pr->u_prev[0]; up[1] = pr->u_prev[1];
    for (int k = 0; k < N; k++) {
        double K = curvature(z[IS], pr->curv);
        double a = 0.5 * (vt - z[IV]);
        a = fmin(fmax(a, up[0] - P->da_max), up[0] + P->da_max);
        a = fmin(fmax(a, P->a_min), P->a_max);
        double dff = atan(2.0 * tan(asin(K * P->l_r)));
        double d = dff - 0.3 * z[IEY] - 0.8 * z[IEPSI];
        d = fmin(fmax(d, up[1] - P->ddf_max), up[1] + P->ddf_max);
        d = fmin(fmax(d, -P->df_max), P->df_max);
        U[2 * k] = a; U[2 * k + 1] = d;
        rk4_step(P, z, U + 2 * k, pr->curv, z, NULL, NULL);
        memcpy(Z + (k + 1) * NZ, z, sizeof(z));
        up[0] = a; up[1] = d;
    }
}

static double guess_merit(const igt_oracle_params *P, const prob_t *pr, const double *Z, const double *U)
{
    int N = P->N;
    double J = 0, su = 0, viol = 0;
    for (int k = 0; k <= N; k++) J += Z[k * NZ + IEPSI] * Z[k * NZ + IEPSI] + Z[k * NZ + IEY] * Z[k * NZ + IEY];
    for (int k = 0; k < N; k++) su += U[2 * k] * U[2 * k] + U[2 * k + 1] * U[2 * k + 1];
    J += P->w_u * su - (Z[N * NZ + IS] - Z[IS]);
    for (int k = 1; k < N; k++) {
        viol += fmax(0.0, Z[k * NZ + IV] - P->v_max) + fmax(0.0, P->v_min - Z[k * NZ + IV]);
    }
    for (int k = 1; k <= N; k++) {
        viol += fmax(0.0, fabs(Z[k * NZ + IEY]) - P->ey_lim);
        if (pr->obs_psi) viol += fmax(0.0, collision_row_ref(P, pr, k, Z + k * NZ));
        else {
            double dx = Z[k * NZ + IX] - pr->obs[2 * k], dy = Z[k * NZ + IY] - pr->obs[2 * k + 1];
            viol += fmax(0.0, P->d_min - sqrt(dx * dx + dy * dy));
        }
    }
    for (int m = 0; m < P->n_cinf; m++)
        viol += fmax(0.0, P->cinf_A[m][0] * Z[(N - 1) * NZ + IV] + P->cinf_A[m][1] * U[2 * (N - 1)] - P->cinf_b[m]);
    return J + GUESS_RHO * viol;
}

/* ------------------------------------------------------------------ solver ------------ */
typedef struct {
    int N, M;                 /* horizon, total rows */
    int *off;                 /* row offset per stage [N+2] */
    double *Z, *U, *Zn, *Un, *Zg, *Ug;
    double *Y, *S, *Yn, *Sn, *Cn;
    row_t *rows, *rtmp;       /* rows of the current iterate (all stages), scratch for a stage */
    double *A, *B;            /* [N][7][7], [N][7][2] */
    double *lam;              /* [N+1][9] adjoint */
    double *ku, *Ku;          /* [N][2], [N][2][9] */
} work_t;

static work_t *work_alloc(const igt_oracle_params *P)
{
    int N = P->N;
    work_t *w = (work_t *)calloc(1, sizeof(work_t));
    w->N = N;
    w->off = (int *)malloc(sizeof(int) * (N + 2));
    int M = 0, mmax = 0;
    for (int k = 0; k <= N; k++) { w->off[k] = M; int n = n_rows_stage(P, k); M += n; if (n > mmax) mmax = n; }
    w->off[N + 1] = M;
    w->M = M;
#define DAL(n) (double *)malloc(sizeof(double) * (size_t)(n))
    w->Z = DAL((N + 1) * NZ); w->Zn = DAL((N + 1) * NZ); w->Zg = DAL((N + 1) * NZ);
    w->U = DAL(N * 2); w->Un = DAL(N * 2); w->Ug = DAL(N * 2);
    w->Y = DAL(M); w->S = DAL(M); w->Yn = DAL(M); w->Sn = DAL(M); w->Cn = DAL(M);
    w->rows = (row_t *)malloc(sizeof(row_t) * M);
    w->rtmp = (row_t *)malloc(sizeof(row_t) * mmax);
    w->A = DAL(N * NZ * NZ); w->B = DAL(N * NZ * 2);
    w->lam = DAL((N + 1) * NA);
    w->ku = DAL(N * 2); w->Ku = DAL(N * 2 * NA);
#undef DAL
    return w;
}

static void work_free(work_t *w)
{
    free(w->off); free(w->Z); free(w->Zn); free(w->Zg); free(w->U); free(w->Un); free(w->Ug);
    free(w->Y); free(w->S); free(w->Yn); free(w->Sn); free(w->Cn); free(w->rows); free(w->rtmp);
    free(w->A); free(w->B); free(w->lam); free(w->ku); free(w->Ku); free(w);
}

/* dt * Hess(lam . f) on (ey, epsi, v, psi, df): oracle/solver.py dyn_hessian.  Adds into the
 * w-space Hessian H[11][11]. */
static void add_dyn_hessian(const igt_oracle_params *P, const double *z, const double *u,
                            const double *curv, const double *lam, double (*H)[NW])
{
    double ey = z[IEY], epsi = z[IEPSI], v = z[IV], psi = z[IPSI], df = u[1];
    double K = curvature(z[IS], curv);
    double rho = P->l_r / (P->l_f + P->l_r), t = tan(df);
    double beta = atan(rho * t), q = 1.0 + rho * rho * t * t;
    double b1 = rho * (1.0 + t * t) / q;
    double b2 = rho * 2.0 * t * (1.0 - rho * rho) / (q * q) * (1.0 + t * t);
    double th = beta + epsi, ph = psi + beta;
    double cth = cos(th), sth = sin(th), cph = cos(ph), sph = sin(ph), cb = cos(beta), sb = sin(beta);
    double Dn = 1.0 - K * ey;
    double m = lam[IS] - K * lam[IEPSI], n = (lam[IEPSI] + lam[IPSI]) / P->l_r;
    double ley = lam[IEY], lx = lam[IX], ly = lam[IY];
    double G_ey = m * cth * K / (Dn * Dn);
    double G_epsi = -m * sth / Dn + ley * cth;
    double G_psi = -lx * sph + ly * cph;
    double G_b = G_epsi + n * cb + G_psi;
    double G_eyey = m * cth * 2 * K * K / (Dn * Dn * Dn);
    double G_eyepsi = -m * sth * K / (Dn * Dn);
    double G_epsiepsi = -m * cth / Dn - ley * sth;
    double G_psipsi = -lx * cph - ly * sph;
    double G_bb = G_epsiepsi - n * sb + G_psipsi;
    double dt = P->dt;
#define ADDS(i, j, val) do { double _v = dt * (val); H[i][j] += _v; if ((i) != (j)) H[j][i] += _v; } while (0)
    ADDS(IEY, IEY, v * G_eyey);
    ADDS(IEY, IEPSI, v * G_eyepsi);
    ADDS(IEPSI, IEPSI, v * G_epsiepsi);
    ADDS(IPSI, IPSI, v * G_psipsi);
    ADDS(IV, IEY, G_ey);
    ADDS(IV, IEPSI, G_epsi);
    ADDS(IV, IPSI, G_psi);
    ADDS(IUD, IEY, v * G_eyepsi * b1);
    ADDS(IUD, IEPSI, v * G_epsiepsi * b1);
    ADDS(IUD, IPSI, v * G_psipsi * b1);
    ADDS(IUD, IV, G_b * b1);
    ADDS(IUD, IUD, v * (G_bb * b1 * b1 + G_b * b2));
#undef ADDS
}

/* status: 0 converged, 1 iteration limit, 2 x0 infeasible, 3 regularisation limit, 4 line search,
 * 5 stalled at an infeasible point, 6 stopped making progress at a point within the reference's tolerances */
static int solve_one(const igt_oracle_params *P, const prob_t *pr, work_t *w,
                     double *Zout, double *Uout, double *cost_out, double *viol_out, int *iters_out)
{
    const int N = P->N;
    int status = 1;
    *iters_out = 0;
    {   /* rows on x0 alone: mpc.py:316-317 and :298-299 at k = 0 */
        double v = pr->x0[IV], ey = pr->x0[IEY], tol = 1e-6;   /* igt_params.x0_tol */
        if (!(v >= P->v_min - tol && v <= P->v_max + tol && fabs(ey) <= P->ey_lim + tol)) {
            for (int i = 0; i < (N + 1) * NZ; i++) Zout[i] = NAN;
            for (int i = 0; i < N * 2; i++) Uout[i] = NAN;
            *cost_out = NAN; *viol_out = INFINITY;
            return 2;
        }
    }
    double *Z = w->Z, *U = w->U, *Y = w->Y, *S = w->S;
    /* initial guess */
    if (pr->u_init) {
        memcpy(U, pr->u_init, sizeof(double) * 2 * N);
        memcpy(Z, pr->x0, sizeof(double) * NZ);
        for (int k = 0; k < N; k++) rk4_step(P, Z + k * NZ, U + 2 * k, pr->curv, Z + (k + 1) * NZ, NULL, NULL);
    } else {
        double best = INFINITY;
        for (int g = 0; g < 5; g++) {
            tracking_rollout(P, pr, GUESS_SPEEDS[g], w->Ug, w->Zg);
            double m = guess_merit(P, pr, w->Zg, w->Ug);
            if (m < best) {
                best = m;
                memcpy(U, w->Ug, sizeof(double) * 2 * N);
                memcpy(Z, w->Zg, sizeof(double) * NZ * (N + 1));
            }
        }
    }
    double mu = P->mu0, reg = 0.0;
    for (int k = 0; k <= N; k++) {
        const double *up = k == 0 ? pr->u_prev : U + 2 * (k - 1);
        int n = rows_eval(P, pr, k, Z + k * NZ, up, k < N ? U + 2 * k : NULL, w->rows + w->off[k]);
        for (int i = 0; i < n; i++) {
            double c = w->rows[w->off[k] + i].c;
            double y = fmax(-c, P->y_init_min);
            Y[w->off[k] + i] = y;
            S[w->off[k] + i] = mu / y;
        }
    }
    int need_jac = 1, ls_fail = 0, trials = 0;
    double stat = 0, rp = 0, s_max = 0, sy_min = 0, sy_max = 0;
    double lxN[NA], lxxN_ss = 0, lxxN_sv = 0, lxxN_vv = 0;
    int it;
    for (it = 0; it <= P->max_iter; it++) {
        *iters_out = it;
        if (need_jac) {
            /* sweep 1: sensitivities + rows of the current trajectory */
            for (int k = 0; k < N; k++) {
                double zn[NZ];
                rk4_step(P, Z + k * NZ, U + 2 * k, pr->curv, zn,
                         (double (*)[NZ])(w->A + k * NZ * NZ), (double (*)[2])(w->B + k * NZ * 2));
            }
            for (int k = 0; k <= N; k++) {
                const double *up = k == 0 ? pr->u_prev : U + 2 * (k - 1);
                rows_eval(P, pr, k, Z + k * NZ, up, k < N ? U + 2 * k : NULL, w->rows + w->off[k]);
            }
            /* terminal cost derivatives */
            memset(lxN, 0, sizeof(lxN));
            const double *zN = Z + N * NZ;
            lxN[IEY] = 2 * zN[IEY]; lxN[IEPSI] = 2 * zN[IEPSI];
            lxxN_ss = lxxN_sv = lxxN_vv = 0;
            if (P->n_layers == 0) lxN[IS] -= 1.0;
            else {
                double V, g[2], H[3];
                mlp_value(P, zN[IS], zN[IV], pr->ctx, &V, g, H);
                lxN[IS] -= g[0]; lxN[IV] -= g[1];
                lxxN_ss = -H[0]; lxxN_sv = -H[1]; lxxN_vv = -H[2];
            }
            /* sweep 2: adjoint + KKT residuals */
            double *lam = w->lam;
            stat = 0; rp = 0; s_max = 0; sy_min = INFINITY; sy_max = 0;
            for (int i = 0; i < NA; i++) lam[N * NA + i] = lxN[i];
            for (int k = N; k >= 0; k--) {
                double gw[NW];
                memset(gw, 0, sizeof(gw));
                int o = w->off[k], n = w->off[k + 1] - o;
                for (int i = 0; i < n; i++) {
                    const row_t *r = &w->rows[o + i];
                    double s = S[o + i], y = Y[o + i];
                    gw[r->i0] += r->g0 * s;
                    if (r->i1 >= 0) gw[r->i1] += r->g1 * s;
                    if (r->i2 >= 0) gw[r->i2] += r->g2 * s;
                    double a = fabs(r->c + y);
                    if (a > rp) rp = a;
                    if (s > s_max) s_max = s;
                    double sy = s * y;
                    if (sy < sy_min) sy_min = sy;
                    if (sy > sy_max) sy_max = sy;
                }
                if (k == N) {
                    for (int i = 0; i < NA; i++) lam[N * NA + i] += gw[i];
                    continue;
                }
                const double *A = w->A + k * NZ * NZ, *B = w->B + k * NZ * 2;
                const double *ln = lam + (k + 1) * NA;
                double *lk = lam + k * NA;
                for (int j = 0; j < 2; j++) {
                    double gu = 2 * P->w_u * U[2 * k + j] + ln[NZ + j] + gw[NA + j];
                    for (int a = 0; a < NZ; a++) gu += B[a * 2 + j] * ln[a];
                    if (fabs(gu) > stat) stat = fabs(gu);
                }
                for (int i = 0; i < NA; i++) lk[i] = gw[i];
                lk[IEY] += 2 * Z[k * NZ + IEY]; lk[IEPSI] += 2 * Z[k * NZ + IEPSI];
                for (int i = 0; i < NZ; i++) {
                    double acc = 0;
                    for (int a = 0; a < NZ; a++) acc += A[a * NZ + i] * ln[a];
                    lk[i] += acc;
                }
            }
            need_jac = 0;
        }
        if (stat <= P->tol * fmax(1.0, s_max) && rp <= P->tol_rp && sy_max <= P->tol_comp) { status = 0; break; }
        const int acceptable = P->acc_tol > 0 && stat <= P->acc_tol * fmax(1.0, s_max) && rp <= P->acc_rp && sy_max <= P->acc_comp;
        if (it == P->max_iter) { if (acceptable) status = 6; break; }
        if (it >= P->stall_iter && rp > P->stall_rp) { status = 5; break; }
        /* barrier update */
        while (mu > P->mu_floor &&
               fmax(fmax(stat, rp), fmax(fabs(sy_max - mu), fabs(sy_min - mu))) <= P->kappa_eps * mu)
            mu = fmax(P->mu_floor, fmin(P->kappa_mu * mu, pow(mu, P->theta_mu)));
        double phi, theta = 0;
        {
            double lg = 0;
            for (int i = 0; i < w->M; i++) { lg += log(Y[i]); theta += fabs(w->rows[i].c + Y[i]); }
            phi = cost_fn(P, pr, Z, U) - mu * lg;
        }
        /* sweep 3: Riccati */
        int ok;
        for (;;) {
            double reg_hint = 0.0;
            ok = 1;
            double Vx[NA], Vxx[NA][NA];
            {
                memset(Vxx, 0, sizeof(Vxx));
                for (int i = 0; i < NA; i++) Vx[i] = lxN[i];
                Vxx[IEY][IEY] = 2.0; Vxx[IEPSI][IEPSI] = 2.0;
                Vxx[IS][IS] += lxxN_ss; Vxx[IS][IV] += lxxN_sv; Vxx[IV][IS] += lxxN_sv; Vxx[IV][IV] += lxxN_vv;
                int o = w->off[N], n = w->off[N + 1] - o;
                for (int i = 0; i < n; i++) {
                    const row_t *r = &w->rows[o + i];
                    double s = S[o + i], y = Y[o + i];
                    double rhat = s * r->c + mu, sig = s / y, gr = s + rhat / y;
                    Vx[r->i0] += r->g0 * gr;
                    Vxx[r->i0][r->i0] += sig * r->g0 * r->g0;
                    if (r->i1 >= 0) {
                        Vx[r->i1] += r->g1 * gr;
                        Vxx[r->i1][r->i1] += sig * r->g1 * r->g1;
                        Vxx[r->i0][r->i1] += sig * r->g0 * r->g1;
                        Vxx[r->i1][r->i0] += sig * r->g0 * r->g1;
                    }
                    if (r->i2 >= 0) {
                        Vx[r->i2] += r->g2 * gr;
                        Vxx[r->i2][r->i2] += sig * r->g2 * r->g2;
                        Vxx[r->i0][r->i2] += sig * r->g0 * r->g2; Vxx[r->i2][r->i0] += sig * r->g0 * r->g2;
                        Vxx[r->i1][r->i2] += sig * r->g1 * r->g2; Vxx[r->i2][r->i1] += sig * r->g1 * r->g2;
                    }
                    if (r->has_h) {
                        Vxx[IX][IX] += s * r->hxx; Vxx[IX][IY] += s * r->hxy;
                        Vxx[IY][IX] += s * r->hxy; Vxx[IY][IY] += s * r->hyy;
                    }
                }
            }
            for (int k = N - 1; k >= 0; k--) {
                const double *A = w->A + k * NZ * NZ, *B = w->B + k * NZ * 2;
                double g[NW], H[NW][NW];
                memset(g, 0, sizeof(g));
                memset(H, 0, sizeof(H));
                /* cost */
                g[IEY] = 2 * Z[k * NZ + IEY]; g[IEPSI] = 2 * Z[k * NZ + IEPSI];
                g[IUA] = 2 * P->w_u * U[2 * k]; g[IUD] = 2 * P->w_u * U[2 * k + 1];
                H[IEY][IEY] = 2.0; H[IEPSI][IEPSI] = 2.0; H[IUA][IUA] = 2 * P->w_u; H[IUD][IUD] = 2 * P->w_u;
                /* next-stage value through zeta+ = [A z + B u; u] : Jacobian F[9][11] */
                double F[NA][NW];
                memset(F, 0, sizeof(F));
                for (int i = 0; i < NZ; i++) {
                    for (int j = 0; j < NZ; j++) F[i][j] = A[i * NZ + j];
                    F[i][IUA] = B[i * 2]; F[i][IUD] = B[i * 2 + 1];
                }
                F[IPA][IUA] = 1.0; F[IPD][IUD] = 1.0;
                double VF[NA][NW];
                for (int i = 0; i < NA; i++)
                    for (int j = 0; j < NW; j++) {
                        double acc = 0;
                        for (int a = 0; a < NA; a++) acc += Vxx[i][a] * F[a][j];
                        VF[i][j] = acc;
                    }
                for (int i = 0; i < NW; i++) {
                    double acc = 0;
                    for (int a = 0; a < NA; a++) acc += F[a][i] * Vx[a];
                    g[i] += acc;
                    for (int j = 0; j < NW; j++) {
                        double h = 0;
                        for (int a = 0; a < NA; a++) h += F[a][i] * VF[a][j];
                        H[i][j] += h;
                    }
                }
                if (P->second_order) add_dyn_hessian(P, Z + k * NZ, U + 2 * k, pr->curv, w->lam + (k + 1) * NA, H);
                /* rows */
                int o = w->off[k], n = w->off[k + 1] - o;
                for (int i = 0; i < n; i++) {
                    const row_t *r = &w->rows[o + i];
                    double s = S[o + i], y = Y[o + i];
                    double rhat = s * r->c + mu, sig = s / y, gr = s + rhat / y;
                    g[r->i0] += r->g0 * gr;
                    H[r->i0][r->i0] += sig * r->g0 * r->g0;
                    if (r->i1 >= 0) {
                        g[r->i1] += r->g1 * gr;
                        H[r->i1][r->i1] += sig * r->g1 * r->g1;
                        H[r->i0][r->i1] += sig * r->g0 * r->g1;
                        H[r->i1][r->i0] += sig * r->g0 * r->g1;
                    }
                    if (r->i2 >= 0) {
                        g[r->i2] += r->g2 * gr;
                        H[r->i2][r->i2] += sig * r->g2 * r->g2;
                        H[r->i0][r->i2] += sig * r->g0 * r->g2; H[r->i2][r->i0] += sig * r->g0 * r->g2;
                        H[r->i1][r->i2] += sig * r->g1 * r->g2; H[r->i2][r->i1] += sig * r->g1 * r->g2;
                    }
                    if (r->has_h) {
                        H[IX][IX] += s * r->hxx; H[IX][IY] += s * r->hxy;
                        H[IY][IX] += s * r->hxy; H[IY][IY] += s * r->hyy;
                    }
                }
                /* solve for the control */
                double q00 = H[IUA][IUA] + reg, q11 = H[IUD][IUD] + reg, q01 = H[IUA][IUD];
                double det = q00 * q11 - q01 * q01;
                if (!(q00 > 0 && det > 1e-12 * q00 * q11)) {
                    /* the shift that makes this stage's Quu positive definite */
                    double a_ = H[IUA][IUA], b_ = H[IUD][IUD];
                    double lmin = 0.5 * (a_ + b_) - sqrt(0.25 * (a_ - b_) * (a_ - b_) + q01 * q01);
                    reg_hint = lmin < 0 ? -lmin * P->reg_jump : 0.0;
                    ok = 0;
                    break;
                }
                double i00 = q11 / det, i11 = q00 / det, i01 = -q01 / det;
                double *ku = w->ku + 2 * k, *Ku = w->Ku + 2 * NA * k;
                ku[0] = -(i00 * g[IUA] + i01 * g[IUD]);
                ku[1] = -(i01 * g[IUA] + i11 * g[IUD]);
                for (int j = 0; j < NA; j++) {
                    Ku[j] = -(i00 * H[IUA][j] + i01 * H[IUD][j]);
                    Ku[NA + j] = -(i01 * H[IUA][j] + i11 * H[IUD][j]);
                }
                /* value update (unregularised Quu) */
                double Quuk[2] = { H[IUA][IUA] * ku[0] + H[IUA][IUD] * ku[1],
                                   H[IUD][IUA] * ku[0] + H[IUD][IUD] * ku[1] };
                double Vxn[NA], Vxxn[NA][NA];
                for (int i = 0; i < NA; i++) {
                    Vxn[i] = g[i] + Ku[i] * (g[IUA] + Quuk[0]) + Ku[NA + i] * (g[IUD] + Quuk[1])
                             + H[IUA][i] * ku[0] + H[IUD][i] * ku[1];
                }
                for (int i = 0; i < NA; i++)
                    for (int j = 0; j < NA; j++) {
                        double QK0 = H[IUA][IUA] * Ku[j] + H[IUA][IUD] * Ku[NA + j];
                        double QK1 = H[IUD][IUA] * Ku[j] + H[IUD][IUD] * Ku[NA + j];
                        Vxxn[i][j] = H[i][j] + Ku[i] * QK0 + Ku[NA + i] * QK1
                                     + Ku[i] * H[IUA][j] + Ku[NA + i] * H[IUD][j]
                                     + H[IUA][i] * Ku[j] + H[IUD][i] * Ku[NA + j];
                    }
                for (int i = 0; i < NA; i++) {
                    Vx[i] = Vxn[i];
                    for (int j = 0; j < NA; j++) Vxx[i][j] = 0.5 * (Vxxn[i][j] + Vxxn[j][i]);
                }
            }
            if (ok) break;
            reg = fmax(fmax(reg * P->reg_up, P->reg_min), reg_hint);
            if (reg > P->reg_max) break;
        }
        if (reg > P->reg_max) { status = acceptable ? 6 : 3; break; }
        /* forward passes */
        double tau = fmax(P->tau_min, 1.0 - mu);
        int accepted = 0;
        double alpha = 1.0;
        if (P->predict_alpha) {
            /* largest step that keeps every slack inside the fraction-to-boundary rule on the
             * LINEARISED closed-loop model (d zeta+ = F [d zeta; d u], d u = ku + Ku d zeta) */
            double dz[NA];
            memset(dz, 0, sizeof(dz));
            for (int k = 0; k <= N; k++) {
                double dw[NW];
                for (int i = 0; i < NA; i++) dw[i] = dz[i];
                dw[IUA] = dw[IUD] = 0;
                if (k < N) {
                    const double *ku = w->ku + 2 * k, *Ku = w->Ku + 2 * NA * k;
                    double d0 = ku[0], d1 = ku[1];
                    for (int j = 0; j < NA; j++) { d0 += Ku[j] * dz[j]; d1 += Ku[NA + j] * dz[j]; }
                    dw[IUA] = d0; dw[IUD] = d1;
                }
                int o = w->off[k], n = w->off[k + 1] - o;
                for (int i = 0; i < n; i++) {
                    const row_t *r = &w->rows[o + i];
                    double y = Y[o + i];
                    double dc = r->g0 * dw[r->i0] + (r->i1 >= 0 ? r->g1 * dw[r->i1] : 0.0) + (r->i2 >= 0 ? r->g2 * dw[r->i2] : 0.0);
                    double dy = -(r->c + y) - dc;
                    if (dy < 0 && -dy * alpha > tau * y) alpha = tau * y / (-dy);
                }
                if (k == N) break;
                const double *A = w->A + k * NZ * NZ, *B = w->B + k * NZ * 2;
                for (int i = 0; i < NZ; i++) {
                    double acc = B[i * 2] * dw[IUA] + B[i * 2 + 1] * dw[IUD];
                    for (int j = 0; j < NZ; j++) acc += A[i * NZ + j] * dw[j];
                    dz[i] = acc;
                }
                dz[IPA] = dw[IUA]; dz[IPD] = dw[IUD];
            }
            alpha *= P->alpha_safety;
        }
        double *Zn = w->Zn, *Un = w->Un, *Yn = w->Yn, *Sn = w->Sn;
        for (int ls = 0; ls < P->n_alpha; ls++, alpha *= 0.5) {
            int fail = 0;
            trials++;
            double thetan = 0, lg = 0;
            memcpy(Zn, Z, sizeof(double) * NZ);
            double upn[2] = { pr->u_prev[0], pr->u_prev[1] };
            for (int k = 0; k <= N && !fail; k++) {
                const double *up = k == 0 ? pr->u_prev : U + 2 * (k - 1);
                double dw[NW];
                for (int i = 0; i < NZ; i++) dw[i] = Zn[k * NZ + i] - Z[k * NZ + i];
                dw[IPA] = upn[0] - up[0]; dw[IPD] = upn[1] - up[1];
                dw[IUA] = dw[IUD] = 0;
                if (k < N) {
                    const double *ku = w->ku + 2 * k, *Ku = w->Ku + 2 * NA * k;
                    double d0 = alpha * ku[0], d1 = alpha * ku[1];
                    for (int j = 0; j < NA; j++) { d0 += Ku[j] * dw[j]; d1 += Ku[NA + j] * dw[j]; }
                    dw[IUA] = d0; dw[IUD] = d1;
                }
                int o = w->off[k], n = w->off[k + 1] - o;
                for (int i = 0; i < n; i++) {
                    const row_t *r = &w->rows[o + i];
                    double s = S[o + i], y = Y[o + i];
                    double dc = r->g0 * dw[r->i0] + (r->i1 >= 0 ? r->g1 * dw[r->i1] : 0.0) + (r->i2 >= 0 ? r->g2 * dw[r->i2] : 0.0);
                    double yn = y - alpha * (r->c + y) - dc;
                    double sn = s + (alpha * (s * r->c + mu) + s * dc) / y;
                    if (yn < (1 - tau) * y) { fail = 1; break; }        /* fraction to the boundary */
                    if (sn < (1 - tau) * s) sn = (1 - tau) * s;        /* multiplier safeguard */
                    Yn[o + i] = yn; Sn[o + i] = sn;
                }
                if (fail || k == N) break;
                Un[2 * k] = U[2 * k] + dw[IUA];
                Un[2 * k + 1] = U[2 * k + 1] + dw[IUD];
                rk4_step(P, Zn + k * NZ, Un + 2 * k, pr->curv, Zn + (k + 1) * NZ, NULL, NULL);
                for (int i = 0; i < NZ; i++) if (!isfinite(Zn[(k + 1) * NZ + i])) fail = 1;
                upn[0] = Un[2 * k]; upn[1] = Un[2 * k + 1];
            }
            if (fail) continue;
            for (int k = 0; k <= N; k++) {
                const double *upk = k == 0 ? pr->u_prev : Un + 2 * (k - 1);
                int o = w->off[k];
                int n = rows_eval(P, pr, k, Zn + k * NZ, upk, k < N ? Un + 2 * k : NULL, w->rtmp);
                for (int i = 0; i < n; i++) {
                    w->Cn[o + i] = w->rtmp[i].c;
                    thetan += fabs(w->rtmp[i].c + Yn[o + i]);
                    lg += log(Yn[o + i]);
                }
            }
            double phin = cost_fn(P, pr, Zn, Un) - mu * lg;
            if (isfinite(phin) &&
                (phin < phi - P->eps_phi * fabs(phi) || thetan < theta * (1 - P->gamma_theta) ||
                 (thetan <= P->theta_small && phin <= phi + P->eps_phi * fmax(1.0, fabs(phi))))) {
                accepted = 1;
                break;
            }
        }
        if (accepted) {
            memcpy(Z, Zn, sizeof(double) * NZ * (N + 1));
            memcpy(U, Un, sizeof(double) * 2 * N);
            memcpy(Y, Yn, sizeof(double) * w->M);
            memcpy(S, Sn, sizeof(double) * w->M);
            reg = reg > P->reg_min ? reg / P->reg_down : 0.0;
            need_jac = 1;
            ls_fail = 0;
        } else {
            reg = fmax(reg * P->reg_up, P->reg_min);
            ls_fail++;
            if (reg > P->reg_max || ls_fail >= P->max_ls_fail) { status = acceptable ? 6 : 4; break; }
        }
        if (trials >= P->max_trials) { status = (!accepted && acceptable) ? 6 : 1; break; }
    }
    memcpy(Zout, Z, sizeof(double) * NZ * (N + 1));
    memcpy(Uout, U, sizeof(double) * 2 * N);
    *cost_out = cost_fn(P, pr, Z, U);
    *viol_out = max_violation(P, pr, Z, U);
    return status;
}

/* ------------------------------------------------------------------ C API ------------- */
void igt_oracle_default_options(igt_oracle_params *P)
{
    P->tol = 1e-6; P->tol_rp = 1e-8; P->tol_comp = 1e-7; P->mu0 = 0.3; P->mu_floor = 1e-8;
    P->kappa_eps = 10.0; P->kappa_mu = 0.2; P->theta_mu = 1.5; P->y_init_min = 0.3;
    P->tau_min = 0.99; P->reg_min = 1e-4; P->reg_up = 10.0; P->reg_down = 10.0; P->reg_max = 1e10;
    P->eps_phi = 1e-12; P->gamma_theta = 1e-6; P->theta_small = 1e-10;
    P->max_iter = 300; P->n_alpha = 6; P->second_order = 1;
    P->max_ls_fail = 1000; P->max_trials = 1000000; P->predict_alpha = 1; P->alpha_safety = 0.99;
    P->reg_jump = 1.1;
    P->stall_iter = 16; P->stall_rp = 1e-2;
    P->acc_tol = 1e-3; P->acc_rp = 1e-6; P->acc_comp = 1e-4;
}

size_t igt_oracle_params_size(void) { return sizeof(igt_oracle_params); }

/* Z[B][N+1][7], A[B][N][7][7], Bm[B][N][7][2]; model 0 = Frenet RK4, 1 = Cartesian Euler
 * (then Z[B][N+1][4], A[B][N][4][4], Bm[B][N][4][2], curv unused) */
int igt_oracle_rollout(const igt_oracle_params *P, int nb, const double *z0, const double *U,
                       const double *curv, double *Z, double *A, double *Bm, int model)
{
    int N = P->N;
    int nz = model == 0 ? NZ : 4;
    for (int b = 0; b < nb; b++) {
        double *Zb = Z + (size_t)b * (N + 1) * nz;
        memcpy(Zb, z0 + (size_t)b * nz, sizeof(double) * nz);
        for (int k = 0; k < N; k++) {
            const double *u = U + ((size_t)b * N + k) * 2;
            if (model == 0)
                rk4_step(P, Zb + k * NZ, u, curv + 3 * (size_t)b, Zb + (k + 1) * NZ,
                         A ? (double (*)[NZ])(A + ((size_t)b * N + k) * NZ * NZ) : NULL,
                         A ? (double (*)[2])(Bm + ((size_t)b * N + k) * NZ * 2) : NULL);
            else
                euler_step(P, Zb + k * 4, u, Zb + (k + 1) * 4,
                           A ? A + ((size_t)b * N + k) * 16 : NULL, A ? Bm + ((size_t)b * N + k) * 8 : NULL);
        }
    }
    return 0;
}

/* cost[B], viol[B] of given trajectories (Z must be the rollout of U) */
int igt_oracle_eval(const igt_oracle_params *P, int nb, const double *x0, const double *u_prev,
                    const double *curv, const double *obs, const double *ctx, const double *Z,
                    const double *U, double *cost, double *viol)
{
    int N = P->N;
    for (int b = 0; b < nb; b++) {
        prob_t pr = { x0 + 7 * (size_t)b, u_prev + 2 * (size_t)b, curv + 3 * (size_t)b,
                      obs + (size_t)b * (N + 1) * 2, ctx ? ctx + 4 * (size_t)b : NULL, NULL };
        cost[b] = cost_fn(P, &pr, Z + (size_t)b * (N + 1) * NZ, U + (size_t)b * N * 2);
        viol[b] = max_violation(P, &pr, Z + (size_t)b * (N + 1) * NZ, U + (size_t)b * N * 2);
    }
    return 0;
}

int igt_oracle_initial_guess(const igt_oracle_params *P, int nb, const double *x0, const double *u_prev,
                             const double *curv, const double *obs, double *U)
{
    int N = P->N;
    work_t *w = work_alloc(P);
    for (int b = 0; b < nb; b++) {
        prob_t pr = { x0 + 7 * (size_t)b, u_prev + 2 * (size_t)b, curv + 3 * (size_t)b,
                      obs + (size_t)b * (N + 1) * 2, NULL, NULL };
        double best = INFINITY;
        for (int g = 0; g < 5; g++) {
            tracking_rollout(P, &pr, GUESS_SPEEDS[g], w->Ug, w->Zg);
            double m = guess_merit(P, &pr, w->Zg, w->Ug);
            if (m < best) { best = m; memcpy(U + (size_t)b * N * 2, w->Ug, sizeof(double) * 2 * N); }
        }
    }
    work_free(w);
    return 0;
}

typedef struct {
    const igt_oracle_params *P;
    int nb;
    const double *x0, *u_prev, *curv, *obs, *ctx, *u_init;
    double *Z, *U, *cost, *viol;
    int *status, *iters;
    int next;                /* shared work counter (chunks of 4 problems) */
    const double *obs_psi;   /* [B][N+1] or NULL */
} batch_t;

static void *batch_worker(void *arg)
{
    batch_t *bt = (batch_t *)arg;
    const igt_oracle_params *P = bt->P;
    int N = P->N;
    work_t *w = work_alloc(P);
    for (;;) {
        int lo = __atomic_fetch_add(&bt->next, 4, __ATOMIC_RELAXED);
        if (lo >= bt->nb) break;
        int hi = lo + 4 < bt->nb ? lo + 4 : bt->nb;
        for (int b = lo; b < hi; b++) {
            prob_t pr = { bt->x0 + 7 * (size_t)b, bt->u_prev + 2 * (size_t)b, bt->curv + 3 * (size_t)b,
                          bt->obs + (size_t)b * (N + 1) * 2, bt->ctx ? bt->ctx + 4 * (size_t)b : NULL,
                          bt->u_init ? bt->u_init + (size_t)b * N * 2 : NULL,
                          bt->obs_psi ? bt->obs_psi + (size_t)b * (N + 1) : NULL };
            bt->status[b] = solve_one(P, &pr, w, bt->Z + (size_t)b * (N + 1) * NZ, bt->U + (size_t)b * N * 2,
                                      bt->cost + b, bt->viol + b, bt->iters + b);
        }
    }
    work_free(w);
    return NULL;
}

/* n_threads <= 0: use every online core.  obs_psi[B][N+1] != NULL selects the OBCA collision rows. */
int igt_oracle_solve_batch_obca(const igt_oracle_params *P, int nb, const double *x0, const double *u_prev,
                                const double *curv, const double *obs, const double *obs_psi, const double *ctx,
                                const double *u_init, double *Z, double *U, double *cost, double *viol,
                                int *status, int *iters, int n_threads)
{
    batch_t bt = { P, nb, x0, u_prev, curv, obs, ctx, u_init, Z, U, cost, viol, status, iters, 0, obs_psi };
    if (n_threads <= 0) n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (n_threads > 256) n_threads = 256;
    if (n_threads > (nb + 3) / 4) n_threads = (nb + 3) / 4;
    if (n_threads <= 1) { batch_worker(&bt); return 0; }
    pthread_t th[256];
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, batch_worker, &bt);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    return 0;
}

int igt_oracle_solve_batch(const igt_oracle_params *P, int nb, const double *x0, const double *u_prev,
                           const double *curv, const double *obs, const double *ctx,
                           const double *u_init, double *Z, double *U, double *cost, double *viol,
                           int *status, int *iters, int n_threads)
{
    return igt_oracle_solve_batch_obca(P, nb, x0, u_prev, curv, obs, NULL, ctx, u_init, Z, U, cost, viol, status, iters, n_threads);
}

/* OBCA mode: max row value (reference units) of given trajectories */
int igt_oracle_viol_obca(const igt_oracle_params *P, int nb, const double *x0, const double *u_prev, const double *curv,
                         const double *obs, const double *obs_psi, const double *Z, const double *U, double *viol)
{
    int N = P->N;
    for (int b = 0; b < nb; b++) {
        prob_t pr = { x0 + 7 * (size_t)b, u_prev + 2 * (size_t)b, curv + 3 * (size_t)b, obs + (size_t)b * (N + 1) * 2, NULL, NULL,
                      obs_psi + (size_t)b * (N + 1) };
        viol[b] = max_violation(P, &pr, Z + (size_t)b * (N + 1) * NZ, U + (size_t)b * N * 2);
    }
    return 0;
}

/* signed rectangle distance and gradient for n pose pairs (tests) */
int igt_oracle_rect_sdist(int n, const double *ego, const double *obs, double *d, double *g)
{
    for (int i = 0; i < n; i++) d[i] = rect_signed_distance(ego + 3 * i, obs + 3 * i, g + 3 * i);
    return 0;
}
