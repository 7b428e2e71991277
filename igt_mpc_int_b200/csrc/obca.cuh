// obca.cuh -- the reference's OBCA collision rows (mpc.py:170-175 dual variables, :211-221 rows,
// common/utils.py:412-442 rotation_translation) in dual-eliminated form.
//
// Per stage k >= 1 the reference adds lambda, mu in R^4 with
//     -g' mu + (A p - b)' lambda >= d_min + 1e-6,  G' mu + R(psi)' A' lambda = 0,  |A' lambda| <= 1,  lambda, mu >= 0
// ({y: A y <= b} the obstacle's rectangle at its forecast pose, {y: G y <= g} the ego rectangle in its own frame).
// The duals appear in no other row and in no cost term, so they can be maximised out stage by stage: the largest
// left-hand side of the first row over the other three is the Euclidean distance between the two rectangles (strong
// duality of the polytope distance problem).  The five rows of a stage are therefore ONE stage-local row
//     d_min + 1e-6 - sd(ego rectangle(x, y, psi), obstacle rectangle_k) <= 0
// with the same KKT points; sd continues the distance into the obstacle as minus the penetration depth (the plain
// distance is flat there and an iterate that jumped inside could not get out).  Gradient in (x, y, psi) from the
// closest pair of features; no Hessian (Gauss-Newton row).  Pinned in tests/test_oracle_obca.py: the reference's
// dual problem solved with SLSQP equals this closed form.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define IGT_OBCA_HD __host__ __device__ inline
#else
#define IGT_OBCA_HD inline
#endif

namespace igt {

constexpr double OBCA_MARGIN = 1e-6;          // mpc.py:216
constexpr double VEH_HL = 4.47 / 2, VEH_HW = 2.0 / 2;   // mpc.py:48-51: half length, half width

// distance from point q to the rectangle at (cx, cy) with heading (c, s): returns the distance, the closest
// point in the rectangle's frame (cl) and in the world (cw), and whether q is inside
template <typename T>
IGT_OBCA_HD T obca_point_rect(T qx, T qy, T cx, T cy, T c, T s, T *cl, T *cw, bool *inside)
{
    const T dx = c * (qx - cx) + s * (qy - cy), dy = -s * (qx - cx) + c * (qy - cy);
    const T hl = T(VEH_HL), hw = T(VEH_HW);
    cl[0] = fmin(fmax(dx, -hl), hl); cl[1] = fmin(fmax(dy, -hw), hw);
    *inside = fabs(dx) <= hl && fabs(dy) <= hw;
    cw[0] = cx + c * cl[0] - s * cl[1]; cw[1] = cy + s * cl[0] + c * cl[1];
    const T ex = dx - cl[0], ey = dy - cl[1];
    return sqrt(ex * ex + ey * ey);
}

// intersecting rectangles: minus the penetration depth (separating-axis test over the four face normals) and its
// gradient with respect to the ego pose
template <typename T>
IGT_OBCA_HD T obca_penetration(T ex, T ey, T ce, T se, T ox, T oy, T co, T so, T *g)
{
    const T hl = T(VEH_HL), hw = T(VEH_HW);
    const T dvx = ex - ox, dvy = ey - oy;
    auto sg = [](T v) { return v >= T(0) ? T(1) : T(-1); };
    const T uxe[2] = { ce, se }, uye[2] = { -se, ce }, uxo[2] = { co, so }, uyo[2] = { -so, co };
    T best = T(1e300), bgx = T(0), bgy = T(0), bgp = T(0);
    for (int f = 0; f < 2; f++) {                                 // ego face normals (they turn with psi)
        const T *a = f == 0 ? uxe : uye;
        const T da[2] = { f == 0 ? uye[0] : -uxe[0], f == 0 ? uye[1] : -uxe[1] };
        const T he = f == 0 ? hl : hw;
        const T axo = a[0] * uxo[0] + a[1] * uxo[1], ayo = a[0] * uyo[0] + a[1] * uyo[1], ad = a[0] * dvx + a[1] * dvy;
        const T ov = he + hl * fabs(axo) + hw * fabs(ayo) - fabs(ad);
        if (ov < best) {
            best = ov;
            bgx = -sg(ad) * a[0]; bgy = -sg(ad) * a[1];
            bgp = hl * sg(axo) * (da[0] * uxo[0] + da[1] * uxo[1]) + hw * sg(ayo) * (da[0] * uyo[0] + da[1] * uyo[1])
                  - sg(ad) * (da[0] * dvx + da[1] * dvy);
        }
    }
    for (int f = 0; f < 2; f++) {                                 // obstacle face normals (fixed)
        const T *a = f == 0 ? uxo : uyo;
        const T ho = f == 0 ? hl : hw;
        const T axe = a[0] * uxe[0] + a[1] * uxe[1], aye = a[0] * uye[0] + a[1] * uye[1], ad = a[0] * dvx + a[1] * dvy;
        const T ov = hl * fabs(axe) + hw * fabs(aye) + ho - fabs(ad);
        if (ov < best) {
            best = ov;
            bgx = -sg(ad) * a[0]; bgy = -sg(ad) * a[1];
            bgp = hl * sg(axe) * aye + hw * sg(aye) * (-axe);
        }
    }
    g[0] = -bgx; g[1] = -bgy; g[2] = -bgp;
    return -best;
}

template <typename T>
IGT_OBCA_HD bool obca_edges_cross(const T (*P)[2], const T (*Q)[2])
{
    for (int i = 0; i < 4; i++) {
        const T a0x = P[i][0], a0y = P[i][1], d1x = P[(i + 1) & 3][0] - a0x, d1y = P[(i + 1) & 3][1] - a0y;
        for (int j = 0; j < 4; j++) {
            const T b0x = Q[j][0], b0y = Q[j][1], d2x = Q[(j + 1) & 3][0] - b0x, d2y = Q[(j + 1) & 3][1] - b0y;
            const T den = d1x * d2y - d1y * d2x;
            if (fabs(den) < T(1e-14)) continue;
            const T rx = b0x - a0x, ry = b0y - a0y;
            const T t = (rx * d2y - ry * d2x) / den, u = (rx * d1y - ry * d1x) / den;
            if (t >= T(0) && t <= T(1) && u >= T(0) && u <= T(1)) return true;
        }
    }
    return false;
}

// signed distance between the ego rectangle at (ex, ey, epsi) and the obstacle rectangle at (ox, oy, opsi) and its
// gradient g[3] with respect to (ex, ey, epsi).  For disjoint convex polygons the minimum is attained between a
// vertex of one and the boundary of the other: the eight vertex-to-rectangle distances cover it.
template <typename T>
IGT_OBCA_HD T obca_rect_sdist(T ex, T ey, T epsi, T ox, T oy, T opsi, T *g)
{
    const T hl = T(VEH_HL), hw = T(VEH_HW);
    const T ce = cos(epsi), se = sin(epsi), co = cos(opsi), so = sin(opsi);
    const T loc[4][2] = { { hl, hw }, { -hl, hw }, { -hl, -hw }, { hl, -hw } };
    T Pe[4][2], Po[4][2];
    for (int i = 0; i < 4; i++) {
        Pe[i][0] = ex + ce * loc[i][0] - se * loc[i][1]; Pe[i][1] = ey + se * loc[i][0] + ce * loc[i][1];
        Po[i][0] = ox + co * loc[i][0] - so * loc[i][1]; Po[i][1] = oy + so * loc[i][0] + co * loc[i][1];
    }
    T best = T(1e300), ax = T(0), ay = T(0), bx = T(0), by = T(0), alx = T(0), aly = T(0);   // a: ego's point, b: obstacle's
    for (int i = 0; i < 4; i++) {                                 // ego vertex -> obstacle rectangle
        T cl[2], cw[2]; bool in;
        const T d = obca_point_rect(Pe[i][0], Pe[i][1], ox, oy, co, so, cl, cw, &in);
        if (in) return obca_penetration(ex, ey, ce, se, ox, oy, co, so, g);
        if (d < best) { best = d; ax = Pe[i][0]; ay = Pe[i][1]; bx = cw[0]; by = cw[1]; alx = loc[i][0]; aly = loc[i][1]; }
    }
    for (int j = 0; j < 4; j++) {                                 // obstacle vertex -> ego rectangle
        T cl[2], cw[2]; bool in;
        const T d = obca_point_rect(Po[j][0], Po[j][1], ex, ey, ce, se, cl, cw, &in);
        if (in) return obca_penetration(ex, ey, ce, se, ox, oy, co, so, g);
        if (d < best) { best = d; ax = cw[0]; ay = cw[1]; bx = Po[j][0]; by = Po[j][1]; alx = cl[0]; aly = cl[1]; }
    }
    if (obca_edges_cross(Pe, Po)) return obca_penetration(ex, ey, ce, se, ox, oy, co, so, g);   // crossing edges, no vertex inside
    const T nx = (ax - bx) / best, ny = (ay - by) / best;         // from the obstacle's closest point to the ego's
    g[0] = nx; g[1] = ny;
    g[2] = nx * (-se * alx - ce * aly) + ny * (ce * alx - se * aly);   // n . d(R a_local)/dpsi
    return best;
}

}  // namespace igt
