/*
 * oracle/geom_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle, float64 geometry).
 *
 * Plain-C restatement of the third-party arithmetic the reference's per-action
 * step calls (DOLFIN / shapely are not vendored under /root/reference and cannot
 * be installed offline).  PARITY UNPINNED: the reference ships no tests, golden
 * vectors or recorded trajectories for this path, so these
 * functions pin the *restatement* in SURVEY.md Appendix A, not a run of FEniCS.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (meshdqn_b200/) never does.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/Makefile) so every
 * double operation is a single IEEE-754 op, as in the CUDA side (-fmad=false).
 *
 * Reference call sites restated here:
 *   orc_smooth            flow_solver.py:67,237      mesh.smooth(50)      (App. A.2)
 *   orc_facet_tags        flow_solver.py:9-30,194-226 SubDomain.mark      (App. A.4)
 *   orc_removable         flow_solver.py:75-78,247-250                    (App. A.3)
 *   orc_polygon_distance  Env2DAirfoil.py:232,240-241 Polygon.distance    (App. A.5)
 *   orc_locate            Env2DAirfoil.py:562,568     Function.interpolate (App. A.7)
 *   orc_eval_p2 / _p1     Env2DAirfoil.py:562,568,516-522                 (App. A.7/A.8)
 *   orc_drag_lift         probes.py:23-31,43-50       assemble(ds(1))     (App. A.9)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define DOLFIN_EPS 3.0e-16

/* ---- A.2: Gauss-Seidel Laplacian smoothing, boundary fixed -------------------- */
void orc_smooth(double *x, int nv, const int *nbr_ptr, const int *nbr_idx,
                const int *vc_ptr, const int *vc_idx, const int *cells,
                const unsigned char *on_boundary, int iters)
{
    for (int it = 0; it < iters; ++it) {
        for (int v = 0; v < nv; ++v) {
            if (on_boundary[v]) continue;
            const double px = x[2 * v], py = x[2 * v + 1];
            double sx = 0.0, sy = 0.0;
            int nn = 0;
            for (int k = nbr_ptr[v]; k < nbr_ptr[v + 1]; ++k) {
                const int o = nbr_idx[k];
                sx += x[2 * o];
                sy += x[2 * o + 1];
                nn += 1;
            }
            if (nn == 0) continue;
            sx /= (double)nn;
            sy /= (double)nn;
            double rmin = 0.0;
            for (int k = vc_ptr[v]; k < vc_ptr[v + 1]; ++k) {
                const int *c = cells + 3 * vc_idx[k];
                int a, b; /* the two other vertices, ascending index order */
                if (c[0] == v) { a = c[1]; b = c[2]; }
                else if (c[1] == v) { a = c[0]; b = c[2]; }
                else { a = c[0]; b = c[1]; }
                const double ax = x[2 * a], ay = x[2 * a + 1];
                const double ex = x[2 * b] - ax, ey = x[2 * b + 1] - ay;
                const double len = sqrt(ex * ex + ey * ey);
                const double cr = ex * (py - ay) - ey * (px - ax);
                const double r = fabs(cr) / len;
                if (rmin == 0.0) rmin = r;
                else rmin = (r < rmin) ? r : rmin;
            }
            const double dx = sx - px, dy = sy - py;
            const double r = sqrt(dx * dx + dy * dy);
            if (r < DOLFIN_EPS) continue;
            const double half = 0.5 * rmin;
            const double step = (half < r) ? half : r;
            x[2 * v] = px + step * dx / r;
            x[2 * v + 1] = py + step * dy / r;
        }
    }
}

/* ---- A.4: exterior-facet tags (4 default, 0 walls, 1 airfoil, 2 inflow, 3 outflow) */
static int in_walls(double x, double y) { (void)x; return (y > 0.5 - 2 * DOLFIN_EPS) || (y < -0.5 + 2 * DOLFIN_EPS); }
static int in_airfoil(double x, double y)
{
    return (x < 3.0 - DOLFIN_EPS) && (x > -0.5 + DOLFIN_EPS) && (y < 0.5 - DOLFIN_EPS) && (y > -0.5 + DOLFIN_EPS);
}
static int in_inflow(double x, double y) { (void)y; return x < -0.5 + DOLFIN_EPS; }
static int in_outflow(double x, double y) { (void)y; return x > 3.0 - 2 * DOLFIN_EPS; }

void orc_facet_tags(const double *x, const int *edges, const int *edge_ncells, int ne, int *tags)
{
    for (int e = 0; e < ne; ++e) {
        int tag = 4;
        if (edge_ncells[e] == 1) {
            const int a = edges[2 * e], b = edges[2 * e + 1];
            const double ax = x[2 * a], ay = x[2 * a + 1], bx = x[2 * b], by = x[2 * b + 1];
            const double mx = (ax + bx) / 2.0, my = (ay + by) / 2.0;
            if (in_walls(ax, ay) && in_walls(bx, by) && in_walls(mx, my)) tag = 0;
            if (in_airfoil(ax, ay) && in_airfoil(bx, by) && in_airfoil(mx, my)) tag = 1;
            if (in_inflow(ax, ay) && in_inflow(bx, by) && in_inflow(mx, my)) tag = 2;
            if (in_outflow(ax, ay) && in_outflow(bx, by) && in_outflow(mx, my)) tag = 3;
        }
        tags[e] = tag;
    }
}

/* ---- A.3: numpy `coord not in bmesh.coordinates()` == not (B == coord).any() --- */
void orc_removable(const double *x, int nv, const int *bverts, int nb, unsigned char *removable)
{
    for (int v = 0; v < nv; ++v) {
        int hit = 0;
        for (int k = 0; k < nb && !hit; ++k) {
            const int b = bverts[k];
            if (x[2 * b] == x[2 * v] || x[2 * b + 1] == x[2 * v + 1]) hit = 1;
        }
        removable[v] = (unsigned char)(!hit);
    }
}

/* ---- A.5: shapely Polygon.distance(Point) -------------------------------------- */
static double pt_seg_dist(double px, double py, double ax, double ay, double bx, double by)
{
    const double dx = bx - ax, dy = by - ay;
    if (dx == 0.0 && dy == 0.0) {
        const double ux = px - ax, uy = py - ay;
        return sqrt(ux * ux + uy * uy);
    }
    const double len2 = dx * dx + dy * dy;
    const double r = ((px - ax) * dx + (py - ay) * dy) / len2;
    if (r <= 0.0) {
        const double ux = px - ax, uy = py - ay;
        return sqrt(ux * ux + uy * uy);
    }
    if (r >= 1.0) {
        const double ux = px - bx, uy = py - by;
        return sqrt(ux * ux + uy * uy);
    }
    const double s = ((ay - py) * dx - (ax - px) * dy) / len2;
    return fabs(s) * sqrt(len2);
}

void orc_polygon_distance(const double *pts, int np, const double *ring, int nr, double *out)
{
    for (int i = 0; i < np; ++i) {
        const double px = pts[2 * i], py = pts[2 * i + 1];
        int inside = 0;
        double best = INFINITY;
        for (int k = 0; k < nr; ++k) {
            const int k2 = (k + 1 == nr) ? 0 : k + 1;
            const double ax = ring[2 * k], ay = ring[2 * k + 1];
            const double bx = ring[2 * k2], by = ring[2 * k2 + 1];
            if ((ay > py) != (by > py)) {
                const double xi = ax + (py - ay) * (bx - ax) / (by - ay);
                if (px < xi) inside = !inside;
            }
            const double d = pt_seg_dist(px, py, ax, ay, bx, by);
            if (d < best) best = d;
        }
        out[i] = inside ? 0.0 : best;
    }
}

/* ---- A.7: point location (brute force == "lowest containing cell index") ------- */
static void bary(const double *x, const int *c, double px, double py, double *l0, double *l1, double *l2)
{
    const double x0 = x[2 * c[0]], y0 = x[2 * c[0] + 1];
    const double x1 = x[2 * c[1]], y1 = x[2 * c[1] + 1];
    const double x2 = x[2 * c[2]], y2 = x[2 * c[2] + 1];
    const double d1x = x1 - x0, d1y = y1 - y0, d2x = x2 - x0, d2y = y2 - y0;
    const double det = d1x * d2y - d2x * d1y;
    const double qx = px - x0, qy = py - y0;
    *l1 = (qx * d2y - d2x * qy) / det;
    *l2 = (d1x * qy - qx * d1y) / det;
    *l0 = 1.0 - *l1 - *l2;
}

static double seg_d2(double px, double py, double ax, double ay, double bx, double by)
{
    const double dx = bx - ax, dy = by - ay;
    const double len2 = dx * dx + dy * dy;
    double t = ((px - ax) * dx + (py - ay) * dy) / len2;
    if (t < 0.0) t = 0.0;
    if (t > 1.0) t = 1.0;
    const double cx = ax + t * dx - px, cy = ay + t * dy - py;
    return cx * cx + cy * cy;
}

static double tri_d2(const double *x, const int *c, double px, double py)
{
    const double x0 = x[2 * c[0]], y0 = x[2 * c[0] + 1];
    const double x1 = x[2 * c[1]], y1 = x[2 * c[1] + 1];
    const double x2 = x[2 * c[2]], y2 = x[2 * c[2] + 1];
    double d = seg_d2(px, py, x0, y0, x1, y1);
    const double d1 = seg_d2(px, py, x1, y1, x2, y2);
    const double d2 = seg_d2(px, py, x0, y0, x2, y2);
    if (d1 < d) d = d1;
    if (d2 < d) d = d2;
    return d;
}

/* returns number of points that needed the closest-cell fallback */
int orc_locate(const double *pts, int np, const double *x, const int *cells, int nc, double tol,
               int *cell_out, double *miss_d2)
{
    int nmiss = 0;
    for (int i = 0; i < np; ++i) {
        const double px = pts[2 * i], py = pts[2 * i + 1];
        int found = -1;
        for (int c = 0; c < nc; ++c) {
            double l0, l1, l2;
            bary(x, cells + 3 * c, px, py, &l0, &l1, &l2);
            double m = l0 < l1 ? l0 : l1;
            m = m < l2 ? m : l2;
            if (m >= -tol) { found = c; break; }
        }
        double bestd = 0.0;
        if (found < 0) {
            bestd = INFINITY;
            for (int c = 0; c < nc; ++c) {
                const double d = tri_d2(x, cells + 3 * c, px, py);
                if (d < bestd) { bestd = d; found = c; }
            }
            nmiss += 1;
        }
        cell_out[i] = found;
        if (miss_d2) miss_d2[i] = bestd;
    }
    return nmiss;
}

/* ---- A.7: P2 vector / P1 scalar evaluation in the located source cell ----------
 * U  : [T][np2_src][2]  (dofs: vertices 0..V0-1, then edges V0+e)
 * P  : [T][V0]
 * out_u : [T][np][2], out_p : [T][np]  (out_p may be NULL, or computed for first np_p points)
 * cell_edges[c][i] = edge opposite local vertex i.
 */
void orc_eval_fields(const double *pts, int np, int np_p, const int *cell_of, const double *x, const int *cells,
                     const int *cell_edges, int nv0, int np2_src, int T, const double *U, const double *P,
                     double *out_u, double *out_p)
{
    for (int i = 0; i < np; ++i) {
        const int c = cell_of[i];
        const int *cv = cells + 3 * c;
        const int *ce = cell_edges + 3 * c;
        double l[3];
        bary(x, cv, pts[2 * i], pts[2 * i + 1], &l[0], &l[1], &l[2]);
        double phi[6];
        phi[0] = l[0] * (2.0 * l[0] - 1.0);
        phi[1] = l[1] * (2.0 * l[1] - 1.0);
        phi[2] = l[2] * (2.0 * l[2] - 1.0);
        phi[3] = 4.0 * l[1] * l[2];
        phi[4] = 4.0 * l[0] * l[2];
        phi[5] = 4.0 * l[0] * l[1];
        int dof[6] = { cv[0], cv[1], cv[2], nv0 + ce[0], nv0 + ce[1], nv0 + ce[2] };
        for (int t = 0; t < T; ++t) {
            const double *Ut = U + (size_t)t * np2_src * 2;
            double ux = 0.0, uy = 0.0;
            for (int k = 0; k < 6; ++k) {
                ux += phi[k] * Ut[2 * dof[k]];
                uy += phi[k] * Ut[2 * dof[k] + 1];
            }
            out_u[((size_t)t * np + i) * 2] = ux;
            out_u[((size_t)t * np + i) * 2 + 1] = uy;
            if (out_p && i < np_p) {
                const double *Pt = P + (size_t)t * nv0;
                double pv = 0.0;
                for (int k = 0; k < 3; ++k) pv += l[k] * Pt[cv[k]];
                out_p[(size_t)t * np_p + i] = pv;
            }
        }
    }
}

/* ---- A.9: drag/lift = sum over tag-1 facets of |f| (sigma(m_f) n) . e_{x,y} -----
 * facets: list of (cell, local index i of the opposite vertex), ascending edge id.
 * U [T][np2][2], P [T][nv]; out drag[T], lift[T].
 */
void orc_drag_lift(const double *x, const int *cells, const int *cell_edges, int nv, int np2, int T,
                   const double *U, const double *P, const int *facet_cell, const int *facet_local, int nf,
                   double mu, double *drag, double *lift)
{
    for (int t = 0; t < T; ++t) {
        const double *Ut = U + (size_t)t * np2 * 2;
        const double *Pt = P + (size_t)t * nv;
        double D = 0.0, L = 0.0;
        for (int f = 0; f < nf; ++f) {
            const int c = facet_cell[f], k = facet_local[f];
            const int *cv = cells + 3 * c;
            const int *ce = cell_edges + 3 * c;
            const double X[3] = { x[2 * cv[0]], x[2 * cv[1]], x[2 * cv[2]] };
            const double Y[3] = { x[2 * cv[0] + 1], x[2 * cv[1] + 1], x[2 * cv[2] + 1] };
            const double det = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0]);
            /* gradients of the barycentric coordinates */
            double gx[3], gy[3];
            gx[0] = (Y[1] - Y[2]) / det; gy[0] = (X[2] - X[1]) / det;
            gx[1] = (Y[2] - Y[0]) / det; gy[1] = (X[0] - X[2]) / det;
            gx[2] = (Y[0] - Y[1]) / det; gy[2] = (X[1] - X[0]) / det;
            double l[3] = { 0.5, 0.5, 0.5 };
            l[k] = 0.0;
            /* P2 basis gradients at the facet midpoint */
            double bx[6], by[6];
            for (int a = 0; a < 3; ++a) {
                const double s = 4.0 * l[a] - 1.0;
                bx[a] = s * gx[a];
                by[a] = s * gy[a];
            }
            bx[3] = 4.0 * (l[1] * gx[2] + l[2] * gx[1]); by[3] = 4.0 * (l[1] * gy[2] + l[2] * gy[1]);
            bx[4] = 4.0 * (l[0] * gx[2] + l[2] * gx[0]); by[4] = 4.0 * (l[0] * gy[2] + l[2] * gy[0]);
            bx[5] = 4.0 * (l[0] * gx[1] + l[1] * gx[0]); by[5] = 4.0 * (l[0] * gy[1] + l[1] * gy[0]);
            const int dof[6] = { cv[0], cv[1], cv[2], nv + ce[0], nv + ce[1], nv + ce[2] };
            double uxx = 0.0, uxy = 0.0, uyx = 0.0, uyy = 0.0; /* u{i}{j} = d u_i / d x_j */
            for (int a = 0; a < 6; ++a) {
                const double u0 = Ut[2 * dof[a]], u1 = Ut[2 * dof[a] + 1];
                uxx += u0 * bx[a]; uxy += u0 * by[a];
                uyx += u1 * bx[a]; uyy += u1 * by[a];
            }
            const int i = (k + 1) % 3, j = (k + 2) % 3;
            const double pm = 0.5 * Pt[cv[i]] + 0.5 * Pt[cv[j]];
            /* facet vector, length and outward normal (points away from vertex k) */
            const double ex = X[j] - X[i], ey = Y[j] - Y[i];
            const double len = sqrt(ex * ex + ey * ey);
            double nx = ey / len, ny = -ex / len;
            const double mx = 0.5 * X[i] + 0.5 * X[j], my = 0.5 * Y[i] + 0.5 * Y[j];
            if (nx * (X[k] - mx) + ny * (Y[k] - my) > 0.0) { nx = -nx; ny = -ny; }
            const double sxx = 2.0 * mu * uxx - pm;
            const double sxy = mu * (uxy + uyx);
            const double syy = 2.0 * mu * uyy - pm;
            D += len * (sxx * nx + sxy * ny);
            L += len * (sxy * nx + syy * ny);
        }
        drag[t] = D;
        lift[t] = L;
    }
}
