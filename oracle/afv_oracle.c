/*
 * oracle/afv_oracle.c -- CPU restatement of PyAFV's build() path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product package (pyafv_b200/) may import, link or execute this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against (i) the reference's own
 * MATLAB golden vectors (tests/golden/init_forces.csv, varying_A0_forces.csv -- copies of the data the
 * reference's tests/test_core.py:5-35 and tests/test_vary_A0.py:4-39 assert to 1e-8) and (ii) outputs of
 * the unmodified reference (v0.4.16, Cython backend) run in the build container and committed as
 * tests/golden/ref_NAME.npz by tests/golden/make_golden.py.
 *
 * What is restated, and from where (all citations into /root/reference/pyafv):
 *   - The tessellation.  The reference calls scipy.spatial.Voronoi (Qhull 2020.2 as vendored by SciPy
 *     >= 1.13.1, finite_voronoi.py:182) -- a third-party dependency whose source is not under
 *     /root/reference.  A finite-Voronoi cell is (Voronoi cell of i) INTERSECT (closed disk of radius r
 *     about r_i), and a bisector with |r_j - r_i| >= 2r cannot touch that disk, so the ring of cell i is
 *     a function of the neighbours within 2r only.  We therefore restate the *published definition*
 *     (Voronoi cell = intersection of bisector half-planes) by clipping a bounding square with the
 *     bisectors of all neighbours within 2r.  The far corners of the square play the role of the
 *     reference's "extension vertices" for hull rays (finite_voronoi.py:224-278): they lie outside the
 *     disk and never enter a ring.
 *   - Ring semantics (entry order, <= vs <, clockwise orientation): finite_voronoi.py:363-459.
 *   - Per-cell area / perimeter / arc length and the per-vertex dA/dh, dP/dh, arc-end sensitivities:
 *     cell_geom.pyx:402-406 (isolated cell), :432-475, :478-518, :521-539.
 *   - Force assembly (scatter over shared vertices, circumcentre Jacobian in its alpha form, outer-vertex
 *     Jacobian with the max(root, delta) regularisation, arc-angle terms): finite_voronoi.py:529-754.
 *   - Connections: finite_voronoi.py:1110-1150 (segment-to-centre distance < r).
 *   - Periodic boundaries: the reference tiles images within min(4.01 r, L) (_connectutil.py:12-96) and
 *     keeps rows [:N]; here the same images are reached by minimum-image displacements (requires L >= 4r
 *     so that at most one image of a cell lies within 2r).
 *
 * The arithmetic deliberately follows the reference's formulation (absolute atan2 angles, alpha-form
 * Jacobians, per-vertex scatter) and NOT the gather formulation used by the CUDA kernels, so that the two
 * are independent computations of the same quantities.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAXV 256 /* polygon capacity per cell; far above any non-degenerate case */
#define ARC (-1)        /* neighbour label of an arc edge */

typedef struct {
    double r, P0, KA, KP, Lambda, delta;
} oracle_params;

/* one ring entry of cell i: vertex position (relative to r_i), type of the edge LEAVING it
 * (1 straight, 0 arc), and the labels (global cell ids, ARC for the disk) of the edges entering/leaving */
typedef struct {
    double x, y;
    int type;
    int64_t nb_in, nb_out;
} ring_entry;

/* one (vertex, cell) incidence used for the scatter over shared vertices */
typedef struct {
    int64_t k0, k1, k2; /* sorted triple for inner vertices; (I, J, -1 - which_end) for outer */
    int64_t cell;
    double gAx, gAy; /* (A_c - A0_c) * dA_c/dh        cell_geom.pyx:515-516 */
    double gPx, gPy; /* (P_c - P0)   * dP_c/dh        cell_geom.pyx:517-518 */
    double da, dl;   /* arc-end weights, outer only    cell_geom.pyx:523-539 */
    double hx, hy;   /* vertex position relative to the cell */
} incidence;

typedef struct {
    int64_t N;
    int64_t *ring_off;  /* N+1 */
    ring_entry *ring;   /* ring_off[N] entries */
    int64_t n_conn;
    int64_t *conn;      /* n_conn x 2, i<j, sorted lexicographically */
    char err[256];
} oracle_ctx;

static double wrap_mi(double d, double L) {
    if (L > 0.0) {
        if (d > 0.5 * L) d -= L;
        else if (d < -0.5 * L) d += L;
    }
    return d;
}

static int cmp_inc(const void *a, const void *b) {
    const incidence *p = (const incidence *)a, *q = (const incidence *)b;
    if (p->k0 != q->k0) return p->k0 < q->k0 ? -1 : 1;
    if (p->k1 != q->k1) return p->k1 < q->k1 ? -1 : 1;
    if (p->k2 != q->k2) return p->k2 < q->k2 ? -1 : 1;
    if (p->cell != q->cell) return p->cell < q->cell ? -1 : 1;
    return 0;
}

static int cmp_pair(const void *a, const void *b) {
    const int64_t *p = (const int64_t *)a, *q = (const int64_t *)b;
    if (p[0] != q[0]) return p[0] < q[0] ? -1 : 1;
    if (p[1] != q[1]) return p[1] < q[1] ? -1 : 1;
    return 0;
}

void *afv_oracle_create(void) { return calloc(1, sizeof(oracle_ctx)); }

static void ctx_clear(oracle_ctx *c) {
    free(c->ring_off); free(c->ring); free(c->conn);
    c->ring_off = NULL; c->ring = NULL; c->conn = NULL; c->n_conn = 0; c->N = 0;
}

void afv_oracle_destroy(void *h) {
    if (!h) return;
    ctx_clear((oracle_ctx *)h);
    free(h);
}

const char *afv_oracle_last_error(void *h) { return ((oracle_ctx *)h)->err; }

/* line n.x = c intersected with line m.x = d */
static void line_isect(double nx, double ny, double c, double mx, double my, double d, double *x, double *y) {
    double det = nx * my - ny * mx;
    *x = (c * my - d * ny) / det;
    *y = (nx * d - mx * c) / det;
}

/*
 * Build everything for N cells.  xy (N,2) row-major; A0 (N,); L = 0 for open boundaries or the side of the
 * periodic square [0,L)^2.  Outputs: forces (N,2), areas, perims, arclens (N,), coord (N,) int64.
 * Returns 0 on success, negative on error (message via afv_oracle_last_error).
 */
int afv_oracle_build(void *h, int64_t N, const double *xy, const double *A0, const oracle_params *prm, double L,
                     double *forces, double *areas, double *perims, double *arclens, int64_t *coord) {
    oracle_ctx *ctx = (oracle_ctx *)h;
    const double r = prm->r, P0 = prm->P0, KA = prm->KA, KP = prm->KP, Lam = prm->Lambda, delta = prm->delta;
    const double PI = 3.14159265358979323846, TWO_PI = 6.2831853071795864769;
    ctx_clear(ctx);
    ctx->err[0] = 0;
    if (N <= 0) { ctx->N = 0; ctx->ring_off = calloc(1, sizeof(int64_t)); return 0; }
    if (L > 0.0 && L < 4.0 * r) { snprintf(ctx->err, sizeof ctx->err, "periodic box L must be >= 4 r"); return -1; }

    /* ---- uniform grid of side >= 2r ---- */
    double xmin = 0, ymin = 0, xmax = L, ymax = L;
    if (L <= 0.0) {
        xmin = xmax = xy[0]; ymin = ymax = xy[1];
        for (int64_t i = 1; i < N; i++) {
            if (xy[2*i] < xmin) xmin = xy[2*i];
            if (xy[2*i] > xmax) xmax = xy[2*i];
            if (xy[2*i+1] < ymin) ymin = xy[2*i+1];
            if (xy[2*i+1] > ymax) ymax = xy[2*i+1];
        }
    }
    const double cut = 2.0 * r;
    int64_t nbx, nby; double bsx, bsy;
    if (L > 0.0) { nbx = nby = (int64_t)floor(L / cut); if (nbx < 1) nbx = nby = 1; bsx = bsy = L / (double)nbx; }
    else {
        nbx = (int64_t)floor((xmax - xmin) / cut) + 1; nby = (int64_t)floor((ymax - ymin) / cut) + 1;
        /* keep the grid modest for pathological spans */
        while (nbx * nby > 4 * N + 64) { if (nbx >= nby) nbx = (nbx + 1) / 2; else nby = (nby + 1) / 2; }
        bsx = (xmax - xmin) / (double)nbx; if (bsx < cut) bsx = cut;
        bsy = (ymax - ymin) / (double)nby; if (bsy < cut) bsy = cut;
    }
    int64_t nb = nbx * nby;
    int64_t *bstart = calloc((size_t)nb + 1, sizeof(int64_t));
    int64_t *bidx = malloc((size_t)N * sizeof(int64_t));
    int64_t *cbin = malloc((size_t)N * sizeof(int64_t));
    for (int64_t i = 0; i < N; i++) {
        double x = xy[2*i], y = xy[2*i+1];
        if (L > 0.0) { x -= L * floor(x / L); y -= L * floor(y / L); }
        int64_t bx = (int64_t)floor((x - xmin) / bsx), by = (int64_t)floor((y - ymin) / bsy);
        if (bx < 0) bx = 0; if (bx >= nbx) bx = nbx - 1;
        if (by < 0) by = 0; if (by >= nby) by = nby - 1;
        cbin[i] = by * nbx + bx;
        bstart[cbin[i] + 1]++;
    }
    for (int64_t b = 0; b < nb; b++) bstart[b + 1] += bstart[b];
    {
        int64_t *fill = malloc((size_t)nb * sizeof(int64_t));
        memcpy(fill, bstart, (size_t)nb * sizeof(int64_t));
        for (int64_t i = 0; i < N; i++) bidx[fill[cbin[i]]++] = i;
        free(fill);
    }

    /* ---- pass 1: rings, per-cell measures, vertex incidences ---- */
    ctx->N = N;
    ctx->ring_off = calloc((size_t)N + 1, sizeof(int64_t));
    int64_t ring_cap = 8 * N + 64, ring_n = 0;
    ctx->ring = malloc((size_t)ring_cap * sizeof(ring_entry));
    int64_t inc_cap = 8 * N + 64, inc_n = 0;
    incidence *inc = malloc((size_t)inc_cap * sizeof(incidence));
    int64_t conn_cap = 4 * N + 64, conn_n = 0;
    int64_t *conn = malloc((size_t)conn_cap * 2 * sizeof(int64_t));
    int rc = 0;

    int64_t cand_cap = 256;
    int64_t *cand_id = malloc((size_t)cand_cap * sizeof(int64_t));
    double *cand_x = malloc((size_t)cand_cap * sizeof(double));
    double *cand_y = malloc((size_t)cand_cap * sizeof(double));

    const double B = 4.0 * r; /* half side of the bounding square (any value > r works) */

    for (int64_t i = 0; i < N && rc == 0; i++) {
        /* candidates within 2r (strict: a bisector at distance exactly r cannot cut the open disk) */
        int64_t nc = 0;
        int64_t bx = cbin[i] % nbx, by = cbin[i] / nbx;
        for (int64_t dy = -1; dy <= 1; dy++) for (int64_t dx = -1; dx <= 1; dx++) {
            int64_t cx = bx + dx, cy = by + dy;
            if (L > 0.0) {
                /* with fewer than 3 bins per side the same bin would be visited twice */
                if (nbx < 3 && (dx == -1 || (nbx == 1 && dx == 1))) continue;
                if (nby < 3 && (dy == -1 || (nby == 1 && dy == 1))) continue;
                cx = (cx % nbx + nbx) % nbx; cy = (cy % nby + nby) % nby;
            } else if (cx < 0 || cy < 0 || cx >= nbx || cy >= nby) continue;
            int64_t b = cy * nbx + cx;
            for (int64_t s = bstart[b]; s < bstart[b + 1]; s++) {
                int64_t j = bidx[s];
                if (j == i) continue;
                double px = wrap_mi(xy[2*j] - xy[2*i], L), py = wrap_mi(xy[2*j+1] - xy[2*i+1], L);
                if (px * px + py * py < cut * cut) {
                    if (nc == cand_cap) {
                        cand_cap *= 2;
                        cand_id = realloc(cand_id, (size_t)cand_cap * sizeof(int64_t));
                        cand_x = realloc(cand_x, (size_t)cand_cap * sizeof(double));
                        cand_y = realloc(cand_y, (size_t)cand_cap * sizeof(double));
                    }
                    cand_id[nc] = j; cand_x[nc] = px; cand_y[nc] = py; nc++;
                }
            }
        }

        /* convex polygon, counter-clockwise, as lines n.x = c; edge k runs from vertex k to vertex k+1 */
        double vx[ORACLE_MAXV], vy[ORACLE_MAXV], wx[ORACLE_MAXV], wy[ORACLE_MAXV];
        double lnx[ORACLE_MAXV], lny[ORACLE_MAXV], lc[ORACLE_MAXV], mnx[ORACLE_MAXV], mny[ORACLE_MAXV], mc[ORACLE_MAXV];
        int64_t lab[ORACLE_MAXV], mlab[ORACLE_MAXV];
        int m = 4;
        vx[0] = -B; vy[0] = -B; lnx[0] = 0;  lny[0] = -1; lc[0] = B; lab[0] = -1; /* bottom */
        vx[1] = B;  vy[1] = -B; lnx[1] = 1;  lny[1] = 0;  lc[1] = B; lab[1] = -1; /* right  */
        vx[2] = B;  vy[2] = B;  lnx[2] = 0;  lny[2] = 1;  lc[2] = B; lab[2] = -1; /* top    */
        vx[3] = -B; vy[3] = B;  lnx[3] = -1; lny[3] = 0;  lc[3] = B; lab[3] = -1; /* left   */
        for (int64_t q = 0; q < nc; q++) {
            double nx = cand_x[q], ny = cand_y[q], c = 0.5 * (nx * nx + ny * ny);
            /* signed distances */
            int any_out = 0, any_in = 0;
            double s[ORACLE_MAXV];
            for (int k = 0; k < m; k++) { s[k] = nx * vx[k] + ny * vy[k] - c; if (s[k] > 0) any_out = 1; else any_in = 1; }
            if (!any_out) continue;
            if (!any_in) { m = 0; break; }
            int mm = 0;
            for (int k = 0; k < m; k++) {
                int k1 = (k + 1) % m;
                int in0 = s[k] <= 0, in1 = s[k1] <= 0;
                if (in0) {
                    /* keep vertex k; the edge leaving it is edge k (possibly shortened) */
                    if (mm >= ORACLE_MAXV - 2) { rc = -2; break; }
                    wx[mm] = vx[k]; wy[mm] = vy[k]; mnx[mm] = lnx[k]; mny[mm] = lny[k]; mc[mm] = lc[k]; mlab[mm] = lab[k]; mm++;
                    if (!in1) {
                        /* leaving: new vertex on edge k, from which the new bisector edge departs */
                        line_isect(lnx[k], lny[k], lc[k], nx, ny, c, &wx[mm], &wy[mm]);
                        mnx[mm] = nx; mny[mm] = ny; mc[mm] = c; mlab[mm] = cand_id[q]; mm++;
                    }
                } else if (in1) {
                    /* re-entering on edge k: new vertex, edge k continues from it */
                    if (mm >= ORACLE_MAXV - 2) { rc = -2; break; }
                    line_isect(lnx[k], lny[k], lc[k], nx, ny, c, &wx[mm], &wy[mm]);
                    mnx[mm] = lnx[k]; mny[mm] = lny[k]; mc[mm] = lc[k]; mlab[mm] = lab[k]; mm++;
                }
            }
            if (rc) break;
            m = mm;
            memcpy(vx, wx, sizeof(double) * m); memcpy(vy, wy, sizeof(double) * m);
            memcpy(lnx, mnx, sizeof(double) * m); memcpy(lny, mny, sizeof(double) * m); memcpy(lc, mc, sizeof(double) * m);
            memcpy(lab, mlab, sizeof(int64_t) * m);
        }
        if (rc) { snprintf(ctx->err, sizeof ctx->err, "polygon capacity exceeded at cell %lld", (long long)i); break; }

        /* ---- ring, clockwise (finite_voronoi.py:363-459).  Clockwise edge e: V1 = v[k+1] -> V2 = v[k] ---- */
        ring_entry ring[3 * ORACLE_MAXV];
        int E = 0;
        for (int kk = 0; kk < m; kk++) {
            int k = m - 1 - kk;                 /* ccw edge index, visited in clockwise order */
            int k1 = (k + 1) % m;
            int64_t j = lab[k];
            if (j < 0) continue;                /* box edge == the reference's ray-ray pair, cell_geom.pyx:221-229 */
            double V1x = vx[k1], V1y = vy[k1], V2x = vx[k], V2y = vy[k];
            /* label of the clockwise-previous edge = ccw edge k+1 */
            int64_t jprev = lab[k1];
            int contact = 0;
            /* inner start vertex: d1 <= r (finite_voronoi.py:375-377) */
            if (sqrt(V1x * V1x + V1y * V1y) <= r) {
                ring[E].x = V1x; ring[E].y = V1y; ring[E].type = 1; ring[E].nb_in = jprev; ring[E].nb_out = j; E++;
                contact = 1;
            }
            /* bisector-circle points (finite_voronoi.py:381-407) */
            double px = lnx[k], py = lny[k];   /* = r_j - r_i */
            double Cx = 0.5 * px, Cy = 0.5 * py;
            double d = sqrt(Cx * Cx + Cy * Cy);
            if (d < r) {
                double dVx = V2x - V1x, dVy = V2y - V1y;
                double segL = sqrt(dVx * dVx + dVy * dVy);
                double den = segL > 0.0 ? segL : 1.0;
                double t1 = ((V1x - 0.0) * dVx + (V1y - 0.0) * dVy) / den; /* = -t of the reference */
                double t2 = t1 + den;
                double tr = sqrt(r * r - d * d);
                double ux = dVx / den, uy = dVy / den;
                if (-tr < t2 && -tr > t1) {
                    ring[E].x = Cx - tr * ux; ring[E].y = Cy - tr * uy; ring[E].type = 1; ring[E].nb_in = ARC; ring[E].nb_out = j; E++;
                    contact = 1;
                }
                if (tr < t2 && tr > t1) {
                    ring[E].x = Cx + tr * ux; ring[E].y = Cy + tr * uy; ring[E].type = 0; ring[E].nb_in = j; ring[E].nb_out = ARC; E++;
                }
            }
            if (contact && i < j) {
                if (conn_n == conn_cap) { conn_cap *= 2; conn = realloc(conn, (size_t)conn_cap * 2 * sizeof(int64_t)); }
                conn[2 * conn_n] = i; conn[2 * conn_n + 1] = j; conn_n++;
            }
        }

        /* store ring */
        if (ring_n + E > ring_cap) { ring_cap = 2 * (ring_n + E); ctx->ring = realloc(ctx->ring, (size_t)ring_cap * sizeof(ring_entry)); }
        memcpy(ctx->ring + ring_n, ring, (size_t)E * sizeof(ring_entry));
        ring_n += E;
        ctx->ring_off[i + 1] = ring_n;

        /* ---- per-cell measures (cell_geom.pyx:402-475) ---- */
        if (E < 2) {
            areas[i] = PI * r * r; perims[i] = 2.0 * PI * r; arclens[i] = 2.0 * PI * r; coord[i] = 0;
            /* a ring of < 2 entries carries no force terms; drop it like the reference does */
            continue;
        }
        /* NB: the reference evaluates atan2 on absolute coordinates minus the centre; the ring here is
         * already relative to the centre, which is the same quantity with less cancellation. */
        double Pstr = 0, Astr = 0, Parc = 0, Aarc = 0, dang[3 * ORACLE_MAXV];
        int z = 0;
        for (int e = 0; e < E; e++) {
            int e2 = (e + 1) % E;
            double a1 = atan2(ring[e].y, ring[e].x), a2 = atan2(ring[e2].y, ring[e2].x);
            double da = a1 - a2; if (da < 0.0) da += TWO_PI;
            dang[e] = da;
            if (ring[e].type == 1) {
                double sx = ring[e].x - ring[e2].x, sy = ring[e].y - ring[e2].y;
                Pstr += sqrt(sx * sx + sy * sy);
                Astr += -0.5 * (ring[e].x * ring[e2].y - ring[e].y * ring[e2].x);
                z++;
            } else { Parc += r * da; Aarc += 0.5 * r * r * da; }
        }
        double Pi_ = Pstr + Parc, Ai_ = Astr + Aarc;
        areas[i] = Ai_; perims[i] = Pi_; arclens[i] = Parc; coord[i] = z;

        /* ---- vertex incidences (cell_geom.pyx:478-539) ---- */
        for (int e = 0; e < E; e++) {
            int e2 = (e + 1) % E, e0 = (e - 1 + E) % E;
            incidence I; memset(&I, 0, sizeof I);
            I.cell = i; I.hx = ring[e].x; I.hy = ring[e].y;
            /* dA/dv1 = -0.5 perp(v2-c) + 0.5 perp(v0-c), perp(u) = (uy, -ux) */
            double dAx = -0.5 * ring[e2].y + 0.5 * ring[e0].y;
            double dAy = 0.5 * ring[e2].x - 0.5 * ring[e0].x;
            double dPx = 0, dPy = 0;
            if (ring[e].type == 1) {
                double sx = ring[e].x - ring[e2].x, sy = ring[e].y - ring[e2].y, l = sqrt(sx * sx + sy * sy);
                if (l > 0.0) { dPx += sx / l; dPy += sy / l; }
            }
            if (ring[e0].type == 1) {
                double sx = ring[e].x - ring[e0].x, sy = ring[e].y - ring[e0].y, l = sqrt(sx * sx + sy * sy);
                if (l > 0.0) { dPx += sx / l; dPy += sy / l; }
            }
            I.gAx = (Ai_ - A0[i]) * dAx; I.gAy = (Ai_ - A0[i]) * dAy;
            I.gPx = (Pi_ - P0) * dPx;    I.gPy = (Pi_ - P0) * dPy;
            if (ring[e].nb_in != ARC && ring[e].nb_out != ARC) {
                /* inner vertex: key = sorted triple */
                int64_t a = i, b = ring[e].nb_in, c = ring[e].nb_out, t;
                if (a > b) { t = a; a = b; b = t; }
                if (b > c) { t = b; b = c; c = t; }
                if (a > b) { t = a; a = b; b = t; }
                I.k0 = a; I.k1 = b; I.k2 = c;
            } else {
                /* outer vertex on the bisector with j; which end: sign((h - r_J) x (r_I - r_J)), I<J
                 * (finite_voronoi.py:665-666) */
                int64_t j = ring[e].nb_in != ARC ? ring[e].nb_in : ring[e].nb_out;
                double pjx = wrap_mi(xy[2*j] - xy[2*i], L), pjy = wrap_mi(xy[2*j+1] - xy[2*i+1], L);
                double rIx, rIy, rJx, rJy; /* relative to cell i */
                int64_t Ii, Jj;
                if (i < j) { Ii = i; Jj = j; rIx = 0; rIy = 0; rJx = pjx; rJy = pjy; }
                else       { Ii = j; Jj = i; rIx = pjx; rIy = pjy; rJx = 0; rJy = 0; }
                double sg = (I.hx - rJx) * (rIy - rJy) - (I.hy - rJy) * (rIx - rJx);
                I.k0 = Ii; I.k1 = Jj; I.k2 = sg > 0 ? -1 : -2;
                /* arc-end weights: the arc leaving this vertex (type 0) or arriving at it */
                if (ring[e].type == 0) { I.da = 0.5 * r * r * (1.0 - cos(dang[e])); I.dl = r; }
                else                   { I.da = -0.5 * r * r * (1.0 - cos(dang[e0])); I.dl = -r; }
            }
            if (inc_n == inc_cap) { inc_cap *= 2; inc = realloc(inc, (size_t)inc_cap * sizeof(incidence)); }
            inc[inc_n++] = I;
        }
    }

    /* ---- pass 2: scatter over shared vertices (finite_voronoi.py:529-754) ---- */
    if (rc == 0) {
        for (int64_t i = 0; i < 2 * N; i++) forces[i] = 0.0;
        double *dEarc = calloc((size_t)2 * N, sizeof(double));
        qsort(inc, (size_t)inc_n, sizeof(incidence), cmp_inc);
        int64_t s = 0;
        while (s < inc_n) {
            int64_t e = s + 1;
            while (e < inc_n && inc[e].k0 == inc[s].k0 && inc[e].k1 == inc[s].k1 && inc[e].k2 == inc[s].k2) e++;
            /* dE_poly/dh = 2 (KA dA_poly_dh + KP dP_poly_dh) summed over the cells sharing h */
            double Gx = 0, Gy = 0;
            for (int64_t q = s; q < e; q++) {
                Gx += 2.0 * (KA * inc[q].gAx + KP * inc[q].gPx);
                Gy += 2.0 * (KA * inc[q].gAy + KP * inc[q].gPy);
            }
            if (inc[s].k2 >= 0) {
                /* inner vertex (i,j,k): frame of cell i = k0 */
                int64_t I = inc[s].k0, J = inc[s].k1, K = inc[s].k2;
                double rix = 0, riy = 0;
                double rjx = wrap_mi(xy[2*J] - xy[2*I], L), rjy = wrap_mi(xy[2*J+1] - xy[2*I+1], L);
                double rkx = wrap_mi(xy[2*K] - xy[2*I], L), rky = wrap_mi(xy[2*K+1] - xy[2*I+1], L);
                double jkx = rjx - rkx, jky = rjy - rky;      /* rj - rk */
                double ijx = rix - rjx, ijy = riy - rjy;      /* ri - rj */
                double ikx = rix - rkx, iky = riy - rky;      /* ri - rk */
                double D0 = ijx * jky - ijy * jkx;
                double D = 2.0 * D0 * D0;
                double jk2 = jkx * jkx + jky * jky, ik2 = ikx * ikx + iky * iky, ij2 = ijx * ijx + ijy * ijy;
                double al_i = jk2 * (ijx * ikx + ijy * iky) / D;
                double al_j = ik2 * ((-ijx) * jkx + (-ijy) * jky) / D;
                double al_k = ij2 * ((-ikx) * (-jkx) + (-iky) * (-jky)) / D;
                /* d/d ri */
                double czx = jky, czy = -jkx;
                double tjx = 2.0 * ikx / ik2 - 2.0 * czx / D0, tjy = 2.0 * iky / ik2 - 2.0 * czy / D0;
                double tkx = 2.0 * ijx / ij2 - 2.0 * czx / D0, tky = 2.0 * ijy / ij2 - 2.0 * czy / D0;
                double daj_x = (ik2 / D) * (-jkx) + al_j * tjx, daj_y = (ik2 / D) * (-jky) + al_j * tjy;
                double dak_x = (ij2 / D) * (jkx) + al_k * tkx,  dak_y = (ij2 / D) * (jky) + al_k * tky;
                /* dh/dxi (2-vector), dh/dyi */
                double hxi_x = al_i + daj_x * (rjx - rix) + dak_x * (rkx - rix), hxi_y = daj_x * (rjy - riy) + dak_x * (rky - riy);
                double hyi_x = daj_y * (rjx - rix) + dak_y * (rkx - rix),        hyi_y = al_i + daj_y * (rjy - riy) + dak_y * (rky - riy);
                /* d/d rj */
                czx = -iky; czy = ikx;
                double tix = 2.0 * jkx / jk2 - 2.0 * czx / D0, tiy = 2.0 * jky / jk2 - 2.0 * czy / D0;
                tkx = 2.0 * (-ijx) / ij2 - 2.0 * czx / D0; tky = 2.0 * (-ijy) / ij2 - 2.0 * czy / D0;
                double dai_x = (jk2 / D) * (-ikx) + al_i * tix, dai_y = (jk2 / D) * (-iky) + al_i * tiy;
                dak_x = (ij2 / D) * (ikx) + al_k * tkx; dak_y = (ij2 / D) * (iky) + al_k * tky;
                double hxj_x = dai_x * (rix - rjx) + al_j + dak_x * (rkx - rjx), hxj_y = dai_x * (riy - rjy) + dak_x * (rky - rjy);
                double hyj_x = dai_y * (rix - rjx) + dak_y * (rkx - rjx),        hyj_y = dai_y * (riy - rjy) + al_j + dak_y * (rky - rjy);
                /* d/d rk */
                czx = ijy; czy = -ijx;
                tix = 2.0 * (-jkx) / jk2 - 2.0 * czx / D0; tiy = 2.0 * (-jky) / jk2 - 2.0 * czy / D0;
                tjx = 2.0 * (-ikx) / ik2 - 2.0 * czx / D0; tjy = 2.0 * (-iky) / ik2 - 2.0 * czy / D0;
                dai_x = (jk2 / D) * (-ijx) + al_i * tix; dai_y = (jk2 / D) * (-ijy) + al_i * tiy;
                daj_x = (ik2 / D) * (ijx) + al_j * tjx;  daj_y = (ik2 / D) * (ijy) + al_j * tjy;
                double hxk_x = dai_x * (rix - rkx) + daj_x * (rjx - rkx) + al_k, hxk_y = dai_x * (riy - rky) + daj_x * (rjy - rky);
                double hyk_x = dai_y * (rix - rkx) + daj_y * (rjx - rkx),        hyk_y = dai_y * (riy - rky) + daj_y * (rjy - rky) + al_k;
                if (e - s == 3) { /* a proper 3-cell vertex; otherwise the topology is degenerate (excluded tie) */
                    forces[2*I]   += Gx * hxi_x + Gy * hxi_y; forces[2*I+1] += Gx * hyi_x + Gy * hyi_y;
                    forces[2*J]   += Gx * hxj_x + Gy * hxj_y; forces[2*J+1] += Gx * hyj_x + Gy * hyj_y;
                    forces[2*K]   += Gx * hxk_x + Gy * hxk_y; forces[2*K+1] += Gx * hyk_x + Gy * hyk_y;
                } else {
                    /* incomplete vertex (cells disagree on a near-cocircular configuration): apply what we have */
                    forces[2*I]   += Gx * hxi_x + Gy * hxi_y; forces[2*I+1] += Gx * hyi_x + Gy * hyi_y;
                    forces[2*J]   += Gx * hxj_x + Gy * hxj_y; forces[2*J+1] += Gx * hyj_x + Gy * hyj_y;
                    forces[2*K]   += Gx * hxk_x + Gy * hxk_y; forces[2*K+1] += Gx * hyk_x + Gy * hyk_y;
                }
            } else {
                /* outer vertex, pair I<J (finite_voronoi.py:653-743) in the frame of cell I */
                int64_t I = inc[s].k0, J = inc[s].k1;
                double rjx = wrap_mi(xy[2*J] - xy[2*I], L), rjy = wrap_mi(xy[2*J+1] - xy[2*I+1], L);
                /* vertex position in the frame of I: take it from I's incidence if present, else from J's */
                double hx = 0, hy = 0; int have = 0;
                double wAi = 0, wAj = 0, wPi = 0, wPj = 0, dli = 0, dlj = 0; /* per-cell arc weights */
                for (int64_t q = s; q < e; q++) {
                    int64_t c = inc[q].cell;
                    /* recover (A_c - A0_c), (P_c - P0) */
                    double dAc = areas[c] - A0[c], dPc = perims[c] - P0;
                    if (c == I) { if (!have) { hx = inc[q].hx; hy = inc[q].hy; have = 1; } wAi = dAc * inc[q].da; wPi = dPc * inc[q].dl; dli = inc[q].dl; }
                    else        { if (!have) { hx = inc[q].hx + rjx; hy = inc[q].hy + rjy; have = 1; } wAj = dAc * inc[q].da; wPj = dPc * inc[q].dl; dlj = inc[q].dl; }
                }
                double rijx = -rjx, rijy = -rjy; /* ri - rj with ri = 0 */
                double rij = sqrt(rijx * rijx + rijy * rijy);
                double root = sqrt(4.0 * r * r - rij * rij);
                double sg = (hx - rjx) * (0.0 - rjy) - (hy - rjy) * (0.0 - rjx);
                sg = (sg > 0) - (sg < 0);
                double czx = rijy, czy = -rijx;
                double den = (root > delta ? root : delta) * rij * rij * rij;
                double dxt_x = -(2.0 * r * r * rijx * czx / den),  dxt_y = -(2.0 * r * r * rijx * czy / den) - root / (2.0 * rij);
                double dyt_x = -(2.0 * r * r * rijy * czx / den) + root / (2.0 * rij), dyt_y = -(2.0 * r * r * rijy * czy / den);
                double hxi_x = 0.5 + sg * dxt_x, hxi_y = sg * dxt_y;
                double hyi_x = sg * dyt_x,       hyi_y = 0.5 + sg * dyt_y;
                double hxj_x = 0.5 - sg * dxt_x, hxj_y = -sg * dxt_y;
                double hyj_x = -sg * dyt_x,      hyj_y = 0.5 - sg * dyt_y;
                forces[2*I]   += Gx * hxi_x + Gy * hxi_y; forces[2*I+1] += Gx * hyi_x + Gy * hyi_y;
                forces[2*J]   += Gx * hxj_x + Gy * hxj_y; forces[2*J+1] += Gx * hyj_x + Gy * hyj_y;
                /* arc-angle sensitivities */
                double uix = hx, uiy = hy, ujx = hx - rjx, ujy = hy - rjy;
                double iui = 1.0 / (uix * uix + uiy * uiy), iuj = 1.0 / (ujx * ujx + ujy * ujy);
                double upix = -uiy * iui, upiy = uix * iui, upjx = -ujy * iuj, upjy = ujx * iuj;
                double dti_xj = hxj_x * upix + hxj_y * upiy, dti_yj = hyj_x * upix + hyj_y * upiy;
                double dti_xi = -dti_xj, dti_yi = -dti_yj;
                double dtj_xi = hxi_x * upjx + hxi_y * upjy, dtj_yi = hyi_x * upjx + hyi_y * upjy;
                double dtj_xj = -dtj_xi, dtj_yj = -dtj_yi;
                /* dE_arc = 2 (KA dA_arc + KP dP_arc) + Lambda dL */
                dEarc[2*I]   += 2.0 * (KA * (wAi * dti_xi + wAj * dtj_xi) + KP * (wPi * dti_xi + wPj * dtj_xi)) + Lam * (dli * dti_xi + dlj * dtj_xi);
                dEarc[2*I+1] += 2.0 * (KA * (wAi * dti_yi + wAj * dtj_yi) + KP * (wPi * dti_yi + wPj * dtj_yi)) + Lam * (dli * dti_yi + dlj * dtj_yi);
                dEarc[2*J]   += 2.0 * (KA * (wAi * dti_xj + wAj * dtj_xj) + KP * (wPi * dti_xj + wPj * dtj_xj)) + Lam * (dli * dti_xj + dlj * dtj_xj);
                dEarc[2*J+1] += 2.0 * (KA * (wAi * dti_yj + wAj * dtj_yj) + KP * (wPi * dti_yj + wPj * dtj_yj)) + Lam * (dli * dti_yj + dlj * dtj_yj);
            }
            s = e;
        }
        for (int64_t i = 0; i < 2 * N; i++) forces[i] = -(forces[i] + dEarc[i]);
        free(dEarc);
        qsort(conn, (size_t)conn_n, 2 * sizeof(int64_t), cmp_pair);
        ctx->conn = conn; ctx->n_conn = conn_n; conn = NULL;
    }

    free(conn); free(inc); free(cand_id); free(cand_x); free(cand_y); free(bstart); free(bidx); free(cbin);
    return rc;
}

int64_t afv_oracle_ring_total(void *h) { oracle_ctx *c = (oracle_ctx *)h; return c->ring_off ? c->ring_off[c->N] : 0; }

/* copy rings out: off (N+1), type (T) int8, nb_in/nb_out (T) int64 (-1 = arc), vxy (T,2) relative to the cell */
void afv_oracle_get_rings(void *h, int64_t *off, int8_t *type, int64_t *nb_in, int64_t *nb_out, double *vxy) {
    oracle_ctx *c = (oracle_ctx *)h;
    memcpy(off, c->ring_off, (size_t)(c->N + 1) * sizeof(int64_t));
    int64_t T = c->ring_off[c->N];
    for (int64_t t = 0; t < T; t++) {
        type[t] = (int8_t)c->ring[t].type; nb_in[t] = c->ring[t].nb_in; nb_out[t] = c->ring[t].nb_out;
        vxy[2*t] = c->ring[t].x; vxy[2*t+1] = c->ring[t].y;
    }
}

int64_t afv_oracle_n_connections(void *h) { return ((oracle_ctx *)h)->n_conn; }
void afv_oracle_get_connections(void *h, int64_t *pairs) {
    oracle_ctx *c = (oracle_ctx *)h;
    if (c->n_conn) memcpy(pairs, c->conn, (size_t)c->n_conn * 2 * sizeof(int64_t));
}
