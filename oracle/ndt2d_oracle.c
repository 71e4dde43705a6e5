/*
 * SPEC ORACLE — test infrastructure, NOT product code, NOT the reference matcher.
 * PARITY UNPINNED: /root/reference holds only README.md:1 ("# GTSAM-NDT"); there is no
 * reference arithmetic to follow. Each function below cites the SPEC.md section it
 * restates. Scalar loops, f32 per point, f64 sums; see ndt2d_oracle.h for who may call it.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off -mfma: fmaf() is one rounding,
 * nothing else is contracted).
 */
#include "ndt2d_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    float res, st, inv_st, ox, oy;
    double inv_std;      /* SPEC 2 (v4): 1.0 / (double)st */
    double qs, qu;       /* SPEC 3 (v4): fixed-point scale 2^22 / res and its unit res * 2^-22 */
    int ov, nhx, nhy, njx, njy;
    uint32_t *n;  /* per cell */
    int64_t *s;   /* 5 per cell: sx sy sxx sxy syy */
    float *cells; /* 8 per cell */
} level_t;

struct oracle_matcher {
    oracle_params prm;
    int nlevels;
    float res[ORACLE_MAX_LEVELS];
    int explicit_grid;
    float gox, goy, gex, gey;
    int has_target;
    level_t lv[ORACLE_MAX_LEVELS];
};

/* ---------------------------------------------------------------- helpers */

static uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

static void level_free(level_t *L)
{
    free(L->n); free(L->s); free(L->cells);
    L->n = NULL; L->s = NULL; L->cells = NULL;
}

void oracle_default_params(oracle_params *p)
{
    /* SPEC 1 defaults */
    p->eig_ratio = 0.01; p->eps_trans = 1e-4; p->eps_rot = 1e-5;
    p->max_step_trans = 0.5; p->max_step_rot = 0.2;
    p->lambda_init = 1e-3; p->lambda_min = 1e-9; p->lambda_max = 1e7;
    p->lambda_up = 10.0; p->lambda_down = 5.0; p->lambda_fail_up = 3.0;
    p->min_points = 3; p->max_iterations = 30; p->overlap = 0; p->reserved = 0;
}

oracle_matcher *oracle_create(void)
{
    oracle_matcher *m = (oracle_matcher *)calloc(1, sizeof(*m));
    if (!m) return NULL;
    oracle_default_params(&m->prm);
    m->nlevels = 1; m->res[0] = 1.0f;
    return m;
}

static void drop_target(oracle_matcher *m)
{
    for (int l = 0; l < ORACLE_MAX_LEVELS; ++l) level_free(&m->lv[l]);
    m->has_target = 0;
}

void oracle_destroy(oracle_matcher *m)
{
    if (!m) return;
    drop_target(m);
    free(m);
}

int oracle_set_params(oracle_matcher *m, const oracle_params *p)
{
    if (p->min_points < 2 || p->max_iterations < 1 || (p->overlap != 0 && p->overlap != 1)) return 1;
    if (!(p->lambda_up > 1.0) || !(p->lambda_fail_up > 1.0) || !(p->lambda_down >= 1.0)) return 1;
    if (m->has_target && p->overlap != m->prm.overlap) drop_target(m);
    m->prm = *p;
    return 0;
}

int oracle_set_resolutions(oracle_matcher *m, const float *res, int nlevels)
{
    if (nlevels < 1 || nlevels > ORACLE_MAX_LEVELS) return 1;
    for (int l = 0; l < nlevels; ++l) if (!(res[l] > 0.0f) || res[l] > 8.0f) return 1;
    drop_target(m);
    m->nlevels = nlevels;
    for (int l = 0; l < nlevels; ++l) m->res[l] = res[l];
    return 0;
}

int oracle_set_grid(oracle_matcher *m, float ox, float oy, float ex, float ey)
{
    drop_target(m);
    m->explicit_grid = (ex > 0.0f && ey > 0.0f);
    m->gox = ox; m->goy = oy; m->gex = ex; m->gey = ey;
    return 0;
}

/* SPEC 2: geometry of one level, explicit or auto-fit */
static int level_setup(oracle_matcher *m, int l, const float *xy, int n)
{
    level_t *L = &m->lv[l];
    level_free(L);
    L->res = m->res[l];
    L->ov = m->prm.overlap;
    L->st = L->ov ? L->res * 0.5f : L->res;
    L->inv_st = 1.0f / L->st;
    L->inv_std = 1.0 / (double)L->st;
    L->qs = 4194304.0 / (double)L->res;
    L->qu = (double)L->res * (1.0 / 4194304.0);
    if (m->explicit_grid) {
        L->ox = m->gox; L->oy = m->goy;
        L->nhx = (int)ceilf(m->gex / L->st);
        L->nhy = (int)ceilf(m->gey / L->st);
    } else {
        float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
        for (int i = 0; i < n; ++i) {
            float x = xy[2 * i], y = xy[2 * i + 1];
            if (!isfinite(x) || !isfinite(y)) continue;
            if (x < xmin) xmin = x;
            if (x > xmax) xmax = x;
            if (y < ymin) ymin = y;
            if (y > ymax) ymax = y;
        }
        if (!(xmin <= xmax)) { xmin = xmax = ymin = ymax = 0.0f; }
        L->ox = floorf(xmin / L->res) * L->res - L->res;
        L->oy = floorf(ymin / L->res) * L->res - L->res;
        L->nhx = (int)ceilf((xmax - L->ox) / L->st) + 2;
        L->nhy = (int)ceilf((ymax - L->oy) / L->st) + 2;
    }
    if (L->nhx < 1 || L->nhy < 1) return 1;
    L->njx = L->nhx + L->ov;
    L->njy = L->nhy + L->ov;
    size_t nc = (size_t)L->njx * (size_t)L->njy;
    L->n = (uint32_t *)calloc(nc, sizeof(uint32_t));
    L->s = (int64_t *)calloc(nc * 5, sizeof(int64_t));
    L->cells = (float *)calloc(nc * 8, sizeof(float));
    return (L->n && L->s && L->cells) ? 0 : 2;
}

/* SPEC 2 (v4): lattice index of a coordinate pair given in CELL UNITS relative to the lattice origin, f64; returns 0
 * when outside. h = floor(f); the fraction f - h (exact in f64) is rounded once to f32 (dfx, dfy in [0, 1]): the
 * per-point algebra of SPEC 4 works on such cell-local coordinates, never on map-frame f32 coordinates. */
static int lattice(const level_t *L, double fx, double fy, int *hx, int *hy, float *dfx, float *dfy)
{
    double nx = floor(fx), ny = floor(fy);
    if (!((nx >= 0.0) && (nx < (double)L->nhx) && (ny >= 0.0) && (ny < (double)L->nhy))) return 0; /* NaN is outside */
    *hx = (int)nx; *hy = (int)ny;
    if (dfx) { *dfx = (float)(fx - nx); *dfy = (float)(fy - ny); }
    return 1;
}

/* SPEC 2 (v4): a target point (no pose): f = ((double)X - (double)origin) * inv_st */
static int lattice_of_point(const level_t *L, float X, float Y, int *hx, int *hy)
{
    return lattice(L, ((double)X - (double)L->ox) * L->inv_std, ((double)Y - (double)L->oy) * L->inv_std, hx, hy, NULL, NULL);
}

/* SPEC 3: integer accumulation of one point into its K cells */
static void accumulate(level_t *L, const float *xy, int n)
{
    const int K = L->ov ? 2 : 1;
    for (int i = 0; i < n; ++i) {
        float X = xy[2 * i], Y = xy[2 * i + 1];
        int hx, hy;
        if (!lattice_of_point(L, X, Y, &hx, &hy)) continue;
        for (int b = 0; b < K; ++b) for (int a = 0; a < K; ++a) {
            int jx = hx + a, jy = hy + b;
            double cx = (double)L->ox + ((double)(jx - L->ov)) * (double)L->st + 0.5 * (double)L->res;
            double cy = (double)L->oy + ((double)(jy - L->ov)) * (double)L->st + 0.5 * (double)L->res;
            double dx = (double)X - cx, dy = (double)Y - cy;
            int64_t qx = llrint(dx * L->qs), qy = llrint(dy * L->qs);
            size_t c = (size_t)jy * (size_t)L->njx + (size_t)jx;
            L->n[c] += 1;
            int64_t *s = L->s + 5 * c;
            s[0] += qx; s[1] += qy; s[2] += qx * qx; s[3] += qx * qy; s[4] += qy * qy;
        }
    }
}

/* SPEC 3: finalisation of every cell */
static void finalize(level_t *L, const oracle_params *P)
{
    const double U = L->qu;
    for (int jy = 0; jy < L->njy; ++jy) for (int jx = 0; jx < L->njx; ++jx) {
        size_t c = (size_t)jy * (size_t)L->njx + (size_t)jx;
        float *rec = L->cells + 8 * c;
        memset(rec, 0, 8 * sizeof(float));
        uint32_t n = L->n[c];
        if (n < (uint32_t)P->min_points) continue;
        const int64_t *s = L->s + 5 * c;
        double N = (double)n;
        double mx = (double)s[0] / N, my = (double)s[1] / N;
        double cxx = ((double)s[2] - (double)s[0] * mx) / (N - 1.0);
        double cxy = ((double)s[3] - (double)s[0] * my) / (N - 1.0);
        double cyy = ((double)s[4] - (double)s[1] * my) / (N - 1.0);
        mx *= U; my *= U; cxx *= U * U; cxy *= U * U; cyy *= U * U;
        double tr = cxx + cyy, hd = 0.5 * (cxx - cyy), rad = sqrt(hd * hd + cxy * cxy);
        double l1 = 0.5 * tr + rad, l2 = 0.5 * tr - rad;
        if (!(l1 > 1e-10)) continue;
        if (l2 < P->eig_ratio * l1) {
            double l2n = P->eig_ratio * l1, vx, vy;
            if (hd >= 0.0) { vx = hd + rad; vy = cxy; } else { vx = cxy; vy = rad - hd; }
            double nn = vx * vx + vy * vy, dl = l1 - l2n;
            cxx = l2n + dl * (vx * vx) / nn;
            cxy = dl * (vx * vy) / nn;
            cyy = l2n + dl * (vy * vy) / nn;
        }
        double det = cxx * cyy - cxy * cxy;
        rec[0] = (float)mx; rec[1] = (float)my;   /* the mean RELATIVE TO THE CELL CENTRE (v4) */
        rec[2] = (float)(cyy / det); rec[3] = (float)(-(cxy / det));
        rec[4] = rec[3]; rec[5] = (float)(cxx / det);
        rec[6] = (float)n; rec[7] = 1.0f;
    }
}

int oracle_set_target(oracle_matcher *m, const float *xy, int n)
{
    drop_target(m);
    for (int l = 0; l < m->nlevels; ++l) {
        int rc = level_setup(m, l, xy, n);
        if (rc) { drop_target(m); return rc; }
        accumulate(&m->lv[l], xy, n);
        finalize(&m->lv[l], &m->prm);
    }
    m->has_target = 1;
    return 0;
}

/* SPEC 7 */
int oracle_add_target(oracle_matcher *m, const float *xy, int n)
{
    if (!m->has_target) return oracle_set_target(m, xy, n);
    for (int l = 0; l < m->nlevels; ++l) {
        accumulate(&m->lv[l], xy, n);
        finalize(&m->lv[l], &m->prm);
    }
    return 0;
}

static const level_t *get_level(const oracle_matcher *m, int level)
{
    if (!m->has_target || level < 0 || level >= m->nlevels) return NULL;
    return &m->lv[level];
}

int oracle_level_geometry(const oracle_matcher *m, int level, float out[5], int32_t dims[4])
{
    const level_t *L = get_level(m, level);
    if (!L) return 1;
    out[0] = L->res; out[1] = L->st; out[2] = L->inv_st; out[3] = L->ox; out[4] = L->oy;
    dims[0] = L->nhx; dims[1] = L->nhy; dims[2] = L->njx; dims[3] = L->njy;
    return 0;
}

int oracle_get_cells(const oracle_matcher *m, int level, float *cells)
{
    const level_t *L = get_level(m, level);
    if (!L) return 1;
    memcpy(cells, L->cells, (size_t)L->njx * L->njy * 8 * sizeof(float));
    return 0;
}

int oracle_get_sums(const oracle_matcher *m, int level, uint32_t *n, int64_t *sums)
{
    const level_t *L = get_level(m, level);
    if (!L) return 1;
    size_t nc = (size_t)L->njx * L->njy;
    memcpy(n, L->n, nc * sizeof(uint32_t));
    memcpy(sums, L->s, nc * 5 * sizeof(int64_t));
    return 0;
}

/* SPEC 4.2: sin and cos of the pose angle as a fixed sequence of f64 operations (two-constant Cody-Waite
 * reduction to [-pi/4, pi/4], the fdlibm kernel polynomials, quadrant from the low bits of the magic sum), so
 * that every implementation gets the same bits; absolute error below 2e-16 for |theta| < 1e5. */
void oracle_sincos(double th, double *sn_out, double *cs_out)
{
    const double MAGIC = 6755399441055744.0;                       /* 1.5 * 2^52 */
    double t = fma(th, 0.63661977236758138, MAGIC);                /* round(th * 2/pi) in the low mantissa bits */
    double k = t - MAGIC;
    uint64_t tb;
    memcpy(&tb, &t, 8);
    int q = (int)(tb & 3u);
    double r = fma(-k, 1.5707963267948966, th);                    /* pi/2 = 1.5707963267948966 + 6.123233995736766e-17 */
    r = fma(-k, 6.123233995736766e-17, r);
    double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    double sn = fma(r * z, ps, r);                                 /* r + r^3 P(z) */
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    double cs = fma(z * z, pc, fma(z, -0.5, 1.0));                 /* 1 - z/2 + z^2 Q(z) */
    double s4 = (q & 1) ? cs : sn, c4 = (q & 1) ? sn : cs;         /* quadrant: q = 0 (s,c) 1 (c,-s) 2 (-s,-c) 3 (-c,s) */
    *sn_out = (q & 2) ? -s4 : s4;
    *cs_out = ((q + 1) & 2) ? -c4 : c4;
}

/* SPEC 4 (v4): the pose as the evaluation uses it. The point-to-cell geometry is f64: (c, s) from SPEC 4.2 and the
 * translation relative to the level's lattice origin; the derivative terms use the f32 roundings (cf, sf). */
typedef struct { double ci, si, txi, tyi; float cf, sf; } pose_t;
static pose_t pose_for_level(const double p[3], const level_t *L)
{
    pose_t q;
    double sn, cs;
    oracle_sincos(p[2], &sn, &cs);
    q.cf = (float)cs; q.sf = (float)sn;
    q.ci = cs * L->inv_std; q.si = sn * L->inv_std;              /* rotation and translation in cell units */
    q.txi = (p[0] - (double)L->ox) * L->inv_std; q.tyi = (p[1] - (double)L->oy) * L->inv_std;
    return q;
}

/* SPEC 4: the per-point quantities shared by its K cells: r = R x and j = dr/dtheta in f32 (derivative terms),
 * the lattice square (hx, hy) and the position (dfx, dfy) in [0, 1] inside it, from the f64 transform */
typedef struct { float rx, ry, jx, jy, dfx, dfy; int hx, hy, inside; } point_t;

static point_t transform_point(const pose_t *q, const level_t *L, float x, float y)
{
    /* SPEC 4: points that cannot lie in any lattice are replaced by a finite far-away point */
    if (!isfinite(x) || !isfinite(y) || fabsf(x) > 1e18f || fabsf(y) > 1e18f) { x = 1e18f; y = 1e18f; }
    const float nc = -q->cf, ns = -q->sf;
    point_t p;
    p.rx = fmaf(q->cf, x, ns * y); p.ry = fmaf(q->sf, x, q->cf * y);
    p.jx = fmaf(ns, x, nc * y);    p.jy = fmaf(q->cf, x, ns * y);
    const double xd = (double)x, yd = (double)y;
    const double fx = fma(q->ci, xd, fma(-q->si, yd, q->txi));
    const double fy = fma(q->si, xd, fma(q->ci, yd, q->tyi));
    p.hx = p.hy = 0; p.dfx = p.dfy = 0.0f;
    p.inside = lattice(L, fx, fy, &p.hx, &p.hy, &p.dfx, &p.dfy);
    return p;
}

int oracle_cell_index(const oracle_matcher *m, int level, const float *xy, int n,
                      const double *pose, int32_t *idx)
{
    const level_t *L = get_level(m, level);
    if (!L) return 1;
    pose_t q;
    if (pose) q = pose_for_level(pose, L);
    for (int i = 0; i < n; ++i) {
        float x = xy[2 * i], y = xy[2 * i + 1];
        int hx, hy, in;
        if (pose) {
            point_t p = transform_point(&q, L, x, y);
            in = p.inside; hx = p.hx; hy = p.hy;
        } else {
            in = lattice_of_point(L, x, y, &hx, &hy);
        }
        idx[i] = in ? hy * L->nhx + hx : -1;
    }
    return 0;
}

/* SPEC 4.1 */
float oracle_expneg(float h)
{
    float t = fmaf(h, 1.44269502f, 12582912.0f);
    float nf = t - 12582912.0f;
    int32_t ni = (int32_t)(f2u(t) - 0x4B400000u);
    float r = fmaf(nf, -0.693145752f, h);
    r = fmaf(nf, -1.42860677e-6f, r);
    float y = -r;
    float p = 1.38888889e-3f;
    p = fmaf(p, y, 8.33333333e-3f);
    p = fmaf(p, y, 4.16666667e-2f);
    p = fmaf(p, y, 1.66666667e-1f);
    p = fmaf(p, y, 0.5f);
    p = fmaf(p, y, 1.0f);
    p = fmaf(p, y, 1.0f);
    return p * u2f((uint32_t)(0x3F800000 - ni * 0x800000));
}

/* SPEC 4: the ten f32 factors (e, c1..c9) of one (point, cell) pair; returns 0 when skipped */
/* (lx, ly): the point relative to the centre of this cell (SPEC 4 v4), f32 */
static int pair_terms(const float *rec, const point_t *p, float lx, float ly, float T[10])
{
    if (rec[7] == 0.0f) return 0;
    float mux = rec[0], muy = rec[1], B00 = rec[2], B01 = rec[3], B11 = rec[5];
    float qx = lx - mux, qy = ly - muy;
    float ux = fmaf(B00, qx, B01 * qy), uy = fmaf(B01, qx, B11 * qy);
    float m1 = qx * ux, m2 = qy * uy;
    float mm = m1 + m2, h = 0.5f * mm;
    if (!(h < 30.0f)) return 0;
    float e = oracle_expneg(h);
    float a1 = ux * p->jx, a1b = uy * p->jy;
    float a2 = a1 + a1b;
    float vx = fmaf(B00, p->jx, B01 * p->jy), vy = fmaf(B01, p->jx, B11 * p->jy);
    float w1 = ux * p->rx, w2 = uy * p->ry;
    float w = w1 + w2;
    float k1 = p->jx * vx, k2 = p->jy * vy;
    float k = k1 + k2;
    k = k - w;
    k = fmaf(-a2, a2, k);
    /* T = (e, c1..c9): the factors SPEC 4 accumulates with acc0 += e, acc_t = fma(e, c_t, acc_t) */
    T[0] = e;
    T[1] = ux; T[2] = uy; T[3] = a2;
    T[4] = fmaf(-ux, ux, B00); T[5] = fmaf(-ux, uy, B01); T[6] = fmaf(-a2, ux, vx);
    T[7] = fmaf(-uy, uy, B11); T[8] = fmaf(-a2, uy, vy); T[9] = k;
    return 1;
}

/* SPEC 4: evaluate with the fixed summation order (64 f32 partials, f64 butterfly); terms_out (optional)
 * gets n*K*10 f32 */
static void evaluate_level(const level_t *L, const float *xy, int n, const double pose[3],
                           double out[10], int32_t *count, float *terms_out)
{
    pose_t q = pose_for_level(pose, L);
    const int K = L->ov ? 2 : 1;
    const float h2 = 0.5f * L->st;
    float part[64][10];
    memset(part, 0, sizeof(part));
    int32_t cnt = 0;
    if (terms_out) memset(terms_out, 0, (size_t)n * K * K * 10 * sizeof(float));
    for (int i = 0; i < n; ++i) {
        point_t p = transform_point(&q, L, xy[2 * i], xy[2 * i + 1]);
        if (!p.inside) continue;
        const int hx = p.hx, hy = p.hy;
        float *acc = part[i & 63];
        for (int b = 0; b < K; ++b) for (int a = 0; a < K; ++a) {
            const float *rec = L->cells + 8 * ((size_t)(hy + b) * L->njx + (size_t)(hx + a));
            /* the cell's centre relative to node (hx, hy): (st/2, st/2) for one grid; node (hx+a, hy+b) itself for the
             * four half-shifted grids */
            const float lx = fmaf(p.dfx, L->st, L->ov ? -((float)a * L->st) : -h2);
            const float ly = fmaf(p.dfy, L->st, L->ov ? -((float)b * L->st) : -h2);
            float T[10];
            if (!pair_terms(rec, &p, lx, ly, T)) continue;
            acc[0] = acc[0] + T[0];
            for (int t = 1; t < 10; ++t) acc[t] = fmaf(T[0], T[t], acc[t]);
            if (terms_out) memcpy(terms_out + ((size_t)i * K * K + (size_t)(b * K + a)) * 10, T, sizeof(T));
            cnt += 1;
        }
    }
    for (int t = 0; t < 10; ++t) {
        double D[32], E[32];
        for (int l = 0; l < 32; ++l) D[l] = (double)part[l][t] + (double)part[l + 32][t];
        for (int o = 16; o > 0; o >>= 1) {
            for (int l = 0; l < 32; ++l) E[l] = D[l] + D[l ^ o];
            memcpy(D, E, sizeof(D));
        }
        out[t] = D[0];
    }
    *count = cnt;
}

int oracle_evaluate(const oracle_matcher *m, int level, const float *xy, int n,
                    const double pose[3], double out10[10], int32_t *count)
{
    const level_t *L = get_level(m, level);
    if (!L) return 1;
    evaluate_level(L, xy, n, pose, out10, count, NULL);
    return 0;
}

int oracle_point_terms(const oracle_matcher *m, int level, const float *xy, int n,
                       const double pose[3], float *terms)
{
    const level_t *L = get_level(m, level);
    if (!L) return 1;
    double out[10]; int32_t cnt;
    evaluate_level(L, xy, n, pose, out, &cnt, terms);
    return 0;
}

/* SPEC 5: damped 3x3 solve in closed form (adjugate / determinant), positive definite by Sylvester's
 * criterion. H6 = {H00,H01,H02,H11,H12,H22}. Returns 1 when solved. Every a*b - c*d below is two rounded
 * products and one rounded difference (the file is compiled with -ffp-contract=off). */
int oracle_solve(const double g[3], const double H6[6], double lambda, double d[3])
{
    double A00 = H6[0] + lambda * fmax(fabs(H6[0]), 1e-9);
    double A11 = H6[3] + lambda * fmax(fabs(H6[3]), 1e-9);
    double A22 = H6[5] + lambda * fmax(fabs(H6[5]), 1e-9);
    double A01 = H6[1], A02 = H6[2], A12 = H6[4];
    double C00 = A11 * A22 - A12 * A12;
    double C01 = A02 * A12 - A01 * A22;
    double C02 = A01 * A12 - A02 * A11;
    double C11 = A00 * A22 - A02 * A02;
    double C12 = A01 * A02 - A00 * A12;
    double C22 = A00 * A11 - A01 * A01;
    double det = (A00 * C00 + A01 * C01) + A02 * C02;
    if (!(A00 > 0.0) || !(C22 > 0.0) || !(det > 0.0)) return 0;
    double r = 1.0 / det;
    d[0] = -(((C00 * g[0] + C01 * g[1]) + C02 * g[2]) * r);
    d[1] = -(((C01 * g[0] + C11 * g[1]) + C12 * g[2]) * r);
    d[2] = -(((C02 * g[0] + C12 * g[1]) + C22 * g[2]) * r);
    return 1;
}

typedef struct { double v[10]; int32_t count; } eval_t;

/* SPEC 5: one pyramid level of Levenberg-Marquardt; updates p, E, returns status */
static int align_level(const level_t *L, const oracle_params *P, const float *xy, int n,
                       double p[3], eval_t *E, int *evals_total)
{
    double lambda = P->lambda_init;
    evaluate_level(L, xy, n, p, E->v, &E->count, NULL);
    int evals = 1, status = 1;
    if (n == 0 || E->count == 0) { *evals_total += evals; return 3; }
    for (;;) {
        if (evals >= P->max_iterations) break;
        double d[3];
        int stalled = 0;
        while (!oracle_solve(&E->v[1], &E->v[4], lambda, d)) {
            lambda = lambda * P->lambda_fail_up;
            if (lambda > P->lambda_max) { stalled = 1; break; }
        }
        if (stalled) { status = 2; break; }
        double n2 = d[0] * d[0] + d[1] * d[1];      /* squared translation step: no sqrt unless the step is clamped */
        if (n2 > P->max_step_trans * P->max_step_trans) {
            double sc = P->max_step_trans / sqrt(n2);
            d[0] *= sc; d[1] *= sc; d[2] *= sc; n2 = P->max_step_trans * P->max_step_trans;
        }
        if (fabs(d[2]) > P->max_step_rot) {
            double sc = P->max_step_rot / fabs(d[2]);
            d[0] *= sc; d[1] *= sc; d[2] *= sc; n2 = n2 * (sc * sc);
        }
        int small = (n2 < P->eps_trans * P->eps_trans) && (fabs(d[2]) < P->eps_rot);
        double pn[3] = {p[0] + d[0], p[1] + d[1], p[2] + d[2]};
        eval_t En;
        evaluate_level(L, xy, n, pn, En.v, &En.count, NULL);
        evals += 1;
        if (En.v[0] > E->v[0]) {
            p[0] = pn[0]; p[1] = pn[1]; p[2] = pn[2];
            *E = En;
            lambda = fmax(lambda / P->lambda_down, P->lambda_min);
            if (small) { status = 0; break; }
        } else {
            if (small) { status = 0; break; }
            lambda = lambda * P->lambda_up;
            if (lambda > P->lambda_max) { status = 2; break; }
        }
    }
    *evals_total += evals;
    return status;
}

int oracle_align(const oracle_matcher *m, const float *xy, int n, const double init[3],
                 oracle_result *res)
{
    if (!m->has_target) return 1;
    double p[3] = {init[0], init[1], init[2]};
    eval_t E; memset(&E, 0, sizeof(E));
    int evals = 0, status = 3;
    for (int l = 0; l < m->nlevels; ++l)
        status = align_level(&m->lv[l], &m->prm, xy, n, p, &E, &evals);
    const double TWO_PI = 6.283185307179586476925286766559;
    res->pose[0] = p[0]; res->pose[1] = p[1];
    res->pose[2] = p[2] - TWO_PI * rint(p[2] / TWO_PI);
    res->score = E.v[0];
    res->grad[0] = E.v[1]; res->grad[1] = E.v[2]; res->grad[2] = E.v[3];
    res->hessian[0] = E.v[4]; res->hessian[1] = E.v[5]; res->hessian[2] = E.v[6];
    res->hessian[3] = E.v[5]; res->hessian[4] = E.v[7]; res->hessian[5] = E.v[8];
    res->hessian[6] = E.v[6]; res->hessian[7] = E.v[8]; res->hessian[8] = E.v[9];
    res->iterations = evals; res->status = status; res->count = E.count; res->reserved = 0;
    return 0;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int oracle_align_batch(const oracle_matcher *m, const float *xy, const int64_t *offsets, int nscans,
                       const double *init, oracle_result *res, int nthreads)
{
    if (!m->has_target) return 1;
    if (nthreads <= 0) nthreads = oracle_num_threads();
    #pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
    for (int b = 0; b < nscans; ++b)
        oracle_align(m, xy + 2 * offsets[b], (int)(offsets[b + 1] - offsets[b]), init + 3 * b, res + b);
    return 0;
}

/* SPEC 6 */
int oracle_sweep(const oracle_matcher *m, int level, const float *xy, int n, const float *hyp,
                 int64_t nhyp, double *scores, int64_t *best_idx, double *best_score, int nthreads)
{
    const level_t *L = get_level(m, level);
    if (!L) return 1;
    if (nthreads <= 0) nthreads = oracle_num_threads();
    int64_t bi = -1; double bs = -1.0;
    #pragma omp parallel num_threads(nthreads)
    {
        int64_t lbi = -1; double lbs = -1.0;
        #pragma omp for schedule(static)
        for (int64_t j = 0; j < nhyp; ++j) {
            double pose[3] = {(double)hyp[3 * j], (double)hyp[3 * j + 1], (double)hyp[3 * j + 2]};
            double out[10]; int32_t cnt;
            evaluate_level(L, xy, n, pose, out, &cnt, NULL);
            if (scores) scores[j] = out[0];
            if (out[0] > lbs) { lbs = out[0]; lbi = j; } /* static schedule: j ascending per thread */
        }
        #pragma omp critical
        {
            if (lbi >= 0 && (lbs > bs || (lbs == bs && lbi < bi))) { bs = lbs; bi = lbi; }
        }
    }
    if (best_idx) *best_idx = bi;
    if (best_score) *best_score = bs;
    return 0;
}

/* SPEC 8 */
int oracle_polar_to_points(const float *ranges_f32, const uint16_t *ranges_u16, int nbeams,
                           double angle_min, double angle_inc, float range_scale,
                           float range_min, float range_max, float *xy_out)
{
    int k = 0;
    for (int i = 0; i < nbeams; ++i) {
        float rho;
        if (ranges_u16) {
            if (ranges_u16[i] == 0) continue;
            rho = (float)ranges_u16[i] * range_scale;
        } else {
            rho = ranges_f32[i];
        }
        if (!((rho >= range_min) && (rho <= range_max))) continue;
        double phi = angle_min + (double)i * angle_inc;
        float cb = (float)cos(phi), sb = (float)sin(phi);
        xy_out[2 * k] = rho * cb; xy_out[2 * k + 1] = rho * sb;
        ++k;
    }
    return k;
}
