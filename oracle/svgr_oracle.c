/*
 * svgr_oracle.c -- CPU restatement of the svgrasterize.py hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * core: it may be built/loaded by tests/, by __graft_entry__.smoke() and by
 * bench.py's cpu_baseline / --impl reference legs, never by the product path.
 *
 * Every function restates one piece of /root/reference/svgrasterize.py in
 * plain double-precision C and cites the lines it follows.  Parity is pinned:
 * tests/test_oracle_pins.py checks these functions against vectors generated
 * by running the unmodified reference (tools/make_golden.py -> tests/golden/),
 * and, when /root/reference is present, against the live reference.
 *
 * Build: gcc -O2 -fPIC -shared -mfma -ffp-contract=off (see oracle/Makefile).
 * -ffp-contract=off matters: the reference is CPython/numpy double arithmetic
 * with one rounding per operation, except inside the BLAS contractions whose
 * fused-multiply-add order is written out explicitly with fma() below.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------
 * Geometry
 * ------------------------------------------------------------------------- */

/* Transform.__call__ (svgrasterize.py:531-534): P @ M[:2,:2].T + M[:2,2].
 * numpy evaluates the 2-term contraction as fma(y, m[c][1], x*m[c][0]) and then
 * adds the translation with a second rounding.  m6 = row-major 2x3. */
ORC_API void orc_transform_points(const double *m6, const double *pts, long n, double *out)
{
    for (long i = 0; i < n; i++) {
        double x = pts[2 * i], y = pts[2 * i + 1];
        out[2 * i] = fma(y, m6[1], x * m6[0]) + m6[2];
        out[2 * i + 1] = fma(y, m6[4], x * m6[3]) + m6[5];
    }
}

/* bezier2_to_bezier3 (svgrasterize.py:2182-2184 with the matrix of :2052-2054):
 * each output = fma(m2,p2, fma(m1,p1, m0*p0)). */
ORC_API void orc_quad_to_cubic(const double *q, double *c)
{
    static const double M[4][3] = {
        {1, 0, 0}, {1.0 / 3, 2.0 / 3, 0}, {0, 2.0 / 3.0, 1.0 / 3}, {0, 0, 1}};
    for (int r = 0; r < 4; r++)
        for (int k = 0; k < 2; k++)
            c[2 * r + k] = fma(M[r][2], q[4 + k], fma(M[r][1], q[2 + k], M[r][0] * q[k]));
}

/* 4-term BLAS contraction used by np.dot(M, batch) in bezier3_split_batch
 * (:2068) and bezier3_flatness_batch (:2087): strided ddot with two
 * accumulators: fma(m0,b0, m2*b2) + fma(m1,b1, m3*b3). */
static inline double dot4(const double *m, double b0, double b1, double b2, double b3)
{
    return fma(m[0], b0, m[2] * b2) + fma(m[1], b1, m[3] * b3);
}

static const double SPLIT[8][4] = {
    {1, 0, 0, 0},       {0.5, 0.5, 0, 0},   {0.25, 0.5, 0.25, 0}, {0.125, 0.375, 0.375, 0.125},
    {0.125, 0.375, 0.375, 0.125}, {0, 0.25, 0.5, 0.25}, {0, 0, 0.5, 0.5},     {0, 0, 0, 1}};
static const double FLAT[2][4] = {{-2, 3, 0, -1}, {-1, 0, 3, -2}};

/* bezier3_flatness_batch (:2071-2088) for one cubic c[4][2]:
 * uv[row][coord] = square(FLAT[row] . c[:,coord]); result = sum_row max_coord uv. */
ORC_API double orc_cubic_flatness(const double *c)
{
    double s = 0.0;
    double v[2];
    for (int r = 0; r < 2; r++) {
        double a = dot4(FLAT[r], c[0], c[2], c[4], c[6]);
        double b = dot4(FLAT[r], c[1], c[3], c[5], c[7]);
        a = a * a;
        b = b * b;
        v[r] = a > b ? a : b;
    }
    s = v[0] + v[1];
    return s;
}

/* bezier3_split_batch (:2066-2068) for one cubic: out[0..7] = left, out[8..15] = right. */
ORC_API void orc_cubic_split(const double *c, double *out)
{
    for (int r = 0; r < 8; r++)
        for (int k = 0; k < 2; k++)
            out[2 * r + k] = dot4(SPLIT[r], c[k], c[2 + k], c[4 + k], c[6 + k]);
}

/* bezier3_flatten_batch (:2091-2098): breadth-first, level by level; a cubic
 * whose flatness is < 16*tol^2 emits the line [p0, p3], the others are split
 * (children ordered left, right per parent).  Output order is the reference's:
 * all depth-0 lines in input order, then depth 1, ...  Returns the number of
 * lines, or -1 if `cap` lines were not enough, or -2 on allocation failure.
 * `max_depth` bounds the loop (the reference has no bound and would spin on
 * NaN input); pass 0 for "no bound". */
ORC_API long orc_flatten_cubics(const double *cubics, long m, double tol, double *lines, long cap,
                                int max_depth)
{
    double thr = (tol * tol) * 16;
    long n = m, count = 0;
    double *cur = (double *)malloc(sizeof(double) * 8 * (size_t)(n > 0 ? n : 1));
    if (!cur)
        return -2;
    memcpy(cur, cubics, sizeof(double) * 8 * (size_t)n);
    int depth = 0;
    while (n > 0) {
        double *next = (double *)malloc(sizeof(double) * 16 * (size_t)n);
        if (!next) {
            free(cur);
            return -2;
        }
        long nn = 0;
        for (long i = 0; i < n; i++) {
            const double *c = cur + 8 * i;
            if (orc_cubic_flatness(c) < thr) {
                if (count >= cap) {
                    free(cur);
                    free(next);
                    return -1;
                }
                double *l = lines + 4 * count++;
                l[0] = c[0], l[1] = c[1], l[2] = c[6], l[3] = c[7];
            } else {
                orc_cubic_split(c, next + 16 * nn);
                nn++;
            }
        }
        free(cur);
        cur = next;
        n = 2 * nn;
        depth++;
        if (max_depth > 0 && depth > max_depth)
            break;
    }
    free(cur);
    return count;
}

/* arc_to_bezier3 (:2355-2394).  Pieces of at most pi/4; np.linspace gives
 * eta_i = i*step + eta (two roundings) with the last one forced to eta+delta.
 * arc(a) = M @ [rx cos a, ry sin a] + center with the 2-term contraction
 * fma(M[k][0], u, M[k][1]*v).  Returns the number of cubics written (<= cap)
 * or -1. */
static inline void arc_mul(const double M[2][2], double u, double v, double *o)
{
    o[0] = fma(M[0][0], u, M[0][1] * v);
    o[1] = fma(M[1][0], u, M[1][1] * v);
}

ORC_API long orc_arc_to_cubics(double cx, double cy, double rx, double ry, double phi, double eta,
                               double eta_delta, double *out, long cap)
{
    const double M[2][2] = {{cos(phi), -sin(phi)}, {sin(phi), cos(phi)}};
    const double max_angle = M_PI / 4;
    long count = (long)ceil(fabs(eta_delta) / max_angle);
    if (count > cap)
        return -1;
    double stop = eta + eta_delta;
    double step = count > 0 ? (stop - eta) / (double)count : 0.0;
    for (long i = 0; i < count; i++) {
        double e1, e2;
        /* np.linspace: y = arange(num) * step + start, y[-1] = stop; when
         * step == 0 numpy multiplies by delta instead (same value: 0). */
        e1 = (double)i * step + eta;
        e2 = (i + 1 == count) ? stop : (double)(i + 1) * step + eta;
        double t = tan((e2 - e1) / 2);
        double sq = sqrt(4 + 3 * (t * t));
        double alpha = sin(e2 - e1) * (sq - 1) / 3;
        double p0[2], p3[2], d1[2], d2[2];
        arc_mul(M, rx * cos(e1), ry * sin(e1), p0);
        p0[0] += cx, p0[1] += cy;
        arc_mul(M, rx * cos(e2), ry * sin(e2), p3);
        p3[0] += cx, p3[1] += cy;
        arc_mul(M, -rx * sin(e1), ry * cos(e1), d1);
        arc_mul(M, -rx * sin(e2), ry * cos(e2), d2);
        double *o = out + 8 * i;
        o[0] = p0[0], o[1] = p0[1];
        o[2] = p0[0] + alpha * d1[0], o[3] = p0[1] + alpha * d1[1];
        o[4] = p3[0] - alpha * d2[0], o[5] = p3[1] - alpha * d2[1];
        o[6] = p3[0], o[7] = p3[1];
    }
    return count;
}

/* ---------------------------------------------------------------------------
 * Coverage
 * ------------------------------------------------------------------------- */

/* line_signed_coverage (svgrasterize.py:2213-2304).  `trace` is (h rows, w
 * cols) row-major; line = {r0, c0, r1, c1} in mask-local pixels (component 0
 * is the row, :2228-2229).  Scalar double arithmetic, one rounding per
 * operation, the column carried incrementally from row to row (:2248). */
ORC_API void orc_line_coverage(double *trace, long h, long w, const double *line)
{
    double r0 = line[0], c0 = line[1], r1 = line[2], c1 = line[3];
    if (r0 == r1)
        return;
    double dir = 1.0;
    if (!(r0 < r1)) {
        double t;
        dir = -1.0;
        t = r0, r0 = r1, r1 = t;
        t = c0, c0 = c1, c1 = t;
    }
    double dxdy = (c1 - c0) / (r1 - r0);
    double x = c0;
    double ys = r0 > 0 ? r0 : 0.0; /* max(0, p0[Y]) */
    long y = (long)ys;             /* int(): truncation, value is >= 0 */
    if (r0 < 0)
        x -= r0 * dxdy;
    double yend_f = ceil(r1);
    long yend = yend_f < (double)h ? (long)yend_f : h;
    double x_next = x;
    for (; y < yend; y++) {
        x = x_next;
        double ytop = (double)(y + 1) < r1 ? (double)(y + 1) : r1;
        double ybot = (double)y > r0 ? (double)y : r0;
        double dy = ytop - ybot;
        double d = dir * dy;
        x_next = x + dxdy * dy;
        double x0 = x, x1 = x_next;
        if (!(x < x_next))
            x0 = x_next, x1 = x;
        double x0_floor = floor(x0);
        long x0i = (long)x0_floor;
        double x1_ceil = ceil(x1);
        long x1i = (long)x1_ceil;
        double *row = trace + y * w;
#define PUT(ix, val)                                   \
    {                                                  \
        long _i = (ix);                                \
        if (_i >= w)                                   \
            continue;                                  \
        row[_i > 0 ? _i : 0] += (val);                 \
    }
        if (x1i <= x0i + 1) {
            double xmf = 0.5 * (x + x_next) - x0_floor;
            PUT(x0i, d * (1 - xmf));
            PUT(x0i + 1, d * xmf);
        } else {
            double s = 1 / (x1 - x0);
            double x0f = x0 - x0_floor;
            double x1f = x1 - x1_ceil + 1.0;
            double a0 = 0.5 * s * pow(1 - x0f, 2.0); /* python ** 2 -> libm pow */
            double am = 0.5 * s * pow(x1f, 2.0);
            PUT(x0i, d * a0);
            if (x1i == x0i + 2) {
                PUT(x0i + 1, d * (1.0 - a0 - am));
            } else {
                double a1 = s * (1.5 - x0f);
                PUT(x0i + 1, d * (a1 - a0));
                for (long xi = x0i + 2; xi < x1i - 1; xi++) {
                    if (xi >= w) /* the reference's `continue` here only skips this xi */
                        continue;
                    row[xi > 0 ? xi : 0] += d * s;
                }
                double a2 = a1 + (double)(x1i - x0i - 3) * s;
                PUT(x1i - 1, d * (1.0 - a2 - am));
            }
            PUT(x1i, d * am);
        }
#undef PUT
    }
}

/* Path.mask :983-990: inclusive prefix sum along columns, fill rule, snap.
 * rule: 0 = nonzero (also the reference's None), 1 = evenodd. */
ORC_API int orc_mask_finish(double *trace, long h, long w, int rule)
{
    if (rule != 0 && rule != 1)
        return -1;
    for (long y = 0; y < h; y++) {
        double *row = trace + y * w;
        double acc = 0.0;
        for (long x = 0; x < w; x++) {
            acc = x == 0 ? row[0] : acc + row[x];
            double m;
            if (rule == 0) {
                m = fabs(acc);
                if (m > 1.0)
                    m = 1.0;
            } else {
                /* np.remainder(acc + 1, 2): floored modulo */
                double a = acc + 1.0;
                double r = fmod(a, 2.0);
                if (r != 0.0) {
                    if (r < 0.0)
                        r += 2.0;
                } else {
                    r = 0.0;
                }
                m = fabs(r - 1.0);
            }
            if (m < 1e-6)
                m = 0.0;
            row[x] = m;
        }
    }
    return 0;
}

/* Whole mask for an edge list already in presentation space (:961-990).
 * edges = n x {r0,c0,r1,c1}; bbox = {min_r, min_c, rows, cols}. */
ORC_API int orc_mask(const double *edges, long n, const long *bbox, int rule, double *out)
{
    long h = bbox[2], w = bbox[3];
    memset(out, 0, sizeof(double) * (size_t)h * (size_t)w);
    double off_r = (double)bbox[0], off_c = (double)bbox[1];
    for (long i = 0; i < n; i++) {
        double l[4] = {edges[4 * i] - off_r, edges[4 * i + 1] - off_c, edges[4 * i + 2] - off_r,
                       edges[4 * i + 3] - off_c};
        orc_line_coverage(out, h, w, l);
    }
    return orc_mask_finish(out, h, w, rule);
}

/* Mask sizing (:961-975).  viewport may be NULL.  Returns 0 and fills bbox,
 * or 1 when the mask is empty (the reference returns None). */
ORC_API int orc_mask_bounds(const double *edges, long n, const long *viewport, long *bbox)
{
    if (n <= 0)
        return 1;
    double mn_r = INFINITY, mn_c = INFINITY, mx_r = -INFINITY, mx_c = -INFINITY;
    for (long i = 0; i < 2 * n; i++) {
        double r = edges[2 * i], c = edges[2 * i + 1];
        if (r < mn_r) mn_r = r;
        if (r > mx_r) mx_r = r;
        if (c < mn_c) mn_c = c;
        if (c > mx_c) mx_c = c;
    }
    long min_r = (long)floor(mn_r) - 1, min_c = (long)floor(mn_c) - 1;
    long max_r = (long)ceil(mx_r) + 1, max_c = (long)ceil(mx_c) + 1;
    if (viewport) {
        long vx = viewport[0], vy = viewport[1], vw = viewport[2], vh = viewport[3];
        if (min_r < vx) min_r = vx;
        if (min_c < vy) min_c = vy;
        if (max_r > vx + vw) max_r = vx + vw;
        if (max_c > vy + vh) max_c = vy + vh;
    }
    bbox[0] = min_r, bbox[1] = min_c, bbox[2] = max_r - min_r, bbox[3] = max_c - min_c;
    return (bbox[2] <= 0 || bbox[3] <= 0) ? 1 : 0;
}
