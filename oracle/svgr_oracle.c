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

/* ---------------------------------------------------------------------------
 * Paint
 * ------------------------------------------------------------------------- */

/* Pixel centre (row + .5, col + .5) of the layer bbox (grad_pixels, :1653-1658)
 * mapped to user space by m1 (the inverse presentation transform, :1022-1027)
 * and then, when present, by m2 (the inverse gradientTransform, :1558-1559 /
 * :1602-1603).  Both are row-major 2x3 and are applied like Transform.__call__. */
static inline void px_to_user(long r, long c, long off_r, long off_c, const double *m1,
                              const double *m2, double *ux, double *uy)
{
    double x = (double)r + ((double)off_r + 0.5), y = (double)c + ((double)off_c + 0.5);
    double ax = fma(y, m1[1], x * m1[0]) + m1[2];
    double ay = fma(y, m1[4], x * m1[3]) + m1[5];
    if (m2) {
        double bx = fma(ay, m2[1], ax * m2[0]) + m2[2];
        double by = fma(ay, m2[4], ax * m2[3]) + m2[5];
        ax = bx, ay = by;
    }
    *ux = ax, *uy = ay;
}

/* GradLinear.fill (:1553-1563): t = ((p - p0) . v) / (v . v). */
ORC_API void orc_grad_linear_t(long rows, long cols, long off_r, long off_c, const double *m1,
                               const double *m2, const double *p0, const double *p1, double *t)
{
    double vx = p1[0] - p0[0], vy = p1[1] - p0[1];
    double vv = fma(vy, vy, vx * vx);
    for (long r = 0; r < rows; r++)
        for (long c = 0; c < cols; c++) {
            double ux, uy;
            px_to_user(r, c, off_r, off_c, m1, m2, &ux, &uy);
            double dx = ux - p0[0], dy = uy - p0[1];
            t[r * cols + c] = fma(dy, vy, dx * vx) / vv;
        }
}

/* GradRadial.fill (:1577-1650).  has_focal = fcenter or fradius given (:1605).
 * valid[] is written only in the focal branch: 1 where det >= 0.  Returns 1 if
 * any det < 0 was seen (the reference then masks those pixels AND, when
 * fradius != radius, also the pixels with t <= fr/(fr-r), :1627-1644; it does
 * neither when no det < 0 exists anywhere in the layer), else 0. */
ORC_API int orc_grad_radial_t(long rows, long cols, long off_r, long off_c, const double *m1,
                              const double *m2, const double *center, double radius, int has_focal,
                              const double *fcenter, double fradius, double *t, uint8_t *valid)
{
    int any_neg = 0;
    if (!has_focal) {
        for (long r = 0; r < rows; r++)
            for (long c = 0; c < cols; c++) {
                double ux, uy;
                px_to_user(r, c, off_r, off_c, m1, m2, &ux, &uy);
                double ox = (ux - center[0]) / radius, oy = (uy - center[1]) / radius;
                t[r * cols + c] = sqrt(ox * ox + oy * oy);
            }
        return 0;
    }
    double cdx = center[0] - fcenter[0], cdy = center[1] - fcenter[1];
    double rd = radius - fradius;
    double a = (cdx * cdx + cdy * cdy) - rd * rd;
    for (long r = 0; r < rows; r++)
        for (long c = 0; c < cols; c++) {
            double ux, uy;
            px_to_user(r, c, off_r, off_c, m1, m2, &ux, &uy);
            double pdx = ux - fcenter[0], pdy = uy - fcenter[1];
            double b = (pdx * cdx + pdy * cdy) + fradius * rd;
            double cc = (pdx * pdx + pdy * pdy) - fradius * fradius;
            double det = b * b - a * cc;
            long i = r * cols + c;
            if (det < 0) { /* NaN det is treated as valid, like `det < 0` in numpy */
                any_neg = 1;
                valid[i] = 0;
                t[i] = 0.0;
            } else {
                double sq = sqrt(det);
                double t1 = (b + sq) / a, t2 = (b - sq) / a;
                valid[i] = 1;
                /* np.maximum propagates NaN */
                t[i] = (t1 != t1 || t2 != t2) ? NAN : (t1 > t2 ? t1 : t2);
            }
        }
    if (any_neg && fradius != radius) {
        double lim = fradius / (fradius - radius);
        for (long i = 0; i < rows * cols; i++)
            if (valid[i] && !(t[i] > lim))
                valid[i] = 0;
    }
    return any_neg;
}

/* grad_spread (:1661-1668) + grad_interpolate (:1671-1683).  spread: 0 pad,
 * 1 repeat (np.modf fractional part: keeps the sign), 2 reflect.  Stops are
 * already in the target colour space (grad_stops_colorspace, :1686-1695),
 * premultiplied, sorted by offset.  valid may be NULL; where valid[i]==0 the
 * pixel is transparent black (:1647-1648). */
ORC_API void orc_grad_colors(long n, const double *t, const uint8_t *valid, int spread, long nstops,
                             const double *offs, const double *cols, double *rgba)
{
    for (long i = 0; i < n; i++) {
        double *o = rgba + 4 * i;
        o[0] = o[1] = o[2] = o[3] = 0.0;
        if (valid && !valid[i])
            continue;
        double v = t[i];
        if (spread == 1) {
            double ip;
            v = modf(v, &ip);
        } else if (spread == 2) {
            double a = v + 1.0;
            double r = fmod(a, 2.0);
            if (r != 0.0) {
                if (r < 0.0)
                    r += 2.0;
            } else {
                r = 0.0;
            }
            v = fabs(r - 1.0);
        }
        if (v <= offs[0]) {
            memcpy(o, cols, 4 * sizeof(double));
            continue;
        }
        if (v > offs[nstops - 1]) {
            memcpy(o, cols + 4 * (nstops - 1), 4 * sizeof(double));
            continue;
        }
        for (long k = 0; k + 1 < nstops; k++) {
            double o0 = offs[k], o1 = offs[k + 1];
            if (v > o0 && v <= o1) {
                double ratio = (v - o0) / (o1 - o0);
                for (int ch = 0; ch < 4; ch++)
                    o[ch] += (1 - ratio) * cols[4 * k + ch] + ratio * cols[4 * (k + 1) + ch];
            }
        }
    }
}

/* ---------------------------------------------------------------------------
 * Filters
 * ------------------------------------------------------------------------- */

/* scipy.signal.convolve(image, kernel[..., None]) in mode "full"
 * (Layer.convolve, :106-115): out[i, j, ch] = sum_{a,b} k[a,b] * in[i-a, j-b, ch],
 * out is (rows+kr-1, cols+kc-1, ch). */
ORC_API void orc_convolve_full(const double *in, long rows, long cols, long ch, const double *k,
                               long kr, long kc, double *out)
{
    long orows = rows + kr - 1, ocols = cols + kc - 1;
    memset(out, 0, sizeof(double) * (size_t)(orows * ocols * ch));
    for (long a = 0; a < kr; a++)
        for (long b = 0; b < kc; b++) {
            double w = k[a * kc + b];
            if (w == 0.0)
                continue;
            for (long i = 0; i < rows; i++) {
                const double *src = in + (i * cols) * ch;
                double *dst = out + ((i + a) * ocols + b) * ch;
                for (long j = 0; j < cols * ch; j++)
                    dst[j] += w * src[j];
            }
        }
}

/* pooling(mat, ksize=(k0,k1), stride=(1,1), method) (:419-468) with no padding:
 * out is (rows-k0+1, cols-k1+1, ch), window anchored at its top-left corner.
 * NaNs are ignored like np.nanmax / np.nanmin. */
ORC_API void orc_pool(const double *in, long rows, long cols, long ch, long k0, long k1, int is_max,
                      double *out)
{
    long orows = rows - k0 + 1, ocols = cols - k1 + 1;
    if (orows <= 0 || ocols <= 0)
        return;
    /* separable: rows of the window first into tmp, then columns */
    double *tmp = (double *)malloc(sizeof(double) * (size_t)(orows * cols * ch));
    for (long i = 0; i < orows; i++)
        for (long j = 0; j < cols * ch; j++) {
            double acc = NAN;
            for (long a = 0; a < k0; a++) {
                double v = in[((i + a) * cols) * ch + j];
                if (v != v)
                    continue;
                if (acc != acc || (is_max ? v > acc : v < acc))
                    acc = v;
            }
            tmp[(i * cols) * ch + j] = acc;
        }
    for (long i = 0; i < orows; i++)
        for (long j = 0; j < ocols; j++)
            for (long c = 0; c < ch; c++) {
                double acc = NAN;
                for (long b = 0; b < k1; b++) {
                    double v = tmp[(i * cols + j + b) * ch + c];
                    if (v != v)
                        continue;
                    if (acc != acc || (is_max ? v > acc : v < acc))
                        acc = v;
                }
                out[(i * ocols + j) * ch + c] = acc;
            }
    free(tmp);
}

/* ---------------------------------------------------------------------------
 * Stroker (Path.stroke, svgrasterize.py:1105-1180 and helpers)
 *
 * Arithmetic notes (probed against numpy 2.3 / OpenBLAS on x86-64 with FMA):
 *   np.dot(u, v) for 2-vectors          = fma(u1, v1, u0*v0)
 *   np.linalg.norm(u)                   = sqrt(fma(u1, u1, u0*u0))
 *   np.cross(u, v) for 2-vectors        = u0*v1 - u1*v0 (no FMA)
 *   [1/8,3/8,3/8,1/8] @ curve           = fma(w0,p0, w1*p1) + fma(w2,p2, w3*p3)
 *   curve.sum(0)                        = ((p0 + p1) + p2) + p3
 *   np.matmul(BEZIER3_SPLIT, curve)     = fma(m3,p3, fma(m2,p2, fma(m1,p1, m0*p0)))
 *   line_offset / line_intersect        = scalar arithmetic, one rounding per op
 * ------------------------------------------------------------------------- */
typedef struct {
    int n;       /* 2 line, 3 quad, 4 cubic */
    double p[8]; /* n points (x, y) */
} Curve;

typedef struct {
    Curve *v;
    long n, cap;
} CurveVec;

static int cv_push(CurveVec *cv, const Curve *c)
{
    if (cv->n == cv->cap) {
        long ncap = cv->cap ? cv->cap * 2 : 64;
        Curve *nv = (Curve *)realloc(cv->v, sizeof(Curve) * (size_t)ncap);
        if (!nv)
            return -1;
        cv->v = nv, cv->cap = ncap;
    }
    cv->v[cv->n++] = *c;
    return 0;
}

static Curve mk_line(const double *a, const double *b)
{
    Curve c;
    memset(&c, 0, sizeof c);
    c.n = 2;
    c.p[0] = a[0], c.p[1] = a[1], c.p[2] = b[0], c.p[3] = b[1];
    return c;
}

/* np.allclose(a, b) on 2-vectors: |a-b| <= atol + rtol*|b|, rtol 1e-5, atol 1e-8 */
static int close1(double a, double b)
{
    if (isfinite(a) && isfinite(b))
        return fabs(a - b) <= 1e-8 + 1e-5 * fabs(b);
    return a == b;
}
static int close2(const double *a, const double *b) { return close1(a[0], b[0]) && close1(a[1], b[1]); }

/* line_offset (:2328-2337); returns 0 when the line is degenerate (None) */
static int line_offset(const double *l, double distance, double *out)
{
    double x1 = l[0], y1 = l[1], x2 = l[2], y2 = l[3];
    double vx = x2 - x1, vy = y2 - y1;
    double len = vx * vx + vy * vy;
    if (len < 2.220446049250313e-16)
        return 0;
    len = sqrt(len);
    double dx = -vy * distance / len;
    double dy = vx * distance / len;
    out[0] = x1 + dx, out[1] = y1 + dy, out[2] = x2 + dx, out[3] = y2 + dy;
    return 1;
}

/* line_intersect (:2307-2325); returns 0 for (None, 0, 0) */
static int line_intersect(const double *l0, const double *l1, double *p, double *t0o, double *t1o)
{
    double x1 = l0[0], y1 = l0[1], x2 = l0[2], y2 = l0[3];
    double x3 = l1[0], y3 = l1[1], x4 = l1[2], y4 = l1[3];
    double det = (x4 - x3) * (y1 - y2) - (x1 - x2) * (y4 - y3);
    if (fabs(det) < 2.220446049250313e-16) {
        *t0o = 0, *t1o = 0;
        return 0;
    }
    double t0 = ((y3 - y4) * (x1 - x3) + (x4 - x3) * (y1 - y3)) / det;
    double t1 = ((y1 - y2) * (x1 - x3) + (x2 - x1) * (y1 - y3)) / det;
    p[0] = x1 * (1 - t0) + x2 * t0;
    p[1] = y1 * (1 - t0) + y2 * t0;
    *t0o = t0, *t1o = t1;
    return 1;
}

#define CIRCLE_K (4 * (M_SQRT2_PY - 1) / 3)
/* math.sqrt(2) */
#define M_SQRT2_PY 1.4142135623730951

/* stroke_line_cap (:1466-1492). cap: 0 butt, 1 round, 2 square. */
static int stroke_cap(const double *p0, const double *p1, int cap, CurveVec *out)
{
    if (close2(p0, p1))
        return 0;
    if (cap == 0) {
        Curve c = mk_line(p0, p1);
        return cv_push(out, &c);
    }
    if (cap == 1) {
        double seg[2] = {p1[0] - p0[0], p1[1] - p0[1]};
        double radius = sqrt(fma(seg[1], seg[1], seg[0] * seg[0])) / 2;
        double two_r = 2 * radius;
        seg[0] /= two_r, seg[1] /= two_r;
        double nrm[2] = {-seg[1], seg[0]};
        double offset = CIRCLE_K * radius;
        double center[2] = {(p0[0] + p1[0]) / 2, (p0[1] + p1[1]) / 2};
        double mid[2] = {center[0] + nrm[0] * radius, center[1] + nrm[1] * radius};
        Curve a, b;
        memset(&a, 0, sizeof a);
        memset(&b, 0, sizeof b);
        a.n = b.n = 4;
        for (int k = 0; k < 2; k++) {
            a.p[0 + k] = p0[k];
            a.p[2 + k] = p0[k] + nrm[k] * offset;
            a.p[4 + k] = mid[k] - seg[k] * offset;
            a.p[6 + k] = mid[k];
            b.p[0 + k] = mid[k];
            b.p[2 + k] = mid[k] + seg[k] * offset;
            b.p[4 + k] = p1[k] + nrm[k] * offset;
            b.p[6 + k] = p1[k];
        }
        if (cv_push(out, &a))
            return -1;
        return cv_push(out, &b);
    }
    if (cap == 2) {
        double seg[2] = {p1[0] - p0[0], p1[1] - p0[1]};
        double nrm[2] = {-seg[1], seg[0]};
        double q0[2] = {p0[0] + nrm[0] / 2, p0[1] + nrm[1] / 2};
        double q1[2] = {p1[0] + nrm[0] / 2, p1[1] + nrm[1] / 2};
        Curve c = mk_line(p0, q0);
        if (cv_push(out, &c))
            return -1;
        c = mk_line(q0, q1);
        if (cv_push(out, &c))
            return -1;
        c = mk_line(q1, p1);
        return cv_push(out, &c);
    }
    return -3;
}

/* stroke_curve_tangent (:1521-1530): first and last non-degenerate control
 * polygon legs; returns 0 when there is none. */
static int curve_tangents(const Curve *c, double *first, double *last)
{
    int found = 0;
    for (int i = 0; i + 1 < c->n; i++) {
        const double *a = c->p + 2 * i, *b = c->p + 2 * i + 2;
        if (close2(a, b))
            continue;
        if (!found)
            memcpy(first, a, 4 * sizeof(double));
        memcpy(last, a, 4 * sizeof(double));
        found = 1;
    }
    return found;
}

/* stroke_line_join (:1495-1518). join: 0 miter, 1 round, 2 bevel, 3 anything else. */
static int stroke_join(const Curve *c0, const Curve *c1, int join, CurveVec *out)
{
    const double *e0 = c0->p + 2 * (c0->n - 1), *s1 = c1->p;
    Curve bevel = mk_line(e0, s1);
    if (join == 2)
        return cv_push(out, &bevel);
    double f0[4], l0[4], f1[4], l1[4];
    int ok0 = curve_tangents(c0, f0, l0);
    int ok1 = curve_tangents(c1, f1, l1);
    if (!ok0 || !ok1)
        return cv_push(out, &bevel);
    if (close2(l0 + 2, f1))
        return 0;
    double p[2], t0, t1;
    int hit = line_intersect(l0, f1, p, &t0, &t1);
    if (!hit || (0 <= t0 && t0 <= 1 && 0 <= t1 && t1 <= 1))
        return cv_push(out, &bevel);
    if (fabs(t0) < 4 && fabs(t1) < 4) {
        if (join == 0) {
            Curve a = mk_line(e0, p), b = mk_line(p, s1);
            if (cv_push(out, &a))
                return -1;
            return cv_push(out, &b);
        }
        if (join == 1) {
            Curve q;
            memset(&q, 0, sizeof q);
            q.n = 3;
            q.p[0] = e0[0], q.p[1] = e0[1], q.p[2] = p[0], q.p[3] = p[1], q.p[4] = s1[0], q.p[5] = s1[1];
            return cv_push(out, &q);
        }
    }
    return cv_push(out, &bevel);
}

/* should_split of bezier3_offset (:2121-2137) */
static int offset_should_split(const double *c)
{
    double bx = c[6] - c[0], by = c[7] - c[1];   /* c3 - c0 */
    double mx = c[4] - c[2], my = c[5] - c[3];   /* c2 - c1 */
    if (fma(by, my, bx * mx) < 0)
        return 1;
    double u1x = c[2] - c[0], u1y = c[3] - c[1]; /* c1 - c0 */
    double u2x = c[4] - c[0], u2y = c[5] - c[1]; /* c2 - c0 */
    double a0 = bx * u1y - by * u1x;
    double a1 = bx * u2y - by * u2x;
    if (a0 * a1 < 0)
        return 1;
    double dist = 0, diag = 0, dd[2], gg[2];
    for (int k = 0; k < 2; k++) {
        double p0 = c[k], p1 = c[2 + k], p2 = c[4 + k], p3 = c[6 + k];
        double mass = (((p0 + p1) + p2) + p3) / 4;
        double mid = fma(0.125, p0, 0.375 * p1) + fma(0.375, p2, 0.125 * p3);
        double mx_ = p0, mn_ = p0;
        if (p1 > mx_) mx_ = p1;
        if (p2 > mx_) mx_ = p2;
        if (p3 > mx_) mx_ = p3;
        if (p1 < mn_) mn_ = p1;
        if (p2 < mn_) mn_ = p2;
        if (p3 < mn_) mn_ = p3;
        double d = mass - mid, g = mx_ - mn_;
        dd[k] = d * d, gg[k] = g * g;
    }
    dist = dd[0] + dd[1];
    diag = gg[0] + gg[1];
    return dist * 100 > diag;
}

/* bezier3_split (:2058-2063): np.matmul(BEZIER3_SPLIT, points) */
static void split_single(const double *c, double *out)
{
    for (int r = 0; r < 8; r++)
        for (int k = 0; k < 2; k++) {
            const double *m = SPLIT[r];
            out[2 * r + k] = fma(m[3], c[6 + k], fma(m[2], c[4 + k], fma(m[1], c[2 + k], m[0] * c[k])));
        }
}

/* bezier3_offset (:2113-2179): Tiller-Hanson offset of one cubic -> cubics appended to out */
static int cubic_offset(const double *curve, double distance, CurveVec *out)
{
    double stack[64][8]; /* depth-first work list; the 16-output cap bounds it */
    int sp = 0;
    long base = out->n;
    memcpy(stack[sp++], curve, 8 * sizeof(double));
    while (sp > 0) {
        double c[8];
        memcpy(c, stack[--sp], sizeof c);
        if (offset_should_split(c) && (out->n - base) < 16 && sp + 2 <= 64) {
            double halves[16];
            split_single(c, halves);
            memcpy(stack[sp++], halves + 8, 8 * sizeof(double)); /* right pushed first */
            memcpy(stack[sp++], halves, 8 * sizeof(double));     /* left popped next */
            continue;
        }
        Curve o;
        memset(&o, 0, sizeof o);
        int np_ = 0, repeat = 0, have_line = 0;
        double line[4] = {0, 0, 0, 0}, o1[2] = {0, 0};
        for (int i = 0; i < 3; i++) {
            const double *p0 = c + 2 * i, *p1 = c + 2 * i + 2;
            if (close2(p0, p1)) {
                repeat++;
                continue;
            }
            double leg[4] = {p0[0], p0[1], p1[0], p1[1]}, off[4];
            if (!line_offset(leg, distance, off))
                return -4; /* the reference raises here (unpacking None) */
            double o0[2] = {off[0], off[1]};
            o1[0] = off[2], o1[1] = off[3];
            if (have_line) {
                double x0[2], t0, t1;
                if (line_intersect(line, off, x0, &t0, &t1))
                    o0[0] = x0[0], o0[1] = x0[1];
                else
                    o0[0] = (line[2] + o0[0]) / 2, o0[1] = (line[3] + o0[1]) / 2;
            }
            for (int r = 0; r < repeat + 1 && np_ < 4; r++)
                o.p[2 * np_] = o0[0], o.p[2 * np_ + 1] = o0[1], np_++;
            repeat = 0;
            line[0] = o0[0], line[1] = o0[1], line[2] = o1[0], line[3] = o1[1];
            have_line = 1;
        }
        if (have_line) {
            for (int r = 0; r < repeat + 1 && np_ < 4; r++)
                o.p[2 * np_] = o1[0], o.p[2 * np_ + 1] = o1[1], np_++;
            o.n = 4;
            if (out->n > base) {
                const Curve *prev = &out->v[out->n - 1];
                const double *pe = prev->p + 2 * (prev->n - 1);
                if (!close2(o.p, pe)) {
                    double pe_copy[2] = {pe[0], pe[1]};
                    int rc = stroke_cap(o.p, pe_copy, 1, out);
                    if (rc)
                        return rc;
                }
            }
            if (cv_push(out, &o))
                return -1;
        }
    }
    return 0;
}

static void curve_reversed(const Curve *c, Curve *r)
{
    memset(r, 0, sizeof *r);
    r->n = c->n;
    for (int i = 0; i < c->n; i++) {
        r->p[2 * i] = c->p[2 * (c->n - 1 - i)];
        r->p[2 * i + 1] = c->p[2 * (c->n - 1 - i) + 1];
    }
}

/* Path.stroke (:1105-1180).  Input: the path arrays of sceneio.path_arrays
 * (tags 0 line, 1 quad, 2 cubic, 3 arc, 4 closed, 5 unclosed; 8 doubles per
 * segment).  Output: same layout, tags 0/1/2 only.  Returns the number of
 * output segments; -1 capacity, -2 memory, -3 bad cap, -4 degenerate leg. */
ORC_API long orc_stroke(const uint8_t *seg_tag, const double *seg_data, const int32_t *sub_off, long nsub,
                        double width, int cap, int join, uint8_t *out_tag, double *out_data, long seg_cap,
                        int32_t *out_sub_off, long sub_cap, long *out_nsub)
{
    double dist = width / 2;
    long nout = 0, nsubout = 0;
    long rc = 0;
    CurveVec fwd = {0, 0, 0}, bwd = {0, 0, 0}, cur = {0, 0, 0};
    out_sub_off[0] = 0;
#define FLUSH()                                                                         \
    {                                                                                   \
        if (nsubout + 1 > sub_cap || nout + cur.n > seg_cap) {                          \
            rc = -1;                                                                    \
            goto done;                                                                  \
        }                                                                               \
        for (long _i = 0; _i < cur.n; _i++) {                                           \
            out_tag[nout] = (uint8_t)(cur.v[_i].n - 2);                                 \
            memset(out_data + 8 * nout, 0, 8 * sizeof(double));                         \
            memcpy(out_data + 8 * nout, cur.v[_i].p, sizeof(double) * 2 * cur.v[_i].n); \
            nout++;                                                                     \
        }                                                                               \
        out_sub_off[++nsubout] = (int32_t)nout;                                         \
    }
#define CHECK(x)          \
    {                     \
        int _r = (x);     \
        if (_r) {         \
            rc = _r;      \
            goto done;    \
        }                 \
    }
    for (long s = 0; s < nsub; s++) {
        long a = sub_off[s], b = sub_off[s + 1];
        if (b <= a)
            continue;
        fwd.n = bwd.n = cur.n = 0;
        int last_tag = -1;
        for (long i = a; i < b; i++) {
            int tag = seg_tag[i];
            const double *d = seg_data + 8 * i;
            last_tag = tag;
            if (tag == 0 || tag == 4) {
                double off[4];
                if (!line_offset(d, dist, off))
                    continue;
                Curve c = mk_line(off, off + 2);
                CHECK(cv_push(&fwd, &c));
                line_offset(d, -dist, off);
                c = mk_line(off, off + 2);
                CHECK(cv_push(&bwd, &c));
            } else if (tag == 2) {
                CHECK(cubic_offset(d, dist, &fwd));
                CHECK(cubic_offset(d, -dist, &bwd));
            } else if (tag == 1) {
                double c3[8];
                orc_quad_to_cubic(d, c3);
                CHECK(cubic_offset(c3, dist, &fwd));
                CHECK(cubic_offset(c3, -dist, &bwd));
            } else if (tag == 3) {
                double cubs[64 * 8];
                long k = orc_arc_to_cubics(d[0], d[1], d[2], d[3], d[4], d[5], d[6], cubs, 64);
                if (k < 0) {
                    rc = -1;
                    goto done;
                }
                for (long j = 0; j < k; j++) {
                    CHECK(cubic_offset(cubs + 8 * j, dist, &fwd));
                    CHECK(cubic_offset(cubs + 8 * j, -dist, &bwd));
                }
            } else if (tag == 5) {
                continue;
            } else {
                rc = -5;
                goto done;
            }
        }
        int closed = last_tag == 4;
        if (fwd.n == 0)
            continue;
        for (long i = 0; i < fwd.n; i++) {
            if (cur.n > 0) {
                Curve prev = cur.v[cur.n - 1];
                CHECK(stroke_join(&prev, &fwd.v[i], join, &cur));
            }
            CHECK(cv_push(&cur, &fwd.v[i]));
        }
        if (closed) {
            Curve prev = cur.v[cur.n - 1], first = cur.v[0];
            CHECK(stroke_join(&prev, &first, join, &cur));
            FLUSH();
            cur.n = 0;
        } else {
            const Curve *lc = &cur.v[cur.n - 1];
            const Curve *lb = &bwd.v[bwd.n - 1];
            double p0[2] = {lc->p[2 * (lc->n - 1)], lc->p[2 * (lc->n - 1) + 1]};
            double p1[2] = {lb->p[2 * (lb->n - 1)], lb->p[2 * (lb->n - 1) + 1]};
            CHECK(stroke_cap(p0, p1, cap, &cur));
        }
        for (long i = bwd.n - 1; i >= 0; i--) {
            Curve rev;
            curve_reversed(&bwd.v[i], &rev);
            if (cur.n > 0) {
                Curve prev = cur.v[cur.n - 1];
                CHECK(stroke_join(&prev, &rev, join, &cur));
            }
            CHECK(cv_push(&cur, &rev));
        }
        {
            Curve prev = cur.v[cur.n - 1], first = cur.v[0];
            if (closed) {
                CHECK(stroke_join(&prev, &first, join, &cur));
            } else {
                double p0[2] = {prev.p[2 * (prev.n - 1)], prev.p[2 * (prev.n - 1) + 1]};
                double p1[2] = {first.p[0], first.p[1]};
                CHECK(stroke_cap(p0, p1, cap, &cur));
            }
        }
        FLUSH();
    }
done:
    free(fwd.v);
    free(bwd.v);
    free(cur.v);
    *out_nsub = nsubout;
    return rc < 0 ? rc : nout;
#undef FLUSH
#undef CHECK
}
