// k_stroke.cu -- stroke outlines on the device (subsystem 1, "stroke outlines").
//
// Replaces Path.stroke (svgrasterize.py:1105-1180) and its helpers:
//   bezier3_offset (Tiller-Hanson)   svgrasterize.py:2113-2179   (+ bezier3_split :2058-2063)
//   line_offset / line_intersect     svgrasterize.py:2328-2337 / :2307-2325
//   stroke_line_cap / _join          svgrasterize.py:1466-1492 / :1495-1518
//   stroke_curve_tangent             svgrasterize.py:1521-1530
//
// The reference strokes one segment after another in a Python loop (9 ms per
// stroke).  Here the work is cut along its real dependencies:
//   phase A  one thread per input segment: offset curves at +w/2 and -w/2
//            (count pass, exclusive scan, emit pass -- no per-thread buffers);
//   phase B  one warp per range of <= 256 segments of a sub-path (stroke_assemble_kernel): every curve owns
//            three fixed outline slots (at most two join curves + itself, unused = SEG_NOP), so what a curve
//            writes depends on its neighbour only -- forward offsets, the cap, the reversed backward offsets,
//            joins in between; slot offsets come from a scanned upper bound per sub-path.
// All arithmetic is float64 with the reference's rounding recipes (SURVEY.md
// appendix B; file compiled with -fmad=false, fma() spelled out), so outlines are
// bit-identical to the reference's and flatten like any other path afterwards.
#include "svgr_kernels.h"

struct DCurve {
    int n;  // 2 line, 3 quad, 4 cubic
    int pad;
    double p[8];
};

#define STROKE_EPS 2.220446049250313e-16
#define STROKE_SQRT2 1.4142135623730951
#define STROKE_CIRCLE_K (4 * (STROKE_SQRT2 - 1) / 3)

__device__ __forceinline__ bool close1(double a, double b)
{
    // np.allclose: |a - b| <= atol + rtol * |b| with rtol 1e-5, atol 1e-8
    if (isfinite(a) && isfinite(b))
        return fabs(a - b) <= 1e-8 + 1e-5 * fabs(b);
    return a == b;
}
__device__ __forceinline__ bool close2(const double *a, const double *b) { return close1(a[0], b[0]) && close1(a[1], b[1]); }

__device__ __forceinline__ DCurve mk_line(const double *a, const double *b)
{
    DCurve c;
    c.n = 2, c.pad = 0;
    c.p[0] = a[0], c.p[1] = a[1], c.p[2] = b[0], c.p[3] = b[1];
    c.p[4] = c.p[5] = c.p[6] = c.p[7] = 0.0;
    return c;
}

// line_offset (svgrasterize.py:2328-2337); false when the line is degenerate (None)
__device__ __forceinline__ bool line_offset(const double *l, double distance, double *out)
{
    double x1 = l[0], y1 = l[1], x2 = l[2], y2 = l[3];
    double vx = x2 - x1, vy = y2 - y1;
    double len = vx * vx + vy * vy;
    if (len < STROKE_EPS)
        return false;
    len = sqrt(len);
    double dx = -vy * distance / len;
    double dy = vx * distance / len;
    out[0] = x1 + dx, out[1] = y1 + dy, out[2] = x2 + dx, out[3] = y2 + dy;
    return true;
}

// line_intersect (svgrasterize.py:2307-2325); false for (None, 0, 0)
__device__ __forceinline__ bool line_intersect(const double *l0, const double *l1, double *p, double *t0o, double *t1o)
{
    double x1 = l0[0], y1 = l0[1], x2 = l0[2], y2 = l0[3];
    double x3 = l1[0], y3 = l1[1], x4 = l1[2], y4 = l1[3];
    double det = (x4 - x3) * (y1 - y2) - (x1 - x2) * (y4 - y3);
    if (fabs(det) < STROKE_EPS) {
        *t0o = 0, *t1o = 0;
        return false;
    }
    double t0 = ((y3 - y4) * (x1 - x3) + (x4 - x3) * (y1 - y3)) / det;
    double t1 = ((y1 - y2) * (x1 - x3) + (x2 - x1) * (y1 - y3)) / det;
    p[0] = x1 * (1 - t0) + x2 * t0;
    p[1] = y1 * (1 - t0) + y2 * t0;
    *t0o = t0, *t1o = t1;
    return true;
}

// A sink receives curves in order and remembers the last one.  Count mode stores nothing.
struct Sink {
    DCurve *out;  // null: count only
    int n;
    int cap;  // emit mode: slots available (writes beyond are dropped, still counted)
    DCurve last;
    __device__ __forceinline__ void push(const DCurve &c)
    {
        if (out && n < cap)
            out[n] = c;
        n++;
        last = c;
    }
};

// stroke_line_cap (svgrasterize.py:1466-1492). cap: 0 butt, 1 round, 2 square; returns false on a bad cap id
__device__ bool stroke_cap(const double *p0, const double *p1, int cap, Sink &out)
{
    if (close2(p0, p1))
        return true;
    if (cap == 0) {
        out.push(mk_line(p0, p1));
        return true;
    }
    if (cap == 1) {
        double seg[2] = {p1[0] - p0[0], p1[1] - p0[1]};
        double radius = sqrt(fma(seg[1], seg[1], seg[0] * seg[0])) / 2;
        double two_r = 2 * radius;
        seg[0] /= two_r, seg[1] /= two_r;
        double nrm[2] = {-seg[1], seg[0]};
        double offset = STROKE_CIRCLE_K * radius;
        double center[2] = {(p0[0] + p1[0]) / 2, (p0[1] + p1[1]) / 2};
        double mid[2] = {center[0] + nrm[0] * radius, center[1] + nrm[1] * radius};
        DCurve a, b;
        a.n = b.n = 4, a.pad = b.pad = 0;
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
        out.push(a);
        out.push(b);
        return true;
    }
    if (cap == 2) {
        double seg[2] = {p1[0] - p0[0], p1[1] - p0[1]};
        double nrm[2] = {-seg[1], seg[0]};
        double q0[2] = {p0[0] + nrm[0] / 2, p0[1] + nrm[1] / 2};
        double q1[2] = {p1[0] + nrm[0] / 2, p1[1] + nrm[1] / 2};
        out.push(mk_line(p0, q0));
        out.push(mk_line(q0, q1));
        out.push(mk_line(q1, p1));
        return true;
    }
    return false;
}

// stroke_curve_tangent (svgrasterize.py:1521-1530)
__device__ bool curve_tangents(const DCurve &c, double *first, double *last)
{
    bool found = false;
    for (int i = 0; i + 1 < c.n; i++) {
        const double *a = c.p + 2 * i, *b = c.p + 2 * i + 2;
        if (close2(a, b))
            continue;
        if (!found)
            first[0] = a[0], first[1] = a[1], first[2] = a[2], first[3] = a[3];
        last[0] = a[0], last[1] = a[1], last[2] = a[2], last[3] = a[3];
        found = true;
    }
    return found;
}

// stroke_line_join (svgrasterize.py:1495-1518). join: 0 miter, 1 round, 2 bevel, 3 anything else
__device__ void stroke_join(const DCurve &c0, const DCurve &c1, int join, Sink &out)
{
    double e0[2] = {c0.p[2 * (c0.n - 1)], c0.p[2 * (c0.n - 1) + 1]};
    double s1[2] = {c1.p[0], c1.p[1]};
    DCurve bevel = mk_line(e0, s1);
    if (join == 2) {
        out.push(bevel);
        return;
    }
    double f0[4], l0[4], f1[4], l1[4];
    bool ok0 = curve_tangents(c0, f0, l0);
    bool ok1 = curve_tangents(c1, f1, l1);
    if (!ok0 || !ok1) {
        out.push(bevel);
        return;
    }
    if (close2(l0 + 2, f1))
        return;
    double p[2], t0, t1;
    bool hit = line_intersect(l0, f1, p, &t0, &t1);
    if (!hit || (0 <= t0 && t0 <= 1 && 0 <= t1 && t1 <= 1)) {
        out.push(bevel);
        return;
    }
    if (fabs(t0) < 4 && fabs(t1) < 4) {
        if (join == 0) {
            out.push(mk_line(e0, p));
            out.push(mk_line(p, s1));
            return;
        }
        if (join == 1) {
            DCurve q;
            q.n = 3, q.pad = 0;
            q.p[0] = e0[0], q.p[1] = e0[1], q.p[2] = p[0], q.p[3] = p[1], q.p[4] = s1[0], q.p[5] = s1[1];
            q.p[6] = q.p[7] = 0.0;
            out.push(q);
            return;
        }
    }
    out.push(bevel);
}

// should_split of bezier3_offset (svgrasterize.py:2121-2137)
__device__ bool offset_should_split(const double *c)
{
    double bx = c[6] - c[0], by = c[7] - c[1];
    double mx = c[4] - c[2], my = c[5] - c[3];
    if (fma(by, my, bx * mx) < 0)
        return true;
    double u1x = c[2] - c[0], u1y = c[3] - c[1];
    double u2x = c[4] - c[0], u2y = c[5] - c[1];
    double a0 = bx * u1y - by * u1x;
    double a1 = bx * u2y - by * u2x;
    if (a0 * a1 < 0)
        return true;
    double dd[2], gg[2];
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
    double dist = dd[0] + dd[1];
    double diag = gg[0] + gg[1];
    return dist * 100 > diag;
}

// bezier3_split (svgrasterize.py:2058-2063): np.matmul(BEZIER3_SPLIT, points), a 4-term fma chain
__device__ __forceinline__ double chain4(double m0, double m1, double m2, double m3, double p0, double p1, double p2,
                                         double p3)
{
    return fma(m3, p3, fma(m2, p2, fma(m1, p1, m0 * p0)));
}

__device__ void split_single(const double *c, double *out)
{
    for (int k = 0; k < 2; k++) {
        double p0 = c[k], p1 = c[2 + k], p2 = c[4 + k], p3 = c[6 + k];
        out[0 + k] = chain4(1, 0, 0, 0, p0, p1, p2, p3);
        out[2 + k] = chain4(0.5, 0.5, 0, 0, p0, p1, p2, p3);
        out[4 + k] = chain4(0.25, 0.5, 0.25, 0, p0, p1, p2, p3);
        out[6 + k] = chain4(0.125, 0.375, 0.375, 0.125, p0, p1, p2, p3);
        out[8 + k] = chain4(0.125, 0.375, 0.375, 0.125, p0, p1, p2, p3);
        out[10 + k] = chain4(0, 0.25, 0.5, 0.25, p0, p1, p2, p3);
        out[12 + k] = chain4(0, 0, 0.5, 0.5, p0, p1, p2, p3);
        out[14 + k] = chain4(0, 0, 0, 1, p0, p1, p2, p3);
    }
}

// bezier3_offset (svgrasterize.py:2113-2179).  Returns false when a control-polygon leg is degenerate
// in a way the reference cannot handle (it raises TypeError there).
__device__ bool cubic_offset(const double *curve, double distance, Sink &out)
{
    double stack[64][8];
    int sp = 0;
    const int base = out.n;
    for (int k = 0; k < 8; k++)
        stack[0][k] = curve[k];
    sp = 1;
    while (sp > 0) {
        double c[8];
        --sp;
        for (int k = 0; k < 8; k++)
            c[k] = stack[sp][k];
        if (offset_should_split(c) && (out.n - base) < 16 && sp + 2 <= 64) {
            double halves[16];
            split_single(c, halves);
            for (int k = 0; k < 8; k++) {
                stack[sp][k] = halves[8 + k];      // right half pushed first
                stack[sp + 1][k] = halves[k];      // left half popped next
            }
            sp += 2;
            continue;
        }
        DCurve o;
        o.n = 4, o.pad = 0;
        for (int k = 0; k < 8; k++)
            o.p[k] = 0.0;
        int np_ = 0, repeat = 0;
        bool have_line = false;
        double line[4] = {0, 0, 0, 0}, o1[2] = {0, 0};
        for (int i = 0; i < 3; i++) {
            const double *p0 = c + 2 * i, *p1 = c + 2 * i + 2;
            if (close2(p0, p1)) {
                repeat++;
                continue;
            }
            double leg[4] = {p0[0], p0[1], p1[0], p1[1]}, off[4];
            if (!line_offset(leg, distance, off))
                return false;
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
            have_line = true;
        }
        if (have_line) {
            for (int r = 0; r < repeat + 1 && np_ < 4; r++)
                o.p[2 * np_] = o1[0], o.p[2 * np_ + 1] = o1[1], np_++;
            if (out.n > base) {
                double pe[2] = {out.last.p[2 * (out.last.n - 1)], out.last.p[2 * (out.last.n - 1) + 1]};
                if (!close2(o.p, pe))
                    stroke_cap(o.p, pe, 1, out);
            }
            out.push(o);
        }
    }
    return true;
}

__device__ __forceinline__ void quad_to_cubic(const double *q, double *c)
{
    const double t1 = 1.0 / 3, t2 = 2.0 / 3;
    for (int k = 0; k < 2; k++) {
        c[0 + k] = fma(0.0, q[4 + k], fma(0.0, q[2 + k], 1.0 * q[k]));
        c[2 + k] = fma(0.0, q[4 + k], fma(t2, q[2 + k], t1 * q[k]));
        c[4 + k] = fma(t1, q[4 + k], fma(t2, q[2 + k], 0.0 * q[k]));
        c[6 + k] = fma(1.0, q[4 + k], fma(0.0, q[2 + k], 0.0 * q[k]));
    }
}

// offsets of one input segment on one side; returns false on the reference's TypeError case
__device__ bool offset_segment(int tag, const double *d, double dist, Sink &out)
{
    if (tag == SEG_LINE || tag == SEG_CLOSED) {
        double off[4];
        if (line_offset(d, dist, off))
            out.push(mk_line(off, off + 2));
        return true;
    }
    if (tag == SEG_CUBIC)
        return cubic_offset(d, dist, out);
    if (tag == SEG_QUAD) {
        double c3[8];
        quad_to_cubic(d, c3);
        return cubic_offset(c3, dist, out);
    }
    return true;  // SEG_UNCLOSED is skipped by the stroker (svgrasterize.py:1142-1143)
}

// seg_job[i] = stroke job of input segment i
// Two threads per input segment, one per side of the stroke: the kernels are one wave of long, branchy float64
// threads, so their time is the latency of the slowest thread -- which is half as long when it offsets one side.
__global__ void stroke_count_kernel(const uint8_t *__restrict__ tag, const double *__restrict__ data,
                                    const int *__restrict__ seg_job, const StrokeRec *__restrict__ jobs, int n_seg,
                                    int *__restrict__ counts, int *__restrict__ err)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 1, side = t & 1;
    if (i >= n_seg)
        return;
    double dist = jobs[seg_job[i]].half_width;
    double d[8];
    for (int k = 0; k < 8; k++)
        d[k] = data[8 * (size_t)i + k];
    Sink q = {nullptr, 0, 0, {}};
    // a degenerate line contributes to neither side (svgrasterize.py:1121-1125): the backward side of a line
    // counts what the forward side counts
    const bool line = tag[i] == SEG_LINE || tag[i] == SEG_CLOSED;
    const bool ok = offset_segment(tag[i], d, (side == 0 || line) ? dist : -dist, q);
    if (!ok && (side == 0 || !line))
        atomicOr(err, 1);
    counts[side * n_seg + i] = q.n;
}

__global__ void stroke_emit_kernel(const uint8_t *__restrict__ tag, const double *__restrict__ data,
                                   const int *__restrict__ seg_job, const StrokeRec *__restrict__ jobs, int n_seg,
                                   const int *__restrict__ offs, DCurve *__restrict__ pool, int pool_cap)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = t >> 1, side = t & 1;
    if (i >= n_seg)
        return;
    double dist = jobs[seg_job[i]].half_width;
    double d[8];
    for (int k = 0; k < 8; k++)
        d[k] = data[8 * (size_t)i + k];
    const int o = offs[side * n_seg + i];
    Sink q = {pool + o, 0, max(0, pool_cap - o), {}};
    if (side == 0) {
        offset_segment(tag[i], d, dist, q);
    } else if (tag[i] == SEG_LINE || tag[i] == SEG_CLOSED) {
        double off[4];
        if (line_offset(d, -dist, off))
            q.push(mk_line(off, off + 2));
    } else {
        offset_segment(tag[i], d, -dist, q);
    }
}

// upper bound of the outline segments of one sub-path: every offset curve, at most two join
// curves before each, two caps of at most three curves
__global__ void stroke_bound_kernel(const int *__restrict__ sub_off, int n_sub, int n_seg, const int *__restrict__ offs,
                                    int *__restrict__ bound)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sub)
        return;
    int a = sub_off[s], b = sub_off[s + 1];
    if (b <= a) {
        bound[s] = 0;
        return;
    }
    // offs = exclusive scan over [forward counts | backward counts | 0]: offs[2 n_seg] is the total
    int f0 = offs[a], f1 = offs[b];
    int b0 = offs[n_seg + a], b1 = offs[n_seg + b];
    int nc = (f1 - f0) + (b1 - b0);
    bound[s] = nc == 0 ? 0 : 3 * nc + 10;
}

// Output of the assembly: every curve of a sub-path owns three consecutive outline slots (at most two join curves
// before it, then the curve itself), the cap / closing join after the forward side and the one at the end own
// three each; unused slots are SEG_NOP (skipped by the flattener and by svgr_read_outline).  With fixed slots a
// curve's output depends on its neighbour only, so the lanes of a warp assemble one sub-path together.
struct OutlineOut {
    uint8_t *tag;
    double *data;
    uint32_t *path;
    int32_t *sub;  // output sub-path id (2 * input sub-path + part), may be null
    long long base, cap;
    int limit;
    uint32_t path_id;
    __device__ __forceinline__ void put(int k, const DCurve &c, int sub_id) const
    {
        const long long i = base + k;
        if (k < limit && i < cap) {
            tag[i] = (uint8_t)(c.n - 2);
            double2 *d = reinterpret_cast<double2 *>(data + 8 * i);
            d[0] = make_double2(c.p[0], c.p[1]);
            d[1] = make_double2(c.p[2], c.p[3]);
            d[2] = make_double2(c.n > 2 ? c.p[4] : 0.0, c.n > 2 ? c.p[5] : 0.0);
            d[3] = make_double2(c.n > 3 ? c.p[6] : 0.0, c.n > 3 ? c.p[7] : 0.0);
            path[i] = path_id;
            if (sub)
                sub[i] = sub_id;
        }
    }
    __device__ __forceinline__ void nop(int k) const
    {
        const long long i = base + k;
        if (k < limit && i < cap) {
            tag[i] = SEG_NOP;
            path[i] = path_id;
            if (sub)
                sub[i] = -1;
        }
    }
    // slots k .. k + 2: the join from `prev` (if any) and then the curve; returns nothing
    __device__ __forceinline__ void curve_with_join(int k, const DCurve *prev, const DCurve &c, int join, int sub_id) const
    {
        int n = 0;
        if (prev) {
            Sink t = {nullptr, 0, 0, {}};
            DCurve tmp[2];
            t.out = tmp, t.cap = 2;
            stroke_join(*prev, c, join, t);
            for (int q = 0; q < t.n && q < 2; q++)
                put(k + n++, tmp[q], sub_id);
        }
        put(k + n++, c, sub_id);
        for (; n < 3; n++)
            nop(k + n);
    }
};

__device__ __forceinline__ DCurve reversed(const DCurve &c)
{
    DCurve r;
    r.n = c.n, r.pad = 0;
    for (int k = 0; k < 8; k++)
        r.p[k] = 0.0;
    for (int i = 0; i < c.n; i++) {
        r.p[2 * i] = c.p[2 * (c.n - 1 - i)];
        r.p[2 * i + 1] = c.p[2 * (c.n - 1 - i) + 1];
    }
    return r;
}

// Phase B: Path.stroke's per-sub-path assembly (svgrasterize.py:1147-1178), one warp per sub-path.  The reference
// walks: forward offset curves with a join before each but the first; then, closed: the join back to the first
// curve (and the backward side starts a new sub-path), open: the cap over to the end of the backward side; the
// backward curves reversed, in reverse order, a join before each (the first one joins to the cap's last curve);
// finally closed: the join back to the first backward curve, open: the cap back to the start.
__global__ void stroke_assemble_kernel(const uint8_t *__restrict__ in_tag, const int *__restrict__ sub_off,
                                       const int *__restrict__ sub_job, const StrokeRec *__restrict__ jobs, int n_sub,
                                       int n_seg, const int *__restrict__ items, int n_items,
                                       const int *__restrict__ offs, const DCurve *__restrict__ pool, const int *__restrict__ bound,
                                       const int *__restrict__ out_off, const int *__restrict__ out_total,
                                       long long out_base, long long out_cap, uint8_t *__restrict__ out_tag,
                                       double *__restrict__ out_data, uint32_t *__restrict__ out_path,
                                       int32_t *__restrict__ out_sub, int *__restrict__ n_out_dev, int *__restrict__ err)
{
    // work item = (sub-path, segment range): a long polyline is cut into ranges of <= 256 segments, a warp each
    const int item = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n_out_dev)
        *n_out_dev = (int)min((long long)*out_total + out_base, out_cap);
    if (item >= n_items)
        return;
    const int s = items[3 * item], ia = items[3 * item + 1], ib = items[3 * item + 2];
    const int a = sub_off[s], b = sub_off[s + 1];
    const int limit = bound[s];
    if (b <= a || limit == 0)
        return;
    const bool first_range = ia == a, last_range = ib == b;
    const StrokeRec &job = jobs[sub_job[s]];
    const int f0 = offs[a], f1 = offs[b];
    const int b0 = offs[n_seg + a], b1 = offs[n_seg + b];
    const int nf = f1 - f0, nb = b1 - b0;
    OutlineOut o;
    o.tag = out_tag, o.data = out_data, o.path = out_path, o.sub = out_sub;
    o.base = out_base + out_off[s], o.cap = out_cap, o.limit = limit, o.path_id = (uint32_t)job.path;
    if (nf == 0) {
        if (first_range)
            for (int k = lane; k < limit; k += 32)
                o.nop(k);
        return;
    }
    // this range's curves on both sides
    const int fi0 = offs[ia] - f0, fi1 = offs[ib] - f0;
    const int bm0 = offs[n_seg + ia] - b0, bm1 = offs[n_seg + ib] - b0;
    const bool closed = in_tag[b - 1] == SEG_CLOSED;
    const int join = job.join, sub_f = 2 * s, sub_b = closed ? 2 * s + 1 : 2 * s;
    bool ok = true;
    // ---- forward side: curve i owns slots 3 i .. 3 i + 2
    for (int i = fi0 + lane; i < fi1; i += 32) {
        const DCurve c = pool[f0 + i];
        if (i > 0) {
            const DCurve prev = pool[f0 + i - 1];
            o.curve_with_join(3 * i, &prev, c, join, sub_f);
        } else {
            o.curve_with_join(0, nullptr, c, join, sub_f);
        }
    }
    // ---- between the sides: slots 3 nf .. 3 nf + 2 (lane 0 of the last range)
    const int kmid = 3 * nf;
    if (lane == 0 && last_range) {
        const DCurve last_fwd = pool[f1 - 1];
        int n = 0;
        if (closed) {
            Sink t = {nullptr, 0, 0, {}};
            DCurve tmp[2];
            t.out = tmp, t.cap = 2;
            stroke_join(last_fwd, pool[f0], join, t);
            for (int q = 0; q < t.n && q < 2; q++)
                o.put(kmid + n++, tmp[q], sub_f);
        } else if (nb > 0) {
            const DCurve lb = pool[b0 + nb - 1];
            const double p0[2] = {last_fwd.p[2 * (last_fwd.n - 1)], last_fwd.p[2 * (last_fwd.n - 1) + 1]};
            const double p1[2] = {lb.p[2 * (lb.n - 1)], lb.p[2 * (lb.n - 1) + 1]};
            Sink t = {nullptr, 0, 0, {}};
            DCurve tmp[3];
            t.out = tmp, t.cap = 3;
            ok = stroke_cap(p0, p1, job.cap, t) && ok;
            for (int q = 0; q < t.n && q < 3; q++)
                o.put(kmid + n++, tmp[q], sub_f);
        }
        for (; n < 3; n++)
            o.nop(kmid + n);
    }
    // ---- backward side, reversed curves in reverse order: the j-th one owns slots 3 nf + 3 + 3 j ..
    for (int m = bm1 - 1 - lane; m >= bm0; m -= 32) {
        const int j = nb - 1 - m;  // position on the backward side
        const DCurve r = reversed(pool[b0 + m]);
        const int k = kmid + 3 + 3 * j;
        if (j > 0) {
            const DCurve prev = reversed(pool[b0 + nb - j]);
            o.curve_with_join(k, &prev, r, join, sub_b);
        } else if (!closed) {
            // the first backward curve joins to the last curve of the cap (recomputed here: whichever lane owns
            // j == 0 -- trailing degenerate segments have no curves -- need not be the one that wrote the cap)
            const DCurve last_fwd = pool[f1 - 1], lb = pool[b0 + nb - 1];
            const double p0[2] = {last_fwd.p[2 * (last_fwd.n - 1)], last_fwd.p[2 * (last_fwd.n - 1) + 1]};
            const double p1[2] = {lb.p[2 * (lb.n - 1)], lb.p[2 * (lb.n - 1) + 1]};
            Sink t = {nullptr, 0, 0, {}};
            DCurve tmp[3];
            t.out = tmp, t.cap = 3;
            stroke_cap(p0, p1, job.cap, t);
            const DCurve before_back = t.n > 0 ? tmp[min(t.n, 3) - 1] : last_fwd;
            o.curve_with_join(k, &before_back, r, join, sub_b);
        } else {
            o.curve_with_join(k, nullptr, r, join, sub_b);  // the backward side of a closed sub-path starts afresh
        }
    }
    // ---- the end: slots 3 (nf + nb) + 3 ..
    const int kend = kmid + 3 + 3 * nb;
    if (lane == 0 && first_range) {
        int n = 0;
        if (closed) {
            if (nb > 0) {
                Sink t = {nullptr, 0, 0, {}};
                DCurve tmp[2];
                t.out = tmp, t.cap = 2;
                stroke_join(reversed(pool[b0]), reversed(pool[b0 + nb - 1]), join, t);
                for (int q = 0; q < t.n && q < 2; q++)
                    o.put(kend + n++, tmp[q], sub_b);
            }
        } else {
            const DCurve last = nb > 0 ? reversed(pool[b0]) : pool[f1 - 1];
            const DCurve first = pool[f0];
            const double p0[2] = {last.p[2 * (last.n - 1)], last.p[2 * (last.n - 1) + 1]};
            const double p1[2] = {first.p[0], first.p[1]};
            Sink t = {nullptr, 0, 0, {}};
            DCurve tmp[3];
            t.out = tmp, t.cap = 3;
            ok = stroke_cap(p0, p1, job.cap, t) && ok;
            for (int q = 0; q < t.n && q < 3; q++)
                o.put(kend + n++, tmp[q], sub_b);
        }
        for (; n < 3; n++)
            o.nop(kend + n);
    }
    if (first_range)
        for (int k = kend + 3 + lane; k < limit; k += 32)
            o.nop(k);
    if (!ok || kend + 3 > limit)
        atomicOr(err, 2);
}

// ---------------------------------------------------------------------------------------------
size_t svgr_stroke_curve_bytes() { return sizeof(DCurve); }

void svgr_launch_stroke_count(const uint8_t *tag, const double *data, const int *seg_job, const StrokeRec *jobs,
                              int n_seg, int *counts, int *err, cudaStream_t s)
{
    if (n_seg > 0)
        stroke_count_kernel<<<(2 * n_seg + 63) / 64, 64, 0, s>>>(tag, data, seg_job, jobs, n_seg, counts, err);
}

void svgr_launch_stroke_emit(const uint8_t *tag, const double *data, const int *seg_job, const StrokeRec *jobs,
                             int n_seg, const int *offs, void *pool, int pool_cap, cudaStream_t s)
{
    if (n_seg > 0)
        stroke_emit_kernel<<<(2 * n_seg + 63) / 64, 64, 0, s>>>(tag, data, seg_job, jobs, n_seg, offs, (DCurve *)pool,
                                                            pool_cap);
}

void svgr_launch_stroke_bound(const int *sub_off, int n_sub, int n_seg, const int *offs, int *bound, cudaStream_t s)
{
    if (n_sub > 0)
        stroke_bound_kernel<<<(n_sub + 127) / 128, 128, 0, s>>>(sub_off, n_sub, n_seg, offs, bound);
}

void svgr_launch_stroke_assemble(const uint8_t *in_tag, const int *sub_off, const int *sub_job, const StrokeRec *jobs,
                                 int n_sub, int n_seg, const int *items, int n_items, const int *offs, const void *pool,
                                 const int *bound, const int *out_off, const int *out_total, long long out_base,
                                 long long out_cap, uint8_t *out_tag, double *out_data, uint32_t *out_path,
                                 int32_t *out_sub, int *n_out_dev, int *err, cudaStream_t s)
{
    if (n_sub > 0 && n_items > 0)
        stroke_assemble_kernel<<<(unsigned)(((long long)n_items * 32 + 127) / 128), 128, 0, s>>>(
            in_tag, sub_off, sub_job, jobs, n_sub, n_seg, items, n_items, offs, (const DCurve *)pool, bound, out_off, out_total,
            out_base, out_cap, out_tag, out_data, out_path, out_sub, n_out_dev, err);
}
