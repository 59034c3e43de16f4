// encode_flat.cpp -- the scene encoder in native code.
//
// The walk Scene.render makes over a scene tree (svgrasterize.py:649-752), which svgrasterize_b200/encode.py
// restates in Python to *record* a scene program instead of rendering, done here on flat scene arrays
// (include/svgr_b200.h: svgr_flat).  Python at ~1.3 ms per icon per core was 90 % of the wall time of a 100 000
// icon job; this walk costs a few microseconds per icon.  The arrays come from svgrasterize_b200/_flatten.c, a
// CPython extension that reads the reference's own Scene / Path / paint / Transform objects directly.
//
// Covered: fills and strokes with solid, linear and radial (incl. two-circle) paints in user space, groups,
// opacity, transforms, clip paths and luminance masks.  Scenes that use objectBoundingBox units, pattern paints or
// filters are encoded by encode.py (the flattener flags them), whose output this code reproduces record for
// record on everything it does cover (tests/test_host_logic.py::test_native_encoder_*).
//
// Arithmetic: matrix products follow numpy's 3 x 3 matmul rounding (fma(a2, b2, fma(a1, b1, a0 b0)), measured),
// small dots numpy's (fma(u1, v1, u0 v0)).  The inverse of a composed transform is an LU solve with partial
// pivoting like numpy's (LAPACK gesv): identical for the scale / swap / translate transforms of real documents,
// within an ulp of it for general matrices (the Python encoder remains the bit-exact path for those).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/svgr_b200.h"
#include "svgr_types.h"

#include "svgr_affine.h"

extern "C" int64_t svgr_arc_to_cubics(double cx, double cy, double rx, double ry, double phi, double eta, double eta_delta,
                                      double *out, int64_t cap);

namespace {

enum { R_FILL = 0, R_STROKE = 1, R_GROUP = 2, R_OPACITY = 3, R_CLIP = 4, R_MASK = 5, R_TRANSFORM = 6, R_FILTER = 7 };

inline M23 matmul(const M23 &a, const M23 &b) { return affine_matmul(a, b); }
inline bool invert(const M23 &t, M23 &out) { return affine_invert(t, out); }

// color.py paint_to_srgb: premultiplied linear RGBA -> premultiplied sRGB RGBA
void paint_to_srgb(const double *c, double *o)
{
    double r = c[0], g = c[1], b = c[2], a = c[3];
    if (a > 0.0001)
        r = r / a, g = g / a, b = b / a;
    auto clip = [](double v) { return fmin(fmax(v, 0.0), 1.0); };
    r = clip(r), g = clip(g), b = clip(b), a = clip(a);
    auto enc = [](double v) { return v <= 0.0031308 ? v * 12.92 : 1.055 * pow(v, 1.0 / 2.4) - 0.055; };
    o[0] = enc(r) * a, o[1] = enc(g) * a, o[2] = enc(b) * a, o[3] = a;
}

struct Canvas {
    int32_t node;
    int64_t off;
    int32_t rows, cols;
};

}  // namespace

struct svgr_encoded {
    std::vector<uint8_t> seg_tag, stroke_tag;
    std::vector<double> seg_data, stroke_data;
    std::vector<uint32_t> seg_path;
    std::vector<PathRec> paths;
    std::vector<StrokeRec> strokes;
    std::vector<int32_t> stroke_sub_off, stroke_sub_job, stroke_seg_job;
    std::vector<PaintRec> paints;
    std::vector<StopRec> stops;
    std::vector<svgr_node> nodes;
    std::vector<int32_t> children;
    std::vector<svgr_bbox_job> bbox_jobs;
    std::vector<svgr_kernel> kernels;
    std::vector<float> weights, matrices;  // matrices: 20 per colour matrix
    std::vector<double> offset_tr;         // 12 per feOffset: forward 2 x 3, inverse 2 x 3
    std::vector<int64_t> canvases;  // node, byte offset, rows, cols per canvas
    std::vector<int32_t> roots;
    int32_t n_focal = 0;
    int64_t canvas_bytes = 0;
    svgr_program prog;
    std::string err;
    int err_code = SVGR_OK;
};

namespace {

struct Encoder {
    const svgr_flat &in;
    svgr_encoded &out;
    // per path of the input: its segments with arcs expanded, built on first use
    struct FlatPath {
        bool ready = false;
        std::vector<uint8_t> tag;
        std::vector<double> data;
        std::vector<int32_t> sub_off;  // relative to the path's first segment
    };
    std::vector<FlatPath> path_cache;
    std::vector<int32_t> kid_stack;

    Encoder(const svgr_flat &i, svgr_encoded &o) : in(i), out(o), path_cache((size_t)i.n_path) { out.stroke_sub_off.push_back(0); }

    bool fail(int code, const char *msg)
    {
        if (out.err_code == SVGR_OK)
            out.err_code = code, out.err = msg;
        return false;
    }

    int node(int tag, int a = 0, int b = 0, int c = 0, int d = 0, const int32_t *kids = nullptr, int n_kids = 0, int flags = 0,
             double f0 = 0.0, double f1 = 0.0, double f2 = 0.0, double f3 = 0.0)
    {
        svgr_node n;
        n.tag = tag, n.a = a, n.b = b, n.c = c, n.d = d;
        n.child_off = (int32_t)out.children.size(), n.child_cnt = n_kids, n.flags = flags;
        n.f[0] = f0, n.f[1] = f1, n.f[2] = f2, n.f[3] = f3;
        for (int k = 0; k < n_kids; k++)
            out.children.push_back(kids[k]);
        out.nodes.push_back(n);
        return (int)out.nodes.size() - 1;
    }
    int empty() { return node(SVGR_N_EMPTY); }
    bool is_empty(int n) const { return out.nodes[(size_t)n].tag == SVGR_N_EMPTY; }

    // encode.py device_path: sub-path structure with arcs expanded to cubics (arc_to_bezier3, :2355)
    const FlatPath *flat_path(int pid)
    {
        if (pid < 0 || pid >= in.n_path) {
            fail(SVGR_E_INVALID, "flat scene: path index out of range");
            return nullptr;
        }
        FlatPath &fp = path_cache[(size_t)pid];
        if (fp.ready)
            return &fp;
        fp.sub_off.push_back(0);
        for (int s = in.path_off[pid]; s < in.path_off[pid + 1]; s++) {
            for (int i = in.sub_off[s]; i < in.sub_off[s + 1]; i++) {
                const double *d = in.seg_data + 8 * (size_t)i;
                if (in.seg_tag[i] == SEG_ARC) {
                    double buf[64 * 8];
                    const int64_t k = svgr_arc_to_cubics(d[0], d[1], d[2], d[3], d[4], d[5], d[6], buf, 64);
                    if (k < 0) {
                        fail(SVGR_E_INVALID, "arc sweeps more than 16 pi");
                        return nullptr;
                    }
                    for (int64_t j = 0; j < k; j++) {
                        fp.tag.push_back(SEG_CUBIC);
                        fp.data.insert(fp.data.end(), buf + 8 * j, buf + 8 * j + 8);
                    }
                } else {
                    fp.tag.push_back(in.seg_tag[i]);
                    fp.data.insert(fp.data.end(), d, d + 8);
                }
            }
            fp.sub_off.push_back((int32_t)fp.tag.size());
        }
        fp.ready = true;
        return &fp;
    }

    int add_path_rec(const M23 &t, const int32_t *viewport, int rule)
    {
        PathRec p;
        memset(&p, 0, sizeof p);
        memcpy(p.m, t.m, sizeof p.m);
        if (viewport) {
            p.has_viewport = 1;
            memcpy(p.viewport, viewport, sizeof p.viewport);
        }
        p.fill_rule = rule;
        out.paths.push_back(p);
        return (int)out.paths.size() - 1;
    }

    int add_fill_path(int path, const M23 &t, int rule, const int32_t *viewport)
    {
        const FlatPath *fp = flat_path(path);
        if (!fp)
            return -1;
        const int pid = add_path_rec(t, viewport, rule);
        out.seg_tag.insert(out.seg_tag.end(), fp->tag.begin(), fp->tag.end());
        out.seg_data.insert(out.seg_data.end(), fp->data.begin(), fp->data.end());
        out.seg_path.insert(out.seg_path.end(), fp->tag.size(), (uint32_t)pid);
        return pid;
    }

    int add_stroke_path(int path, const M23 &t, double width, int cap, int join, const int32_t *viewport)
    {
        const FlatPath *fp = flat_path(path);
        if (!fp)
            return -1;
        const int pid = add_path_rec(t, viewport, 0);
        const int job = (int)out.strokes.size();
        const int sub_begin = (int)out.stroke_sub_job.size();
        const int n_sub = (int)fp->sub_off.size() - 1;
        const int32_t base = (int32_t)out.stroke_tag.size();
        for (int s = 1; s <= n_sub; s++)
            out.stroke_sub_off.push_back(fp->sub_off[(size_t)s] + base);
        out.stroke_sub_job.insert(out.stroke_sub_job.end(), (size_t)n_sub, job);
        out.stroke_tag.insert(out.stroke_tag.end(), fp->tag.begin(), fp->tag.end());
        out.stroke_data.insert(out.stroke_data.end(), fp->data.begin(), fp->data.end());
        out.stroke_seg_job.insert(out.stroke_seg_job.end(), fp->tag.size(), job);
        StrokeRec r;
        memset(&r, 0, sizeof r);
        r.half_width = width / 2, r.sub_begin = sub_begin, r.sub_end = sub_begin + n_sub, r.cap = cap, r.join = join, r.path = pid;
        out.strokes.push_back(r);
        return pid;
    }

    int paint_record(PaintRec &rec)
    {
        out.paints.push_back(rec);
        return (int)out.paints.size() - 1;
    }

    // Path.fill / Path.mask on path `pid` (svgrasterize.py:995-1103; encode.py _leaf)
    int leaf(int pid, int paint, const M23 &t, bool mask_only, bool linear_rgb)
    {
        if (mask_only)
            return node(SVGR_N_LEAF, pid, -1, 1, -1);
        if (paint < 0)
            return empty();
        if (paint >= in.n_paint) {
            fail(SVGR_E_INVALID, "flat scene: paint index out of range");
            return -1;
        }
        const svgr_flat_paint &p = in.paints[paint];
        PaintRec rec;
        memset(&rec, 0, sizeof rec);
        rec.pat_node = -1;
        if (p.kind == 1) {
            double c[4];
            if (linear_rgb)
                memcpy(c, p.p, sizeof c);
            else
                paint_to_srgb(p.p, c);
            rec.kind = PAINT_SOLID;
            for (int k = 0; k < 4; k++)
                rec.color[k] = (float)c[k];
            return node(SVGR_N_LEAF, pid, paint_record(rec), linear_rgb ? 1 : 0, -1);
        }
        if (p.kind != 2 && p.kind != 3) {
            fail(SVGR_E_UNSUPPORTED, "flat scene: paint kind needs the Python encoder");
            return -1;
        }
        if (p.spread < 0 || p.spread > 2) {
            fail(SVGR_E_INVALID, "invalid spread method");
            return -1;
        }
        const bool lin = p.lin < 0 ? linear_rgb : (p.lin != 0);
        // stops in the target colour space (grad_stops_colorspace, :1686) + the reciprocal spans
        if (p.stop_cnt < 1 || p.stop_off < 0 || p.stop_off + p.stop_cnt > in.n_stop) {
            fail(SVGR_E_INVALID, "gradient without stops");
            return -1;
        }
        rec.spread = p.spread, rec.stop_off = (int32_t)out.stops.size(), rec.stop_cnt = p.stop_cnt;
        for (int k = 0; k < p.stop_cnt; k++) {
            const svgr_flat_stop &s = in.stops[p.stop_off + k];
            double c[4];
            if (lin)
                memcpy(c, s.color, sizeof c);
            else
                paint_to_srgb(s.color, c);
            StopRec sr;
            memset(&sr, 0, sizeof sr);
            sr.offset = s.offset;
            for (int q = 0; q < 4; q++)
                sr.color[q] = (float)c[q];
            sr.inv_span = k + 1 < p.stop_cnt ? 1.0 / (in.stops[p.stop_off + k + 1].offset - s.offset) : 0.0;
            out.stops.push_back(sr);
        }
        // the gradient's own geometry: what does not depend on the transform
        double geom[6] = {0, 0, 0, 0, 0, 0};
        if (p.kind == 2) {
            const double v0 = p.p[2] - p.p[0], v1 = p.p[3] - p.p[1];
            rec.kind = PAINT_LINEAR;
            geom[0] = p.p[0], geom[1] = p.p[1], geom[2] = v0, geom[3] = v1, geom[4] = fma(v1, v1, v0 * v0);
        } else if (!(p.focal & 3)) {
            rec.kind = PAINT_RADIAL;
            geom[0] = p.p[0], geom[1] = p.p[1], geom[2] = p.p[2];
        } else {
            const double cx = p.p[0], cy = p.p[1], r = p.p[2];
            const double fx = (p.focal & 1) ? p.p[3] : cx, fy = (p.focal & 1) ? p.p[4] : cy;
            const double fr = (p.focal & 2) ? p.p[5] : 0.0;
            const double cd0 = cx - fx, cd1 = cy - fy, rd = r - fr;
            // encode.py: (cd ** 2).sum() - rd ** 2 -- numpy squares by multiplying, Python's float ** 2 is libm's pow,
            // which is not always the rounded product (the exponent is volatile so that the compiler keeps the call)
            volatile double two = 2.0;
            const double a = (cd0 * cd0 + cd1 * cd1) - pow(rd, two);
            rec.kind = PAINT_RADIAL_FOCAL;
            geom[0] = fx, geom[1] = fy;
            rec.g[0] = cd0, rec.g[1] = cd1, rec.g[2] = fr * rd, rec.g[3] = a, rec.g[4] = fr * fr;
            rec.g[5] = fr / (fr - r), rec.g[6] = fr != r ? 1.0 : 0.0, rec.g[7] = 1.0 / a;
            rec.flag = out.n_focal++;
        }
        // pixel centre -> gradient space: transform.invert, then the inverse gradientTransform (:1022-1031, :1558)
        M23 to_user;
        if (!invert(t, to_user)) {
            fail(SVGR_E_INVALID, "Singular matrix");
            return -1;
        }
        if (p.bbox_units) {
            // the map needs ConvexHull.bbox_transform of the flattened leaf (:1023-1026): the render completes it
            svgr_bbox_job job;
            memset(&job, 0, sizeof job);
            job.path = pid, job.has_grad_tr = p.has_transform ? 1 : 0;
            memcpy(job.inv, to_user.m, sizeof job.inv);
            if (p.has_transform)
                memcpy(job.grad_inv, p.inv, sizeof job.grad_inv);
            memcpy(job.geom, geom, sizeof job.geom);
            job.paint = paint_record(rec);
            out.bbox_jobs.push_back(job);
            return node(SVGR_N_LEAF, pid, job.paint, lin ? 1 : 0, -1);
        }
        if (p.has_transform) {
            M23 gi;
            memcpy(gi.m, p.inv, sizeof gi.m);
            to_user = matmul(gi, to_user);
        }
        gradient_coefficients(rec.kind, geom, to_user, rec.g, rec.m1);
        return node(SVGR_N_LEAF, pid, paint_record(rec), lin ? 1 : 0, -1);
    }

    // Transform.__call__ on rows of points (svgrasterize.py:531-534): points @ m[:2, :2].T + m[:2, 2]
    static void apply(const M23 &m, double x, double y, double &ox, double &oy)
    {
        ox = fma(y, m.m[1], x * m.m[0]) + m.m[2];
        oy = fma(y, m.m[4], x * m.m[3]) + m.m[5];
    }

    // numpy's pairwise summation of a contiguous run (np.add.reduce): blocks of 8 accumulators below 128 elements
    static double pairwise_sum(const double *a, size_t n, size_t stride)
    {
        if (n < 8) {
            double r = 0.0;  // numpy starts from the first element; adding it to -0.0 / 0.0 is the same value here
            for (size_t i = 0; i < n; i++)
                r = i == 0 ? a[0] : r + a[i * stride];
            return r;
        }
        if (n <= 128) {
            double r[8];
            for (int k = 0; k < 8; k++)
                r[k] = a[k * stride];
            size_t i = 8;
            for (; i + 8 <= n; i += 8)
                for (int k = 0; k < 8; k++)
                    r[k] += a[(i + k) * stride];
            double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
            for (; i < n; i++)
                res += a[i * stride];
            return res;
        }
        size_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2, stride) + pairwise_sum(a + n2 * stride, n - n2, stride);
    }

    // blur_kernel (svgrasterize.py:1903-1944; encode.py blur_kernel): the Gaussian of feGaussianBlur under `t`.
    // false: the blur is a no-op (< half a pixel both ways).  rows x cols weights, row-major.
    bool blur_kernel(const M23 &t, double sx, double sy, int &rows, int &cols, std::vector<double> &w)
    {
        double ox, oy;
        apply(t, 0.0, 0.0, ox, oy);
        // basis = transform(np.eye(2)) - origin; scale = np.linalg.norm(basis, axis=1)
        double b00, b01, b10, b11;
        apply(t, 1.0, 0.0, b00, b01);
        apply(t, 0.0, 1.0, b10, b11);
        b00 -= ox, b01 -= oy, b10 -= ox, b11 -= oy;
        const double scale_x = sqrt(b00 * b00 + b01 * b01), scale_y = sqrt(b10 * b10 + b11 * b11);
        if (scale_x * sx < 0.5 && scale_y * sy < 0.5)
            return false;
        if (scale_x * sx < 0.5)
            sx = 0.5 / scale_x;
        else if (scale_y * sy < 0.5)
            sy = 0.5 / scale_y;
        const double reach = 2.5;
        const double bx[4] = {-reach * sx, -reach * sx, reach * sx, reach * sx};
        const double by[4] = {-reach * sy, reach * sy, reach * sy, -reach * sy};
        double lo0 = 0, lo1 = 0, hi0 = 0, hi1 = 0;
        for (int k = 0; k < 4; k++) {
            double px, py;
            apply(t, bx[k], by[k], px, py);
            px -= ox, py -= oy;
            if (k == 0 || px < lo0) lo0 = px;
            if (k == 0 || px > hi0) hi0 = px;
            if (k == 0 || py < lo1) lo1 = py;
            if (k == 0 || py > hi1) hi1 = py;
        }
        // .astype(int): truncation toward zero
        const long long l0 = (long long)lo0, l1 = (long long)lo1, h0 = (long long)hi0, h1 = (long long)hi1;
        long long kw = h0 - l0, kh = h1 - l1;
        kw += 1 - (kw & 1), kh += 1 - (kh & 1);
        if (kw <= 0 || kh <= 0 || kw > 1 << 15 || kh > 1 << 15 || kw * kh > (1ll << 26)) {
            fail(SVGR_E_UNSUPPORTED, "blur kernel too large");
            return false;
        }
        M23 inv;
        if (!invert(t, inv)) {
            fail(SVGR_E_INVALID, "Singular matrix");
            return false;
        }
        double zx, zy;
        apply(inv, 0.0, 0.0, zx, zy);
        const double cx = -(double)kw / 2 + 0.5, cy = -(double)kh / 2 + 0.5;
        const double dx = 2 * (sx * sx), dy = 2 * (sy * sy);
        rows = (int)kw, cols = (int)kh;
        w.resize((size_t)rows * cols);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++) {
                double px, py;
                apply(inv, (double)r + cx, (double)c + cy, px, py);
                px -= zx, py -= zy;
                w[(size_t)r * cols + c] = exp(-(px * px) / dx) * exp(-(py * py) / dy);
            }
        const double total = pairwise_sum(w.data(), w.size(), 1);
        for (double &v : w)
            v = v / total;
        return true;
    }

    // encode.py _compose: Layer.compose as a node (mode code, or 5 + the arithmetic coefficients)
    int compose(const int32_t *kids, int n_kids, int mode, const double *k4)
    {
        if (k4)
            return node(SVGR_N_COMPOSE, 5, 0, 0, 0, kids, n_kids, 1, k4[0], k4[1], k4[2], k4[3]);
        return node(SVGR_N_COMPOSE, mode, 0, 0, 0, kids, n_kids, 1);
    }

    // Filter.__call__ (svgrasterize.py:1801-1831; encode.py _filter) lowered to nodes
    int filter(int first, int count, const M23 &t, int source)
    {
        std::vector<int> stack;
        {
            const int32_t c[1] = {source};
            stack.push_back(node(SVGR_N_SRC_ALPHA, 0, 0, 0, 0, c, 1));
            stack.push_back(node(SVGR_N_CONVERT, 0, 1, 0, 0, c, 1));
        }
        for (int i = first; i < first + count; i++) {
            const svgr_flat_fe &fe = in.fes[i];
            if (fe.n_in < 0 || fe.in_off < 0 || (int64_t)fe.in_off + fe.n_in > in.n_fe_input) {
                fail(SVGR_E_INVALID, "flat scene: filter inputs outside the table");
                return -1;
            }
            std::vector<int32_t> args;
            for (int k = 0; k < fe.n_in; k++) {
                const int slot = in.fe_inputs[fe.in_off + k];
                if (slot < 0 || slot >= (int)stack.size()) {
                    fail(SVGR_E_INVALID, "list index out of range");  // the reference's IndexError on stack[i]
                    return -1;
                }
                args.push_back(stack[(size_t)slot]);
            }
            auto need = [&](int n) {
                if ((int)args.size() < n) {
                    fail(SVGR_E_INVALID, "filter primitive: missing input");
                    return false;
                }
                return true;
            };
            int outn = -1;
            switch (fe.tag) {
            case 11: {  // feOffset
                if (!need(1))
                    return -1;
                M23 inv;
                if (!invert(t, inv)) {
                    fail(SVGR_E_INVALID, "Singular matrix");
                    return -1;
                }
                out.offset_tr.insert(out.offset_tr.end(), t.m, t.m + 6);
                out.offset_tr.insert(out.offset_tr.end(), inv.m, inv.m + 6);
                outn = node(SVGR_N_OFFSET, (int)(out.offset_tr.size() / 12) - 1, 0, 0, 0, args.data(), 1, 0, fe.a[0], fe.a[1]);
                break;
            }
            case 9:  // feMerge
                outn = compose(args.data(), (int)args.size(), 0, nullptr);
                break;
            case 0: {  // feBlend: a plain over-merge (:1877)
                if (!need(2))
                    return -1;
                const int32_t c[2] = {args[1], args[0]};
                outn = compose(c, 2, 0, nullptr);
                break;
            }
            case 3: {  // feComposite
                if (!need(2))
                    return -1;
                if (fe.flag < 0) {
                    fail(SVGR_E_INVALID, "invalid compose mode");
                    return -1;
                }
                const int32_t c[2] = {args[1], args[0]};
                outn = fe.flag == 1 ? compose(c, 2, 5, fe.a) : compose(c, 2, (int)fe.a[0], nullptr);
                break;
            }
            case 8: {  // feGaussianBlur
                if (!need(1))
                    return -1;
                int rows = 0, cols = 0;
                std::vector<double> w;
                if (!blur_kernel(t, fe.a[0], fe.a[1], rows, cols, w)) {
                    if (out.err_code != SVGR_OK)
                        return -1;
                    outn = args[0];
                    break;
                }
                // encode.py _blur: separable when the outer product of the marginals reproduces the kernel
                std::vector<double> a((size_t)rows), b((size_t)cols);
                for (int r = 0; r < rows; r++)
                    a[(size_t)r] = pairwise_sum(w.data() + (size_t)r * cols, (size_t)cols, 1);
                for (int c = 0; c < cols; c++) {
                    // numpy reduces axis 0 of a C-contiguous array row by row: a plain running sum per column
                    double acc = w[(size_t)c];
                    for (int r = 1; r < rows; r++)
                        acc += w[(size_t)r * cols + c];
                    b[(size_t)c] = acc;
                }
                double worst = 0.0;
                for (int r = 0; r < rows; r++)
                    for (int c = 0; c < cols; c++)
                        worst = fmax(worst, fabs(a[(size_t)r] * b[(size_t)c] - w[(size_t)r * cols + c]));
                svgr_kernel k;
                k.rows = rows, k.cols = cols, k.separable = worst < 1e-12 ? 1 : 0, k.weight_off = (int32_t)out.weights.size();
                if (k.separable) {
                    for (double v : a)
                        out.weights.push_back((float)v);
                    for (double v : b)
                        out.weights.push_back((float)v);
                } else {
                    for (double v : w)
                        out.weights.push_back((float)v);
                }
                out.kernels.push_back(k);
                outn = node(SVGR_N_BLUR, (int)out.kernels.size() - 1, 0, 0, 0, args.data(), 1);
                break;
            }
            case 1: {  // feColorMatrix
                if (!need(1))
                    return -1;
                if (fe.flag != 1) {
                    outn = args[0];  // "invalid color matrix" (:1857): the primitive is skipped
                    break;
                }
                for (int k = 0; k < 20; k++)
                    out.matrices.push_back((float)fe.a[k]);
                outn = node(SVGR_N_CMATRIX, (int)(out.matrices.size() / 20) - 1, 0, 0, 0, args.data(), 1);
                break;
            }
            case 10: {  // feMorphology
                if (!need(1))
                    return -1;
                // u = transform([[rx, 0], [0, ry]]) - transform(zeros): the window is twice the length of its rows
                double zx, zy, u00, u01, u10, u11;
                apply(t, 0.0, 0.0, zx, zy);
                apply(t, fe.a[0], 0.0, u00, u01);
                apply(t, 0.0, fe.a[1], u10, u11);
                u00 -= zx, u01 -= zy, u10 -= zx, u11 -= zy;
                const int k0 = (int)(sqrt(fma(u01, u01, u00 * u00)) * 2), k1 = (int)(sqrt(fma(u11, u11, u10 * u10)) * 2);
                if (k0 < 1 || k1 < 1) {
                    outn = args[0];
                } else {
                    if (fe.flag < 0) {
                        fail(SVGR_E_INVALID, "invalid poll method");
                        return -1;
                    }
                    outn = node(SVGR_N_MORPH, k0, k1, fe.flag, 0, args.data(), 1);
                }
                break;
            }
            default:
                fail(SVGR_E_INVALID, "unsupported filter type");
                return -1;
            }
            stack.push_back(outn);
        }
        return stack.back();
    }

    // Scene.render (svgrasterize.py:649-752; encode.py Encoder.encode)
    int encode(int ni, const M23 &t, bool mask_only, const int32_t *viewport, bool linear_rgb, int depth = 0)
    {
        if (out.err_code != SVGR_OK)
            return -1;
        if (ni < 0 || ni >= in.n_node || depth > 4096) {
            fail(SVGR_E_INVALID, "flat scene: node index out of range / tree too deep");
            return -1;
        }
        const svgr_flat_node &n = in.nodes[ni];
        if (n.child_off < 0 || n.child_cnt < 0 || (int64_t)n.child_off + n.child_cnt > in.n_child) {
            fail(SVGR_E_INVALID, "flat scene: child range outside the children table");
            return -1;
        }
        const int32_t *kids = in.children + n.child_off;
        const int lin = linear_rgb ? 1 : 0;
        switch (n.tag) {
        case R_FILL: {
            if (!mask_only && n.b < 0)
                return empty();
            if (n.c < 0 || n.c > 1) {
                fail(SVGR_E_INVALID, "Invalid fill rule");
                return -1;
            }
            const int pid = add_fill_path(n.a, t, n.c, viewport);
            return pid < 0 ? -1 : leaf(pid, n.b, t, mask_only, linear_rgb);
        }
        case R_STROKE: {
            if (n.c < 0) {
                fail(SVGR_E_INVALID, "unkown line cap type");
                return -1;
            }
            if (!mask_only && n.b < 0)
                return empty();
            const int pid = add_stroke_path(n.a, t, n.f[0], n.c, n.d, viewport);
            return pid < 0 ? -1 : leaf(pid, n.b, t, mask_only, linear_rgb);
        }
        case R_GROUP: {
            const size_t base = kid_stack.size();
            for (int k = 0; k < n.child_cnt; k++) {
                const int kid = encode(kids[k], t, mask_only, viewport, linear_rgb, depth + 1);
                if (kid < 0)
                    return -1;
                if (!is_empty(kid))
                    kid_stack.push_back(kid);
            }
            const int cnt = (int)(kid_stack.size() - base);
            int res;
            if (cnt == 0)
                res = empty();
            else if (cnt == 1)
                res = kid_stack[base];
            else
                res = node(SVGR_N_GROUP, 0, 0, 0, 0, kid_stack.data() + base, cnt, lin);
            kid_stack.resize(base);
            return res;
        }
        case R_OPACITY: {
            if (n.child_cnt != 1)
                break;
            const int tn = encode(kids[0], t, mask_only, viewport, linear_rgb, depth + 1);
            if (tn < 0 || is_empty(tn))
                return tn;
            const int32_t c[1] = {tn};
            return node(SVGR_N_OPACITY, 0, 0, 0, 0, c, 1, lin, n.f[0]);
        }
        case R_TRANSFORM: {
            if (n.child_cnt != 1 || n.a < 0 || n.a >= in.n_tr)
                break;
            M23 tr;
            memcpy(tr.m, in.tr + 6 * (size_t)n.a, sizeof tr.m);
            return encode(kids[0], matmul(t, tr), mask_only, viewport, linear_rgb, depth + 1);
        }
        case R_CLIP:
        case R_MASK: {
            if (n.child_cnt != 2)
                break;
            const int tn = encode(kids[0], t, mask_only, viewport, linear_rgb, depth + 1);
            if (tn < 0 || is_empty(tn))
                return tn;
            if (n.a) {
                fail(SVGR_E_UNSUPPORTED, "flat scene: objectBoundingBox clip / mask needs the Python encoder");
                return -1;
            }
            int stencil;
            if (n.tag == R_CLIP) {
                stencil = encode(kids[1], t, true, viewport, linear_rgb, depth + 1);
                if (stencil < 0 || is_empty(stencil))
                    return stencil;
            } else {
                const int sub = encode(kids[1], t, mask_only, viewport, linear_rgb, depth + 1);
                if (sub < 0 || is_empty(sub))
                    return sub;
                const int32_t c[1] = {sub};
                stencil = node(SVGR_N_LUMA, 0, 0, 0, 0, c, 1, lin);
            }
            const int32_t c[2] = {stencil, tn};
            return node(SVGR_N_IN, 0, 0, 0, 0, c, 2, lin);
        }
        case R_FILTER: {
            if (n.child_cnt != 1 || n.a < 0 || n.b < 0 || (int64_t)n.a + n.b > in.n_fe)
                break;
            const int tn = encode(kids[0], t, mask_only, viewport, linear_rgb, depth + 1);
            if (tn < 0 || is_empty(tn))
                return tn;
            return filter(n.a, n.b, t, tn);
        }
        default:
            fail(SVGR_E_INVALID, "unhandled scene type");
            return -1;
        }
        fail(SVGR_E_INVALID, "flat scene: malformed node");
        return -1;
    }

    // main() of the reference (svgrasterize.py:3854-3881; encode.py add_scene)
    bool add_scene(const svgr_flat_scene &sc)
    {
        const M23 swap = {{0.0, 1.0, 0.0, 1.0, 0.0, 0.0}};  // the x / y swap every render starts from (:246, :3823)
        const int32_t viewport[4] = {0, 0, sc.height, sc.width};
        const int root = encode(sc.root, swap, false, viewport, sc.linear_rgb != 0);
        if (root < 0)
            return false;
        int nd;
        if (is_empty(root)) {
            nd = node(SVGR_N_CANVAS, sc.height, sc.width, 0, 0, nullptr, 0, sc.linear_rgb ? 1 : 0, (double)out.canvas_bytes);
        } else {
            const int32_t c[1] = {root};
            nd = node(SVGR_N_CANVAS, sc.height, sc.width, 0, 0, c, 1, sc.linear_rgb ? 1 : 0, (double)out.canvas_bytes);
        }
        out.canvases.push_back(nd), out.canvases.push_back(out.canvas_bytes);
        out.canvases.push_back(sc.height), out.canvases.push_back(sc.width);
        out.canvas_bytes += 4ll * sc.height * sc.width;
        out.roots.push_back(root);
        return true;
    }
};

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int svgr_encode_flat(const svgr_flat *in, svgr_encoded **out_handle)
{
    if (!in || !out_handle)
        return SVGR_E_INVALID;
    svgr_encoded *e = new svgr_encoded();
    *out_handle = e;
    if (in->n_path < 0 || in->n_sub < 0 || in->n_seg < 0 || in->n_node < 0 || in->n_scene < 0 || in->n_paint < 0 ||
        in->n_stop < 0 || in->n_child < 0 || in->n_tr < 0) {
        e->err_code = SVGR_E_INVALID, e->err = "flat scene: negative count";
        return e->err_code;
    }
    for (int p = 0; p < in->n_path; p++)
        if (in->path_off[p] < 0 || in->path_off[p] > in->path_off[p + 1] || in->path_off[p + 1] > in->n_sub) {
            e->err_code = SVGR_E_INVALID, e->err = "flat scene: path offsets outside the sub-path table";
            return e->err_code;
        }
    for (int s = 0; s < in->n_sub; s++)
        if (in->sub_off[s] < 0 || in->sub_off[s] > in->sub_off[s + 1] || in->sub_off[s + 1] > in->n_seg) {
            e->err_code = SVGR_E_INVALID, e->err = "flat scene: sub-path offsets outside the segments";
            return e->err_code;
        }
    Encoder enc(*in, *e);
    for (int s = 0; s < in->n_scene; s++)
        if (!enc.add_scene(in->scenes[s]))
            return e->err_code;
    svgr_program &p = e->prog;
    memset(&p, 0, sizeof p);
    p.n_seg = (int64_t)e->seg_tag.size(), p.seg_tag = e->seg_tag.data(), p.seg_data = e->seg_data.data(), p.seg_path = e->seg_path.data();
    p.n_path = (int32_t)e->paths.size(), p.paths = e->paths.data();
    p.n_stroke = (int32_t)e->strokes.size(), p.strokes = e->strokes.data();
    p.n_stroke_sub = (int32_t)e->stroke_sub_job.size(), p.stroke_sub_off = e->stroke_sub_off.data(), p.stroke_sub_job = e->stroke_sub_job.data();
    p.n_stroke_seg = (int64_t)e->stroke_tag.size(), p.stroke_tag = e->stroke_tag.data(), p.stroke_data = e->stroke_data.data();
    p.stroke_seg_job = e->stroke_seg_job.data();
    p.n_paint = (int32_t)e->paints.size(), p.paints = e->paints.data();
    p.n_stop = (int32_t)e->stops.size(), p.stops = e->stops.data();
    p.n_focal = e->n_focal;
    p.n_node = (int32_t)e->nodes.size(), p.nodes = e->nodes.data();
    p.n_child = (int32_t)e->children.size(), p.children = e->children.data();
    p.canvas_bytes = e->canvas_bytes;
    p.n_bbox_job = (int32_t)e->bbox_jobs.size(), p.bbox_jobs = e->bbox_jobs.data();
    p.n_kernel = (int32_t)e->kernels.size(), p.kernels = e->kernels.data();
    p.n_weight = (int32_t)e->weights.size(), p.weights = e->weights.data();
    p.n_matrix = (int32_t)(e->matrices.size() / 20), p.matrices = e->matrices.data();
    p.n_offset_tr = (int32_t)(e->offset_tr.size() / 12), p.offset_tr = e->offset_tr.data();
    return SVGR_OK;
}

const svgr_program *svgr_encoded_program(const svgr_encoded *e) { return e ? &e->prog : nullptr; }
const char *svgr_encoded_error(const svgr_encoded *e) { return e ? e->err.c_str() : "null handle"; }
/* canvases: 4 int64 per canvas (node, byte offset, rows, cols); roots: one node index per scene */
int64_t svgr_encoded_canvases(const svgr_encoded *e, const int64_t **canvases, const int32_t **roots)
{
    if (!e)
        return 0;
    if (canvases)
        *canvases = e->canvases.data();
    if (roots)
        *roots = e->roots.data();
    return (int64_t)e->roots.size();
}
void svgr_encoded_free(svgr_encoded *e) { delete e; }

#pragma GCC visibility pop
}
