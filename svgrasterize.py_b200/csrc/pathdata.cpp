// pathdata.cpp -- the SVG path-data reader in native code.
//
// Path.from_svg (svgrasterize.py:1252-1430): tokenises a `d` attribute / glyph outline string and builds the
// sub-path list.  Here the result goes straight into the flat segment arrays the encoder consumes (tags,
// 8 doubles per segment, sub-path offsets), with the reference's exact behaviour: its number grammar (FLOAT_RE,
// :41), whitespace / command sets, relative-coordinate rules, smooth-curve reflection, the `A` command's
// conversion to the parametric arc (arc_svg_to_parametric, :2397-2450) and its quirks -- arguments before the
// first command are dropped, a zero-radius arc becomes the degenerate line [dst, dst] (:1373-1376), a path that
// ends without `z` gets a PATH_UNCLOSED segment back to its start -- and its ValueErrors.
//
// Arithmetic of the arc conversion follows what numpy / CPython compute, measured on this image: matmul(M, v) =
// fma(M[k,0], v0, M[k,1] v1); matmul(M.T, v) = fma(M[1,k], v1, M[0,k] v0); x ** 2 = libm pow(x, 2) for Python floats
// and numpy scalars alike; np.dot of two 2-vectors = fma(a1, b1, a0 b0); np.cross = a0 b1 - a1 b0; np.linalg.norm =
// sqrt(dot(v, v)).  tests/test_host_logic.py checks the segments bit for bit against vectors recorded from the
// unmodified reference (tests/golden_eager/pathdata.npz).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/svgr_b200.h"
#include "svgr_types.h"

namespace {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == ','; }
inline bool is_command(char c) { return c != 0 && strchr("MmZzLlHhVvCcSsQqTtAa", c) != nullptr; }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// FLOAT_RE = [-+]?(?:(?:\d*\.\d+)|(?:\d+\.?))(?:[Ee][+-]?\d+)?  -> length of the match at s (0: no match)
size_t match_float(const char *s, size_t n)
{
    size_t i = 0;
    if (i < n && (s[i] == '-' || s[i] == '+'))
        i++;
    size_t d0 = i;
    while (i < n && is_digit(s[i]))
        i++;
    const size_t int_digits = i - d0;
    size_t end = 0;
    if (i < n && s[i] == '.') {
        size_t j = i + 1;
        while (j < n && is_digit(s[j]))
            j++;
        if (j > i + 1)
            end = j;  // \d*\.\d+
        else if (int_digits > 0)
            end = i + 1;  // \d+\.
    } else if (int_digits > 0) {
        end = i;  // \d+
    }
    if (end == 0)
        return 0;
    // optional exponent: only taken when complete
    if (end < n && (s[end] == 'e' || s[end] == 'E')) {
        size_t j = end + 1;
        if (j < n && (s[j] == '-' || s[j] == '+'))
            j++;
        size_t k = j;
        while (k < n && is_digit(s[k]))
            k++;
        if (k > j)
            end = k;
    }
    return end;
}

double libm_sq(double x)
{
    volatile double two = 2.0;  // Python's and numpy's `** 2` call pow(); keep the call (it is not always x * x)
    return pow(x, two);
}

double angle_between(const double *v0, const double *v1)
{
    // svgrasterize.py:2472-2478
    const double dot = fma(v0[1], v1[1], v0[0] * v1[0]);
    const double n0 = sqrt(fma(v0[1], v0[1], v0[0] * v0[0])), n1 = sqrt(fma(v1[1], v1[1], v1[0] * v1[0]));
    double c = dot / (n0 * n1);
    c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);  // np.clip (NaN passes through)
    double angle = acos(c);
    if (v0[0] * v1[1] - v0[1] * v1[0] < 0)
        angle = -angle;
    return angle;
}

// arc_svg_to_parametric (svgrasterize.py:2397-2450) -> row = cx, cy, rx, ry, phi, eta, eta_delta
void arc_to_parametric(const double *src, const double *dst, double rx, double ry, double x_axis_rot, bool large, bool sweep,
                       double *row)
{
    rx = fabs(rx), ry = fabs(ry);
    const double phi = x_axis_rot * M_PI / 180;
    const double cp = cos(phi), sp = sin(phi);
    const double M[2][2] = {{cp, sp}, {-sp, cp}};
    const double h0 = (src[0] - dst[0]) / 2, h1 = (src[1] - dst[1]) / 2;
    const double x1 = fma(M[0][0], h0, M[0][1] * h1), y1 = fma(M[1][0], h0, M[1][1] * h1);
    double s = libm_sq(x1 / rx) + libm_sq(y1 / ry);
    if (s > 1) {
        s = sqrt(s);
        rx *= s, ry *= s;
    }
    const double q = libm_sq(rx * ry) / (libm_sq(rx * y1) + libm_sq(ry * x1)) - 1;
    double sq = sqrt(q > 0 ? q : 0.0);  // max(0, nan) is 0 in Python: (0 if nan > 0 ...) -- max keeps its first argument
    if (large == sweep)
        sq = -sq;
    const double cx = sq * (rx * y1 / ry), cy = sq * (-ry * x1 / rx);
    // np.matmul(M.T, center) + (dst + src) / 2
    const double c0 = fma(M[1][0], cy, M[0][0] * cx) + (dst[0] + src[0]) / 2;
    const double c1 = fma(M[1][1], cy, M[0][1] * cx) + (dst[1] + src[1]) / 2;
    const double v0[2] = {1.0, 0.0};
    const double v1[2] = {(x1 - cx) / rx, (y1 - cy) / ry};
    const double v2[2] = {(-x1 - cx) / rx, (-y1 - cy) / ry};
    const double eta = angle_between(v0, v1);
    double eta_delta = fmod(angle_between(v1, v2), 2 * M_PI);
    if (!sweep && eta_delta > 0)
        eta_delta -= 2 * M_PI;
    if (sweep && eta_delta < 0)
        eta_delta += 2 * M_PI;
    row[0] = c0, row[1] = c1, row[2] = rx, row[3] = ry, row[4] = phi, row[5] = eta, row[6] = eta_delta, row[7] = 0.0;
}

struct Reader {
    std::vector<uint8_t> tags;
    std::vector<double> data;
    std::vector<int32_t> sub_off;
    std::string err;
    size_t sub_start = 0;  // first segment of the sub-path under construction

    void seg(uint8_t tag, const double *pts, int n_doubles)
    {
        tags.push_back(tag);
        const size_t at = data.size();
        data.resize(at + 8, 0.0);
        memcpy(data.data() + at, pts, (size_t)n_doubles * 8);
    }
    void seg2(uint8_t tag, const double *a, const double *b)
    {
        const double p[4] = {a[0], a[1], b[0], b[1]};
        seg(tag, p, 4);
    }
    bool open() const { return tags.size() > sub_start; }
    void end_sub()
    {
        sub_off.push_back((int32_t)tags.size());
        sub_start = tags.size();
    }
    bool fail(const std::string &m)
    {
        err = m;
        return false;
    }
    // chunk(vs, size) (:1533-1538): ValueError unless vs splits into at least one full chunk
    bool chunkable(const std::vector<double> &a, size_t size)
    {
        if (a.empty() || a.size() % size != 0)
            return fail("list can not be chunked in " + std::to_string(size) + "s");
        return true;
    }

    bool run(const char *in, size_t len)
    {
        sub_off.push_back(0);
        std::vector<double> args;
        char cmd = 0;
        double pos[2] = {0.0, 0.0}, start[2] = {0.0, 0.0};
        bool has_sc = false, has_sq = false;
        double sc[2] = {0, 0}, sqd[2] = {0, 0};  // reflected control points of the last C/S and Q/T
        size_t off = 0;
        while (off <= len) {
            const char ch = off < len ? in[off] : 0;
            if (off < len && is_space(ch)) {
                off++;
            } else if (off == len || is_command(ch)) {
                std::vector<double> a;
                a.swap(args);
                auto rel = [&](bool r, const double *d, double *o) { o[0] = r ? pos[0] + d[0] : d[0], o[1] = r ? pos[1] + d[1] : d[1]; };
                if (cmd == 0) {
                    // arguments before the first command are dropped (:1292-1293)
                } else if (cmd == 'M' || cmd == 'm') {
                    if (open()) {
                        seg2(SEG_UNCLOSED, pos, start);
                        end_sub();
                    }
                    if (!chunkable(a, 2))
                        return false;
                    // `first` is already False when the first command is processed (:1417), so "m" is relative
                    const bool r = cmd == 'm';
                    double p[2];
                    rel(r, &a[0], p);
                    pos[0] = start[0] = p[0], pos[1] = start[1] = p[1];
                    for (size_t i = 2; i < a.size(); i += 2) {
                        double d[2];
                        rel(r, &a[i], d);
                        seg2(SEG_LINE, pos, d);
                        pos[0] = d[0], pos[1] = d[1];
                    }
                } else if (cmd == 'L' || cmd == 'l') {
                    if (!chunkable(a, 2))
                        return false;
                    for (size_t i = 0; i < a.size(); i += 2) {
                        double d[2];
                        rel(cmd == 'l', &a[i], d);
                        seg2(SEG_LINE, pos, d);
                        pos[0] = d[0], pos[1] = d[1];
                    }
                } else if (cmd == 'V' || cmd == 'v' || cmd == 'H' || cmd == 'h') {
                    if (a.empty())
                        return fail(std::string("command '") + cmd + "' expects at least one argument");
                    const bool r = cmd == 'v' || cmd == 'h', vert = cmd == 'V' || cmd == 'v';
                    for (double v : a) {
                        // position(is_relative, pos, [0 if is_relative else pos[0], dst]) (:1321, :1330)
                        double raw[2], d[2];
                        if (vert)
                            raw[0] = r ? 0.0 : pos[0], raw[1] = v;
                        else
                            raw[0] = v, raw[1] = r ? 0.0 : pos[1];
                        rel(r, raw, d);
                        seg2(SEG_LINE, pos, d);
                        pos[0] = d[0], pos[1] = d[1];
                    }
                } else if (cmd == 'C' || cmd == 'c') {
                    if (!chunkable(a, 6))
                        return false;
                    for (size_t i = 0; i < a.size(); i += 6) {
                        double p[8] = {pos[0], pos[1], 0, 0, 0, 0, 0, 0};
                        for (int k = 0; k < 3; k++)
                            rel(cmd == 'c', &a[i + 2 * k], p + 2 + 2 * k);
                        seg(SEG_CUBIC, p, 8);
                        pos[0] = p[6], pos[1] = p[7];
                        sc[0] = p[6] * 2 - p[4], sc[1] = p[7] * 2 - p[5], has_sc = true;
                    }
                } else if (cmd == 'S' || cmd == 's') {
                    if (!chunkable(a, 4))
                        return false;
                    for (size_t i = 0; i < a.size(); i += 4) {
                        double p[8] = {pos[0], pos[1], 0, 0, 0, 0, 0, 0};
                        rel(cmd == 's', &a[i], p + 4);
                        rel(cmd == 's', &a[i + 2], p + 6);
                        if (!has_sc)
                            sc[0] = pos[0], sc[1] = pos[1], has_sc = true;
                        p[2] = sc[0], p[3] = sc[1];
                        seg(SEG_CUBIC, p, 8);
                        pos[0] = p[6], pos[1] = p[7];
                        sc[0] = p[6] * 2 - p[4], sc[1] = p[7] * 2 - p[5];
                    }
                } else if (cmd == 'Q' || cmd == 'q') {
                    if (!chunkable(a, 4))
                        return false;
                    for (size_t i = 0; i < a.size(); i += 4) {
                        double p[6] = {pos[0], pos[1], 0, 0, 0, 0};
                        rel(cmd == 'q', &a[i], p + 2);
                        rel(cmd == 'q', &a[i + 2], p + 4);
                        seg(SEG_QUAD, p, 6);
                        pos[0] = p[4], pos[1] = p[5];
                        sqd[0] = p[4] * 2 - p[2], sqd[1] = p[5] * 2 - p[3], has_sq = true;
                    }
                } else if (cmd == 'T' || cmd == 't') {
                    if (!chunkable(a, 2))
                        return false;
                    for (size_t i = 0; i < a.size(); i += 2) {
                        double p[6] = {pos[0], pos[1], 0, 0, 0, 0};
                        rel(cmd == 't', &a[i], p + 4);
                        if (!has_sq)
                            sqd[0] = pos[0], sqd[1] = pos[1], has_sq = true;
                        p[2] = sqd[0], p[3] = sqd[1];
                        seg(SEG_QUAD, p, 6);
                        pos[0] = p[4], pos[1] = p[5];
                        sqd[0] = p[4] * 2 - p[2], sqd[1] = p[5] * 2 - p[3];
                    }
                } else if (cmd == 'A' || cmd == 'a') {
                    if (!chunkable(a, 7))
                        return false;
                    for (size_t i = 0; i < a.size(); i += 7) {
                        double d[2], src[2] = {pos[0], pos[1]};
                        rel(cmd == 'a', &a[i + 5], d);
                        pos[0] = d[0], pos[1] = d[1];
                        if (a[i] == 0 || a[i + 1] == 0) {
                            seg2(SEG_LINE, pos, d);  // the reference has already moved `pos`: a degenerate line
                        } else {
                            double row[8];
                            arc_to_parametric(src, d, a[i], a[i + 1], a[i + 2], a[i + 3] > 0.001, a[i + 4] > 0.001, row);
                            seg(SEG_ARC, row, 8);
                        }
                    }
                } else if (cmd == 'Z' || cmd == 'z') {
                    if (!a.empty())
                        return fail("`z` command does not accept any argmuents");
                    seg2(SEG_CLOSED, pos, start);
                    end_sub();
                    pos[0] = start[0], pos[1] = start[1];
                }
                if (cmd != 0 && !strchr("CcSs", cmd))
                    has_sc = false;
                if (cmd != 0 && !strchr("QqTt", cmd))
                    has_sq = false;
                off++;
                cmd = ch;
            } else {
                const size_t m = match_float(in + off, len - off);
                if (m == 0)
                    return fail(std::string("not recognized command '") + ch + "' at: " + std::to_string(off));
                char buf[64];
                double v;
                if (m < sizeof buf) {
                    memcpy(buf, in + off, m);
                    buf[m] = 0;
                    v = strtod(buf, nullptr);
                } else {
                    std::string tmp(in + off, m);
                    v = strtod(tmp.c_str(), nullptr);
                }
                args.push_back(v);
                off += m;
            }
        }
        if (open()) {
            seg2(SEG_UNCLOSED, pos, start);
            end_sub();
        }
        return true;
    }
};

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

/* Path.from_svg (svgrasterize.py:1252-1430).  Writes at most cap_seg segments (tag + 8 doubles each) and cap_sub + 1
 * sub-path offsets; n_seg / n_sub receive the counts needed (call again with larger buffers when they exceed the
 * capacities: SVGR_E_NOMEM).  SVGR_E_INVALID = the reference raises ValueError (text in err, when err_cap > 0). */
int svgr_path_from_svg(const char *d, int64_t len, uint8_t *tags, double *data, int64_t cap_seg, int32_t *sub_off,
                       int64_t cap_sub, int64_t *n_seg, int64_t *n_sub, char *err, int32_t err_cap)
{
    if (!d || len < 0 || !n_seg || !n_sub)
        return SVGR_E_INVALID;
    Reader r;
    if (!r.run(d, (size_t)len)) {
        if (err && err_cap > 0)
            snprintf(err, (size_t)err_cap, "%s", r.err.c_str());
        return SVGR_E_INVALID;
    }
    *n_seg = (int64_t)r.tags.size();
    *n_sub = (int64_t)r.sub_off.size() - 1;
    if (*n_seg > cap_seg || *n_sub > cap_sub)
        return SVGR_E_NOMEM;
    if (*n_seg > 0) {
        memcpy(tags, r.tags.data(), r.tags.size());
        memcpy(data, r.data.data(), r.data.size() * 8);
    }
    memcpy(sub_off, r.sub_off.data(), r.sub_off.size() * 4);
    return SVGR_OK;
}

#pragma GCC visibility pop
}
