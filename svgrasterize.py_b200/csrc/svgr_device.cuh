// svgr_device.cuh -- device helpers shared by the compose and filter kernels:
// colour-space / alpha conversions (svgrasterize.py:471-503, Layer.convert :129-164),
// paint evaluation (svgrasterize.py:1553-1650, :1661-1683, pattern gather :1074-1094)
// and the source fetch that turns any layer-like value into an RGBA pixel.
#pragma once
#include "svgr_kernels.h"


__device__ __forceinline__ float4 f4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }

__device__ __forceinline__ float clip01(float v) { return __saturatef(v); }  // one instruction; NaN -> 0 like fmax/fmin

// Packed float32 pairs (Blackwell FFMA2 / FMUL2): an RGBA pixel is two instructions instead of four.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;}\n"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;}\n"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

// v * m
__device__ __forceinline__ float4 scale4(float4 v, float m)
{
    const float2 mm = make_float2(m, m);
    const float2 lo = mul2(make_float2(v.x, v.y), mm), hi = mul2(make_float2(v.z, v.w), mm);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// s + d * q
__device__ __forceinline__ float4 madd4(float4 d, float q, float4 s)
{
    const float2 qq = make_float2(q, q);
    const float2 lo = fma2(make_float2(d.x, d.y), qq, make_float2(s.x, s.y));
    const float2 hi = fma2(make_float2(d.z, d.w), qq, make_float2(s.z, s.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

// x^y for x in (0, 1], y a constant: exp2(y log2 x) on the special function unit.  __log2f is accurate to
// 2^-21.4 absolute near 1 and a few ulp elsewhere, so the result is within ~2e-6 relative of powf -- inside
// the 1e-5 layer tolerance and far inside +-1 LSB -- at a fifth of its instruction count.
__device__ __forceinline__ float pow_unit(float x, float y) { return exp2f(y * __log2f(x)); }

__device__ __forceinline__ float lin_to_srgb1(float v)
{
    return v <= 0.0031308f ? v * 12.92f : 1.055f * pow_unit(v, 1.0f / 2.4f) - 0.055f;
}

__device__ __forceinline__ float srgb_to_lin1(float v)
{
    return v <= 0.04045f ? v * (1.0f / 12.92f) : pow_unit((v + 0.055f) * (1.0f / 1.055f), 2.4f);
}

// color_pre_to_straight_alpha (svgrasterize.py:471-477): divide where alpha > 1e-4, clip all to [0, 1]
__device__ __forceinline__ float4 unpremultiply(float4 v)
{
    if (v.w > 0.0001f) {
        float inv = __frcp_rn(v.w);  // one reciprocal instead of three divides (<= 1 ulp apart)
        v.x = v.x * inv;
        v.y = v.y * inv;
        v.z = v.z * inv;
    }
    return f4(clip01(v.x), clip01(v.y), clip01(v.z), clip01(v.w));
}

__device__ __forceinline__ float4 premultiply(float4 v) { return f4(v.x * v.w, v.y * v.w, v.z * v.w, v.w); }

// Layer.convert (svgrasterize.py:129-164) on one RGBA pixel; code = SVGR_CONV(src pre, src lin, dst pre, dst lin)
__device__ __forceinline__ float4 convert_px(float4 v, int code)
{
    bool pre = code & 1, lin = code & 2, tpre = code & 4, tlin = code & 8;
    if (lin != tlin) {
        if (pre) {
            v = unpremultiply(v);
            pre = false;
        }
        if (tlin)
            v = f4(srgb_to_lin1(v.x), srgb_to_lin1(v.y), srgb_to_lin1(v.z), v.w);
        else
            v = f4(lin_to_srgb1(v.x), lin_to_srgb1(v.y), lin_to_srgb1(v.z), v.w);
    }
    if (pre != tpre)
        v = tpre ? premultiply(v) : unpremultiply(v);
    return v;
}

__device__ __forceinline__ bool conv_is_identity(int code) { return (code & 3) == ((code >> 2) & 3); }

// np.remainder for doubles (floored modulo)
__device__ __forceinline__ double floored_mod(double a, double b)
{
    double r = fmod(a, b);
    if (r != 0.0) {
        if ((r < 0.0) != (b < 0.0))
            r += b;
    } else {
        r = copysign(0.0, b);
    }
    return r;
}

// np.remainder(u, 2.0) without the fmod call: u / 2 and 2 floor(u / 2) are exact, so u - 2 floor(u / 2) is the
// exact floored remainder rounded once -- the same single rounding np.remainder makes (fmod is exact, its
// "+ 2.0" for negative u is the one inexact step there).  NaN and infinities give NaN like fmod.
__device__ __forceinline__ double floored_mod2(double u)
{
    const double r = u - 2.0 * floor(u * 0.5);
    return (u - u == 0.0) ? r : __longlong_as_double(0x7ff8000000000000ll);
}

// grad_spread (svgrasterize.py:1661-1668)
__device__ __forceinline__ double grad_spread(double t, int spread)
{
    if (spread == 1)
        // np.modf's fractional part: exact, keeps the sign (repeat degenerates to pad for t < 0), 0 for +-inf
        return fabs(t) == __longlong_as_double(0x7ff0000000000000ll) ? 0.0 : t - trunc(t);
    if (spread == 2)
        return fabs(floored_mod2(t + 1.0) - 1.0);
    return t;
}

// grad_interpolate (svgrasterize.py:1671-1683); stops premultiplied in the target colour space
__device__ __forceinline__ float4 grad_color(double v, const StopRec *__restrict__ st, int n)
{
    if (v <= st[0].offset)
        return f4(st[0].color[0], st[0].color[1], st[0].color[2], st[0].color[3]);
    if (v > st[n - 1].offset)
        return f4(st[n - 1].color[0], st[n - 1].color[1], st[n - 1].color[2], st[n - 1].color[3]);
    float4 o = f4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k + 1 < n; k++) {
        double o0 = st[k].offset, o1 = st[k + 1].offset;
        if (v > o0 && v <= o1) {
            float ratio = (float)((v - o0) * st[k].inv_span);
            o.x += fmaf(ratio, st[k + 1].color[0] - st[k].color[0], st[k].color[0]);
            o.y += fmaf(ratio, st[k + 1].color[1] - st[k].color[1], st[k].color[1]);
            o.z += fmaf(ratio, st[k + 1].color[2] - st[k].color[2], st[k].color[2]);
            o.w += fmaf(ratio, st[k + 1].color[3] - st[k].color[3], st[k].color[3]);
        }
    }
    return o;  // NaN offsets match no interval: transparent black
}

// Gradient parameter of a pixel centre (x, y) = (row + .5, col + .5).  The reference maps the centre through
// transform.invert and the inverse gradientTransform (svgrasterize.py:1022-1031, :1558, :1602) and then applies
// the gradient's own formula; both maps are affine, so the host encoder composes them (float64) into the
// coefficients below and the device evaluates one fused expression per pixel:
//   linear        t  = g[0] x + g[1] y + g[2]                              (svgrasterize.py:1561-1562)
//   radial        o  = (m1[0] x + m1[1] y + m1[2],  m1[3] x + m1[4] y + m1[5]) = (p - c) / r,  t = |o|   (:1605-1607)
//   two-circle    pd = (m1[0] x + m1[1] y + m1[2],  m1[3] x + m1[4] y + m1[5]) = p - f
//                 g = [cdx, cdy, fr * rd, a, fr^2, fr / (fr - r), fr != r, 1 / a]                          (:1612-1644)
__device__ __forceinline__ double focal_det(const PaintRec &p, double x, double y, double *b_out)
{
    double pdx = fma(p.m1[0], x, fma(p.m1[1], y, p.m1[2]));
    double pdy = fma(p.m1[3], x, fma(p.m1[4], y, p.m1[5]));
    double b = fma(pdx, p.g[0], fma(pdy, p.g[1], p.g[2]));
    double cc = fma(pdx, pdx, fma(pdy, pdy, -p.g[4]));
    *b_out = b;
    return fma(b, b, -p.g[3] * cc);
}

// Paint colour at the pixel centre (x, y): premultiplied RGBA.  `p` may live in shared memory.
// pat / pat_stride: the pattern tile image (PAINT_PATTERN only).
// `stops`: the paint's StopRec array (global memory, or a shared-memory copy staged by the caller).
__device__ __forceinline__ float4 paint_eval(const RenderTables &T, const PaintRec &p, const StopRec *stops, double x,
                                             double y, const float4 *pat, int pat_stride)
{
    if (p.kind == PAINT_SOLID)
        return f4(p.color[0], p.color[1], p.color[2], p.color[3]);
    if (p.kind == PAINT_PATTERN) {
        // Path.fill pattern branch (svgrasterize.py:1074-1094); Transform.__call__ roundings (:531-534)
        double ax = fma(y, p.m1[1], x * p.m1[0]) + p.m1[2];
        double ay = fma(y, p.m1[4], x * p.m1[3]) + p.m1[5];
        ax = floored_mod(ax - p.g[0], p.g[2]);
        ay = floored_mod(ay - p.g[1], p.g[3]);
        double bx = fma(ay, p.m2[1], ax * p.m2[0]) + p.m2[2];
        double by = fma(ay, p.m2[4], ax * p.m2[3]) + p.m2[5];
        long long ir = (long long)bx - p.pat_r0, ic = (long long)by - p.pat_c0;  // astype(int): truncation
        if (ir < 0) ir += p.pat_rows;  // numpy negative indices wrap
        if (ic < 0) ic += p.pat_cols;
        if (ir < 0 || ic < 0 || ir >= p.pat_rows || ic >= p.pat_cols)
            return f4(0.f, 0.f, 0.f, 0.f);
        return __ldg(pat + ir * pat_stride + ic);
    }
    double t;
    if (p.kind == PAINT_LINEAR) {
        t = fma(p.g[0], x, fma(p.g[1], y, p.g[2]));
    } else if (p.kind == PAINT_RADIAL) {
        double ox = fma(p.m1[0], x, fma(p.m1[1], y, p.m1[2]));
        double oy = fma(p.m1[3], x, fma(p.m1[4], y, p.m1[5]));
        t = sqrt(fma(ox, ox, oy * oy));
    } else {
        double b;
        double det = focal_det(p, x, y, &b);
        bool any_neg = T.focal_flags[p.flag] != 0;
        if (any_neg && det < 0)
            return f4(0.f, 0.f, 0.f, 0.f);
        double sq = sqrt(det);
        double t1 = (b + sq) * p.g[7], t2 = (b - sq) * p.g[7];
        t = (t1 != t1 || t2 != t2) ? __longlong_as_double(0x7ff8000000000000ll) : (t1 > t2 ? t1 : t2);
        if (any_neg && p.g[6] != 0.0 && !(t > p.g[5]))
            return f4(0.f, 0.f, 0.f, 0.f);
    }
    return grad_color(grad_spread(t, p.spread), stops, p.stop_cnt);
}

// Gradient parameter only (after spread), or NaN-free "invalid" -> *ok = false (transparent pixel)
__device__ __forceinline__ double grad_param(const RenderTables &T, const PaintRec &p, double x, double y, bool *ok)
{
    *ok = true;
    double t;
    if (p.kind == PAINT_LINEAR) {
        t = fma(p.g[0], x, fma(p.g[1], y, p.g[2]));
    } else if (p.kind == PAINT_RADIAL) {
        double ox = fma(p.m1[0], x, fma(p.m1[1], y, p.m1[2]));
        double oy = fma(p.m1[3], x, fma(p.m1[4], y, p.m1[5]));
        t = sqrt(fma(ox, ox, oy * oy));
    } else {
        double b;
        double det = focal_det(p, x, y, &b);
        bool any_neg = T.focal_flags[p.flag] != 0;
        if (any_neg && det < 0) {
            *ok = false;
            return 0.0;
        }
        double sq = sqrt(det);
        double t1 = (b + sq) * p.g[7], t2 = (b - sq) * p.g[7];
        t = (t1 != t1 || t2 != t2) ? __longlong_as_double(0x7ff8000000000000ll) : (t1 > t2 ? t1 : t2);
        if (any_neg && p.g[6] != 0.0 && !(t > p.g[5])) {
            *ok = false;
            return 0.0;
        }
    }
    return grad_spread(t, p.spread);
}

// grad_interpolate for N pixels at once with the stop loop outermost: a stop's offsets and colours are
// loaded once for all N pixels of the thread instead of once per pixel.  t: spread parameter, need[k]:
// pixel k wants a colour.  out[k] is written for every k (zero when !need[k] or no interval matches).
template <int N>
__device__ __forceinline__ void grad_colors(const double *t, const bool *need, const StopRec *__restrict__ st, int n,
                                            float4 *out)
{
    bool open[N];  // still looking for its interval
#pragma unroll
    for (int k = 0; k < N; k++) {
        out[k] = f4(0.f, 0.f, 0.f, 0.f);
        open[k] = need[k];
    }
    {
        const double o_min = st[0].offset, o_max = st[n - 1].offset;
        const float4 c_min = f4(st[0].color[0], st[0].color[1], st[0].color[2], st[0].color[3]);
        const float4 c_max = f4(st[n - 1].color[0], st[n - 1].color[1], st[n - 1].color[2], st[n - 1].color[3]);
#pragma unroll
        for (int k = 0; k < N; k++) {
            if (open[k] && t[k] <= o_min)
                out[k] = c_min, open[k] = false;
            else if (open[k] && t[k] > o_max)
                out[k] = c_max, open[k] = false;
        }
    }
    for (int j = 0; j + 1 < n; j++) {
        const double o0 = st[j].offset, o1 = st[j + 1].offset, inv = st[j].inv_span;
        const float4 c0 = f4(st[j].color[0], st[j].color[1], st[j].color[2], st[j].color[3]);
        const float4 dc = f4(st[j + 1].color[0] - c0.x, st[j + 1].color[1] - c0.y, st[j + 1].color[2] - c0.z,
                             st[j + 1].color[3] - c0.w);
#pragma unroll
        for (int k = 0; k < N; k++)
            if (open[k] && t[k] > o0 && t[k] <= o1) {
                const float ratio = (float)((t[k] - o0) * inv);
                float4 q = madd4(dc, ratio, c0);
                out[k] = f4(out[k].x + q.x, out[k].y + q.y, out[k].z + q.z, out[k].w + q.w);
            }
    }
}

// ---- source access, split so that a thread can issue the loads of several pixels before using them
__device__ __forceinline__ bool src_hits(const SrcRec &s, int r, int c)
{
    return r >= s.r0 && r < s.r0 + s.rows && c >= s.c0 && c < s.c0 + s.cols;
}

// raw memory value of the source at (r, c) (inside its bbox): RGBA for SRC_L4, (a, a, a, a) otherwise
__device__ __forceinline__ float4 src_load(const RenderTables &T, const SrcRec &s, int r, int c)
{
    long long idx = (long long)(r - s.br0) * s.stride + (c - s.bc0);
    if (s.kind == SRC_L4)
        return __ldg(reinterpret_cast<const float4 *>(T.layers + s.off) + idx);
    const float *base = (s.kind == SRC_L1) ? T.layers : T.cov;
    float a = __ldg(base + s.off + idx);
    return f4(a, a, a, a);
}

// paint, folded opacity and Layer.convert applied to a loaded value; (x, y) = pixel centre
__device__ __forceinline__ float4 src_finish(const RenderTables &T, const SrcRec &s, const PaintRec *paint, float4 v,
                                             double x, double y)
{
    if (s.kind == SRC_COVPAINT) {
        if (v.w == 0.f)
            return f4(0.f, 0.f, 0.f, 0.f);
        float a = v.w;
        float4 p = paint_eval(T, *paint, T.stops + paint->stop_off, x, y,
                              reinterpret_cast<const float4 *>(T.layers + s.off2), s.stride2);
        v = f4(p.x * a, p.y * a, p.z * a, p.w * a);
    }
    if (s.mul != 1.0f)
        v = f4(v.x * s.mul, v.y * s.mul, v.z * s.mul, v.w * s.mul);
    if (s.kind != SRC_L1 && s.kind != SRC_COV && !conv_is_identity(s.conv))
        v = convert_px(v, s.conv);
    return v;
}

// Value of a source at global pixel (r, c), which must lie inside its bbox.
__device__ __forceinline__ float4 fetch_src(const RenderTables &T, const SrcRec &s, int r, int c)
{
    float4 v = src_load(T, s, r, c);
    const PaintRec *p = s.kind == SRC_COVPAINT ? T.paints + s.paint : nullptr;
    return src_finish(T, s, p, v, (double)r + 0.5, (double)c + 0.5);
}

// one-channel value of a stencil modifier entry at (r, c) (inside the owner's valid region)
__device__ __forceinline__ float mod_value(const RenderTables &T, const SrcRec &s, int r, int c)
{
    long long idx = (long long)(r - s.br0) * s.stride + (c - s.bc0);
    float m;
    if (s.kind == SRC_MOD_COV) {
        m = __ldg(T.cov + s.off + idx);
    } else if (s.kind == SRC_MOD_L1) {
        m = __ldg(T.layers + s.off + idx);
    } else if (s.kind == SRC_MOD_L4A) {
        m = __ldg(T.layers + s.off + 4 * idx + 3);
    } else {
        float4 v = __ldg(reinterpret_cast<const float4 *>(T.layers + s.off) + idx);
        if (s.mul != 1.0f)
            v = f4(v.x * s.mul, v.y * s.mul, v.z * s.mul, v.w * s.mul);
        v = convert_px(v, s.conv);
        return (v.x * 0.2125f + v.y * 0.7154f + v.z * 0.072f) * v.w;  // svgrasterize.py:734-736
    }
    return m * s.mul;
}
