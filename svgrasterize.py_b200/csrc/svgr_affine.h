// svgr_affine.h -- the float64 matrix arithmetic of the scene walk, written once for the host encoder
// (encode_flat.cpp) and for the device kernel that completes objectBoundingBox gradients (k_flatten.cu).
// Every function follows numpy's rounding on this platform, so that the host path produces the
// same bits as encode.py does with numpy.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define SVGR_HD __host__ __device__ __forceinline__
#else
#define SVGR_HD inline
#endif

struct M23 {
    double m[6];  // row-major 2 x 3; the third row is (0, 0, 1)
};

// numpy: a @ b for 3 x 3 float64 (measured on this OpenBLAS: fma(a2, b2, fma(a1, b1, a0 * b0)) per element)
SVGR_HD M23 affine_matmul(const M23 &a, const M23 &b)
{
    M23 c;
    for (int i = 0; i < 2; i++) {
        const double a0 = a.m[3 * i], a1 = a.m[3 * i + 1], a2 = a.m[3 * i + 2];
        c.m[3 * i + 0] = fma(a2, 0.0, fma(a1, b.m[3], a0 * b.m[0]));
        c.m[3 * i + 1] = fma(a2, 0.0, fma(a1, b.m[4], a0 * b.m[1]));
        c.m[3 * i + 2] = fma(a2, 1.0, fma(a1, b.m[5], a0 * b.m[2]));
    }
    return c;
}

// np.linalg.inv of the 3 x 3 affine matrix: LU with partial pivoting, columns of the identity solved one by one
// (right-looking elimination, multipliers scaled by the reciprocal pivot, fused substitution steps)
SVGR_HD bool affine_invert(const M23 &t, M23 &out)
{
    double a[3][3] = {{t.m[0], t.m[1], t.m[2]}, {t.m[3], t.m[4], t.m[5]}, {0.0, 0.0, 1.0}};
    int perm[3] = {0, 1, 2};
    for (int k = 0; k < 3; k++) {
        int p = k;
        for (int i = k + 1; i < 3; i++)
            if (fabs(a[i][k]) > fabs(a[p][k]))
                p = i;
        if (a[p][k] == 0.0)
            return false;
        if (p != k) {
            for (int j = 0; j < 3; j++) {
                const double tmp = a[k][j];
                a[k][j] = a[p][j], a[p][j] = tmp;
            }
            const int tp = perm[k];
            perm[k] = perm[p], perm[p] = tp;
        }
        const double rp = 1.0 / a[k][k];
        for (int i = k + 1; i < 3; i++)
            a[i][k] = a[i][k] * rp;
        for (int i = k + 1; i < 3; i++)
            for (int j = k + 1; j < 3; j++)
                a[i][j] = a[i][j] - a[i][k] * a[k][j];
    }
    double x[3][3];
    for (int c = 0; c < 3; c++) {
        double y[3];
        for (int i = 0; i < 3; i++)
            y[i] = perm[i] == c ? 1.0 : 0.0;
        for (int i = 0; i < 3; i++)
            for (int q = 0; q < i; q++)
                y[i] = fma(-a[i][q], y[q], y[i]);
        for (int i = 2; i >= 0; i--) {
            for (int q = i + 1; q < 3; q++)
                y[i] = fma(-a[i][q], y[q], y[i]);
            y[i] = y[i] * (1.0 / a[i][i]);
        }
        for (int i = 0; i < 3; i++)
            x[i][c] = y[i];
    }
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 3; j++)
            out.m[3 * i + j] = x[i][j];
    return true;
}

// The pixel -> gradient-space coefficients of a gradient paint from `to_user` (pixel centre -> the paint's user
// space, the inverse gradientTransform already applied): what encode.py's _leaf stores in PaintRec.g / .m1
// (svgrasterize.py:1561-1562, :1605-1607, :1612-1644).  kind = PAINT_*: 1 linear (geom = p0x p0y vx vy vv),
// 2 radial (geom = cx cy r), 3 two-circle (geom = fx fy).
SVGR_HD void gradient_coefficients(int kind, const double *geom, const M23 &to_user, double *g, double *m1)
{
    const double A00 = to_user.m[0], A01 = to_user.m[1], T0 = to_user.m[2];
    const double A10 = to_user.m[3], A11 = to_user.m[4], T1 = to_user.m[5];
    if (kind == 1) {
        const double v0 = geom[2], v1 = geom[3], vv = geom[4];
        g[0] = fma(v1, A10, v0 * A00) / vv;
        g[1] = fma(v1, A11, v0 * A01) / vv;
        const double d0 = T0 - geom[0], d1 = T1 - geom[1];
        g[2] = fma(d1, v1, d0 * v0) / vv;
    } else if (kind == 2) {
        const double r = geom[2];
        m1[0] = A00 / r, m1[1] = A01 / r, m1[3] = A10 / r, m1[4] = A11 / r;
        m1[2] = (T0 - geom[0]) / r, m1[5] = (T1 - geom[1]) / r;
    } else {
        m1[0] = A00, m1[1] = A01, m1[3] = A10, m1[4] = A11;
        m1[2] = T0 - geom[0], m1[5] = T1 - geom[1];
    }
}
