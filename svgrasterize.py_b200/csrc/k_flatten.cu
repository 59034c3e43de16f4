// k_flatten.cu -- curve flattening + path bounds (subsystem 1 of the hot path).
//
// Replaces, for a whole batch of paths in one launch:
//   Path.mask segment loop           svgrasterize.py:930-945
//   bezier2_to_bezier3               svgrasterize.py:2182-2184 (matrix :2052)
//   Transform.__call__               svgrasterize.py:531-534
//   bezier3_flatness_batch           svgrasterize.py:2071-2088
//   bezier3_split_batch              svgrasterize.py:2066-2068
//   bezier3_flatten_batch            svgrasterize.py:2091-2098
//   mask sizing                      svgrasterize.py:961-975
//
// Bit-exactness: the reference's numpy contractions run through OpenBLAS with
// a fixed fused-multiply-add order.  This file is
// compiled with -fmad=false and spells every fma() explicitly, so each control
// point, flatness value and emitted end point has the reference's bits.  The
// edge ORDER differs (the reference is breadth-first per path, this kernel is
// warp-cooperative depth-first); parity is defined on the per-path multiset.
//
// Kernel shape: one warp owns a LIFO work list of cubics in shared memory.
// Each round every lane pops one cubic, tests flatness, and either emits the
// chord [p0, p3] (warp-aggregated append to the global edge list) or splits it
// at t = 1/2 and pushes both halves.  Lanes therefore stay busy regardless of
// how unevenly the subdivision depth is distributed over the input curves.
#include "svgr_kernels.h"
#include "svgr_affine.h"
#include "../../include/svgr_b200.h"
#include <algorithm>

#define FLAT_WARPS 4
#define FLAT_Q 160       // work-list slots per warp
#define FLAT_MAXD 24     // subdivision depth cap (the reference has none and spins on NaN)
#define FLAT_QCAP (FLAT_Q - FLAT_MAXD)

__device__ __forceinline__ double dot4(double m0, double m1, double m2, double m3, double b0, double b1, double b2,
                                       double b3)
{
    // strided ddot with two accumulators: fma(m0,b0, m2*b2) + fma(m1,b1, m3*b3)
    return fma(m0, b0, m2 * b2) + fma(m1, b1, m3 * b3);
}

__device__ __forceinline__ unsigned long long key_of(double v)
{
    // order-preserving map double -> u64
    long long b = __double_as_longlong(v);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}

__device__ __forceinline__ double of_key(unsigned long long k)
{
    long long b = (k & 0x8000000000000000ull) ? (long long)(k & 0x7fffffffffffffffull) : (long long)~k;
    return __longlong_as_double(b);
}

// minmax[path] = {min r, min c, max r, max c} as ordered keys
__device__ __forceinline__ void bounds_update(unsigned long long *minmax, uint32_t path, double r0, double c0,
                                              double r1, double c1)
{
    unsigned long long *mm = minmax + 4ull * path;
    atomicMin(mm + 0, key_of(r0 < r1 ? r0 : r1));
    atomicMin(mm + 1, key_of(c0 < c1 ? c0 : c1));
    atomicMax(mm + 2, key_of(r0 < r1 ? r1 : r0));
    atomicMax(mm + 3, key_of(c0 < c1 ? c1 : c0));
}

// Per-lane running bounds of the path the lane is emitting edges for: the global min / max atomics are issued
// when the lane moves on to another path (and at the end), not per edge.  A path whose outline explodes into
// thousands of edges (a stroke offset near a cusp) would otherwise serialise 4 atomics per edge on the same four
// addresses.
struct LaneBounds {
    uint32_t path;
    double mn_r, mn_c, mx_r, mx_c;
};

__device__ __forceinline__ void bounds_flush(unsigned long long *minmax, const LaneBounds &b)
{
    if (b.path == 0xffffffffu)
        return;
    unsigned long long *mm = minmax + 4ull * b.path;
    atomicMin(mm + 0, key_of(b.mn_r));
    atomicMin(mm + 1, key_of(b.mn_c));
    atomicMax(mm + 2, key_of(b.mx_r));
    atomicMax(mm + 3, key_of(b.mx_c));
}

__device__ __forceinline__ void emit_edges(bool has, double r0, double c0, double r1, double c1, uint32_t path,
                                           double *edges, uint32_t *edge_path, unsigned long long cap,
                                           unsigned long long *n_edges, unsigned long long *minmax, int lane,
                                           LaneBounds &lb)
{
    unsigned m = __ballot_sync(0xffffffffu, has);
    if (!m)
        return;
    unsigned long long base = 0;
    int leader = __ffs(m) - 1;
    if (lane == leader)
        base = atomicAdd(n_edges, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (has) {
        unsigned long long idx = base + __popc(m & ((1u << lane) - 1));
        if (idx < cap) {
            double2 *e = reinterpret_cast<double2 *>(edges + 4 * idx);
            e[0] = make_double2(r0, c0);
            e[1] = make_double2(r1, c1);
            edge_path[idx] = path;
        }
        const double lo_r = r0 < r1 ? r0 : r1, hi_r = r0 < r1 ? r1 : r0;
        const double lo_c = c0 < c1 ? c0 : c1, hi_c = c0 < c1 ? c1 : c0;
        if (lb.path != path) {
            bounds_flush(minmax, lb);
            lb.path = path, lb.mn_r = lo_r, lb.mn_c = lo_c, lb.mx_r = hi_r, lb.mx_c = hi_c;
        } else {
            // key order: the comparisons below pick what atomicMin / atomicMax on the ordered keys would keep
            lb.mn_r = key_of(lo_r) < key_of(lb.mn_r) ? lo_r : lb.mn_r;
            lb.mn_c = key_of(lo_c) < key_of(lb.mn_c) ? lo_c : lb.mn_c;
            lb.mx_r = key_of(hi_r) > key_of(lb.mx_r) ? hi_r : lb.mx_r;
            lb.mx_c = key_of(hi_c) > key_of(lb.mx_c) ? hi_c : lb.mx_c;
        }
    }
}

// A warp works through the subdivision tree of its 32 segments alone.  That is fine for ordinary curves (depth
// 3-5), but one curve that needs depth 12 (a stroke outline blown up by a cusp: thousands of edges) kept a single
// warp busy for 0.8 ms while the rest of the GPU had finished.  So the tree is cut every FLAT_PASS_DEPTH levels:
// a node that is still not flat at the pass's depth limit is appended to a global overflow list instead of being
// split locally, and the next pass (same kernel, reading that list: `in`) spreads those nodes over all warps
// again.  The splits are the same de Casteljau steps in the same arithmetic, so the edges are bit-identical.
struct FlatOverflow {
    double *cubic;    // 8 doubles per node (already in presentation space)
    uint32_t *path;
    uint8_t *depth;
    int *count;       // nodes appended (may exceed cap: the excess was split locally instead)
    int cap;
};

__global__ void __launch_bounds__(FLAT_WARPS * 32)
flatten_kernel(const uint8_t *__restrict__ seg_tag, const double *__restrict__ seg_data,
               const uint32_t *__restrict__ seg_path, long long n_seg_host, const int *__restrict__ n_seg_dev,
               const PathRec *__restrict__ paths, double thr, double *__restrict__ edges,
               uint32_t *__restrict__ edge_path, unsigned long long cap, unsigned long long *n_edges,
               unsigned long long *minmax, FlatOverflow in, FlatOverflow out, int depth_limit)
{
    __shared__ double st[FLAT_WARPS][8][FLAT_Q];
    __shared__ uint32_t st_path[FLAT_WARPS][FLAT_Q];
    __shared__ uint8_t st_depth[FLAT_WARPS][FLAT_Q];

    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    long long n_seg = n_seg_host;
    if (n_seg_dev) {
        long long nd = *n_seg_dev;
        n_seg = nd < n_seg ? nd : n_seg;
    }
    if (in.cubic) {
        long long nd = *in.count;
        n_seg = nd < in.cap ? nd : in.cap;
    }
    // segments per warp and refill: 32, except that a short overflow list is spread one node (or a few) per warp --
    // 32 deferred nodes in one warp would be the same lonely tree again, 5 levels further down
    int per = 32;
    if (in.cubic) {
        const long long warps = (long long)gridDim.x * FLAT_WARPS;
        const long long fair = (n_seg + warps - 1) / warps;
        per = (int)(fair < 1 ? 1 : (fair > 32 ? 32 : fair));
    }
    const long long n_chunks = (n_seg + per - 1) / per;
    long long chunk = (long long)blockIdx.x * FLAT_WARPS + w;
    const long long chunk_step = (long long)gridDim.x * FLAT_WARPS;
    int count = 0;
    LaneBounds lb;
    lb.path = 0xffffffffu, lb.mn_r = lb.mn_c = lb.mx_r = lb.mx_c = 0.0;

    for (;;) {
        // ---- refill: take the next 32 segments while there is room for 32 cubics
        if (count <= FLAT_QCAP - 64 && chunk < n_chunks) {
            // input segments: 32 consecutive ones per warp (coalesced).  Deferred nodes: lane l of chunk c takes node
            // c + l * n_chunks -- the children of one exploding curve sit next to each other in the list and must
            // not end up in the same warp again (a node is a 64-byte record: the strided read costs nothing)
            long long i = in.cubic ? chunk + (long long)lane * n_chunks : (chunk << 5) + lane;
            if (lane >= per)
                i = n_seg;
            chunk += chunk_step;
            int tag = SEG_NOP;
            double p[8];
            uint32_t path = 0;
            int depth0 = 0;
            if (i < n_seg)
                tag = in.cubic ? -1 : seg_tag[i];
            bool is_line = tag == SEG_LINE || tag == SEG_CLOSED || tag == SEG_UNCLOSED;
            bool is_curve = tag == SEG_QUAD || tag == SEG_CUBIC || tag == -1;
            if (tag == -1) {  // a node deferred by the pass before: transformed cubic, its path and depth
                const double2 *d = reinterpret_cast<const double2 *>(in.cubic + 8 * i);
                double2 a = d[0], b = d[1], c = d[2], e = d[3];
                p[0] = a.x, p[1] = a.y, p[2] = b.x, p[3] = b.y, p[4] = c.x, p[5] = c.y, p[6] = e.x, p[7] = e.y;
                path = in.path[i];
                depth0 = in.depth[i];
            } else if (is_line || is_curve) {
                const double2 *d = reinterpret_cast<const double2 *>(seg_data + 8 * i);
                double2 a = d[0], b = d[1];
                p[0] = a.x, p[1] = a.y, p[2] = b.x, p[3] = b.y;
                if (is_curve) {
                    double2 c = d[2], e = d[3];
                    p[4] = c.x, p[5] = c.y, p[6] = e.x, p[7] = e.y;
                }
                path = seg_path[i];
                if (tag == SEG_QUAD) {
                    // bezier2_to_bezier3: out = fma(m2,p2, fma(m1,p1, m0*p0)), rows of
                    // [[1,0,0],[1/3,2/3,0],[0,2/3,1/3],[0,0,1]]
                    const double t1 = 1.0 / 3, t2 = 2.0 / 3;
                    double q[6] = {p[0], p[1], p[2], p[3], p[4], p[5]};
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        p[0 + k] = fma(0.0, q[4 + k], fma(0.0, q[2 + k], 1.0 * q[k]));
                        p[2 + k] = fma(0.0, q[4 + k], fma(t2, q[2 + k], t1 * q[k]));
                        p[4 + k] = fma(t1, q[4 + k], fma(t2, q[2 + k], 0.0 * q[k]));
                        p[6 + k] = fma(1.0, q[4 + k], fma(0.0, q[2 + k], 0.0 * q[k]));
                    }
                }
                // Transform.__call__: out_c = fma(y, M[c][1], x*M[c][0]) + M[c][2]
                const double *m = paths[path].m;
                double m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4], m5 = m[5];
                int np = is_line ? 2 : 4;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (k < np) {
                        double x = p[2 * k], y = p[2 * k + 1];
                        p[2 * k] = fma(y, m1, x * m0) + m2;
                        p[2 * k + 1] = fma(y, m4, x * m3) + m5;
                    }
                }
            }
            emit_edges(is_line, p[0], p[1], p[2], p[3], path, edges, edge_path, cap, n_edges, minmax, lane, lb);
            unsigned mc = __ballot_sync(0xffffffffu, is_curve);
            if (is_curve) {
                int pos = count + __popc(mc & ((1u << lane) - 1));
#pragma unroll
                for (int k = 0; k < 8; k++)
                    st[w][k][pos] = p[k];
                st_path[w][pos] = path;
                st_depth[w][pos] = (uint8_t)depth0;
            }
            count += __popc(mc);
            __syncwarp();
            continue;
        }
        if (count == 0)
            break;

        // ---- pop up to 32 cubics from the top of the list
        int room = FLAT_QCAP - count;
        int k = count < 32 ? count : 32;
        if (k > room)
            k = room > 1 ? room : 1;  // depth-first on one entry may use the MAXD head-room
        bool active = lane < k;
        double c[8];
        uint32_t path = 0;
        int depth = 0;
        if (active) {
            int pos = count - 1 - lane;
#pragma unroll
            for (int j = 0; j < 8; j++)
                c[j] = st[w][j][pos];
            path = st_path[w][pos];
            depth = st_depth[w][pos];
        }
        count -= k;
        __syncwarp();

        bool flat = false;
        if (active) {
            // bezier3_flatness_batch: rows [-2,3,0,-1] and [-1,0,3,-2]
            double ux = dot4(-2.0, 3.0, 0.0, -1.0, c[0], c[2], c[4], c[6]);
            double uy = dot4(-2.0, 3.0, 0.0, -1.0, c[1], c[3], c[5], c[7]);
            double vx = dot4(-1.0, 0.0, 3.0, -2.0, c[0], c[2], c[4], c[6]);
            double vy = dot4(-1.0, 0.0, 3.0, -2.0, c[1], c[3], c[5], c[7]);
            ux = ux * ux, uy = uy * uy, vx = vx * vx, vy = vy * vy;
            double f = (ux > uy ? ux : uy) + (vx > vy ? vx : vy);
            flat = (f < thr) || depth >= FLAT_MAXD;
        }
        emit_edges(active && flat, c[0], c[1], c[6], c[7], path, edges, edge_path, cap, n_edges, minmax, lane, lb);

        bool split = active && !flat;
        // at the pass's depth limit a node goes to the overflow list (if there is room) instead of being split here
        if (out.cubic) {
            const bool defer = split && depth >= depth_limit;
            const unsigned md = __ballot_sync(0xffffffffu, defer);
            if (md) {
                int base = 0;
                const int leader = __ffs(md) - 1;
                if (lane == leader)
                    base = atomicAdd(out.count, __popc(md));
                base = __shfl_sync(0xffffffffu, base, leader);
                const int idx = base + __popc(md & ((1u << lane) - 1));
                if (defer && idx < out.cap) {
                    double2 *d = reinterpret_cast<double2 *>(out.cubic + 8ll * idx);
                    d[0] = make_double2(c[0], c[1]), d[1] = make_double2(c[2], c[3]);
                    d[2] = make_double2(c[4], c[5]), d[3] = make_double2(c[6], c[7]);
                    out.path[idx] = path;
                    out.depth[idx] = (uint8_t)depth;
                    split = false;
                }
            }
        }
        unsigned ms = __ballot_sync(0xffffffffu, split);
        if (split) {
            int pos = count + 2 * __popc(ms & ((1u << lane) - 1));
            // bezier3_split_batch: rows of the 8x4 de Casteljau matrix, dot4 recipe
#pragma unroll
            for (int j = 0; j < 2; j++) {
                double b0 = c[j], b1 = c[2 + j], b2 = c[4 + j], b3 = c[6 + j];
                double mid = dot4(0.125, 0.375, 0.375, 0.125, b0, b1, b2, b3);
                st[w][0 + j][pos] = dot4(1.0, 0.0, 0.0, 0.0, b0, b1, b2, b3);
                st[w][2 + j][pos] = dot4(0.5, 0.5, 0.0, 0.0, b0, b1, b2, b3);
                st[w][4 + j][pos] = dot4(0.25, 0.5, 0.25, 0.0, b0, b1, b2, b3);
                st[w][6 + j][pos] = mid;
                st[w][0 + j][pos + 1] = mid;
                st[w][2 + j][pos + 1] = dot4(0.0, 0.25, 0.5, 0.25, b0, b1, b2, b3);
                st[w][4 + j][pos + 1] = dot4(0.0, 0.0, 0.5, 0.5, b0, b1, b2, b3);
                st[w][6 + j][pos + 1] = dot4(0.0, 0.0, 0.0, 1.0, b0, b1, b2, b3);
            }
            st_path[w][pos] = path;
            st_path[w][pos + 1] = path;
            st_depth[w][pos] = (uint8_t)(depth + 1);
            st_depth[w][pos + 1] = (uint8_t)(depth + 1);
        }
        count += 2 * __popc(ms);
        __syncwarp();
    }
    bounds_flush(minmax, lb);
}

__global__ void minmax_init_kernel(unsigned long long *minmax, int n_paths)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_paths) {
        minmax[4 * i + 0] = ~0ull;
        minmax[4 * i + 1] = ~0ull;
        minmax[4 * i + 2] = 0ull;
        minmax[4 * i + 3] = 0ull;
    }
}

// Mask sizing (svgrasterize.py:961-975): floor(min) - 1, ceil(max) + 1, clipped to the viewport.
__global__ void bounds_kernel(const unsigned long long *__restrict__ minmax, const PathRec *__restrict__ paths,
                              int n_paths, PathBox *__restrict__ boxes, PathBox *__restrict__ full_boxes,
                              double *__restrict__ minmax_f64)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_paths)
        return;
    PathBox b = {0, 0, 0, 0}, fb = {0, 0, 0, 0};
    if (minmax[4 * i + 0] != ~0ull) {
        double mn_r = of_key(minmax[4 * i + 0]), mn_c = of_key(minmax[4 * i + 1]);
        double mx_r = of_key(minmax[4 * i + 2]), mx_c = of_key(minmax[4 * i + 3]);
        if (minmax_f64) {
            minmax_f64[4 * i + 0] = mn_r, minmax_f64[4 * i + 1] = mn_c;
            minmax_f64[4 * i + 2] = mx_r, minmax_f64[4 * i + 3] = mx_c;
        }
        const double lim = 1.0e9;  // keep the int conversion defined for wild inputs
        mn_r = fmax(-lim, fmin(lim, mn_r)), mn_c = fmax(-lim, fmin(lim, mn_c));
        mx_r = fmax(-lim, fmin(lim, mx_r)), mx_c = fmax(-lim, fmin(lim, mx_c));
        long long min_r = (long long)floor(mn_r) - 1, min_c = (long long)floor(mn_c) - 1;
        long long max_r = (long long)ceil(mx_r) + 1, max_c = (long long)ceil(mx_c) + 1;
        const PathRec &p = paths[i];
        auto clip_box = [&](bool has, const int32_t *vp, PathBox &o) {
            long long a_r = min_r, a_c = min_c, z_r = max_r, z_c = max_c;
            if (has) {
                long long vx = vp[0], vy = vp[1], vw = vp[2], vh = vp[3];
                if (a_r < vx) a_r = vx;
                if (a_c < vy) a_c = vy;
                if (z_r > vx + vw) z_r = vx + vw;
                if (z_c > vy + vh) z_c = vy + vh;
            }
            long long rows = z_r - a_r, cols = z_c - a_c;
            if (rows > 0 && cols > 0) {
                o.r0 = (int32_t)a_r, o.c0 = (int32_t)a_c;
                o.rows = (int32_t)(rows < 0x7fffffff ? rows : 0x7fffffff);
                o.cols = (int32_t)(cols < 0x7fffffff ? cols : 0x7fffffff);
            }
        };
        clip_box(p.has_viewport != 0, p.viewport, b);
        if (p.has_full)
            clip_box(true, p.full_viewport, fb);  // the box the whole-canvas render would give this mask
        else
            fb = b;
    } else if (minmax_f64) {
        minmax_f64[4 * i + 0] = minmax_f64[4 * i + 1] = minmax_f64[4 * i + 2] = minmax_f64[4 * i + 3] = 0.0;
    }
    boxes[i] = b;
    if (full_boxes)
        full_boxes[i] = fb;
}

// User-space bounding box of a set of paths' end points: ConvexHull.bbox
// (svgrasterize.py:2002-2007) -- the extremes of transform.invert(points) are
// attained on hull vertices, so the Graham scan is not needed.  One query =
// (inverse matrix, list of paths); q_of_path is a CSR path -> queries.
__global__ void cloud_bounds_kernel(const double *__restrict__ edges, const uint32_t *__restrict__ edge_path,
                                    const unsigned long long *__restrict__ n_edges, unsigned long long cap,
                                    const int *__restrict__ pq_off, const int *__restrict__ pq_idx,
                                    const double *__restrict__ q_inv, unsigned long long *q_minmax)
{
    unsigned long long n = *n_edges < cap ? *n_edges : cap;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t path = edge_path[i];
        int a = pq_off[path], b = pq_off[path + 1];
        if (a == b)
            continue;
        const double2 *e = reinterpret_cast<const double2 *>(edges + 4 * i);
        double2 p0 = e[0], p1 = e[1];
        for (int j = a; j < b; j++) {
            int q = pq_idx[j];
            const double *m = q_inv + 6 * q;
            double x0 = fma(p0.y, m[1], p0.x * m[0]) + m[2], y0 = fma(p0.y, m[4], p0.x * m[3]) + m[5];
            double x1 = fma(p1.y, m[1], p1.x * m[0]) + m[2], y1 = fma(p1.y, m[4], p1.x * m[3]) + m[5];
            bounds_update(q_minmax, (uint32_t)q, x0, y0, x1, y1);
        }
    }
}

// Completes the PaintRec of every objectBoundingBox gradient (svgr_bbox_job): ConvexHull.bbox of the leaf's end
// points in user space (the min / max keys cloud_bounds_kernel has just accumulated for the job) ->
// ConvexHull.bbox_transform = transform.translate(x, y).scale(w, h), or the transform itself for an empty box
// (svgrasterize.py:2002-2023) -> inverse -> inverse gradientTransform -> the coefficients paint_eval reads.  The same
// float64 recipes as the host encoder (svgr_affine.h); a singular matrix (a box of zero width xor height) is the
// reference's LinAlgError and raises *err.
__global__ void resolve_paints_kernel(const svgr_bbox_job *__restrict__ jobs, int n_jobs, const PathRec *__restrict__ paths,
                                      const unsigned long long *__restrict__ keys, PaintRec *paints, int *err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_jobs)
        return;
    const svgr_bbox_job j = jobs[i];
    if (keys[4 * i] == ~0ull)
        return;  // no end points: the leaf renders nothing and the paint is never evaluated
    const double x = of_key(keys[4 * i + 0]), y = of_key(keys[4 * i + 1]);
    const double w = of_key(keys[4 * i + 2]) - x, h = of_key(keys[4 * i + 3]) - y;
    M23 t, to_user;
    for (int k = 0; k < 6; k++)
        t.m[k] = paths[j.path].m[k];
    if (w <= 0 && h <= 0) {
        for (int k = 0; k < 6; k++)
            to_user.m[k] = j.inv[k];  // transform.invert as numpy computed it
    } else {
        const M23 tr = {{1.0, 0.0, x, 0.0, 1.0, y}}, sc = {{w, 0.0, 0.0, 0.0, h, 0.0}};
        t = affine_matmul(affine_matmul(t, tr), sc);
        if (!affine_invert(t, to_user)) {
            atomicOr(err, 1);
            return;
        }
    }
    if (j.has_grad_tr) {
        M23 gi;
        for (int k = 0; k < 6; k++)
            gi.m[k] = j.grad_inv[k];
        to_user = affine_matmul(gi, to_user);
    }
    PaintRec &p = paints[j.paint];
    gradient_coefficients(p.kind, j.geom, to_user, p.g, p.m1);
}

__global__ void keys_to_f64_kernel(const unsigned long long *__restrict__ keys, double *__restrict__ out, int n4)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        bool empty = keys[4 * i] == ~0ull;
        for (int k = 0; k < 4; k++)
            out[4 * i + k] = empty ? 0.0 : of_key(keys[4 * i + k]);
    }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
void svgr_launch_minmax_init(unsigned long long *minmax, int n, cudaStream_t s)
{
    if (n > 0)
        minmax_init_kernel<<<(n + 255) / 256, 256, 0, s>>>(minmax, n);
}

void svgr_launch_flatten(const uint8_t *seg_tag, const double *seg_data, const uint32_t *seg_path, long long n_seg,
                         const int *n_seg_dev, const PathRec *paths, double thr, double *edges, uint32_t *edge_path,
                         unsigned long long cap, unsigned long long *n_edges, unsigned long long *minmax, int sm_count,
                         const svgr_flat_overflow *ovf, cudaStream_t s)
{
    if (n_seg <= 0)
        return;
    long long chunks = (n_seg + 31) / 32;
    long long blocks = (chunks + FLAT_WARPS - 1) / FLAT_WARPS;
    long long max_blocks = (long long)sm_count * 8;
    if (blocks > max_blocks)
        blocks = max_blocks;
    FlatOverflow none = {nullptr, nullptr, nullptr, nullptr, 0}, out = none;
    if (ovf)
        out = {ovf->cubic[0], ovf->path[0], ovf->depth[0], ovf->counts + 0, ovf->cap};
    flatten_kernel<<<(unsigned)blocks, FLAT_WARPS * 32, 0, s>>>(seg_tag, seg_data, seg_path, n_seg, n_seg_dev, paths, thr,
                                                                edges, edge_path, cap, n_edges, minmax, none, out,
                                                                SVGR_FLAT_PASS_DEPTH);
}

// The passes behind the first one: pass k reads the nodes pass k - 1 deferred (counts[k - 1], buffers alternate) and
// defers what is still not flat SVGR_FLAT_PASS_DEPTH levels further down.  The host does not know the counts: every
// pass is launched, an empty one costs a launch of idle blocks.
void svgr_launch_flatten_overflow(const PathRec *paths, double thr, double *edges, uint32_t *edge_path,
                                  unsigned long long cap, unsigned long long *n_edges, unsigned long long *minmax,
                                  int sm_count, const svgr_flat_overflow *ovf, cudaStream_t s)
{
    for (int k = 1; k < SVGR_FLAT_PASSES; k++) {
        FlatOverflow in = {ovf->cubic[(k - 1) & 1], ovf->path[(k - 1) & 1], ovf->depth[(k - 1) & 1], ovf->counts + (k - 1),
                           ovf->cap};
        FlatOverflow out = {ovf->cubic[k & 1], ovf->path[k & 1], ovf->depth[k & 1], ovf->counts + k, ovf->cap};
        if (k == SVGR_FLAT_PASSES - 1)
            out.cubic = nullptr;  // the last pass finishes everything locally (depth cap FLAT_MAXD)
        int blocks = sm_count * 4;
        flatten_kernel<<<blocks, FLAT_WARPS * 32, 0, s>>>(nullptr, nullptr, nullptr, (long long)ovf->cap, nullptr, paths, thr,
                                                          edges, edge_path, cap, n_edges, minmax, in, out,
                                                          SVGR_FLAT_PASS_DEPTH * (k + 1));
    }
}

void svgr_launch_bounds(const unsigned long long *minmax, const PathRec *paths, int n_paths, PathBox *boxes,
                        PathBox *full_boxes,
                        double *minmax_f64, cudaStream_t s)
{
    if (n_paths > 0)
        bounds_kernel<<<(n_paths + 255) / 256, 256, 0, s>>>(minmax, paths, n_paths, boxes, full_boxes, minmax_f64);
}

void svgr_launch_resolve_paints(const double *edges, const uint32_t *edge_path, const unsigned long long *n_edges,
                                unsigned long long cap, const int *pq_off, const int *pq_idx, const double *q_inv,
                                unsigned long long *keys, const svgr_bbox_job *jobs, int n_jobs, const PathRec *paths,
                                PaintRec *paints, int *err, int sm_count, cudaStream_t s)
{
    if (n_jobs <= 0)
        return;
    minmax_init_kernel<<<(n_jobs + 255) / 256, 256, 0, s>>>(keys, n_jobs);
    cloud_bounds_kernel<<<sm_count * 4, 256, 0, s>>>(edges, edge_path, n_edges, cap, pq_off, pq_idx, q_inv, keys);
    resolve_paints_kernel<<<(n_jobs + 255) / 256, 256, 0, s>>>(jobs, n_jobs, paths, keys, paints, err);
}

void svgr_launch_cloud_bounds(const double *edges, const uint32_t *edge_path, const unsigned long long *n_edges,
                              unsigned long long cap, const int *pq_off, const int *pq_idx, const double *q_inv,
                              unsigned long long *q_minmax, double *q_out, int n_q, int sm_count, cudaStream_t s)
{
    if (n_q <= 0)
        return;
    minmax_init_kernel<<<(n_q + 255) / 256, 256, 0, s>>>(q_minmax, n_q);
    cloud_bounds_kernel<<<sm_count * 4, 256, 0, s>>>(edges, edge_path, n_edges, cap, pq_off, pq_idx, q_inv, q_minmax);
    keys_to_f64_kernel<<<(n_q + 255) / 256, 256, 0, s>>>(q_minmax, q_out, n_q);
}
