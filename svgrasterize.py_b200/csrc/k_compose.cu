// k_compose.cu -- paint + Porter-Duff compositing, one HBM pass per output layer
// (subsystem 3 of the hot path), and the final canvas quantisation.
//
// Replaces:
//   canvas_compose               svgrasterize.py:277-298
//   canvas_merge_union           svgrasterize.py:330-379  (OVER fast path and the padded `full` path)
//   canvas_merge_intersect       svgrasterize.py:382-416
//   canvas_merge_at              svgrasterize.py:304-327
//   Layer.convert / color_*      svgrasterize.py:129-164, :471-503
//   Layer.opacity                svgrasterize.py:171-175
//   Path.fill paint evaluation   svgrasterize.py:1014-1097 (solid, gradients, pattern gather)
//   SourceAlpha, luminance mask  svgrasterize.py:1803-1806, :734-736
//   Layer.color_matrix           svgrasterize.py:95-104
//   canvas_to_png quantisation   svgrasterize.py:263 (after Layer.write_png's convert, :212)
//
// One launch executes every op of one dependency level of a whole batch.  An op
// folds a list of sources into one output layer; a Path.fill layer is never
// materialised -- its source is (coverage mask, paint record) and the paint is
// evaluated in registers while compositing, so a child costs 4 B/px of reads and
// the output is written exactly once (16 B/px).
//
// Tiling: CTA = 32 rows x 32 columns of output, 256 threads, 4 pixels per thread (rows 8 apart), a warp covers
// 512 contiguous bytes of an RGBA row.  cull_kernel (one warp per tile) first writes, for every tile, the list of
// the op's sources that touch it, order preserved; the compose CTA copies its op, those sources and their paints
// into shared memory in one pass and folds them in registers.
#include <algorithm>

#include "svgr_device.cuh"

#define CMP_TR SVGR_CMP_TR   // 32 rows: 4 pixels per thread, 8 rows apart
#define CMP_TC SVGR_CMP_TC   // 32 columns: one warp = 512 contiguous bytes of an RGBA row
#define CMP_PX (CMP_TR / 8)
#ifndef CMP_CAP
#define CMP_CAP 48           // sources staged per round
#endif

__device__ __forceinline__ float4 blend_px(int mode, const float *k, float4 d, float4 s)
{
    float sa = s.w, da = d.w;
    switch (mode) {
    case MODE_OVER: {
        float q = 1.0f - sa;
        return f4(s.x + d.x * q, s.y + d.y * q, s.z + d.z * q, s.w + d.w * q);
    }
    case MODE_OUT: {
        float q = 1.0f - da;
        return f4(s.x * q, s.y * q, s.z * q, s.w * q);
    }
    case MODE_IN:
        return f4(s.x * da, s.y * da, s.z * da, s.w * da);
    case MODE_ATOP: {
        float q = 1.0f - sa;
        return f4(s.x * da + d.x * q, s.y * da + d.y * q, s.z * da + d.z * q, s.w * da + d.w * q);
    }
    case MODE_XOR: {
        float q = 1.0f - sa, p = 1.0f - da;
        return f4(s.x * p + d.x * q, s.y * p + d.y * q, s.z * p + d.z * q, s.w * p + d.w * q);
    }
    default: {
        float k1 = k[0], k2 = k[1], k3 = k[2], k4 = k[3];
        return f4(clip01(k1 * s.x * d.x + k2 * s.x + k3 * d.x + k4), clip01(k1 * s.y * d.y + k2 * s.y + k3 * d.y + k4),
                  clip01(k1 * s.z * d.z + k2 * s.z + k3 * d.z + k4), clip01(k1 * s.w * d.w + k2 * s.w + k3 * d.w + k4));
    }
    }
}

template <class T>
__device__ __forceinline__ void copy16(T *dst, const T *src, int part)
{
    reinterpret_cast<uint4 *>(dst)[part] = __ldg(reinterpret_cast<const uint4 *>(src) + part);
}

// Source pixels are read once per tile: they do not need a place in L1 (which the 64-register fold uses for its
// spilled loop invariants and the paint tables); L2 keeps them for the neighbouring tile that shares the line
// (measured: 4.34 -> 4.30 ms per 2048 icons).
__device__ __forceinline__ float ld_stream(const float *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_stream(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// One kernel for layer ops (float output) and canvas ops (RGBA8 output).  Per CTA:
//   1. the tile's head (cull_kernel) names its op and its list of sources; the op record, up to CMP_CAP
//      SrcRecs and the PaintRecs they need are copied into shared memory by all threads at once -- one
//      barrier, nobody chases records through global memory afterwards;
//   2. every thread folds the staged sources into its 4 pixels, issuing the 4 loads of a source back to back
//      before using any of them;
//   3. tiles with more than CMP_CAP sources (the 935-layer group of demo/material-design.svg) repeat 1-2.
#ifndef SVGR_CMP_OCC
#define SVGR_CMP_OCC 4
#endif
// SIMPLE: every op of the launch is an OVER fold whose sources need no colour conversion and no pattern gather
// (what an icon batch consists of); the rare branches are compiled out, which keeps the hot loop's code compact.
template <bool SIMPLE>
__global__ void __launch_bounds__(256, SVGR_CMP_OCC)
compose_kernel(RenderTables T, const OpRec *__restrict__ ops, const TileHead *__restrict__ heads,
               const TileEntry *__restrict__ list, float *__restrict__ layers_out, uint8_t *__restrict__ canvas_out)
{
    __shared__ __align__(16) OpRec s_op;
    __shared__ __align__(16) SrcRec s_src[CMP_CAP + 1];
    __shared__ __align__(16) PaintRec s_paint[CMP_CAP + 1];
    // per staged source, relative to this tile: bit r of s_rm = tile row r is inside the source (and the op), bit c
    // of s_cm likewise for columns, s_toff = element offset of the tile's top-left pixel in the source's storage
    __shared__ unsigned s_rm[CMP_CAP + 1], s_cm[CMP_CAP + 1];
    __shared__ int s_toff[CMP_CAP + 1];

    const int tid = threadIdx.x;
    const int4 head = __ldg(reinterpret_cast<const int4 *>(heads) + blockIdx.x);  // op, entries, first entry
    const int n_list = head.y;
    const TileEntry *entries = list + head.z;
    static_assert(sizeof(OpRec) % 8 == 0, "OpRec must be a multiple of 8 bytes");
    static_assert(sizeof(PaintRec) == 224, "PaintRec layout");
    static_assert(4 * (CMP_CAP + 1) <= 256, "one staging pass");

    // `take` entries from `start` on go to shared memory; a stencil modifier stays with the source before it
    auto round_size = [&](int start) {
        int take = min(CMP_CAP, n_list - start);
        if (start + take < n_list && __ldg(&entries[start + take].paint) == -2)
            take--;
        return take;
    };
    auto stage = [&](int start, int take) {
        if (tid < 4 * take) {
            const int j = tid >> 2, part = tid & 3;
            const TileEntry *e = entries + start + j;
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(e));
            copy16(&s_src[j], T.srcs + (int)a.x, part);
            if (part == 0) {
                s_rm[j] = a.z, s_cm[j] = a.w;
                s_toff[j] = __ldg(&e->toff);
            }
        }
        for (int q = tid; q < take * 14; q += 256) {
            const int j = q / 14, part = q - j * 14;
            const int paint = __ldg(&entries[start + j].paint);
            if (paint >= 0)
                copy16(&s_paint[j], T.paints + paint, part);
        }
    };

    if (tid < (int)(sizeof(OpRec) / 8))
        reinterpret_cast<uint2 *>(&s_op)[tid] = __ldg(reinterpret_cast<const uint2 *>(ops + head.x) + tid);
    int start = 0, n = round_size(0);
    stage(0, n);
    __syncthreads();
    const OpRec &op = s_op;
    const int local = blockIdx.x - op.tile_base;
    const int tr = local / op.ntile_c, tc = local - tr * op.ntile_c;
    const int ty = tid >> 5, tx = tid & 31;
    const int lr0 = tr * CMP_TR + ty, lc = tc * CMP_TC + tx;  // output-local; pixel k is 8 k rows further down
    const int r0 = op.r0 + lr0, c = op.c0 + lc;
    const bool col_live = lc < op.cols;
    const int mode = op.mode;
    const bool skip_outside = SIMPLE || (mode == MODE_OVER);  // blending a zero source is the identity for OVER
    const double x0 = (double)r0 + 0.5, y0 = (double)c + 0.5;  // pixel centre of pixel 0

    float4 acc[CMP_PX];
#pragma unroll
    for (int k = 0; k < CMP_PX; k++)
        acc[k] = f4(0.f, 0.f, 0.f, 0.f);

    for (;;) {
        // ---- fold: everything that does not depend on the pixel is hoisted out of the 4-pixel loops
        if (col_live) {
            for (int j = 0; j < n; j++) {
                const SrcRec &s = s_src[j];
                if (s.kind >= SRC_MOD_COV)
                    continue;  // consumed together with its owner below
                const bool first = start + j == 0;  // source 0 of the op is always listed first
                // pixel k of this thread (row ty + 8 k of the tile) is inside the source: bit 8 k of `live`
                const unsigned live = ((s_rm[j] >> ty) & 0x01010101u) * ((s_cm[j] >> tx) & 1u);
                if (!live && skip_outside && !first)
                    continue;
                const int base = s_toff[j] + ty * s.stride + tx, step = 8 * s.stride;  // a layer has < 2^31 px
                // a stencil modifier (clip / mask fused into this source) is one more dependent load per pixel: its
                // scalar kinds are requested here, together with the source's own data, instead of after the paint
                const bool has_mod = j + 1 < n && s_src[j + 1].kind >= SRC_MOD_COV;
                const int mod_kind = has_mod ? s_src[j + 1].kind : 0;
                float mval[CMP_PX];
                if (has_mod && mod_kind != SRC_MOD_LUMA) {
                    const SrcRec &md = s_src[j + 1];
                    const float *mp = (mod_kind == SRC_MOD_COV ? T.cov : T.layers) + md.off;
                    const int esz = mod_kind == SRC_MOD_L4A ? 4 : 1, eoff = mod_kind == SRC_MOD_L4A ? 3 : 0;
#pragma unroll
                    for (int k = 0; k < CMP_PX; k++) {
                        const long long idx = (long long)(r0 + 8 * k - md.br0) * md.stride + (c - md.bc0);
                        mval[k] = (live >> (8 * k) & 1) ? ld_stream(mp + esz * idx + eoff) : 0.f;
                    }
                }
                float4 v[CMP_PX];
                if (s.kind == SRC_L4) {
                    const float4 *p = reinterpret_cast<const float4 *>(T.layers + s.off);
#pragma unroll
                    for (int k = 0; k < CMP_PX; k++)
                        v[k] = (live >> (8 * k) & 1) ? ld_stream(p + (base + k * step)) : f4(0.f, 0.f, 0.f, 0.f);
                } else {
                    const float *p = (s.kind == SRC_L1 ? T.layers : T.cov) + s.off;
                    float a[CMP_PX];
#pragma unroll
                    for (int k = 0; k < CMP_PX; k++)
                        a[k] = (live >> (8 * k) & 1) ? ld_stream(p + (base + k * step)) : 0.f;
                    // nothing of the path in these four pixels (the inside of a stroked ring, the corners of a
                    // blob's box): an all-zero source is the identity of the over blend
                    if (skip_outside && !first && a[0] == 0.f && a[1] == 0.f && a[2] == 0.f && a[3] == 0.f)
                        continue;
                    if (s.kind == SRC_COVPAINT) {
                        const PaintRec &pr = s_paint[j];
                        if (pr.kind == PAINT_SOLID) {
                            const float4 col = f4(pr.color[0], pr.color[1], pr.color[2], pr.color[3]);
#pragma unroll
                            for (int k = 0; k < CMP_PX; k++)
                                v[k] = scale4(col, a[k]);
                        } else {
                            const StopRec *st = T.stops + pr.stop_off;
                            if (!SIMPLE && pr.kind == PAINT_PATTERN) {
                                const float4 *pat = reinterpret_cast<const float4 *>(T.layers + s.off2);
#pragma unroll
                                for (int k = 0; k < CMP_PX; k++) {
                                    v[k] = f4(0.f, 0.f, 0.f, 0.f);
                                    if (a[k] != 0.f)
                                        v[k] = scale4(paint_eval(T, pr, st, x0 + 8.0 * k, y0, pat, s.stride2), a[k]);
                                }
                            } else {
                                // gradient: parameters of the covered pixels first, then one pass over the stops
                                double t[CMP_PX];
                                bool need[CMP_PX];
#pragma unroll
                                for (int k = 0; k < CMP_PX; k++) {
                                    t[k] = 0.0, need[k] = false;
                                    if (a[k] != 0.f)
                                        t[k] = grad_param(T, pr, x0 + 8.0 * k, y0, &need[k]);
                                }
                                grad_colors<CMP_PX>(t, need, st, pr.stop_cnt, v);
#pragma unroll
                                for (int k = 0; k < CMP_PX; k++)
                                    v[k] = scale4(v[k], a[k]);
                            }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < CMP_PX; k++)
                            v[k] = f4(a[k], a[k], a[k], a[k]);
                    }
                }
                if (s.mul != 1.0f) {
                    const float m = s.mul;
#pragma unroll
                    for (int k = 0; k < CMP_PX; k++)
                        v[k] = scale4(v[k], m);
                }
                if (!SIMPLE && s.kind != SRC_L1 && s.kind != SRC_COV && !conv_is_identity(s.conv)) {
#pragma unroll
                    for (int k = 0; k < CMP_PX; k++)
                        if (live >> (8 * k) & 1)
                            v[k] = convert_px(v[k], s.conv);
                }
                if (has_mod) {
                    // stencil of a clip / luminance mask that was never written out as a layer
                    const SrcRec &md = s_src[j + 1];
                    if (mod_kind != SRC_MOD_LUMA) {
                        const float mm = md.mul;
#pragma unroll
                        for (int k = 0; k < CMP_PX; k++)
                            v[k] = scale4(v[k], mval[k] * mm);  // dead pixels have v = 0
                    } else {
#pragma unroll
                        for (int k = 0; k < CMP_PX; k++)
                            if (live >> (8 * k) & 1) {
                                const float m = mod_value(T, md, r0 + 8 * k, c);
                                v[k] = scale4(v[k], m);
                            }
                    }
                }
                if (first) {
#pragma unroll
                    for (int k = 0; k < CMP_PX; k++)
                        acc[k] = v[k];
                } else if (SIMPLE || mode == MODE_OVER) {
#pragma unroll
                    for (int k = 0; k < CMP_PX; k++) {
                        // a dead pixel has v = 0: the blend is the identity
                        acc[k] = madd4(acc[k], 1.0f - v[k].w, v[k]);
                    }
                } else if (mode == MODE_IN) {
#pragma unroll
                    for (int k = 0; k < CMP_PX; k++) {
                        acc[k] = scale4(v[k], acc[k].w);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < CMP_PX; k++)
                        acc[k] = blend_px(mode, op.k, acc[k], v[k]);
                }
            }
        }
        start += n;
        if (start >= n_list)
            break;
        __syncthreads();  // everybody is done with the staged records
        n = round_size(start);
        stage(start, n);
        __syncthreads();
    }
    if (!col_live)
        return;

#pragma unroll
    for (int k = 0; k < CMP_PX; k++) {
        const int lr = lr0 + 8 * k;
        if (lr >= op.rows)
            continue;
        float4 a = acc[k];
        if (op.mul != 1.0f)
            a = scale4(a, op.mul);
        if (op.post & POST_CLIP01)
            a = f4(clip01(a.x), clip01(a.y), clip01(a.z), clip01(a.w));
        if (!SIMPLE && (op.post & POST_ALPHA))
            a = f4(0.f, 0.f, 0.f, a.w);
        if (!SIMPLE && (op.post & POST_MATRIX)) {
            const float *M = T.matrices + 20 * op.aux;  // row-major 4x5
            float4 v = a;
            a.x = clip01(M[0] * v.x + M[1] * v.y + M[2] * v.z + M[3] * v.w + M[4]);
            a.y = clip01(M[5] * v.x + M[6] * v.y + M[7] * v.z + M[8] * v.w + M[9]);
            a.z = clip01(M[10] * v.x + M[11] * v.y + M[12] * v.z + M[13] * v.w + M[14]);
            a.w = clip01(M[15] * v.x + M[16] * v.y + M[17] * v.z + M[18] * v.w + M[19]);
        }
        const long long idx = (long long)lr * op.stride + lc;
        if (op.kind == OP_CANVAS) {
            // main() :3870-3881: the premultiplied result over a zero canvas, clipped to [0, 1]
            // (canvas_merge_at :326), converted to straight-alpha sRGB (Layer.write_png :212) and quantised
            // with round-half-even (np.round, :263).  op.k1 carries the render's linear_rgb flag.
            if (!SIMPLE && (op.post & POST_MATRIX))
                // a colour matrix folded into the canvas pass: its result is straight-alpha linear; main() first
                // converts the layer to premultiplied render space (:3871)
                a = convert_px(a, SVGR_CONV(0, 1, 1, op.k1 != 0));
            a = f4(clip01(a.x), clip01(a.y), clip01(a.z), clip01(a.w));
            if (!SIMPLE && op.k1 != 0)
                a = convert_px(a, SVGR_CONV(1, 1, 0, 0));
            else
                a = unpremultiply(a);  // sRGB render mode: Layer.convert is the alpha division only
            // x * 255 + 1.5 * 2^23 leaves round-half-even(x * 255) in the low mantissa bits
            uchar4 q;
            q.x = (unsigned char)(__float_as_uint(a.x * 255.0f + 12582912.0f) & 0xffu);
            q.y = (unsigned char)(__float_as_uint(a.y * 255.0f + 12582912.0f) & 0xffu);
            q.z = (unsigned char)(__float_as_uint(a.z * 255.0f + 12582912.0f) & 0xffu);
            q.w = (unsigned char)(__float_as_uint(a.w * 255.0f + 12582912.0f) & 0xffu);
            reinterpret_cast<uchar4 *>(canvas_out + op.out_off)[idx] = q;
        } else if (!SIMPLE && (op.post & POST_LUMA)) {
            // Scene.render mask branch: luma . rgb * alpha on the straight-alpha image (svgrasterize.py:734-736)
            layers_out[op.out_off + idx] = (a.x * 0.2125f + a.y * 0.7154f + a.z * 0.072f) * a.w;
        } else if (!SIMPLE && op.out_ch == 1) {
            layers_out[op.out_off + idx] = a.w;
        } else {
            reinterpret_cast<float4 *>(layers_out + op.out_off)[idx] = a;
        }
    }
}

// any(det < 0) over the mask bbox of a two-circle gradient fill (svgrasterize.py:1621-1622): the
// reference masks invalid pixels only when at least one exists anywhere in the layer.
__global__ void focal_flag_kernel(RenderTables T, const FocalJob *__restrict__ jobs, int n_jobs, int block0,
                                  int *__restrict__ flags)
{
    const int bid = (int)blockIdx.x + block0;  // block numbers are global over the render, launches cover ranges
    __shared__ int s_job;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_jobs - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].block_base <= bid)
                lo = mid;
            else
                hi = mid - 1;
        }
        s_job = lo;
    }
    __syncthreads();
    const FocalJob &j = jobs[s_job];
    const PaintRec &p = T.paints[j.paint];
    long long base = (long long)(bid - j.block_base) * 1024;
    long long n = (long long)j.rows * j.cols;
    bool neg = false;
    for (int k = 0; k < 4; k++) {
        long long i = base + threadIdx.x + 256 * k;
        if (i < n) {
            int r = j.r0 + (int)(i / j.cols), c = j.c0 + (int)(i % j.cols);
            double b;
            if (focal_det(p, (double)r + 0.5, (double)c + 0.5, &b) < 0)
                neg = true;
        }
    }
    if (__syncthreads_or(neg) && threadIdx.x == 0)
        atomicOr(flags + p.flag, 1);
}

// ---------------------------------------------------------------------------------------------
void svgr_launch_compose(const RenderTables &T, const OpRec *ops, const TileHead *heads, const TileEntry *list, int n_tiles,
                         bool simple, float *layers_out, uint8_t *canvas_out, cudaStream_t s)
{
    if (n_tiles <= 0)
        return;
    if (simple)
        compose_kernel<true><<<n_tiles, 256, 0, s>>>(T, ops, heads, list, layers_out, canvas_out);
    else
        compose_kernel<false><<<n_tiles, 256, 0, s>>>(T, ops, heads, list, layers_out, canvas_out);
}

// Source lists of the tiles of one compose launch, one warp per tile: the op of the tile (binary search over the
// launch's ops, their tile_base column stays in L1/L2), then an order-preserving cull of the op's sources against
// the tile rectangle (ballots).  What the fold needs per (source, tile) and not per pixel -- row / column masks,
// the storage offset of the tile's corner, the paint index -- is computed here once instead of by 256 threads.
// A blend other than OVER keeps every source (a zero source is not its identity); source 0 is always kept.
__global__ void __launch_bounds__(256)
cull_kernel(RenderTables T, const OpRec *__restrict__ ops, int n_ops, int n_tiles, TileHead *__restrict__ heads,
            TileEntry *__restrict__ list)
{
    const int t = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t >= n_tiles)
        return;
    int lo = 0, hi = n_ops - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(&ops[mid].tile_base) <= t)
            lo = mid;
        else
            hi = mid - 1;
    }
    const OpRec *op = ops + lo;
    const int o_r0 = __ldg(&op->r0), o_c0 = __ldg(&op->c0), o_rows = __ldg(&op->rows), o_cols = __ldg(&op->cols);
    const int src_off = __ldg(&op->src_off), src_cnt = __ldg(&op->src_cnt), ntile_c = __ldg(&op->ntile_c);
    const bool keep_all = __ldg(&op->mode) != MODE_OVER;
    const int local = t - __ldg(&op->tile_base);
    const int tr = local / ntile_c, tc = local - tr * ntile_c;
    const int tile_r0 = o_r0 + tr * CMP_TR, tile_c0 = o_c0 + tc * CMP_TC;
    const int tile_r1 = min(tile_r0 + CMP_TR, o_r0 + o_rows), tile_c1 = min(tile_c0 + CMP_TC, o_c0 + o_cols);
    const int first = __ldg(&op->list_base) + local * src_cnt;
    TileEntry *out = list + first;
    int n = 0;
    for (int k = 0; k < src_cnt; k += 32) {
        const int i = k + lane;
        bool hit = false;
        TileEntry e;
        e.src = src_off + i, e.paint = -1, e.rm = 0u, e.cm = 0u, e.toff = 0, e.pad[0] = e.pad[1] = e.pad[2] = 0;
        if (i < src_cnt) {
            const int4 *g = reinterpret_cast<const int4 *>(T.srcs + src_off + i);
            const int4 a = __ldg(g), b = __ldg(g + 1), d = __ldg(g + 3);  // kind r0 c0 rows | cols stride paint conv | .. br0 bc0
            const int ra = max(a.y, tile_r0), rb = min(a.y + a.w, tile_r1);
            const int ca = max(a.z, tile_c0), cb = min(a.z + b.x, tile_c1);
            e.rm = rb > ra ? (0xffffffffu >> (32 - (rb - ra))) << (ra - tile_r0) : 0u;
            e.cm = cb > ca ? (0xffffffffu >> (32 - (cb - ca))) << (ca - tile_c0) : 0u;
            e.toff = (tile_r0 - d.z) * b.y + (tile_c0 - d.w);
            e.paint = a.x >= SRC_MOD_COV ? -2 : (a.x == SRC_COVPAINT ? b.z : -1);
            // a modifier carries its owner's rectangle: the two are kept or dropped together
            hit = i == 0 || keep_all || (e.rm != 0u && e.cm != 0u);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) {
            uint4 *dst = reinterpret_cast<uint4 *>(out + n + __popc(m & ((1u << lane) - 1)));
            dst[0] = make_uint4((unsigned)e.src, (unsigned)e.paint, e.rm, e.cm);
            dst[1] = make_uint4((unsigned)e.toff, 0u, 0u, 0u);
        }
        n += __popc(m);
    }
    if (lane == 0)
        reinterpret_cast<int4 *>(heads)[t] = make_int4(lo, n, first, 0);
}

void svgr_launch_cull(const RenderTables &T, const OpRec *ops, int n_ops, int n_tiles, TileHead *heads, TileEntry *list,
                      cudaStream_t s)
{
    if (n_ops > 0 && n_tiles > 0)
        cull_kernel<<<(n_tiles + 7) / 8, 256, 0, s>>>(T, ops, n_ops, n_tiles, heads, list);
}

// tile -> op table of one launch: op i owns tiles [ops[i].tile_base, ops[i + 1].tile_base).  One thread per
// tile, binary search over the ops (their tile_base column stays in L1/L2): a single 8192 x 8192 op has 65 536
// tiles, so this must not be a per-op loop.
__global__ void expand_ops_kernel(const OpRec *__restrict__ ops, int n_ops, int n_tiles, int *__restrict__ tile_op)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles)
        return;
    int lo = 0, hi = n_ops - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(&ops[mid].tile_base) <= t)
            lo = mid;
        else
            hi = mid - 1;
    }
    tile_op[t] = lo;
}

void svgr_launch_expand_ops(const OpRec *ops, int n_ops, int n_tiles, int *tile_op, cudaStream_t s)
{
    if (n_ops > 0 && n_tiles > 0)
        expand_ops_kernel<<<(n_tiles + 255) / 256, 256, 0, s>>>(ops, n_ops, n_tiles, tile_op);
}

void svgr_launch_focal_flags(const RenderTables &T, const void *jobs, int n_jobs, int block0, int n_blocks, int *flags,
                             cudaStream_t s)
{
    if (n_blocks > 0)
        focal_flag_kernel<<<n_blocks, 256, 0, s>>>(T, (const FocalJob *)jobs, n_jobs, block0, flags);
}

// ---------------------------------------------------------------------------------------------
// eager element-wise entry points of the reference's call surface (svgr_grad_* / svgr_quantize_u8)
// ---------------------------------------------------------------------------------------------
// grad_pixels (svgrasterize.py:1653-1658): centres of the pixels of a viewport
__global__ void grad_pixels_kernel(int r0, int c0, long long n, int cols, double2 *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols, c = i - r * cols;
        out[i] = make_double2((double)r + ((double)r0 + 0.5), (double)c + ((double)c0 + 0.5));
    }
}

// grad_spread (svgrasterize.py:1661-1668)
__global__ void grad_spread_kernel(const double *__restrict__ in, long long n, int spread, double *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = grad_spread(in[i], spread);
}

// grad_interpolate (svgrasterize.py:1671-1683)
__global__ void grad_interpolate_kernel(const double *__restrict__ in, long long n, const StopRec *__restrict__ stops,
                                        int n_stops, float4 *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = grad_color(in[i], stops, n_stops);
}

// np.round(x * 255).astype(np.uint8) (svgrasterize.py:263): round half even; values outside [0, 255] wrap like the
// integer cast does on x86-64 (the low byte of the rounded integer), NaN -> 0
__global__ void quantize_kernel(const float *__restrict__ in, long long n, uint8_t *__restrict__ out)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = rintf(in[i] * 255.0f);
        out[i] = (v == v) ? (uint8_t)((long long)v & 0xff) : (uint8_t)0;
    }
}

static unsigned eager_blocks(long long n) { return (unsigned)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148 * 32)); }

void svgr_launch_grad_pixels(int r0, int c0, int rows, int cols, double *out, cudaStream_t s)
{
    const long long n = (long long)rows * cols;
    if (n > 0)
        grad_pixels_kernel<<<eager_blocks(n), 256, 0, s>>>(r0, c0, n, cols, reinterpret_cast<double2 *>(out));
}

void svgr_launch_grad_spread(const double *in, long long n, int spread, double *out, cudaStream_t s)
{
    if (n > 0)
        grad_spread_kernel<<<eager_blocks(n), 256, 0, s>>>(in, n, spread, out);
}

void svgr_launch_grad_interpolate(const double *in, long long n, const StopRec *stops, int n_stops, float *out,
                                  cudaStream_t s)
{
    if (n > 0)
        grad_interpolate_kernel<<<eager_blocks(n), 256, 0, s>>>(in, n, stops, n_stops, reinterpret_cast<float4 *>(out));
}

void svgr_launch_quantize(const float *in, long long n, uint8_t *out, cudaStream_t s)
{
    if (n > 0)
        quantize_kernel<<<eager_blocks(n), 256, 0, s>>>(in, n, out);
}
