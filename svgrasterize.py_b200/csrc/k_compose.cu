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
// Tiling: CTA = 8 rows x 32 columns, one pixel (one float4, 16 B) per thread, a
// warp covers 512 contiguous bytes of a row.  When an op has many sources (the
// 935-layer group of demo/material-design.svg) warp 0 first culls the source list
// against the tile rectangle into shared memory, preserving order.
#include "svgr_device.cuh"

#define CMP_TR 8
#define CMP_TC 32
#define CMP_CULL_MIN 8
#define CMP_LIST 1024

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

__device__ __forceinline__ int find_op(const OpRec *__restrict__ ops, int n_ops, int tile)
{
    int lo = 0, hi = n_ops - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (ops[mid].tile_base <= tile)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ bool src_hits(const SrcRec &s, int r, int c)
{
    return r >= s.r0 && r < s.r0 + s.rows && c >= s.c0 && c < s.c0 + s.cols;
}

__global__ void __launch_bounds__(CMP_TR *CMP_TC)
compose_kernel(RenderTables T, const OpRec *__restrict__ ops, int n_ops, float *__restrict__ layers_out)
{
    __shared__ int s_op;
    __shared__ int s_list[CMP_LIST];
    __shared__ int s_n, s_next;

    if (threadIdx.x == 0)
        s_op = find_op(ops, n_ops, blockIdx.x);
    __syncthreads();
    const OpRec &op = ops[s_op];
    const int local = blockIdx.x - op.tile_base;
    const int tr = local / op.ntile_c, tc = local - tr * op.ntile_c;
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const int lr = tr * CMP_TR + ty, lc = tc * CMP_TC + tx;  // output-local
    const int r = op.r0 + lr, c = op.c0 + lc;
    const bool live = lr < op.rows && lc < op.cols;
    const SrcRec *srcs = T.srcs + op.src_off;
    const int mode = op.mode;
    const bool skip_outside = (mode == MODE_OVER);  // blending a zero source is the identity for OVER

    float4 acc = f4(0.f, 0.f, 0.f, 0.f);
    if (op.src_cnt <= CMP_CULL_MIN) {
        if (live) {
            for (int k = 0; k < op.src_cnt; k++) {
                const SrcRec &s = srcs[k];
                bool in = src_hits(s, r, c);
                if (k > 0 && !in && skip_outside)
                    continue;
                float4 v = in ? fetch_src(T, s, r, c) : f4(0.f, 0.f, 0.f, 0.f);
                acc = (k == 0) ? v : blend_px(mode, op.k, acc, v);
            }
        }
    } else {
        // order-preserving cull of the source list against the tile rectangle (warp 0), in rounds
        const int tr0 = op.r0 + tr * CMP_TR, tc0 = op.c0 + tc * CMP_TC;
        if (threadIdx.x == 0)
            s_next = 0;
        __syncthreads();
        for (;;) {
            int start = s_next;
            __syncthreads();
            if (start >= op.src_cnt)
                break;
            if (threadIdx.x < 32) {
                int n = 0, k = start;
                for (; k < op.src_cnt && n <= CMP_LIST - 32; k += 32) {
                    int i = k + tx;
                    bool hit = false;
                    if (i < op.src_cnt) {
                        const SrcRec &s = srcs[i];
                        hit = (i == 0 || !skip_outside) ||
                              (s.r0 < tr0 + CMP_TR && s.r0 + s.rows > tr0 && s.c0 < tc0 + CMP_TC && s.c0 + s.cols > tc0);
                    }
                    unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (hit)
                        s_list[n + __popc(m & ((1u << tx) - 1))] = i;
                    n += __popc(m);
                }
                if (tx == 0) {
                    s_n = n;
                    s_next = k < op.src_cnt ? k : op.src_cnt;
                }
            }
            __syncthreads();
            if (live) {
                int n = s_n;
                for (int j = 0; j < n; j++) {
                    int k = s_list[j];
                    const SrcRec &s = srcs[k];
                    bool in = src_hits(s, r, c);
                    if (k > 0 && !in && skip_outside)
                        continue;
                    float4 v = in ? fetch_src(T, s, r, c) : f4(0.f, 0.f, 0.f, 0.f);
                    acc = (k == 0) ? v : blend_px(mode, op.k, acc, v);
                }
            }
            __syncthreads();
        }
    }
    if (!live)
        return;

    if (op.mul != 1.0f)
        acc = f4(acc.x * op.mul, acc.y * op.mul, acc.z * op.mul, acc.w * op.mul);
    if (op.post & POST_CLIP01)
        acc = f4(clip01(acc.x), clip01(acc.y), clip01(acc.z), clip01(acc.w));
    if (op.post & POST_ALPHA)
        acc = f4(0.f, 0.f, 0.f, acc.w);
    if (op.post & POST_MATRIX) {
        const float *M = T.matrices + 20 * op.aux;  // row-major 4x5
        float4 v = acc;
        acc.x = clip01(M[0] * v.x + M[1] * v.y + M[2] * v.z + M[3] * v.w + M[4]);
        acc.y = clip01(M[5] * v.x + M[6] * v.y + M[7] * v.z + M[8] * v.w + M[9]);
        acc.z = clip01(M[10] * v.x + M[11] * v.y + M[12] * v.z + M[13] * v.w + M[14]);
        acc.w = clip01(M[15] * v.x + M[16] * v.y + M[17] * v.z + M[18] * v.w + M[19]);
    }
    long long idx = (long long)lr * op.stride + lc;
    if (op.post & POST_LUMA) {
        // Scene.render mask branch: luma . rgb * alpha on the straight-alpha image (svgrasterize.py:734-736)
        float l = (acc.x * 0.2125f + acc.y * 0.7154f + acc.z * 0.072f) * acc.w;
        layers_out[op.out_off + idx] = l;
    } else if (op.out_ch == 1) {
        layers_out[op.out_off + idx] = acc.w;
    } else {
        reinterpret_cast<float4 *>(layers_out + op.out_off)[idx] = acc;
    }
}

// Final step of a render (svgrasterize.py:3870-3881): the root layer, converted to the render's
// premultiplied colour space, is blitted over a zero canvas and clipped to [0, 1] (canvas_merge_at),
// converted to straight-alpha sRGB and quantised.  op.mode carries the render's linear_rgb flag,
// op.src_cnt == 0 means "nothing rendered" (the canvas stays transparent).
__global__ void __launch_bounds__(CMP_TR *CMP_TC)
canvas_kernel(RenderTables T, const OpRec *__restrict__ ops, int n_ops, uint8_t *__restrict__ out)
{
    __shared__ int s_op;
    if (threadIdx.x == 0)
        s_op = find_op(ops, n_ops, blockIdx.x);
    __syncthreads();
    const OpRec &op = ops[s_op];
    const int local = blockIdx.x - op.tile_base;
    const int tr = local / op.ntile_c, tc = local - tr * op.ntile_c;
    const int lr = tr * CMP_TR + (threadIdx.x >> 5), lc = tc * CMP_TC + (threadIdx.x & 31);
    if (lr >= op.rows || lc >= op.cols)
        return;
    const int r = op.r0 + lr, c = op.c0 + lc;
    float4 v = f4(0.f, 0.f, 0.f, 0.f);
    if (op.src_cnt > 0) {
        const SrcRec &s = T.srcs[op.src_off];
        if (src_hits(s, r, c)) {
            v = fetch_src(T, s, r, c);
            v = f4(clip01(v.x), clip01(v.y), clip01(v.z), clip01(v.w));
        }
    }
    v = convert_px(v, SVGR_CONV(1, op.mode != 0, 0, 0));
    uchar4 q;
    q.x = (unsigned char)__float2int_rn(v.x * 255.0f);
    q.y = (unsigned char)__float2int_rn(v.y * 255.0f);
    q.z = (unsigned char)__float2int_rn(v.z * 255.0f);
    q.w = (unsigned char)__float2int_rn(v.w * 255.0f);
    reinterpret_cast<uchar4 *>(out + op.out_off)[(long long)lr * op.stride + lc] = q;
}

// any(det < 0) over the mask bbox of a two-circle gradient fill (svgrasterize.py:1621-1622): the
// reference masks invalid pixels only when at least one exists anywhere in the layer.
__global__ void focal_flag_kernel(RenderTables T, const FocalJob *__restrict__ jobs, int n_jobs, int *__restrict__ flags)
{
    __shared__ int s_job;
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_jobs - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (jobs[mid].block_base <= (int)blockIdx.x)
                lo = mid;
            else
                hi = mid - 1;
        }
        s_job = lo;
    }
    __syncthreads();
    const FocalJob &j = jobs[s_job];
    const PaintRec &p = T.paints[j.paint];
    long long base = (long long)(blockIdx.x - j.block_base) * 1024;
    long long n = (long long)j.rows * j.cols;
    bool neg = false;
    for (int k = 0; k < 4; k++) {
        long long i = base + threadIdx.x + 256 * k;
        if (i < n) {
            int r = j.r0 + (int)(i / j.cols), c = j.c0 + (int)(i % j.cols);
            double ux, uy, b, a;
            px_to_user(p, r, c, &ux, &uy);
            if (focal_det(p, ux, uy, &b, &a) < 0)
                neg = true;
        }
    }
    if (__syncthreads_or(neg) && threadIdx.x == 0)
        atomicOr(flags + p.flag, 1);
}

// ---------------------------------------------------------------------------------------------
void svgr_launch_compose(const RenderTables &T, const OpRec *ops, int n_ops, int n_tiles, float *layers_out,
                         cudaStream_t s)
{
    if (n_tiles > 0)
        compose_kernel<<<n_tiles, CMP_TR * CMP_TC, 0, s>>>(T, ops, n_ops, layers_out);
}

void svgr_launch_canvas(const RenderTables &T, const OpRec *ops, int n_ops, int n_tiles, uint8_t *out, cudaStream_t s)
{
    if (n_tiles > 0)
        canvas_kernel<<<n_tiles, CMP_TR * CMP_TC, 0, s>>>(T, ops, n_ops, out);
}

void svgr_launch_focal_flags(const RenderTables &T, const void *jobs, int n_jobs, int n_blocks, int *flags,
                             cudaStream_t s)
{
    if (n_blocks > 0)
        focal_flag_kernel<<<n_blocks, 256, 0, s>>>(T, (const FocalJob *)jobs, n_jobs, flags);
}
