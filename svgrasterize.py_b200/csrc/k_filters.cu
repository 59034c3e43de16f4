// k_filters.cu -- filter stencils (subsystem 4 of the hot path).
//
// Replaces:
//   Layer.convolve = scipy.signal.convolve(image, kernel[..., None]), mode "full"
//                                   svgrasterize.py:106-118 (kernel from blur_kernel, :1903-1944)
//   Layer.morphology + pooling      svgrasterize.py:120-127, :419-468 (nanmax / nanmin, stride 1, no padding)
//
// A Gaussian whose quadratic form is diagonal is an outer product
// g_row x g_col, and a k0 x k1 max/min window is separable by construction, so
// both run as two shared-memory staged 1-D passes (columns first, then rows):
//   OP_STENCIL_H   out[r, j] = OP_t  w[t] . in[r, j -/+ t]     tile 8 rows x 128 cols
//   OP_STENCIL_V   out[i, c] = OP_t  w[t] . in[i -/+ t, c]     tile 64 rows x 32 cols
// ("-" for the full-mode convolution, "+" for the top-left anchored pooling window).
// The first pass reads its source through fetch_src, so the reference's
// Layer.convert (straight alpha + linear RGB for blur, premultiplied linear for
// morphology) is fused into the load.  A rotated anisotropic Gaussian is not
// separable; OP_CONV2D applies the full kw x kh kernel directly.
#include <atomic>

#include "svgr_device.cuh"

__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// np.nanmax / np.nanmin of two values: IEEE maxNum / minNum (one FMNMX): the operand that is not NaN wins
__device__ __forceinline__ float nanmax1(float a, float v) { return fmaxf(a, v); }
__device__ __forceinline__ float nanmin1(float a, float v) { return fminf(a, v); }

template <int ST>
__device__ __forceinline__ float4 stencil_init()
{
    if (ST == STENCIL_CONV)
        return zero4();
    float n = __int_as_float(0x7fc00000);
    return make_float4(n, n, n, n);
}

template <int ST>
__device__ __forceinline__ void stencil_acc(float4 &a, float w, float4 v)
{
    if (ST == STENCIL_CONV) {
        a = madd4(v, w, a);  // two FFMA2
    } else if (ST == STENCIL_MAX) {
        a.x = nanmax1(a.x, v.x), a.y = nanmax1(a.y, v.y), a.z = nanmax1(a.z, v.z), a.w = nanmax1(a.w, v.w);
    } else {
        a.x = nanmin1(a.x, v.x), a.y = nanmin1(a.y, v.y), a.z = nanmin1(a.z, v.z), a.w = nanmin1(a.w, v.w);
    }
}

// source pixel at source-local (i, j), zero outside
__device__ __forceinline__ float4 load_local(const RenderTables &T, const SrcRec &s, int i, int j)
{
    if (i < 0 || j < 0 || i >= s.rows || j >= s.cols)
        return zero4();
    return fetch_src(T, s, s.r0 + i, s.c0 + j);
}

// Horizontal pass.  Shared tile: SVGR_STH_TR rows x (SVGR_STH_TC + k - 1) pixels.
template <int ST>
__device__ __forceinline__ void stencil_h_body(const RenderTables &T, const OpRec &op, int local, float4 *sm,
                                               float *__restrict__ layers_out)
{
    const SrcRec &s = T.srcs[op.src_off];
    const int k = op.k0;
    const int tr = local / op.ntile_c, tc = local - tr * op.ntile_c;
    const int row0 = tr * SVGR_STH_TR, col0 = tc * SVGR_STH_TC;
    const int span = SVGR_STH_TC + k - 1;
    // conv: out j reads in j - t, t in [0, k)  -> in range [col0 - k + 1, col0 + TC)
    // pool: out j reads in j + t               -> in range [col0, col0 + TC + k - 1)
    const int in0 = (ST == STENCIL_CONV) ? col0 - k + 1 : col0;
    for (int i = threadIdx.x; i < SVGR_STH_TR * span; i += 256) {
        int rr = i / span, cc = i - rr * span;
        sm[i] = load_local(T, s, row0 + rr, in0 + cc);
    }
    __syncthreads();
    const float *w = T.weights + op.aux;
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const int lr = row0 + ty;
    if (lr >= op.rows)
        return;
    const float4 *row = sm + ty * span;
#pragma unroll
    for (int q = 0; q < SVGR_STH_TC / 32; q++) {
        int jc = tx + 32 * q;  // tile-local output column
        int lc = col0 + jc;
        if (lc >= op.cols)
            continue;
        float4 a = stencil_init<ST>();
        if (ST == STENCIL_CONV) {
            for (int t = 0; t < k; t++)
                stencil_acc<ST>(a, w[t], row[jc + (k - 1) - t]);
        } else {
            for (int t = 0; t < k; t++)
                stencil_acc<ST>(a, 0.f, row[jc + t]);
        }
        reinterpret_cast<float4 *>(layers_out + op.out_off)[(long long)lr * op.stride + lc] = a;
    }
}

// Vertical pass.  Shared tile: (SVGR_STV_TR + k - 1) rows x SVGR_STV_TC pixels.
template <int ST>
__device__ __forceinline__ void stencil_v_body(const RenderTables &T, const OpRec &op, int local, float4 *sm,
                                               float *__restrict__ layers_out)
{
    const SrcRec &s = T.srcs[op.src_off];
    const int k = op.k0;
    const int tr = local / op.ntile_c, tc = local - tr * op.ntile_c;
    const int row0 = tr * SVGR_STV_TR, col0 = tc * SVGR_STV_TC;
    const int span = SVGR_STV_TR + k - 1;
    const int in0 = (ST == STENCIL_CONV) ? row0 - k + 1 : row0;
    for (int i = threadIdx.x; i < span * SVGR_STV_TC; i += 256) {
        int rr = i / SVGR_STV_TC, cc = i - rr * SVGR_STV_TC;
        sm[i] = load_local(T, s, in0 + rr, col0 + cc);
    }
    __syncthreads();
    const float *w = T.weights + op.aux;
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const int lc = col0 + tx;
    if (lc >= op.cols)
        return;
#pragma unroll
    for (int q = 0; q < SVGR_STV_TR / 8; q++) {
        int ir = ty + 8 * q;  // tile-local output row
        int lr = row0 + ir;
        if (lr >= op.rows)
            continue;
        float4 a = stencil_init<ST>();
        if (ST == STENCIL_CONV) {
            for (int t = 0; t < k; t++)
                stencil_acc<ST>(a, w[t], sm[(ir + (k - 1) - t) * SVGR_STV_TC + tx]);
        } else {
            for (int t = 0; t < k; t++)
                stencil_acc<ST>(a, 0.f, sm[(ir + t) * SVGR_STV_TC + tx]);
        }
        reinterpret_cast<float4 *>(layers_out + op.out_off)[(long long)lr * op.stride + lc] = a;
    }
}

__global__ void __launch_bounds__(256)
stencil_kernel(RenderTables T, const OpRec *__restrict__ ops, const int *__restrict__ tile_op,
               float *__restrict__ layers_out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *sm = reinterpret_cast<float4 *>(smem_raw);
    const OpRec &op = ops[__ldg(tile_op + blockIdx.x)];
    const int local = blockIdx.x - op.tile_base;
    if (op.kind == OP_STENCIL_H) {
        if (op.stencil == STENCIL_CONV)
            stencil_h_body<STENCIL_CONV>(T, op, local, sm, layers_out);
        else if (op.stencil == STENCIL_MAX)
            stencil_h_body<STENCIL_MAX>(T, op, local, sm, layers_out);
        else
            stencil_h_body<STENCIL_MIN>(T, op, local, sm, layers_out);
    } else {
        if (op.stencil == STENCIL_CONV)
            stencil_v_body<STENCIL_CONV>(T, op, local, sm, layers_out);
        else if (op.stencil == STENCIL_MAX)
            stencil_v_body<STENCIL_MAX>(T, op, local, sm, layers_out);
        else
            stencil_v_body<STENCIL_MIN>(T, op, local, sm, layers_out);
    }
}

// Direct 2-D "full" convolution: out[i, j] = sum_{a, b} k[a, b] . in[i - a, j - b]; tile 8 x 32.  The kernel rows
// are walked in groups of `ga` rows whose input rows (8 + ga - 1 of them) fit the staging buffer, so the kernel
// height is unbounded (ga = kr when everything fits at once, the common case).
__global__ void __launch_bounds__(256)
conv2d_kernel(RenderTables T, const OpRec *__restrict__ ops, const int *__restrict__ tile_op,
              float *__restrict__ layers_out, int smem_bytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *sm = reinterpret_cast<float4 *>(smem_raw);
    const OpRec &op = ops[__ldg(tile_op + blockIdx.x)];
    const SrcRec &s = T.srcs[op.src_off];
    const int local = blockIdx.x - op.tile_base;
    const int tr = local / op.ntile_c, tc = local - tr * op.ntile_c;
    const int kr = op.k0, kc = op.k1;
    const int row0 = tr * SVGR_C2D_TR, col0 = tc * SVGR_C2D_TC;
    const int scols = SVGR_C2D_TC + kc - 1;
    const int ga = min(kr, max(1, smem_bytes / (scols * 16) - SVGR_C2D_TR + 1));  // kernel rows per group
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const int lr = row0 + ty, lc = col0 + tx;
    const float *w = T.weights + op.aux;
    float4 acc = zero4();
    for (int a0 = 0; a0 < kr; a0 += ga) {
        const int na = min(ga, kr - a0);        // kernel rows a0 .. a0 + na - 1
        const int srows = SVGR_C2D_TR + na - 1;  // input rows row0 - (a0 + na - 1) .. row0 + 7 - a0
        if (a0 > 0)
            __syncthreads();
        for (int i = threadIdx.x; i < srows * scols; i += 256) {
            int rr = i / scols, cc = i - rr * scols;
            sm[i] = load_local(T, s, row0 - (a0 + na - 1) + rr, col0 - kc + 1 + cc);
        }
        __syncthreads();
        for (int a = 0; a < na; a++)
            for (int b = 0; b < kc; b++) {
                float wt = w[(a0 + a) * kc + b];
                float4 v = sm[(ty + na - 1 - a) * scols + (tx + kc - 1 - b)];
                acc = madd4(v, wt, acc);
            }
    }
    if (lr >= op.rows || lc >= op.cols)
        return;
    reinterpret_cast<float4 *>(layers_out + op.out_off)[(long long)lr * op.stride + lc] = acc;
}

// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// the opt-in is per device; setting it twice is harmless, so the flags only need to be atomic, not locked
static std::atomic<bool> g_attr_set[64];

static void ensure_attrs()
{
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!g_attr_set[dev].load(std::memory_order_acquire)) {
        cudaFuncSetAttribute(stencil_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SVGR_MAX_DYN_SMEM);
        cudaFuncSetAttribute(conv2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SVGR_MAX_DYN_SMEM);
        g_attr_set[dev].store(true, std::memory_order_release);
    }
}

int svgr_launch_stencil(const RenderTables &T, const OpRec *ops, const int *tile_op, int n_tiles, size_t smem_bytes,
                        float *layers_out, cudaStream_t s)
{
    if (n_tiles <= 0)
        return 0;
    if (smem_bytes > SVGR_MAX_DYN_SMEM)
        return -1;
    ensure_attrs();
    stencil_kernel<<<n_tiles, 256, smem_bytes, s>>>(T, ops, tile_op, layers_out);
    return 0;
}

int svgr_launch_conv2d(const RenderTables &T, const OpRec *ops, const int *tile_op, int n_tiles, size_t smem_bytes,
                       float *layers_out, cudaStream_t s)
{
    if (n_tiles <= 0)
        return 0;
    if (smem_bytes > SVGR_MAX_DYN_SMEM)
        return -1;
    ensure_attrs();
    conv2d_kernel<<<n_tiles, 256, smem_bytes, s>>>(T, ops, tile_op, layers_out, (int)smem_bytes);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// pooling(mat, ksize, stride, method, pad) (svgrasterize.py:419-468) in its general form: strided windows, NaN
// padding to ceil(n / s) outputs, max / min / mean that ignore NaN (np.nanmax / nanmin / nanmean).  The hot
// path (Layer.morphology: stride 1, no padding) runs the separable stencils above; this is the call surface.
// ---------------------------------------------------------------------------------------------
__global__ void pooling_kernel(const float *__restrict__ in, int rows, int cols, int ch, int ky, int kx, int sy, int sx,
                               int method, float *__restrict__ out, int orows, int ocols)
{
    const long long n = (long long)orows * ocols * ch;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % ch);
        const long long px = i / ch;
        const int oc = (int)(px % ocols), orow = (int)(px / ocols);
        float acc = __int_as_float(0x7fc00000), sum = 0.f;
        int cnt = 0;
        for (int a = 0; a < ky; a++) {
            const int r = orow * sy + a;
            if (r >= rows)
                break;  // NaN padding: ignored
            for (int b = 0; b < kx; b++) {
                const int q = oc * sx + b;
                if (q >= cols)
                    break;
                const float v = in[((long long)r * cols + q) * ch + c];
                if (method == 0)
                    acc = nanmax1(acc, v);
                else if (method == 1)
                    acc = nanmin1(acc, v);
                else if (v == v)
                    sum += v, cnt++;
            }
        }
        out[i] = method == 2 ? (cnt ? sum / (float)cnt : __int_as_float(0x7fc00000)) : acc;
    }
}

void svgr_launch_pooling(const float *in, int rows, int cols, int ch, int ky, int kx, int sy, int sx, int method,
                         float *out, int orows, int ocols, cudaStream_t s)
{
    const long long n = (long long)orows * ocols * ch;
    if (n <= 0)
        return;
    const long long blocks = (n + 255) / 256;
    pooling_kernel<<<(unsigned)(blocks < 148 * 32 ? blocks : 148 * 32), 256, 0, s>>>(in, rows, cols, ch, ky, kx, sy, sx,
                                                                                  method, out, orows, ocols);
}
