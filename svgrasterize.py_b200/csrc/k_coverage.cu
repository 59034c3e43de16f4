// k_coverage.cu -- edge binning + anti-aliased coverage (subsystem 2 of the hot path).
//
// Replaces, for every path of a batch at once:
//   line_signed_coverage   svgrasterize.py:2213-2304 (pure-Python scalar loop, the #1 cost)
//   np.cumsum + fill rule  svgrasterize.py:983-990
//
// Layout: every path owns a float32 mask (rows x stride, stride = cols rounded
// up to 4 floats so that rows start 16-byte aligned) inside the coverage arena.
// A mask is cut into bands of SVGR_BAND_ROWS rows; edges are binned to the bands
// whose rows they cross (count -> scan -> fill), and one CTA renders one tile =
// (band, chunk of SVGR_TILE_COLS columns):
//   1. signed-area deltas of the band's edges are accumulated into a shared-memory
//      trace tile (32-bit fixed-point shared atomics: deterministic; the edge arithmetic itself is float64),
//   2. each warp prefix-sums rows with register scans + warp shuffles,
//   3. the fill rule and the 1e-6 snap are applied and the row leaves as 128-bit stores.
// The trace never touches HBM: traffic is 36 B per binned edge in, 4 B per pixel out.
//
// Column chunks are independent because of the reference's own clamp rule
// (SURVEY A7): coverage that falls left of column 0 is accumulated INTO column
// 0; a chunk treats everything left of it the same way, which is exactly the
// running winding number the prefix sum would have carried in.
#include <algorithm>
#include <atomic>

#include "svgr_kernels.h"

#ifndef COV_THREADS
#define COV_THREADS 128
#endif

// ---------------------------------------------------------------------------------------------
// generic exclusive scan of int32 (count -> offset), three phases
// ---------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_BLOCK (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ int block_exclusive_scan(int v, int *total, int *sh)
{
    // sh: 32 ints
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += t;
    }
    if (lane == 31)
        sh[w] = incl;
    __syncthreads();
    if (w == 0) {
        int s = lane < (blockDim.x >> 5) ? sh[lane] : 0;
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o)
                si += t;
        }
        sh[lane] = si - s;
        if (lane == 31)
            sh[32] = si;
    }
    __syncthreads();
    int res = incl - v + sh[w];
    *total = sh[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_local_kernel(const int *__restrict__ in, int *__restrict__ out,
                                                                  int *__restrict__ block_sums, long long n)
{
    __shared__ int sh[33];
    long long base = (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    int total;
    int ex = block_exclusive_scan(sum, &total, sh);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n)
            out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0)
        block_sums[blockIdx.x] = total;
}

// single block: scans up to any number of block sums sequentially in chunks
__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_kernel(int *__restrict__ sums, int nb, int *__restrict__ total_out)
{
    __shared__ int sh[33];
    int carry = 0;
    for (int base = 0; base < nb; base += SCAN_THREADS) {
        int i = base + threadIdx.x;
        int v = i < nb ? sums[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, &total, sh);
        if (i < nb)
            sums[i] = ex + carry;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out)
        *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(int *__restrict__ out, const int *__restrict__ block_sums,
                                                                long long n)
{
    long long base = (long long)blockIdx.x * SCAN_BLOCK + (long long)threadIdx.x * SCAN_ITEMS;
    int add = block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n)
            out[base + k] += add;
}

// out[i] = sum(in[0..i)); out has n entries; *total_out = sum of all (may be null). tmp: ceil(n / SCAN_BLOCK) ints.
void svgr_launch_exclusive_scan(const int *in, int *out, long long n, int *tmp, int *total_out, cudaStream_t s)
{
    if (n <= 0) {
        if (total_out)
            cudaMemsetAsync(total_out, 0, sizeof(int), s);
        return;
    }
    int nb = (int)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    scan_local_kernel<<<nb, SCAN_THREADS, 0, s>>>(in, out, tmp, n);
    scan_sums_kernel<<<1, SCAN_THREADS, 0, s>>>(tmp, nb, total_out);
    if (nb > 1)
        scan_add_kernel<<<nb, SCAN_THREADS, 0, s>>>(out, tmp, n);
}

// ---------------------------------------------------------------------------------------------
// binning
// ---------------------------------------------------------------------------------------------
struct EdgeRows {
    int y0, y1;  // mask-local row range [y0, y1) the edge writes to (empty if y0 >= y1)
};

// Rows of the mask that line_signed_coverage visits for this edge (svgrasterize.py:2232-2243),
// plus a conservative cull of edges that lie entirely right of the mask (every write would hit
// the `index >= w: continue` rule).
__device__ __forceinline__ EdgeRows edge_rows(double er0, double ec0, double er1, double ec1, const MaskRec &m)
{
    EdgeRows out = {0, 0};
    if (m.rows <= 0 || m.cols <= 0)
        return out;
    double r0 = er0 - (double)m.r0, r1 = er1 - (double)m.r0;
    if (r0 == r1 || !(r0 == r0) || !(r1 == r1))
        return out;
    if (r0 > r1) {
        double t = r0;
        r0 = r1, r1 = t;
    }
    double c0 = ec0 - (double)m.c0, c1 = ec1 - (double)m.c0;
    double cmin = c0 < c1 ? c0 : c1;
    if (cmin >= (double)m.cols + 1.0)
        return out;
    double ys = r0 > 0.0 ? r0 : 0.0;
    if (ys >= (double)m.rows)
        return out;
    double ye = ceil(r1);
    out.y0 = (int)ys;
    out.y1 = ye < (double)m.rows ? (int)ye : m.rows;
    if (out.y1 < 0)
        out.y1 = 0;
    return out;
}

__global__ void bin_count_kernel(const double *__restrict__ edges, const uint32_t *__restrict__ edge_path,
                                 unsigned long long n_edges, const MaskRec *__restrict__ masks,
                                 int *__restrict__ band_count)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_edges;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const double2 *e = reinterpret_cast<const double2 *>(edges + 4 * i);
        double2 a = e[0], b = e[1];
        const MaskRec &m = masks[edge_path[i]];
        EdgeRows er = edge_rows(a.x, a.y, b.x, b.y, m);
        if (er.y0 >= er.y1)
            continue;
        int b0 = er.y0 / SVGR_BAND_ROWS, b1 = (er.y1 - 1) / SVGR_BAND_ROWS;
        for (int k = b0; k <= b1; k++)
            atomicAdd(band_count + m.band_base + k, 1);
    }
}

__global__ void bin_fill_kernel(const double *__restrict__ edges, const uint32_t *__restrict__ edge_path,
                                unsigned long long n_edges, const MaskRec *__restrict__ masks,
                                const int *__restrict__ band_off, int *__restrict__ band_cursor,
                                uint32_t *__restrict__ bin_edges, double2 *__restrict__ bin_data, long long cap)
{
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_edges;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const double2 *e = reinterpret_cast<const double2 *>(edges + 4 * i);
        double2 a = e[0], b = e[1];
        const MaskRec &m = masks[edge_path[i]];
        EdgeRows er = edge_rows(a.x, a.y, b.x, b.y, m);
        if (er.y0 >= er.y1)
            continue;
        int b0 = er.y0 / SVGR_BAND_ROWS, b1 = (er.y1 - 1) / SVGR_BAND_ROWS;
        for (int k = b0; k <= b1; k++) {
            int band = m.band_base + k;
            long long pos = (long long)band_off[band] + atomicAdd(band_cursor + band, 1);
            if (pos < cap) {  // an undersized list is detected and repaired by the host after the fact
                bin_edges[pos] = (uint32_t)i;
                // the edge itself travels with its index: the coverage CTA reads its band's edges as one
                // contiguous run (32 B each) instead of gathering them through the index list
                bin_data[2 * pos] = a, bin_data[2 * pos + 1] = b;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// coverage
// ---------------------------------------------------------------------------------------------
// Where the signed-area deltas go.  The tile kernel accumulates in shared memory in 32-bit fixed point: integer adds
// commute exactly, so a mask no longer depends on the order in which the threads arrive (float atomics made masks
// differ in their last bit from run to run), and an integer shared-memory atomic is one instruction (ATOMS.ADD)
// where the float one is a compare-and-swap loop (ATOMS.CAST.SPIN).  Scale 2^22: a delta is rounded to 1.2e-7 (a
// float32 near 1.0: 6e-8), the row's prefix sum is then exact; sums wrap modulo 2^32 and are right as long as the
// final winding-weighted coverage stays below 512 in magnitude.  line_coverage_kernel (the eager entry point
// onto a caller's float trace) keeps float atomics.
#define COV_FIX_SCALE 4194304.0
#define COV_FIX_INV (1.0f / 4194304.0f)
struct FixedTrace {
    typedef int cell;
    static __device__ __forceinline__ void add(int *p, double v) { atomicAdd(p, __double2int_rn(v * COV_FIX_SCALE)); }
    static __device__ __forceinline__ void add_run(int *p, long long lo, long long hi, double v)
    {
        const int q = __double2int_rn(v * COV_FIX_SCALE);
        for (long long xi = lo; xi < hi; xi++)
            atomicAdd(p + xi, q);
    }
};
struct FloatTrace {
    typedef float cell;
    static __device__ __forceinline__ void add(float *p, double v) { atomicAdd(p, (float)v); }
    static __device__ __forceinline__ void add_run(float *p, long long lo, long long hi, double v)
    {
        const float q = (float)v;
        for (long long xi = lo; xi < hi; xi++)
            atomicAdd(p + xi, q);
    }
};

template <class A>
__device__ __forceinline__ void put(typename A::cell *row, int w, long long ix, double val)
{
    // index >= w dropped, index < 0 accumulated into column 0 (svgrasterize.py:2262 etc.)
    if (ix >= w)
        return;
    A::add(row + (ix > 0 ? (int)ix : 0), val);
}

// One row crossing of one edge (svgrasterize.py:2244-2303), columns local to the tile.
template <class A>
__device__ __forceinline__ void edge_row(typename A::cell *row, int w, double x, double x_next, double d)
{
    double x0 = x, x1 = x_next;
    if (!(x < x_next))
        x0 = x_next, x1 = x;
    double x0_floor = floor(x0), x1_ceil = ceil(x1);
    if (x1_ceil <= 0.0) {  // whole span left of the tile: the deltas sum to d, all into column 0
        A::add(row, d);
        return;
    }
    if (x0_floor >= (double)w)
        return;
    long long x0i = (long long)x0_floor, x1i = (long long)x1_ceil;
    if (x1i <= x0i + 1) {
        double xmf = 0.5 * (x + x_next) - x0_floor;
        put<A>(row, w, x0i, d * (1.0 - xmf));
        put<A>(row, w, x0i + 1, d * xmf);
        return;
    }
    double s = 1.0 / (x1 - x0);
    double x0f = x0 - x0_floor;
    double x1f = x1 - x1_ceil + 1.0;
    double a0 = 0.5 * s * ((1.0 - x0f) * (1.0 - x0f));
    double am = 0.5 * s * (x1f * x1f);
    put<A>(row, w, x0i, d * a0);
    if (x1i == x0i + 2) {
        put<A>(row, w, x0i + 1, d * (1.0 - a0 - am));
    } else {
        double a1 = s * (1.5 - x0f);
        put<A>(row, w, x0i + 1, d * (a1 - a0));
        double ds = d * s;
        long long lo = x0i + 2, hi = x1i - 1;  // interior columns [lo, hi)
        if (lo < 0) {
            long long neg_end = hi < 0 ? hi : 0;
            if (neg_end > lo)
                A::add(row, (double)(neg_end - lo) * ds);
            lo = neg_end > lo ? neg_end : lo;
        }
        if (hi > w)
            hi = w;
        A::add_run(row, lo, hi, ds);
        double a2 = a1 + (double)(x1i - x0i - 3) * s;
        put<A>(row, w, x1i - 1, d * (1.0 - a2 - am));
    }
    put<A>(row, w, x1i, d * am);
}

__device__ __forceinline__ float fill_rule_apply(float m, int rule)
{
    float v;
    if (rule == 0) {
        v = fminf(fabsf(m), 1.0f);
    } else {
        float r = fmodf(m + 1.0f, 2.0f);  // floored modulo of np.remainder
        if (r < 0.0f)
            r += 2.0f;
        v = fabsf(r - 1.0f);
    }
    return v < 1e-6f ? 0.0f : v;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Persistent: the grid is one wave of CTAs, each walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  A tile is a
// short job (a few microseconds) that starts with two dependent global loads (its record, then its edges);
// walking lets the next tile's record (cp.async into shared memory) and the first 128 of its edges (registers)
// arrive while the current tile is being scanned and stored.
__global__ void __launch_bounds__(COV_THREADS)
coverage_kernel(const TileRec *__restrict__ tiles, int n_tiles, const double2 *__restrict__ bin_data,
                float *__restrict__ cov)
{
    __shared__ __align__(16) int trace[SVGR_BAND_ROWS][SVGR_TILE_COLS];  // fixed point, see FixedTrace
    __shared__ double s_r0[COV_THREADS], s_r1[COV_THREADS], s_c0[COV_THREADS], s_dxdy[COV_THREADS];
    __shared__ float s_dir[COV_THREADS];
    __shared__ uint16_t s_pairs[COV_THREADS * SVGR_BAND_ROWS];  // (edge << 4) | row of the band, compacted
    __shared__ int s_npairs[2];
    __shared__ __align__(16) TileRec s_tile[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the trace tile starts zeroed and every tile leaves it zeroed again (the scan clears what it reads)
    for (int i = tid; i < SVGR_BAND_ROWS * SVGR_TILE_COLS / 4; i += COV_THREADS)
        reinterpret_cast<int4 *>(&trace[0][0])[i] = make_int4(0, 0, 0, 0);
    int tile = blockIdx.x, cur = 0, chunk_no = 0;
    if (tid < 2)
        s_npairs[tid] = 0;
    if (tile < n_tiles && tid < 3)
        reinterpret_cast<uint4 *>(&s_tile[0])[tid] = __ldg(reinterpret_cast<const uint4 *>(tiles + tile) + tid);
    __syncthreads();
    double2 pa = make_double2(0.0, 0.0), pb = pa;  // this thread's edge of the first chunk of the tile
    if (tile < n_tiles && tid < s_tile[0].e_cnt) {
        const double2 *e = bin_data + 2ull * (unsigned)(s_tile[0].e_off + tid);
        pa = __ldg(e), pb = __ldg(e + 1);
    }

    for (; tile < n_tiles; tile += gridDim.x, cur ^= 1) {
        // ---- the tile record (expand_masks_kernel): mask geometry, tile position, the band's slice of the bins
        const TileRec m = s_tile[cur];
        const int next = tile + gridDim.x;
        if (next < n_tiles && tid < 3)
            cp_async16(reinterpret_cast<uint4 *>(&s_tile[cur ^ 1]) + tid, reinterpret_cast<const uint4 *>(tiles + next) + tid);
        const int yb = m.yb, col0 = m.col0;
        const int nrows = min(SVGR_BAND_ROWS, m.rows - yb);
        const int w = min(SVGR_TILE_COLS, m.cols - col0);     // columns that exist in the mask
        const int wpad = min(SVGR_TILE_COLS, m.stride - col0);  // columns that exist in memory (multiple of 4)

        // ---- 1. accumulate the signed areas of this band's edges.  Per chunk of COV_THREADS edges: (a) one
        // thread per edge orients it, stores slope / start in shared memory and counts the rows of the band it
        // crosses, (b) a shared counter hands out the slots of a compact (edge, row) pair list, (c) the pairs are
        // spread over all threads.  Flattened edges are short (2-3 rows): without the compaction a warp would
        // run the row arithmetic with 4 of its 32 lanes active.
        const int e_off = m.e_off, e_cnt = m.e_cnt;
        for (int chunk0 = 0; chunk0 < e_cnt; chunk0 += COV_THREADS) {
            const int n_chunk = min(COV_THREADS, e_cnt - chunk0);
            if (chunk0 > 0)
                __syncthreads();  // previous chunk consumed
            int row_a = 0, row_n = 0;
            if (tid < n_chunk) {
                if (chunk0 > 0) {
                    const double2 *e = bin_data + 2ull * (unsigned)(e_off + chunk0 + tid);
                    pa = __ldg(e), pb = __ldg(e + 1);
                }
                double r0 = pa.x - (double)m.r0, c0 = pa.y - (double)m.c0;
                double r1 = pb.x - (double)m.r0, c1 = pb.y - (double)m.c0;
                float dir = 1.0f;
                if (!(r0 < r1)) {
                    double t;
                    dir = -1.0f;
                    t = r0, r0 = r1, r1 = t;
                    t = c0, c0 = c1, c1 = t;
                }
                int ya = 0, yz = 0;
                double dxdy = 0.0;
                if (r0 != r1) {
                    dxdy = (c1 - c0) / (r1 - r0);
                    c0 -= (double)col0;
                    double ys = r0 > 0.0 ? r0 : 0.0;
                    int y_first = (int)ys;
                    double yend_f = ceil(r1);
                    int y_last = yend_f < (double)m.rows ? (int)yend_f : m.rows;
                    ya = max(y_first, yb), yz = min(y_last, yb + nrows);
                }
                s_r0[tid] = r0, s_r1[tid] = r1, s_c0[tid] = c0, s_dxdy[tid] = dxdy;
                s_dir[tid] = dir;
                row_a = ya - yb, row_n = max(0, yz - ya);
            }
            // slots of the pair list are handed out by a shared counter (the order of the pairs does not matter:
            // they are accumulated with atomics anyway); the counters of consecutive chunks alternate so that the
            // reset of one never races with the readers of the other
            int *counter = &s_npairs[chunk_no & 1];
            if (row_n > 0) {
                const int pair0 = atomicAdd(counter, row_n);
                for (int k = 0; k < row_n; k++)
                    s_pairs[pair0 + k] = (uint16_t)((tid << 4) | (row_a + k));
            }
            if (tid == 0)
                s_npairs[(chunk_no + 1) & 1] = 0;
            __syncthreads();
            const int n_pairs = *counter;
            chunk_no++;
            for (int p = tid; p < n_pairs; p += COV_THREADS) {
                const int pr = s_pairs[p];
                const int ei = pr >> 4, yl = pr & (SVGR_BAND_ROWS - 1);
                const double r0 = s_r0[ei], r1 = s_r1[ei], dxdy = s_dxdy[ei];
                const int y = yb + yl;
                double ytop = (double)(y + 1) < r1 ? (double)(y + 1) : r1;
                double ybot = (double)y > r0 ? (double)y : r0;
                double dy = ytop - ybot;
                double x = s_c0[ei] + dxdy * (ybot - r0);
                double x_next = x + dxdy * dy;
                edge_row<FixedTrace>(trace[yl], w, x, x_next, (double)s_dir[ei] * dy);
            }
        }
        cp_async_wait_all();
        __syncthreads();  // trace complete, the next tile's record has landed

        // ---- the first chunk of the next tile's edges starts its trip now and lands during the scan
        if (next < n_tiles && tid < s_tile[cur ^ 1].e_cnt) {
            const double2 *e = bin_data + 2ull * (unsigned)(s_tile[cur ^ 1].e_off + tid);
            pa = __ldg(e), pb = __ldg(e + 1);
        }

        // ---- 2. prefix sum along columns, fill rule, 128-bit stores; the trace is cleared as it is read
        for (int y = warp; y < nrows; y += COV_THREADS / 32) {
            int carry = 0;
            float *dst = cov + m.off + (long long)(yb + y) * m.stride + col0;
            for (int cb = 0; cb < wpad; cb += 128) {
                int c = cb + 4 * lane;
                int4 v = make_int4(0, 0, 0, 0);
                if (c < wpad) {
                    v = *reinterpret_cast<const int4 *>(&trace[y][c]);
                    *reinterpret_cast<int4 *>(&trace[y][c]) = make_int4(0, 0, 0, 0);
                }
                v.y += v.x;
                v.z += v.y;
                v.w += v.z;
                int incl = v.w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o)
                        incl += t;
                }
                int ex = incl - v.w + carry;
                carry += __shfl_sync(0xffffffffu, incl, 31);
                if (c < wpad) {
                    float4 o4;  // exact integer prefix sums -> coverage
                    o4.x = fill_rule_apply((float)(v.x + ex) * COV_FIX_INV, m.fill_rule);
                    o4.y = fill_rule_apply((float)(v.y + ex) * COV_FIX_INV, m.fill_rule);
                    o4.z = fill_rule_apply((float)(v.z + ex) * COV_FIX_INV, m.fill_rule);
                    o4.w = fill_rule_apply((float)(v.w + ex) * COV_FIX_INV, m.fill_rule);
                    *reinterpret_cast<float4 *>(dst + c) = o4;
                }
            }
        }
        __syncthreads();  // trace cleared before the next tile accumulates into it
    }
}

// line_signed_coverage(canvas, line) (svgrasterize.py:2213-2304) as an entry point of its own: one thread per
// line walks the rows the line crosses and adds its signed-area deltas to a trace in global memory (the same
// per-row arithmetic as the tile kernel above, mask origin (0, 0), no column chunking).
__global__ void line_coverage_kernel(const double *__restrict__ lines, long long n, float *__restrict__ trace, int rows,
                                     int cols)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double r0 = lines[4 * i], c0 = lines[4 * i + 1], r1 = lines[4 * i + 2], c1 = lines[4 * i + 3];
        if (r0 == r1 || !(r0 == r0) || !(r1 == r1))
            continue;
        double dir = 1.0;
        if (!(r0 < r1)) {
            double t;
            dir = -1.0;
            t = r0, r0 = r1, r1 = t;
            t = c0, c0 = c1, c1 = t;
        }
        const double dxdy = (c1 - c0) / (r1 - r0);
        const double ys = r0 > 0.0 ? r0 : 0.0;
        if (ys >= (double)rows)
            continue;
        const double yend_f = ceil(r1);
        const int y_first = (int)ys, y_last = yend_f < (double)rows ? (int)yend_f : rows;
        for (int y = y_first; y < y_last; y++) {
            const double ytop = (double)(y + 1) < r1 ? (double)(y + 1) : r1;
            const double ybot = (double)y > r0 ? (double)y : r0;
            const double dy = ytop - ybot;
            const double x = c0 + dxdy * (ybot - r0);
            const double x_next = x + dxdy * dy;
            edge_row<FloatTrace>(trace + (long long)y * cols, cols, x, x_next, dir * dy);
        }
    }
}

void svgr_launch_line_coverage(const double *lines, long long n, float *trace, int rows, int cols, cudaStream_t s)
{
    if (n <= 0 || rows <= 0 || cols <= 0)
        return;
    const long long blocks = std::min<long long>((n + 127) / 128, 148 * 16);
    line_coverage_kernel<<<(unsigned)blocks, 128, 0, s>>>(lines, n, trace, rows, cols);
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
void svgr_launch_bin_count(const double *edges, const uint32_t *edge_path, unsigned long long n_edges,
                           const MaskRec *masks, int *band_count, int sm_count, cudaStream_t s)
{
    if (n_edges == 0)
        return;
    unsigned long long blocks = (n_edges + 255) / 256;
    if (blocks > (unsigned long long)sm_count * 16)
        blocks = (unsigned long long)sm_count * 16;
    bin_count_kernel<<<(unsigned)blocks, 256, 0, s>>>(edges, edge_path, n_edges, masks, band_count);
}

void svgr_launch_bin_fill(const double *edges, const uint32_t *edge_path, unsigned long long n_edges,
                          const MaskRec *masks, const int *band_off, int *band_cursor, uint32_t *bin_edges,
                          double *bin_data, long long cap, int sm_count, cudaStream_t s)
{
    if (n_edges == 0)
        return;
    unsigned long long blocks = (n_edges + 255) / 256;
    if (blocks > (unsigned long long)sm_count * 16)
        blocks = (unsigned long long)sm_count * 16;
    bin_fill_kernel<<<(unsigned)blocks, 256, 0, s>>>(edges, edge_path, n_edges, masks, band_off, band_cursor, bin_edges,
                                                     reinterpret_cast<double2 *>(bin_data), cap);
}

__global__ void expand_masks_kernel(const MaskRec *__restrict__ masks, int n_masks, int n_tiles,
                                    const int *__restrict__ band_off, const int *__restrict__ band_cnt, long long bin_cap,
                                    TileRec *__restrict__ tiles)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles)
        return;
    int lo = 0, hi = n_masks - 1;
    while (lo < hi) {  // last mask with tile_base <= t (empty masks share their successor's base and sort before it)
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(&masks[mid].tile_base) <= t)
            lo = mid;
        else
            hi = mid - 1;
    }
    const MaskRec m = masks[lo];
    const int local = t - m.tile_base;
    const int band_local = local / m.ntile_c;
    const int band = m.band_base + band_local;
    TileRec r;
    r.r0 = m.r0, r.c0 = m.c0, r.rows = m.rows, r.cols = m.cols;
    r.stride = m.stride, r.fill_rule = m.fill_rule;
    r.yb = band_local * SVGR_BAND_ROWS, r.col0 = (local - band_local * m.ntile_c) * SVGR_TILE_COLS;
    r.off = m.off;
    r.e_off = band_off[band];
    r.e_cnt = (int)max(0ll, min((long long)band_cnt[band], bin_cap - r.e_off));
    tiles[t] = r;
}

void svgr_launch_coverage(const MaskRec *masks, int n_masks, int n_tiles, TileRec *tiles, const int *band_off,
                          const int *band_cnt, const double *bin_data, long long bin_cap, float *cov, cudaStream_t s)
{
    if (n_tiles <= 0)
        return;
    expand_masks_kernel<<<(n_tiles + 255) / 256, 256, 0, s>>>(masks, n_masks, n_tiles, band_off, band_cnt, bin_cap, tiles);
    // CTAs of one full wave, per device (contexts of several devices and threads share this table: the value is
    // a pure function of the device, so a racing double initialisation writes the same number)
    static std::atomic<int> wave_of[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    int wave = wave_of[dev].load(std::memory_order_relaxed);
    if (wave == 0) {
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coverage_kernel, COV_THREADS, 0);
        wave = std::max(1, sms * std::max(1, per_sm));
        wave_of[dev].store(wave, std::memory_order_relaxed);
    }
    coverage_kernel<<<std::min(n_tiles, wave), COV_THREADS, 0, s>>>(tiles, n_tiles, reinterpret_cast<const double2 *>(bin_data),
                                                                  cov);
}
