// k_stencil_tma.cu -- the separable filter stencils fed by the Tensor Memory Accelerator.
//
// Same operations as k_filters.cu (Layer.convolve = scipy "full" convolution, svgrasterize.py:106-118, with the
// kernel of blur_kernel, :1903-1944; Layer.morphology / pooling, :120-127, :419-468), for the common case of a
// materialised RGBA float32 source layer: such a layer *is* a 2-D tensor, and a stencil tile plus its halo is a
// box of it whose out-of-range part must read as zero -- exactly what a tiled tensor map delivers.
//
//   * persistent CTAs (one wave) walk the (tile, tap group) units of a launch;
//   * one elected thread issues cp.async.bulk.tensor loads of the next unit's box(es) into the other half of a
//     two-stage shared-memory ring while all warps work on the current one (mbarrier complete_tx);
//   * every thread owns R consecutive outputs along the filter axis and slides a register window over them: a tap
//     costs one LDS.128 and R x 2 packed FMAs instead of R loads (R = 7 along rows: lanes 7 pixels apart keep the
//     128-bit shared-memory reads conflict free; R = 8 down columns);
//   * kernels longer than a stage are cut into tap groups that accumulate into the same registers, so there is no
//     upper bound on the kernel length (a 1001-tap blur is 4 groups along rows, 16 down columns);
//   * the horizontal pass hands its tile back through cp.async.bulk.tensor stores (coalescing and the clip at the
//     layer edge are the TMA unit's job), the vertical pass stores 512-byte rows directly.
//
// The pixel type is 16 bytes and a box dimension is limited to 256 elements, so the maps describe a layer as a
// matrix of uint64 pairs: dim0 = 2 x columns, dim1 = rows; a row box holds 128 pixels.
#include <cuda.h>

#include <atomic>

#include "svgr_device.cuh"

#define TH_TR SVGR_TMA_H_TR      // 8 rows per tile: one per warp
#define TH_R 7                   // outputs per thread along the row
#define TH_TC SVGR_TMA_H_TC      // 224 = 32 lanes x 7
#define TH_CHUNK 128             // pixels per input box
#define TH_OCHUNK 112            // pixels per output box (two per tile)
#define TV_TR SVGR_TMA_V_TR      // 64 rows per tile: 8 per warp
#define TV_R 8
#define TV_TC SVGR_TMA_V_TC      // 32 columns: one per lane

static_assert(TH_TC == 32 * TH_R && TH_TC == 2 * TH_OCHUNK, "horizontal tile shape");
static_assert(TV_TR == 8 * TV_R, "vertical tile shape");

namespace {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int x, int y, const void *src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(x), "r"(y)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// np.nanmax / np.nanmin of two values: IEEE maxNum / minNum (one FMNMX): the operand that is not NaN wins
__device__ __forceinline__ float nanmax_t(float a, float v) { return fmaxf(a, v); }
__device__ __forceinline__ float nanmin_t(float a, float v) { return fminf(a, v); }

template <int ST>
__device__ __forceinline__ void tap(float4 &a, float w, const float4 &v)
{
    if (ST == STENCIL_CONV) {
        a = madd4(v, w, a);
    } else if (ST == STENCIL_MAX) {
        a.x = nanmax_t(a.x, v.x), a.y = nanmax_t(a.y, v.y), a.z = nanmax_t(a.z, v.z), a.w = nanmax_t(a.w, v.w);
    } else {
        a.x = nanmin_t(a.x, v.x), a.y = nanmin_t(a.y, v.y), a.z = nanmin_t(a.z, v.z), a.w = nanmin_t(a.w, v.w);
    }
}

// What a thread needs to know about a unit = (tile, tap group) of the launch.
struct Unit {
    int tile, g, ng;   // tile of the launch, tap group, groups of the tile's op
    int opi;           // op index inside the launch
    int k, kc, u0;     // taps of the op, taps of this group, first tap of this group
    int row0, col0;    // tile origin, output-local
    int st;            // STENCIL_*
};

template <bool HORIZ>
__device__ __forceinline__ void unit_load(Unit &u, const OpRec *__restrict__ ops, const int *__restrict__ tile_op, int kc_max)
{
    u.opi = __ldg(tile_op + u.tile);
    const OpRec *op = ops + u.opi;
    u.k = __ldg(&op->k0), u.st = __ldg(&op->stencil);
    const int local = u.tile - __ldg(&op->tile_base), ntc = __ldg(&op->ntile_c);
    const int tr = local / ntc, tc = local - tr * ntc;
    u.row0 = tr * (HORIZ ? TH_TR : TV_TR), u.col0 = tc * (HORIZ ? TH_TC : TV_TC);
    u.ng = (u.k + kc_max - 1) / kc_max;
}
__device__ __forceinline__ void unit_group(Unit &u, int kc_max)
{
    u.u0 = u.g * kc_max;
    u.kc = min(kc_max, u.k - u.u0);
}

// Sliding register window: acc[r] (op)= w[u] . in[i + r] for the kc taps of the group; `load(p)` returns window
// position p of this thread (0 <= p <= kc + R - 2), `wp[-i]` the weight of tap i of the group.
template <int ST, int R, class Load>
__device__ __forceinline__ void slide(float4 (&acc)[R], int kc, const float *__restrict__ wp, Load load)
{
    float4 x[R];
#pragma unroll
    for (int r = 0; r < R - 1; r++)
        x[r] = load(r);
    for (int i = 0; i < kc; i += R) {
#pragma unroll
        for (int ii = 0; ii < R; ii++) {
            if (i + ii < kc) {
                x[(ii + R - 1) % R] = load(i + ii + R - 1);
                const float w = ST == STENCIL_CONV ? __ldg(wp - (i + ii)) : 0.f;
#pragma unroll
                for (int r = 0; r < R; r++)
                    tap<ST>(acc[r], w, x[(ii + r) % R]);
            }
        }
    }
}

template <int R>
__device__ __forceinline__ void acc_init(float4 (&acc)[R], int st)
{
    const float v = st == STENCIL_CONV ? 0.f : __int_as_float(0x7fc00000);
#pragma unroll
    for (int r = 0; r < R; r++)
        acc[r] = make_float4(v, v, v, v);
}

}  // namespace

// One kernel per direction.  Dynamic shared memory: two input stages of `stage_bytes`, then (horizontal only) the
// output staging tile.  tmaps[2 * op + 0] = source map, tmaps[2 * op + 1] = destination map (horizontal only).
template <bool HORIZ>
__global__ void __launch_bounds__(256, 2)
stencil_tma_kernel(RenderTables T, const OpRec *__restrict__ ops, const int *__restrict__ tile_op, int n_tiles,
                   const CUtensorMap *__restrict__ tmaps, float *__restrict__ layers_out, int stage_bytes, int kc_max)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full[2];
    constexpr int R = HORIZ ? TH_R : TV_R;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto stage_ptr = [&](int st) { return smem + (size_t)st * stage_bytes; };
    float4 *out_stage = reinterpret_cast<float4 *>(smem + 2 * (size_t)stage_bytes);

    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_async_smem();
    }
    __syncthreads();

    // thread 0: start the loads of unit `u` into stage `s`
    auto issue = [&](const Unit &u, int s) {
        const CUtensorMap *map = tmaps + 2 * u.opi;
        const int base = u.st == STENCIL_CONV ? -(u.k - 1) : 0;
        if (HORIZ) {
            const int n_box = (u.kc + TH_TC - 1 + TH_CHUNK - 1) / TH_CHUNK;
            mbar_expect_tx(&full[s], (unsigned)n_box * TH_TR * TH_CHUNK * 16);
            for (int c = 0; c < n_box; c++)
                tma_load_2d(stage_ptr(s) + (size_t)c * TH_TR * TH_CHUNK * 16, map, 2 * (u.col0 + base + u.u0 + c * TH_CHUNK),
                            u.row0, &full[s]);
        } else {
            const int box_rows = min(u.k, kc_max) + TV_TR - 1;  // the map's box: fixed per op
            mbar_expect_tx(&full[s], (unsigned)box_rows * TV_TC * 16);
            tma_load_2d(stage_ptr(s), map, 2 * u.col0, u.row0 + base + u.u0, &full[s]);
        }
    };

    Unit cur;
    cur.tile = blockIdx.x, cur.g = 0;
    if (cur.tile >= n_tiles)
        return;
    unit_load<HORIZ>(cur, ops, tile_op, kc_max);
    unit_group(cur, kc_max);
    if (tid == 0)
        issue(cur, 0);

    float4 acc[R];
    bool stored = false;  // horizontal: a bulk store of this CTA may still be reading the staging tile
    for (unsigned n = 0;; n++) {
        const int s = n & 1;
        // ---- the unit after this one: its box(es) start travelling now, into the stage freed one iteration ago
        Unit nxt = cur;
        bool has_next = true;
        if (cur.g + 1 < cur.ng) {
            nxt.g = cur.g + 1;
        } else {
            nxt.tile = cur.tile + gridDim.x, nxt.g = 0;
            has_next = nxt.tile < n_tiles;
            if (has_next)
                unit_load<HORIZ>(nxt, ops, tile_op, kc_max);
        }
        if (has_next) {
            unit_group(nxt, kc_max);
            if (tid == 0)
                issue(nxt, s ^ 1);
        }
        // ---- this unit
        const OpRec *op = ops + cur.opi;
        const SrcRec *src = T.srcs + __ldg(&op->src_off);
        const int conv = __ldg(&src->conv);
        const float mul = __ldg(&src->mul);
        const bool fix = mul != 1.0f || !conv_is_identity(conv);
        if (cur.g == 0)
            acc_init<R>(acc, cur.st);
        mbar_wait(&full[s], (n >> 1) & 1);
        const float *wp = T.weights + __ldg(&op->aux) + (cur.k - 1 - cur.u0);  // conv: tap u has weight w[k - 1 - u]
        const unsigned char *sb = stage_ptr(s);
        auto fetch = [&](const float4 *p) {
            float4 v = *p;
            if (fix) {
                if (mul != 1.0f)
                    v = scale4(v, mul);
                v = convert_px(v, conv);  // zero stays zero: the out-of-range fill needs no special case
            }
            return v;
        };
        if (HORIZ) {
            const int p0 = lane * TH_R;
            auto load = [&](int p) {
                const int q = p0 + p;
                return fetch(reinterpret_cast<const float4 *>(sb + ((size_t)((q >> 7) * TH_TR + warp) * TH_CHUNK + (q & 127)) * 16));
            };
            if (cur.st == STENCIL_CONV)
                slide<STENCIL_CONV, R>(acc, cur.kc, wp, load);
            else if (cur.st == STENCIL_MAX)
                slide<STENCIL_MAX, R>(acc, cur.kc, wp, load);
            else
                slide<STENCIL_MIN, R>(acc, cur.kc, wp, load);
        } else {
            const int r0 = warp * TV_R;
            auto load = [&](int p) { return fetch(reinterpret_cast<const float4 *>(sb + ((size_t)(r0 + p) * TV_TC + lane) * 16)); };
            if (cur.st == STENCIL_CONV)
                slide<STENCIL_CONV, R>(acc, cur.kc, wp, load);
            else if (cur.st == STENCIL_MAX)
                slide<STENCIL_MAX, R>(acc, cur.kc, wp, load);
            else
                slide<STENCIL_MIN, R>(acc, cur.kc, wp, load);
        }
        // ---- last group of the tile: the outputs leave
        if (cur.g + 1 == cur.ng) {
            if (HORIZ) {
                if (stored) {
                    if (tid == 0)
                        tma_store_wait_read();  // the previous tile's stores have read the staging tile
                    __syncthreads();
                }
                float4 *o = out_stage + ((size_t)((lane >> 4) * TH_TR + warp) * TH_OCHUNK + (lane & 15) * TH_R);
#pragma unroll
                for (int r = 0; r < R; r++)
                    o[r] = acc[r];
                fence_async_smem();
                __syncthreads();
                if (tid == 0) {
                    const CUtensorMap *omap = tmaps + 2 * cur.opi + 1;
                    tma_store_2d(omap, 2 * cur.col0, cur.row0, out_stage);
                    tma_store_2d(omap, 2 * (cur.col0 + TH_OCHUNK), cur.row0, out_stage + (size_t)TH_TR * TH_OCHUNK);
                    tma_store_commit();
                }
                stored = true;
            } else {
                const int rows = __ldg(&op->rows), cols = __ldg(&op->cols), stride = __ldg(&op->stride);
                const int lc = cur.col0 + lane;
                if (lc < cols) {
                    float4 *dst = reinterpret_cast<float4 *>(layers_out + __ldg(&op->out_off));
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int lr = cur.row0 + warp * TV_R + r;
                        if (lr < rows)
                            dst[(long long)lr * stride + lc] = acc[r];
                    }
                }
            }
        }
        if (!has_next)
            break;
        __syncthreads();  // everybody is done with stage s: the next iteration may refill it
        cur = nxt;
    }
    if (HORIZ && tid == 0)
        tma_store_wait_all();
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*svgr_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                         const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static svgr_encode_tiled_fn encode_fn()
{
    static std::atomic<void *> cached{nullptr};
    void *p = cached.load(std::memory_order_acquire);
    if (!p) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cached.store(p, std::memory_order_release);
    }
    return (svgr_encode_tiled_fn)p;
}

// Tensor map of a rows x cols RGBA float32 region whose rows are `stride_px` pixels apart, cut into boxes of
// box_px x box_rows pixels.  `out` must be 64-byte aligned host memory of 128 bytes.  Returns 0 on success.
int svgr_encode_layer_map(void *out, const float *base, long long rows, long long cols, long long stride_px, int box_px,
                          int box_rows)
{
    svgr_encode_tiled_fn fn = encode_fn();
    if (!fn || rows <= 0 || cols <= 0 || box_px < 1 || box_px > 128 || box_rows < 1 || box_rows > 256)
        return -1;
    const cuuint64_t dims[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)stride_px * 16};
    const cuuint32_t box[2] = {(cuuint32_t)box_px * 2, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn((CUtensorMap *)out, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void *)base, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -2;
}

size_t svgr_stencil_tma_smem(bool horiz, int k_max, int *kc_out)
{
    // taps per group: along rows a stage holds at most 4 boxes of 128 pixels (kc + 223 <= 512); down columns a box
    // holds at most 128 rows (kc + 63 <= 128)
    const int kc = std::min(k_max, horiz ? SVGR_TMA_H_KC : SVGR_TMA_V_KC);
    *kc_out = kc;
    if (horiz) {
        const size_t n_box = (size_t)(kc + TH_TC - 1 + TH_CHUNK - 1) / TH_CHUNK;
        return 2 * n_box * TH_TR * TH_CHUNK * 16 + (size_t)TH_TR * TH_TC * 16;
    }
    return 2 * (size_t)(kc + TV_TR - 1) * TV_TC * 16;
}

int svgr_launch_stencil_tma(const RenderTables &T, const OpRec *ops, const int *tile_op, int n_tiles, const void *tmaps,
                            bool horiz, int k_max, int sm_count, float *layers_out, cudaStream_t s)
{
    if (n_tiles <= 0)
        return 0;
    int kc = 0;
    const size_t smem = svgr_stencil_tma_smem(horiz, k_max, &kc);
    if (smem > SVGR_MAX_DYN_SMEM)
        return -1;
    static std::atomic<bool> attr_set[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!attr_set[dev].load(std::memory_order_acquire)) {
        cudaFuncSetAttribute(stencil_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SVGR_MAX_DYN_SMEM);
        cudaFuncSetAttribute(stencil_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SVGR_MAX_DYN_SMEM);
        attr_set[dev].store(true, std::memory_order_release);
    }
    const int stage = (int)((smem - (horiz ? (size_t)TH_TR * TH_TC * 16 : 0)) / 2);
    const int per_sm = smem > 110 * 1024 ? 1 : 2;
    const int grid = std::min(n_tiles, sm_count * per_sm);
    if (horiz)
        stencil_tma_kernel<true><<<grid, 256, smem, s>>>(T, ops, tile_op, n_tiles, (const CUtensorMap *)tmaps, layers_out, stage, kc);
    else
        stencil_tma_kernel<false><<<grid, 256, smem, s>>>(T, ops, tile_op, n_tiles, (const CUtensorMap *)tmaps, layers_out, stage, kc);
    return 0;
}
