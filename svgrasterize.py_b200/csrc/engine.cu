// engine.cu -- host side of libsvgr_b200.so: context, device buffers, the render
// pipeline and the planner that turns a scene program + device-computed path
// boxes into per-level op tables (the work Scene.render does node by node in the
// reference, svgrasterize.py:649-752, done once per batch here).
//
// Pipeline of one svgr_render call (one stream, one host synchronisation):
//   H2D program -> stroke (count/scan/emit/assemble) -> flatten + bounds
//   -> D2H path boxes  [sync]  -> plan (host) -> H2D plan tables
//   -> bin count / scan / fill -> coverage -> focal flags -> compose levels
//   -> canvas quantise -> (D2H RGBA8)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/svgr_b200.h"
#include "svgr_kernels.h"

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                         \
            return SVGR_E_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define FAIL(code, msg)    \
    do {                   \
        ctx->err = (msg);  \
        return (code);     \
    } while (0)

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes, bool keep = false)
    {
        if (bytes <= cap)
            return cudaSuccess;
        size_t ncap = bytes + bytes / 4 + 256;
        void *np_ = nullptr;
        cudaError_t e = cudaMalloc(&np_, ncap);
        if (e != cudaSuccess)
            return e;
        if (keep && p && cap)
            cudaMemcpy(np_, p, cap, cudaMemcpyDeviceToDevice);
        if (p)
            cudaFree(p);
        p = np_, cap = ncap;
        return cudaSuccess;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr, cap = 0;
    }
    template <class T>
    T *as() const
    {
        return (T *)p;
    }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap)
            return cudaSuccess;
        size_t ncap = bytes + bytes / 4 + 256;
        if (p)
            cudaFreeHost(p);
        p = nullptr, cap = 0;
        cudaError_t e = cudaMallocHost(&p, ncap);
        if (e == cudaSuccess)
            cap = ncap;
        return e;
    }
    void release()
    {
        if (p)
            cudaFreeHost(p);
        p = nullptr, cap = 0;
    }
};

// Value of a node: VAL_EMPTY, one of SRC_L4 / SRC_L1 / SRC_COV / SRC_COVPAINT, or VAL_LUMA (planner only:
// the luminance mask of a materialised RGBA layer that has not been written out as a layer of its own).
enum { VAL_EMPTY = 0, VAL_LUMA = 100 };

struct Val {
    int kind = VAL_EMPTY;
    int r0 = 0, c0 = 0, rows = 0, cols = 0;  // valid region
    int br0 = 0, bc0 = 0;                    // global coordinates of storage element 0
    int stride = 0;
    int paint = -1;
    float mul = 1.0f;
    int pre = 1, lin = 1;
    long long off = 0;
    long long off2 = 0;
    int stride2 = 0;
    int level = 0;
    bool missed = false;  // MERGE_AT whose overlay does not touch the box (canvas_merge_at returns None, :315-322)
    int luma_conv = 0;  // VAL_LUMA: Layer.convert code applied before the luminance
    // stencil: the value is multiplied by a one-channel view (a fused compose IN)
    int st_kind = 0;  // 0 none, else SRC_MOD_*
    int st_br0 = 0, st_bc0 = 0, st_stride = 0, st_conv = 0;
    long long st_off = 0;
    float st_mul = 1.0f;
    // producer (for the canvas fusion)
    int op_index = -1, owner = -1;
    bool one_channel() const { return kind == SRC_L1 || kind == SRC_COV || kind == VAL_LUMA; }
    bool is_virtual() const { return kind == SRC_COVPAINT || kind == VAL_LUMA || mul != 1.0f || st_kind != 0; }
};

struct PlannedOp {
    OpRec op;
    int level;
    int cls;  // 0 compose, 1 stencil, 2 conv2d, 3 canvas, 4 / 5 TMA-fed stencil along rows / down columns
};

struct Launch {
    int cls;
    int op_begin, op_count, n_tiles;
    size_t smem;
    long long list_slots;  // compose launches: TileEntry slots (sum over ops of tiles x sources)
    long long head_off = 0, list_off = 0;  // where this launch's tile heads / entries start in the chunk's buffers
    bool simple = true;  // compose launches: only OVER folds without conversions, patterns or special outputs
    int k_max = 0;       // TMA stencil launches: longest kernel of the launch
    long long tmap_off = 0;  // TMA stencil launches: first tensor map of the launch in the chunk's map table
};

struct StatusBlock {  // device -> host after flatten
    unsigned long long n_edges;
    int stroke_err;
    int stroke_curves;   // total offset curves (pool use)
    int outline_total;   // outline segments incl. padding
    int outline_count;   // segments visible to flatten
    int binned_total;
    int paint_err;       // resolve_paints_kernel: bit 0 = singular objectBoundingBox matrix
};

}  // namespace

struct svgr_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // device -> host copies of finished canvases, overlapped with later chunks
    cudaEvent_t ev_copy[16] = {nullptr};
    cudaStream_t up_stream = nullptr;  // plan tables of the next chunk go up (and its focal flags are computed)
    cudaEvent_t ev_up[16] = {nullptr}; // while the current chunk composes
    cudaEvent_t ev_up_begin = nullptr;
    cudaEvent_t ev_geo[2] = {nullptr, nullptr};  // fill-segment flatten on up_stream beside the stroke kernels
    cudaEvent_t ev_done[16] = {nullptr};  // a chunk's launches have finished (its list buffers may be rewritten)
    std::string err;

    // ---- resident program (device) + host copies of what the planner needs
    bool have_program = false;
    const svgr_program *pending = nullptr;
    bool structure_ready = false;  // node_uses / chunk_bounds computed for the loaded program
    std::vector<int> node_uses, chunk_bounds;  // host tables of this program still to be copied (finish_load)
    long long n_seg = 0;
    int n_path = 0, n_stroke = 0, n_stroke_sub = 0, n_paint = 0, n_stop = 0, n_focal = 0, n_node = 0;
    long long n_stroke_seg = 0;
    long long canvas_bytes = 0;
    double flatness = 0.0;
    DevBuf d_eager[3];  // scratch of the eager element-wise entry points
    // PNG encoding (k_png.cu): segment / canvas tables, worst-case slots, sizes, packed files
    DevBuf d_png_segs, d_png_canvases, d_png_scratch, d_png_seg_bytes, d_png_seg_adler, d_png_file_bytes, d_png_file_off,
        d_png_out, d_png_in, d_png_code;
    PinBuf pin_png;
    cudaEvent_t ev_png[2] = {nullptr, nullptr};
    DevBuf d_tmaps;     // tensor maps of the TMA-fed stencil launches of the render in flight (128 B each)
    PinBuf pin_tmaps;
    int n_filter_nodes = 0;  // blur / morphology nodes of the loaded program (bounds the map table)
    DevBuf d_seg_tag, d_seg_data, d_seg_path, d_paths, d_strokes, d_ssub_off, d_ssub_job, d_stag, d_sdata, d_sseg_job;
    DevBuf d_sitems;            // stroke assembly work items: (sub-path, first segment, end segment), <= 256 segments each
    std::vector<int> h_sitems;
    int n_sitems = 0;
    DevBuf d_paints, d_stops, d_matrices, d_weights;
    DevBuf d_bbox_jobs, d_bbox_csr, d_bbox_inv, d_bbox_keys;  // objectBoundingBox gradients completed after flattening
    int n_bbox_job = 0;
    std::vector<int> h_bbox_csr;
    std::vector<double> h_bbox_inv;
    std::vector<PathRec> h_paths;
    std::vector<PaintRec> h_paints;
    std::vector<svgr_node> h_nodes;
    std::vector<int32_t> h_children;
    std::vector<svgr_kernel> h_kernels;
    std::vector<float> h_matrices;
    std::vector<double> h_offset_tr;
    std::vector<svgr_external> h_ext;
    std::vector<std::vector<float>> h_ext_data;
    int n_weight = 0, n_matrix = 0;

    // ---- stroke state
    DevBuf d_scounts, d_soffs, d_scan_tmp, d_pool, d_sbound, d_sout_off, d_sout_total;
    DevBuf d_otag, d_odata, d_opath, d_osub, d_ocount;
    long long pool_cap = 0, outline_cap = 0;

    // ---- flatten state
    DevBuf d_edges, d_edge_path, d_minmax, d_boxes, d_minmax_f64, d_status;
    long long edge_cap = 0;
    long long n_edges = 0;
    std::vector<PathBox> h_boxes;
    std::vector<PathBox> h_full_boxes;  // band renders: the boxes the whole-canvas render would give (else empty)
    bool band_mode = false;             // some path carries a full viewport
    DevBuf d_full_boxes;
    PinBuf pin_full_boxes;
    PinBuf pin_boxes, pin_status, pin_plan, pin_out, pin_masks;

    // ---- plan state
    MaskRec *h_masks = nullptr;  // n_path records, built by plan_masks directly in pinned memory (pin_masks) ...
    std::vector<MaskRec> h_masks_store;  // ... or here, where no copy to the device follows (plan-only calls)
    std::vector<Val> vals;
    std::vector<PlannedOp> ops;
    std::vector<SrcRec> srcs;
    std::vector<FocalJob> focal_jobs;
    std::vector<Launch> launches;
    int n_focal_blocks = 0;
    long long n_bands = 0, n_cov_tiles = 0, cov_floats = 0, layer_floats = 0;
    long long mask_pixels = 0, layer_pixels = 0, compose_bytes = 0, compose_bytes_8d = 0, canvas_pixels = 0;
    int n_levels = 0;
    DevBuf d_masks, d_band_cnt, d_band_off, d_band_cur, d_bin_edges, d_cov, d_layers, d_ops, d_srcs, d_focal_jobs,
        d_focal_flags, d_canvas, d_q, d_tile_map, d_tile_rec, d_bin_data, d_heads, d_lists, d_ovf_cubic[2],
        d_ovf_path[2], d_ovf_depth[2], d_ovf_counts;
    long long bin_cap = 0;
    long long n_binned = 0;
    bool planned = false, covered = false, composed = false;
    // Plan cache: the plan is a function of (program, path boxes, arena addresses).  A re-render of the resident
    // program that finds the same boxes keeps every table the last render left on the device (masks, ops, sources,
    // per-tile source lists, focal flags, tensor maps) and only repeats the launches: no host planning, no uploads,
    // no culls.
    bool plan_cached = false;
    std::vector<PathBox> cache_boxes, cache_full_boxes;
    const void *cache_layers = nullptr;
    std::vector<int> chunk_launch_begin, chunk_op_begin;  // per planned chunk (+ the end), into launches / ops
    int plan_cache_hits = 0;

    cudaEvent_t ev[16] = {nullptr};
    cudaEvent_t ev_chunk[16][2] = {{nullptr}};
};

// ---------------------------------------------------------------------------------------------
// planner
// ---------------------------------------------------------------------------------------------
namespace {

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long align4(long long v) { return (v + 3) & ~3ll; }

struct Planner {
    svgr_ctx *ctx;
    std::string err;
    int err_code = SVGR_E_INVALID;
    long long layer_top = 0;
    const std::vector<int> *uses_p = nullptr;  // how many nodes read each node (shared, read-only while planning)
    std::vector<PlannedOp> sorted_ops;
    std::vector<Val> scratch_vals;
    // source lists under construction: ops nest (a consumer may have to materialise one of its inputs first),
    // so the scratch vectors form a small stack instead of being allocated per op
    std::vector<std::vector<SrcRec>> ss_pool;
    int ss_depth = 0;
    struct SrcList {
        Planner &pl;
        std::vector<SrcRec> &v;
        explicit SrcList(Planner &p) : pl(p), v(p.take()) {}
        ~SrcList() { pl.ss_depth--; }
    };
    std::vector<SrcRec> &take()
    {
        if (ss_depth >= (int)ss_pool.size())
            ss_depth = (int)ss_pool.size() - 1;  // cannot happen (nesting is two deep); never grow while lists are live
        std::vector<SrcRec> &v = ss_pool[ss_depth++];
        v.clear();
        return v;
    }
    // outputs: the context's tables, or thread-local ones when the node range is planned by several threads
    std::vector<PlannedOp> *ops_p;
    std::vector<SrcRec> *srcs_p;
    std::vector<FocalJob> *focal_p;
    int n_focal_blocks = 0;
    long long layer_pixels = 0;

    explicit Planner(svgr_ctx *c) : ctx(c), ops_p(&c->ops), srcs_p(&c->srcs), focal_p(&c->focal_jobs)
    {
        ss_pool.resize(64);  // deeper nesting than this does not occur (materialise -> unary is two levels)
    }

    Val alloc(int kind, int r0, int c0, int rows, int cols, int pre, int lin, int level)
    {
        Val v;
        v.kind = kind, v.r0 = r0, v.c0 = c0, v.rows = rows, v.cols = cols, v.pre = pre, v.lin = lin, v.level = level;
        v.br0 = r0, v.bc0 = c0;
        layer_top = align4(layer_top);
        v.off = layer_top;
        if (kind == SRC_L4) {
            v.stride = cols;
            layer_top += 4ll * rows * cols;
        } else {
            v.stride = (int)align4(cols);
            layer_top += (long long)rows * v.stride;
        }
        layer_pixels += (long long)rows * cols;
        return v;
    }

    SrcRec src_of(const Val &v, int want_pre, int want_lin)
    {
        SrcRec s;
        memset(&s, 0, sizeof s);
        s.kind = v.kind;
        s.r0 = v.r0, s.c0 = v.c0, s.rows = v.rows, s.cols = v.cols;
        s.br0 = v.br0, s.bc0 = v.bc0;
        s.stride = v.stride;
        s.paint = v.paint;
        s.conv = v.one_channel() ? 0 : SVGR_CONV(v.pre, v.lin, want_pre, want_lin);
        s.mul = v.mul;
        s.off = v.off;
        s.off2 = v.off2, s.stride2 = v.stride2;
        return s;
    }

    // Appends the source entries of `v` as seen by a consumer that wants (want_pre, want_lin): one entry,
    // plus a stencil modifier entry when v carries a fused compose-IN.  A stencil cannot be combined with a
    // pending colour conversion (the reference converts the already clipped layer), and a deferred luminance
    // is only usable as a stencil: both cases are written out as layers first.  Returns the level of v.
    int push_src(std::vector<SrcRec> &ss, Val v, int want_pre, int want_lin)
    {
        if (v.kind == VAL_LUMA)
            v = materialize(v);
        if (v.st_kind && !v.one_channel() && !(v.pre == want_pre && v.lin == want_lin))
            v = materialize(v);
        ss.push_back(src_of(v, want_pre, want_lin));
        if (v.st_kind) {
            SrcRec m;
            memset(&m, 0, sizeof m);
            m.kind = v.st_kind;
            m.r0 = v.r0, m.c0 = v.c0, m.rows = v.rows, m.cols = v.cols;  // same rectangle as the owner: culled together
            m.br0 = v.st_br0, m.bc0 = v.st_bc0, m.stride = v.st_stride;
            m.off = v.st_off, m.mul = v.st_mul, m.conv = v.st_conv;
            ss.push_back(m);
        }
        return v.level;
    }

    // a view the single-source stencil kernels can read: no modifier, no deferred luminance
    Val plain(const Val &v) { return (v.st_kind || v.kind == VAL_LUMA) ? materialize(v) : v; }

    // emits an op producing `out` from srcs
    void emit(int cls, int kind, const Val &out, const std::vector<SrcRec> &ss, int mode, int post, float mul, int level,
              const float *k = nullptr, int aux = 0, int k0 = 0, int k1 = 0, int stencil = 0)
    {
        PlannedOp p;
        memset(&p, 0, sizeof p);
        p.level = level, p.cls = cls;
        OpRec &o = p.op;
        o.kind = kind, o.mode = mode;
        o.r0 = out.r0, o.c0 = out.c0, o.rows = out.rows, o.cols = out.cols;
        o.stride = out.stride;
        o.out_ch = out.kind == SRC_L4 ? 4 : 1;
        o.out_off = out.off;
        o.src_off = (int)srcs_p->size(), o.src_cnt = (int)ss.size();
        o.post = post, o.aux = aux, o.k0 = k0, o.k1 = k1, o.stencil = stencil;
        o.mul = mul;
        if (k)
            for (int i = 0; i < 4; i++)
                o.k[i] = k[i];
        for (auto &s : ss)
            srcs_p->push_back(s);
        ops_p->push_back(p);
    }

    // single-source pass: convert / multiply / post-process into a fresh layer
    Val unary(const Val &v, int out_kind, int want_pre, int want_lin, int out_pre, int out_lin, int post, float mul,
              int aux = 0)
    {
        SrcList sl(*this);
        std::vector<SrcRec> &ss = sl.v;
        int level = push_src(ss, v, want_pre, want_lin) + 1;
        Val out = alloc(out_kind, v.r0, v.c0, v.rows, v.cols, out_pre, out_lin, level);
        out.op_index = (int)ops_p->size();
        emit(0, OP_COMPOSE, out, ss, MODE_OVER, post, mul, out.level, nullptr, aux);
        return out;
    }

    // the region of a node its consumers can observe (demand_range below)
    struct DBox {
        bool any = false, inf = false;
        int r0 = 0, c0 = 0, r1 = 0, c1 = 0;
        void add_all() { any = inf = true; }
        void add(int a0, int b0, int a1, int b1)
        {
            if (a1 <= a0 || b1 <= b0 || inf)
                return;
            if (!any)
                r0 = a0, c0 = b0, r1 = a1, c1 = b1, any = true;
            else
                r0 = std::min(r0, a0), c0 = std::min(c0, b0), r1 = std::max(r1, a1), c1 = std::max(c1, b1);
        }
    };

    Val materialize(const Val &v)
    {
        if (v.kind == VAL_EMPTY || !v.is_virtual())
            return v;
        if (v.kind == VAL_LUMA) {
            // Scene.render mask branch (svgrasterize.py:733-737): luminance of the straight-alpha image
            Val l4 = v;
            l4.kind = SRC_L4;
            SrcRec s = src_of(l4, 0, 0);
            s.conv = v.luma_conv;
            Val out = alloc(SRC_L1, v.r0, v.c0, v.rows, v.cols, v.pre, v.lin, v.level + 1);
            emit(0, OP_COMPOSE, out, {s}, MODE_OVER, POST_LUMA, 1.0f, out.level);
            return out;
        }
        SrcList sl(*this);
        std::vector<SrcRec> &ss = sl.v;
        push_src(ss, v, v.pre, v.lin);  // identity conversion: never recurses into materialize
        Val out = alloc(v.one_channel() ? SRC_L1 : SRC_L4, v.r0, v.c0, v.rows, v.cols, v.pre, v.lin, v.level + 1);
        emit(0, OP_COMPOSE, out, ss, MODE_OVER, POST_NONE, 1.0f, out.level);
        return out;
    }

    // Layer.compose (svgrasterize.py:178-207)
    // intersect: blend on the intersection of the boxes whatever the mode (canvas_merge_intersect, :382-416);
    // raw: the layers are plain arrays, blended as they are (canvas_compose, :277-298, has no Layer.convert)
    Val compose(const std::vector<Val> &layers, int mode, const float *k, int lin, bool intersect = false,
                bool raw = false, const DBox *clip = nullptr)
    {
        if (layers.empty())
            return Val();
        if (layers.size() == 1)
            return layers[0];
        int pre = mode != MODE_ARITH;
        int r0, c0, r1, c1;
        if (mode == MODE_IN || intersect) {
            r0 = c0 = INT32_MIN, r1 = c1 = INT32_MAX;
            for (auto &l : layers) {
                r0 = std::max(r0, l.r0), c0 = std::max(c0, l.c0);
                r1 = std::min(r1, l.r0 + l.rows), c1 = std::min(c1, l.c0 + l.cols);
            }
            if (r1 - r0 <= 0 || c1 - c0 <= 0)
                return Val();
        } else {
            r0 = c0 = INT32_MAX, r1 = c1 = INT32_MIN;
            for (auto &l : layers) {
                r0 = std::min(r0, l.r0), c0 = std::min(c0, l.c0);
                r1 = std::max(r1, l.r0 + l.rows), c1 = std::max(c1, l.c0 + l.cols);
            }
        }
        if (clip && clip->any && !clip->inf) {
            // only this much of the result is ever looked at (demand_range)
            r0 = std::max(r0, clip->r0), c0 = std::max(c0, clip->c0);
            r1 = std::min(r1, clip->r1), c1 = std::min(c1, clip->c1);
            if (r1 - r0 <= 0 || c1 - c0 <= 0)
                return Val();
        }
        SrcList sl(*this);
        std::vector<SrcRec> &ss = sl.v;
        int level = 0;
        for (auto &l : layers)
            level = std::max(level, raw ? push_src(ss, l, l.pre, l.lin) : push_src(ss, l, pre, lin));
        level += 1;
        Val out = alloc(SRC_L4, r0, c0, r1 - r0, c1 - c0, pre, lin, level);
        out.op_index = (int)ops_p->size();
        emit(0, OP_COMPOSE, out, ss, mode, POST_NONE, 1.0f, level, k);
        return out;
    }

    static int py_int(double v) { return (int)v; }  // Python int(): truncation toward zero

    // ---- row-band renders.  Masks are clipped to the band, so a layer's origin differs from the one
    // the whole-canvas render would compute -- and Layer.convolve (:114) / filter_offset (:1849) truncate *toward
    // zero* from that origin: int(r0 - kw / 2) is a ceiling for a layer that starts within half a kernel of row 0
    // and a floor below it.  A band must place its filter results where the full render does, or the bands do not
    // join up.  The planner therefore also tracks, per node, the box of the full render (from the paths' boxes
    // against the full viewport, same box algebra, no ops) and takes the truncations from it.
    struct SBox {
        bool live = false;
        int r0 = 0, c0 = 0, rows = 0, cols = 0;
    };
    std::vector<SBox> shadow;  // empty unless the program is a band render

    static SBox shadow_union(const std::vector<SBox> &shadow, const svgr_node &n, const int32_t *ch, bool intersect,
                             bool need_all)
    {
        SBox o;
        int r0 = 0, c0 = 0, r1 = 0, c1 = 0;
        for (int k = 0; k < n.child_cnt; k++) {
            const SBox &b = shadow[ch[k]];
            if (!b.live) {
                if (need_all)
                    return SBox();
                continue;
            }
            if (!o.live) {
                r0 = b.r0, c0 = b.c0, r1 = b.r0 + b.rows, c1 = b.c0 + b.cols;
                o.live = true;
            } else if (intersect) {
                r0 = std::max(r0, b.r0), c0 = std::max(c0, b.c0);
                r1 = std::min(r1, b.r0 + b.rows), c1 = std::min(c1, b.c0 + b.cols);
            } else {
                r0 = std::min(r0, b.r0), c0 = std::min(c0, b.c0);
                r1 = std::max(r1, b.r0 + b.rows), c1 = std::max(c1, b.c0 + b.cols);
            }
        }
        if (!o.live || r1 <= r0 || c1 <= c0)
            return SBox();
        o.r0 = r0, o.c0 = c0, o.rows = r1 - r0, o.cols = c1 - c0;
        return o;
    }

    // the box node i has in the whole-canvas render (mirrors the box logic of node() below)
    void shadow_node(int i) { box_node(i, ctx->h_full_boxes, shadow); }

    // the box of node i from the path boxes `pb`, children's boxes taken from `shadow` (= the vector being filled)
    void box_node(int i, const std::vector<PathBox> &pb, std::vector<SBox> &shadow)
    {
        const svgr_node &n = ctx->h_nodes[i];
        const int32_t *ch = ctx->h_children.data() + n.child_off;
        SBox o;
        auto child = [&](int k) -> SBox { return (k < n.child_cnt && ch[k] >= 0 && ch[k] < i) ? shadow[ch[k]] : SBox(); };
        switch (n.tag) {
        case SVGR_N_LEAF:
            if (n.a >= 0 && n.a < ctx->n_path) {
                const PathBox &b = pb[n.a];
                if (b.rows > 0 && b.cols > 0) {
                    o.live = true, o.r0 = b.r0, o.c0 = b.c0, o.rows = b.rows, o.cols = b.cols;
                    if (n.b >= 0 && n.b < ctx->n_paint && ctx->h_paints[n.b].kind == PAINT_PATTERN &&
                        !(n.d >= 0 && n.d < i && shadow[n.d].live))
                        o = SBox();
                }
            }
            break;
        case SVGR_N_GROUP:
            o = shadow_union(shadow, n, ch, false, false);
            break;
        case SVGR_N_IN:
            o = shadow_union(shadow, n, ch, true, true);
            break;
        case SVGR_N_COMPOSE:
            o = shadow_union(shadow, n, ch, n.a == MODE_IN || (n.flags & 4), true);
            break;
        case SVGR_N_OPACITY:
        case SVGR_N_LUMA:
        case SVGR_N_SRC_ALPHA:
        case SVGR_N_CONVERT:
        case SVGR_N_CMATRIX:
        case SVGR_N_CANVAS:
            o = child(0);
            break;
        case SVGR_N_BLUR: {
            o = child(0);
            if (o.live && n.a >= 0 && n.a < (int)ctx->h_kernels.size()) {
                const svgr_kernel &kn = ctx->h_kernels[n.a];
                const int r0 = py_int((double)o.r0 - kn.rows / 2.0), c0 = py_int((double)o.c0 - kn.cols / 2.0);
                o.rows += kn.rows - 1, o.cols += kn.cols - 1, o.r0 = r0, o.c0 = c0;
            }
            break;
        }
        case SVGR_N_MORPH:
            o = child(0);
            if (o.live) {
                o.rows -= n.a - 1, o.cols -= n.b - 1;
                if (o.rows <= 0 || o.cols <= 0)
                    o = SBox();
            }
            break;
        case SVGR_N_OFFSET:
            o = child(0);
            if (o.live && n.a >= 0 && (size_t)(n.a + 1) * 12 <= ctx->h_offset_tr.size()) {
                int dr, dc;
                offset_shift(n, o.r0, o.c0, &dr, &dc);
                o.r0 += dr, o.c0 += dc;
            }
            break;
        case SVGR_N_MERGE_AT:
            if (child(0).live && n.c > 0 && n.d > 0)
                o.live = true, o.r0 = n.a, o.c0 = n.b, o.rows = n.c, o.cols = n.d;
            break;
        case SVGR_N_EXTERNAL:
            if (n.a >= 0 && n.a < (int)ctx->h_ext.size() && ctx->h_ext[n.a].rows > 0 && ctx->h_ext[n.a].cols > 0) {
                const svgr_external &e = ctx->h_ext[n.a];
                o.live = true, o.r0 = e.r0, o.c0 = e.c0, o.rows = e.rows, o.cols = e.cols;
            }
            break;
        default:
            break;
        }
        shadow[i] = o;
    }

    // ---- demand boxes.  compose([stencil, image], IN) lives on the intersection of its two boxes (:382-416), so
    // whatever a group under a clip or a mask renders outside the other operand's box is never seen -- in the c5 icon
    // the four blobs are composited over their whole union box and then clipped to the circle, the mask group is
    // rendered over 192 x 192 px for a stroke that covers a part of it.  Per plan range (scenes do not refer across
    // ranges): the natural box of every node (box_node, the algebra band renders already use), then top-down the
    // region of every node its consumers can observe.  Every operator between a group and its consumer here is
    // pointwise (pixel p of the input only reaches pixel p of the output), so the observable region of a child is
    // that of its parent cut to the parent's own box; filters with reach (blur, morphology, offset), merge_at,
    // pattern tiles and anything read back by the caller observe everything.  A group's fold is then emitted on its
    // box cut to its demand: same sources, same order, same arithmetic per pixel, fewer tiles.
    std::vector<SBox> natural;
    std::vector<DBox> demand;
    bool use_demand = true;

    void demand_range(int a, int b)
    {
        svgr_ctx *c = ctx;
        natural.resize((size_t)c->n_node);
        demand.resize((size_t)c->n_node);
        for (int i = a; i < b; i++) {
            box_node(i, c->h_boxes, natural);
            demand[i] = DBox();
        }
        for (int i = b - 1; i >= a; i--) {
            const svgr_node &n = c->h_nodes[i];
            DBox &d = demand[i];
            if (!d.any || (n.flags & 2) || n.tag == SVGR_N_CANVAS || i == c->n_node - 1)
                d.add_all();  // nobody above restricts it (or the caller reads it)
            const int32_t *ch = c->h_children.data() + n.child_off;
            const bool pointwise = n.tag == SVGR_N_GROUP || n.tag == SVGR_N_IN || n.tag == SVGR_N_COMPOSE ||
                                   n.tag == SVGR_N_OPACITY || n.tag == SVGR_N_LUMA || n.tag == SVGR_N_SRC_ALPHA ||
                                   n.tag == SVGR_N_CONVERT || n.tag == SVGR_N_CMATRIX || n.tag == SVGR_N_CANVAS;
            const SBox &nb = natural[i];
            for (int k = 0; k < n.child_cnt; k++) {
                const int kid = ch[k];
                if (kid < a || kid >= i)
                    continue;
                if (!pointwise || !nb.live)
                    demand[kid].add_all();
                else if (d.inf)
                    demand[kid].add(nb.r0, nb.c0, nb.r0 + nb.rows, nb.c0 + nb.cols);
                else
                    demand[kid].add(std::max(d.r0, nb.r0), std::max(d.c0, nb.c0), std::min(d.r1, nb.r0 + nb.rows),
                                    std::min(d.c1, nb.c0 + nb.cols));
            }
            if (n.tag == SVGR_N_LEAF && n.d >= a && n.d < i)
                demand[n.d].add_all();  // pattern tile: gathered from anywhere
        }
        if (getenv("SVGR_PLAN_DEBUG")) {
            long long nat_px = 0, seen_px = 0;
            for (int i = a; i < b; i++) {
                const svgr_node &n = c->h_nodes[i];
                if (n.tag != SVGR_N_LEAF || !natural[i].live)
                    continue;
                const SBox &nb = natural[i];
                const DBox &d = demand[i];
                nat_px += (long long)nb.rows * nb.cols;
                if (d.inf)
                    seen_px += (long long)nb.rows * nb.cols;
                else if (d.any) {
                    const long long rr = std::min(d.r1, nb.r0 + nb.rows) - std::max(d.r0, nb.r0);
                    const long long cc = std::min(d.c1, nb.c0 + nb.cols) - std::max(d.c0, nb.c0);
                    seen_px += std::max(0ll, rr) * std::max(0ll, cc);
                }
            }
            fprintf(stderr, "[plan] leaf mask pixels %lld, observable %lld (%.1f%%)\n", nat_px, seen_px,
                    100.0 * (double)seen_px / (double)std::max(1ll, nat_px));
        }
    }

    // filter_offset (svgrasterize.py:1844-1850) for a layer whose origin is (r0, c0): the integer shift.  1-D points
    // go through numpy's vector @ matrix path, whose rounding is fma(x, m0, y * m1) + m2.
    void offset_shift(const svgr_node &n, int r0, int c0, int *dr, int *dc) const
    {
        const double *f = ctx->h_offset_tr.data() + 12 * n.a, *inv = f + 6;
        double x = r0, y = c0;
        double ux = fma(x, inv[0], y * inv[1]) + inv[2], uy = fma(x, inv[3], y * inv[4]) + inv[5];
        ux = ux + n.f[0], uy = uy + n.f[1];
        double tx = fma(ux, f[0], uy * f[1]) + f[2], ty = fma(ux, f[3], uy * f[4]) + f[5];
        *dr = py_int(tx) - r0, *dc = py_int(ty) - c0;
    }

    // A separable stencil pass whose source is a materialised RGBA layer runs on the TMA-fed kernels
    // (k_stencil_tma.cu): the layer is a tensor, tile + halo a zero-filled box of it.  Anything else (coverage x
    // paint, one-channel layers) is read element by element through fetch_src by the generic kernel.
    bool tma_ok = getenv("SVGR_NO_TMA") == nullptr;
    int stencil_cls(const Val &src, bool horiz) const { return (tma_ok && src.kind == SRC_L4) ? (horiz ? 4 : 5) : 1; }

    // The generic stencil kernel stages a tile plus the whole kernel length in shared memory, which bounds the
    // length; the TMA-fed kernels do not have that bound but need an RGBA layer.  A long kernel over anything else
    // (coverage x paint, a one-channel layer) therefore gets its source written out as RGBA first.
    Val stencil_source(const Val &v, int k_rows, int k_cols, int want_pre, int want_lin)
    {
        const bool fits = (size_t)SVGR_STH_TR * (SVGR_STH_TC + k_cols - 1) * 16 <= SVGR_MAX_DYN_SMEM &&
                          (size_t)(SVGR_STV_TR + k_rows - 1) * SVGR_STV_TC * 16 <= SVGR_MAX_DYN_SMEM;
        if (v.kind == SRC_L4 || v.kind == VAL_EMPTY || (fits && !getenv("SVGR_FORCE_RGBA_STENCIL")) || !tma_ok)
            return v;
        // written out in the flags the stencil wants (a one-channel value is only relabelled, like Layer.convert)
        return unary(v, SRC_L4, want_pre, want_lin, want_pre, want_lin, POST_NONE, 1.0f);
    }

    bool node(int i, Val &out)
    {
        const svgr_node &n = ctx->h_nodes[i];
        const int32_t *ch = ctx->h_children.data() + n.child_off;
        const int lin = n.flags & 1;
        auto child = [&](int k) -> const Val & { return ctx->vals[ch[k]]; };
        for (int k = 0; k < n.child_cnt; k++)
            if (ch[k] < 0 || ch[k] >= i) {
                err = "scene program: children must precede their parent";
                return false;
            }
        out = Val();
        switch (n.tag) {
        case SVGR_N_EMPTY:
            break;
        case SVGR_N_LEAF: {
            if (n.a < 0 || n.a >= ctx->n_path) {
                err = "leaf: bad path index";
                return false;
            }
            const MaskRec &m = ctx->h_masks[n.a];
            if (m.rows <= 0 || m.cols <= 0)
                break;
            out.kind = n.b < 0 ? SRC_COV : SRC_COVPAINT;
            out.r0 = m.r0, out.c0 = m.c0, out.rows = m.rows, out.cols = m.cols, out.stride = m.stride;
            out.br0 = m.r0, out.bc0 = m.c0;
            out.off = m.off;
            out.paint = n.b;
            out.pre = 1, out.lin = n.b < 0 ? 1 : (n.c != 0);
            out.level = 0;
            if (n.b >= 0) {
                if (n.b >= ctx->n_paint) {
                    err = "leaf: bad paint index";
                    return false;
                }
                const PaintRec &p = ctx->h_paints[n.b];
                if (p.kind == PAINT_PATTERN) {
                    if (n.d < 0 || n.d >= i) {
                        err = "pattern leaf: bad pattern node";
                        return false;
                    }
                    const Val &pat = ctx->vals[n.d];
                    if (pat.kind != SRC_L4) {  // tile rendered to nothing: Path.fill returns None (:1064-1065)
                        out = Val();
                        break;
                    }
                    if (pat.missed) {
                        // the tile misses its repeat cell: canvas_merge_at returns None (:315-322) and the gather
                        // at :1094 subscripts it -- the reference raises TypeError
                        err = "'NoneType' object is not subscriptable";
                        err_code = SVGR_E_TYPE;
                        return false;
                    }
                    out.off2 = pat.off, out.stride2 = pat.stride;
                    out.pre = pat.pre, out.lin = pat.lin;
                    out.level = pat.level;
                } else if (p.kind == PAINT_RADIAL_FOCAL) {
                    FocalJob j;
                    j.paint = n.b, j.r0 = m.r0, j.c0 = m.c0, j.rows = m.rows, j.cols = m.cols;
                    if (!ctx->h_full_boxes.empty()) {
                        // "any det < 0" (:1621) is a property of the whole mask box, not of the band's part of it
                        const PathBox &fb = ctx->h_full_boxes[n.a];
                        if (fb.rows > 0 && fb.cols > 0)
                            j.r0 = fb.r0, j.c0 = fb.c0, j.rows = fb.rows, j.cols = fb.cols;
                    }
                    j.block_base = n_focal_blocks;
                    n_focal_blocks += (int)(((long long)j.rows * j.cols + 1023) / 1024);
                    focal_p->push_back(j);
                }
            }
            break;
        }
        case SVGR_N_GROUP: {
            std::vector<Val> &ls = scratch_vals;
            ls.clear();
            for (int k = 0; k < n.child_cnt; k++)
                if (child(k).kind != VAL_EMPTY)
                    ls.push_back(child(k));
            out = compose(ls, MODE_OVER, nullptr, lin, false, false, use_demand ? &demand[i] : nullptr);
            break;
        }
        case SVGR_N_OPACITY: {
            Val v = child(0);
            if (v.kind == VAL_EMPTY)
                break;
            if (v.kind == VAL_LUMA)
                v = materialize(v);
            float value = (float)n.f[0];
            if (v.one_channel() || (v.pre == 1 && v.lin == lin)) {
                out = v;  // Layer.opacity: convert is a relabel here, the multiply folds into the source
                out.mul = v.mul * value;
                out.pre = 1, out.lin = lin;
            } else {
                out = unary(v, SRC_L4, 1, lin, 1, lin, POST_NONE, value);
            }
            break;
        }
        case SVGR_N_IN: {
            if (n.child_cnt != 2) {
                err = "IN node needs two children";
                return false;
            }
            if (child(0).kind == VAL_EMPTY || child(1).kind == VAL_EMPTY)
                break;
            {
                // compose([stencil, image], IN) = image x stencil alpha on the intersection
                // (svgrasterize.py:382-416 with :290).  When the image needs no conversion the product is not
                // written out: the image keeps its storage and gains a stencil that its consumer applies.
                const Val &S = child(0), &I = child(1);
                bool img_ok = (I.kind == SRC_L4 || I.kind == SRC_COVPAINT) && I.st_kind == 0 && I.pre == 1 && I.lin == lin;
                bool st_ok = S.st_kind == 0 && (S.kind == SRC_COV || S.kind == SRC_L1 || S.kind == VAL_LUMA ||
                                                (S.kind == SRC_L4 && S.pre == 1 && S.lin == lin));
                if (img_ok && st_ok) {
                    int r0 = std::max(S.r0, I.r0), c0 = std::max(S.c0, I.c0);
                    int r1 = std::min(S.r0 + S.rows, I.r0 + I.rows), c1 = std::min(S.c0 + S.cols, I.c0 + I.cols);
                    if (r1 - r0 <= 0 || c1 - c0 <= 0)
                        break;
                    out = I;
                    out.r0 = r0, out.c0 = c0, out.rows = r1 - r0, out.cols = c1 - c0;
                    out.st_kind = S.kind == SRC_COV ? SRC_MOD_COV
                                  : S.kind == SRC_L1 ? SRC_MOD_L1
                                  : S.kind == SRC_L4 ? SRC_MOD_L4A
                                                     : SRC_MOD_LUMA;
                    out.st_br0 = S.br0, out.st_bc0 = S.bc0, out.st_stride = S.stride, out.st_off = S.off;
                    out.st_mul = S.mul, out.st_conv = S.luma_conv;
                    out.level = std::max(I.level, S.level);
                    out.op_index = -1;
                    break;
                }
            }
            out = compose({child(0), child(1)}, MODE_IN, nullptr, lin);
            break;
        }
        case SVGR_N_LUMA: {
            const Val &v = child(0);
            if (v.kind == VAL_EMPTY)
                break;
            if (v.one_channel()) {
                err = "luminance mask of a one-channel layer (the reference raises here)";
                return false;
            }
            if (v.kind == SRC_L4 && v.st_kind == 0 && !(n.flags & 2)) {
                out = v;  // deferred: becomes a stencil modifier of the masked layer, or a layer on demand
                out.kind = VAL_LUMA;
                out.luma_conv = SVGR_CONV(v.pre, v.lin, 0, lin);
                out.pre = 0, out.lin = lin;
                out.op_index = -1;
                break;
            }
            out = unary(v, SRC_L1, 0, lin, 0, lin, POST_LUMA, 1.0f);
            break;
        }
        case SVGR_N_COMPOSE: {
            std::vector<Val> ls;
            for (int k = 0; k < n.child_cnt; k++) {
                if (child(k).kind == VAL_EMPTY)
                    return true;  // a None input: the reference would fail later; propagate "nothing"
                ls.push_back(child(k));
            }
            float kk[4] = {(float)n.f[0], (float)n.f[1], (float)n.f[2], (float)n.f[3]};
            if (n.a < 0 || n.a > MODE_ARITH) {
                err = "invalid compose mode";
                return false;
            }
            out = compose(ls, n.a, kk, lin, (n.flags & 4) != 0, (n.flags & 8) != 0);
            break;
        }
        case SVGR_N_SRC_ALPHA: {
            const Val &v = child(0);
            if (v.kind == VAL_EMPTY)
                break;
            // alpha is the last channel whatever the flags; no conversion (svgrasterize.py:1803-1808)
            out = unary(v, SRC_L4, v.pre, v.lin, 1, 1, POST_ALPHA, 1.0f);
            break;
        }
        case SVGR_N_CONVERT: {
            const Val &v = child(0);
            if (v.kind == VAL_EMPTY)
                break;
            int pre = n.a < 0 ? v.pre : (n.a != 0), l = n.b < 0 ? v.lin : (n.b != 0);
            if (v.kind == VAL_LUMA) {
                out = materialize(v);
                out.pre = pre, out.lin = l;
            } else if (v.one_channel() || (v.pre == pre && v.lin == l)) {
                out = v;
                out.pre = pre, out.lin = l;
            } else {
                out = unary(v, SRC_L4, pre, l, pre, l, POST_NONE, 1.0f);
            }
            break;
        }
        case SVGR_N_BLUR: {
            const Val v_in = plain(child(0));
            if (v_in.kind == VAL_EMPTY)
                break;
            if (n.a < 0 || n.a >= (int)ctx->h_kernels.size()) {
                err = "blur: bad kernel index";
                return false;
            }
            const svgr_kernel &kn = ctx->h_kernels[n.a];
            const Val v = kn.separable ? stencil_source(v_in, kn.rows, kn.cols, 0, 1) : v_in;
            // Layer.convolve (svgrasterize.py:106-115): full convolution on straight-alpha linear RGBA
            int orows = v.rows + kn.rows - 1, ocols = v.cols + kn.cols - 1;
            // the truncation toward zero is taken where the whole-canvas render takes it (band renders: shadow box)
            const int fr0 = (!shadow.empty() && shadow[ch[0]].live) ? shadow[ch[0]].r0 : v.r0;
            const int fc0 = (!shadow.empty() && shadow[ch[0]].live) ? shadow[ch[0]].c0 : v.c0;
            int r0 = v.r0 + (py_int((double)fr0 - kn.rows / 2.0) - fr0), c0 = v.c0 + (py_int((double)fc0 - kn.cols / 2.0) - fc0);
            if (kn.separable) {
                Val tmp = alloc(SRC_L4, v.r0, c0, v.rows, ocols, 0, 1, v.level + 1);
                emit(stencil_cls(v, true), OP_STENCIL_H, tmp, {src_of(v, 0, 1)}, 0, 0, 1.0f, tmp.level, nullptr,
                     kn.weight_off + kn.rows, kn.cols, 0, STENCIL_CONV);
                out = alloc(SRC_L4, r0, c0, orows, ocols, 0, 1, tmp.level + 1);
                emit(stencil_cls(tmp, false), OP_STENCIL_V, out, {src_of(tmp, 0, 1)}, 0, 0, 1.0f, out.level, nullptr,
                     kn.weight_off, kn.rows, 0, STENCIL_CONV);
            } else {
                out = alloc(SRC_L4, r0, c0, orows, ocols, 0, 1, v.level + 1);
                emit(2, OP_CONV2D, out, {src_of(v, 0, 1)}, 0, 0, 1.0f, out.level, nullptr, kn.weight_off, kn.rows, kn.cols);
            }
            break;
        }
        case SVGR_N_MORPH: {
            const Val v_in = plain(child(0));
            if (v_in.kind == VAL_EMPTY)
                break;
            int k0 = n.a, k1 = n.b;
            const Val v = stencil_source(v_in, std::max(k0, 1), std::max(k1, 1), 1, 1);
            int st = n.c ? STENCIL_MAX : STENCIL_MIN;
            int orows = v.rows - k0 + 1, ocols = v.cols - k1 + 1;
            if (k0 < 1 || k1 < 1) {
                err = "morphology: window must be >= 1";
                return false;
            }
            if (orows <= 0 || ocols <= 0)
                break;
            // Layer.morphology (svgrasterize.py:120-127): premultiplied linear, top-left anchored, offset unchanged
            Val tmp = alloc(SRC_L4, v.r0, v.c0, v.rows, ocols, 1, 1, v.level + 1);
            emit(stencil_cls(v, true), OP_STENCIL_H, tmp, {src_of(v, 1, 1)}, 0, 0, 1.0f, tmp.level, nullptr, 0, k1, 0, st);
            out = alloc(SRC_L4, v.r0, v.c0, orows, ocols, 1, 1, tmp.level + 1);
            emit(stencil_cls(tmp, false), OP_STENCIL_V, out, {src_of(tmp, 1, 1)}, 0, 0, 1.0f, out.level, nullptr, 0, k0, 0,
                 st);
            break;
        }
        case SVGR_N_CMATRIX: {
            const Val &v = child(0);
            if (v.kind == VAL_EMPTY)
                break;
            if (n.a < 0 || n.a >= ctx->n_matrix) {
                err = "color matrix: bad index";
                return false;
            }
            out = unary(v, SRC_L4, 0, 1, 0, 1, POST_MATRIX, 1.0f, n.a);
            break;
        }
        case SVGR_N_OFFSET: {
            const Val &v = child(0);
            if (v.kind == VAL_EMPTY)
                break;
            if (n.a < 0 || (size_t)(n.a + 1) * 12 > ctx->h_offset_tr.size()) {
                err = "offset: bad transform index";
                return false;
            }
            // filter_offset (svgrasterize.py:1844-1850); the shift is computed from the origin the layer has in the
            // whole-canvas render (band renders: shadow box)
            // paint is evaluated at global pixel centres and stencils are anchored: move only real layers
            out = (v.kind == SRC_COVPAINT || v.kind == VAL_LUMA || v.st_kind) ? materialize(v) : v;
            {
                const bool sh = !shadow.empty() && shadow[ch[0]].live;
                int dr, dc;
                offset_shift(n, sh ? shadow[ch[0]].r0 : v.r0, sh ? shadow[ch[0]].c0 : v.c0, &dr, &dc);
                out.r0 += dr, out.c0 += dc, out.br0 += dr, out.bc0 += dc;
            }
            break;
        }
        case SVGR_N_MERGE_AT: {
            const Val &v = child(0);
            if (v.kind == VAL_EMPTY)
                break;
            if (n.c <= 0 || n.d <= 0)
                break;
            // canvas_merge_at onto zeros (svgrasterize.py:304-327): no conversion, result clipped to [0, 1]
            {
                Val vv = v.kind == VAL_LUMA ? materialize(v) : v;
                SrcList sl(*this);
        std::vector<SrcRec> &ss = sl.v;
                int level = push_src(ss, vv, vv.pre, vv.lin) + 1;
                Val o2 = alloc(SRC_L4, n.a, n.b, n.c, n.d, vv.pre, vv.lin, level);
                emit(0, OP_COMPOSE, o2, ss, MODE_OVER, POST_CLIP01, 1.0f, o2.level);
                out = o2;
                out.missed = std::min(vv.r0 + vv.rows, n.a + n.c) <= std::max(vv.r0, n.a) ||
                             std::min(vv.c0 + vv.cols, n.b + n.d) <= std::max(vv.c0, n.b);
            }
            break;
        }
        case SVGR_N_CANVAS: {
            Val v = n.child_cnt > 0 ? child(0) : Val();
            if (v.one_channel()) {
                err = "Only RGBA layers are supported";
                return false;
            }
            long long out_off = (long long)n.f[0];
            if (n.a < 0 || n.b < 0 || out_off < 0 || out_off + 4ll * n.a * n.b > ctx->canvas_bytes) {
                err = "canvas node outside the output buffer";
                return false;
            }
            // When the root is an over-group that nobody else reads, its fold writes the canvas directly
            // (clip, straight-alpha sRGB, RGBA8) instead of a float layer that is read back once.
            // The same for a feColorMatrix at the root (a filter that ends in a colour matrix): its pass applies the
            // matrix and then goes on to the canvas steps -- as long as the matrix has no bias column, because the
            // canvas is larger than the layer and M . 0 + bias would paint the outside.
            if (v.kind == SRC_L4 && !v.is_virtual() && v.op_index >= 0 && v.owner == ch[0] && (*uses_p)[ch[0]] == 1) {
                PlannedOp &po = (*ops_p)[v.op_index];
                const bool plain_over = v.pre == 1 && v.lin == lin && po.op.post == POST_NONE;
                bool matrix_root = false;
                if (po.op.post == POST_MATRIX && v.pre == 0 && v.lin == 1 && po.op.aux >= 0 &&
                    (size_t)(po.op.aux + 1) * 20 <= ctx->h_matrices.size()) {
                    const float *M = ctx->h_matrices.data() + 20 * (size_t)po.op.aux;
                    matrix_root = M[4] == 0.f && M[9] == 0.f && M[14] == 0.f && M[19] == 0.f;
                }
                if (po.cls == 0 && po.op.kind == OP_COMPOSE && po.op.mode == MODE_OVER && (plain_over || matrix_root) &&
                    po.op.mul == 1.0f) {
                    po.cls = 3;
                    OpRec &o = po.op;
                    o.kind = OP_CANVAS, o.k1 = lin;
                    o.r0 = n.c, o.c0 = n.d, o.rows = n.a, o.cols = n.b, o.stride = n.b, o.out_ch = 4;
                    o.out_off = out_off;
                    layer_pixels -= (long long)v.rows * v.cols;
                    out = Val();
                    break;
                }
            }
            PlannedOp p;
            memset(&p, 0, sizeof p);
            p.cls = 3;
            OpRec &o = p.op;
            o.kind = OP_CANVAS, o.mode = MODE_OVER, o.k1 = lin;
            o.r0 = n.c, o.c0 = n.d, o.rows = n.a, o.cols = n.b, o.stride = n.b, o.out_ch = 4;
            o.out_off = out_off;
            o.mul = 1.0f;
            SrcList sl(*this);
        std::vector<SrcRec> &ss = sl.v;
            int level = 0;
            if (v.kind != VAL_EMPTY)
                level = push_src(ss, v, 1, lin);
            p.level = level + 1;
            o.src_off = (int)srcs_p->size();
            o.src_cnt = (int)ss.size();
            for (auto &sr : ss)
                srcs_p->push_back(sr);
            ops_p->push_back(p);
            out = v;
            break;
        }
        case SVGR_N_EXTERNAL: {
            if (n.a < 0 || n.a >= (int)ctx->h_ext.size()) {
                err = "external: bad index";
                return false;
            }
            const svgr_external &e = ctx->h_ext[n.a];
            if (e.rows <= 0 || e.cols <= 0)
                break;
            out = alloc(e.channels == 1 ? SRC_L1 : SRC_L4, e.r0, e.c0, e.rows, e.cols, e.pre_alpha != 0,
                        e.linear_rgb != 0, 0);
            break;
        }
        default:
            err = "unknown node tag";
            return false;
        }
        if ((n.flags & 2) && out.kind != VAL_EMPTY)
            out = materialize(out);
        if (out.op_index >= 0 && out.owner < 0)
            out.owner = i;
        return true;
    }

    bool run() { return plan_masks() && plan_nodes(); }

    // masks, bands and coverage tiles: all the binning and coverage launches need
    bool plan_masks()
    {
        svgr_ctx *c = ctx;
        c->mask_pixels = 0;
        long long cov_top = 0, band = 0, tile = 0;
        for (int i = 0; i < c->n_path; i++) {
            MaskRec &m = c->h_masks[i];
            memset(&m, 0, sizeof m);
            const PathBox &b = c->h_boxes[i];
            m.band_base = (int)band, m.tile_base = (int)tile;
            m.fill_rule = c->h_paths[i].fill_rule;
            if (b.rows <= 0 || b.cols <= 0)
                continue;
            m.r0 = b.r0, m.c0 = b.c0, m.rows = b.rows, m.cols = b.cols;
            m.stride = (int)align4(b.cols);
            m.off = cov_top;
            cov_top += (long long)m.rows * m.stride;
            m.ntile_c = ceil_div(m.cols, SVGR_TILE_COLS);
            int nb = ceil_div(m.rows, SVGR_BAND_ROWS);
            band += nb;
            tile += (long long)nb * m.ntile_c;
            c->mask_pixels += (long long)m.rows * m.cols;
            if (band > 0x7fffff00ll || tile > 0x7fffff00ll) {
                err = "too many coverage tiles in one batch";
                return false;
            }
        }
        c->n_bands = band, c->n_cov_tiles = tile, c->cov_floats = cov_top;
        return true;
    }

    // ---- the scene program.  Planning is incremental: the node table is cut at scene boundaries (right after
    // canvas nodes) into chunks, and every chunk yields its own ops and launches, so that the pipeline can
    // upload and launch chunk k while the host plans chunk k + 1.
    std::vector<int> chunk_bounds;

    bool begin_nodes()
    {
        svgr_ctx *c = ctx;
        c->ops.clear(), c->srcs.clear(), c->focal_jobs.clear(), c->launches.clear();
        c->n_focal_blocks = 0, c->layer_pixels = 0, c->n_levels = 0;
        c->compose_bytes = 0, c->compose_bytes_8d = 0, c->canvas_pixels = 0;
        n_focal_blocks = 0, layer_pixels = 0, layer_top = 0;
        // reference counts and chunk boundaries depend on the program only: computed once per program
        if (!c->structure_ready)
            analyse_structure();
        uses_p = &c->node_uses;
        chunk_bounds = c->chunk_bounds;
        c->vals.resize(c->n_node);  // every entry is assigned by plan_range before anyone reads it
        shadow.clear();
        if (!c->h_full_boxes.empty())
            shadow.resize((size_t)c->n_node);
        return true;
    }

    void analyse_structure()
    {
        svgr_ctx *c = ctx;
        std::vector<int> &uses = c->node_uses;
        uses.assign(c->n_node, 0);
        std::vector<int> cuts;  // indices right after a canvas node: candidate scene boundaries
        int n_filter = 0;
        for (int i = 0; i < c->n_node; i++) {
            const svgr_node &n = c->h_nodes[i];
            for (int k = 0; k < n.child_cnt; k++) {
                int ch = c->h_children[n.child_off + k];
                if (ch >= 0 && ch < c->n_node)
                    uses[ch]++;
            }
            if (n.tag == SVGR_N_LEAF && n.d >= 0 && n.d < c->n_node)
                uses[n.d]++;
            if (n.flags & 2)
                uses[i]++;  // read back by the caller
            if (n.tag == SVGR_N_CANVAS)
                cuts.push_back(i + 1);
            if (n.tag == SVGR_N_BLUR || n.tag == SVGR_N_MORPH)
                n_filter++;
        }
        c->n_filter_nodes = n_filter;
        // chunks of at least 4096 nodes, at most 8 of them, cut where no reference crosses
        std::vector<int> &bounds = c->chunk_bounds;
        bounds.assign(1, 0);
        int max_chunks = 8;
        if (const char *e = getenv("SVGR_CHUNKS"))  // 1: plan everything before the first compose launch (clean kernel timings)
            max_chunks = std::max(1, atoi(e));
        const int want_chunks = std::min(max_chunks, c->n_node / 4096);
        if (want_chunks >= 2 && !cuts.empty()) {
            std::vector<int> min_ref(c->n_node + 1, c->n_node);  // smallest index referenced at or after i
            for (int i = c->n_node - 1; i >= 0; i--) {
                const svgr_node &n = c->h_nodes[i];
                int m = min_ref[i + 1];
                for (int k = 0; k < n.child_cnt; k++)
                    m = std::min(m, c->h_children[n.child_off + k]);
                if (n.tag == SVGR_N_LEAF && n.d >= 0)
                    m = std::min(m, n.d);
                min_ref[i] = m;
            }
            for (int t = 1; t < want_chunks; t++) {
                int want = (int)((long long)c->n_node * t / want_chunks);
                auto it = std::lower_bound(cuts.begin(), cuts.end(), want);
                if (it == cuts.end())
                    break;
                int cut = *it;
                if (cut > bounds.back() && cut < c->n_node && min_ref[cut] >= cut)
                    bounds.push_back(cut);
            }
        }
        bounds.push_back(c->n_node);
        c->structure_ready = true;
    }

    int n_chunks() const { return (int)chunk_bounds.size() - 1; }

    bool plan_range(int a, int b)
    {
        svgr_ctx *c = ctx;
        use_demand = !getenv("SVGR_NO_DEMAND");
        if (use_demand)
            demand_range(a, b);
        for (int i = a; i < b; i++) {
            if (!shadow.empty())
                shadow_node(i);
            // a value nobody reads is not computed (SourceAlpha of a filter that only uses SourceGraphic,
            // svgrasterize.py:1803-1809, is the common case)
            const svgr_node &n = c->h_nodes[i];
            if ((*uses_p)[i] == 0 && n.tag != SVGR_N_CANVAS && n.tag != SVGR_N_LEAF && i != c->n_node - 1) {
                c->vals[i] = Val();
                continue;
            }
            if (!node(i, c->vals[i]))
                return false;
        }
        return true;
    }

    // Plans chunk k.  On return ctx->ops[op_begin ..] are the chunk's ops, grouped by (level, class), and
    // ctx->launches[launch_begin ..] its launches (op_begin absolute).
    bool plan_chunk(int k, int *op_begin_out, int *launch_begin_out)
    {
        svgr_ctx *c = ctx;
        const size_t op_begin = c->ops.size();
        const size_t launch_begin = c->launches.size();
        *op_begin_out = (int)op_begin, *launch_begin_out = (int)launch_begin;
        if (!plan_range(chunk_bounds[k], chunk_bounds[k + 1]))
            return false;
        c->n_focal_blocks = n_focal_blocks, c->layer_pixels = layer_pixels;
        c->layer_floats = align4(layer_top);
        const size_t n_new = c->ops.size() - op_begin;
        if (n_new == 0)
            return true;
        // ---- order the chunk's ops by (level, class)
        {
            int max_level = 0;
            for (size_t q = op_begin; q < c->ops.size(); q++)
                max_level = std::max(max_level, c->ops[q].level);
            if (max_level < (1 << 20)) {  // counting sort: a chunk has a handful of levels and many ops
                std::vector<int> start((size_t)(max_level + 1) * 8 + 1, 0);
                for (size_t q = op_begin; q < c->ops.size(); q++)
                    start[(size_t)c->ops[q].level * 8 + c->ops[q].cls + 1]++;
                for (size_t q = 1; q < start.size(); q++)
                    start[q] += start[q - 1];
                sorted_ops.resize(n_new);
                for (size_t q = op_begin; q < c->ops.size(); q++)
                    sorted_ops[start[(size_t)c->ops[q].level * 8 + c->ops[q].cls]++] = c->ops[q];
                std::copy(sorted_ops.begin(), sorted_ops.end(), c->ops.begin() + op_begin);
            } else {
                std::stable_sort(c->ops.begin() + op_begin, c->ops.end(), [](const PlannedOp &a, const PlannedOp &b) {
                    return a.level != b.level ? a.level < b.level : a.cls < b.cls;
                });
            }
        }
        int nl = 0, last_level = -1;
        size_t i = op_begin;
        while (i < c->ops.size()) {
            size_t j = i;
            Launch L;
            L.cls = c->ops[i].cls, L.op_begin = (int)i, L.n_tiles = 0, L.smem = 0, L.list_slots = 0;
            long long tiles = 0, slots = 0;
            while (j < c->ops.size() && c->ops[j].level == c->ops[i].level && c->ops[j].cls == c->ops[i].cls) {
                OpRec &o = c->ops[j].op;
                int tr = SVGR_CMP_TR, tc = SVGR_CMP_TC;
                size_t smem = 0;
                if (c->ops[j].cls == 4) {
                    tr = SVGR_TMA_H_TR, tc = SVGR_TMA_H_TC;
                    L.k_max = std::max(L.k_max, o.k0);
                } else if (c->ops[j].cls == 5) {
                    tr = SVGR_TMA_V_TR, tc = SVGR_TMA_V_TC;
                    L.k_max = std::max(L.k_max, o.k0);
                } else if (o.kind == OP_STENCIL_H) {
                    tr = SVGR_STH_TR, tc = SVGR_STH_TC;
                    smem = (size_t)SVGR_STH_TR * (SVGR_STH_TC + o.k0 - 1) * 16;
                } else if (o.kind == OP_STENCIL_V) {
                    tr = SVGR_STV_TR, tc = SVGR_STV_TC;
                    smem = (size_t)(SVGR_STV_TR + o.k0 - 1) * SVGR_STV_TC * 16;
                } else if (o.kind == OP_CONV2D) {
                    tr = SVGR_C2D_TR, tc = SVGR_C2D_TC;
                    // the kernel walks the kernel rows in groups that fit the staging buffer (at least one row)
                    const size_t row_bytes = (size_t)(SVGR_C2D_TC + o.k1 - 1) * 16;
                    smem = (size_t)(SVGR_C2D_TR + o.k0 - 1) * row_bytes;
                    if (smem > SVGR_MAX_DYN_SMEM)
                        smem = std::max<size_t>(SVGR_C2D_TR, SVGR_MAX_DYN_SMEM / row_bytes) * row_bytes;
                }
                o.ntile_c = std::max(1, ceil_div(o.cols, tc));
                o.tile_base = (int)tiles;
                const long long op_tiles = (long long)ceil_div(o.rows, tr) * ceil_div(o.cols, tc);
                tiles += op_tiles;
                o.list_base = (int)slots;
                if (o.kind == OP_COMPOSE || o.kind == OP_CANVAS)
                    slots += op_tiles * o.src_cnt;
                L.smem = std::max(L.smem, smem);
                // algorithmic traffic: every output pixel written once, every source pixel read once
                const PlannedOp &po = c->ops[j];
                // ... and the same work in SURVEY 8(d)'s units: 36 B per layer pixel composited (4 B coverage +
                // 16 B destination read + 16 B written, one pass per layer), 20 B per quantised canvas pixel;
                // stencil passes move what they must either way
                const bool fold = po.cls == 0 || po.cls == 3;
                if (fold) {
                    if (o.mode != MODE_OVER || (o.post & (POST_ALPHA | POST_MATRIX | POST_LUMA)) ||
                        (po.cls == 3 ? o.k1 != 0 : o.out_ch != 4))
                        L.simple = false;
                    for (int q = 0; q < o.src_cnt && L.simple; q++) {
                        const SrcRec &sr = c->srcs[o.src_off + q];
                        if (sr.kind == SRC_L4 || sr.kind == SRC_COVPAINT) {
                            if ((sr.conv & 3) != ((sr.conv >> 2) & 3))
                                L.simple = false;
                            if (sr.kind == SRC_COVPAINT && c->h_paints[sr.paint].kind == PAINT_PATTERN)
                                L.simple = false;
                        }
                    }
                }
                if (po.cls == 3) {
                    c->canvas_pixels += (long long)o.rows * o.cols;
                    c->compose_bytes += (long long)o.rows * o.cols * 4;  // RGBA8 out
                    c->compose_bytes_8d += (long long)o.rows * o.cols * 20;
                } else {
                    c->compose_bytes += (long long)o.rows * o.cols * o.out_ch * 4;
                    if (!fold)
                        c->compose_bytes_8d += (long long)o.rows * o.cols * o.out_ch * 4;
                }
                for (int q = 0; q < o.src_cnt; q++) {
                    const SrcRec &sr = c->srcs[o.src_off + q];
                    long long rr, cc;
                    if (po.cls == 0 || po.cls == 3) {
                        rr = std::min(o.r0 + o.rows, sr.r0 + sr.rows) - std::max(o.r0, sr.r0);
                        cc = std::min(o.c0 + o.cols, sr.c0 + sr.cols) - std::max(o.c0, sr.c0);
                    } else {
                        rr = sr.rows, cc = sr.cols;
                    }
                    if (rr > 0 && cc > 0) {
                        const long long moved =
                            rr * cc * ((sr.kind == SRC_L4 || sr.kind == SRC_MOD_L4A || sr.kind == SRC_MOD_LUMA) ? 16 : 4);
                        c->compose_bytes += moved;
                        c->compose_bytes_8d += fold ? rr * cc * 36 : moved;  // a fused clip / mask is a layer pass too
                    }
                }
                j++;
            }
            if (tiles > 0x7fffff00ll || slots > 0x7fffff00ll) {
                err = "too many tiles in one launch";
                return false;
            }
            L.list_slots = slots;
            if (L.smem > SVGR_MAX_DYN_SMEM) {
                // only the generic kernels stage a whole kernel length at once (sources that are not plain RGBA
                // layers); the TMA-fed kernels cut long kernels into tap groups
                err = "stencil too long for shared-memory staging";
                err_code = SVGR_E_UNSUPPORTED;
                return false;
            }
            L.op_count = (int)(j - i), L.n_tiles = (int)tiles;
            c->launches.push_back(L);
            if (c->ops[i].level != last_level)
                nl++, last_level = c->ops[i].level;
            i = j;
        }
        c->n_levels = std::max(c->n_levels, nl);
        return true;
    }

    // everything at once (taps, CPU tests)
    bool plan_nodes()
    {
        if (!begin_nodes())
            return false;
        for (int k = 0; k < n_chunks(); k++) {
            int a, b;
            if (!plan_chunk(k, &a, &b))
                return false;
        }
        return true;
    }
};

template <class T>
cudaError_t upload(DevBuf &b, const T *src, size_t n, cudaStream_t s)
{
    cudaError_t e = b.ensure(std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess || n == 0)
        return e;
    return cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, s);
}

float ev_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// pipeline
// ---------------------------------------------------------------------------------------------
// Host copies of the tables the planner reads.  Deferred: svgr_render enqueues the uploads and the geometry
// kernels first and makes these copies while the GPU is busy (ctx->pending points at the caller's program,
// which outlives the call).
static int finish_load(svgr_ctx *ctx)
{
    const svgr_program *p = ctx->pending;
    if (!p)
        return SVGR_OK;
    ctx->pending = nullptr;
    ctx->h_paths.assign(p->paths, p->paths + p->n_path);
    ctx->h_paints.assign(p->paints, p->paints + p->n_paint);
    ctx->h_nodes.assign(p->nodes, p->nodes + p->n_node);
    ctx->h_children.assign(p->children, p->children + p->n_child);
    ctx->h_kernels.assign(p->kernels, p->kernels + p->n_kernel);
    ctx->h_matrices.assign(p->matrices, p->matrices + (size_t)p->n_matrix * 20);
    ctx->h_offset_tr.assign(p->offset_tr, p->offset_tr + (size_t)p->n_offset_tr * 12);
    ctx->h_ext.assign(p->externals, p->externals + p->n_external);
    ctx->h_ext_data.clear();
    for (int i = 0; i < p->n_external; i++) {
        const svgr_external &e = p->externals[i];
        if (e.channels != 1 && e.channels != 4)
            FAIL(SVGR_E_INVALID, "external layer must have 1 or 4 channels");
        size_t n = (size_t)std::max(e.rows, 0) * std::max(e.cols, 0) * e.channels;
        ctx->h_ext_data.emplace_back(e.image, e.image + n);
    }
    return SVGR_OK;
}

static int load_program(svgr_ctx *ctx, const svgr_program *p, cudaStream_t s, bool host_only = false)
{
    if (!p)
        FAIL(SVGR_E_INVALID, "null program");
    // ---- validate everything the device will index with before anything of the context changes: a rejected
    // program leaves the resident one (and svgr_render_resident) intact
    if (p->n_seg < 0 || p->n_path < 0 || p->n_node < 0 || p->n_stroke_seg < 0 || p->n_stroke < 0 || p->n_stroke_sub < 0 ||
        p->n_paint < 0 || p->n_stop < 0 || p->n_focal < 0 || p->n_child < 0 || p->n_kernel < 0 || p->n_weight < 0 ||
        p->n_matrix < 0 || p->n_offset_tr < 0 || p->n_external < 0 || p->canvas_bytes < 0)
        FAIL(SVGR_E_INVALID, "negative count in program");
    if (!(p->flatness >= 0.0))
        FAIL(SVGR_E_INVALID, "flatness must be >= 0 (0 = the reference's 0.1)");
    if (p->n_seg > 0x3fffffffll || p->n_stroke_seg > 0x3fffffffll)
        FAIL(SVGR_E_UNSUPPORTED, "more than 2^30 segments in one program");
    for (int i = 0; i < p->n_node; i++) {
        const svgr_node &n = p->nodes[i];
        if (n.child_cnt < 0 || n.child_off < 0 || (long long)n.child_off + n.child_cnt > p->n_child)
            FAIL(SVGR_E_INVALID, "node child range outside the children table");
    }
    for (long long i = 0; i < p->n_seg; i++)
        if (p->seg_path[i] >= (uint32_t)p->n_path)
            FAIL(SVGR_E_INVALID, "segment refers to a path outside the path table");
    for (int i = 0; i < p->n_stroke; i++) {
        const StrokeRec &q = p->strokes[i];
        if (q.path < 0 || q.path >= p->n_path)
            FAIL(SVGR_E_INVALID, "stroke refers to a path outside the path table");
        if (q.sub_begin < 0 || q.sub_end < q.sub_begin || q.sub_end > p->n_stroke_sub)
            FAIL(SVGR_E_INVALID, "stroke sub-path range outside the sub-path table");
    }
    if (p->n_stroke_sub > 0) {
        if (p->stroke_sub_off[0] < 0 || (long long)p->stroke_sub_off[p->n_stroke_sub] > p->n_stroke_seg)
            FAIL(SVGR_E_INVALID, "stroke sub-path offsets outside the stroke segments");
        for (int q = 0; q < p->n_stroke_sub; q++) {
            if (p->stroke_sub_off[q] > p->stroke_sub_off[q + 1])
                FAIL(SVGR_E_INVALID, "stroke sub-path offsets must not decrease");
            if (p->stroke_sub_job[q] < 0 || p->stroke_sub_job[q] >= p->n_stroke)
                FAIL(SVGR_E_INVALID, "stroke sub-path refers to a job outside the stroke table");
        }
    }
    for (long long i = 0; i < p->n_stroke_seg; i++)
        if (p->stroke_seg_job[i] < 0 || p->stroke_seg_job[i] >= p->n_stroke)
            FAIL(SVGR_E_INVALID, "stroke segment refers to a job outside the stroke table");
    for (int i = 0; i < p->n_paint; i++) {
        const PaintRec &q = p->paints[i];
        if (q.kind != PAINT_SOLID && q.kind != PAINT_PATTERN &&
            (q.stop_cnt < 1 || q.stop_off < 0 || q.stop_off + q.stop_cnt > p->n_stop))
            FAIL(SVGR_E_INVALID, "gradient without stops");
        if (q.kind == PAINT_RADIAL_FOCAL && (q.flag < 0 || q.flag >= p->n_focal))
            FAIL(SVGR_E_INVALID, "focal flag index out of range");
    }
    for (int i = 0; i < p->n_kernel; i++) {
        const svgr_kernel &k = p->kernels[i];
        if (k.rows < 1 || k.cols < 1 || k.weight_off < 0)
            FAIL(SVGR_E_INVALID, "convolution kernel with an empty extent");
        const long long need = k.separable ? (long long)k.rows + k.cols : (long long)k.rows * k.cols;
        if ((long long)k.weight_off + need > p->n_weight)
            FAIL(SVGR_E_INVALID, "convolution kernel weights outside the weight table");
    }
    for (int i = 0; i < p->n_external; i++) {
        const svgr_external &e = p->externals[i];
        if (e.channels != 1 && e.channels != 4)
            FAIL(SVGR_E_INVALID, "external layer must have 1 or 4 channels");
    }
    if (p->n_bbox_job < 0 || (p->n_bbox_job > 0 && !p->bbox_jobs))
        FAIL(SVGR_E_INVALID, "bad objectBoundingBox job table");
    for (int i = 0; i < p->n_bbox_job; i++) {
        const svgr_bbox_job &j = p->bbox_jobs[i];
        if (j.path < 0 || j.path >= p->n_path || j.paint < 0 || j.paint >= p->n_paint)
            FAIL(SVGR_E_INVALID, "objectBoundingBox job refers outside the path / paint tables");
        const int kind = p->paints[j.paint].kind;
        if (kind != PAINT_LINEAR && kind != PAINT_RADIAL && kind != PAINT_RADIAL_FOCAL)
            FAIL(SVGR_E_INVALID, "objectBoundingBox job on a paint that is not a gradient");
    }
    // ---- accepted: from here on the context describes the new program
    ctx->have_program = false;
    ctx->plan_cached = false;
    ctx->n_seg = p->n_seg, ctx->n_path = p->n_path, ctx->n_stroke = p->n_stroke, ctx->n_stroke_sub = p->n_stroke_sub;
    ctx->n_stroke_seg = p->n_stroke_seg, ctx->n_paint = p->n_paint, ctx->n_stop = p->n_stop, ctx->n_focal = p->n_focal;
    ctx->n_node = p->n_node, ctx->canvas_bytes = p->canvas_bytes;
    ctx->n_weight = p->n_weight, ctx->n_matrix = p->n_matrix;
    ctx->flatness = p->flatness;
    ctx->band_mode = false;
    for (int i = 0; i < p->n_path && !ctx->band_mode; i++)
        ctx->band_mode = p->paths[i].has_full != 0;
    if (!host_only) {
        CK(upload(ctx->d_seg_tag, p->seg_tag, (size_t)p->n_seg, s));
        CK(upload(ctx->d_seg_data, p->seg_data, (size_t)p->n_seg * 8, s));
        CK(upload(ctx->d_seg_path, p->seg_path, (size_t)p->n_seg, s));
        CK(upload(ctx->d_paths, p->paths, (size_t)p->n_path, s));
        CK(upload(ctx->d_strokes, p->strokes, (size_t)p->n_stroke, s));
        CK(upload(ctx->d_ssub_off, p->stroke_sub_off, p->n_stroke_sub > 0 ? (size_t)p->n_stroke_sub + 1 : 0, s));
        CK(upload(ctx->d_ssub_job, p->stroke_sub_job, (size_t)p->n_stroke_sub, s));
        {   // one warp assembles the outline of at most 256 consecutive segments of a sub-path
            ctx->h_sitems.clear();
            for (int q = 0; q < p->n_stroke_sub; q++) {
                const int a = p->stroke_sub_off[q], b = p->stroke_sub_off[q + 1];
                for (int x = a; x < b; x += 256) {
                    ctx->h_sitems.push_back(q);
                    ctx->h_sitems.push_back(x);
                    ctx->h_sitems.push_back(std::min(x + 256, b));
                }
            }
            ctx->n_sitems = (int)(ctx->h_sitems.size() / 3);
            CK(upload(ctx->d_sitems, ctx->h_sitems.data(), ctx->h_sitems.size(), s));
        }
        CK(upload(ctx->d_stag, p->stroke_tag, (size_t)p->n_stroke_seg, s));
        CK(upload(ctx->d_sdata, p->stroke_data, (size_t)p->n_stroke_seg * 8, s));
        CK(upload(ctx->d_sseg_job, p->stroke_seg_job, (size_t)p->n_stroke_seg, s));
        CK(upload(ctx->d_paints, p->paints, (size_t)p->n_paint, s));
        CK(upload(ctx->d_stops, p->stops, (size_t)p->n_stop, s));
        ctx->n_bbox_job = p->n_bbox_job;
        if (p->n_bbox_job > 0) {
            // CSR path -> jobs for cloud_bounds_kernel (one query per job, its path list = the leaf's own path)
            std::vector<int> &csr = ctx->h_bbox_csr;
            csr.assign((size_t)p->n_path + 1 + p->n_bbox_job, 0);
            for (int i = 0; i < p->n_bbox_job; i++)
                csr[p->bbox_jobs[i].path + 1]++;
            for (int i = 0; i < p->n_path; i++)
                csr[i + 1] += csr[i];
            std::vector<int> cur(csr.begin(), csr.begin() + p->n_path);
            for (int i = 0; i < p->n_bbox_job; i++)
                csr[(size_t)p->n_path + 1 + cur[p->bbox_jobs[i].path]++] = i;
            std::vector<double> &inv = ctx->h_bbox_inv;
            inv.resize((size_t)p->n_bbox_job * 6);
            for (int i = 0; i < p->n_bbox_job; i++)
                memcpy(&inv[6 * (size_t)i], p->bbox_jobs[i].inv, 48);
            CK(ctx->d_bbox_csr.ensure(csr.size() * sizeof(int)));
            CK(ctx->d_bbox_inv.ensure(inv.size() * sizeof(double)));
            CK(ctx->d_bbox_keys.ensure((size_t)p->n_bbox_job * 32));
            CK(cudaMemcpyAsync(ctx->d_bbox_csr.p, csr.data(), csr.size() * sizeof(int), cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(ctx->d_bbox_inv.p, inv.data(), inv.size() * sizeof(double), cudaMemcpyHostToDevice, s));
            CK(upload(ctx->d_bbox_jobs, p->bbox_jobs, (size_t)p->n_bbox_job, s));
        }
        CK(upload(ctx->d_matrices, p->matrices, (size_t)p->n_matrix * 20, s));
        CK(upload(ctx->d_weights, p->weights, (size_t)p->n_weight, s));
    }
    ctx->have_program = true;
    ctx->planned = ctx->covered = ctx->composed = false;
    ctx->structure_ready = false;
    ctx->pending = p;
    if (host_only)
        return finish_load(ctx);
    return SVGR_OK;
}

static int run_pipeline(svgr_ctx *ctx, cudaStream_t s, int stop_after, uint8_t *out, int out_on_device, int timing,
                        svgr_stats *stats, int depth = 0)
{
    if (!ctx->have_program)
        FAIL(SVGR_E_INVALID, "no program loaded");
    const int SM = ctx->sm_count;
    int retries = 0;
    long long n_kernels = 0;
    float ms_stroke = 0, ms_flatten = 0, ms_plan = 0, ms_bin = 0, ms_cov = 0, ms_cmp = 0, ms_canvas = 0, ms_d2h = 0;
    ctx->planned = ctx->covered = ctx->composed = false;
    auto mark = [&](int i) {
        if (timing)
            cudaEventRecord(ctx->ev[i], s);
    };

    // ---- capacities (grow-only, kept between calls)
    const long long S = ctx->n_stroke_seg;
    if (ctx->edge_cap == 0)
        ctx->edge_cap = 1024;
    {
        long long want = ctx->n_seg * 6 + S * 48 + 1024;
        if (ctx->edge_cap < want)
            ctx->edge_cap = want;
    }
    if (S > 0) {
        ctx->pool_cap = std::max(ctx->pool_cap, S * 8 + 256);
        ctx->outline_cap = std::max(ctx->outline_cap, S * 24 + 16ll * ctx->n_stroke_sub + 256);
    }
    CK(ctx->d_status.ensure(sizeof(StatusBlock)));
    CK(ctx->pin_status.ensure(sizeof(StatusBlock)));
    CK(ctx->d_minmax.ensure((size_t)std::max(ctx->n_path, 1) * 32));
    CK(ctx->d_boxes.ensure((size_t)std::max(ctx->n_path, 1) * sizeof(PathBox)));
    CK(ctx->d_minmax_f64.ensure((size_t)std::max(ctx->n_path, 1) * 32));
    CK(ctx->pin_boxes.ensure((size_t)std::max(ctx->n_path, 1) * sizeof(PathBox)));
    StatusBlock *d_st = ctx->d_status.as<StatusBlock>();
    StatusBlock st;

    for (;;) {
        CK(ctx->d_edges.ensure((size_t)ctx->edge_cap * 32));
        CK(ctx->d_edge_path.ensure((size_t)ctx->edge_cap * 4));
        CK(cudaMemsetAsync(d_st, 0, sizeof(StatusBlock), s));
        svgr_launch_minmax_init(ctx->d_minmax.as<unsigned long long>(), ctx->n_path, s);
        n_kernels += 1;
        mark(0);
        // Path.mask's literal flatness 0.1 (svgrasterize.py:955/:957) unless the program carries another one
        // (bezier3_flatten_batch(batch, flatness), :2091-2093: threshold = flatness^2 * 16)
        const double tol = ctx->flatness > 0.0 ? ctx->flatness : 0.1;
        const double thr = (tol * tol) * 16;
        // overflow lists of the multi-pass flatten (deep subdivision trees are re-spread over all warps)
        svgr_flat_overflow ovf;
        ovf.cap = 1 << 18;
        const bool flatten_wanted = stop_after != SVGR_STOP_STROKE;
        // the fill segments do not depend on the stroke outlines: their flatten runs on the side stream while the
        // stroke kernels (short, latency-bound) run here; both append to the same edge list with atomics
        const bool fill_on_side = flatten_wanted && S > 0 && ctx->n_seg > 0;
        if (flatten_wanted) {
            for (int q = 0; q < 2; q++) {
                CK(ctx->d_ovf_cubic[q].ensure((size_t)ovf.cap * 64));
                CK(ctx->d_ovf_path[q].ensure((size_t)ovf.cap * 4));
                CK(ctx->d_ovf_depth[q].ensure((size_t)ovf.cap));
                ovf.cubic[q] = ctx->d_ovf_cubic[q].as<double>(), ovf.path[q] = ctx->d_ovf_path[q].as<uint32_t>();
                ovf.depth[q] = ctx->d_ovf_depth[q].as<uint8_t>();
            }
            CK(ctx->d_ovf_counts.ensure(SVGR_FLAT_PASSES * sizeof(int)));
            ovf.counts = ctx->d_ovf_counts.as<int>();
            CK(cudaMemsetAsync(ovf.counts, 0, SVGR_FLAT_PASSES * sizeof(int), s));
        }
        if (fill_on_side) {
            CK(cudaEventRecord(ctx->ev_geo[0], s));
            CK(cudaStreamWaitEvent(ctx->up_stream, ctx->ev_geo[0], 0));
            svgr_launch_flatten(ctx->d_seg_tag.as<uint8_t>(), ctx->d_seg_data.as<double>(), ctx->d_seg_path.as<uint32_t>(),
                                ctx->n_seg, nullptr, ctx->d_paths.as<PathRec>(), thr, ctx->d_edges.as<double>(),
                                ctx->d_edge_path.as<uint32_t>(), (unsigned long long)ctx->edge_cap, &d_st->n_edges,
                                ctx->d_minmax.as<unsigned long long>(), SM, &ovf, ctx->up_stream);
            CK(cudaEventRecord(ctx->ev_geo[1], ctx->up_stream));
        }
        // ---- stroke outlines
        if (S > 0) {
            if (S > 0x3fffffff)
                FAIL(SVGR_E_INVALID, "too many stroke segments");
            const int nS = (int)S, nSub = ctx->n_stroke_sub;
            CK(ctx->d_scounts.ensure((size_t)(2 * nS + 1) * 4));
            CK(ctx->d_soffs.ensure((size_t)(2 * nS + 1) * 4));
            CK(ctx->d_scan_tmp.ensure((size_t)((2 * nS + nSub) / 2048 + 4) * 4));
            CK(ctx->d_pool.ensure((size_t)ctx->pool_cap * svgr_stroke_curve_bytes()));
            CK(ctx->d_sbound.ensure((size_t)(nSub + 1) * 4));
            CK(ctx->d_sout_off.ensure((size_t)(nSub + 1) * 4));
            CK(ctx->d_otag.ensure((size_t)ctx->outline_cap));
            CK(ctx->d_odata.ensure((size_t)ctx->outline_cap * 64));
            CK(ctx->d_opath.ensure((size_t)ctx->outline_cap * 4));
            CK(ctx->d_osub.ensure((size_t)ctx->outline_cap * 4));
            CK(cudaMemsetAsync(ctx->d_scounts.p, 0, (size_t)(2 * nS + 1) * 4, s));
            svgr_launch_stroke_count(ctx->d_stag.as<uint8_t>(), ctx->d_sdata.as<double>(), ctx->d_sseg_job.as<int>(),
                                     ctx->d_strokes.as<StrokeRec>(), nS, ctx->d_scounts.as<int>(), &d_st->stroke_err, s);
            svgr_launch_exclusive_scan(ctx->d_scounts.as<int>(), ctx->d_soffs.as<int>(), 2 * nS + 1,
                                       ctx->d_scan_tmp.as<int>(), &d_st->stroke_curves, s);
            svgr_launch_stroke_emit(ctx->d_stag.as<uint8_t>(), ctx->d_sdata.as<double>(), ctx->d_sseg_job.as<int>(),
                                    ctx->d_strokes.as<StrokeRec>(), nS, ctx->d_soffs.as<int>(), ctx->d_pool.p,
                                    (int)std::min<long long>(ctx->pool_cap, 0x7fffffff), s);
            svgr_launch_stroke_bound(ctx->d_ssub_off.as<int>(), nSub, nS, ctx->d_soffs.as<int>(), ctx->d_sbound.as<int>(), s);
            svgr_launch_exclusive_scan(ctx->d_sbound.as<int>(), ctx->d_sout_off.as<int>(), nSub, ctx->d_scan_tmp.as<int>(),
                                       &d_st->outline_total, s);
            svgr_launch_stroke_assemble(ctx->d_stag.as<uint8_t>(), ctx->d_ssub_off.as<int>(), ctx->d_ssub_job.as<int>(),
                                        ctx->d_strokes.as<StrokeRec>(), nSub, nS, ctx->d_sitems.as<int>(), ctx->n_sitems,
                                        ctx->d_soffs.as<int>(), ctx->d_pool.p,
                                        ctx->d_sbound.as<int>(), ctx->d_sout_off.as<int>(), &d_st->outline_total, 0,
                                        ctx->outline_cap, ctx->d_otag.as<uint8_t>(), ctx->d_odata.as<double>(),
                                        ctx->d_opath.as<uint32_t>(), ctx->d_osub.as<int32_t>(), &d_st->outline_count,
                                        &d_st->stroke_err, s);
            // count, 2 x (scan: local, sums[, add]), emit, bound, assemble
            n_kernels += 4 + 2 + 2 + ((2 * nS + 1) > 2048 ? 1 : 0) + (nSub > 2048 ? 1 : 0);
        }
        mark(1);
        // ---- flatten + bounds
        if (flatten_wanted) {
            if (fill_on_side)
                CK(cudaStreamWaitEvent(s, ctx->ev_geo[1], 0));  // the fill segments' edges are in the list
            else
                svgr_launch_flatten(ctx->d_seg_tag.as<uint8_t>(), ctx->d_seg_data.as<double>(), ctx->d_seg_path.as<uint32_t>(),
                                    ctx->n_seg, nullptr, ctx->d_paths.as<PathRec>(), thr, ctx->d_edges.as<double>(),
                                    ctx->d_edge_path.as<uint32_t>(), (unsigned long long)ctx->edge_cap, &d_st->n_edges,
                                    ctx->d_minmax.as<unsigned long long>(), SM, &ovf, s);
            if (S > 0)
                svgr_launch_flatten(ctx->d_otag.as<uint8_t>(), ctx->d_odata.as<double>(), ctx->d_opath.as<uint32_t>(),
                                    ctx->outline_cap, &d_st->outline_count, ctx->d_paths.as<PathRec>(), thr,
                                    ctx->d_edges.as<double>(), ctx->d_edge_path.as<uint32_t>(),
                                    (unsigned long long)ctx->edge_cap, &d_st->n_edges,
                                    ctx->d_minmax.as<unsigned long long>(), SM, &ovf, s);
            if (ctx->n_seg > 0 || S > 0) {
                svgr_launch_flatten_overflow(ctx->d_paths.as<PathRec>(), thr, ctx->d_edges.as<double>(),
                                             ctx->d_edge_path.as<uint32_t>(), (unsigned long long)ctx->edge_cap,
                                             &d_st->n_edges, ctx->d_minmax.as<unsigned long long>(), SM, &ovf, s);
                n_kernels += SVGR_FLAT_PASSES - 1;
            }
            if (ctx->band_mode) {
                CK(ctx->d_full_boxes.ensure((size_t)std::max(ctx->n_path, 1) * sizeof(PathBox)));
                CK(ctx->pin_full_boxes.ensure((size_t)std::max(ctx->n_path, 1) * sizeof(PathBox)));
            }
            svgr_launch_bounds(ctx->d_minmax.as<unsigned long long>(), ctx->d_paths.as<PathRec>(), ctx->n_path,
                               ctx->d_boxes.as<PathBox>(), ctx->band_mode ? ctx->d_full_boxes.as<PathBox>() : nullptr,
                               ctx->d_minmax_f64.as<double>(), s);
            if (ctx->band_mode && ctx->n_path > 0)
                CK(cudaMemcpyAsync(ctx->pin_full_boxes.p, ctx->d_full_boxes.p, (size_t)ctx->n_path * sizeof(PathBox),
                                   cudaMemcpyDeviceToHost, s));
            if (ctx->n_bbox_job > 0) {
                // objectBoundingBox gradients: the boxes of the flattened leaves exist now
                svgr_launch_resolve_paints(ctx->d_edges.as<double>(), ctx->d_edge_path.as<uint32_t>(), &d_st->n_edges,
                                           (unsigned long long)ctx->edge_cap, ctx->d_bbox_csr.as<int>(),
                                           ctx->d_bbox_csr.as<int>() + ctx->n_path + 1, ctx->d_bbox_inv.as<double>(),
                                           ctx->d_bbox_keys.as<unsigned long long>(),
                                           ctx->d_bbox_jobs.as<svgr_bbox_job>(), ctx->n_bbox_job,
                                           ctx->d_paths.as<PathRec>(), ctx->d_paints.as<PaintRec>(), &d_st->paint_err,
                                           SM, s);
                n_kernels += 3;
            }
            n_kernels += (ctx->n_seg > 0) + (S > 0) + (ctx->n_path > 0);
            if (ctx->n_path > 0)
                CK(cudaMemcpyAsync(ctx->pin_boxes.p, ctx->d_boxes.p, (size_t)ctx->n_path * sizeof(PathBox),
                                   cudaMemcpyDeviceToHost, s));
        }
        CK(cudaMemcpyAsync(ctx->pin_status.p, d_st, sizeof(StatusBlock), cudaMemcpyDeviceToHost, s));
        mark(2);
        {
            int rc_load = finish_load(ctx);  // host-side table copies overlap the geometry kernels
            if (rc_load != SVGR_OK)
                return rc_load;
        }
        CK(cudaStreamSynchronize(s));
        st = *(StatusBlock *)ctx->pin_status.p;
        if (timing) {
            ms_stroke += ev_ms(ctx->ev[0], ctx->ev[1]);
            ms_flatten += ev_ms(ctx->ev[1], ctx->ev[2]);
        }
        bool again = false;
        if ((long long)st.n_edges > ctx->edge_cap) {
            ctx->edge_cap = (long long)st.n_edges + 1024;
            again = true;
        }
        if (S > 0 && st.stroke_curves > ctx->pool_cap) {
            ctx->pool_cap = (long long)st.stroke_curves + 256;
            again = true;
        }
        if (S > 0 && st.outline_total > ctx->outline_cap) {
            ctx->outline_cap = (long long)st.outline_total + 256;
            again = true;
        }
        if (!again)
            break;
        if (++retries > 4)
            FAIL(SVGR_E_NOMEM, "capacity retry limit reached");
    }
    if (st.stroke_err & 1)
        FAIL(SVGR_E_STROKE, "cannot unpack non-iterable NoneType object");
    if (st.stroke_err & 2)
        FAIL(SVGR_E_INVALID, "stroke assembly overflow or unknown line cap");
    if (st.paint_err & 1)
        FAIL(SVGR_E_INVALID, "Singular matrix");  // numpy.linalg.LinAlgError in ConvexHull.bbox_transform(...).invert
    ctx->n_edges = (long long)st.n_edges;
    if (ctx->n_edges > 0x7fffffffll / 4)
        FAIL(SVGR_E_UNSUPPORTED, "more than 2^29 edges in one program: the band scans are 32-bit");
    ctx->h_boxes.assign((PathBox *)ctx->pin_boxes.p, (PathBox *)ctx->pin_boxes.p + ctx->n_path);
    if (ctx->band_mode && stop_after != SVGR_STOP_STROKE)
        ctx->h_full_boxes.assign((PathBox *)ctx->pin_full_boxes.p, (PathBox *)ctx->pin_full_boxes.p + ctx->n_path);
    else
        ctx->h_full_boxes.clear();

    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->n_edges = ctx->n_edges;
        stats->n_outline_segs = st.outline_count;
        stats->retries = retries;
        stats->plan_cached = 0;
        stats->ms_stroke = ms_stroke, stats->ms_flatten = ms_flatten;
    }
    if (stop_after == SVGR_STOP_STROKE || stop_after == SVGR_STOP_FLATTEN)
        return SVGR_OK;
    if (stop_after == SVGR_STOP_PLAN) {
        ctx->h_masks_store.resize((size_t)ctx->n_path + 1);
        ctx->h_masks = ctx->h_masks_store.data();
        Planner pl0(ctx);
        if (!pl0.run())
            FAIL(pl0.err_code, pl0.err);
        ctx->planned = true;
        return SVGR_OK;
    }

    // ---- plan, part 1: masks -> binning + coverage are launched before the node plan is made
    mark(3);
    Planner pl(ctx);
    // the plan cache holds when this is the program the cached plan was made for, the device found the same boxes
    // and nothing the tables point into has moved
    const bool cached = ctx->plan_cached && stop_after == SVGR_STOP_NONE && ctx->h_ext.empty() &&
                        ctx->cache_layers == ctx->d_layers.p && ctx->cache_boxes.size() == ctx->h_boxes.size() &&
                        (ctx->h_boxes.empty() || memcmp(ctx->cache_boxes.data(), ctx->h_boxes.data(),
                                                        ctx->h_boxes.size() * sizeof(PathBox)) == 0) &&
                        ctx->cache_full_boxes.size() == ctx->h_full_boxes.size() &&
                        (ctx->h_full_boxes.empty() || memcmp(ctx->cache_full_boxes.data(), ctx->h_full_boxes.data(),
                                                             ctx->h_full_boxes.size() * sizeof(PathBox)) == 0) &&
                        !getenv("SVGR_NO_PLAN_CACHE");
    ctx->plan_cached = false;  // until this render has gone through
    const size_t b_masks = (size_t)ctx->n_path * sizeof(MaskRec);
    auto t_h0 = std::chrono::steady_clock::now();
    if (!cached) {
        CK(ctx->pin_masks.ensure(b_masks + 64));  // the GPU waits for this table: it is written where it is copied from
        ctx->h_masks = (MaskRec *)ctx->pin_masks.p;
        if (!pl.plan_masks())
            FAIL(SVGR_E_INVALID, pl.err);
    }
    auto t_h1 = std::chrono::steady_clock::now();
    if (!cached) {
        CK(ctx->d_masks.ensure(std::max<size_t>(b_masks, 16)));
        if (b_masks)
            CK(cudaMemcpyAsync(ctx->d_masks.p, ctx->pin_masks.p, b_masks, cudaMemcpyHostToDevice, s));
    }
    CK(ctx->d_cov.ensure((size_t)std::max<long long>(ctx->cov_floats, 4) * 4));
    mark(4);

    // ---- binning (count -> scan -> fill).  The bin list capacity is a grow-only guess that is checked
    // after the fact (status block read at the end of the call): no host round trip in the middle.
    const long long NB = ctx->n_bands;
    ctx->bin_cap = std::max(ctx->bin_cap, 3 * ctx->n_edges + 1024);
    CK(ctx->d_band_cnt.ensure((size_t)(NB + 1) * 4));
    CK(ctx->d_band_off.ensure((size_t)(NB + 1) * 4));
    CK(ctx->d_band_cur.ensure((size_t)(NB + 1) * 4));
    CK(ctx->d_scan_tmp.ensure((size_t)(NB / 2048 + 4) * 4));
    CK(ctx->d_bin_edges.ensure((size_t)ctx->bin_cap * 4));
    CK(ctx->d_bin_data.ensure((size_t)ctx->bin_cap * 32));
    CK(ctx->d_tile_rec.ensure((size_t)std::max<long long>(ctx->n_cov_tiles, 1) * sizeof(TileRec)));
    CK(cudaMemsetAsync(ctx->d_band_cnt.p, 0, (size_t)(NB + 1) * 4, s));
    CK(cudaMemsetAsync(ctx->d_band_cur.p, 0, (size_t)(NB + 1) * 4, s));
    if (NB > 0 && ctx->n_edges > 0) {
        svgr_launch_bin_count(ctx->d_edges.as<double>(), ctx->d_edge_path.as<uint32_t>(), (unsigned long long)ctx->n_edges,
                              ctx->d_masks.as<MaskRec>(), ctx->d_band_cnt.as<int>(), SM, s);
        svgr_launch_exclusive_scan(ctx->d_band_cnt.as<int>(), ctx->d_band_off.as<int>(), NB + 1, ctx->d_scan_tmp.as<int>(),
                                   &d_st->binned_total, s);
        svgr_launch_bin_fill(ctx->d_edges.as<double>(), ctx->d_edge_path.as<uint32_t>(), (unsigned long long)ctx->n_edges,
                             ctx->d_masks.as<MaskRec>(), ctx->d_band_off.as<int>(), ctx->d_band_cur.as<int>(),
                             ctx->d_bin_edges.as<uint32_t>(), ctx->d_bin_data.as<double>(), ctx->bin_cap, SM, s);
        n_kernels += 4 + ((NB + 1) > 2048 ? 1 : 0);
    } else {
        CK(cudaMemsetAsync(ctx->d_band_off.p, 0, (size_t)(NB + 1) * 4, s));
    }
    mark(5);
    // ---- coverage
    svgr_launch_coverage(ctx->d_masks.as<MaskRec>(), ctx->n_path, (int)ctx->n_cov_tiles, ctx->d_tile_rec.as<TileRec>(),
                         ctx->d_band_off.as<int>(), ctx->d_band_cnt.as<int>(), ctx->d_bin_data.as<double>(), ctx->bin_cap,
                         ctx->d_cov.as<float>(), s);
    ctx->covered = true;
    n_kernels += ctx->n_cov_tiles > 0 ? 2 : 0;
    mark(6);

    // ---- plan, part 2 + pixels, interleaved: the node table is planned chunk by chunk (host); as soon as a
    // chunk's op tables exist they are copied up and its launches are issued, so the GPU composes chunk k
    // while the host plans chunk k + 1 (and bins / rasterises coverage while it plans chunk 0).
    auto t_h2 = std::chrono::steady_clock::now();
    float host_nodes_ms = 0.f;
    int n_launches = 0, timed_chunks = 0;
    bool early_copy_ok = true;
    long long copied_bytes = 0, copied_hi = 0;
    if (!cached && !pl.begin_nodes())
        FAIL(SVGR_E_INVALID, pl.err);
    {
        // capacities that hold for every chunk: device tables must not move while launches are in flight
        const size_t ops_cap = (size_t)ctx->n_node * 3 + 64, srcs_cap = ((size_t)ctx->h_children.size() + ctx->n_node) * 3 + 64;
        const size_t focal_cap = (size_t)ctx->n_path + 16;
        CK(ctx->d_ops.ensure(ops_cap * sizeof(OpRec)));
        CK(ctx->d_srcs.ensure(srcs_cap * sizeof(SrcRec)));
        CK(ctx->d_focal_jobs.ensure(focal_cap * sizeof(FocalJob)));
        CK(ctx->pin_plan.ensure(ops_cap * sizeof(OpRec) + srcs_cap * sizeof(SrcRec) + focal_cap * sizeof(FocalJob) + 64));
        CK(ctx->d_focal_flags.ensure((size_t)std::max(ctx->n_focal, 1) * 4));
        if (ctx->n_focal > 0 && !cached)
            CK(cudaMemsetAsync(ctx->d_focal_flags.p, 0, (size_t)ctx->n_focal * 4, s));
        // tensor maps: two passes per filter node, a source and a destination map per pass
        const size_t tmap_cap = (size_t)ctx->n_filter_nodes * 4 + 4;
        if (ctx->n_filter_nodes > 0) {
            CK(ctx->d_tmaps.ensure(tmap_cap * 128 + 128));
            CK(ctx->pin_tmaps.ensure(tmap_cap * 128 + 128));
        }
        size_t tmap_top = 0;
        char *pin_ops = (char *)ctx->pin_plan.p;
        char *pin_srcs = pin_ops + ops_cap * sizeof(OpRec);
        char *pin_focal = pin_srcs + srcs_cap * sizeof(SrcRec);
        uint8_t *canvas = nullptr;
        if (ctx->canvas_bytes > 0) {
            if (out_on_device && out) {
                canvas = out;
            } else {
                CK(ctx->d_canvas.ensure((size_t)std::max<long long>(ctx->canvas_bytes, 4)));
                canvas = ctx->d_canvas.as<uint8_t>();
            }
        }
        size_t up_ops = 0, up_srcs = 0, up_focal = 0;  // records already uploaded
        int focal_blocks_done = 0;
        bool externals_up = false;
        mark(10);
        // uploads and focal flags of a chunk run on their own stream, behind everything issued so far (paints,
        // the zeroed flags), so that they overlap the compose launches of the chunk before
        CK(cudaEventRecord(ctx->ev_up_begin, s));
        CK(cudaStreamWaitEvent(ctx->up_stream, ctx->ev_up_begin, 0));
        if (!cached) {
            ctx->chunk_launch_begin.clear(), ctx->chunk_op_begin.clear();
        }
        const int n_chunks = cached ? (int)ctx->chunk_launch_begin.size() - 1 : pl.n_chunks();
        long long heads_top = 0, lists_top = 0;  // tile heads / source lists of the chunks so far
        RenderTables T;
        for (int k = 0; k < n_chunks; k++) {
            int op_begin = 0, launch_begin = 0;
            size_t launch_end = 0, op_end = 0;
            if (cached) {
                op_begin = ctx->chunk_op_begin[k], launch_begin = ctx->chunk_launch_begin[k];
                op_end = (size_t)ctx->chunk_op_begin[k + 1], launch_end = (size_t)ctx->chunk_launch_begin[k + 1];
                T.srcs = ctx->d_srcs.as<SrcRec>(), T.paints = ctx->d_paints.as<PaintRec>(), T.stops = ctx->d_stops.as<StopRec>();
                T.focal_flags = ctx->d_focal_flags.as<int>(), T.cov = ctx->d_cov.as<float>(), T.layers = ctx->d_layers.as<float>();
                T.matrices = ctx->d_matrices.as<float>(), T.weights = ctx->d_weights.as<float>();
            } else {
            auto t_c0 = std::chrono::steady_clock::now();
            if (!pl.plan_chunk(k, &op_begin, &launch_begin))
                FAIL(pl.err_code, pl.err);
            host_nodes_ms += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_c0).count();
            ctx->chunk_launch_begin.push_back(launch_begin), ctx->chunk_op_begin.push_back(op_begin);
            launch_end = ctx->launches.size(), op_end = ctx->ops.size();
            if (stop_after == SVGR_STOP_COVERAGE)
                continue;
            if (ctx->ops.size() > ops_cap || ctx->srcs.size() > srcs_cap || ctx->focal_jobs.size() > focal_cap)
                FAIL(SVGR_E_NOMEM, "plan tables larger than their bound");
            // the layer arena may only grow while nothing is running (first render of a new size)
            if ((size_t)ctx->layer_floats * 4 > ctx->d_layers.cap) {
                CK(cudaStreamSynchronize(s));
                CK(ctx->d_layers.ensure((size_t)ctx->layer_floats * 4 + ((size_t)ctx->layer_floats * 4) / 2, true));
            }
            // tile -> op map of the stencil launches; tile heads and source lists of the compose launches.  The
            // lists of a whole chunk are written on the side stream while the chunk before it composes, so they
            // live in two buffers used by alternate chunks.
            // Every chunk gets its own stretch of the two buffers (they are kept for the plan cache, and a chunk's
            // lists are written on the side stream while the chunk before it composes).
            long long max_tiles = 1;
            for (size_t q = launch_begin; q < ctx->launches.size(); q++) {
                Launch &L = ctx->launches[q];
                if (L.cls == 0 || L.cls == 3) {
                    L.head_off = heads_top, L.list_off = lists_top;
                    heads_top += L.n_tiles, lists_top += L.list_slots;
                } else {
                    max_tiles = std::max<long long>(max_tiles, L.n_tiles);
                }
            }
            if ((size_t)max_tiles * 4 > ctx->d_tile_map.cap) {
                CK(cudaStreamSynchronize(s));
                CK(ctx->d_tile_map.ensure((size_t)max_tiles * 4 * 2));
            }
            if ((size_t)heads_top * sizeof(TileHead) > ctx->d_heads.cap || (size_t)lists_top * sizeof(TileEntry) > ctx->d_lists.cap) {
                // grow with the earlier chunks' lists preserved; nothing may be in flight while the buffers move
                CK(cudaStreamSynchronize(s));
                CK(cudaStreamSynchronize(ctx->up_stream));
                // room for the chunks still to come: they are about as large as the ones so far
                const double grow = (double)n_chunks / (double)(k + 1);
                CK(ctx->d_heads.ensure((size_t)((double)std::max<long long>(heads_top, 1) * sizeof(TileHead) * grow), true));
                CK(ctx->d_lists.ensure((size_t)((double)std::max<long long>(lists_top, 1) * sizeof(TileEntry) * grow), true));
            }
            // stage + upload the new records
            const size_t n_ops = ctx->ops.size() - up_ops, n_srcs = ctx->srcs.size() - up_srcs;
            const size_t n_focal = ctx->focal_jobs.size() - up_focal;
            OpRec *po = (OpRec *)pin_ops + up_ops;
            for (size_t q = 0; q < n_ops; q++)
                po[q] = ctx->ops[up_ops + q].op;
            cudaStream_t us = k < 16 ? ctx->up_stream : s;
            if (n_ops)
                CK(cudaMemcpyAsync(ctx->d_ops.as<OpRec>() + up_ops, po, n_ops * sizeof(OpRec), cudaMemcpyHostToDevice, us));
            if (n_srcs) {
                memcpy(pin_srcs + up_srcs * sizeof(SrcRec), ctx->srcs.data() + up_srcs, n_srcs * sizeof(SrcRec));
                CK(cudaMemcpyAsync(ctx->d_srcs.as<SrcRec>() + up_srcs, pin_srcs + up_srcs * sizeof(SrcRec),
                                   n_srcs * sizeof(SrcRec), cudaMemcpyHostToDevice, us));
            }
            if (n_focal) {
                memcpy(pin_focal + up_focal * sizeof(FocalJob), ctx->focal_jobs.data() + up_focal, n_focal * sizeof(FocalJob));
                CK(cudaMemcpyAsync(ctx->d_focal_jobs.as<FocalJob>() + up_focal, pin_focal + up_focal * sizeof(FocalJob),
                                   n_focal * sizeof(FocalJob), cudaMemcpyHostToDevice, us));
            }
            // tensor maps of this chunk's TMA-fed stencil launches: the layer arena cannot move any more before
            // they run (it only grows above, with the stream drained), so device addresses are final here
            {
                const size_t first = tmap_top;
                char *pin = (char *)(((uintptr_t)ctx->pin_tmaps.p + 63) & ~(uintptr_t)63);
                char *dev = (char *)(((uintptr_t)ctx->d_tmaps.p + 63) & ~(uintptr_t)63);
                for (size_t q = launch_begin; q < ctx->launches.size(); q++) {
                    Launch &L = ctx->launches[q];
                    if (L.cls != 4 && L.cls != 5)
                        continue;
                    L.tmap_off = (long long)tmap_top;
                    int kc = 0;
                    svgr_stencil_tma_smem(L.cls == 4, L.k_max, &kc);
                    for (int j = 0; j < L.op_count; j++) {
                        const OpRec &o = ctx->ops[L.op_begin + j].op;
                        const SrcRec &sr = ctx->srcs[o.src_off];
                        if (tmap_top + 2 > tmap_cap)
                            FAIL(SVGR_E_NOMEM, "tensor map table larger than its bound");
                        const float *src = ctx->d_layers.as<float>() + sr.off +
                                           4 * ((long long)(sr.r0 - sr.br0) * sr.stride + (sr.c0 - sr.bc0));
                        const int box_px = L.cls == 4 ? 128 : SVGR_TMA_V_TC;
                        const int box_rows = L.cls == 4 ? SVGR_TMA_H_TR : std::min(o.k0, kc) + SVGR_TMA_V_TR - 1;
                        if (svgr_encode_layer_map(pin + tmap_top * 128, src, sr.rows, sr.cols, sr.stride, box_px, box_rows))
                            FAIL(SVGR_E_CUDA, "cuTensorMapEncodeTiled failed for a stencil source");
                        if (L.cls == 4 &&
                            svgr_encode_layer_map(pin + (tmap_top + 1) * 128, ctx->d_layers.as<float>() + o.out_off, o.rows,
                                                  o.cols, o.stride, SVGR_TMA_H_TC / 2, SVGR_TMA_H_TR))
                            FAIL(SVGR_E_CUDA, "cuTensorMapEncodeTiled failed for a stencil destination");
                        tmap_top += 2;
                    }
                }
                if (tmap_top > first)
                    CK(cudaMemcpyAsync(dev + first * 128, pin + first * 128, (tmap_top - first) * 128, cudaMemcpyHostToDevice,
                                       us));
            }
            T.srcs = ctx->d_srcs.as<SrcRec>(), T.paints = ctx->d_paints.as<PaintRec>(), T.stops = ctx->d_stops.as<StopRec>();
            T.focal_flags = ctx->d_focal_flags.as<int>(), T.cov = ctx->d_cov.as<float>(), T.layers = ctx->d_layers.as<float>();
            T.matrices = ctx->d_matrices.as<float>(), T.weights = ctx->d_weights.as<float>();
            if (n_focal) {
                // block numbers of this chunk's jobs start at focal_blocks_done: rebase through the launch offset
                svgr_launch_focal_flags(T, ctx->d_focal_jobs.as<FocalJob>() + up_focal, (int)n_focal, focal_blocks_done,
                                        ctx->n_focal_blocks - focal_blocks_done, ctx->d_focal_flags.as<int>(), us);
                n_kernels += 1;
                focal_blocks_done = ctx->n_focal_blocks;
            }
            // per-tile source lists of this chunk's compose launches: they depend on the tables only
            for (size_t q = launch_begin; q < ctx->launches.size(); q++) {
                const Launch &L = ctx->launches[q];
                if (L.cls == 0 || L.cls == 3)
                    svgr_launch_cull(T, ctx->d_ops.as<OpRec>() + L.op_begin, L.op_count, L.n_tiles,
                                     ctx->d_heads.as<TileHead>() + L.head_off, ctx->d_lists.as<TileEntry>() + L.list_off, us);
            }
            if (us != s) {
                CK(cudaEventRecord(ctx->ev_up[k], us));
                CK(cudaStreamWaitEvent(s, ctx->ev_up[k], 0));
            }
            up_ops = ctx->ops.size(), up_srcs = ctx->srcs.size(), up_focal = ctx->focal_jobs.size();
            }  // !cached
            if (!externals_up) {
                externals_up = true;  // programs with external layers are planned as a single chunk
                for (int i = 0; i < ctx->n_node; i++) {
                    const svgr_node &n = ctx->h_nodes[i];
                    if (n.tag != SVGR_N_EXTERNAL || ctx->vals[i].kind == VAL_EMPTY)
                        continue;
                    const Val &v = ctx->vals[i];
                    const std::vector<float> &src = ctx->h_ext_data[n.a];
                    if (v.kind == SRC_L4)
                        CK(cudaMemcpyAsync(ctx->d_layers.as<float>() + v.off, src.data(), src.size() * 4,
                                           cudaMemcpyHostToDevice, s));
                    else
                        CK(cudaMemcpy2DAsync(ctx->d_layers.as<float>() + v.off, (size_t)v.stride * 4, src.data(),
                                             (size_t)v.cols * 4, (size_t)v.cols * 4, v.rows, cudaMemcpyHostToDevice, s));
                }
            }
            if (timing && k < 16)
                cudaEventRecord(ctx->ev_chunk[k][0], s);
            for (size_t q = launch_begin; q < launch_end; q++) {
                const Launch &L = ctx->launches[q];
                const OpRec *ops = ctx->d_ops.as<OpRec>() + L.op_begin;
                const int *tile_op = ctx->d_tile_map.as<int>();
                const TileHead *heads = ctx->d_heads.as<TileHead>() + L.head_off;
                const TileEntry *list = ctx->d_lists.as<TileEntry>() + L.list_off;
                if (L.cls == 1 || L.cls == 2 || L.cls == 4 || L.cls == 5)
                    svgr_launch_expand_ops(ops, L.op_count, L.n_tiles, ctx->d_tile_map.as<int>(), s);
                if (L.cls == 4 || L.cls == 5) {
                    const char *dev = (const char *)(((uintptr_t)ctx->d_tmaps.p + 63) & ~(uintptr_t)63);
                    if (svgr_launch_stencil_tma(T, ops, tile_op, L.n_tiles, dev + L.tmap_off * 128, L.cls == 4, L.k_max, SM,
                                                ctx->d_layers.as<float>(), s))
                        FAIL(SVGR_E_UNSUPPORTED, "stencil needs more shared memory than available");
                } else if (L.cls == 0)
                    svgr_launch_compose(T, ops, heads, list, L.n_tiles, L.simple, ctx->d_layers.as<float>(), nullptr, s);
                else if (L.cls == 1) {
                    if (svgr_launch_stencil(T, ops, tile_op, L.n_tiles, L.smem, ctx->d_layers.as<float>(), s))
                        FAIL(SVGR_E_UNSUPPORTED, "stencil needs more shared memory than available");
                } else if (L.cls == 2) {
                    if (svgr_launch_conv2d(T, ops, tile_op, L.n_tiles, L.smem, ctx->d_layers.as<float>(), s))
                        FAIL(SVGR_E_UNSUPPORTED, "convolution needs more shared memory than available");
                } else
                    svgr_launch_compose(T, ops, heads, list, L.n_tiles, L.simple, ctx->d_layers.as<float>(), canvas, s);
                n_launches++;
                n_kernels += L.n_tiles > 0 ? 2 : 0;
            }
            if (timing && k < 16)
                cudaEventRecord(ctx->ev_chunk[k][1], s);
            if (k < 16)
                CK(cudaEventRecord(ctx->ev_done[k], s));
            timed_chunks = std::min(k + 1, 16);
            // host output: the canvases this chunk finished go down on the copy stream while the next chunk
            // is planned and composed (chunks own increasing, disjoint byte ranges of the output)
            if (canvas && out && !out_on_device && early_copy_ok && k < 16) {
                long long lo = -1, hi = -1;
                for (size_t q = (size_t)op_begin; q < op_end; q++) {
                    const PlannedOp &po = ctx->ops[q];
                    if (po.cls != 3)
                        continue;
                    long long a = po.op.out_off, b = a + 4ll * po.op.rows * po.op.cols;
                    lo = lo < 0 ? a : std::min(lo, a);
                    hi = std::max(hi, b);
                }
                if (lo >= 0) {
                    if (lo < copied_hi) {
                        early_copy_ok = false;  // canvases not laid out in node order: one copy at the end
                    } else {
                        CK(cudaEventRecord(ctx->ev_copy[k], s));
                        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_copy[k], 0));
                        CK(cudaMemcpyAsync(out + lo, canvas + lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost,
                                           ctx->copy_stream));
                        copied_bytes += hi - lo;
                        copied_hi = hi;
                    }
                }
            }
        }
        if (!cached)
            ctx->chunk_launch_begin.push_back((int)ctx->launches.size()), ctx->chunk_op_begin.push_back((int)ctx->ops.size());
        ctx->planned = true;
        if (stop_after != SVGR_STOP_COVERAGE) {
            mark(7);
            mark(8);
            if (canvas && out && !out_on_device) {
                if (early_copy_ok && copied_bytes == ctx->canvas_bytes)
                    CK(cudaStreamSynchronize(ctx->copy_stream));  // everything already went down chunk by chunk
                else {
                    CK(cudaStreamSynchronize(ctx->copy_stream));
                    CK(cudaMemcpyAsync(out, canvas, (size_t)ctx->canvas_bytes, cudaMemcpyDeviceToHost, s));
                }
            }
            mark(9);
            ctx->composed = true;
        }
    }
    auto t_h3 = std::chrono::steady_clock::now();
    (void)t_h2, (void)t_h3;
    CK(cudaMemcpyAsync(ctx->pin_status.p, d_st, sizeof(StatusBlock), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    ctx->n_binned = ((StatusBlock *)ctx->pin_status.p)->binned_total;
    if (getenv("SVGR_FLAT_DEBUG") && ctx->d_ovf_counts.p) {  // nodes deferred by each flatten pass
        int counts[SVGR_FLAT_PASSES] = {0};
        cudaMemcpy(counts, ctx->d_ovf_counts.p, sizeof counts, cudaMemcpyDeviceToHost);
        fprintf(stderr, "flatten overflow per pass:");
        for (int q = 0; q < SVGR_FLAT_PASSES; q++)
            fprintf(stderr, " %d", counts[q]);
        fprintf(stderr, "\n");
    }
    if (ctx->n_binned > ctx->bin_cap) {
        // the guess was too small: the masks of the overflowing bands are wrong.  Grow and run again.
        ctx->bin_cap = ctx->n_binned + ctx->n_binned / 8 + 1024;
        if (depth >= 2)
            FAIL(SVGR_E_NOMEM, "bin list capacity retry limit reached");
        int rc = run_pipeline(ctx, s, stop_after, out, out_on_device, timing, stats, depth + 1);
        if (rc == SVGR_OK && stats)
            stats->retries += 1;
        return rc;
    }
    if (stop_after == SVGR_STOP_NONE) {
        // everything this render left on the device describes (program, boxes): the next render may reuse it
        ctx->plan_cached = true;
        ctx->cache_boxes = ctx->h_boxes, ctx->cache_full_boxes = ctx->h_full_boxes;
        ctx->cache_layers = ctx->d_layers.p;
        ctx->plan_cache_hits += cached ? 1 : 0;
    }
    if (timing) {
        ms_plan = ev_ms(ctx->ev[3], ctx->ev[4]);
        ms_bin = ev_ms(ctx->ev[4], ctx->ev[5]);
        ms_cov = ev_ms(ctx->ev[5], ctx->ev[6]);
        if (stop_after != SVGR_STOP_COVERAGE) {
            ms_cmp = ev_ms(ctx->ev[6], ctx->ev[7]);  // includes whatever host planning the launches had to wait for
            ms_canvas = ev_ms(ctx->ev[7], ctx->ev[8]);
            ms_d2h = ev_ms(ctx->ev[8], ctx->ev[9]);
        }
    }
    if (stats) {
        stats->n_bands = ctx->n_bands, stats->n_cov_tiles = ctx->n_cov_tiles, stats->n_binned = ctx->n_binned;
        stats->cov_floats = ctx->cov_floats, stats->layer_floats = ctx->layer_floats;
        stats->n_ops = (int64_t)ctx->ops.size(), stats->n_levels = ctx->n_levels;
        stats->n_launches = n_launches;
        stats->mask_pixels = ctx->mask_pixels, stats->layer_pixels = ctx->layer_pixels;
        stats->coverage_bytes = ctx->cov_floats * 4 + ctx->n_binned * 36;
        stats->compose_bytes = ctx->compose_bytes, stats->canvas_pixels = ctx->canvas_pixels;
        stats->compose_bytes_8d = ctx->compose_bytes_8d;
        stats->n_kernels = n_kernels;
        stats->host_plan_masks_ms = std::chrono::duration<float, std::milli>(t_h1 - t_h0).count();
        stats->host_plan_nodes_ms = host_nodes_ms;
        stats->plan_cached = cached ? 1 : 0;
        stats->ms_compose_busy = 0.f;
        if (timing && stop_after != SVGR_STOP_COVERAGE)
            for (int k = 0; k < timed_chunks; k++)
                stats->ms_compose_busy += ev_ms(ctx->ev_chunk[k][0], ctx->ev_chunk[k][1]);
        stats->ms_plan = ms_plan, stats->ms_bin = ms_bin, stats->ms_coverage = ms_cov, stats->ms_compose = ms_cmp;
        stats->ms_canvas = ms_canvas, stats->ms_d2h = ms_d2h;
        stats->ms_total = ms_stroke + ms_flatten + ms_plan + ms_bin + ms_cov + ms_cmp + ms_canvas + ms_d2h;
    }
    return SVGR_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {
#pragma GCC visibility push(default)

int svgr_version(void) { return SVGR_VERSION; }

int svgr_sizeof(int what)
{
    switch (what) {
    case 0: return (int)sizeof(PathRec);
    case 1: return (int)sizeof(StrokeRec);
    case 2: return (int)sizeof(PaintRec);
    case 3: return (int)sizeof(StopRec);
    case 4: return (int)sizeof(svgr_node);
    case 5: return (int)sizeof(svgr_kernel);
    case 6: return (int)sizeof(svgr_external);
    case 7: return (int)sizeof(svgr_program);
    case 8: return (int)sizeof(svgr_stats);
    case 9: return (int)sizeof(MaskRec);
    case 10: return (int)sizeof(svgr_bbox_job);
    default: return -1;
    }
}

int svgr_create(int device, svgr_ctx **out)
{
    if (!out)
        return SVGR_E_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n)
        return SVGR_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess)
        return SVGR_E_CUDA;
    svgr_ctx *ctx = new svgr_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess)
        ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return SVGR_E_CUDA;
    }
    for (auto &e : ctx->ev)
        cudaEventCreate(&e);
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (auto &e : ctx->ev_copy)
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto &e : ctx->ev_chunk)
        cudaEventCreate(&e[0]), cudaEventCreate(&e[1]);
    cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking);
    for (auto &e : ctx->ev_up)
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->ev_up_begin, cudaEventDisableTiming);
    for (auto &e : ctx->ev_geo)
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    cudaEventCreate(&ctx->ev_png[0]), cudaEventCreate(&ctx->ev_png[1]);
    for (auto &e : ctx->ev_done)
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    *out = ctx;
    return SVGR_OK;
}

void svgr_destroy(svgr_ctx *ctx)
{
    if (!ctx)
        return;
    cudaSetDevice(ctx->device);
    DevBuf *bufs[] = {&ctx->d_sitems, &ctx->d_seg_tag, &ctx->d_seg_data, &ctx->d_seg_path, &ctx->d_paths, &ctx->d_strokes, &ctx->d_ssub_off,
                      &ctx->d_ssub_job, &ctx->d_stag, &ctx->d_sdata, &ctx->d_sseg_job, &ctx->d_paints, &ctx->d_stops,
                      &ctx->d_matrices, &ctx->d_weights, &ctx->d_scounts, &ctx->d_soffs, &ctx->d_scan_tmp, &ctx->d_pool,
                      &ctx->d_sbound, &ctx->d_sout_off, &ctx->d_sout_total, &ctx->d_otag, &ctx->d_odata, &ctx->d_opath,
                      &ctx->d_osub, &ctx->d_ocount, &ctx->d_edges, &ctx->d_edge_path, &ctx->d_minmax, &ctx->d_boxes,
                      &ctx->d_minmax_f64, &ctx->d_status, &ctx->d_masks, &ctx->d_band_cnt, &ctx->d_band_off,
                      &ctx->d_band_cur, &ctx->d_bin_edges, &ctx->d_cov, &ctx->d_layers, &ctx->d_ops, &ctx->d_srcs,
                      &ctx->d_focal_jobs, &ctx->d_focal_flags, &ctx->d_canvas, &ctx->d_q, &ctx->d_tile_map, &ctx->d_tile_rec, &ctx->d_bin_data, &ctx->d_heads, &ctx->d_lists, &ctx->d_ovf_cubic[0], &ctx->d_ovf_cubic[1], &ctx->d_ovf_path[0], &ctx->d_ovf_path[1],
                      &ctx->d_ovf_depth[0], &ctx->d_ovf_depth[1], &ctx->d_ovf_counts, &ctx->d_eager[0], &ctx->d_eager[1],
                      &ctx->d_eager[2], &ctx->d_tmaps, &ctx->d_full_boxes, &ctx->d_png_segs, &ctx->d_png_canvases, &ctx->d_png_scratch,
                      &ctx->d_png_seg_bytes, &ctx->d_png_seg_adler, &ctx->d_png_file_bytes, &ctx->d_png_file_off,
                      &ctx->d_png_out, &ctx->d_png_in, &ctx->d_png_code, &ctx->d_bbox_jobs, &ctx->d_bbox_csr, &ctx->d_bbox_inv,
                      &ctx->d_bbox_keys};
    for (DevBuf *b : bufs)
        b->release();
    ctx->pin_boxes.release(), ctx->pin_status.release(), ctx->pin_plan.release(), ctx->pin_out.release();
    ctx->pin_masks.release();
    ctx->pin_tmaps.release();
    ctx->pin_full_boxes.release();
    ctx->pin_png.release();
    for (auto &e : ctx->ev_png)
        if (e)
            cudaEventDestroy(e);
    for (auto &e : ctx->ev)
        if (e)
            cudaEventDestroy(e);
    for (auto &e : ctx->ev_chunk) {
        if (e[0])
            cudaEventDestroy(e[0]);
        if (e[1])
            cudaEventDestroy(e[1]);
    }
    for (auto &e : ctx->ev_copy)
        if (e)
            cudaEventDestroy(e);
    if (ctx->copy_stream)
        cudaStreamDestroy(ctx->copy_stream);
    for (auto &e : ctx->ev_up)
        if (e)
            cudaEventDestroy(e);
    if (ctx->ev_up_begin)
        cudaEventDestroy(ctx->ev_up_begin);
    for (auto &e : ctx->ev_geo)
        if (e)
            cudaEventDestroy(e);
    for (auto &e : ctx->ev_done)
        if (e)
            cudaEventDestroy(e);
    if (ctx->up_stream)
        cudaStreamDestroy(ctx->up_stream);
    if (ctx->own_stream)
        cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *svgr_last_error(svgr_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

// A call that failed half way may have left copies and kernels in flight on the side streams: nothing of it
// may still be running when the caller reuses its buffers or renders again.
// ---- PNG files from RGBA8 canvases that are in device memory (k_png.cu) -----------------------------------------
namespace {
struct PngSegH {
    int32_t canvas, row0, rows, last;
    int64_t slot;
};
struct PngCanvasH {
    int64_t src;
    int32_t rows, cols, seg0, nseg;
};
struct PngImage {
    int64_t src;
    int32_t rows, cols;
};
}  // namespace

// images: canvases inside `canvas_buf` (device).  out / offsets as in svgr_render_png.
static int encode_png(svgr_ctx *ctx, cudaStream_t s, const uint8_t *canvas_buf, const std::vector<PngImage> &images, uint8_t *out,
                      int64_t out_cap, int out_on_device, int64_t *offsets, int timing, svgr_stats *stats)
{
    const int n = (int)images.size();
    if (!offsets)
        FAIL(SVGR_E_INVALID, "png: offsets must not be null");
    offsets[0] = 0;
    if (n == 0)
        return SVGR_OK;
    std::vector<PngSegH> segs;
    std::vector<PngCanvasH> cvs((size_t)n);
    int64_t slot_top = 0;
    for (int i = 0; i < n; i++) {
        const PngImage &im = images[(size_t)i];
        if (im.rows <= 0 || im.cols <= 0)
            FAIL(SVGR_E_INVALID, "png: empty canvas");
        const int per = svgr_png_rows_per_segment(im.cols);
        cvs[(size_t)i] = {im.src, im.rows, im.cols, (int32_t)segs.size(), (im.rows + per - 1) / per};
        for (int r = 0; r < im.rows; r += per) {
            const int rows = std::min(per, im.rows - r);
            segs.push_back({i, r, rows, r + rows >= im.rows ? 1 : 0, slot_top});
            slot_top += svgr_png_slot_bytes(rows, im.cols);
        }
    }
    const size_t n_seg = segs.size();
    CK(ctx->d_png_segs.ensure(n_seg * sizeof(PngSegH)));
    CK(ctx->d_png_canvases.ensure((size_t)n * sizeof(PngCanvasH)));
    CK(ctx->d_png_scratch.ensure((size_t)slot_top + 64));
    CK(ctx->d_png_seg_bytes.ensure(n_seg * 4));
    CK(ctx->d_png_seg_adler.ensure(n_seg * 8));
    CK(ctx->d_png_file_bytes.ensure((size_t)n * 4));
    CK(ctx->d_png_file_off.ensure((size_t)n * 8));
    CK(ctx->pin_png.ensure((size_t)n * 4 + (size_t)n * 8 + 64));
    if (timing)
        cudaEventRecord(ctx->ev_png[0], s);
    CK(cudaMemcpyAsync(ctx->d_png_segs.p, segs.data(), n_seg * sizeof(PngSegH), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_png_canvases.p, cvs.data(), (size_t)n * sizeof(PngCanvasH), cudaMemcpyHostToDevice, s));
    CK(ctx->d_png_code.ensure(svgr_png_code_scratch_bytes()));
    const int n_deflate_kernels =
        svgr_launch_png_deflate(ctx->d_png_segs.p, (int)n_seg, ctx->d_png_canvases.p, canvas_buf,
                                ctx->d_png_scratch.as<uint8_t>(), ctx->d_png_seg_bytes.as<int>(),
                                ctx->d_png_seg_adler.as<unsigned>(), ctx->d_png_code.p, s);
    svgr_launch_png_sizes(ctx->d_png_canvases.p, n, ctx->d_png_seg_bytes.as<int>(), ctx->d_png_file_bytes.as<int>(), s);
    int *h_sizes = (int *)ctx->pin_png.p;
    long long *h_off = (long long *)((char *)ctx->pin_png.p + (((size_t)n * 4 + 15) & ~(size_t)15));
    CK(cudaMemcpyAsync(h_sizes, ctx->d_png_file_bytes.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));  // the file sizes decide where every file goes
    int64_t total = 0;
    for (int i = 0; i < n; i++) {
        h_off[i] = total;
        offsets[i] = total;
        total += h_sizes[i];
    }
    offsets[n] = total;
    if (out && total > out_cap)
        FAIL(SVGR_E_NOMEM, "png: output buffer too small (offsets[n_canvas] holds the size needed)");
    uint8_t *dst = (out && out_on_device) ? out : nullptr;
    if (!dst) {
        CK(ctx->d_png_out.ensure((size_t)total + 64));
        dst = ctx->d_png_out.as<uint8_t>();
    }
    CK(cudaMemcpyAsync(ctx->d_png_file_off.p, h_off, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    svgr_launch_png_pack(ctx->d_png_canvases.p, n, ctx->d_png_segs.p, ctx->d_png_seg_bytes.as<int>(),
                         ctx->d_png_seg_adler.as<unsigned>(), ctx->d_png_scratch.as<uint8_t>(),
                         ctx->d_png_file_off.as<long long>(), dst, s);
    if (timing)
        cudaEventRecord(ctx->ev_png[1], s);
    if (out && !out_on_device)
        CK(cudaMemcpyAsync(out, dst, (size_t)total, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (stats) {
        stats->png_bytes = total;
        stats->n_kernels += 2 + n_deflate_kernels;
        if (timing) {
            stats->ms_png = ev_ms(ctx->ev_png[0], ctx->ev_png[1]);
            stats->ms_total += stats->ms_png;
        }
    }
    return SVGR_OK;
}

// the canvases of the resident program, in canvas node order
static int program_canvases(svgr_ctx *ctx, std::vector<PngImage> &images)
{
    images.clear();
    for (int i = 0; i < ctx->n_node; i++) {
        const svgr_node &n = ctx->h_nodes[(size_t)i];
        if (n.tag == SVGR_N_CANVAS)
            images.push_back({(int64_t)n.f[0], n.a, n.b});
    }
    return SVGR_OK;
}

static void drain_after_failure(svgr_ctx *ctx, cudaStream_t s)
{
    cudaStreamSynchronize(s);
    if (ctx->up_stream)
        cudaStreamSynchronize(ctx->up_stream);
    if (ctx->copy_stream)
        cudaStreamSynchronize(ctx->copy_stream);
    cudaGetLastError();
}

int svgr_render(svgr_ctx *ctx, const svgr_program *prog, void *stream, int stop_after, uint8_t *out, int out_on_device,
                int timing, svgr_stats *stats)
{
    if (!ctx)
        return SVGR_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->own_stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timing) {
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
    }
    int rc = load_program(ctx, prog, s);
    if (timing)
        cudaEventRecord(e1, s);
    if (rc == SVGR_OK)
        rc = run_pipeline(ctx, s, stop_after, out, out_on_device, timing, stats);
    if (rc != SVGR_OK)
        drain_after_failure(ctx, s);
    if (ctx->pending) {  // the call failed before the host tables were copied: nothing resident to re-render
        ctx->pending = nullptr;
        ctx->have_program = false;
    }
    if (rc != SVGR_OK && rc != SVGR_E_INVALID)
        ctx->have_program = false;  // a render that died half way leaves nothing to re-render either
    if (timing) {
        if (rc == SVGR_OK && stats) {
            stats->ms_h2d = ev_ms(e0, e1);
            stats->ms_total += stats->ms_h2d;
        }
        cudaEventDestroy(e0), cudaEventDestroy(e1);
    }
    return rc;
}

int svgr_render_png(svgr_ctx *ctx, const svgr_program *prog, void *stream, uint8_t *out, int64_t out_cap, int out_on_device,
                    int64_t *offsets, int timing, svgr_stats *stats)
{
    if (!ctx)
        return SVGR_E_INVALID;
    // the canvases stay in the context's device buffer; only the files are copied out
    int rc = svgr_render(ctx, prog, stream, SVGR_STOP_NONE, nullptr, 1, timing, stats);
    if (rc != SVGR_OK)
        return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->own_stream;
    std::vector<PngImage> images;
    program_canvases(ctx, images);
    rc = encode_png(ctx, s, ctx->d_canvas.as<uint8_t>(), images, out, out_cap, out_on_device, offsets, timing, stats);
    if (rc != SVGR_OK)
        drain_after_failure(ctx, s);
    return rc;
}

int svgr_render_resident_png(svgr_ctx *ctx, void *stream, uint8_t *out, int64_t out_cap, int out_on_device,
                             int64_t *offsets, int timing, svgr_stats *stats)
{
    if (!ctx)
        return SVGR_E_INVALID;
    int rc = svgr_render_resident(ctx, stream, nullptr, timing, stats);
    if (rc != SVGR_OK)
        return rc;
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->own_stream;
    std::vector<PngImage> images;
    program_canvases(ctx, images);
    rc = encode_png(ctx, s, ctx->d_canvas.as<uint8_t>(), images, out, out_cap, out_on_device, offsets, timing, stats);
    if (rc != SVGR_OK)
        drain_after_failure(ctx, s);
    return rc;
}

int svgr_png_encode(svgr_ctx *ctx, const uint8_t *images, int32_t n_images, const int32_t *rows, const int32_t *cols,
                    uint8_t *out, int64_t out_cap, int64_t *offsets)
{
    if (!ctx)
        return SVGR_E_INVALID;
    if (n_images < 0 || (n_images > 0 && (!images || !rows || !cols)))
        FAIL(SVGR_E_INVALID, "png_encode: bad arguments");
    CK(cudaSetDevice(ctx->device));
    std::vector<PngImage> list;
    int64_t top = 0;
    for (int i = 0; i < n_images; i++) {
        if (rows[i] <= 0 || cols[i] <= 0)
            FAIL(SVGR_E_INVALID, "png_encode: empty image");
        list.push_back({top, rows[i], cols[i]});
        top += 4ll * rows[i] * cols[i];
    }
    cudaStream_t s = ctx->own_stream;
    CK(ctx->d_png_in.ensure((size_t)std::max<int64_t>(top, 4)));
    if (top > 0)
        CK(cudaMemcpyAsync(ctx->d_png_in.p, images, (size_t)top, cudaMemcpyHostToDevice, s));
    const int rc = encode_png(ctx, s, ctx->d_png_in.as<uint8_t>(), list, out, out_cap, 0, offsets, 0, nullptr);
    if (rc != SVGR_OK)
        drain_after_failure(ctx, s);
    return rc;
}

// Host-only: runs the planner `reps` times on a program and given path boxes without touching CUDA
// (profiling and CPU tests of the planner).  info = {ops, srcs, launches, levels, layer floats (low 31 bits)}.
int svgr_debug_plan(const svgr_program *prog, const int32_t *boxes, int reps, float *ms_masks, float *ms_nodes,
                    int64_t *info)
{
    svgr_ctx *ctx = new svgr_ctx();
    int rc = load_program(ctx, prog, nullptr, true);
    if (rc == SVGR_OK) {
        ctx->h_boxes.assign((const PathBox *)boxes, (const PathBox *)boxes + ctx->n_path);
        ctx->h_full_boxes.clear();
        ctx->h_masks_store.resize((size_t)ctx->n_path + 1);
        ctx->h_masks = ctx->h_masks_store.data();
        float best_m = 1e30f, best_n = 1e30f;
        for (int r = 0; r < reps && rc == SVGR_OK; r++) {
            Planner pl(ctx);
            auto t0 = std::chrono::steady_clock::now();
            bool ok = pl.plan_masks();
            auto t1 = std::chrono::steady_clock::now();
            auto tb0 = std::chrono::steady_clock::now();
            ok = ok && pl.begin_nodes();
            auto tb1 = std::chrono::steady_clock::now();
            for (int k = 0; ok && k < pl.n_chunks(); k++) {
                int a, b;
                ok = pl.plan_chunk(k, &a, &b);
            }
            auto t2 = std::chrono::steady_clock::now();
            if (getenv("SVGR_PLAN_DEBUG"))
                fprintf(stderr, "[plan] begin_nodes %.3f ms, chunks %.3f ms (%d chunks)\n",
                        std::chrono::duration<float, std::milli>(tb1 - tb0).count(),
                        std::chrono::duration<float, std::milli>(t2 - tb1).count(), pl.n_chunks());
            if (!ok)
                rc = SVGR_E_INVALID;
            best_m = std::min(best_m, std::chrono::duration<float, std::milli>(t1 - t0).count());
            best_n = std::min(best_n, std::chrono::duration<float, std::milli>(t2 - t1).count());
        }
        if (ms_masks)
            *ms_masks = best_m;
        if (ms_nodes)
            *ms_nodes = best_n;
        if (info) {
            info[0] = (int64_t)ctx->ops.size(), info[1] = (int64_t)ctx->srcs.size();
            info[2] = (int64_t)ctx->launches.size(), info[3] = ctx->n_levels, info[4] = ctx->layer_floats;
            info[5] = ctx->compose_bytes;
        }
        if (getenv("SVGR_PLAN_DEBUG")) {
            // tiles per compose op / sources per op, weighted by tiles
            int64_t by_tiles[8] = {0}, by_srcs[8] = {0}, total = 0;
            for (const PlannedOp &po : ctx->ops) {
                const OpRec &o = po.op;
                if (o.kind != OP_COMPOSE && o.kind != OP_CANVAS)
                    continue;
                const int64_t nt = (int64_t)((o.rows + SVGR_CMP_TR - 1) / SVGR_CMP_TR) * o.ntile_c;
                int b = 0, c = 0;
                while ((1 << b) < nt && b < 7)
                    b++;
                while ((1 << c) < o.src_cnt && c < 7)
                    c++;
                by_tiles[b] += nt, by_srcs[c] += nt, total += nt;
            }
            fprintf(stderr, "[plan] compose tiles %lld; share by tiles/op (<=1,2,4,..,128+):", (long long)total);
            for (int i = 0; i < 8; i++)
                fprintf(stderr, " %.3f", (double)by_tiles[i] / (double)std::max<int64_t>(total, 1));
            fprintf(stderr, "\n[plan] share by sources/op (<=1,2,4,..,128+):");
            for (int i = 0; i < 8; i++)
                fprintf(stderr, " %.3f", (double)by_srcs[i] / (double)std::max<int64_t>(total, 1));
            fprintf(stderr, "\n");
        }
    }
    delete ctx;
    return rc;
}

int svgr_render_resident(svgr_ctx *ctx, void *stream, uint8_t *out_device, int timing, svgr_stats *stats)
{
    if (!ctx)
        return SVGR_E_INVALID;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->own_stream;
    const int rc = run_pipeline(ctx, s, SVGR_STOP_NONE, out_device, 1, timing, stats);
    if (rc != SVGR_OK)
        drain_after_failure(ctx, s);
    return rc;
}

int svgr_read_edges(svgr_ctx *ctx, double *edges, uint32_t *edge_path, int64_t cap, int64_t *n_edges)
{
    if (!ctx || !n_edges)
        return SVGR_E_INVALID;
    *n_edges = ctx->n_edges;
    int64_t n = std::min<int64_t>(cap, ctx->n_edges);
    if (n > 0 && edges)
        CK(cudaMemcpy(edges, ctx->d_edges.p, (size_t)n * 32, cudaMemcpyDeviceToHost));
    if (n > 0 && edge_path)
        CK(cudaMemcpy(edge_path, ctx->d_edge_path.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return SVGR_OK;
}

int svgr_read_boxes(svgr_ctx *ctx, int32_t *boxes, double *minmax)
{
    if (!ctx || !boxes)
        return SVGR_E_INVALID;
    if ((int)ctx->h_boxes.size() != ctx->n_path)
        FAIL(SVGR_E_INVALID, "no flatten result to read");
    if (ctx->n_path > 0)
        memcpy(boxes, ctx->h_boxes.data(), (size_t)ctx->n_path * sizeof(PathBox));
    if (minmax && ctx->n_path > 0)
        CK(cudaMemcpy(minmax, ctx->d_minmax_f64.p, (size_t)ctx->n_path * 32, cudaMemcpyDeviceToHost));
    return SVGR_OK;
}

int svgr_read_mask(svgr_ctx *ctx, int32_t path, float *out)
{
    if (!ctx || !out)
        return SVGR_E_INVALID;
    if (!ctx->covered || path < 0 || path >= ctx->n_path)
        FAIL(SVGR_E_INVALID, "no coverage result for this path");
    const MaskRec &m = ctx->h_masks[path];
    if (m.rows <= 0 || m.cols <= 0)
        return SVGR_OK;
    CK(cudaMemcpy2D(out, (size_t)m.cols * 4, ctx->d_cov.as<float>() + m.off, (size_t)m.stride * 4, (size_t)m.cols * 4,
                    m.rows, cudaMemcpyDeviceToHost));
    return SVGR_OK;
}

int svgr_read_bins(svgr_ctx *ctx, int32_t path, int32_t *band_off, uint32_t *bin_edges, int64_t cap, int64_t *n_binned)
{
    if (!ctx || !band_off || !n_binned)
        return SVGR_E_INVALID;
    if (!ctx->covered || path < 0 || path >= ctx->n_path)
        FAIL(SVGR_E_INVALID, "no binning result for this path");
    const MaskRec &m = ctx->h_masks[path];
    int nb = m.rows > 0 ? ceil_div(m.rows, SVGR_BAND_ROWS) : 0;
    std::vector<int> off(nb + 1, 0);
    if (nb > 0)
        CK(cudaMemcpy(off.data(), ctx->d_band_off.as<int>() + m.band_base, (size_t)(nb + 1) * 4, cudaMemcpyDeviceToHost));
    int base = off[0];
    for (int i = 0; i <= nb; i++)
        band_off[i] = off[i] - base;
    *n_binned = off[nb] - base;
    int64_t n = std::min<int64_t>(cap, *n_binned);
    if (n > 0 && bin_edges)
        CK(cudaMemcpy(bin_edges, ctx->d_bin_edges.as<uint32_t>() + base, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return SVGR_OK;
}

int svgr_read_outline(svgr_ctx *ctx, uint8_t *tag, double *data, uint32_t *path, int32_t *sub, int64_t cap,
                      int64_t *n_segs)
{
    if (!ctx || !n_segs)
        return SVGR_E_INVALID;
    StatusBlock st = *(StatusBlock *)ctx->pin_status.p;
    int64_t total = ctx->n_stroke_seg > 0 ? st.outline_count : 0;
    *n_segs = total;
    int64_t n = std::min<int64_t>(cap, total);
    if (n > 0) {
        if (tag)
            CK(cudaMemcpy(tag, ctx->d_otag.p, (size_t)n, cudaMemcpyDeviceToHost));
        if (data)
            CK(cudaMemcpy(data, ctx->d_odata.p, (size_t)n * 64, cudaMemcpyDeviceToHost));
        if (path)
            CK(cudaMemcpy(path, ctx->d_opath.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        if (sub)
            CK(cudaMemcpy(sub, ctx->d_osub.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    }
    return SVGR_OK;
}

int svgr_node_info(svgr_ctx *ctx, int32_t node, int32_t *info)
{
    if (!ctx || !info)
        return SVGR_E_INVALID;
    if (!ctx->planned || node < 0 || node >= ctx->n_node)
        FAIL(SVGR_E_INVALID, "no plan / bad node index");
    const Val &v = ctx->vals[node];
    info[0] = v.kind == VAL_EMPTY ? 0 : (v.one_channel() ? 2 : 1);
    info[1] = v.r0, info[2] = v.c0, info[3] = v.rows, info[4] = v.cols;
    info[5] = v.pre, info[6] = v.lin;
    info[7] = v.is_virtual() ? 1 : 0;  // 1: not materialised, svgr_read_node refuses
    return SVGR_OK;
}

int svgr_read_node(svgr_ctx *ctx, int32_t node, float *out)
{
    if (!ctx || !out)
        return SVGR_E_INVALID;
    if (!ctx->composed || node < 0 || node >= ctx->n_node)
        FAIL(SVGR_E_INVALID, "no render result / bad node index");
    const Val &v = ctx->vals[node];
    if (v.kind == VAL_EMPTY)
        return SVGR_OK;
    if (v.is_virtual())
        FAIL(SVGR_E_INVALID, "node is not materialised (set flags bit 1 on it)");
    if (v.kind == SRC_L4) {
        CK(cudaMemcpy(out, ctx->d_layers.as<float>() + v.off, (size_t)v.rows * v.cols * 16, cudaMemcpyDeviceToHost));
    } else {
        const float *base = (v.kind == SRC_COV ? ctx->d_cov.as<float>() : ctx->d_layers.as<float>()) + v.off;
        CK(cudaMemcpy2D(out, (size_t)v.cols * 4, base, (size_t)v.stride * 4, (size_t)v.cols * 4, v.rows,
                        cudaMemcpyDeviceToHost));
    }
    return SVGR_OK;
}

int svgr_cloud_bounds(svgr_ctx *ctx, int32_t n_query, const int32_t *q_off, const int32_t *q_paths, const double *q_inv,
                      double *out)
{
    if (!ctx || n_query < 0 || (n_query > 0 && (!q_off || !q_paths || !q_inv || !out)))
        return SVGR_E_INVALID;
    if (n_query == 0)
        return SVGR_OK;
    if ((int)ctx->h_boxes.size() != ctx->n_path)
        FAIL(SVGR_E_INVALID, "no flatten result");
    CK(cudaSetDevice(ctx->device));
    // invert the query -> paths lists into a CSR path -> queries
    std::vector<int> pq_off(ctx->n_path + 1, 0);
    for (int q = 0; q < n_query; q++)
        for (int j = q_off[q]; j < q_off[q + 1]; j++) {
            if (q_paths[j] < 0 || q_paths[j] >= ctx->n_path)
                FAIL(SVGR_E_INVALID, "cloud query: bad path index");
            pq_off[q_paths[j] + 1]++;
        }
    for (int i = 0; i < ctx->n_path; i++)
        pq_off[i + 1] += pq_off[i];
    std::vector<int> pq_idx(std::max(pq_off[ctx->n_path], 1)), cur(pq_off.begin(), pq_off.end() - 1);
    for (int q = 0; q < n_query; q++)
        for (int j = q_off[q]; j < q_off[q + 1]; j++)
            pq_idx[cur[q_paths[j]]++] = q;
    size_t b_off = pq_off.size() * 4, b_idx = pq_idx.size() * 4, b_inv = (size_t)n_query * 48, b_mm = (size_t)n_query * 32;
    size_t a_idx = (b_off + 15) & ~15ull, a_inv = (a_idx + b_idx + 15) & ~15ull, a_mm = a_inv + b_inv, a_out = a_mm + b_mm;
    CK(ctx->d_q.ensure(a_out + b_mm));
    char *d = (char *)ctx->d_q.p;
    cudaStream_t s = ctx->own_stream;
    CK(cudaMemcpyAsync(d, pq_off.data(), b_off, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d + a_idx, pq_idx.data(), b_idx, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d + a_inv, q_inv, b_inv, cudaMemcpyHostToDevice, s));
    svgr_launch_cloud_bounds(ctx->d_edges.as<double>(), ctx->d_edge_path.as<uint32_t>(),
                             &ctx->d_status.as<StatusBlock>()->n_edges, (unsigned long long)ctx->edge_cap, (int *)d,
                             (int *)(d + a_idx), (double *)(d + a_inv), (unsigned long long *)(d + a_mm),
                             (double *)(d + a_out), n_query, ctx->sm_count, s);
    CK(cudaMemcpyAsync(out, d + a_out, b_mm, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return SVGR_OK;
}

// ---- eager element-wise entry points: host in, one kernel, host out -------------------------
int svgr_line_signed_coverage(svgr_ctx *ctx, float *trace, int32_t rows, int32_t cols, const double *lines, int64_t n)
{
    if (!ctx)
        return SVGR_E_INVALID;
    if (!trace || rows < 0 || cols < 0 || n < 0 || (n > 0 && !lines))
        FAIL(SVGR_E_INVALID, "line_signed_coverage: bad arguments");
    if (rows == 0 || cols == 0 || n == 0)
        return SVGR_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->own_stream;
    const size_t b_trace = (size_t)rows * cols * 4, b_lines = (size_t)n * 32;
    CK(ctx->d_eager[0].ensure(b_trace));
    CK(ctx->d_eager[1].ensure(b_lines));
    CK(cudaMemcpyAsync(ctx->d_eager[0].p, trace, b_trace, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_eager[1].p, lines, b_lines, cudaMemcpyHostToDevice, s));
    svgr_launch_line_coverage(ctx->d_eager[1].as<double>(), n, ctx->d_eager[0].as<float>(), rows, cols, s);
    CK(cudaMemcpyAsync(trace, ctx->d_eager[0].p, b_trace, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return SVGR_OK;
}

int svgr_grad_pixels(svgr_ctx *ctx, int32_t r0, int32_t c0, int32_t rows, int32_t cols, double *out)
{
    if (!ctx)
        return SVGR_E_INVALID;
    if (rows < 0 || cols < 0 || (!out && rows > 0 && cols > 0))
        FAIL(SVGR_E_INVALID, "grad_pixels: bad arguments");
    const size_t bytes = (size_t)rows * cols * 16;
    if (bytes == 0)
        return SVGR_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->own_stream;
    CK(ctx->d_eager[0].ensure(bytes));
    svgr_launch_grad_pixels(r0, c0, rows, cols, ctx->d_eager[0].as<double>(), s);
    CK(cudaMemcpyAsync(out, ctx->d_eager[0].p, bytes, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return SVGR_OK;
}

int svgr_grad_spread(svgr_ctx *ctx, const double *offsets, int64_t n, int32_t spread, double *out)
{
    if (!ctx)
        return SVGR_E_INVALID;
    if (n < 0 || (n > 0 && (!offsets || !out)))
        FAIL(SVGR_E_INVALID, "grad_spread: bad arguments");
    if (spread < 0 || spread > 2)
        FAIL(SVGR_E_INVALID, "invalid spread method");
    if (n == 0)
        return SVGR_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->own_stream;
    CK(ctx->d_eager[0].ensure((size_t)n * 8));
    CK(ctx->d_eager[1].ensure((size_t)n * 8));
    CK(cudaMemcpyAsync(ctx->d_eager[0].p, offsets, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    svgr_launch_grad_spread(ctx->d_eager[0].as<double>(), n, spread, ctx->d_eager[1].as<double>(), s);
    CK(cudaMemcpyAsync(out, ctx->d_eager[1].p, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return SVGR_OK;
}

int svgr_grad_interpolate(svgr_ctx *ctx, const double *offsets, int64_t n, const struct StopRec *stops, int32_t n_stops,
                          float *out)
{
    if (!ctx)
        return SVGR_E_INVALID;
    if (n < 0 || (n > 0 && (!offsets || !out)))
        FAIL(SVGR_E_INVALID, "grad_interpolate: bad arguments");
    if (n_stops < 1 || !stops)
        FAIL(SVGR_E_INVALID, "gradient without stops");
    if (n == 0)
        return SVGR_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->own_stream;
    CK(ctx->d_eager[0].ensure((size_t)n * 8));
    CK(ctx->d_eager[1].ensure((size_t)n * 16));
    CK(ctx->d_eager[2].ensure((size_t)n_stops * sizeof(StopRec)));
    CK(cudaMemcpyAsync(ctx->d_eager[0].p, offsets, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_eager[2].p, stops, (size_t)n_stops * sizeof(StopRec), cudaMemcpyHostToDevice, s));
    svgr_launch_grad_interpolate(ctx->d_eager[0].as<double>(), n, ctx->d_eager[2].as<StopRec>(), n_stops,
                                 ctx->d_eager[1].as<float>(), s);
    CK(cudaMemcpyAsync(out, ctx->d_eager[1].p, (size_t)n * 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return SVGR_OK;
}

int svgr_quantize_u8(svgr_ctx *ctx, const float *values, int64_t n, uint8_t *out)
{
    if (!ctx)
        return SVGR_E_INVALID;
    if (n < 0 || (n > 0 && (!values || !out)))
        FAIL(SVGR_E_INVALID, "quantize: bad arguments");
    if (n == 0)
        return SVGR_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->own_stream;
    CK(ctx->d_eager[0].ensure((size_t)n * 4));
    CK(ctx->d_eager[1].ensure((size_t)n));
    CK(cudaMemcpyAsync(ctx->d_eager[0].p, values, (size_t)n * 4, cudaMemcpyHostToDevice, s));
    svgr_launch_quantize(ctx->d_eager[0].as<float>(), n, ctx->d_eager[1].as<uint8_t>(), s);
    CK(cudaMemcpyAsync(out, ctx->d_eager[1].p, (size_t)n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return SVGR_OK;
}

int svgr_pooling(svgr_ctx *ctx, const float *mat, int32_t rows, int32_t cols, int32_t ch, int32_t ky, int32_t kx,
                 int32_t sy, int32_t sx, int32_t method, int32_t pad, float *out, int32_t out_rows, int32_t out_cols)
{
    if (!ctx)
        return SVGR_E_INVALID;
    if (!mat || !out || rows < 0 || cols < 0 || ch < 1 || ky < 1 || kx < 1 || sy < 1 || sx < 1)
        FAIL(SVGR_E_INVALID, "pooling: bad arguments");
    if (method < 0 || method > 2)
        FAIL(SVGR_E_INVALID, "invalid poll method");
    // output extent of the reference (:435-455): padded -> ceil(n / s); cropped -> (n - k) / s + 1
    const int want_r = pad ? (rows + sy - 1) / sy : (rows >= ky ? (rows - ky) / sy + 1 : 0);
    const int want_c = pad ? (cols + sx - 1) / sx : (cols >= kx ? (cols - kx) / sx + 1 : 0);
    if (out_rows != want_r || out_cols != want_c)
        FAIL(SVGR_E_INVALID, "pooling: output extent does not match ksize / stride / pad");
    const size_t b_in = (size_t)rows * cols * ch * 4, b_out = (size_t)out_rows * out_cols * ch * 4;
    if (b_out == 0)
        return SVGR_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->own_stream;
    CK(ctx->d_eager[0].ensure(b_in));
    CK(ctx->d_eager[1].ensure(b_out));
    CK(cudaMemcpyAsync(ctx->d_eager[0].p, mat, b_in, cudaMemcpyHostToDevice, s));
    svgr_launch_pooling(ctx->d_eager[0].as<float>(), rows, cols, ch, ky, kx, sy, sx, method, ctx->d_eager[1].as<float>(),
                        out_rows, out_cols, s);
    CK(cudaMemcpyAsync(out, ctx->d_eager[1].p, b_out, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    return SVGR_OK;
}

int64_t svgr_arc_to_cubics(double cx, double cy, double rx, double ry, double phi, double eta, double eta_delta,
                           double *out, int64_t cap);

// Expands every SEG_ARC row of a segment list into its cubic pieces in one call (arc rows hold
// cx, cy, rx, ry, phi, eta, eta_delta).  new_index[i] = position of input segment i in the output
// (n + 1 entries, so sub-path offsets can be mapped).  Returns the output length, or a negative code.
int64_t svgr_expand_arcs(const uint8_t *tags, const double *data, int64_t n, uint8_t *out_tags, double *out_data,
                         int64_t cap, int64_t *new_index)
{
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        new_index[i] = m;
        const double *d = data + 8 * i;
        if (tags[i] == SEG_ARC) {
            if (m + 64 > cap)
                return SVGR_E_NOMEM;
            int64_t k = svgr_arc_to_cubics(d[0], d[1], d[2], d[3], d[4], d[5], d[6], out_data + 8 * m, 64);
            if (k < 0)
                return k;
            for (int64_t j = 0; j < k; j++)
                out_tags[m + j] = SEG_CUBIC;
            m += k;
        } else {
            if (m + 1 > cap)
                return SVGR_E_NOMEM;
            out_tags[m] = tags[i];
            memcpy(out_data + 8 * m, d, 64);
            m++;
        }
    }
    new_index[n] = m;
    return m;
}

int64_t svgr_arc_to_cubics(double cx, double cy, double rx, double ry, double phi, double eta, double eta_delta,
                           double *out, int64_t cap)
{
    // arc_to_bezier3 (svgrasterize.py:2355-2394): pieces of at most pi/4; np.linspace gives
    // eta_i = i*step + eta with the last one forced to eta + delta; 2x2 . vec = fma(M0, u, M1*v)
    const double c = cos(phi), sn = sin(phi);
    const double M[2][2] = {{c, -sn}, {sn, c}};
    const double max_angle = M_PI / 4;
    double cnt_f = ceil(fabs(eta_delta) / max_angle);
    if (!(cnt_f >= 0) || cnt_f > (double)cap)
        return SVGR_E_INVALID;
    int64_t count = (int64_t)cnt_f;
    double stop = eta + eta_delta;
    double step = count > 0 ? (stop - eta) / (double)count : 0.0;
    for (int64_t i = 0; i < count; i++) {
        double e1 = (double)i * step + eta;
        double e2 = (i + 1 == count) ? stop : (double)(i + 1) * step + eta;
        double t = tan((e2 - e1) / 2);
        double sq = sqrt(4 + 3 * (t * t));
        double alpha = sin(e2 - e1) * (sq - 1) / 3;
        double u, v, p0[2], p3[2], d1[2], d2[2];
        u = rx * cos(e1), v = ry * sin(e1);
        p0[0] = fma(M[0][0], u, M[0][1] * v) + cx, p0[1] = fma(M[1][0], u, M[1][1] * v) + cy;
        u = rx * cos(e2), v = ry * sin(e2);
        p3[0] = fma(M[0][0], u, M[0][1] * v) + cx, p3[1] = fma(M[1][0], u, M[1][1] * v) + cy;
        u = -rx * sin(e1), v = ry * cos(e1);
        d1[0] = fma(M[0][0], u, M[0][1] * v), d1[1] = fma(M[1][0], u, M[1][1] * v);
        u = -rx * sin(e2), v = ry * cos(e2);
        d2[0] = fma(M[0][0], u, M[0][1] * v), d2[1] = fma(M[1][0], u, M[1][1] * v);
        double *o = out + 8 * i;
        o[0] = p0[0], o[1] = p0[1];
        o[2] = p0[0] + alpha * d1[0], o[3] = p0[1] + alpha * d1[1];
        o[4] = p3[0] - alpha * d2[0], o[5] = p3[1] - alpha * d2[1];
        o[6] = p3[0], o[7] = p3[1];
    }
    return count;
}

#pragma GCC visibility pop
}  // extern "C"
