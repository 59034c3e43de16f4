/*
 * svgr_b200.h -- C ABI of the svgrasterize B200 core (libsvgr_b200.so).
 *
 * The reference (aslpavel/svgrasterize.py) has no FFI: its rasterizer is Python + numpy
 * called from Scene.render (svgrasterize.py:649-752) and main() (:3854-3881).  This
 * header is the boundary a maintainer binds with ctypes to replace that hot path:
 *
 *   reference call                                           replaced by
 *   -------------------------------------------------------  -----------------------------
 *   Path.mask / bezier3_flatten_batch (:922, :2091)          svgr_render(.., SVGR_STOP_FLATTEN) + svgr_read_edges / svgr_read_boxes
 *   line_signed_coverage + cumsum + fill rule (:2213, :983)  svgr_render(.., SVGR_STOP_COVERAGE) + svgr_read_mask / svgr_read_bins
 *   Path.stroke (:1105)                                      svgr_render(.., SVGR_STOP_STROKE) + svgr_read_outline
 *   Path.fill, Layer.compose/convert/opacity, canvas_* ,
 *   Filter.__call__, Layer.convolve/morphology/color_matrix,
 *   Scene.render (:995, :178, :277-416, :1801, :95-127, :649) svgr_render + svgr_node_info / svgr_read_node
 *   canvas_merge_at + write_png quantisation (:3870-3881,:263) svgr_render canvas nodes -> RGBA8
 *   ConvexHull.bbox (:2002)                                  svgr_cloud_bounds
 *   arc_to_bezier3 (:2355)  [host, libm]                     svgr_arc_to_cubics
 *   line_signed_coverage(canvas, line) (:2213)               svgr_line_signed_coverage
 *   grad_pixels / grad_spread / grad_interpolate (:1653-1683) svgr_grad_pixels / svgr_grad_spread / svgr_grad_interpolate
 *   canvas_to_png quantisation of a float image (:263)       svgr_quantize_u8
 *   pooling with stride / padding / mean (:419-468)          svgr_pooling
 *
 * Conventions: every function returns 0 on success, a negative SVGR_E_* code on
 * failure (svgr_last_error gives the text).  Nothing throws.  No torch types: plain
 * pointers, sizes and a cudaStream_t passed as void*.  All `program` arrays are HOST
 * pointers; svgr_render copies them to the device itself (the copies are part of an
 * end-to-end timing).  A context is bound to one device and is not thread-safe.
 */
#ifndef SVGR_B200_H
#define SVGR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVGR_VERSION 210

enum {
    SVGR_OK = 0,
    SVGR_E_INVALID = -1,  /* bad argument / malformed program: Python raises ValueError */
    SVGR_E_CUDA = -2,     /* CUDA runtime error: Python raises RuntimeError */
    SVGR_E_NOMEM = -3,
    SVGR_E_UNSUPPORTED = -4,
    SVGR_E_STROKE = -5, /* degenerate control polygon: the reference raises TypeError here (:2157) */
    SVGR_E_TYPE = -6,   /* another place where the reference raises TypeError: a pattern tile that misses its repeat
                           cell (canvas_merge_at returns None, :1093-1094) */
};

/* where svgr_render stops (taps for the stage-level parity tests and the eager API) */
enum {
    SVGR_STOP_NONE = 0,     /* full pipeline */
    SVGR_STOP_STROKE = 1,   /* stroke outlines only */
    SVGR_STOP_FLATTEN = 2,  /* + flatten, bounds */
    SVGR_STOP_COVERAGE = 3, /* + binning, coverage masks */
    SVGR_STOP_PLAN = 4,     /* flatten, bounds and the host plan only (node boxes; no pixels) */
};

/* scene-program node tags (the host encoder lowers the reference's Scene tree to these) */
enum {
    SVGR_N_EMPTY = 0,     /* renders nothing (reference returns None) */
    SVGR_N_LEAF = 1,      /* a = path, b = paint (-1: mask only), c = paint linear_rgb flag, d = pattern `pat` node or -1 */
    SVGR_N_GROUP = 2,     /* children over-composed (Layer.compose OVER); flags bit0 = linear_rgb */
    SVGR_N_OPACITY = 3,   /* child 0 x f[0]; flags bit0 = linear_rgb */
    SVGR_N_IN = 4,        /* compose([child0 (stencil), child1 (image)], IN); flags bit0 = linear_rgb */
    SVGR_N_LUMA = 5,      /* luminance x alpha of child 0 as one channel; flags bit0 = linear_rgb */
    SVGR_N_COMPOSE = 6,   /* Layer.compose(children, mode a, arithmetic k = f[0..3]); flags bit0 = linear_rgb,
                             bit2 = blend on the intersection of the boxes whatever the mode (canvas_merge_intersect
                             with a blend other than `in`), bit3 = children are raw arrays: no Layer.convert
                             (canvas_compose / canvas_merge_* on plain images) */
    SVGR_N_SRC_ALPHA = 7, /* (0,0,0,alpha) of child 0, premultiplied linear */
    SVGR_N_CONVERT = 8,   /* Layer.convert(pre_alpha = a, linear_rgb = b) of child 0 */
    SVGR_N_BLUR = 9,      /* Layer.convolve with kernel a of the kernel table */
    SVGR_N_MORPH = 10,    /* Layer.morphology: window a x b, c = 1 max / 0 min */
    SVGR_N_CMATRIX = 11,  /* Layer.color_matrix with matrix a */
    SVGR_N_OFFSET = 12,   /* feOffset: f[0..1] = (dx, dy), a = index into the offset-transform table */
    SVGR_N_MERGE_AT = 13, /* canvas_merge_at of child 0 onto zeros of bbox (a, b, c, d) */
    SVGR_N_CANVAS = 14,   /* final canvas: a = rows, b = cols, (c, d) = first row / column (row-band renders),
                             flags bit0 = linear_rgb, f[0] = byte offset in the output */
    SVGR_N_EXTERNAL = 15, /* a = index into the external layer table (host images handed in by the eager API) */
};

typedef struct svgr_node {
    int32_t tag;
    int32_t a, b, c, d;
    int32_t child_off, child_cnt;
    int32_t flags;
    double f[4];
} svgr_node;

/* Gaussian kernel of one feGaussianBlur under one transform (blur_kernel, :1903-1944) */
typedef struct svgr_kernel {
    int32_t rows, cols;  /* kernel extent along image rows / columns */
    int32_t separable;   /* 1: weights = row vector (rows) then column vector (cols); 0: rows x cols matrix */
    int32_t weight_off;  /* offset into the weight table */
} svgr_kernel;

/* layer handed in from the host (eager Layer API) */
typedef struct svgr_external {
    const float *image; /* rows x cols x channels, float32, C-contiguous */
    int32_t r0, c0, rows, cols;
    int32_t channels; /* 1 or 4 */
    int32_t pre_alpha, linear_rgb;
    int32_t pad;
} svgr_external;

/* A gradient paint in objectBoundingBox units (paint.bbox_units, svgrasterize.py:1023-1026): its pixel -> gradient
 * map depends on ConvexHull.bbox_transform (:2002-2023) of the leaf's own end points, which only exist once the path
 * is flattened.  The encoder leaves the map out of the PaintRec and records this job; the render completes the record
 * on the device (bounds of transform.invert(points) -> transform.translate(x, y).scale(w, h) -> inverse ->
 * coefficients) between flattening and compositing, without a host round trip. */
typedef struct svgr_bbox_job {
    int32_t paint;       /* PaintRec to complete (kind, stops, spread, the two-circle constants are already there) */
    int32_t path;        /* the leaf's path */
    int32_t has_grad_tr; /* 1: grad_inv holds the inverse gradientTransform */
    int32_t pad;
    double inv[6];       /* transform.invert of the leaf's transform, 2 x 3 row-major (numpy's) */
    double grad_inv[6];  /* inverse gradientTransform, 2 x 3 row-major */
    double geom[6];      /* linear: p0x p0y vx vy vv | radial: cx cy r | two-circle: fx fy */
} svgr_bbox_job;

/* The records below are declared in svgr_types.h (same layout on host and device). */
struct PathRec;
struct StrokeRec;
struct PaintRec;
struct StopRec;

typedef struct svgr_program {
    /* fill geometry: arcs already expanded to cubics (svgr_arc_to_cubics) */
    int64_t n_seg;
    const uint8_t *seg_tag;   /* n_seg */
    const double *seg_data;   /* n_seg x 8 */
    const uint32_t *seg_path; /* n_seg */
    /* one record per Path.mask call */
    int32_t n_path;
    const struct PathRec *paths;
    /* stroke jobs: user-space source segments, outlines are produced on the device */
    int32_t n_stroke;
    const struct StrokeRec *strokes;
    int32_t n_stroke_sub;
    const int32_t *stroke_sub_off; /* n_stroke_sub + 1, into the stroke segments */
    const int32_t *stroke_sub_job; /* n_stroke_sub */
    int64_t n_stroke_seg;
    const uint8_t *stroke_tag;
    const double *stroke_data;
    const int32_t *stroke_seg_job; /* n_stroke_seg */
    /* paint */
    int32_t n_paint;
    const struct PaintRec *paints;
    int32_t n_stop;
    const struct StopRec *stops;
    int32_t n_focal; /* number of two-circle gradient flags */
    /* scene program, children before parents */
    int32_t n_node;
    const svgr_node *nodes;
    int32_t n_child;
    const int32_t *children;
    /* filter tables */
    int32_t n_kernel;
    const svgr_kernel *kernels;
    int32_t n_weight;
    const float *weights;
    int32_t n_matrix;
    const float *matrices;   /* n_matrix x 20 (4x5 row-major) */
    int32_t n_offset_tr;
    const double *offset_tr; /* n_offset_tr x 12: forward 2x3 then inverse 2x3 */
    /* external layers */
    int32_t n_external;
    const svgr_external *externals;
    /* output */
    int64_t canvas_bytes; /* total RGBA8 bytes written by the canvas nodes */
    double flatness;      /* bezier3_flatten_batch(batch, flatness) (:2091); 0 = Path.mask's literal 0.1 (:955) */
    /* objectBoundingBox gradients completed on the device */
    int32_t n_bbox_job;
    int32_t pad_bbox;
    const svgr_bbox_job *bbox_jobs;
} svgr_program;

typedef struct svgr_stats {
    int64_t n_edges;
    int64_t n_outline_segs;
    int64_t n_bands, n_cov_tiles, n_binned;
    int64_t cov_floats, layer_floats; /* arena sizes */
    int64_t n_ops, n_levels, n_launches;
    int64_t mask_pixels, layer_pixels;
    int64_t coverage_bytes; /* algorithmic bytes of the coverage launch: 4 B per mask pixel + 36 B per binned edge */
    int64_t compose_bytes;  /* algorithmic bytes of the compose launches: every source pixel read once, every
                               output pixel written once */
    int64_t canvas_pixels;  /* pixels quantised by the canvas launch (16 B read + 4 B written each) */
    int64_t n_kernels;      /* kernel launches issued by this call */
    float ms_total, ms_h2d, ms_stroke, ms_flatten, ms_plan, ms_bin, ms_coverage, ms_compose, ms_canvas, ms_d2h;
    int32_t retries;
    int32_t plan_cached; /* 1: the render reused the plan (tables, tile lists, tensor maps) of the one before: same
                            program, same path boxes */
    float host_plan_masks_ms, host_plan_nodes_ms; /* wall time of the two host planning phases */
    float ms_compose_busy; /* sum over plan chunks of first-launch -> last-launch device time: ms_compose minus
                              the waits for the host planner */
    float pad2;
    int64_t compose_bytes_8d; /* the same compose work in SURVEY.md 8(d)'s units: 36 B per layer pixel composited
                                 (one HBM pass per layer: 4 B coverage + 16 B destination read + 16 B written) + 20 B
                                 per quantised canvas pixel; compose_bytes is less because the fold keeps the
                                 destination in registers */
    int64_t png_bytes;        /* svgr_render_png: total bytes of the PNG files */
    float ms_png;             /* device time of the PNG encoding kernels */
    float pad3;
} svgr_stats;

typedef struct svgr_ctx svgr_ctx;

int svgr_version(void);
/* sizeof of the ABI records, for binding self-checks: 0 PathRec, 1 StrokeRec, 2 PaintRec, 3 StopRec,
 * 4 svgr_node, 5 svgr_kernel, 6 svgr_external, 7 svgr_program, 8 svgr_stats, 9 MaskRec, 10 svgr_bbox_job */
int svgr_sizeof(int what);
int svgr_create(int device, svgr_ctx **out);
void svgr_destroy(svgr_ctx *ctx);
const char *svgr_last_error(svgr_ctx *ctx);

/* Render a program.  `out` receives canvas_bytes of RGBA8 (may be NULL when the program has no
 * canvas nodes or the caller only wants taps); out_on_device != 0 means `out` is a device pointer.
 * `stream` is a cudaStream_t (NULL = the context's own stream).  With timing != 0 the stage times
 * in `stats` are measured with CUDA events (adds synchronisation). */
int svgr_render(svgr_ctx *ctx, const svgr_program *prog, void *stream, int stop_after, uint8_t *out,
                int out_on_device, int timing, svgr_stats *stats);

/* Re-run the device part of the last svgr_render with the program already resident in HBM
 * (no host->device copy of the program); used for kernel-only throughput measurements. */
int svgr_render_resident(svgr_ctx *ctx, void *stream, uint8_t *out_device, int timing, svgr_stats *stats);

/* Render `prog` and encode every canvas as a PNG file ON THE DEVICE (8-bit RGBA, Paeth filter, one zlib stream of
 * dynamic-Huffman deflate blocks); only the files cross PCIe.  Replaces Layer.write_png / canvas_to_png
 * (svgrasterize.py:209-213, :249-274) for batches: the files decode to exactly the canvas bytes svgr_render
 * would have returned, but are not the reference's bytes (its filter-0 + zlib level 9 stream stays available
 * through the host path).  `out` receives the files back to back (host pointer, or device when out_on_device);
 * offsets (host, n_canvas + 1 entries, canvas node order) their byte ranges.  When out_cap is too small the call
 * fails with SVGR_E_NOMEM and offsets[n_canvas] holds the size needed. */
int svgr_render_png(svgr_ctx *ctx, const svgr_program *prog, void *stream, uint8_t *out, int64_t out_cap, int out_on_device,
                    int64_t *offsets, int timing, svgr_stats *stats);
/* The same on the program left resident by the last svgr_render / svgr_render_png. */
int svgr_render_resident_png(svgr_ctx *ctx, void *stream, uint8_t *out, int64_t out_cap, int out_on_device,
                             int64_t *offsets, int timing, svgr_stats *stats);
/* PNG-encode RGBA8 images that are already in host memory (rows x cols x 4 each, back to back) on the device:
 * the batch form of canvas_to_png's device path. */
int svgr_png_encode(svgr_ctx *ctx, const uint8_t *images, int32_t n_images, const int32_t *rows, const int32_t *cols,
                    uint8_t *out, int64_t out_cap, int64_t *offsets);

/* Host-only: plan `prog` for the given path boxes (n_path x 4 int32) `reps` times without touching CUDA;
 * reports the best wall time of the two planning phases and info = {ops, sources, launches, levels,
 * layer arena floats, compose bytes}.  For profiling and CPU tests of the planner. */
int svgr_debug_plan(const svgr_program *prog, const int32_t *boxes, int reps, float *ms_masks, float *ms_nodes,
                    int64_t *info);

/* ---- taps on the state left by the last svgr_render ------------------------------------- */
int svgr_read_edges(svgr_ctx *ctx, double *edges, uint32_t *edge_path, int64_t cap, int64_t *n_edges);
int svgr_read_boxes(svgr_ctx *ctx, int32_t *boxes /* n_path x 4 */, double *minmax /* n_path x 4 or NULL */);
int svgr_read_mask(svgr_ctx *ctx, int32_t path, float *out /* rows x cols */);
int svgr_read_bins(svgr_ctx *ctx, int32_t path, int32_t *band_off /* nbands + 1 */, uint32_t *bin_edges, int64_t cap,
                   int64_t *n_binned);
int svgr_read_outline(svgr_ctx *ctx, uint8_t *tag, double *data, uint32_t *path, int32_t *sub, int64_t cap,
                      int64_t *n_segs);
/* info[0] = kind (0 empty, 1 RGBA, 2 one channel), info[1..4] = r0, c0, rows, cols, info[5] = pre_alpha,
 * info[6] = linear_rgb */
int svgr_node_info(svgr_ctx *ctx, int32_t node, int32_t *info);
int svgr_read_node(svgr_ctx *ctx, int32_t node, float *out /* rows x cols x channels */);

/* ConvexHull.bbox(transform) for sets of paths of the last flattened program: query q covers the
 * paths listed in q_paths[q_off[q] .. q_off[q+1]) and maps end points by q_inv[q] (2x3 inverse
 * transform); out[q] = min x, min y, max x, max y in user space. */
int svgr_cloud_bounds(svgr_ctx *ctx, int32_t n_query, const int32_t *q_off, const int32_t *q_paths,
                      const double *q_inv, double *out);

/* arc_to_bezier3 (svgrasterize.py:2355-2394) on the host with libm, bit-identical to the reference.
 * Returns the number of cubics written (8 doubles each), or a negative code. */
int64_t svgr_arc_to_cubics(double cx, double cy, double rx, double ry, double phi, double eta, double eta_delta,
                           double *out, int64_t cap);

/* The same for a whole segment list: every SEG_ARC row (cx, cy, rx, ry, phi, eta, eta_delta) is replaced by
 * its cubic pieces, all other rows are copied.  new_index has n + 1 entries (input position -> output
 * position).  Returns the output length or a negative code (SVGR_E_NOMEM: cap too small). */
int64_t svgr_expand_arcs(const uint8_t *tags, const double *data, int64_t n, uint8_t *out_tags, double *out_data,
                         int64_t cap, int64_t *new_index);

/* ---- native scene encoder (csrc/encode_flat.cpp): the walk of Scene.render (svgrasterize.py:649-752) over flat
 * scene arrays -> a scene program.  The arrays are what svgrasterize_b200/_flatten.c reads out of the reference's own
 * Scene / Path / paint / Transform objects; many scenes share one set of tables (a batch). */
typedef struct svgr_flat_paint {
    int32_t kind;          /* 1 solid, 2 linear gradient, 3 radial gradient */
    int32_t spread;        /* 0 pad, 1 repeat, 2 reflect */
    int32_t bbox_units;    /* objectBoundingBox units: the encoder records a svgr_bbox_job, the render completes it */
    int32_t lin;           /* paint.linear_rgb: -1 None, 0, 1 */
    int32_t has_transform; /* 1: inv holds the inverse gradientTransform */
    int32_t stop_off, stop_cnt;
    int32_t focal;         /* radial: bit0 fcenter given, bit1 fradius given */
    double p[8];           /* solid: r g b a | linear: p0x p0y p1x p1y | radial: cx cy r fx fy fr */
    double inv[6];         /* inverse gradientTransform, 2 x 3 row-major (numpy's: Transform.invert) */
} svgr_flat_paint;
typedef struct svgr_flat_stop {
    double offset;
    double color[4]; /* premultiplied linear RGBA */
} svgr_flat_stop;
typedef struct svgr_flat_node {
    int32_t tag; /* RENDER_* (svgrasterize.py:576-583): 0 fill (a path, b paint or -1, c rule 0 nonzero / 1 evenodd),
                    1 stroke (a path, b paint, c cap 0 butt 1 round 2 square, d join 0 miter 1 round 2 bevel 3 other,
                    f[0] width), 2 group, 3 opacity (f[0]), 4 clip / 5 mask (children target, other; a = bbox units),
                    6 transform (a = index into tr), 7 filter (child = the target, a = first primitive in fes,
                    b = number of primitives) */
    int32_t a, b, c, d;
    int32_t child_off, child_cnt;
    int32_t pad;
    double f[2];
} svgr_flat_node;
/* one filter primitive (svgrasterize.py:1733-1799, Filter.__call__ :1801-1831) */
typedef struct svgr_flat_fe {
    int32_t tag;    /* FE_*: 0 blend, 1 colour matrix, 3 composite, 8 gaussian blur, 9 merge, 10 morphology, 11 offset */
    int32_t n_in;   /* inputs: slots of the filter's stack (0 SourceAlpha, 1 SourceGraphic, i + 2 primitive i) */
    int32_t in_off; /* into fe_inputs */
    int32_t flag;   /* composite: 1 arithmetic (a = k1..k4), 0 a[0] = mode, -1 invalid mode | colour matrix: 1 a = the
                       4 x 5 matrix, 0 not a 4 x 5 array (the primitive is skipped) | morphology: 1 max, 0 min, -1 other */
    double a[20];   /* offset: dx dy | blur: std_x std_y (std_y = std_x when None) | morphology: rx ry */
} svgr_flat_fe;
typedef struct svgr_flat_scene {
    int32_t root;          /* node index */
    int32_t width, height; /* canvas size: main() renders with viewport [0, 0, height, width] (:3857-3861) */
    int32_t linear_rgb;
} svgr_flat_scene;
typedef struct svgr_flat {
    int64_t n_seg;
    const uint8_t *seg_tag;  /* PATH_* tags; arcs are expanded by the encoder */
    const double *seg_data;  /* n_seg x 8 */
    int32_t n_sub;
    const int32_t *sub_off;  /* n_sub + 1: sub-path -> segments (empty sub-paths already dropped) */
    int32_t n_path;
    const int32_t *path_off; /* n_path + 1: path -> sub-paths */
    int32_t n_tr;
    const double *tr;        /* n_tr x 6: 2 x 3 row-major matrices of the transform nodes */
    int32_t n_paint;
    const svgr_flat_paint *paints;
    int32_t n_stop;
    const svgr_flat_stop *stops;
    int32_t n_node;
    const svgr_flat_node *nodes;
    int32_t n_child;
    const int32_t *children;
    int32_t n_scene;
    const svgr_flat_scene *scenes;
    int32_t n_fe;
    const svgr_flat_fe *fes;
    int32_t n_fe_input;
    const int32_t *fe_inputs;
} svgr_flat;
typedef struct svgr_encoded svgr_encoded;
/* Encodes every scene of `in` (in order, one canvas each) into one program owned by the returned handle (also
 * returned on failure, for svgr_encoded_error; free it with svgr_encoded_free).  SVGR_E_UNSUPPORTED: a scene uses
 * a feature this encoder leaves to the Python one. */
int svgr_encode_flat(const svgr_flat *in, svgr_encoded **out);
const svgr_program *svgr_encoded_program(const svgr_encoded *e);
const char *svgr_encoded_error(const svgr_encoded *e);
/* canvases: 4 int64 per scene (canvas node, byte offset in the output, rows, cols); roots: the scenes' root nodes;
 * returns the number of scenes */
int64_t svgr_encoded_canvases(const svgr_encoded *e, const int64_t **canvases, const int32_t **roots);
void svgr_encoded_free(svgr_encoded *e);

/* ---- path-data reader (csrc/pathdata.cpp).  Path.from_svg (svgrasterize.py:1252-1430): tokenises a `d` attribute
 * or glyph outline and builds the sub-paths, straight into flat segment arrays (PATH_* tag + 8 doubles per segment:
 * the points of a line / quad / cubic, or cx cy rx ry phi eta eta_delta of an arc -- arc_svg_to_parametric, :2397).
 * n_seg / n_sub receive the counts; SVGR_E_NOMEM when they exceed cap_seg / cap_sub (call again), SVGR_E_INVALID
 * where the reference raises ValueError (message in err). */
int svgr_path_from_svg(const char *d, int64_t len, uint8_t *tags, double *data, int64_t cap_seg, int32_t *sub_off,
                       int64_t cap_sub, int64_t *n_seg, int64_t *n_sub, char *err, int32_t err_cap);

/* ---- eager element-wise entry points of the reference's call surface.  Host pointers in and
 * out; every call copies up, launches one kernel on the context's stream and copies down (one synchronisation). */

/* line_signed_coverage(canvas, line) (svgrasterize.py:2213-2304) for n lines (r0, c0, r1, c1 each) accumulated
 * into `trace` (rows x cols float32, updated in place): the signed-area deltas before cumsum and fill rule. */
int svgr_line_signed_coverage(svgr_ctx *ctx, float *trace, int32_t rows, int32_t cols, const double *lines, int64_t n);
/* grad_pixels(viewport) (:1653-1658): out[(i * cols + j) * 2 + {0, 1}] = (r0 + i + 0.5, c0 + j + 0.5) */
int svgr_grad_pixels(svgr_ctx *ctx, int32_t r0, int32_t c0, int32_t rows, int32_t cols, double *out);
/* grad_spread(offsets, spread) (:1661-1668) on n float64 offsets; spread 0 pad, 1 repeat, 2 reflect */
int svgr_grad_spread(svgr_ctx *ctx, const double *offsets, int64_t n, int32_t spread, double *out);
/* grad_interpolate(offset, stops, linear_rgb) (:1671-1683): n offsets -> n x 4 float32 premultiplied RGBA; the
 * stops arrive already converted to the target colour space (grad_stops_colorspace, :1686) */
int svgr_grad_interpolate(svgr_ctx *ctx, const double *offsets, int64_t n, const struct StopRec *stops, int32_t n_stops,
                          float *out);
/* np.round(canvas * 255).astype(uint8) of canvas_to_png (:263): n float32 values -> n bytes (round half even) */
int svgr_quantize_u8(svgr_ctx *ctx, const float *values, int64_t n, uint8_t *out);
/* pooling(mat, ksize, stride, method, pad) (:419-468) on a rows x cols x ch float32 array: method 0 max, 1 min,
 * 2 mean (nan-ignoring like np.nanmax / nanmin / nanmean); out has out_rows x out_cols x ch elements with
 * out = (n - k) / s + 1 without padding, ceil(n / s) with it. */
int svgr_pooling(svgr_ctx *ctx, const float *mat, int32_t rows, int32_t cols, int32_t ch, int32_t ky, int32_t kx,
                 int32_t sy, int32_t sx, int32_t method, int32_t pad, float *out, int32_t out_rows, int32_t out_cols);

#ifdef __cplusplus
}
#endif
#endif
