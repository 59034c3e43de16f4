// svgr_types.h -- plain-old-data records shared by the host planner and the
// sm_100a kernels of the svgrasterize B200 core.  Everything that crosses the
// host/device boundary is one of these structs or a flat array of scalars.
//
// Coordinates follow the reference (svgrasterize.py:246, :3823): after the
// canvas transform component 0 of a point is the image ROW, component 1 the
// COLUMN.  A layer bbox is (r0, c0, rows, cols) (svgrasterize.py:88-89).
#pragma once
#include <stdint.h>

// ---- path segments (svgrasterize.py:865-873) -------------------------------
enum : uint8_t {
    SEG_LINE = 0,
    SEG_QUAD = 1,
    SEG_CUBIC = 2,
    SEG_ARC = 3,  // host side only: arcs are expanded to cubics before upload
    SEG_CLOSED = 4,
    SEG_UNCLOSED = 5,
    SEG_NOP = 255,  // padding written by the device stroker
};

// One entry per Path.mask call of the reference (svgrasterize.py:922).
struct PathRec {
    double m[6];          // presentation transform, row-major 2x3
    int32_t viewport[4];  // (r0, c0, rows, cols) clip, svgrasterize.py:968-971
    int32_t has_viewport;
    int32_t fill_rule;  // 0 nonzero (also None), 1 evenodd
    // Row-band renders clip masks to the band (+ halo) but place filter results where the full-canvas render would
    // (Layer.convolve / filter_offset truncate toward zero from the layer's origin, :114 / :1849): has_full = 1 adds
    // the viewport of the whole canvas, against which a second box is computed.
    int32_t has_full;
    int32_t pad;
    int32_t full_viewport[4];
};

// Per-path result of flatten + bounds (svgrasterize.py:961-975).
struct PathBox {
    int32_t r0, c0, rows, cols;  // rows <= 0 or cols <= 0: empty (reference returns None)
};

// Stroke job: one Path.stroke call (svgrasterize.py:1105).  Its input segments
// live in the stroke segment buffer [sub_off[sub_begin] .. sub_off[sub_end]).
struct StrokeRec {
    double half_width;
    int32_t sub_begin, sub_end;  // range in the stroke sub-path table
    int32_t cap;                 // 0 butt, 1 round, 2 square
    int32_t join;                // 0 miter, 1 round, 2 bevel, 3 other
    int32_t path;                // destination path id (outline segments get this id)
    int32_t pad;
};

// ---- coverage ----------------------------------------------------------------
#define SVGR_BAND_ROWS 16    // rows per coverage band (bin granularity)
#define SVGR_TILE_COLS 256   // columns per coverage tile

// Per path, written by the planner: where its mask lives and its bands.
struct MaskRec {
    int32_t r0, c0, rows, cols;
    int32_t stride;     // floats per row (cols rounded up to 4)
    int32_t band_base;  // first global band id
    int32_t tile_base;  // first global coverage tile id
    int32_t ntile_c;    // column tiles per band
    int64_t off;        // float offset into the coverage arena
    int32_t fill_rule;
    int32_t pad;
};

// One coverage tile (band of SVGR_BAND_ROWS rows x chunk of SVGR_TILE_COLS columns of one mask), written on the
// device after binning so that a coverage CTA starts from a single 48-byte load instead of chasing
// tile -> mask -> band -> bin through four dependent ones.
struct TileRec {
    int32_t r0, c0, rows, cols;  // of the mask
    int32_t stride, fill_rule;
    int32_t yb, col0;            // first row / column of the tile inside the mask
    int64_t off;                 // float offset of the mask in the coverage arena
    int32_t e_off, e_cnt;        // the band's slice of the bin arrays
};
static_assert(sizeof(TileRec) == 48, "TileRec is loaded as three 16-byte pieces");

// ---- paint (svgrasterize.py:995-1103, :1544-1695) ---------------------------------
enum : int32_t {
    PAINT_SOLID = 0,
    PAINT_LINEAR = 1,
    PAINT_RADIAL = 2,        // simple: t = |p - c| / r
    PAINT_RADIAL_FOCAL = 3,  // two-circle
    PAINT_PATTERN = 4,
};

struct PaintRec {
    int32_t kind;
    int32_t spread;  // 0 pad, 1 repeat, 2 reflect
    int32_t stop_off, stop_cnt;
    int32_t has_m2;
    int32_t flag;  // PAINT_RADIAL_FOCAL: index into the any-negative-det flag array
    int32_t pat_r0, pat_c0, pat_rows, pat_cols;  // PAINT_PATTERN: bbox of the `pat` image (lo .. hi)
    int32_t pat_node;                            // PAINT_PATTERN: node holding the pat image
    int32_t pad;
    float color[4];  // PAINT_SOLID: premultiplied colour in the layer's colour space
    // Gradients: the affine maps pixel centre -> user space -> gradient space are composed on the host, see
    // paint_eval in svgr_device.cuh for the meaning of m1 / g per kind.  Pattern: m1 = inverse `rep` matrix,
    // m2 = forward `rep` matrix, g = x y w h of the tile (svgrasterize.py:1074-1094).
    double m1[6];
    double m2[6];
    double g[8];
};

struct StopRec {
    double offset;
    float color[4];   // premultiplied, already in the target colour space
    double inv_span;  // 1 / (next offset - offset): the interpolation ratio is a multiply on the device
};

// ---- layers and ops ----------------------------------------------------------------
enum : int32_t {
    SRC_NONE = 0,
    SRC_L4 = 1,        // RGBA f32 layer
    SRC_L1 = 2,        // one-channel f32 layer (alpha is the image, svgrasterize.py:283-284)
    SRC_COV = 3,       // coverage mask used as a one-channel layer (mask_only leaves)
    SRC_COVPAINT = 4,  // coverage x paint: a Path.fill layer that is never materialised
    // stencil modifiers: an entry of these kinds multiplies the source entry before it by a one-channel
    // value (compose IN of a clip / mask that is never materialised, svgrasterize.py:698-741)
    SRC_MOD_COV = 8,    // coverage mask (clip path rendered mask-only)
    SRC_MOD_L1 = 9,     // one-channel layer
    SRC_MOD_L4A = 10,   // alpha of an RGBA layer
    SRC_MOD_LUMA = 11,  // luminance x alpha of an RGBA layer after Layer.convert (conv)
};

// conversion codes: bit0 source pre_alpha, bit1 source linear, bit2 target pre, bit3 target linear
#define SVGR_CONV(sp, sl, tp, tl) ((int)(sp) | ((int)(sl) << 1) | ((int)(tp) << 2) | ((int)(tl) << 3))

struct SrcRec {
    int32_t kind;
    int32_t r0, c0, rows, cols;  // valid region (global coordinates)
    int32_t stride;  // pixels per row
    int32_t paint;   // SRC_COVPAINT
    int32_t conv;    // SVGR_CONV code; identity when source == target flags
    float mul;       // folded opacity
    int32_t stride2; // pattern paint: pixels per row of the `pat` image
    int64_t off;     // float offset into the arena (coverage arena for SRC_COV*)
    int64_t off2;    // pattern paint: float offset of the `pat` image in the layer arena
    int32_t br0, bc0;  // global coordinates of element 0 of the storage (>= valid region origin for clipped views)
};  // 64 bytes: staged into shared memory as four 16-byte pieces
static_assert(sizeof(SrcRec) == 64, "SrcRec must be 64 bytes");

enum : int32_t {
    MODE_OVER = 0,
    MODE_OUT = 1,
    MODE_IN = 2,
    MODE_ATOP = 3,
    MODE_XOR = 4,
    MODE_ARITH = 5,
};

enum : int32_t {
    POST_NONE = 0,
    POST_CLIP01 = 1,     // canvas_merge_at clips to [0, 1] (svgrasterize.py:326)
    POST_ALPHA = 2,      // SourceAlpha: (0, 0, 0, a) (svgrasterize.py:1803-1806)
    POST_LUMA = 4,       // luminance mask: dot(rgb, LUMA) * a -> one channel (svgrasterize.py:734-736)
    POST_MATRIX = 8,     // feColorMatrix: M v + b, clip (svgrasterize.py:95-104)
};

enum : int32_t {
    OP_COMPOSE = 0,    // fold sources with a Porter-Duff mode
    OP_STENCIL_H = 1,  // separable stencil along columns (axis 1)
    OP_STENCIL_V = 2,  // separable stencil along rows (axis 0)
    OP_CONV2D = 3,     // direct 2-D convolution (non-separable blur)
    OP_CANVAS = 4,     // merge_at onto the canvas + straight sRGB + RGBA8 quantise
};

enum : int32_t { STENCIL_CONV = 0, STENCIL_MAX = 1, STENCIL_MIN = 2 };

struct OpRec {
    int32_t kind;
    int32_t mode;
    int32_t r0, c0, rows, cols;  // output bbox
    int32_t stride;              // output pixels per row
    int32_t out_ch;              // 1 or 4
    int64_t out_off;             // float offset into the layer arena (byte offset for OP_CANVAS)
    int32_t src_off, src_cnt;
    int32_t post;
    int32_t aux;     // POST_MATRIX: matrix index | stencil: weight offset | CONV2D: weight offset
    int32_t k0, k1;  // stencil length | CONV2D kernel rows, cols
    int32_t stencil; // STENCIL_*
    int32_t tile_base;  // first tile id of this op inside its launch
    int32_t ntile_c;    // tiles per tile-row
    int32_t list_base;  // compose ops: first TileEntry slot of this op's tiles (src_cnt slots per tile)
    float mul;    // opacity (applied after the fold)
    float k[4];   // arithmetic coefficients
    float pad2;
};

// Per-tile source list of a compose launch (cull_kernel -> compose_kernel): the sources of the tile's op that
// touch the tile, in order, with everything about the pair (source, tile) that does not depend on the pixel.
struct TileEntry {
    int32_t src;     // index into the SrcRec table
    int32_t paint;   // PaintRec index of a SRC_COVPAINT source, -1 none, -2 the entry is a stencil modifier
    uint32_t rm, cm; // bit r / c: tile row r / column c lies inside the source's valid region (and the op)
    int32_t toff;    // element offset of the tile's top-left pixel in the source's storage
    int32_t pad[3];
};
static_assert(sizeof(TileEntry) == 32, "TileEntry is read as two 16-byte pieces");
struct TileHead {
    int32_t op;      // index of the tile's op inside the launch
    int32_t n;       // entries
    int32_t first;   // first entry in the launch's list
    int32_t pad;
};

// any(det < 0) pre-pass of a two-circle gradient fill (svgrasterize.py:1621-1622)
struct FocalJob {
    int32_t paint;
    int32_t r0, c0, rows, cols;
    int32_t block_base;  // 1024 pixels per block
};

// tile shapes (rows x cols of output pixels per CTA)
#ifndef SVGR_CMP_TR
#define SVGR_CMP_TR 32
#endif
#define SVGR_CMP_TC 32
#define SVGR_C2D_TR 8
#define SVGR_C2D_TC 32
#define SVGR_STH_TR 8
#define SVGR_STH_TC 128
#define SVGR_STV_TR 64
#define SVGR_STV_TC 32
// TMA-fed stencils (k_stencil_tma.cu): tile shapes and taps per shared-memory stage
#define SVGR_TMA_H_TR 8
#define SVGR_TMA_H_TC 224
#define SVGR_TMA_H_KC 289
#define SVGR_TMA_V_TR 64
#define SVGR_TMA_V_TC 32
#define SVGR_TMA_V_KC 65
