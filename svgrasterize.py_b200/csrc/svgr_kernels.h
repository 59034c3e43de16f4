// svgr_kernels.h -- internal: launcher prototypes shared by the kernel files and the host engine.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "svgr_types.h"

#define SVGR_MAX_DYN_SMEM (200 * 1024)

struct RenderTables {
    const SrcRec *srcs;
    const PaintRec *paints;
    const StopRec *stops;
    const int *focal_flags;  // any(det < 0) per focal radial paint (svgrasterize.py:1622)
    const float *cov;        // coverage arena
    const float *layers;     // layer arena (read side)
    const float *matrices;   // 20 floats per feColorMatrix
    const float *weights;    // stencil / convolution weights
};

// k_flatten.cu
void svgr_launch_minmax_init(unsigned long long *minmax, int n, cudaStream_t s);
// overflow lists of the multi-pass flatten (k_flatten.cu): two buffers used alternately, one counter per pass
#define SVGR_FLAT_PASS_DEPTH 5
#define SVGR_FLAT_PASSES 5  // depth limits 5, 10, 15, 20, then the depth cap
struct svgr_flat_overflow {
    double *cubic[2];
    uint32_t *path[2];
    uint8_t *depth[2];
    int *counts;  // SVGR_FLAT_PASSES ints, zeroed before the first pass
    int cap;
};
void svgr_launch_flatten(const uint8_t *seg_tag, const double *seg_data, const uint32_t *seg_path, long long n_seg,
                         const int *n_seg_dev, const PathRec *paths, double thr, double *edges, uint32_t *edge_path,
                         unsigned long long cap, unsigned long long *n_edges, unsigned long long *minmax, int sm_count,
                         const svgr_flat_overflow *ovf, cudaStream_t s);
void svgr_launch_flatten_overflow(const PathRec *paths, double thr, double *edges, uint32_t *edge_path,
                                  unsigned long long cap, unsigned long long *n_edges, unsigned long long *minmax,
                                  int sm_count, const svgr_flat_overflow *ovf, cudaStream_t s);
void svgr_launch_bounds(const unsigned long long *minmax, const PathRec *paths, int n_paths, PathBox *boxes,
                        PathBox *full_boxes, double *minmax_f64, cudaStream_t s);
struct svgr_bbox_job;
void svgr_launch_resolve_paints(const double *edges, const uint32_t *edge_path, const unsigned long long *n_edges,
                                unsigned long long cap, const int *pq_off, const int *pq_idx, const double *q_inv,
                                unsigned long long *keys, const svgr_bbox_job *jobs, int n_jobs, const PathRec *paths,
                                PaintRec *paints, int *err, int sm_count, cudaStream_t s);
void svgr_launch_cloud_bounds(const double *edges, const uint32_t *edge_path, const unsigned long long *n_edges,
                              unsigned long long cap, const int *pq_off, const int *pq_idx, const double *q_inv,
                              unsigned long long *q_minmax, double *q_out, int n_q, int sm_count, cudaStream_t s);

// k_coverage.cu
void svgr_launch_exclusive_scan(const int *in, int *out, long long n, int *tmp, int *total_out, cudaStream_t s);
void svgr_launch_bin_count(const double *edges, const uint32_t *edge_path, unsigned long long n_edges,
                           const MaskRec *masks, int *band_count, int sm_count, cudaStream_t s);
void svgr_launch_bin_fill(const double *edges, const uint32_t *edge_path, unsigned long long n_edges,
                          const MaskRec *masks, const int *band_off, int *band_cursor, uint32_t *bin_edges,
                          double *bin_data,
                          long long cap, int sm_count, cudaStream_t s);
void svgr_launch_coverage(const MaskRec *masks, int n_masks, int n_tiles, TileRec *tiles, const int *band_off,
                          const int *band_cnt, const double *bin_data, long long bin_cap, float *cov, cudaStream_t s);

void svgr_launch_line_coverage(const double *lines, long long n, float *trace, int rows, int cols, cudaStream_t s);

// k_compose.cu
void svgr_launch_grad_pixels(int r0, int c0, int rows, int cols, double *out, cudaStream_t s);
void svgr_launch_grad_spread(const double *in, long long n, int spread, double *out, cudaStream_t s);
void svgr_launch_grad_interpolate(const double *in, long long n, const StopRec *stops, int n_stops, float *out,
                                  cudaStream_t s);
void svgr_launch_quantize(const float *in, long long n, uint8_t *out, cudaStream_t s);
void svgr_launch_expand_ops(const OpRec *ops, int n_ops, int n_tiles, int *tile_op, cudaStream_t s);
void svgr_launch_cull(const RenderTables &T, const OpRec *ops, int n_ops, int n_tiles, TileHead *heads, TileEntry *list,
                      cudaStream_t s);
void svgr_launch_compose(const RenderTables &T, const OpRec *ops, const TileHead *heads, const TileEntry *list, int n_tiles,
                         bool simple, float *layers_out, uint8_t *canvas_out, cudaStream_t s);
void svgr_launch_focal_flags(const RenderTables &T, const void *jobs, int n_jobs, int block0, int n_blocks, int *flags,
                             cudaStream_t s);

// k_filters.cu
int svgr_launch_stencil(const RenderTables &T, const OpRec *ops, const int *tile_op, int n_tiles, size_t smem_bytes,
                        float *layers_out, cudaStream_t s);
int svgr_launch_conv2d(const RenderTables &T, const OpRec *ops, const int *tile_op, int n_tiles, size_t smem_bytes,
                       float *layers_out, cudaStream_t s);

void svgr_launch_pooling(const float *in, int rows, int cols, int ch, int ky, int kx, int sy, int sx, int method,
                         float *out, int orows, int ocols, cudaStream_t s);

// k_stencil_tma.cu
int svgr_encode_layer_map(void *out, const float *base, long long rows, long long cols, long long stride_px, int box_px,
                          int box_rows);
size_t svgr_stencil_tma_smem(bool horiz, int k_max, int *kc_out);
int svgr_launch_stencil_tma(const RenderTables &T, const OpRec *ops, const int *tile_op, int n_tiles, const void *tmaps,
                            bool horiz, int k_max, int sm_count, float *layers_out, cudaStream_t s);

// k_png.cu
struct PngSeg;
struct PngCanvas;
long long svgr_png_slot_bytes(int rows, int cols);
int svgr_png_rows_per_segment(int cols);
size_t svgr_png_code_scratch_bytes();
int svgr_launch_png_deflate(const void *segs, int n_seg, const void *canvases, const unsigned char *canvas_buf,
                            unsigned char *scratch, int *seg_bytes, unsigned *seg_adler, void *code_scratch, cudaStream_t s);
void svgr_launch_png_sizes(const void *canvases, int n_canvas, const int *seg_bytes, int *file_bytes, cudaStream_t s);
void svgr_launch_png_pack(const void *canvases, int n_canvas, const void *segs, const int *seg_bytes, const unsigned *seg_adler,
                          const unsigned char *scratch, const long long *file_off, unsigned char *out, cudaStream_t s);

// k_stroke.cu
size_t svgr_stroke_curve_bytes();
void svgr_launch_stroke_count(const uint8_t *tag, const double *data, const int *seg_job, const StrokeRec *jobs,
                              int n_seg, int *counts, int *err, cudaStream_t s);
void svgr_launch_stroke_emit(const uint8_t *tag, const double *data, const int *seg_job, const StrokeRec *jobs,
                             int n_seg, const int *offs, void *pool, int pool_cap, cudaStream_t s);
void svgr_launch_stroke_bound(const int *sub_off, int n_sub, int n_seg, const int *offs, int *bound, cudaStream_t s);
void svgr_launch_stroke_assemble(const uint8_t *in_tag, const int *sub_off, const int *sub_job, const StrokeRec *jobs,
                                 int n_sub, int n_seg, const int *items, int n_items, const int *offs, const void *pool,
                                 const int *bound, const int *out_off, const int *out_total, long long out_base,
                                 long long out_cap, uint8_t *out_tag, double *out_data, uint32_t *out_path,
                                 int32_t *out_sub, int *n_out_dev, int *err, cudaStream_t s);
