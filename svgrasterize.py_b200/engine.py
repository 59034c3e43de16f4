"""Engine: one svgr_ctx (one GPU, one stream) behind a small Python class.

PyTorch is used only for what it is good at here: picking the device, handing
over the current CUDA stream and owning output tensors that stay on the GPU.
Every pixel is produced by libsvgr_b200.so; if the library or the GPU is missing
the constructor raises (there is no CPU fallback).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .encode import Program

_ERRORS = {
    _lib.E_INVALID: ValueError,
    _lib.E_CUDA: RuntimeError,
    _lib.E_NOMEM: MemoryError,
    _lib.E_UNSUPPORTED: NotImplementedError,
    _lib.E_STROKE: TypeError,
    _lib.E_TYPE: TypeError,
}


class Engine:
    def __init__(self, device: int = 0):
        self.L = _lib.lib()
        ctx = C.c_void_p()
        rc = self.L.svgr_create(int(device), C.byref(ctx))
        if rc != 0 or not ctx.value:
            raise RuntimeError(
                f"svgr_create(device={device}) failed ({rc}): the svgrasterize B200 core needs a CUDA device "
                "(no CPU fallback)")
        self.ctx = ctx
        self.device = int(device)
        self.program = None
        self.last_stats = None
        self._keep = None

    def close(self):
        if getattr(self, "ctx", None):
            self.L.svgr_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self.L.svgr_last_error(self.ctx)
            msg = msg.decode() if msg else f"error {rc}"
            raise _ERRORS.get(rc, RuntimeError)(msg)

    @staticmethod
    def _stream(stream):
        """None -> the context's own (non-blocking) stream.  A torch stream or a raw handle -> that stream; torch's
        default stream has handle 0, which the C-ABI would read as "own stream", so it is passed as
        cudaStreamLegacy (0x1): the work is then ordered with everything else the caller has on that stream."""
        if stream is None:
            return None
        handle = stream if isinstance(stream, int) else stream.cuda_stream  # torch.cuda.Stream
        return C.c_void_p(handle if handle else 1)

    # -- rendering -----------------------------------------------------------------------------
    def render(self, program: Program, stop: int = _lib.STOP_NONE, out=None, timing: bool = False, stream=None):
        """Run the pipeline on `program` (host arrays are copied to the device inside the call).

        out: None | numpy uint8 array (host, canvas_bytes long) | torch CUDA uint8 tensor.  Returns the
        stats dict; with out=None and canvases present the RGBA8 bytes are returned as `stats["canvas"]`."""
        cprog, keep = program.to_c()
        self.program, self._keep = program, keep
        stats = _lib.Stats()
        out_ptr, on_dev, host_out = None, 0, None
        if stop == _lib.STOP_NONE and program.canvas_bytes > 0:
            if out is None:
                host_out = np.empty(program.canvas_bytes, dtype=np.uint8)
                out_ptr = host_out.ctypes.data
            elif isinstance(out, np.ndarray):
                if out.dtype != np.uint8 or not out.flags.c_contiguous or out.nbytes < program.canvas_bytes:
                    raise ValueError(f"out must be a C-contiguous uint8 array of >= {program.canvas_bytes} bytes")
                out_ptr = out.ctypes.data
            else:  # torch tensor
                self._check_device_tensor(out, program.canvas_bytes)
                out_ptr, on_dev = out.data_ptr(), 1
        rc = self.L.svgr_render(self.ctx, C.byref(cprog), self._stream(stream), int(stop), out_ptr, on_dev,
                                int(bool(timing)), C.byref(stats))
        self._check(rc)
        res = stats.as_dict()
        if host_out is not None:
            res["canvas"] = host_out
        self.last_stats = res
        return res

    def _check_device_tensor(self, t, nbytes):
        if not t.is_cuda or t.device.index != self.device:
            raise ValueError(f"out must live on cuda:{self.device} (the engine's device), not {t.device}")
        if not t.is_contiguous() or t.numel() * t.element_size() < nbytes:
            raise ValueError(f"out must be contiguous and hold >= {nbytes} bytes")

    def render_resident(self, out_device=None, timing: bool = False, stream=None):
        """Re-run the device pipeline on the program left resident by the last render()."""
        if self.program is None:
            raise ValueError("no program resident: call render() first")
        stats = _lib.Stats()
        if out_device is not None:
            self._check_device_tensor(out_device, self.program.canvas_bytes)
        ptr = None if out_device is None else out_device.data_ptr()
        self._check(self.L.svgr_render_resident(self.ctx, self._stream(stream), ptr, int(bool(timing)), C.byref(stats)))
        return stats.as_dict()

    # -- PNG files made on the device ------------------------------------------------------------------
    def _png_call(self, fn, head, n_canvas, raw_bytes, out, timing):
        """Shared tail of render_png / render_resident_png: out = None (a host buffer is allocated and trimmed),
        a numpy uint8 array (host, e.g. pinned) or a torch CUDA uint8 tensor."""
        stats = _lib.Stats()
        offsets = np.zeros(n_canvas + 1, dtype=np.int64)
        own = None
        if out is None:
            # PNG of flat vector art is a small fraction of the raw bytes; a pathological canvas can exceed them
            own = out = np.empty(raw_bytes // 2 + 4096 * n_canvas + 65536, dtype=np.uint8)
        for attempt in range(2):
            if isinstance(out, np.ndarray):
                if out.dtype != np.uint8 or not out.flags.c_contiguous:
                    raise ValueError("out must be a C-contiguous uint8 array")
                ptr, cap, on_dev = out.ctypes.data, out.nbytes, 0
            else:
                self._check_device_tensor(out, 0)
                ptr, cap, on_dev = out.data_ptr(), out.numel() * out.element_size(), 1
            rc = fn(*head, ptr, cap, on_dev, offsets.ctypes.data, int(bool(timing)), C.byref(stats))
            if rc == _lib.E_NOMEM and own is not None and attempt == 0 and offsets[-1] > cap:
                own = out = np.empty(int(offsets[-1]), dtype=np.uint8)  # rare: run again with the exact size
                continue
            self._check(rc)
            break
        res = stats.as_dict()
        res["offsets"] = offsets
        if own is not None:
            res["png"] = own[: int(offsets[-1])]
        self.last_stats = res
        return res

    def render_png(self, program: Program, out=None, timing: bool = False, stream=None):
        """svgr_render_png: render and PNG-encode every canvas on the device.  -> stats with "offsets" (int64,
        n_canvas + 1) and, when out is None, "png" (uint8 array holding the files back to back)."""
        cprog, keep = program.to_c()
        self.program, self._keep = program, keep
        return self._png_call(self.L.svgr_render_png, (self.ctx, C.byref(cprog), self._stream(stream)),
                              len(program.canvases), program.canvas_bytes, out, timing)

    def render_resident_png(self, out=None, timing: bool = False, stream=None):
        if self.program is None:
            raise ValueError("no program resident: call render() / render_png() first")
        return self._png_call(self.L.svgr_render_resident_png, (self.ctx, self._stream(stream)),
                              len(self.program.canvases), self.program.canvas_bytes, out, timing)

    def png_encode(self, images):
        """PNG-encode host RGBA8 images (list of (rows, cols, 4) uint8 arrays) on the device -> list of bytes."""
        images = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        for im in images:
            if im.ndim != 3 or im.shape[2] != 4:
                raise ValueError("png_encode expects (rows, cols, 4) uint8 images")
        if not images:
            return []
        flat = np.concatenate([im.reshape(-1) for im in images])
        rows = np.asarray([im.shape[0] for im in images], dtype=np.int32)
        cols = np.asarray([im.shape[1] for im in images], dtype=np.int32)
        offsets = np.zeros(len(images) + 1, dtype=np.int64)
        out = np.empty(flat.nbytes // 2 + 4096 * len(images) + 65536, dtype=np.uint8)
        for attempt in range(2):
            rc = self.L.svgr_png_encode(self.ctx, flat.ctypes.data, len(images), rows.ctypes.data, cols.ctypes.data,
                                        out.ctypes.data, out.nbytes, offsets.ctypes.data)
            if rc == _lib.E_NOMEM and attempt == 0 and offsets[-1] > out.nbytes:
                out = np.empty(int(offsets[-1]), dtype=np.uint8)
                continue
            self._check(rc)
            break
        return [out[offsets[i]: offsets[i + 1]].tobytes() for i in range(len(images))]

    def canvas(self, program: Program, raw: np.ndarray, index: int = 0) -> np.ndarray:
        _node, off, rows, cols = program.canvases[index]
        return raw[off: off + 4 * rows * cols].reshape(rows, cols, 4)

    # -- taps ---------------------------------------------------------------------------------
    def edges(self):
        """-> (edges (E, 4) float64 [r0, c0, r1, c1], edge_path (E,) uint32) of the last render."""
        n = C.c_int64()
        self._check(self.L.svgr_read_edges(self.ctx, None, None, 0, C.byref(n)))
        e = np.empty((n.value, 4), dtype=np.float64)
        p = np.empty(n.value, dtype=np.uint32)
        if n.value:
            self._check(self.L.svgr_read_edges(self.ctx, e.ctypes.data, p.ctypes.data, n.value, C.byref(n)))
        return e, p

    def boxes(self, minmax: bool = False):
        n = len(self.program.paths)
        b = np.zeros((n, 4), dtype=np.int32)
        mm = np.zeros((n, 4), dtype=np.float64) if minmax else None
        self._check(self.L.svgr_read_boxes(self.ctx, b.ctypes.data, None if mm is None else mm.ctypes.data))
        return (b, mm) if minmax else b

    def mask(self, path: int, box) -> np.ndarray:
        rows, cols = int(box[2]), int(box[3])
        out = np.zeros((max(rows, 0), max(cols, 0)), dtype=np.float32)
        if out.size:
            self._check(self.L.svgr_read_mask(self.ctx, int(path), out.ctypes.data))
        return out

    def bins(self, path: int, box):
        nb = (max(int(box[2]), 0) + 15) // 16
        off = np.zeros(nb + 1, dtype=np.int32)
        n = C.c_int64()
        self._check(self.L.svgr_read_bins(self.ctx, int(path), off.ctypes.data, None, 0, C.byref(n)))
        ids = np.zeros(n.value, dtype=np.uint32)
        if n.value:
            self._check(self.L.svgr_read_bins(self.ctx, int(path), off.ctypes.data, ids.ctypes.data, n.value, C.byref(n)))
        return off, ids

    def outline(self):
        """Outline segments of the last render's strokes: (tag, data (n, 8), path, sub) without padding."""
        n = C.c_int64()
        self._check(self.L.svgr_read_outline(self.ctx, None, None, None, None, 0, C.byref(n)))
        tag = np.empty(n.value, dtype=np.uint8)
        data = np.empty((n.value, 8), dtype=np.float64)
        path = np.empty(n.value, dtype=np.uint32)
        sub = np.empty(n.value, dtype=np.int32)
        if n.value:
            self._check(self.L.svgr_read_outline(self.ctx, tag.ctypes.data, data.ctypes.data, path.ctypes.data,
                                                 sub.ctypes.data, n.value, C.byref(n)))
        keep = tag != _lib.SEG_NOP
        return tag[keep], data[keep], path[keep], sub[keep]

    def node_info(self, node: int):
        """-> (kind 0 empty / 1 RGBA / 2 one channel, r0, c0, rows, cols, pre_alpha, linear_rgb, virtual)"""
        info = (C.c_int32 * 8)()
        self._check(self.L.svgr_node_info(self.ctx, int(node), info))
        return tuple(info)

    def node(self, node: int):
        """Layer of a node after a full render: (image float32 (rows, cols, ch), (r0, c0), pre, lin) or None."""
        kind, r0, c0, rows, cols, pre, lin, _virtual = self.node_info(node)
        if kind == 0:
            return None
        ch = 4 if kind == 1 else 1
        img = np.empty((rows, cols, ch), dtype=np.float32)
        self._check(self.L.svgr_read_node(self.ctx, int(node), img.ctypes.data))
        return img, (r0, c0), bool(pre), bool(lin)

    # -- eager element-wise entry points ---------------------------------------------
    def line_signed_coverage(self, trace: np.ndarray, lines: np.ndarray) -> np.ndarray:
        """trace (rows, cols) float32 C-contiguous, updated in place; lines (n, 4) float64 [r0, c0, r1, c1]."""
        lines = np.ascontiguousarray(lines, dtype=np.float64).reshape(-1, 4)
        self._check(self.L.svgr_line_signed_coverage(self.ctx, trace.ctypes.data, trace.shape[0], trace.shape[1],
                                                     lines.ctypes.data if len(lines) else None, len(lines)))
        return trace

    def grad_pixels(self, r0, c0, rows, cols) -> np.ndarray:
        out = np.empty((int(rows), int(cols), 2), dtype=np.float64)
        self._check(self.L.svgr_grad_pixels(self.ctx, int(r0), int(c0), int(rows), int(cols), _lib.ptr(out)))
        return out

    def grad_spread(self, offsets: np.ndarray, spread: int) -> np.ndarray:
        offsets = np.ascontiguousarray(offsets, dtype=np.float64)
        out = np.empty_like(offsets)
        self._check(self.L.svgr_grad_spread(self.ctx, _lib.ptr(offsets), offsets.size, int(spread), _lib.ptr(out)))
        return out

    def grad_interpolate(self, offsets: np.ndarray, stops: np.ndarray) -> np.ndarray:
        """stops: _lib.STOP_DT records already in the target colour space -> (..., 4) float32"""
        offsets = np.ascontiguousarray(offsets, dtype=np.float64)
        out = np.empty((*offsets.shape, 4), dtype=np.float32)
        stops = np.ascontiguousarray(stops)
        self._check(self.L.svgr_grad_interpolate(self.ctx, _lib.ptr(offsets), offsets.size, _lib.ptr(stops), len(stops),
                                                 _lib.ptr(out)))
        return out

    def quantize_u8(self, values: np.ndarray) -> np.ndarray:
        values = np.ascontiguousarray(values, dtype=np.float32)
        out = np.empty(values.shape, dtype=np.uint8)
        self._check(self.L.svgr_quantize_u8(self.ctx, _lib.ptr(values), values.size, _lib.ptr(out)))
        return out

    def pooling(self, mat: np.ndarray, ksize, stride, method: int, pad: bool) -> np.ndarray:
        mat = np.ascontiguousarray(mat, dtype=np.float32)
        rows, cols, ch = mat.shape
        (ky, kx), (sy, sx) = ksize, stride
        if pad:
            orows, ocols = -(-rows // sy), -(-cols // sx)
        else:
            orows = (rows - ky) // sy + 1 if rows >= ky else 0
            ocols = (cols - kx) // sx + 1 if cols >= kx else 0
        out = np.empty((max(orows, 0), max(ocols, 0), ch), dtype=np.float32)
        if out.size == 0 or mat.size == 0:
            return out
        self._check(self.L.svgr_pooling(self.ctx, _lib.ptr(mat), rows, cols, ch, int(ky), int(kx), int(sy), int(sx),
                                        int(method), int(bool(pad)), _lib.ptr(out), orows, ocols))
        return out

    def cloud_bounds(self, path_lists, inverses):
        """ConvexHull.bbox reductions: per query (list of path ids, inverse 2x3) -> (minx, miny, maxx, maxy)."""
        nq = len(path_lists)
        off = np.zeros(nq + 1, dtype=np.int32)
        off[1:] = np.cumsum([len(p) for p in path_lists])
        paths = np.asarray([p for ps in path_lists for p in ps], dtype=np.int32)
        inv = np.ascontiguousarray(np.asarray(inverses, dtype=np.float64).reshape(nq, 6))
        out = np.zeros((nq, 4), dtype=np.float64)
        self._check(self.L.svgr_cloud_bounds(self.ctx, nq, off.ctypes.data, paths.ctypes.data if len(paths) else None,
                                             inv.ctypes.data, out.ctypes.data))
        return out


_default = {}


def default_engine(device: int | None = None) -> Engine:
    """Process-wide engine for the eager API (one per device)."""
    if device is None:
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("svgrasterize B200 core: no CUDA device available (there is no CPU fallback)")
        device = torch.cuda.current_device()
    if device not in _default:
        _default[device] = Engine(device)
    return _default[device]
