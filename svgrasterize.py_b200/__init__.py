"""svgrasterize B200 core: the rasterizer hot path of aslpavel/svgrasterize.py
(flatten, coverage, paint + compose, filters) as sm_100a CUDA kernels behind a
C-ABI, with a host-side mirror of the reference's Path / Layer / Scene call
surface.  See DESIGN.md and INTEGRATION.md at the repository root."""
from . import scene, sceneio  # noqa: F401
from .scene import *  # noqa: F401,F403
from . import api  # noqa: F401,E402  (attaches mask / fill / stroke / render to Path, Scene and Filter)
from .api import (Layer, canvas_compose, canvas_create, canvas_merge_at, canvas_merge_intersect,  # noqa: F401,E402
                  canvas_merge_union, canvas_to_png, pooling, bezier3_flatten_batch, blur_kernel, render_canvas,
                  line_signed_coverage, grad_pixels, grad_spread, grad_interpolate, install, uninstall,
                  render_png, render_png_batch)

__version__ = "0.1.0"
