"""svgrasterize B200 core: the rasterizer hot path of aslpavel/svgrasterize.py
(flatten, coverage, paint + compose, filters) as sm_100a CUDA kernels behind a
C-ABI, with a host-side mirror of the reference's Path / Layer / Scene call
surface.  See DESIGN.md and INTEGRATION.md at the repository root."""
from . import scene, sceneio  # noqa: F401
from .scene import *  # noqa: F401,F403

__version__ = "0.1.0"
