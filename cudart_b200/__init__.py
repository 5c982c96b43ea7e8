"""cudart_b200 - B200-native (sm_100a) implementation of CudaRT's per-pixel path-tracing hot path.

Only what that path needs lives here: ``csrc/`` (CUDA kernels + host scene build + the C ABI of
``include/crt_abi.h``), ``capi`` (ctypes binding), ``renderer`` (mirror of the reference's ``Renderer`` seam),
``scenebuild`` (host scene assembly through the C ABI), ``sharding`` (row-tile multi-GPU plumbing) and
``scenefile`` (flattened scene container).
"""
from .scenefile import FlatScene, load_crts, save_crts  # noqa: F401

__all__ = ["FlatScene", "load_crts", "save_crts"]
