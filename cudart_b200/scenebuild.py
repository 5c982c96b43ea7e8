"""Host scene assembly through the C ABI: the reference's ``constructBvh`` + ``constructLightList``
(``scene.cpp:140-165``, ``bvh.cpp:65-323``) rebuilt by ``libcrt_b200.so`` (``csrc/host_scene.cpp``)."""
from __future__ import annotations

import ctypes as C
import dataclasses

import numpy as np

from . import capi
from .scenefile import NODE_DTYPE, TRIANGLE_DTYPE, FlatScene


def build_bvh(tris: np.ndarray, positions: np.ndarray):
    """tris: TRIANGLE_DTYPE[n] in OBJ order; positions: f32[n, 9].  Returns (tris in BVH order, nodes)."""
    tris = np.ascontiguousarray(tris, TRIANGLE_DTYPE).copy()
    positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 9)
    if positions.shape[0] != tris.shape[0]:
        raise ValueError("positions must hold 9 floats per triangle")
    nodes = np.zeros(max(2 * tris.shape[0], 1), NODE_DTYPE)
    n = C.c_uint32(0)
    capi.check(capi.lib().crt_build_bvh(tris.ctypes.data, positions.ctypes.data, tris.shape[0], nodes.ctypes.data, C.byref(n)))
    return tris, nodes[: n.value].copy()


def build_lights(tris: np.ndarray, mat_floats: np.ndarray, mat_indices: np.ndarray):
    """Returns (tris with _triId rewritten for lights, ka-scaled mat_floats, light_indices, lights_area)."""
    tris = np.ascontiguousarray(tris, TRIANGLE_DTYPE).copy()
    mf = np.ascontiguousarray(mat_floats, np.float32).copy()
    mi = np.ascontiguousarray(mat_indices, np.int32)
    lights = np.zeros(max(tris.shape[0], 1), np.uint32)
    n, area = C.c_uint32(0), C.c_float(0)
    capi.check(capi.lib().crt_build_lights(tris.ctypes.data, tris.shape[0], mf.ctypes.data, mi.ctypes.data, mi.shape[0],
                                           lights.ctypes.data, C.byref(n), C.byref(area)))
    return tris, mf, lights[: n.value].copy(), float(area.value)


def wide_bvh(nodes: np.ndarray) -> np.ndarray:
    """The 128-byte 4-ary nodes the kernels traverse, as raw u32[n, 32] (introspection for tests)."""
    nodes = np.ascontiguousarray(nodes, NODE_DTYPE)
    n = C.c_uint32(0)
    capi.check(capi.lib().crt_wide_bvh(nodes.ctypes.data, nodes.shape[0], None, 0, C.byref(n)))
    out = np.zeros((n.value, 32), np.uint32)
    capi.check(capi.lib().crt_wide_bvh(nodes.ctypes.data, nodes.shape[0], out.ctypes.data, n.value, C.byref(n)))
    return out


PAD_AUTO, PAD_STATIC, PAD_DYNAMIC = 0, 1, 2


def tight_bvh(tris: np.ndarray, nodes: np.ndarray, region_lo, region_hi, sah: bool = False, pad_mode: int = PAD_STATIC,
              pads: bool = False):
    """The box a ray (starting inside the region) must pass before each triangle is tested in the tight BVH (f32[n, 6],
    reference triangle order), plus (individually boxed triangles, binary nodes, wide nodes).  Introspection for tests.
    pads=True appends (f32[n, 2] run-time pad terms (a1, b2) per triangle, whether the dynamic-pad tree was built)."""
    tris = np.ascontiguousarray(tris, TRIANGLE_DTYPE)
    nodes = np.ascontiguousarray(nodes, NODE_DTYPE)
    region = np.asarray(list(region_lo) + list(region_hi), np.float32)
    boxes = np.zeros((tris.shape[0], 6), np.float32)
    nb, nn, nw = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    cost = np.zeros(4, np.float64)      # SAH (nodes, triangles) per random line: tight tree, reference-leaf tree
    pad = np.zeros((tris.shape[0], 2), np.float32)
    dyn = C.c_uint32(0)
    capi.check(capi.lib().crt_tight_bvh(tris.ctypes.data, tris.shape[0], nodes.ctypes.data, nodes.shape[0], region.ctypes.data,
                                        int(pad_mode), boxes.ctypes.data, pad.ctypes.data, C.byref(dyn),
                                        C.byref(nb), C.byref(nn), C.byref(nw), cost.ctypes.data))
    info = (int(nb.value), int(nn.value), int(nw.value))
    out = (boxes, info, cost) if sah else (boxes, info)
    return out + (pad, bool(dyn.value)) if pads else out


def derived_fingerprint(tris: np.ndarray, nodes: np.ndarray, region_lo, region_hi, pad_mode: int = PAD_AUTO):
    """128-bit hash of every array crt_create derives on the host (trees, flattenings, triangle records)."""
    tris = np.ascontiguousarray(tris, TRIANGLE_DTYPE)
    nodes = np.ascontiguousarray(nodes, NODE_DTYPE)
    region = np.asarray(list(region_lo) + list(region_hi), np.float32)
    out = (C.c_uint64 * 2)()
    capi.check(capi.lib().crt_derived_fingerprint(tris.ctypes.data, tris.shape[0], nodes.ctypes.data, nodes.shape[0], region.ctypes.data,
                                                  int(pad_mode), C.byref(out)))
    return (int(out[0]), int(out[1]))


def assemble(tris, positions, mat_floats, mat_indices, camera, tex_dims=None, tex_offsets=None, tex_data=None,
             width=0, height=0) -> FlatScene:
    """OBJ-order triangles + unscaled material table -> the flat scene the renderers consume."""
    t, nodes = build_bvh(tris, positions)
    t, mf, lights, area = build_lights(t, mat_floats, mat_indices)
    z = lambda a, dt, shape: np.zeros(shape, dt) if a is None else np.ascontiguousarray(a, dt)
    return FlatScene(t, nodes, mf.reshape(-1, 4, 3), np.ascontiguousarray(mat_indices, np.int32).reshape(-1, 2), lights, area,
                     camera, z(tex_dims, np.uint32, (0,)), z(tex_offsets, np.uint32, (0,)),
                     z(tex_data, np.float32, (0, 3)).reshape(-1, 3), width, height)
