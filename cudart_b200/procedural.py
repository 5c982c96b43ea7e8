"""Procedural scenes assembled with the product's own host scene build (no reference code involved):
the height-field of BASELINE.json config 4 ("synthetic procedurally tessellated 1M-triangle scene").

``terrain_scene(n)`` produces the same flat arrays the reference's loader + builder produce from the OBJ that
``oracle/make_assets.py`` writes for the same ``n`` (checked bit for bit against the ``terrain64`` golden
fixture in tests/test_host_logic.py): vertices are rounded through the OBJ's ``%.6f`` text form, triangles
follow ``scene.cpp:93-134`` (v1, e1 = v2 - v1, e2 = v3 - v1, area = |e1 x e2| / 2 in fp32), materials follow
``scene.cpp:175-190`` (ka = Ka * LIGHTS_GAIN), and the BVH / light list come from ``crt_build_bvh`` /
``crt_build_lights``.
"""
from __future__ import annotations

import math

import numpy as np

from . import scenebuild
from .scenefile import CAMERA_DTYPE, DIFFUSE, EMISSIVE, TRIANGLE_DTYPE, FlatScene

LIGHTS_GAIN = np.float32(3.0)       # scene.cpp:19


def _through_obj_text(a: np.ndarray) -> np.ndarray:
    """What std::stof reads back from '%.6f' (obj_load.h parses vertex lines with stof)."""
    return np.char.mod("%.6f", a.astype(np.float64)).astype(np.float64).astype(np.float32)


def _tanf(x: np.float32) -> np.float32:
    """The C library's tanf (json_loader.cpp:35 calls it on a float); numpy's float32 tan may differ in the last bit."""
    import ctypes
    import ctypes.util
    libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    libm.tanf.restype = ctypes.c_float
    libm.tanf.argtypes = [ctypes.c_float]
    return np.float32(libm.tanf(float(x)))


def camera_from_json(cam: dict, width: int, height: int) -> np.ndarray:
    """CameraJsonLoader::getCamera (json_loader.cpp:25-47)."""
    f32 = np.float32
    fov_deg, focal, fstop = f32(cam["fieldOfView"]), f32(cam["focalLength"]), f32(cam["fStop"])
    eye = np.asarray(cam["eye"], np.float64).astype(np.float32)
    d = np.asarray(cam["viewDirection"], np.float64).astype(np.float32)
    u = np.asarray(cam["upDirection"], np.float64).astype(np.float32)

    def length(v):
        return f32(np.sqrt(f32(f32(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])))

    def normalize(v):
        n = length(v)
        return np.array([v[0] / n, v[1] / n, v[2] / n], np.float32)

    def cross(a, b):
        return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]], np.float32)

    out = np.zeros(1, CAMERA_DTYPE)
    out["xpixels"], out["ypixels"] = width, height
    out["fov"] = _tanf(f32(float(f32(fov_deg * f32(0.5))) * math.pi / 180.0))
    out["eye"] = eye
    out["focusDistance"] = length(d)
    nd = normalize(d)
    out["dir"] = nd
    nu = normalize(u)
    out["up"] = nu
    out["right"] = normalize(cross(nd, nu))
    out["apertureWidth"] = focal / fstop
    out["aspect"] = f32(width) / f32(height)
    return out


TERRAIN_CAMERA = {
    "fStop": 1000.0, "eye": [0.0, 12.0, 40.0], "focalLength": 0.032, "fieldOfView": 60.0,
    "upDirection": [0.0, 1.0, 0.0], "viewDirection": [0.0, -13.68, -37.59],
}


def terrain_triangles(n: int, extent: float = 30.0):
    """OBJ-order triangles (TRIANGLE_DTYPE) and their vertex positions f32[m, 9] for the n x n quad height field
    plus the 12 x 12 emissive quad at y = 14 (oracle/make_assets.py:terrain_obj writes the same geometry)."""
    step = 2.0 * extent / n
    idx = np.arange(n + 1, dtype=np.float64)
    x = -extent + idx * step
    z = -extent + idx * step
    X, Z = np.meshgrid(x, z)                      # row j = z index, column i = x index
    Y = 1.5 * np.sin(0.35 * X) * np.cos(0.35 * Z) + 0.25 * np.sin(3.1 * X + 1.7 * Z)
    V = np.stack([_through_obj_text(X.ravel()), _through_obj_text(Y.ravel()), _through_obj_text(Z.ravel())], axis=1)
    j, i = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a = (j * (n + 1) + i).ravel()
    b, c = a + 1, a + (n + 1)
    d = c + 1
    # f a c b / f b c d per quad, quads in row-major order
    tri_idx = np.empty((2 * n * n, 3), np.int64)
    tri_idx[0::2] = np.stack([a, c, b], axis=1)
    tri_idx[1::2] = np.stack([b, c, d], axis=1)
    L = np.array([[-6, 14, -6], [6, 14, -6], [-6, 14, 6], [6, 14, 6]], np.float32)
    P = np.concatenate([V[tri_idx].reshape(-1, 9), np.stack([L[[0, 1, 2]].ravel(), L[[2, 1, 3]].ravel()])]).astype(np.float32)
    m = P.shape[0]
    t = np.zeros(m, TRIANGLE_DTYPE)
    v1, v2, v3 = P[:, 0:3], P[:, 3:6], P[:, 6:9]
    e1, e2 = v2 - v1, v3 - v1
    t["triId"] = np.arange(m)
    t["materialId"][-2:] = 1
    t["v1"], t["e1"], t["e2"] = v1, e1, e2
    up = np.array([0, 1, 0], np.float32)
    for k in ("n1", "n2", "n3"):
        t[k] = up
        t[k][-2:] = np.array([0, -1, 0], np.float32)
    cx = e1[:, 1] * e2[:, 2] - e1[:, 2] * e2[:, 1]
    cy = e1[:, 2] * e2[:, 0] - e1[:, 0] * e2[:, 2]
    cz = e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
    t["area"] = np.sqrt((cx * cx + cy * cy) + cz * cz) / np.float32(2.0)
    return t, P


def terrain_scene(n: int, width: int, height: int) -> FlatScene:
    tris, positions = terrain_triangles(n)
    mat_floats = np.zeros((2, 4, 3), np.float32)
    mat_floats[0, 0] = [0.55, 0.50, 0.42]                       # ground: Kd
    mat_floats[0, 3] = [0.0, 1.0, 0.0]                          # (Ns, Ni, diffuse): obj_load.h defaults, Ni from the mtl
    mat_floats[1, 0] = [1.0, 1.0, 1.0]                          # lamp: Kd
    mat_floats[1, 1] = np.array([1.0, 0.95, 0.85], np.float32) * LIGHTS_GAIN
    mat_floats[1, 3] = [0.0, 1.0, 0.0]
    mat_indices = np.array([[DIFFUSE, -1], [EMISSIVE, -1]], np.int32)
    cam = camera_from_json(TERRAIN_CAMERA, width, height)
    return scenebuild.assemble(tris, positions, mat_floats, mat_indices, cam, width=width, height=height)
